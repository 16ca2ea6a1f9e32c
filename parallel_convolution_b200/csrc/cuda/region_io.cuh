// Global memory <-> register movement of CTA regions (band loads/stores), owner lookup across the
// 3x3 tile neighbourhood, and the system-scope flag primitives of the in-kernel halo protocol.
//
// This is where the halo exchange is *fused* into the stencil kernel: a region row/column that
// belongs to a neighbouring tile is loaded straight from that tile's buffer through its
// peer-mapped pointer (NVLink P2P load when the neighbour is another GPU, a plain load when it is a
// virtual rank on the same device). There is no send side, no packing and no staging copy — the
// reference needs MPI_Isend/Irecv + MPI_Type_vector for the same job (mpi/mpi_conv.c:80-88,
// 166-200) and never moves the diagonal cells (defect D1); here corners are just the part of the
// region that happens to fall into a diagonal neighbour's tile.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pconv_cuda.h"
#include "stencil_core.cuh"

namespace pconv {

// 128-bit WEAK global store. (__stcg compiles to STG.E.128.STRONG.GPU on sm_100a — a gpu-scope strong store
// per chunk; measured, a burst of them costs a CTA ~2 us where weak stores cost ~0.2 us. Visibility does not
// need it: every consumer of these bytes synchronises through a release/acquire flag pair or a launch
// boundary, both of which order weak stores as well, and L1 is write-through.)
__device__ __forceinline__ void st_weak(uint4* p, const uint4& v) {
    asm volatile("st.weak.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct SourceSet {
    const uint8_t* p[9];
    int64_t pitch[9];
};

__device__ __forceinline__ int owner_of(int64_t v, const int64_t (&lo)[4]) {
    return v < lo[1] ? 0 : (v < lo[2] ? 1 : 2);
}

// Slow path of load_chunk: a chunk that straddles an ownership boundary / the image edge, or is
// misaligned in its owner's buffer, is assembled byte by byte. Kept out of line: it is rare (only
// odd tile widths hit it) and would otherwise be unrolled into every band row.
static __device__ __noinline__ uint4 load_chunk_bytes(const pconv_geometry_t& g, const SourceSet& s,
                                               int orow, int64_t lrow, int64_t gb) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int t = 0; t < 16; ++t) {
        const int64_t b = gb + t;
        if (b < g.cb_lo[0] || b >= g.cb_lo[3]) continue;   // outside the 3x3 tile window: zero
        const int oc = owner_of(b, g.cb_lo);
        const int d = orow * 3 + oc;
        const uint8_t* base = s.p[d];
        if (base == nullptr) continue;
        const uint32_t v = __ldcg(base + lrow * s.pitch[d] + (b - g.cb_lo[oc]));
        w[t >> 2] |= v << (8 * (t & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// A 16-byte chunk that lies inside ONE owner's buffer but not on a 16-byte boundary there (tiles whose width
// is not a whole number of chunks shift everything the neighbour sees): two aligned 128-bit loads and a
// 128-bit funnel shift instead of sixteen single-byte loads. The extra bytes read belong to the same
// allocation (every tile buffer ends with >= 64 bytes of slack) and are discarded.
static __device__ __noinline__ uint4 load_chunk_unaligned(const uint8_t* a) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(a);
    const uint4* q = reinterpret_cast<const uint4*>(u & ~(uintptr_t)15);
    const uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
    const uint32_t bs = (uint32_t)(u & 3u) * 8u;
    uint32_t w0, w1, w2, w3, w4;
    switch ((u >> 2) & 3u) {       // word shift: uniform over the warp (pitches and chunk steps are multiples of 16)
        case 0: w0 = lo.x; w1 = lo.y; w2 = lo.z; w3 = lo.w; w4 = hi.x; break;
        case 1: w0 = lo.y; w1 = lo.z; w2 = lo.w; w3 = hi.x; w4 = hi.y; break;
        case 2: w0 = lo.z; w1 = lo.w; w2 = hi.x; w3 = hi.y; w4 = hi.z; break;
        default: w0 = lo.w; w1 = hi.x; w2 = hi.y; w3 = hi.z; w4 = hi.w; break;
    }
    return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs),
                      __funnelshift_r(w3, w4, bs));
}

// One 16-byte chunk of the image starting at global byte column gb of global row gr: a single
// 128-bit load through the owning tile's (possibly peer-mapped) pointer — L2-coherent, L1 bypass,
// because the data was written by other CTAs / other GPUs. Zero outside the image.
__device__ __forceinline__ uint4 load_chunk(const pconv_geometry_t& g, const SourceSet& s,
                                            int64_t gr, int64_t gb) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    // (the window [row_lo[0], row_lo[3]) x [cb_lo[0], cb_lo[3]) never sticks out of the image; what lies
    //  beyond it but inside the image belongs to a tile two steps away and is never needed: zero)
    if (gr < g.row_lo[0] || gr >= g.row_lo[3] || gb + 16 <= g.cb_lo[0] || gb >= g.cb_lo[3]) return z;
    const int orow = owner_of(gr, g.row_lo);
    const int64_t lrow = gr - g.row_lo[orow];
    if (gb >= g.cb_lo[0] && gb + 16 <= g.cb_lo[3]) {
        const int oc = owner_of(gb, g.cb_lo);
        if (owner_of(gb + 15, g.cb_lo) == oc) {
            const int d = orow * 3 + oc;
            const uint8_t* base = s.p[d];
            if (base == nullptr) return z;
            const uint8_t* a = base + lrow * s.pitch[d] + (gb - g.cb_lo[oc]);
            if ((reinterpret_cast<uintptr_t>(a) & 15u) == 0)
                return __ldcg(reinterpret_cast<const uint4*>(a));
        }
    }
    return load_chunk_bytes(g, s, orow, lrow, gb);
}

// Load R consecutive rows of one 16-byte column strip (a warp's band) into registers. Everything
// that does not depend on the row is hoisted: the column classification (owner, offset, fast/slow),
// the address of band row 0 in each of the three owning tile-rows' buffers, and the band-relative
// 32-bit row ranges (inside the image / owned by tile-row 0, 1, 2). Per row that leaves two 32-bit
// compares, a pointer select, one 32x32+64 multiply-add and one 128-bit load — the region load is
// executed by every thread of every launch, so its instruction count is launch overhead.
// (Pitches are < 2^31: checked by the launchers.)
template <int R, bool MASKED = false>
__device__ __forceinline__ void load_band(uint4 (&rows)[R], const pconv_geometry_t& g,
                                          const SourceSet& s, int64_t gr0, int64_t gb,
                                          uint32_t mask = 0xffffffffu) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    // Everything is clipped to the 3x3 tile window [row_lo[0], row_lo[3]) x [cb_lo[0], cb_lo[3]): a region
    // overhang that reaches past the adjacent tile (small tiles, >= 3 tile rows/columns) is never needed
    // and must not be read — it would fall outside that neighbour's buffer.
    const bool col_out = gb + 16 <= g.cb_lo[0] || gb >= g.cb_lo[3];
    const bool col_whole = gb >= g.cb_lo[0] && gb + 16 <= g.cb_lo[3];
    const int oc = owner_of(gb < g.cb_lo[0] ? g.cb_lo[0] : gb, g.cb_lo);
    const bool one_owner = col_whole && owner_of(gb + 15, g.cb_lo) == oc;
    const int64_t cofs = gb - g.cb_lo[oc];
    // band-relative row boundaries, clamped to [0, R]
    auto rel = [&](int64_t v) { v -= gr0; return (int)(v < 0 ? 0 : (v > R ? R : v)); };
    const int i_lo = rel(g.row_lo[0]), i_hi = rel(g.row_lo[3]);  // rows inside the tile window
    const int n1 = rel(g.row_lo[1]), n2 = rel(g.row_lo[2]);     // tile-row 0 below n1, 1 below n2, else 2
    // address of (image row gr0, my column) in the buffer of owning tile-row o, and whether 128-bit
    // loads are legal there
    auto origin = [&](int o, int& pitch, bool& fast) -> const uint8_t* {
        pitch = (int)s.pitch[3 * o + oc];
        const uint8_t* base = s.p[3 * o + oc];
        const uint8_t* a = base ? base + cofs + (gr0 - g.row_lo[o]) * pitch : nullptr;
        fast = one_owner && a && ((reinterpret_cast<uintptr_t>(a) | (uintptr_t)pitch) & 15u) == 0;
        return a;
    };
    // Common case first (warp-uniform except for the alignment test): the whole band lies inside the
    // image and inside ONE owning tile-row — per row a multiply-add and a 128-bit load.
    if (i_lo == 0 && i_hi == R && !col_out && (n1 == R || (n1 == 0 && n2 == R) || n2 == 0)) {
        int pitch;
        bool fast;
        const uint8_t* a = origin(n1 == R ? 0 : (n2 == R ? 1 : 2), pitch, fast);
        if (fast) {
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (!MASKED || ((mask >> i) & 1u))
                    rows[i] = __ldcg(reinterpret_cast<const uint4*>(a + (int64_t)i * pitch));
            return;
        }
    }
    int p0, p1, p2;
    bool f0, f1, f2;
    const uint8_t* a0 = origin(0, p0, f0);
    const uint8_t* a1 = origin(1, p1, f1);
    const uint8_t* a2 = origin(2, p2, f2);
    uint32_t slow = 0;   // rows that need the byte-wise path (rare: odd tile widths, image edge)
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (MASKED && !((mask >> i) & 1u)) continue;   // keep this row (resident halo refresh)
        if (col_out || i < i_lo || i >= i_hi) { rows[i] = z; continue; }
        const bool in0 = i < n1, in1 = i < n2;
        const uint8_t* a = in0 ? a0 : (in1 ? a1 : a2);
        const int pitch = in0 ? p0 : (in1 ? p1 : p2);
        const bool fast = in0 ? f0 : (in1 ? f1 : f2);
        if (fast) rows[i] = __ldcg(reinterpret_cast<const uint4*>(a + (int64_t)i * pitch));
        else slow |= 1u << i;
    }
    if (slow) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (!((slow >> i) & 1u)) continue;
            const int64_t gr = gr0 + i;
            const int o = owner_of(gr, g.row_lo);
            const uint8_t* base = s.p[3 * o + oc];
            if (one_owner && base != nullptr)      // whole chunk in one buffer, only misaligned there
                rows[i] = load_chunk_unaligned(base + cofs + (gr - g.row_lo[o]) * s.pitch[3 * o + oc]);
            else
                rows[i] = load_chunk_bytes(g, s, o, gr - g.row_lo[o], gb);
        }
    }
}

// Slow path of store_band: a chunk that straddles the tile edge or is misaligned goes out byte by
// byte. Out of line for the same reason as load_chunk_bytes.
static __device__ __noinline__ void store_chunk_bytes(uint8_t* a, int64_t lb, int64_t tile_cb, uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int t = 0; t < 16; ++t)
        if (lb + t >= 0 && lb + t < tile_cb) a[t] = (uint8_t)(w[t >> 2] >> (8 * (t & 3)));
}

// Store the rows of a band selected by `mask` (bit i = row i) into my tile, clipped to the tile. The
// column test, the row clipping (as a bit range) and the address of band row 0 are hoisted out of the
// row loop; per selected row: one multiply-add and one 128-bit store.
template <int R>
__device__ __forceinline__ void store_band(const uint4 (&rows)[R], uint32_t mask,
                                           uint8_t* __restrict__ dst, const pconv_geometry_t& g,
                                           int64_t gr0, int64_t gb) {
    const int64_t lb = gb - g.cb_lo[1], lr0 = gr0 - g.row_lo[1];
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1], tile_cb = g.cb_lo[2] - g.cb_lo[1];
    if (mask == 0u || lb + 16 <= 0 || lb >= tile_cb) return;
    // band rows [i_lo, i_hi) lie inside the tile
    const int i_lo = (int)(lr0 < 0 ? (lr0 < -R ? R : -lr0) : 0);
    const int64_t left = tile_rows - lr0;
    const int i_hi = (int)(left < 0 ? 0 : (left > R ? R : left));
    if (i_lo >= i_hi) return;
    const uint32_t range = (i_hi >= 32 ? 0xffffffffu : ((1u << i_hi) - 1u)) & ~((1u << i_lo) - 1u);
    const uint32_t m = mask & range;
    const int pitch = (int)g.pitch;
    uint8_t* a0 = dst + lr0 * g.pitch + lb;
    const bool fast = lb >= 0 && lb + 16 <= tile_cb && (reinterpret_cast<uintptr_t>(a0) & 15u) == 0 &&
                      (pitch & 15) == 0;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (!((m >> i) & 1u)) continue;
        uint8_t* a = a0 + (int64_t)i * pitch;
        if (fast) st_weak(reinterpret_cast<uint4*>(a), rows[i]);
        else store_chunk_bytes(a, lb, tile_cb, rows[i]);
    }
}

// Byte mask of a lane's 16-byte strip: 0xff where the byte lies inside the image. Only CTAs whose
// region sticks out of the image sideways (*need, CTA-uniform) pay for the per-byte classification.
__device__ __forceinline__ uint4 column_mask(const pconv_geometry_t& g, int64_t gb0, bool* need) {
    *need = (gb0 < 0) || (gb0 + kSP > g.img_cb);
    if (!*need) return make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    const int64_t gb = gb0 + 16 * (threadIdx.x & 31);
    uint32_t m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t w = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int64_t b = gb + 4 * j + t;
            if (b >= 0 && b < g.img_cb) w |= 0xffu << (8 * t);
        }
        m[j] = w;
    }
    return make_uint4(m[0], m[1], m[2], m[3]);
}

// ---- system-scope flags ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// Relaxed system-scope store: after ONE fence.acq_rel.sys a burst of these is a release sequence to every
// flag — a st.release.sys per flag would drain the NVLink write queue once per neighbour (8 x ~1-2 us).
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin until *flag >= want (epochs are monotonically increasing, wrap-safe compare) or the
// watchdog fires. Returns the ns spent waiting; on timeout records error 1 in status[0].
// SYS: the flag is written by another GPU (system scope); otherwise by a CTA of this GPU (gpu
// scope — an order of magnitude cheaper than a system-scope acquire/fence).
template <bool SYS = true>
__device__ __forceinline__ uint64_t wait_flag(const uint32_t* flag, uint32_t want,
                                              uint64_t timeout_ns, uint64_t* status) {
    auto ld = [&]() { return SYS ? ld_acquire_sys(flag) : ld_acquire_gpu(flag); };
    if ((int32_t)(ld() - want) >= 0) return 0;
    const uint64_t t0 = globaltimer_ns();
    uint64_t dt = 0;
    while ((int32_t)(ld() - want) < 0) {
        __nanosleep(32);
        dt = globaltimer_ns() - t0;
        if (dt > timeout_ns) {
            if (status) atomicExch(reinterpret_cast<unsigned long long*>(status), 1ull);
            break;
        }
    }
    return dt;
}

__device__ __forceinline__ void record_wait(uint64_t* status, uint64_t ns) {
    if (!status) return;
    atomicMax(reinterpret_cast<unsigned long long*>(status + 1), (unsigned long long)ns);
    atomicAdd(reinterpret_cast<unsigned long long*>(status + 2), (unsigned long long)ns);
    atomicAdd(reinterpret_cast<unsigned long long*>(status + 3), 1ull);
}

}  // namespace pconv
