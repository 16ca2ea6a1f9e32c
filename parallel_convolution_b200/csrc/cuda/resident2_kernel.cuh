// Split-phase resident kernel (see resident2_kernel.cu for the schedule and the dependency argument);
// one translation unit per arithmetic path instantiates it.
#pragma once

#include <cuda_runtime.h>

#include "pconv_cuda.h"
#include "region_io.cuh"
#include "resident_kernel.cuh"
#include "runtime_internal.h"
#include "stencil_core.cuh"

namespace pconv {

// Per-thread description of one of the CTA's two virtual tiles. Everything here is fixed for the
// whole run; the pair is swapped (together with the register band) at the end of every half-period.
struct VTile {
    int64_t grw;         // image row of this warp's band row 0
    int vi;              // virtual tile row (0 .. nty-1)
    uint32_t m_publish;  // band rows/chunks the neighbours' halos read
    uint32_t m_main;     // halo refreshed while the sibling computes (slack >= 1/4 period)
    uint32_t m_late;     // halo rows facing the sibling row: refreshed right before computing
    uint32_t m_final;    // my interior
    uint32_t rowmask;    // band rows inside the image
    uint32_t w_main;     // neighbour directions (bit d = dy*3+dx) awaited before m_main / m_late
    uint32_t w_late;
    bool masked;         // CTA-uniform: the region touches the image border
    bool first;          // CTA-uniform: computed in the first half of a period
};

template <int R>
__device__ __forceinline__ void swap_bands(uint4 (&a)[R], uint4 (&b)[R], VTile& ta, VTile& tb) {
#pragma unroll
    for (int i = 0; i < R; ++i) { const uint4 t = a[i]; a[i] = b[i]; b[i] = t; }
    const VTile t = ta; ta = tb; tb = t;
}

template <int KIND, int C, int R>
__global__ void __launch_bounds__(kThreads, 1)
resident2_kernel(const __grid_constant__ ResidentArgs A) {
    __shared__ EdgeBuffers eb;
    __shared__ unsigned long long s_wait_ns;
    const pconv_geometry_t& g = A.g;
    constexpr int RH = kWarps * R;
    const int k = A.k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tip = blockIdx.x / A.ntx, tj = blockIdx.x % A.ntx;   // physical CTA row / column
    const int64_t gb0 = g.cb_lo[1] + (int64_t)tj * A.TWB - A.HB;
    const int64_t gb = gb0 + 16 * lane;
    const int fw = A.ntx + 2;  // width of the extended flag grid
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1];

    if (threadIdx.x == 0) s_wait_ns = 0ull;
    __syncthreads();

    // ---- launch-level protocol: the neighbours' input tiles are in place ----------------------------
    const int64_t gr_top = g.row_lo[1] + (int64_t)(2 * tip) * A.TH - k;          // both regions together
    const int64_t gr_end = g.row_lo[1] + (int64_t)(2 * tip + 1) * A.TH - k + RH;
    if (A.sync.wait_value != 0u && threadIdx.x < 9 && threadIdx.x != 4) {
        const int dy = (int)threadIdx.x / 3, dx = (int)threadIdx.x % 3;
        const bool rows_hit = dy == 0 ? gr_top < g.row_lo[1] : dy == 2 ? gr_end > g.row_lo[2] : true;
        const bool cols_hit = dx == 0 ? gb0 < g.cb_lo[1] : dx == 2 ? gb0 + kSP > g.cb_lo[2] : true;
        if (rows_hit && cols_hit && A.src[0].p[threadIdx.x] != nullptr) {
            const uint64_t ns = wait_flag(A.sync.local_flags + threadIdx.x, A.sync.wait_value,
                                          A.sync.timeout_ns, A.sync.status);
            atomicMax(&s_wait_ns, (unsigned long long)ns);
        }
    }
    __syncthreads();

    // ---- the two virtual tiles ------------------------------------------------------------------------
    // cur = the one computed in the first half of a period. Even CTA rows start with their top tile,
    // odd rows with their bottom tile, so virtual tile rows are phased F S S F F S S F ...: the only
    // zero-slack dependencies are between the two tiles of one CTA (and the sibling row's laterals).
    bool need_cols;
    const uint4 cmask = column_mask(g, gb0, &need_cols);
    const int vcb = (int)min((int64_t)A.TWB, g.cb_lo[2] - g.cb_lo[1] - (int64_t)tj * A.TWB);
    const bool col_int = (16 * lane >= A.HB) && (16 * lane < A.HB + vcb);
    const bool col_full = (16 * lane >= A.HB) && (16 * lane + 16 <= A.HB + vcb);
    const bool col_edge = col_int && (16 * lane < 2 * A.HB || 16 * lane + 16 > vcb);

    auto describe = [&](int s, bool first) {
        VTile t;
        t.vi = 2 * tip + s;
        const int64_t gr0 = g.row_lo[1] + (int64_t)t.vi * A.TH - k;
        t.grw = gr0 + (int64_t)warp * R;
        const int vr = (int)max((int64_t)0, min((int64_t)A.TH, tile_rows - (int64_t)t.vi * A.TH));
        const int sd = s == 0 ? 1 : -1;            // where the sibling row lies: +1 below, -1 above
        t.m_publish = t.m_main = t.m_late = t.m_final = t.rowmask = 0u;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int rr = warp * R + i;
            const bool row_int = rr >= k && rr < k + vr;
            const bool row_edge = rr < 2 * k || rr >= vr;
            const bool sib_side = sd > 0 ? rr >= k + vr : rr < k;
            if (row_int && col_int && (row_edge || col_edge)) t.m_publish |= 1u << i;
            if (!(row_int && col_full)) {
                if (first && sib_side) t.m_late |= 1u << i;
                else t.m_main |= 1u << i;
            }
            if (row_int && col_int) t.m_final |= 1u << i;
            if (t.grw + i >= 0 && t.grw + i < g.img_rows) t.rowmask |= 1u << i;
        }
        t.masked = need_cols || gr0 < 0 || gr0 + RH > g.img_rows;
        t.first = first;
        // wait sets, bit d = (dy+1)*3 + (dx+1): the sibling itself is ordered by this CTA's own program
        const uint32_t row_up = 0x007u, row_mid = 0x028u, row_dn = 0x1c0u;   // d = 0..2 | 3,5 | 6..8
        const uint32_t sib_row = sd > 0 ? row_dn : row_up, far_row = sd > 0 ? row_up : row_dn;
        const uint32_t sib_self = sd > 0 ? 0x080u : 0x002u;
        if (first) { t.w_main = far_row | row_mid; t.w_late = sib_row & ~sib_self; }
        else       { t.w_main = (far_row | row_mid | sib_row) & ~sib_self; t.w_late = 0u; }
        return t;
    };
    const int s_first = tip & 1;
    VTile tc = describe(s_first, true), to = describe(1 - s_first, false);

    uint4 cur[R], oth[R];
    load_band<R>(cur, g, A.src[0], tc.grw, gb);
    load_band<R>(oth, g, A.src[0], to.grw, gb);

    // wait for the flags of the neighbours in `wset` around virtual tile row vi to reach `epoch`
    auto wait_set = [&](int vi, uint32_t wset, uint32_t epoch) {
        if (threadIdx.x >= 32 && threadIdx.x < 41 && ((wset >> (threadIdx.x - 32)) & 1u)) {
            const int d = (int)threadIdx.x - 32;
            const int ei = vi + d / 3, ej = tj + d % 3;       // extended-grid coordinates (+1 -1 folded)
            const int sy = ei < 1 ? 0 : (ei > A.nty ? 2 : 1), sx = ej < 1 ? 0 : (ej > A.ntx ? 2 : 1);
            const bool exists = (sy == 1 && sx == 1) || A.src[0].p[sy * 3 + sx] != nullptr;
            if (exists) {
                const uint32_t* fl = A.cta_flags + ei * fw + ej;
                const uint64_t ns = (sy == 1 && sx == 1)
                    ? wait_flag<false>(fl, epoch, A.sync.timeout_ns, A.sync.status)   // local CTA
                    : wait_flag<true>(fl, epoch, A.sync.timeout_ns, A.sync.status);   // other GPU
                atomicMax(&s_wait_ns, (unsigned long long)ns);
                if (ns > A.sync.timeout_ns && A.sync.status != nullptr) {
                    // post-mortem of the first wait that timed out: who waited for whom, and what was there
                    unsigned long long* st = reinterpret_cast<unsigned long long*>(A.sync.status);
                    const unsigned long long what = 1ull << 63 | (unsigned long long)vi << 48 |
                        (unsigned long long)tj << 40 | (unsigned long long)d << 32 | (unsigned long long)epoch;
                    if (atomicCAS(st + 9, 0ull, what) == 0ull) st[10] = (unsigned long long)ld_acquire_sys(fl);
                }
            }
        }
        __syncthreads();
    };

    const bool probe = A.sync.status != nullptr && threadIdx.x == 0 &&
                       blockIdx.x == (unsigned)((A.nty / 4) * A.ntx + A.ntx / 2);
    uint64_t t_pub = 0, t_late = 0, t_main = 0, t_comp = 0, t_mark = probe ? globaltimer_ns() : 0;
    auto lap = [&](uint64_t& acc) {
        if (probe) { const uint64_t t = globaltimer_ns(); acc += t - t_mark; t_mark = t; }
    };

    const int nblocks = (A.loops + k - 1) / k;
    for (int hp = 0; hp < 2 * nblocks; ++hp) {
        const int b = hp >> 1;                               // block computed on `cur` now
        const int kb = min(k, A.loops - b * k);
        // ---- zero-slack part of my halo: the rows facing the sibling row (published a moment ago) ---
        if (b > 0 && tc.first) {
            wait_set(tc.vi, tc.w_late, A.epoch_base + (uint32_t)b);
            load_band<R, true>(cur, g, A.src[b & 1], tc.grw, gb, tc.m_late);
            lap(t_late);
        }
        for (int part = 0; part < 2; ++part) {
            const int kk = part == 0 ? kb / 2 : kb - kb / 2;
            if (tc.masked) region_iterate<KIND, C, R, true>(cur, eb, kk, cmask, tc.rowmask, need_cols, A.f);
            else region_iterate<KIND, C, R, false>(cur, eb, kk, cmask, tc.rowmask, need_cols, A.f);
            __syncthreads();  // edge buffers: the next call restarts at parity 0
            lap(t_comp);
            if (part == 0) {
                // ---- refresh the sibling's halo for ITS next block while I keep computing: its
                // registers are dead until then, so the loads land in the background -------------
                const int bo = (hp + 1) >> 1;
                if (bo > 0 && bo < nblocks) {
                    wait_set(to.vi, to.w_main, A.epoch_base + (uint32_t)bo);
                    load_band<R, true>(oth, g, A.src[bo & 1], to.grw, gb, to.m_main);
                    lap(t_main);
                }
            }
        }
        // ---- publish block b of `cur`: strips + epoch b+1 (the last block goes out as the result) ----
        if (b < nblocks - 1) {
            const uint32_t epoch = A.epoch_base + (uint32_t)(b + 1);
            store_band<R>(cur, tc.m_publish, A.self[(b + 1) & 1], g, tc.grw, gb);
            __syncthreads();
            if (threadIdx.x == 0)
                st_release_gpu(A.cta_flags + (tc.vi + 1) * fw + (tj + 1), epoch);
            if (threadIdx.x < 9 && threadIdx.x != 4) {
                const int sy = (int)threadIdx.x / 3 - 1, sx = (int)threadIdx.x % 3 - 1;
                const bool on_y = sy == 0 || (sy < 0 ? tc.vi == 0 : tc.vi == A.nty - 1);
                const bool on_x = sx == 0 || (sx < 0 ? tj == 0 : tj == A.ntx - 1);
                uint32_t* remote = A.remote_cta_flags[threadIdx.x];
                if (on_y && on_x && remote != nullptr)
                    st_release_sys(remote + (tc.vi + 1 - sy * A.nty) * fw + (tj + 1 - sx * A.ntx), epoch);
            }
            lap(t_pub);
        } else {
            store_band<R>(cur, tc.m_final, A.self[nblocks & 1], g, tc.grw, gb);
        }
        swap_bands<R>(cur, oth, tc, to);
    }
    if (probe) {
        unsigned long long* st = reinterpret_cast<unsigned long long*>(A.sync.status);
        atomicAdd(st + 4, (unsigned long long)t_pub);
        atomicAdd(st + 5, (unsigned long long)t_late);
        atomicAdd(st + 6, (unsigned long long)t_main);
        atomicAdd(st + 7, (unsigned long long)t_comp);
        atomicAdd(st + 8, (unsigned long long)(2 * (nblocks - 1)));
    }
    if (threadIdx.x == 0 && s_wait_ns) record_wait(A.sync.status, s_wait_ns);

    // ---- launch-level protocol: publish "run done" -------------------------------------------------
    if (A.sync.signal_value != 0u) {
        __syncthreads();
        if (threadIdx.x == 0) {   // acq_rel ticket: releases this CTA's stores, acquires the others'
            const uint32_t ticket = atom_add_acq_rel_gpu(A.sync.done_counter, 1u);
            if (ticket == gridDim.x - 1) {
                *A.sync.done_counter = 0u;
                if (!A.sync.inject_drop_signal) {
                    fence_acq_rel_sys();
#pragma unroll
                    for (int d = 0; d < 9; ++d)
                        if (d != 4 && A.sync.remote_flags[d] != nullptr)
                            st_release_sys(A.sync.remote_flags[d], A.sync.signal_value);
                }
            }
        }
    }
}

#define PCONV_RES2_R_SEP(X) X(2) X(4) X(5) X(7)
#define PCONV_RES2_R_INT(X) X(2) X(4) X(5)
#define PCONV_RES2_R_F32(X) X(2) X(4) X(5)

template <int KIND, int C>
static ResidentFn pick_resident2_r(int R) {
#define CASE(RR) if (R == RR) return resident2_kernel<KIND, C, RR>;
    if constexpr (KIND == PCONV_KIND_F32) { PCONV_RES2_R_F32(CASE) }
    else if constexpr (KIND == PCONV_KIND_INT) { PCONV_RES2_R_INT(CASE) }
    else { PCONV_RES2_R_SEP(CASE) }
#undef CASE
    return nullptr;
}

template <int KIND>
static ResidentFn pick_resident2_kind(int C, int R) {
    switch (C) {
        case 1: return pick_resident2_r<KIND, 1>(R);
        case 3: return pick_resident2_r<KIND, 3>(R);
        case 4: return pick_resident2_r<KIND, 4>(R);
        default: return nullptr;
    }
}

ResidentFn pick_resident2_f32(int C, int R);
ResidentFn pick_resident2_int(int C, int R);
ResidentFn pick_resident2_sep(int C, int R);

}  // namespace pconv
