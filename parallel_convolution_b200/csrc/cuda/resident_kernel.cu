// Resident kernel: one persistent launch per run. The image tile is spread over the CTAs of the
// GPU (one region per CTA, held in registers — stencil_core.cuh) and stays on chip for ALL
// iterations; only k-deep border strips move, every k iterations:
//
//   exchange b (b = 1, 2, ...):
//     1. each CTA stores the border strips of its interior into tile buffer parity (b & 1)
//     2. st.release (gpu scope) of epoch_base+b to its own flag — and, for CTAs
//        on the edge of the GPU's tile, to the mirrored ring entry in the neighbour GPU's flag
//        array (peer store over NVLink)
//     3. eight threads spin (ld.acquire.sys) on the flags of the eight surrounding CTAs — local
//        CTAs or ring entries written by the neighbour GPU
//     4. the halo rows / columns of the region are re-loaded from the surrounding owners' buffers
//        (local L2 or peer HBM over NVLink) straight into registers
//     5. k more iterations in registers
//
// Two buffer parities make the strips double-buffered: writing version b never races with a
// neighbour still reading version b-1, and passing step 3 of exchange b proves every neighbour is
// done with version b-1. The final interior goes to the parity that holds no live strips.
// This is the north-star design of BASELINE.json: persistent, double-buffered, P2P loads of
// neighbour borders inside the stencil kernel, no NCCL/MPI on the path. It replaces the reference's
// per-iteration MPI_Isend/Irecv/Wait schedule (mpi/mpi_conv.c:164-248) and its launch-per-
// iteration CUDA loop (cuda/cuda_convolute.cu:66-87).
#include <cuda_runtime.h>

#include <cstring>

#include "pconv_cuda.h"
#include "region_io.cuh"
#include "runtime_internal.h"
#include "stencil_core.cuh"

namespace pconv {

struct ResidentArgs {
    pconv_geometry_t g;
    SourceSet src[2];        // by parity
    uint8_t* self[2];
    FilterDev f;
    int loops, k, TH, TWB, HB, nty, ntx;
    uint32_t* cta_flags;
    uint32_t* remote_cta_flags[9];
    uint32_t epoch_base;
    pconv_sync_t sync;
};

template <int KIND, int C, int R>
__global__ void __launch_bounds__(kThreads, 1)
resident_kernel(const __grid_constant__ ResidentArgs A) {
    __shared__ EdgeBuffers eb;
    __shared__ unsigned long long s_wait_ns;
    const pconv_geometry_t& g = A.g;
    constexpr int RH = kWarps * R;
    const int k = A.k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x / A.ntx, tj = blockIdx.x % A.ntx;
    const int64_t gr0 = g.row_lo[1] + (int64_t)ti * A.TH - k;
    const int64_t gb0 = g.cb_lo[1] + (int64_t)tj * A.TWB - A.HB;
    const int64_t grw = gr0 + (int64_t)warp * R;
    const int64_t gb = gb0 + 16 * lane;
    const int fw = A.ntx + 2;  // width of the extended flag grid

    if (threadIdx.x == 0) s_wait_ns = 0ull;
    __syncthreads();

    // ---- launch-level protocol: the neighbours' input tiles are in place ----------------------------
    if (A.sync.wait_value != 0u && threadIdx.x < 9 && threadIdx.x != 4) {
        const int dy = (int)threadIdx.x / 3, dx = (int)threadIdx.x % 3;
        const bool rows_hit = dy == 0 ? gr0 < g.row_lo[1] : dy == 2 ? gr0 + RH > g.row_lo[2] : true;
        const bool cols_hit = dx == 0 ? gb0 < g.cb_lo[1] : dx == 2 ? gb0 + kSP > g.cb_lo[2] : true;
        if (rows_hit && cols_hit && A.src[0].p[threadIdx.x] != nullptr) {
            const uint64_t ns = wait_flag(A.sync.local_flags + threadIdx.x, A.sync.wait_value,
                                          A.sync.timeout_ns, A.sync.status);
            atomicMax(&s_wait_ns, (unsigned long long)ns);
        }
    }
    __syncthreads();

    // ---- load the region once ---------------------------------------------------------------------
    uint4 rows[R];
    load_band<R>(rows, g, A.src[0], grw, gb);
    bool need_cols;
    const uint4 cmask = column_mask(g, gb0, &need_cols);
    uint32_t rowmask = 0;
#pragma unroll
    for (int i = 0; i < R; ++i)
        if (grw + i >= 0 && grw + i < g.img_rows) rowmask |= 1u << i;
    const bool masked = need_cols || gr0 < 0 || gr0 + RH > g.img_rows;
    // What this CTA is responsible for ("interior"): its TH x TWB block clipped to my tile. Every
    // other chunk of the region is halo — owned by another CTA, another GPU, or nobody (outside the
    // image) — and is refreshed at each exchange so that staleness never travels more than k pixels.
    const int vr = (int)min((int64_t)A.TH, g.row_lo[2] - g.row_lo[1] - (int64_t)ti * A.TH);   // valid rows
    const int vcb = (int)min((int64_t)A.TWB, g.cb_lo[2] - g.cb_lo[1] - (int64_t)tj * A.TWB);  // valid bytes
    const bool col_int = (16 * lane >= A.HB) && (16 * lane < A.HB + vcb);        // overlaps my interior
    const bool col_full = (16 * lane >= A.HB) && (16 * lane + 16 <= A.HB + vcb); // entirely my interior
    const bool col_edge = col_int && (16 * lane < 2 * A.HB || 16 * lane + 16 > vcb);
    // Row masks of this thread's band (bit i = band row i), fixed for the whole run:
    //   m_publish : rows/chunks whose current value the eight neighbours' halos will read
    //   m_reload  : everything that is not entirely mine (a chunk straddling the tile edge holds
    //               some neighbour bytes: it is re-read whole — my own bytes come back from the strip
    //               I just published)
    //   m_final   : my interior
    uint32_t m_publish = 0, m_reload = 0, m_final = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int rr = warp * R + i;
        const bool row_int = rr >= k && rr < k + vr;
        const bool row_edge = rr < 2 * k || rr >= vr;
        if (row_int && col_int && (row_edge || col_edge)) m_publish |= 1u << i;
        if (!(row_int && col_full)) m_reload |= 1u << i;
        if (row_int && col_int) m_final |= 1u << i;
    }

    // Phase probe: one interior CTA accumulates where its time goes (publish / flag+wait / reload /
    // compute) into status[4..8] — the "exposed halo-exchange time" of the run, measured on the
    // device with %globaltimer.
    const bool probe = A.sync.status != nullptr && threadIdx.x == 0 &&
                       blockIdx.x == (unsigned)((A.nty / 2) * A.ntx + A.ntx / 2);
    uint64_t t_pub = 0, t_wait = 0, t_load = 0, t_comp = 0, t_mark = probe ? globaltimer_ns() : 0;
    auto lap = [&](uint64_t& acc) {
        if (probe) { const uint64_t t = globaltimer_ns(); acc += t - t_mark; t_mark = t; }
    };

    const int nblocks = (A.loops + k - 1) / k;
    for (int b = 0; b < nblocks; ++b) {
        if (b > 0) {
            const int par = b & 1;
            uint8_t* mine = A.self[par];
            // 1. publish the border strips of my interior (what the eight neighbours' halos read)
            store_band<R>(rows, m_publish, mine, g, grw, gb);
            __syncthreads();
            lap(t_pub);
            // 2. announce epoch b: own flag + mirrored ring entries on neighbouring GPUs
            const uint32_t epoch = A.epoch_base + (uint32_t)b;
            // The release store alone publishes the strips: the CTA's stores happen-before it through
            // the barrier above (causality is cumulative), so no separate MEMBAR.SC is issued.
            if (threadIdx.x == 0)     // local consumers: gpu scope is enough (and much cheaper)
                st_release_gpu(A.cta_flags + (ti + 1) * fw + (tj + 1), epoch);
            if (threadIdx.x < 9 && threadIdx.x != 4) {
                const int sy = (int)threadIdx.x / 3 - 1, sx = (int)threadIdx.x % 3 - 1;
                const bool on_y = sy == 0 || (sy < 0 ? ti == 0 : ti == A.nty - 1);
                const bool on_x = sx == 0 || (sx < 0 ? tj == 0 : tj == A.ntx - 1);
                uint32_t* remote = A.remote_cta_flags[threadIdx.x];
                if (on_y && on_x && remote != nullptr)
                    st_release_sys(remote + (ti + 1 - sy * A.nty) * fw + (tj + 1 - sx * A.ntx), epoch);
            }
            // 3. wait for the eight surrounding CTAs (local, or ring entries written by a neighbour GPU)
            if (threadIdx.x >= 32 && threadIdx.x < 41 && threadIdx.x != 36) {
                const int d = (int)threadIdx.x - 32;
                const int ei = ti + d / 3, ej = tj + d % 3;  // extended-grid coordinates (+1 -1 folded)
                const int sy = ei < 1 ? 0 : (ei > A.nty ? 2 : 1), sx = ej < 1 ? 0 : (ej > A.ntx ? 2 : 1);
                const bool exists = (sy == 1 && sx == 1) || A.src[0].p[sy * 3 + sx] != nullptr;
                if (exists) {
                    const uint32_t* fl = A.cta_flags + ei * fw + ej;
                    const uint64_t ns = (sy == 1 && sx == 1)
                        ? wait_flag<false>(fl, epoch, A.sync.timeout_ns, A.sync.status)   // local CTA
                        : wait_flag<true>(fl, epoch, A.sync.timeout_ns, A.sync.status);   // other GPU
                    atomicMax(&s_wait_ns, (unsigned long long)ns);
                }
            }
            __syncthreads();
            lap(t_wait);
            // 4. re-load my halo (whole rows above/below the interior, edge chunks of interior rows)
            if (m_reload) load_band<R, true>(rows, g, A.src[par], grw, gb, m_reload);
            if (probe) { asm volatile("" ::"r"(rows[0].x), "r"(rows[R - 1].x)); }
            lap(t_load);
        }
        // 5. iterate
        const int kb = min(k, A.loops - b * k);
        if (masked) region_iterate<KIND, C, R, true>(rows, eb, kb, cmask, rowmask, need_cols, A.f);
        else region_iterate<KIND, C, R, false>(rows, eb, kb, cmask, rowmask, need_cols, A.f);
        __syncthreads();  // edge buffers: next block restarts at parity 0
        lap(t_comp);
    }
    if (probe) {
        unsigned long long* st = reinterpret_cast<unsigned long long*>(A.sync.status);
        atomicAdd(st + 4, (unsigned long long)t_pub);
        atomicAdd(st + 5, (unsigned long long)t_wait);
        atomicAdd(st + 6, (unsigned long long)t_load);
        atomicAdd(st + 7, (unsigned long long)t_comp);
        atomicAdd(st + 8, (unsigned long long)(nblocks - 1));
    }

    // ---- final interior -> the parity without live strips -------------------------------------------
    store_band<R>(rows, m_final, A.self[nblocks & 1], g, grw, gb);
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

using ResidentFn = void (*)(const ResidentArgs);

#define PCONV_FOR_R_SEP(X) X(2) X(4) X(7) X(10) X(14)
#define PCONV_FOR_R_INT(X) X(2) X(4) X(7) X(10)
#define PCONV_FOR_R_F32(X) X(2) X(4) X(6) X(8) X(10)

template <int KIND, int C>
static ResidentFn pick_r(int R) {
#define CASE(RR) if (R == RR) return resident_kernel<KIND, C, RR>;
    if constexpr (KIND == PCONV_KIND_F32) { PCONV_FOR_R_F32(CASE) }
    else if constexpr (KIND == PCONV_KIND_INT) { PCONV_FOR_R_INT(CASE) }
    else { PCONV_FOR_R_SEP(CASE) }
#undef CASE
    return nullptr;
}

static ResidentFn pick_kernel(int kind, int C, int R) {
    switch (kind * 8 + C) {
        case PCONV_KIND_SEP121 * 8 + 1: return pick_r<PCONV_KIND_SEP121, 1>(R);
        case PCONV_KIND_SEP121 * 8 + 3: return pick_r<PCONV_KIND_SEP121, 3>(R);
        case PCONV_KIND_INT * 8 + 1: return pick_r<PCONV_KIND_INT, 1>(R);
        case PCONV_KIND_INT * 8 + 3: return pick_r<PCONV_KIND_INT, 3>(R);
        case PCONV_KIND_F32 * 8 + 1: return pick_r<PCONV_KIND_F32, 1>(R);
        case PCONV_KIND_F32 * 8 + 3: return pick_r<PCONV_KIND_F32, 3>(R);
        default: return nullptr;
    }
}

}  // namespace pconv

using namespace pconv;

extern "C" int pconv_resident_launch(const pconv_resident_params_t* p, void* stream) {
    if (!p) return pconv_rt::fail("null params");
    const pconv_geometry_t& g = p->g;
    if (p->loops < 1 || p->k < 1 || p->TH < 1 || p->TWB < 16 || (p->TWB & 15) || (p->HB & 15) ||
        p->HB < p->k * g.C || p->TWB + 2 * p->HB > kSP || p->TH + 2 * p->k > kWarps * p->R ||
        p->TH < p->k || p->TWB < p->HB)
        return pconv_rt::fail("bad resident geometry (loops/k/TH/TWB/HB/R)");
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1], tile_cb = g.cb_lo[2] - g.cb_lo[1];
    if ((int64_t)p->nty * p->TH < tile_rows || (int64_t)p->ntx * p->TWB < tile_cb)
        return pconv_rt::fail("CTA grid does not cover the tile");
    const int grid = p->nty * p->ntx;
    if (grid < 1 || grid > pconv_rt::sm_count()) return pconv_rt::fail("resident grid exceeds the SM count");
    for (int s = 0; s < 3; s += 2) {
        const int64_t nr = g.row_lo[s + 1] - g.row_lo[s], nc = g.cb_lo[s + 1] - g.cb_lo[s];
        if ((nr > 0 && nr < p->k) || (nc > 0 && nc < (int64_t)p->k * g.C))
            return pconv_rt::fail("halo depth exceeds a neighbouring tile");
    }
    ResidentFn fn = pick_kernel(p->filt.kind, g.C, p->R);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height) combination");

    ResidentArgs a;
    std::memset(&a, 0, sizeof(a));
    a.g = g;
    for (int d = 0; d < 9; ++d) {
        a.src[0].p[d] = p->src_a[d]; a.src[0].pitch[d] = g.pitch;
        a.src[1].p[d] = p->src_b[d]; a.src[1].pitch[d] = g.pitch;
        a.remote_cta_flags[d] = p->remote_cta_flags[d];
    }
    a.self[0] = p->self_a; a.self[1] = p->self_b;
    pconv_rt::make_filter_dev(p->filt, &a.f);
    a.loops = p->loops; a.k = p->k; a.TH = p->TH; a.TWB = p->TWB; a.HB = p->HB;
    a.nty = p->nty; a.ntx = p->ntx;
    a.cta_flags = p->cta_flags;
    a.epoch_base = p->epoch_base;
    a.sync = p->sync;
    if (a.sync.timeout_ns == 0) a.sync.timeout_ns = 10ull * 1000 * 1000 * 1000;

    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)fn, dim3(grid), dim3(kThreads), args, 0,
                                                (cudaStream_t)stream);
    pconv_rt::count_launch();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "resident_kernel launch");
    return 0;
}

extern "C" int pconv_resident_prepare(int kind, int C, int R) {
    ResidentFn fn = pick_kernel(kind, C, R);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height) combination");
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, (const void*)fn);
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "cudaFuncGetAttributes");
    return 0;
}
