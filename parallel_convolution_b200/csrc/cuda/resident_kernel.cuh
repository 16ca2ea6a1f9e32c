// Resident kernel template (see resident_kernel.cu for the protocol description); one translation
// unit per arithmetic path instantiates it (resident_kind_*.cu) so the build parallelises.
#pragma once

#include <cuda_runtime.h>

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

#define PCONV_RES_R_SEP(X) X(2) X(4) X(7) X(10) X(14)
#define PCONV_RES_R_INT(X) X(2) X(4) X(7) X(10)
#define PCONV_RES_R_F32(X) X(2) X(4) X(6) X(8) X(10)

template <int KIND, int C>
static ResidentFn pick_resident_r(int R) {
#define CASE(RR) if (R == RR) return resident_kernel<KIND, C, RR>;
    if constexpr (KIND == PCONV_KIND_F32) { PCONV_RES_R_F32(CASE) }
    else if constexpr (KIND == PCONV_KIND_INT) { PCONV_RES_R_INT(CASE) }
    else { PCONV_RES_R_SEP(CASE) }
#undef CASE
    return nullptr;
}

template <int KIND>
static ResidentFn pick_resident_kind(int C, int R) {
    switch (C) {
        case 1: return pick_resident_r<KIND, 1>(R);
        case 3: return pick_resident_r<KIND, 3>(R);
        case 4: return pick_resident_r<KIND, 4>(R);
        default: return nullptr;
    }
}

ResidentFn pick_resident_f32(int C, int R);
ResidentFn pick_resident_int(int C, int R);
ResidentFn pick_resident_sep(int C, int R);

}  // namespace pconv
