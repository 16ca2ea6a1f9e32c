// Streaming kernel template (see stream_kernel.cu for the protocol description). Lives in a header
// so that each arithmetic path is instantiated in its own translation unit (stream_kind_*.cu): the
// three paths x three channel strides x five band heights compile in parallel instead of in one
// three-minute nvcc invocation.
#pragma once

#include <cuda_runtime.h>

#include "pconv_cuda.h"
#include "region_io.cuh"
#include "runtime_internal.h"
#include "stencil_core.cuh"

namespace pconv {

struct StreamArgs {
    pconv_geometry_t g;
    SourceSet src;
    uint8_t* dst;
    FilterDev f;
    int k, TH, TWB, HB, ntx, nty;
    pconv_sync_t sync;
};

template <int KIND, int C, int R>
__global__ void __launch_bounds__(kThreads, 1)
stream_kernel(const __grid_constant__ StreamArgs A) {
    __shared__ EdgeBuffers eb;
    __shared__ unsigned long long s_wait_ns;
    const pconv_geometry_t& g = A.g;
    constexpr int RH = kWarps * R;
    const int k = A.k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = A.ntx * A.nty;
    const int64_t tile_row_end = g.row_lo[2], tile_cb_end = g.cb_lo[2];

    // Programmatic dependent launch: this grid may be scheduled while the previous launch of the
    // run is still draining (its CTAs occupy whole SMs, so ours start one by one as theirs exit);
    // nothing above touches memory the previous launch writes. Let the next launch do the same to
    // us, then wait for the previous grid to complete and flush before reading its output.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ti = tile / A.ntx, tj = tile % A.ntx;
        const int64_t gr0 = g.row_lo[1] + (int64_t)ti * A.TH - k;     // image row of region row 0
        const int64_t gb0 = g.cb_lo[1] + (int64_t)tj * A.TWB - A.HB;  // image byte of region byte 0

        // ---- halo protocol: wait for the neighbours this region reaches into --------------------
        if (A.sync.wait_value != 0u) {
            if (threadIdx.x == 0) s_wait_ns = 0ull;
            __syncthreads();
            if (threadIdx.x < 9 && threadIdx.x != 4) {
                const int dy = (int)threadIdx.x / 3, dx = (int)threadIdx.x % 3;
                const bool rows_hit = dy == 0 ? gr0 < g.row_lo[1]
                                     : dy == 2 ? gr0 + RH > tile_row_end : true;
                const bool cols_hit = dx == 0 ? gb0 < g.cb_lo[1]
                                     : dx == 2 ? gb0 + kSP > tile_cb_end : true;
                if (rows_hit && cols_hit && A.src.p[threadIdx.x] != nullptr) {
                    const uint64_t ns = wait_flag(A.sync.local_flags + threadIdx.x, A.sync.wait_value,
                                                  A.sync.timeout_ns, A.sync.status);
                    atomicMax(&s_wait_ns, (unsigned long long)ns);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0 && s_wait_ns) record_wait(A.sync.status, s_wait_ns);
        }

        // ---- load my band: R independent 128-bit loads per thread, straight into registers ------
        const int64_t grw = gr0 + (int64_t)warp * R;   // image row of my band row 0
        const int64_t gb = gb0 + 16 * lane;            // image byte of my strip
        uint4 rows[R];
        load_band<R>(rows, g, A.src, grw, gb);

        bool need_cols;
        const uint4 cmask = column_mask(g, gb0, &need_cols);
        uint32_t rowmask = 0;
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (grw + i >= 0 && grw + i < g.img_rows) rowmask |= 1u << i;

        // ---- k fused iterations in registers ------------------------------------------------------
        region_iterate<KIND, C, R>(rows, eb, k, cmask, rowmask, need_cols, A.f);

        // ---- store the interior straight from registers ------------------------------------------
        const bool col_in = (16 * lane >= A.HB) && (16 * lane < A.HB + A.TWB);
        uint32_t m_store = 0;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int rr = warp * R + i;
            if (col_in && rr >= k && rr < k + A.TH) m_store |= 1u << i;
        }
        store_band<R>(rows, m_store, A.dst, g, grw, gb);
        __syncthreads();  // edge buffers are reused by the next tile
    }

    // ---- halo protocol: publish "launch done" to the neighbours -----------------------------------
    if (A.sync.signal_value != 0u) {
        __syncthreads();
        if (threadIdx.x == 0) {   // acq_rel ticket: releases this CTA's stores, acquires the others'
            const uint32_t ticket = atom_add_acq_rel_gpu(A.sync.done_counter, 1u);
            if (ticket == gridDim.x - 1) {
                *A.sync.done_counter = 0u;  // self-reset for the next launch (stream ordered)
                if (A.sync.inject_delay_ns > 0) {
                    const uint64_t t0 = globaltimer_ns();
                    while (globaltimer_ns() - t0 < (uint64_t)A.sync.inject_delay_ns) __nanosleep(256);
                }
                if (!A.sync.inject_drop_signal) {
                    fence_acq_rel_sys();
#pragma unroll
                    for (int d = 0; d < 9; ++d)
                        if (d != 4 && A.sync.remote_flags[d] != nullptr)
                            st_relaxed_sys(A.sync.remote_flags[d], A.sync.signal_value);
                }
            }
        }
    }
}

using KernelFn = void (*)(const StreamArgs);

// Band heights compiled in. Larger R = taller region = less halo recompute; smaller R = more CTAs
// for small tiles (strong scaling). (register budget: 128/thread at 512 threads; the lists below are
// what fits without spilling in the hot loop)
#define PCONV_FOR_R_SEP(X) X(2) X(4) X(7) X(10) X(14)
#define PCONV_FOR_R_INT(X) X(2) X(4) X(7) X(10)
#define PCONV_FOR_R_F32(X) X(2) X(4) X(6) X(8) X(10)

template <int KIND, int C>
static KernelFn pick_stream_r(int R) {
#define CASE(RR) if (R == RR) return stream_kernel<KIND, C, RR>;
    if constexpr (KIND == PCONV_KIND_F32) { PCONV_FOR_R_F32(CASE) }
    else if constexpr (KIND == PCONV_KIND_INT) { PCONV_FOR_R_INT(CASE) }
    else { PCONV_FOR_R_SEP(CASE) }
#undef CASE
    return nullptr;
}

template <int KIND>
static KernelFn pick_stream_kind(int C, int R) {
    switch (C) {
        case 1: return pick_stream_r<KIND, 1>(R);
        case 3: return pick_stream_r<KIND, 3>(R);
        case 4: return pick_stream_r<KIND, 4>(R);
        default: return nullptr;
    }
}

// one definition per translation unit stream_kind_{f32,int,sep}.cu
KernelFn pick_stream_f32(int C, int R);
KernelFn pick_stream_int(int C, int R);
KernelFn pick_stream_sep(int C, int R);

}  // namespace pconv
