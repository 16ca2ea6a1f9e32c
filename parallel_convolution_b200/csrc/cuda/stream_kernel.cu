// Streaming kernel: k fused stencil iterations per launch over a tile of any size, with the halo
// exchange fused in (regions that overlap a neighbouring tile are loaded through that tile's
// peer-mapped pointer, region_io.cuh) and a launch-granular cross-GPU flag protocol:
//
//   start of launch n : CTAs whose region touches neighbour d spin on local_flags[d] >= n-1
//                        (RAW: d finished writing the buffer I am about to read; WAR: d finished
//                        reading the buffer I am about to overwrite — the same CTAs, see below)
//   end of launch n   : the last CTA to finish (acq_rel ticket) stores n to every neighbour's flag
//                        slot with st.release.sys.
//
// Launch n reads buffer X and writes buffer Y; launch n+1 swaps them. A neighbour only ever reads
// the k-deep border of my tile, and the CTAs that write that border are exactly the CTAs whose
// input region reaches into the neighbour, i.e. the ones that wait — interior CTAs never wait.
// No NCCL, no host round trip between launches; the host enqueues all launches of a run at once.
//
// The persistent grid (one CTA per SM) walks the CTA regions of the tile round-robin.
// Replaces, for the GPU path, the reference's launch-per-iteration thread-per-pixel kernels
// (cuda/cuda_convolute.cu:66-87) and, for the multi-GPU path, its MPI iteration loop
// (mpi/mpi_conv.c:164-248).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>

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

template <int KIND, int C>
__global__ void __launch_bounds__(kThreads, 1)
stream_kernel(const __grid_constant__ StreamArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t s_wait_ns;
    const pconv_geometry_t& g = A.g;
    const int k = A.k, RH = A.TH + 2 * A.k;
    uint8_t* buf0 = smem;
    uint8_t* buf1 = smem + (size_t)RH * kSP;
    const int ntiles = A.ntx * A.nty;
    const int64_t tile_row_end = g.row_lo[2], tile_cb_end = g.cb_lo[2];

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ti = tile / A.ntx, tj = tile % A.ntx;
        const int64_t gr0 = g.row_lo[1] + (int64_t)ti * A.TH - k;     // image row of region row 0
        const int64_t gb0 = g.cb_lo[1] + (int64_t)tj * A.TWB - A.HB;  // image byte of region byte 0

        // ---- halo protocol: wait for the neighbours this region reaches into --------------------
        if (A.sync.wait_value != 0u) {
            if (threadIdx.x == 0) s_wait_ns = 0;
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
                    atomicMax(reinterpret_cast<unsigned long long*>(&s_wait_ns), (unsigned long long)ns);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0 && s_wait_ns) record_wait(A.sync.status, s_wait_ns);
        }

        // ---- load region (+ zero rows outside the image in both buffers) ------------------------
        load_rect(buf0, buf1, g, A.src, gr0, gb0, 0, RH, 0, kChunksPerRow);
        bool need_mask;
        const uint4 mask = column_mask(g, gb0, &need_mask);
        // region rows that lie inside the image
        const int img_lo = (int)max((int64_t)0, -gr0);
        const int img_hi = (int)min((int64_t)RH, g.img_rows - gr0);
        __syncthreads();

        // ---- k fused iterations, ping-pong in shared memory -------------------------------------
        for (int it = 1; it <= k; ++it) {
            const uint8_t* s_src = (it & 1) ? buf0 : buf1;
            uint8_t* s_dst = (it & 1) ? buf1 : buf0;
            const int lo = max(max(it, img_lo), 1);
            const int hi = min(min(RH - it, img_hi), RH - 1);
            if (lo < hi) iterate_region<KIND, C>(s_src, s_dst, lo, hi, mask, need_mask, A.f);
            __syncthreads();
        }

        // ---- store the interior ------------------------------------------------------------------
        const uint8_t* fin = (k & 1) ? buf1 : buf0;
        store_rect(fin, A.dst, g, gr0, gb0, k, k + A.TH, A.HB / 16, (A.HB + A.TWB) / 16);
        __syncthreads();
    }

    // ---- halo protocol: publish "launch done" to the neighbours -----------------------------------
    if (A.sync.signal_value != 0u) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const uint32_t ticket = atom_add_acq_rel_gpu(A.sync.done_counter, 1u);
            if (ticket == gridDim.x - 1) {
                *A.sync.done_counter = 0u;  // self-reset for the next launch (stream ordered)
                if (A.sync.inject_delay_ns > 0) {
                    const uint64_t t0 = globaltimer_ns();
                    while (globaltimer_ns() - t0 < (uint64_t)A.sync.inject_delay_ns) __nanosleep(256);
                }
                if (!A.sync.inject_drop_signal) {
                    __threadfence_system();
#pragma unroll
                    for (int d = 0; d < 9; ++d)
                        if (d != 4 && A.sync.remote_flags[d] != nullptr)
                            st_release_sys(A.sync.remote_flags[d], A.sync.signal_value);
                }
            }
        }
    }
}

using KernelFn = void (*)(const StreamArgs);

template <int KIND, int C>
static KernelFn get_kernel() { return stream_kernel<KIND, C>; }

static KernelFn pick_kernel(int kind, int C) {
    switch (kind * 8 + C) {
        case PCONV_KIND_SEP121 * 8 + 1: return get_kernel<PCONV_KIND_SEP121, 1>();
        case PCONV_KIND_SEP121 * 8 + 3: return get_kernel<PCONV_KIND_SEP121, 3>();
        case PCONV_KIND_INT * 8 + 1: return get_kernel<PCONV_KIND_INT, 1>();
        case PCONV_KIND_INT * 8 + 3: return get_kernel<PCONV_KIND_INT, 3>();
        case PCONV_KIND_F32 * 8 + 1: return get_kernel<PCONV_KIND_F32, 1>();
        case PCONV_KIND_F32 * 8 + 3: return get_kernel<PCONV_KIND_F32, 3>();
        default: return nullptr;
    }
}

}  // namespace pconv

using namespace pconv;

extern "C" int pconv_stream_max_region_rows(int ctas_per_sm) {
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int budget = (pconv_rt::max_optin_smem() - 1024) / ctas_per_sm - 1024;
    return budget / (2 * kSP);
}

extern "C" int pconv_stream_launch(const pconv_stream_params_t* p, void* stream) {
    if (!p) return pconv_rt::fail("null params");
    const pconv_geometry_t& g = p->g;
    if (p->k < 1 || p->TH < 1 || p->TWB < 16 || (p->TWB & 15) || (p->HB & 15) ||
        p->HB < p->k * g.C || p->TWB + 2 * p->HB > kSP)
        return pconv_rt::fail("bad region geometry (k/TH/TWB/HB)");
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1], tile_cb = g.cb_lo[2] - g.cb_lo[1];
    if (tile_rows <= 0 || tile_cb <= 0) return pconv_rt::fail("empty tile");
    // the halo may only reach into the *adjacent* tiles
    for (int s = 0; s < 3; s += 2) {
        const int64_t nr = g.row_lo[s + 1] - g.row_lo[s], nc = g.cb_lo[s + 1] - g.cb_lo[s];
        if ((nr > 0 && nr < p->k) || (nc > 0 && nc < (int64_t)p->k * g.C))
            return pconv_rt::fail("halo depth exceeds a neighbouring tile");
    }
    KernelFn fn = pick_kernel(p->filt.kind, g.C);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels) combination");

    StreamArgs a;
    std::memset(&a, 0, sizeof(a));
    a.g = g;
    for (int d = 0; d < 9; ++d) {
        a.src.p[d] = p->src[d];
        a.src.pitch[d] = p->src_pitch[d] ? p->src_pitch[d] : g.pitch;
    }
    a.dst = p->dst;
    pconv_rt::make_filter_dev(p->filt, &a.f);
    a.k = p->k; a.TH = p->TH; a.TWB = p->TWB; a.HB = p->HB;
    a.nty = (int)((tile_rows + p->TH - 1) / p->TH);
    a.ntx = (int)((tile_cb + p->TWB - 1) / p->TWB);
    a.sync = p->sync;
    if (a.sync.timeout_ns == 0) a.sync.timeout_ns = 10ull * 1000 * 1000 * 1000;

    const int RH = p->TH + 2 * p->k;
    const size_t smem = (size_t)2 * RH * kSP;
    if ((int)smem > pconv_rt::max_optin_smem() - 1024) return pconv_rt::fail("region does not fit in shared memory");
    cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "cudaFuncSetAttribute");
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fn, kThreads, smem);
    if (e != cudaSuccess || occ < 1) return pconv_rt::fail_cuda(e, "occupancy");
    const int ntiles = a.ntx * a.nty;
    int grid = p->grid > 0 ? p->grid : pconv_rt::sm_count() * occ;
    if (grid > ntiles) grid = ntiles;
    fn<<<grid, kThreads, smem, (cudaStream_t)stream>>>(a);
    pconv_rt::count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "stream_kernel launch");
    return 0;
}
