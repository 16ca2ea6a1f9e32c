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
// A CTA region (16*R rows x 512 B) is loaded from global memory straight into registers, iterated
// k times there (stencil_core.cuh) and its interior stored straight back: shared memory only
// carries two boundary rows per warp per iteration.
// The persistent grid (one CTA per SM) walks the CTA regions of the tile round-robin.
// Replaces, for the GPU path, the reference's launch-per-iteration thread-per-pixel kernels
// (cuda/cuda_convolute.cu:66-87) and, for the multi-GPU path, its MPI iteration loop
// (mpi/mpi_conv.c:164-248).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>

#include "stream_kernel.cuh"

namespace pconv {

__global__ void sync_kernel(const pconv_sync_t S) {
    if (S.wait_value != 0u && threadIdx.x < 9 && threadIdx.x != 4 && S.remote_flags[threadIdx.x] != nullptr) {
        const uint64_t ns = wait_flag(S.local_flags + threadIdx.x, S.wait_value, S.timeout_ns, S.status);
        if (ns) record_wait(S.status, ns);
    }
    __syncthreads();
    if (S.signal_value != 0u && threadIdx.x == 0 && !S.inject_drop_signal) {
        __threadfence_system();
#pragma unroll
        for (int d = 0; d < 9; ++d)
            if (d != 4 && S.remote_flags[d] != nullptr) st_relaxed_sys(S.remote_flags[d], S.signal_value);
    }
}

static KernelFn pick_kernel(int kind, int C, int R) {
    switch (kind) {
        case PCONV_KIND_SEP121: return pick_stream_sep(C, R);
        case PCONV_KIND_INT: return pick_stream_int(C, R);
        case PCONV_KIND_F32: return pick_stream_f32(C, R);
        default: return nullptr;
    }
}

}  // namespace pconv

using namespace pconv;

// Band heights available for a filter kind (rows per warp; region rows = 16 * R).
extern "C" int pconv_stream_band_heights(int kind, int* out, int cap) {
    int n = 0;
#define PUSH(RR) { if (n < cap) out[n] = RR; ++n; }
    if (kind == PCONV_KIND_F32) { PCONV_FOR_R_F32(PUSH) }
    else if (kind == PCONV_KIND_INT) { PCONV_FOR_R_INT(PUSH) }
    else { PCONV_FOR_R_SEP(PUSH) }
#undef PUSH
    return n;
}

extern "C" int pconv_stream_launch(const pconv_stream_params_t* p, void* stream) {
    if (!p) return pconv_rt::fail("null params");
    const pconv_geometry_t& g = p->g;
    if (p->k < 1 || p->TH < 1 || p->TWB < 16 || (p->TWB & 15) || (p->HB & 15) ||
        p->HB < p->k * g.C || p->TWB + 2 * p->HB > kSP || p->TH + 2 * p->k > kWarps * p->R)
        return pconv_rt::fail("bad region geometry (k/TH/TWB/HB/R)");
    const int64_t tile_rows = g.row_lo[2] - g.row_lo[1], tile_cb = g.cb_lo[2] - g.cb_lo[1];
    if (tile_rows <= 0 || tile_cb <= 0) return pconv_rt::fail("empty tile");
    constexpr int64_t kMaxPitch = (int64_t)1 << 30;   // region_io.cuh does 32-bit row * pitch products
    if (g.pitch <= 0 || g.pitch >= kMaxPitch) return pconv_rt::fail("tile pitch out of range");
    for (int d = 0; d < 9; ++d)
        if (p->src_pitch[d] < 0 || p->src_pitch[d] >= kMaxPitch) return pconv_rt::fail("source pitch out of range");
    // the halo may only reach into the *adjacent* tiles
    for (int s = 0; s < 3; s += 2) {
        const int64_t nr = g.row_lo[s + 1] - g.row_lo[s], nc = g.cb_lo[s + 1] - g.cb_lo[s];
        if ((nr > 0 && nr < p->k) || (nc > 0 && nc < (int64_t)p->k * g.C))
            return pconv_rt::fail("halo depth exceeds a neighbouring tile");
    }
    KernelFn fn = pick_kernel(p->filt.kind, g.C, p->R);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height) combination");

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

    const int ntiles = a.ntx * a.nty;
    int grid = p->grid > 0 ? p->grid : pconv_rt::sm_count();
    if (grid > ntiles) grid = ntiles;
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pconv_rt::pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, a);
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "stream_kernel launch");
    pconv_rt::count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "stream_kernel launch");
    return 0;
}

extern "C" int pconv_stream_run(const pconv_stream_params_t* pp, int nranks, const int32_t* ks,
                                int nlaunch, uint32_t epoch0, int use_flags, uint32_t first_wait,
                                uint32_t* const* first_flags, void* stream) {
    if (!pp || !ks || nranks < 1 || nlaunch < 0) return pconv_rt::fail("bad run arguments");
    for (int i = 0; i < nlaunch; ++i) {
        for (int r = 0; r < nranks; ++r) {
            pconv_stream_params_t p = pp[2 * r + (i & 1)];
            p.k = ks[i];
            if (use_flags) {
                p.sync.signal_value = epoch0 + (uint32_t)i + 1u;
                p.sync.wait_value = i ? epoch0 + (uint32_t)i : first_wait;
                if (i == 0 && first_wait && first_flags && first_flags[r]) p.sync.local_flags = first_flags[r];
            } else {
                p.sync.signal_value = p.sync.wait_value = 0u;
            }
            const int rc = pconv_stream_launch(&p, stream);
            if (rc != 0) return rc;
        }
    }
    return 0;
}

extern "C" int pconv_sync_launch(const pconv_sync_t* s, void* stream) {
    if (!s) return pconv_rt::fail("null sync");
    pconv_sync_t a = *s;
    if (a.timeout_ns == 0) a.timeout_ns = 10ull * 1000 * 1000 * 1000;
    sync_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    pconv_rt::count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "sync_kernel launch");
    return 0;
}

extern "C" int pconv_kernel_prepare(int kind, int C, int R) {
    KernelFn fn = pick_kernel(kind, C, R);
    if (!fn) return pconv_rt::fail("unsupported (filter kind, channels, band height) combination");
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, (const void*)fn);
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "cudaFuncGetAttributes");
    return 0;
}
