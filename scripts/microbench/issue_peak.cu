// Measured issue peak of the Gaussian SWAR kernel's exact instruction mix (the denominator of the
// roofline fraction bench.py reports): band_sep121 from csrc/cuda/stencil_core.cuh — the same code the
// kernels run — iterated on registers only: no shared-memory edge exchange, no barrier, no global memory,
// 16 warps per SM (one 512-thread CTA at 128 registers, exactly the kernels' occupancy).
// Output: region bytes per clock per SM at the sampled SM clock. The production kernels reach ~23 B/clk/SM;
// the modelled dual-pipe bound is 30-34 (csrc/host/planner.cpp, utils/roofline.py).
//   nvcc -O3 -std=c++17 --fmad=false -gencode arch=compute_100a,code=sm_100a -I../../parallel_convolution_b200/csrc/cuda issue_peak.cu -o issue_peak
#include <cstdio>
#include <cuda_runtime.h>
#include "stencil_core.cuh"

using namespace pconv;

template <int C, int R>
__global__ void __launch_bounds__(512, 1) peak_kernel(uint4* out, int iters, unsigned seed) {
    uint4 rows[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const unsigned v = seed * 2654435761u + threadIdx.x * 97u + i * 131u + blockIdx.x;
        rows[i] = make_uint4(v, v * 3u, v * 5u, v * 7u);
    }
    FilterDev f{};
    const uint4 cmask = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (int it = 0; it < iters; ++it) {
        // the rows a neighbouring warp would deliver: any register values do (timing only)
        const uint4 top = rows[R - 1], bottom = rows[0];
        band_sep121<C, R, kMaskNone>(rows, top, bottom, cmask, 0xffffffffu, false, f);
    }
    uint4 acc = rows[0];
#pragma unroll
    for (int i = 1; i < R; ++i) { acc.x ^= rows[i].x; acc.y ^= rows[i].y; acc.z ^= rows[i].z; acc.w ^= rows[i].w; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int C, int R>
void run(const char* name, int sms, double mhz) {
    uint4* out;
    cudaMalloc(&out, sizeof(uint4) * sms * 512);
    const int iters = 2000;
    peak_kernel<C, R><<<sms, 512>>>(out, 10, 1u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        peak_kernel<C, R><<<sms, 512>>>(out, iters, rep + 2u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double bytes_per_sm = 512.0 * R * 16.0 * iters;
    const double clk = best * 1e-3 * mhz * 1e6;
    printf("{\"kernel\": \"%s\", \"R\": %d, \"C\": %d, \"ms\": %.4f, \"bytes_per_clk_per_sm\": %.2f, \"sm_mhz\": %.0f, \"err\": \"%s\"}\n",
           name, R, C, best, bytes_per_sm / clk, mhz, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double mhz = khz / 1e3;
    run<3, 14>("sep121 rgb", sms, mhz);
    run<1, 14>("sep121 grey", sms, mhz);
    run<3, 7>("sep121 rgb", sms, mhz);
    return 0;
}
