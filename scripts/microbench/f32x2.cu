// Microbenchmark: issue rate of packed FP32 (FADD2 / FMUL2 / FFMA2) against scalar FADD / FMUL on
// sm_100a, to decide whether the exact-fp32 path (separately rounded mul and add) should move to
// f32x2 operations. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kChains = 8;

template <int MODE>
__global__ void bench(float* out, unsigned long long* cycles, float seed) {
    // kChains independent dependency chains per thread so the pipe, not the latency, is the limit
    float a[kChains], b[kChains];
    unsigned long long p[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = seed + (float)(threadIdx.x + i);
        b[i] = seed * 0.5f + (float)i;
        asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
    }
    const float c = seed * 1.0001f;
    unsigned long long cc;
    asm("mov.b64 %0, {%1,%1};" : "=l"(cc) : "f"(c));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (MODE == 0) { a[i] = __fadd_rn(a[i], c); b[i] = __fadd_rn(b[i], c); }          // 2 FADD
            if (MODE == 1) { a[i] = __fmul_rn(a[i], c); b[i] = __fmul_rn(b[i], c); }          // 2 FMUL
            if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));   // 1 FADD2
            if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));   // 1 FMUL2
            if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc)); // 1 FFMA2
            if (MODE == 5) { a[i] = __fmaf_rn(a[i], c, c); b[i] = __fmaf_rn(b[i], c, c); }    // 2 FFMA
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        float x, y;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i]));
        s += a[i] + b[i] + x + y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int elems_per_op) {
    float* out; unsigned long long* cyc;
    const int blocks = 148, threads = 512;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(unsigned long long));
    bench<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    bench<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
    const double pairs = (double)kIters * kChains;            // each "pair step" produces 2 fp32 results
    const double warps = threads / 32.0;
    // fp32 results per clock per SM
    printf("%-8s cycles=%.0f  fp32 results/clk/SM=%.1f  (%d warps/SM, %s)\n", name, avg,
           pairs * 2 * threads / avg, (int)warps, elems_per_op == 2 ? "1 packed instr per pair" : "2 scalar instrs per pair");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FADD", 1); run<1>("FMUL", 1); run<5>("FFMA", 1);
    run<2>("FADD2", 2); run<3>("FMUL2", 2); run<4>("FFMA2", 2);
    return 0;
}
