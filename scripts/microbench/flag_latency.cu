// Microbenchmark: the latency floor of one in-kernel halo exchange on one GPU.
//   1. flag ping-pong between two CTAs (st.release.gpu / ld.acquire.gpu): one-way signalling latency
//   2. the same with a payload: producer stores `bytes` (16 B per thread), barrier, release; consumer
//      acquires, then loads the payload: publish -> visible-data latency as the resident kernel pays it
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o flag_latency flag_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid = 2 CTAs * stride so that the two active CTAs land on different SMs (other CTAs idle-exit)
__global__ void pingpong(unsigned* flags, uint4* payload, int rounds, int payload_threads,
                         unsigned long long* cycles, int partner_block) {
    const bool is_a = blockIdx.x == 0, is_b = blockIdx.x == (unsigned)partner_block;
    if (!is_a && !is_b) return;
    unsigned* mine = flags + (is_a ? 0 : 32);
    unsigned* theirs = flags + (is_a ? 32 : 0);
    uint4* out = payload + (is_a ? 0 : 4096);
    const uint4* in = payload + (is_a ? 4096 : 0);
    uint4 acc = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r) {
        if (is_a) {
            if ((int)threadIdx.x < payload_threads) out[threadIdx.x] = make_uint4(r, r, r, r);
            __syncthreads();
            if (threadIdx.x == 0) st_release(theirs, (unsigned)r);
            if (threadIdx.x == 0) { long long spins = 0; while (ld_acquire(mine) < (unsigned)r && ++spins < (1ll << 24)) {} }
            __syncthreads();
            if ((int)threadIdx.x < payload_threads) { uint4 v = in[threadIdx.x]; acc.x += v.x; }
        } else {
            if (threadIdx.x == 0) { long long spins = 0; while (ld_acquire(mine) < (unsigned)r && ++spins < (1ll << 24)) {} }
            __syncthreads();
            if ((int)threadIdx.x < payload_threads) { uint4 v = in[threadIdx.x]; acc.x += v.x; }
            if ((int)threadIdx.x < payload_threads) out[threadIdx.x] = make_uint4(r, r, r, r);
            __syncthreads();
            if (threadIdx.x == 0) st_release(theirs, (unsigned)r);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && is_a) cycles[0] = (unsigned long long)(t1 - t0);
    if (acc.x == 0xffffffffu) cycles[1] = acc.x;
}

int main() {
    unsigned* flags; uint4* payload; unsigned long long* cyc;
    cudaMalloc(&flags, 256); cudaMalloc(&payload, 8192 * sizeof(uint4)); cudaMalloc(&cyc, 16);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int rounds = 2000;
    for (int partner : {1, 74, 147}) {                 // neighbouring SM / other die / far end
        for (int pt : {0, 32, 160, 512}) {             // payload: 0, 512 B, 2.5 KB, 8 KB
            cudaMemset(flags, 0, 256);
            pingpong<<<148, 512>>>(flags, payload, rounds, pt, cyc, partner);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("partner CTA %3d payload %5d B: round trip %.0f cycles = one-way %.2f us (at %d MHz nominal) %s\n",
                   partner, pt * 16, (double)h / rounds, (double)h / rounds / 2.0 / (clk_khz / 1e3), clk_khz / 1000,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
