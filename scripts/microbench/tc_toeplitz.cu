// Would the 5th-generation tensor cores help a 3x3 uint8 stencil? A bounded measurement instead of prose.
//
// The only GEMM in this workload is artificial: the vertical [1 2 1] pass of the Gaussian as a banded
// (Toeplitz) matrix product  D[128 x N] (s32, TMEM) = T[128 x K] (u8 weights) * X[K x N] (u8 image rows),
// K = 128 + 2 halo rows rounded up to the MMA granule 32 -> 160, i.e. FIVE tcgen05.mma.kind::i8
// (M=128, N=256, K=32) per 128 x 256-byte output tile, 160 multiply-adds per output byte for a 3-tap sum.
// The horizontal pass cannot follow on the tensor cores (its input is 10-bit), so every s32 accumulator has
// to come back through tcgen05.ld (4 bytes of TMEM traffic per output byte) for ~5 CUDA-core ops per byte —
// more than the ~3.75 ops per byte the whole SWAR kernel spends on BOTH passes.
//
// This program measures the two tensor-side terms on B200: (a) cycles per tcgen05.mma.kind::i8 m128n256k32
// issued back to back by one thread per CTA, on every SM at once; (b) cycles for the four epilogue warps to
// read one 128 x 256 accumulator tile with tcgen05.ld. Output: the bytes/clk/SM ceiling each term puts on
// the vertical pass alone, next to the SWAR kernel's measured ~23 B/clk/SM for the complete filter.
// Operand contents are irrelevant for the timing (the descriptors are valid, the data is noise).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tc_toeplitz.cu -o tc_toeplitz
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// kind::i8 instruction descriptor: D = s32, A = B = u8, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, no swizzle, K-major: 8 x 16-byte core matrices, LBO between the two
// 16-byte K chunks, SBO between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

constexpr int kM = 128, kN = 256, kKB = 32;     // one MMA: 128 x 256 x 32 (bytes of K)
constexpr int kMmas = 640;                      // 128 output tiles' worth (5 MMAs each)

__global__ void __launch_bounds__(128, 1) tc_kernel(unsigned long long* out) {
    __shared__ __align__(1024) uint8_t sA[kM * kKB];     // 4 KB
    __shared__ __align__(1024) uint8_t sB[kN * kKB];     // 8 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kM * kKB; i += 128) sA[i] = (uint8_t)(i * 7 + 1);
    for (int i = threadIdx.x; i < kN * kKB; i += 128) sB[i] = (uint8_t)(i * 13 + 5);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");      // operand bytes -> visible to the tensor-core proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const uint64_t adesc = make_desc(smem_u32(sA), kM * 16, 128), bdesc = make_desc(smem_u32(sB), kN * 16, 128);
    const uint32_t idesc = make_idesc(kM, kN);

    long long t0 = 0, t1 = 0, t2 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int i = 0; i < kMmas; ++i) {
            const uint32_t acc = (i % 5) ? 1u : 0u;               // five K chunks accumulate into one tile
            const uint32_t d = tmem + (uint32_t)((i / 5) & 1) * 256u;   // two accumulator tiles, alternating
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    // everybody waits for the last MMA
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x == 0) t1 = clock64();
    // epilogue: warp w reads TMEM lanes [32w, 32w+32) x 256 columns, 32 columns per instruction
    uint32_t chk = 0;
    for (int rep = 0; rep < 16; ++rep) {
#pragma unroll
        for (int c = 0; c < kN; c += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) chk ^= v[j];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        t2 = clock64();
        out[blockIdx.x * 4 + 0] = (unsigned long long)(t1 - t0);
        out[blockIdx.x * 4 + 1] = (unsigned long long)(t2 - t1);
    }
    if (lane == 0) atomicXor(&out[blockIdx.x * 4 + 2], (unsigned long long)chk);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned long long* out;
    cudaMalloc(&out, sizeof(unsigned long long) * 4 * sms);
    cudaMemset(out, 0, sizeof(unsigned long long) * 4 * sms);
    for (int rep = 0; rep < 3; ++rep) tc_kernel<<<sms, 128>>>(out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    unsigned long long* h = new unsigned long long[4 * sms];
    cudaMemcpy(h, out, sizeof(unsigned long long) * 4 * sms, cudaMemcpyDeviceToHost);
    double mma = 0, epi = 0;
    for (int i = 0; i < sms; ++i) { mma += (double)h[4 * i]; epi += (double)h[4 * i + 1]; }
    mma /= sms; epi /= sms;
    const double clk_per_mma = mma / kMmas, clk_per_tile_epi = epi / 16.0;
    const double tile_bytes = 128.0 * 256.0;
    printf("{\"mma\": \"tcgen05.mma.cta_group::1.kind::i8 m128n256k32\", \"sms\": %d, \"clk_per_mma\": %.1f, "
           "\"macs_per_clk_per_sm\": %.0f, \"vertical_pass_bytes_per_clk_per_sm_mma_only\": %.1f, "
           "\"tcgen05_ld_clk_per_128x256_tile\": %.0f, \"vertical_pass_bytes_per_clk_per_sm_tmem_read_only\": %.1f, "
           "\"swar_kernel_bytes_per_clk_per_sm_both_passes\": 23.4}\n",
           sms, clk_per_mma, 128.0 * 256.0 * 32.0 / clk_per_mma, tile_bytes / (5.0 * clk_per_mma),
           clk_per_tile_epi, tile_bytes / clk_per_tile_epi);
    return 0;
}
