// Pack / unpack kernels for the NCCL baseline backend (parallel/halo_nccl.py): copy a
// rows x nbytes rectangle between a pitched tile and a dense staging buffer. The fused backend
// never runs these — it has no staging — which is the point of the comparison. They play the
// role of the MPI library's internal (un)packing of MPI_Type_vector columns
// (mpi/mpi_conv.c:80-82,172-173).
#include <cuda_runtime.h>

#include "pconv_cuda.h"
#include "runtime_internal.h"

namespace {

template <bool PACK>
__global__ void rect_copy_kernel(uint8_t* tile, long long pitch, long long row0, long long byte0,
                                 long long rows, long long nbytes, uint8_t* dense) {
    const long long total = rows * nbytes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / nbytes, c = i % nbytes;
        uint8_t* t = tile + (row0 + r) * pitch + byte0 + c;
        if (PACK) dense[i] = *t; else *t = dense[i];
    }
}

int launch(bool pack, uint8_t* tile, long long pitch, long long row0, long long byte0,
           long long rows, long long nbytes, uint8_t* dense, void* stream) {
    if (rows <= 0 || nbytes <= 0) return 0;
    const long long total = rows * nbytes;
    int blocks = (int)((total + 255) / 256);
    if (blocks > pconv_rt::sm_count() * 8) blocks = pconv_rt::sm_count() * 8;
    if (pack) rect_copy_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(tile, pitch, row0, byte0, rows, nbytes, dense);
    else rect_copy_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(tile, pitch, row0, byte0, rows, nbytes, dense);
    pconv_rt::count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "rect_copy_kernel");
    return 0;
}

// All strips of one exchange in ONE launch (blockIdx.y = strip), 128-bit copies where the rectangle
// allows it — what a tuned send/recv halo exchange does; the per-strip kernels above remain for tests.
struct RectBatch {
    pconv_rect_t r[8];
    int n;
};

template <bool PACK>
__global__ void rect_batch_kernel(const RectBatch B) {
    const pconv_rect_t& q = B.r[blockIdx.y];
    uint8_t* tile = q.tile + q.row0 * q.pitch + q.byte0;
    const bool vec = ((q.nbytes | q.pitch) & 15) == 0 && ((reinterpret_cast<uintptr_t>(tile) | reinterpret_cast<uintptr_t>(q.dense)) & 15) == 0;
    if (vec) {
        const long long n16 = q.nbytes >> 4, total = q.rows * n16;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / n16, c = i % n16;
            uint4* t = reinterpret_cast<uint4*>(tile + r * q.pitch) + c;
            uint4* d = reinterpret_cast<uint4*>(q.dense) + i;
            if (PACK) *d = *t; else *t = *d;
        }
    } else {
        const long long total = q.rows * q.nbytes;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / q.nbytes, c = i % q.nbytes;
            uint8_t* t = tile + r * q.pitch + c;
            if (PACK) q.dense[i] = *t; else *t = q.dense[i];
        }
    }
}

}  // namespace

extern "C" int pconv_cuda_rect_batch(const pconv_rect_t* rects, int n, int pack, void* stream) {
    if (n <= 0) return 0;
    if (!rects || n > 8) return pconv_rt::fail("rect batch: 1..8 rectangles");
    RectBatch b;
    long long most = 0;
    for (int i = 0; i < n; ++i) {
        b.r[i] = rects[i];
        if (rects[i].rows < 0 || rects[i].nbytes < 0) return pconv_rt::fail("rect batch: negative extent");
        most = most > rects[i].rows * rects[i].nbytes ? most : rects[i].rows * rects[i].nbytes;
    }
    b.n = n;
    int bx = (int)((most / 16 + 255) / 256);
    bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
    const dim3 grid(bx, n);
    if (pack) rect_batch_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(b);
    else rect_batch_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(b);
    pconv_rt::count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return pconv_rt::fail_cuda(e, "rect_batch_kernel");
    return 0;
}

extern "C" int pconv_cuda_pack_rect(const uint8_t* tile, int64_t pitch, int64_t row0, int64_t byte0,
                                    int64_t rows, int64_t nbytes, uint8_t* dense, void* stream) {
    return launch(true, const_cast<uint8_t*>(tile), pitch, row0, byte0, rows, nbytes, dense, stream);
}
extern "C" int pconv_cuda_unpack_rect(uint8_t* tile, int64_t pitch, int64_t row0, int64_t byte0,
                                      int64_t rows, int64_t nbytes, const uint8_t* dense, void* stream) {
    return launch(false, tile, pitch, row0, byte0, rows, nbytes, const_cast<uint8_t*>(dense), stream);
}
