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

}  // namespace

extern "C" int pconv_cuda_pack_rect(const uint8_t* tile, int64_t pitch, int64_t row0, int64_t byte0,
                                    int64_t rows, int64_t nbytes, uint8_t* dense, void* stream) {
    return launch(true, const_cast<uint8_t*>(tile), pitch, row0, byte0, rows, nbytes, dense, stream);
}
extern "C" int pconv_cuda_unpack_rect(uint8_t* tile, int64_t pitch, int64_t row0, int64_t byte0,
                                      int64_t rows, int64_t nbytes, const uint8_t* dense, void* stream) {
    return launch(false, tile, pitch, row0, byte0, rows, nbytes, const_cast<uint8_t*>(dense), stream);
}
