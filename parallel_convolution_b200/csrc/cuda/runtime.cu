// Runtime around the kernels: device queries, IPC-exportable allocations for peer memory,
// pinned host memory, pitched copies, the device-side synthetic image generator, an L2 flush and
// error reporting. The reference's equivalent is a dozen cudaMalloc/cudaMemcpy calls inlined in
// gpuConvolute (cuda/cuda_convolute.cu:49-102) guarded by CUDA_SAFE_CALL -> exit()
// (cuda/funcs.h:10-16); here every entry point returns a code and keeps a message instead of
// killing the process.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "pconv_cuda.h"
#include "runtime_internal.h"

namespace {

std::mutex g_err_mu;
std::string g_err = "";
std::atomic<long long> g_launches{0};

struct DevInfo {
    int dev = -1, sms = 0, smem_optin = 0;
    long long l2 = 0;
};
thread_local DevInfo t_info;

const DevInfo& info() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (t_info.dev != dev) {
        t_info.dev = dev;
        cudaDeviceGetAttribute(&t_info.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&t_info.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        int l2 = 0;
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
        t_info.l2 = l2;
    }
    return t_info;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// One thread per byte-octet would need alignment games; one thread per byte is plenty for a
// generator that runs once per benchmark.
__global__ void synth_kernel(unsigned long long seed, long long row_bytes, long long row0,
                             long long nrows, long long byte0, long long nbytes, uint8_t* dst,
                             long long pitch) {
    const long long total = nrows * nbytes;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / nbytes, c = i % nbytes;
        const unsigned long long gidx = (unsigned long long)((row0 + r) * row_bytes + byte0 + c);
        const unsigned long long word = splitmix64(seed ^ (gidx >> 3));
        dst[r * pitch + c] = (uint8_t)(word >> (8 * (gidx & 7)));
    }
}

__global__ void flush_kernel(uint4* p, long long n16, unsigned v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = make_uint4(v, v + 1, v + 2, v + 3);
}

}  // namespace

namespace pconv_rt {

int fail(const char* msg) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = msg;
    return -1;
}
int fail_cuda(cudaError_t e, const char* where) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return -2;
}
int sm_count() { return info().sms; }
int max_optin_smem() { return info().smem_optin; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
    static const bool on = [] { const char* v = std::getenv("PCONV_PDL"); return !(v && v[0] == '0'); }();
    return on;
}

void make_filter_dev(const pconv_filter_t& f, pconv::FilterDev* out) {
    std::memset(out, 0, sizeof(*out));
    double pos = 0.0, neg = 0.0;
    long long ipos = 0, ineg = 0;
    for (int i = 0; i < 9; ++i) {
        out->h[i] = f.h[i];
        out->w[i] = f.w[i];
        if (f.h[i] > 0) pos += (double)f.h[i]; else neg -= (double)f.h[i];
        if (f.w[i] > 0) ipos += f.w[i]; else ineg -= f.w[i];
    }
    out->shift = f.shift;
    out->symmetric = f.h[0] == f.h[2] && f.h[3] == f.h[5] && f.h[6] == f.h[8] &&   // left == right
                     f.h[0] == f.h[6] && f.h[1] == f.h[7];                             // top == bottom
    if (f.kind == PCONV_KIND_F32) {
        out->clamp_lo = neg > 0.0;
        out->clamp_hi = pos * 255.0 >= 255.5;  // below 256 truncation already yields <= 255
    } else {
        const unsigned bias = (unsigned)(ineg * 255);  // -S_min; bias + S_max <= 32767 (classifier)
        out->bias2 = (bias << 16) | bias;
        const unsigned nb = (unsigned)(-(int)bias) & 0xffffu;
        out->negbias2 = (nb << 16) | nb;
        const unsigned mv = (256u << f.shift) - 1u;
        out->maxv2 = (mv << 16) | mv;
        out->clamp_lo = ineg > 0;
        out->clamp_hi = ((ipos * 255) >> f.shift) > 255;
    }
}

}  // namespace pconv_rt

using namespace pconv_rt;

#define RT_CHECK(call, what)                              \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return fail_cuda(e__, what); \
    } while (0)

extern "C" {

int pconv_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int pconv_cuda_set_device(int dev) { RT_CHECK(cudaSetDevice(dev), "cudaSetDevice"); return 0; }
int pconv_cuda_sm_count(void) { return sm_count(); }
int64_t pconv_cuda_l2_bytes(void) { return info().l2; }
const char* pconv_cuda_last_error(void) {
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mu);
    copy = g_err;
    return copy.c_str();
}
const char* pconv_cuda_version(void) { return "pconv-cuda 0.1 (sm_100a)"; }

int pconv_cuda_malloc(void** p, int64_t bytes) {
    RT_CHECK(cudaMalloc(p, (size_t)bytes), "cudaMalloc");
    RT_CHECK(cudaMemset(*p, 0, (size_t)bytes), "cudaMemset");
    return 0;
}
int pconv_cuda_free(void* p) { RT_CHECK(cudaFree(p), "cudaFree"); return 0; }

int pconv_cuda_ipc_export(void* p, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    RT_CHECK(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
    std::memcpy(handle, &h, 64);
    return 0;
}
int pconv_cuda_ipc_open(const uint8_t handle[64], void** p) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    RT_CHECK(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return 0;
}
int pconv_cuda_ipc_close(void* p) { RT_CHECK(cudaIpcCloseMemHandle(p), "cudaIpcCloseMemHandle"); return 0; }

int pconv_cuda_enable_peer(int peer_dev) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    RT_CHECK(e, "cudaDeviceEnablePeerAccess");
    return 0;
}

int pconv_cuda_memset_async(void* p, int value, int64_t bytes, void* stream) {
    RT_CHECK(cudaMemsetAsync(p, value, (size_t)bytes, (cudaStream_t)stream), "cudaMemsetAsync");
    return 0;
}
int pconv_cuda_memcpy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch,
                              int64_t width_bytes, int64_t rows, int kind, void* stream) {
    RT_CHECK(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes,
                               (size_t)rows, (cudaMemcpyKind)kind, (cudaStream_t)stream),
             "cudaMemcpy2DAsync");
    return 0;
}
int pconv_cuda_stream_sync(void* stream) {
    RT_CHECK(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
    return 0;
}
int pconv_cuda_memcpy_async(void* dst, const void* src, int64_t bytes, int kind, void* stream) {
    RT_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, (cudaMemcpyKind)kind, (cudaStream_t)stream),
             "cudaMemcpyAsync");
    return 0;
}
int pconv_cuda_device_sync(void) { RT_CHECK(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); return 0; }
int pconv_cuda_stream_create(void** stream) {
    cudaStream_t s;
    RT_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    *stream = (void*)s;
    return 0;
}
int pconv_cuda_stream_destroy(void* stream) { RT_CHECK(cudaStreamDestroy((cudaStream_t)stream), "cudaStreamDestroy"); return 0; }
int pconv_cuda_event_create(void** ev) {
    cudaEvent_t e;
    RT_CHECK(cudaEventCreate(&e), "cudaEventCreate");
    *ev = (void*)e;
    return 0;
}
int pconv_cuda_event_destroy(void* ev) { RT_CHECK(cudaEventDestroy((cudaEvent_t)ev), "cudaEventDestroy"); return 0; }
int pconv_cuda_event_record(void* ev, void* stream) {
    RT_CHECK(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream), "cudaEventRecord");
    return 0;
}
int pconv_cuda_event_sync(void* ev) { RT_CHECK(cudaEventSynchronize((cudaEvent_t)ev), "cudaEventSynchronize"); return 0; }
int pconv_cuda_stream_wait_event(void* stream, void* ev) {
    RT_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0), "cudaStreamWaitEvent");
    return 0;
}
int pconv_cuda_event_elapsed_ms(void* start, void* stop, float* ms) {
    RT_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop), "cudaEventElapsedTime");
    return 0;
}
int pconv_cuda_host_alloc(void** p, int64_t bytes) {
    RT_CHECK(cudaHostAlloc(p, (size_t)bytes, cudaHostAllocPortable), "cudaHostAlloc");
    return 0;
}
int pconv_cuda_host_free(void* p) { RT_CHECK(cudaFreeHost(p), "cudaFreeHost"); return 0; }

int pconv_cuda_synth_rect(uint64_t seed, int64_t row_bytes, int64_t row0, int64_t nrows,
                          int64_t byte0, int64_t nbytes, uint8_t* dst, int64_t pitch, void* stream) {
    if (nrows <= 0 || nbytes <= 0) return 0;
    const long long total = nrows * nbytes;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    synth_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, row_bytes, row0, nrows, byte0,
                                                          nbytes, dst, pitch);
    count_launch();
    RT_CHECK(cudaGetLastError(), "synth_kernel");
    return 0;
}

int pconv_cuda_flush_l2(void* scratch, int64_t bytes, void* stream) {
    static std::atomic<unsigned> salt{1};
    flush_kernel<<<sm_count() * 4, 512, 0, (cudaStream_t)stream>>>((uint4*)scratch, bytes / 16,
                                                                  salt.fetch_add(4));
    RT_CHECK(cudaGetLastError(), "flush_kernel");
    return 0;
}

int64_t pconv_cuda_launch_count(void) { return g_launches.load(); }

}  // extern "C"
