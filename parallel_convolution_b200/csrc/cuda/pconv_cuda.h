// C ABI of libpconv_cuda.so — sm_100a kernels + the small runtime around them.
// Bound from Python with ctypes (parallel_convolution_b200/ops/cuda.py); structs here are mirrored
// there field-for-field.
#ifndef PCONV_CUDA_H
#define PCONV_CUDA_H

#include <stdint.h>
#include "../common/pconv_types.h"

#ifdef __cplusplus
extern "C" {
#endif

// Directions are row-major (dy,dx) in {-1,0,1}^2: 0=NW 1=N 2=NE 3=W 4=self 5=E 6=SW 7=S 8=SE.
// The tile seen from neighbour d looks back at me through direction 8-d.

// Everything a kernel needs to know about where pixels live. One rank == one tile.
typedef struct pconv_geometry {
    int64_t img_rows;      // H
    int64_t img_cb;        // W * C  (bytes per image row)
    int64_t row_lo[4];     // global row boundaries: start of tile-row ty-1, ty, ty+1, end of ty+1
    int64_t cb_lo[4];      // same in column *bytes* for tile-col tx-1, tx, tx+1
    int64_t pitch;         // byte pitch of every tile buffer (multiple of 16)
    int32_t C;             // bytes per pixel == horizontal neighbour distance
    int32_t pad_;
} pconv_geometry_t;

// Cross-GPU (or virtual-rank) signalling: all memory is plain device memory; for multi-GPU runs
// local_flags lives in peer-visible memory and remote_flags[] are peer-mapped pointers.
typedef struct pconv_sync {
    uint32_t* local_flags;      // [9] slot d: last epoch completed by neighbour d (they write it)
    uint32_t* remote_flags[9];  // where I announce my epoch to neighbour d (NULL if absent)
    uint32_t* done_counter;     // device scratch: CTAs finished in this launch (self-resetting)
    uint32_t wait_value;        // border CTAs spin until local_flags[d] >= wait_value (0: no wait)
    uint32_t signal_value;      // last CTA stores this to every remote flag (0: no signal)
    uint64_t timeout_ns;        // watchdog for the spin (0: default 10 s)
    uint64_t* status;           // [0] error code (1 = watchdog fired), [1] max wait ns, [2] sum wait
                                // ns, [3] number of waits; resident kernel probe CTA: [4] publish ns,
                                // [5] flag+wait ns, [6] halo reload ns, [7] compute ns, [8] exchanges
    int32_t inject_delay_ns;    // fault injection: delay my signal by this much (tests)
    int32_t inject_drop_signal; // fault injection: never signal (the peer's watchdog must fire)
} pconv_sync_t;

// ---- streaming kernel: k fused iterations per launch over any tile size ------------------------
typedef struct pconv_stream_params {
    pconv_geometry_t g;
    const uint8_t* src[9];   // source tile base pointers by direction (NULL: no such neighbour)
    int64_t src_pitch[9];    // byte pitch of each source (peer tiles: g.pitch; NCCL ghost strips:
                             // their own dense pitch, with src[d] a *virtual* base so that
                             // src[d] + local_row*pitch + local_byte lands inside the strip)
    uint8_t* dst;            // my destination tile (pitch g.pitch)
    pconv_filter_t filt;
    int32_t k;               // fused iterations in this launch (>= 1)
    int32_t TH;              // interior rows per CTA region (TH + 2k <= 16*R)
    int32_t TWB;             // interior bytes per CTA region (multiple of 16, <= 512 - 2*HB)
    int32_t HB;              // column halo bytes (multiple of 16, >= k*C)
    int32_t grid;            // CTAs to launch (0: one per SM)
    int32_t R;               // band height: rows per warp held in registers (region = 16*R rows)
    pconv_sync_t sync;
} pconv_stream_params_t;

int pconv_stream_launch(const pconv_stream_params_t* p, void* stream);
// Enqueue a whole run from native code (one call instead of one Python round trip per launch):
// `pp` holds 2*nranks parameter blocks, [2*r + 0] reading buffer parity A and [2*r + 1] parity B of
// (virtual) rank r; launch i uses block (i & 1) with k = ks[i]; with use_flags the epochs are
// signal = epoch0 + i + 1 and wait = (i ? epoch0 + i : first_wait).
// first_flags (optional, [nranks]): flag array the FIRST launch waits on instead of
// sync.local_flags — the "input ready" bank of the pipelined step protocol.
int pconv_stream_run(const pconv_stream_params_t* pp, int nranks, const int32_t* ks, int nlaunch,
                     uint32_t epoch0, int use_flags, uint32_t first_wait, uint32_t* const* first_flags,
                     void* stream);
// Stand-alone halo-protocol step (1 CTA): wait until every existing neighbour reached wait_value,
// then announce signal_value to them. Brackets host/device copies into a tile buffer between runs
// so that back-to-back steps need no host barrier (bench.py, streaming ingest).
int pconv_sync_launch(const pconv_sync_t* s, void* stream);
// Force the lazy module load of the kernel a plan will use (cudaFuncGetAttributes), so that the
// first timed run does not pay ~10 ms of one-time SASS loading. resident: 0 stream kernel,
// 1 resident kernel, 2 split-phase resident kernel, 3 split-phase + protocol warp (experimental).
int pconv_kernel_prepare(int kind, int C, int R, int resident);
// Band heights compiled for a filter kind; returns the count, fills out[0..cap).
int pconv_stream_band_heights(int kind, int* out, int cap);

// ---- resident kernel: the whole tile stays in registers for ALL iterations of a run ------------
// One persistent launch: CTA (ti,tj) keeps its region in registers, and every k iterations the CTAs
// exchange k-deep border strips through the (L2-resident / peer-mapped) tile buffers with per-CTA
// epoch flags — between CTAs of one GPU and, for border CTAs, straight across NVLink.
// Requires nty*ntx <= #SMs (all CTAs co-resident; launched cooperatively) and identical tile
// shapes on every rank.
typedef struct pconv_resident_params {
    pconv_geometry_t g;
    const uint8_t* src_a[9]; // parity-0 buffers by direction: the input image, then even exchanges
    const uint8_t* src_b[9]; // parity-1 buffers by direction: odd exchanges
    uint8_t* self_a;         // == src_a[4], writable
    uint8_t* self_b;         // == src_b[4], writable
    pconv_filter_t filt;
    int32_t loops;           // total iterations in this launch (>= 1)
    int32_t k;               // iterations between exchanges
    int32_t TH, TWB, HB;     // region interior / halo geometry (as in the stream kernel)
    int32_t R;               // band height
    int32_t nty, ntx;        // CTA grid (same on every rank)
    uint32_t* cta_flags;     // [(nty+2)*(ntx+2)] epochs, incl. a ring mirrored from neighbour GPUs
    uint32_t* remote_cta_flags[9];  // the same array on neighbour d (peer-mapped) or NULL
    uint32_t epoch_base;     // exchange b (1-based) uses epoch_base + b; host advances it per run
    uint32_t dual;           // 1: split-phase kernel (two virtual tile rows per CTA, nty counts them; even);
                             // 2: the same with a protocol warp (resident3_kernel.cuh, experimental)
    pconv_sync_t sync;       // launch-level protocol, identical to the stream kernel's
} pconv_resident_params_t;
// The result lands in parity (ceil(loops/k) & 1): 1 -> self_b, 0 -> self_a.
int pconv_resident_launch(const pconv_resident_params_t* p, void* stream);
// Band heights (rows per warp per virtual tile) compiled for the split-phase kernel.
int pconv_resident2_band_heights(int kind, int* out, int cap);
int pconv_resident3_band_heights(int kind, int* out, int cap);

// ---- runtime -----------------------------------------------------------------------------------
int pconv_cuda_device_count(void);
int pconv_cuda_set_device(int dev);
int pconv_cuda_sm_count(void);
int64_t pconv_cuda_l2_bytes(void);
const char* pconv_cuda_last_error(void);
const char* pconv_cuda_version(void);

int pconv_cuda_malloc(void** p, int64_t bytes);   // cudaMalloc (IPC-exportable, zero-filled)
int pconv_cuda_free(void* p);
int pconv_cuda_ipc_export(void* p, uint8_t handle[64]);
int pconv_cuda_ipc_open(const uint8_t handle[64], void** p);
int pconv_cuda_ipc_close(void* p);
int pconv_cuda_enable_peer(int peer_dev);          // single-process multi-device mode
int pconv_cuda_memset_async(void* p, int value, int64_t bytes, void* stream);
int pconv_cuda_memcpy2d_async(void* dst, int64_t dpitch, const void* src, int64_t spitch,
                              int64_t width_bytes, int64_t rows, int kind, void* stream);
int pconv_cuda_memcpy_async(void* dst, const void* src, int64_t bytes, int kind, void* stream);
int pconv_cuda_stream_sync(void* stream);
int pconv_cuda_device_sync(void);
int pconv_cuda_stream_create(void** stream);
int pconv_cuda_stream_destroy(void* stream);
int pconv_cuda_event_create(void** ev);
int pconv_cuda_event_destroy(void* ev);
int pconv_cuda_event_record(void* ev, void* stream);
int pconv_cuda_event_sync(void* ev);
int pconv_cuda_stream_wait_event(void* stream, void* ev);
int pconv_cuda_event_elapsed_ms(void* start, void* stop, float* ms);
int pconv_cuda_host_alloc(void** p, int64_t bytes);  // pinned
int pconv_cuda_host_free(void* p);

// Synthetic image: same function of (seed, global byte index) as pconv_synth_rect on the host.
int pconv_cuda_synth_rect(uint64_t seed, int64_t row_bytes, int64_t row0, int64_t nrows,
                          int64_t byte0, int64_t nbytes, uint8_t* dst, int64_t pitch, void* stream);
// Write `bytes` of zeros-then-ones garbage through L2 so the next kernel starts cold.
int pconv_cuda_flush_l2(void* scratch, int64_t bytes, void* stream);

// ---- pack/unpack kernels for the NCCL baseline backend -----------------------------------------
// Copy a rows x nbytes rectangle between a pitched tile and a dense staging buffer.
int pconv_cuda_pack_rect(const uint8_t* tile, int64_t pitch, int64_t row0, int64_t byte0,
                         int64_t rows, int64_t nbytes, uint8_t* dense, void* stream);
int pconv_cuda_unpack_rect(uint8_t* tile, int64_t pitch, int64_t row0, int64_t byte0,
                           int64_t rows, int64_t nbytes, const uint8_t* dense, void* stream);

// Number of kernels this library has launched since load (bench.py reports it as gpu_launches).
int64_t pconv_cuda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
