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
                                // ns, [3] number of waits; persistent kernel probe CTA: [4] push ns,
                                // [5] signal+wait ns, [6] inbox->register ns, [7] compute ns, [8] exchanges
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
// first timed run does not pay ~10 ms of one-time SASS loading (stream kernel; the persistent kernel
// is loaded by pconv_persist_capacity).
int pconv_kernel_prepare(int kind, int C, int R);
// Band heights compiled for a filter kind; returns the count, fills out[0..cap).
int pconv_stream_band_heights(int kind, int* out, int cap);

// ---- persistent kernel: the whole tile stays in registers for ALL iterations of a run ---------
// One cooperative launch: CTA (ti,tj) keeps its region (W*R rows x 512 B) in registers; every k
// iterations the CTAs PUSH k-deep border blocks into each other's inbox slots (an exchange arena per
// rank, peer-mapped across GPUs), signal with one release-store per neighbour and fetch their own inbox
// with bulk async copies on an mbarrier (persist_kernel.cuh, common/xplan.h). W = 8 runs two CTAs per
// SM (one CTA's exchange hides behind the other's arithmetic), W = 16 one.
// Requires identical tile shapes on every rank and all nty*ntx CTAs co-resident.
typedef struct pconv_persist_params {
    pconv_geometry_t g;
    const uint8_t* src[9];   // input tiles by direction (NULL: no such neighbour); [4] is mine
    uint8_t* dst;            // my output tile (the other buffer of the ping-pong pair)
    pconv_filter_t filt;
    int32_t loops;           // total iterations in this launch (>= 1)
    int32_t k;               // iterations between exchanges
    int32_t TH, TWB, HB;     // region interior / halo geometry (TWB + 2*HB == 512)
    int32_t R, W;            // band height, warps per CTA (8 or 16)
    int32_t nty, ntx;        // CTA grid (same on every rank)
    uint8_t* arena[9];       // exchange arenas by rank direction: [4] mine, others peer-mapped or NULL
    int64_t arena_bytes;     // size of every arena
    uint32_t epoch_base;     // exchange b (1-based) carries epoch_base + b; the host advances it by
                             // ceil(loops/k) - 1 per run (epochs stay contiguous: the inbox parity alternates)
    uint32_t peer_scope;     // 1: neighbouring ranks are other GPUs (system-scope flags)
    uint32_t debug_skip;     // timing experiments only (wrong results): 1 no push, 2 no wait/fetch, 4 no refresh
    uint32_t pad_;
    pconv_sync_t sync;       // launch-level protocol, identical to the stream kernel's
} pconv_persist_params_t;
int pconv_persist_launch(const pconv_persist_params_t* p, void* stream);
// Bytes of exchange arena this plan needs (flags + two inbox parities for every CTA); <0: bad plan.
int64_t pconv_persist_arena_bytes(const pconv_persist_params_t* p);
// CTAs of the plan's kernel instance the device keeps resident (also loads the module); <0: error.
int pconv_persist_capacity(const pconv_persist_params_t* p);
// Band heights compiled for a filter kind (same list for W = 8 and W = 16).
int pconv_persist_band_heights(int kind, int* out, int cap);

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

// The same for up to 8 rectangles in ONE launch (128-bit copies when pitch, width and both pointers are
// 16-byte aligned): pack != 0 copies tile -> dense, else dense -> tile.
typedef struct pconv_rect {
    uint8_t* tile;
    int64_t pitch, row0, byte0, rows, nbytes;
    uint8_t* dense;
} pconv_rect_t;
int pconv_cuda_rect_batch(const pconv_rect_t* rects, int n, int pack, void* stream);

// Number of kernels this library has launched since load (bench.py reports it as gpu_launches).
int64_t pconv_cuda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
