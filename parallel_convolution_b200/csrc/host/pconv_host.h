// C ABI of libpconv_host.so — the CPU side of the framework (no CUDA dependency).
// Loaded from Python through ctypes (parallel_convolution_b200/ops/host.py) and linked into the
// native CLI (csrc/native_main.cpp).
#ifndef PCONV_HOST_H
#define PCONV_HOST_H

#include <stdint.h>
#include "../common/pconv_types.h"

#ifdef __cplusplus
extern "C" {
#endif

// ---- filters (models/filters.py is the Python face of these) --------------------------------
// Build a filter from integer numerators and a divisor the way the reference does
// (h = (float)(w / (double)div), mpi/mpi_conv.c:98-107) and classify the exact fast path.
int pconv_filter_from_int(const int32_t w[9], double div, pconv_filter_t* out);
// Build from arbitrary fp32 taps (classified INT when every tap is n/2^s with small n).
int pconv_filter_from_float(const float h[9], pconv_filter_t* out);

// ---- decomposition (reference: divide_rows, mpi/mpi_conv.c:363-377) --------------------------
// Reference-compatible chooser: returns rows-of-tiles, 0 when no exact division exists.
int pconv_divide_rows(int64_t rows, int64_t cols, int workers);
// Extended chooser: exact division when possible (identical to the reference), else the
// perimeter-minimising grid with remainder tiles. Returns 0 on success.
int pconv_choose_grid(int64_t rows, int64_t cols, int workers, int* row_div, int* col_div,
                      int* exact);
// Balanced 1-D split with remainder: part idx of `parts` over [0, extent).
void pconv_split(int64_t extent, int parts, int idx, int64_t* start, int64_t* size);

// ---- launch planner for the GPU kernels (csrc/host/planner.cpp) ------------------------------
typedef struct pconv_plan_request {
    int64_t tile_rows, tile_cb;     // largest tile of the grid: rows, bytes per row
    int64_t min_nbr_rows, min_nbr_cols;  // smallest neighbouring tile extent in pixels (0: none)
    int32_t C, kind, loops, sms;
    int32_t bands[16];              // band heights R compiled into the CUDA library
    int32_t nbands;
    int32_t force_k, force_R;       // 0: planner decides
    int32_t resident;               // 1: plan the persistent kernel (all CTAs co-resident)
    int32_t peers;                  // 1: neighbours sit on other GPUs (dearer exchanges)
    int32_t warps;                  // persistent kernel: warps per CTA, 8 (two CTAs per SM) or 16
} pconv_plan_request_t;
typedef struct pconv_plan_result {
    int32_t k, R, TH, TWB, HB, ntx, nty, waves;
    double est_us;                  // modelled time per launch (stream) / per run (resident)
} pconv_plan_t;
// 0: ok, 1: no feasible plan, <0: bad request.
int pconv_plan(const pconv_plan_request_t* req, pconv_plan_t* out);

// ---- CPU emulation of the persistent kernel's exchange protocol (csrc/host/xplan_emu.cpp) ----------
// Replays every (rank, CTA, warp, lane) of a persist_kernel launch with the device code's own index
// functions (common/xplan.h); out must equal the oracle. 0 ok, <0 rejected geometry / slot overflow.
int pconv_xplan_emulate(const uint8_t* img, uint8_t* out, int64_t H, int64_t Wpx, int C,
                        const pconv_filter_t* f, int loops, int gy, int gx, int k, int R, int W,
                        int TH, uint32_t epoch_base);

// ---- oracle: whole image, zero padded, scalar, deliberately naive ----------------------------
// src/dst: H x (W*C) bytes, dense. Applies the filter `loops` times; result is in dst.
// scratch must hold H*W*C bytes (ping-pong partner). Returns 0.
int pconv_oracle_run(const uint8_t* src, uint8_t* dst, uint8_t* scratch, int64_t H, int64_t W,
                     int C, const pconv_filter_t* f, int loops);

// ---- CPU engine: one iteration over a rectangle of a halo-padded tile ------------------------
// `src0`/`dst0` point at pixel (0,0) of the tile interior; every computed pixel must have a
// valid 1-pixel halo around it in src (the caller owns the padding). Rows [r0,r1), pixel
// columns [c0,c1). `threads` <= 0 means "use OpenMP default".
int pconv_cpu_convolve(const uint8_t* src0, uint8_t* dst0, int64_t pitch, int C,
                       const pconv_filter_t* f, int64_t r0, int64_t r1, int64_t c0, int64_t c1,
                       int threads);
// Whole dense image, `loops` iterations, zero padded, multi-threaded; dst receives the result.
int pconv_cpu_run_image(const uint8_t* src, uint8_t* dst, int64_t H, int64_t W, int C,
                        const pconv_filter_t* f, int loops, int threads);
// The loop alone, on caller-owned padded images (what a timer should bracket; the reference times its loop the
// same way, after the calloc and the file read: mpi/mpi_conv.c:113-162): pa/pb = pixel (0,0) of two (H+2) x
// pitch images with a zero 1-pixel frame, pitch = pconv_cpu_padded_pitch(W, C), input in pa. k <= 1 plain
// passes, k > 1 temporally blocked. Returns 0 / 1: the result is in pa / pb; < 0 on bad arguments.
int pconv_cpu_run_padded(uint8_t* pa, uint8_t* pb, int64_t pitch, int64_t H, int64_t W, int C,
                         const pconv_filter_t* f, int loops, int threads, int k);
int64_t pconv_cpu_padded_pitch(int64_t W, int C);
// The depth pconv_cpu_run_image picks: 1 while both padded images fit in the last-level cache, else up to 6
// (PCONV_CPU_BLOCK_DEPTH overrides).
int pconv_cpu_blocking_depth(int64_t H, int64_t W, int C, int loops);
// Same result, temporally blocked: row bands advance k iterations in L2-sized private buffers
// (pconv_cpu_run_image picks this automatically for memory-bound sizes).
int pconv_cpu_run_image_blocked(const uint8_t* src, uint8_t* dst, int64_t H, int64_t W, int C,
                                const pconv_filter_t* f, int loops, int threads, int k);
// The building block of every temporally blocked CPU path: rows [r0,r1) x pixel columns [c0,c1) of an
// H x W image advance kk iterations inside two scratch buffers (each pconv_cpu_band_scratch_bytes(r1-r0,
// c1-c0, C, kk) bytes); the kk-deep halo is read from src0, only the band itself is written to dst0. Both
// images share `pitch` and carry a zero frame; src0/dst0 point at pixel (0,0).
int64_t pconv_cpu_band_scratch_bytes(int64_t band_rows, int64_t cols, int C, int k);
void pconv_cpu_band_advance(const uint8_t* src0, uint8_t* dst0, int64_t pitch, int64_t H, int64_t W, int C,
                            const pconv_filter_t* f, int64_t r0, int64_t r1, int64_t c0, int64_t c1, int kk,
                            uint8_t* scratch_a, uint8_t* scratch_b);
// Out-of-core: file -> file, `loops` iterations, never holding more than ~mem_bytes of the image
// (row strips with k-deep halos, ceil(loops/k) passes; intermediate passes use "<out>.pass").
// k <= 0 / mem_bytes <= 0 pick defaults (8 iterations per pass, 256 MiB). 0 or -errno.
int pconv_cpu_run_file(const char* in_path, const char* out_path, int64_t H, int64_t W, int C,
                       const pconv_filter_t* f, int loops, int threads, int k, int64_t mem_bytes);
int pconv_omp_max_threads(void);

// ---- CPU ranks of one node over a shared image (csrc/host/shm_ranks.cpp) -------------------------
// One segment (a file under /dev/shm) holds both ping-pong images, zero frame included, and one epoch
// counter per rank; ranks read their halos directly from the neighbours' part of the image.
typedef struct pconv_shm pconv_shm_t;
// Rank 0 creates (fails if the path exists), the others attach after it; NULL on error. The path may be
// unlinked as soon as every rank has attached.
pconv_shm_t* pconv_shm_create(const char* path, int64_t H, int64_t W, int C, int nranks);
pconv_shm_t* pconv_shm_attach(const char* path);
void pconv_shm_close(pconv_shm_t* s);
uint8_t* pconv_shm_pixels(pconv_shm_t* s, int which);   // pixel (0,0) of image 0 / 1
int64_t pconv_shm_pitch(const pconv_shm_t* s);
int64_t pconv_shm_bytes(const pconv_shm_t* s);
int64_t pconv_shm_epoch(const pconv_shm_t* s, int rank);   // epochs completed; newest data in image[epoch & 1]
void pconv_shm_reset(pconv_shm_t* s, int rank);           // epoch := 0 (callers synchronise around it)
void pconv_shm_abort(pconv_shm_t* s, int rank);           // make ranks waiting on `rank` fail with -ECANCELED
// For launchers without a process group (pconv_native -np N; the mapping survives fork()): a barrier over the
// segment's nranks processes (0, -ETIMEDOUT, -ECANCELED when a rank aborted) and one double per rank.
int pconv_shm_barrier(pconv_shm_t* s, double timeout_s);
void pconv_shm_set_value(pconv_shm_t* s, int rank, double v);
double pconv_shm_get_value(const pconv_shm_t* s, int rank);
// Advance this rank's tile (rows [r0,r1), pixel columns [c0,c1)) by `loops` iterations, k per epoch.
// nbr[9]: neighbour ranks in row-major 3x3 order (index 4 = self), -1 where the image ends; k must not
// exceed the smallest neighbouring tile extent. 0, -EINVAL, -ETIMEDOUT (a neighbour did not advance within
// timeout_s; <= 0 waits forever) or -ECANCELED (a neighbour aborted).
int pconv_shm_run(pconv_shm_t* s, int rank, int64_t r0, int64_t r1, int64_t c0, int64_t c1,
                  const int32_t nbr[9], const pconv_filter_t* f, int loops, int k, int threads,
                  double timeout_s, int64_t* epochs_run);

// ---- CPU ranks with private tiles and TCP messages (csrc/host/net_ranks.cpp) --------------------------
// For ranks that do not share a node: full mesh of TCP connections, rendezvous through rank 0
// (master_host:master_port; master_fd >= 0 hands rank 0 an already listening socket), k-deep halo exchange
// with the reference's schedule (send -> halo-free part of the first iteration -> wait -> frame).
typedef struct pconv_net pconv_net_t;
pconv_net_t* pconv_net_init(int rank, int world, const char* master_host, int master_port, int master_fd,
                            double timeout_s);
void pconv_net_close(pconv_net_t* n);
const char* pconv_net_error(const pconv_net_t* n);
int pconv_net_barrier(pconv_net_t* n);
int pconv_net_max(pconv_net_t* n, double* v);            // in place: the maximum over all ranks, on every rank
int64_t pconv_net_tile_pitch(int64_t cols, int C, int k);
// buf_a / buf_b: two ZEROED buffers of (rows + 2k) * pitch bytes, the tile's pixel (0,0) at row k, byte k*C, the
// input tile in buf_a; nbr[9] as in pconv_shm_run; k <= the smallest tile extent. Returns 0 / 1 = the result is in
// buf_a / buf_b, or a negative errno (pconv_net_error says why).
int pconv_net_run(pconv_net_t* n, int64_t H, int64_t W, int C, int64_t row0, int64_t rows, int64_t col0,
                  int64_t cols, const int32_t nbr[9], uint8_t* buf_a, uint8_t* buf_b, const pconv_filter_t* f,
                  int loops, int k, int threads, int64_t* exchanges);

// ---- raw image I/O (64-bit offsets; reference: MPI_File_* mpi/mpi_conv.c:133-147,253-271) ----
int64_t pconv_file_size(const char* path);
// Read a rectangle: rows [row0,row0+nrows), bytes [byte0,byte0+nbytes) of every row of a file
// whose rows are `row_bytes` long, into dst with `dst_pitch`. Returns 0 or -errno (-EIO on a
// short file — the reference spins forever there, cuda/funcs.c:39-45).
int pconv_read_rect(const char* path, int64_t row_bytes, int64_t row0, int64_t nrows,
                    int64_t byte0, int64_t nbytes, uint8_t* dst, int64_t dst_pitch);
// Create/truncate the output to its full size (rank 0 calls this once before the tiles land).
int pconv_create_file(const char* path, int64_t total_bytes);
int pconv_write_rect(const char* path, int64_t row_bytes, int64_t row0, int64_t nrows,
                     int64_t byte0, int64_t nbytes, const uint8_t* src, int64_t src_pitch);

// ---- synthetic images: value of byte i of the whole image = hash(seed, i) ---------------------
// Independent of the decomposition, so any rank can generate exactly its own tile.
void pconv_synth_rect(uint64_t seed, int64_t row_bytes, int64_t row0, int64_t nrows,
                      int64_t byte0, int64_t nbytes, uint8_t* dst, int64_t dst_pitch);

const char* pconv_host_version(void);

#ifdef __cplusplus
}
#endif
#endif
