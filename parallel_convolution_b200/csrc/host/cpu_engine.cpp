// CPU engine: the B200 framework's stand-in for the reference's mpi_conv / mpi_omp_conv compute
// (convolute/convolute_grey/convolute_rgb, mpi/mpi_conv.c:299-333 and the OpenMP worksharing of
// mpi_omp/mpi_omp_conv.c:297-310). Differences by design, not translation:
//   * grey and RGB are the same loop: an interleaved image is a byte stream whose horizontal
//     neighbour is C bytes away, so one kernel serves any channel count;
//   * row kernels are written to auto-vectorise (AVX2/AVX-512 clones) instead of a call per pixel
//     through float** (SURVEY.md §3.1);
//   * the integer path is used exactly when it is provably identical to the fp32 oracle
//     (filters.cpp); the fp32 path keeps the oracle's tap order and is compiled with
//     -ffp-contract=off, so both are byte-exact;
//   * OpenMP parallelises rows with one parallel region per pass (the reference opens one per
//     convolute() call, including 1x1 corner calls, and its pragma does not compile: defect D7).
#include <omp.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cstdio>
#include <string>

#include "pconv_host.h"

#define PCONV_CLONES __attribute__((target_clones("default", "avx2", "arch=x86-64-v4")))

namespace {

// ---- one output row, n bytes, neighbour distance C bytes -------------------------------------

PCONV_CLONES void row_int(const uint8_t* __restrict up, const uint8_t* __restrict mid,
                          const uint8_t* __restrict dn, uint8_t* __restrict out, int64_t n,
                          int C, const int32_t* w, int shift) {
    const short w0 = (short)w[0], w1 = (short)w[1], w2 = (short)w[2];
    const short w3 = (short)w[3], w4 = (short)w[4], w5 = (short)w[5];
    const short w6 = (short)w[6], w7 = (short)w[7], w8 = (short)w[8];
    for (int64_t x = 0; x < n; ++x) {
        // |S| <= sum|w| * 255 <= 32640: fits 16-bit lanes (filters.cpp guarantees sum|w| <= 128)
        short s = (short)(w0 * up[x - C] + w1 * up[x] + w2 * up[x + C] + w3 * mid[x - C] +
                          w4 * mid[x] + w5 * mid[x + C] + w6 * dn[x - C] + w7 * dn[x] +
                          w8 * dn[x + C]);
        short v = (short)(s >> shift);  // arithmetic shift: floor; negatives clamp to 0 below
        v = v < 0 ? (short)0 : v;
        v = v > 255 ? (short)255 : v;
        out[x] = (uint8_t)v;
    }
}

PCONV_CLONES void row_sep121(const uint8_t* __restrict up, const uint8_t* __restrict mid,
                             const uint8_t* __restrict dn, uint8_t* __restrict out, int64_t n,
                             int C) {
    // vertical [1 2 1] first into a 16-bit row buffer (<= 1020), then horizontal (<= 4080) and >> 4:
    // exact Gaussian/16. Two simple loops vectorise (u8 -> u16 widening adds); a fused 9-tap
    // expression does not.
    constexpr int64_t kChunk = 4096;
    unsigned short v[kChunk + 16];
    for (int64_t x0 = 0; x0 < n; x0 += kChunk) {
        const int64_t m = (n - x0 < kChunk ? n - x0 : kChunk);
        const uint8_t* u = up + x0 - C;
        const uint8_t* c = mid + x0 - C;
        const uint8_t* d = dn + x0 - C;
        for (int64_t x = 0; x < m + 2 * C; ++x)
            v[x] = (unsigned short)((unsigned short)u[x] + (unsigned short)(2 * c[x]) + (unsigned short)d[x]);
        uint8_t* o = out + x0;
        for (int64_t x = 0; x < m; ++x)
            o[x] = (uint8_t)((unsigned short)(v[x] + 2 * v[x + C] + v[x + 2 * C]) >> 4);
    }
}

PCONV_CLONES void row_f32(const uint8_t* __restrict up, const uint8_t* __restrict mid,
                          const uint8_t* __restrict dn, uint8_t* __restrict out, int64_t n,
                          int C, const float* h) {
    const float h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3], h4 = h[4], h5 = h[5], h6 = h[6],
                h7 = h[7], h8 = h[8];
    for (int64_t x = 0; x < n; ++x) {
        float v = 0.0f;  // oracle order; -ffp-contract=off keeps mul and add separately rounded
        v = v + (float)up[x - C] * h0;
        v = v + (float)up[x] * h1;
        v = v + (float)up[x + C] * h2;
        v = v + (float)mid[x - C] * h3;
        v = v + (float)mid[x] * h4;
        v = v + (float)mid[x + C] * h5;
        v = v + (float)dn[x - C] * h6;
        v = v + (float)dn[x] * h7;
        v = v + (float)dn[x + C] * h8;
        v = v > 0.0f ? v : 0.0f;
        v = v < 255.0f ? v : 255.0f;
        out[x] = (uint8_t)(int)v;
    }
}

inline void one_row(const uint8_t* src0, uint8_t* dst0, int64_t pitch, int C,
                    const pconv_filter_t* f, int64_t y, int64_t c0, int64_t c1) {
    const int64_t b0 = c0 * C, n = (c1 - c0) * C;
    const uint8_t* mid = src0 + y * pitch + b0;
    const uint8_t* up = mid - pitch;
    const uint8_t* dn = mid + pitch;
    uint8_t* out = dst0 + y * pitch + b0;
    switch (f->kind) {
        case PCONV_KIND_SEP121: row_sep121(up, mid, dn, out, n, C); break;
        case PCONV_KIND_INT: row_int(up, mid, dn, out, n, C, f->w, f->shift); break;
        default: row_f32(up, mid, dn, out, n, C, f->h); break;
    }
}

}  // namespace

extern "C" int pconv_omp_max_threads(void) { return omp_get_max_threads(); }

extern "C" int pconv_cpu_convolve(const uint8_t* src0, uint8_t* dst0, int64_t pitch, int C,
                                  const pconv_filter_t* f, int64_t r0, int64_t r1, int64_t c0,
                                  int64_t c1, int threads) {
    if (!src0 || !dst0 || !f || C <= 0) return -1;
    if (r1 <= r0 || c1 <= c0) return 0;
    const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt) if (nt > 1 && (r1 - r0) > 1)
    for (int64_t y = r0; y < r1; ++y) one_row(src0, dst0, pitch, C, f, y, c0, c1);
    return 0;
}

// Temporal blocking on the CPU (the same idea as the GPU kernels, one level up the hierarchy): a
// thread takes a band of rows plus a k-deep halo into cache-sized private ping-pong buffers, advances it
// k iterations there (the halo absorbs the staleness; rows and columns on the image border keep their
// zero padding and never go stale), and writes the band back — two passes over memory per k iterations
// instead of 2k. Exactly the oracle's arithmetic, so the result is bit-identical.

extern "C" int64_t pconv_cpu_band_scratch_bytes(int64_t band_rows, int64_t cols, int C, int k) {
    const int64_t lpitch = ((cols + 2 * k + 2) * C + 63) / 64 * 64;
    return (band_rows + 2 * k + 2) * lpitch;
}

// One band: rows [r0,r1) x pixel columns [c0,c1) of an H x W image advance kk iterations. src0/dst0 point
// at pixel (0,0) of two images with the same pitch and a valid zero frame; the band's halo is read from src
// (it may belong to other bands, threads or ranks), only the band itself is written to dst. xa/xb: two
// scratch buffers of pconv_cpu_band_scratch_bytes(r1 - r0, c1 - c0, C, kk) bytes each.
extern "C" void pconv_cpu_band_advance(const uint8_t* src0, uint8_t* dst0, int64_t pitch, int64_t H, int64_t W,
                                       int C, const pconv_filter_t* f, int64_t r0, int64_t r1, int64_t c0,
                                       int64_t c1, int kk, uint8_t* sa, uint8_t* sb) {
    if (kk == 1) {   // nothing to block: straight from image to image, the halo is read in place
        for (int64_t y = r0; y < r1; ++y) one_row(src0, dst0, pitch, C, f, y, c0, c1);
        return;
    }
    const int64_t lo = std::max<int64_t>(0, r0 - kk), hi = std::min(H, r1 + kk);
    const int64_t bl = std::max<int64_t>(0, c0 - kk), bh = std::min(W, c1 + kk);
    const int64_t rows = hi - lo, cols = bh - bl;
    const int64_t lpitch = ((c1 - c0 + 2 * kk + 2) * C + 63) / 64 * 64;
    uint8_t* xa = sa + lpitch + C;    // local pixel (0,0) == image pixel (lo, bl)
    uint8_t* xb = sb + lpitch + C;
    // 1-pixel frame = the zero padding wherever the band touches the image border; at cut edges the frame is
    // only read by outputs that are discarded (the valid region shrinks by one per iteration)
    for (uint8_t* x : {xa, xb}) {
        std::memset(x - lpitch - C, 0, (size_t)((cols + 2) * C));
        std::memset(x + rows * lpitch - C, 0, (size_t)((cols + 2) * C));
        for (int64_t y = 0; y < rows; ++y) {
            std::memset(x + y * lpitch - C, 0, (size_t)C);
            std::memset(x + y * lpitch + cols * C, 0, (size_t)C);
        }
    }
    for (int64_t y = 0; y < rows; ++y)
        std::memcpy(xa + y * lpitch, src0 + (lo + y) * pitch + bl * C, (size_t)(cols * C));
    for (int j = 1; j <= kk; ++j) {
        // still exact after j iterations: j away from a cut edge, all the way to an image edge
        const int64_t ra = lo > 0 ? j : 0, re = hi < H ? rows - j : rows;
        const int64_t ca = bl > 0 ? j : 0, ce = bh < W ? cols - j : cols;
        for (int64_t y = ra; y < re; ++y) one_row(xa, xb, lpitch, C, f, y, ca, ce);
        std::swap(xa, xb);
    }
    for (int64_t y = r0; y < r1; ++y)
        std::memcpy(dst0 + y * pitch + c0 * C, xa + (y - lo) * lpitch + (c0 - bl) * C, (size_t)((c1 - c0) * C));
}

// Advance a padded image pair: pa/pb point at pixel (0,0) of two (H + 2) x pitch images whose 1-pixel frame
// is zero (and stays zero); the input is in pa. k <= 1: one pass over the whole image per iteration; k > 1:
// bands advance k iterations in cache-sized private buffers. Returns 0 / 1 = the result is in pa / pb.
extern "C" int pconv_cpu_run_padded(uint8_t* pa, uint8_t* pb, int64_t pitch, int64_t H, int64_t W, int C,
                                    const pconv_filter_t* f, int loops, int threads, int k) {
    if (!pa || !pb || !f || H <= 0 || W <= 0 || C <= 0 || loops < 0) return -1;
    const int nt = threads > 0 ? threads : omp_get_max_threads();
    if (k <= 1) {
        for (int t = 0; t < loops; ++t) {
            pconv_cpu_convolve(pa, pb, pitch, C, f, 0, H, 0, W, nt);
            std::swap(pa, pb);
        }
        return loops & 1;
    }
    // band height: both private buffers (band + 2k halo rows) should sit in ~1 MB of L2
    int64_t band = (int64_t)(1 << 20) / (2 * pitch) - 2 * k;
    band = std::max<int64_t>(band, 8);
    band = std::min<int64_t>(band, std::max<int64_t>((H + nt - 1) / nt, 8));
    const int64_t nbands = (H + band - 1) / band;
    const size_t sbytes = (size_t)pconv_cpu_band_scratch_bytes(band, W, C, k);
    std::vector<std::vector<uint8_t>> scratch((size_t)nt * 2);   // allocated once: no malloc in the passes
#pragma omp parallel num_threads(nt) if (nt > 1)
    {
        const int tid = omp_get_thread_num();
        scratch[(size_t)tid * 2].assign(sbytes, 0);
        scratch[(size_t)tid * 2 + 1].assign(sbytes, 0);
        for (int done = 0; done < loops;) {
            const int kk = std::min(k, loops - done);
            const uint8_t* from = (done / k) & 1 ? pb : pa;
            uint8_t* to = (done / k) & 1 ? pa : pb;
#pragma omp for schedule(dynamic)
            for (int64_t b = 0; b < nbands; ++b)
                pconv_cpu_band_advance(from, to, pitch, H, W, C, f, b * band, std::min(H, b * band + band), 0, W,
                                       kk, scratch[(size_t)tid * 2].data(), scratch[(size_t)tid * 2 + 1].data());
            done += kk;      // the implicit barrier of the loop separates the passes
        }
    }
    return ((loops + k - 1) / k) & 1;
}

extern "C" int64_t pconv_cpu_padded_pitch(int64_t W, int C) { return ((W * (int64_t)C + 2 * C + 63) / 64) * 64; }

// Dense image in, dense image out: padded ping-pong images around pconv_cpu_run_padded.
static int run_image_k(const uint8_t* src, uint8_t* dst, int64_t H, int64_t W, int C,
                       const pconv_filter_t* f, int loops, int threads, int k) {
    const int64_t rb = W * (int64_t)C;
    const int64_t pitch = pconv_cpu_padded_pitch(W, C);
    // Two zero-initialised padded buffers: the 1-pixel frame stays 0 forever = zero padding
    // (same device the reference uses: calloc'd (rows+2)x(cols+2) tiles, mpi/mpi_conv.c:113-124).
    std::vector<uint8_t> ga((size_t)((H + 2) * pitch), 0), gb((size_t)((H + 2) * pitch), 0);
    uint8_t* pa = ga.data() + pitch + C;   // pixel (0,0) of the global ping-pong images
    uint8_t* pb = gb.data() + pitch + C;
    for (int64_t y = 0; y < H; ++y) std::memcpy(pa + y * pitch, src + y * rb, (size_t)rb);
    const int which = pconv_cpu_run_padded(pa, pb, pitch, H, W, C, f, loops, threads, k);
    if (which < 0) return which;
    const uint8_t* last = which ? pb : pa;
    for (int64_t y = 0; y < H; ++y) std::memcpy(dst + y * rb, last + y * pitch, (size_t)rb);
    return 0;
}

static int run_image_blocked(const uint8_t* src, uint8_t* dst, int64_t H, int64_t W, int C,
                             const pconv_filter_t* f, int loops, int threads, int k) {
    return run_image_k(src, dst, H, W, C, f, loops, threads, k);
}

extern "C" int pconv_cpu_run_image_blocked(const uint8_t* src, uint8_t* dst, int64_t H, int64_t W,
                                           int C, const pconv_filter_t* f, int loops, int threads,
                                           int k) {
    if (!src || !dst || !f || H <= 0 || W <= 0 || C <= 0 || loops < 0 || k < 1) return -1;
    if (loops == 0) {
        std::memcpy(dst, src, (size_t)(H * W * C));
        return 0;
    }
    return run_image_blocked(src, dst, H, W, C, f, loops, threads, k);
}

// Out-of-core: the image never has to fit in memory. Each pass streams the file through row strips
// (strip + k-row halo, two ping-pong buffers sized by `mem_bytes`), advances a strip k iterations
// (same staleness argument as the temporal blocking above) and writes it to the other file of a
// ping-pong pair; ceil(loops/k) passes over the data in total. The reference cannot even address a
// 32768x32768 RGB image (32-bit offsets, mpi/mpi_conv.c:120,142); this path only needs disk space.
// Intermediate passes go to "<out>.pass" and the last one lands in `out_path`.
extern "C" int pconv_cpu_run_file(const char* in_path, const char* out_path, int64_t H, int64_t W,
                                  int C, const pconv_filter_t* f, int loops, int threads, int k,
                                  int64_t mem_bytes) {
    if (!in_path || !out_path || !f || H <= 0 || W <= 0 || C <= 0 || loops < 0) return -1;
    if (std::strcmp(in_path, out_path) == 0) return -22;   // -EINVAL: strips are written while others are still unread
    const int64_t rb = W * (int64_t)C;
    if (pconv_file_size(in_path) < H * rb) return -5;   // -EIO: short input
    const int64_t pitch = ((rb + 2 * C + 63) / 64) * 64;
    if (k < 1) k = 8;
    k = std::max(1, std::min(k, std::max(loops, 1)));
    // strip height from the memory budget: two buffers of (strip + 2k + 2) rows
    int64_t strip = (mem_bytes > 0 ? mem_bytes : ((int64_t)256 << 20)) / (2 * pitch) - 2 * k - 2;
    if (strip < 4 * k) strip = 4 * k;                   // never let the halo dominate
    strip = std::min(strip, H);
    const int npass = loops == 0 ? 1 : (loops + k - 1) / k;
    const std::string tmp = std::string(out_path) + ".pass";
    std::vector<uint8_t> la((size_t)((strip + 2 * k + 2) * pitch)), lb(la.size());
    int rc = pconv_create_file(out_path, H * rb);
    if (rc == 0 && npass > 1) rc = pconv_create_file(tmp.c_str(), H * rb);
    for (int pass = 0, done = 0; pass < npass && rc == 0; ++pass) {
        const int kk = loops == 0 ? 0 : std::min(k, loops - done);
        // the last pass must land in out_path: passes alternate backwards from there
        const bool to_out = ((npass - 1 - pass) & 1) == 0;
        const char* src_path = pass == 0 ? in_path : (to_out ? tmp.c_str() : out_path);
        const char* dst_path = to_out ? out_path : tmp.c_str();
        for (int64_t r0 = 0; r0 < H && rc == 0; r0 += strip) {
            const int64_t r1 = std::min(H, r0 + strip);
            const int64_t lo = std::max<int64_t>(0, r0 - kk), hi = std::min(H, r1 + kk);
            const int64_t rows = hi - lo;
            std::memset(la.data(), 0, (size_t)((rows + 2) * pitch));
            std::memset(lb.data(), 0, (size_t)((rows + 2) * pitch));
            uint8_t* xa = la.data() + pitch + C;   // local pixel (0,0) == image row lo
            uint8_t* xb = lb.data() + pitch + C;
            rc = pconv_read_rect(src_path, rb, lo, rows, 0, rb, xa, pitch);
            if (rc != 0) break;
            for (int j = 1; j <= kk; ++j) {
                const int64_t a = lo > 0 ? j : 0, e = hi < H ? rows - j : rows;
                pconv_cpu_convolve(xa, xb, pitch, C, f, a, e, 0, W, threads);
                std::swap(xa, xb);
            }
            rc = pconv_write_rect(dst_path, rb, r0, r1 - r0, 0, rb, xa + (r0 - lo) * pitch, pitch);
        }
        done += kk;
    }
    if (npass > 1) std::remove(tmp.c_str());
    return rc;
}

extern "C" int pconv_cpu_blocking_depth(int64_t H, int64_t W, int C, int loops) {
    // Temporal blocking pays when the two padded images do not fit in the last-level cache: then every plain pass
    // streams the image through memory, and k iterations per pass divide that traffic by k. Cache-resident images
    // are faster with plain passes (no band copies, no redundant halo rows).
    static const int64_t llc = [] {
        long v = sysconf(_SC_LEVEL3_CACHE_SIZE);
        if (v <= 0) v = sysconf(_SC_LEVEL2_CACHE_SIZE);
        return (int64_t)(v > 0 ? v : (32 << 20));
    }();
    if (const char* e = std::getenv("PCONV_CPU_BLOCK_DEPTH")) return std::max(1, std::atoi(e));
    const int64_t bytes = 2 * (H + 2) * pconv_cpu_padded_pitch(W, C);
    return (loops >= 4 && bytes > llc) ? std::min(loops, 6) : 1;
}

extern "C" int pconv_cpu_run_image(const uint8_t* src, uint8_t* dst, int64_t H, int64_t W, int C,
                                   const pconv_filter_t* f, int loops, int threads) {
    if (!src || !dst || !f || H <= 0 || W <= 0 || C <= 0 || loops < 0) return -1;
    if (loops == 0) {
        std::memcpy(dst, src, (size_t)(H * W * C));
        return 0;
    }
    return run_image_k(src, dst, H, W, C, f, loops, threads, pconv_cpu_blocking_depth(H, W, C, loops));
}
