// Device-side stencil core shared by the streaming and the resident kernels.
//
// Geometry (all kernels): a CTA of 512 threads (16 warps) owns a *region* of RH = 16*R rows x 512
// bytes (kSP) that lives in REGISTERS: warp w holds rows [w*R, (w+1)*R), lane l the 16-byte strip
// [16 l, 16 l + 16) of each of them (R uint4 per thread). Horizontal neighbours that fall outside
// a lane's strip come from the adjacent lane with warp shuffles; the row above / below a warp's
// band comes from the adjacent warp through a 32 KB shared-memory edge buffer (2 rows per warp per
// iteration) — the bulk of the data never touches shared memory between iterations. Rows are
// *streamed*: pushing row r+1 completes output row r. k iterations are fused per region (temporal
// blocking): the outer ring of the region goes stale by one pixel per iteration and the halo
// (k rows, HB >= k*C bytes) absorbs it, so the interior is bit-identical to k global passes.
//
// An interleaved image is a byte stream whose horizontal neighbour is C bytes away (C = 1 grey,
// 3 RGB): grey and RGB share every line of code below except the lane-shuffle patterns.
//
// Arithmetic paths (selected by the host-side classifier, csrc/host/filters.cpp; each is
// bit-identical to the fp32 oracle of csrc/host/oracle.cpp for the filters routed to it):
//   Sep121Path : [1 2 1]x[1 2 1]/16, 16-bit SWAR lanes, vertical pass first       ~15 ALU ops / 4 B
//   IntPath    : any integer taps, divisor 2^s, biased 16-bit SWAR IMAD lanes     ~30 ALU ops / 4 B
//   F32Path    : any fp32 taps, __fmul_rn/__fadd_rn in the oracle's tap order     ~20 ALU ops / B
// (Reference kernels for contrast: one thread per pixel, 9-27 scalar byte loads and FP64 DFMAs,
//  cuda/cuda_convolute.cu:9-47.)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/pconv_types.h"

namespace pconv {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kSP = 512;            // region width in bytes (32 lanes x 16-byte strips)
constexpr int kChunksPerRow = kSP / 16;  // 32 == one 16-byte chunk per lane
constexpr unsigned kFull = 0xffffffffu;

// Device copy of the filter with the constants the epilogues need (built by the launcher).
struct FilterDev {
    float h[9];
    int32_t w[9];
    int32_t shift;
    uint32_t bias2;     // IntPath: (bias << 16) | bias, bias = -S_min
    uint32_t negbias2;  // IntPath: packed (-bias) as s16x2
    uint32_t maxv2;     // IntPath: packed ((256 << shift) - 1)
    int32_t clamp_lo;   // result can go negative  -> clamp at 0
    int32_t clamp_hi;   // result can exceed 255   -> clamp at 255
    int32_t symmetric;  // F32Path: taps symmetric in both axes -> shared products (taps_symmetric)
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    return __byte_perm(a, b, s);
}
// packed signed 16-bit lane arithmetic (PTX ISA 8.0, sm_90+; native on sm_100a)
__device__ __forceinline__ uint32_t add_s16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t min_s16x2_relu(uint32_t a, uint32_t b) {  // max(min(a,b),0)
    uint32_t r;
    asm("min.s16x2.relu %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// 16-bit lanes: E = bytes (2,0) of a word, O = bytes (3,1); value in the low byte of each lane.
__device__ __forceinline__ uint32_t lanes_even(uint32_t w) { return w & 0x00ff00ffu; }
__device__ __forceinline__ uint32_t lanes_odd(uint32_t w) { return prmt(w, 0u, 0x4341u); }
// (hi lane of a) -> lo lane, (lo lane of b) -> hi lane: "the next 16-bit lane pair to the right".
__device__ __forceinline__ uint32_t lanes_shift(uint32_t a, uint32_t b) { return prmt(a, b, 0x5432u); }

// =============================================================================================
// Sep121Path — the reference's default Gaussian (mpi/mpi_conv.c:96,104; cuda_convolute.cu:12).
// Exactness: every partial sum is an integer <= 4080 < 2^16, result = sum >> 4.
//   vertical   V_r(x)  = p_{r-1}(x) + 2 p_r(x) + p_{r+1}(x)  as (U_{r-1}+U_r) + (U_r+U_{r+1})
//   horizontal out(x)  = V(x-C) + 2 V(x) + V(x+C)            on the 16-bit V lanes
// =============================================================================================
template <int C>
struct Sep121Path {
    static_assert(C == 1 || C == 3 || C == 4, "channel strides 1 (grey), 3 (RGB) and 4 (RGBA) are implemented");
    uint32_t uE[4], uO[4];  // unpacked lanes of the last pushed row
    uint32_t pE[4], pO[4];  // pair sum U_{r-1} + U_r

    __device__ __forceinline__ Sep121Path(const FilterDev&) {}

    // first row of a segment (row a-1): nothing to emit, nothing to accumulate yet
    __device__ __forceinline__ void prime0(const uint4& w) {
        const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { uE[j] = lanes_even(x[j]); uO[j] = lanes_odd(x[j]); }
    }
    // second row (row a)
    __device__ __forceinline__ void prime1(const uint4& w) {
        const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t nE = lanes_even(x[j]), nO = lanes_odd(x[j]);
            pE[j] = uE[j] + nE; pO[j] = uO[j] + nO;
            uE[j] = nE; uO[j] = nO;
        }
    }
    // push row r+1 when output row r lies outside the image (it stays zero): state update only
    __device__ __forceinline__ void skip(const uint4& w) { prime1(w); }
    // Result pack: sum << 4 puts the result byte into the high half of each 16-bit lane, one PRMT packs four.
    // LANES (kMaskLanes): the shift is a multiply by a per-lane factor, 16 in lanes whose chunk lies inside
    // the image and 0 in lanes outside — dead lanes then produce zero rows, at exactly the instruction count
    // of the unmasked code (IMAD with a register factor instead of IMAD.SHL).
    // push row r+1, return output row r
    __device__ __forceinline__ uint4 push(const uint4& w) {
        uint4 o;
        push_impl<false>(w, o, 16u);
        return o;
    }
    // push row r+1, output row r with the per-lane factor m16 (16: alive, 0: the lane is outside the image)
    __device__ __forceinline__ uint4 push_lanes(const uint4& w, uint32_t m16) {
        uint4 o;
        push_impl<true>(w, o, m16);
        return o;
    }

    template <bool LANES>
    __device__ __forceinline__ void push_impl(const uint4& w, uint4& dst, uint32_t m16) {
        const uint32_t x[4] = {w.x, w.y, w.z, w.w};
        uint32_t vE[4], vO[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t nE = lanes_even(x[j]), nO = lanes_odd(x[j]);
            const uint32_t qE = uE[j] + nE, qO = uO[j] + nO;  // U_r + U_{r+1}
            vE[j] = pE[j] + qE; vO[j] = pO[j] + qO;           // V_r  (<= 1020 per lane)
            pE[j] = qE; pO[j] = qO; uE[j] = nE; uO[j] = nO;
        }
        // Horizontal pass straight on the V lanes. vX[0] is the left neighbour lane's word 3,
        // vX[1..4] this lane's words, vX[5] the right neighbour lane's word 0: the shuffles only
        // depend on V, so all of them are in flight before any horizontal arithmetic starts.
        //   grey (C=1): out(x) = V(x-1) + 2V(x) + V(x+1)      RGB (C=3): out(x) = V(x-3) + 2V(x) + V(x+3)
        // The sums (<= 4080) are shifted LEFT by 4 (<= 65280, still inside the 16-bit lane) so the
        // result byte is the lane's high byte and one PRMT packs four of them: no masks, and the
        // shifts run on the FMA pipe (IMAD.SHL), which the profile shows is the idle one.
        uint32_t o[4];
        if (C == 1) {
            const uint32_t oL = __shfl_up_sync(kFull, vO[3], 1);    // V(x-1) for this lane's byte 0
            const uint32_t eR = __shfl_down_sync(kFull, vE[0], 1);  // V(x+1) for this lane's byte 15
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t oPrev = j ? vO[j - 1] : oL, eNext = j < 3 ? vE[j + 1] : eR;
                const uint32_t oE = 2u * vE[j] + lanes_shift(oPrev, vO[j]) + vO[j];
                const uint32_t oO = 2u * vO[j] + vE[j] + lanes_shift(vE[j], eNext);
                o[j] = LANES ? prmt(oE * m16, oO * m16, 0x7351u) : prmt(oE << 4, oO << 4, 0x7351u);
            }
        } else if (C == 4) {
            // RGBA: x +- 4 is the same lane of the adjacent word — no lane shifts at all
            const uint32_t eL = __shfl_up_sync(kFull, vE[3], 1), oL = __shfl_up_sync(kFull, vO[3], 1);
            const uint32_t eR = __shfl_down_sync(kFull, vE[0], 1), oR = __shfl_down_sync(kFull, vO[0], 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t oE = 2u * vE[j] + (j ? vE[j - 1] : eL) + (j < 3 ? vE[j + 1] : eR);
                const uint32_t oO = 2u * vO[j] + (j ? vO[j - 1] : oL) + (j < 3 ? vO[j + 1] : oR);
                o[j] = LANES ? prmt(oE * m16, oO * m16, 0x7351u) : prmt(oE << 4, oO << 4, 0x7351u);
            }
        } else {
            const uint32_t eL = __shfl_up_sync(kFull, vE[3], 1), oL = __shfl_up_sync(kFull, vO[3], 1);
            const uint32_t eR = __shfl_down_sync(kFull, vE[0], 1), oR = __shfl_down_sync(kFull, vO[0], 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ePrev = j ? vE[j - 1] : eL, oPrev = j ? vO[j - 1] : oL;
                const uint32_t eNext = j < 3 ? vE[j + 1] : eR, oNext = j < 3 ? vO[j + 1] : oR;
                // even x = 4j, 4j+2: x-3 -> (4j-3, 4j-1) = O lanes of word j-1; x+3 -> (4j+3, 4j+5)
                const uint32_t oE = 2u * vE[j] + oPrev + lanes_shift(vO[j], oNext);
                // odd x = 4j+1, 4j+3: x-3 -> (4j-2, 4j); x+3 -> (4j+4, 4j+6) = E lanes of word j+1
                const uint32_t oO = 2u * vO[j] + lanes_shift(ePrev, vE[j]) + eNext;
                o[j] = LANES ? prmt(oE * m16, oO * m16, 0x7351u) : prmt(oE << 4, oO << 4, 0x7351u);
            }
        }
        dst = make_uint4(o[0], o[1], o[2], o[3]);
    }
};

// =============================================================================================
// IntPath — arbitrary integer taps over 2^shift (sharpen, laplacian, emboss, sobel, ...).
// Scatter form: input row r+1 contributes taps k=2 to output row r, k=1 to row r+1, k=0 to row
// r+2; three rotating accumulators per 16-bit lane pair. Negative taps are handled by running
// the 32-bit IMAD on both lanes at once: inter-lane borrows cancel mod 2^32 as long as each final
// lane total is in [0, 65535], which the bias guarantees (bias = -S_min, bias + S_max <= 32767).
// =============================================================================================
template <int C>
struct IntPath {
    static_assert(C == 1 || C == 3 || C == 4, "channel strides 1 (grey), 3 (RGB) and 4 (RGBA) are implemented");
    uint32_t aE[3][4], aO[3][4];  // [slot][word]; slot roles rotate with the row phase
    const FilterDev& f;

    __device__ __forceinline__ IntPath(const FilterDev& fd) : f(fd) {}

    template <int K>
    __device__ __forceinline__ void tap_row(uint32_t (&accE)[4], uint32_t (&accO)[4],
                                            const uint32_t (&E)[6], const uint32_t (&O)[6],
                                            bool fresh) {
        // E/O index: 0 = left neighbour's word 3, 1..4 = own words, 5 = right neighbour's word 0
        const uint32_t wl = (uint32_t)f.w[3 * K + 0], wc = (uint32_t)f.w[3 * K + 1],
                       wr = (uint32_t)f.w[3 * K + 2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t e = fresh ? f.bias2 : accE[j];
            uint32_t o = fresh ? f.bias2 : accO[j];
            // lanes at x-C / x+C for the even / odd positions of word j; the lane shifts are only
            // built for taps that are not zero (warp-uniform branches on constant-bank weights)
            if (wl) {
                const uint32_t lE = C == 1 ? lanes_shift(O[j], O[j + 1]) : (C == 3 ? O[j] : E[j]);
                const uint32_t lO = C == 1 ? E[j + 1] : (C == 3 ? lanes_shift(E[j], E[j + 1]) : O[j]);
                e += wl * lE; o += wl * lO;
            }
            if (wc) { e += wc * E[j + 1]; o += wc * O[j + 1]; }
            if (wr) {
                const uint32_t rE = C == 1 ? O[j + 1] : (C == 3 ? lanes_shift(O[j + 1], O[j + 2]) : E[j + 2]);
                const uint32_t rO = C == 1 ? lanes_shift(E[j + 1], E[j + 2]) : (C == 3 ? E[j + 2] : O[j + 2]);
                e += wr * rE; o += wr * rO;
            }
            accE[j] = e; accO[j] = o;
        }
    }

    __device__ __forceinline__ void unpack(const uint4& w, uint32_t (&E)[6], uint32_t (&O)[6]) {
        const uint32_t x[4] = {w.x, w.y, w.z, w.w};
        const uint32_t xl = __shfl_up_sync(kFull, w.w, 1);    // left neighbour's last word
        const uint32_t xr = __shfl_down_sync(kFull, w.x, 1);  // right neighbour's first word
        E[0] = lanes_even(xl); O[0] = lanes_odd(xl);
#pragma unroll
        for (int j = 0; j < 4; ++j) { E[j + 1] = lanes_even(x[j]); O[j + 1] = lanes_odd(x[j]); }
        E[5] = lanes_even(xr); O[5] = lanes_odd(xr);
    }

    // PH = row phase (0,1,2): slot PH is completed by this row, (PH+1)%3 gets k=1, (PH+2)%3 k=0.
    template <int PH>
    __device__ __forceinline__ void prime0(const uint4& w) {  // row a-1: only k=0 -> out row a
        uint32_t E[6], O[6];
        unpack(w, E, O);
        tap_row<0>(aE[(PH + 2) % 3], aO[(PH + 2) % 3], E, O, true);
    }
    template <int PH>
    __device__ __forceinline__ void prime1(const uint4& w) {  // row a: k=1 -> a, k=0 -> a+1
        uint32_t E[6], O[6];
        unpack(w, E, O);
        tap_row<1>(aE[(PH + 1) % 3], aO[(PH + 1) % 3], E, O, false);
        tap_row<0>(aE[(PH + 2) % 3], aO[(PH + 2) % 3], E, O, true);
    }
    template <int PH>
    __device__ __forceinline__ uint4 push(const uint4& w) {
        uint32_t E[6], O[6];
        unpack(w, E, O);
        tap_row<2>(aE[PH], aO[PH], E, O, false);
        tap_row<1>(aE[(PH + 1) % 3], aO[(PH + 1) % 3], E, O, false);
        tap_row<0>(aE[(PH + 2) % 3], aO[(PH + 2) % 3], E, O, true);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t e = aE[PH][j], q = aO[PH][j];
            // per 16-bit lane: clamp(acc - bias, 0, (256<<s)-1) in two packed ops (sm_90+ s16x2 ISA)
            e = min_s16x2_relu(add_s16x2(e, f.negbias2), f.maxv2);
            q = min_s16x2_relu(add_s16x2(q, f.negbias2), f.maxv2);
            e >>= f.shift; q >>= f.shift;  // lane-crossing bits land in byte 1/3, never selected
            o[j] = prmt(e, q, 0x6240u);
        }
        return make_uint4(o[0], o[1], o[2], o[3]);
    }
};

// =============================================================================================
// F32Path — exact emulation of the oracle's fp32 sequence for arbitrary taps (box/9, edge/28,
// user floats). Scatter form keeps the oracle's order: row r-1 adds taps (0,0),(0,1),(0,2) to
// output row r first, then row r adds (1,*), then row r+1 adds (2,*). mul and add are separately
// rounded (__fmul_rn/__fadd_rn: never contracted to FMA). Truncation uses the 2^23 trick with a
// round-toward-zero add. (The reference's GPU kernel promotes every tap to FP64,
// cuda/cuda_convolute.cu:21,38-40; that is only exact for /16.)
// =============================================================================================
template <int C>
struct F32Path {
    static_assert(C == 1 || C == 3 || C == 4, "channel strides 1 (grey), 3 (RGB) and 4 (RGBA) are implemented");
    float acc[3][16];
    const FilterDev& f;

    __device__ __forceinline__ F32Path(const FilterDev& fd) : f(fd) {}

    __device__ __forceinline__ static float byte_to_float(uint32_t w, int i) {
        // place byte i in the low mantissa bits of 2^23, subtract 2^23: exact u8 -> f32 in 2 ops
        return __uint_as_float(prmt(w, 0x4b000000u, 0x7650u | (uint32_t)i)) - 8388608.0f;
    }

    // v index: 0..C-1 = left neighbour's last C bytes, C..C+15 = own bytes, then C right bytes
    __device__ __forceinline__ void convert(const uint4& w, float (&v)[16 + 2 * C]) {
        const uint32_t x[4] = {w.x, w.y, w.z, w.w};
        const uint32_t xl = __shfl_up_sync(kFull, w.w, 1);
        const uint32_t xr = __shfl_down_sync(kFull, w.x, 1);
#pragma unroll
        for (int i = 0; i < C; ++i) v[i] = byte_to_float(xl, 4 - C + i);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[C + i] = byte_to_float(x[i >> 2], i & 3);
#pragma unroll
        for (int i = 0; i < C; ++i) v[C + 16 + i] = byte_to_float(xr, i);
    }

    template <int K>
    __device__ __forceinline__ void tap_row(float (&a)[16], const float (&v)[16 + 2 * C], bool fresh) {
        const float hl = f.h[3 * K], hc = f.h[3 * K + 1], hr = f.h[3 * K + 2];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float ql = __fmul_rn(v[i], hl);             // byte at x - C
            const float qc = __fmul_rn(v[i + C], hc);         // x
            const float qr = __fmul_rn(v[i + 2 * C], hr);     // x + C
            float s = fresh ? ql : __fadd_rn(a[i], ql);       // 0 + q == q (sign of zero is moot)
            s = __fadd_rn(s, qc);
            a[i] = __fadd_rn(s, qr);
        }
    }

    template <int PH>
    __device__ __forceinline__ void prime0(const uint4& w) {
        float v[16 + 2 * C];
        convert(w, v);
        tap_row<0>(acc[(PH + 2) % 3], v, true);
    }
    template <int PH>
    __device__ __forceinline__ void prime1(const uint4& w) {
        float v[16 + 2 * C];
        convert(w, v);
        tap_row<1>(acc[(PH + 1) % 3], v, false);
        tap_row<0>(acc[(PH + 2) % 3], v, true);
    }
    // Filters symmetric in both axes (box, the reference's "edge"/28, gaussian_soft, any isotropic
    // kernel): taps (0,*) and (2,*) multiply the same pixels by the same weights, and so do the left
    // and right tap of a row, so each product is computed once and added where the oracle adds it —
    // same operands, same order, same roundings: 88 instead of 144 multiplies per 16 bytes.
    template <int PH>
    __device__ __forceinline__ void taps_symmetric(const float (&v)[16 + 2 * C]) {
        const float hAs = f.h[0], hAc = f.h[1], hBs = f.h[3], hBc = f.h[4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int b0 = 8 * half;
            {
                float t[8 + 2 * C], c[8];
#pragma unroll
                for (int i = 0; i < 8 + 2 * C; ++i) t[i] = __fmul_rn(v[b0 + i], hAs);
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = __fmul_rn(v[b0 + i + C], hAc);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = __fadd_rn(acc[PH][b0 + i], t[i]);                  // taps (2,0) (2,1) (2,2)
                    a = __fadd_rn(a, c[i]);
                    acc[PH][b0 + i] = __fadd_rn(a, t[i + 2 * C]);
                    const float z = __fadd_rn(t[i], c[i]);                       // taps (0,0) (0,1) (0,2), fresh
                    acc[(PH + 2) % 3][b0 + i] = __fadd_rn(z, t[i + 2 * C]);
                }
            }
            {
                float t[8 + 2 * C], c[8];
#pragma unroll
                for (int i = 0; i < 8 + 2 * C; ++i) t[i] = __fmul_rn(v[b0 + i], hBs);
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = __fmul_rn(v[b0 + i + C], hBc);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = __fadd_rn(acc[(PH + 1) % 3][b0 + i], t[i]);        // taps (1,0) (1,1) (1,2)
                    a = __fadd_rn(a, c[i]);
                    acc[(PH + 1) % 3][b0 + i] = __fadd_rn(a, t[i + 2 * C]);
                }
            }
        }
    }

    template <int PH>
    __device__ __forceinline__ uint4 push(const uint4& w) {
        float v[16 + 2 * C];
        convert(w, v);
        if (f.symmetric) {
            taps_symmetric<PH>(v);
        } else {
            tap_row<2>(acc[PH], v, false);
            tap_row<1>(acc[(PH + 1) % 3], v, false);
            tap_row<0>(acc[(PH + 2) % 3], v, true);
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t t[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float s = acc[PH][4 * j + b];
                if (f.clamp_lo) s = fmaxf(s, 0.0f);
                if (f.clamp_hi) s = fminf(s, 255.0f);
                t[b] = __float_as_uint(__fadd_rz(s, 8388608.0f));  // low byte = trunc(s)
            }
            o[j] = prmt(prmt(t[0], t[1], 0x0040u), prmt(t[2], t[3], 0x0040u), 0x5410u);
        }
        return make_uint4(o[0], o[1], o[2], o[3]);
    }
};

// ---------------------------------------------------------------------------------------------
// Band iteration. Warp w of a CTA owns region rows [w*R, (w+1)*R) and keeps them in REGISTERS
// (rows[i] = this lane's 16-byte strip of band row i) for all fused iterations; only the row above
// (`top`, last row of warp w-1) and the row below (`bottom`, first row of warp w+1) come from
// shared memory. One call advances the whole band by one iteration in place: pushing row i+1
// completes row i, whose raw value is no longer needed (the path state holds what is).
// Zero padding must stay zero, so rows / bytes outside the image are forced back to zero after every
// iteration — but only where that can happen. MODE is chosen PER WARP (warp-uniform):
//   0  every row and byte of the band lies inside the image: no masking at all (the common case)
//   1  rows all inside, some columns of the CTA outside (left / right image edge) and the edge cuts through
//      a 16-byte chunk: byte-wise column mask (four LOP3 per row)
//   2  some band rows outside the image (first / last warp of a CTA at the top / bottom image edge):
//      those rows are not computed at all (so that warp has LESS work), plus the column mask if needed
//   3  rows all inside, the image edge is chunk-aligned: every lane's chunk is either entirely inside or
//      entirely outside the image. The Gaussian path multiplies its sums by a per-lane factor (16 inside,
//      0 outside) where the unmasked code shifts by 4, so out-of-image lanes produce zero rows at no extra
//      instruction; the other paths keep their out-of-image lanes with one select per word.
// A CTA on the image border used to run the fully masked code in all of its warps and was 15-25 %
// slower than an interior CTA — and in a persistent grid everybody waits for the slowest CTA.
// ---------------------------------------------------------------------------------------------
constexpr int kMaskNone = 0, kMaskCols = 1, kMaskRows = 2, kMaskLanes = 3;

template <int MODE>
__device__ __forceinline__ uint4 apply_mask(uint4 o, const uint4& cmask, uint32_t rowmask, int i) {
    if (MODE == kMaskCols) {
        o.x &= cmask.x; o.y &= cmask.y; o.z &= cmask.z; o.w &= cmask.w;
    } else if (MODE == kMaskRows) {
        const bool v = (rowmask >> i) & 1u;
        o.x = v ? (o.x & cmask.x) : 0u; o.y = v ? (o.y & cmask.y) : 0u;
        o.z = v ? (o.z & cmask.z) : 0u; o.w = v ? (o.w & cmask.w) : 0u;
    }
    return o;
}

template <int C, int R, int MODE>
__device__ __forceinline__ void band_sep121(uint4 (&rows)[R], const uint4& top, const uint4& bottom,
                                            const uint4& cmask, uint32_t rowmask, bool need_cols,
                                            const FilterDev& f) {
    Sep121Path<C> p(f);
    p.prime0(top);
    p.prime1(rows[0]);
    const uint32_t m16 = cmask.x & 16u;    // kMaskLanes: a lane's mask is all ones (factor 16) or all zeros (0)
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const uint4& nxt = i + 1 < R ? rows[i + 1] : bottom;
        if (MODE == kMaskLanes) {
            rows[i] = p.push_lanes(nxt, m16);
        } else if (MODE != kMaskRows || ((rowmask >> i) & 1u)) {
            uint4 o = p.push(nxt);
            if (MODE == kMaskCols || (MODE == kMaskRows && need_cols)) {
                o.x &= cmask.x; o.y &= cmask.y; o.z &= cmask.z; o.w &= cmask.w;
            }
            rows[i] = o;
        } else {
            p.skip(nxt);
            rows[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

template <template <int> class Path, int C, int R, int MODE, int I>
struct PhasedSteps {
    static __device__ __forceinline__ void run(Path<C>& p, uint4 (&rows)[R], const uint4& bottom,
                                               const uint4& cmask, uint32_t rowmask) {
        if constexpr (I < R) {
            const uint4 nxt = (I + 1 < R) ? rows[(I + 1 < R) ? I + 1 : 0] : bottom;
            const uint4 o = p.template push<I % 3>(nxt);
            if constexpr (MODE == kMaskLanes) { if (cmask.x != 0u) rows[I] = o; }
            else rows[I] = apply_mask<MODE>(o, cmask, rowmask, I);
            PhasedSteps<Path, C, R, MODE, I + 1>::run(p, rows, bottom, cmask, rowmask);
        }
    }
};

template <template <int> class Path, int C, int R, int MODE>
__device__ __forceinline__ void band_phased(uint4 (&rows)[R], const uint4& top, const uint4& bottom,
                                            const uint4& cmask, uint32_t rowmask, const FilterDev& f) {
    Path<C> p(f);
    // priming rows behave like pushes with phases -2 and -1 (== 1 and 2 mod 3)
    p.template prime0<1>(top);
    p.template prime1<2>(rows[0]);
    PhasedSteps<Path, C, R, MODE, 0>::run(p, rows, bottom, cmask, rowmask);
}

template <int KIND, int C, int R, int MODE>
__device__ __forceinline__ void band_iterate(uint4 (&rows)[R], const uint4& top, const uint4& bottom,
                                             const uint4& cmask, uint32_t rowmask, bool need_cols,
                                             const FilterDev& f) {
    if constexpr (KIND == PCONV_KIND_SEP121) band_sep121<C, R, MODE>(rows, top, bottom, cmask, rowmask, need_cols, f);
    else if constexpr (KIND == PCONV_KIND_INT) band_phased<IntPath, C, R, MODE>(rows, top, bottom, cmask, rowmask, f);
    else band_phased<F32Path, C, R, MODE>(rows, top, bottom, cmask, rowmask, f);
}

// Shared-memory exchange of band boundary rows between adjacent warps: edge[parity][warp][0|1][lane]
template <int W>
struct EdgeBuffersT {
    uint4 row[2][W][2][32];
};
using EdgeBuffers = EdgeBuffersT<kWarps>;

// k fused iterations over a CTA region held in registers (W warps, R rows each). One __syncthreads
// per iteration: iteration `it` publishes into parity it&1 and everybody reads that parity after the
// barrier; the next iteration publishes into the other parity, so a fast warp never overwrites rows
// a slow warp is still reading. The masking mode is picked per warp (see band_iterate).
template <int KIND, int C, int R, int W = kWarps>
__device__ __forceinline__ void region_iterate(uint4 (&rows)[R], EdgeBuffersT<W>& eb, int k,
                                               const uint4& cmask, uint32_t rowmask, bool need_cols,
                                               const FilterDev& f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    constexpr uint32_t kAll = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
    // Does the image edge cut through some lane's chunk (CTA-uniform)? Only CTAs at a left / right image edge ask.
    bool partial = false;
    if (need_cols) {
        const uint32_t all = cmask.x & cmask.y & cmask.z & cmask.w, any = cmask.x | cmask.y | cmask.z | cmask.w;
        partial = __syncthreads_or(all != 0xffffffffu && any != 0u) != 0;
    }
    const int mode = (rowmask & kAll) != kAll ? kMaskRows
                     : (need_cols ? (partial ? kMaskCols : kMaskLanes) : kMaskNone);   // warp-uniform
    for (int it = 0; it < k; ++it) {
        const int par = it & 1;
        eb.row[par][warp][0][lane] = rows[0];
        eb.row[par][warp][1][lane] = rows[R - 1];
        __syncthreads();
        const uint4 top = warp > 0 ? eb.row[par][warp - 1][1][lane] : zero;
        const uint4 bottom = warp < W - 1 ? eb.row[par][warp + 1][0][lane] : zero;
        if (mode == kMaskNone) band_iterate<KIND, C, R, kMaskNone>(rows, top, bottom, cmask, rowmask, need_cols, f);
        else if (mode == kMaskLanes) band_iterate<KIND, C, R, kMaskLanes>(rows, top, bottom, cmask, rowmask, need_cols, f);
        else if (mode == kMaskCols) band_iterate<KIND, C, R, kMaskCols>(rows, top, bottom, cmask, rowmask, need_cols, f);
        else band_iterate<KIND, C, R, kMaskRows>(rows, top, bottom, cmask, rowmask, need_cols, f);
    }
}

}  // namespace pconv
