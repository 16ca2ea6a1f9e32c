// Filter construction + classification of the provably-exact fast path.
//
// The reference fixes its filter at compile time by commenting lines (mpi/mpi_conv.c:95-107);
// here filters are runtime values. A filter is always *defined* by its nine fp32 taps and the
// oracle's fp32 sequence (pconv_types.h). The INT / SEP121 kinds are only selected when the
// integer evaluation is mathematically identical to that sequence:
//   taps are n_i / 2^s with integer n_i  =>  every product p*h and every partial sum is a multiple
//   of 2^-s with magnitude < 2^24 * 2^-s, hence exactly representable, hence the fp32 sequence
//   equals S / 2^s with S = sum n_i p_i, regardless of FMA contraction or association.
//   trunc-toward-zero then saturate-to-[0,255] equals clamp(S >> s, 0, 255) (negatives clamp to 0
//   under both roundings).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pconv_host.h"

namespace {

const int32_t kGauss[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};

void classify(pconv_filter_t* f, bool integral, int shift) {
    f->kind = PCONV_KIND_F32;
    f->shift = 0;
    if (!integral) {
        std::memset(f->w, 0, sizeof(f->w));
        return;
    }
    int64_t sum_abs = 0;
    for (int i = 0; i < 9; ++i) sum_abs += std::llabs((long long)f->w[i]);
    // 16-bit SWAR lanes need (S_max - S_min) = sum|w|*255 <= 32767, and the packed clamp constant
    // (256 << shift) - 1 must fit a signed 16-bit lane: shift <= 7.
    if (sum_abs > 128 || shift < 0 || shift > 7) return;
    f->shift = shift;
    f->kind = PCONV_KIND_INT;
    if (shift == 4 && std::memcmp(f->w, kGauss, sizeof(kGauss)) == 0) f->kind = PCONV_KIND_SEP121;
}

}  // namespace

extern "C" int pconv_filter_from_int(const int32_t w[9], double div, pconv_filter_t* out) {
    if (!out || div == 0.0) return -1;
    for (int i = 0; i < 9; ++i) {
        out->w[i] = w[i];
        out->h[i] = (float)((double)w[i] / div);  // int / double -> float, as the reference
    }
    int shift = -1;
    for (int s = 0; s <= 7; ++s)
        if (div == (double)(1 << s)) shift = s;
    classify(out, shift >= 0, shift);
    return 0;
}

extern "C" int pconv_filter_from_float(const float h[9], pconv_filter_t* out) {
    if (!out) return -1;
    for (int i = 0; i < 9; ++i) out->h[i] = h[i];
    // smallest s such that every tap * 2^s is an integer
    for (int s = 0; s <= 7; ++s) {
        bool ok = true;
        for (int i = 0; i < 9 && ok; ++i) {
            float v = h[i] * (float)(1 << s);  // exact: power-of-two scaling
            if (v != std::nearbyint(v) || std::fabs(v) > 1024.f) ok = false;
        }
        if (ok) {
            for (int i = 0; i < 9; ++i) out->w[i] = (int32_t)(h[i] * (float)(1 << s));
            classify(out, true, s);
            return 0;
        }
    }
    classify(out, false, 0);
    return 0;
}
