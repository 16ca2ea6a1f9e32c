// The golden oracle: the reference's single-process semantics (zero-padded, every pixel
// computed, true 8-neighbour dependence) written as plainly as possible.
// Compiled with -O1 -ffp-contract=off so that the fp32 sequence is exactly
//   val = val + (float)p * h   (one rounding for the product, one for the sum)
// which is what x86-64 gcc produces for mpi/mpi_conv.c:312-333 without -march.
// Nothing here is fast on purpose; the CPU engine (cpu_engine.cpp) and every CUDA path are
// tested byte-for-byte against this function.
#include <cstdint>
#include <cstring>

#include "pconv_host.h"

namespace {

inline uint8_t saturate_trunc(float v) {
    if (!(v > 0.0f)) return 0;  // negatives, -0, NaN
    if (v >= 255.0f) return 255;
    return (uint8_t)(int)v;  // trunc toward zero
}

void one_pass(const uint8_t* src, uint8_t* dst, int64_t H, int64_t W, int C, const float* h) {
    const int64_t rb = W * (int64_t)C;
    for (int64_t y = 0; y < H; ++y) {
        for (int64_t x = 0; x < W; ++x) {
            for (int ch = 0; ch < C; ++ch) {
                volatile float val = 0.0f;  // volatile: forbid re-association / vector tricks
                for (int k = 0; k < 3; ++k) {
                    for (int l = 0; l < 3; ++l) {
                        const int64_t yy = y + k - 1, xx = x + l - 1;
                        float p = 0.0f;
                        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                            p = (float)src[yy * rb + xx * C + ch];
                        volatile float prod = p * h[3 * k + l];
                        val = val + prod;
                    }
                }
                dst[y * rb + x * C + ch] = saturate_trunc(val);
            }
        }
    }
}

}  // namespace

extern "C" int pconv_oracle_run(const uint8_t* src, uint8_t* dst, uint8_t* scratch, int64_t H,
                                int64_t W, int C, const pconv_filter_t* f, int loops) {
    if (!src || !dst || !scratch || !f || H <= 0 || W <= 0 || C <= 0 || loops < 0) return -1;
    const int64_t n = H * W * (int64_t)C;
    if (loops == 0) {
        std::memcpy(dst, src, (size_t)n);
        return 0;
    }
    // Arrange the ping-pong so the final pass lands in dst (any parity; cf. reference defect D4).
    uint8_t* a = (loops % 2 == 1) ? dst : scratch;
    uint8_t* b = (loops % 2 == 1) ? scratch : dst;
    one_pass(src, a, H, W, C, f->h);
    for (int t = 1; t < loops; ++t) {
        one_pass(a, b, H, W, C, f->h);
        uint8_t* tmp = a;
        a = b;
        b = tmp;
    }
    return 0;
}
