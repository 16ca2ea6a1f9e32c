// Shared plain-C types used by the host (CPU) library and the CUDA library.
//
// Numerics contract (the oracle; see SURVEY.md §2.4 and reference mpi/mpi_conv.c:312-333):
//   val = 0.0f; for k in 0..2, l in 0..2 (row-major): val = fadd_rn(val, fmul_rn((float)src, h[k][l]))
//   out = saturate_u8(trunc_toward_zero(val))          (saturation is OUR defined extension: the
//   reference leaves float->uint8 of out-of-range values undefined)
// Pixels outside the image are 0 at every iteration (zero padding, reference np=1 semantics).
#ifndef PCONV_TYPES_H
#define PCONV_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Which arithmetic path is *provably bit-identical* to the fp32 oracle for this filter.
enum pconv_filter_kind {
    PCONV_KIND_F32 = 0,     // generic: exact fp32 emulation (mul_rn + add_rn in tap order)
    PCONV_KIND_INT = 1,     // integer taps, divisor 2^shift, sum|w| <= 128: exact integer SWAR path
    PCONV_KIND_SEP121 = 2,  // [1 2 1]x[1 2 1]/16 (the reference's default Gaussian): separable SWAR
};

typedef struct pconv_filter {
    float h[9];     // fp32 taps, row-major, exactly as the oracle multiplies them
    int32_t w[9];   // integer numerators (valid when kind != F32)
    int32_t shift;  // divisor = 2^shift (valid when kind != F32)
    int32_t kind;   // pconv_filter_kind
} pconv_filter_t;

#ifdef __cplusplus
}
#endif
#endif
