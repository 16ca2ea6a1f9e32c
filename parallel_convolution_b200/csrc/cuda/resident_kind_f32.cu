// Instantiates resident_kernel<PCONV_KIND_F32, C, R> for every compiled channel stride and band height.
#include "resident_kernel.cuh"

namespace pconv {
ResidentFn pick_resident_f32(int C, int R) { return pick_resident_kind<PCONV_KIND_F32>(C, R); }
}  // namespace pconv
