// Instantiates persist_kernel<PCONV_KIND_F32, C, R, 16> for every compiled channel stride and band height.
#include "persist_kernel.cuh"

namespace pconv {
PersistFn pick_persist_f32_w16(int C, int R) { return pick_persist_c<PCONV_KIND_F32, 16>(C, R); }
}  // namespace pconv
