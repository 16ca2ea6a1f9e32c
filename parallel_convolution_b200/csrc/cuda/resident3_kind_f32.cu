// Split-phase resident kernels with a protocol warp, f32 path (experimental; see resident3_kernel.cuh).
#include "resident3_kernel.cuh"
namespace pconv {
ResidentFn pick_resident3_f32(int C, int R) { return pick_resident3_kind<PCONV_KIND_F32>(C, R); }
}  // namespace pconv
