// Split-phase resident kernels of the exact-fp32 path (all channel strides and band heights).
#include "resident2_kernel.cuh"
namespace pconv {
ResidentFn pick_resident2_f32(int C, int R) { return pick_resident2_kind<PCONV_KIND_F32>(C, R); }
}  // namespace pconv
