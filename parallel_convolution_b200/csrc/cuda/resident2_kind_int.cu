// Split-phase resident kernels of the packed-integer path (all channel strides and band heights).
#include "resident2_kernel.cuh"
namespace pconv {
ResidentFn pick_resident2_int(int C, int R) { return pick_resident2_kind<PCONV_KIND_INT>(C, R); }
}  // namespace pconv
