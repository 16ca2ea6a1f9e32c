// Split-phase resident kernels of the separable-Gaussian path (all channel strides and band heights).
#include "resident2_kernel.cuh"
namespace pconv {
ResidentFn pick_resident2_sep(int C, int R) { return pick_resident2_kind<PCONV_KIND_SEP121>(C, R); }
}  // namespace pconv
