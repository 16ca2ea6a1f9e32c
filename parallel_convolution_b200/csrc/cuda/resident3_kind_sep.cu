// Split-phase resident kernels with a protocol warp, sep path (experimental; see resident3_kernel.cuh).
#include "resident3_kernel.cuh"
namespace pconv {
ResidentFn pick_resident3_sep(int C, int R) { return pick_resident3_kind<PCONV_KIND_SEP121>(C, R); }
}  // namespace pconv
