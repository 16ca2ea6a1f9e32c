// Split-phase resident kernels with a protocol warp, int path (experimental; see resident3_kernel.cuh).
#include "resident3_kernel.cuh"
namespace pconv {
ResidentFn pick_resident3_int(int C, int R) { return pick_resident3_kind<PCONV_KIND_INT>(C, R); }
}  // namespace pconv
