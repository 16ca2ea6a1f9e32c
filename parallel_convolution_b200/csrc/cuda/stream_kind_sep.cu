// Instantiates stream_kernel<PCONV_KIND_SEP121, C, R> for every compiled channel stride and band height.
#include "stream_kernel.cuh"

namespace pconv {
KernelFn pick_stream_sep(int C, int R) { return pick_stream_kind<PCONV_KIND_SEP121>(C, R); }
}  // namespace pconv
