// Internal helpers shared by the .cu translation units of libpconv_cuda.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pconv_cuda.h"
#include "stencil_core.cuh"

namespace pconv_rt {

int fail(const char* msg);                          // records msg, returns -1
int fail_cuda(cudaError_t e, const char* where);    // records "where: cudaGetErrorString", returns -2
int sm_count();
int max_optin_smem();
void count_launch();
bool pdl_enabled();   // PCONV_PDL=0 disables programmatic dependent launch
void make_filter_dev(const pconv_filter_t& f, pconv::FilterDev* out);

}  // namespace pconv_rt
