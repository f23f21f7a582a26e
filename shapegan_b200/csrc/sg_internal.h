// Internal (non-ABI) helpers shared by the translation units of libsg_b200.
#pragma once
#include <cuda_runtime.h>

#include "../../include/sg_b200.h"

namespace sg {
int sg_fail(int code, const char* msg);   // records msg for sg_last_error(), returns code
int* sg_error_word();                     // device word written by kernel watchdogs
void sg_count_launch();                   // bookkeeping for bench.py's gpu_launches
}  // namespace sg

#define SG_CUDA_CHECK_LAUNCH()                                        \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return sg::sg_fail((int)e__, cudaGetErrorString(e__)); \
    sg::sg_count_launch();                                            \
  } while (0)
