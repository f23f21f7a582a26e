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

// Per-DEVICE one-time initialisation.  cudaFuncSetAttribute, __constant__ uploads and device-symbol addresses belong to the current
// device, and one process may drive several (nn.DataParallel replicas run as threads of one process: train_hybrid_progressive_gan.py:62-71).
// Usage:  static sg::PerDevice once;  if (once.first()) { ...; once.done(); }     (first() is true until done() was called on this device)
namespace sg {
struct PerDevice {
  bool flag[64] = {};
  static int cur() { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64) ? d : 0; }
  bool first() const { return !flag[cur()]; }
  void done() { flag[cur()] = true; }
};
}  // namespace sg

