// TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: libsg_b200 does not link libcuda)
// and the cp.async.bulk.tensor PTX wrappers.  Activations are NDHWC bf16 planes, so an implicit-GEMM A tile
// (128 output voxels x 64 channels of one filter tap) is ONE 5-D box [64 c, bw, bh, bd, bn]:
//   * stride-2 convolution  -> elementStrides {1,2,2,2,1}, box start (2*o0 - 1 + k) per axis
//   * transposed convolution -> unit strides, box start (q0 + tap offset)
//   * padding and ragged tails -> out-of-bounds coordinates (incl. negative) are zero-filled by the TMA unit
// and lands in shared memory as 128 rows x 128 B with the 128B swizzle the UMMA descriptors expect.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 tensor map with 128B swizzle.  dims/box/estr are innermost-first; strides_bytes has rank-1 entries (dims 1..).
inline bool tma_make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, const uint32_t* estr, bool swizzle = true) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn) return false;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

#ifdef __CUDACC__
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(m), "r"(c0), "r"(c1), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"((uint32_t)__cvta_generic_to_shared(bar))
      : "memory");
}
// multicast forms: the box lands at the same shared-memory offset in every CTA of `mask` and completes bytes on each one's barrier
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
      "l"(m), "r"(c0), "r"(c1), "r"((uint32_t)__cvta_generic_to_shared(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_mc(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint64_t* bar, unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;" ::"r"(dst),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"((uint32_t)__cvta_generic_to_shared(bar)), "h"(mask)
      : "memory");
}
// CTA-pair form: the transaction bytes are reported to the mbarrier at the same offset in the pair's LEADER CTA (rank 0)
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(bar) & 0xFEFFFFFFu;      // clear the peer bit: CTA 0's window
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(mbar)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(bar) & 0xFEFFFFFFu;
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(m), "r"(c0), "r"(c1), "r"(mbar)
               : "memory");
}
// shared -> global tile store (bulk async-group completion): the box is read from 128B-swizzled shared memory
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(m), "r"(c0), "r"(c1), "r"(c2),
               "r"(src)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed stores of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have been written out completely
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

// Decomposition of a tile of `rows` consecutive (n, z, y, x) grid points (x fastest) into a box; needs power-of-two
// grid extents.  Returns false when the tile is not a box.
struct TileBox { int bx, by, bz, bn; };
inline bool tile_box(int rows, int gx, int gy, int gz, TileBox* b) {
  if (!is_pow2(gx) || !is_pow2(gy) || !is_pow2(gz) || !is_pow2(rows)) return false;
  int r = rows;
  b->bx = gx < r ? gx : r; r /= b->bx;
  b->by = gy < r ? gy : r; r /= b->by;
  b->bz = gz < r ? gz : r; r /= b->bz;
  b->bn = r;
  return b->bx <= 128 && b->by <= 128 && b->bz <= 128 && b->bn <= 256;
}

}  // namespace sg
