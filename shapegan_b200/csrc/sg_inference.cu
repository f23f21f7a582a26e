// HBM-bound helpers around the SDFNet inference consumers and the voxel data path:
//   sg_grid_sphere_index   cells of the util.get_voxel_coordinates(R) grid inside the 1.1 sphere (model/sdf_net.py:7-19)
//   sg_voxel_ingest        VoxelDataset.__getitem__'s clamp + rescale of a raw SDF voxel batch (datasets.py:16-23)
#include <algorithm>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

// One thread per grid cell s = (ix * r + iy) * r + iz.  The test is numpy's float32 arithmetic, operation by operation:
// np.linalg.norm(p, axis=1) = sqrt((x*x + y*y) + z*z) with every product / sum / root rounded to float32, compared with float32(1.1).
__global__ void sg_grid_sphere_index_kernel(int r, const float* __restrict__ axis, float radius, int* __restrict__ index_out, int* __restrict__ count) {
  const long long total = (long long)r * r * r;
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
    const int iz = (int)(s % r); const long long t = s / r;
    const int iy = (int)(t % r), ix = (int)(t / r);
    const float x = __ldg(axis + ix), y = __ldg(axis + r + iy), z = __ldg(axis + 2 * r + iz);
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    const bool inside = __fsqrt_rn(n2) < radius;
    // warp-aggregated append (order is irrelevant: the list addresses a scatter)
    const unsigned m = __ballot_sync(__activemask(), inside);
    if (inside) {
      const unsigned lane = threadIdx.x & 31u;
      const int leader = __ffs(m) - 1;
      int base = 0;
      if ((int)lane == leader) base = atomicAdd(count, __popc(m));
      base = __shfl_sync(m, base, leader);
      index_out[base + __popc(m & ((1u << lane) - 1u))] = (int)s;
    }
  }
}

// dst = clamp(src, -c, c) [/ c]   -- the reference does result.clamp_(-c, c); result /= c in float32 (datasets.py:19-22)
__global__ void sg_voxel_ingest_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long n4, float c, int rescale,
                                       const float* __restrict__ src_tail, float* __restrict__ dst_tail, int tail) {
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long i = i0; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldcs(src + i);
    v.x = fminf(fmaxf(v.x, -c), c); v.y = fminf(fmaxf(v.y, -c), c); v.z = fminf(fmaxf(v.z, -c), c); v.w = fminf(fmaxf(v.w, -c), c);
    if (rescale) { v.x = __fdiv_rn(v.x, c); v.y = __fdiv_rn(v.y, c); v.z = __fdiv_rn(v.z, c); v.w = __fdiv_rn(v.w, c); }
    dst[i] = v;
  }
  if (i0 < tail) {
    float v = fminf(fmaxf(src_tail[i0], -c), c);
    dst_tail[i0] = rescale ? __fdiv_rn(v, c) : v;
  }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_grid_sphere_index(int r, const float* axis, float radius, int32_t* index_out, int32_t* count, void* stream) {
  if (r <= 0 || r > 1024 || !axis || !index_out || !count) return sg_fail(-1, "sg_grid_sphere_index: bad arguments");
  const long long total = (long long)r * r * r;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  sg_grid_sphere_index_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(r, axis, radius, index_out, count);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_voxel_ingest(const float* src, float* dst, int64_t n, float clamp, int rescale, void* stream) {
  if (n <= 0) return 0;
  if (!src || !dst || !(clamp > 0.f)) return sg_fail(-1, "sg_voxel_ingest: bad arguments");
  if (((uintptr_t)src | (uintptr_t)dst) & 15) return sg_fail(-2, "sg_voxel_ingest: 16-byte aligned buffers");
  const long long n4 = n / 4;
  const int tail = (int)(n - n4 * 4);
  const int grid = (int)std::max<long long>(1, std::min<long long>((n4 + 255) / 256, 148 * 16));
  sg_voxel_ingest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (float4*)dst, n4, clamp, rescale, src + n4 * 4, dst + n4 * 4, tail);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
