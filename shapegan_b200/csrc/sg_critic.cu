// Small HBM-bound kernels of the hand-scheduled critic update (shapegan_b200/critic.py): the pieces of the WGAN-GP
// (train_hybrid_progressive_gan.py:102-111) that are not GEMMs.
//   sg_gp_interp   x^ = alpha real + (1 - alpha) fake per sample                                           (:103-105)
//   sg_gp_seed     per sample n = |g|_2 ; gp += weight (n - 1)^2 / B ; v = d gp / d g = 2 weight (n - 1) / (B n) g   (:110-111)
//   sg_critic_loss loss = mean s[0:B] - mean s[B:2B] (+ gp)                                                (train_wgan.py:66-68, :163)
#include <algorithm>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

__global__ void sg_gp_interp_kernel(const float4* __restrict__ real, const float4* __restrict__ fake, const float* __restrict__ alpha,
                                    float4* __restrict__ out, long long m4, long long total4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float a = __ldg(alpha + i / m4);
    const float4 r = __ldg(real + i), f = __ldg(fake + i);
    // the reference's expression order: alpha * real + ((1 - alpha) * fake)
    const float na = 1.f - a;
    float4 o;
    o.x = __fadd_rn(__fmul_rn(a, r.x), __fmul_rn(na, f.x)); o.y = __fadd_rn(__fmul_rn(a, r.y), __fmul_rn(na, f.y));
    o.z = __fadd_rn(__fmul_rn(a, r.z), __fmul_rn(na, f.z)); o.w = __fadd_rn(__fmul_rn(a, r.w), __fmul_rn(na, f.w));
    out[i] = o;
  }
}

__device__ __forceinline__ double block_sum_256(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += sm[w];
  __syncthreads();
  return t;
}

// one block per sample; the sample (m floats) is read twice, the second time out of L2
__global__ void __launch_bounds__(256) sg_gp_seed_kernel(const float4* __restrict__ g, float4* __restrict__ v, long long m4, int b, float weight,
                                                        double* __restrict__ gp_sum) {
  __shared__ double sm[8];
  const float4* gs = g + (long long)blockIdx.x * m4;
  float4* vs = v + (long long)blockIdx.x * m4;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < m4; i += 256) {
    const float4 x = gs[i];
    acc += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
  }
  const double n = sqrt(block_sum_256(acc, sm));
  const float coef = n > 0.0 ? (float)(2.0 * (double)weight * (n - 1.0) / ((double)b * n)) : 0.f;
  for (long long i = threadIdx.x; i < m4; i += 256) {
    const float4 x = gs[i];
    vs[i] = make_float4(coef * x.x, coef * x.y, coef * x.z, coef * x.w);
  }
  if (threadIdx.x == 0) atomicAdd(gp_sum, (double)weight * (n - 1.0) * (n - 1.0) / (double)b);
}

__global__ void __launch_bounds__(256) sg_critic_loss_kernel(const float* __restrict__ s, int b, const double* __restrict__ gp_sum, float* __restrict__ out) {
  __shared__ double sm[8];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < b; i += 256) { a += (double)s[i]; c += (double)s[b + i]; }
  a = block_sum_256(a, sm);
  c = block_sum_256(c, sm);
  if (threadIdx.x == 0) {
    const double gp = gp_sum ? *gp_sum : 0.0;
    out[0] = (float)(a / b - c / b + gp);      // mean D(fake) - mean D(real) + gp
    out[1] = (float)gp;
    out[2] = (float)(a / b);
    out[3] = (float)(c / b);
  }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_gp_interp(const float* real, const float* fake, const float* alpha, float* out, int b, int64_t m, void* stream) {
  if (b <= 0 || m <= 0) return 0;
  if (!real || !fake || !alpha || !out || (m & 3) || (((uintptr_t)real | (uintptr_t)fake | (uintptr_t)out) & 15)) return sg_fail(-1, "sg_gp_interp: bad arguments");
  const long long total4 = (long long)b * (m / 4);
  const int grid = (int)std::min<long long>((total4 + 255) / 256, 148 * 16);
  sg_gp_interp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)real, (const float4*)fake, alpha, (float4*)out, m / 4, total4);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_gp_seed(const float* g, float* v, int b, int64_t m, float weight, double* gp_sum, void* stream) {
  if (b <= 0 || m <= 0) return 0;
  if (!g || !v || !gp_sum || (m & 3) || (((uintptr_t)g | (uintptr_t)v) & 15)) return sg_fail(-1, "sg_gp_seed: bad arguments");
  sg_gp_seed_kernel<<<b, 256, 0, (cudaStream_t)stream>>>((const float4*)g, (float4*)v, m / 4, b, weight, gp_sum);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_critic_loss(const float* scores, int b, const double* gp_sum, float* out4, void* stream) {
  if (b <= 0 || !scores || !out4) return sg_fail(-1, "sg_critic_loss: bad arguments");
  sg_critic_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scores, b, gp_sum, out4);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
