// HBM-bound kernels of the point-set GAN path (model/point_sdf_net.py): everything in SDFGenerator / PointNet that is
// not a Linear layer (those run on sg_igemm / sg_wgrad like every other dense contraction of the library).
//   sg_ln_act_fwd / _bwd      LayerNorm(C) (+ ReLU) over the channels of each row             point_sdf_net.py:64,110-112
//   sg_rows_add_vec           x[row, :] += v[row / seg_len, :]  (the per-shape z_lin(z) term)  point_sdf_net.py:105-109
//   sg_segment_colsum         out[s, :] = sum of the rows of segment s (its backward)
//   sg_segmax_fwd             PointNet's max over the points of a shape + argmax               point_sdf_net.py:40-41
//   sg_segmax_scatter / _gather   its backward (gradient to the arg-max rows) and the backward of that (WGAN-GP double backward)
// Tensors are bf16 plane tensors [P][rows][C] (P = 2: hi/lo split, value = hi + lo), C % 8 == 0.
#include <algorithm>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16lo_to_f(u.x); v[1] = bf16hi_to_f(u.x); v[2] = bf16lo_to_f(u.y); v[3] = bf16hi_to_f(u.y);
  v[4] = bf16lo_to_f(u.z); v[5] = bf16hi_to_f(u.z); v[6] = bf16lo_to_f(u.w); v[7] = bf16hi_to_f(u.w);
}
__device__ __forceinline__ void ld8p(const bf16* p, long long ps, int planes, float (&v)[8]) {
  ld8(p, v);
  if (planes == 2) {
    float w[8];
    ld8(p + ps, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += w[i];
  }
}
__device__ __forceinline__ void st8p(bf16* p, long long ps, int planes, const float (&v)[8]) {
  uint4 hi;
  hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]); hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = hi;
  if (planes == 2) {
    uint4 lo;
    lo.x = pack_bf16x2(v[0] - bf16lo_to_f(hi.x), v[1] - bf16hi_to_f(hi.x)); lo.y = pack_bf16x2(v[2] - bf16lo_to_f(hi.y), v[3] - bf16hi_to_f(hi.y));
    lo.z = pack_bf16x2(v[4] - bf16lo_to_f(hi.z), v[5] - bf16hi_to_f(hi.z)); lo.w = pack_bf16x2(v[6] - bf16lo_to_f(hi.w), v[7] - bf16hi_to_f(hi.w));
    *reinterpret_cast<uint4*>(p + ps) = lo;
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- LayerNorm (+ activation): one warp per row, C <= 1024 (C / 8 <= 128 pieces, 4 per lane)
constexpr int kLnMaxPieces = 4;

__global__ void __launch_bounds__(256) sg_ln_act_fwd_kernel(const bf16* x, long long x_ps, bf16* y, long long y_ps, int planes, long long rows, int c,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
                                                            float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int pieces = c >> 3;
  float v[kLnMaxPieces][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kLnMaxPieces; ++j) {
    const int pc = lane + 32 * j;
    if (pc < pieces) {
      ld8p(x + row * c + pc * 8, x_ps, planes, v[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[j][i];
    }
  }
  const float mean = warp_sum(s) / (float)c;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kLnMaxPieces; ++j)
    if (lane + 32 * j < pieces) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; q += d * d; }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)c + eps);      // biased variance, like torch.nn.LayerNorm
  if (lane == 0 && stats) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
#pragma unroll
  for (int j = 0; j < kLnMaxPieces; ++j) {
    const int pc = lane + 32 * j;
    if (pc < pieces) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = apply_act((v[j][i] - mean) * rstd * __ldg(gamma + pc * 8 + i) + __ldg(beta + pc * 8 + i), act);
      st8p(y + row * c + pc * 8, y_ps, planes, o);
    }
  }
}

// gx = rstd (ghat - mean_c(ghat) - xhat mean_c(ghat xhat)), ghat = gamma * (gy * act'(y)); sums[0..c) += gy' (dbeta), [c..2c) += gy' xhat (dgamma)
__global__ void __launch_bounds__(256) sg_ln_act_bwd_kernel(const bf16* gy, long long gy_ps, const bf16* y, long long y_ps, const bf16* x, long long x_ps,
                                                            bf16* gx, long long gx_ps, int planes, long long rows, int c, const float* __restrict__ gamma,
                                                            int act, const float* __restrict__ stats, float* __restrict__ part) {
  extern __shared__ float sm[];            // [2][c] per-block partial sums of dbeta / dgamma
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int pieces = c >> 3;
  for (long long row = blockIdx.x * 8LL + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float gh[kLnMaxPieces][8], xh[kLnMaxPieces][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxPieces; ++j) {
      const int pc = lane + 32 * j;
      if (pc < pieces) {
        float g[8], yy[8], xx[8];
        ld8p(gy + row * c + pc * 8, gy_ps, planes, g);
        ld8p(x + row * c + pc * 8, x_ps, planes, xx);
        if (act != ACT_NONE) ld8p(y + row * c + pc * 8, y_ps, planes, yy);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float gp = act != ACT_NONE ? g[i] * act_grad_from_output(yy[i], act) : g[i];
          const float xhat = (xx[i] - mean) * rstd;
          atomicAdd(&sm[pc * 8 + i], gp);
          atomicAdd(&sm[c + pc * 8 + i], gp * xhat);
          const float ghat = gp * __ldg(gamma + pc * 8 + i);
          gh[j][i] = ghat; xh[j][i] = xhat;
          s1 += ghat; s2 += ghat * xhat;
        }
      }
    }
    const float m1 = warp_sum(s1) / (float)c, m2 = warp_sum(s2) / (float)c;
#pragma unroll
    for (int j = 0; j < kLnMaxPieces; ++j) {
      const int pc = lane + 32 * j;
      if (pc < pieces) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rstd * (gh[j][i] - m1 - xh[j][i] * m2);
        st8p(gx + row * c + pc * 8, gx_ps, planes, o);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) part[(long long)blockIdx.x * 2 * c + i] = sm[i];     // deterministic: reduced by sg_ln_reduce
}

__global__ void sg_ln_reduce_kernel(const float* __restrict__ part, int blocks, int c2, float* __restrict__ gbeta, float* __restrict__ ggamma, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c2) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += (double)part[(long long)b * c2 + i];
  if (i < c) gbeta[i] = (float)s; else ggamma[i - c] = (float)s;
}

// ---- per-segment vector add and its backward
__global__ void sg_rows_add_vec_kernel(const bf16* x, long long x_ps, const float* __restrict__ v, bf16* y, long long y_ps, int planes, long long rows, int c,
                                       long long seg_len) {
  const int pieces = c >> 3;
  const long long total = rows * pieces;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / pieces;
    const int pc = (int)(i - row * pieces);
    float a[8];
    ld8p(x + row * c + pc * 8, x_ps, planes, a);
    const float* vv = v + (row / seg_len) * c + pc * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += __ldg(vv + k);
    st8p(y + row * c + pc * 8, y_ps, planes, a);
  }
}

// out[s][c] = sum over the seg_len rows of segment s.  grid (segments, c / 64, row splits): few long segments (the hybrid GAN's 64^3 grids:
// 2 segments of 262144 rows) are split over rows, every block writes its partial to part[split][s][c], sg_segment_colsum_fold adds the
// splits in a fixed order (deterministic, no atomics).  splits == 1 writes `out` directly.
__global__ void __launch_bounds__(256) sg_segment_colsum_kernel(const bf16* x, long long x_ps, int planes, int c, long long seg_len, float* __restrict__ out,
                                                                int segs) {
  __shared__ float sm[8][64];
  const int seg = blockIdx.x, c0 = blockIdx.y * 64;
  const int lane8 = threadIdx.x & 7, rgrp = threadIdx.x >> 3;      // 8 pieces of 8 channels x 32 row groups
  const long long per = (seg_len + gridDim.z - 1) / gridDim.z;
  const long long r0 = blockIdx.z * per, r1 = min(seg_len, r0 + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int ch = c0 + lane8 * 8;
  if (ch < c)
    for (long long r = r0 + rgrp; r < r1; r += 32) {
      float a[8];
      ld8p(x + ((long long)seg * seg_len + r) * c + ch, x_ps, planes, a);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += a[k];
    }
  // reduce the 32 row groups: 4 per warp (shuffle over lanes with the same lane8), then 8 warps through shared memory
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8); acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16); }
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l < 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[warp][l * 8 + k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < c) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
    out[((long long)blockIdx.z * segs + seg) * c + c0 + threadIdx.x] = s;
  }
}

__global__ void sg_segment_colsum_fold_kernel(const float* __restrict__ part, int splits, long long n, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(long long)k * n + i];
  out[i] = s;
}

// ---- segment max (PointNet pooling): out[s][c] = max over the rows of segment s, arg[s][c] = the (first) row that attains it
__global__ void __launch_bounds__(256) sg_segmax_fwd_kernel(const bf16* x, long long x_ps, int planes, int c, long long seg_len, bf16* out, long long out_ps,
                                                            int* __restrict__ arg) {
  __shared__ float smv[8][64];
  __shared__ int smi[8][64];
  const int seg = blockIdx.x, c0 = blockIdx.y * 64;
  const int lane8 = threadIdx.x & 7, rgrp = threadIdx.x >> 3;
  float best[8]; int bidx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bidx[k] = 0x7fffffff; }
  const int ch = c0 + lane8 * 8;
  if (ch < c)
    for (long long r = rgrp; r < seg_len; r += 32) {
      float a[8];
      ld8p(x + ((long long)seg * seg_len + r) * c + ch, x_ps, planes, a);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (a[k] > best[k] || (a[k] == best[k] && (int)r < bidx[k])) { best[k] = a[k]; bidx[k] = (int)r; }
    }
  auto merge = [&](float ov, int oi, float& v, int& i) { if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; } };
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best[k], o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx[k], o);
      merge(ov, oi, best[k], bidx[k]);
    }
  }
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l < 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { smv[warp][l * 8 + k] = best[k]; smi[warp][l * 8 + k] = bidx[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < c) {
    float v = smv[0][threadIdx.x]; int i = smi[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 8; ++w) merge(smv[w][threadIdx.x], smi[w][threadIdx.x], v, i);
    const long long o = (long long)seg * c + c0 + threadIdx.x;
    arg[o] = i;
    // copy the planes of the winning row (exact: no re-rounding of hi / lo)
    const long long src = ((long long)seg * seg_len + i) * c + c0 + threadIdx.x;
    out[o] = x[src];
    if (planes == 2) out[o + out_ps] = x[src + x_ps];
  }
}

// scatter: big[seg*seg_len + arg[s][c]][c] = small[s][c], zero elsewhere (big must be zero-filled);  gather: small[s][c] = big[...][c]
__global__ void sg_segmax_move_kernel(bf16* big, long long big_ps, bf16* small, long long small_ps, int planes, int segs, int c, long long seg_len,
                                      const int* __restrict__ arg, int gather) {
  const long long total = (long long)segs * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / c;
    const int ch = (int)(i - s * c);
    const long long b = (s * seg_len + arg[i]) * c + ch;
    for (int pl = 0; pl < planes; ++pl) {
      if (gather) small[i + pl * small_ps] = big[b + pl * big_ps];
      else big[b + pl * big_ps] = small[i + pl * small_ps];
    }
  }
}

}  // namespace sg

using namespace sg;
#define ST(s) ((cudaStream_t)(s))

extern "C" int sg_ln_act_fwd(const void* x, int64_t x_ps, void* y, int64_t y_ps, int planes, int64_t rows, int c, const float* gamma, const float* beta,
                             float eps, int act, float* stats, void* stream) {
  if (rows <= 0) return 0;
  if (!x || !y || !gamma || !beta || (c & 7) || c > 8 * 32 * kLnMaxPieces) return sg_fail(-1, "sg_ln_act_fwd: C must be a multiple of 8, <= 1024");
  sg_ln_act_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>((const bf16*)x, x_ps, (bf16*)y, y_ps, planes, rows, c, gamma, beta, eps, act, stats);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_ln_act_bwd(const void* gy, int64_t gy_ps, const void* y, int64_t y_ps, const void* x, int64_t x_ps, void* gx, int64_t gx_ps, int planes,
                             int64_t rows, int c, const float* gamma, int act, const float* stats, float* gbeta, float* ggamma, float* workspace,
                             int64_t workspace_floats, void* stream) {
  if (rows <= 0) return 0;
  if (!gy || !x || !gx || !gamma || !stats || !gbeta || !ggamma || !workspace || (c & 7) || c > 8 * 32 * kLnMaxPieces) return sg_fail(-1, "sg_ln_act_bwd: bad arguments");
  int blocks = (int)std::min<long long>((rows + 7) / 8, 296);
  blocks = (int)std::min<long long>(blocks, workspace_floats / (2 * c));
  if (blocks < 1) return sg_fail(-2, "sg_ln_act_bwd: workspace too small (>= 2*C floats per block)");
  sg_ln_act_bwd_kernel<<<blocks, 256, 2 * c * sizeof(float), ST(stream)>>>((const bf16*)gy, gy_ps, (const bf16*)y, y_ps, (const bf16*)x, x_ps, (bf16*)gx, gx_ps,
                                                                          planes, rows, c, gamma, act, stats, workspace);
  SG_CUDA_CHECK_LAUNCH();
  sg_ln_reduce_kernel<<<(2 * c + 255) / 256, 256, 0, ST(stream)>>>(workspace, blocks, 2 * c, gbeta, ggamma, c);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_rows_add_vec(const void* x, int64_t x_ps, const float* v, void* y, int64_t y_ps, int planes, int64_t rows, int c, int64_t seg_len, void* stream) {
  if (rows <= 0) return 0;
  if (!x || !v || !y || (c & 7) || seg_len <= 0) return sg_fail(-1, "sg_rows_add_vec: bad arguments");
  const long long total = rows * (c >> 3);
  sg_rows_add_vec_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, ST(stream)>>>((const bf16*)x, x_ps, v, (bf16*)y, y_ps, planes, rows, c, seg_len);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t sg_segment_colsum_workspace(int segs, int c, int64_t seg_len) {
  const long long blocks = (long long)segs * ((c + 63) / 64);
  long long splits = std::max<long long>(1, std::min<long long>(296 / std::max<long long>(blocks, 1), seg_len / 2048));
  return splits > 1 ? (size_t)splits * segs * c * sizeof(float) : 0;
}

extern "C" int sg_segment_colsum(const void* x, int64_t x_ps, int planes, int segs, int c, int64_t seg_len, float* out, float* workspace, void* stream) {
  if (segs <= 0) return 0;
  if (!x || !out || (c & 7) || seg_len <= 0) return sg_fail(-1, "sg_segment_colsum: bad arguments");
  const size_t ws = sg_segment_colsum_workspace(segs, c, seg_len);
  const int splits = ws ? (int)(ws / ((size_t)segs * c * sizeof(float))) : 1;
  if (splits > 1 && !workspace) return sg_fail(-2, "sg_segment_colsum: few long segments need a workspace (sg_segment_colsum_workspace)");
  dim3 grid(segs, (c + 63) / 64, splits);
  sg_segment_colsum_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)x, x_ps, planes, c, seg_len, splits > 1 ? workspace : out, segs);
  SG_CUDA_CHECK_LAUNCH();
  if (splits > 1) {
    const long long n = (long long)segs * c;
    sg_segment_colsum_fold_kernel<<<(int)((n + 255) / 256), 256, 0, ST(stream)>>>(workspace, splits, n, out);
    SG_CUDA_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int sg_segmax_fwd(const void* x, int64_t x_ps, int planes, int segs, int c, int64_t seg_len, void* out, int64_t out_ps, int32_t* arg, void* stream) {
  if (segs <= 0) return 0;
  if (!x || !out || !arg || (c & 7) || seg_len <= 0 || seg_len >= (1LL << 31)) return sg_fail(-1, "sg_segmax_fwd: bad arguments");
  dim3 grid(segs, (c + 63) / 64);
  sg_segmax_fwd_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)x, x_ps, planes, c, seg_len, (bf16*)out, out_ps, arg);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_segmax_move(void* big, int64_t big_ps, void* small, int64_t small_ps, int planes, int segs, int c, int64_t seg_len, const int32_t* arg, int gather,
                              void* stream) {
  if (segs <= 0) return 0;
  if (!big || !small || !arg) return sg_fail(-1, "sg_segmax_move: null");
  const long long total = (long long)segs * c;
  sg_segmax_move_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, ST(stream)>>>((bf16*)big, big_ps, (bf16*)small, small_ps, planes, segs, c, seg_len, arg, gather);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
