// HBM-bound kernels of the hot path: activation backward, train-mode BatchNorm (statistics / apply / backward),
// the single-channel col2im of ConvTranspose3d(C->1), row dot products (C->1 linear / k4s1 conv on 4^3),
// plane <-> fp32 conversion, SDFNet input assembly, fade-in blend, and the fused optimizer updates.
// All tensors are NDHWC "plane" tensors viewed as [rows, C] (C % 8 == 0): every access is a coalesced 16-byte piece.
#include <algorithm>

#include <cstdlib>
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

__device__ __forceinline__ void load8(const bf16* base, long long ps, int planes, long long off, float (&v)[8]) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(base + off));
  v[0] = bf16lo_to_f(h.x); v[1] = bf16hi_to_f(h.x); v[2] = bf16lo_to_f(h.y); v[3] = bf16hi_to_f(h.y);
  v[4] = bf16lo_to_f(h.z); v[5] = bf16hi_to_f(h.z); v[6] = bf16lo_to_f(h.w); v[7] = bf16hi_to_f(h.w);
  if (planes == 2) {
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(base + ps + off));
    v[0] += bf16lo_to_f(l.x); v[1] += bf16hi_to_f(l.x); v[2] += bf16lo_to_f(l.y); v[3] += bf16hi_to_f(l.y);
    v[4] += bf16lo_to_f(l.z); v[5] += bf16hi_to_f(l.z); v[6] += bf16lo_to_f(l.w); v[7] += bf16hi_to_f(l.w);
  }
}
__device__ __forceinline__ void store8(bf16* base, long long ps, int planes, long long off, const float (&v)[8]) {
  uint4 hi;
  hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
  hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(base + off) = hi;
  if (planes == 2) {
    uint4 lo;
    lo.x = pack_bf16x2(v[0] - bf16lo_to_f(hi.x), v[1] - bf16hi_to_f(hi.x));
    lo.y = pack_bf16x2(v[2] - bf16lo_to_f(hi.y), v[3] - bf16hi_to_f(hi.y));
    lo.z = pack_bf16x2(v[4] - bf16lo_to_f(hi.z), v[5] - bf16hi_to_f(hi.z));
    lo.w = pack_bf16x2(v[6] - bf16lo_to_f(hi.w), v[7] - bf16hi_to_f(hi.w));
    *reinterpret_cast<uint4*>(base + ps + off) = lo;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Column-reducing kernels: thread (r, p) owns piece column p (8 channels) and strides over rows, so the per-channel
// sums stay in registers; one shared-memory fold + one double atomicAdd per channel per block.
// blockDim.x = P8 * R with P8 = C/8.
template <int NACC>
__device__ __forceinline__ void fold_and_emit(float (&acc)[NACC][8], int p8, int r, int P8, int R, double* out, int c_total) {
  extern __shared__ float red[];   // [NACC][R][P8*8]
  const int C = P8 * 8;
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(a * R + r) * C + p8 * 8 + j] = acc[a][j];
  __syncthreads();
  for (int i = threadIdx.x; i < NACC * C; i += blockDim.x) {
    const int a = i / C, c = i - a * C;
    float s = 0.f;
    for (int rr = 0; rr < R; ++rr) s += red[(a * R + rr) * C + c];
    atomicAdd(out + (size_t)a * c_total + c, (double)s);
  }
}

struct ColArgs {
  const bf16 *ga, *y, *x; long long ga_ps, y_ps, x_ps;
  bf16* g; long long g_ps;
  int planes; long long rows; int c, act;
  const float *mean, *invstd;
  double* sums;
};

// mode 0: bn stats (sum x, sum x^2 of `x`)      mode 1: act-bwd  g = ga*act'(y) (+ sum g)
// mode 2: bn-bwd reduce (sum g, sum g*xhat)
template <int MODE>
__global__ void sg_colreduce_kernel(const ColArgs a, int P8, int R, long long rows_per_block) {
  const int p8 = threadIdx.x % P8, r = threadIdx.x / P8;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(a.rows, r0 + rows_per_block);
  float mean[8], istd[8];
  if (MODE == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean[j] = a.mean[p8 * 8 + j]; istd[j] = a.invstd[p8 * 8 + j]; }
  }
  for (long long row = r0 + r; row < r1; row += R) {
    const long long off = row * a.c + p8 * 8;
    if (MODE == 0) {
      float v[8];
      load8(a.x, a.x_ps, a.planes, off, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += v[j]; acc[1][j] += v[j] * v[j]; }
    } else {
      float g[8], y[8];
      load8(a.ga, a.ga_ps, a.planes, off, g);
      if (a.act != ACT_NONE) {
        load8(a.y, a.y_ps, a.planes, off, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= act_grad_from_output(y[j], a.act);
      }
      if (MODE == 1) {
        if (a.g) store8(a.g, a.g_ps, a.planes, off, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[0][j] += g[j];
      } else {
        float x[8];
        load8(a.x, a.x_ps, a.planes, off, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[0][j] += g[j]; acc[1][j] += g[j] * (x[j] - mean[j]) * istd[j]; }
      }
    }
  }
  if (a.sums) fold_and_emit<2>(acc, p8, r, P8, R, a.sums, a.c);
}

__global__ void sg_bn_finalize_kernel(const double* sums, long long rows, int c, float eps, float momentum, float* mean,
                                      float* invstd, float* running_mean, float* running_var) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double m = sums[i] / (double)rows;
  double var = sums[c + i] / (double)rows - m * m;
  if (var < 0) var = 0;
  mean[i] = (float)m;
  invstd[i] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (float)m;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
  }
}

// y = act((x - mean) * invstd * gamma + beta)
__global__ void sg_bn_apply_kernel(const bf16* x, long long x_ps, bf16* y, long long y_ps, int planes, long long rows, int c,
                                   const float* mean, const float* invstd, const float* gamma, const float* beta, int act) {
  const int P8 = c / 8;
  const long long pieces = rows * P8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int p8 = (int)(i % P8);
    const long long off = (i / P8) * c + p8 * 8;
    float v[8];
    load8(x, x_ps, planes, off, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = p8 * 8 + j;
      v[j] = apply_act((v[j] - mean[ch]) * invstd[ch] * gamma[ch] + beta[ch], act);
    }
    store8(y, y_ps, planes, off, v);
  }
}

// gx = gamma*invstd*(g - sum_g/n - xhat*sum_gxhat/n),  g = ga*act'(y).  sums == NULL: eval mode (running statistics are constants of the
// graph): gx = gamma*invstd*g
__global__ void sg_bn_bwd_apply_kernel(const bf16* ga, long long ga_ps, const bf16* y, long long y_ps, const bf16* x, long long x_ps,
                                       bf16* gx, long long gx_ps, int planes, long long rows, int c, int act, const float* mean,
                                       const float* invstd, const float* gamma, const double* sums) {
  const int P8 = c / 8;
  const long long pieces = rows * P8;
  const double inv_n = 1.0 / (double)rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int p8 = (int)(i % P8);
    const long long off = (i / P8) * c + p8 * 8;
    float g[8], yy[8], xx[8];
    load8(ga, ga_ps, planes, off, g);
    load8(y, y_ps, planes, off, yy);
    load8(x, x_ps, planes, off, xx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = p8 * 8 + j;
      const float gg = g[j] * act_grad_from_output(yy[j], act);
      const float xhat = (xx[j] - mean[ch]) * invstd[ch];
      g[j] = sums ? gamma[ch] * invstd[ch] * (gg - (float)(sums[ch] * inv_n) - xhat * (float)(sums[c + ch] * inv_n)) : gamma[ch] * invstd[ch] * gg;
    }
    store8(gx, gx_ps, planes, off, g);
  }
}

// dst[i] (+)= scale * (float)src[i]   (double sums -> fp32 parameter gradients)
__global__ void sg_emit_sums_kernel(const double* src, float* dst, int n, int accumulate, float scale, int wc, long long s_t, long long s_c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float* o = dst + (long long)(i / wc) * s_t + (long long)(i % wc) * s_c;
    *o = (accumulate ? *o : 0.f) + scale * (float)src[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ConvTranspose3d(C -> 1, k4, s2, p1) second stage: P[v][64] holds x[v,:] . W[:, tap] for the 64 taps (tensor-core
// GEMM, N = 64); every output voxel sums the 8 taps that reach it.   out = act(bias + sum)
__global__ void sg_col2im_c1_kernel(const bf16* P, long long p_ps, int planes, int n, int d, int h, int w, const float* bias,
                                    int act, float* out) {
  const int OD = 2 * d, OH = 2 * h, OW = 2 * w;
  const long long total = (long long)n * OD * OH * OW;
  const float b = bias ? __ldg(bias) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % OW); long long t = i / OW;
    const int oh = (int)(t % OH); t /= OH;
    const int od = (int)(t % OD); const long long nn = t / OD;
    float s = b;
#pragma unroll
    for (int td = 0; td < 2; ++td) {
      const int pd = od & 1, id = (od >> 1) + (pd ? 1 - td : -td), kd = pd ? 2 * td : 1 + 2 * td;
      if (id < 0 || id >= d) continue;
#pragma unroll
      for (int th = 0; th < 2; ++th) {
        const int ph = oh & 1, ih = (oh >> 1) + (ph ? 1 - th : -th), kh = ph ? 2 * th : 1 + 2 * th;
        if (ih < 0 || ih >= h) continue;
#pragma unroll
        for (int tw = 0; tw < 2; ++tw) {
          const int pw = ow & 1, iw = (ow >> 1) + (pw ? 1 - tw : -tw), kw = pw ? 2 * tw : 1 + 2 * tw;
          if (iw < 0 || iw >= w) continue;
          const long long v = ((nn * d + id) * h + ih) * w + iw;
          const long long off = v * 64 + kd * 16 + kh * 4 + kw;
          float x = __bfloat162float(P[off]);
          if (planes == 2) x += __bfloat162float(P[p_ps + off]);
          s += x;
        }
      }
    }
    out[i] = apply_act(s, act);
  }
}

// Tiled variant (bf16, output extents multiples of 16 x 8 x 8): a block stages the (8+2) x (4+2) x (4+2) input rows (64 taps = 128 B
// each) that reach its 16 x 8 x 8 output tile with coalesced 16-byte loads (zero rows outside the volume), then every output voxel
// sums its 8 taps out of shared memory (row pitch 33 words: the 8 taps of the 32 lanes fall into distinct banks).  The direct kernel
// above issues eight scattered 2-byte global loads per output -- 33 us for the 2 M voxels of a B = 64 batch, 4 launches per WGAN step.
constexpr int kC2iPitchW = 33;
__global__ void __launch_bounds__(256) sg_col2im_c1_tiled_kernel(const bf16* P, int n, int d, int h, int w, const float* bias, int act, float* out) {
  extern __shared__ uint32_t c2i_rows[];                 // [6 z][6 y][10 x][33 words]
  const int OD = 2 * d, OH = 2 * h, OW = 2 * w;
  const int tx = blockIdx.x, ty = blockIdx.y, tzn = blockIdx.z;
  const int tzs = OD / 8, nn = tzn / tzs, tz = tzn - nn * tzs;
  const int ix0 = tx * 8 - 1, iy0 = ty * 4 - 1, iz0 = tz * 4 - 1;
  const int t = threadIdx.x;
  for (int i = t; i < 360 * 8; i += 256) {
    const int r = i >> 3, ch = i & 7;
    const int rx = r % 10, ry = (r / 10) % 6, rz = r / 60;
    const int ix = ix0 + rx, iy = iy0 + ry, iz = iz0 + rz;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ix >= 0 && ix < w && iy >= 0 && iy < h && iz >= 0 && iz < d)
      v = __ldg(reinterpret_cast<const uint4*>(P + ((((long long)nn * d + iz) * h + iy) * w + ix) * 64) + ch);
    uint32_t* dst = c2i_rows + r * kC2iPitchW + ch * 4;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  __syncthreads();
  const float b = bias ? __ldg(bias) : 0.f;
  const int lx = t & 15, ly = (t >> 4) & 7, lz0 = t >> 7;
  const int ow = tx * 16 + lx, oh = ty * 8 + ly;
  const int pw = lx & 1, ph = ly & 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int lz = lz0 + 2 * k, od = tz * 8 + lz, pd = lz & 1;
    float s = b;
#pragma unroll
    for (int td = 0; td < 2; ++td) {
      const int rz = (lz >> 1) + (pd ? 1 - td : -td) + 1, kd = pd ? 2 * td : 1 + 2 * td;
#pragma unroll
      for (int th = 0; th < 2; ++th) {
        const int ry = (ly >> 1) + (ph ? 1 - th : -th) + 1, kh = ph ? 2 * th : 1 + 2 * th;
#pragma unroll
        for (int tw = 0; tw < 2; ++tw) {
          const int rx = (lx >> 1) + (pw ? 1 - tw : -tw) + 1, kw = pw ? 2 * tw : 1 + 2 * tw;
          const int tap = kd * 16 + kh * 4 + kw;
          const uint32_t wd = c2i_rows[((rz * 6 + ry) * 10 + rx) * kC2iPitchW + (tap >> 1)];
          s += (tap & 1) ? bf16hi_to_f(wd) : bf16lo_to_f(wd);       // rows outside the volume were staged as zeros
        }
      }
    }
    out[(((long long)nn * OD + od) * OH + oh) * OW + ow] = apply_act(s, act);
  }
}

__global__ void sg_unary_bwd_f32_kernel(const float* gy, const float* y, float* gx, long long n, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    gx[i] = gy[i] * act_grad_from_output(y[i], act);
}
__global__ void sg_unary_f32_kernel(const float* x, float* y, long long n, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = apply_act(x[i], act);
}

// ---------------------------------------------------------------------------------------------------------------
// y[r] = act(x[r,:] . w + b): one warp per row.   (Linear(256->1)+tanh sdf_net.py:50-51; Conv3d(256->1,k4,s1) gan.py:55;
// Linear(128->1) progressive_gan.py:30)
__global__ void sg_rowdot_fwd_kernel(const bf16* x, long long x_ps, int planes, long long rows, int c, const float* w,
                                     int wc, long long s_t, long long s_c, const float* bias, int act, float* y) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const float b = bias ? __ldg(bias) : 0.f;
  for (long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    float s = 0.f;
    for (int p8 = lane; p8 < c / 8; p8 += 32) {
      float v[8];
      load8(x, x_ps, planes, row * c + p8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int k = p8 * 8 + j; s += v[j] * __ldg(w + (long long)(k / wc) * s_t + (long long)(k % wc) * s_c); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = apply_act(s + b, act);
  }
}
// few, very wide rows (Conv3d(256->1,k4,s1): 64 rows x 16384): one block per row
__global__ void sg_rowdot_fwd_block_kernel(const bf16* x, long long x_ps, int planes, long long rows, int c, const float* w,
                                           int wc, long long s_t, long long s_c, const float* bias, int act, float* y) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  float s = 0.f;
  for (int p8 = threadIdx.x; p8 < c / 8; p8 += blockDim.x) {
    float v[8];
    load8(x, x_ps, planes, row * c + p8 * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int k = p8 * 8 + j; s += v[j] * __ldg(w + (long long)(k / wc) * s_t + (long long)(k % wc) * s_c); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) y[row] = apply_act(t + (bias ? __ldg(bias) : 0.f), act);
  }
}
// g[r] = gy[r]*act'(y[r]);  gx[r,:] = g[r]*w (planes);  gw[c] += sum_r g[r]*x[r,c];  gb += sum_r g[r]
__global__ void sg_rowdot_bwd_kernel(const float* gy, const float* y, int act, const bf16* x, long long x_ps, int planes,
                                     long long rows, int c, const float* w, int wc, long long s_t, long long s_c, bf16* gx,
                                     long long gx_ps, double* sums, int P8, int R, long long rows_per_block, int x_mask_act) {
  // blockIdx.y selects a slab of P8 piece columns (wide rows: C up to 16384)
  const int pl = threadIdx.x % P8, r = threadIdx.x / P8;
  const int p8 = blockIdx.y * P8 + pl;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  float wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int k = p8 * 8 + j; wv[j] = w[(long long)(k / wc) * s_t + (long long)(k % wc) * s_c]; }
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (long long row = r0 + r; row < r1; row += R) {
    const float g = gy[row] * act_grad_from_output(y[row], act);
    const long long off = row * c + p8 * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 1.f;
    if (x) {
      load8(x, x_ps, planes, off, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] += g * v[j];
    }
    if (p8 == 0) acc[1][0] += g;
    if (gx) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = g * wv[j] * (x_mask_act != ACT_NONE ? act_grad_from_output(v[j], x_mask_act) : 1.f);
      store8(gx, gx_ps, planes, off, o);
    }
  }
  if (sums) {   // sums[0..c) = gw (this slab's columns), sums[c] = gb
    extern __shared__ float red[];
    const int CS = P8 * 8;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(a * R + r) * CS + pl * 8 + j] = acc[a][j];
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CS; i += blockDim.x) {
      const int a = i / CS, cc = i - a * CS;
      if (a == 1 && (cc != 0 || blockIdx.y != 0)) continue;
      float s = 0.f;
      for (int rr = 0; rr < R; ++rr) s += red[(a * R + rr) * CS + cc];
      atomicAdd(sums + (a == 0 ? (size_t)blockIdx.y * CS + cc : (size_t)c), (double)s);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 [rows, src_ld] (first c_src columns) -> planes [rows, c_dst] (zero padded)
__global__ void sg_to_planes_kernel(const float* src, long long src_ld, long long rows, int c_src, bf16* dst, long long dst_ps,
                                    int planes, int c_dst) {
  const int P8 = c_dst / 8;
  const long long pieces = rows * P8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int p8 = (int)(i % P8);
    const long long row = i / P8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = p8 * 8 + j;
      v[j] = ch < c_src ? src[row * src_ld + ch] : 0.f;
    }
    store8(dst, dst_ps, planes, row * c_dst + p8 * 8, v);
  }
}
// planes [rows, c_src] -> fp32 [rows, dst_ld] (first c_take columns), dst (+)= scale*src
__global__ void sg_from_planes_kernel(const bf16* src, long long src_ps, int planes, long long rows, int c_src, int c_take,
                                      float* dst, long long dst_ld, int accumulate, float scale) {
  const int P8 = (c_take + 7) / 8;
  const long long pieces = rows * P8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int p8 = (int)(i % P8);
    const long long row = i / P8;
    float v[8];
    load8(src, src_ps, planes, row * c_src + p8 * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = p8 * 8 + j;
      if (ch < c_take) {
        float* o = dst + row * dst_ld + ch;
        *o = (accumulate ? *o : 0.f) + scale * v[j];
      }
    }
  }
}

// SDFNet input rows: [x, y, z, latent(L), 0...] (sdf_net.py:57 cat(points, latent_codes)); latent row = index ? table[index[i]] : latent[i]
__global__ void sg_sdf_pack_input_kernel(const float* points, const float* latent, const int* index, int L, long long n,
                                         bf16* dst, long long dst_ps, int planes, int c_dst) {
  const int P8 = c_dst / 8;
  const long long pieces = n * P8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int p8 = (int)(i % P8);
    const long long row = i / P8;
    const long long lrow = index ? (long long)index[row] : row;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = p8 * 8 + j;
      float x = 0.f;
      if (ch < 3) x = points[row * 3 + ch];
      else if (ch < 3 + L) x = latent[lrow * L + (ch - 3)];
      v[j] = x;
    }
    store8(dst, dst_ps, planes, row * c_dst + p8 * 8, v);
  }
}
// gradient wrt the input rows = ga (+ gb): scatter to gpoints [n,3] and glatent ([n,L] rows, or atomicAdd into table rows)
__global__ void sg_sdf_unpack_grad_kernel(const bf16* ga, long long ga_ps, const bf16* gb, long long gb_ps, int planes, long long n,
                                          int c_src, int L, const int* index, float* gpoints, float* glatent) {
  const int P8 = (3 + L + 7) / 8;
  const long long pieces = n * P8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int p8 = (int)(i % P8);
    const long long row = i / P8;
    float v[8];
    load8(ga, ga_ps, planes, row * c_src + p8 * 8, v);
    if (gb) {
      float u[8];
      load8(gb, gb_ps, planes, row * c_src + p8 * 8, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += u[j];
    }
    const long long lrow = index ? (long long)index[row] : row;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = p8 * 8 + j;
      if (ch < 3) { if (gpoints) gpoints[row * 3 + ch] = v[j]; }
      else if (ch < 3 + L && glatent) {
        if (index) atomicAdd(glatent + lrow * L + (ch - 3), v[j]);
        else glatent[lrow * L + (ch - 3)] = v[j];
      }
    }
  }
}
// indexed variant: each thread owns one piece column and walks a contiguous run of rows, accumulating in registers while
// the shape index stays the same (points of a shape are contiguous in every caller of the reference) and flushing one
// atomicAdd per channel per run -- instead of one atomic per point and channel.
__global__ void sg_sdf_unpack_grad_runs_kernel(const bf16* ga, long long ga_ps, const bf16* gb, long long gb_ps, int planes, long long n,
                                               int c_src, int L, const int* index, float* gpoints, float* glatent, int rows_per_thread) {
  const int P8 = (3 + L + 7) / 8;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int p8 = (int)(t % P8);
  const long long r0 = (t / P8) * rows_per_thread, r1 = min(n, r0 + rows_per_thread);
  if (r0 >= n) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  int cur = index[r0];
  for (long long row = r0; row < r1; ++row) {
    const int s = index[row];
    if (s != cur) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int ch = p8 * 8 + j; if (ch >= 3 && ch < 3 + L) atomicAdd(glatent + (long long)cur * L + (ch - 3), acc[j]); acc[j] = 0.f; }
      cur = s;
    }
    float v[8];
    load8(ga, ga_ps, planes, row * c_src + p8 * 8, v);
    if (gb) {
      float u[8];
      load8(gb, gb_ps, planes, row * c_src + p8 * 8, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += u[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = p8 * 8 + j;
      if (ch < 3) { if (gpoints) gpoints[row * 3 + ch] = v[j]; }
      else acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int ch = p8 * 8 + j; if (ch >= 3 && ch < 3 + L) atomicAdd(glatent + (long long)cur * L + (ch - 3), acc[j]); }
}

// progressive_gan.py:48-50   x = f*x + (1-f)*from_SDF(x_in[:, ::2, ::2, ::2])   (x: [B,r,r,r,C] planes, vol: [B,2r,2r,2r] fp32)
__global__ void sg_fade_fwd_kernel(const bf16* x, long long x_ps, bf16* y, long long y_ps, int planes, int b, int r, int c,
                                   const float* vol, float f) {
  const int P8 = c / 8;
  const long long pieces = (long long)b * r * r * r * P8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int p8 = (int)(i % P8);
    const long long v = i / P8;
    float a[8];
    load8(x, x_ps, planes, v * c + p8 * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= f;
    if (p8 == 0) {
      const int w = (int)(v % r); long long t = v / r;
      const int h = (int)(t % r); t /= r;
      const int d = (int)(t % r); const long long n = t / r;
      const int R2 = 2 * r;
      a[0] += (1.f - f) * vol[((n * R2 + 2 * d) * R2 + 2 * h) * R2 + 2 * w];
    }
    store8(y, y_ps, planes, v * c + p8 * 8, a);
  }
}
// gvol[b, 2d, 2h, 2w] += (1-f) * g[b,d,h,w,0]    (gx = f*g is a plain scale done by sg_scale_planes)
__global__ void sg_fade_bwd_vol_kernel(const bf16* g, long long g_ps, int planes, int b, int r, int c, float f, float* gvol) {
  const long long total = (long long)b * r * r * r;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    float x = __bfloat162float(g[v * c]);
    if (planes == 2) x += __bfloat162float(g[g_ps + v * c]);
    const int w = (int)(v % r); long long t = v / r;
    const int h = (int)(t % r); t /= r;
    const int d = (int)(t % r); const long long n = t / r;
    const int R2 = 2 * r;
    gvol[((n * R2 + 2 * d) * R2 + 2 * h) * R2 + 2 * w] += (1.f - f) * x;
  }
}
// y = alpha*a (+ beta*b)
__global__ void sg_axpby_planes_kernel(const bf16* a, long long a_ps, const bf16* b, long long b_ps, bf16* y, long long y_ps,
                                       int planes, long long pieces, float alpha, float beta) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    float u[8];
    load8(a, a_ps, planes, i * 8, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] *= alpha;
    if (b) {
      float v[8];
      load8(b, b_ps, planes, i * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] += beta * v[j];
    }
    store8(y, y_ps, planes, i * 8, u);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused optimizer updates over a flat fp32 arena (torch.optim.RMSprop / Adam defaults; train_wgan.py:45-46,
// train_gan.py:28-31, train_sdf_autodecoder.py:44-45).  grad_scale folds the 1/world_size of the gradient all-reduce;
// clip > 0 folds Discriminator.clip_weights (gan.py:67-69).
__global__ void sg_rmsprop_kernel(float* p, const float* g, float* sq, long long n, float lr, float alpha, float eps,
                                  float grad_scale, float clip) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * grad_scale;
    const float s = alpha * sq[i] + (1.f - alpha) * gr * gr;
    sq[i] = s;
    float v = p[i] - lr * gr / (sqrtf(s) + eps);
    if (clip > 0.f) v = fminf(fmaxf(v, -clip), clip);
    p[i] = v;
  }
}
__global__ void sg_adam_kernel(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                               float bc1, float bc2_sqrt, float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * grad_scale;
    const float mm = b1 * m[i] + (1.f - b1) * gr;
    const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mm; v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mm / denom);
  }
}
__global__ void sg_clamp_kernel(float* p, long long n, float lo, float hi) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = fminf(fmaxf(p[i], lo), hi);
}

// L1 data term of the auto-decoder loss (train_sdf_autodecoder.py:88): loss += sum|out - target| / n ; gout = sign(out-target)/n
__global__ void sg_l1_loss_grad_kernel(const float* out, const float* target, float* gout, long long n, double* loss_sum) {
  float local = 0.f;
  const float inv = 1.f / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = out[i] - target[i];
    local += fabsf(d);
    if (gout) gout[i] = (d > 0.f ? inv : (d < 0.f ? -inv : 0.f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && loss_sum) atomicAdd(loss_sum, (double)local * inv);
}

__global__ void sg_sum_f32_kernel(const float* x, long long n, double* out) {
  float local = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) local += x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, (double)local);
}

static inline int ew_grid(long long items, int block) {
  return (int)std::max<long long>(1, std::min<long long>((items + block - 1) / block, 148 * 16));
}

struct ColCfg { int P8, R, block; long long rpb; int grid; size_t smem; };
static int col_cfg(long long rows, int c, ColCfg& k, const char* who) {
  if (c <= 0 || (c & 7) || c > 2048) { return sg_fail(-30, who); }
  k.P8 = c / 8;
  k.R = std::max(1, 256 / k.P8);
  k.block = k.P8 * k.R;
  long long per = (long long)k.R * 4;      // >= 4 rows per thread; small tensors still spread over many blocks
  long long blocks = std::max<long long>(1, std::min<long long>((rows + per - 1) / per, 148 * 8));
  k.rpb = (rows + blocks - 1) / blocks;
  k.grid = (int)((rows + k.rpb - 1) / std::max<long long>(k.rpb, 1));
  if (k.grid < 1) k.grid = 1;
  k.smem = (size_t)2 * k.R * c * sizeof(float);
  return 0;
}

}  // namespace sg

using namespace sg;
#define ST(s) ((cudaStream_t)(s))

extern "C" int sg_act_bwd(const void* ga, int64_t ga_ps, const void* y, int64_t y_ps, void* g, int64_t g_ps, int planes,
                          int64_t rows, int c, int act, double* sums, void* stream) {
  if (rows <= 0) return 0;
  ColCfg k;
  int rc = col_cfg(rows, c, k, "sg_act_bwd: C must be a multiple of 8, <= 2048");
  if (rc) return rc;
  ColArgs a = {};
  a.ga = (const bf16*)ga; a.ga_ps = ga_ps; a.y = (const bf16*)y; a.y_ps = y_ps; a.g = (bf16*)g; a.g_ps = g_ps;
  a.planes = planes; a.rows = rows; a.c = c; a.act = act; a.sums = sums;
  sg_colreduce_kernel<1><<<k.grid, k.block, k.smem, ST(stream)>>>(a, k.P8, k.R, k.rpb);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_bn_stats(const void* x, int64_t x_ps, int planes, int64_t rows, int c, double* sums, void* stream) {
  if (rows <= 0) return 0;
  ColCfg k;
  int rc = col_cfg(rows, c, k, "sg_bn_stats: C must be a multiple of 8, <= 2048");
  if (rc) return rc;
  ColArgs a = {};
  a.x = (const bf16*)x; a.x_ps = x_ps; a.planes = planes; a.rows = rows; a.c = c; a.sums = sums;
  sg_colreduce_kernel<0><<<k.grid, k.block, k.smem, ST(stream)>>>(a, k.P8, k.R, k.rpb);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_bn_finalize(const double* sums, int64_t rows, int c, float eps, float momentum, float* mean, float* invstd,
                              float* running_mean, float* running_var, void* stream) {
  sg_bn_finalize_kernel<<<(c + 127) / 128, 128, 0, ST(stream)>>>(sums, rows, c, eps, momentum, mean, invstd, running_mean, running_var);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_bn_apply(const void* x, int64_t x_ps, void* y, int64_t y_ps, int planes, int64_t rows, int c, const float* mean,
                           const float* invstd, const float* gamma, const float* beta, int act, void* stream) {
  if (rows <= 0) return 0;
  if (c & 7) return sg_fail(-31, "sg_bn_apply: C % 8");
  sg_bn_apply_kernel<<<ew_grid(rows * (c / 8), 256), 256, 0, ST(stream)>>>((const bf16*)x, x_ps, (bf16*)y, y_ps, planes, rows, c, mean,
                                                                          invstd, gamma, beta, act);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_bn_bwd_reduce(const void* ga, int64_t ga_ps, const void* y, int64_t y_ps, const void* x, int64_t x_ps, int planes,
                                int64_t rows, int c, int act, const float* mean, const float* invstd, double* sums, void* stream) {
  if (rows <= 0) return 0;
  ColCfg k;
  int rc = col_cfg(rows, c, k, "sg_bn_bwd_reduce: C must be a multiple of 8, <= 2048");
  if (rc) return rc;
  ColArgs a = {};
  a.ga = (const bf16*)ga; a.ga_ps = ga_ps; a.y = (const bf16*)y; a.y_ps = y_ps; a.x = (const bf16*)x; a.x_ps = x_ps;
  a.planes = planes; a.rows = rows; a.c = c; a.act = act; a.mean = mean; a.invstd = invstd; a.sums = sums;
  sg_colreduce_kernel<2><<<k.grid, k.block, k.smem, ST(stream)>>>(a, k.P8, k.R, k.rpb);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_bn_bwd_apply(const void* ga, int64_t ga_ps, const void* y, int64_t y_ps, const void* x, int64_t x_ps, void* gx,
                               int64_t gx_ps, int planes, int64_t rows, int c, int act, const float* mean, const float* invstd,
                               const float* gamma, const double* sums, void* stream) {
  if (rows <= 0) return 0;
  sg_bn_bwd_apply_kernel<<<ew_grid(rows * (c / 8), 256), 256, 0, ST(stream)>>>((const bf16*)ga, ga_ps, (const bf16*)y, y_ps,
                                                                              (const bf16*)x, x_ps, (bf16*)gx, gx_ps, planes, rows, c,
                                                                              act, mean, invstd, gamma, sums);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_emit_sums(const double* src, float* dst, int n, int accumulate, float scale, int wc, int64_t s_t, int64_t s_c, void* stream) {
  if (n <= 0) return 0;
  if (wc <= 0) { wc = n; s_t = 0; s_c = 1; }
  sg_emit_sums_kernel<<<(n + 127) / 128, 128, 0, ST(stream)>>>(src, dst, n, accumulate, scale, wc, s_t, s_c);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_col2im_c1(const void* P, int64_t p_ps, int planes, int n, int d, int h, int w, const float* bias, int act,
                            float* out, void* stream) {
  const long long total = (long long)n * d * h * w * 8;
  if (total <= 0) return 0;
  {
    const char* nt = getenv("SG_B200_NO_TILED_COL2IM");
    if (planes == 1 && (w % 8) == 0 && (h % 4) == 0 && (d % 4) == 0 && ((uintptr_t)P & 15) == 0 && (long long)(d / 4) * n <= 65535 &&
        !(nt && nt[0] == '1')) {
      const size_t smem = 360 * kC2iPitchW * sizeof(uint32_t);
      static PerDevice attr;
      if (attr.first()) { cudaFuncSetAttribute(sg_col2im_c1_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr.done(); }
      dim3 grid((unsigned)(w / 8), (unsigned)(h / 4), (unsigned)((d / 4) * n));
      sg_col2im_c1_tiled_kernel<<<grid, 256, smem, ST(stream)>>>((const bf16*)P, n, d, h, w, bias, act, out);
      SG_CUDA_CHECK_LAUNCH();
      return 0;
    }
  }
  sg_col2im_c1_kernel<<<ew_grid(total, 256), 256, 0, ST(stream)>>>((const bf16*)P, p_ps, planes, n, d, h, w, bias, act, out);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_unary_f32(const float* x, float* y, int64_t n, int act, void* stream) {
  if (n <= 0) return 0;
  sg_unary_f32_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(x, y, n, act);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_unary_bwd_f32(const float* gy, const float* y, float* gx, int64_t n, int act, void* stream) {
  if (n <= 0) return 0;
  sg_unary_bwd_f32_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(gy, y, gx, n, act);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_rowdot_fwd(const void* x, int64_t x_ps, int planes, int64_t rows, int c, const float* w, int wc, int64_t s_t,
                             int64_t s_c, const float* bias, int act, float* y, void* stream) {
  if (rows <= 0) return 0;
  if (c & 7) return sg_fail(-32, "sg_rowdot_fwd: C % 8");
  if (wc <= 0) { wc = c; s_t = 0; s_c = 1; }
  if (rows <= 2048 && c >= 2048) {
    sg_rowdot_fwd_block_kernel<<<(int)rows, 256, 0, ST(stream)>>>((const bf16*)x, x_ps, planes, rows, c, w, wc, s_t, s_c, bias, act, y);
    SG_CUDA_CHECK_LAUNCH();
    return 0;
  }
  sg_rowdot_fwd_kernel<<<ew_grid(rows * 32, 256), 256, 0, ST(stream)>>>((const bf16*)x, x_ps, planes, rows, c, w, wc, s_t, s_c, bias, act, y);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_rowdot_bwd(const float* gy, const float* y, int act, const void* x, int64_t x_ps, int planes, int64_t rows, int c,
                             const float* w, int wc, int64_t s_t, int64_t s_c, void* gx, int64_t gx_ps, double* sums, int x_mask_act,
                             void* stream) {
  if (rows <= 0) return 0;
  if (wc <= 0) { wc = c; s_t = 0; s_c = 1; }
  // wide rows are processed in slabs of <= 2048 columns (grid.y); c must divide evenly into slabs
  int slab = c, slabs = 1;
  while (slab > 2048) { if (slab & 1) return sg_fail(-37, "sg_rowdot_bwd: C"); slab >>= 1; slabs <<= 1; }
  ColCfg k;
  int rc = col_cfg(rows, slab, k, "sg_rowdot_bwd: C must be a multiple of 8");
  if (rc) return rc;
  dim3 grid(k.grid, slabs);
  sg_rowdot_bwd_kernel<<<grid, k.block, k.smem, ST(stream)>>>(gy, y, act, (const bf16*)x, x_ps, planes, rows, c, w, wc, s_t, s_c,
                                                            (bf16*)gx, gx_ps, sums, k.P8, k.R, k.rpb, x_mask_act);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_to_planes(const float* src, int64_t src_ld, int64_t rows, int c_src, void* dst, int64_t dst_ps, int planes, int c_dst,
                            void* stream) {
  if (rows <= 0) return 0;
  if (c_dst & 7) return sg_fail(-33, "sg_to_planes: c_dst % 8");
  sg_to_planes_kernel<<<ew_grid(rows * (c_dst / 8), 256), 256, 0, ST(stream)>>>(src, src_ld, rows, c_src, (bf16*)dst, dst_ps, planes, c_dst);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_from_planes(const void* src, int64_t src_ps, int planes, int64_t rows, int c_src, int c_take, float* dst,
                              int64_t dst_ld, int accumulate, float scale, void* stream) {
  if (rows <= 0) return 0;
  if (c_src & 7) return sg_fail(-34, "sg_from_planes: c_src % 8");
  sg_from_planes_kernel<<<ew_grid(rows * ((c_take + 7) / 8), 256), 256, 0, ST(stream)>>>((const bf16*)src, src_ps, planes, rows, c_src,
                                                                                        c_take, dst, dst_ld, accumulate, scale);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_sdf_pack_input(const float* points, const float* latent, const int32_t* index, int L, int64_t n, void* dst,
                                 int64_t dst_ps, int planes, int c_dst, void* stream) {
  if (n <= 0) return 0;
  if ((c_dst & 7) || c_dst < 3 + L) return sg_fail(-35, "sg_sdf_pack_input: c_dst");
  sg_sdf_pack_input_kernel<<<ew_grid(n * (c_dst / 8), 256), 256, 0, ST(stream)>>>(points, latent, index, L, n, (bf16*)dst, dst_ps, planes, c_dst);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_sdf_unpack_grad(const void* ga, int64_t ga_ps, const void* gb, int64_t gb_ps, int planes, int64_t n, int c_src, int L,
                                  const int32_t* index, float* gpoints, float* glatent, void* stream) {
  if (n <= 0) return 0;
  if (index && glatent) {
    const int P8 = (3 + L + 7) / 8, rpt = 64;
    const long long threads = ((n + rpt - 1) / rpt) * P8;
    sg_sdf_unpack_grad_runs_kernel<<<(int)((threads + 255) / 256), 256, 0, ST(stream)>>>((const bf16*)ga, ga_ps, (const bf16*)gb, gb_ps, planes, n,
                                                                                      c_src, L, index, gpoints, glatent, rpt);
    SG_CUDA_CHECK_LAUNCH();
    return 0;
  }
  sg_sdf_unpack_grad_kernel<<<ew_grid(n * ((3 + L + 7) / 8), 256), 256, 0, ST(stream)>>>((const bf16*)ga, ga_ps, (const bf16*)gb, gb_ps,
                                                                                       planes, n, c_src, L, index, gpoints, glatent);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_fade_fwd(const void* x, int64_t x_ps, void* y, int64_t y_ps, int planes, int b, int r, int c, const float* vol, float f,
                           void* stream) {
  const long long pieces = (long long)b * r * r * r * (c / 8);
  if (pieces <= 0) return 0;
  sg_fade_fwd_kernel<<<ew_grid(pieces, 256), 256, 0, ST(stream)>>>((const bf16*)x, x_ps, (bf16*)y, y_ps, planes, b, r, c, vol, f);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_fade_bwd_vol(const void* g, int64_t g_ps, int planes, int b, int r, int c, float f, float* gvol, void* stream) {
  const long long total = (long long)b * r * r * r;
  if (total <= 0) return 0;
  sg_fade_bwd_vol_kernel<<<ew_grid(total, 256), 256, 0, ST(stream)>>>((const bf16*)g, g_ps, planes, b, r, c, f, gvol);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_axpby_planes(const void* a, int64_t a_ps, const void* b, int64_t b_ps, void* y, int64_t y_ps, int planes,
                               int64_t elems, float alpha, float beta, void* stream) {
  if (elems <= 0) return 0;
  if (elems & 7) return sg_fail(-36, "sg_axpby_planes: elems % 8");
  sg_axpby_planes_kernel<<<ew_grid(elems / 8, 256), 256, 0, ST(stream)>>>((const bf16*)a, a_ps, (const bf16*)b, b_ps, (bf16*)y, y_ps,
                                                                         planes, elems / 8, alpha, beta);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_rmsprop(float* p, const float* g, float* sq, int64_t n, float lr, float alpha, float eps, float grad_scale,
                          float clip, void* stream) {
  if (n <= 0) return 0;
  sg_rmsprop_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(p, g, sq, n, lr, alpha, eps, grad_scale, clip);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, int step,
                       float grad_scale, void* stream) {
  if (n <= 0) return 0;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  sg_adam_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, sqrtf(bc2), grad_scale);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_clamp(float* p, int64_t n, float lo, float hi, void* stream) {
  if (n <= 0) return 0;
  sg_clamp_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(p, n, lo, hi);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_sum_f32(const float* x, int64_t n, double* out, void* stream) {
  if (n <= 0) return 0;
  sg_sum_f32_kernel<<<(int)std::min<long long>((n + 255) / 256, 592), 256, 0, ST(stream)>>>(x, n, out);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
extern "C" int sg_l1_loss_grad(const float* out, const float* target, float* gout, int64_t n, double* loss_sum, void* stream) {
  if (n <= 0) return 0;
  sg_l1_loss_grad_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(out, target, gout, n, loss_sum);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
