// Fused persistent DeepSDF MLP (model/sdf_net.py:56-61), bf16 operands / fp32 accumulation, activations never leave the SM.
//
// FORWARD  sg_sdfnet_fwd_kernel: all 8 layers of a pair of 128-point tiles inside one CTA.
//   TMEM  : two 128 x 256 fp32 accumulators (tile A: columns 0-255, tile B: 256-511).  Before the MMAs of a layer run, the
//           epilogue warps PRE-LOAD the accumulator with that layer's fp32 bias (+ the K = 3 xyz columns of layers 1 and 5)
//           through tcgen05.st, so every MMA accumulates and the epilogue is only  tcgen05.ld -> cvt.rn.relu.bf16x2 -> st.shared.
//   SMEM  : hA, hB   64 KB each  hidden activations of the two tiles, K-major 128B-swizzled (4 chunks of [128][64] bf16),
//                                 overwritten in place by the epilogue of every layer (it runs after the layer's MMAs retire)
//           lat      32 KB       latent operand of ONE tile (2 chunks), gathered by index from the fp32 table
//           w[2]     64 KB       two stages of streamed weights, one chunk = [256 n][64 k] bf16 = 32 KB, 30 chunks / pair:
//                                 L1 latent(2) L2(4) L3(4) L4(4) L5 latent(2, tile A) L5 hidden(4) L5 latent(2 again, tile B)
//                                 L6(4) L7(4); the hidden chunks feed BOTH tiles
//   CONST : biases, xyz columns, the 256 -> 1 head (14 KB, uploaded per launch): uniform reads that do not touch the
//           2 KB of L1 this kernel's shared-memory carve-out leaves
//   warps : 0 weight loader (1-D TMA bulk copies) | 1 MMA issuer (tcgen05.mma M128 N256 K16) | 2-5 latent gather |
//           6-13 epilogue of tile A | 14-21 epilogue of tile B; an epilogue warp owns TMEM lane quadrant warp%4 (32 rows) and one
//           column half (128 columns): four epilogue warps per scheduler hide each other's TMEM / constant-bank latencies
//   Layer 8 (256 -> 1) + tanh is a per-row dot product folded into the epilogue of layer 7.
//
// BACKWARD sg_sdfnet_bwd_kernel: the input-gradient chain g7 -> g6 -> ... -> g1 of a tile pair inside one CTA
//   (g_l = gradient w.r.t. the pre-activation of layer l).  Head: g7 = gout (1 - out^2) w8 [h7 > 0]; layers 7..2:
//   g_{l-1} = (g_l W_l[:, :256]) [h_{l-1} > 0] with g_l as the SMEM A operand, W_l^T streamed (24 chunks / pair, 3 stages),
//   the ReLU mask read as 1 bit per element from the forward's mask stash (32 B per row and layer), one step ahead of its use.
//   Every g_l is written to `gstash`
//   (bf16 [7][n][256]) for the weight-gradient GEMMs, the bias sums and the two input-gradient GEMMs that follow.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_tma.h"

namespace sg {

constexpr int kSdfThreads = 704;
constexpr uint32_t kOffDot = 1024;           // float[2][128] inside the header block: column-half partial sums of the 256 -> 1 head
constexpr int kSdfImgChunks = 28;       // chunks in the weight image
constexpr int kSdfStreamChunks = 30;    // chunks streamed per tile pair (the 2 latent chunks of layers2.0 twice)
constexpr uint32_t kChunkBytes = 32768;
constexpr uint32_t kSdfHdr = 2048;
constexpr uint32_t kOffHA = kSdfHdr;
constexpr uint32_t kOffHB = kOffHA + 65536;
constexpr uint32_t kOffLat = kOffHB + 65536;
constexpr uint32_t kOffW = kOffLat + 32768;
constexpr uint32_t kSdfSmem = kOffW + 2 * kChunkBytes;      // 231424 B

struct SdfHdr {
  uint64_t w_full[3], w_empty[3];
  uint64_t lat_full, lat_empty;
  uint64_t acc_full[2], h_ready[2];
  uint32_t tmem_base;
};

struct SdfP {
  const float* points; const float* latent; const int* index; long long n;
  const char* w_img; float* out; bf16* stash; uint32_t* mstash;
  long long pairs;
  int* err;
  int tma_stash;            // stash rows of layers 1..6 leave through TMA tile stores out of the activation tiles in SMEM
  CUtensorMap tm_stash;     // bf16 [7][n][256], box [1][128][64], 128B swizzle
  // ---- single-latent inference (kernel instantiation FOLDED = true): the latent part of layers1.0 / layers2.0 is the constant
  // vector W[:, latent] z, folded into the bias columns of the aux block by the host, so those layers stream no latent chunks
  const int* ray_index;     // compact row -> slot of the point / output arrays (NULL: identity)
  const int* n_ptr;         // device-resident row count (NULL: n) -- sphere tracing compacts its ray list on the device
  int grid_r;               // > 0: slot s is grid cell (s / r^2, s / r % r, s % r); coordinates come from grid_axis[3][r]
  const float* grid_axis;
  // sphere-tracing step (rendering/raymarching.py:106-120 and :48-61): points advance in place, survivors are re-listed
  float* trace_points; const float* trace_dirs; unsigned char* trace_hit; int* next_index; int* next_count;
  float sdf_offset, trace_clamp, trace_threshold, trace_radius; int trace_miss_y;
};

// aux layout (floats)
constexpr int kAuxXb1 = 0;          // layers1.0, per column PAIR (c, c+1): {wx_c, wx_c1, wy_c, wy_c1, wz_c, wz_c1, b_c, b_c1} (fma.rn.f32x2 operands)
constexpr int kAuxXb5 = 1024;       // same for layers2.0
constexpr int kAuxBias = 2048;      // [5][256]: layers1.2, 1.4, 1.6, layers2.2, 2.4
constexpr int kAuxW8 = 2048 + 5 * 256;
constexpr int kAuxB8 = kAuxW8 + 256;
constexpr int kAuxFloats = kAuxB8 + 1;

__constant__ float c_sdf_aux[kAuxFloats];     // forward: see above; uploaded (device -> device) by sg_sdfnet_fwd on its stream
__constant__ float c_sdf_w8[256];             // backward: layers2.6 weight row

__device__ __forceinline__ unsigned long long f32x2_bcast(float a) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(a));
  return d;
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// fp32 value the accumulator of layer `l` starts from, columns [col, col+32) of the row at (px, py, pz)
__device__ __forceinline__ void sdf_acc_init(int l, int col, float px, float py, float pz, uint32_t (&v)[32]) {
  if (l == 1 || l == 5) {
    const unsigned long long* xb = reinterpret_cast<const unsigned long long*>(c_sdf_aux + (l == 1 ? kAuxXb1 : kAuxXb5)) + (col >> 1) * 4;
    const unsigned long long x2 = f32x2_bcast(px), y2 = f32x2_bcast(py), z2 = f32x2_bcast(pz);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      unsigned long long d = fma_f32x2(xb[4 * j + 2], z2, xb[4 * j + 3]);
      d = fma_f32x2(xb[4 * j + 1], y2, d);
      d = fma_f32x2(xb[4 * j], x2, d);
      v[2 * j] = (uint32_t)d; v[2 * j + 1] = (uint32_t)(d >> 32);
    }
  } else {
    const int bslot = (l < 5) ? l - 2 : l - 3;            // layers 2,3,4 -> 0,1,2 ; layers 6,7 -> 3,4
    const float* b = c_sdf_aux + kAuxBias + bslot * 256 + col;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(b[j]);
  }
}

// where the point of compact row `gr` lives / where its result goes
template <bool FOLDED>
__device__ __forceinline__ long long sdf_slot(const SdfP& p, long long gr) {
  if (FOLDED && p.ray_index != nullptr) return (long long)__ldg(p.ray_index + gr);
  return gr;
}
template <bool FOLDED>
__device__ __forceinline__ void sdf_load_point(const SdfP& p, long long slot, float& x, float& y, float& z) {
  if (FOLDED && p.grid_r > 0) {
    const uint32_t s = (uint32_t)slot, r = (uint32_t)p.grid_r;
    const uint32_t iz = s % r, t = s / r, iy = t % r, ix = t / r;
    x = __ldg(p.grid_axis + ix); y = __ldg(p.grid_axis + r + iy); z = __ldg(p.grid_axis + 2 * r + iz);
    return;
  }
  if (FOLDED && p.trace_points != nullptr) {      // advanced in place by this very kernel (other rays): plain loads, not the read-only path
    const float* src = p.trace_points + slot * 3;
    x = src[0]; y = src[1]; z = src[2];
    return;
  }
  x = __ldg(p.points + slot * 3); y = __ldg(p.points + slot * 3 + 1); z = __ldg(p.points + slot * 3 + 2);
}

template <bool FOLDED>
__global__ void __launch_bounds__(kSdfThreads, 1) sg_sdfnet_fwd_kernel(const __grid_constant__ SdfP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SdfHdr* hdr = reinterpret_cast<SdfHdr*>(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const long long n_rows = (FOLDED && p.n_ptr != nullptr) ? (long long)__ldg(p.n_ptr) : p.n;
  const long long n_pairs = (FOLDED && p.n_ptr != nullptr) ? ((n_rows + kTileRows - 1) / kTileRows + 1) / 2 : p.pairs;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // broadcast: the compiler can keep everything derived from it uniform
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->w_full[i], 1); mbar_init(&hdr->w_empty[i], 1);
      mbar_init(&hdr->acc_full[i], 1); mbar_init(&hdr->h_ready[i], 256);
    }
    mbar_init(&hdr->lat_full, 128); mbar_init(&hdr->lat_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0) atomicExch(p.err, kErrSmemAlign);
    __trap();
  }
  const uint32_t s_base = smem_u32(smem);

  if (warp == 0) {
    // ================================================================ weight loader
    uint32_t g = 0;
    for (long long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
      for (int j = 0; j < (FOLDED ? 24 : kSdfStreamChunks); ++j, ++g) {
        // stream position -> image chunk: [14,16) and [20,22) are the latent chunks 18,19 of layers2.0, [16,20) its hidden chunks 14-17
        // folded latent: only the hidden chunks 2..17 and 20..27
        const int ic = FOLDED ? (j < 16 ? j + 2 : j + 4) : (j < 14 ? j : (j < 16 ? j + 4 : j - 2));
        const uint32_t st = g & 1u;
        mbar_wait(&hdr->w_empty[st], ((g >> 1) & 1u) ^ 1u, p.err);
        if (elect_one()) {
          mbar_arrive_expect_tx(&hdr->w_full[st], kChunkBytes);
          bulk_g2s(s_base + kOffW + st * kChunkBytes, p.w_img + (size_t)ic * kChunkBytes, kChunkBytes, &hdr->w_full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (every MMA accumulates: the epilogue warps
    // pre-load the accumulators with the bias terms before they arrive on h_ready)
    const uint32_t idesc = umma_idesc(128, 256, false, false);
    uint32_t g = 0, lat_n = 0, hr_n[2] = {0, 0};
    // latent part of a layer for tile t: 2 resident weight chunks (stages g, g+1) x the shared latent tile
    auto latent_mma = [&](int t, bool commit_acc) {
      mbar_wait(&hdr->lat_full, lat_n & 1u, p.err); ++lat_n;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)t * 256u;
        for (int c = 0; c < 2; ++c) {
          const uint32_t a = s_base + kOffLat + c * kTileBytes, b = s_base + kOffW + ((g + c) & 1u) * kChunkBytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(d, umma_desc(a + kk * 32, 16, 1024), umma_desc(b + kk * 32, 16, 1024), idesc, 1u);
        }
        umma_commit(&hdr->lat_empty);
        if (commit_acc) umma_commit(&hdr->acc_full[t]);
      }
      __syncwarp();
    };
    auto wait_two_chunks = [&]() {
      mbar_wait(&hdr->w_full[g & 1u], (g >> 1) & 1u, p.err);
      mbar_wait(&hdr->w_full[(g + 1) & 1u], ((g + 1) >> 1) & 1u, p.err);
    };
    auto release_two_chunks = [&]() {
      if (elect_one()) { umma_commit(&hdr->w_empty[g & 1u]); umma_commit(&hdr->w_empty[(g + 1) & 1u]); }
      __syncwarp();
      g += 2;
    };
    for (long long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
      for (int l = 1; l <= 7; ++l) {
        // h[t] holds the activations of layer l-1 and accumulator t is drained and re-initialised
        for (int t = 0; t < 2; ++t) { mbar_wait(&hdr->h_ready[t], hr_n[t] & 1u, p.err); ++hr_n[t]; }
        tc_fence_after();
        if (p.tma_stash && l >= 2 && elect_one()) {      // h_{l-1} of both tiles: SMEM tiles -> stash[l-2] (rows >= n are clipped)
          for (int t = 0; t < 2; ++t)
            for (int c = 0; c < 4; ++c)
              tma_store_3d(&p.tm_stash, s_base + (t ? kOffHB : kOffHA) + c * kTileBytes, c * 64, (int)((pr * 2 + t) * kTileRows), l - 2);
          tma_store_commit();
        }
        if (l == 1) {
          if (FOLDED) {       // layers1.0 is entirely in the accumulator pre-load (xyz columns + bias + W[:, latent] z): no MMA
            if (elect_one()) { umma_commit(&hdr->acc_full[0]); umma_commit(&hdr->acc_full[1]); }
            __syncwarp();
            continue;
          }
          wait_two_chunks();
          latent_mma(0, true);
          latent_mma(1, true);
          release_two_chunks();
          continue;
        }
        if (l == 5 && !FOLDED) {   // latent part of layers2.0 for tile A first: tile B's latent rows are gathered under the hidden chunks
          wait_two_chunks();
          latent_mma(0, false);
          release_two_chunks();
        }
        for (int c = 0; c < 4; ++c, ++g) {
          mbar_wait(&hdr->w_full[g & 1u], (g >> 1) & 1u, p.err);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b = s_base + kOffW + (g & 1u) * kChunkBytes;
            for (int t = 0; t < 2; ++t) {
              const uint32_t d = tmem_base + (uint32_t)t * 256u;
              const uint32_t a = s_base + (t ? kOffHB : kOffHA) + c * kTileBytes;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_bf16(d, umma_desc(a + kk * 32, 16, 1024), umma_desc(b + kk * 32, 16, 1024), idesc, 1u);
              if (c == 3 && t == 0 && p.tma_stash) tma_store_wait_read();   // the epilogue overwrites the tiles after acc_full
              if (c == 3 && (FOLDED || l != 5 || t == 0)) umma_commit(&hdr->acc_full[t]);
            }
            umma_commit(&hdr->w_empty[g & 1u]);
          }
          __syncwarp();
        }
        if (l == 5 && !FOLDED) {
          wait_two_chunks();
          latent_mma(1, true);
          release_two_chunks();
        }
      }
    }
    if (p.tma_stash && elect_one()) tma_store_wait_all();
  } else if (warp >= 2 && warp < 6) {
    // ================================================================ latent gather: fp32 table rows -> bf16 swizzled A tile
    const int gw = warp - 2;
    uint32_t use = 0;
    if (!FOLDED)
    for (long long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
      for (int u = 0; u < 4; ++u, ++use) {          // (A,L1) (B,L1) (A,L5) (B,L5)
        const long long tile = pr * 2 + (u & 1);
        if (use > 0) mbar_wait(&hdr->lat_empty, (use - 1) & 1u, p.err);
        const int chunk = lane >> 4, piece = (lane & 15) >> 1, half = lane & 1;
        // table row of each of this warp's 32 points: one coalesced index load, broadcast by shuffle (no dependent load chain)
        const long long gr_l = tile * kTileRows + gw * 32 + lane;
        long long my_row = -1;
        if (gr_l < n_rows) my_row = p.index ? (long long)__ldg(p.index + gr_l) : gr_l;
#pragma unroll 16
        for (int i = 0; i < 32; ++i) {
          const int r = gw * 32 + i;
          const long long lrow = __shfl_sync(0xffffffffu, my_row, i);
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (lrow >= 0) f = __ldg(reinterpret_cast<const float4*>(p.latent + lrow * 128) + lane);
          uint2 v;
          v.x = pack_bf16x2(f.x, f.y); v.y = pack_bf16x2(f.z, f.w);
          *reinterpret_cast<uint2*>(smem + kOffLat + chunk * kTileBytes + sw128((uint32_t)r, (uint32_t)piece) + half * 8) = v;
        }
        fence_proxy_async();
        mbar_arrive(&hdr->lat_full);
      }
    }
  } else if (warp >= 6) {
    // ================================================================ epilogue: tile t, lane quadrant q (32 rows), column half
    const int e = warp - 6;
    const int t = e >> 3, half = (e >> 2) & 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int cbase = half * 128;
    uint8_t* hbuf = smem + (t ? kOffHB : kOffHA);
    float* dotbuf = reinterpret_cast<float*>(smem + kOffDot) + t * 128;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + cbase);
    uint32_t af_n = 0;
    float px = 0.f, py = 0.f, pz = 0.f;
    long long cur_slot = 0;
    {
      const long long gr = ((long long)blockIdx.x * 2 + t) * kTileRows + r;
      if (gr < n_rows) { cur_slot = sdf_slot<FOLDED>(p, gr); sdf_load_point<FOLDED>(p, cur_slot, px, py, pz); }
      // accumulator of layer 1 of the first pair
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t iv[32];
        sdf_acc_init(1, cbase + c0, px, py, pz, iv);
        tmem_st32(t_addr + c0, iv);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&hdr->h_ready[t]);
    }
    for (long long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
      const long long gr = (pr * 2 + t) * kTileRows + r;
      const bool valid = gr < n_rows;
      // points of this CTA's next pair (its layer-1 accumulator is initialised at the end of this pair)
      const long long npr = pr + gridDim.x;
      const long long ngr = (npr * 2 + t) * kTileRows + r;
      float nx = 0.f, ny = 0.f, nz = 0.f;
      long long next_slot = 0;
      if (npr < n_pairs && ngr < n_rows) { next_slot = sdf_slot<FOLDED>(p, ngr); sdf_load_point<FOLDED>(p, next_slot, nx, ny, nz); }
      for (int l = 1; l <= 7; ++l) {
        mbar_wait(&hdr->acc_full[t], af_n & 1u, p.err); ++af_n;
        tc_fence_after();
        const int nl = (l < 7) ? l + 1 : 1;                        // layer whose accumulator start value goes in behind the drain
        const bool init_next = (l < 7) || (npr < n_pairs);
        const float ix = (l < 7) ? px : nx, iy = (l < 7) ? py : ny, iz = (l < 7) ? pz : nz;
        float dot = 0.f;
        bf16* srow = (p.stash != nullptr && valid && (l == 7 || !p.tma_stash)) ? p.stash + ((size_t)(l - 1) * (size_t)p.n + (size_t)gr) * 256 + cbase : nullptr;
        uint32_t* mrow = (p.mstash != nullptr && valid) ? p.mstash + ((size_t)(l - 1) * (size_t)p.n + (size_t)gr) * 8 + half * 4 : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          const int col = cbase + c0;
          uint32_t acc[32];
          tmem_ld32(t_addr + c0, acc);
          tmem_ld_wait();
          uint4 pk[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            pk[u].x = pack_relu_bf16x2(__uint_as_float(acc[u * 8]), __uint_as_float(acc[u * 8 + 1]));
            pk[u].y = pack_relu_bf16x2(__uint_as_float(acc[u * 8 + 2]), __uint_as_float(acc[u * 8 + 3]));
            pk[u].z = pack_relu_bf16x2(__uint_as_float(acc[u * 8 + 4]), __uint_as_float(acc[u * 8 + 5]));
            pk[u].w = pack_relu_bf16x2(__uint_as_float(acc[u * 8 + 6]), __uint_as_float(acc[u * 8 + 7]));
          }
          if (l < 7) {
            const uint32_t chunk = (uint32_t)col >> 6, pbase = ((uint32_t)col & 63u) >> 3;
#pragma unroll
            for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(hbuf + chunk * kTileBytes + sw128((uint32_t)r, pbase + u)) = pk[u];
          } else {
            // layers2.6 (256 -> 1) on the bf16-rounded activations (what the backward sees in the stash)
            const float* w8 = c_sdf_aux + kAuxW8 + col;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              dot = fmaf(bf16lo_to_f(pk[u].x), w8[u * 8], dot); dot = fmaf(bf16hi_to_f(pk[u].x), w8[u * 8 + 1], dot);
              dot = fmaf(bf16lo_to_f(pk[u].y), w8[u * 8 + 2], dot); dot = fmaf(bf16hi_to_f(pk[u].y), w8[u * 8 + 3], dot);
              dot = fmaf(bf16lo_to_f(pk[u].z), w8[u * 8 + 4], dot); dot = fmaf(bf16hi_to_f(pk[u].z), w8[u * 8 + 5], dot);
              dot = fmaf(bf16lo_to_f(pk[u].w), w8[u * 8 + 6], dot); dot = fmaf(bf16hi_to_f(pk[u].w), w8[u * 8 + 7], dot);
            }
          }
          if (srow) { stg_256(srow + c0, pk[0], pk[1]); stg_256(srow + c0 + 16, pk[2], pk[3]); }
          if (mrow) {
            // ReLU mask of these 32 columns, 1 bit each (bit k = column 2k, bit 16+k = column 2k+1): the backward chain reads
            // 32 B per row and layer instead of the 512 B activation row.  h >= 0, so (h + 0x7fff) carries into bit 15 iff h != 0.
            const uint32_t w[16] = {pk[0].x, pk[0].y, pk[0].z, pk[0].w, pk[1].x, pk[1].y, pk[1].z, pk[1].w,
                                    pk[2].x, pk[2].y, pk[2].z, pk[2].w, pk[3].x, pk[3].y, pk[3].z, pk[3].w};
            uint32_t bits = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) bits |= ((w[k] + 0x7fff7fffu) >> (15 - k)) & (0x00010001u << k);
            mrow[c0 >> 5] = bits;
          }
          if (init_next) {
            uint32_t iv[32];
            sdf_acc_init(nl, col, ix, iy, iz, iv);
            tmem_st32(t_addr + c0, iv);
          }
        }
        if (l == 7) {       // the two column halves of a row meet through shared memory (named barrier of the tile's 8 warps)
          if (half == 1) dotbuf[r] = dot;
          asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
          if (half == 0 && valid) {
            const float sdf = tanhf(dot + dotbuf[r] + c_sdf_aux[kAuxB8]);
            if (FOLDED && p.trace_points != nullptr) {
              // one sphere-tracing step of ray `cur_slot` (rendering/raymarching.py:108-118): advance by the clamped distance, classify
              const float d = fminf(fmaxf(sdf + p.sdf_offset, -p.trace_clamp), p.trace_clamp);
              const float* dir = p.trace_dirs + cur_slot * 3;
              const float qx = px + __ldg(dir) * d, qy = py + __ldg(dir + 1) * d, qz = pz + __ldg(dir + 2) * d;
              float* dst = p.trace_points + cur_slot * 3;
              dst[0] = qx; dst[1] = qy; dst[2] = qz;
              const bool hit = d > 0.f && d < p.trace_threshold;
              const bool miss = p.trace_miss_y ? (qy > p.trace_radius) : (sqrtf(qx * qx + qy * qy + qz * qz) > p.trace_radius);
              if (hit) p.trace_hit[cur_slot] = 1;
              else if (!miss) p.next_index[atomicAdd(p.next_count, 1)] = (int)cur_slot;
              if (p.out != nullptr) p.out[cur_slot] = d;
            } else {
              p.out[FOLDED ? cur_slot : gr] = sdf;
            }
          }
        }
        tmem_st_wait();
        if (l < 7) fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core's async proxy
        tc_fence_before();
        mbar_arrive(&hdr->h_ready[t]);
      }
      px = nx; py = ny; pz = nz; cur_slot = next_slot;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------------------ backward
constexpr int kSdfBwdThreads = 576;     // 0 weight loader | 1 MMA issuer | 2-9 epilogue of tile A | 10-17 epilogue of tile B (quadrant x column half)
constexpr int kSdfBwdChunks = 24;       // W7^T, W6^T, W5[:, :256]^T, W4^T, W3^T, W2^T, 4 chunks each
constexpr int kSdfBwdStages = 3;
constexpr uint32_t kOffBW = kOffHB + 65536;
constexpr uint32_t kSdfBwdSmem = kOffBW + kSdfBwdStages * kChunkBytes;      // 231424 B

struct SdfBwdP {
  const float* gout; const float* out; const uint32_t* mstash; const char* wt_img; bf16* gstash;
  long long n, pairs;
  int* err;
  int tma_stash;            // g_7..g_2 leave through TMA tile stores out of the operand tiles in SMEM
  CUtensorMap tm_stash;     // bf16 [7][n][256], box [1][128][64], 128B swizzle
  float* gpoints;           // XYZ instantiation: d out / d xyz per point [n][3] = g_1 W1[:, 0:3] + g_5 W5[:, 256:259] (model/sdf_net.py:57,59)
};

__constant__ float c_sdf_bwd_xyz[2 * 3 * 256];   // XYZ: [layers1.0 | layers2.0][x, y, z][256 out-features]

// XYZ = true: additionally accumulates the gradient w.r.t. the point coordinates inside the drains of g_5 and g_1 (SDFNet.get_normals,
// model/sdf_net.py:118-128) -- no K = 131 GEMMs, and with gstash == NULL nothing but [n][3] floats leaves the SM.
template <bool XYZ>
__global__ void __launch_bounds__(kSdfBwdThreads, 1) sg_sdfnet_bwd_kernel(const __grid_constant__ SdfBwdP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SdfHdr* hdr = reinterpret_cast<SdfHdr*>(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // broadcast: the compiler can keep everything derived from it uniform
  if (tid == 0) {
    for (int i = 0; i < kSdfBwdStages; ++i) { mbar_init(&hdr->w_full[i], 1); mbar_init(&hdr->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&hdr->acc_full[i], 1); mbar_init(&hdr->h_ready[i], 256); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0) atomicExch(p.err, kErrSmemAlign);
    __trap();
  }
  const uint32_t s_base = smem_u32(smem);

  if (warp == 0) {
    // ================================================================ weight loader
    uint32_t st = 0, ph = 0;
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      for (int j = 0; j < kSdfBwdChunks; ++j) {
        mbar_wait(&hdr->w_empty[st], ph ^ 1u, p.err);
        if (elect_one()) {
          mbar_arrive_expect_tx(&hdr->w_full[st], kChunkBytes);
          bulk_g2s(s_base + kOffBW + st * kChunkBytes, p.wt_img + (size_t)j * kChunkBytes, kChunkBytes, &hdr->w_full[st]);
        }
        __syncwarp();
        if (++st == kSdfBwdStages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = umma_idesc(128, 256, false, false);
    uint32_t st = 0, ph = 0, hr_n[2] = {0, 0};
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      for (int l = 7; l >= 2; --l) {
        for (int t = 0; t < 2; ++t) { mbar_wait(&hdr->h_ready[t], hr_n[t] & 1u, p.err); ++hr_n[t]; }   // g_l in SMEM, accumulator drained
        tc_fence_after();
        if (p.tma_stash && elect_one()) {             // g_l of both tiles: SMEM operand tiles -> gstash[l-1] (rows >= n are clipped)
          for (int t = 0; t < 2; ++t)
            for (int c = 0; c < 4; ++c)
              tma_store_3d(&p.tm_stash, s_base + (t ? kOffHB : kOffHA) + c * kTileBytes, c * 64, (int)((pr * 2 + t) * kTileRows), l - 1);
          tma_store_commit();
        }
        for (int c = 0; c < 4; ++c) {
          mbar_wait(&hdr->w_full[st], ph, p.err);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b = s_base + kOffBW + st * kChunkBytes;
            for (int t = 0; t < 2; ++t) {
              const uint32_t d = tmem_base + (uint32_t)t * 256u;
              const uint32_t a = s_base + (t ? kOffHB : kOffHA) + c * kTileBytes;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_bf16(d, umma_desc(a + kk * 32, 16, 1024), umma_desc(b + kk * 32, 16, 1024), idesc, (c | kk) ? 1u : 0u);
              if (c == 3 && t == 0 && p.tma_stash) tma_store_wait_read();   // the epilogue overwrites the tiles after acc_full
              if (c == 3) umma_commit(&hdr->acc_full[t]);
            }
            umma_commit(&hdr->w_empty[st]);
          }
          __syncwarp();
          if (++st == kSdfBwdStages) { st = 0; ph ^= 1u; }
        }
      }
    }
    if (p.tma_stash && elect_one()) tma_store_wait_all();
  } else if (warp >= 2) {
    // ================================================================ epilogue: tile t, lane quadrant q (32 rows), column half
    const int e = warp - 2;
    const int t = e >> 3, half = (e >> 2) & 1;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int cbase = half * 128;
    uint8_t* hbuf = smem + (t ? kOffHB : kOffHA);
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + cbase);
    // step s of a pair reads mstash[6 - s] (s = 0: head, s = 8 - l: the mask of layer l's epilogue): 16 bytes per thread, fetched
    // one step ahead of its use
    auto load_mask = [&](long long pair, int sidx) -> uint4 {
      const long long g = (pair * 2 + t) * kTileRows + r;
      if (pair >= p.pairs || g >= p.n) return make_uint4(0u, 0u, 0u, 0u);
      return __ldg(reinterpret_cast<const uint4*>(p.mstash + ((size_t)(6 - sidx) * (size_t)p.n + (size_t)g) * 8 + half * 4));
    };
    uint32_t af_n = 0;
    float gxyz[3] = {0.f, 0.f, 0.f};
    float* xyzbuf = reinterpret_cast<float*>(hbuf);     // [128 rows][3], column half 1 -> column half 0: this tile's operand region is idle
                                                        // between the MMAs of layer 2 (acc_full) and the head of the CTA's next pair
    uint4 mcur = load_mask(blockIdx.x, 0);
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      const long long gr = (pr * 2 + t) * kTileRows + r;
      const bool valid = gr < p.n;
      const long long npr = pr + gridDim.x;
      // ---------- head: g7 = gout * tanh'(out) * w8 where h7 > 0 (model/sdf_net.py:50-51)
      {
        const uint4 mnext = load_mask(pr, 1);
        float s = 0.f;
        if (valid) { const float o = __ldg(p.out + gr); s = __ldg(p.gout + gr) * (1.f - o * o); }
        bf16* grow = p.gstash + ((size_t)6 * (size_t)p.n + (size_t)(valid ? gr : 0)) * 256 + cbase;
        const uint32_t mb[4] = {mcur.x, mcur.y, mcur.z, mcur.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = cbase + c * 32;
          const float* w8 = c_sdf_w8 + col;
          const uint32_t bits = mb[c];
          uint32_t o[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float lo = ((bits >> k) & 1u) ? s * w8[2 * k] : 0.f;
            const float hi = ((bits >> (16 + k)) & 1u) ? s * w8[2 * k + 1] : 0.f;
            o[k] = pack_bf16x2(lo, hi);
          }
          uint4 pk[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) pk[u] = make_uint4(o[u * 4], o[u * 4 + 1], o[u * 4 + 2], o[u * 4 + 3]);
          const uint32_t chunk = (uint32_t)col >> 6, pbase = ((uint32_t)col & 63u) >> 3;
#pragma unroll
          for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(hbuf + chunk * kTileBytes + sw128((uint32_t)r, pbase + u)) = pk[u];
          if (valid && !p.tma_stash && p.gstash != nullptr) { stg_256(grow + c * 32, pk[0], pk[1]); stg_256(grow + c * 32 + 16, pk[2], pk[3]); }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&hdr->h_ready[t]);
        mcur = mnext;
      }
      // ---------- layers 7..2: g_{l-1} = (g_l W_l) [h_{l-1} > 0]
      for (int l = 7; l >= 2; --l) {
        const int sn = (8 - l) + 1;                                   // next step: of this pair, or the head of this CTA's next pair
        const uint4 mnext = (sn <= 6) ? load_mask(pr, sn) : load_mask(npr, 0);
        mbar_wait(&hdr->acc_full[t], af_n & 1u, p.err); ++af_n;
        tc_fence_after();
        bf16* grow = p.gstash + ((size_t)(l - 2) * (size_t)p.n + (size_t)(valid ? gr : 0)) * 256 + cbase;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int c0 = c * 32, col = cbase + c0;
          uint32_t acc[32];
          tmem_ld32(t_addr + c0, acc);
          tmem_ld_wait();
          const uint32_t bits = c == 0 ? mcur.x : (c == 1 ? mcur.y : (c == 2 ? mcur.z : mcur.w));
          uint4 pk[4];
          const float* wxyz = c_sdf_bwd_xyz + (l == 6 ? 768 : 0) + col;     // XYZ: g_5 (drained at l = 6) meets layers2.0, g_1 (l = 2) layers1.0
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int wd = u * 4 + k;                              // bf16x2 word: columns 2*wd (bit wd) and 2*wd+1 (bit 16+wd)
              const float lo = ((bits >> wd) & 1u) ? __uint_as_float(acc[2 * wd]) : 0.f;
              const float hi = ((bits >> (16 + wd)) & 1u) ? __uint_as_float(acc[2 * wd + 1]) : 0.f;
              o[k] = pack_bf16x2(lo, hi);
              if (XYZ && (l == 6 || l == 2)) {
#pragma unroll
                for (int d = 0; d < 3; ++d) gxyz[d] = fmaf(hi, wxyz[d * 256 + 2 * wd + 1], fmaf(lo, wxyz[d * 256 + 2 * wd], gxyz[d]));
              }
            }
            pk[u] = make_uint4(o[0], o[1], o[2], o[3]);
          }
          if (l > 2) {
            const uint32_t chunk = (uint32_t)col >> 6, pbase = ((uint32_t)col & 63u) >> 3;
#pragma unroll
            for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(hbuf + chunk * kTileBytes + sw128((uint32_t)r, pbase + u)) = pk[u];
          }
          if (valid && (l == 2 || !p.tma_stash) && p.gstash != nullptr) { stg_256(grow + c0, pk[0], pk[1]); stg_256(grow + c0 + 16, pk[2], pk[3]); }
        }
        if (l > 2) {
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&hdr->h_ready[t]);
        } else {
          tc_fence_before();      // orders the accumulator reads before the head of the next pair arrives on h_ready
          if (XYZ) {              // the two column halves of a row meet through shared memory (named barrier of the tile's 8 warps)
            if (half == 1) { xyzbuf[r * 3] = gxyz[0]; xyzbuf[r * 3 + 1] = gxyz[1]; xyzbuf[r * 3 + 2] = gxyz[2]; }
            asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
            if (half == 0 && valid) {
              float* dst = p.gpoints + gr * 3;
              dst[0] = gxyz[0] + xyzbuf[r * 3]; dst[1] = gxyz[1] + xyzbuf[r * 3 + 1]; dst[2] = gxyz[2] + xyzbuf[r * 3 + 2];
            }
            asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");      // xyzbuf is reused by this CTA's next pair
            gxyz[0] = gxyz[1] = gxyz[2] = 0.f;
          }
        }
        mcur = mnext;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace sg

using namespace sg;

// bf16 [7][n][256] stash as a 3-D tensor map, box = one 128-row x 64-column operand tile
static bool sdf_stash_map(CUtensorMap* m, const void* base, long long n) {
  if (getenv("SG_B200_NO_TMA_STASH") != nullptr) return false;
  if (n >= (1LL << 31) || ((uintptr_t)base & 15)) return false;
  const uint64_t dims[3] = {256, (uint64_t)n, 7};
  const uint64_t str[2] = {512, (uint64_t)n * 512};
  const uint32_t box[3] = {64, 128, 1};
  const uint32_t one[3] = {1, 1, 1};
  return tma_make_map(m, base, 3, dims, str, box, one);
}

extern "C" int sg_sdfnet_fwd(const sg_sdfnet_fwd_args* a, void* stream) {
  if (!a || !a->points || !a->latent || !a->w_img || !a->aux || !a->out) return sg_fail(-1, "sg_sdfnet_fwd: null");
  if (a->n <= 0) return 0;
  SdfP p;
  memset(&p, 0, sizeof(p));
  p.points = a->points; p.latent = a->latent; p.index = a->index; p.n = a->n;
  p.w_img = (const char*)a->w_img; p.out = a->out; p.stash = (bf16*)a->stash; p.mstash = (uint32_t*)a->mask_stash;
  const long long tiles = (a->n + kTileRows - 1) / kTileRows;
  p.pairs = (tiles + 1) / 2;
  p.err = sg_error_word();
  p.tma_stash = (p.stash != nullptr && sdf_stash_map(&p.tm_stash, p.stash, a->n)) ? 1 : 0;
  static PerDevice attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(sg_sdfnet_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfSmem);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    attr.done();
  }
  // biases / xyz columns / head -> constant bank (stream ordered, capturable; one SDFNet per stream at a time)
  cudaError_t e = cudaMemcpyToSymbolAsync(c_sdf_aux, a->aux, sizeof(float) * kAuxFloats, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  const int grid = (int)std::min<long long>(p.pairs, sg_num_sms());
  sg_sdfnet_fwd_kernel<false><<<grid, kSdfThreads, kSdfSmem, (cudaStream_t)stream>>>(p);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

// Single-latent inference (evaluate_in_batches / get_voxels / get_normals / sphere tracing): see sg_b200.h
extern "C" int sg_sdfnet_infer(const sg_sdfnet_infer_args* a, void* stream) {
  if (!a || !a->w_img || !a->aux) return sg_fail(-1, "sg_sdfnet_infer: null weights");
  const bool trace = a->trace_points != nullptr;
  if (!trace && !a->out) return sg_fail(-1, "sg_sdfnet_infer: null out");
  if (!trace && a->grid_r <= 0 && !a->points) return sg_fail(-1, "sg_sdfnet_infer: no point source (points, grid_r or trace_points)");
  if (a->grid_r > 0 && (!a->grid_axis || a->grid_r > 1024)) return sg_fail(-2, "sg_sdfnet_infer: grid needs grid_axis[3][r], r <= 1024");
  if (trace && (!a->trace_dirs || !a->trace_hit || !a->next_index || !a->next_count || !a->ray_index))
    return sg_fail(-3, "sg_sdfnet_infer: a tracing step needs trace_dirs, trace_hit, ray_index, next_index, next_count");
  if (a->n <= 0) return 0;
  if (a->n >= (1LL << 31)) return sg_fail(-4, "sg_sdfnet_infer: n must be < 2^31");
  SdfP p;
  memset(&p, 0, sizeof(p));
  p.points = a->points; p.n = a->n; p.n_ptr = a->n_ptr; p.ray_index = a->ray_index;
  p.grid_r = a->grid_r; p.grid_axis = a->grid_axis;
  p.w_img = (const char*)a->w_img; p.out = a->out; p.mstash = (uint32_t*)a->mask_stash;
  p.trace_points = a->trace_points; p.trace_dirs = a->trace_dirs; p.trace_hit = a->trace_hit;
  p.next_index = a->next_index; p.next_count = a->next_count;
  p.sdf_offset = a->sdf_offset; p.trace_clamp = a->trace_clamp; p.trace_threshold = a->trace_threshold; p.trace_radius = a->trace_radius;
  p.trace_miss_y = a->trace_miss_y;
  const long long tiles = (a->n + kTileRows - 1) / kTileRows;
  p.pairs = (tiles + 1) / 2;
  p.err = sg_error_word();
  static PerDevice attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(sg_sdfnet_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfSmem);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    attr.done();
  }
  cudaError_t e = cudaMemcpyToSymbolAsync(c_sdf_aux, a->aux, sizeof(float) * kAuxFloats, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  const int grid = (int)std::min<long long>(p.pairs, sg_num_sms());
  sg_sdfnet_fwd_kernel<true><<<grid, kSdfThreads, kSdfSmem, (cudaStream_t)stream>>>(p);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_sdfnet_fwd_layout(int32_t* chunks, int32_t* chunk_bytes, int32_t* aux_floats) {
  if (chunks) *chunks = kSdfImgChunks;
  if (chunk_bytes) *chunk_bytes = (int32_t)kChunkBytes;
  if (aux_floats) *aux_floats = kAuxFloats;
  return 0;
}

extern "C" int sg_sdfnet_bwd(const sg_sdfnet_bwd_args* a, void* stream) {
  if (!a || !a->gout || !a->out || !a->mask_stash || !a->wt_img || !a->w8 || (!a->gstash && !a->gpoints)) return sg_fail(-1, "sg_sdfnet_bwd: null");
  if (a->gpoints && !a->xyz_w) return sg_fail(-1, "sg_sdfnet_bwd: gpoints needs xyz_w");
  if (a->n <= 0) return 0;
  SdfBwdP p;
  memset(&p, 0, sizeof(p));
  p.gout = a->gout; p.out = a->out; p.mstash = (const uint32_t*)a->mask_stash; p.wt_img = (const char*)a->wt_img; p.gstash = (bf16*)a->gstash;
  p.n = a->n;
  const long long tiles = (a->n + kTileRows - 1) / kTileRows;
  p.pairs = (tiles + 1) / 2;
  p.err = sg_error_word();
  p.gpoints = a->gpoints;
  p.tma_stash = (p.gstash != nullptr && sdf_stash_map(&p.tm_stash, p.gstash, a->n)) ? 1 : 0;
  static PerDevice attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(sg_sdfnet_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfBwdSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sg_sdfnet_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfBwdSmem);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    attr.done();
  }
  cudaError_t e = cudaMemcpyToSymbolAsync(c_sdf_w8, a->w8, sizeof(float) * 256, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  const int grid = (int)std::min<long long>(p.pairs, sg_num_sms());
  if (p.gpoints != nullptr) {
    e = cudaMemcpyToSymbolAsync(c_sdf_bwd_xyz, a->xyz_w, sizeof(float) * 2 * 3 * 256, 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    sg_sdfnet_bwd_kernel<true><<<grid, kSdfBwdThreads, kSdfBwdSmem, (cudaStream_t)stream>>>(p);
  } else {
    sg_sdfnet_bwd_kernel<false><<<grid, kSdfBwdThreads, kSdfBwdSmem, (cudaStream_t)stream>>>(p);
  }
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
