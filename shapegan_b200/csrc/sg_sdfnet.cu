// Fused persistent DeepSDF MLP forward (model/sdf_net.py:56-61): all 8 layers of a pair of 128-point tiles inside one
// CTA, bf16 operands / fp32 accumulation, activations never leave the SM.
//
//   TMEM  : two 128 x 256 fp32 accumulators (tile A: columns 0-255, tile B: 256-511)
//   SMEM  : hA, hB   64 KB each  hidden activations of the two tiles, K-major 128B-swizzled (4 chunks of [128][64] bf16),
//                                 overwritten in place by the epilogue of every layer (it runs after the layer's MMAs retire)
//           lat      32 KB       latent operand of ONE tile (2 chunks), gathered by index from the fp32 table
//           w[2]     64 KB       two stages of streamed weights, one chunk = [256 n][64 k] bf16 = 32 KB, 28 chunks / pair:
//                                 L1 latent(2) L2(4) L3(4) L4(4) L5 hidden(4) L5 latent(2) L6(4) L7(4); every chunk feeds BOTH tiles
//   warps : 0 weight loader (1-D TMA bulk copies) | 1 MMA issuer (tcgen05.mma M128 N256 K16) | 4-7 latent gather |
//           8-11 epilogue of tile A | 12-15 epilogue of tile B  (TMEM -> +bias/xyz -> ReLU -> bf16 -> SMEM [+ HBM stash])
//   The xyz columns of layers 1 and 5 (K = 3) and all biases are applied in fp32 in the epilogue; layer 8 (256 -> 1) + tanh is a
//   per-row dot product folded into the epilogue of layer 7.
#include <algorithm>
#include <cstring>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

constexpr int kSdfThreads = 512;
constexpr int kSdfChunks = 28;
constexpr uint32_t kChunkBytes = 32768;
constexpr uint32_t kSdfHdr = 2048;
constexpr uint32_t kOffHA = kSdfHdr;
constexpr uint32_t kOffHB = kOffHA + 65536;
constexpr uint32_t kOffLat = kOffHB + 65536;
constexpr uint32_t kOffW = kOffLat + 32768;
constexpr uint32_t kSdfSmem = kOffW + 2 * kChunkBytes;      // 231424 B

struct SdfHdr {
  uint64_t w_full[2], w_empty[2];
  uint64_t lat_full, lat_empty;
  uint64_t acc_full[2], h_ready[2];
  uint32_t tmem_base;
};

struct SdfP {
  const float* points; const float* latent; const int* index; long long n;
  const char* w_img; const float* aux; float* out; bf16* stash;
  long long pairs;
  int* err;
};

// aux layout (floats)
constexpr int kAuxXb1 = 0;          // float4[256] {w_x, w_y, w_z, bias} of layers1.0
constexpr int kAuxXb5 = 1024;       // float4[256] of layers2.0
constexpr int kAuxBias = 2048;      // [5][256]: layers1.2, 1.4, 1.6, layers2.2, 2.4
constexpr int kAuxW8 = 2048 + 5 * 256;
constexpr int kAuxB8 = kAuxW8 + 256;

__global__ void __launch_bounds__(kSdfThreads, 1) sg_sdfnet_fwd_kernel(const __grid_constant__ SdfP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SdfHdr* hdr = reinterpret_cast<SdfHdr*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->w_full[i], 1); mbar_init(&hdr->w_empty[i], 1);
      mbar_init(&hdr->acc_full[i], 1); mbar_init(&hdr->h_ready[i], 128);
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
    if (lane == 0) {
      uint32_t g = 0;
      for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
        for (int j = 0; j < kSdfChunks; ++j, ++g) {
          const uint32_t st = g & 1u;
          mbar_wait(&hdr->w_empty[st], ((g >> 1) & 1u) ^ 1u, p.err);
          mbar_arrive_expect_tx(&hdr->w_full[st], kChunkBytes);
          bulk_g2s(s_base + kOffW + st * kChunkBytes, p.w_img + (size_t)j * kChunkBytes, kChunkBytes, &hdr->w_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = umma_idesc(128, 256, false, false);
    uint32_t g = 0, lat_n = 0, hr_n[2] = {0, 0};
    bool first = true;
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      // ---------- layer 1: latent part (2 resident weight chunks, one tile after the other through the shared latent tile)
      if (!first) {
        for (int t = 0; t < 2; ++t) { mbar_wait(&hdr->h_ready[t], hr_n[t] & 1u, p.err); ++hr_n[t]; }   // accumulators drained
      }
      first = false;
      mbar_wait(&hdr->w_full[g & 1u], (g >> 1) & 1u, p.err);
      mbar_wait(&hdr->w_full[(g + 1) & 1u], ((g + 1) >> 1) & 1u, p.err);
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&hdr->lat_full, lat_n & 1u, p.err); ++lat_n;
        tc_fence_after();
        if (lane == 0) {
          const uint32_t d = tmem_base + (uint32_t)t * 256u;
          for (int c = 0; c < 2; ++c) {
            const uint32_t a = s_base + kOffLat + c * kTileBytes, b = s_base + kOffW + ((g + c) & 1u) * kChunkBytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d, umma_desc(a + kk * 32, 16, 1024), umma_desc(b + kk * 32, 16, 1024), idesc, (c | kk) ? 1u : 0u);
          }
          umma_commit(&hdr->lat_empty);
          umma_commit(&hdr->acc_full[t]);
        }
        __syncwarp();
      }
      if (lane == 0) { umma_commit(&hdr->w_empty[g & 1u]); umma_commit(&hdr->w_empty[(g + 1) & 1u]); }
      __syncwarp();
      g += 2;
      // ---------- layers 2..7
      for (int l = 2; l <= 7; ++l) {
        for (int t = 0; t < 2; ++t) { mbar_wait(&hdr->h_ready[t], hr_n[t] & 1u, p.err); ++hr_n[t]; }     // h[t] = activations of layer l-1
        tc_fence_after();
        for (int c = 0; c < 4; ++c, ++g) {
          mbar_wait(&hdr->w_full[g & 1u], (g >> 1) & 1u, p.err);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b = s_base + kOffW + (g & 1u) * kChunkBytes;
            for (int t = 0; t < 2; ++t) {
              const uint32_t d = tmem_base + (uint32_t)t * 256u;
              const uint32_t a = s_base + (t ? kOffHB : kOffHA) + c * kTileBytes;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_bf16(d, umma_desc(a + kk * 32, 16, 1024), umma_desc(b + kk * 32, 16, 1024), idesc, (c | kk) ? 1u : 0u);
              if (c == 3 && l != 5) umma_commit(&hdr->acc_full[t]);
            }
            umma_commit(&hdr->w_empty[g & 1u]);
          }
          __syncwarp();
        }
        if (l == 5) {   // + latent part of layers2.0 on top of the hidden part
          mbar_wait(&hdr->w_full[g & 1u], (g >> 1) & 1u, p.err);
          mbar_wait(&hdr->w_full[(g + 1) & 1u], ((g + 1) >> 1) & 1u, p.err);
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&hdr->lat_full, lat_n & 1u, p.err); ++lat_n;
            tc_fence_after();
            if (lane == 0) {
              const uint32_t d = tmem_base + (uint32_t)t * 256u;
              for (int c = 0; c < 2; ++c) {
                const uint32_t a = s_base + kOffLat + c * kTileBytes, b = s_base + kOffW + ((g + c) & 1u) * kChunkBytes;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16(d, umma_desc(a + kk * 32, 16, 1024), umma_desc(b + kk * 32, 16, 1024), idesc, 1u);
              }
              umma_commit(&hdr->lat_empty);
              umma_commit(&hdr->acc_full[t]);
            }
            __syncwarp();
          }
          if (lane == 0) { umma_commit(&hdr->w_empty[g & 1u]); umma_commit(&hdr->w_empty[(g + 1) & 1u]); }
          __syncwarp();
          g += 2;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================================================ latent gather: fp32 table rows -> bf16 swizzled A tile
    const int gw = warp - 4;
    uint32_t use = 0;
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      for (int u = 0; u < 4; ++u, ++use) {          // (A,L1) (B,L1) (A,L5) (B,L5)
        const long long tile = pr * 2 + (u & 1);
        if (use > 0) mbar_wait(&hdr->lat_empty, (use - 1) & 1u, p.err);
        const int chunk = lane >> 4, piece = (lane & 15) >> 1, half = lane & 1;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          const int r = gw * 32 + i;
          const long long gr = tile * kTileRows + r;
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gr < p.n) {
            const long long lrow = p.index ? (long long)__ldg(p.index + gr) : gr;
            f = __ldg(reinterpret_cast<const float4*>(p.latent + lrow * 128) + lane);
          }
          uint2 v;
          v.x = pack_bf16x2(f.x, f.y); v.y = pack_bf16x2(f.z, f.w);
          *reinterpret_cast<uint2*>(smem + kOffLat + chunk * kTileBytes + sw128((uint32_t)r, (uint32_t)piece) + half * 8) = v;
        }
        fence_proxy_async();
        mbar_arrive(&hdr->lat_full);
      }
    }
  } else if (warp >= 8) {
    // ================================================================ epilogue of tile t (t = 0: warps 8-11, t = 1: warps 12-15)
    const int t = (warp - 8) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint8_t* hbuf = smem + (t ? kOffHB : kOffHA);
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
    const float4* xb1 = reinterpret_cast<const float4*>(p.aux + kAuxXb1);
    const float4* xb5 = reinterpret_cast<const float4*>(p.aux + kAuxXb5);
    const float4* w8 = reinterpret_cast<const float4*>(p.aux + kAuxW8);
    const float b8 = __ldg(p.aux + kAuxB8);
    uint32_t af_n = 0;
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      const long long gr = (pr * 2 + t) * kTileRows + r;
      const bool valid = gr < p.n;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (valid) { px = __ldg(p.points + gr * 3); py = __ldg(p.points + gr * 3 + 1); pz = __ldg(p.points + gr * 3 + 2); }
      for (int l = 1; l <= 7; ++l) {
        mbar_wait(&hdr->acc_full[t], af_n & 1u, p.err); ++af_n;
        tc_fence_after();
        const bool xyz = (l == 1 || l == 5);
        const float4* xb = (l == 1) ? xb1 : xb5;
        const int bslot = (l < 5) ? l - 2 : l - 3;            // layers 2,3,4 -> 0,1,2 ; layers 6,7 -> 3,4
        const float4* bias = reinterpret_cast<const float4*>(p.aux + kAuxBias + bslot * 256);
        float dot = 0.f;
        bf16* srow = (p.stash != nullptr && valid) ? p.stash + ((size_t)(l - 1) * (size_t)p.n + (size_t)gr) * 256 : nullptr;
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t acc[32];
          __syncwarp();
          tmem_ld32(t_addr + c0, acc);
          tmem_ld_wait();
          float v[32];
          if (xyz) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float4 w = __ldg(xb + c0 + j);
              v[j] = fmaxf(__uint_as_float(acc[j]) + w.w + w.x * px + w.y * py + w.z * pz, 0.f);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(bias + ((c0 + j) >> 2));
              v[j] = fmaxf(__uint_as_float(acc[j]) + b.x, 0.f);
              v[j + 1] = fmaxf(__uint_as_float(acc[j + 1]) + b.y, 0.f);
              v[j + 2] = fmaxf(__uint_as_float(acc[j + 2]) + b.z, 0.f);
              v[j + 3] = fmaxf(__uint_as_float(acc[j + 3]) + b.w, 0.f);
            }
          }
          uint4 pk[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            pk[u].x = pack_bf16x2(v[u * 8], v[u * 8 + 1]); pk[u].y = pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]);
            pk[u].z = pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]); pk[u].w = pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]);
          }
          if (l < 7) {
            const uint32_t chunk = (uint32_t)c0 >> 6, pbase = ((uint32_t)c0 & 63u) >> 3;
#pragma unroll
            for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(hbuf + chunk * kTileBytes + sw128((uint32_t)r, pbase + u)) = pk[u];
          } else {
            // layers2.6 (256 -> 1) on the bf16-rounded activations (what the backward sees in the stash)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 w = __ldg(w8 + ((c0 + j) >> 2));
              dot += bf16_round(v[j]) * w.x + bf16_round(v[j + 1]) * w.y + bf16_round(v[j + 2]) * w.z + bf16_round(v[j + 3]) * w.w;
            }
          }
          if (srow) {
#pragma unroll
            for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(srow + c0 + u * 8) = pk[u];
          }
        }
        if (l == 7 && valid) p.out[gr] = tanhf(dot + b8);
        if (l < 7) fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core's async proxy
        tc_fence_before();
        mbar_arrive(&hdr->h_ready[t]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_sdfnet_fwd(const sg_sdfnet_fwd_args* a, void* stream) {
  if (!a || !a->points || !a->latent || !a->w_img || !a->aux || !a->out) return sg_fail(-1, "sg_sdfnet_fwd: null");
  if (a->n <= 0) return 0;
  SdfP p;
  memset(&p, 0, sizeof(p));
  p.points = a->points; p.latent = a->latent; p.index = a->index; p.n = a->n;
  p.w_img = (const char*)a->w_img; p.aux = a->aux; p.out = a->out; p.stash = (bf16*)a->stash;
  const long long tiles = (a->n + kTileRows - 1) / kTileRows;
  p.pairs = (tiles + 1) / 2;
  p.err = sg_error_word();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sg_sdfnet_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfSmem);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = (int)std::min<long long>(p.pairs, sg_num_sms());
  sg_sdfnet_fwd_kernel<<<grid, kSdfThreads, kSdfSmem, (cudaStream_t)stream>>>(p);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_sdfnet_fwd_layout(int32_t* chunks, int32_t* chunk_bytes, int32_t* aux_floats) {
  if (chunks) *chunks = kSdfChunks;
  if (chunk_bytes) *chunk_bytes = (int32_t)kChunkBytes;
  if (aux_floats) *aux_floats = kAuxB8 + 1;
  return 0;
}
