// Fused persistent DeepSDF MLP forward (model/sdf_net.py:56-61): all 8 layers of a pair of 128-point tiles inside one
// CTA, bf16 operands / fp32 accumulation, activations never leave the SM.
//
//   TMEM  : two 128 x 256 fp32 accumulators (tile A: columns 0-255, tile B: 256-511)
//   SMEM  : hA, hB   64 KB each  hidden activations of the two tiles, K-major 128B-swizzled (4 chunks of [128][64] bf16),
//                                 overwritten in place by the epilogue of every layer (it runs after the layer's MMAs retire);
//                                 chunks 0-1 double as the latent operand (gathered by index from the fp32 table) for layer 1
//                                 and, once the hidden part of layer 5 has consumed h4, for the latent part of layer 5
//           w[3]     96 KB       three stages of streamed weights, one chunk = [256 n][64 k] bf16 = 32 KB, 28 chunks / pair:
//                                 L1 latent(2) L2(4) L3(4) L4(4) L5 hidden(4) L5 latent(2) L6(4) L7(4); every chunk feeds BOTH tiles
//   warps : 0 weight loader (1-D TMA bulk copies) | 1 MMA issuer (tcgen05.mma M128 N256 K16) | 4-7 latent gather |
//           8-11 epilogue of tile A | 12-15 epilogue of tile B  (TMEM -> +bias/xyz -> ReLU -> bf16 -> SMEM [+ HBM stash])
//   The xyz columns of layers 1 and 5 (K = 3) and all biases are applied in fp32 in the epilogue; layer 8 (256 -> 1) + tanh is a
//   per-row dot product folded into the epilogue of layer 7.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

constexpr int kSdfThreads = 512;
constexpr int kSdfChunks = 28;
constexpr int kSdfStages = 3;
constexpr uint32_t kChunkBytes = 32768;
constexpr uint32_t kSdfHdr = 2048;
constexpr uint32_t kOffHA = kSdfHdr;
constexpr uint32_t kOffHB = kOffHA + 65536;
constexpr uint32_t kOffW = kOffHB + 65536;
constexpr uint32_t kSdfSmem = kOffW + kSdfStages * kChunkBytes;      // 231424 B

struct SdfHdr {
  uint64_t w_full[kSdfStages], w_empty[kSdfStages];
  uint64_t lat_full[2], lat_free[2];     // per tile: latent rows staged in h[t] chunks 0-1 / those chunks may be overwritten
  uint64_t acc_full[2], h_ready[2];
  uint32_t tmem_base;
};

struct SdfP {
  const float* points; const float* latent; const int* index; long long n;
  const char* w_img; const float* aux; float* out; bf16* stash;
  long long pairs;
  int* err;
};

// aux layout (floats), resident in __constant__ memory: the epilogue reads it with warp-uniform addresses
constexpr int kAuxXb1 = 0;          // float4[256] {w_x, w_y, w_z, bias} of layers1.0
constexpr int kAuxXb5 = 1024;       // float4[256] of layers2.0
constexpr int kAuxBias = 2048;      // [5][256]: layers1.2, 1.4, 1.6, layers2.2, 2.4
constexpr int kAuxW8 = 2048 + 5 * 256;
constexpr int kAuxB8 = kAuxW8 + 256;
constexpr int kAuxFloats = kAuxB8 + 4;
__constant__ float c_sdf_aux[kAuxFloats];

// one 16-column block of the epilogue: +bias (/ +xyz) -> ReLU -> bf16 -> SMEM operand tile (+ HBM stash) or the layer-8 dot.
// kConstAux: aux block read from __constant__ memory (indexed LDC) or from global memory (broadcast LDG.128).
template <bool kConstAux>
__device__ __forceinline__ void sdf_epi16(const uint32_t (&acc)[16], int c, int l, bool xyz, int xoff, int boff, float px, float py,
                                          float pz, uint8_t* hbuf, int r, bf16* srow, float& dot, const float* gaux) {
  float v[16];
  if (xyz) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 w = kConstAux ? *reinterpret_cast<const float4*>(&c_sdf_aux[xoff + 4 * (c + j)])
                                 : __ldg(reinterpret_cast<const float4*>(gaux + xoff) + c + j);
      v[j] = fmaxf(__uint_as_float(acc[j]) + w.w + w.x * px + w.y * py + w.z * pz, 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = kConstAux ? *reinterpret_cast<const float4*>(&c_sdf_aux[boff + c + j])
                                 : __ldg(reinterpret_cast<const float4*>(gaux + boff + c + j));
      v[j] = fmaxf(__uint_as_float(acc[j]) + b.x, 0.f);
      v[j + 1] = fmaxf(__uint_as_float(acc[j + 1]) + b.y, 0.f);
      v[j + 2] = fmaxf(__uint_as_float(acc[j + 2]) + b.z, 0.f);
      v[j + 3] = fmaxf(__uint_as_float(acc[j + 3]) + b.w, 0.f);
    }
  }
  uint4 pk[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    pk[u].x = pack_bf16x2(v[u * 8], v[u * 8 + 1]); pk[u].y = pack_bf16x2(v[u * 8 + 2], v[u * 8 + 3]);
    pk[u].z = pack_bf16x2(v[u * 8 + 4], v[u * 8 + 5]); pk[u].w = pack_bf16x2(v[u * 8 + 6], v[u * 8 + 7]);
  }
  if (l < 7) {
    const uint32_t chunk = (uint32_t)c >> 6, pbase = ((uint32_t)c & 63u) >> 3;
    *reinterpret_cast<uint4*>(hbuf + chunk * kTileBytes + sw128((uint32_t)r, pbase)) = pk[0];
    *reinterpret_cast<uint4*>(hbuf + chunk * kTileBytes + sw128((uint32_t)r, pbase + 1)) = pk[1];
  } else {
    // layers2.6 (256 -> 1) on the bf16-rounded activations (what the backward sees in the stash)
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 w = kConstAux ? *reinterpret_cast<const float4*>(&c_sdf_aux[kAuxW8 + c + j])
                                 : __ldg(reinterpret_cast<const float4*>(gaux + kAuxW8 + c + j));
      dot += bf16_round(v[j]) * w.x + bf16_round(v[j + 1]) * w.y + bf16_round(v[j + 2]) * w.z + bf16_round(v[j + 3]) * w.w;
    }
  }
  if (srow) {
    *reinterpret_cast<uint4*>(srow + c) = pk[0];
    *reinterpret_cast<uint4*>(srow + c + 8) = pk[1];
  }
}

template <bool kConstAux>
__global__ void __launch_bounds__(kSdfThreads, 1) sg_sdfnet_fwd_kernel(const __grid_constant__ SdfP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SdfHdr* hdr = reinterpret_cast<SdfHdr*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kSdfStages; ++i) { mbar_init(&hdr->w_full[i], 1); mbar_init(&hdr->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->acc_full[i], 1); mbar_init(&hdr->h_ready[i], 128);
      mbar_init(&hdr->lat_full[i], 128); mbar_init(&hdr->lat_free[i], 1);
    }
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
    // ================================================================ weight loader: 28 chunks per pair through 3 stages
    if (lane == 0) {
      uint32_t g = 0;
      for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
        for (int j = 0; j < kSdfChunks; ++j, ++g) {
          const uint32_t st = g % kSdfStages, use = g / kSdfStages;
          mbar_wait(&hdr->w_empty[st], (use & 1u) ^ 1u, p.err);
          mbar_arrive_expect_tx(&hdr->w_full[st], kChunkBytes);
          bulk_g2s(s_base + kOffW + st * kChunkBytes, p.w_img + (size_t)j * kChunkBytes, kChunkBytes, &hdr->w_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer: every weight chunk feeds both tiles
    const uint32_t idesc = umma_idesc(128, 256, false, false);
    uint32_t g = 0, lat_n = 0, hr_n = 0;
    bool first = true;
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      for (int j = 0; j < kSdfChunks; ++j, ++g) {
        // chunk schedule: [0,2) L1 latent | [2,14) L2-L4 | [14,18) L5 hidden | [18,20) L5 latent | [20,28) L6-L7
        const bool lat_chunk = (j < 2) || (j == 18 || j == 19);
        int a_chunk, layer_first, layer_last;
        if (j < 2) { a_chunk = j; layer_first = (j == 0); layer_last = (j == 1); }
        else if (j < 18) { a_chunk = (j - 2) & 3; layer_first = (a_chunk == 0); layer_last = (a_chunk == 3) && (j < 14); }
        else if (j < 20) { a_chunk = j - 18; layer_first = 0; layer_last = (j == 19); }
        else { a_chunk = (j - 20) & 3; layer_first = (a_chunk == 0); layer_last = (a_chunk == 3); }
        if (j == 0) {
          if (!first) { for (int t = 0; t < 2; ++t) mbar_wait(&hdr->h_ready[t], hr_n & 1u, p.err); ++hr_n; }   // accumulators drained (layer 7 of the previous pair)
          first = false;
          for (int t = 0; t < 2; ++t) mbar_wait(&hdr->lat_full[t], lat_n & 1u, p.err);
          ++lat_n;
        } else if (j == 18) {
          for (int t = 0; t < 2; ++t) mbar_wait(&hdr->lat_full[t], lat_n & 1u, p.err);
          ++lat_n;
        } else if (!lat_chunk && layer_first) {
          for (int t = 0; t < 2; ++t) mbar_wait(&hdr->h_ready[t], hr_n & 1u, p.err);      // h[t] = activations of the previous layer
          ++hr_n;
        }
        const uint32_t st = g % kSdfStages, use = g / kSdfStages;
        mbar_wait(&hdr->w_full[st], use & 1u, p.err);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b = s_base + kOffW + st * kChunkBytes;
          for (int t = 0; t < 2; ++t) {
            const uint32_t d = tmem_base + (uint32_t)t * 256u;
            const uint32_t a = s_base + (t ? kOffHB : kOffHA) + a_chunk * kTileBytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(d, umma_desc(a + kk * 32, 16, 1024), umma_desc(b + kk * 32, 16, 1024), idesc, (layer_first && kk == 0) ? 0u : 1u);
            if (layer_last) umma_commit(&hdr->acc_full[t]);
            if (j == 17 || j == 27) umma_commit(&hdr->lat_free[t]);      // h[t] chunks 0-1 may now receive latent rows
          }
          umma_commit(&hdr->w_empty[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================================================ latent gather: fp32 table rows -> bf16 rows of h[t] chunks 0-1
    const int gw = warp - 4;
    uint32_t fr_n = 0;
    bool first = true;
    for (long long pr = blockIdx.x; pr < p.pairs; pr += gridDim.x) {
      for (int u = 0; u < 2; ++u) {                // u = 0: for layer 1, u = 1: for the latent part of layer 5
        if (!(first && u == 0)) { for (int t = 0; t < 2; ++t) mbar_wait(&hdr->lat_free[t], fr_n & 1u, p.err); ++fr_n; }
        for (int t = 0; t < 2; ++t) {
          const long long tile = pr * 2 + t;
          uint8_t* hb = smem + (t ? kOffHB : kOffHA);
          const int chunk = lane >> 4, piece = (lane & 15) >> 1, half = lane & 1;
#pragma unroll 8
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
            *reinterpret_cast<uint2*>(hb + chunk * kTileBytes + sw128((uint32_t)r, (uint32_t)piece) + half * 8) = v;
          }
          fence_proxy_async();
          mbar_arrive(&hdr->lat_full[t]);
        }
      }
      first = false;
    }
  } else if (warp >= 8) {
    // ================================================================ epilogue of tile t (t = 0: warps 8-11, t = 1: warps 12-15)
    const int t = (warp - 8) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint8_t* hbuf = smem + (t ? kOffHB : kOffHA);
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
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
        const int xoff = (l == 1) ? kAuxXb1 : kAuxXb5;
        const int boff = kAuxBias + ((l < 5) ? l - 2 : l - 3) * 256;      // layers 2,3,4 -> slots 0,1,2 ; layers 6,7 -> 3,4
        float dot = 0.f;
        bf16* srow = (p.stash != nullptr && valid) ? p.stash + ((size_t)(l - 1) * (size_t)p.n + (size_t)gr) * 256 : nullptr;
        uint32_t a0[16], a1[16];
        __syncwarp();
        tmem_ld16(t_addr, a0);
        for (int c0 = 0; c0 < 256; c0 += 32) {
          tmem_ld_wait();
          tmem_ld16(t_addr + c0 + 16, a1);                 // next 16 columns in flight while these are processed
          sdf_epi16<kConstAux>(a0, c0, l, xyz, xoff, boff, px, py, pz, hbuf, r, srow, dot, p.aux);
          tmem_ld_wait();
          if (c0 + 32 < 256) tmem_ld16(t_addr + c0 + 32, a0);
          sdf_epi16<kConstAux>(a1, c0 + 16, l, xyz, xoff, boff, px, py, pz, hbuf, r, srow, dot, p.aux);
        }
        if (l == 7 && valid) p.out[gr] = tanhf(dot + (kConstAux ? c_sdf_aux[kAuxB8] : __ldg(p.aux + kAuxB8)));
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
  static int const_aux = -1;
  if (const_aux < 0) { const char* e = getenv("SG_B200_SDF_CONST_AUX"); const_aux = (e && e[0] == '1') ? 1 : 0; }
  if (const_aux) {  // aux block (16 KB) in __constant__ memory: stream-ordered device-to-device copy, graph capturable
    cudaError_t e = cudaMemcpyToSymbolAsync(c_sdf_aux, a->aux, (size_t)(kAuxB8 + 1) * sizeof(float), 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  }
  const long long tiles = (a->n + kTileRows - 1) / kTileRows;
  p.pairs = (tiles + 1) / 2;
  p.err = sg_error_word();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sg_sdfnet_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sg_sdfnet_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfSmem);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = (int)std::min<long long>(p.pairs, sg_num_sms());
  if (const_aux) sg_sdfnet_fwd_kernel<true><<<grid, kSdfThreads, kSdfSmem, (cudaStream_t)stream>>>(p);
  else sg_sdfnet_fwd_kernel<false><<<grid, kSdfThreads, kSdfSmem, (cudaStream_t)stream>>>(p);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_sdfnet_fwd_layout(int32_t* chunks, int32_t* chunk_bytes, int32_t* aux_floats) {
  if (chunks) *chunks = kSdfChunks;
  if (chunk_bytes) *chunk_bytes = (int32_t)kChunkBytes;
  if (aux_floats) *aux_floats = kAuxB8 + 1;
  return 0;
}
