// C-ABI plumbing of libsg_b200: error reporting, device error word, SM count, launch counter, weight packing.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

static thread_local char g_err[512] = "";
static long long g_launches = 0;

int sg_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "unknown error");
  return code;
}
void sg_count_launch() { ++g_launches; }

__device__ int g_error_word = 0;

int* sg_error_word() {
  static int* ptr[64] = {};                    // the symbol has one instance per device
  const int d = PerDevice::cur();
  if (!ptr[d]) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_error_word) == cudaSuccess) ptr[d] = (int*)p;
  }
  return ptr[d];
}

// ---------------------------------------------------------------------------------------------- weight packing
// One thread per 16-byte piece (8 consecutive K elements of one B row).
__global__ void sg_pack_b_kernel(const sg_pack_b_args a) {
  const int kchunks = a.k_pad / 64;
  const long long pieces = (long long)a.classes * kchunks * a.n_pad * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pieces; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    long long t = i >> 3;
    const int n = (int)(t % a.n_pad); t /= a.n_pad;
    const int kc = (int)(t % kchunks);
    const int cls = (int)(t / kchunks);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 64 + j * 8 + e;
      const int tap = k / a.c_count, c = k - tap * a.c_count;
      float x = 0.f;
      if (n < a.n_valid && tap < a.taps && c < a.c_valid) {
        int st = tap;
        if (a.classes == 8) {
          // ConvTranspose3d(k4,s2,p1) output parity p uses kernel taps {1,3} (p=0) or {0,2} (p=1): kk = p ? 2t : 1+2t
          const int td = (tap >> 2) & 1, th = (tap >> 1) & 1, tw = tap & 1;
          const int kd = ((cls >> 2) & 1) ? 2 * td : 1 + 2 * td;
          const int kh = ((cls >> 1) & 1) ? 2 * th : 1 + 2 * th;
          const int kw = (cls & 1) ? 2 * tw : 1 + 2 * tw;
          st = kd * 16 + kh * 4 + kw;
        }
        const int n1 = n / a.n0_count, n0 = n - n1 * a.n0_count;
        x = a.w[(long long)n1 * a.s_n1 + (long long)n0 * a.s_n0 + (long long)st * a.s_tap + (long long)c * a.s_c];
      }
      v[e] = x;
    }
    uint4 hi;
    hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
    hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
    uint8_t* base = reinterpret_cast<uint8_t*>(a.image);
    const size_t blk = (size_t)a.n_pad * 128;
    const size_t off = (((size_t)cls * kchunks + kc) * a.planes) * blk + (size_t)n * 128 + (size_t)((j ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(base + off) = hi;
    if (a.planes == 2) {
      uint4 lo;
      lo.x = pack_bf16x2(v[0] - bf16lo_to_f(hi.x), v[1] - bf16hi_to_f(hi.x));
      lo.y = pack_bf16x2(v[2] - bf16lo_to_f(hi.y), v[3] - bf16hi_to_f(hi.y));
      lo.z = pack_bf16x2(v[4] - bf16lo_to_f(hi.z), v[5] - bf16hi_to_f(hi.z));
      lo.w = pack_bf16x2(v[6] - bf16lo_to_f(hi.w), v[7] - bf16hi_to_f(hi.w));
      *reinterpret_cast<uint4*>(base + off + blk) = lo;
    }
  }
}

// Conv / ConvTranspose weights (the 64 kernel taps of a (n, c) pair are contiguous in the source: s_tap == 1), c_count % 64 == 0.
// One block per (n, 64-channel chunk): the [64 c][64 taps] fp32 tile is read with coalesced 16-byte loads into shared memory and the
// 512 16-byte pieces (8 consecutive c of one tap) are cut out of it.  The generic kernel above reads every element with its own
// strided 4-byte load behind a chain of integer divisions: ~10 us per weight inside the step graph, 12 re-packs per WGAN-GP step
// = 120 us of a 2.8 ms step (profiles/r02e: bq_base vs bq_stale).  Same bytes out, bit for bit.
__global__ void __launch_bounds__(256) sg_pack_b_taps_kernel(const sg_pack_b_args a) {
  __shared__ float tile[64][65];
  const int n = blockIdx.x, cb = blockIdx.y, t = threadIdx.x;
  {
    const int cl = t >> 2, q4 = t & 3, c = cb * 64 + cl;
    const bool ok = n < a.n_valid && c < a.c_valid;
    const float4* src = reinterpret_cast<const float4*>(a.w + (long long)n * a.s_n0 + (long long)c * a.s_c) + q4 * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float* d = &tile[cl][q4 * 16 + i * 4];
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  }
  __syncthreads();
  const int cchunks = a.c_count >> 6, kchunks = a.k_pad >> 6;
  uint8_t* base = reinterpret_cast<uint8_t*>(a.image);
  const size_t blk = (size_t)a.n_pad * 128;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int pc = t + 256 * h, j = pc & 7;
    int cls = 0, tap = pc >> 3, st = tap;
    if (a.classes == 8) {
      cls = pc >> 6; tap = (pc >> 3) & 7;
      const int td = (tap >> 2) & 1, th = (tap >> 1) & 1, tw = tap & 1;
      const int kd = ((cls >> 2) & 1) ? 2 * td : 1 + 2 * td;
      const int kh = ((cls >> 1) & 1) ? 2 * th : 1 + 2 * th;
      const int kw = (cls & 1) ? 2 * tw : 1 + 2 * tw;
      st = kd * 16 + kh * 4 + kw;
    }
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tile[j * 8 + e][st];
    uint4 hi;
    hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
    hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
    const int kc = tap * cchunks + cb;
    const size_t off = (((size_t)cls * kchunks + kc) * a.planes) * blk + (size_t)n * 128 + (size_t)((j ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(base + off) = hi;
    if (a.planes == 2) {
      uint4 lo;
      lo.x = pack_bf16x2(v[0] - bf16lo_to_f(hi.x), v[1] - bf16hi_to_f(hi.x));
      lo.y = pack_bf16x2(v[2] - bf16lo_to_f(hi.y), v[3] - bf16hi_to_f(hi.y));
      lo.z = pack_bf16x2(v[4] - bf16lo_to_f(hi.z), v[5] - bf16hi_to_f(hi.z));
      lo.w = pack_bf16x2(v[6] - bf16lo_to_f(hi.w), v[7] - bf16hi_to_f(hi.w));
      *reinterpret_cast<uint4*>(base + off + blk) = lo;
    }
  }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_abi_version(void) { return SG_ABI_VERSION; }
extern "C" const char* sg_last_error(void) { return g_err; }
extern "C" long long sg_launch_count(void) { return g_launches; }

extern "C" int sg_num_sms(void) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

// 0 when `stream` is not being captured into a CUDA graph, else the capture sequence's unique id: lets host-side caches tell
// "packed in THIS capture" (its pack kernel is a node of the graph and re-runs on every replay) from "packed earlier".
extern "C" int sg_stream_capture_id(void* stream, unsigned long long* id_out) {
  if (!id_out) return sg_fail(-1, "sg_stream_capture_id: null");
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  cudaError_t e = cudaStreamGetCaptureInfo((cudaStream_t)stream, &st, &id);
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  *id_out = (st == cudaStreamCaptureStatusActive) ? id : 0ull;
  return 0;
}

extern "C" int sg_device_error_word(int32_t** dev_ptr) {
  int* p = sg_error_word();
  if (!p) return sg_fail(-1, "sg_device_error_word: no device symbol (is a CUDA device present?)");
  if (dev_ptr) *dev_ptr = p;
  return 0;
}

// Diagnostic only (synchronises): reads and clears the watchdog word.  0 = healthy.
extern "C" int sg_check_device_error(void) {
  int* p = sg_error_word();
  if (!p) return sg_fail(-1, "sg_check_device_error: no device symbol");
  int v = 0, zero = 0;
  cudaError_t e = cudaMemcpy(&v, p, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  if (v != 0) cudaMemcpy(p, &zero, sizeof(int), cudaMemcpyHostToDevice);
  return v;
}

extern "C" size_t sg_pack_b_bytes(const sg_pack_b_args* a) {
  if (!a) return 0;
  return (size_t)a->classes * (a->k_pad / 64) * a->planes * a->n_pad * 128;
}

extern "C" int sg_pack_b(const sg_pack_b_args* a, void* stream) {
  if (!a || !a->w || !a->image) return sg_fail(-1, "sg_pack_b: null");
  if (a->planes != 1 && a->planes != 2) return sg_fail(-2, "sg_pack_b: planes");
  if (a->classes != 1 && a->classes != 8) return sg_fail(-3, "sg_pack_b: classes must be 1 or 8");
  if (a->k_pad <= 0 || (a->k_pad & 63) || a->n_pad <= 0 || (a->n_pad & 15)) return sg_fail(-4, "sg_pack_b: padding");
  if (a->c_count <= 0 || a->n0_count <= 0 || a->taps <= 0) return sg_fail(-5, "sg_pack_b: counts");
  {
    const char* np = getenv("SG_B200_NO_FAST_PACK");
    const bool taps_ok = (a->classes == 1 && a->taps == 64) || (a->classes == 8 && a->taps == 8);
    if (taps_ok && a->s_tap == 1 && (a->c_count & 63) == 0 && a->k_pad == a->taps * a->c_count && a->n0_count >= a->n_pad &&
        (a->s_n0 & 3) == 0 && (a->s_c & 3) == 0 && ((uintptr_t)a->w & 15) == 0 && a->n_pad <= 65535 && !(np && np[0] == '1')) {
      sg_pack_b_taps_kernel<<<dim3((unsigned)a->n_pad, (unsigned)(a->c_count >> 6)), 256, 0, (cudaStream_t)stream>>>(*a);
      SG_CUDA_CHECK_LAUNCH();
      return 0;
    }
  }
  const long long pieces = (long long)a->classes * (a->k_pad / 64) * a->n_pad * 8;
  const int block = 256;
  const int grid = (int)std::min<long long>((pieces + block - 1) / block, 148 * 32);
  sg_pack_b_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
