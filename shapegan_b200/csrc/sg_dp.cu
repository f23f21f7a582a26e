// Data-parallel optimizer step as ONE kernel over NVLink peer memory (SURVEY §8 e): the gradient all-reduce, the fused RMSprop / Adam
// update (+ 1/world scaling, + the critic's weight clip) and the parameter broadcast of a flat arena, without NCCL and without a
// second pass over the gradients.
//
// Every rank owns the contiguous shard [rank * chunk, (rank + 1) * chunk) of the arena AND the optimizer state of that shard only:
//   in-barrier   rank r tells every peer "my gradients are complete and nobody on my GPU reads my parameters any more"
//   reduce       for its shard: g = sum over peers of peer_grad[p][i]  (float4 loads straight out of the peers' HBM over NVLink /
//                NVSwitch; a reduce-scatter with no staging buffer), fixed summation order = bit-identical parameters on all ranks
//   update       RMSprop / Adam on the shard (torch.optim defaults, train_wgan.py:45-46, train_gan.py:28-31), state stays local
//   broadcast    the new parameter values are stored into EVERY peer's parameter arena (an all-gather by peer stores)
//   out-barrier  rank r leaves the kernel only when every peer has finished reading r's gradients and writing r's parameters
// One persistent CTA per SM; CTA 0 runs the cross-GPU barriers (system-scope release/acquire on small signal pads in symmetric
// memory), the other CTAs wait on a local flag.  All spins are bounded: a lost peer sets the error word and traps instead of hanging.
#include <algorithm>

#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

constexpr int kErrDpTimeout = 0x61;
constexpr int kDpMaxWorld = 16;

struct DpP {
  float* peer_grad[kDpMaxWorld];
  float* peer_param[kDpMaxWorld];
  uint32_t* peer_pad[kDpMaxWorld];     // each: uint32 [2][world] (in-barrier, out-barrier), indexed by SOURCE rank
  int rank, world;
  long long n, chunk;
  float* s1; float* s2;                // optimizer state of the local shard only
  int kind;                            // 0 RMSprop, 1 Adam
  float lr, a, b2, eps, clip, grad_scale, bc1, bc2s;
  uint32_t* sync;                      // local: [0] epoch, [1] release flag, [2] arrival counter
  int* err;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {       // peer HBM: never through the (incoherent) L1
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

template <class F>
__device__ __forceinline__ void dp_spin(F done, int* err) {
  uint64_t t0 = 0;
  for (uint32_t spins = 1; !done(); ++spins) {
    if ((spins & 0x3ffu) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ull) { if (err) atomicExch(err, kErrDpTimeout); __trap(); }
    }
  }
}

// cross-GPU barrier `which` (0 in, 1 out) of epoch e, executed by warp 0 of CTA 0: lane p talks to peer p
__device__ __forceinline__ void dp_peer_barrier(const DpP& p, int which, uint32_t e) {
  const int lane = threadIdx.x & 31;
  if (lane < p.world) {
    __threadfence_system();
    st_release_sys(p.peer_pad[lane] + which * p.world + p.rank, e);
    const uint32_t* mine = p.peer_pad[p.rank] + which * p.world + lane;
    dp_spin([&] { return ld_acquire_sys(mine) >= e; }, p.err);
  }
  __syncwarp();
}

__device__ __forceinline__ float dp_update(const DpP& p, float w, float g, float& s1, float& s2) {
  g *= p.grad_scale;
  if (p.kind == 0) {                       // torch.optim.RMSprop defaults: alpha 0.99, eps 1e-8, no momentum, not centered
    s1 = p.a * s1 + (1.f - p.a) * g * g;
    w -= p.lr * g / (sqrtf(s1) + p.eps);
    if (p.clip > 0.f) w = fminf(fmaxf(w, -p.clip), p.clip);
  } else {                                 // torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8
    s1 = p.a * s1 + (1.f - p.a) * g;
    s2 = p.b2 * s2 + (1.f - p.b2) * g * g;
    w -= (p.lr / p.bc1) * s1 / (sqrtf(s2) / p.bc2s + p.eps);
  }
  return w;
}

__global__ void __launch_bounds__(512, 1) sg_dp_step_kernel(const DpP p) {
  const uint32_t e = p.sync[0] + 1;                        // epoch of this launch (bumped at the end by CTA 0)
  // ---- in-barrier
  if (blockIdx.x == 0) {
    if (threadIdx.x < 32) {
      dp_peer_barrier(p, 0, e);
      if (threadIdx.x == 0) { __threadfence(); atomicExch(&p.sync[1], e); }
    }
  }
  if (threadIdx.x == 0) dp_spin([&] { return ld_acquire_gpu(p.sync + 1) >= e; }, p.err);
  __syncthreads();
  // ---- reduce-scatter + update + all-gather of this rank's shard
  const long long lo = (long long)p.rank * p.chunk, hi = min(p.n, lo + p.chunk);
  const long long n4 = hi > lo ? (hi - lo) / 4 : 0;       // chunk % 4 == 0; the arena's tail (n % 4) belongs to the last rank
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long off = lo + i * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < p.world; ++r) {                    // fixed order: every rank would compute the same sum
      const float4 x = ld_peer_f4(p.peer_grad[r] + off);
      g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
    }
    float4 w = *reinterpret_cast<const float4*>(p.peer_param[p.rank] + off);
    float4 a = *reinterpret_cast<float4*>(p.s1 + i * 4);
    float4 b = p.kind ? *reinterpret_cast<float4*>(p.s2 + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    w.x = dp_update(p, w.x, g.x, a.x, b.x); w.y = dp_update(p, w.y, g.y, a.y, b.y);
    w.z = dp_update(p, w.z, g.z, a.z, b.z); w.w = dp_update(p, w.w, g.w, a.w, b.w);
    *reinterpret_cast<float4*>(p.s1 + i * 4) = a;
    if (p.kind) *reinterpret_cast<float4*>(p.s2 + i * 4) = b;
    for (int r = 0; r < p.world; ++r) *reinterpret_cast<float4*>(p.peer_param[r] + off) = w;
  }
  // scalar tail of the last shard
  if (blockIdx.x == 0 && threadIdx.x < (int)((hi - lo) - n4 * 4)) {
    const long long off = lo + n4 * 4 + threadIdx.x, si = n4 * 4 + threadIdx.x;
    float g = 0.f;
    for (int r = 0; r < p.world; ++r) g += *reinterpret_cast<volatile float*>(p.peer_grad[r] + off);
    float s1 = p.s1[si], s2 = p.kind ? p.s2[si] : 0.f;
    const float w = dp_update(p, p.peer_param[p.rank][off], g, s1, s2);
    p.s1[si] = s1;
    if (p.kind) p.s2[si] = s2;
    for (int r = 0; r < p.world; ++r) p.peer_param[r][off] = w;
  }
  // ---- out-barrier: every CTA's peer stores are out before CTA 0 tells the peers
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&p.sync[2], 1u);
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    if (threadIdx.x == 0) dp_spin([&] { return ld_acquire_gpu(p.sync + 2) >= e * gridDim.x; }, p.err);
    __syncwarp();
    dp_peer_barrier(p, 1, e);
    if (threadIdx.x == 0) { __threadfence(); p.sync[0] = e; }
  }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_dp_step(const sg_dp_step_args* a, void* stream) {
  if (!a || a->world < 2 || a->world > kDpMaxWorld || a->rank < 0 || a->rank >= a->world) return sg_fail(-1, "sg_dp_step: bad rank / world (2..16)");
  if (!a->peer_grad || !a->peer_param || !a->peer_pad || !a->s1 || !a->sync || a->n <= 0) return sg_fail(-2, "sg_dp_step: null");
  if (a->kind != 0 && a->kind != 1) return sg_fail(-3, "sg_dp_step: kind must be 0 (rmsprop) or 1 (adam)");
  if (a->kind == 1 && !a->s2) return sg_fail(-2, "sg_dp_step: adam needs s2");
  if (a->chunk <= 0 || (a->chunk & 3) || a->chunk * a->world < a->n) return sg_fail(-4, "sg_dp_step: chunk must be a multiple of 4 covering n");
  DpP p;
  for (int r = 0; r < a->world; ++r) {
    p.peer_grad[r] = (float*)a->peer_grad[r]; p.peer_param[r] = (float*)a->peer_param[r]; p.peer_pad[r] = (uint32_t*)a->peer_pad[r];
    if (!p.peer_grad[r] || !p.peer_param[r] || !p.peer_pad[r]) return sg_fail(-2, "sg_dp_step: null peer pointer");
    if (((uintptr_t)p.peer_grad[r] | (uintptr_t)p.peer_param[r]) & 15) return sg_fail(-5, "sg_dp_step: arenas must be 16-byte aligned");
  }
  p.rank = a->rank; p.world = a->world; p.n = a->n; p.chunk = a->chunk; p.s1 = a->s1; p.s2 = a->s2; p.kind = a->kind;
  p.lr = a->lr; p.eps = a->eps; p.clip = a->clip; p.grad_scale = a->grad_scale;
  if (a->kind == 0) { p.a = a->beta1; p.b2 = 0.f; p.bc1 = 1.f; p.bc2s = 1.f; }
  else { p.a = a->beta1; p.b2 = a->beta2; p.bc1 = 1.f - powf(a->beta1, (float)a->step); p.bc2s = sqrtf(1.f - powf(a->beta2, (float)a->step)); }
  p.sync = (uint32_t*)a->sync; p.err = sg_error_word();
  sg_dp_step_kernel<<<sg_num_sms(), 512, 0, (cudaStream_t)stream>>>(p);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
