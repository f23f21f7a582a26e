// sm_100a primitives shared by every kernel in libsg_b200: mbarrier, cp.async, 1-D TMA bulk copy,
// tcgen05 (TMEM alloc / MMA / commit / ld) and UMMA descriptor construction.  Hand-written inline PTX;
// descriptor bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg {

typedef __nv_bfloat16 bf16;

constexpr int kTileRows = 128;         // rows of one operand tile == TMEM lanes
constexpr int kChunkK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kTileBytes = 16384;      // 128 rows x 128 B
constexpr float kLreluSlope = 0.2f;    // model/gan.py:11 (every LeakyReLU on the path uses 0.2)

// error codes written to the device-side error word (see sg_abi.cu)
constexpr int kErrMbarTimeout = 0x51;
constexpr int kErrSmemAlign = 0x52;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
// One lane of a CONVERGED warp.  Guarding the single-thread tcgen05 / TMA issue code with elect.sync (instead of lane == 0)
// tells ptxas the region runs on exactly one thread: descriptors stay in uniform registers and every UTCHMMA / UTMALDG is a
// straight-line instruction instead of an elect-and-branch loop over the active lanes (measured: 120 -> <64 cycles per MMA issue).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU box.  The clock is only consulted every 4096
// failed polls (each poll already suspends the thread for the hardware's try_wait window): nothing but the poll on the fast path.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
  uint64_t t0 = 0;
  for (uint32_t spins = 1; !mbar_try_wait(bar, parity); ++spins) {
    if ((spins & 0xfffu) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        if (err_flag) atomicExch(err_flag, kErrMbarTimeout);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- thread-block clusters (CTA pairs)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of THIS CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// arrive on an mbarrier that may live in another CTA of the cluster.  Default (cta-scope) release semantics on purpose: what the pair
// kernels hand over through these barriers is produced and consumed by the async proxies (cp.async / TMA writes fenced with
// fence.proxy.async, read by each CTA's own tensor core); only the SIGNAL crosses the CTA boundary.  The cluster-scope forms
// (.release.cluster / try_wait.acquire.cluster) compile to MEMBAR + an L1 invalidate (CCTL.IVALL) per poll and halved the kernel's
// speed (profiles/r02c_ncu_prof_conv.txt: 36 % of all stall samples on CCTL.IVALL).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait on a barrier whose arrivals may come from the peer CTA (threads, TMA, tensor core); cta-scope acquire, see above
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* err_flag) {
  uint64_t t0 = 0;
  for (uint32_t spins = 1; !mbar_try_wait(bar, parity); ++spins) {
    if ((spins & 0xfffu) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        if (err_flag) atomicExch(err_flag, kErrMbarTimeout);
        __trap();
      }
    }
  }
}
// cta_group::1 commit that arrives on the barrier at the same offset in every CTA of `mask` (a stage is free only when BOTH CTAs that
// receive multicast tiles into it have finished reading it)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, unsigned short mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// ---- cta_group::2 flavours of the tcgen05 primitives (a CTA pair = two SMs of one TPC acting on one 256-row accumulator)
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's own 128 rows] * B[each CTA's own N/2 rows]; issued by ONE thread of the LEADER CTA
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((unsigned short)3)
               : "memory");
}

// ---------------------------------------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- cp.async (LDGSTS)
// 16-byte global->shared copy; src_bytes == 0 zero-fills the destination (used for padding / ragged tails).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// wait until at most `n` of this thread's committed cp.async groups are still in flight (n is a runtime value)
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------------- 1-D TMA bulk copy (UBLKCP)
// global -> shared, completion reported as transaction bytes on an mbarrier.  size % 16 == 0, 16 B aligned.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (warp%4)*32+t, columns [c, c+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same lane/column mapping as tmem_ld32 (used to pre-load accumulators with bias terms)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
//   K-major  tile [rows][64 bf16]: 8-row groups SBO = 1024 B apart; LBO unused (encoded 1); K step of 16 = +32 B.
//   MN-major tile [k rows][64 bf16 of M/N]: 8-k-row groups SBO = 1024 B apart, next 64-wide M/N atom LBO bytes away;
//            K step of 16 rows = +2048 B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 with BF16 operands and FP32 accumulation.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                       // D format: F32
         | (1u << 7)                     // A format: BF16
         | (1u << 10)                    // B format: BF16
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(n >> 3) << 17)
         | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued MMAs of this thread arrive (once) on the mbarrier when they retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------- 256-bit global access (sm_100+)
// one full 32-byte sector per thread and instruction: row-per-thread streams stay sector-exact without relying on L1
struct u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ldg_nc_256(const void* p) {
  u32x8 r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w), "l"(p)
               : "memory");
}

// pull `bytes` (multiple of 16) of global memory into L2 ahead of the loads that will need them
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// byte offset of 16-byte piece `piece` of row `row` inside a 128B-swizzled [rows][128 B] tile (tile base 1024 B aligned)
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t piece) { return row * 128u + ((piece ^ (row & 7u)) << 4); }

// ---------------------------------------------------------------------------------------------- bf16 helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16lo_to_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// {lo, hi} -> bf16x2 with ReLU folded into the conversion (one instruction per two elements)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// Conv3d(1->C,k4,s2,p1) im2col for EIGHT consecutive output-x voxels (same n, od, oh) of piece g = taps (kd = g>>1,
// kh = 2(g&1)+{0,1}, kw = 0..3): two input lines, each fetched with 10 aligned float2 loads, feed all 8 rows x 8 taps.
// Needs W % 16 == 0 (so 8 outputs never straddle a line and float2 pairs never straddle the border).
__device__ __forceinline__ void patch_load8(const float* vol_n, int D, int H, int W, int od, int oh, int ow0, int g, bool valid, float (&f)[2][20]) {
  const int kd = g >> 1, kh0 = (g & 1) * 2;
  const int d = 2 * od - 1 + kd;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int h = 2 * oh - 1 + kh0 + hh;
    const bool lv = valid && d >= 0 && d < D && h >= 0 && h < H;
    const float2* line = reinterpret_cast<const float2*>(vol_n + ((size_t)(lv ? d : 0) * H + (lv ? h : 0)) * W) + (ow0 - 1);
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const int w = 2 * ow0 - 2 + 2 * j;
      float2 v = make_float2(0.f, 0.f);
      if (lv && w >= 0 && w < W) v = __ldg(line + j);
      f[hh][2 * j] = v.x; f[hh][2 * j + 1] = v.y;
    }
  }
}
__device__ __forceinline__ void patch_store8(const float (&f)[2][20], int g, uint8_t* tile_hi, uint8_t* tile_lo, int row0);
__device__ __forceinline__ void patch_fill8(const float* vol_n, int D, int H, int W, int od, int oh, int ow0, int g, bool valid,
                                            uint8_t* tile_hi, uint8_t* tile_lo, int row0) {
  float f[2][20];
  patch_load8(vol_n, D, H, W, od, oh, ow0, g, valid, f);
  patch_store8(f, g, tile_hi, tile_lo, row0);
}
__device__ __forceinline__ void patch_store8(const float (&f)[2][20], int g, uint8_t* tile_hi, uint8_t* tile_lo, int row0) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a0 = f[0][2 * i + 1], a1 = f[0][2 * i + 2], a2 = f[0][2 * i + 3], a3 = f[0][2 * i + 4];
    const float b0 = f[1][2 * i + 1], b1 = f[1][2 * i + 2], b2 = f[1][2 * i + 3], b3 = f[1][2 * i + 4];
    uint4 hi;
    hi.x = pack_bf16x2(a0, a1); hi.y = pack_bf16x2(a2, a3); hi.z = pack_bf16x2(b0, b1); hi.w = pack_bf16x2(b2, b3);
    const uint32_t off = sw128((uint32_t)(row0 + i), (uint32_t)g);
    *reinterpret_cast<uint4*>(tile_hi + off) = hi;
    if (tile_lo) {
      uint4 lo;
      lo.x = pack_bf16x2(a0 - bf16lo_to_f(hi.x), a1 - bf16hi_to_f(hi.x));
      lo.y = pack_bf16x2(a2 - bf16lo_to_f(hi.y), a3 - bf16hi_to_f(hi.y));
      lo.z = pack_bf16x2(b0 - bf16lo_to_f(hi.z), b1 - bf16hi_to_f(hi.z));
      lo.w = pack_bf16x2(b2 - bf16lo_to_f(hi.w), b3 - bf16hi_to_f(hi.w));
      *reinterpret_cast<uint4*>(tile_lo + off) = lo;
    }
  }
}

// activations used on the path
enum Act { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };
// `act` is uniform over a launch.  The piecewise-linear activations (everything on the training path except the generator's tanh and
// the GAN discriminator's sigmoid) are written as selects on loop-invariant predicates, not as a switch: inside the unrolled 8-element
// loops of the HBM-bound kernels and the GEMM epilogues a switch cost ~19 instructions per element (profiles/r02e_ncu_patch.txt).
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act <= ACT_RELU) {
    const float neg = (act == ACT_LRELU) ? v * kLreluSlope : 0.f;
    return (act == ACT_NONE || v > 0.f) ? v : neg;
  }
  return act == ACT_TANH ? tanhf(v) : 1.f / (1.f + __expf(-v));
}
// derivative expressed through the stored OUTPUT y of the activation
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  if (act <= ACT_RELU) {
    const float neg = (act == ACT_LRELU) ? kLreluSlope : 0.f;
    return (act == ACT_NONE || y > 0.f) ? 1.f : neg;
  }
  return act == ACT_TANH ? 1.f - y * y : y * (1.f - y);
}

}  // namespace sg
