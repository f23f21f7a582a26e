// Implicit-GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM) for every dense contraction of
// the shapegan hot path: Conv3d / ConvTranspose3d (k4,s2,p1) forward + input gradient, nn.Linear, and the
// weight-gradient GEMMs.  One persistent, warp-specialised CTA per SM:
//
//   warps 0-3  producers   A operand: 16-byte cp.async gathers from NDHWC bf16 planes straight into the
//                          128B-swizzled K-major (or MN-major for wgrad) UMMA layout, zero-filled padding;
//                          B operand (pre-packed weights): 1-D TMA bulk copies (cp.async.bulk) on an mbarrier.
//   warp  4    MMA issuer  one thread issues tcgen05.mma (M=128, N=bn, K=16) and tcgen05.commit.
//   warps 5-8  epilogue    tcgen05.ld TMEM -> registers -> bias / activation / mask -> bf16 planes or fp32.
//
// fp32x mode (planes=2) feeds hi/lo bf16 splits and issues three MMAs per K step (hi*hi + hi*lo + lo*hi).
#include <algorithm>
#include <cstring>

#include <cstdlib>

#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_tma.h"

namespace sg {

constexpr int kProducerThreads = 128;
constexpr int kBLoaderThreads = 96;    // TMA mode: warps 1-3 stream the weight tile with cp.async (keeps it off the TMA unit)
constexpr int kIgemmThreads = 288;      // 4 producer warps + 1 MMA warp + 4 epilogue warps
constexpr int kMaxStages = 6;
constexpr int kSmemHeader = 2048;       // barriers + tmem pointer + staged bias
constexpr int kTraceLen = 1024;
__device__ long long g_igemm_trace[3 * kTraceLen];   // SG_B200_IGEMM_DIAG & 128 in a --trace build: clock64 stamps of CTA 0 (see tr())

// halo kernels, SG_B200_IGEMM_DIAG == 128: 8 channels x 128 slots of clock64 from CTA 0 (see tools/trace_igemm.py)
constexpr int kTr = 128;
// compiled in only with -DSG_IGEMM_TRACE (python -m shapegan_b200.build --trace): the stamps sit on the single-thread issue loops
__device__ __forceinline__ void tr(bool on, int ch, int i) {
#ifdef SG_IGEMM_TRACE
  if (on && i < kTr) g_igemm_trace[ch * kTr + i] = clock64();
#endif
}

struct IgemmP {
  int mode, planes;
  const char* a_ptr; long long a_ps; int aN, aD, aH, aW, aC;
  const char* a2_ptr; long long a2_ps; int a2C;
  long long rows; int kchunks, n_pad, n_valid, bn, mt, ksplit, classes;
  const char* b;
  const float* bias; int act, bias_mod;
  const bf16* mask; long long mask_ps; int mask_act;
  char* out; long long out_ps; int out_kind, out_ld, oD, oH, oW;
  long long ks_stride;   // elements between the fp32 partial slabs of consecutive K splits (0: atomics / no split)
  int stages, m_tiles, n_tiles; long long work_total;
  int diag;              // measurement only (SG_B200_IGEMM_DIAG): 1 skip A loads, 2 skip B loads, 4 skip MMAs, 8 skip the epilogue stores,
                         // 16 skip the whole K loop and epilogue (launch + prologue + teardown only)
  int acc_bufs, acc_slot;
  unsigned ktab_bytes, stage_bytes, a_stage_bytes;
  int* err;
  // TMA gather (C % 64 == 0, power-of-two grids): a tile of 128 rows is the box (bx,by,bz,bn) of the row grid (gx,gy,gz)
  int use_tma, bx, by, bz, bnn, gx, gy, gz;
  CUtensorMap tmA[2];    // per plane
  CUtensorMap tmA2[2];   // second DENSE source
  // halo variant (sg_igemm_halo_kernel): one strided box of (8+1) x 8 x (2 mt + 1) grid points serves the 4 taps (qz, qx)
  int halo; unsigned blk_bytes; int b_stages;
  int a_blocks;          // halo blocks in flight (2 or 3)
  int b_tma;             // halo kernels: weight tiles by the TMA unit (1-D bulk copy; pair: 2-D box reported to the leader) instead of cp.async
  int b_depth;           // cp.async weight tiles in flight per loader thread
  CUtensorMap tmB;       // pair kernel, b_tma: the packed weight image as [rows][64] bf16, unswizzled (the image is pre-swizzled)
  unsigned epi_off;      // halo kernels: byte offset of the epilogue's staging tile (128 rows x bn bf16) in shared memory, 0 = none
  int pair;              // sg_igemm_halo2_kernel: CTA pairs, tcgen05.mma.cta_group::2 (each CTA: its own 128-row tiles + half of every weight tile)
  CUtensorMap tmH;
};

struct SmemHeader {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t accfull[2];
  uint64_t accempty[2];
  uint32_t tmem_base;
  uint32_t pad_[3];
  float sbias[256];      // bias of the current N tile (epilogue broadcast reads)
  uint64_t blk_full[3];  // halo kernels: A blocks (2 or 3 of them)
  uint64_t blk_empty[3];
};

template <class P>
__device__ __forceinline__ void decode_work(const P& p, long long w64, int& cls, int& nt, int& mtile, int& ks) {
  uint32_t w = (uint32_t)w64;              // work_total < 2^31 (checked on the host): 32-bit division only
  if (p.ksplit > 1) { ks = (int)(w % (uint32_t)p.ksplit); w /= (uint32_t)p.ksplit; } else ks = 0;
  mtile = (int)(w % (uint32_t)p.m_tiles); w /= (uint32_t)p.m_tiles;
  if (p.n_tiles > 1) { nt = (int)(w % (uint32_t)p.n_tiles); cls = (int)(w / (uint32_t)p.n_tiles); } else { nt = 0; cls = (int)w; }
}

// v[0..8) *= act'(mask) for 8 packed bf16 mask values (the stored OUTPUT of the consumer's activation).  The piecewise-linear
// activations are decided on the bf16 bit patterns (y > 0 <=> the pattern is a positive integer); the generic switch cost ~19
// instructions per element and made every mask-fused launch 2.5-3x slower than its unmasked twin (profiles/r02e_ncu_patch.txt).
__device__ __forceinline__ void mask_mul8(float* v, const uint4& m, int mask_act) {
  const uint32_t w[4] = {m.x, m.y, m.z, m.w};
  if (mask_act == ACT_LRELU || mask_act == ACT_RELU) {
    const float neg = (mask_act == ACT_LRELU) ? kLreluSlope : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] *= ((int)(w[i] << 16) > 0) ? 1.f : neg;
      v[2 * i + 1] *= ((int)(w[i] & 0xffff0000u) > 0) ? 1.f : neg;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] *= act_grad_from_output(bf16lo_to_f(w[i]), mask_act);
      v[2 * i + 1] *= act_grad_from_output(bf16hi_to_f(w[i]), mask_act);
    }
  }
}

// this warp's 32 rows x 8 chunks (bn = 64) of the mask tile, coalesced: lane i + 32 k reads chunk (i & 7) of row (i >> 3) + 4 k
__device__ __forceinline__ void mask_tile_load8(const bf16* mask, long long my_ob, int lane, uint4 (&m)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = lane + 32 * k, rr = i >> 3, ch = i & 7;
    const long long ob = __shfl_sync(0xffffffffu, my_ob, rr);
    m[k] = make_uint4(0u, 0u, 0u, 0u);
    if (ob >= 0) m[k] = *reinterpret_cast<const uint4*>(mask + ob + ch * 8);      // plain load: `mask` may alias `out`
  }
}

template <int ACT>
__device__ __forceinline__ float act_t(float v, int runtime_act) {
  if (ACT == ACT_NONE) return v;
  if (ACT == ACT_LRELU) return v > 0.f ? v : v * kLreluSlope;
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  return apply_act(v, runtime_act);
}

// Epilogue warps: TMEM -> registers -> (+bias from smem) -> activation -> bf16 planes / fp32.
// Fast path (whole N tile valid, 16-byte aligned rows, no mask): ~3 instructions per element, straight-line.
struct EpiP {   // the fields the epilogue needs, copied ONCE into registers: the kernel parameter is reached through a
                // reference here, and every p.field access would otherwise be a generic load with a long-scoreboard stall
  int mode, planes, n_valid, bn, mt, ksplit, m_tiles, n_tiles, bias_mod, act, mask_act, out_kind, out_ld, oD, oH, oW, aD, aH, aW, acc_bufs, acc_slot;
  long long rows, work_total, out_ps, ks_stride;
  int pair;               // CTA-pair kernel: work items are pairs of M tiles, this CTA owns tile 2*mtile + cluster rank
  unsigned epi_off;
  const float* bias; const bf16* mask; char* out; int* err;
};

// 32 accumulator columns of this lane's row -> (+ bias) -> activation -> (x act'(mask), read from the staging slot) -> bf16 -> staging slot
template <int ACT, bool MASKED>
__device__ __forceinline__ void staged_convert32(const uint32_t (&r)[32], const float* sb, bool add_bias, int act, int mask_act, uint32_t srow, int cb,
                                                 int lane) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = add_bias ? *reinterpret_cast<const float4*>(sb + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[j] = act_t<ACT>(__uint_as_float(r[j]) + b.x, act);
    v[j + 1] = act_t<ACT>(__uint_as_float(r[j + 1]) + b.y, act);
    v[j + 2] = act_t<ACT>(__uint_as_float(r[j + 2]) + b.z, act);
    v[j + 3] = act_t<ACT>(__uint_as_float(r[j + 3]) + b.w, act);
  }
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const uint32_t slot = srow + (uint32_t)((((cb + j) >> 3) ^ (lane & 7)) << 4);
    if (MASKED) {
      const uint4 m = ld_shared_v4(slot);
      mask_mul8(v + j, m, mask_act);
    }
    uint4 hi;
    hi.x = pack_bf16x2(v[j], v[j + 1]); hi.y = pack_bf16x2(v[j + 2], v[j + 3]);
    hi.z = pack_bf16x2(v[j + 4], v[j + 5]); hi.w = pack_bf16x2(v[j + 6], v[j + 7]);
    st_shared_v4(slot, hi);
  }
}

template <int ACT>
__device__ __noinline__ void epilogue_role(const IgemmP& gp, SmemHeader* hdr, uint32_t tmem_base, int cps) {
  EpiP p;
  p.mode = gp.mode; p.planes = gp.planes; p.n_valid = gp.n_valid; p.bn = gp.bn; p.mt = gp.mt; p.ksplit = gp.ksplit;
  p.m_tiles = gp.m_tiles; p.n_tiles = gp.n_tiles; p.bias_mod = gp.bias_mod; p.act = gp.act; p.mask_act = gp.mask_act;
  p.out_kind = gp.out_kind; p.out_ld = gp.out_ld; p.oD = gp.oD; p.oH = gp.oH; p.oW = gp.oW; p.aD = gp.aD; p.aH = gp.aH; p.aW = gp.aW;
  p.pair = gp.pair; p.epi_off = gp.epi_off; p.acc_bufs = gp.acc_bufs; p.acc_slot = gp.acc_slot; p.rows = gp.rows; p.work_total = gp.work_total; p.out_ps = gp.out_ps; p.ks_stride = gp.ks_stride;
  p.bias = gp.bias; p.mask = gp.mask; p.out = gp.out; p.err = gp.err;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;                 // TMEM lane quadrant this warp may access
  const int trow = q * 32 + lane;
  const int etid = threadIdx.x - 5 * 32;  // 0..127 among the epilogue threads
  float* sbias = hdr->sbias;
  int it = 0, staged_nt = -1;
  const uint32_t crank = p.pair ? cluster_ctarank() : 0u;
  const long long w0 = p.pair ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long wstep = p.pair ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  // pair kernel: the accumulator-drained signal goes to the LEADER's barrier (its MMA thread feeds both CTAs' accumulators)
  const uint32_t accempty_addr[2] = {p.pair ? mapa_u32(smem_u32(&hdr->accempty[0]), 0) : 0u, p.pair ? mapa_u32(smem_u32(&hdr->accempty[1]), 0) : 0u};
  for (long long w = w0; w < p.work_total; w += wstep, ++it) {
    int cls, nt, mtile, ks;
    decode_work(p, w, cls, nt, mtile, ks);
    if (p.pair) mtile = mtile * 2 + (int)crank;
    const int ab = (p.acc_bufs == 2) ? (it & 1) : 0;
    const uint32_t aph = (uint32_t)((it / p.acc_bufs) & 1);
    const bool add_bias = (p.bias != nullptr) && (ks == 0);
    if (staged_nt != nt) {               // bias of this N tile -> shared memory (broadcast reads afterwards)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int j = etid; j < p.bn; j += 128) {
        const int n = nt * p.bn + j;
        sbias[j] = (p.bias != nullptr && n < p.n_valid) ? __ldg(p.bias + (p.bias_mod > 0 ? n % p.bias_mod : n)) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      staged_nt = nt;
    }
    const bool tile_full = (nt * p.bn + p.bn <= p.n_valid);
    const bool fast = tile_full && (p.bn & 31) == 0 &&
                      ((p.out_kind == SG_OUT_BF16 && (p.out_ld & 7) == 0) || (p.out_kind == SG_OUT_F32 && (p.out_ld & 3) == 0 && p.mask == nullptr));
    const bool staged = fast && p.epi_off != 0u && p.out_kind == SG_OUT_BF16 && p.planes == 1 && (p.bn == 64 || p.bn == 128);
    // this lane's output row (element offset of its first column, -1 = beyond the last row) in each sub-tile
    long long ob_sub[2] = {-1, -1};
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      if (sub < p.mt) {
        const long long gr = ((long long)mtile * p.mt + sub) * kTileRows + trow;
        if (gr < p.rows) {
          long long orow = gr;
          if (p.mode == SG_MODE_CONVT) {
            int qw = (int)(gr % p.aW); long long t = gr / p.aW;
            int qh = (int)(t % p.aH); t /= p.aH;
            int qd = (int)(t % p.aD); long long n = t / p.aD;
            orow = ((n * p.oD + 2 * qd + ((cls >> 2) & 1)) * p.oH + 2 * qh + ((cls >> 1) & 1)) * p.oW + 2 * qw + (cls & 1);
          }
          ob_sub[sub] = orow * p.out_ld + (long long)nt * p.bn + (long long)ks * p.ks_stride;   // ks_stride: split-K partial slabs
        }
      }
    }
    // bn = 64 (8 chunks per row): the mask tile of a sub-tile is fetched one phase ahead -- sub-tile 0's while this warp still waits for
    // the accumulator, sub-tile 1's while sub-tile 0 is converted -- so the ~2 k clock round trip to HBM (the mask is a 100 MB activation
    // of the batched critic) is never exposed.  (bn = 128 would need 64 registers for the tile in flight: loaded in phase, below.)
    const bool mpref = staged && p.mask != nullptr && p.bn == 64;
    uint4 mpre[8];
    if (mpref) mask_tile_load8(p.mask, ob_sub[0], lane, mpre);
    if (p.pair) mbar_wait_cluster(&hdr->accfull[ab], aph, p.err); else mbar_wait(&hdr->accfull[ab], aph, p.err);
    tc_fence_after();
    if (etid == 0) tr((gp.diag & 128) && blockIdx.x == 0, 6, 2 * it);
    for (int sub = 0; sub < p.mt; ++sub) {
      const bool valid = ob_sub[sub] >= 0;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ab * p.mt + sub) * p.acc_slot);
      const long long obase = ob_sub[sub];
      if (staged) {
        // ---- staged stores.  A thread owns an accumulator ROW (TMEM lane), so direct stores put 16 bytes per lane at a row-pitch stride:
        // 32 half-written sectors per instruction; the tile store took 14.7 k clocks (profiles/r02e_trace_halo.txt).  Instead each warp
        // parks its 32 rows in shared memory (16-byte chunks XOR-swizzled by row: conflict-free both ways) and writes them out
        // cooperatively, whole rows per instruction.  A mask tile travels the other way first, so its loads are coalesced too.
        const uint32_t pitch = (uint32_t)p.bn * 2u;
        const int cpr = p.bn >> 3, lgc = (cpr == 16) ? 4 : 3;           // 16-byte chunks per row (bn = 64 or 128)
        const uint32_t stg = smem_u32(reinterpret_cast<uint8_t*>(hdr) + p.epi_off) + (uint32_t)q * 32u * pitch;
        const uint32_t srow = stg + (uint32_t)lane * pitch;
        const long long my_ob = valid ? obase : -1;
        __syncwarp();
        if (mpref) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = lane + 32 * k, rr = i >> 3, ch = i & 7;
            st_shared_v4(stg + (uint32_t)rr * pitch + (uint32_t)((ch ^ (rr & 7)) << 4), mpre[k]);
          }
          if (sub + 1 < p.mt) mask_tile_load8(p.mask, ob_sub[sub + 1], lane, mpre);
          __syncwarp();
        } else if (p.mask != nullptr) {
          // all of this warp's mask loads are issued before the first use (a rolled load -> store loop serialises the round trips:
          // 8 x ~2 k clocks per sub-tile made the masked D1 launch 3x slower than the unmasked one, profiles/r02e_trace_patch.txt)
          for (int k0 = 0; k0 < cpr; k0 += 8) {            // eight loads in flight per round (bn = 128: two rounds)
            uint4 mreg[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int i = lane + 32 * (k0 + k), rr = i >> lgc, ch = i & (cpr - 1);
              const long long ob = __shfl_sync(0xffffffffu, my_ob, rr);
              mreg[k] = make_uint4(0u, 0u, 0u, 0u);
              if (ob >= 0) mreg[k] = *reinterpret_cast<const uint4*>(p.mask + ob + ch * 8);      // plain load: `mask` may alias `out`
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int i = lane + 32 * (k0 + k), rr = i >> lgc, ch = i & (cpr - 1);
              st_shared_v4(stg + (uint32_t)rr * pitch + (uint32_t)((ch ^ (rr & 7)) << 4), mreg[k]);
            }
          }
          __syncwarp();
        }
        // accumulator -> bias / activation (/ mask) -> bf16 -> this lane's staging row.  Unmasked: two 32-column TMEM loads in flight per
        // wait; masked: one (the mask tile in flight already holds 32 registers)
        if (p.mask == nullptr) {
          for (int c0 = 0; c0 < p.bn; c0 += 64) {
            uint32_t r0[32], r1[32];
            __syncwarp();
            tmem_ld32(t_addr + c0, r0);
            tmem_ld32(t_addr + c0 + 32, r1);
            tmem_ld_wait();
            staged_convert32<ACT, false>(r0, sbias + c0, add_bias, p.act, 0, srow, c0, lane);
            staged_convert32<ACT, false>(r1, sbias + c0 + 32, add_bias, p.act, 0, srow, c0 + 32, lane);
          }
        } else {
          for (int c0 = 0; c0 < p.bn; c0 += 32) {
            uint32_t r0[32];
            __syncwarp();
            tmem_ld32(t_addr + c0, r0);
            tmem_ld_wait();
            staged_convert32<ACT, true>(r0, sbias + c0, add_bias, p.act, p.mask_act, srow, c0, lane);
          }
        }
        __syncwarp();
        if (!(gp.diag & 8)) {
          bf16* const outp = reinterpret_cast<bf16*>(p.out);
          for (int i = lane; i < 32 * cpr; i += 32) {
            const int rr = i >> lgc, ch = i & (cpr - 1);
            const long long ob = __shfl_sync(0xffffffffu, my_ob, rr);
            const uint4 val = ld_shared_v4(stg + (uint32_t)rr * pitch + (uint32_t)((ch ^ (rr & 7)) << 4));
            if (ob >= 0) *reinterpret_cast<uint4*>(outp + ob + ch * 8) = val;
          }
        }
        __syncwarp();
      } else if (fast) {
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(t_addr + c0, r);
          tmem_ld_wait();
          if (!valid || (gp.diag & 8)) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = add_bias ? *reinterpret_cast<const float4*>(sbias + c0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j] = act_t<ACT>(__uint_as_float(r[j]) + b.x, p.act);
            v[j + 1] = act_t<ACT>(__uint_as_float(r[j + 1]) + b.y, p.act);
            v[j + 2] = act_t<ACT>(__uint_as_float(r[j + 2]) + b.z, p.act);
            v[j + 3] = act_t<ACT>(__uint_as_float(r[j + 3]) + b.w, p.act);
          }
          if (p.out_kind == SG_OUT_BF16) {
            bf16* o = reinterpret_cast<bf16*>(p.out) + obase + c0;
            if (p.mask != nullptr) {       // out *= act'(mask): the activation backward of the consumer layer, fused
              const bf16* mk = p.mask + obase + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 m = *reinterpret_cast<const uint4*>(mk + j);     // plain load: `mask` may alias `out` (in-place activation backward)
                mask_mul8(v + j, m, p.mask_act);
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 hi;
              hi.x = pack_bf16x2(v[j], v[j + 1]); hi.y = pack_bf16x2(v[j + 2], v[j + 3]);
              hi.z = pack_bf16x2(v[j + 4], v[j + 5]); hi.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = hi;
              if (p.planes == 2) {
                uint4 lo;
                lo.x = pack_bf16x2(v[j] - bf16lo_to_f(hi.x), v[j + 1] - bf16hi_to_f(hi.x));
                lo.y = pack_bf16x2(v[j + 2] - bf16lo_to_f(hi.y), v[j + 3] - bf16hi_to_f(hi.y));
                lo.z = pack_bf16x2(v[j + 4] - bf16lo_to_f(hi.z), v[j + 5] - bf16hi_to_f(hi.z));
                lo.w = pack_bf16x2(v[j + 6] - bf16lo_to_f(hi.w), v[j + 7] - bf16hi_to_f(hi.w));
                *reinterpret_cast<uint4*>(o + p.out_ps + j) = lo;
              }
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + obase + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      } else {
        // generic path: ragged N, masks, atomics, unaligned rows
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          uint32_t r[32];
          __syncwarp();
          if (c0 + 32 <= p.bn) {
            tmem_ld32(t_addr + c0, r);
          } else {
            uint32_t r16[16];
            tmem_ld16(t_addr + c0, r16);
#pragma unroll
            for (int j = 0; j < 16; ++j) { r[j] = r16[j]; r[16 + j] = 0; }
          }
          tmem_ld_wait();
          if (!valid) continue;
          const int ncols = min(32, p.bn - c0);
          for (int j = 0; j < ncols; ++j) {
            const int n = nt * p.bn + c0 + j;
            if (n >= p.n_valid) break;
            float x = __uint_as_float(r[j]);
            if (add_bias) x += sbias[c0 + j];
            x = act_t<ACT>(x, p.act);
            const long long eoff = obase + c0 + j;
            if (p.mask) x *= act_grad_from_output(__bfloat162float(p.mask[eoff]), p.mask_act);
            if (p.out_kind == SG_OUT_BF16) {
              bf16* o = reinterpret_cast<bf16*>(p.out) + eoff;
              const bf16 h = __float2bfloat16_rn(x);
              *o = h;
              if (p.planes == 2) o[p.out_ps] = __float2bfloat16_rn(x - __bfloat162float(h));
            } else if (p.out_kind == SG_OUT_F32) {
              reinterpret_cast<float*>(p.out)[eoff] = x;
            } else {
              atomicAdd(reinterpret_cast<float*>(p.out) + eoff, x);
            }
          }
        }
      }
    }
    if (etid == 0) tr((gp.diag & 128) && blockIdx.x == 0, 6, 2 * it + 1);
    tc_fence_before();
    if (p.pair) {            // one arrival per CTA: the 128 epilogue threads meet first
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (etid == 0) mbar_arrive_cluster(accempty_addr[ab]);
    } else {
      mbar_arrive(&hdr->accempty[ab]);
    }
  }
}

__global__ void __launch_bounds__(kIgemmThreads, 1) sg_igemm_kernel(const __grid_constant__ IgemmP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  uint32_t* ktab = reinterpret_cast<uint32_t*>(smem + kSmemHeader);
  uint8_t* stage0 = smem + kSmemHeader + p.ktab_bytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int S = p.stages;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&hdr->full[s], (p.use_tma ? (p.b_tma ? 0 : kBLoaderThreads) : kProducerThreads) + 1); mbar_init(&hdr->empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&hdr->accfull[i], 1); mbar_init(&hdr->accempty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&hdr->tmem_base, 512);
  if (p.mode == SG_MODE_CONV || p.mode == SG_MODE_CONVT) {
    // ktab[k/8] = channel | tap offsets; K index k = tap*C + c (c fastest)
    const int n8 = p.kchunks * 8;
    for (int i = tid; i < n8; i += kIgemmThreads) {
      int k = i * 8, tap = k / p.aC, c = k - tap * p.aC;
      uint32_t dd, dh, dw;
      if (p.mode == SG_MODE_CONV) { dd = tap >> 4; dh = (tap >> 2) & 3; dw = tap & 3; }
      else { dd = tap >> 2; dh = (tap >> 1) & 1; dw = tap & 1; }
      ktab[i] = (uint32_t)c | (dd << 16) | (dh << 19) | (dw << 22);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  if ((smem_u32(smem) & 1023u) != 0) {       // swizzled operand tiles need a 1024-byte aligned base
    if (tid == 0) atomicExch(p.err, kErrSmemAlign);
    __trap();
  }

  const int cps = (p.kchunks + p.ksplit - 1) / p.ksplit;   // K chunks per split
  if (p.diag & 16) {
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
    return;
  }

  if (warp < 4 && p.use_tma) {
    // ================================================================ TMA PRODUCER: ONE elected thread runs the whole loop (waits
    // included): no per-stage elect / warp re-convergence on the issue path
    if (warp == 0 && elect_one()) {
      tma_prefetch_desc(&p.tmA[0]);
      if (p.planes == 2) tma_prefetch_desc(&p.tmA[1]);
      int s = 0; uint32_t ph = 0; int tcnt = 0;
      const int c1chunks = (p.aC + 63) >> 6;
      // power-of-two row grid: tile origin by shifts/masks, once per tile (no 64-bit division in the K loop)
      const int lgx = 31 - __clz(max(p.gx, 1)), lgy = 31 - __clz(max(p.gy, 1)), lgz = 31 - __clz(max(p.gz, 1));
      for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x) {
        int cls, nt, mtile, ks;
        decode_work(p, w, cls, nt, mtile, ks);
        const int pd = (cls >> 2) & 1, phh = (cls >> 1) & 1, pw = cls & 1;
        const int k0 = ks * cps, k1 = min(p.kchunks, k0 + cps);
        const int mul = (p.mode == SG_MODE_CONV) ? 2 : 1;
        int tx0[2], ty0[2], tz0[2], tn0[2], trow[2];
        for (int sub = 0; sub < p.mt; ++sub) {
          const uint32_t row0 = (uint32_t)((mtile * p.mt + sub) * kTileRows);
          trow[sub] = (int)row0;
          tx0[sub] = mul * (int)(row0 & (uint32_t)(p.gx - 1));
          ty0[sub] = mul * (int)((row0 >> lgx) & (uint32_t)(p.gy - 1));
          tz0[sub] = mul * (int)((row0 >> (lgx + lgy)) & (uint32_t)(p.gz - 1));
          tn0[sub] = (int)(row0 >> (lgx + lgy + lgz));
        }
        int tap = 0, c0 = 0;
        if (p.mode != SG_MODE_DENSE) { tap = (k0 * 64) / p.aC; c0 = k0 * 64 - tap * p.aC; }
        for (int kc = k0; kc < k1; ++kc) {
          mbar_wait(&hdr->empty[s], ph ^ 1, p.err);
          tr((p.diag & 128) && blockIdx.x == 0, 0, tcnt++);
          uint64_t* bar = &hdr->full[s];
          const uint32_t a_base = smem_u32(stage0 + (size_t)s * p.stage_bytes);
          int ox = 0, oy = 0, oz = 0;
          if (p.mode == SG_MODE_CONV) { oz = (tap >> 4) - 1; oy = ((tap >> 2) & 3) - 1; ox = (tap & 3) - 1; }
          else if (p.mode == SG_MODE_CONVT) {
            const int td = (tap >> 2) & 1, th = (tap >> 1) & 1, tw = tap & 1;
            oz = pd ? 1 - td : -td; oy = phh ? 1 - th : -th; ox = pw ? 1 - tw : -tw;
          }
          {
            const uint32_t b_bytes = (p.b_tma && !(p.diag & 2)) ? (uint32_t)(p.planes * p.bn) * 128u : 0u;
            mbar_arrive_expect_tx(bar, ((p.diag & 1) ? 0u : p.a_stage_bytes) + b_bytes);
            if (b_bytes) {        // the packed weight tile of this K chunk: one linear bulk copy per plane
              const char* bsrc = p.b + ((((size_t)cls * p.kchunks + kc) * p.planes) * p.n_pad + (size_t)nt * p.bn) * 128;
              for (int pl = 0; pl < p.planes; ++pl)
                bulk_g2s(a_base + p.a_stage_bytes + (uint32_t)(pl * p.bn) * 128u, bsrc + (size_t)pl * p.n_pad * 128, (uint32_t)p.bn * 128u, bar);
            }
            for (int sub = 0; sub < ((p.diag & 1) ? 0 : p.mt); ++sub) {
              for (int pl = 0; pl < p.planes; ++pl) {
                const uint32_t dst = a_base + (uint32_t)(sub * p.planes + pl) * kTileBytes;
                if (p.mode == SG_MODE_DENSE) {
                  if (kc < c1chunks) tma_load_2d(dst, &p.tmA[pl], kc * 64, trow[sub], bar);
                  else tma_load_2d(dst, &p.tmA2[pl], (kc - c1chunks) * 64, trow[sub], bar);
                } else {
                  tma_load_5d(dst, &p.tmA[pl], c0, tx0[sub] + ox, ty0[sub] + oy, tz0[sub] + oz, tn0[sub], bar);
                }
              }
            }
          }
          c0 += 64;
          if (c0 >= p.aC) { c0 = 0; ++tap; }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
    // ================================================================ WEIGHT-TILE LOADERS (warps 1-3): the packed B tile of every stage
    // is a linear copy; moving it with 16-byte cp.async (LSU path) instead of a bulk copy leaves the TMA unit -- this kernel's
    // per-SM load limit, ~3.5 clk per gathered 128-byte row -- to the A tiles alone (loads-only 48 -> 44 us on Conv3d 64->128).
    // Gathering A rows this way as well was measured 3x slower than TMA.  Completion: wait_group (one stage behind),
    // generic->async proxy fence, arrive.
    if (warp >= 1 && !p.b_tma) {
      const int bt = tid - 32;                                   // 0..95
      int s = 0; uint32_t ph = 0;
      int pending = 0, oldest = 0;
      const int depth = min(S - 1, p.b_depth);                               // weight tiles in flight per loader thread (see the halo kernels)
      const uint32_t pieces = (uint32_t)p.planes * (uint32_t)p.bn * 8u;      // 16-byte pieces per stage
      const size_t bplane = (size_t)p.n_pad * 128;
      for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x) {
        int cls, nt, mtile, ks;
        decode_work(p, w, cls, nt, mtile, ks);
        const int k0 = ks * cps, k1 = min(p.kchunks, k0 + cps);
        const char* bsrc = p.b + (((size_t)cls * p.kchunks + k0) * p.planes * p.n_pad + (size_t)nt * p.bn) * 128;
        for (int kc = k0; kc < k1; ++kc) {
          mbar_wait(&hdr->empty[s], ph ^ 1, p.err);
          const uint32_t b_dst = smem_u32(stage0 + (size_t)s * p.stage_bytes) + p.a_stage_bytes;
          if (!(p.diag & 2)) {
            if (p.planes == 1) {
              for (uint32_t i = (uint32_t)bt; i < pieces; i += kBLoaderThreads) cp_async16(b_dst + i * 16u, bsrc + (size_t)i * 16u, 16u);
            } else {
              const uint32_t per_plane = (uint32_t)p.bn * 8u;
              for (uint32_t i = (uint32_t)bt; i < pieces; i += kBLoaderThreads) {
                const uint32_t pl = i / per_plane, j = i - pl * per_plane;
                cp_async16(b_dst + i * 16u, bsrc + pl * bplane + (size_t)j * 16u, 16u);
              }
            }
          }
          cp_async_commit();
          if (++pending > depth) {
            cp_async_wait_dyn(depth); fence_proxy_async(); mbar_arrive(&hdr->full[oldest]);
            if (++oldest == S) oldest = 0;
            --pending;
          }
          bsrc += (size_t)p.planes * bplane;
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
      while (pending > 0) {
        cp_async_wait_dyn(pending - 1); fence_proxy_async(); mbar_arrive(&hdr->full[oldest]);
        if (++oldest == S) oldest = 0;
        --pending;
      }
    }
  } else if (warp < 4) {
    // ================================================================ PRODUCERS (cp.async gather fallback)
    const int g = tid & 7, rb = tid >> 3;
    int s = 0; uint32_t ph = 0;
    // loads run `lag` stages ahead of the completion signal: keeps (lag+1) x stage bytes in flight per SM
    const int lag = max(1, S - 2);
    int pending = 0, oldest = 0, tslot = 0, tsig = 0;
    const uint32_t b_tile_bytes = (uint32_t)p.bn * 128u;
    const bool patch_fast = p.mode == SG_MODE_PATCH && (p.aW & 15) == 0;
    if (patch_fast) {
      // ---- Conv3d(1 -> N) on the fp32 volume, W % 16 == 0 (every BASELINE resolution): K = 64 taps = ONE chunk per work item.
      // Thread (rb = tid >> 3, g = tid & 7) fills 8 consecutive output-x rows of 16-byte piece g (taps kd = g >> 1, kh = 2 (g & 1) + {0, 1},
      // kw = 0..3) from two register-blocked input lines.  The loads of BOTH sub-tiles are issued before any conversion (one exposed
      // memory round trip per stage), and the stage is signalled at once: it was written with ordinary stores, nothing is in flight.
      const float* vol = reinterpret_cast<const float*>(p.a_ptr);
      const int oW = p.aW >> 1, oH = p.aH >> 1, oD = p.aD >> 1;
      for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x) {
        int cls, nt, mtile, ks;
        decode_work(p, w, cls, nt, mtile, ks);
        float f[2][2][20];
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub >= p.mt) break;
          const long long gr0 = ((long long)mtile * p.mt + sub) * kTileRows + rb * 8;
          const bool v = gr0 < p.rows;
          const uint32_t r32 = (uint32_t)(v ? gr0 : 0);
          const int ow0 = (int)(r32 % (uint32_t)oW); uint32_t t2 = r32 / (uint32_t)oW;
          const int oh = (int)(t2 % (uint32_t)oH); t2 /= (uint32_t)oH;
          const int od = (int)(t2 % (uint32_t)oD); const uint32_t n = t2 / (uint32_t)oD;
          patch_load8(vol + (size_t)n * p.aD * p.aH * p.aW, p.aD, p.aH, p.aW, od, oh, ow0, g, v, f[sub]);
        }
        mbar_wait(&hdr->empty[s], ph ^ 1, p.err);            // (the loads above are already in flight)
        if (tid == 0) tr((p.diag & 128) && blockIdx.x == 0, 1, tslot++);
        uint8_t* st = stage0 + (size_t)s * p.stage_bytes;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub >= p.mt) break;
          uint8_t* tile = st + (size_t)(sub * p.planes) * kTileBytes;
          patch_store8(f[sub], g, tile, p.planes == 2 ? tile + kTileBytes : nullptr, rb * 8);
        }
        if (warp == 0 && elect_one()) {
          mbar_arrive_expect_tx(&hdr->full[s], b_tile_bytes * p.planes);
          const uint32_t b_dst = smem_u32(st) + p.a_stage_bytes;
          for (int pl = 0; pl < p.planes; ++pl) {
            const char* src = p.b + ((((size_t)cls * p.kchunks) * p.planes + pl) * p.n_pad + (size_t)nt * p.bn) * 128;
            bulk_g2s(b_dst + pl * b_tile_bytes, src, b_tile_bytes, &hdr->full[s]);
          }
        }
        fence_proxy_async(); mbar_arrive(&hdr->full[s]);
        if (tid == 0) tr((p.diag & 128) && blockIdx.x == 0, 2, tsig++);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    } else
    for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x) {
      int cls, nt, mtile, ks;
      decode_work(p, w, cls, nt, mtile, ks);
      const int pd = (cls >> 2) & 1, phh = (cls >> 1) & 1, pw = cls & 1;
      // per-thread row descriptors: 8 rows per M sub-tile
      int ri_n[2][8]; uint32_t ri_c[2][8];
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ri_n[sub][i] = -1; ri_c[sub][i] = 0;
          if (sub < p.mt) {
            long long gr = ((long long)mtile * p.mt + sub) * kTileRows + rb + 16 * i;
            if (gr < p.rows) {
              if (p.mode == SG_MODE_DENSE) {
                ri_n[sub][i] = (int)gr;
              } else if (p.mode == SG_MODE_CONVT) {
                const uint32_t g32 = (uint32_t)gr;                                  // rows < 2^31 (host-checked)
                int qw = (int)(g32 % (uint32_t)p.aW); uint32_t t = g32 / (uint32_t)p.aW;
                int qh = (int)(t % (uint32_t)p.aH); t /= (uint32_t)p.aH;
                int qd = (int)(t % (uint32_t)p.aD); int n = (int)(t / (uint32_t)p.aD);
                ri_n[sub][i] = n * p.aD * p.aH * p.aW;
                ri_c[sub][i] = (uint32_t)(qd + 1) | ((uint32_t)(qh + 1) << 10) | ((uint32_t)(qw + 1) << 20);
              } else {  // CONV / PATCH: rows enumerate the stride-2 output grid
                const int oW = p.aW >> 1, oH = p.aH >> 1, oD = p.aD >> 1;
                const uint32_t g32 = (uint32_t)gr;
                int ow = (int)(g32 % (uint32_t)oW); uint32_t t = g32 / (uint32_t)oW;
                int oh = (int)(t % (uint32_t)oH); t /= (uint32_t)oH;
                int od = (int)(t % (uint32_t)oD); int n = (int)(t / (uint32_t)oD);
                ri_n[sub][i] = n * p.aD * p.aH * p.aW;
                ri_c[sub][i] = (uint32_t)(2 * od) | ((uint32_t)(2 * oh) << 10) | ((uint32_t)(2 * ow) << 20);  // (2o-1)+1
              }
            }
          }
        }
      }
      const int k0 = ks * cps, k1 = min(p.kchunks, k0 + cps);
      for (int kc = k0; kc < k1; ++kc) {
        mbar_wait(&hdr->empty[s], ph ^ 1, p.err);
        if (tid == 0) tr((p.diag & 128) && blockIdx.x == 0, 1, tslot++);
        uint8_t* st = stage0 + (size_t)s * p.stage_bytes;
        const uint32_t a_base = smem_u32(st);
        if (p.mode == SG_MODE_PATCH) {
          // single-channel fp32 volume: K = 64 taps, piece g = taps [8g, 8g+8) = (kd = g>>1, kh = 2(g&1)+{0,1}, kw = 0..3)
          const float* vol = reinterpret_cast<const float*>(p.a_ptr);
          const int kd = g >> 1, khb = (g & 1) * 2;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            if (sub >= p.mt) break;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = 0.f;
              if (ri_n[sub][i] >= 0) {
                const uint32_t c = ri_c[sub][i];
                const int d = (int)(c & 1023) - 1 + kd;
                const int w0 = (int)((c >> 20) & 1023) - 1;
                if (d >= 0 && d < p.aD) {
#pragma unroll
                  for (int hh = 0; hh < 2; ++hh) {
                    const int h = (int)((c >> 10) & 1023) - 1 + khb + hh;
                    if (h >= 0 && h < p.aH) {
                      const float* rowp = vol + (size_t)ri_n[sub][i] + ((size_t)d * p.aH + h) * p.aW;
#pragma unroll
                      for (int ww = 0; ww < 4; ++ww) {
                        const int x = w0 + ww;
                        if (x >= 0 && x < p.aW) v[hh * 4 + ww] = __ldg(rowp + x);
                      }
                    }
                  }
                }
              }
              const uint32_t row = rb + 16 * i;
              uint4 hi;
              hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
              hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
              uint8_t* tile = st + (size_t)(sub * p.planes) * kTileBytes;
              *reinterpret_cast<uint4*>(tile + sw128(row, g)) = hi;
              if (p.planes == 2) {
                uint4 lo;
                lo.x = pack_bf16x2(v[0] - bf16lo_to_f(hi.x), v[1] - bf16hi_to_f(hi.x));
                lo.y = pack_bf16x2(v[2] - bf16lo_to_f(hi.y), v[3] - bf16hi_to_f(hi.y));
                lo.z = pack_bf16x2(v[4] - bf16lo_to_f(hi.z), v[5] - bf16hi_to_f(hi.z));
                lo.w = pack_bf16x2(v[6] - bf16lo_to_f(hi.w), v[7] - bf16hi_to_f(hi.w));
                *reinterpret_cast<uint4*>(tile + kTileBytes + sw128(row, g)) = lo;
              }
            }
          }
        } else {
          // bf16 plane gathers
          const char* src_base; long long src_ps; int rowC, coff;
          int dd = 0, dh = 0, dw = 0;
          if (p.mode == SG_MODE_DENSE) {
            const int c1chunks = (p.aC + 63) >> 6;      // chunks served by the first source (its tail chunk is zero-filled)
            if (kc < c1chunks) { src_base = p.a_ptr; src_ps = p.a_ps; rowC = p.aC; coff = kc * 64 + g * 8; }
            else { src_base = p.a2_ptr; src_ps = p.a2_ps; rowC = p.a2C; coff = (kc - c1chunks) * 64 + g * 8; }
          } else {
            const uint32_t e = ktab[kc * 8 + g];
            src_base = p.a_ptr; src_ps = p.a_ps; rowC = p.aC; coff = (int)(e & 0xffff);
            dd = (int)((e >> 16) & 7); dh = (int)((e >> 19) & 7); dw = (int)((e >> 22) & 7);
            if (p.mode == SG_MODE_CONVT) { dd = pd ? 1 - dd : -dd; dh = phh ? 1 - dh : -dh; dw = pw ? 1 - dw : -dw; }
          }
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            if (sub >= p.mt) break;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              long long off = 0; uint32_t nbytes = 0;
              if (ri_n[sub][i] >= 0) {
                if (p.mode == SG_MODE_DENSE) {
                  if (coff < rowC) { off = (long long)ri_n[sub][i] * rowC + coff; nbytes = 16; }
                } else {
                  const uint32_t c = ri_c[sub][i];
                  const int d = (int)(c & 1023) - 1 + dd, h = (int)((c >> 10) & 1023) - 1 + dh, x = (int)((c >> 20) & 1023) - 1 + dw;
                  if (d >= 0 && d < p.aD && h >= 0 && h < p.aH && x >= 0 && x < p.aW) {
                    off = ((long long)ri_n[sub][i] + ((long long)d * p.aH + h) * p.aW + x) * rowC + coff; nbytes = 16;
                  }
                }
              }
              const uint32_t row = rb + 16 * i;
              const uint32_t dst = a_base + (uint32_t)(sub * p.planes) * kTileBytes + sw128(row, g);
              cp_async16(dst, src_base + off * 2, nbytes);
              if (p.planes == 2) cp_async16(dst + kTileBytes, src_base + (off + src_ps) * 2, nbytes);
            }
          }
        }
        if (warp == 0 && elect_one()) {
          mbar_arrive_expect_tx(&hdr->full[s], b_tile_bytes * p.planes);
          const uint32_t b_dst = a_base + p.a_stage_bytes;
          for (int pl = 0; pl < p.planes; ++pl) {
            const char* src = p.b + ((((size_t)cls * p.kchunks + kc) * p.planes + pl) * p.n_pad + (size_t)nt * p.bn) * 128;
            bulk_g2s(b_dst + pl * b_tile_bytes, src, b_tile_bytes, &hdr->full[s]);
          }
        }
        if (p.mode == SG_MODE_PATCH) {
          // the im2col tile was written with ordinary stores: the stage is complete now (running `lag` stages behind, as the cp.async
          // gathers must, kept the first MMA waiting for FOUR tiles -- 27 k of the layer's 77 k clocks, profiles/r02e_trace_patch.txt)
          fence_proxy_async(); mbar_arrive(&hdr->full[s]);
          if (tid == 0) tr((p.diag & 128) && blockIdx.x == 0, 2, tsig++);
        } else {
          cp_async_commit();
          if (++pending > lag) {
            cp_async_wait_dyn(lag); fence_proxy_async(); mbar_arrive(&hdr->full[oldest]);
            if (tid == 0) tr((p.diag & 128) && blockIdx.x == 0, 2, tsig++);
            if (++oldest == S) oldest = 0;
            --pending;
          }
        }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    while (pending > 0) {
      cp_async_wait_dyn(pending - 1); fence_proxy_async(); mbar_arrive(&hdr->full[oldest]);
      if (++oldest == S) oldest = 0;
      --pending;
    }
  } else if (warp == 4) {
    // ================================================================ MMA ISSUER
    // ONE elected thread runs the whole loop (barrier polls included)
    const uint32_t idesc = umma_idesc(128, p.bn, false, false);
    int s = 0; uint32_t ph = 0; int it = 0; int tcnt = 0;
    if (elect_one())
    for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x, ++it) {
      int cls, nt, mtile, ks;
      decode_work(p, w, cls, nt, mtile, ks);
      const int ab = (p.acc_bufs == 2) ? (it & 1) : 0;
      const uint32_t aph = (uint32_t)((it / p.acc_bufs) & 1);
      mbar_wait(&hdr->accempty[ab], aph ^ 1, p.err);
      tc_fence_after();
      const int k0 = ks * cps, k1 = min(p.kchunks, k0 + cps);
      for (int kc = k0; kc < k1; ++kc) {
        mbar_wait(&hdr->full[s], ph, p.err);
        tc_fence_after();
        tr((p.diag & 128) && blockIdx.x == 0, 4, tcnt);
        {
          const uint32_t a_base = smem_u32(stage0 + (size_t)s * p.stage_bytes);
          const uint32_t b_base = a_base + p.a_stage_bytes;
          const uint32_t b_tile_bytes = (uint32_t)p.bn * 128u;
          for (int sub = 0; sub < ((p.diag & 4) ? 0 : p.mt); ++sub) {
            const uint32_t d_addr = tmem_base + (uint32_t)((ab * p.mt + sub) * p.acc_slot);
            const uint32_t a_hi = a_base + (uint32_t)(sub * p.planes) * kTileBytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t da = umma_desc(a_hi + kk * 32, 16, 1024);
              const uint64_t db = umma_desc(b_base + kk * 32, 16, 1024);
              umma_bf16(d_addr, da, db, idesc, (kc > k0 || kk > 0) ? 1u : 0u);
              if (p.planes == 2) {
                const uint64_t da_lo = umma_desc(a_hi + kTileBytes + kk * 32, 16, 1024);
                const uint64_t db_lo = umma_desc(b_base + b_tile_bytes + kk * 32, 16, 1024);
                umma_bf16(d_addr, da, db_lo, idesc, 1u);
                umma_bf16(d_addr, da_lo, db, idesc, 1u);
              }
            }
          }
          umma_commit(&hdr->empty[s]);
          tr((p.diag & 128) && blockIdx.x == 0, 5, tcnt++);
        }
        if (++s == S) { s = 0; ph ^= 1; }
      }
      umma_commit(&hdr->accfull[ab]);
    }
    __syncwarp();
  } else {
    // ================================================================ EPILOGUE
    switch (p.act) {
      case ACT_NONE: epilogue_role<ACT_NONE>(p, hdr, tmem_base, cps); break;
      case ACT_LRELU: epilogue_role<ACT_LRELU>(p, hdr, tmem_base, cps); break;
      case ACT_RELU: epilogue_role<ACT_RELU>(p, hdr, tmem_base, cps); break;
      default: epilogue_role<-1>(p, hdr, tmem_base, cps); break;
    }
  }
  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ================================================================================================ halo-reuse variant
// Conv3d k4 s2 p1 forward gather (SG_MODE_CONV) and ConvTranspose3d k4 s2 p1 per-class gather (SG_MODE_CONVT) on an 8 x 8 x gz
// row grid, bf16 (planes = 1), C % 64 == 0.  The plain kernel loads one 128-row A tile per (tap, 64-channel chunk): the TMA unit
// needs ~3.5 clk per gathered 128-byte row, which bounds those layers.  Here the taps that differ only by +1 in
// x and/or z of the stride-2 (or unit-stride) sample grid -- (qz, qx) in {0,1}^2 for a fixed (kh | th, parity pz, px) -- read ONE
// halo block [(2 mt + 1) z][8 y][9 x] rows of 128 B: tap (qz, qx) is the block viewed from row (qz*8*9 + qx) on, 8-row groups
// 9 rows apart.  tcgen05 applies the 128B swizzle to absolute shared-memory address bits (tools/exp_umma_view.cu), so such
// shifted, non-1024-aligned views with SBO = 9*128 are valid A operands.  2.4x (mt = 1) to 2.8x (mt = 2) fewer TMA rows.
//   smem: [header][2 A blocks][b_stages weight tiles]; roles as in sg_igemm_kernel; ksplit = 1.
struct HaloSeq {     // position in the K sequence: group (kh|th, pz, px) -> channel chunk -> 4 taps
  int grp, cc, t4;
};

__device__ __forceinline__ int halo_tap(const IgemmP& p, int grp, int t4) {
  const int qz = t4 >> 1, qx = t4 & 1;
  if (p.mode == SG_MODE_CONV) {          // grp = kh*4 + pz*2 + px ; k = 2 q + parity
    const int kh = grp >> 2, pz = (grp >> 1) & 1, px = grp & 1;
    return (2 * qz + pz) * 16 + kh * 4 + (2 * qx + px);
  }
  const int th = grp;                      // CONVT: shift q = 1 - t  (offset = (parity ? 0 : -1) + q)
  return (1 - qz) * 4 + th * 2 + (1 - qx);
}

// One filter tap of the halo kernels: mt sub-tiles x 4 K steps (K = 16 each) of one 64-channel chunk.  The single issuing thread is the
// pace-setter of the narrow-N layers (ConvT 128->64: 8 MMAs of 32 clk per tap), so the descriptors are not rebuilt per MMA: their
// constant fields live in `a_hi` / `b_hi`, the start-address field (14 bits of address >> 4) is advanced by plain adds.
template <bool PAIR>
__device__ __forceinline__ void halo_issue_tap(uint32_t d0, uint32_t acc_slot, int mt, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                               uint32_t idesc, bool first) {
  constexpr uint32_t kSubStep = (2u * 8u * 9u * 128u) >> 4;        // two z planes of the halo block per sub-tile
#pragma unroll
  for (int sub = 0; sub < 2; ++sub) {
    if (sub < mt) {
      const uint32_t d_addr = d0 + (uint32_t)sub * acc_slot;
      const uint32_t al = a_lo + (uint32_t)sub * kSubStep;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(al + 2u * kk);
        const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 2u * kk);
        if (PAIR) umma2_bf16(d_addr, da, db, idesc, (first && kk == 0) ? 0u : 1u);
        else umma_bf16(d_addr, da, db, idesc, (first && kk == 0) ? 0u : 1u);
      }
    }
  }
}

__global__ void __launch_bounds__(kIgemmThreads, 1) sg_igemm_halo_kernel(const __grid_constant__ IgemmP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  uint8_t* blk0 = smem + kSmemHeader;
  uint8_t* bst0 = blk0 + (size_t)p.a_blocks * p.blk_bytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool tron = (p.diag & 128) && blockIdx.x == 0;
  if (tid == 0) tr(tron, 7, 0);
  const int SB = p.b_stages;
  const uint32_t b_tile_bytes = (uint32_t)p.bn * 128u;
  const int cchunks = p.aC >> 6;
  const int ngroups = (p.mode == SG_MODE_CONV) ? 16 : 2;
  if (tid == 0) {
    for (int s = 0; s < SB; ++s) { mbar_init(&hdr->full[s], p.b_tma ? 1 : kBLoaderThreads); mbar_init(&hdr->empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&hdr->accfull[i], 1); mbar_init(&hdr->accempty[i], 128); }
    for (int i = 0; i < 3; ++i) { mbar_init(&hdr->blk_full[i], 1); mbar_init(&hdr->blk_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  if (tid == 0) tr(tron, 7, 1);
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0) atomicExch(p.err, kErrSmemAlign);
    __trap();
  }

  if (warp == 0) {
    // ================================================================ TMA producer: one halo block per (group, channel chunk)
    if (elect_one()) {
      tma_prefetch_desc(&p.tmH);
      int bi = 0; uint32_t bph = 0; int tb = 0;
      const int lgz = 31 - __clz(max(p.gz, 1));
      for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x) {
        int cls, nt, mtile, ks;
        decode_work(p, w, cls, nt, mtile, ks);
        const int pd = (cls >> 2) & 1, phh = (cls >> 1) & 1, pw = cls & 1;
        const uint32_t row0 = (uint32_t)(mtile * p.mt) * kTileRows;        // rows = ((n*gz + z)*8 + y)*8 + x
        const int z0 = (int)((row0 >> 6) & (uint32_t)(p.gz - 1)), n0 = (int)(row0 >> (6 + lgz));
        for (int grp = 0; grp < ngroups; ++grp) {
          int x, y, z;
          if (p.mode == SG_MODE_CONV) {
            const int kh = grp >> 2, pz = (grp >> 1) & 1, px = grp & 1;
            x = -1 + px; y = -1 + kh; z = 2 * z0 - 1 + pz;
          } else {
            const int th = grp;
            x = pw ? 0 : -1; y = phh ? 1 - th : -th; z = z0 + (pd ? 0 : -1);
          }
          for (int cc = 0; cc < cchunks; ++cc) {
            mbar_wait(&hdr->blk_empty[bi], bph ^ 1, p.err);
            tr(tron, 0, tb++);
            mbar_arrive_expect_tx(&hdr->blk_full[bi], p.blk_bytes);
            tma_load_5d(smem_u32(blk0 + (size_t)bi * p.blk_bytes), &p.tmH, cc * 64, x, y, z, n0, &hdr->blk_full[bi]);
            if (++bi == p.a_blocks) { bi = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp < 4 && p.b_tma) {
    // ================================================================ weight tiles by 1-D bulk copies (one elected thread of warp 1)
    if (warp == 1 && elect_one()) {
      int s = 0; uint32_t ph = 0; int tl = 0;
      for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x) {
        int cls, nt, mtile, ks;
        decode_work(p, w, cls, nt, mtile, ks);
        const char* bcls = p.b + ((size_t)cls * p.kchunks * p.n_pad + (size_t)nt * p.bn) * 128;
        for (int grp = 0; grp < ngroups; ++grp)
          for (int cc = 0; cc < cchunks; ++cc)
            for (int t4 = 0; t4 < 4; ++t4) {
              const int kc = halo_tap(p, grp, t4) * cchunks + cc;
              mbar_wait(&hdr->empty[s], ph ^ 1, p.err);
              tr(tron, 1, tl++);
              mbar_arrive_expect_tx(&hdr->full[s], b_tile_bytes);
              bulk_g2s(smem_u32(bst0 + (size_t)s * b_tile_bytes), bcls + (size_t)kc * p.n_pad * 128, b_tile_bytes, &hdr->full[s]);
              if (++s == SB) { s = 0; ph ^= 1; }
            }
      }
    }
  } else if (warp < 4) {
    // ================================================================ weight-tile loaders (warps 1-3, cp.async), K sequence order
    const int bt = tid - 32;
    int s = 0; uint32_t ph = 0;
    int pending = 0, oldest = 0, tl = 0, ta = 0;
    // committed weight tiles a loader thread runs ahead of its completion signal.  One: the 16-byte LDGSTS path moves ~23 B/clk per SM however
    // many are queued; with two ahead ConvT 128->64 went from 55 to 70 us, with five the landing time of a tile went from ~1400 to ~9000 clk
    // (profiles/r02e_trace_halo.txt, r02e_sweep_cfgs.txt) -- deeper is slower, not faster.
    const int depth = min(SB - 1, p.b_depth);
    const uint32_t pieces = (uint32_t)p.bn * 8u;
    for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x) {
      int cls, nt, mtile, ks;
      decode_work(p, w, cls, nt, mtile, ks);
      const char* bcls = p.b + ((size_t)cls * p.kchunks * p.n_pad + (size_t)nt * p.bn) * 128;
      for (int grp = 0; grp < ngroups; ++grp)
        for (int cc = 0; cc < cchunks; ++cc)
          for (int t4 = 0; t4 < 4; ++t4) {
            const int kc = halo_tap(p, grp, t4) * cchunks + cc;
            const char* bsrc = bcls + (size_t)kc * p.n_pad * 128;
            mbar_wait(&hdr->empty[s], ph ^ 1, p.err);
            if (bt == 0) tr(tron, 1, tl++);
            const uint32_t b_dst = smem_u32(bst0 + (size_t)s * b_tile_bytes);
            for (uint32_t i = (uint32_t)bt; i < pieces; i += kBLoaderThreads) cp_async16(b_dst + i * 16u, bsrc + (size_t)i * 16u, 16u);
            cp_async_commit();
            if (++pending > depth) {
              cp_async_wait_dyn(depth); fence_proxy_async(); mbar_arrive(&hdr->full[oldest]);
              if (bt == 0) tr(tron, 2, ta++);
              if (++oldest == SB) oldest = 0;
              --pending;
            }
            if (++s == SB) { s = 0; ph ^= 1; }
          }
    }
    while (pending > 0) {
      cp_async_wait_dyn(pending - 1); fence_proxy_async(); mbar_arrive(&hdr->full[oldest]);
      if (++oldest == SB) oldest = 0;
      --pending;
    }
  } else if (warp == 4) {
    // ================================================================ MMA issuer (one elected thread)
    const uint32_t idesc = umma_idesc(128, p.bn, false, false);
    const uint32_t a_hi = (uint32_t)(umma_desc(0, 16, 9 * 128) >> 32), b_hi = (uint32_t)(umma_desc(0, 16, 1024) >> 32);
    const uint32_t a_lbo = (16u >> 4) << 16;      // LBO field (bits 16..29) rides in the low word
    const int mt_ = p.mt; const uint32_t acc_slot_ = (uint32_t)p.acc_slot;
    int s = 0; uint32_t ph = 0; int it = 0;
    int bi = 0; uint32_t bph = 0; int tb = 0, tt = 0;
    if (elect_one())
    for (long long w = blockIdx.x; w < p.work_total; w += gridDim.x, ++it) {
      const int ab = (p.acc_bufs == 2) ? (it & 1) : 0;
      const uint32_t aph = (uint32_t)((it / p.acc_bufs) & 1);
      mbar_wait(&hdr->accempty[ab], aph ^ 1, p.err);
      tr(tron, 7, 2);
      tc_fence_after();
      bool first = true;
      for (int grp = 0; grp < ngroups; ++grp)
        for (int cc = 0; cc < cchunks; ++cc) {
          mbar_wait(&hdr->blk_full[bi], bph, p.err);
          tr(tron, 3, tb++);
          const uint32_t blk = smem_u32(blk0 + (size_t)bi * p.blk_bytes);
          for (int t4 = 0; t4 < 4; ++t4) {
            const int qz = t4 >> 1, qx = t4 & 1;
            mbar_wait(&hdr->full[s], ph, p.err);
            tr(tron, 4, tt);
            tc_fence_after();
            const uint32_t b_base = smem_u32(bst0 + (size_t)s * b_tile_bytes);
            halo_issue_tap<false>(tmem_base + (uint32_t)(ab * mt_ * acc_slot_), acc_slot_, mt_,
                                  (((blk & 0x3FFFFu) >> 4) + (uint32_t)((qz * 72 + qx) * 8)) | a_lbo, ((b_base & 0x3FFFFu) >> 4) | a_lbo, a_hi, b_hi, idesc, first);
            first = false;
            umma_commit(&hdr->empty[s]);
            tr(tron, 5, tt++);
            if (++s == SB) { s = 0; ph ^= 1; }
          }
          umma_commit(&hdr->blk_empty[bi]);
          if (++bi == p.a_blocks) { bi = 0; bph ^= 1; }
        }
      umma_commit(&hdr->accfull[ab]);
    }
    __syncwarp();
  } else {
    switch (p.act) {
      case ACT_NONE: epilogue_role<ACT_NONE>(p, hdr, tmem_base, p.kchunks); break;
      case ACT_LRELU: epilogue_role<ACT_LRELU>(p, hdr, tmem_base, p.kchunks); break;
      case ACT_RELU: epilogue_role<ACT_RELU>(p, hdr, tmem_base, p.kchunks); break;
      default: epilogue_role<-1>(p, hdr, tmem_base, p.kchunks); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) tr(tron, 7, 3);
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ================================================================================================ halo variant on CTA PAIRS
// Same K sequence and halo views as sg_igemm_halo_kernel, on a 2-CTA cluster with tcgen05.mma.cta_group::2: every MMA spans 256 rows
// (this CTA's 128-row sub-tile + the peer's) and the weight tile is SPLIT between the two CTAs -- each stages only bn / 2 of its rows.
// What this buys: the single-CTA M128 x N128 x K16 MMA reads 8 KB of shared-memory operands for 64 clocks of math and is paced by
// those reads (~108 clk measured); in the pair each SM reads its 4 KB A slice + 2 KB B slice for the same math.
//   leader (cluster rank 0): its MMA thread issues for both CTAs; its barriers collect the pair's arrivals
//     blk_full[i]   count 1 + transaction bytes of BOTH CTAs' halo blocks (peer's TMA reports to the leader: .cta_group::2 form)
//     full[s]       2 x 96 weight-loader threads (the peer's arrive through shared::cluster)
//     accempty[ab]  2 x 128 epilogue threads
//   both CTAs: blk_empty / empty / accfull are signalled by tcgen05.commit ... multicast::cluster to the same offset in each CTA.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kIgemmThreads, 1) sg_igemm_halo2_kernel(const __grid_constant__ IgemmP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  uint8_t* blk0 = smem + kSmemHeader;
  uint8_t* bst0 = blk0 + (size_t)p.a_blocks * p.blk_bytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool tron = (p.diag & 128) && blockIdx.x == 0;
  if (tid == 0) tr(tron, 7, 0);
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int SB = p.b_stages;
  const uint32_t b_half_bytes = (uint32_t)(p.bn / 2) * 128u;      // this CTA's half of a weight tile
  const int cchunks = p.aC >> 6;
  const int ngroups = (p.mode == SG_MODE_CONV) ? 16 : 2;
  const long long w0 = blockIdx.x >> 1, wstep = gridDim.x >> 1;
  if (tid == 0) {
    for (int s = 0; s < SB; ++s) { mbar_init(&hdr->full[s], p.b_tma ? 1 : 2); mbar_init(&hdr->empty[s], 1); }      // cp.async: one arrival per CTA (see the loaders)
    for (int i = 0; i < 2; ++i) { mbar_init(&hdr->accfull[i], 1); mbar_init(&hdr->accempty[i], 2); }
    for (int i = 0; i < 3; ++i) { mbar_init(&hdr->blk_full[i], 1); mbar_init(&hdr->blk_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc2(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();          // the allocator's shared-memory write vs this CTA's readers (compute-sanitizer racecheck does not take the cluster
  cluster_sync_all();       // barrier below as a shared-memory synchronisation point: profiles/r02f_sanitizer.txt); then both CTAs are set up
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  if (tid == 0) tr(tron, 7, 1);
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0) atomicExch(p.err, kErrSmemAlign);
    __trap();
  }

  if (warp == 0) {
    // ================================================================ TMA producer: this CTA's halo blocks; bytes reported to the leader
    if (elect_one()) {
      tma_prefetch_desc(&p.tmH);
      int bi = 0; uint32_t bph = 0; int tb = 0;
      const int lgz = 31 - __clz(max(p.gz, 1));
      for (long long w = w0; w < p.work_total; w += wstep) {
        int cls, nt, mtile, ks;
        decode_work(p, w, cls, nt, mtile, ks);
        mtile = mtile * 2 + (int)crank;
        const int pd = (cls >> 2) & 1, phh = (cls >> 1) & 1, pw = cls & 1;
        const uint32_t row0 = (uint32_t)(mtile * p.mt) * kTileRows;
        const int z0 = (int)((row0 >> 6) & (uint32_t)(p.gz - 1)), n0 = (int)(row0 >> (6 + lgz));
        for (int grp = 0; grp < ngroups; ++grp) {
          int x, y, z;
          if (p.mode == SG_MODE_CONV) {
            const int kh = grp >> 2, pz = (grp >> 1) & 1, px = grp & 1;
            x = -1 + px; y = -1 + kh; z = 2 * z0 - 1 + pz;
          } else {
            const int th = grp;
            x = pw ? 0 : -1; y = phh ? 1 - th : -th; z = z0 + (pd ? 0 : -1);
          }
          for (int cc = 0; cc < cchunks; ++cc) {
            mbar_wait_cluster(&hdr->blk_empty[bi], bph ^ 1, p.err);
            tr(tron, 0, tb++);
            if (leader) mbar_arrive_expect_tx(&hdr->blk_full[bi], 2u * p.blk_bytes);
            tma_load_5d_2sm(smem_u32(blk0 + (size_t)bi * p.blk_bytes), &p.tmH, cc * 64, x, y, z, n0, &hdr->blk_full[bi]);
            if (++bi == p.a_blocks) { bi = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp < 4 && p.b_tma) {
    // ================================================================ weight tiles by TMA: this CTA's half of every tap's tile, bytes reported to the leader
    if (warp == 1 && elect_one()) {
      tma_prefetch_desc(&p.tmB);
      int s = 0; uint32_t ph = 0; int tl = 0;
      for (long long w = w0; w < p.work_total; w += wstep) {
        int cls, nt, mtile, ks;
        decode_work(p, w, cls, nt, mtile, ks);
        const int row0 = cls * p.kchunks * p.n_pad + nt * p.bn + (int)crank * (p.bn / 2);
        for (int grp = 0; grp < ngroups; ++grp)
          for (int cc = 0; cc < cchunks; ++cc)
            for (int t4 = 0; t4 < 4; ++t4) {
              const int kc = halo_tap(p, grp, t4) * cchunks + cc;
              mbar_wait_cluster(&hdr->empty[s], ph ^ 1, p.err);
              tr(tron, 1, tl++);
              if (leader) mbar_arrive_expect_tx(&hdr->full[s], 2u * b_half_bytes);
              tma_load_2d_2sm(smem_u32(bst0 + (size_t)s * b_half_bytes), &p.tmB, 0, row0 + kc * p.n_pad, &hdr->full[s]);
              if (++s == SB) { s = 0; ph ^= 1; }
            }
      }
    }
  } else if (warp < 4) {
    // ================================================================ weight-tile loaders: this CTA's HALF (rows crank * bn/2 ...) of every tap's tile
    const int bt = tid - 32;
    int s = 0; uint32_t ph = 0;
    int pending = 0, oldest = 0, tl = 0, ta = 0;
    // committed weight tiles a loader thread runs ahead of its completion signal.  One: the 16-byte LDGSTS path moves ~23 B/clk per SM however
    // many are queued; with two ahead ConvT 128->64 went from 55 to 70 us, with five the landing time of a tile went from ~1400 to ~9000 clk
    // (profiles/r02e_trace_halo.txt, r02e_sweep_cfgs.txt) -- deeper is slower, not faster.
    const int depth = min(SB - 1, p.b_depth);
    const uint32_t pieces = (uint32_t)(p.bn / 2) * 8u;
    uint32_t full_addr[kMaxStages];
    for (int i = 0; i < kMaxStages; ++i) full_addr[i] = mapa_u32(smem_u32(&hdr->full[i]), 0);
    for (long long w = w0; w < p.work_total; w += wstep) {
      int cls, nt, mtile, ks;
      decode_work(p, w, cls, nt, mtile, ks);
      const char* bcls = p.b + ((size_t)cls * p.kchunks * p.n_pad + (size_t)nt * p.bn + (size_t)crank * (p.bn / 2)) * 128;
      for (int grp = 0; grp < ngroups; ++grp)
        for (int cc = 0; cc < cchunks; ++cc)
          for (int t4 = 0; t4 < 4; ++t4) {
            const int kc = halo_tap(p, grp, t4) * cchunks + cc;
            const char* bsrc = bcls + (size_t)kc * p.n_pad * 128;
            mbar_wait_cluster(&hdr->empty[s], ph ^ 1, p.err);
            if (bt == 0) tr(tron, 1, tl++);
            const uint32_t b_dst = smem_u32(bst0 + (size_t)s * b_half_bytes);
            for (uint32_t i = (uint32_t)bt; i < pieces; i += kBLoaderThreads) cp_async16(b_dst + i * 16u, bsrc + (size_t)i * 16u, 16u);
            cp_async_commit();
            if (++pending > depth) {
              // the 96 loader threads meet on a named barrier and ONE of them arrives at the leader: 192 cluster-scope arrivals per
              // tap on a single mbarrier (half of them remote) serialised the pair to half the single-CTA rate
              cp_async_wait_dyn(depth); fence_proxy_async();
              asm volatile("bar.sync 2, 96;" ::: "memory");
              if (bt == 0) { mbar_arrive_cluster(full_addr[oldest]); tr(tron, 2, ta++); }
              if (++oldest == SB) oldest = 0;
              --pending;
            }
            if (++s == SB) { s = 0; ph ^= 1; }
          }
    }
    while (pending > 0) {
      cp_async_wait_dyn(pending - 1); fence_proxy_async();
      asm volatile("bar.sync 2, 96;" ::: "memory");
      if (bt == 0) mbar_arrive_cluster(full_addr[oldest]);
      if (++oldest == SB) oldest = 0;
      --pending;
    }
  } else if (warp == 4) {
    // ================================================================ MMA issuer: one elected thread of the LEADER CTA
    if (leader && elect_one()) {
      const uint32_t idesc = umma_idesc(256, p.bn, false, false);
      const uint32_t a_hi = (uint32_t)(umma_desc(0, 16, 9 * 128) >> 32), b_hi = (uint32_t)(umma_desc(0, 16, 1024) >> 32);
      const uint32_t a_lbo = (16u >> 4) << 16;
      const int mt_ = p.mt; const uint32_t acc_slot_ = (uint32_t)p.acc_slot;
      int s = 0; uint32_t ph = 0; int it = 0;
      int bi = 0; uint32_t bph = 0; int tb = 0, tt = 0;
      for (long long w = w0; w < p.work_total; w += wstep, ++it) {
        const int ab = (p.acc_bufs == 2) ? (it & 1) : 0;
        const uint32_t aph = (uint32_t)((it / p.acc_bufs) & 1);
        mbar_wait_cluster(&hdr->accempty[ab], aph ^ 1, p.err);
      tr(tron, 7, 2);
        tc_fence_after();
        bool first = true;
        for (int grp = 0; grp < ngroups; ++grp)
          for (int cc = 0; cc < cchunks; ++cc) {
            mbar_wait_cluster(&hdr->blk_full[bi], bph, p.err);
          tr(tron, 3, tb++);
            const uint32_t blk = smem_u32(blk0 + (size_t)bi * p.blk_bytes);
            for (int t4 = 0; t4 < 4; ++t4) {
              const int qz = t4 >> 1, qx = t4 & 1;
              mbar_wait_cluster(&hdr->full[s], ph, p.err);
            tr(tron, 4, tt);
              tc_fence_after();
              const uint32_t b_base = smem_u32(bst0 + (size_t)s * b_half_bytes);
              halo_issue_tap<true>(tmem_base + (uint32_t)(ab * mt_ * acc_slot_), acc_slot_, mt_,
                                   (((blk & 0x3FFFFu) >> 4) + (uint32_t)((qz * 72 + qx) * 8)) | a_lbo, ((b_base & 0x3FFFFu) >> 4) | a_lbo, a_hi, b_hi, idesc, first);
              first = false;
              umma2_commit(&hdr->empty[s]);
            tr(tron, 5, tt++);
              if (++s == SB) { s = 0; ph ^= 1; }
            }
            umma2_commit(&hdr->blk_empty[bi]);
            if (++bi == p.a_blocks) { bi = 0; bph ^= 1; }
          }
        umma2_commit(&hdr->accfull[ab]);
      }
    }
    __syncwarp();
  } else {
    switch (p.act) {
      case ACT_NONE: epilogue_role<ACT_NONE>(p, hdr, tmem_base, p.kchunks); break;
      case ACT_LRELU: epilogue_role<ACT_LRELU>(p, hdr, tmem_base, p.kchunks); break;
      case ACT_RELU: epilogue_role<ACT_RELU>(p, hdr, tmem_base, p.kchunks); break;
      default: epilogue_role<-1>(p, hdr, tmem_base, p.kchunks); break;
    }
  }
  // neither CTA may leave while the other can still touch its shared memory, barriers or tensor memory
  tc_fence_before();
  cluster_sync_all();
  if (tid == 0) tr(tron, 7, 3);
  if (warp == 4) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

// ================================================================================================ host launcher
static int igemm_validate(const sg_igemm_args* a) {
  if (!a) return sg_fail(-1, "sg_igemm: null args");
  if (a->planes != 1 && a->planes != 2) return sg_fail(-2, "sg_igemm: planes must be 1 or 2");
  if (a->mode < 0 || a->mode > 3) return sg_fail(-3, "sg_igemm: bad mode");
  if (a->k <= 0 || (a->k & 63)) return sg_fail(-4, "sg_igemm: K must be a positive multiple of 64");
  if (a->n_pad <= 0 || (a->n_pad & 15) || a->n_valid <= 0 || a->n_valid > a->n_pad) return sg_fail(-5, "sg_igemm: bad N");
  if (a->rows < 0) return sg_fail(-6, "sg_igemm: negative rows");
  if (a->mode != SG_MODE_DENSE && a->rows >= (1LL << 31)) return sg_fail(-6, "sg_igemm: conv rows must be < 2^31");
  if (!a->a.ptr || !a->b_packed || !a->out) return sg_fail(-7, "sg_igemm: null tensor");
  if (a->mode == SG_MODE_DENSE) {
    if (a->a.c & 7) return sg_fail(-8, "sg_igemm: DENSE needs C % 8 == 0");
    int kk = a->a2.ptr ? a->a.c + ((a->a2.c + 63) & ~63) : ((a->a.c + 63) & ~63);
    if (kk != a->k) return sg_fail(-8, "sg_igemm: DENSE needs K == roundup64(a.c) (or a.c + roundup64(a2.c))");
    if (a->a2.ptr && ((a->a.c & 63) || (a->a2.c & 7))) return sg_fail(-9, "sg_igemm: two-source DENSE needs a.c % 64 == 0, a2.c % 8 == 0");
  } else if (a->mode == SG_MODE_CONV) {
    if ((a->a.c & 7) || a->k != 64 * a->a.c) return sg_fail(-10, "sg_igemm: CONV needs C%8==0 and K==64*C");
    if ((a->a.d | a->a.h | a->a.w) & 1) return sg_fail(-11, "sg_igemm: CONV needs even dims");
  } else if (a->mode == SG_MODE_CONVT) {
    if ((a->a.c & 7) || a->k != 8 * a->a.c) return sg_fail(-12, "sg_igemm: CONVT needs C%8==0 and K==8*C");
    if (a->out_d != 2 * a->a.d || a->out_h != 2 * a->a.h || a->out_w != 2 * a->a.w) return sg_fail(-13, "sg_igemm: CONVT out dims");
  } else {
    if (a->a.c != 1 || a->k != 64) return sg_fail(-14, "sg_igemm: PATCH needs C==1, K==64");
    if ((a->a.d | a->a.h | a->a.w) & 1) return sg_fail(-11, "sg_igemm: PATCH needs even dims");
  }
  if (a->a.d > 1000 || a->a.h > 1000 || a->a.w > 1000) return sg_fail(-15, "sg_igemm: dims too large");
  if (a->out_kind < 0 || a->out_kind > 2) return sg_fail(-16, "sg_igemm: bad out_kind");
  return 0;
}


// ================================================================================================ split-K finish
// out[row, n] = act( sum_s P[s][row][n] + bias[n] ) (* act'(mask)) -> bf16 planes / fp32.  One thread per 8 consecutive columns.
struct FinishP {
  const float* ws; int ksplit; long long slab, rows; int n_pad, n_valid;
  const float* bias; int bias_mod, act;
  const bf16* mask; long long mask_ps; int mask_act;
  char* out; long long out_ps; int out_kind, out_ld, planes;
};

__global__ void __launch_bounds__(256) sg_splitk_finish_kernel(const FinishP p) {
  const int groups = p.n_pad >> 3;
  const long long total = p.rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / groups;
    const int n0 = (int)(i - row * groups) * 8;
    if (n0 >= p.n_valid) continue;
    float v[8];
    {
      const float4* src = reinterpret_cast<const float4*>(p.ws + row * p.n_pad + n0);
      float4 a = __ldcs(src), b = __ldcs(src + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    for (int s = 1; s < p.ksplit; ++s) {
      const float4* src = reinterpret_cast<const float4*>(p.ws + (long long)s * p.slab + row * p.n_pad + n0);
      float4 a = __ldcs(src), b = __ldcs(src + 1);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    const int nv = min(8, p.n_valid - n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = v[j];
      if (p.bias != nullptr && j < nv) { const int n = n0 + j; x += __ldg(p.bias + (p.bias_mod > 0 ? n % p.bias_mod : n)); }
      v[j] = apply_act(x, p.act);
    }
    const long long eoff = row * p.out_ld + n0;
    if (p.mask != nullptr) {
      for (int j = 0; j < nv; ++j) v[j] *= act_grad_from_output(__bfloat162float(p.mask[eoff + j]), p.mask_act);
    }
    if (p.out_kind == SG_OUT_BF16) {
      bf16* o = reinterpret_cast<bf16*>(p.out) + eoff;
      if (nv == 8 && (p.out_ld & 7) == 0) {
        uint4 hi;
        hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]); hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(o) = hi;
        if (p.planes == 2) {
          uint4 lo;
          lo.x = pack_bf16x2(v[0] - bf16lo_to_f(hi.x), v[1] - bf16hi_to_f(hi.x));
          lo.y = pack_bf16x2(v[2] - bf16lo_to_f(hi.y), v[3] - bf16hi_to_f(hi.y));
          lo.z = pack_bf16x2(v[4] - bf16lo_to_f(hi.z), v[5] - bf16hi_to_f(hi.z));
          lo.w = pack_bf16x2(v[6] - bf16lo_to_f(hi.w), v[7] - bf16hi_to_f(hi.w));
          *reinterpret_cast<uint4*>(o + p.out_ps) = lo;
        }
      } else {
        for (int j = 0; j < nv; ++j) {
          const bf16 h = __float2bfloat16_rn(v[j]);
          o[j] = h;
          if (p.planes == 2) o[p.out_ps + j] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
        }
      }
    } else {
      float* o = reinterpret_cast<float*>(p.out) + eoff;
      for (int j = 0; j < nv; ++j) o[j] = v[j];
    }
  }
}

// Tile configuration shared by sg_igemm and sg_igemm_plan.  `ws_ok`: a split-K workspace may be used.
// Split-K (fp32 partial slabs + sg_splitk_finish_kernel) is chosen when the output has too few tiles to fill the machine even at
// the widest N tile and K is long: Conv3d(128->256) on 8^3 -> 4^3 at B=64 is 32 row tiles x K = 8192; narrowing N to 64 (the
// alternative) fills 128 SMs with 128x64 MMAs that run at a third of the tensor rate (A-operand shared-memory reads per MMA).
struct TileCfg { int bn, mt, ksplit; size_t ws_bytes; long long out_rows; };

static int igemm_tiles(const sg_igemm_args* a, bool ws_ok, TileCfg* t) {
  const int sms = sg_num_sms();
  const long long row_tiles = (a->rows + 127) / 128;
  const int classes = (a->mode == SG_MODE_CONVT) ? 8 : 1;
  const int kchunks = a->k / 64;
  int bn = a->bn, mt = a->mt, ksplit = a->ksplit;
  t->out_rows = a->rows * classes;
  t->ws_bytes = 0;
  bool auto_split = false;
  if (bn <= 0) {
    bn = a->n_pad;
    if (bn > 256) {                       // largest divisor of n_pad that is a multiple of 16 and <= 256
      bn = 256;
      while (a->n_pad % bn) bn -= 16;
    }
    const long long items = row_tiles * classes * (a->n_pad / bn);
    const char* nsk = getenv("SG_B200_NO_SPLITK");
    const bool no_split = nsk && nsk[0] == '1';
    if (ksplit <= 0 && ws_ok && !no_split && a->out_kind != SG_OUT_F32_ATOMIC && items < sms && kchunks >= 32 && (bn & 31) == 0) {
      // less than one wave of tiles and a long K: split K, and choose the number of M sub-tiles per CTA together with the split so that
      // the machine is as full as possible (ties: fewer splits).  Conv3d(128->256) on the batched critic's 3 x 64 samples is 96 tiles:
      // unsplit 96 CTAs ran 91 us, two sub-tiles x 3 splits = 144 CTAs run 66 us (profiles/r02e_sweep_d3_cfgs.txt).
      int best_ks = 1, best_mt = 1; long long best_fill = 0;
      for (int mtc = 1; mtc <= ((mt <= 0) ? 2 : 1); ++mtc) {
        const int m_try = (mt > 0) ? mt : mtc;
        if (m_try * bn > 512) continue;
        const long long it0 = ((row_tiles + m_try - 1) / m_try) * classes * (a->n_pad / bn);
        if (it0 >= sms) continue;
        int ks = (int)(sms / it0);
        while (ks > 1 && kchunks / ks < 16) --ks;
        const long long fill = it0 * ks;
        if (fill > best_fill || (fill == best_fill && ks < best_ks)) { best_fill = fill; best_ks = ks; best_mt = m_try; }
      }
      if (best_ks > 1) { ksplit = best_ks; if (mt <= 0) mt = best_mt; auto_split = true; }
    }
    // not enough work for the machine: narrow the N tile (keeps tensor throughput, multiplies CTAs)
    if (!auto_split)
      while (bn > 64 && (bn % 32) == 0 && row_tiles * classes * (a->n_pad / bn) < sms) bn /= 2;
  }
  if (bn > 256 || (bn & 15) || a->n_pad % bn) return sg_fail(-20, "sg_igemm: bad bn");
  const int n_tiles = a->n_pad / bn;
  // two M sub-tiles share every B tile (halves the weight traffic out of L2) once there is more than one wave of work
  // (not when that would cost the accumulator double buffer: with bn = 256 two sub-tiles fill TMEM and the epilogue serialises)
  if (mt <= 0) mt = (row_tiles * classes * n_tiles > sms && bn * 4 <= 512) ? 2 : 1;
  if (mt < 1 || mt > 2 || mt * bn > 512) return sg_fail(-21, "sg_igemm: bad mt");
  if (ksplit <= 0) ksplit = 1;
  if (ksplit > kchunks) ksplit = kchunks;
  {  // no empty split
    int cps = (kchunks + ksplit - 1) / ksplit;
    ksplit = (kchunks + cps - 1) / cps;
  }
  if (ksplit > 1 && a->out_kind != SG_OUT_F32_ATOMIC) {
    if (!ws_ok) {
      if (auto_split) ksplit = 1; else return sg_fail(-22, "sg_igemm: split-K needs SG_OUT_F32_ATOMIC or a splitk_ws workspace (sg_igemm_plan)");
    } else {
      t->ws_bytes = (size_t)ksplit * (size_t)t->out_rows * (size_t)a->n_pad * sizeof(float);
    }
  }
  t->bn = bn; t->mt = mt; t->ksplit = ksplit;
  return 0;
}

// Shared-memory plan of the halo kernels: [header][a_blocks halo blocks][b_stages weight tiles][epilogue staging tile].
// Preference order (measured, profiles/r02e_trace_halo.txt): the staging tile (the direct tile store cost 18 % of the kernel), then a third
// halo block (a 46 KB block takes ~3400 clk to land, exactly the time its predecessor is consumed in: two blocks left the MMA thread
// waiting ~330 clk per block), as long as at least 4 weight stages remain.  Returns the dynamic shared memory size.
static size_t halo_smem_plan(IgemmP& p, unsigned b_tile, bool bf16_out) {
  const long long total = 227LL * 1024 - kSmemHeader;
  const char* ns = getenv("SG_B200_NO_EPI_STAGE");
  unsigned epi = (bf16_out && p.planes == 1 && (p.bn == 64 || p.bn == 128) && !(ns && ns[0] == '1')) ? 128u * (unsigned)p.bn * 2u : 0u;
  int blocks = 3;
  { const char* ab = getenv("SG_B200_HALO_BLOCKS"); if (ab && atoi(ab) == 2) blocks = 2; }
  if (blocks == 3 && (total - 3LL * p.blk_bytes - epi) / b_tile < 4) blocks = 2;
  if ((total - (long long)blocks * p.blk_bytes - epi) / b_tile < 3) epi = 0;
  p.a_blocks = blocks;
  p.b_stages = (int)std::min<long long>(kMaxStages, (total - (long long)blocks * p.blk_bytes - epi) / b_tile);
  if (p.b_stages < 0) p.b_stages = 0;
  const size_t used = kSmemHeader + (size_t)blocks * p.blk_bytes + (size_t)p.b_stages * b_tile;
  p.epi_off = epi ? (unsigned)used : 0u;
  return used + epi;
}

}  // namespace sg

using namespace sg;

extern "C" int sg_igemm(const sg_igemm_args* a, void* stream) {
  int rc = igemm_validate(a);
  if (rc) return rc;
  if (a->rows == 0) return 0;
  IgemmP p;
  memset(&p, 0, sizeof(p));
  p.mode = a->mode; p.planes = a->planes;
  p.a_ptr = (const char*)a->a.ptr; p.a_ps = a->a.plane_stride;
  p.aN = a->a.n; p.aD = a->a.d; p.aH = a->a.h; p.aW = a->a.w; p.aC = a->a.c;
  p.a2_ptr = (const char*)a->a2.ptr; p.a2_ps = a->a2.plane_stride; p.a2C = a->a2.c;
  p.rows = a->rows; p.kchunks = a->k / 64; p.n_pad = a->n_pad; p.n_valid = a->n_valid;
  p.classes = (a->mode == SG_MODE_CONVT) ? 8 : 1;
  p.b = (const char*)a->b_packed; p.bias = a->bias; p.act = a->act; p.bias_mod = a->bias_mod;
  p.mask = (const bf16*)a->mask; p.mask_ps = a->mask_plane_stride; p.mask_act = a->mask_act;
  p.out = (char*)a->out; p.out_ps = a->out_plane_stride; p.out_kind = a->out_kind; p.out_ld = a->out_ld;
  p.oD = a->out_d; p.oH = a->out_h; p.oW = a->out_w;
  p.err = sg_error_word();

  const int sms = sg_num_sms();
  const long long row_tiles = (a->rows + 127) / 128;
  // ---- tile configuration
  TileCfg tc;
  {
    const bool ws_ok = a->splitk_ws != nullptr;
    rc = igemm_tiles(a, ws_ok, &tc);
    if (rc) return rc;
    if (tc.ws_bytes > (size_t)a->splitk_ws_bytes) {
      if (a->ksplit > 1) return sg_fail(-22, "sg_igemm: splitk_ws too small (see sg_igemm_plan)");
      rc = igemm_tiles(a, false, &tc);            // auto split without room: fall back to the unsplit configuration
      if (rc) return rc;
    }
  }
  int bn = tc.bn, mt = tc.mt, ksplit = tc.ksplit;
  const int n_tiles = a->n_pad / bn;
  const bool split_ws = ksplit > 1 && a->out_kind != SG_OUT_F32_ATOMIC;
  // explicit/auto mt=2 that would leave fewer than 2 pipeline stages falls back to one M sub-tile
  if (mt == 2) {
    unsigned sb = ((unsigned)(2 * a->planes * kTileBytes + a->planes * bn * 128) + 1023u) & ~1023u;
    unsigned kt = (a->mode == SG_MODE_CONV || a->mode == SG_MODE_CONVT) ? (unsigned)((p.kchunks * 8 * 4 + 1023) & ~1023) : 0u;
    if ((227u * 1024u - kSmemHeader - kt) / sb < 2) mt = 1;
  }
  p.bn = bn; p.mt = mt; p.ksplit = ksplit;
  if (split_ws) {
    // K splits write fp32 partial slabs [ks][out rows][n_pad]; bias / activation / mask / bf16 conversion happen in the finish kernel
    p.out = (char*)a->splitk_ws; p.out_ps = 0; p.out_kind = SG_OUT_F32; p.out_ld = a->n_pad;
    p.ks_stride = tc.out_rows * (long long)a->n_pad;
    p.bias = nullptr; p.bias_mod = 0; p.act = ACT_NONE; p.mask = nullptr;
  }
  p.n_tiles = n_tiles;
  p.m_tiles = (int)((row_tiles + mt - 1) / mt);
  p.work_total = (long long)p.classes * n_tiles * p.m_tiles * ksplit;
  if (p.work_total >= (1LL << 31)) return sg_fail(-24, "sg_igemm: too many tiles");
  p.acc_slot = (bn + 31) & ~31;
  p.acc_bufs = (2 * mt * p.acc_slot <= 512) ? 2 : 1;
  p.a_stage_bytes = (unsigned)(mt * a->planes * kTileBytes);
  p.stage_bytes = p.a_stage_bytes + (unsigned)(a->planes * bn * 128);
  p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
  p.ktab_bytes = (a->mode == SG_MODE_CONV || a->mode == SG_MODE_CONVT) ? (unsigned)((p.kchunks * 8 * 4 + 1023) & ~1023) : 0u;
  const unsigned budget = 227u * 1024u - kSmemHeader - p.ktab_bytes;
  int stages = (int)(budget / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return sg_fail(-23, "sg_igemm: tile does not fit shared memory");
  // epilogue staging tile behind the stages (see epilogue_role): worth a stage where the output is the larger stream (bn <= 128)
  unsigned plain_epi = 0;
  {
    const char* ns = getenv("SG_B200_NO_EPI_STAGE");
    if (a->out_kind == SG_OUT_BF16 && !split_ws && a->planes == 1 && (bn == 64 || bn == 128) && !(ns && ns[0] == '1')) {
      const unsigned epi = 128u * (unsigned)bn * 2u;
      int st2 = budget > epi ? (int)((budget - epi) / p.stage_bytes) : 0;
      if (st2 > kMaxStages) st2 = kMaxStages;
      if (st2 >= std::min(stages, 4)) { stages = st2; plain_epi = epi; }
    }
  }
  p.stages = stages;
  p.epi_off = plain_epi ? (unsigned)(kSmemHeader + p.ktab_bytes + (size_t)stages * p.stage_bytes) : 0u;
  // ---- TMA gather: the whole A tile of a (tap, 64-channel) chunk is one tensor-map box
  p.use_tma = 0;
  {
    const char* no_tma = getenv("SG_B200_NO_TMA");
    const bool want = !(no_tma && no_tma[0] == '1') && a->mode != SG_MODE_PATCH && (a->a.c % 64) == 0 &&
                      (!a->a2.ptr || (a->a2.c % 64) == 0) && ((uintptr_t)a->a.ptr % 16) == 0 && a->rows < (1LL << 31);
    if (want) {
      bool ok = true;
      const uint32_t one[5] = {1, 1, 1, 1, 1};
      if (a->mode == SG_MODE_DENSE) {
        const uint32_t box[2] = {64, 128};
        for (int pl = 0; pl < a->planes && ok; ++pl) {
          uint64_t dims[2] = {(uint64_t)a->a.c, (uint64_t)a->rows};
          uint64_t str[1] = {(uint64_t)a->a.c * 2};
          ok = tma_make_map(&p.tmA[pl], (const char*)a->a.ptr + (size_t)pl * a->a.plane_stride * 2, 2, dims, str, box, one);
          if (ok && a->a2.ptr) {
            uint64_t dims2[2] = {(uint64_t)a->a2.c, (uint64_t)a->rows};
            uint64_t str2[1] = {(uint64_t)a->a2.c * 2};
            ok = tma_make_map(&p.tmA2[pl], (const char*)a->a2.ptr + (size_t)pl * a->a2.plane_stride * 2, 2, dims2, str2, box, one);
          }
        }
      } else {
        const bool conv = a->mode == SG_MODE_CONV;
        const int gx = conv ? a->a.w / 2 : a->a.w, gy = conv ? a->a.h / 2 : a->a.h, gz = conv ? a->a.d / 2 : a->a.d;
        TileBox tb;
        ok = tile_box(kTileRows, gx, gy, gz, &tb) && a->rows == (long long)a->a.n * gx * gy * gz;
        if (ok) {
          const uint64_t C = (uint64_t)a->a.c, W = (uint64_t)a->a.w, H = (uint64_t)a->a.h, D = (uint64_t)a->a.d;
          uint64_t dims[5] = {C, W, H, D, (uint64_t)a->a.n};
          uint64_t str[4] = {C * 2, W * C * 2, H * W * C * 2, D * H * W * C * 2};
          const uint32_t m = conv ? 2u : 1u;
          uint32_t box[5] = {64, m * tb.bx, m * tb.by, m * tb.bz, (uint32_t)tb.bn};
          uint32_t es[5] = {1, m, m, m, 1};
          for (int pl = 0; pl < a->planes && ok; ++pl)
            ok = tma_make_map(&p.tmA[pl], (const char*)a->a.ptr + (size_t)pl * a->a.plane_stride * 2, 5, dims, str, box, es);
          p.gx = gx; p.gy = gy; p.gz = gz; p.bx = tb.bx; p.by = tb.by; p.bz = tb.bz; p.bnn = tb.bn;
        }
      }
      p.use_tma = ok ? 1 : 0;
    }
  }
  { const char* dg = getenv("SG_B200_IGEMM_DIAG"); p.diag = dg ? atoi(dg) : 0; }
  // weight tiles through the TMA unit: the CTA-pair halo kernel only.  There the LSU path's ~23 B/clk per SM (bn = 128) resp. its latency
  // with two tiles in flight (bn = 64) starved the MMAs: Conv3d 64->128 45 -> 37 us, ConvT 128->64 61 -> 49 us at B = 64, 156 -> 121 us on
  // the batched critic (profiles/r02e_sweep_btma.txt).  Everywhere else it measured slower -- the single-CTA halo kernel and the plain
  // kernel (Conv3d 128->256 39.9 -> 42.0 us): profiles/r02e_sweep_cfgs*.txt.  SG_B200_B_TMA=0/1 forces it for measurements.
  { const char* e = getenv("SG_B200_B_TMA"); p.b_tma = e ? (e[0] != '0') : 0; }
  { const char* e = getenv("SG_B200_B_DEPTH"); p.b_depth = e ? std::max(1, atoi(e)) : 1; }
  // ---- halo-reuse variant (see sg_igemm_halo_kernel): 8 x 8 x gz row grids, bf16
  {
    const bool conv = a->mode == SG_MODE_CONV;
    const char* no_halo = getenv("SG_B200_NO_HALO");
    if (p.use_tma && !(no_halo && no_halo[0] == '1') && a->planes == 1 && (conv || a->mode == SG_MODE_CONVT) && p.gx == 8 && p.gy == 8 &&
        p.bx == 8 && p.by == 8 && is_pow2(p.gz) && p.gz % (2 * mt) == 0 && ksplit == 1 && (a->a.c % 64) == 0 && (p.diag == 0 || p.diag == 128)) {
      const uint64_t C = (uint64_t)a->a.c, W = (uint64_t)a->a.w, H = (uint64_t)a->a.h, D = (uint64_t)a->a.d;
      uint64_t dims[5] = {C, W, H, D, (uint64_t)a->a.n};
      uint64_t str[4] = {C * 2, W * C * 2, H * W * C * 2, D * H * W * C * 2};
      const uint32_t m = conv ? 2u : 1u;
      const uint32_t zp = (uint32_t)(2 * mt + 1);
      uint32_t box[5] = {64, m * 9, m * 8, m * zp, 1};
      uint32_t es[5] = {1, m, m, m, 1};
      p.blk_bytes = 9u * 8u * zp * 128u;
      halo_smem_plan(p, (unsigned)bn * 128u, a->out_kind == SG_OUT_BF16 && !split_ws);
      if (p.b_stages >= 3 && tma_make_map(&p.tmH, a->a.ptr, 5, dims, str, box, es)) p.halo = 1;
    }
  }
  // CTA-pair variant of the halo kernel: an even number of M tiles per (class, N tile), whole pairs inside one sample's z range
  if (p.halo) {
    const char* no_pair = getenv("SG_B200_NO_PAIR");
    const long long pair_tiles = p.m_tiles;
    if (!(no_pair && no_pair[0] == '1') && (pair_tiles & 1) == 0 && (bn & 31) == 0 && bn >= 32 && p.gz % (4 * mt) == 0) {
      IgemmP q = p;
      q.pair = 1;
      q.m_tiles = p.m_tiles / 2;
      q.work_total = (long long)q.classes * q.n_tiles * q.m_tiles;
      { const char* e = getenv("SG_B200_B_TMA"); q.b_tma = e ? (e[0] != '0') : 1; }
      const size_t psmem = halo_smem_plan(q, (unsigned)(bn / 2) * 128u, a->out_kind == SG_OUT_BF16 && !split_ws);
      if (q.b_tma) {
        const uint64_t brows = (uint64_t)q.classes * (uint64_t)q.kchunks * (uint64_t)q.n_pad;
        uint64_t bd[2] = {64, brows};
        uint64_t bs[1] = {128};
        uint32_t bb[2] = {64, (uint32_t)(bn / 2)};
        uint32_t be[2] = {1, 1};
        if (brows >= (1ull << 31) || !tma_make_map(&q.tmB, a->b_packed, 2, bd, bs, bb, be, false)) q.b_tma = 0;
      }
      static PerDevice pattr;
      if (pattr.first()) {
        cudaError_t e = cudaFuncSetAttribute(sg_igemm_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
        pattr.done();
      }
      const int clusters = (int)std::min<long long>(q.work_total, sms / 2);
      sg_igemm_halo2_kernel<<<2 * clusters, kIgemmThreads, psmem, (cudaStream_t)stream>>>(q);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
      sg_count_launch();
      return 0;
    }
  }
  if (p.halo) {
    const size_t hsmem = halo_smem_plan(p, (unsigned)bn * 128u, a->out_kind == SG_OUT_BF16 && !split_ws);
    static PerDevice hattr;
    if (hattr.first()) {
      cudaError_t e = cudaFuncSetAttribute(sg_igemm_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
      hattr.done();
    }
    const int hgrid = (int)std::min<long long>(p.work_total, sms);
    sg_igemm_halo_kernel<<<hgrid, kIgemmThreads, hsmem, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    sg_count_launch();
    return 0;
  }
  p.epi_off = plain_epi ? (unsigned)(kSmemHeader + p.ktab_bytes + (size_t)stages * p.stage_bytes) : 0u;      // (the halo plans above overwrote it)
  const size_t smem = kSmemHeader + p.ktab_bytes + (size_t)stages * p.stage_bytes + plain_epi;
  static PerDevice attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(sg_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    attr.done();
  }
  int grid = (int)std::min<long long>(p.work_total, sms);
  { const char* gg = getenv("SG_B200_IGEMM_GRID"); if (gg && atoi(gg) > 0) grid = std::min(grid, atoi(gg)); }   // measurement only
  sg_igemm_kernel<<<grid, kIgemmThreads, smem, (cudaStream_t)stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  sg_count_launch();
  if (split_ws) {
    FinishP f;
    f.ws = (const float*)a->splitk_ws; f.ksplit = ksplit; f.slab = p.ks_stride; f.rows = tc.out_rows; f.n_pad = a->n_pad; f.n_valid = a->n_valid;
    f.bias = a->bias; f.bias_mod = a->bias_mod; f.act = a->act;
    f.mask = (const bf16*)a->mask; f.mask_ps = a->mask_plane_stride; f.mask_act = a->mask_act;
    f.out = (char*)a->out; f.out_ps = a->out_plane_stride; f.out_kind = a->out_kind; f.out_ld = a->out_ld; f.planes = a->planes;
    const long long total = tc.out_rows * (a->n_pad >> 3);
    const int fgrid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
    sg_splitk_finish_kernel<<<fgrid, 256, 0, (cudaStream_t)stream>>>(f);
    e = cudaGetLastError();
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    sg_count_launch();
  }
  return 0;
}

// Workspace (bytes of fp32 partial slabs) sg_igemm would use for these arguments when `splitk_ws` is provided; 0 = no split-K.
extern "C" int sg_igemm_plan(const sg_igemm_args* a, size_t* ws_bytes) {
  if (!ws_bytes) return sg_fail(-1, "sg_igemm_plan: null");
  *ws_bytes = 0;
  int rc = igemm_validate(a);
  if (rc) return rc;
  if (a->rows == 0) return 0;
  TileCfg tc;
  rc = igemm_tiles(a, true, &tc);
  if (rc) return rc;
  *ws_bytes = tc.ws_bytes;
  return 0;
}

// diagnostics: clock64 trace of CTA 0 of the last sg_igemm launched with SG_B200_IGEMM_DIAG & 128 (synchronises)
extern "C" int sg_debug_igemm_trace(long long* host_out, int n) {
  if (!host_out || n <= 0 || n > 3 * kTraceLen) return sg_fail(-1, "sg_debug_igemm_trace: bad args");
  cudaError_t e = cudaMemcpyFromSymbol(host_out, g_igemm_trace, sizeof(long long) * (size_t)n, 0, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  return 0;
}
