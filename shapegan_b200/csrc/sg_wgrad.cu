// Weight-gradient GEMM on tcgen05:  P[m, n] = sum_rows A[row, m] * gather(B)[row (+tap), n]
// Both operands are MN-major in shared memory (the reduction dimension -- voxels/points/samples -- is the slow
// axis of the NDHWC planes), gathered by cp.async into 128B-swizzled [rows][64 ch] atoms.  The 128 x (<=256)
// fp32 accumulator lives in TMEM; each CTA owns one (m-tile, n-group, row-split) and writes an fp32 partial that
// sg_wgrad_reduce folds (deterministically, no atomics) into the torch-layout .grad.
#include <algorithm>
#include <cstring>

#include <cstdlib>

#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_tma.h"

namespace sg {

constexpr int kWgThreads = 288;
constexpr int kWgMaxStages = 4;
constexpr int kWgHeader = 1024;
constexpr int kAtomsA = 2;     // 128 M rows  = 2 atoms of 64 channels
constexpr int kAtomsB = 4;     // <=256 N cols = 4 atoms of 64 columns

struct WgradP {
  int b_mode, planes;
  const char* a_ptr; long long a_ps; int Ca;
  const char* b_ptr; long long b_ps; int bD, bH, bW, Cb;
  long long rows; int row_tiles, rs;          // rs = rows per stage (128 / planes)
  int taps, n_total, n_groups, m_tiles, m_pad, ksplit, merge_n;
  float* partials;
  float* bias_partials;      // optional [ksplit][m_pad]: column sums of A (= the bias gradient when A is dY), one extra N = 16 MMA per K step
                             // against a constant tile of ones -- the activations are not read a second time for it
  int stages; unsigned stage_bytes, tile_bytes;
  long long work_total;
  int* err;
  int use_tma, gx, gy, gz;
  int pair;                   // CTA pairs (2-CTA cluster): the two M tiles of one (N group, row split) share every B tile through TMA multicast
  CUtensorMap tmA[2], tmB[2];
};

struct WgHeader {
  uint64_t full[kWgMaxStages];
  uint64_t empty[kWgMaxStages];
  uint64_t accfull[2];
  uint64_t accempty[2];
  uint32_t tmem_base;
};

// PAIR: launched as 2-CTA clusters.  The pair owns M tiles (2j, 2j+1) of one (N group, row split): both need the SAME gathered B tiles, so
// each CTA issues only half of the B boxes, multicast into both CTAs' stages (the weight-gradient GEMM is bound by the bytes / TMA rows
// it pulls per stage: 2 A atoms + 4 B atoms per CTA become 2 + 2).  A stage is recycled when BOTH CTAs' MMAs have read it: the commit
// that frees it is multicast to both empty barriers (count 2).
template <bool PAIR>
__global__ void __launch_bounds__(kWgThreads, 1) sg_wgrad_kernel(const __grid_constant__ WgradP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  WgHeader* hdr = reinterpret_cast<WgHeader*>(smem);
  uint8_t* stage0 = smem + kWgHeader;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const long long w0 = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long wstep = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&hdr->full[s], p.use_tma ? 1 : 128); mbar_init(&hdr->empty[s], PAIR ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&hdr->accfull[i], 1); mbar_init(&hdr->accempty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (tid == 0) atomicExch(p.err, kErrSmemAlign);
    __trap();
  }
  const int tps = (p.row_tiles + p.ksplit - 1) / p.ksplit;    // row tiles (stages) per split
  const int rpi = p.rs / 16;                                   // rows handled per producer thread per atom
  // bias mode: one accumulator buffer (columns 0..255) + the 16 bias columns at 256; a 2 KB block of bf16 ones behind the stages
  // serves every K step as the MN-major B operand (all elements equal: the swizzle is irrelevant)
  const bool bias_mode = p.bias_partials != nullptr;
  uint8_t* ones_tile = stage0 + (size_t)S * p.stage_bytes;
  if (bias_mode) {
    for (int i = tid; i < 512; i += kWgThreads) reinterpret_cast<uint32_t*>(ones_tile)[i] = 0x3f803f80u;
    fence_proxy_async();
    __syncthreads();
  }

  if (warp < 4 && p.use_tma) {
    // ================================================================ TMA PRODUCER (warp 0, one elected lane issues): every atom is one box
    if (warp == 0) {
      if (elect_one()) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmB[0]);
      }
      int s = 0; uint32_t ph = 0;
      const int lgx = 31 - __clz(max(p.gx, 1)), lgy = 31 - __clz(max(p.gy, 1)), lgz = 31 - __clz(max(p.gz, 1));
      for (long long w = w0; w < p.work_total; w += wstep) {
        const uint32_t w32 = (uint32_t)w;
        const int ks = (int)(w32 % (uint32_t)p.ksplit);
        const int ng = (int)((w32 / (uint32_t)p.ksplit) % (uint32_t)p.n_groups);
        const int mtile = (int)(w32 / ((uint32_t)p.ksplit * (uint32_t)p.n_groups)) * (PAIR ? 2 : 1) + (int)crank;
        const int m0 = mtile * 128;
        const int nb_atoms = min(kAtomsB, (p.n_total - ng * 256) / 64);
        const uint32_t tx_bytes = (uint32_t)((kAtomsA + nb_atoms) * p.planes) * p.tile_bytes;
        int bcb[kAtomsB], bkx[kAtomsB], bky[kAtomsB], bkz[kAtomsB];
        for (int j = 0; j < kAtomsB; ++j) {
          const int n = ng * 256 + j * 64;
          const int tap = n / p.Cb;
          bcb[j] = n - tap * p.Cb; bkx[j] = (tap & 3) - 1; bky[j] = ((tap >> 2) & 3) - 1; bkz[j] = (tap >> 4) - 1;
        }
        const int t0 = ks * tps, t1 = min(p.row_tiles, t0 + tps);
        for (int rt = t0; rt < t1; ++rt) {
          mbar_wait(&hdr->empty[s], ph ^ 1, p.err);
          uint64_t* bar = &hdr->full[s];
          const uint32_t base = smem_u32(stage0 + (size_t)s * p.stage_bytes);
          const uint32_t row0 = (uint32_t)rt * (uint32_t)p.rs;
          const uint32_t bbase = base + (uint32_t)(kAtomsA * p.planes) * p.tile_bytes;
          int x0 = 0, y0 = 0, z0 = 0, n0 = 0;
          if (p.b_mode == SG_MODE_CONV) {
            x0 = 2 * (int)(row0 & (uint32_t)(p.gx - 1));
            y0 = 2 * (int)((row0 >> lgx) & (uint32_t)(p.gy - 1));
            z0 = 2 * (int)((row0 >> (lgx + lgy)) & (uint32_t)(p.gz - 1));
            n0 = (int)(row0 >> (lgx + lgy + lgz));
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(bar, tx_bytes);
            for (int at = 0; at < kAtomsA; ++at)
              for (int pl = 0; pl < p.planes; ++pl)
                tma_load_2d(base + (uint32_t)(at * p.planes + pl) * p.tile_bytes, &p.tmA[pl], m0 + at * 64, (int)row0, bar);
            for (int j = 0; j < nb_atoms; ++j) {
              if (PAIR && (uint32_t)(j & 1) != crank) continue;       // the peer issues this atom, multicast into both CTAs
              for (int pl = 0; pl < p.planes; ++pl) {
                const uint32_t dst = bbase + (uint32_t)(j * p.planes + pl) * p.tile_bytes;
                if (PAIR) {
                  if (p.b_mode == SG_MODE_DENSE) tma_load_2d_mc(dst, &p.tmB[pl], bcb[j], (int)row0, bar, 3);
                  else tma_load_5d_mc(dst, &p.tmB[pl], bcb[j], x0 + bkx[j], y0 + bky[j], z0 + bkz[j], n0, bar, 3);
                } else {
                  if (p.b_mode == SG_MODE_DENSE) tma_load_2d(dst, &p.tmB[pl], bcb[j], (int)row0, bar);
                  else tma_load_5d(dst, &p.tmB[pl], bcb[j], x0 + bkx[j], y0 + bky[j], z0 + bkz[j], n0, bar);
                }
              }
            }
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    // ================================================================ PRODUCERS (cp.async gather fallback)
    const int g = tid & 7, rb = tid >> 3;
    int s = 0; uint32_t ph = 0;
    const int lag = max(1, S - 2);
    int pending = 0, oldest = 0;
    const int oW = p.bW >> 1, oH = p.bH >> 1, oD = p.bD >> 1;   // CONV/PATCH: rows enumerate the stride-2 output grid
    for (long long w = w0; w < p.work_total; w += wstep) {
      const int ks = (int)(w % p.ksplit);
      const int ng = (int)((w / p.ksplit) % p.n_groups);
      const int mtile = (int)(w / ((long long)p.ksplit * p.n_groups));
      const int m0 = mtile * 128;
      const int nb_atoms = min(kAtomsB, (p.n_total - ng * 256) / 64);
      // per-thread column descriptors of the B atoms: tap offsets + channel
      int bc[kAtomsB], bdd[kAtomsB], bdh[kAtomsB], bdw[kAtomsB];
#pragma unroll
      for (int j = 0; j < kAtomsB; ++j) {
        const int n = ng * 256 + j * 64 + g * 8;
        const int tap = n / p.Cb;
        bc[j] = n - tap * p.Cb; bdd[j] = tap >> 4; bdh[j] = (tap >> 2) & 3; bdw[j] = tap & 3;
      }
      const int t0 = ks * tps, t1 = min(p.row_tiles, t0 + tps);
      for (int rt = t0; rt < t1; ++rt) {
        mbar_wait(&hdr->empty[s], ph ^ 1, p.err);
        uint8_t* st = stage0 + (size_t)s * p.stage_bytes;
        const uint32_t base = smem_u32(st);
        const bool patch8 = (p.b_mode == SG_MODE_PATCH) && ((p.bW & 15) == 0);
        if (patch8 && rb * 8 < p.rs) {
          // B atom 0 = im2col patches: 8 consecutive output-x rows of piece g per thread (see patch_fill8)
          const long long gr0 = (long long)rt * p.rs + rb * 8;
          const bool v = gr0 < p.rows;
          const uint32_t r32 = (uint32_t)(v ? gr0 : 0);
          const int ow0 = (int)(r32 % (uint32_t)oW); uint32_t t2 = r32 / (uint32_t)oW;
          const int oh = (int)(t2 % (uint32_t)oH); t2 /= (uint32_t)oH;
          const int od = (int)(t2 % (uint32_t)oD); const uint32_t n = t2 / (uint32_t)oD;
          uint8_t* tile = st + (size_t)(kAtomsA * p.planes) * p.tile_bytes;
          patch_fill8(reinterpret_cast<const float*>(p.b_ptr) + (size_t)n * p.bD * p.bH * p.bW, p.bD, p.bH, p.bW, od, oh, ow0, g, v, tile,
                      p.planes == 2 ? tile + p.tile_bytes : nullptr, rb * 8);
        }
        for (int i = 0; i < rpi; ++i) {
          const int row = rb + 16 * i;
          const long long gr = (long long)rt * p.rs + row;
          const bool rvalid = gr < p.rows;
          // ---- A atoms (dense rows of ca channels)
#pragma unroll
          for (int at = 0; at < kAtomsA; ++at) {
            const int ch = m0 + at * 64 + g * 8;
            const bool v = rvalid && ch < p.Ca;
            const long long off = v ? gr * p.Ca + ch : 0;
            const uint32_t dst = base + (uint32_t)(at * p.planes) * p.tile_bytes + sw128(row, g);
            cp_async16(dst, p.a_ptr + off * 2, v ? 16u : 0u);
            if (p.planes == 2) cp_async16(dst + p.tile_bytes, p.a_ptr + (off + p.a_ps) * 2, v ? 16u : 0u);
          }
          // ---- B atoms
          const uint32_t bbase = base + (uint32_t)(kAtomsA * p.planes) * p.tile_bytes;
          if (p.b_mode == SG_MODE_PATCH && patch8) {
            // filled above
          } else if (p.b_mode == SG_MODE_PATCH) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (rvalid) {
              int ow = (int)(gr % oW); long long t = gr / oW;
              int oh = (int)(t % oH); t /= oH;
              int od = (int)(t % oD); long long n = t / oD;
              const float* vol = reinterpret_cast<const float*>(p.b_ptr) + n * p.bD * p.bH * p.bW;
              const int d = 2 * od - 1 + (g >> 1), w0 = 2 * ow - 1;
              if (d >= 0 && d < p.bD) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int h = 2 * oh - 1 + (g & 1) * 2 + hh;
                  if (h >= 0 && h < p.bH) {
                    const float* rowp = vol + ((size_t)d * p.bH + h) * p.bW;
#pragma unroll
                    for (int ww = 0; ww < 4; ++ww) {
                      const int x = w0 + ww;
                      if (x >= 0 && x < p.bW) v[hh * 4 + ww] = __ldg(rowp + x);
                    }
                  }
                }
              }
            }
            uint4 hi;
            hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]);
            hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(st + (size_t)(kAtomsA * p.planes) * p.tile_bytes + sw128(row, g)) = hi;
            if (p.planes == 2) {
              uint4 lo;
              lo.x = pack_bf16x2(v[0] - bf16lo_to_f(hi.x), v[1] - bf16hi_to_f(hi.x));
              lo.y = pack_bf16x2(v[2] - bf16lo_to_f(hi.y), v[3] - bf16hi_to_f(hi.y));
              lo.z = pack_bf16x2(v[4] - bf16lo_to_f(hi.z), v[5] - bf16hi_to_f(hi.z));
              lo.w = pack_bf16x2(v[6] - bf16lo_to_f(hi.w), v[7] - bf16hi_to_f(hi.w));
              *reinterpret_cast<uint4*>(st + (size_t)(kAtomsA * p.planes + 1) * p.tile_bytes + sw128(row, g)) = lo;
            }
          } else {
            int n = 0, od = 0, oh = 0, ow = 0;
            if (p.b_mode == SG_MODE_CONV && rvalid) {
              ow = (int)(gr % oW); long long t = gr / oW;
              oh = (int)(t % oH); t /= oH;
              od = (int)(t % oD); n = (int)(t / oD);
            }
#pragma unroll
            for (int j = 0; j < kAtomsB; ++j) {
              if (j >= nb_atoms) break;
              long long off = 0; bool v = false;
              if (rvalid) {
                if (p.b_mode == SG_MODE_DENSE) { off = gr * p.Cb + bc[j]; v = true; }
                else {
                  const int d = 2 * od - 1 + bdd[j], h = 2 * oh - 1 + bdh[j], x = 2 * ow - 1 + bdw[j];
                  if (d >= 0 && d < p.bD && h >= 0 && h < p.bH && x >= 0 && x < p.bW) {
                    off = ((((long long)n * p.bD + d) * p.bH + h) * p.bW + x) * p.Cb + bc[j]; v = true;
                  }
                }
              }
              const uint32_t dst = bbase + (uint32_t)(j * p.planes) * p.tile_bytes + sw128(row, g);
              cp_async16(dst, p.b_ptr + off * 2, v ? 16u : 0u);
              if (p.planes == 2) cp_async16(dst + p.tile_bytes, p.b_ptr + (off + p.b_ps) * 2, v ? 16u : 0u);
            }
          }
        }
        cp_async_commit();
        if (++pending > lag) {
          cp_async_wait_dyn(lag); fence_proxy_async(); mbar_arrive(&hdr->full[oldest]);
          if (++oldest == S) oldest = 0;
          --pending;
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
    int s = 0; uint32_t ph = 0; int it = 0;
    const uint32_t atom_stride = (uint32_t)p.planes * p.tile_bytes;     // LBO: next 64-wide M/N atom (same plane)
    for (long long w = w0; w < p.work_total; w += wstep, ++it) {
      const int ks = (int)(w % p.ksplit);
      const int ng = (int)((w / p.ksplit) % p.n_groups);
      const int nb_atoms = min(kAtomsB, (p.n_total - ng * 256) / 64);
      const int ab = bias_mode ? 0 : (it & 1);
      const uint32_t aph = (uint32_t)(bias_mode ? (it & 1) : ((it >> 1) & 1));
      mbar_wait(&hdr->accempty[ab], aph ^ 1, p.err);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)(ab * 256);
      const uint32_t idesc_full = umma_idesc(128, nb_atoms * 64, true, true);
      const uint32_t idesc_atom = umma_idesc(128, 64, true, true);
      const uint32_t idesc_bias = umma_idesc(128, 16, true, true);
      const bool do_bias = bias_mode && ng == 0;
      const int t0 = ks * tps, t1 = min(p.row_tiles, t0 + tps);
      for (int rt = t0; rt < t1; ++rt) {
        mbar_wait(&hdr->full[s], ph, p.err);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_base = smem_u32(stage0 + (size_t)s * p.stage_bytes);
          const uint32_t b_base = a_base + (uint32_t)(kAtomsA * p.planes) * p.tile_bytes;
          const int ksteps = p.rs / 16;
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint32_t acc = (rt > t0 || kk > 0) ? 1u : 0u;
            const uint64_t da = umma_desc(a_base + kk * 2048, atom_stride, 1024);
            const uint64_t da_lo = umma_desc(a_base + p.tile_bytes + kk * 2048, atom_stride, 1024);
            if (p.merge_n) {
              const uint64_t db = umma_desc(b_base + kk * 2048, atom_stride, 1024);
              umma_bf16(d_addr, da, db, idesc_full, acc);
              if (p.planes == 2) {
                const uint64_t db_lo = umma_desc(b_base + p.tile_bytes + kk * 2048, atom_stride, 1024);
                umma_bf16(d_addr, da, db_lo, idesc_full, 1u);
                umma_bf16(d_addr, da_lo, db, idesc_full, 1u);
              }
            } else {
              for (int j = 0; j < nb_atoms; ++j) {
                const uint32_t bj = b_base + (uint32_t)j * atom_stride;
                const uint64_t db = umma_desc(bj + kk * 2048, atom_stride, 1024);
                umma_bf16(d_addr + j * 64, da, db, idesc_atom, acc);
                if (p.planes == 2) {
                  const uint64_t db_lo = umma_desc(bj + p.tile_bytes + kk * 2048, atom_stride, 1024);
                  umma_bf16(d_addr + j * 64, da, db_lo, idesc_atom, 1u);
                  umma_bf16(d_addr + j * 64, da_lo, db, idesc_atom, 1u);
                }
              }
            }
            if (do_bias) {
              const uint64_t dones = umma_desc(smem_u32(ones_tile), atom_stride, 1024);
              umma_bf16(tmem_base + 256, da, dones, idesc_bias, acc);
              if (p.planes == 2) umma_bf16(tmem_base + 256, da_lo, dones, idesc_bias, 1u);
            }
          }
          if (PAIR) umma_commit_mc(&hdr->empty[s], 3); else umma_commit(&hdr->empty[s]);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(&hdr->accfull[ab]);
      __syncwarp();
    }
  } else {
    // ================================================================ EPILOGUE: TMEM -> fp32 partials
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    int it = 0;
    for (long long w = w0; w < p.work_total; w += wstep, ++it) {
      const int ks = (int)(w % p.ksplit);
      const int ng = (int)((w / p.ksplit) % p.n_groups);
      const int mtile = (int)(w / ((long long)p.ksplit * p.n_groups)) * (PAIR ? 2 : 1) + (int)crank;
      const int nb_atoms = min(kAtomsB, (p.n_total - ng * 256) / 64);
      const int ab = bias_mode ? 0 : (it & 1);
      const uint32_t aph = (uint32_t)(bias_mode ? (it & 1) : ((it >> 1) & 1));
      mbar_wait(&hdr->accfull[ab], aph, p.err);
      tc_fence_after();
      float* orow = p.partials + ((size_t)ks * p.m_pad + (size_t)mtile * 128 + trow) * p.n_total + (size_t)ng * 256;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 256);
      // an empty split (no row tiles) never issued an MMA: its partial is defined as zero
      const int t0 = ks * tps, t1 = min(p.row_tiles, t0 + tps);
      for (int c0 = 0; c0 < nb_atoms * 64; c0 += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(t_addr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          if (t1 <= t0) v = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(orow + c0 + j) = v;
        }
      }
      if (bias_mode && ng == 0) {
        uint32_t rb16[16];
        __syncwarp();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + 256u, rb16);
        tmem_ld_wait();
        p.bias_partials[(size_t)ks * p.m_pad + (size_t)mtile * 128 + trow] = (t1 <= t0) ? 0.f : __uint_as_float(rb16[0]);
      }
      tc_fence_before();
      mbar_arrive(&hdr->accempty[ab]);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // the peer may still multicast into / commit onto this CTA's shared memory
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------ reduce + scatter
__global__ void sg_wgrad_reduce_kernel(const sg_wgrad_reduce_args a) {
  const long long n_total = (long long)a.taps * a.cb;
  const long long total = (long long)a.m_valid * n_total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / n_total);
    const int n = (int)(i - (long long)m * n_total);
    const int tap = n / a.cb, c = n - tap * a.cb;
    if (a.c_valid > 0 && c >= a.c_valid) continue;
    float acc = 0.f;
    const float* p = a.partials + (size_t)m * n_total + n;
    const size_t sstride = (size_t)a.m_pad * n_total;
    for (int s0 = 0; s0 < a.ksplit; s0 += 8) {        // eight loads in flight, added in split order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (s0 + u < a.ksplit) ? p[(size_t)(s0 + u) * sstride] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    float* g = a.grad + (long long)m * a.sm + (long long)tap * a.st + (long long)c * a.sc;
    acc *= a.scale;
    *g = a.accumulate ? (*g + acc) : acc;
  }
}

static int wgrad_geometry(const sg_wgrad_args* a, WgradP& p) {
  if (!a) return sg_fail(-1, "sg_wgrad: null args");
  if (a->planes != 1 && a->planes != 2) return sg_fail(-2, "sg_wgrad: planes must be 1 or 2");
  if (a->b_mode != SG_MODE_DENSE && a->b_mode != SG_MODE_CONV && a->b_mode != SG_MODE_PATCH) return sg_fail(-3, "sg_wgrad: bad b_mode");
  if (a->a.c <= 0 || (a->a.c & 7)) return sg_fail(-4, "sg_wgrad: a.c must be a positive multiple of 8");
  memset(&p, 0, sizeof(p));
  p.b_mode = a->b_mode; p.planes = a->planes;
  p.Ca = a->a.c; p.Cb = a->b.c;
  p.bD = a->b.d; p.bH = a->b.h; p.bW = a->b.w;
  if (a->b_mode == SG_MODE_PATCH) {
    if (a->b.c != 1) return sg_fail(-5, "sg_wgrad: PATCH needs b.c == 1");
    p.taps = 1; p.Cb = 64; p.n_total = 64;                 // the 64 taps play the role of channels
  } else {
    if (a->b.c <= 0 || (a->b.c & 7)) return sg_fail(-6, "sg_wgrad: b.c must be a positive multiple of 8");
    p.taps = (a->b_mode == SG_MODE_CONV) ? 64 : 1;
    p.n_total = p.taps * a->b.c;
    if (p.n_total & 63) return sg_fail(-7, "sg_wgrad: taps*b.c must be a multiple of 64");
  }
  if (a->b_mode != SG_MODE_DENSE && ((a->b.d | a->b.h | a->b.w) & 1)) return sg_fail(-8, "sg_wgrad: gathered dims must be even");
  p.rows = a->rows;
  p.rs = 128 / a->planes;
  p.row_tiles = (int)((a->rows + p.rs - 1) / p.rs);
  p.m_tiles = (p.Ca + 127) / 128;
  p.m_pad = p.m_tiles * 128;
  p.n_groups = (p.n_total + 255) / 256;
  p.tile_bytes = (unsigned)(p.rs * 128);
  p.stage_bytes = (unsigned)((kAtomsA + kAtomsB) * a->planes) * p.tile_bytes;
  p.stages = std::min<int>(kWgMaxStages, (int)((227u * 1024u - kWgHeader - 2048u) / p.stage_bytes));   // 2 KB: the bias mode's tile of ones
  if (p.stages < 2) return sg_fail(-9, "sg_wgrad: stage does not fit");
  return 0;
}

}  // namespace sg

using namespace sg;

extern "C" int sg_wgrad_plan(sg_wgrad_args* a, size_t* workspace_bytes) {
  WgradP p;
  int rc = wgrad_geometry(a, p);
  if (rc) return rc;
  int ksplit = a->ksplit;
  if (ksplit <= 0) {
    const int sms = sg_num_sms();
    const int base = p.m_tiles * p.n_groups;
    ksplit = std::max(1, sms / base);                 // floor: all work items in ONE wave (a 2nd, mostly empty wave doubles the time)
    ksplit = std::min(ksplit, std::max(1, p.row_tiles));
    // keep the fp32 partial workspace modest (<= 256 MB)
    while (ksplit > 1 && (size_t)ksplit * p.m_pad * p.n_total * 4 > (256u << 20)) --ksplit;
  }
  if (p.row_tiles > 0) {   // no empty split
    int tps = (p.row_tiles + ksplit - 1) / ksplit;
    ksplit = (p.row_tiles + tps - 1) / tps;
  } else {
    ksplit = 1;
  }
  a->ksplit_out = ksplit;
  if (workspace_bytes) *workspace_bytes = (size_t)ksplit * p.m_pad * p.n_total * sizeof(float);
  a->bias_ws_floats = ksplit * p.m_pad;
  return 0;
}

extern "C" int sg_wgrad(const sg_wgrad_args* a, void* stream) {
  WgradP p;
  int rc = wgrad_geometry(a, p);
  if (rc) return rc;
  if (!a->a.ptr || !a->b.ptr || !a->partials) return sg_fail(-10, "sg_wgrad: null tensor");
  if (a->ksplit_out <= 0) return sg_fail(-11, "sg_wgrad: call sg_wgrad_plan first");
  p.a_ptr = (const char*)a->a.ptr; p.a_ps = a->a.plane_stride;
  p.b_ptr = (const char*)a->b.ptr; p.b_ps = a->b.plane_stride;
  p.ksplit = a->ksplit_out; p.merge_n = a->merge_n;
  p.partials = a->partials;
  p.bias_partials = a->bias_partials;
  p.work_total = (long long)p.m_tiles * p.n_groups * p.ksplit;
  p.err = sg_error_word();
  p.use_tma = 0;
  {
    const char* no_tma = getenv("SG_B200_NO_TMA");
    const bool want = !(no_tma && no_tma[0] == '1') && a->b_mode != SG_MODE_PATCH && (p.Ca % 64) == 0 && (p.Cb % 64) == 0 &&
                      ((uintptr_t)a->a.ptr % 16) == 0 && ((uintptr_t)a->b.ptr % 16) == 0 && a->rows < (1LL << 31);
    if (want) {
      bool ok = true;
      const uint32_t one[5] = {1, 1, 1, 1, 1};
      const uint32_t box2[2] = {64, (uint32_t)p.rs};
      TileBox tb = {0, 0, 0, 0};
      if (a->b_mode == SG_MODE_CONV) {
        p.gx = a->b.w / 2; p.gy = a->b.h / 2; p.gz = a->b.d / 2;
        ok = tile_box(p.rs, p.gx, p.gy, p.gz, &tb) && a->rows == (long long)a->b.n * p.gx * p.gy * p.gz;
      }
      for (int pl = 0; pl < a->planes && ok; ++pl) {
        uint64_t dA[2] = {(uint64_t)p.Ca, (uint64_t)a->rows};
        uint64_t sA[1] = {(uint64_t)p.Ca * 2};
        ok = tma_make_map(&p.tmA[pl], p.a_ptr + (size_t)pl * p.a_ps * 2, 2, dA, sA, box2, one);
        if (!ok) break;
        if (a->b_mode == SG_MODE_DENSE) {
          uint64_t dB[2] = {(uint64_t)p.Cb, (uint64_t)a->rows};
          uint64_t sB[1] = {(uint64_t)p.Cb * 2};
          ok = tma_make_map(&p.tmB[pl], p.b_ptr + (size_t)pl * p.b_ps * 2, 2, dB, sB, box2, one);
        } else {
          const uint64_t C = (uint64_t)p.Cb, W = (uint64_t)a->b.w, H = (uint64_t)a->b.h, D = (uint64_t)a->b.d;
          uint64_t dB[5] = {C, W, H, D, (uint64_t)a->b.n};
          uint64_t sB[4] = {C * 2, W * C * 2, H * W * C * 2, D * H * W * C * 2};
          uint32_t box[5] = {64, 2u * tb.bx, 2u * tb.by, 2u * tb.bz, (uint32_t)tb.bn};
          uint32_t es[5] = {1, 2, 2, 2, 1};
          ok = tma_make_map(&p.tmB[pl], p.b_ptr + (size_t)pl * p.b_ps * 2, 5, dB, sB, box, es);
        }
      }
      p.use_tma = ok ? 1 : 0;
    }
  }
  const size_t smem = kWgHeader + (size_t)p.stages * p.stage_bytes + 2048;
  static PerDevice attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(sg_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sg_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    attr.done();
  }
  const int sms = sg_num_sms();
  const char* no_pair = getenv("SG_B200_NO_WGRAD_PAIR");
  // measured (profiles/r02e_sweep_wgrad_pair.txt): +3 % on the SDFNet layers (M = 512, millions of rows), -3 % on Conv3d(128->256) at B = 64
  // (4096 rows per split: the cluster launch and the two extra cluster barriers are not amortised) -> long contractions only
  long long pair_min_rows = 1LL << 18;
  { const char* mr = getenv("SG_B200_WGRAD_PAIR_MIN_ROWS"); if (mr) pair_min_rows = atoll(mr); }      // tests force the pair on small problems
  if (p.use_tma && (p.m_tiles & 1) == 0 && a->rows >= pair_min_rows && !(no_pair && no_pair[0] == '1')) {
    p.pair = 1;
    p.work_total = (long long)(p.m_tiles / 2) * p.n_groups * p.ksplit;            // in pairs
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2u * (unsigned)std::min<long long>(p.work_total, sms / 2));
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, sg_wgrad_kernel<true>, p);
    if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
    sg_count_launch();
    return 0;
  }
  const int grid = (int)std::min<long long>(p.work_total, sms);
  sg_wgrad_kernel<false><<<grid, kWgThreads, smem, (cudaStream_t)stream>>>(p);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

// conv weights (taps = 64, grad index = m*sm + c*64 + tap): sum the splits and transpose (tap <-> c) through shared memory so that
// both the partial reads (c contiguous) and the gradient writes (tap contiguous) are coalesced.  grid = (m_valid, cb/32).
__global__ void sg_wgrad_reduce_t64_kernel(const sg_wgrad_reduce_args a) {
  __shared__ float tile[32][65];
  const int m = blockIdx.x, c0 = blockIdx.y * 32;
  const long long n_total = 64LL * a.cb;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;          // 8 warps
  {
    // a warp owns taps wrp, wrp + 8, ..., wrp + 56: eight independent accumulation streams, every split's eight loads in flight together
    // (one tap at a time left 8 x ceil(ksplit / 8) dependent round trips per warp: 14-20 us per launch for 19 MB of L2-resident partials)
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    const int c = c0 + lane;
    if (c < a.cb) {
      const float* p = a.partials + (size_t)m * n_total + (size_t)wrp * a.cb + c;
      const size_t sstride = (size_t)a.m_pad * n_total, tstride = (size_t)8 * a.cb;
#pragma unroll 2
      for (int s = 0; s < a.ksplit; ++s) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(size_t)s * sstride + (size_t)u * tstride];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += v[u];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) tile[lane][wrp + 8 * u] = acc[u] * a.scale;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
    const int c = i >> 6, tap = i & 63;
    if (c0 + c < a.cb && (a.c_valid <= 0 || c0 + c < a.c_valid)) {
      float* g = a.grad + (long long)m * a.sm + (long long)(c0 + c) * 64 + tap;
      *g = a.accumulate ? (*g + tile[c][tap]) : tile[c][tap];
    }
  }
}

// many splits, few outputs (e.g. the 64x64 filter of Conv3d(1->64) reduced over 148 row splits): one warp per output element
__global__ void sg_wgrad_reduce_warp_kernel(const sg_wgrad_reduce_args a) {
  const long long n_total = (long long)a.taps * a.cb;
  const long long total = (long long)a.m_valid * n_total;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; i < total; i += warps) {
    const int m = (int)(i / n_total);
    const int n = (int)(i - (long long)m * n_total);
    const int tap = n / a.cb, c = n - tap * a.cb;
    if (a.c_valid > 0 && c >= a.c_valid) continue;
    float acc = 0.f;
    for (int s = lane; s < a.ksplit; s += 32) acc += a.partials[((size_t)s * a.m_pad + m) * n_total + n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float* g = a.grad + (long long)m * a.sm + (long long)tap * a.st + (long long)c * a.sc;
      acc *= a.scale;
      *g = a.accumulate ? (*g + acc) : acc;
    }
  }
}

extern "C" int sg_wgrad_reduce(const sg_wgrad_reduce_args* a, void* stream) {
  if (!a || !a->partials || !a->grad) return sg_fail(-1, "sg_wgrad_reduce: null");
  const long long total = (long long)a->m_valid * a->taps * a->cb;
  if (total <= 0) return 0;
  const int block = 256;
  if (a->taps == 64 && a->st == 1 && a->sc == 64 && total > (1 << 16)) {
    dim3 grid(a->m_valid, (a->cb + 31) / 32);
    sg_wgrad_reduce_t64_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a);
    SG_CUDA_CHECK_LAUNCH();
    return 0;
  }
  if (a->ksplit >= 16 && total <= (1 << 16)) {
    const int grid = (int)std::min<long long>((total * 32 + block - 1) / block, 148 * 16);
    sg_wgrad_reduce_warp_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a);
    SG_CUDA_CHECK_LAUNCH();
    return 0;
  }
  const int grid = (int)std::min<long long>((total + block - 1) / block, 148 * 16);
  sg_wgrad_reduce_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
