// Experiment (round-2 groundwork): can a tcgen05 A operand be a SHIFTED, NON-1024-ALIGNED VIEW into a larger 128B-swizzled
// shared-memory block?  If yes, one TMA halo box per stride-2 parity class can serve several filter taps of the implicit GEMM
// (tap kw+2 = the kw box shifted by one 128-byte row, 8-row groups 9 rows apart), cutting the TMA row count -- the measured
// per-SM limit of sg_igemm -- by up to 4x.
//
// Block layout: rows j = 0..143 of 128 B (64 bf16), written like TMA writes a swizzled box: 16-byte chunk c of row j lands at
// j*128 + ((c ^ (j & 7)) << 4) from a 1024-aligned base.  Logical A row r = 8g + i  <->  block row 9g + i + shift.
// Descriptor: start = base + shift*128, SBO = 9*128, SWIZZLE_128B; variants of the base_offset field.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/exp_umma_view tools/exp_umma_view.cu && /tmp/exp_umma_view
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../shapegan_b200/csrc/sg_common.cuh"

using namespace sg;

constexpr int kRowsBlock = 144, kN = 64;

__global__ void __launch_bounds__(128, 1) exp_kernel(const bf16* a_block, const bf16* b_tile, float* out, int shift, int sbo_rows,
                                                     int base_off_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* a_s = smem;                 // 144 rows x 128 B = 18 KB (1024-aligned)
  uint8_t* b_s = smem + 20480;         // 64 rows x 128 B, canonical tile (1024-aligned)
  for (int i = tid; i < kRowsBlock * 8; i += 128) {
    const int j = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + j * 128 + ((c ^ (j & 7)) << 4)) = *reinterpret_cast<const uint4*>(a_block + j * 64 + c * 8);
  }
  for (int i = tid; i < kN * 8; i += 128) {
    const int j = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + j * 128 + ((c ^ (j & 7)) << 4)) = *reinterpret_cast<const uint4*>(b_tile + j * 64 + c * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    const uint32_t a0 = smem_u32(a_s) + (uint32_t)shift * 128u, b0 = smem_u32(b_s);
    const uint32_t idesc = umma_idesc(128, kN, false, false);
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t da = umma_desc(a0 + kk * 32, 16, (uint32_t)sbo_rows * 128u);
      if (base_off_mode == 1) da |= (uint64_t)((a0 >> 7) & 7u) << 49;
      const uint64_t db = umma_desc(b0 + kk * 32, 16, 1024);
      umma_bf16(tmem, da, db, idesc, kk > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0, nullptr);
  tc_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < kN; c0 += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)tid * kN + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  std::vector<float> ha(kRowsBlock * 64), hb(kN * 64);
  for (int j = 0; j < kRowsBlock; ++j) for (int k = 0; k < 64; ++k) ha[j * 64 + k] = (float)((j * 7 + k * 3) % 13 - 6);
  for (int n = 0; n < kN; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = (float)((n * 5 + k) % 7 - 3);
  std::vector<bf16> ba(ha.size()), bb(hb.size());
  for (size_t i = 0; i < ha.size(); ++i) ba[i] = __float2bfloat16(ha[i]);
  for (size_t i = 0; i < hb.size(); ++i) bb[i] = __float2bfloat16(hb[i]);
  bf16 *da, *db; float* dout;
  cudaMalloc(&da, ba.size() * 2); cudaMalloc(&db, bb.size() * 2); cudaMalloc(&dout, 128 * kN * 4);
  cudaMemcpy(da, ba.data(), ba.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, bb.data(), bb.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> hout(128 * kN);
  struct Cfg { int shift, sbo_rows, mode; const char* what; };
  const Cfg cfgs[] = {{0, 8, 0, "control: dense tile, SBO 1024"},
                      {0, 9, 0, "groups 9 rows apart, shift 0, base_offset 0"},
                      {0, 9, 1, "groups 9 rows apart, shift 0, base_offset (start>>7)&7"},
                      {1, 9, 0, "groups 9 rows apart, shift 1, base_offset 0"},
                      {1, 9, 1, "groups 9 rows apart, shift 1, base_offset (start>>7)&7"},
                      {1, 8, 0, "dense groups, shift 1 (start +128 B), base_offset 0"},
                      {1, 8, 1, "dense groups, shift 1 (start +128 B), base_offset 1"}};
  for (const Cfg& c : cfgs) {
    cudaMemset(dout, 0, 128 * kN * 4);
    exp_kernel<<<1, 128, 32768>>>(da, db, dout, c.shift, c.sbo_rows, c.mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-62s CUDA error %s\n", c.what, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int r = 0; r < 128; ++r) {
      const int j = (r >> 3) * c.sbo_rows + (r & 7) + c.shift;
      for (int n = 0; n < kN; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)ha[j * 64 + k] * hb[n * 64 + k];
        const double d = fabs(ref - hout[r * kN + n]);
        if (d > 1e-3) ++bad;
        if (d > worst) worst = d;
      }
    }
    printf("%-62s max|err| %.3g  mismatches %d / %d  -> %s\n", c.what, worst, bad, 128 * kN, bad ? "WRONG" : "OK");
  }
  return 0;
}
