// GPU marching cubes: replaces skimage.measure.marching_cubes_lewiner at model/sdf_net.py:103 (SDFNet.get_mesh).
// HBM-bound stream compaction: the volume is read a few times, nothing is sorted, output order is deterministic (grid order).
//
//   sg_mc_count   one thread per grid point p: v = number of its three owned edges (p -> p + e_axis) that cross the level,
//                 t = triangles of the cell with origin p (case table, sg_mc_tables.h); per-block sums of (v, t) as one uint64.
//   sg_mc_scan    one block: exclusive scan of the block sums (decides where every block's vertices / faces start); totals.
//   sg_mc_emit_vertices  re-derives v per point, block-local scan, writes vbase[p] and the vertices (linear interpolation of the
//                 crossing, float32 operation order of the numpy oracle) and normals (central-difference gradient, interpolated).
//   sg_mc_emit_faces     re-derives the case per cell, block-local scan, writes faces through vbase of the owning points.
#include <algorithm>

#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_mc_tables.h"

namespace sg {

__constant__ uint8_t c_mc_count[256];
__constant__ int8_t c_mc_edges[256 * 16];

constexpr int kMcBlock = 256;

struct McP {
  const float* vol; int nx, ny, nz; float level; float sx, sy, sz;
  unsigned long long* block_sums;   // [blocks + 1]: v in the low 32 bits, t in the high 32 bits (exclusive-scanned in place)
  int* vbase;                       // [points]
  float* verts; float* normals; int* faces;
};

__device__ __forceinline__ float mc_at(const McP& p, int x, int y, int z) { return __ldg(p.vol + ((long long)x * p.ny + y) * p.nz + z); }

// crossing flags of the three edges owned by point (x, y, z) and the triangle count of the cell it is the origin of
__device__ __forceinline__ void mc_point(const McP& p, int x, int y, int z, int& vmask, int& cs) {
  const float lvl = p.level;
  const bool i0 = mc_at(p, x, y, z) < lvl;
  const bool hx = x + 1 < p.nx, hy = y + 1 < p.ny, hz = z + 1 < p.nz;
  const bool ix = hx ? mc_at(p, x + 1, y, z) < lvl : i0;
  const bool iy = hy ? mc_at(p, x, y + 1, z) < lvl : i0;
  const bool iz = hz ? mc_at(p, x, y, z + 1) < lvl : i0;
  vmask = (int)(ix != i0) | ((int)(iy != i0) << 1) | ((int)(iz != i0) << 2);
  cs = -1;
  if (hx && hy && hz) {
    cs = (int)i0 | ((int)ix << 1) | ((int)iy << 2) | ((int)(mc_at(p, x + 1, y + 1, z) < lvl) << 3) | ((int)iz << 4) |
         ((int)(mc_at(p, x + 1, y, z + 1) < lvl) << 5) | ((int)(mc_at(p, x, y + 1, z + 1) < lvl) << 6) |
         ((int)(mc_at(p, x + 1, y + 1, z + 1) < lvl) << 7);
  }
}

__device__ __forceinline__ void mc_decode(const McP& p, long long i, int& x, int& y, int& z) {
  z = (int)(i % p.nz); const long long t = i / p.nz;
  y = (int)(t % p.ny); x = (int)(t / p.ny);
}

// exclusive block scan of one 64-bit value per thread (256 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ unsigned long long mc_block_scan(unsigned long long v, unsigned long long* smem8, unsigned long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) smem8[warp] = inc;
  __syncthreads();
  unsigned long long wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kMcBlock / 32; ++w) { const unsigned long long s = smem8[w]; if (w < warp) wbase += s; tot += s; }
  __syncthreads();
  if (total) *total = tot;
  return wbase + inc - v;
}

__global__ void __launch_bounds__(kMcBlock) sg_mc_count_kernel(const McP p, long long points) {
  __shared__ unsigned long long sm[8];
  const long long i = blockIdx.x * (long long)kMcBlock + threadIdx.x;
  unsigned long long c = 0;
  if (i < points) {
    int x, y, z, vm, cs;
    mc_decode(p, i, x, y, z);
    mc_point(p, x, y, z, vm, cs);
    c = (unsigned long long)__popc(vm) | ((unsigned long long)(cs >= 0 ? c_mc_count[cs] : 0) << 32);
  }
  unsigned long long tot;
  mc_block_scan(c, sm, &tot);
  if (threadIdx.x == 0) p.block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kMcBlock) sg_mc_scan_kernel(unsigned long long* sums, int blocks) {
  __shared__ unsigned long long sm[8];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < blocks; base += kMcBlock) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < blocks ? sums[i] : 0ull;
    unsigned long long tot;
    const unsigned long long ex = mc_block_scan(v, sm, &tot);
    if (i < blocks) sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[blocks] = carry;     // totals: vertices (low), faces (high)
}

__device__ __forceinline__ float mc_grad(const McP& p, int x, int y, int z, int axis) {
  // np.gradient: central differences inside, one-sided first differences at the borders (unit spacing)
  const int n = axis == 0 ? p.nx : (axis == 1 ? p.ny : p.nz);
  const int c = axis == 0 ? x : (axis == 1 ? y : z);
  if (n < 2) return 0.f;
  int lo = c - 1, hi = c + 1; float div = 2.f;
  if (c == 0) { lo = 0; div = 1.f; }
  if (c == n - 1) { hi = n - 1; div = 1.f; }
  const float a = axis == 0 ? mc_at(p, hi, y, z) : (axis == 1 ? mc_at(p, x, hi, z) : mc_at(p, x, y, hi));
  const float b = axis == 0 ? mc_at(p, lo, y, z) : (axis == 1 ? mc_at(p, x, lo, z) : mc_at(p, x, y, lo));
  return __fdiv_rn(__fsub_rn(a, b), div);
}

__global__ void __launch_bounds__(kMcBlock) sg_mc_emit_vertices_kernel(const McP p, long long points) {
  __shared__ unsigned long long sm[8];
  const long long i = blockIdx.x * (long long)kMcBlock + threadIdx.x;
  int x = 0, y = 0, z = 0, vm = 0, cs = -1;
  if (i < points) { mc_decode(p, i, x, y, z); mc_point(p, x, y, z, vm, cs); }
  const unsigned long long ex = mc_block_scan((unsigned long long)__popc(vm), sm, nullptr);
  if (i >= points) return;
  int vid = (int)((p.block_sums[blockIdx.x] & 0xffffffffull) + ex);
  p.vbase[i] = vid;
  if (!vm) return;
  const float v0 = mc_at(p, x, y, z);
  const float g0[3] = {mc_grad(p, x, y, z, 0), mc_grad(p, x, y, z, 1), mc_grad(p, x, y, z, 2)};
  const float sp[3] = {p.sx, p.sy, p.sz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (!((vm >> a) & 1)) continue;
    const int qx = x + (a == 0), qy = y + (a == 1), qz = z + (a == 2);
    const float v1 = mc_at(p, qx, qy, qz);
    const float t = __fdiv_rn(__fsub_rn(p.level, v0), __fsub_rn(v1, v0));
    float pos[3] = {(float)x, (float)y, (float)z};
    pos[a] = __fadd_rn(pos[a], t);
    float* vo = p.verts + (long long)vid * 3;
    vo[0] = __fmul_rn(pos[0], sp[0]); vo[1] = __fmul_rn(pos[1], sp[1]); vo[2] = __fmul_rn(pos[2], sp[2]);
    float n[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float g1 = mc_grad(p, qx, qy, qz, d);
      n[d] = __fdiv_rn(__fadd_rn(g0[d], __fmul_rn(__fsub_rn(g1, g0[d]), t)), sp[d]);
    }
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])), __fmul_rn(n[2], n[2])));
    float* no = p.normals + (long long)vid * 3;
    no[0] = len > 0.f ? __fdiv_rn(n[0], len) : 0.f; no[1] = len > 0.f ? __fdiv_rn(n[1], len) : 0.f; no[2] = len > 0.f ? __fdiv_rn(n[2], len) : 0.f;
    ++vid;
  }
}

__global__ void __launch_bounds__(kMcBlock) sg_mc_emit_faces_kernel(const McP p, long long points) {
  __shared__ unsigned long long sm[8];
  const long long i = blockIdx.x * (long long)kMcBlock + threadIdx.x;
  int x = 0, y = 0, z = 0, vm = 0, cs = -1;
  if (i < points) { mc_decode(p, i, x, y, z); mc_point(p, x, y, z, vm, cs); }
  const int nt = cs >= 0 ? c_mc_count[cs] : 0;
  const unsigned long long ex = mc_block_scan((unsigned long long)nt, sm, nullptr);
  if (nt == 0) return;
  int* fo = p.faces + ((long long)(p.block_sums[blockIdx.x] >> 32) + (long long)ex) * 3;
  const float lvl = p.level;
  for (int k = 0; k < nt * 3; ++k) {
    const int e = c_mc_edges[cs * 16 + k];
    const int axis = e >> 2, a = e & 1, b = (e >> 1) & 1;
    // owner point of edge e: the cell origin shifted by (a, b) on the two other axes (ascending axis order)
    int ox = x, oy = y, oz = z;
    if (axis == 0) { oy += a; oz += b; } else if (axis == 1) { ox += a; oz += b; } else { ox += a; oy += b; }
    const long long oi = ((long long)ox * p.ny + oy) * p.nz + oz;
    // rank of `axis` among the owner's crossing edges
    const bool o0 = mc_at(p, ox, oy, oz) < lvl;
    int rank = 0;
    if (axis > 0 && ox + 1 < p.nx && ((mc_at(p, ox + 1, oy, oz) < lvl) != o0)) ++rank;
    if (axis > 1 && oy + 1 < p.ny && ((mc_at(p, ox, oy + 1, oz) < lvl) != o0)) ++rank;
    fo[k] = p.vbase[oi] + rank;
  }
}

}  // namespace sg

using namespace sg;

static int mc_upload_tables(cudaStream_t st) {
  static PerDevice once;
  if (!once.first()) return 0;
  cudaError_t e = cudaMemcpyToSymbolAsync(c_mc_count, kMcTriCountHost, sizeof(kMcTriCountHost), 0, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_mc_edges, kMcTriEdgesHost, sizeof(kMcTriEdgesHost), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return sg_fail((int)e, cudaGetErrorString(e));
  once.done();
  return 0;
}

static int mc_fill(const sg_mc_args* a, McP* p, long long* points, int* blocks) {
  if (!a || !a->volume || !a->block_sums) return sg_fail(-1, "sg_mc: null");
  if (a->nx < 2 || a->ny < 2 || a->nz < 2 || (long long)a->nx * a->ny * a->nz >= (1LL << 31)) return sg_fail(-2, "sg_mc: grid must be 2..(2^31 points)");
  p->vol = a->volume; p->nx = a->nx; p->ny = a->ny; p->nz = a->nz; p->level = a->level;
  p->sx = a->spacing[0]; p->sy = a->spacing[1]; p->sz = a->spacing[2];
  p->block_sums = (unsigned long long*)a->block_sums; p->vbase = a->vbase; p->verts = a->vertices; p->normals = a->normals; p->faces = a->faces;
  *points = (long long)a->nx * a->ny * a->nz;
  *blocks = (int)((*points + kMcBlock - 1) / kMcBlock);
  return 0;
}

extern "C" size_t sg_mc_workspace_entries(int nx, int ny, int nz) {
  const long long points = (long long)nx * ny * nz;
  return (size_t)((points + kMcBlock - 1) / kMcBlock + 1);
}

extern "C" int sg_mc_count(const sg_mc_args* a, void* stream) {
  McP p; long long points; int blocks;
  int rc = mc_fill(a, &p, &points, &blocks);
  if (rc) return rc;
  rc = mc_upload_tables((cudaStream_t)stream);
  if (rc) return rc;
  sg_mc_count_kernel<<<blocks, kMcBlock, 0, (cudaStream_t)stream>>>(p, points);
  SG_CUDA_CHECK_LAUNCH();
  sg_mc_scan_kernel<<<1, kMcBlock, 0, (cudaStream_t)stream>>>(p.block_sums, blocks);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int sg_mc_emit(const sg_mc_args* a, void* stream) {
  McP p; long long points; int blocks;
  int rc = mc_fill(a, &p, &points, &blocks);
  if (rc) return rc;
  if (!a->vbase || !a->vertices || !a->normals || !a->faces) return sg_fail(-3, "sg_mc_emit: null outputs");
  sg_mc_emit_vertices_kernel<<<blocks, kMcBlock, 0, (cudaStream_t)stream>>>(p, points);
  SG_CUDA_CHECK_LAUNCH();
  sg_mc_emit_faces_kernel<<<blocks, kMcBlock, 0, (cudaStream_t)stream>>>(p, points);
  SG_CUDA_CHECK_LAUNCH();
  return 0;
}
