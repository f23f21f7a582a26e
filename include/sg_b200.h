/* libsg_b200 — C ABI of the B200-native shapegan hot path.
 *
 * The reference (marian42/shapegan) has no FFI/plugin layer: every FLOP on its hot path is issued through
 * torch.nn modules (model/gan.py:8-23,48-57; model/progressive_gan.py:26-42; model/autoencoder.py:15-64;
 * model/sdf_net.py:26-52).  This header is the boundary we define UNDER those classes: each entry point names the
 * torch call it replaces.  Host code (shapegan_b200/*.py) binds it with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, <0 for an argument/shape error (nothing launched), >0 = cudaError_t.
 *     sg_last_error() returns a thread-local message.  Nothing here allocates, frees, synchronises or retains
 *     pointers: all buffers (incl. workspaces) belong to the caller; work is ordered on `stream`
 *     (a cudaStream_t passed as void*) and is CUDA-graph capturable.
 *   - activations are NDHWC bf16 "plane" tensors: planes=1 -> plain bf16 (throughput mode); planes=2 -> hi/lo bf16
 *     split of an fp32 value (value = hi + lo), consumed by the tensor cores as three bf16 products (fp32x mode,
 *     the <=1e-3 parity mode).  Single-channel volumes (voxel grids) and vectors at module boundaries are fp32.
 *   - weights live in fp32 torch layout (nn.Parameter) and are re-packed into tensor-core operand images
 *     (K-major, 128B-swizzled, zero padded) by sg_pack_b whenever they change.
 */
#ifndef SG_B200_H
#define SG_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_ABI_VERSION 1

/* gather modes of the implicit GEMM (A operand) */
#define SG_MODE_DENSE 0 /* rows x K contiguous (nn.Linear; k4/s1 convs on 4^3 or 1^3 grids)                    */
#define SG_MODE_CONV 1  /* Conv3d k4 s2 p1 forward gather == ConvTranspose3d k4 s2 p1 input-gradient gather     */
#define SG_MODE_CONVT 2 /* ConvTranspose3d k4 s2 p1 forward (8 output-parity classes) == Conv3d input gradient  */
#define SG_MODE_PATCH 3 /* Conv3d k4 s2 p1 over a single-channel fp32 volume (first discriminator/encoder layer) */

#define SG_ACT_NONE 0
#define SG_ACT_LRELU 1 /* slope 0.2, model/gan.py:11 */
#define SG_ACT_RELU 2
#define SG_ACT_TANH 3
#define SG_ACT_SIGMOID 4

#define SG_OUT_BF16 0       /* bf16 planes, row stride out_ld */
#define SG_OUT_F32 1        /* fp32 [rows][out_ld] */
#define SG_OUT_F32_ATOMIC 2 /* fp32 atomicAdd (split-K) */

typedef struct {
  const void* ptr;      /* bf16 planes (NDHWC) or fp32 (SG_MODE_PATCH source) */
  int64_t plane_stride; /* elements between the hi and lo plane */
  int32_t n, d, h, w, c;
} sg_tensor;

/* out[rows, n] = act( gather(A)[rows, K] . B[n, K]^T + bias ) (* act'(mask))
 * replaces: F.conv3d / F.conv_transpose3d / F.linear forward and input-gradient calls issued by
 * model/gan.py:9-21,49-55, model/progressive_gan.py:28-38, model/autoencoder.py:16-63, model/sdf_net.py:27-50 */
typedef struct {
  int32_t mode, planes;
  sg_tensor a;  /* source tensor */
  sg_tensor a2; /* optional second DENSE source appended along K (ptr NULL = unused); sdf_net.py:59 skip-concat */
  int64_t rows; /* GEMM rows (per parity class in SG_MODE_CONVT) */
  int32_t k;    /* padded K, multiple of 64 */
  int32_t n_pad, n_valid;
  int32_t bn, mt, ksplit; /* tile config: N tile (mult of 16, <=256), M sub-tiles per CTA (1|2), K splits; 0 = auto */
  const void* b_packed;   /* sg_pack_b image */
  const float* bias;      /* [n_valid] (or [bias_mod]) or NULL */
  int32_t act;
  int32_t bias_mod;       /* >0: bias index = n % bias_mod (ConvTranspose3d on a 1^3 grid: n = pos*Cout + co) */
  const void* mask; /* optional bf16 planes, same layout as out: out *= act'(mask) with mask_act */
  int64_t mask_plane_stride;
  int32_t mask_act;
  void* out;
  int64_t out_plane_stride;
  int32_t out_kind, out_ld;
  int32_t out_d, out_h, out_w; /* SG_MODE_CONVT: output grid (2d,2h,2w) */
  void* splitk_ws;             /* optional fp32 workspace for split-K with a non-atomic output: [ksplit][out rows][n_pad]; the K splits
                                  store partial slabs, a finish kernel sums them and applies bias / act / mask.  NULL = never split. */
  int64_t splitk_ws_bytes;
} sg_igemm_args;
int sg_igemm(const sg_igemm_args* a, void* stream);
/* bytes of splitk_ws sg_igemm would use for `a` (0 = it will not split K): few output tiles x long K, e.g. Conv3d(128->256) 8^3->4^3 */
int sg_igemm_plan(const sg_igemm_args* a, size_t* ws_bytes);

/* dW partials: P[split][m, n] = sum_rows A[row, m] * gather(B)[row (+tap), n]
 * replaces: the weight-gradient (bwd-filter) calls autograd issues for the layers above.
 * A = rows x ca dense (e.g. dY), B = tensor gathered like SG_MODE_CONV (taps=64), DENSE (taps=1) or PATCH. */
typedef struct {
  int32_t b_mode, planes; /* SG_MODE_DENSE | SG_MODE_CONV | SG_MODE_PATCH */
  sg_tensor a;            /* dense rows: n*d*h*w rows of c channels */
  sg_tensor b;            /* gathered tensor */
  int64_t rows;
  int32_t ksplit;     /* number of row splits (0 = auto) */
  int32_t merge_n;    /* 1: one MMA spans all N atoms (LBO stride); 0: one MMA per 64-column atom */
  float* partials;    /* [ksplit][m_pad][n_total] fp32 workspace, m_pad = roundup(a.c,128), n_total = taps*b.c */
  int32_t ksplit_out; /* filled by sg_wgrad_plan */
  float* bias_partials;   /* optional [ksplit][m_pad] fp32: column sums of A per split (A = dY: the bias gradient), produced by one extra
                             N = 16 MMA per K step against a tile of ones; fold with sg_wgrad_reduce (taps = 1, cb = 1, sm = 1) */
  int32_t bias_ws_floats; /* filled by sg_wgrad_plan: floats bias_partials must hold */
} sg_wgrad_args;
int sg_wgrad_plan(sg_wgrad_args* a, size_t* workspace_bytes); /* fills ksplit, returns bytes for partials */
int sg_wgrad(const sg_wgrad_args* a, void* stream);

/* Reduce wgrad partials over splits and scatter into a torch-layout fp32 gradient:
 * grad[(m*sm + tap*st + c*sc)] (+)= sum_s P[s][m][tap*cb + c]        (m < m_valid, c < cb) */
typedef struct {
  const float* partials;
  int32_t ksplit, m_pad, m_valid, taps, cb;
  int64_t sm, st, sc;
  float* grad;
  int32_t accumulate; /* 1: += (torch .grad accumulation semantics) */
  float scale;
  int32_t c_valid;    /* >0: only columns c < c_valid of each tap are emitted (zero-padded operands) */
} sg_wgrad_reduce_args;
int sg_wgrad_reduce(const sg_wgrad_reduce_args* a, void* stream);

/* Pack fp32 torch-layout weights into the K-major swizzled operand image consumed by sg_igemm.
 * image[class][kc][plane][n_pad][64]; logical B[n, k] with n = n1*n0_count + n0 and k = tap*kc_count + c:
 *   src index = n1*s_n1 + n0*s_n0 + tapmap(class, tap)*s_tap + c*s_c ; zero outside (n < n_valid, c < c_valid). */
typedef struct {
  const float* w;
  void* image;
  int32_t planes, classes; /* classes = 8 selects the k4/s2/p1 transposed tap map, else 1 */
  int32_t n_pad, n_valid, n0_count;
  int64_t s_n1, s_n0;
  int32_t k_pad, taps, c_count, c_valid; /* k = tap*c_count + c */
  int64_t s_tap, s_c;
} sg_pack_b_args;
size_t sg_pack_b_bytes(const sg_pack_b_args* a);
int sg_pack_b(const sg_pack_b_args* a, void* stream);

/* ---- HBM-bound companions (all tensors are plane tensors viewed as [rows, C], C % 8 == 0) ----
 * `*_ps` = plane stride in elements; `sums` = zero-initialised double workspace.                                  */

/* g = ga * act'(y) (y = stored activation output); sums[0..c) += column sums of g (= bias gradient).
 * replaces: LeakyReLU/ReLU backward + bias gradient reductions (model/gan.py:11..52, model/sdf_net.py:28..47)      */
int sg_act_bwd(const void* ga, int64_t ga_ps, const void* y, int64_t y_ps, void* g, int64_t g_ps, int planes, int64_t rows,
               int c, int act, double* sums, void* stream);
/* train-mode BatchNorm3d/1d (model/gan.py:10,14,18; model/autoencoder.py:17..60): sums[0..c)=sum x, [c..2c)=sum x^2 */
int sg_bn_stats(const void* x, int64_t x_ps, int planes, int64_t rows, int c, double* sums, void* stream);
int sg_bn_finalize(const double* sums, int64_t rows, int c, float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, void* stream);
int sg_bn_apply(const void* x, int64_t x_ps, void* y, int64_t y_ps, int planes, int64_t rows, int c, const float* mean,
                const float* invstd, const float* gamma, const float* beta, int act, void* stream);
int sg_bn_bwd_reduce(const void* ga, int64_t ga_ps, const void* y, int64_t y_ps, const void* x, int64_t x_ps, int planes,
                     int64_t rows, int c, int act, const float* mean, const float* invstd, double* sums, void* stream);
int sg_bn_bwd_apply(const void* ga, int64_t ga_ps, const void* y, int64_t y_ps, const void* x, int64_t x_ps, void* gx,
                    int64_t gx_ps, int planes, int64_t rows, int c, int act, const float* mean, const float* invstd,
                    const float* gamma, const double* sums, void* stream);
/* dst[(i / wc)*s_t + (i % wc)*s_c] (+)= scale*src[i] */
int sg_emit_sums(const double* src, float* dst, int n, int accumulate, float scale, int wc, int64_t s_t, int64_t s_c, void* stream);
/* second stage of ConvTranspose3d(C->1,k4,s2,p1) (model/gan.py:21, model/autoencoder.py:63) and of the input
 * gradient of Conv3d(1->C): out[n,2d,2h,2w] = act(bias + sum of the 8 taps of P[n*d*h*w, 64])                       */
int sg_col2im_c1(const void* P, int64_t p_ps, int planes, int n, int d, int h, int w, const float* bias, int act, float* out,
                 void* stream);
int sg_unary_f32(const float* x, float* y, int64_t n, int act, void* stream);
int sg_unary_bwd_f32(const float* gy, const float* y, float* gx, int64_t n, int act, void* stream);
/* y[r] = act(x[r,:].w + b)  (Linear(256->1)+Tanh model/sdf_net.py:50-51; Conv3d(256->1,k4,s1) model/gan.py:55;
 * Linear(128->1) model/progressive_gan.py:30) and its backward: gx = g*w, sums[0..c) += g*x, sums[c] += g            */
/* weight element of column k lives at w[(k / wc)*s_t + (k % wc)*s_c] (wc % 8 == 0) */
int sg_rowdot_fwd(const void* x, int64_t x_ps, int planes, int64_t rows, int c, const float* w, int wc, int64_t s_t, int64_t s_c,
                  const float* bias, int act, float* y, void* stream);
/* x_mask_act != SG_ACT_NONE additionally multiplies gx by act'(x) (x = stored activation feeding this layer) */
int sg_rowdot_bwd(const float* gy, const float* y, int act, const void* x, int64_t x_ps, int planes, int64_t rows, int c,
                  const float* w, int wc, int64_t s_t, int64_t s_c, void* gx, int64_t gx_ps, double* sums, int x_mask_act, void* stream);
int sg_to_planes(const float* src, int64_t src_ld, int64_t rows, int c_src, void* dst, int64_t dst_ps, int planes, int c_dst,
                 void* stream);
int sg_from_planes(const void* src, int64_t src_ps, int planes, int64_t rows, int c_src, int c_take, float* dst, int64_t dst_ld,
                   int accumulate, float scale, void* stream);
/* rows [x,y,z,latent,0..] of cat(points, latent_codes) (model/sdf_net.py:57); latent row i = index ? table[index[i]] : latent[i] */
int sg_sdf_pack_input(const float* points, const float* latent, const int32_t* index, int L, int64_t n, void* dst, int64_t dst_ps,
                      int planes, int c_dst, void* stream);
int sg_sdf_unpack_grad(const void* ga, int64_t ga_ps, const void* gb, int64_t gb_ps, int planes, int64_t n, int c_src, int L,
                       const int32_t* index, float* gpoints, float* glatent, void* stream);
/* fade-in blend model/progressive_gan.py:48-50 */
int sg_fade_fwd(const void* x, int64_t x_ps, void* y, int64_t y_ps, int planes, int b, int r, int c, const float* vol, float f,
                void* stream);
int sg_fade_bwd_vol(const void* g, int64_t g_ps, int planes, int b, int r, int c, float f, float* gvol, void* stream);
int sg_axpby_planes(const void* a, int64_t a_ps, const void* b, int64_t b_ps, void* y, int64_t y_ps, int planes, int64_t elems,
                    float alpha, float beta, void* stream);
/* fused optimizer updates over flat fp32 arenas (torch.optim.RMSprop/Adam defaults: train_wgan.py:45-46,
 * train_gan.py:28-31, train_sdf_autodecoder.py:44-45); clip>0 folds Discriminator.clip_weights (model/gan.py:67-69) */
int sg_rmsprop(float* p, const float* g, float* sq, int64_t n, float lr, float alpha, float eps, float grad_scale, float clip,
               void* stream);
int sg_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, int step,
            float grad_scale, void* stream);
int sg_clamp(float* p, int64_t n, float lo, float hi, void* stream);
/* loss_sum += mean|out-target| ; gout = sign(out-target)/n   (train_sdf_autodecoder.py:88 data term) */
/* out[0] += sum x (double) */
int sg_sum_f32(const float* x, int64_t n, double* out, void* stream);
int sg_l1_loss_grad(const float* out, const float* target, float* gout, int64_t n, double* loss_sum, void* stream);

/* Fused persistent DeepSDF MLP forward (model/sdf_net.py:56-61, latent_code_size 128, bf16 operands): all 8 layers of a
 * tile pair in one CTA.  w_img = 28 weight chunks of 32 KB in stream order (sg_pack_b images of
 * L1[:,3:131] | L2 | L3 | L4 | L5[:,0:256] | L5[:,259:387] | L6 | L7); aux = fp32 {float4[256] (wx,wy,wz,b) of L1,
 * same of L5, bias[5][256] of L2,L3,L4,L6,L7, w8[256], b8}; stash (optional) receives the 7 hidden activations as
 * bf16 [7][n][256] for the backward pass.  replaces: the 8 F.linear + 2 torch.cat + ReLU/Tanh launches of SDFNet.forward */
typedef struct {
  const float* points;
  const float* latent;
  const int32_t* index; /* NULL: latent row i belongs to point i */
  int64_t n;
  const void* w_img;
  const float* aux;
  float* out;
  void* stash;
  void* mask_stash; /* optional uint32 [7][n][8]: ReLU mask of the 7 hidden activations, 1 bit per element (bit k of word j =
                       column 32j+2k, bit 16+k = column 32j+2k+1); what sg_sdfnet_bwd reads instead of the activations */
} sg_sdfnet_fwd_args;
int sg_sdfnet_fwd(const sg_sdfnet_fwd_args* a, void* stream);
int sg_sdfnet_fwd_layout(int32_t* chunks, int32_t* chunk_bytes, int32_t* aux_floats);

/* Single-latent inference with the same fused kernel (SDFNet.evaluate_in_batches / get_voxels / get_normals, model/sdf_net.py:63-128, and
 * the sphere tracer's step, rendering/raymarching.py:48-61,106-120): one latent code z for every point, so the latent part of layers1.0 and
 * layers2.0 is the constant vector W[:, latent] z -- the host folds it into those layers' bias slots of `aux`, and the kernel streams no
 * latent chunks at all (layers1.0 needs no MMA: it is entirely the fp32 accumulator pre-load).  0.79 of the forward FLOPs per point.
 * Point source, one of: points[slot][3]; grid_r > 0: slot s is cell (s / r^2, s / r % r, s % r) of the util.get_voxel_coordinates grid and
 * its coordinates come from grid_axis[3][r] (the float32(float64 linspace) values, bit-exact); trace_points (a tracing step).
 * ray_index (optional) maps compact row i to its slot: rows outside the 1.1 sphere are simply not listed and the result lands in
 * out[slot] of a pre-filled grid (model/sdf_net.py:7-19,77-95).  n_ptr (optional) is a device-resident row count <= n.
 * Tracing step: p = trace_points[slot]; d = clamp(sdf(p) + sdf_offset, +-trace_clamp); p += trace_dirs[slot] * d (in place);
 * hit = 0 < d < trace_threshold -> trace_hit[slot] = 1; miss = |p| > trace_radius (or p.y > trace_radius when trace_miss_y);
 * every other ray is appended to next_index[atomicAdd(next_count, 1)].  out (optional) receives d. */
typedef struct {
  const float* points;
  int64_t n;
  const int32_t* n_ptr;
  const int32_t* ray_index;
  int32_t grid_r;
  const float* grid_axis;
  const void* w_img; /* the forward image of sg_sdfnet_fwd (its latent chunks are skipped) */
  const float* aux;  /* sg_sdfnet_fwd's aux block with b + W[:, latent] z in the bias slots of layers1.0 / layers2.0 */
  float* out;
  void* mask_stash;  /* optional uint32 [7][n][8] ReLU masks (compact row order) for sg_sdfnet_bwd's xyz gradient */
  float* trace_points;
  const float* trace_dirs;
  uint8_t* trace_hit;
  int32_t* next_index;
  int32_t* next_count;
  float sdf_offset, trace_clamp, trace_threshold, trace_radius;
  int32_t trace_miss_y;
} sg_sdfnet_infer_args;
int sg_sdfnet_infer(const sg_sdfnet_infer_args* a, void* stream);

/* Fused input-gradient chain of the same MLP (what autograd runs for SDFNet.forward, model/sdf_net.py:56-61, between the
 * tanh head and layers1.0): g7 = gout*(1-out^2)*w8*[h7>0], then g_{l-1} = (g_l . W_l[:, :256]) * [h_{l-1}>0] for l = 7..2,
 * one persistent CTA per tile pair.  mask_stash = the forward's 1-bit ReLU masks; wt_img = 24 chunks of 32 KB = sg_pack_b images of the
 * TRANSPOSED weights B[n = in-feature][k = out-feature] of layers2.4, layers2.2, layers2.0[:, :256], layers1.6, layers1.4,
 * layers1.2 (in that order); w8 = layers2.6.weight [256].  gstash receives g_1..g_7 as bf16 [7][n][256] (index l-1): the
 * operands of the weight-gradient GEMMs, bias sums and the two input-gradient GEMMs (sg_wgrad / sg_act_bwd / sg_igemm).
 * Both fused kernels keep their small fp32 tables in __constant__ memory, uploaded on `stream` by the call itself: do not
 * run two DIFFERENT SDFNets concurrently on two streams of one process. */
typedef struct {
  const float* gout; /* [n] gradient w.r.t. the tanh output */
  const float* out;  /* [n] the forward's output */
  const void* mask_stash;
  const void* wt_img;
  const float* w8;
  int64_t n;
  void* gstash;        /* may be NULL when only gpoints is wanted (SDFNet.get_normals: nothing but [n][3] floats leaves the SM) */
  float* gpoints;      /* optional [n][3]: gradient w.r.t. the point coordinates, accumulated inside the drains of g_5 and g_1 */
  const float* xyz_w;  /* with gpoints: fp32 [2][3][256] = layers1.0.weight[:, 0:3]^T, layers2.0.weight[:, 256:259]^T */
} sg_sdfnet_bwd_args;
int sg_sdfnet_bwd(const sg_sdfnet_bwd_args* a, void* stream);

/* Data-parallel optimizer step as ONE kernel over NVLink peer memory (replaces nn.DataParallel's gradient reduction,
 * train_hybrid_progressive_gan.py:62-71, + torch.optim.RMSprop/Adam .step(), train_wgan.py:45-46,70): every rank owns a shard of the flat
 * arena and its optimizer state; in-barrier -> reduce its shard straight out of all peers' gradient arenas (float4 loads over NVLink) ->
 * update -> store the new parameters into every peer's parameter arena -> out-barrier.  peer_grad / peer_param / peer_pad are HOST
 * arrays of `world` device pointers into symmetric (peer-mapped) memory: gradient arenas, parameter arenas, and uint32 [2][world] signal
 * pads (zero-initialised once).  sync = local uint32 [3] (zero-initialised once).  chunk = shard length (multiple of 4, chunk*world >= n);
 * s1 / s2 hold the state of the LOCAL shard only.  kind 0 = RMSprop (beta1 = alpha), 1 = Adam (`step` = 1-based step count).
 * Capturable; every rank must launch it in the same order. */
typedef struct {
  const void* const* peer_grad;
  const void* const* peer_param;
  const void* const* peer_pad;
  int32_t rank, world;
  int64_t n, chunk;
  float* s1;
  float* s2;
  int32_t kind;
  float lr, beta1, beta2, eps, clip, grad_scale;
  int32_t step;
  void* sync;
} sg_dp_step_args;
int sg_dp_step(const sg_dp_step_args* a, void* stream);

/* ---- non-GEMM pieces of the hand-scheduled critic update (shapegan_b200/critic.py) ----
 * sg_gp_interp : out[b] = alpha[b] * real[b] + (1 - alpha[b]) * fake[b]          (train_hybrid_progressive_gan.py:103-105), m floats per sample
 * sg_gp_seed   : n_b = |g_b|_2 ; *gp_sum += weight (n_b - 1)^2 / B ; v_b = 2 weight (n_b - 1) / (B n_b) g_b   (:110-111 and its derivative w.r.t. g)
 * sg_critic_loss : out4 = {mean s[0:B] - mean s[B:2B] + gp, gp, mean s[0:B], mean s[B:2B]}   (train_wgan.py:66-68; :163) */
int sg_gp_interp(const float* real, const float* fake, const float* alpha, float* out, int b, int64_t m, void* stream);
int sg_gp_seed(const float* g, float* v, int b, int64_t m, float weight, double* gp_sum, void* stream);
int sg_critic_loss(const float* scores, int b, const double* gp_sum, float* out4, void* stream);

/* ---- point-set GAN path (model/point_sdf_net.py): everything that is not a Linear layer; plane tensors [P][rows][C], C % 8 == 0 ----
 * sg_ln_act_fwd/bwd : y = act(LayerNorm_C(x) gamma + beta) (eps inside the root, biased variance), stats = fp32 [rows][2] (mean, rstd);
 *                     bwd also returns dbeta / dgamma (workspace: >= 2*C floats per block, up to 296 blocks; deterministic)
 * sg_rows_add_vec   : y[row] = x[row] + v[row / seg_len]   (z_lin(z).unsqueeze(1) + x, :105-109); sg_segment_colsum = its backward w.r.t. v
 * sg_segmax_fwd     : out[s] = max over the seg_len rows of segment s, arg[s][c] = first arg-max row (x.max(dim=-2)[0], :40-41)
 * sg_segmax_move    : gather = 0: big[s*seg_len + arg[s][c]][c] = small[s][c] (big pre-zeroed) ; gather = 1: the reverse */
int sg_ln_act_fwd(const void* x, int64_t x_ps, void* y, int64_t y_ps, int planes, int64_t rows, int c, const float* gamma, const float* beta,
                  float eps, int act, float* stats, void* stream);
int sg_ln_act_bwd(const void* gy, int64_t gy_ps, const void* y, int64_t y_ps, const void* x, int64_t x_ps, void* gx, int64_t gx_ps, int planes,
                  int64_t rows, int c, const float* gamma, int act, const float* stats, float* gbeta, float* ggamma, float* workspace,
                  int64_t workspace_floats, void* stream);
int sg_rows_add_vec(const void* x, int64_t x_ps, const float* v, void* y, int64_t y_ps, int planes, int64_t rows, int c, int64_t seg_len, void* stream);
size_t sg_segment_colsum_workspace(int segs, int c, int64_t seg_len); /* bytes; 0 = none (few long segments are split over rows) */
int sg_segment_colsum(const void* x, int64_t x_ps, int planes, int segs, int c, int64_t seg_len, float* out, float* workspace, void* stream);
int sg_segmax_fwd(const void* x, int64_t x_ps, int planes, int segs, int c, int64_t seg_len, void* out, int64_t out_ps, int32_t* arg, void* stream);
int sg_segmax_move(void* big, int64_t big_ps, void* small, int64_t small_ps, int planes, int segs, int c, int64_t seg_len, const int32_t* arg,
                   int gather, void* stream);

/* ---- SDFNet inference consumers / voxel data path ---- */
/* Cells s = (ix*r + iy)*r + iz of the util.get_voxel_coordinates(r) grid (util.py:60-74) with |p| < radius, the sphere mask of
 * SDFVoxelizationHelperData (model/sdf_net.py:12-14), evaluated in numpy's float32 arithmetic operation by operation (bit-exact set).
 * axis = [3][r] coordinate tables; index_out has room for r^3 entries (unordered); *count must be zero on entry. */
int sg_grid_sphere_index(int r, const float* axis, float radius, int32_t* index_out, int32_t* count, void* stream);
/* VoxelDataset.__getitem__ (datasets.py:16-23) on a whole raw batch: dst = clamp(src, -clamp, clamp) [/ clamp], float32, bit-exact */
int sg_voxel_ingest(const float* src, float* dst, int64_t n, float clamp, int rescale, void* stream);

/* GPU marching cubes (replaces skimage.measure.marching_cubes_lewiner at model/sdf_net.py:103): volume fp32 [nx][ny][nz] (z fastest),
 * inside = value < level, vertices on the crossing cell edges (linear interpolation, index space x spacing), faces oriented towards
 * increasing value, per-vertex normals from the interpolated central-difference gradient.  Two calls around one host read:
 *   sg_mc_count : fills block_sums (uint64 [sg_mc_workspace_entries]); its LAST entry = #vertices (low 32 bits) | #faces << 32
 *   sg_mc_emit  : with vertices [V][3], normals [V][3], faces [F][3] and vbase int32 [nx*ny*nz] allocated by the caller
 * Output order is deterministic (grid order); the case table is generated (oracle/mc_tables.py -> csrc/sg_mc_tables.h). */
typedef struct {
  const float* volume;
  int32_t nx, ny, nz;
  float level;
  float spacing[3];
  void* block_sums;
  int32_t* vbase;
  float* vertices;
  float* normals;
  int32_t* faces;
} sg_mc_args;
size_t sg_mc_workspace_entries(int nx, int ny, int nz);
int sg_mc_count(const sg_mc_args* a, void* stream);
int sg_mc_emit(const sg_mc_args* a, void* stream);

/* ---- misc ---- */
int sg_abi_version(void);
const char* sg_last_error(void);
int sg_device_error_word(int32_t** dev_ptr); /* device word set by a kernel watchdog (mbarrier timeout) */
int sg_check_device_error(void);               /* diagnostic, synchronises: returns and clears that word */
int sg_num_sms(void);
/* 0 when `stream` is not being captured into a CUDA graph, else the unique id of that capture (host-side weight-image
 * caches must not reuse an image packed OUTSIDE the capture: the graph would replay with frozen weights) */
int sg_stream_capture_id(void* stream, unsigned long long* id_out);
long long sg_launch_count(void);               /* kernels launched by this library since load (bench gpu_launches) */
/* diagnostics (tools/diag_conv.py): clock64 trace [3][1024] {producer got a stage, MMA warp got its data, MMAs issued} of CTA 0
 * of the last sg_igemm launched with env SG_B200_IGEMM_DIAG & 128; synchronises */
int sg_debug_igemm_trace(long long* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* SG_B200_H */
