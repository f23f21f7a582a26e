"""Thin torch-tensor front end over the C ABI (shapegan_b200/_lib.py).  torch is used here only for device
memory (caching allocator) and the current stream; every FLOP happens inside libsg_b200."""
import ctypes

import torch

from . import _lib as L


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def capture_id():
    """0 outside CUDA-graph capture, else the unique id of the capture running on the current stream"""
    if not torch.cuda.is_current_stream_capturing():
        return 0
    cid = ctypes.c_ulonglong(0)
    L.check(L.lib().sg_stream_capture_id(stream(), ctypes.byref(cid)), 'sg_stream_capture_id')
    return int(cid.value)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('shapegan_b200: tensors must live on a CUDA device (no CPU fallback on the hot path)')


def round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------- plane tensors
def to_planes(x, planes):
    """fp32 -> bf16 plane tensor [P, ...] (value = sum of planes)."""
    hi = x.to(torch.bfloat16)
    if planes == 1:
        return hi.unsqueeze(0).contiguous()
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack((hi, lo), 0).contiguous()


def from_planes(t):
    return t.float().sum(0) if t.shape[0] > 1 else t[0].float()


def tdesc(t, n=1, d=1, h=1, w=1, c=1):
    """Descriptor of a plane tensor [P, ...] (bf16) or an fp32 volume (P-less)."""
    ps = t.stride(0) if (t.dtype == torch.bfloat16 and t.shape[0] == 2) else 0
    return L.SgTensor(ctypes.c_void_p(t.data_ptr()), ps, n, d, h, w, c)


def null_tensor():
    return L.SgTensor(None, 0, 0, 0, 0, 0, 0)


# ------------------------------------------------------------------------------------------------- weight packing
def pack_b(w, planes, n_valid, k_pad, taps, c_count, c_valid, s_n0, s_tap, s_c, classes=1, n0_count=None, s_n1=0,
           n_pad=None, out=None):
    _require_cuda(w)
    assert w.dtype == torch.float32      # may be a strided view: all addressing goes through the explicit strides
    n_pad = round_up(n_valid, 16) if n_pad is None else n_pad
    a = L.SgPackBArgs(ctypes.c_void_p(w.data_ptr()), None, planes, classes, n_pad, n_valid,
                      n0_count if n0_count is not None else n_pad, s_n1, s_n0, k_pad, taps, c_count, c_valid, s_tap, s_c)
    nbytes = L.lib().sg_pack_b_bytes(ctypes.byref(a))
    if out is not None:
        assert out.numel() == nbytes and out.dtype == torch.uint8
        img = out
    else:
        img = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    a.image = ctypes.c_void_p(img.data_ptr())
    L.check(L.lib().sg_pack_b(ctypes.byref(a), stream()), 'sg_pack_b')
    return img


def pack_conv_fwd(w, planes):
    """nn.Conv3d weight [Cout,Cin,4,4,4] -> B[n=cout, k=(tap,cin)] for SG_MODE_CONV (or PATCH when Cin==1)."""
    cout, cin = w.shape[0], w.shape[1]
    return pack_b(w, planes, cout, 64 * cin, 64, cin, cin, s_n0=cin * 64, s_tap=1, s_c=64)


def pack_conv_dgrad(w, planes):
    """nn.Conv3d weight -> B[class][n=cin, k=(t,cout)] for SG_MODE_CONVT applied to dY."""
    cout, cin = w.shape[0], w.shape[1]
    return pack_b(w, planes, cin, 8 * cout, 8, cout, cout, s_n0=64, s_tap=1, s_c=cin * 64, classes=8)


def pack_convt_fwd(w, planes):
    """nn.ConvTranspose3d weight [Cin,Cout,4,4,4] -> B[class][n=cout, k=(t,cin)] for SG_MODE_CONVT."""
    cin, cout = w.shape[0], w.shape[1]
    return pack_b(w, planes, cout, 8 * cin, 8, cin, cin, s_n0=64, s_tap=1, s_c=cout * 64, classes=8)


def pack_convt_dgrad(w, planes):
    """nn.ConvTranspose3d weight -> B[n=cin, k=(tap,cout)] for SG_MODE_CONV applied to dY."""
    cin, cout = w.shape[0], w.shape[1]
    return pack_b(w, planes, cin, 64 * cout, 64, cout, cout, s_n0=cout * 64, s_tap=1, s_c=64)


def pack_linear(w, planes, k_pad=None):
    """nn.Linear weight [out,in] -> B[n=out, k=in]."""
    out_f, in_f = w.shape
    k_pad = round_up(in_f, 64) if k_pad is None else k_pad
    return pack_b(w, planes, out_f, k_pad, 1, k_pad, in_f, s_n0=in_f, s_tap=0, s_c=1)


def pack_linear_dgrad(w, planes, n_pad=None):
    """nn.Linear weight [out,in] -> B[n=in, k=out] (dX = dY . W)."""
    out_f, in_f = w.shape
    return pack_b(w, planes, in_f, round_up(out_f, 64), 1, round_up(out_f, 64), out_f, s_n0=1, s_tap=0, s_c=in_f, n_pad=n_pad)


# ------------------------------------------------------------------------------------------------- implicit GEMM
_SPLITK_MAX_ELEMS = 148 * 128 * 256          # split-K only pays when the output has fewer tiles than SMs: skip the plan call otherwise


def igemm(mode, planes, a, a_dims, rows, k, b_img, n_valid, out, out_ld, out_kind=L.OUT_BF16, bias=None, act=L.ACT_NONE,
          a2=None, a2_c=0, mask=None, mask_act=L.ACT_NONE, out_dims=(0, 0, 0), bn=0, mt=0, ksplit=0, n_pad=None, bias_mod=0):
    _require_cuda(a, b_img, out)
    n_pad = round_up(n_valid, 16) if n_pad is None else n_pad
    n, d, h, w, c = a_dims
    args = L.SgIgemmArgs()
    args.mode, args.planes = mode, planes
    args.a = tdesc(a, n, d, h, w, c)
    args.a2 = tdesc(a2, 1, 1, 1, 1, a2_c) if a2 is not None else null_tensor()
    args.rows, args.k, args.n_pad, args.n_valid = rows, k, n_pad, n_valid
    args.bn, args.mt, args.ksplit = bn, mt, ksplit
    args.b_packed = ctypes.c_void_p(b_img.data_ptr())
    args.bias = ctypes.c_void_p(bias.data_ptr()) if bias is not None else None
    args.act = act
    args.bias_mod = bias_mod
    if mask is not None:
        args.mask = ctypes.c_void_p(mask.data_ptr())
        args.mask_plane_stride = mask.stride(0) if mask.shape[0] == 2 else 0
        args.mask_act = mask_act
    args.out = ctypes.c_void_p(out.data_ptr())
    args.out_plane_stride = out.stride(0) if (out.dtype == torch.bfloat16 and out.shape[0] == 2) else 0
    args.out_kind, args.out_ld = out_kind, out_ld
    args.out_d, args.out_h, args.out_w = out_dims
    if out_kind != L.OUT_F32_ATOMIC and rows * (8 if mode == L.MODE_CONVT else 1) * n_pad <= _SPLITK_MAX_ELEMS:
        # few output tiles x long K (e.g. Conv3d(128->256) 8^3 -> 4^3): the library may split K over fp32 partial slabs
        nbytes = ctypes.c_size_t(0)
        L.check(L.lib().sg_igemm_plan(ctypes.byref(args), ctypes.byref(nbytes)), 'sg_igemm_plan')
        if nbytes.value:
            ws = torch.empty(nbytes.value // 4, dtype=torch.float32, device=out.device)
            args.splitk_ws, args.splitk_ws_bytes = ctypes.c_void_p(ws.data_ptr()), nbytes.value
    L.check(L.lib().sg_igemm(ctypes.byref(args), stream()), 'sg_igemm')
    return out


def wgrad(b_mode, planes, a, a_c, b, b_dims, rows, grad, sm, st, sc, m_valid, cb=None, taps=None, accumulate=False,
          scale=1.0, merge_n=1, ksplit=0, c_valid=0, bias_grad=None, bias_accumulate=False):
    """grad[m*sm + tap*st + c*sc] (+)= sum_rows A[row,m] * gather(B)[row(+tap), c].
    bias_grad (fp32 [m_valid]): additionally (+)= the column sums of A -- the bias gradient when A is dY -- from the same pass over A."""
    _require_cuda(a, b, grad)
    n, d, h, w, c = b_dims
    args = L.SgWgradArgs()
    args.b_mode, args.planes = b_mode, planes
    args.a = tdesc(a, 1, 1, 1, 1, a_c)
    args.b = tdesc(b, n, d, h, w, c)
    args.rows, args.ksplit, args.merge_n = rows, ksplit, merge_n
    nbytes = ctypes.c_size_t(0)
    L.check(L.lib().sg_wgrad_plan(ctypes.byref(args), ctypes.byref(nbytes)), 'sg_wgrad_plan')
    ws = torch.empty(max(nbytes.value // 4, 1), dtype=torch.float32, device=grad.device)
    args.partials = ctypes.c_void_p(ws.data_ptr())
    bws = None
    if bias_grad is not None:
        bws = torch.empty(max(args.bias_ws_floats, 1), dtype=torch.float32, device=grad.device)
        args.bias_partials = ctypes.c_void_p(bws.data_ptr())
    L.check(L.lib().sg_wgrad(ctypes.byref(args), stream()), 'sg_wgrad')
    if b_mode == L.MODE_PATCH:
        taps_r, cb_r = 1, 64
    else:
        taps_r = 64 if b_mode == L.MODE_CONV else 1
        cb_r = c
    if taps is not None:
        taps_r, cb_r = taps, cb
    r = L.SgWgradReduceArgs(ctypes.c_void_p(ws.data_ptr()), args.ksplit_out, round_up(a_c, 128), m_valid, taps_r, cb_r,
                            sm, st, sc, ctypes.c_void_p(grad.data_ptr()), 1 if accumulate else 0, scale, c_valid)
    L.check(L.lib().sg_wgrad_reduce(ctypes.byref(r), stream()), 'sg_wgrad_reduce')
    if bws is not None:
        rb = L.SgWgradReduceArgs(ctypes.c_void_p(bws.data_ptr()), args.ksplit_out, round_up(a_c, 128), m_valid, 1, 1, 1, 0, 0,
                                 ctypes.c_void_p(bias_grad.data_ptr()), 1 if bias_accumulate else 0, scale, 0)
        L.check(L.lib().sg_wgrad_reduce(ctypes.byref(rb), stream()), 'sg_wgrad_reduce(bias)')
    return grad


def device_error_word():
    p = ctypes.c_void_p()
    L.check(L.lib().sg_device_error_word(ctypes.byref(p)), 'sg_device_error_word')
    return p.value


# ------------------------------------------------------------------------------------------------- HBM-bound companions
def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _ps(t):
    return t.stride(0) if (t is not None and t.shape[0] == 2) else 0


def _call(name, *args):
    L.check(getattr(L.lib(), name)(*args, stream()), name)


class _SumArena:
    """Zero-initialised double workspaces for the column-reducing kernels, cut from one memset'ed arena instead of one
    `torch.zeros` (= one fill launch) per reduction.  Every slice is handed out once.  An arena created while a CUDA graph
    is being captured belongs to that graph (its memset is a node of the graph and re-runs on every replay); it is never
    shared with work outside the capture or with another capture."""
    SIZE = 1 << 15

    def __init__(self):
        self.buf, self.off, self.captured, self.device = None, 0, 0, None

    def take(self, n, device):
        n_al = (n + 7) & ~7
        if n_al > self.SIZE:
            return torch.zeros(n, dtype=torch.float64, device=device)
        capturing = capture_id() if torch.device(device).type == 'cuda' else 0      # the capture's id, 0 = eager
        if self.buf is None or self.device != device or self.off + n_al > self.SIZE or self.captured != capturing:
            self.buf = torch.zeros(self.SIZE, dtype=torch.float64, device=device)
            self.off, self.captured, self.device = 0, capturing, device
        out = self.buf[self.off:self.off + n]
        self.off += n_al
        return out


_SUMS = _SumArena()


def dsums(n, device):
    return _SUMS.take(n, device)


def act_bwd(ga, y, act, c, want_sums=False, want_g=True):
    """g = ga * act'(y) over plane tensors viewed as [rows, c]; optional column sums (bias gradient)."""
    planes = ga.shape[0]
    rows = ga[0].numel() // c
    g = torch.empty_like(ga) if want_g else None
    sums = dsums(2 * c, ga.device) if want_sums else None
    _call('sg_act_bwd', _p(ga), _ps(ga), _p(y), _ps(y) if y is not None else 0, _p(g), _ps(g), planes, rows, c, act, _p(sums))
    return g, sums


def emit_sums(sums, dst, n, accumulate=False, scale=1.0, wc=0, s_t=0, s_c=1):
    _call('sg_emit_sums', _p(sums), _p(dst), n, 1 if accumulate else 0, scale, wc, s_t, s_c)
    return dst


def bn_forward(x, c, gamma, beta, act, running_mean, running_var, eps, momentum, training):
    """train: batch statistics (+ running-stat update); eval: running statistics.  Returns y, mean, invstd."""
    planes = x.shape[0]
    rows = x[0].numel() // c
    dev = x.device
    mean = torch.empty(c, dtype=torch.float32, device=dev)
    invstd = torch.empty(c, dtype=torch.float32, device=dev)
    if training:
        sums = dsums(2 * c, dev)
        _call('sg_bn_stats', _p(x), _ps(x), planes, rows, c, _p(sums))
        _call('sg_bn_finalize', _p(sums), rows, c, eps, momentum, _p(mean), _p(invstd), _p(running_mean), _p(running_var))
    else:
        # eval-mode statistics are tiny [C] vectors prepared with torch scalar math
        mean.copy_(running_mean)
        invstd.copy_(torch.rsqrt(running_var + eps))
    y = torch.empty_like(x)
    _call('sg_bn_apply', _p(x), _ps(x), _p(y), _ps(y), planes, rows, c, _p(mean), _p(invstd), _p(gamma), _p(beta), act)
    return y, mean, invstd


def bn_backward(ga, y, x, c, act, mean, invstd, gamma, training=True):
    """returns gx (planes), ggamma, gbeta (fp32 [c]); training=False: mean / invstd are the running statistics (constants)"""
    planes = x.shape[0]
    rows = x[0].numel() // c
    sums = dsums(2 * c, x.device)
    _call('sg_bn_bwd_reduce', _p(ga), _ps(ga), _p(y), _ps(y), _p(x), _ps(x), planes, rows, c, act, _p(mean), _p(invstd), _p(sums))
    gx = torch.empty_like(x)
    _call('sg_bn_bwd_apply', _p(ga), _ps(ga), _p(y), _ps(y), _p(x), _ps(x), _p(gx), _ps(gx), planes, rows, c, act, _p(mean),
          _p(invstd), _p(gamma), _p(sums) if training else None)
    gbeta = torch.empty(c, dtype=torch.float32, device=x.device)
    ggamma = torch.empty(c, dtype=torch.float32, device=x.device)
    emit_sums(sums, gbeta, c)
    emit_sums(sums[c:], ggamma, c)
    return gx, ggamma, gbeta


def col2im_c1(pm, n, d, h, w, bias, act):
    out = torch.empty((n, 2 * d, 2 * h, 2 * w), dtype=torch.float32, device=pm.device)
    _call('sg_col2im_c1', _p(pm), _ps(pm), pm.shape[0], n, d, h, w, _p(bias), act, _p(out))
    return out


def unary_f32(x, act):
    y = torch.empty_like(x)
    _call('sg_unary_f32', _p(x), _p(y), x.numel(), act)
    return y


def unary_bwd_f32(gy, y, act):
    gx = torch.empty_like(y)
    _call('sg_unary_bwd_f32', _p(gy.contiguous()), _p(y), _p(gx), y.numel(), act)
    return gx


def rowdot_fwd(x, c, w, bias, act, wc=0, s_t=0, s_c=1):
    rows = x[0].numel() // c
    y = torch.empty(rows, dtype=torch.float32, device=x.device)
    _call('sg_rowdot_fwd', _p(x), _ps(x), x.shape[0], rows, c, _p(w), wc, s_t, s_c, _p(bias), act, _p(y))
    return y


def rowdot_bwd(gy, y, act, x, c, w, need_gx, need_gw, planes, rows, wc=0, s_t=0, s_c=1, x_mask_act=L.ACT_NONE):
    dev = gy.device
    gx = torch.empty((planes, rows, c), dtype=torch.bfloat16, device=dev) if need_gx else None
    sums = dsums(c + 8, dev) if need_gw else None
    use_x = need_gw or x_mask_act != L.ACT_NONE
    _call('sg_rowdot_bwd', _p(gy), _p(y), act, _p(x) if use_x else None, _ps(x) if x is not None else 0, planes, rows, c, _p(w),
          wc, s_t, s_c, _p(gx), _ps(gx), _p(sums), x_mask_act)
    return gx, sums


def f32_to_planes(src, planes, c_dst=None):
    """fp32 [rows, c_src] (row stride = src.stride(0)) -> planes [P, rows, c_dst] (zero padded)."""
    rows, c_src = src.shape
    c_dst = round_up(c_src, 8) if c_dst is None else c_dst
    assert src.stride(1) == 1
    dst = torch.empty((planes, rows, c_dst), dtype=torch.bfloat16, device=src.device)
    _call('sg_to_planes', _p(src), src.stride(0), rows, c_src, _p(dst), _ps(dst), planes, c_dst)
    return dst


def planes_to_f32(src, c_src, c_take=None, out=None, accumulate=False, scale=1.0):
    rows = src[0].numel() // c_src
    c_take = c_src if c_take is None else c_take
    if out is None:
        out = torch.empty((rows, c_take), dtype=torch.float32, device=src.device)
    _call('sg_from_planes', _p(src), _ps(src), src.shape[0], rows, c_src, c_take, _p(out), out.stride(0), 1 if accumulate else 0, scale)
    return out


def sdf_pack_input(points, latent, index, L_, planes, c_dst):
    n = points.shape[0]
    dst = torch.empty((planes, n, c_dst), dtype=torch.bfloat16, device=points.device)
    _call('sg_sdf_pack_input', _p(points), _p(latent), _p(index), L_, n, _p(dst), _ps(dst), planes, c_dst)
    return dst


def sdf_unpack_grad(ga, gb, c_src, L_, index, gpoints, glatent):
    n = ga.shape[1]
    _call('sg_sdf_unpack_grad', _p(ga), _ps(ga), _p(gb), _ps(gb), ga.shape[0], n, c_src, L_, _p(index), _p(gpoints), _p(glatent))


def fade_fwd(x, b, r, c, vol, f):
    y = torch.empty_like(x)
    _call('sg_fade_fwd', _p(x), _ps(x), _p(y), _ps(y), x.shape[0], b, r, c, _p(vol), f)
    return y


def fade_bwd_vol(g, b, r, c, f, gvol):
    _call('sg_fade_bwd_vol', _p(g), _ps(g), g.shape[0], b, r, c, f, _p(gvol))
    return gvol


def axpby_planes(a, alpha, b=None, beta=0.0):
    y = torch.empty_like(a)
    _call('sg_axpby_planes', _p(a), _ps(a), _p(b), _ps(b), _p(y), _ps(y), a.shape[0], a[0].numel(), alpha, beta)
    return y


def rmsprop(p, g, sq, lr, alpha=0.99, eps=1e-8, grad_scale=1.0, clip=0.0):
    _call('sg_rmsprop', _p(p), _p(g), _p(sq), p.numel(), lr, alpha, eps, grad_scale, clip)


def adam(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0):
    _call('sg_adam', _p(p), _p(g), _p(m), _p(v), p.numel(), lr, b1, b2, eps, step, grad_scale)


def dp_step(peers, n, s1, s2, kind, lr, step, grad_scale=1.0, clip=0.0, alpha=0.99, b1=0.9, b2=0.999, eps=1e-8):
    """fused data-parallel optimizer step over peer memory (sg_dp_step); `peers` = train._PeerArenas"""
    a = L.SgDpStepArgs()
    a.peer_grad = ctypes.cast(peers.peer_grad, ctypes.POINTER(ctypes.c_void_p))
    a.peer_param = ctypes.cast(peers.peer_param, ctypes.POINTER(ctypes.c_void_p))
    a.peer_pad = ctypes.cast(peers.peer_pad, ctypes.POINTER(ctypes.c_void_p))
    a.rank, a.world, a.n, a.chunk = peers.rank, peers.world, n, peers.chunk
    a.s1, a.s2 = _p(s1), _p(s2)
    a.kind = 0 if kind == 'rmsprop' else 1
    a.lr, a.beta1, a.beta2, a.eps, a.clip, a.grad_scale = lr, (alpha if kind == 'rmsprop' else b1), b2, eps, clip, grad_scale
    a.step, a.sync = step, _p(peers.sync)
    L.check(L.lib().sg_dp_step(ctypes.byref(a), stream()), 'sg_dp_step')


def clamp_(p, lo, hi):
    _call('sg_clamp', _p(p), p.numel(), lo, hi)


def l1_loss_grad(out, target, want_grad=True):
    gout = torch.empty_like(out) if want_grad else None
    loss = dsums(1, out.device)
    _call('sg_l1_loss_grad', _p(out), _p(target), _p(gout), out.numel(), _p(loss))
    return loss, gout


def sum_f32(x):
    out = dsums(1, x.device)
    _call('sg_sum_f32', _p(x), x.numel(), _p(out))
    return out


def sdfnet_fwd(points, latent, index, w_img, aux, stash=None, mask_stash=None):
    n = points.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=points.device)
    a = L.SgSdfnetFwdArgs(_p(points), _p(latent), _p(index), n, _p(w_img), _p(aux), _p(out), _p(stash), _p(mask_stash))
    L.check(L.lib().sg_sdfnet_fwd(ctypes.byref(a), stream()), 'sg_sdfnet_fwd')
    return out


def sdfnet_bwd(gout, out, mask_stash, wt_img, w8, xyz_w=None, want_gstash=True):
    """Fused input-gradient chain (sg_sdfnet.cu): returns gstash bf16 [7, n, 256] = gradients w.r.t. the pre-activations
    of layers 1..7 (None when want_gstash is False) and, with xyz_w [2,3,256], the gradient w.r.t. the points [n, 3]."""
    n = out.shape[0]
    gstash = torch.empty((7, n, 256), dtype=torch.bfloat16, device=out.device) if want_gstash else None
    gpoints = torch.empty((n, 3), dtype=torch.float32, device=out.device) if xyz_w is not None else None
    a = L.SgSdfnetBwdArgs(_p(gout), _p(out), _p(mask_stash), _p(wt_img), _p(w8), n, _p(gstash), _p(gpoints), _p(xyz_w))
    L.check(L.lib().sg_sdfnet_bwd(ctypes.byref(a), stream()), 'sg_sdfnet_bwd')
    return (gstash, gpoints) if xyz_w is not None else gstash


def sdfnet_infer(w_img, aux, n, out=None, points=None, ray_index=None, n_ptr=None, grid_r=0, grid_axis=None, mask_stash=None, trace=None):
    """Single-latent fused forward (sg_sdfnet_infer).  `trace` = dict(points, dirs, hit, next_index, next_count, sdf_offset, clamp,
    threshold, radius, miss_y) turns the call into one sphere-tracing step."""
    a = L.SgSdfnetInferArgs()
    a.points, a.n, a.n_ptr, a.ray_index = _p(points), n, _p(n_ptr), _p(ray_index)
    a.grid_r, a.grid_axis = grid_r, _p(grid_axis)
    a.w_img, a.aux, a.out, a.mask_stash = _p(w_img), _p(aux), _p(out), _p(mask_stash)
    if trace is not None:
        a.trace_points, a.trace_dirs, a.trace_hit = _p(trace['points']), _p(trace['dirs']), _p(trace['hit'])
        a.next_index, a.next_count = _p(trace['next_index']), _p(trace['next_count'])
        a.sdf_offset, a.trace_clamp, a.trace_threshold, a.trace_radius = trace['sdf_offset'], trace['clamp'], trace['threshold'], trace['radius']
        a.trace_miss_y = 1 if trace.get('miss_y') else 0
    L.check(L.lib().sg_sdfnet_infer(ctypes.byref(a), stream()), 'sg_sdfnet_infer')
    return out


def segment_colsum(x, segs, seg_len):
    """fp32 [segs, C] column sums of the consecutive seg_len-row segments of a plane tensor [P, segs * seg_len, C]"""
    p, _, c = x.shape
    out = torch.empty((segs, c), dtype=torch.float32, device=x.device)
    nbytes = L.lib().sg_segment_colsum_workspace(segs, c, seg_len)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x.device) if nbytes else None
    _call('sg_segment_colsum', _p(x), _ps(x), p, segs, c, seg_len, _p(out), _p(ws))
    return out


def grid_sphere_index(r, axis, radius):
    """int32 list (sorted) of the cells of the R^3 grid inside the sphere; `axis` fp32 [3, r] on the device"""
    idx = torch.empty(r * r * r, dtype=torch.int32, device=axis.device)
    count = torch.zeros(1, dtype=torch.int32, device=axis.device)
    _call('sg_grid_sphere_index', r, _p(axis), float(radius), _p(idx), _p(count))
    return torch.sort(idx[:int(count.item())]).values.contiguous()


def voxel_ingest(src, clamp=0.1, rescale=True, out=None):
    out = torch.empty_like(src) if out is None else out
    _call('sg_voxel_ingest', _p(src), _p(out), src.numel(), float(clamp), 1 if rescale else 0)
    return out
