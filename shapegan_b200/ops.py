"""Layer operators of the shapegan hot path as twice-differentiable torch.autograd Functions over libsg_b200.

Every linear layer kind is a `LinOp` with three primitives implemented by the tcgen05 kernels:
    fwd(x, W)   y  = L_W x          (sg_igemm)
    tr(g, W)    gx = L_W^T g        (sg_igemm, the transposed gather)
    wgrad(x,g)  gW = dL/dW          (sg_wgrad + sg_wgrad_reduce)
and three generic Functions (`_Fwd`, `_Tr`, `_Wgrad`) wire them into autograd so that
`autograd.grad(..., create_graph=True)` followed by `.backward()` (the WGAN-GP of
train_hybrid_progressive_gan.py:102-111) composes out of the same kernels: the discriminators are
piecewise linear (Conv/Linear + LeakyReLU, no BatchNorm), so the double backward is
    d/dg  of L_W^T g  -> L_W          d/dW of <c, L_W^T g> -> wgrad(c, g)
with the LeakyReLU masks held fixed.

Tensors between layers are NDHWC bf16 plane tensors [P, B, D, H, W, C] (P = 1 bf16, P = 2 hi/lo fp32x);
single-channel voxel volumes at module boundaries are fp32 [B, D, H, W]."""
import os
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L
from . import config, raw

ACT_NONE, ACT_LRELU, ACT_RELU, ACT_TANH, ACT_SIGMOID = L.ACT_NONE, L.ACT_LRELU, L.ACT_RELU, L.ACT_TANH, L.ACT_SIGMOID
r64 = lambda v: raw.round_up(v, 64)   # noqa: E731


# ------------------------------------------------------------------------------------------------- packed-weight cache
# measurement only (results are WRONG): never re-pack a weight once packed -- the step-time difference is the cost of the per-step re-packs
_STALE_PACK = os.environ.get('SG_B200_DIAG_STALE_PACK') == '1'


class _PackCache:
    """fp32 parameter -> tensor-core operand image, cached per parameter OBJECT (a uid stamped on the tensor; a
    device address can be recycled by the caching allocator, so it is not an identity) and validated by
    (`_version`, data_ptr, shape) AND the CUDA-graph capture the image was packed in (0 = eager).  An optimizer step or
    load_state_dict bumps `_version`; `.data` mutations do not: call invalidate() (our clip_weights does).
    An image packed outside a capture is never reused inside one (and vice versa): the pack kernel has to be a node of
    the graph, or every replay would run on the weights of capture time while the optimizer keeps updating the real ones.
    Entries die with their parameter (weakref.finalize)."""

    def __init__(self):
        self.store = {}
        self.next_uid = 1

    def _uid(self, w):
        uid = getattr(w, '_sg_uid', None)
        if uid is None:
            uid = self.next_uid
            self.next_uid += 1
            try:
                w._sg_uid = uid
                weakref.finalize(w, self._drop, uid)
            except Exception:               # un-stampable tensor: never cached
                return None
        return uid

    def _drop(self, uid):
        for key in [k for k in self.store if k[0] == uid]:
            self.store.pop(key, None)

    def lookup(self, key, sig):
        """cached payload for `key` if it was built from `sig` in the current capture context, else None"""
        hit = self.store.get(key)
        if hit is not None and (_STALE_PACK or (hit[0] == sig and hit[2] == raw.capture_id())):
            return hit[1]
        return None

    def put(self, key, sig, payload):
        self.store[key] = (sig, payload, raw.capture_id())

    def get(self, w, kind, planes, fn):
        uid = self._uid(w)
        if uid is None:
            return fn(w.detach(), planes)
        key = (uid, kind, planes)
        sig = (w._version, w.data_ptr(), tuple(w.shape))
        img = self.lookup(key, sig)
        if img is None:
            img = fn(w.detach(), planes)
            self.put(key, sig, img)
        return img

    def invalidate(self):
        self.store.clear()


PACK_CACHE = _PackCache()


def invalidate_weight_cache():
    PACK_CACHE.invalidate()


def _planes():
    return config.planes()


# First-order backward passes of the Step classes (train.py) let the layer Functions accumulate weight / bias gradients
# straight into the allocated `.grad` views of the flat gradient arena (what AccumulateGrad would do with a returned tensor,
# minus the temporary and the add kernel).  Opt-in: `torch.autograd.grad(out, x)` or `backward(inputs=[x])` through a layer
# must NOT touch parameter gradients, and ctx.needs_input_grad cannot tell those calls from a plain `.backward()`.
_ARENA_ACCUMULATE = [False]


class arena_backward:
    """with ops.arena_backward(): loss.backward()   -- parameter gradients accumulate in place into their `.grad`"""

    def __enter__(self):
        self.prev = _ARENA_ACCUMULATE[0]
        _ARENA_ACCUMULATE[0] = True

    def __exit__(self, *exc):
        _ARENA_ACCUMULATE[0] = self.prev
        return False


def _into_grad(p, numel=None):
    """the parameter's allocated .grad if in-place accumulation is active and legal for this backward, else None"""
    if not _ARENA_ACCUMULATE[0] or torch.is_grad_enabled():
        return None
    g = p.grad
    if g is None or not g.is_contiguous() or (numel is not None and g.numel() != numel):
        return None
    return g


def _new(shape, device):
    return torch.empty((_planes(),) + tuple(shape), dtype=torch.bfloat16, device=device)


# ------------------------------------------------------------------------------------------------- linear operators
class LinOp:
    """Shape bookkeeping + the three kernel-backed primitives of one linear layer kind."""
    name = 'linop'

    def fwd(self, x, w, bias, act):
        raise NotImplementedError

    def tr(self, g, w):
        raise NotImplementedError

    def wgrad(self, x, g, w_shape):
        raise NotImplementedError

    def out_channels(self):
        raise NotImplementedError


class ConvOp(LinOp):
    """nn.Conv3d(cin, cout, 4, stride 2, padding 1), cin % 8 == 0   (model/gan.py:51,53; progressive_gan.py:38)."""

    def __init__(self, cin, cout):
        self.cin, self.cout = cin, cout

    def fwd(self, x, w, bias, act, out=None, mask=None, mask_act=ACT_NONE):
        """`out` (optional) receives the result; `mask`: out *= act'(mask) (an activation backward fused into the epilogue; may alias out)"""
        p, b, d, h, wd, c = x.shape
        rows = b * (d // 2) * (h // 2) * (wd // 2)
        y = out if out is not None else _new((b, d // 2, h // 2, wd // 2, self.cout), x.device)
        img = PACK_CACHE.get(w, 'conv_fwd', p, raw.pack_conv_fwd)
        raw.igemm(L.MODE_CONV, p, x, (b, d, h, wd, c), rows, 64 * c, img, self.cout, y, self.cout, bias=bias, act=act, mask=mask, mask_act=mask_act)
        return y

    def tr(self, g, w, mask=None, mask_act=ACT_NONE):
        p, b, d, h, wd, c = g.shape          # g: [P,B,Do,Ho,Wo,Cout]
        gx = _new((b, 2 * d, 2 * h, 2 * wd, self.cin), g.device)
        img = PACK_CACHE.get(w, 'conv_dgrad', p, raw.pack_conv_dgrad)
        raw.igemm(L.MODE_CONVT, p, g, (b, d, h, wd, c), b * d * h * wd, 8 * c, img, self.cin, gx, self.cin,
                  out_dims=(2 * d, 2 * h, 2 * wd), mask=mask, mask_act=mask_act)
        return gx

    bias_in_wgrad = True     # A operand of the weight-gradient GEMM is dY: its column sums (the bias gradient) come out of the same pass

    def wgrad(self, x, g, w_shape, into=None, bias_into=None):
        p, b, d, h, wd, c = x.shape
        gw = into if into is not None else torch.empty(w_shape, dtype=torch.float32, device=x.device)
        raw.wgrad(L.MODE_CONV, p, g, self.cout, x, (b, d, h, wd, c), g[0].numel() // self.cout, gw,
                  sm=self.cin * 64, st=1, sc=64, m_valid=self.cout, accumulate=into is not None,
                  bias_grad=bias_into, bias_accumulate=True)
        return gw

    def wgrad_into(self, x, g, grad, bias_into=None):
        self.wgrad(x, g, None, into=grad, bias_into=bias_into)

    def out_channels(self):
        return self.cout


class Conv1Op(LinOp):
    """nn.Conv3d(1, cout, 4, 2, 1) over an fp32 voxel volume [B,D,H,W] (model/gan.py:49, progressive_gan.py:38 after
    from_SDF zero-padding, autoencoder.py:16).  `w` is the [cout, 1, 4,4,4] slice of the weight."""

    def __init__(self, cout, w_cin=1):
        # w_cin > 1: the full [cout, w_cin, 4,4,4] weight of a from_SDF layer, of which only input channel 0 sees
        # non-zero data (progressive_gan.py:15); addressing uses its strides, the gradient of the other channels is 0.
        self.cout, self.w_cin = cout, w_cin

    def fwd(self, x, w, bias, act, out=None, mask=None, mask_act=ACT_NONE):
        b, d, h, wd = x.shape
        p = _planes()
        rows = b * (d // 2) * (h // 2) * (wd // 2)
        y = out if out is not None else _new((b, d // 2, h // 2, wd // 2, self.cout), x.device)
        img = PACK_CACHE.get(w, 'conv1_fwd', p, lambda t, pl: raw.pack_b(
            t, pl, self.cout, 64, 64, 1, 1, s_n0=self.w_cin * 64, s_tap=1, s_c=0))
        raw.igemm(L.MODE_PATCH, p, x, (b, d, h, wd, 1), rows, 64, img, self.cout, y, self.cout, bias=bias, act=act, mask=mask, mask_act=mask_act)
        return y

    def tr(self, g, w):
        # dX = col2im( dY[rows, cout] . W[cout, 64 taps] )
        p, b, d, h, wd, c = g.shape
        rows = b * d * h * wd
        img = PACK_CACHE.get(w, 'c1_taps', p, lambda t, pl: raw.pack_b(
            t, pl, 64, r64(c), 1, r64(c), c, s_n0=1, s_tap=0, s_c=self.w_cin * 64))
        pm = _new((rows, 64), g.device)
        raw.igemm(L.MODE_DENSE, p, g, (1, 1, 1, 1, c), rows, r64(c), img, 64, pm, 64)
        return raw.col2im_c1(pm, b, d, h, wd, None, ACT_NONE)

    def wgrad(self, x, g, w_shape, into=None):
        """into: accumulate into an existing gradient (channels > 0 of a from_SDF weight receive nothing, i.e. stay as they are)"""
        b, d, h, wd = x.shape
        gw = into if into is not None else (torch.zeros if self.w_cin > 1 else torch.empty)(w_shape, dtype=torch.float32, device=x.device)
        raw.wgrad(L.MODE_PATCH, g.shape[0], g, self.cout, x, (b, d, h, wd, 1), g[0].numel() // self.cout, gw,
                  sm=self.w_cin * 64, st=0, sc=1, m_valid=self.cout, accumulate=into is not None)
        return gw

    def out_channels(self):
        return self.cout


class ConvTOp(LinOp):
    """nn.ConvTranspose3d(cin, cout, 4, stride 2, padding 1), cout % 8 == 0  (model/gan.py:13,17; autoencoder.py:55,59)."""

    def __init__(self, cin, cout):
        self.cin, self.cout = cin, cout

    def fwd(self, x, w, bias, act):
        p, b, d, h, wd, c = x.shape
        y = _new((b, 2 * d, 2 * h, 2 * wd, self.cout), x.device)
        img = PACK_CACHE.get(w, 'convt_fwd', p, raw.pack_convt_fwd)
        raw.igemm(L.MODE_CONVT, p, x, (b, d, h, wd, c), b * d * h * wd, 8 * c, img, self.cout, y, self.cout, bias=bias, act=act,
                  out_dims=(2 * d, 2 * h, 2 * wd))
        return y

    def tr(self, g, w):
        p, b, d, h, wd, c = g.shape          # [P,B,2D,2H,2W,Cout]
        rows = b * (d // 2) * (h // 2) * (wd // 2)
        gx = _new((b, d // 2, h // 2, wd // 2, self.cin), g.device)
        img = PACK_CACHE.get(w, 'convt_dgrad', p, raw.pack_convt_dgrad)
        raw.igemm(L.MODE_CONV, p, g, (b, d, h, wd, c), rows, 64 * c, img, self.cin, gx, self.cin)
        return gx

    def wgrad(self, x, g, w_shape, into=None):
        p, b, d, h, wd, c = g.shape
        gw = into if into is not None else torch.empty(w_shape, dtype=torch.float32, device=x.device)
        raw.wgrad(L.MODE_CONV, p, x, self.cin, g, (b, d, h, wd, c), x[0].numel() // self.cin, gw,
                  sm=self.cout * 64, st=1, sc=64, m_valid=self.cin, accumulate=into is not None)
        return gw

    def wgrad_into(self, x, g, grad):
        self.wgrad(x, g, None, into=grad)

    def out_channels(self):
        return self.cout


class ConvT1Op(LinOp):
    """nn.ConvTranspose3d(cin, 1, 4, 2, 1) -> fp32 volume [B,2D,2H,2W]  (model/gan.py:21, autoencoder.py:63):
    tensor-core projection onto the 64 taps, then the 8-tap col2im gather (+bias, +tanh)."""

    def __init__(self, cin):
        self.cin = cin

    def fwd(self, x, w, bias, act):
        p, b, d, h, wd, c = x.shape
        rows = b * d * h * wd
        img = PACK_CACHE.get(w, 'c1_taps', p, lambda t, pl: raw.pack_b(
            t, pl, 64, r64(c), 1, r64(c), c, s_n0=1, s_tap=0, s_c=64))
        pm = _new((rows, 64), x.device)
        raw.igemm(L.MODE_DENSE, p, x, (1, 1, 1, 1, c), rows, r64(c), img, 64, pm, 64)
        return raw.col2im_c1(pm, b, d, h, wd, bias, act)

    def tr(self, g, w):
        b, d, h, wd = g.shape                # fp32 volume [B,2D,2H,2W]
        p = _planes()
        rows = b * (d // 2) * (h // 2) * (wd // 2)
        gx = _new((b, d // 2, h // 2, wd // 2, self.cin), g.device)
        img = PACK_CACHE.get(w, 'conv_fwd', p, raw.pack_conv_fwd)      # [cin,1,4,4,4] has the Conv3d(1->cin) layout
        raw.igemm(L.MODE_PATCH, p, g, (b, d, h, wd, 1), rows, 64, img, self.cin, gx, self.cin)
        return gx

    def wgrad(self, x, g, w_shape):
        b, d, h, wd = g.shape
        gw = torch.empty(w_shape, dtype=torch.float32, device=x.device)
        raw.wgrad(L.MODE_PATCH, x.shape[0], x, self.cin, g, (b, d, h, wd, 1), x[0].numel() // self.cin, gw,
                  sm=64, st=0, sc=1, m_valid=self.cin)
        return gw

    def out_channels(self):
        return 1


class DenseOp(LinOp):
    """y[row, (n1,n0)] = sum_{t,c} x[row, (t,c)] * W.flat[n1*s_n1 + n0*s_n0 + t*s_t + c*s_c]

    Covers nn.Linear, ConvTranspose3d(k4,s1) on a 1^3 grid (gan.py:9: n1 = position, n0 = cout), Conv3d(k4,s1) on a
    4^3 grid (autoencoder.py:28: t = position, c = cin) and the NCDHW-flatten + Linear head of progressive_gan.py:27-28
    (t = position, c = channel).  Input/output are plane tensors [P, rows, K] / [P, rows, N]."""

    def __init__(self, n1, n0, s_n1, s_n0, t, c, s_t, s_c, tag, c_valid=None):
        """c_valid < c: the input rows carry c physical columns of which only the first c_valid exist in the weight (zero padding of
        narrow inputs such as xyz / xyz+dist up to the kernels' K granularity); only with t == 1 and n1 == 1"""
        self.n1, self.n0, self.s_n1, self.s_n0 = n1, n0, s_n1, s_n0
        self.t, self.c, self.s_t, self.s_c = t, c, s_t, s_c
        self.n, self.k = n1 * n0, t * c
        self.tag = tag
        self.c_valid = c if c_valid is None else c_valid
        assert self.c_valid == c or (t == 1 and n1 == 1)
        assert self.n % 8 == 0 and self.k % 8 == 0, 'DenseOp needs N, K multiples of 8'
        assert n1 == 1 or t == 1, 'DenseOp: only one side may be two-level'

    def fwd(self, x, w, bias, act, out=None, mask=None, mask_act=ACT_NONE):
        p, rows = x.shape[0], x.shape[1]
        y = out if out is not None else _new((rows, self.n), x.device)
        img = PACK_CACHE.get(w, self.tag + '_f', p, lambda tt, pl: raw.pack_b(
            tt, pl, self.n, r64(self.k), self.t, self.c, self.c_valid, s_n0=self.s_n0, s_tap=self.s_t, s_c=self.s_c,
            n0_count=self.n0, s_n1=self.s_n1))
        raw.igemm(L.MODE_DENSE, p, x, (1, 1, 1, 1, self.k), rows, r64(self.k), img, self.n, y, self.n, bias=bias, act=act,
                  bias_mod=self.n0 if (self.n1 > 1 and bias is not None) else 0, mask=mask, mask_act=mask_act)
        return y

    def tr(self, g, w, mask=None, mask_act=ACT_NONE):
        p, rows = g.shape[0], g.shape[1]
        gx = _new((rows, self.k), g.device)
        kv = self.c_valid if self.c_valid != self.c else self.k         # rows >= c_valid of the image are zero: those gx columns come out 0
        img = PACK_CACHE.get(w, self.tag + '_t', p, lambda tt, pl: raw.pack_b(
            tt, pl, kv, r64(self.n), self.n1, self.n0, self.n0, s_n0=self.s_c, s_tap=self.s_n1, s_c=self.s_n0,
            n0_count=self.c, s_n1=self.s_t, n_pad=raw.round_up(self.k, 16)))
        raw.igemm(L.MODE_DENSE, p, g, (1, 1, 1, 1, self.n), rows, r64(self.n), img, self.k, gx, self.k, mask=mask, mask_act=mask_act,
                  n_pad=raw.round_up(self.k, 16))
        return gx

    def wgrad(self, x, g, w_shape, into=None):
        p, rows = x.shape[0], x.shape[1]
        gw = into if into is not None else torch.empty(w_shape, dtype=torch.float32, device=x.device)
        acc = into is not None
        if self.n1 == 1:      # A = g (one-level n0), B = x (t, c)
            raw.wgrad(L.MODE_DENSE, p, g, self.n, x, (1, 1, 1, 1, self.k), rows, gw, sm=self.s_n0, st=self.s_t, sc=self.s_c,
                      m_valid=self.n, taps=self.t, cb=self.c, accumulate=acc, c_valid=self.c_valid if self.c_valid != self.c else 0)
        else:                 # A = x (one-level c), B = g (n1, n0)
            raw.wgrad(L.MODE_DENSE, p, x, self.k, g, (1, 1, 1, 1, self.n), rows, gw, sm=self.s_c, st=self.s_n1, sc=self.s_n0,
                      m_valid=self.k, taps=self.n1, cb=self.n0, accumulate=acc)
        return gw

    def out_channels(self):
        return self.n0


def linear_op(in_f, out_f, tag='lin'):
    """nn.Linear(in_f, out_f): weight [out, in]."""
    return DenseOp(1, out_f, 0, in_f, 1, in_f, 0, 1, '%s_%d_%d' % (tag, in_f, out_f))


# ------------------------------------------------------------------------------------------------- generic Functions
def _bias_grad(g, c, into=None):
    """column sums of a plane tensor viewed as [rows, c] -> fp32 [c] (accumulated into `into` when given)"""
    _, sums = raw.act_bwd(g, None, ACT_NONE, c, want_sums=True, want_g=False)
    if into is not None:
        return raw.emit_sums(sums, into, c, accumulate=True)
    out = torch.empty(c, dtype=torch.float32, device=g.device)
    return raw.emit_sums(sums, out, c)


class _MaskMul(Function):
    """g = ga * act'(y) with y the stored activation OUTPUT (piecewise-linear acts: the mask is a constant of the
    double backward, so this Function is its own derivative).  The same pass can emit the column sums of g (the bias
    gradient), returned as a non-differentiable second output."""

    @staticmethod
    def forward(ctx, ga, y, act, c, want_bias, bias_into=None):
        ctx.act, ctx.c = act, c
        ctx.save_for_backward(y)
        g, sums = raw.act_bwd(ga.contiguous(), y, act, c, want_sums=want_bias)
        if want_bias and bias_into is not None:       # accumulate straight into an allocated .grad (flat gradient arena)
            raw.emit_sums(sums, bias_into, c, accumulate=True)
            gb = torch.empty(0, dtype=torch.float32, device=ga.device)
        else:
            gb = torch.empty(c if want_bias else 0, dtype=torch.float32, device=ga.device)
            if want_bias:
                raw.emit_sums(sums, gb, c)
        ctx.mark_non_differentiable(gb)
        return g, gb

    @staticmethod
    def backward(ctx, gg, _ggb):
        (y,) = ctx.saved_tensors
        return _MaskMul.apply(gg, y, ctx.act, ctx.c, False)[0], None, None, None, None, None


class _Fwd(Function):
    """y = act(L_W x + b)"""

    @staticmethod
    def forward(ctx, op, x, w, bias, act):
        x = x.contiguous()
        y = op.fwd(x, w, bias, act)
        ctx.op, ctx.act, ctx.has_bias = op, act, bias is not None
        ctx.bias_obj = bias
        ctx.w_obj = w                       # identity for the pack cache; saved_tensors still does the version check
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, _w_checked, y = ctx.saved_tensors
        w = ctx.w_obj
        op = ctx.op
        gy = gy.contiguous()
        c = op.out_channels()
        want_b = ctx.has_bias and ctx.needs_input_grad[3]
        gb_fused = None
        # first-order backward with an allocated bias .grad (flat gradient arena): the bias sums accumulate in place, like the
        # weight gradient below -- no temporary, no AccumulateGrad add kernel
        b_into = _into_grad(ctx.bias_obj, c) if want_b else None
        # arena backward of a layer whose weight-gradient GEMM has dY as its A operand: the bias column sums ride that GEMM
        bias_via_wgrad = (b_into is not None and getattr(op, 'bias_in_wgrad', False) and ctx.needs_input_grad[2]
                          and _into_grad(w) is not None)
        if ctx.act != ACT_NONE:
            g, gb_fused = _MaskMul.apply(gy, y, ctx.act, c, want_b and not bias_via_wgrad, b_into)      # activation backward (+ bias column sums)
        else:
            g = gy
        gx = _Tr.apply(op, g, w) if ctx.needs_input_grad[1] else None
        gw = None
        if ctx.needs_input_grad[2]:
            w_into = _into_grad(w) if hasattr(op, 'wgrad_into') else None
            if w_into is not None:
                # first-order backward of a Step class with an allocated .grad (the flat gradient arena): accumulate in place,
                # exactly what AccumulateGrad would do with the returned tensor, minus the temporary and the extra add kernel
                if bias_via_wgrad:
                    op.wgrad_into(x, g, w_into, bias_into=b_into)
                else:
                    op.wgrad_into(x, g, w_into)
            else:
                gw = _Wgrad.apply(op, x, g, w)
        gb = None
        if want_b and not bias_via_wgrad:
            # the bias gradient is never differentiated again on this path (the GP contributes exactly zero to biases)
            if gb_fused is not None:
                gb = None if b_into is not None else gb_fused
            elif b_into is not None:
                _bias_grad(g.detach(), c, into=b_into)
            else:
                gb = _bias_grad(g.detach(), c)
        return None, gx, gw, gb, None


class _Tr(Function):
    """gx = L_W^T g"""

    @staticmethod
    def forward(ctx, op, g, w):
        g = g.contiguous()
        ctx.op = op
        ctx.w_obj = w
        ctx.save_for_backward(g, w)
        return op.tr(g, w)

    @staticmethod
    def backward(ctx, ggx):
        g, _w_checked = ctx.saved_tensors
        w = ctx.w_obj
        op = ctx.op
        ggx = ggx.contiguous()
        g_g = _Fwd.apply(op, ggx, w, None, ACT_NONE) if ctx.needs_input_grad[1] else None
        g_w = _Wgrad.apply(op, ggx, g, w) if ctx.needs_input_grad[2] else None
        return None, g_g, g_w


class _Wgrad(Function):
    """gW = wgrad(x, g).  Differentiating THROUGH a weight gradient is not needed by any script of the reference;
    it fails loudly instead of silently returning zero."""

    @staticmethod
    def forward(ctx, op, x, g, w):
        return op.wgrad(x.contiguous(), g.contiguous(), tuple(w.shape))

    @staticmethod
    def backward(ctx, ggw):
        raise NotImplementedError('shapegan_b200: third-order path (gradient of a weight gradient) is not implemented')


def linear_layer(op, x, w, bias, act=ACT_NONE):
    return _Fwd.apply(op, x, w, bias, act)


# ------------------------------------------------------------------------------------------------- BatchNorm (+act)
class _BatchNormAct(Function):
    """train/eval BatchNorm{1,3}d over a plane tensor viewed as [rows, C] followed by an activation
    (model/gan.py:10-11 etc.).  Running statistics are updated in place (momentum 0.1, unbiased variance)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, eps, momentum, act, c):
        x = x.contiguous()
        y, mean, invstd = raw.bn_forward(x, c, gamma.detach(), beta.detach(), act, running_mean, running_var, eps, momentum, training)
        ctx.act, ctx.c, ctx.training = act, c, training
        ctx.save_for_backward(x, y, mean, invstd, gamma)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, y, mean, invstd, gamma = ctx.saved_tensors
        gx, ggamma, gbeta = raw.bn_backward(gy.contiguous(), y, x, ctx.c, ctx.act, mean, invstd, gamma.detach(), training=ctx.training)
        return gx, ggamma, gbeta, None, None, None, None, None, None, None


def batchnorm_act(x, bn, act, c):
    """`bn` is the nn.BatchNorm{1,3}d parameter container."""
    training = bn.training
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1          # int64 scalar bookkeeping, kept for state_dict parity
    return _BatchNormAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, bn.eps,
                               bn.momentum if bn.momentum is not None else 0.1, act, c)


# ------------------------------------------------------------------------------------------------- N = 1 layers
class _RowDot(Function):
    """y[r] = act(x[r,:].w + b) for a plane tensor x [P, rows, C]; w is addressed through (wc, s_t, s_c) so torch
    weight layouts need no permuted copy.  Twice differentiable wrt x and w for ACT_NONE (the critic/GP path)."""

    @staticmethod
    def forward(ctx, x, w, bias, act, wc, s_t, s_c):
        x = x.contiguous()
        c = x.shape[2]
        y = raw.rowdot_fwd(x, c, w.detach(), bias.detach() if bias is not None else None, act, wc, s_t, s_c)
        ctx.meta = (act, wc, s_t, s_c, c, bias is not None)
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        act, wc, s_t, s_c, c, has_bias = ctx.meta
        if act != ACT_NONE and torch.is_grad_enabled():
            raise NotImplementedError('shapegan_b200: double backward through a non-linear output activation')
        gx, gw, gb = _RowDotBwd.apply(gy.contiguous(), y, x, w, act, wc, s_t, s_c, ctx.needs_input_grad[0],
                                      ctx.needs_input_grad[1], has_bias)
        return (gx if ctx.needs_input_grad[0] else None, gw if ctx.needs_input_grad[1] else None,
                gb if (has_bias and ctx.needs_input_grad[2]) else None, None, None, None, None)


class _RowDotBwd(Function):
    @staticmethod
    def forward(ctx, gy, y, x, w, act, wc, s_t, s_c, need_gx, need_gw, has_bias):
        planes, rows, c = x.shape
        gx, sums = raw.rowdot_bwd(gy, y, act, x, c, w.detach(), need_gx, need_gw or has_bias, planes, rows, wc, s_t, s_c)
        gw = gb = None
        if need_gw or has_bias:
            gw = torch.empty_like(w)
            raw.emit_sums(sums, gw, c, wc=wc if wc > 0 else 0, s_t=s_t, s_c=s_c)
            gb = torch.empty(1, dtype=torch.float32, device=x.device)
            raw.emit_sums(sums[c:], gb, 1)
        ctx.meta = (act, wc, s_t, s_c)
        ctx.save_for_backward(gy, x, w)
        if gx is None:
            gx = torch.zeros(0, device=x.device)
        if gw is None:
            gw = torch.zeros(0, device=x.device)
            gb = torch.zeros(0, device=x.device)
        ctx.mark_non_differentiable(gb)
        return gx, gw, gb

    @staticmethod
    def backward(ctx, ggx, ggw, ggb):
        # gx[r,:] = gy[r] * w  (act == NONE):  d/dgy = ggx[r,:].w   d/dw = sum_r gy[r]*ggx[r,:]
        gy, x, w = ctx.saved_tensors
        act, wc, s_t, s_c = ctx.meta
        g_gy = g_w = None
        if ggx is not None and ggx.numel():
            ggx = ggx.contiguous()
            if ctx.needs_input_grad[0]:
                g_gy = _RowDot.apply(ggx, w, None, ACT_NONE, wc, s_t, s_c)
            if ctx.needs_input_grad[3]:
                ones = torch.ones_like(gy)
                _, g_w, _ = _RowDotBwd.apply(gy, ones, ggx, w, ACT_NONE, wc, s_t, s_c, False, True, False)
        return g_gy, None, None, g_w, None, None, None, None, None, None, None


def rowdot(x, w, bias, act=ACT_NONE, wc=0, s_t=0, s_c=1):
    return _RowDot.apply(x, w, bias, act, wc, s_t, s_c)


# ------------------------------------------------------------------------------------------------- boundary conversions
class _ToPlanes(Function):
    """fp32 [rows, C] -> plane tensor [P, rows, C8]"""

    @staticmethod
    def forward(ctx, x, c_dst):
        ctx.c_src = x.shape[1]
        return raw.f32_to_planes(x.contiguous(), _planes(), c_dst)

    @staticmethod
    def backward(ctx, g):
        return _FromPlanes.apply(g, ctx.c_src), None


class _FromPlanes(Function):
    """plane tensor [P, rows, C] -> fp32 [rows, c_take]"""

    @staticmethod
    def forward(ctx, x, c_take):
        ctx.c_src = x.shape[2]
        return raw.planes_to_f32(x.contiguous(), x.shape[2], c_take)

    @staticmethod
    def backward(ctx, g):
        return _ToPlanes.apply(g, ctx.c_src), None


def to_planes(x, c_dst=None):
    return _ToPlanes.apply(x, raw.round_up(x.shape[1], 8) if c_dst is None else c_dst)


def from_planes(x, c_take=None):
    return _FromPlanes.apply(x, x.shape[2] if c_take is None else c_take)


class _UnaryF32(Function):
    """tanh / sigmoid on small fp32 outputs (model/gan.py:22,56)."""

    @staticmethod
    def forward(ctx, x, act):
        y = raw.unary_f32(x.contiguous(), act)
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return raw.unary_bwd_f32(gy, y, ctx.act), None


def unary_f32(x, act):
    return _UnaryF32.apply(x, act)


class _ConvT1Act(Function):
    """ConvTranspose3d(C->1) + bias + tanh in one pass (model/gan.py:21-22): y = act(col2im(x.W) + b)."""

    @staticmethod
    def forward(ctx, op, x, w, bias, act):
        x = x.contiguous()
        y = op.fwd(x, w, bias.detach() if bias is not None else None, act)
        ctx.op, ctx.act, ctx.has_bias = op, act, bias is not None
        ctx.w_obj = w
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, _w_checked, y = ctx.saved_tensors
        w = ctx.w_obj
        op = ctx.op
        g = raw.unary_bwd_f32(gy.contiguous(), y, ctx.act) if ctx.act != ACT_NONE else gy.contiguous()
        gx = op.tr(g, w) if ctx.needs_input_grad[1] else None
        gw = op.wgrad(x, g, tuple(w.shape)) if ctx.needs_input_grad[2] else None
        gb = None
        if ctx.has_bias and ctx.needs_input_grad[3]:
            gb = raw.emit_sums(raw.sum_f32(g), torch.empty(1, dtype=torch.float32, device=g.device), 1)
        return None, gx, gw, gb, None


def convt1_act(op, x, w, bias, act):
    return _ConvT1Act.apply(op, x, w, bias, act)


# ------------------------------------------------------------------------------------------------- fade-in blend
class _Fade(Function):
    """progressive_gan.py:48-50: f*x + (1-f)*from_SDF(x_in[:, ::2, ::2, ::2]) ; linear in (x, x_in)."""

    @staticmethod
    def forward(ctx, x, vol, f):
        p, b, r, _, _, c = x.shape
        ctx.f = f
        ctx.vol_shape = vol.shape
        return raw.fade_fwd(x.contiguous(), b, r, c, vol.contiguous(), f)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        gx = _Scale.apply(g, ctx.f) if ctx.needs_input_grad[0] else None
        gvol = _FadeVolGrad.apply(g, ctx.f, ctx.vol_shape) if ctx.needs_input_grad[1] else None
        return gx, gvol, None


class _Scale(Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return raw.axpby_planes(x.contiguous(), alpha)

    @staticmethod
    def backward(ctx, g):
        return _Scale.apply(g, ctx.alpha), None


class _FadeVolGrad(Function):
    """gvol[b, 2d, 2h, 2w] = (1-f) * g[b,d,h,w,0], zero elsewhere"""

    @staticmethod
    def forward(ctx, g, f, vol_shape):
        p, b, r, _, _, c = g.shape
        gvol = torch.zeros(vol_shape, dtype=torch.float32, device=g.device)
        ctx.f, ctx.c, ctx.shape = f, c, g.shape
        return raw.fade_bwd_vol(g.contiguous(), b, r, c, f, gvol)

    @staticmethod
    def backward(ctx, gg):
        # adjoint: planes tensor with channel 0 = (1-f) * gg[::2], other channels zero  == fade_fwd(0, gg, f=0-scaled)
        p, b, r, _, _, c = ctx.shape
        zero = torch.zeros(ctx.shape, dtype=torch.bfloat16, device=gg.device)
        return raw.fade_fwd(zero, b, r, c, gg.contiguous(), ctx.f), None, None


def fade(x, vol, f):
    return _Fade.apply(x, vol, f)


class _Fanout2(Function):
    """A plane tensor consumed by two layers.  autograd would add the two incoming plane gradients plane-wise in
    bf16 (destroying the hi/lo split); this sums them value-wise and re-splits."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, g1, g2):
        if g1 is None:
            return g2
        if g2 is None:
            return g1
        return raw.axpby_planes(g1.contiguous(), 1.0, g2.contiguous(), 1.0)


def fanout2(x):
    return _Fanout2.apply(x)
