"""autograd Functions of the point-set GAN path (model/point_sdf_net.py) over csrc/sg_pointnet.cu: LayerNorm(+ReLU), the per-shape
vector add of SDFGenerator, and PointNet's max pooling with the scatter / gather pair that makes it twice differentiable (the
WGAN-GP of train_point_gan.py:61-71 differentiates through the critic's input gradient)."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L
from . import raw
from .raw import _call, _p, _ps


class _LayerNormAct(Function):
    """y = act(LayerNorm_C(x) * gamma + beta) over a plane tensor [P, rows, C]   (point_sdf_net.py:64, :110-112)"""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, act):
        x = x.contiguous()
        p, rows, c = x.shape
        y = torch.empty_like(x)
        stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
        _call('sg_ln_act_fwd', _p(x), _ps(x), _p(y), _ps(y), p, rows, c, _p(gamma.detach()), _p(beta.detach()), float(eps), act, _p(stats))
        ctx.act = act
        ctx.save_for_backward(x, y, gamma, stats)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, y, gamma, stats = ctx.saved_tensors
        p, rows, c = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        gbeta = torch.empty(c, dtype=torch.float32, device=x.device)
        ggamma = torch.empty(c, dtype=torch.float32, device=x.device)
        ws = torch.empty(296 * 2 * c, dtype=torch.float32, device=x.device)
        _call('sg_ln_act_bwd', _p(gy), _ps(gy), _p(y), _ps(y), _p(x), _ps(x), _p(gx), _ps(gx), p, rows, c, _p(gamma.detach()), ctx.act, _p(stats),
              _p(gbeta), _p(ggamma), _p(ws), ws.numel())
        return gx, ggamma, gbeta, None, None


def layernorm_act(x, ln, act):
    """`ln` is the nn.LayerNorm parameter container"""
    return _LayerNormAct.apply(x, ln.weight, ln.bias, ln.eps, act)


class _RowsAddVec(Function):
    """y[row, :] = x[row, :] + v[row // seg_len, :]  with v fp32 [segments, C]   (point_sdf_net.py:105-109: z_lin(z).unsqueeze(1) + x)"""

    @staticmethod
    def forward(ctx, x, v, seg_len):
        x = x.contiguous()
        p, rows, c = x.shape
        y = torch.empty_like(x)
        _call('sg_rows_add_vec', _p(x), _ps(x), _p(v.contiguous()), _p(y), _ps(y), p, rows, c, seg_len)
        ctx.meta = (seg_len, v.shape[0])
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        seg_len, segs = ctx.meta
        g = g.contiguous()
        p, rows, c = g.shape
        return g, raw.segment_colsum(g, segs, seg_len), None


def rows_add_vec(x, v, seg_len):
    return _RowsAddVec.apply(x, v, seg_len)


def _segmax_move(big, small, arg, seg_len, gather):
    p, segs, c = small.shape
    _call('sg_segmax_move', _p(big), _ps(big), _p(small), _ps(small), p, segs, c, seg_len, _p(arg), 1 if gather else 0)


class _SegMax(Function):
    """out[s, :] = max over the seg_len rows of segment s (x.max(dim=-2)[0], point_sdf_net.py:40-41).  Piecewise linear: its backward
    scatters to the arg-max rows, the backward of that gathers them again -- `_SegScatter` / `_SegGather` are each other's derivative."""

    @staticmethod
    def forward(ctx, x, seg_len):
        x = x.contiguous()
        p, rows, c = x.shape
        segs = rows // seg_len
        out = torch.empty((p, segs, c), dtype=torch.bfloat16, device=x.device)
        arg = torch.empty((segs, c), dtype=torch.int32, device=x.device)
        _call('sg_segmax_fwd', _p(x), _ps(x), p, segs, c, seg_len, _p(out), _ps(out), _p(arg))
        ctx.seg_len, ctx.rows = seg_len, rows
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g, _garg):
        (arg,) = ctx.saved_tensors
        return _SegScatter.apply(g, arg, ctx.seg_len, ctx.rows), None


class _SegScatter(Function):
    @staticmethod
    def forward(ctx, small, arg, seg_len, rows):
        small = small.contiguous()
        p, segs, c = small.shape
        big = torch.zeros((p, rows, c), dtype=torch.bfloat16, device=small.device)
        _segmax_move(big, small, arg, seg_len, gather=False)
        ctx.seg_len = seg_len
        ctx.save_for_backward(arg)
        return big

    @staticmethod
    def backward(ctx, gbig):
        (arg,) = ctx.saved_tensors
        return _SegGather.apply(gbig, arg, ctx.seg_len), None, None, None


class _SegGather(Function):
    @staticmethod
    def forward(ctx, big, arg, seg_len):
        big = big.contiguous()
        p, rows, c = big.shape
        segs = arg.shape[0]
        small = torch.empty((p, segs, c), dtype=torch.bfloat16, device=big.device)
        _segmax_move(big, small, arg, seg_len, gather=True)
        ctx.meta = (seg_len, rows)
        ctx.save_for_backward(arg)
        return small

    @staticmethod
    def backward(ctx, gsmall):
        (arg,) = ctx.saved_tensors
        seg_len, rows = ctx.meta
        return _SegScatter.apply(gsmall, arg, seg_len, rows), None, None


def segment_max(x, seg_len):
    return _SegMax.apply(x, seg_len)[0]
