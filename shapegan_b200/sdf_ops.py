"""DeepSDF MLP (model/sdf_net.py:23-61) as one autograd Function over libsg_b200.

forward : cat(points, latent) -> 4x(Linear+ReLU) -> cat(x, input) -> 3x(Linear+ReLU) -> Linear -> tanh
Layer-by-layer variant: every Linear is one tcgen05 implicit-GEMM launch with the bias+ReLU epilogue fused; the
skip concat (sdf_net.py:59) is never materialised (two-source K loop); the 256->1 head + tanh is a row-dot kernel.
The backward is hand scheduled (no autograd graph between layers): act-bwd+bias-sum, wgrad, dgrad per layer.
`fused=True` swaps the forward for the persistent fused kernel (sg_sdfnet.cu) when available."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L
from . import config, raw
from .ops import PACK_CACHE, r64

HID = 256          # SDF_NET_BREADTH, model/sdf_net.py:21


def _pack_lin(w, planes, c_valid, k_pad):
    return raw.pack_b(w, planes, w.shape[0], k_pad, 1, k_pad, c_valid, s_n0=w.stride(0), s_tap=0, s_c=1)


def _pack_lin_t(w, planes, n_valid, n_pad, col0=0):
    """B[n = in-feature (col0 + n), k = out-feature] = W[k, col0 + n]"""
    out_f = w.shape[0]
    view = w[:, col0:]
    return raw.pack_b(view, planes, n_valid, r64(out_f), 1, r64(out_f), out_f, s_n0=1, s_tap=0, s_c=w.stride(0), n_pad=n_pad)


def fused_enabled():
    import os
    return os.environ.get('SG_B200_NO_FUSED_SDF') != '1'


def _fused_pack(w, b):
    """28 weight chunks (stream order of sg_sdfnet.cu) + fp32 aux block; cached on the parameter versions."""
    key = (PACK_CACHE._uid(w[0]), 'sdf_fused_fwd', 1)
    sig = tuple((t._version, t.data_ptr()) for t in list(w) + list(b))
    hit = PACK_CACHE.lookup(key, sig)
    if hit is not None:
        return hit
    dev = w[0].device
    chunk = 32768
    img = torch.empty(28 * chunk, dtype=torch.uint8, device=dev)
    off = 0

    def put(view, k):
        nonlocal off
        nbytes = (k // 64) * chunk
        raw.pack_b(view, 1, 256, k, 1, k, k, s_n0=view.stride(0), s_tap=0, s_c=1, out=img[off:off + nbytes])
        off += nbytes
    w1, w5 = w[0].detach(), w[4].detach()
    put(w1[:, 3:131], 128)
    for i in (1, 2, 3):
        put(w[i].detach(), 256)
    put(w5[:, 0:256], 256)
    put(w5[:, 259:387], 128)
    for i in (5, 6):
        put(w[i].detach(), 256)
    assert off == 28 * chunk
    aux = _aux_block(w, b, b[0].detach(), b[4].detach())
    PACK_CACHE.put(key, sig, (img, aux))
    return img, aux


def _aux_block(w, b, bias1, bias5):
    """fp32 aux block of sg_sdfnet.cu; `bias1` / `bias5` are the bias vectors layers1.0 / layers2.0 start from."""
    w1, w5 = w[0].detach(), w[4].detach()
    # per column pair (c, c+1): {wx_c, wx_c1, wy_c, wy_c1, wz_c, wz_c1, b_c, b_c1} -- the fma.rn.f32x2 operand order of sg_sdfnet.cu
    xb1 = torch.cat((w1[:, 0:3], bias1.unsqueeze(1)), 1).reshape(128, 2, 4).permute(0, 2, 1)
    xb5 = torch.cat((w5[:, 256:259], bias5.unsqueeze(1)), 1).reshape(128, 2, 4).permute(0, 2, 1)
    return torch.cat([xb1.reshape(-1), xb5.reshape(-1)] + [b[i].detach() for i in (1, 2, 3, 5, 6)] +
                     [w[7].detach().reshape(-1), b[7].detach().reshape(-1)]).contiguous().float()


def folded_enabled(latent_size):
    """single-latent inference rides the fused kernel: bf16 mode, latent_code_size 128 (the fused kernel's weight image)"""
    return config.planes() == 1 and latent_size == 128 and fused_enabled()


def folded_weights(params, latent_code):
    """(weight image, aux) for single-latent inference: the latent part of layers1.0 / layers2.0 is the constant vector
    W[:, latent] z (model/sdf_net.py:57,59 with one z for every point), folded in fp32 into those layers' bias slots -- the kernel
    then streams no latent chunks and layers1.0 needs no MMA at all."""
    w = [params[2 * i] for i in range(8)]
    b = [params[2 * i + 1] for i in range(8)]
    img, _ = _fused_pack(w, b)
    z = latent_code.detach().reshape(-1).float()
    w1, w5 = w[0].detach(), w[4].detach()
    bias1 = b[0].detach() + (w1[:, 3:131] * z).sum(1)
    bias5 = b[4].detach() + (w5[:, 259:387] * z).sum(1)
    return img, _aux_block(w, b, bias1, bias5)


def infer_single_latent(params, latent_code, **kw):
    img, aux = folded_weights(params, latent_code)
    return raw.sdfnet_infer(img, aux, **kw)


def normals_single_latent(params, latent_code, points, normalize=True):
    """SDFNet.get_normals (model/sdf_net.py:118-128) without autograd: folded forward that keeps only the 1-bit ReLU masks, then the
    fused backward chain with the xyz gradient accumulated inside its drains.  Returns (sdf [n], d sdf / d xyz [n, 3])."""
    w = [params[2 * i] for i in range(8)]
    n = points.shape[0]
    dev = points.device
    img, aux = folded_weights(params, latent_code)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    mstash = torch.empty((7, n, 8), dtype=torch.int32, device=dev)
    raw.sdfnet_infer(img, aux, n, out=out, points=points, mask_stash=mstash)
    xyz_w = torch.stack((w[0].detach()[:, 0:3].t(), w[4].detach()[:, 256:259].t())).contiguous().float()
    gout = torch.ones(n, dtype=torch.float32, device=dev)
    _, g = raw.sdfnet_bwd(gout, out, mstash, _fused_pack_t(w), w[7].detach().reshape(-1), xyz_w=xyz_w, want_gstash=False)
    if normalize:
        g /= torch.norm(g, dim=1).unsqueeze(dim=1)
    return out, g


class SDFNetFunction(Function):
    """args: points [N,3] fp32, latent fp32 ([N,L] rows, or a [S,L] table when `index` int32 [N] is given),
    then the 16 parameters in state_dict order (layers1.{0,2,4,6}.{weight,bias}, layers2.{0,2,4,6}.{weight,bias})."""

    @staticmethod
    def forward(ctx, points, latent, index, want_grad, seg_len, *params):
        planes = config.planes()
        n = points.shape[0]
        dev = points.device
        lat = latent.shape[1]
        cin = 3 + lat
        cin8 = r64(cin)            # input rows are physically zero-padded to a 64-multiple (K chunks, wgrad atoms)
        w = [params[2 * i] for i in range(8)]
        b = [params[2 * i + 1] for i in range(8)]
        points = points.contiguous()
        latent = latent.contiguous()
        need_graph = bool(want_grad)      # decided by the caller: grad mode is always off inside Function.forward
        if planes == 1 and lat == 128 and fused_enabled():
            # fused persistent kernel: all 8 layers per tile pair in one CTA (sg_sdfnet.cu)
            img, aux = _fused_pack(w, b)
            stash = torch.empty((7, n, HID), dtype=torch.bfloat16, device=dev) if need_graph else None
            mstash = torch.empty((7, n, 8), dtype=torch.int32, device=dev) if need_graph else None      # 1-bit ReLU masks
            out = raw.sdfnet_fwd(points, latent, index, img, aux, stash, mstash)
            if need_graph:
                # uniform segments (`seg_len` consecutive points per table row, the layout of BASELINE configs[2]): the backward needs the
                # xyz columns only (zero-padded to one 64-column atom); the latent columns are handled per SHAPE, not per point
                ctx.seg_len = int(seg_len) if (index is not None and seg_len and n == seg_len * latent.shape[0]) else 0
                if ctx.seg_len:
                    x_in = raw.f32_to_planes(points, planes, 64)
                else:
                    x_in = raw.sdf_pack_input(points, latent, index, lat, planes, cin8)
                ctx.meta = (planes, n, lat, cin, cin8, index is not None, latent.shape[0])
                ctx.w_objs = w
                ctx.fused = True
                ctx.mstash = mstash
                ctx.save_for_backward(x_in, out, index if index is not None else torch.empty(0, device=dev), stash, latent, *w)
            return out
        ctx.fused = False
        x_in = raw.sdf_pack_input(points, latent, index, lat, planes, cin8)
        hs = []
        h = x_in
        for i in range(4):                                   # layers1 (sdf_net.py:26-38)
            kin = cin8 if i == 0 else HID
            img = PACK_CACHE.get(w[i], 'sdf_f%d' % i, planes, lambda t, pl, i=i, kin=kin: _pack_lin(
                t, pl, cin if i == 0 else HID, r64(kin)))
            y = torch.empty((planes, n, HID), dtype=torch.bfloat16, device=dev)
            raw.igemm(L.MODE_DENSE, planes, h, (1, 1, 1, 1, kin), n, r64(kin), img, HID, y, HID, bias=b[i], act=L.ACT_RELU)
            hs.append(y)
            h = y
        # layers2.0 on cat(x, input) (sdf_net.py:59, :41): two-source K loop, K = 256 + roundup64(cin8)
        k5 = HID + r64(cin8)
        img = PACK_CACHE.get(w[4], 'sdf_f4', planes, lambda t, pl: _pack_lin(t, pl, HID + cin, k5))
        y = torch.empty((planes, n, HID), dtype=torch.bfloat16, device=dev)
        raw.igemm(L.MODE_DENSE, planes, h, (1, 1, 1, 1, HID), n, k5, img, HID, y, HID, bias=b[4], act=L.ACT_RELU,
                  a2=x_in, a2_c=cin8)
        hs.append(y)
        h = y
        for i in (5, 6):
            img = PACK_CACHE.get(w[i], 'sdf_f%d' % i, planes, lambda t, pl: _pack_lin(t, pl, HID, HID))
            y = torch.empty((planes, n, HID), dtype=torch.bfloat16, device=dev)
            raw.igemm(L.MODE_DENSE, planes, h, (1, 1, 1, 1, HID), n, HID, img, HID, y, HID, bias=b[i], act=L.ACT_RELU)
            hs.append(y)
            h = y
        out = raw.rowdot_fwd(h, HID, w[7], b[7], L.ACT_TANH)                  # sdf_net.py:50-51
        ctx.meta = (planes, n, lat, cin, cin8, index is not None, latent.shape[0])
        ctx.w_objs = w                       # parameter identities for the pack cache
        ctx.save_for_backward(x_in, out, index if index is not None else torch.empty(0, device=dev), *hs, *w)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        planes, n, lat, cin, cin8, indexed, lat_rows = ctx.meta
        saved = ctx.saved_tensors
        x_in, out, index = saved[0], saved[1], saved[2]
        hstash = saved[3] if ctx.fused else None                                   # the fused forward's [7, n, 256] stash
        latent_saved = saved[4] if ctx.fused else None
        hs = [hstash[i].unsqueeze(0) for i in range(7)] if ctx.fused else list(saved[3:10])
        w = ctx.w_objs                       # the saved parameter tensors were version-checked by autograd
        dev = gout.device
        need_points = ctx.needs_input_grad[0]
        need_latent = ctx.needs_input_grad[1]
        need_w = [ctx.needs_input_grad[5 + 2 * i] for i in range(8)]
        need_b = [ctx.needs_input_grad[6 + 2 * i] for i in range(8)]
        gw = [None] * 8
        gb = [None] * 8
        gout = gout.contiguous()

        def f32(shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        if ctx.fused and fused_bwd_enabled():
            return _backward_fused(planes, n, lat, cin, cin8, indexed, lat_rows, x_in, out, index, ctx.mstash, hs, w, gout,
                                   need_points, need_latent, need_w, need_b, seg_len=getattr(ctx, 'seg_len', 0), latent=latent_saved)

        # ---- head: Linear(256->1) + tanh
        gh, sums = raw.rowdot_bwd(gout, out, L.ACT_TANH, hs[6], HID, w[7], True, need_w[7] or need_b[7], planes, n)
        if need_w[7] or need_b[7]:
            gw[7] = raw.emit_sums(sums, f32(w[7].shape), HID)
            gb[7] = raw.emit_sums(sums[HID:], f32((1,)), 1)
        gx_a = gx_b = None
        for i in (6, 5, 4, 3, 2, 1, 0):
            g, sums = raw.act_bwd(gh, hs[i], L.ACT_RELU, HID, want_sums=need_b[i])
            if need_b[i]:
                gb[i] = raw.emit_sums(sums, f32((HID,)), HID)
            src = hs[i - 1] if i > 0 else x_in
            if i == 4:
                # W5 = [hidden 256 | xyz 3 | latent L]  (sdf_net.py:59)
                if need_w[4]:
                    gw[4] = f32(w[4].shape)
                    raw.wgrad(L.MODE_DENSE, planes, g, HID, hs[3], (1, 1, 1, 1, HID), n, gw[4], sm=HID + cin, st=0, sc=1, m_valid=HID)
                    _wgrad_input(planes, g, x_in, cin, cin8, n, gw[4][:, HID:], HID + cin)
                if need_points or need_latent:
                    img = PACK_CACHE.get(w[4], 'sdf_t4in', planes, lambda t, pl: _pack_lin_t(t, pl, cin, cin8, col0=HID))
                    gx_b = torch.empty((planes, n, cin8), dtype=torch.bfloat16, device=dev)
                    raw.igemm(L.MODE_DENSE, planes, g, (1, 1, 1, 1, HID), n, HID, img, cin8, gx_b, cin8, n_pad=cin8)   # rows >= cin of the image are zero
                img = PACK_CACHE.get(w[4], 'sdf_t4h', planes, lambda t, pl: _pack_lin_t(t, pl, HID, HID))
                gh = torch.empty((planes, n, HID), dtype=torch.bfloat16, device=dev)
                raw.igemm(L.MODE_DENSE, planes, g, (1, 1, 1, 1, HID), n, HID, img, HID, gh, HID)
            elif i == 0:
                if need_w[0]:
                    gw[0] = f32(w[0].shape)
                    _wgrad_input(planes, g, x_in, cin, cin8, n, gw[0], cin)
                if need_points or need_latent:
                    img = PACK_CACHE.get(w[0], 'sdf_t0', planes, lambda t, pl: _pack_lin_t(t, pl, cin, cin8))
                    gx_a = torch.empty((planes, n, cin8), dtype=torch.bfloat16, device=dev)
                    raw.igemm(L.MODE_DENSE, planes, g, (1, 1, 1, 1, HID), n, HID, img, cin8, gx_a, cin8, n_pad=cin8)   # rows >= cin of the image are zero
            else:
                if need_w[i]:
                    gw[i] = f32(w[i].shape)
                    raw.wgrad(L.MODE_DENSE, planes, g, HID, src, (1, 1, 1, 1, HID), n, gw[i], sm=HID, st=0, sc=1, m_valid=HID)
                img = PACK_CACHE.get(w[i], 'sdf_t%d' % i, planes, lambda t, pl: _pack_lin_t(t, pl, HID, HID))
                gh = torch.empty((planes, n, HID), dtype=torch.bfloat16, device=dev)
                raw.igemm(L.MODE_DENSE, planes, g, (1, 1, 1, 1, HID), n, HID, img, HID, gh, HID)
        gpoints = glatent = None
        if need_points or need_latent:
            gpoints = f32((n, 3)) if need_points else None
            if need_latent:
                glatent = torch.zeros((lat_rows, lat), dtype=torch.float32, device=dev) if indexed else f32((n, lat))
            raw.sdf_unpack_grad(gx_a, gx_b, cin8, lat, index if indexed else None, gpoints, glatent)
        grads = [gpoints, glatent, None, None, None]
        for i in range(8):
            grads.append(gw[i])
            grads.append(gb[i])
        return tuple(grads)


def fused_bwd_enabled():
    import os
    return os.environ.get('SG_B200_NO_FUSED_SDF_BWD') != '1'


def _fused_pack_t(w):
    """24 transposed weight chunks in the stream order of sg_sdfnet_bwd_kernel: layers2.4, 2.2, 2.0[:, :256], layers1.6, 1.4, 1.2."""
    key = (PACK_CACHE._uid(w[1]), 'sdf_fused_bwd', 1)
    order = (6, 5, 4, 3, 2, 1)
    sig = tuple((w[i]._version, w[i].data_ptr()) for i in order)
    hit = PACK_CACHE.lookup(key, sig)
    if hit is not None:
        return hit
    chunk = 32768
    img = torch.empty(24 * chunk, dtype=torch.uint8, device=w[1].device)
    for j, i in enumerate(order):
        t = w[i].detach()
        # B[n = in-feature][k = out-feature] = W[k, n]  (first 256 in-features: layers2.0 keeps its xyz/latent columns out)
        raw.pack_b(t, 1, HID, HID, 1, HID, HID, s_n0=1, s_tap=0, s_c=t.stride(0), n_pad=HID, out=img[j * 4 * chunk:(j + 1) * 4 * chunk])
    PACK_CACHE.put(key, sig, img)
    return img


def _backward_fused(planes, n, lat, cin, cin8, indexed, lat_rows, x_in, out, index, mstash, hs, w, gout, need_points, need_latent, need_w,
                    need_b, seg_len=0, latent=None):
    """Backward of the fused forward: ONE persistent kernel runs the whole input-gradient chain (sg_sdfnet_bwd) and leaves
    g_1..g_7 in a bf16 stash; the weight gradients, bias sums and the two input-gradient GEMMs read that stash."""
    dev = gout.device

    def f32(shape):
        return torch.empty(shape, dtype=torch.float32, device=dev)

    gw = [None] * 8
    gb = [None] * 8
    assert mstash.shape == (7, n, 8) and mstash.is_contiguous()
    if need_w[7] or need_b[7]:
        _, sums = raw.rowdot_bwd(gout, out, L.ACT_TANH, hs[6], HID, w[7], False, True, planes, n)
        gw[7] = raw.emit_sums(sums, f32(w[7].shape), HID)
        gb[7] = raw.emit_sums(sums[HID:], f32((1,)), 1)
    gst = raw.sdfnet_bwd(gout, out, mstash, _fused_pack_t(w), w[7].detach().reshape(-1))
    g = [gst[i].unsqueeze(0) for i in range(7)]                                     # g[i]: gradient w.r.t. the pre-activation of layer i+1
    for i in range(7):
        # the bias gradient (column sums of g_i) rides the weight-gradient GEMM of the same layer: one extra N = 16 MMA per K step against
        # a tile of ones instead of a second pass over the 512 B/point g stash
        bias_here = need_b[i] and need_w[i]
        if need_b[i]:
            gb[i] = f32((HID,))
            if not bias_here:
                _, sums = raw.act_bwd(g[i], None, L.ACT_NONE, HID, want_sums=True, want_g=False)
                raw.emit_sums(sums, gb[i], HID)
        if not need_w[i]:
            continue
        gw[i] = f32(w[i].shape)
        bg = gb[i] if bias_here else None
        if i == 0:
            if seg_len:
                raw.wgrad(L.MODE_DENSE, planes, g[0], HID, x_in, (1, 1, 1, 1, 64), n, gw[0], sm=cin, st=0, sc=1, m_valid=HID, c_valid=3, bias_grad=bg)
            else:
                _wgrad_input(planes, g[0], x_in, cin, cin8, n, gw[0], cin, bias_grad=bg)
        elif i == 4:       # W5 = [hidden 256 | xyz 3 | latent L]  (sdf_net.py:59)
            raw.wgrad(L.MODE_DENSE, planes, g[4], HID, hs[3], (1, 1, 1, 1, HID), n, gw[4], sm=HID + cin, st=0, sc=1, m_valid=HID, bias_grad=bg)
            if seg_len:
                raw.wgrad(L.MODE_DENSE, planes, g[4], HID, x_in, (1, 1, 1, 1, 64), n, gw[4][:, HID:], sm=HID + cin, st=0, sc=1, m_valid=HID, c_valid=3)
            else:
                _wgrad_input(planes, g[4], x_in, cin, cin8, n, gw[4][:, HID:], HID + cin)
        else:
            raw.wgrad(L.MODE_DENSE, planes, g[i], HID, hs[i - 1], (1, 1, 1, 1, HID), n, gw[i], sm=HID, st=0, sc=1, m_valid=HID, bias_grad=bg)
    gpoints = glatent = None
    if seg_len and not need_points:
        # every point of a segment shares its latent row: the latent columns of dW1 / dW5 and the latent-table gradient are products of
        # PER-SHAPE sums of g_1 / g_5 ([S, 256], one pass over two of the seven g planes) with [S, L] / [256, L] matrices -- no K = 131
        # GEMMs over the points, no [n, 131] input rows, no gradient scatter
        segs = lat_rows
        gsum = [raw.segment_colsum(g[li], segs, seg_len) for li in (0, 4)]
        z = latent.detach().float()
        w1, w5 = w[0].detach(), w[4].detach()
        if need_w[0]:
            gw[0][:, 3:3 + lat] = gsum[0].t() @ z                                          # [256, S] x [S, L]: per-shape, ~1e-5 of the step's FLOPs
        if need_w[4]:
            gw[4][:, HID + 3:HID + 3 + lat] = gsum[1].t() @ z
        if need_latent:
            glatent = gsum[0] @ w1[:, 3:3 + lat] + gsum[1] @ w5[:, HID + 3:HID + 3 + lat]
        grads = [None, glatent, None, None, None]
        for i in range(8):
            grads.append(gw[i])
            grads.append(gb[i])
        return tuple(grads)
    if need_points or need_latent:
        img = PACK_CACHE.get(w[0], 'sdf_t0', planes, lambda t, pl: _pack_lin_t(t, pl, cin, cin8))
        gx_a = torch.empty((planes, n, cin8), dtype=torch.bfloat16, device=dev)
        raw.igemm(L.MODE_DENSE, planes, g[0], (1, 1, 1, 1, HID), n, HID, img, cin8, gx_a, cin8, n_pad=cin8)
        img = PACK_CACHE.get(w[4], 'sdf_t4in', planes, lambda t, pl: _pack_lin_t(t, pl, cin, cin8, col0=HID))
        gx_b = torch.empty((planes, n, cin8), dtype=torch.bfloat16, device=dev)
        raw.igemm(L.MODE_DENSE, planes, g[4], (1, 1, 1, 1, HID), n, HID, img, cin8, gx_b, cin8, n_pad=cin8)
        gpoints = f32((n, 3)) if need_points else None
        if need_latent:
            glatent = torch.zeros((lat_rows, lat), dtype=torch.float32, device=dev) if indexed else f32((n, lat))
        raw.sdf_unpack_grad(gx_a, gx_b, cin8, lat, index if indexed else None, gpoints, glatent)
    grads = [gpoints, glatent, None, None, None]
    for i in range(8):
        grads.append(gw[i])
        grads.append(gb[i])
    return tuple(grads)


def _wgrad_input(planes, g, x_in, cin, cin8, n, grad_view, ld, bias_grad=None):
    """dW[:, input columns] = g^T x_in  where x_in has cin8 (multiple of 8) physical columns, cin valid.
    sg_wgrad needs a column count that is a multiple of 64; the reduce only emits the valid columns."""
    raw.wgrad(L.MODE_DENSE, planes, g, HID, x_in, (1, 1, 1, 1, cin8), n, grad_view, sm=ld, st=0, sc=1, m_valid=HID, c_valid=cin,
              bias_grad=bias_grad)


def sdfnet_apply(points, latent, index, params, seg_len=0):
    """seg_len > 0: the caller guarantees that table row s owns the points [s * seg_len, (s + 1) * seg_len) (index[i] == i // seg_len)"""
    want_grad = torch.is_grad_enabled() and (points.requires_grad or latent.requires_grad or any(p.requires_grad for p in params))
    if points.requires_grad:
        seg_len = 0          # the per-point xyz gradient takes the general path
    return SDFNetFunction.apply(points, latent, index, want_grad, seg_len, *params)
