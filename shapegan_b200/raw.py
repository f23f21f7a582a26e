"""Thin torch-tensor front end over the C ABI (shapegan_b200/_lib.py).  torch is used here only for device
memory (caching allocator) and the current stream; every FLOP happens inside libsg_b200."""
import ctypes

import torch

from . import _lib as L


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


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
           n_pad=None):
    _require_cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous()
    n_pad = round_up(n_valid, 16) if n_pad is None else n_pad
    a = L.SgPackBArgs(ctypes.c_void_p(w.data_ptr()), None, planes, classes, n_pad, n_valid,
                      n0_count if n0_count is not None else n_pad, s_n1, s_n0, k_pad, taps, c_count, c_valid, s_tap, s_c)
    nbytes = L.lib().sg_pack_b_bytes(ctypes.byref(a))
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
def igemm(mode, planes, a, a_dims, rows, k, b_img, n_valid, out, out_ld, out_kind=L.OUT_BF16, bias=None, act=L.ACT_NONE,
          a2=None, a2_c=0, mask=None, mask_act=L.ACT_NONE, out_dims=(0, 0, 0), bn=0, mt=0, ksplit=0, n_pad=None):
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
    if mask is not None:
        args.mask = ctypes.c_void_p(mask.data_ptr())
        args.mask_plane_stride = mask.stride(0) if mask.shape[0] == 2 else 0
        args.mask_act = mask_act
    args.out = ctypes.c_void_p(out.data_ptr())
    args.out_plane_stride = out.stride(0) if (out.dtype == torch.bfloat16 and out.shape[0] == 2) else 0
    args.out_kind, args.out_ld = out_kind, out_ld
    args.out_d, args.out_h, args.out_w = out_dims
    L.check(L.lib().sg_igemm(ctypes.byref(args), stream()), 'sg_igemm')
    return out


def wgrad(b_mode, planes, a, a_c, b, b_dims, rows, grad, sm, st, sc, m_valid, cb=None, taps=None, accumulate=False,
          scale=1.0, merge_n=1, ksplit=0):
    """grad[m*sm + tap*st + c*sc] (+)= sum_rows A[row,m] * gather(B)[row(+tap), c]."""
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
    L.check(L.lib().sg_wgrad(ctypes.byref(args), stream()), 'sg_wgrad')
    if b_mode == L.MODE_PATCH:
        taps_r, cb_r = 1, 64
    else:
        taps_r = 64 if b_mode == L.MODE_CONV else 1
        cb_r = c
    if taps is not None:
        taps_r, cb_r = taps, cb
    r = L.SgWgradReduceArgs(ctypes.c_void_p(ws.data_ptr()), args.ksplit_out, round_up(a_c, 128), m_valid, taps_r, cb_r,
                            sm, st, sc, ctypes.c_void_p(grad.data_ptr()), 1 if accumulate else 0, scale)
    L.check(L.lib().sg_wgrad_reduce(ctypes.byref(r), stream()), 'sg_wgrad_reduce')
    return grad


def device_error_word():
    p = ctypes.c_void_p()
    L.check(L.lib().sg_device_error_word(ctypes.byref(p)), 'sg_device_error_word')
    return p.value
