"""ctypes binding of libsg_b200.so (include/sg_b200.h).  No torch types cross this boundary: only raw device
pointers, sizes and a cudaStream_t.  The product path fails loudly when the CUDA library is missing."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_longlong, c_size_t, c_void_p

from . import build as _build

MODE_DENSE, MODE_CONV, MODE_CONVT, MODE_PATCH = 0, 1, 2, 3
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
OUT_BF16, OUT_F32, OUT_F32_ATOMIC = 0, 1, 2


class SgTensor(ctypes.Structure):
    _fields_ = [('ptr', c_void_p), ('plane_stride', c_int64),
                ('n', c_int32), ('d', c_int32), ('h', c_int32), ('w', c_int32), ('c', c_int32)]


class SgIgemmArgs(ctypes.Structure):
    _fields_ = [('mode', c_int32), ('planes', c_int32), ('a', SgTensor), ('a2', SgTensor), ('rows', c_int64),
                ('k', c_int32), ('n_pad', c_int32), ('n_valid', c_int32),
                ('bn', c_int32), ('mt', c_int32), ('ksplit', c_int32),
                ('b_packed', c_void_p), ('bias', c_void_p), ('act', c_int32), ('bias_mod', c_int32),
                ('mask', c_void_p), ('mask_plane_stride', c_int64), ('mask_act', c_int32),
                ('out', c_void_p), ('out_plane_stride', c_int64), ('out_kind', c_int32), ('out_ld', c_int32),
                ('out_d', c_int32), ('out_h', c_int32), ('out_w', c_int32),
                ('splitk_ws', c_void_p), ('splitk_ws_bytes', c_int64)]


class SgWgradArgs(ctypes.Structure):
    _fields_ = [('b_mode', c_int32), ('planes', c_int32), ('a', SgTensor), ('b', SgTensor), ('rows', c_int64),
                ('ksplit', c_int32), ('merge_n', c_int32), ('partials', c_void_p), ('ksplit_out', c_int32),
                ('bias_partials', c_void_p), ('bias_ws_floats', c_int32)]


class SgWgradReduceArgs(ctypes.Structure):
    _fields_ = [('partials', c_void_p), ('ksplit', c_int32), ('m_pad', c_int32), ('m_valid', c_int32),
                ('taps', c_int32), ('cb', c_int32), ('sm', c_int64), ('st', c_int64), ('sc', c_int64),
                ('grad', c_void_p), ('accumulate', c_int32), ('scale', c_float), ('c_valid', c_int32)]


class SgPackBArgs(ctypes.Structure):
    _fields_ = [('w', c_void_p), ('image', c_void_p), ('planes', c_int32), ('classes', c_int32),
                ('n_pad', c_int32), ('n_valid', c_int32), ('n0_count', c_int32), ('s_n1', c_int64), ('s_n0', c_int64),
                ('k_pad', c_int32), ('taps', c_int32), ('c_count', c_int32), ('c_valid', c_int32),
                ('s_tap', c_int64), ('s_c', c_int64)]


class SgSdfnetFwdArgs(ctypes.Structure):
    _fields_ = [('points', c_void_p), ('latent', c_void_p), ('index', c_void_p), ('n', c_int64), ('w_img', c_void_p),
                ('aux', c_void_p), ('out', c_void_p), ('stash', c_void_p), ('mask_stash', c_void_p)]


class SgSdfnetBwdArgs(ctypes.Structure):
    _fields_ = [('gout', c_void_p), ('out', c_void_p), ('mask_stash', c_void_p), ('wt_img', c_void_p), ('w8', c_void_p),
                ('n', c_int64), ('gstash', c_void_p), ('gpoints', c_void_p), ('xyz_w', c_void_p)]


class SgDpStepArgs(ctypes.Structure):
    _fields_ = [('peer_grad', ctypes.POINTER(c_void_p)), ('peer_param', ctypes.POINTER(c_void_p)), ('peer_pad', ctypes.POINTER(c_void_p)),
                ('rank', c_int32), ('world', c_int32), ('n', c_int64), ('chunk', c_int64), ('s1', c_void_p), ('s2', c_void_p),
                ('kind', c_int32), ('lr', c_float), ('beta1', c_float), ('beta2', c_float), ('eps', c_float), ('clip', c_float),
                ('grad_scale', c_float), ('step', c_int32), ('sync', c_void_p)]


class SgMcArgs(ctypes.Structure):
    _fields_ = [('volume', c_void_p), ('nx', c_int32), ('ny', c_int32), ('nz', c_int32), ('level', c_float), ('spacing', c_float * 3),
                ('block_sums', c_void_p), ('vbase', c_void_p), ('vertices', c_void_p), ('normals', c_void_p), ('faces', c_void_p)]


class SgSdfnetInferArgs(ctypes.Structure):
    _fields_ = [('points', c_void_p), ('n', c_int64), ('n_ptr', c_void_p), ('ray_index', c_void_p), ('grid_r', c_int32),
                ('grid_axis', c_void_p), ('w_img', c_void_p), ('aux', c_void_p), ('out', c_void_p), ('mask_stash', c_void_p),
                ('trace_points', c_void_p), ('trace_dirs', c_void_p), ('trace_hit', c_void_p), ('next_index', c_void_p),
                ('next_count', c_void_p), ('sdf_offset', c_float), ('trace_clamp', c_float), ('trace_threshold', c_float),
                ('trace_radius', c_float), ('trace_miss_y', c_int32)]


# every symbol include/sg_b200.h declares, with (restype, argtypes); tests check that all of them resolve
SYMBOLS = {
    'sg_abi_version': (c_int32, []),
    'sg_last_error': (c_char_p, []),
    'sg_device_error_word': (c_int32, [ctypes.POINTER(c_void_p)]),
    'sg_check_device_error': (c_int32, []),
    'sg_num_sms': (c_int32, []),
    'sg_launch_count': (c_longlong, []),
    'sg_stream_capture_id': (c_int32, [c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]),
    'sg_debug_igemm_trace': (c_int32, [ctypes.POINTER(c_longlong), c_int32]),
    'sg_igemm': (c_int32, [ctypes.POINTER(SgIgemmArgs), c_void_p]),
    'sg_igemm_plan': (c_int32, [ctypes.POINTER(SgIgemmArgs), ctypes.POINTER(c_size_t)]),
    'sg_wgrad_plan': (c_int32, [ctypes.POINTER(SgWgradArgs), ctypes.POINTER(c_size_t)]),
    'sg_wgrad': (c_int32, [ctypes.POINTER(SgWgradArgs), c_void_p]),
    'sg_wgrad_reduce': (c_int32, [ctypes.POINTER(SgWgradReduceArgs), c_void_p]),
    'sg_pack_b_bytes': (c_size_t, [ctypes.POINTER(SgPackBArgs)]),
    'sg_pack_b': (c_int32, [ctypes.POINTER(SgPackBArgs), c_void_p]),
    'sg_sdfnet_fwd': (c_int32, [ctypes.POINTER(SgSdfnetFwdArgs), c_void_p]),
    'sg_sdfnet_fwd_layout': (c_int32, [ctypes.POINTER(c_int32), ctypes.POINTER(c_int32), ctypes.POINTER(c_int32)]),
    'sg_sdfnet_bwd': (c_int32, [ctypes.POINTER(SgSdfnetBwdArgs), c_void_p]),
    'sg_sdfnet_infer': (c_int32, [ctypes.POINTER(SgSdfnetInferArgs), c_void_p]),
    'sg_dp_step': (c_int32, [ctypes.POINTER(SgDpStepArgs), c_void_p]),
    'sg_mc_workspace_entries': (c_size_t, [c_int32, c_int32, c_int32]),
    'sg_mc_count': (c_int32, [ctypes.POINTER(SgMcArgs), c_void_p]),
    'sg_mc_emit': (c_int32, [ctypes.POINTER(SgMcArgs), c_void_p]),
    # p = pointer, l = int64, i = int32, f = float (see _sig)
    'sg_act_bwd': 'plplpliliipp',
    'sg_bn_stats': 'plilipp',
    'sg_bn_finalize': 'pliffppppp',
    'sg_bn_apply': 'plplilippppip',
    'sg_bn_bwd_reduce': 'plplpliliipppp',
    'sg_bn_bwd_apply': 'plplplpliliippppp',
    'sg_emit_sums': 'ppiifillp',
    'sg_col2im_c1': 'pliiiiipipp',
    'sg_unary_f32': 'pplip',
    'sg_unary_bwd_f32': 'ppplip',
    'sg_rowdot_fwd': 'plilipillpipp',
    'sg_rowdot_bwd': 'ppiplilipillplpip',
    'sg_to_planes': 'pllipliip',
    'sg_from_planes': 'pliliiplifp',
    'sg_sdf_pack_input': 'pppilpliip',
    'sg_sdf_unpack_grad': 'plpliliipppp',
    'sg_fade_fwd': 'plpliiiipfp',
    'sg_fade_bwd_vol': 'pliiiifpp',
    'sg_axpby_planes': 'plplplilffp',
    'sg_rmsprop': 'ppplfffffp',
    'sg_adam': 'pppplffffifp',
    'sg_clamp': 'plffp',
    'sg_sum_f32': 'plpp',
    'sg_l1_loss_grad': 'ppplpp',
    'sg_ln_act_fwd': 'plplilippfipp',
    'sg_ln_act_bwd': 'plplplplilipipppplp',
    'sg_rows_add_vec': 'plpplililp',
    'sg_segment_colsum': 'pliiilppp',
    'sg_segment_colsum_workspace': (c_size_t, [c_int32, c_int32, c_int64]),
    'sg_segmax_fwd': 'pliiilplpp',
    'sg_segmax_move': 'plpliiilpip',
    'sg_gp_interp': 'ppppilp',
    'sg_gp_seed': 'ppilfpp',
    'sg_critic_loss': 'pippp',
    'sg_grid_sphere_index': 'ipfppp',
    'sg_voxel_ingest': 'pplfip',
}

_SIG = {'p': c_void_p, 'l': c_int64, 'i': c_int32, 'f': c_float}
for _k, _v in list(SYMBOLS.items()):
    if isinstance(_v, str):
        SYMBOLS[_k] = (c_int32, [_SIG[ch] for ch in _v])

_lib = None


def lib():
    """Load (building if stale and nvcc is available) the in-tree shared library.  Raises if it cannot be loaded:
    there is no CPU or eager-PyTorch fallback for the hot path."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or (os.path.exists(_build.NVCC) and os.environ.get('SG_B200_NO_REBUILD') != '1'):
        if os.path.exists(_build.NVCC):
            path = _build.build_lib()
    if not os.path.exists(path):
        raise RuntimeError('libsg_b200.so is missing (%s) and nvcc is unavailable: the shapegan_b200 hot path has no '
                           'fallback. Run `python -m shapegan_b200.build`.' % path)
    h = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(h, name)           # AttributeError here == ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if h.sg_abi_version() != 1:
        raise RuntimeError('libsg_b200 ABI version mismatch')
    _lib = h
    return h


def check(rc, what=''):
    if rc != 0:
        msg = lib().sg_last_error()
        raise RuntimeError('libsg_b200 %s failed (%d): %s' % (what, rc, msg.decode() if msg else '?'))
