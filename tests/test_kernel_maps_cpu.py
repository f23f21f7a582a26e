"""Index maps of the round-2 kernels, restated in Python/numpy and checked against the definitions they implement (the GPU tests check the
numbers against the slow kernels bit for bit; these pin the maps that produce them, on CPU):

  * the staged epilogue of the implicit-GEMM kernels (csrc/sg_igemm.cu, `staged` path): row-owner writes and cooperative reads of the
    XOR-swizzled staging tile cover every chunk exactly once and are shared-memory-bank-conflict free per quarter warp;
  * the descriptor arithmetic of halo_issue_tap(): adding to the packed 14-bit start-address field equals rebuilding the descriptor;
  * the tiled second stage of ConvTranspose3d(C -> 1) (csrc/sg_elementwise.cu, sg_col2im_c1_tiled_kernel) against the gather definition
    (model/gan.py:21: output voxel o sums taps k with (o + 1 - k) even, input (o + 1 - k) / 2);
  * the tap-contiguous weight pack (csrc/sg_abi.cu, sg_pack_b_taps_kernel) against the generic pack's index formula for the four conv layouts.
"""
import itertools

import numpy as np


# ------------------------------------------------------------------------------------------------ staged epilogue
def _slot(row, chunk, pitch):
    """byte offset of 16-byte chunk `chunk` of staging row `row` (mirror of `srow + ((chunk ^ (lane & 7)) << 4)`)"""
    return row * pitch + ((chunk ^ (row & 7)) << 4)


def test_staged_epilogue_swizzle_is_a_bijection_and_conflict_free():
    for bn in (64, 128):
        pitch, cpr = bn * 2, bn // 8
        lgc = 4 if cpr == 16 else 3
        # row owner: lane l owns row l and writes its cpr chunks
        written = {_slot(l, c, pitch) for l in range(32) for c in range(cpr)}
        assert len(written) == 32 * cpr and max(written) + 16 <= 32 * pitch
        # cooperative pass: element i = lane + 32 k -> row i >> lgc, chunk i & (cpr - 1): reads every slot exactly once
        read = []
        for k in range(cpr):
            for lane in range(32):
                i = lane + 32 * k
                read.append(_slot(i >> lgc, i & (cpr - 1), pitch))
        assert sorted(read) == sorted(written)
        # 16-byte accesses are served a quarter warp (8 lanes) at a time: the 8 addresses must fall into 8 distinct 16-byte bank groups
        for c in range(cpr):                                                   # row-owner store of chunk c
            for q in range(4):
                groups = {(_slot(l, c, pitch) // 16) % 8 for l in range(8 * q, 8 * q + 8)}
                assert len(groups) == 8
        for k in range(cpr):                                                   # cooperative load
            for q in range(4):
                lanes = range(8 * q, 8 * q + 8)
                groups = {(_slot((l + 32 * k) >> lgc, (l + 32 * k) & (cpr - 1), pitch) // 16) % 8 for l in lanes}
                assert len(groups) == 8
        # whole rows per instruction: the 32 lanes of one cooperative step touch 32 / cpr complete rows
        for k in range(cpr):
            rows = {(lane + 32 * k) >> lgc for lane in range(32)}
            assert len(rows) == 32 // cpr


# ------------------------------------------------------------------------------------------------ UMMA descriptors
def _umma_desc(saddr, lbo, sbo):
    """mirror of umma_desc() in sg_common.cuh"""
    return ((saddr & 0x3FFFF) >> 4) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (2 << 61)


def test_issue_loop_descriptor_adds_equal_rebuilt_descriptors():
    a_hi = _umma_desc(0, 16, 9 * 128) >> 32
    b_hi = _umma_desc(0, 16, 1024) >> 32
    lbo = (16 >> 4) << 16
    sub_step = (2 * 8 * 9 * 128) >> 4
    rng = np.random.default_rng(0)
    for _ in range(200):
        blk = int(rng.integers(2, 180)) * 1024               # 1024-aligned halo block somewhere in 227 KB of shared memory
        b_base = int(rng.integers(2, 220)) * 1024
        for qz, qx, sub, kk in itertools.product((0, 1), (0, 1), (0, 1), range(4)):
            a_view = blk + (((sub * 2 + qz) * 8) * 9 + qx) * 128                # what the kernel used to rebuild per MMA
            ref_a = _umma_desc(a_view + kk * 32, 16, 9 * 128)
            ref_b = _umma_desc(b_base + kk * 32, 16, 1024)
            a_lo = ((((blk & 0x3FFFF) >> 4) + (qz * 72 + qx) * 8) | lbo) + sub * sub_step
            b_lo = ((b_base & 0x3FFFF) >> 4) | lbo
            assert (a_hi << 32) | (a_lo + 2 * kk) == ref_a
            assert (b_hi << 32) | (b_lo + 2 * kk) == ref_b


# ------------------------------------------------------------------------------------------------ tiled col2im
def _col2im_direct(P, n, d, h, w):
    """definition: ConvTranspose3d(C -> 1, k4, s2, p1) second stage; P[v, tap] with tap = kd*16 + kh*4 + kw"""
    out = np.zeros((n, 2 * d, 2 * h, 2 * w), dtype=np.float64)
    Pv = P.reshape(n, d, h, w, 4, 4, 4)
    for kd, kh, kw in itertools.product(range(4), repeat=3):
        for idd, ih, iw in itertools.product(range(d), range(h), range(w)):
            od, oh, ow = 2 * idd - 1 + kd, 2 * ih - 1 + kh, 2 * iw - 1 + kw
            if 0 <= od < 2 * d and 0 <= oh < 2 * h and 0 <= ow < 2 * w:
                out[:, od, oh, ow] += Pv[:, idd, ih, iw, kd, kh, kw]
    return out


def _col2im_tiled(P, n, d, h, w):
    """mirror of sg_col2im_c1_tiled_kernel: 16 x 8 x 8 output tiles fed from a (8+2) x (4+2) x (4+2) block of input rows"""
    OD, OH, OW = 2 * d, 2 * h, 2 * w
    out = np.zeros((n, OD, OH, OW), dtype=np.float64)
    Pv = P.reshape(n, d, h, w, 64)
    for nn, tz, ty, tx in itertools.product(range(n), range(OD // 8), range(OH // 8), range(OW // 16)):
        ix0, iy0, iz0 = tx * 8 - 1, ty * 4 - 1, tz * 4 - 1
        rows = np.zeros((6, 6, 10, 64))
        for rz, ry, rx in itertools.product(range(6), range(6), range(10)):
            ix, iy, iz = ix0 + rx, iy0 + ry, iz0 + rz
            if 0 <= ix < w and 0 <= iy < h and 0 <= iz < d:
                rows[rz, ry, rx] = Pv[nn, iz, iy, ix]
        for lz, ly, lx in itertools.product(range(8), range(8), range(16)):
            pd, ph, pw = lz & 1, ly & 1, lx & 1
            s = 0.0
            for td, th, tw in itertools.product((0, 1), repeat=3):
                rz = (lz >> 1) + ((1 - td) if pd else -td) + 1
                ry = (ly >> 1) + ((1 - th) if ph else -th) + 1
                rx = (lx >> 1) + ((1 - tw) if pw else -tw) + 1
                kd = 2 * td if pd else 1 + 2 * td
                kh = 2 * th if ph else 1 + 2 * th
                kw = 2 * tw if pw else 1 + 2 * tw
                s += rows[rz, ry, rx, kd * 16 + kh * 4 + kw]
            out[nn, tz * 8 + lz, ty * 8 + ly, tx * 16 + lx] = s
    return out


def test_tiled_col2im_matches_the_transposed_convolution_definition():
    rng = np.random.default_rng(1)
    n, d, h, w = 1, 4, 8, 8
    P = rng.standard_normal((n * d * h * w, 64))
    assert np.allclose(_col2im_tiled(P, n, d, h, w), _col2im_direct(P, n, d, h, w), rtol=0, atol=1e-12)


# ------------------------------------------------------------------------------------------------ weight pack
def _pack_generic(wflat, n_pad, n_valid, k_pad, taps, c_count, c_valid, s_n0, s_tap, s_c, classes):
    """mirror of sg_pack_b_kernel (planes = 1): image[class][kc][n][64] as float for comparison, -1 = never written"""
    kch = k_pad // 64
    img = np.zeros((classes, kch, n_pad, 64))
    for cls, kc, n, j, e in itertools.product(range(classes), range(kch), range(n_pad), range(8), range(8)):
        k = kc * 64 + j * 8 + e
        tap, c = divmod(k, c_count)
        x = 0.0
        if n < n_valid and tap < taps and c < c_valid:
            st = tap
            if classes == 8:
                td, th, tw = (tap >> 2) & 1, (tap >> 1) & 1, tap & 1
                kd = 2 * td if (cls >> 2) & 1 else 1 + 2 * td
                kh = 2 * th if (cls >> 1) & 1 else 1 + 2 * th
                kw = 2 * tw if cls & 1 else 1 + 2 * tw
                st = kd * 16 + kh * 4 + kw
            x = wflat[n * s_n0 + st * s_tap + c * s_c]
        img[cls, kc, n, ((j ^ (n & 7)) << 3) + e] = x            # 16-byte piece j lands at chunk j ^ (n & 7)
    return img


def _pack_taps(wflat, n_pad, n_valid, k_pad, taps, c_count, c_valid, s_n0, s_c, classes):
    """mirror of sg_pack_b_taps_kernel: one block per (n, 64-channel chunk), 512 pieces cut out of a [64 c][64 taps] tile"""
    kch, cch = k_pad // 64, c_count // 64
    img = np.zeros((classes, kch, n_pad, 64))
    for n, cb in itertools.product(range(n_pad), range(cch)):
        tile = np.zeros((64, 64))
        for cl in range(64):
            c = cb * 64 + cl
            if n < n_valid and c < c_valid:
                tile[cl] = wflat[n * s_n0 + c * s_c: n * s_n0 + c * s_c + 64]
        for pc in range(512):
            j = pc & 7
            cls, tap, st = 0, pc >> 3, pc >> 3
            if classes == 8:
                cls, tap = pc >> 6, (pc >> 3) & 7
                td, th, tw = (tap >> 2) & 1, (tap >> 1) & 1, tap & 1
                kd = 2 * td if (cls >> 2) & 1 else 1 + 2 * td
                kh = 2 * th if (cls >> 1) & 1 else 1 + 2 * th
                kw = 2 * tw if cls & 1 else 1 + 2 * tw
                st = kd * 16 + kh * 4 + kw
            kc = tap * cch + cb
            for e in range(8):
                img[cls, kc, n, ((j ^ (n & 7)) << 3) + e] = tile[j * 8 + e, st]
    return img


def test_tap_contiguous_pack_equals_the_generic_pack():
    rng = np.random.default_rng(2)
    cout, cin = 24, 64                                   # n_pad = 32 > n_valid for the layouts whose N is cout
    layouts = [
        # (weight shape, n_valid, k_pad, taps, c_count, s_n0, s_c, classes)          mirrors raw.pack_conv_fwd / _dgrad / convt_fwd / _dgrad
        ((cout, cin, 64), cout, 64 * cin, 64, cin, cin * 64, 64, 1),
        ((cin, 16, 64), 16, 8 * cin, 8, cin, 64, 16 * 64, 8),
        ((cin, cout, 64), cout, 8 * cin, 8, cin, 64, cout * 64, 8),
        ((64, 16, 64), 64, 64 * 64, 64, 64, 16 * 64, 64, 1) if False else ((16, 64, 64), 16, 64 * 64, 64, 64, 64 * 64, 64, 1),
    ]
    for shape, n_valid, k_pad, taps, c_count, s_n0, s_c, classes in layouts:
        wflat = rng.standard_normal(int(np.prod(shape)))
        n_pad = (n_valid + 15) // 16 * 16
        a = _pack_generic(wflat, n_pad, n_valid, k_pad, taps, c_count, c_count, s_n0, 1, s_c, classes)
        b = _pack_taps(wflat, n_pad, n_valid, k_pad, taps, c_count, c_count, s_n0, s_c, classes)
        assert np.array_equal(a, b), (shape, classes)
