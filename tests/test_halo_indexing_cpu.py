"""Index algebra of sg_igemm_halo_kernel (shapegan_b200/csrc/sg_igemm.cu), restated in Python and checked against the
definition of the two gathers it serves:

  Conv3d(k4,s2,p1) forward (model/gan.py:49-53): output voxel o reads input 2*o - 1 + k per axis, k in 0..3;
  ConvTranspose3d(k4,s2,p1) forward by output-parity class (model/gan.py:13-21): class bit p, tap bit t read input q + (p ? 1-t : -t).

The kernel loads one halo block per (group, 64-channel chunk) -- rows [(2 mt + 1) z'][8 y][9 x'] of the stride-2 (conv) or
unit-stride (convT) sample grid -- and views tap (q_z, q_x) of the group as rows starting at (q_z*8*9 + q_x), 8-row groups 9 rows apart.
This file pins: every tap is issued exactly once per channel chunk, and every operand row of every tap is the voxel the
convolution definition asks for.  (The GPU tests check the numbers; this checks the map that produces them.)"""
import itertools


def halo_tap(mode, grp, t4):
    """mirror of halo_tap() in sg_igemm.cu"""
    qz, qx = t4 >> 1, t4 & 1
    if mode == 'conv':
        kh, pz, px = grp >> 2, (grp >> 1) & 1, grp & 1
        return (2 * qz + pz) * 16 + kh * 4 + (2 * qx + px)
    th = grp
    return (1 - qz) * 4 + th * 2 + (1 - qx)


def block_origin(mode, grp, z0, cls):
    """first sample of the halo block, in input coordinates (x, y, z), and the sample stride; mirrors the TMA producer"""
    if mode == 'conv':
        kh, pz, px = grp >> 2, (grp >> 1) & 1, grp & 1
        return (-1 + px, -1 + kh, 2 * z0 - 1 + pz), 2
    pd, ph, pw = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
    th = grp
    return ((0 if pw else -1), (1 - th if ph else -th), z0 + (0 if pd else -1)), 1


def view_row(sub, qz, qx, r):
    """block row of operand row r (0..127) of M sub-tile `sub` for tap (qz, qx): start (sub*2+qz)*72 + qx, groups 9 rows apart"""
    return ((sub * 2 + qz) * 8) * 9 + qx + (r >> 3) * 9 + (r & 7)


def block_row_to_sample(row):
    z, rem = divmod(row, 72)
    y, x = divmod(rem, 9)
    return x, y, z


def test_every_tap_once_per_chunk():
    for mode, ngroups, ntaps in (('conv', 16, 64), ('convt', 2, 8)):
        taps = sorted(halo_tap(mode, g, t) for g in range(ngroups) for t in range(4))
        assert taps == list(range(ntaps))


def test_operand_rows_are_the_voxels_the_convolution_reads():
    for mt in (1, 2):
        for z0 in (0, 2 * mt):                                   # tile origin along z (multiples of the tile's z extent)
            # ---- Conv3d k4 s2 p1: tile rows enumerate output voxels (z0 + zz, y, x), x fastest
            for grp, t4 in itertools.product(range(16), range(4)):
                tap = halo_tap('conv', grp, t4)
                kd, kh, kw = tap >> 4, (tap >> 2) & 3, tap & 3
                (bx, by, bz), stride = block_origin('conv', grp, z0, 0)
                for sub, r in itertools.product(range(mt), range(128)):
                    oz, oy, ox = z0 + 2 * sub + (r >> 6), (r >> 3) & 7, r & 7
                    sx, sy, sz = block_row_to_sample(view_row(sub, t4 >> 1, t4 & 1, r))
                    assert sz <= 2 * mt and sy <= 7 and sx <= 8          # inside the block [(2 mt + 1)][8][9]
                    got = (bx + stride * sx, by + stride * sy, bz + stride * sz)
                    assert got == (2 * ox - 1 + kw, 2 * oy - 1 + kh, 2 * oz - 1 + kd)
            # ---- ConvTranspose3d k4 s2 p1, output-parity class cls: tile rows enumerate INPUT voxels (z0 + zz, y, x)
            for cls, grp, t4 in itertools.product(range(8), range(2), range(4)):
                tap = halo_tap('convt', grp, t4)
                td, th, tw = tap >> 2, (tap >> 1) & 1, tap & 1
                pd, ph, pw = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
                (bx, by, bz), stride = block_origin('convt', grp, z0, cls)
                for sub, r in itertools.product(range(mt), range(128)):
                    qz, qy, qx = z0 + 2 * sub + (r >> 6), (r >> 3) & 7, r & 7
                    sx, sy, sz = block_row_to_sample(view_row(sub, t4 >> 1, t4 & 1, r))
                    assert sz <= 2 * mt and sy <= 7 and sx <= 8
                    got = (bx + stride * sx, by + stride * sy, bz + stride * sz)
                    want = (qx + (1 - tw if pw else -tw), qy + (1 - th if ph else -th), qz + (1 - td if pd else -td))
                    assert got == want


def test_group_stride_is_uniform_across_the_tile():
    """what makes one UMMA descriptor (single SBO) describe a tap: consecutive 8-row groups are always 9 block rows apart,
    also across the z planes of the tile (8 y-rows x 9 = 72 = 8 groups x 9)"""
    for sub, qz, qx in itertools.product(range(2), range(2), range(2)):
        starts = [view_row(sub, qz, qx, 8 * g) for g in range(16)]
        assert all(b - a == 9 for a, b in zip(starts, starts[1:]))
        assert all(view_row(sub, qz, qx, 8 * g + i) == starts[g] + i for g in range(16) for i in range(8))
    # both M sub-tiles of an mt = 2 tile are one run of 32 uniformly spaced groups (z stride 72 rows = 8 groups)
    assert view_row(1, 0, 0, 0) - view_row(0, 0, 0, 120) == 9
