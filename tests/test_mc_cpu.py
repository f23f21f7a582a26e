"""CPU: the generated marching-cubes case table (oracle/mc_tables.py) and its numpy implementation -- the oracle of the GPU
marching cubes.  skimage's marching_cubes_lewiner (what the reference calls, model/sdf_net.py:103) is absent from
this image: its triangulation is unpinned; these properties are what is pinned."""
import os

import numpy as np

from oracle import mc_tables as M

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def manifold_defects(faces):
    de = {}
    for tri in faces:
        for i in range(3):
            k = (int(tri[i]), int(tri[(i + 1) % 3]))
            de[k] = de.get(k, 0) + 1
    return sum(1 for (a, b), c in de.items() if c != 1 or de.get((b, a), 0) != 1)


def test_header_in_sync_with_generator():
    import sys
    sys.path.insert(0, os.path.join(REPO, 'tools'))
    import gen_mc_tables
    with open(os.path.join(REPO, 'shapegan_b200', 'csrc', 'sg_mc_tables.h')) as f:
        assert f.read() == gen_mc_tables.render()


def test_table_structure():
    count, edges, ec = M.tables()
    assert count[0] == 0 and count[255] == 0 and count.max() == 5 and int(count.sum()) == 820      # same budget as the classic table
    for case in range(256):
        used = edges[case][edges[case] >= 0]
        assert len(used) == 3 * count[case]
        # a triangle's vertices sit on edges whose two corners have different signs
        for e in used:
            c0, c1 = ec[e]
            assert ((case >> c0) & 1) != ((case >> c1) & 1)
        # complementary configuration: the same surface with the opposite orientation
        comp = edges[255 - case][edges[255 - case] >= 0]
        assert sorted(set(used.tolist())) == sorted(set(comp.tolist()))


def test_sphere_is_closed_oriented_and_on_the_level_set():
    r = 20
    ax = np.linspace(-1, 1, r)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    vol = (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6).astype(np.float32)
    h = 2 / (r - 1)
    v, f, n = M.marching_cubes(vol, 0.0, spacing=(h, h, h))
    assert manifold_defects(f) == 0
    assert len(v) - 3 * len(f) // 2 + len(f) == 2                                   # Euler characteristic of a sphere
    c = v - 1.0
    rad = np.linalg.norm(c, axis=1)
    assert rad.min() > 0.59 and rad.max() < 0.601
    tri = c[f]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert np.all(np.einsum('ij,ij->i', fn, tri.mean(1)) > 0)                       # faces point towards increasing value (outward)
    assert np.all(np.einsum('ij,ij->i', n, c) > 0.9 * rad)                          # vertex normals ~ radial
    area = 0.5 * np.linalg.norm(fn, axis=1).sum()
    assert abs(area - 4 * np.pi * 0.36) / (4 * np.pi * 0.36) < 0.02


def test_random_fields_are_watertight():
    rng = np.random.default_rng(0)
    for _ in range(3):
        vol = np.pad(rng.standard_normal((10, 11, 12)), 1, constant_values=5.0).astype(np.float32)
        v, f, n = M.marching_cubes(vol, 0.0)
        assert len(f) > 1000 and manifold_defects(f) == 0
        # every vertex lies on exactly one grid edge
        frac = np.abs(v - np.round(v))
        assert np.all((frac > 1e-7).sum(1) <= 1)
