"""GPU (-m gpu): sg_mc_* against the numpy oracle (oracle/mc_tables.py): same vertices (bit-exact), same faces (identical indices in
identical order), normals to 1e-5; and SDFNet.get_mesh on the reference's chairs checkpoint yields a closed surface."""
import numpy as np
import pytest
import torch

from oracle import mc_tables as M

pytestmark = pytest.mark.gpu


def _vol(kind, shape, seed=0):
    rng = np.random.default_rng(seed)
    if kind == 'noise':
        return np.pad(rng.standard_normal(tuple(s - 2 for s in shape)), 1, constant_values=5.0).astype(np.float32)
    ax = [np.linspace(-1, 1, s) for s in shape]
    X, Y, Z = np.meshgrid(*ax, indexing='ij')
    return (np.sqrt(X ** 2 + 0.7 * Y ** 2 + Z ** 2) - 0.55 + 0.05 * np.sin(7 * X) * np.cos(5 * Z)).astype(np.float32)


@pytest.mark.parametrize('kind,shape,level,spacing', [('blob', (24, 24, 24), 0.0, (1.0, 1.0, 1.0)), ('noise', (13, 17, 19), 0.0, (0.5, 0.25, 2.0)),
                                                     ('blob', (33, 20, 27), 0.1, (2 / 32,) * 3), ('noise', (40, 9, 66), -0.3, (1.0, 1.0, 1.0))])
def test_gpu_marching_cubes_equals_oracle(kind, shape, level, spacing):
    from shapegan_b200.mesh import marching_cubes
    vol = _vol(kind, shape)
    v_ref, f_ref, n_ref = M.marching_cubes(vol, level, spacing)
    v, f, n, vals = marching_cubes(vol, level, spacing)
    assert v.shape == v_ref.shape and f.shape == f_ref.shape
    assert np.array_equal(f, f_ref)
    assert np.array_equal(v, v_ref)                                 # same float32 operations in the same order
    assert np.abs(n - n_ref).max() < 1e-5
    assert np.all(vals == np.float32(level))


def test_empty_volume_raises_like_skimage():
    from shapegan_b200.mesh import marching_cubes
    with pytest.raises(ValueError):
        marching_cubes(np.ones((8, 8, 8), dtype=np.float32), 0.0)


def test_get_mesh_on_chairs_checkpoint():
    from conftest import load_golden
    from model.sdf_net import SDFNet
    from test_mc_cpu import manifold_defects
    g = load_golden('sdfnet_chairs')
    net = SDFNet()
    net.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('w.')}, strict=True)
    mesh = net.get_mesh(torch.from_numpy(g['z']).cuda(), voxel_resolution=48)
    assert mesh is not None and len(mesh.faces) > 500
    assert manifold_defects(mesh.faces) == 0                        # the padded grid closes the surface
    v = np.asarray(mesh.vertices)
    assert v.min() > -1.1 and v.max() < 1.1
    pts = mesh.sample(256) if hasattr(mesh, 'sample') else v[:256]
    assert pts.shape == (256, 3)
