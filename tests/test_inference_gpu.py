"""GPU (-m gpu): the SDFNet inference consumers and the voxel ingest (f2) against the CPU oracle (oracle/ref_render.py),
on the reference's own checkpoint (examples/gan_generator_voxels_chairs.to, shipped inside tests/golden/sdfnet_chairs.npz).

bf16 mode is the mode these kernels exist in (single-latent fused kernel); tolerances are that mode's: ~3e-3 relative on SDF values
(measured, profiles/r02*_parity_errors.txt).  Integer work (the sphere index, the ingest clamp/scale) is bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import ref_render as RR
from oracle import ref_torch as R

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _bf16():
    from shapegan_b200 import config
    old = config.precision()
    config.set_precision('bf16')
    yield
    config.set_precision(old)


def chairs():
    from model.sdf_net import SDFNet
    g = load_golden('sdfnet_chairs')
    sd = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('w.')}
    net = SDFNet()
    net.load_state_dict(sd, strict=True)
    return net, sd, torch.from_numpy(g['z'])


@pytest.mark.parametrize('r,count', [(8, None), (16, None), (32, 20360), (64, 169648)])
def test_sphere_index_is_bit_exact(r, count):
    """the device-built list of cells inside the 1.1 sphere == numpy's float32 mask (model/sdf_net.py:12), SURVEY App. A.6 counts"""
    from shapegan_b200.nn.sdf_net import _GridHelper
    mask, _ = RR.sphere_mask(r)
    h = _GridHelper('cuda', r, True)
    want = np.nonzero(mask)[0].astype(np.int32)
    assert np.array_equal(h.index.cpu().numpy(), want)
    if count is not None:
        assert h.count == count
    lin = np.linspace(-1, 1, r).astype(np.float32)
    assert np.array_equal(h.axis.cpu().numpy(), np.stack([lin] * 3))


def test_folded_single_latent_forward():
    """latent folded into the bias (no latent MMAs) == the general fused kernel with the latent broadcast by index == the oracle"""
    net, sd, z = chairs()
    g = torch.Generator().manual_seed(3)
    n = 256 * 150 + 19                                           # more tile pairs than SMs + a ragged tail
    pts = (torch.rand((n, 3), generator=g) * 2 - 1)
    ref = R.sdfnet_forward(sd, pts, z.reshape(1, -1).repeat(n, 1))
    with torch.no_grad():
        folded = net.evaluate_in_batches(pts.cuda(), z.cuda(), return_cpu_tensor=False)
        general = net(pts.cuda(), z.cuda().reshape(1, -1), torch.zeros(n, dtype=torch.int32, device='cuda'))
    assert rel_l2(folded, ref) < 6e-3 and rel_l2(general, ref) < 6e-3
    assert rel_l2(folded, general) < 4e-3                        # the fold keeps W z in fp32, the general path rounds z and W to bf16


@pytest.mark.parametrize('r,sphere_only', [(32, True), (16, False), (64, True)])
def test_get_voxels_fused(r, sphere_only):
    net, sd, z = chairs()
    vox = net.get_voxels(z.cuda(), r, sphere_only=sphere_only)
    ref = RR.voxelise(sd, z, r, sphere_only=sphere_only)
    assert vox.shape == ref.shape and vox.dtype == np.float32
    mask, _ = RR.sphere_mask(r)
    if sphere_only:
        outside = ~mask.reshape((r,) * 3)
        assert np.all(vox[outside] == 1.0)                       # untouched cells of the ones-filled grid: exact
    assert rel_l2(torch.from_numpy(vox), torch.from_numpy(ref)) < 6e-3
    assert np.mean(np.sign(vox) == np.sign(ref)) > 0.999         # inside / outside classification


def test_normals_against_oracle_autograd():
    """d sdf / d xyz out of the fused backward chain vs autograd through the CPU oracle.  Compared where the gradient means something:
    the SDF is clamped to +-0.1 in training, far from the surface it is flat and its 'normal' is the direction of rounding noise in
    the reference as well."""
    net, sd, z = chairs()
    g = torch.Generator().manual_seed(5)
    n = 128 * 37 + 5
    pts = (torch.rand((n, 3), generator=g) * 1.6 - 0.8)
    sdf_ref, n_ref = RR.normals(sd, z, pts)
    p = pts.cuda()
    normals = net.get_normals(z.cuda(), p)
    assert p.requires_grad and p.grad is normals                 # the reference's side effects on the caller's tensor (:121,:124)
    assert torch.allclose(normals.norm(dim=1), torch.ones(n, device='cuda'), atol=1e-4)
    near = sdf_ref.abs() < 0.05
    assert near.sum().item() > 300
    cos = (normals.cpu() * n_ref).sum(1)[near]
    assert cos.mean().item() > 0.998 and (cos > 0.99).float().mean().item() > 0.98, (cos.mean().item(), (cos > 0.99).float().mean().item())
    # the un-normalised gradient itself
    from shapegan_b200 import sdf_ops
    with torch.no_grad():
        sdf_dev, grad_dev = sdf_ops.normals_single_latent(net._params(), z.cuda(), pts.cuda(), normalize=False)
    pr = pts.clone().requires_grad_(True)
    R.sdfnet_forward(sd, pr, z.reshape(1, -1).repeat(n, 1)).sum().backward()
    assert rel_l2(sdf_dev, sdf_ref) < 6e-3
    assert rel_l2(grad_dev.cpu()[near], pr.grad[near]) < 6e-2      # seven bf16 layers forward + seven backward
    # no parameter gradients are produced on this path
    assert all(q.grad is None for q in net.parameters())


def test_surface_points_land_on_the_surface():
    net, sd, z = chairs()
    torch.manual_seed(0)
    pts, normals = net.get_surface_points(z.cuda(), sample_size=40000, sdf_cutoff=0.02, return_normals=True)
    assert pts.shape[0] > 500 and pts.shape == normals.shape
    with torch.no_grad():
        d = R.sdfnet_forward(sd, pts.cpu(), z.reshape(1, -1).repeat(pts.shape[0], 1))
    assert d.abs().median().item() < 5e-3                        # one projection step along the normal (:141) from within 0.02 of the surface


def _rays(res, radius=1.0):
    """a small pinhole bundle aimed at the origin, entry points on the bounding sphere (raymarching.py:70-99 in miniature)"""
    cam = np.array([1.4, 0.9, 1.7])
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0])); right /= np.linalg.norm(right)
    up = np.cross(fwd, right); up /= np.linalg.norm(up)
    u, v = np.meshgrid(np.linspace(-1, 1, res), np.linspace(-1, 1, res))
    focal = 1.0 / np.tan(np.arcsin(radius / np.linalg.norm(cam)))
    d = u.reshape(-1, 1) * right + v.reshape(-1, 1) * up + focal * fwd
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    p = np.tile(cam, (d.shape[0], 1)).astype(np.float32)
    b = np.einsum('ij,ij->i', p, d) * 2
    c = np.dot(cam, cam) - radius * radius
    with np.errstate(invalid='ignore'):
        t = (-b - np.sqrt(np.power(b, 2) - 4 * c)) / 2
    idx = np.argwhere(np.isfinite(t)).reshape(-1)
    p[idx] += d[idx] * t[idx, np.newaxis]
    return torch.from_numpy(p), torch.from_numpy(d), torch.from_numpy(idx)


def test_sphere_tracer_against_oracle_loop():
    """render_image's marching loop (raymarching.py:106-121) on the device vs the CPU restatement: same rays hit, same positions"""
    from shapegan_b200.rendering import trace_camera_rays
    net, sd, z = chairs()
    p0, d, idx = _rays(48)
    iters, thr = 150, 0.0005
    m_ref, p_ref = RR.march(sd, z, p0, d, idx.clone(), iters, 0.02, thr, 1.0)
    p_dev = p0.clone().cuda()
    m_dev = trace_camera_rays(net, z.cuda(), p_dev, d.cuda(), idx.cuda(), iterations=iters, threshold=thr)
    m_dev, p_dev = m_dev.cpu(), p_dev.cpu()
    agree = (m_dev == m_ref).float().mean().item()
    assert agree > 0.985, agree                                  # silhouette rays may flip under the bf16 SDF error
    both = (m_dev == 1) & (m_ref == 1)
    assert both.sum().item() > 200
    dist = (p_dev[both] - p_ref[both]).norm(dim=1)
    assert dist.median().item() < 2e-3 and (dist < 2e-2).float().mean().item() > 0.97, (dist.median().item(), dist.max().item())
    untouched = torch.ones(p0.shape[0], dtype=torch.bool); untouched[idx] = False
    assert torch.equal(p_dev[untouched], p0[untouched])          # rays that miss the bounding sphere are never touched


def test_shadow_tracer_against_oracle_loop():
    from shapegan_b200.rendering import trace_shadow_rays
    net, sd, z = chairs()
    torch.manual_seed(1)
    surf = net.get_surface_points(z.cuda(), sample_size=4000).cpu()[:1500]
    light = torch.tensor([2.0, 4.0, 3.0])
    d = light.unsqueeze(0) - surf
    d = d / d.norm(dim=1, keepdim=True)
    start = surf + d * 0.1                                       # raymarching.py:41
    m_ref, _ = RR.march(sd, z, start, d, torch.arange(start.shape[0]), 60, 0.1, 0.001, 1.0, miss_y=True)
    m_dev = trace_shadow_rays(net, z.cuda(), start.clone().cuda(), d.cuda(), iterations=60)
    assert (m_dev.cpu() == m_ref).float().mean().item() > 0.97


def test_voxel_ingest_bit_exact():
    from shapegan_b200 import raw
    g = torch.Generator().manual_seed(7)
    x = torch.randn((5, 32, 32, 32), generator=g) * 0.08
    x[0, 0, 0, :5] = torch.tensor([0.1, -0.1, 0.0999999, 0.3, -7.0])
    for clamp, rescale in ((0.1, True), (0.1, False), (0.05, True)):
        want = RR.ingest(x.numpy(), clamp, rescale)
        got = raw.voxel_ingest(x.cuda(), clamp, rescale).cpu()
        assert torch.equal(got, want), (clamp, rescale)
    odd = torch.randn((1003,), generator=g)                      # length not a multiple of 4
    assert torch.equal(raw.voxel_ingest(odd.cuda()).cpu(), RR.ingest(odd.numpy()))


def test_voxel_batch_stream_matches_dataset(tmp_path):
    """VoxelBatchStream (pinned staging -> side-stream H2D -> sg_voxel_ingest) delivers exactly what VoxelDataset.__getitem__ +
    DataLoader collation deliver (datasets.py:16-23), incl. the short last batch."""
    from shapegan_b200.data import VoxelDataset
    g = torch.Generator().manual_seed(11)
    files = []
    for i in range(11):
        f = str(tmp_path / ('%03d.npy' % i))
        np.save(f, (torch.randn((16, 16, 16), generator=g) * 0.07).numpy())
        files.append(f)
    ds = VoxelDataset(files)
    want = torch.stack([ds[i] for i in range(len(ds))])
    got = torch.cat([b.clone() for b in ds.stream(4, shuffle=False)]).cpu()
    assert got.shape == want.shape and torch.equal(got, want)
    assert len(ds.stream(4)) == 3 and len(ds.stream(4, drop_last=True)) == 2
    seen = torch.cat([b.clone() for b in ds.stream(4, shuffle=True, seed=3)]).cpu()
    assert torch.equal(seen.sum(dim=(1, 2, 3)).sort().values, want.sum(dim=(1, 2, 3)).sort().values)      # a permutation of the same items
