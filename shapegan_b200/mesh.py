"""GPU marching cubes: the mesh extraction behind SDFNet.get_mesh (model/sdf_net.py:97-113), which the reference
delegates to skimage.measure.marching_cubes_lewiner (third party, absent from this image, removed from current scikit-image).

`marching_cubes(volume, level, spacing)` keeps that routine's calling convention and return tuple (verts, faces, normals, values);
the triangulation comes from a generated case table (oracle/mc_tables.py) -- closed and consistently oriented, one vertex per
crossing cell edge (shared between the triangles around it, like skimage's indexed output), deterministic order."""
import ctypes

import numpy as np
import torch

from . import _lib as L
from . import raw


def marching_cubes(volume, level=0.0, spacing=(1.0, 1.0, 1.0), return_torch=False):
    """volume: [nx, ny, nz] float32 (numpy array or CUDA tensor).  Returns (verts [V,3], faces [F,3] int32, normals [V,3], values [V]).
    Raises ValueError when the level set does not cross the volume (skimage's behaviour, relied on at model/sdf_net.py:104-108)."""
    vol = torch.as_tensor(volume, dtype=torch.float32)
    if not vol.is_cuda:
        vol = vol.cuda()
    vol = vol.contiguous()
    if vol.dim() != 3:
        raise ValueError('marching_cubes: volume must be 3-D')
    nx, ny, nz = (int(v) for v in vol.shape)
    dev = vol.device
    entries = L.lib().sg_mc_workspace_entries(nx, ny, nz)
    sums = torch.empty(entries, dtype=torch.int64, device=dev)
    a = L.SgMcArgs()
    a.volume, a.nx, a.ny, a.nz, a.level = ctypes.c_void_p(vol.data_ptr()), nx, ny, nz, float(level)
    a.spacing[0], a.spacing[1], a.spacing[2] = (float(s) for s in spacing)
    a.block_sums = ctypes.c_void_p(sums.data_ptr())
    L.check(L.lib().sg_mc_count(ctypes.byref(a), raw.stream()), 'sg_mc_count')
    totals = int(sums[-1].item())                               # the one host read: output sizes
    nv, nf = totals & 0xffffffff, (totals >> 32) & 0xffffffff
    if nf == 0:
        raise ValueError('Surface level must be within volume data range.')
    verts = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    normals = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((nf, 3), dtype=torch.int32, device=dev)
    vbase = torch.empty(nx * ny * nz, dtype=torch.int32, device=dev)
    a.vbase, a.vertices = ctypes.c_void_p(vbase.data_ptr()), ctypes.c_void_p(verts.data_ptr())
    a.normals, a.faces = ctypes.c_void_p(normals.data_ptr()), ctypes.c_void_p(faces.data_ptr())
    L.check(L.lib().sg_mc_emit(ctypes.byref(a), raw.stream()), 'sg_mc_emit')
    values = torch.full((nv,), float(level), dtype=torch.float32, device=dev)
    if return_torch:
        return verts, faces, normals, values
    return verts.cpu().numpy(), faces.cpu().numpy(), normals.cpu().numpy(), values.cpu().numpy()


class Mesh:
    """Minimal stand-in for trimesh.Trimesh when trimesh is not installed: `.vertices`, `.faces`, `.vertex_normals`."""

    def __init__(self, vertices, faces, vertex_normals):
        self.vertices, self.faces, self.vertex_normals = vertices, faces, vertex_normals

    def sample(self, count):
        """area-weighted uniform surface samples (what trimesh.Trimesh.sample does; model/sdf_net.py:115-116)"""
        tri = self.vertices[self.faces]
        area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
        pick = np.random.choice(len(area), size=count, p=area / area.sum())
        u = np.random.rand(count, 2)
        flip = u.sum(1) > 1
        u[flip] = 1 - u[flip]
        t = tri[pick]
        return t[:, 0] + (t[:, 1] - t[:, 0]) * u[:, :1] + (t[:, 2] - t[:, 0]) * u[:, 1:]
