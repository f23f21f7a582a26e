"""DeepSDF MLP — drop-in for model/sdf_net.py: `SDFNet(latent_code_size=128, device='cuda')`, state_dict keys
`layers1.{0,2,4,6}.*` / `layers2.{0,2,4,6}.*`, `forward(points, latent_codes)` and the inference helpers.

Extension (kept out of the reference signature's way): `forward(points, latent_codes, shape_index=None, points_per_shape=0)` — when
`shape_index` (int [N]) is given, `latent_codes` is a [S, L] table and row i uses `latent_codes[shape_index[i]]`.
This is what the reference's callers compute with `latent_codes[model_indices, :]` (train_sdf_autodecoder.py:80) or
`.repeat(...)` (sdf_net.py:64, train_hybrid_progressive_gan.py:92) without materialising the [N, L] copy.  `points_per_shape` = P
promises `shape_index[i] == i // P` (consecutive blocks of P points per shape, the layout of BASELINE configs[2] and of the hybrid
GAN's grids): the backward then forms the latent gradients per shape instead of per point."""
import os

import numpy as np
import torch
import torch.nn as nn

from .. import raw, sdf_ops
from ..sdf_ops import sdfnet_apply
from . import LATENT_CODE_SIZE, SavableModule, _require_cuda

SDF_NET_BREADTH = 256

sdf_voxelization_helper = dict()


def get_voxel_coordinates(resolution=32, size=1, center=0, return_torch_tensor=False, device=None):
    """util.py:60-74 semantics: row i*R^2 + j*R + k <-> (x_i, y_j, z_k); inclusive float64 linspace cast to float32."""
    if type(center) == int:
        center = (center, center, center)
    axes = [np.linspace(c - size, c + size, resolution) for c in center]
    grid = np.stack(np.meshgrid(*axes, indexing='ij')).reshape(3, -1).transpose()
    if return_torch_tensor:
        return torch.tensor(grid, dtype=torch.float32, device=device)
    return grid.astype(np.float32)


def get_points_in_unit_sphere(n, device):
    """util.py:32-39: rejection sampling from the cube."""
    x = torch.rand(int(n * 2.5), 3, device=device) * 2 - 1
    keep = (torch.norm(x, dim=1) < 1).nonzero().squeeze()[:n]
    x = x[keep, :]
    if x.shape[0] < n:
        print("Warning: Did not find enough points.")
    return x


class _GridHelper:
    """Device-side replacement of SDFVoxelizationHelperData for the fused voxelisation: axis tables [3, R] (the float32(float64
    linspace) values of util.py:60-74) and the sorted list of cells with |p| < 1.1 (model/sdf_net.py:12), built on the device."""

    def __init__(self, device, r, sphere_only):
        axis = np.stack([np.linspace(-1, 1, r)] * 3).astype(np.float32)
        self.axis = torch.tensor(axis, device=device)
        self.index = raw.grid_sphere_index(r, self.axis, 1.1) if sphere_only else None
        self.count = int(self.index.shape[0]) if sphere_only else r * r * r


_grid_helpers = dict()


def _grid_helper(device, r, sphere_only):
    key = (str(device), r, sphere_only)
    if key not in _grid_helpers:
        _grid_helpers[key] = _GridHelper(device, r, sphere_only)
    return _grid_helpers[key]


class SDFVoxelizationHelperData():
    """model/sdf_net.py:7-19"""

    def __init__(self, device, voxel_resolution, sphere_only=True):
        sample_points = get_voxel_coordinates(voxel_resolution)
        if sphere_only:
            unit_sphere_mask = np.linalg.norm(sample_points, axis=1) < 1.1
            sample_points = sample_points[unit_sphere_mask, :]
            self.unit_sphere_mask = unit_sphere_mask.reshape(voxel_resolution, voxel_resolution, voxel_resolution)
        self.sample_points = torch.tensor(sample_points, device=device)
        self.point_count = self.sample_points.shape[0]


class SDFNet(SavableModule):
    def __init__(self, latent_code_size=LATENT_CODE_SIZE, device='cuda'):
        super().__init__(filename="sdf_net.to")
        self.latent_code_size = latent_code_size
        w = SDF_NET_BREADTH
        first = [nn.Linear(3 + latent_code_size, w), nn.ReLU(inplace=True)]
        for _ in range(3):
            first += [nn.Linear(w, w), nn.ReLU(inplace=True)]
        self.layers1 = nn.Sequential(*first)
        second = [nn.Linear(w + latent_code_size + 3, w), nn.ReLU(inplace=True)]
        for _ in range(2):
            second += [nn.Linear(w, w), nn.ReLU(inplace=True)]
        second += [nn.Linear(w, 1), nn.Tanh()]
        self.layers2 = nn.Sequential(*second)
        if device == 'cuda' and not torch.cuda.is_available():
            device = 'cpu'       # parameters stay inspectable on a CPU-only host; forward() still requires CUDA
        self.to(device)

    def _params(self):
        ps = []
        for seq in (self.layers1, self.layers2):
            for i in (0, 2, 4, 6):
                ps += [seq[i].weight, seq[i].bias]
        return ps

    def forward(self, points, latent_codes, shape_index=None, points_per_shape=0):
        _require_cuda(points, 'SDFNet.forward')
        n = points.shape[0]
        if n == 0:
            return torch.zeros((0,), dtype=torch.float32, device=points.device)
        # the kernels take raw pointers: shape errors the reference would get from torch.cat (model/sdf_net.py:57) are raised here
        if points.dim() != 2 or points.shape[1] != 3:
            raise RuntimeError('SDFNet.forward: points must be [N, 3], got %s' % (tuple(points.shape),))
        if latent_codes.dim() != 2 or latent_codes.shape[1] != self.latent_code_size:
            raise RuntimeError('SDFNet.forward: latent_codes must be [*, %d], got %s' % (self.latent_code_size, tuple(latent_codes.shape)))
        idx = None
        if shape_index is not None:
            if shape_index.numel() != n:
                raise RuntimeError('SDFNet.forward: shape_index has %d entries for %d points' % (shape_index.numel(), n))
            idx = shape_index.to(device=points.device, dtype=torch.int32).contiguous()
            if os.environ.get('SG_B200_CHECK_INDEX') == '1' and n > 0:      # debug only: synchronises
                lo, hi = int(idx.min()), int(idx.max())
                if lo < 0 or hi >= latent_codes.shape[0]:
                    raise RuntimeError('SDFNet.forward: shape_index out of range [0, %d)' % latent_codes.shape[0])
        elif latent_codes.shape[0] != n:
            raise RuntimeError('SDFNet.forward: %d latent rows for %d points (sizes must match, model/sdf_net.py:57)' % (latent_codes.shape[0], n))
        out = sdfnet_apply(points.float(), latent_codes.float(), idx, self._params(), seg_len=points_per_shape if idx is not None else 0)
        return out.squeeze()

    # ------------------------------------------------------------------ inference helpers (model/sdf_net.py:63-168)
    def evaluate_in_batches(self, points, latent_code, batch_size=100000, return_cpu_tensor=True):
        """One latent code for all points (:63-75).  bf16 mode: ONE launch of the folded single-latent kernel over all points (the
        latent is a constant of the call: W[:, latent] z goes into the bias, see sdf_ops.folded_weights) instead of chunks of
        `batch_size` with `.repeat(batch_size, 1)`; fp32x mode: chunked like the reference, latent broadcast by index."""
        n = points.shape[0]
        _require_cuda(points, 'SDFNet.evaluate_in_batches')
        with torch.no_grad():
            if n > 0 and sdf_ops.folded_enabled(self.latent_code_size):
                result = torch.empty((n,), dtype=torch.float32, device=points.device)
                sdf_ops.infer_single_latent(self._params(), latent_code.to(points.device), n=n, out=result, points=points.float().contiguous())
            else:
                table = latent_code.reshape(1, -1).to(points.device)
                result = torch.zeros((n,), device=points.device)
                for start in range(0, n, batch_size):
                    chunk = points[start:start + batch_size, :]
                    zeros = torch.zeros((chunk.shape[0],), dtype=torch.int32, device=points.device)
                    result[start:start + batch_size] = self(chunk, table, zeros)
        return result.cpu() if return_cpu_tensor else result

    def get_voxels(self, latent_code, voxel_resolution, sphere_only=True, pad=True):
        """:77-95.  bf16 mode: the grid never exists as a point list -- the kernel derives the coordinates of cell s from the axis
        tables, only the cells inside the 1.1 sphere are listed (device-built index, cached per resolution) and every result is
        scattered straight into the ones-filled grid."""
        r = voxel_resolution
        if sdf_ops.folded_enabled(self.latent_code_size):
            helper = _grid_helper(self.device, r, sphere_only)
            with torch.no_grad():
                grid = torch.ones((r * r * r,), dtype=torch.float32, device=self.device)
                sdf_ops.infer_single_latent(self._params(), latent_code.to(self.device), n=helper.count, out=grid, ray_index=helper.index,
                                            grid_r=r, grid_axis=helper.axis)
            voxels = grid.reshape(r, r, r).cpu().numpy()
            if not sphere_only and pad:
                voxels = np.pad(voxels, 1, mode='constant', constant_values=1)
            return voxels
        key = (voxel_resolution, sphere_only)
        if key not in sdf_voxelization_helper:
            sdf_voxelization_helper[key] = SDFVoxelizationHelperData(self.device, voxel_resolution, sphere_only)
        helper_data = sdf_voxelization_helper[key]
        with torch.no_grad():
            distances = self.evaluate_in_batches(helper_data.sample_points, latent_code).numpy()
        if sphere_only:
            voxels = np.ones((voxel_resolution, voxel_resolution, voxel_resolution), dtype=np.float32)
            voxels[helper_data.unit_sphere_mask] = distances
        else:
            voxels = distances.reshape(voxel_resolution, voxel_resolution, voxel_resolution)
            if pad:
                voxels = np.pad(voxels, 1, mode='constant', constant_values=1)
        return voxels

    def get_mesh(self, latent_code, voxel_resolution=64, sphere_only=True, raise_on_empty=False, level=0):
        """:97-113 with the GPU marching cubes of shapegan_b200.mesh in place of skimage.measure.marching_cubes_lewiner.  Returns a
        trimesh.Trimesh when trimesh is installed (the reference's return type), else a minimal object with the same three fields."""
        from ..mesh import Mesh, marching_cubes
        size = 2
        voxels = self.get_voxels(latent_code, voxel_resolution=voxel_resolution, sphere_only=sphere_only)
        voxels = np.pad(voxels, 1, mode='constant', constant_values=1)
        spacing = (size / voxel_resolution,) * 3
        try:
            vertices, faces, normals, _ = marching_cubes(voxels, level=level, spacing=spacing)
        except ValueError as value_error:
            if raise_on_empty:
                raise value_error
            return None
        vertices -= size / 2
        try:
            import trimesh
        except ImportError:
            return Mesh(vertices, faces, normals)
        return trimesh.Trimesh(vertices=vertices, faces=faces, vertex_normals=normals)

    def get_uniform_surface_points(self, latent_code, point_count=1000, voxel_resolution=64, sphere_only=True, level=0):
        mesh = self.get_mesh(latent_code, voxel_resolution=voxel_resolution, sphere_only=sphere_only, level=level)
        return mesh.sample(point_count)

    def _sdf_and_gradient(self, latent_code, points):
        """(sdf [n], d sdf / d xyz [n,3]) of one latent code; `points.grad` receives the gradient like the reference's backward()."""
        if sdf_ops.folded_enabled(self.latent_code_size) and points.shape[0] > 0:
            # analytic path: folded forward (ReLU masks only) + fused backward chain with the xyz gradient in its drains.  Unlike
            # the reference's sdf.backward() (:122-126) this does NOT accumulate into the parameters' .grad -- nobody reads those.
            with torch.no_grad():
                sdf, g = sdf_ops.normals_single_latent(self._params(), latent_code.to(points.device), points.detach().float().contiguous(),
                                                       normalize=False)
            points.requires_grad = True
            points.grad = g
            return sdf, g
        points.requires_grad = True
        zeros = torch.zeros((points.shape[0],), dtype=torch.int32, device=points.device)
        sdf = self(points, latent_code.reshape(1, -1), zeros)
        sdf.backward(torch.ones(sdf.shape[0], device=self.device))
        return sdf.detach(), points.grad

    def get_normals(self, latent_code, points):
        if latent_code.requires_grad or points.requires_grad:
            raise Exception('get_normals may only be called with tensors that don\'t require grad.')
        _, normals = self._sdf_and_gradient(latent_code, points)
        normals /= torch.norm(normals, dim=1).unsqueeze(dim=1)
        return normals

    def get_surface_points(self, latent_code, sample_size=100000, sdf_cutoff=0.1, return_normals=False, use_unit_sphere=True):
        if use_unit_sphere:
            points = get_points_in_unit_sphere(n=sample_size, device=self.device) * 1.1
        else:
            points = torch.rand((sample_size, 3), device=self.device) * 2.2 - 1
        sdf, normals = self._sdf_and_gradient(latent_code, points)
        normals /= torch.norm(normals, dim=1).unsqueeze(dim=1)
        points.requires_grad = False
        points -= normals * sdf.detach().unsqueeze(dim=1)        # project onto the surface along the normal
        mask = (torch.abs(sdf) < sdf_cutoff) & torch.all(torch.isfinite(points), dim=1)
        points = points[mask, :]
        normals = normals[mask, :]
        return (points, normals) if return_normals else points

    def get_surface_points_in_batches(self, latent_code, amount=1000):
        result = torch.zeros((amount, 3), device=self.device)
        position = 0
        iteration_limit = 20
        while position < amount and iteration_limit > 0:
            points = self.get_surface_points(latent_code, sample_size=amount * 6)
            amount_used = min(amount - position, points.shape[0])
            result[position:position + amount_used, :] = points[:amount_used, :]
            position += amount_used
            iteration_limit -= 1
        return result
