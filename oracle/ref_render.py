"""ORACLE — test infrastructure, NOT product code.

CPU restatement (numpy / torch CPU fp32 on top of oracle/ref_torch.py) of the SDFNet inference consumers of the reference:
  voxelise()        SDFVoxelizationHelperData + SDFNet.get_voxels            model/sdf_net.py:7-19, :77-95
  normals()         SDFNet.get_normals (autograd through the MLP)            model/sdf_net.py:118-128
  march()           the sphere-tracing loops of render_image / get_shadows   rendering/raymarching.py:106-121, :48-64
  ingest()          VoxelDataset.__getitem__                                 datasets.py:16-23
The reference has no tests for these: parity is pinned by running these restatements against the CUDA path on the
reference's own checkpoint (examples/gan_generator_voxels_chairs.to, stored in tests/golden/sdfnet_chairs.npz)."""
import numpy as np
import torch

from . import ref_torch as R


def sphere_mask(resolution):
    """model/sdf_net.py:9-14: float32 grid points, float32 norm, < 1.1"""
    pts = R.voxel_coordinates(resolution).numpy()
    return np.linalg.norm(pts, axis=1) < 1.1, pts


def voxelise(sd, latent, resolution, sphere_only=True, pad=True):
    mask, pts = sphere_mask(resolution)
    if sphere_only:
        pts = pts[mask, :]
    p = torch.from_numpy(pts)
    with torch.no_grad():
        d = R.sdfnet_forward(sd, p, latent.reshape(1, -1).repeat(p.shape[0], 1)).numpy()         # :63-75
    if sphere_only:
        vox = np.ones((resolution,) * 3, dtype=np.float32)                                     # :88
        vox[mask.reshape((resolution,) * 3)] = d                                                # :89
    else:
        vox = d.reshape((resolution,) * 3)
        if pad:
            vox = np.pad(vox, 1, mode='constant', constant_values=1)
    return vox


def normals(sd, latent, points):
    p = points.clone().requires_grad_(True)                                                     # :121
    sdf = R.sdfnet_forward(sd, p, latent.reshape(1, -1).repeat(p.shape[0], 1))                  # :122
    sdf.backward(torch.ones(sdf.shape[0]))                                                      # :123
    g = p.grad
    return sdf.detach(), g / torch.norm(g, dim=1).unsqueeze(1)                                  # :125-126


def march(sd, latent, points, directions, indices, iterations, clamp, threshold, radius, sdf_offset=0.0, miss_y=False):
    """Returns (hit mask uint8 [N], points after marching).  indices: int64 list of the rays to trace."""
    points = points.clone()
    mask = torch.zeros(points.shape[0], dtype=torch.uint8)
    z = latent.reshape(1, -1)
    with torch.no_grad():
        for _ in range(iterations):
            test = points[indices, :]
            sdf = R.sdfnet_forward(sd, test, z.repeat(test.shape[0], 1)).reshape(-1) + sdf_offset       # :108 / :51
            sdf = torch.clamp(sdf, -clamp, clamp)                                                       # :109 / :52
            points[indices, :] += directions[indices, :] * sdf.unsqueeze(1)                             # :110 / :53
            hits = (sdf > 0) & (sdf < threshold)                                                        # :112 / :55
            mask[indices[hits]] = 1
            indices = indices[~hits]
            if miss_y:
                misses = points[indices, 1] > radius                                                    # :59
            else:
                misses = torch.norm(points[indices, :], dim=1) > radius                                 # :116
            indices = indices[~misses]
            if indices.shape[0] < 2:                                                                    # :119 / :62
                break
    mask[indices] = 1                                                                                   # :121 / :64
    return mask, points


def ingest(raw, clamp=0.1, rescale_sdf=True):
    """datasets.py:16-23 on an in-memory float32 array"""
    result = torch.from_numpy(np.array(raw, dtype=np.float32, copy=True))
    if clamp is not None:
        result.clamp_(-clamp, clamp)
        if rescale_sdf:
            result /= clamp
    return result
