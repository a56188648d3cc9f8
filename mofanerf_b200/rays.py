"""Host-side ray generation (PyTorch plumbing)."""
from __future__ import annotations

import numpy as np
import torch


def get_rays(H, W, K, c2w):
    """tools/run_nerf_helpers.py:153-168.  Row-major (H, W): ray index = row * W + col."""
    dev = c2w.device if torch.is_tensor(c2w) else None
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev),
                          indexing="ij")
    i, j = i.t(), j.t()
    k = [[float(K[a][b]) for b in range(3)] for a in range(3)] if not torch.is_tensor(K) else K
    dirs = torch.stack([(i - k[0][2]) / k[0][0], -(j - k[1][2]) / k[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """tools/run_nerf_helpers.py:181-199 (forward-facing scenes; unused by the MoFaNeRF configs)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def pose_spherical(phi, theta, radius):
    """tools/load_facescape.py:9-38."""
    def trans_t(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], np.float32)

    def rot_y(p):
        return np.array([[np.cos(p), 0, -np.sin(p), 0], [0, 1, 0, 0], [np.sin(p), 0, np.cos(p), 0],
                         [0, 0, 0, 1]], np.float32)

    def rot_x(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0],
                         [0, 0, 0, 1]], np.float32)

    c2w = rot_y(phi / 180. * np.pi) @ (rot_x(theta / 180. * np.pi) @ trans_t(radius))
    return torch.tensor(c2w, dtype=torch.float32)


def pack_rays(rays_o, rays_d, near, far, viewdirs=None):
    """[N, 8 or 11] ray batch: o d near far (viewdir)   models/render_class.py:173-179."""
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near_t = near * torch.ones_like(rays_d[..., :1])
    far_t = far * torch.ones_like(rays_d[..., :1])
    rays = torch.cat([rays_o, rays_d, near_t, far_t], -1)
    if viewdirs is not None:
        # 11 floats per ray in the reference; one zero column pads the row to 48 bytes so the kernels read it as
        # three 128-bit loads (the engine accepts any row stride >= 11)
        rays = torch.cat([rays, viewdirs, torch.zeros_like(near_t)], -1)
    return rays
