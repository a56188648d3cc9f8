"""CPU/fp32 ORACLE for the MoFaNeRF ray-marching hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain device-agnostic PyTorch fp32, the algorithm of the
reference's hot path.  It exists so tests / smoke() / bench.py's cpu_baseline leg can
check the CUDA engine; nothing in ``mofanerf_b200/`` may import it.

Parity pin: the reference ships no tests or golden vectors, so this
restatement is pinned against outputs of the *reference itself* imported in the build
container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``); see
``tests/test_oracle_golden.py``.

Every function cites the reference file:line (relative to /root/reference) it follows.
Differences from the reference, all results-neutral:
  * no ``.cuda()`` calls, no global default-device assumptions;
  * random draws are explicit inputs (``t_rand``, ``u``, ``noise``) so a CUDA kernel can
    be fed the same numbers (reference draws them from torch's global generator:
    models/render_class.py:305,463; tools/run_nerf_helpers.py:216).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# Positional encoding                                   models/model.py:15-63
# ----------------------------------------------------------------------------------------
def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    """cat[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (model.py:24-45).

    freq_bands = 2 ** linspace(0, L-1, L) (log_sampling=True, model.py:31-32)."""
    freqs = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)
    out = [x]
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def embed_dim(multires: int) -> int:
    return 3 + 3 * 2 * multires


# ----------------------------------------------------------------------------------------
# Networks (same module tree => same state_dict keys and same init RNG order)
# ----------------------------------------------------------------------------------------
class SkipMLP(nn.Module):
    """models/model.py:202-230."""

    def __init__(self, D=8, W=256, input_ch=256, skip=None):
        super().__init__()
        self.skips = skip
        self.linears1 = nn.Sequential()
        self.linears1.add_module("Linear0", nn.Linear(input_ch, W))
        self.linears1.add_module("relu0", nn.ReLU())
        self.linears2 = nn.Sequential()
        if skip is not None:
            for i in range(skip):
                self.linears1.add_module(f"Linear{i + 1}", nn.Linear(W, W))
                self.linears1.add_module(f"relu{i + 1}", nn.ReLU())
            self.linears2.add_module("Linear0", nn.Linear(W + input_ch, W))
            self.linears2.add_module("relu0", nn.ReLU())
            for i in range(D - skip - 2):
                self.linears2.add_module(f"Linear{i + 1}", nn.Linear(W, W))
                self.linears2.add_module(f"relu{i + 1}", nn.ReLU())
        else:
            for i in range(D):
                self.linears1.add_module(f"Linear{i + 1}", nn.Linear(W, W))
                self.linears1.add_module(f"relu{i + 1}", nn.ReLU())
        _xavier_relu(self)

    def forward(self, x):
        h = self.linears1(x)
        if self.skips is not None:
            h = self.linears2(torch.cat([x, h], dim=1))  # model.py:229 order [x_in, h]
        return h


def _xavier_relu(mod: nn.Module) -> None:
    """model.py:139-142, 190-193, 232-244: xavier_uniform(gain=relu) on every Linear weight."""
    for m in mod.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight.data, gain=nn.init.calculate_gain("relu"))


class NeRF(nn.Module):
    """models/model.py:80-137 (use_viewdirs=True branch, the only one the configs use)."""

    def __init__(self, D=8, W=256, input_ch=93, input_ch_views=27, input_ch_textureCodes=256,
                 input_ch_shapeCodes=50):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.input_ch_shapeCodes = input_ch_shapeCodes
        self.input_ch_textureCodes = input_ch_textureCodes
        self.xyzEncode = SkipMLP(D=3, W=W, input_ch=input_ch, skip=None)
        self.linear_BiM_xyz = SkipMLP(D=D, W=W, input_ch=input_ch_shapeCodes + W, skip=4)
        self.linear_uv_xyzBiM = SkipMLP(D=D, W=W, input_ch=input_ch_textureCodes + W, skip=4)
        self.linear_view_xyBMuv = nn.Sequential(nn.Linear(input_ch_views + W, W // 2), nn.ReLU())
        self.alpha_linear = nn.Sequential(nn.Linear(W, 1))
        self.rgb_linear = nn.Linear(W // 2, 3)
        _xavier_relu(self)

    def forward(self, input_pts, input_bmCodes, input_views, input_uvCodes):
        xyz_code = self.xyzEncode(input_pts)                                        # :127
        sigmaCodes = self.linear_BiM_xyz(torch.cat([input_bmCodes, xyz_code], 1))   # :129
        alpha = self.alpha_linear(sigmaCodes)                                       # :130
        rgbCodes = self.linear_uv_xyzBiM(torch.cat([input_uvCodes, sigmaCodes], 1))  # :132
        rgbCodes = self.linear_view_xyBMuv(torch.cat([input_views, rgbCodes], 1))   # :133
        rgb = self.rgb_linear(rgbCodes)                                             # :134
        return torch.cat([rgb, alpha], -1)                                          # :135


class StyleModule(nn.Module):
    """models/model.py:174-199: shape code -> (expression scale, expression bias)."""

    def __init__(self, D=4, W=256, input_ch_bm=50, out_ch=30):
        super().__init__()
        self.linears1 = nn.Sequential()
        for i in range(D):
            self.linears1.add_module(f"Linear{i}", nn.Linear(input_ch_bm if i == 0 else W, W))
            self.linears1.add_module(f"relu{i}", nn.ReLU())
        self.linears_scale = nn.Linear(W, out_ch)
        self.linears_bias = nn.Linear(W, out_ch)
        _xavier_relu(self)

    def forward(self, bmcodes):
        feature = self.linears1(bmcodes)
        return self.linears_scale(feature), self.linears_bias(feature)


def build_nets(seed: int = 0, W_c: int = 256, D_c: int = 8, W_f: int = 1024, D_f: int = 10,
               multires: int = 10, multires_views: int = 4, n_shape: int = 50, n_tex: int = 256,
               n_exp: int = 30):
    """Seeded random-init nets in the construction order of tools/create_model_condition.py:16-50
    (coarse NeRF, fine NeRF, then the renderer's StyleModule).  Returns (coarse, fine, style)."""
    torch.manual_seed(seed)
    in_ch = embed_dim(multires) + n_exp
    in_v = embed_dim(multires_views)
    coarse = NeRF(D_c, W_c, in_ch, in_v, n_tex, n_shape)
    fine = NeRF(D_f, W_f, in_ch, in_v, n_tex, n_shape) if W_f else None
    style = StyleModule(input_ch_bm=n_shape, out_ch=n_exp)
    for m in (coarse, fine, style):
        if m is not None:
            m.eval()
    return coarse, fine, style


# ----------------------------------------------------------------------------------------
# Ray generation                                   tools/run_nerf_helpers.py:153-168
# ----------------------------------------------------------------------------------------
def get_rays(H: int, W: int, K, c2w: torch.Tensor):
    """Pinhole rays; row-major (H, W): ray index = row * W + col."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - float(K[0][2])) / float(K[0][0]),
                        -(j - float(K[1][2])) / float(K[1][1]),
                        -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def pose_spherical(phi: float, theta: float, radius: float) -> torch.Tensor:
    """tools/load_facescape.py:9-38."""
    def trans_t(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], np.float32)

    def rot_zyy(p):
        return np.array([[np.cos(p), 0, -np.sin(p), 0], [0, 1, 0, 0],
                         [np.sin(p), 0, np.cos(p), 0], [0, 0, 0, 1]], np.float32)

    def rot_phi(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0],
                         [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], np.float32)

    c2w = trans_t(radius)
    c2w = rot_phi(theta / 180.0 * np.pi) @ c2w
    c2w = rot_zyy(phi / 180.0 * np.pi) @ c2w
    return torch.tensor(c2w, dtype=torch.float32)


def make_ray_batch(rays_o, rays_d, near: float, far: float) -> torch.Tensor:
    """[N, 11] = o(3) d(3) near far viewdir(3)   (render_class.py:158-179, 393-415)."""
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    viewdirs = viewdirs.reshape(-1, 3).float()
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    nr = near * torch.ones_like(rays_d[..., :1])
    fr = far * torch.ones_like(rays_d[..., :1])
    return torch.cat([rays_o, rays_d, nr, fr, viewdirs], -1)


# ----------------------------------------------------------------------------------------
# Compositing                                          models/render_class.py:440-482
# ----------------------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, noise=None, white_bkgd=False):
    """raw [N,S,4] (pre-activation r,g,b,sigma) -> rgb_map, disp_map, acc_map, weights, depth_map.

    ``noise`` is the already-scaled additive sigma noise ([N,S]) or None (:461-468)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]                                        # :454
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)             # :455
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)                           # :458
    rgb = torch.sigmoid(raw[..., :3])                                                  # :460
    sig = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1.0 - torch.exp(-F.relu(sig) * dists)                                      # :453,470
    trans = torch.cumprod(
        torch.cat([torch.ones((alpha.shape[0], 1), dtype=alpha.dtype), 1.0 - alpha + 1e-10], -1),
        -1)[:, :-1]                                                                    # :471
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -2)                                  # :473
    depth_map = torch.sum(weights * z_vals, -1)                                        # :475
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map),
                               depth_map / torch.sum(weights, -1))                    # :476
    acc_map = torch.sum(weights, -1)                                                   # :477
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])                                 # :480
    return rgb_map, disp_map, acc_map, weights, depth_map


# ----------------------------------------------------------------------------------------
# Hierarchical sampling                              tools/run_nerf_helpers.py:203-247
# ----------------------------------------------------------------------------------------
def sample_pdf(bins, weights, N_samples, det=False, u=None):
    weights = weights + 1e-5                                                           # :205
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)                         # :208
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, steps=N_samples)                              # :212
            u = u.expand(list(cdf.shape[:-1]) + [N_samples])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [N_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                                      # :231
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    matched_shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(matched_shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(matched_shape), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)                   # :243
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])                      # :245


# ----------------------------------------------------------------------------------------
# Network query                                        models/render_class.py:69-109
# ----------------------------------------------------------------------------------------
def expression_mod(style: nn.Module, shape_codes: torch.Tensor, exp_code: torch.Tensor):
    """exp_scale * expCodes_Sigma[expType] + exp_bias    (render_class.py:75,80-81)."""
    scale, bias = style(shape_codes[0, :].reshape(1, -1))
    return scale * exp_code.reshape(1, -1) + bias


def run_network(pts, viewdirs, net: NeRF, shape_codes, exp_mod, tex_codes,
                multires=10, multires_views=4, netchunk: Optional[int] = 65536, forward_fn=None):
    """pts [N,S,3], viewdirs [N,3] -> raw [N,S,4].  forward_fn(net, emb, shp, emb_dirs, tex) replaces net(...)
    (used by the tests' reduced-precision emulation)."""
    flat = pts.reshape(-1, pts.shape[-1])
    P = flat.shape[0]
    emb = torch.cat([embed(flat, multires), exp_mod.reshape(1, -1).expand(P, -1)], -1)   # :77-83
    shp = shape_codes[0, :].expand(P, shape_codes.shape[-1])                             # :74
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)                            # :88-89
    emb_dirs = embed(dirs, multires_views)                                               # :90
    tex = tex_codes.reshape(1, -1).expand(P, -1)                                         # :104
    step = P if netchunk is None else netchunk
    call = net if forward_fn is None else (lambda *a: forward_fn(net, *a))
    out = torch.cat([call(emb[i:i + step], shp[i:i + step], emb_dirs[i:i + step], tex[i:i + step])
                     for i in range(0, P, step)], 0)                                     # :105
    return out.reshape(list(pts.shape[:-1]) + [out.shape[-1]])


# ----------------------------------------------------------------------------------------
# render_rays                                           models/render_class.py:239-352
# ----------------------------------------------------------------------------------------
def coarse_z_vals(near, far, N_samples, lindisp=False, perturb=0.0, t_rand=None):
    """:291-313.  near/far [N,1]."""
    t_vals = torch.linspace(0.0, 1.0, steps=N_samples)
    if not lindisp:
        z_vals = near * (1.0 - t_vals) + far * t_vals
    else:
        z_vals = 1.0 / (1.0 / near * (1.0 - t_vals) + 1.0 / far * t_vals)
    z_vals = z_vals.expand([near.shape[0], N_samples])
    if perturb > 0.0:
        mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(z_vals.shape)
        z_vals = lower + (upper - lower) * t_rand
    return z_vals


def render_rays(rays: torch.Tensor, net_coarse: NeRF, net_fine: Optional[NeRF], shape_codes, exp_mod,
                tex_codes, N_samples=64, N_importance=64, perturb=0.0, lindisp=False,
                white_bkgd=False, retraw=False, t_rand=None, u=None, noise_c=None, noise_f=None,
                multires=10, multires_views=4, netchunk=65536, run_fine=True,
                z_fine_override=None, forward_fn=None) -> Dict[str, torch.Tensor]:
    """rays [N,11] -> dict with the reference's keys (:338-345).  Differentiable (call under torch.no_grad()
    for inference).  z_fine_override: use these fine-pass depths instead of resampling (gradient tests compare
    two implementations at identical sample points; the reference detaches z_samples anyway, :326)."""
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    viewdirs = rays[:, 8:11]
    near, far = rays[:, 6:7], rays[:, 7:8]
    z_vals = coarse_z_vals(near, far, N_samples, lindisp, perturb, t_rand)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]              # :315
    raw = run_network(pts, viewdirs, net_coarse, shape_codes, exp_mod, tex_codes,
                      multires, multires_views, netchunk, forward_fn)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, noise_c, white_bkgd)
    ret = {}
    if N_importance > 0 and run_fine:
        rgb0, disp0, acc0 = rgb_map, disp_map, acc_map
        z_mid = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])                                 # :324
        z_samples = sample_pdf(z_mid, weights[..., 1:-1], N_importance, det=(perturb == 0.0), u=u)
        z_samples = z_samples.detach()                                                     # :326
        z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)                     # :328
        if z_fine_override is not None:
            z_vals = z_fine_override
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        run_fn = net_coarse if net_fine is None else net_fine                              # :332
        raw = run_network(pts, viewdirs, run_fn, shape_codes, exp_mod, tex_codes,
                          multires, multires_views, netchunk, forward_fn)
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, noise_f, white_bkgd)
        ret.update(rgb0=rgb0, disp0=disp0, acc0=acc0,
                   z_std=torch.std(z_samples, dim=-1, unbiased=False))                     # :345
        ret["z_vals_fine"] = z_vals            # extra (not in the reference dict): parity aid
    ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map)
    ret["weights"] = weights                   # extra: parity aid
    if retraw:
        ret["raw"] = raw
    return ret


def flops_per_ray(W_c=256, D_c=8, W_f=1024, D_f=10, S_c=64, S_f=128, in_ch=93, in_v=27, n_shape=50,
                  n_tex=256) -> float:
    """Algorithmic FLOPs (2*in*out over every nn.Linear as the reference executes them),
    SURVEY.md §8(d)."""
    def net(W, D):
        mac = in_ch * W + 3 * W * W                                   # xyzEncode
        for lat in (n_shape, n_tex):                                  # two skipMLPs
            mac += (lat + W) * W + 4 * W * W + (lat + 2 * W) * W + (D - 6) * W * W
        mac += W                                                      # alpha
        mac += (in_v + W) * (W // 2) + (W // 2) * 3                   # view + rgb
        return 2.0 * mac
    f = S_c * net(W_c, D_c)
    if W_f:
        f += S_f * net(W_f, D_f)
    return f


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)
