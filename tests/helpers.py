"""Shared helpers for the parity tests (oracle side).  Test infrastructure only."""
import os

import numpy as np
import torch

from oracle import mofa_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith("meta_")}
    out = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out_")}
    inp = {k: torch.from_numpy(z[k]) for k in ("rays_o", "rays_d", "shape", "tex", "exp")}
    return meta, inp, out


def build_case_nets(meta):
    c, f, s = O.build_nets(int(meta["seed"]), int(meta["W_c"]), int(meta["D_c"]), int(meta["W_f"]),
                           int(meta["D_f"]))
    if not np.isnan(meta["sigma_bias"]):
        for n in (c, f):
            if n is not None:
                n.alpha_linear[0].bias.data.fill_(float(meta["sigma_bias"]))
    return c, f, s


def case_randoms(meta, n_rays):
    """The reference's ``pytest=True`` hook draws (render_class.py:308-311,465-468;
    run_nerf_helpers.py:218-226): np.random.seed(0) before every draw."""
    S, Ni = int(meta["N_samples"]), int(meta["N_importance"])
    rnd = dict(t_rand=None, u=None, noise_c=None, noise_f=None)
    if meta["pytest"]:
        if meta["perturb"] > 0:
            np.random.seed(0)
            rnd["t_rand"] = torch.Tensor(np.random.rand(n_rays, S))
            np.random.seed(0)
            rnd["u"] = torch.Tensor(np.random.rand(n_rays, Ni))
        if meta["raw_noise_std"] > 0:
            np.random.seed(0)
            rnd["noise_c"] = torch.Tensor(np.random.rand(n_rays, S) * meta["raw_noise_std"])
            np.random.seed(0)
            rnd["noise_f"] = torch.Tensor(np.random.rand(n_rays, S + Ni) * meta["raw_noise_std"])
    return rnd


def oracle_render(meta, inp, nets=None, **over):
    c, f, s = nets if nets is not None else build_case_nets(meta)
    rays = O.make_ray_batch(inp["rays_o"], inp["rays_d"], float(meta["near"]), float(meta["far"]))
    em = O.expression_mod(s, inp["shape"], inp["exp"])
    rnd = case_randoms(meta, rays.shape[0])
    kw = dict(N_samples=int(meta["N_samples"]), N_importance=int(meta["N_importance"]),
              perturb=float(meta["perturb"]), lindisp=bool(meta["lindisp"]),
              white_bkgd=bool(meta["white_bkgd"]), retraw=True, **rnd)
    kw.update(over)
    with torch.no_grad():
        return O.render_rays(rays, c, f, inp["shape"], em, inp["tex"], **kw), rays, em


def assert_close_nan(a, b, atol, rtol=0.0, what=""):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    na, nb = torch.isnan(a), torch.isnan(b)
    assert torch.equal(na, nb), f"{what}: NaN masks differ ({int(na.sum())} vs {int(nb.sum())})"
    d = (a[~na] - b[~nb]).abs()
    lim = atol + rtol * b[~nb].abs()
    assert bool((d <= lim).all()), f"{what}: max|d|={d.max().item():.3e} (atol {atol}, rtol {rtol})"
