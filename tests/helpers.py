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
    for k in ("ray_index", "K", "c2w", "exp_table", "exp_slot", "uv_seed", "angle", "sigma_last", "sigma_abs_min"):   # round-2 frame crops / samples
        if k in z.files:
            inp[k] = torch.from_numpy(z[k]) if z[k].ndim else z[k].item()
    return meta, inp, out


FRAME_CROPS_CFG4 = ["cfg4_800_exp9", "cfg4_800_exp14", "cfg4_800_exp2"]
FRAME_CROPS_CFG5 = ["cfg5_800_id0", "cfg5_800_id1", "cfg5_800_id2"]
# 1024 rays spread over the 800x800 frame, unmodified reference: a dense field (the crops' nets) and bench.py's synthetic
# frame (seed-0 nets: an almost empty fine field, where the reference's last-interval step decides single rays)
FRAME_SAMPLES = ["frame_sample_1024", "frame_sample_bench_1024"]


def build_reference_like(seed, W_c=256, D_c=8, W_f=1024, D_f=10):
    """(coarse, fine, style, renderer) in the RNG order of oracle/ref_loader.build_reference — which follows
    tools/create_model_condition.py:16-50 — using the oracle's nets and the product's B200Renderer (whose constructor
    consumes the generator exactly like myRenderer.__init__, models/render_class.py:41-58: texture encoder, StyleModule,
    20 expression codes; oracle/make_golden.py asserts this against the reference before writing the cfg5 fixtures)."""
    from mofanerf_b200 import B200Renderer
    torch.manual_seed(seed)
    in_ch, in_v = O.embed_dim(10) + 30, O.embed_dim(4)
    coarse = O.NeRF(D_c, W_c, in_ch, in_v, 256, 50).eval()
    fine = O.NeRF(D_f, W_f, in_ch, in_v, 256, 50).eval()
    state = torch.random.get_rng_state()
    style = O.StyleModule(input_ch_bm=50, out_ch=30).eval()
    torch.random.set_rng_state(state)
    renderer = B200Renderer(expCodesLen=30)
    renderer.idSpecificMod.load_state_dict(style.state_dict())
    return coarse, fine, style, renderer.eval()


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
