"""GPU tests of the reference-facing surface beyond render_fitting: render() (texture-encoder path),
render_path() (PNG writing, skip-if-exists), latent sweeps (BASELINE configs #4/#5 in miniature), rank-sharded
rendering on one GPU ("virtual ranks"), and BASELINE config #2 (400x400 FULL frame) through size-independent
properties."""
import os

import numpy as np
import pytest
import torch

from oracle import mofa_oracle as O
from tests import parity_log
from tests.helpers import build_case_nets, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _renderer(style):
    from mofanerf_b200 import B200Renderer
    torch.manual_seed(123)
    r = B200Renderer(expCodesLen=30).to(DEV)
    r.idSpecificMod.load_state_dict(style.state_dict())
    return r


def _kw(c, f, **over):
    kw = dict(near=8.0, far=26.0, use_viewdirs=True, ndc=False, network_fn=c.to(DEV), network_fine=f.to(DEV),
              N_samples=64, N_importance=64, perturb=0.0, raw_noise_std=0.0)
    kw.update(over)
    return kw


def _cam(H, W, angle=10.0):
    focal = 1200.0 * H / 512.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    return K, O.pose_spherical(angle, 0.0, 16.0)


def test_render_with_uvmap_matches_oracle():
    """myRenderer.render (models/render_class.py:125-197): c2w ray generation + texEncoder(uvMap) + expType."""
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    r = _renderer(s)
    H, W = 6, 5
    K, c2w = _cam(H, W)
    uv = torch.rand(512, 512, 3, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        # oracle side: same texture encoder weights, same expression slot
        tex, _ = r.texEncoder(uv.permute(2, 0, 1).unsqueeze(0).to(DEV))
        exp = r.expCodes_Sigma[7].detach().cpu()
        ro, rd = O.get_rays(H, W, K, c2w[:3, :4])
        rays = O.make_ray_batch(ro, rd, 8.0, 26.0)
        ref = O.render_rays(rays, c, f, inp["shape"], O.expression_mod(s, inp["shape"], exp), tex.cpu().reshape(-1))
        rgb, disp, acc, extras = r.render(H, W, K, chunk=1 << 20, c2w=c2w[:3, :4].to(DEV), shapeCodes=inp["shape"].to(DEV),
                                          uvMap=uv.to(DEV), expType=7, retraw=True, **_kw(c, f))
    assert rgb.shape == (H, W, 3) and disp.shape == (H, W) and acc.shape == (H, W)
    assert extras["raw"].shape == (H, W, 128, 4) and extras["rgb0"].shape == (H, W, 3) and extras["losses"] == 0
    d = (rgb.cpu().reshape(-1, 3) - ref["rgb_map"]).abs().max().item()
    d0 = (extras["rgb0"].cpu().reshape(-1, 3) - ref["rgb0"]).abs().max().item()
    print(f"[parity] render(): rgb {d:.2e} rgb0 {d0:.2e}")
    parity_log.record("render()[c2w + texEncoder, 6x5]", rgb_map_max=d, rgb0_max=d0)
    assert d <= 3e-2 and d0 <= 5e-5


def test_render_path_writes_png_and_skips_existing(tmp_path):
    """myRenderer.render_path (models/render_class.py:199-237)."""
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    r = _renderer(s)
    H, W = 4, 4
    K, c2w = _cam(H, W)
    poses = torch.stack([c2w, O.pose_spherical(-20.0, 0.0, 16.0)]).to(DEV)
    uv = torch.rand(2, 512, 512, 3, generator=torch.Generator().manual_seed(5)).to(DEV)
    shapes = inp["shape"].repeat(2, 1).to(DEV)
    kw = _kw(c, f)
    with torch.no_grad():
        rgbs, disps = r.render_path(poses, [H, W, 1200.0 * H / 512], K, 1 << 20, kw, uvMap=uv, expType=[3, 11],
                                    savedir=str(tmp_path), shapeCodes=shapes, name=None)
    assert rgbs.shape == (2, H, W, 3) and disps.shape == (2, H, W)
    assert os.path.exists(tmp_path / "000.png") and os.path.exists(tmp_path / "001.png")
    with torch.no_grad():   # named output that already exists -> (0, 0), nothing rendered (:212-216)
        (tmp_path / "done.png").write_bytes(b"x")
        out = r.render_path(poses[:1], [H, W, 1.0], K, 1 << 20, kw, uvMap=uv, expType=[3], savedir=str(tmp_path),
                            shapeCodes=shapes, name="done")
    assert out == (0, 0)


def test_latent_sweep_refolds_per_call():
    """Configs #4/#5 in miniature: consecutive calls with different expression / shape / texture codes must each
    match the oracle (the latent fold is per call) and differ from one another."""
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    r = _renderer(s)
    rays_o, rays_d = inp["rays_o"][:24], inp["rays_d"][:24]
    rays = O.make_ray_batch(rays_o, rays_d, 8.0, 26.0)
    g = torch.Generator().manual_seed(21)
    cases, outs = [], []
    for i in range(3):      # oracle first (CPU), then the engine (the same modules move to the GPU)
        shape = inp["shape"] + 0.02 * torch.randn(1, 50, generator=g) * (i > 0)
        tex = inp["tex"] + 0.1 * torch.randn(256, generator=g) * (i > 1)
        exp = torch.rand(1, 30, generator=g)
        with torch.no_grad():
            ref = O.render_rays(rays, c, f, shape, O.expression_mod(s, shape, exp), tex)
        cases.append((shape, tex, exp, ref))
    kw = _kw(c, f)
    for shape, tex, exp, ref in cases:
        with torch.no_grad():
            rgb, _, _, ex = r.render_fitting(1, 24, None, rays=(rays_o.to(DEV), rays_d.to(DEV)), shapeCodes=shape.to(DEV),
                                             uvCodes=tex.to(DEV), expType=20, expCodes=exp.to(DEV), **kw)
        assert (ex["rgb0"].cpu() - ref["rgb0"]).abs().max().item() <= 5e-5      # coarse maps: split precision, fp32-class
        d = (rgb.cpu() - ref["rgb_map"]).abs()
        assert d.max().item() <= 3e-2 and d.mean().item() <= 1e-3, f"mean {d.mean().item():.2e} max {d.max().item():.2e}"
        outs.append(rgb.cpu())
    assert (outs[0] - outs[1]).abs().max().item() > 1e-3 and (outs[1] - outs[2]).abs().max().item() > 1e-3


def test_virtual_rank_sharding_is_exact():
    """rendering each rank's contiguous ray range separately and concatenating equals the
    single-GPU result bit-for-bit, for R = 2, 4, 8."""
    from mofanerf_b200.distributed import shard_range
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    r = _renderer(s)
    n = inp["rays_o"].shape[0]
    args = dict(shapeCodes=inp["shape"].to(DEV), uvCodes=inp["tex"].to(DEV), expType=20, expCodes=inp["exp"].to(DEV))
    with torch.no_grad():
        full = r.render_fitting(1, n, None, rays=(inp["rays_o"].to(DEV), inp["rays_d"].to(DEV)), **args, **_kw(c, f))[0]
        for R in (2, 4, 8):
            parts = []
            for k in range(R):
                lo, hi = shard_range(n, k, R)
                parts.append(r.render_fitting(1, hi - lo, None, rays=(inp["rays_o"][lo:hi].to(DEV),
                                                                     inp["rays_d"][lo:hi].to(DEV)), **args, **_kw(c, f))[0])
            assert torch.equal(torch.cat(parts, 0), full), f"R={R}"


def test_config2_400x400_full_frame_properties():
    """BASELINE config #2 (400x400, 64+128 samples, one identity, full-width nets) at full size, via
    size-independent properties: finite colours in [0,1], acc in [0,1+eps], sorted fine depths, row-major ray order
    (a vertically flipped camera image equals the flipped render of the rays), and a 97-ray subsample that matches
    a stand-alone render of just those rays bit-for-bit."""
    meta, inp, _ = load_case("full_w1024")
    c, f, s = build_case_nets(meta)
    r = _renderer(s)
    H = W = 400
    K, c2w = _cam(H, W, 60.0)
    args = dict(shapeCodes=inp["shape"].to(DEV), uvCodes=inp["tex"].to(DEV), expType=20, expCodes=inp["exp"].to(DEV))
    with torch.no_grad():
        rgb, disp, acc, ex = r.render_fitting(H, W, K, chunk=1 << 20, c2w=c2w[:3, :4].to(DEV), **args, **_kw(c, f))
        assert rgb.shape == (H, W, 3)
        assert bool(torch.isfinite(rgb).all()) and float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5
        assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-4
        assert bool(torch.isfinite(ex["z_std"]).all())
        from mofanerf_b200.rays import get_rays
        ro, rd = get_rays(H, W, K, c2w[:3, :4].to(DEV))
        idx = torch.linspace(0, H * W - 1, 97).long().to(DEV)
        # the frame's rays come from the engine's generate_rays kernel; render the SAME ray rows on their own
        eng = r.engine(DEV)
        all_rays = eng.generate_rays(H, W, K, c2w[:3, :4], 8.0, 26.0)
        rd_cpu = get_rays(H, W, K, c2w[:3, :4])[1]       # torch on the CPU (the fixtures' arithmetic), not torch's CUDA reduction
        assert torch.equal(all_rays[:, 3:6].cpu(), rd_cpu.reshape(-1, 3)), "generate_rays vs get_rays"
        sub = eng.render_rays(all_rays[idx], 64, 64)["rgb_map"]
        assert torch.equal(sub, rgb.reshape(-1, 3)[idx]), "row-major ray index / per-ray independence"
        ref_rays = O.make_ray_batch(ro.reshape(-1, 3)[idx[:12]].cpu(), rd.reshape(-1, 3)[idx[:12]].cpu(), 8.0, 26.0)
        ref = O.render_rays(ref_rays, c.cpu(), f.cpu(), inp["shape"], O.expression_mod(s, inp["shape"], inp["exp"]), inp["tex"])
        d12 = (sub[:12].cpu() - ref["rgb_map"]).abs()
        parity_log.record("config2 400x400 frame, 12 rays vs oracle", rgb_map_max=d12.max().item(), rgb_map_mean=d12.mean().item())
        assert d12.max().item() <= 3e-2 and d12.mean().item() <= 1e-3


@pytest.mark.parametrize("n,S,Ni,wb,lindisp", [(1, 64, 64, False, False), (3, 32, 64, True, False), (130, 16, 16, False, True),
                                               (5, 128, 128, False, False), (2, 2, 0, False, False), (7, 256, 0, True, False)])
def test_edge_sizes_match_oracle(n, S, Ni, wb, lindisp):
    """Ragged / extreme sizes: single ray, point counts that are not multiples of the 128-row tile, S_f = 96 / 256, the
    smallest and largest sample counts, coarse-only passes, white background and inverse-depth sampling.  Coarse maps are
    compared tightly (no resampling feedback); final maps at the end-to-end bound."""
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    ro, rd = inp["rays_o"][:n], inp["rays_d"][:n]
    rays = O.make_ray_batch(ro, rd, 8.0, 26.0)
    with torch.no_grad():
        ref = O.render_rays(rays, c, f, inp["shape"], O.expression_mod(s, inp["shape"], inp["exp"]), inp["tex"],
                            N_samples=S, N_importance=Ni, white_bkgd=wb, lindisp=lindisp)
    r = _renderer(s)
    with torch.no_grad():
        rgb, disp, acc, ex = r.render_fitting(1, n, None, rays=(ro.to(DEV), rd.to(DEV)), shapeCodes=inp["shape"].to(DEV),
                                              uvCodes=inp["tex"].to(DEV), expType=20, expCodes=inp["exp"].to(DEV),
                                              **_kw(c, f, N_samples=S, N_importance=Ni, white_bkgd=wb, lindisp=lindisp))
    assert rgb.shape == (1, n, 3) or rgb.shape == (n, 3)
    rgb = rgb.reshape(n, 3).cpu()
    if Ni > 0:
        d0 = (ex["rgb0"].reshape(n, 3).cpu() - ref["rgb0"]).abs().max().item()
        assert d0 <= 5e-5, f"rgb0 {d0:.2e}"
        d = (rgb - ref["rgb_map"]).abs().flatten()
        parity_log.record(f"edge n={n} S={S} Ni={Ni}", rgb0_max=d0, rgb_map_max=d.max().item(), rgb_map_mean=d.mean().item())
        # 16 + 16 samples in inverse depth: sample_pdf's `denom < 1e-5 -> 1` guard (run_nerf_helpers.py:243) sits exactly
        # on the pdf of an empty bin ((0 + 1e-5) / sum with sum ~ 1), so a 1e-7 change of the cdf flips it and moves one of
        # only 32 samples by a whole (huge, inverse-depth) bin: a cliff of the reference algorithm itself (measured: 2 of
        # 130 rays, everything else < 1e-4).  There the bound is the mean and the 90th percentile; everywhere else the
        # per-ray maximum.
        q90 = torch.quantile(d, 0.9).item()
        assert d.mean().item() <= 5e-3 and q90 <= 1e-3, f"mean {d.mean().item():.2e} q90 {q90:.2e} max {d.max().item():.2e}"
        if S >= 32:
            assert d.max().item() <= 3e-2, f"max {d.max().item():.2e}"
    else:
        assert "rgb0" not in ex
        dm = (rgb - ref["rgb_map"]).abs().max().item()
        parity_log.record(f"edge n={n} S={S} Ni={Ni}", rgb_map_max=dm)
        assert dm <= 5e-5
        assert (acc.reshape(n).cpu() - ref["acc_map"]).abs().max().item() <= 5e-5


def test_invalid_arguments_fail_loudly():
    from mofanerf_b200 import get_engine
    eng = get_engine(DEV)
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    eng.load_network(0, c.to(DEV), force=True)
    eng.load_network(1, f.to(DEV), force=True)
    eng.set_latents(inp["shape"], inp["exp"] * 0, inp["tex"])
    rays = torch.zeros(4, 11, device=DEV)
    with pytest.raises(RuntimeError, match="n_samples"):
        eng.render_rays(rays, 1, 0)                       # the reference degenerates at one sample; refused
    with pytest.raises(RuntimeError, match="out of range"):
        eng.render_rays(rays, 200, 100)                   # S_c + N_i > 256
    with pytest.raises(ValueError):
        eng.set_latents(torch.zeros(49), torch.zeros(30), torch.zeros(256))
    with pytest.raises(ValueError, match="engine on"):
        eng.render_rays(torch.zeros(4, 11), 8, 0)         # host tensor handed to the device entry point
