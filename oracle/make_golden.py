"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The reference has no golden vectors of its own, so these fixtures — outputs of
the reference's own functions at fixed seeds — are what pins the oracle (and through it the
CUDA engine).  Nets are NOT stored (110 MB): they are rebuilt from the seed by
oracle.mofa_oracle.build_nets(); this script asserts that rebuild is bit-identical to the
reference's own construction.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import mofa_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _latents(seed):
    """Latents drawn from the reference's priors (configs/texShpDistribution.npy, consumed by
    tools/wild_fit_base.py:21-45); expression code ~ U[0,1) as render_class.py:53-56."""
    d = np.load(os.path.join(ref_loader.REF_ROOT, "configs", "texShpDistribution.npy"), allow_pickle=True).item()
    g = torch.Generator().manual_seed(seed)
    f = lambda k: torch.as_tensor(np.asarray(d[k]), dtype=torch.float32)
    shape = f("shape_mean").reshape(1, 50) + f("shape_std").reshape(1, 50) * torch.randn(1, 50, generator=g)
    tex = f("texture_mean").reshape(256) + f("texture_std").reshape(256) * torch.randn(256, generator=g)
    exp = torch.rand(1, 30, generator=g)
    return shape, tex, exp


def _check_same_nets(ref_net, ora_net):
    a, b = ref_net.state_dict(), ora_net.state_dict()
    assert list(a.keys()) == list(b.keys()), "state_dict key order differs"
    for k in a:
        assert torch.equal(a[k], b[k]), f"init differs at {k}"


def _camera(H, W, angle):
    focal = 1200.0 * H / 512.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], np.float64)
    c2w = O.pose_spherical(float(angle), 0.0, 16.0)
    return K, c2w


def render_case(name, seed, W_c, D_c, W_f, D_f, H, W, angle, N_samples, N_importance, crop=None,
                perturb=0.0, raw_noise_std=0.0, pytest=False, white_bkgd=False, sigma_bias=None,
                lindisp=False):
    ref = ref_loader.load()
    coarse, fine, renderer = ref_loader.build_reference(seed, W_c, D_c, W_f, D_f)
    oc, of, ostyle = O.build_nets(seed, W_c, D_c, W_f, D_f)
    _check_same_nets(coarse, oc)
    if fine is not None:
        _check_same_nets(fine, of)
    _check_same_nets(renderer.idSpecificMod, ostyle)
    if sigma_bias is not None:
        for n in (coarse, fine):
            if n is not None:
                n.alpha_linear[0].bias.data.fill_(sigma_bias)
    shape, tex, exp = _latents(seed + 100)
    K, c2w = _camera(H, W, angle)
    rays_o, rays_d = ref.helpers.get_rays(H, W, K, c2w[:3, :4])
    rays_o = rays_o.reshape(-1, 3)
    rays_d = rays_d.reshape(-1, 3)
    if crop is not None:  # deterministic subset of rays (keeps fixtures and CPU time small)
        idx = torch.linspace(0, H * W - 1, crop).long()
        rays_o, rays_d = rays_o[idx], rays_d[idx]
    kwargs = dict(network_fn=coarse, network_fine=fine, N_samples=N_samples, N_importance=N_importance,
                  perturb=perturb, raw_noise_std=raw_noise_std, white_bkgd=white_bkgd, lindisp=lindisp,
                  use_viewdirs=True, ndc=False, near=8.0, far=26.0, pytest=pytest, retraw=True)
    with torch.no_grad():
        rgb, disp, acc, extras = renderer.render_fitting(H, W, K, chunk=4096, rays=(rays_o, rays_d),
                                                         shapeCodes=shape, uvCodes=tex, expType=20,
                                                         expCodes=exp, **kwargs)
    out = dict(rgb_map=rgb, disp_map=disp, acc_map=acc)
    for k in ("rgb0", "disp0", "acc0", "z_std", "raw"):
        if k in extras:
            out[k] = extras[k]
    meta = dict(seed=seed, W_c=W_c, D_c=D_c, W_f=W_f, D_f=D_f, H=H, W=W, N_samples=N_samples,
                N_importance=N_importance, perturb=perturb, raw_noise_std=raw_noise_std,
                pytest=int(pytest), white_bkgd=int(white_bkgd), lindisp=int(lindisp),
                sigma_bias=(np.nan if sigma_bias is None else sigma_bias), near=8.0, far=26.0)
    arrays = {f"out_{k}": v.detach().numpy().astype(np.float32) for k, v in out.items()}
    arrays.update(rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), shape=shape.numpy(), tex=tex.numpy(),
                  exp=exp.numpy(), K=K, c2w=c2w.numpy())
    for k, v in meta.items():
        arrays[f"meta_{k}"] = np.asarray(v)
    if "out_raw" in arrays and arrays["out_raw"].size > 200_000:
        del arrays["out_raw"]
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
    nan = int(np.isnan(arrays["out_disp_map"]).sum())
    print(f"[golden] {name}: rays={rays_o.shape[0]} rgb mean={arrays['out_rgb_map'].mean():.4f} "
          f"acc mean={arrays['out_acc_map'].mean():.4f} disp NaNs={nan}")


def op_cases():
    """Op-level known answers from the reference's own functions."""
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(7)
    arrays = {}
    # positional encoding (models/model.py:15-63)
    x = (torch.rand(64, 3, generator=g) * 2 - 1) * 20.0
    e10, _ = ref.model.get_embedder(10, 0)
    e4, _ = ref.model.get_embedder(4, 0)
    arrays.update(pe_x=x.numpy(), pe_out10=e10(x).numpy(), pe_out4=e4(x).numpy())
    # raw2outputs (render_class.py:440-482): includes zero-density rays (NaN disp) and white bkgd
    N, S = 48, 64
    raw = torch.randn(N, S, 4, generator=g) * 2.0
    raw[:8, :, 3] = -5.0  # relu -> 0 density everywhere: acc == 0, disp = NaN
    z = torch.sort(torch.rand(N, S, generator=g) * 18 + 8, -1)[0]
    d = torch.randn(N, 3, generator=g)
    for wb in (0, 1):
        o = ref.render_class.raw2outputs(raw, z, d, 0, bool(wb))
        for nm, v in zip(("rgb", "disp", "acc", "weights", "depth"), o):
            arrays[f"r2o_wb{wb}_{nm}"] = v.numpy()
    arrays.update(r2o_raw=raw.numpy(), r2o_z=z.numpy(), r2o_d=d.numpy())
    # sample_pdf (run_nerf_helpers.py:203-247): det and explicit-u (pytest hook) variants
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    w = o[3][:, 1:-1]
    arrays["pdf_det"] = ref.helpers.sample_pdf(bins, w, 64, det=True).numpy()
    arrays["pdf_rand_pytest"] = ref.helpers.sample_pdf(bins, w, 64, det=False, pytest=True).numpy()
    np.random.seed(0)
    arrays["pdf_u"] = np.random.rand(N, 64).astype(np.float32)
    arrays.update(pdf_bins=bins.numpy(), pdf_w=w.numpy())
    # NeRF.forward (models/model.py:121-137) on a small seeded net
    torch.manual_seed(3)
    net = ref.model.NeRF(D=8, W=256, input_ch_shapeCodes=50, input_ch_textureCodes=256, input_ch=93,
                         output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True).eval()
    torch.manual_seed(3)
    onet = O.NeRF(8, 256, 93, 27, 256, 50).eval()
    _check_same_nets(net, onet)
    P = 32
    a = torch.randn(P, 93, generator=g)
    b = torch.randn(P, 50, generator=g) * 0.03
    c = torch.randn(P, 27, generator=g)
    t = torch.randn(P, 256, generator=g) * 0.3
    with torch.no_grad():
        arrays.update(net_in_pts=a.numpy(), net_in_shape=b.numpy(), net_in_views=c.numpy(),
                      net_in_tex=t.numpy(), net_out=net(a, b, c, t).numpy())
    # get_rays (run_nerf_helpers.py:153-168)
    K, c2w = _camera(6, 10, 30.0)
    ro, rd = ref.helpers.get_rays(6, 10, K, c2w[:3, :4])
    arrays.update(rays_K=K, rays_c2w=c2w.numpy(), rays_o=ro.numpy(), rays_d=rd.numpy())
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **arrays)
    print("[golden] ops.npz written")


def frame_crop_cases():
    """Round-2 fixtures: crops of the 800x800 frame at the real widths (W_f = 1024) for BASELINE configs #4 and #5.

    cfg4 (run_fit.py:394-403, rendering_modulation): render_fitting with expCodes = render.expCodes_Sigma[e] for three
    expression slots, same identity, target_pose = pose_spherical(0, 0, 16).
    cfg5 (render_refine_trainSet.py:245-304 -> render_path -> render, models/render_class.py:125-197): three identities,
    each with its own shape code, its own 512x512 UV map through the (seeded random-init) texture encoder, its own
    expression slot and view.
    The rays are a seeded random subset of the frame's 640 000 rays (indices stored: the engine's in-kernel ray
    generation must reproduce them from (c2w, K, H, W))."""
    ref = ref_loader.load()
    seed, W_c, D_c, W_f, D_f, H, W, n_crop = 5, 256, 8, 1024, 10, 800, 800, 192
    coarse, fine, renderer = ref_loader.build_reference(seed, W_c, D_c, W_f, D_f)
    oc, of, ostyle = O.build_nets(seed, W_c, D_c, W_f, D_f)
    _check_same_nets(coarse, oc)
    _check_same_nets(fine, of)
    _check_same_nets(renderer.idSpecificMod, ostyle)
    exp_table = torch.cat([c.detach().reshape(1, 30) for c in renderer.expCodes_Sigma[:20]], 0)
    # the GPU tests rebuild the texture encoder and the expression table from the seed (tests/helpers.py
    # build_reference_like): make sure that rebuild is bit-identical to the reference's own construction
    from tests.helpers import build_reference_like
    _, _, _, mine = build_reference_like(seed, W_c, D_c, W_f, D_f)
    _check_same_nets(renderer.texEncoder, mine.texEncoder)
    _check_same_nets(renderer.idSpecificMod, mine.idSpecificMod)
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(renderer.expCodes_Sigma[:20], mine.expCodes_Sigma[:20]))
    kwargs = dict(network_fn=coarse, network_fine=fine, N_samples=64, N_importance=64, perturb=0.0, raw_noise_std=0.0,
                  white_bkgd=False, lindisp=False, use_viewdirs=True, ndc=False, near=8.0, far=26.0, retraw=False)
    meta = dict(seed=seed, W_c=W_c, D_c=D_c, W_f=W_f, D_f=D_f, H=H, W=W, N_samples=64, N_importance=64, perturb=0.0,
                raw_noise_std=0.0, pytest=0, white_bkgd=0, lindisp=0, sigma_bias=np.nan, near=8.0, far=26.0)

    def save(name, out, extras, rays_o, rays_d, idx, K, c2w, shape, tex, exp, **more):
        res = dict(rgb_map=out[0], disp_map=out[1], acc_map=out[2])
        for k in ("rgb0", "disp0", "acc0", "z_std"):
            res[k] = extras[k]
        arrays = {f"out_{k}": v.detach().numpy().astype(np.float32) for k, v in res.items()}
        arrays.update(rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), ray_index=idx.numpy(), K=K, c2w=c2w.numpy(),
                      shape=shape.numpy(), tex=tex.detach().numpy(), exp=exp.detach().numpy(), exp_table=exp_table.numpy())
        for k, v in more.items():
            arrays[k] = np.asarray(v)
        for k, v in meta.items():
            arrays[f"meta_{k}"] = np.asarray(v)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
        print(f"[golden] {name}: rays={rays_o.shape[0]} rgb mean={arrays['out_rgb_map'].mean():.4f} "
              f"acc mean={arrays['out_acc_map'].mean():.4f}")

    # ---- config #4: expression sweep (three of the slots run_fit.py:394 renders)
    shape, tex, exp_unused = _latents(seed + 100)
    K, c2w = _camera(H, W, 0.0)
    ro, rd = ref.helpers.get_rays(H, W, K, c2w[:3, :4])
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    for e in (9, 14, 2):
        idx = torch.randperm(H * W, generator=torch.Generator().manual_seed(1000 + e))[:n_crop].sort()[0]
        with torch.no_grad():
            out = renderer.render_fitting(H, W, K, chunk=4096, rays=(ro[idx], rd[idx]), shapeCodes=shape, uvCodes=tex,
                                          expType=20, expCodes=renderer.expCodes_Sigma[e], **kwargs)
        save(f"cfg4_800_exp{e}", out[:3], out[3], ro[idx], rd[idx], idx, K, c2w, shape, tex,
             renderer.expCodes_Sigma[e], exp_slot=e)

    # ---- config #5: identities through render() with a UV map -> texEncoder
    for i, (angle, e) in enumerate(((-35.0, 3), (0.0, 11), (50.0, 17))):
        shape_i, _, _ = _latents(seed + 200 + i)
        uv = torch.rand(512, 512, 3, generator=torch.Generator().manual_seed(300 + i))
        K, c2w = _camera(H, W, angle)
        ro, rd = ref.helpers.get_rays(H, W, K, c2w[:3, :4])
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        idx = torch.randperm(H * W, generator=torch.Generator().manual_seed(2000 + i))[:n_crop].sort()[0]
        with torch.no_grad():
            out = renderer.render(H, W, K, chunk=4096, rays=(ro[idx], rd[idx]), shapeCodes=shape_i, uvMap=uv, expType=e,
                                  **kwargs)
            tex_i = renderer.decoding_texCodes.reshape(-1)
        save(f"cfg5_800_id{i}", out[:3], out[3], ro[idx], rd[idx], idx, K, c2w, shape_i, tex_i,
             renderer.expCodes_Sigma[e], exp_slot=e, uv_seed=300 + i, angle=angle)


def frame_sample_case(n_rays=1024, bench_regime=False):
    """Round-2 (late) fixture: a LARGE sample of the 800x800 frame at the real widths — every 625th ray — through the
    unmodified reference's render_fitting (config #4 set-up, expression slot 9).  The 192-ray crops never met the
    reference's own last-interval discontinuity (raw2outputs gives the last sample an interval of 1e10,
    models/render_class.py:449: alpha_last is a step in sigma_last); a sample of this size does, so the fixture also
    keeps the reference's pre-activation sigma of every ray's last fine sample (`sigma_last`) and the smallest
    |sigma| along the ray."""
    ref = ref_loader.load()
    seed, W_c, D_c, W_f, D_f, H, W = (0 if bench_regime else 5), 256, 8, 1024, 10, 800, 800
    coarse, fine, renderer = ref_loader.build_reference(seed, W_c, D_c, W_f, D_f)
    oc, of, ostyle = O.build_nets(seed, W_c, D_c, W_f, D_f)
    _check_same_nets(coarse, oc)
    _check_same_nets(fine, of)
    _check_same_nets(renderer.idSpecificMod, ostyle)
    e = 9
    if bench_regime:
        # bench.py's synthetic frame (seed-0 nets, bench.synth_inputs latents and camera): the random-init fine net of
        # this seed leaves the volume almost empty (median fine acc 0.03) — the regime in which the last-interval step
        # and near-zero sigma decide single rays
        import bench
        shape, tex, exp_b, _, _ = bench.synth_inputs(H, W)
        renderer.expCodes_Sigma[e] = exp_b
        K, c2w = _camera(H, W, 30.0)
    else:
        shape, tex, _ = _latents(seed + 100)
        K, c2w = _camera(H, W, 0.0)
    ro, rd = ref.helpers.get_rays(H, W, K, c2w[:3, :4])
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    idx = torch.linspace(0, H * W - 1, n_rays).long()
    kwargs = dict(network_fn=coarse, network_fine=fine, N_samples=64, N_importance=64, perturb=0.0, raw_noise_std=0.0,
                  white_bkgd=False, lindisp=False, use_viewdirs=True, ndc=False, near=8.0, far=26.0, retraw=True)
    with torch.no_grad():
        out = renderer.render_fitting(H, W, K, chunk=4096, rays=(ro[idx], rd[idx]), shapeCodes=shape, uvCodes=tex,
                                      expType=20, expCodes=renderer.expCodes_Sigma[e], **kwargs)
    res = dict(rgb_map=out[0], disp_map=out[1], acc_map=out[2])
    for k in ("rgb0", "disp0", "acc0", "z_std"):
        res[k] = out[3][k]
    raw = out[3]["raw"].reshape(n_rays, -1, 4)
    arrays = {f"out_{k}": v.detach().numpy().astype(np.float32) for k, v in res.items()}
    arrays.update(sigma_last=raw[:, -1, 3].numpy().astype(np.float32),
                  sigma_abs_min=raw[..., 3].abs().min(dim=1).values.numpy().astype(np.float32),
                  rays_o=ro[idx].numpy(), rays_d=rd[idx].numpy(), ray_index=idx.numpy(), K=K, c2w=c2w.numpy(),
                  shape=shape.numpy(), tex=tex.numpy(), exp=renderer.expCodes_Sigma[e].detach().numpy(), exp_slot=np.asarray(e))
    meta = dict(seed=seed, W_c=W_c, D_c=D_c, W_f=W_f, D_f=D_f, H=H, W=W, N_samples=64, N_importance=64, perturb=0.0,
                raw_noise_std=0.0, pytest=0, white_bkgd=0, lindisp=0, sigma_bias=np.nan, near=8.0, far=26.0)
    for k, v in meta.items():
        arrays[f"meta_{k}"] = np.asarray(v)
    name = f"frame_sample_{'bench_' if bench_regime else ''}{n_rays}"
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
    print(f"[golden] {name}: rgb mean={arrays['out_rgb_map'].mean():.4f} acc median="
          f"{np.median(arrays['out_acc_map']):.4f} min|sigma_last|={np.abs(arrays['sigma_last']).min():.2e}")


def main():
    torch.set_num_threads(os.cpu_count())
    if "--r02" in sys.argv:      # only the round-2 frame crops (the round-1 fixtures are unchanged)
        frame_crop_cases()
        return
    if "--frame-sample" in sys.argv:
        frame_sample_case()
        frame_sample_case(bench_regime=True)
        return
    op_cases()
    # config #1 (plumbing): 64x64, 32 samples, coarse only
    render_case("cfg1_64x64_s32", 0, 256, 8, 1024, 10, 64, 64, 0.0, 32, 0)
    # small two-pass case with a narrow fine net (fast on CPU)
    render_case("small_w256", 1, 256, 8, 256, 10, 20, 20, -30.0, 64, 64, crop=160)
    # the real configuration (coarse 256x8 + fine 1024x10), 64+128 evaluations per ray
    render_case("full_w1024", 5, 256, 8, 1024, 10, 400, 400, 60.0, 64, 64, crop=48)
    # pytest-hook mode: seeded stratified jitter, random u, sigma noise
    render_case("perturb_pytest", 2, 256, 8, 256, 10, 16, 16, 0.0, 64, 64, crop=64, perturb=1.0,
                raw_noise_std=1.0, pytest=True)
    # empty space: negative sigma bias => acc ~ 0, NaN disparity; white background; lindisp
    render_case("empty_white", 4, 256, 8, 256, 10, 16, 16, 20.0, 64, 64, crop=64, white_bkgd=True,
                sigma_bias=-30.0, lindisp=True)
    frame_crop_cases()


if __name__ == "__main__":
    main()
