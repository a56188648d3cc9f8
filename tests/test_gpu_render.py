"""GPU parity, end to end: B200Renderer (-> C ABI -> sm_100a kernels) against
  (1) the fixtures produced by the unmodified reference (tests/golden/*.npz) and
  (2) the oracle run on the same inputs.

Stated tolerance (north_star: "within a stated fp32 tolerance"; SURVEY.md §7 / BASELINE.md §4.4 ask for
max |Δrgb| <= 3e-2, mean <= 3e-3, PSNR >= 45 dB): the coarse net runs in split precision (fp16 hi+lo operands, three
tensor-core products per layer, fp32 accumulation: fp32-class), the fine net's dense layers use single fp16 operands
with fp32 accumulation, every other stage is fp32.  For rendered maps in [0,1]:
  * coarse maps (rgb0 / acc0): max <= 5e-5 (measured 2e-6 .. 5e-6: the reference's own fp32 noise floor);
  * stage-wise, on identical sample points (test_teacher_forced_stages): per-point fine-net outputs within
    2e-2 * max(1, |raw|max), composited rgb within 2e-3 (measured 5e-4) — the precision of the fp16 dense chain;
  * end to end: max |Δrgb| <= 3e-2, mean |Δrgb| <= 1e-3, PSNR(engine, reference) >= 50 dB (measured over all fixtures:
    max <= 1.5e-2, mean <= 2.9e-4, PSNR 58 .. 112 dB; table in profiles/parity_r02.json), acc within 3e-2; disparity
    where acc > 1e-3 within 5e-2 relative; NaN positions coincide where the reference's acc is exactly 0.  The per-ray
    maximum is looser than the mean because inverse-CDF resampling is discontinuous: an fp32-rounding-level change of a
    coarse weight can move one fine sample across a bin of a random-init field with 2^9 positional frequencies (the
    reference shows the same sensitivity between its own CPU and CUDA runs).  The SIMT verification kernel (single fp16
    operands, different accumulation order) is held to the same end-to-end bounds and agrees with the default path to
    2e-3 on the coarse maps (its own fp16 error).
  * on a LARGE sample of the bench frame (test_frame_sample_error_distribution_with_the_references_own_yardsticks, 4096
    rays) the bound is a distribution, not one maximum: the reference's own last-interval step (relu(sigma) * 1e10,
    models/render_class.py:449) flips the opacity of ~3 in 8192 rays of the synthetic frame for ANY finite-precision
    arithmetic — the reference's default TF32 mode flips the same rays — so the test asserts p50 / p99 / the fraction of
    rays within 3e-2, that every flip sits on |sigma_last| < 2e-2 in the reference, and that the engine is closer to the
    fp32 reference than the reference run with TF32 matmuls is (measured 3x closer).
Ray order is bit-exact by construction and tested (output row i <-> input ray i).
"""
import numpy as np
import pytest
import torch

from oracle import mofa_oracle as O
from tests import parity_log
from tests.helpers import (FRAME_CROPS_CFG4, FRAME_CROPS_CFG5, FRAME_SAMPLES, build_case_nets, build_reference_like, case_randoms, load_case,
                           oracle_render)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def engine_render(meta, inp, nets, gemm_simt=False, chunk=1024 * 32, want_aux=False, engine_chunk=None):
    from mofanerf_b200 import B200Renderer
    c, f, s = nets
    r = B200Renderer(expCodesLen=30).to(DEV)
    r.idSpecificMod.load_state_dict(s.state_dict())
    if engine_chunk is not None:
        r.engine(DEV).chunk_rays = engine_chunk
    n = inp["rays_o"].shape[0]
    rnd = case_randoms(meta, n)
    kw = dict(network_fn=c.to(DEV), network_fine=None if f is None else f.to(DEV), N_samples=int(meta["N_samples"]),
              N_importance=int(meta["N_importance"]), perturb=float(meta["perturb"]),
              raw_noise_std=float(meta["raw_noise_std"]), white_bkgd=bool(meta["white_bkgd"]),
              lindisp=bool(meta["lindisp"]), retraw=True, pytest=bool(meta["pytest"]), gemm_simt=gemm_simt,
              want_aux=want_aux)
    with torch.no_grad():
        rgb, disp, acc, extras = r.render_fitting(
            int(meta["H"]), int(meta["W"]), None, chunk=chunk, rays=(inp["rays_o"].to(DEV), inp["rays_d"].to(DEV)),
            shapeCodes=inp["shape"].to(DEV), uvCodes=inp["tex"].to(DEV), expType=20, expCodes=inp["exp"].to(DEV),
            near=float(meta["near"]), far=float(meta["far"]), use_viewdirs=True, ndc=False, **kw)
    torch.cuda.synchronize()
    r.engine(DEV).chunk_rays = 0
    out = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, **{k: v for k, v in extras.items() if torch.is_tensor(v)})
    return {k: v.float().cpu() for k, v in out.items()}


def check_maps(name, got, ref, max_rgb=3e-2, mean_rgb=1e-3, min_psnr=50.0, max_acc=3e-2, disp_min_acc=1e-3,
               max_rgb0=5e-5):
    msgs = []
    for k in ("rgb_map", "rgb0"):   # measured numbers go to profiles/parity_r02.json before anything is asserted
        if k in ref:
            d = (got[k] - ref[k]).abs()
            parity_log.record(name, **{f"{k}_max": d.max().item(), f"{k}_mean": d.mean().item(),
                                       f"{k}_psnr_db": O.psnr(got[k], ref[k]), "rays": ref[k].shape[0]})
    for k in ("acc_map", "acc0"):
        if k in ref:
            parity_log.record(name, **{f"{k}_max": (got[k] - ref[k]).abs().max().item()})
    for k in ("rgb_map", "rgb0"):
        if k in ref:
            d = (got[k] - ref[k]).abs()
            ps = O.psnr(got[k], ref[k])
            msgs.append(f"{k}: max {d.max().item():.2e} mean {d.mean().item():.2e} psnr {ps:.1f}")
            assert d.max().item() <= max_rgb and d.mean().item() <= mean_rgb and ps >= min_psnr, f"{name} {msgs[-1]}"
            if k == "rgb0" and max_rgb0 is not None:      # the coarse pass is fp32-class
                assert d.max().item() <= max_rgb0, f"{name} {msgs[-1]}"
    for k in ("acc_map", "acc0"):
        if k in ref:
            d = (got[k] - ref[k]).abs().max().item()
            msgs.append(f"{k}: max {d:.2e}")
            assert d <= max_acc, f"{name} {msgs[-1]}"
    for k, ka in (("disp_map", "acc_map"), ("disp0", "acc0")):
        if k in ref:
            empty = ref[ka] == 0
            # NaN disparity (0/0) wherever BOTH see no density at all; a ray the reference renders as exactly empty may
            # pick up a ~1e-6 weight in the fp16 chain (sigma + noise crossing zero): then acc must stay negligible
            assert bool(torch.isnan(got[k][empty & (got[ka] == 0)]).all()), f"{name} {k}: NaN expected where acc == 0"
            assert bool((got[ka][empty] <= 1e-3).all()), f"{name} {ka}: density where the reference has none"
            solid = (ref[ka] > disp_min_acc) & (got[ka] > disp_min_acc)
            if solid.any():
                rel = ((got[k][solid] - ref[k][solid]).abs() / ref[k][solid].abs().clamp_min(1e-6)).max().item()
                msgs.append(f"{k}: max rel {rel:.2e}")
                assert rel <= 5e-2, f"{name} {msgs[-1]}"
    if "z_std" in ref:
        d = (got["z_std"] - ref["z_std"]).abs().max().item()
        msgs.append(f"z_std: max {d:.2e}")
        assert d <= 0.25, f"{name} {msgs[-1]}"     # one coarse bin (18/63) — resampling moves with the coarse weights
    print(f"[parity] {name}: " + "; ".join(msgs))


@pytest.mark.parametrize("name", ["small_w256", "full_w1024", "perturb_pytest", "empty_white", "cfg1_64x64_s32"])
def test_against_reference_fixtures(name):
    meta, inp, gold = load_case(name)
    got = engine_render(meta, inp, build_case_nets(meta))
    check_maps(name, got, gold)     # one bound for every fixture, the seeded-random (pytest=True) case included


def _crop_kwargs(meta, c, f):
    return dict(network_fn=c.to(DEV), network_fine=f.to(DEV), N_samples=int(meta["N_samples"]),
                N_importance=int(meta["N_importance"]), perturb=0.0, raw_noise_std=0.0, white_bkgd=False, lindisp=False,
                near=float(meta["near"]), far=float(meta["far"]), use_viewdirs=True, ndc=False)


def _maps(rgb, disp, acc, extras):
    out = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, **{k: v for k, v in extras.items() if torch.is_tensor(v)})
    return {k: v.float().cpu() for k, v in out.items()}


@pytest.mark.parametrize("name", FRAME_CROPS_CFG4)
def test_frame_crop_config4_expression_sweep(name):
    """BASELINE config #4 (run_fit.py:394-403): render_fitting with expCodes = expCodes_Sigma[e] on 192 rays of the
    800x800 frame, real widths (W_c 256 / W_f 1024), against the unmodified reference's output."""
    meta, inp, gold = load_case(name)
    c, f, s, r = build_reference_like(int(meta["seed"]))
    r = r.to(DEV)
    e = int(inp["exp_slot"])
    assert torch.equal(r.expCodes_Sigma[e].detach().cpu().reshape(-1), inp["exp"].reshape(-1))
    with torch.no_grad():
        out = r.render_fitting(int(meta["H"]), int(meta["W"]), inp["K"].numpy(), chunk=1 << 20,
                               rays=(inp["rays_o"].to(DEV), inp["rays_d"].to(DEV)), shapeCodes=inp["shape"].to(DEV),
                               uvCodes=inp["tex"].to(DEV), expType=20, expCodes=r.expCodes_Sigma[e], **_crop_kwargs(meta, c, f))
    torch.cuda.synchronize()
    check_maps(name, _maps(*out), gold)


@pytest.mark.parametrize("name", FRAME_CROPS_CFG5)
def test_frame_crop_config5_identities_through_render(name):
    """BASELINE config #5 (render_refine_trainSet.py:245-304 -> render_path -> render): per identity a shape code, a
    512x512 UV map through the texture encoder, an expression slot and a view; 192 rays of the 800x800 frame."""
    meta, inp, gold = load_case(name)
    c, f, s, r = build_reference_like(int(meta["seed"]))
    r = r.to(DEV)
    uv = torch.rand(512, 512, 3, generator=torch.Generator().manual_seed(int(inp["uv_seed"])))
    # the fixture's texture code comes from the reference's fp32 CPU convolutions; PyTorch's default on GPU lets cuDNN
    # use TF32 for convolutions (4e-4 on the code, measured) — the engine is not involved in that step, so pin it to fp32
    # here to feed both sides the same latent ("identical rays and latents")
    tf32_was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        out = r.render(int(meta["H"]), int(meta["W"]), inp["K"].numpy(), chunk=1 << 20,
                       rays=(inp["rays_o"].to(DEV), inp["rays_d"].to(DEV)), shapeCodes=inp["shape"].to(DEV),
                       uvMap=uv.to(DEV), expType=int(inp["exp_slot"]), **_crop_kwargs(meta, c, f))
        tex = r.decoding_texCodes.reshape(-1).cpu()
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = tf32_was
    d_tex = (tex - inp["tex"]).abs().max().item()
    parity_log.record(name, tex_code_max=d_tex)
    assert d_tex <= 1e-4, f"texture encoder (cuDNN) vs reference (CPU): {d_tex:.2e}"
    check_maps(name, _maps(*out), gold)


def test_psnr_to_target_within_0p1_db_of_the_reference():
    """north_star: "rendered PSNR within 0.1 dB of reference on the fit/render demo".  Without a trained checkpoint the
    demo is emulated with the reference's own outputs: the image rendered with expression slot 9 is the TARGET, the
    renders with slots 14 and 2 (and of other identities) play the fitted result.  PSNR(engine render, target) must equal
    PSNR(reference render, target) to 0.1 dB — the reference renders are the fixtures produced by the unmodified
    reference."""
    meta, inp9, gold9 = load_case("cfg4_800_exp9")
    target = gold9["rgb_map"]
    c, f, s, r = build_reference_like(int(meta["seed"]))
    r = r.to(DEV)
    for name in ("cfg4_800_exp14", "cfg4_800_exp2"):
        m2, inp, gold = load_case(name)
        e = int(inp["exp_slot"])
        # same camera and ray subset as the target?  the crops use different random ray subsets: render the TARGET's rays
        with torch.no_grad():
            out = r.render_fitting(int(meta["H"]), int(meta["W"]), inp9["K"].numpy(), chunk=1 << 20,
                                   rays=(inp9["rays_o"].to(DEV), inp9["rays_d"].to(DEV)), shapeCodes=inp["shape"].to(DEV),
                                   uvCodes=inp["tex"].to(DEV), expType=20, expCodes=r.expCodes_Sigma[e],
                                   **_crop_kwargs(meta, c, f))[0].float().cpu()
            rays = O.make_ray_batch(inp9["rays_o"], inp9["rays_d"], 8.0, 26.0)
            ref = O.render_rays(rays[:64], c.cpu(), f.cpu(), inp["shape"], O.expression_mod(s, inp["shape"], inp["exp"]),
                                inp["tex"])["rgb_map"]
        p_eng = O.psnr(out[:64], target[:64])
        p_ref = O.psnr(ref, target[:64])
        parity_log.record(f"psnr-to-target[{name} vs exp9 target]", engine_db=p_eng, reference_db=p_ref, delta_db=abs(p_eng - p_ref))
        print(f"[parity] PSNR to target: engine {p_eng:.3f} dB, reference {p_ref:.3f} dB")
        assert abs(p_eng - p_ref) <= 0.1, f"{name}: engine {p_eng:.3f} dB vs reference {p_ref:.3f} dB"


def test_simt_and_tensor_core_paths_agree():
    meta, inp, gold = load_case("small_w256")
    nets = build_case_nets(meta)
    a = engine_render(meta, inp, nets, gemm_simt=False)
    b = engine_render(meta, inp, nets, gemm_simt=True)
    check_maps("small_w256[simt]", b, gold, max_rgb0=2e-3)    # single fp16 on the coarse pass as well
    d0 = (a["rgb0"] - b["rgb0"]).abs().max().item()
    d = (a["rgb_map"] - b["rgb_map"]).abs().max().item()
    print(f"[parity] tcgen05 vs SIMT: rgb0 {d0:.2e} rgb {d:.2e}")
    parity_log.record("small_w256[simt vs tcgen05]", rgb0_max=d0, rgb_map_max=d)
    # the default coarse path is split precision (fp32-class), the SIMT verification path single fp16: they differ by the
    # fp16 chain's own error on the coarse maps (~1e-3)
    assert d0 <= 2e-3 and d <= 6e-2, f"tcgen05 vs SIMT dense kernels disagree: coarse {d0:.3e} final {d:.3e}"


@pytest.mark.parametrize("name", ["small_w256", "full_w1024", "perturb_pytest"])
def test_teacher_forced_stages(name):
    """Feed the engine the ORACLE's sample points for both passes (run_network), composite with the engine's
    raw2outputs on the oracle's depths: isolates the fp16 dense chain from resampling sensitivity."""
    from mofanerf_b200 import B200Renderer
    meta, inp, _ = load_case(name)
    nets = build_case_nets(meta)
    c, f, s = nets
    ref, rays, em = oracle_render(meta, inp, nets)
    with torch.no_grad():
        rnd = case_randoms(meta, rays.shape[0])
        z_c = O.coarse_z_vals(rays[:, 6:7], rays[:, 7:8], int(meta["N_samples"]), perturb=float(meta["perturb"]),
                              t_rand=rnd["t_rand"])
        pts_c = rays[:, None, 0:3] + rays[:, None, 3:6] * z_c[..., None]
        z_f = ref["z_vals_fine"]
        pts_f = rays[:, None, 0:3] + rays[:, None, 3:6] * z_f[..., None]
        raw_c_ref = O.run_network(pts_c, rays[:, 8:11], c, inp["shape"], em, inp["tex"])
        r = B200Renderer(expCodesLen=30).to(DEV)
        r.idSpecificMod.load_state_dict(s.state_dict())
        r.shapeCodes, r.expType, r.decoding_texCodes = inp["shape"].to(DEV), 20, inp["tex"].to(DEV)
        r.expCodes_Sigma.append(inp["exp"].to(DEV))
        eng = r.engine(DEV)
        eng.load_network(0, c.to(DEV))
        eng.load_network(1, f.to(DEV))
        raw_c = r.run_network(pts_c.to(DEV), rays[:, 8:11].to(DEV), c)
        raw_f = r.run_network(pts_f.to(DEV), rays[:, 8:11].to(DEV), f)
        rgb_c = eng.raw2outputs(raw_c, z_c, rays[:, 3:6], rnd["noise_c"])[0].cpu()
        rgb_f, _, acc_f, w_f, _ = eng.raw2outputs(raw_f, z_f, rays[:, 3:6], rnd["noise_f"])
    for nm, a, b in (("coarse", raw_c.cpu(), raw_c_ref), ("fine", raw_f.cpu(), ref["raw"])):
        d = (a - b).abs().max().item()
        sc = max(1.0, b.abs().max().item())
        print(f"[parity] {name} teacher-forced raw {nm}: max {d:.3e} (|ref|max {sc:.1f})")
        parity_log.record(f"{name}[teacher-forced]", **{f"raw_{nm}_max": d, f"raw_{nm}_refmax": sc})
        assert d <= 2e-2 * sc
    d0 = (rgb_c - ref["rgb0"]).abs().max().item()
    d1 = (rgb_f.cpu() - ref["rgb_map"]).abs().max().item()
    dw = (w_f.cpu() - ref["weights"]).abs().max().item()
    print(f"[parity] {name} teacher-forced rgb: coarse {d0:.2e} fine {d1:.2e} weights {dw:.2e}")
    parity_log.record(f"{name}[teacher-forced]", rgb0_max=d0, rgb_map_max=d1, weights_max=dw)
    assert d0 <= 5e-5 and d1 <= 2e-3


def test_stagewise_vs_oracle_and_ray_order():
    """Per-stage comparison with the oracle on the full-width net, plus ray-order and chunk invariance."""
    meta, inp, gold = load_case("full_w1024")
    nets = build_case_nets(meta)
    ref, rays, _ = oracle_render(meta, inp, nets)
    got = engine_render(meta, inp, nets, want_aux=True)
    # coarse sample depths are pure fp32 arithmetic: the first 64 of the merged fine depths include them
    zf = got["z_vals"]
    assert bool((zf[:, 1:] >= zf[:, :-1]).all())
    dz = (zf - ref["z_vals_fine"]).abs()
    assert dz.mean().item() < 2e-2, f"fine sample depths drift: mean {dz.mean().item():.3e}"
    raw_d = (got["raw"] - ref["raw"]).abs()
    print(f"[parity] raw(fine): max {raw_d.max().item():.3e} mean {raw_d.mean().item():.3e}")
    parity_log.record("full_w1024[stagewise]", z_fine_mean=dz.mean().item(), z_fine_max=dz.max().item(),
                      raw_fine_max=raw_d.max().item(), raw_fine_mean=raw_d.mean().item())
    # ray order: permute the input rays; outputs must permute identically (bit-exact)
    perm = torch.randperm(inp["rays_o"].shape[0], generator=torch.Generator().manual_seed(0))
    inp2 = dict(inp, rays_o=inp["rays_o"][perm], rays_d=inp["rays_d"][perm])
    got2 = engine_render(meta, inp2, nets)
    assert torch.equal(got2["rgb_map"], got["rgb_map"][perm]), "ray order / per-ray independence violated"
    # chunk invariance (engine-internal chunk and Python-level chunk): bit-exact
    got3 = engine_render(meta, inp, nets, chunk=7, engine_chunk=3)
    assert torch.equal(got3["rgb_map"], got["rgb_map"])
    assert torch.equal(got3["z_std"], got["z_std"])


def test_run_network_matches_nerf_forward():
    """models/model.py:121-137 through myRenderer.run_network (render_class.py:69-94)."""
    from mofanerf_b200 import B200Renderer
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    g = torch.Generator().manual_seed(9)
    pts = (torch.rand(50, 7, 3, generator=g) * 2 - 1) * 10
    vd = torch.nn.functional.normalize(torch.randn(50, 3, generator=g), dim=-1)
    em = O.expression_mod(s, inp["shape"], inp["exp"])
    with torch.no_grad():
        ref = O.run_network(pts, vd, f, inp["shape"], em, inp["tex"])
        r = B200Renderer(expCodesLen=30).to(DEV)
        r.idSpecificMod.load_state_dict(s.state_dict())
        r.shapeCodes, r.expType, r.decoding_texCodes = inp["shape"].to(DEV), 20, inp["tex"].to(DEV)
        r.expCodes_Sigma.append(inp["exp"].to(DEV))
        out = r.run_network(pts.to(DEV), vd.to(DEV), f.to(DEV)).cpu()
    d = (out - ref).abs()
    scale = ref.abs().max().item()
    print(f"[parity] run_network: max {d.max().item():.3e} (|ref|max {scale:.2f})")
    parity_log.record("run_network[fine W=256]", raw_max=d.max().item(), raw_refmax=scale)
    assert d.max().item() <= 2e-2 * max(1.0, scale)


def test_split_precision_coarse_kernel_is_fp32_class():
    """The coarse net (W = 256) runs on the tensor cores in split precision (fp16 hi + lo operands, three products per
    layer, coarse_split.cu): its per-point outputs must sit at the fp32 noise floor of the reference itself (the oracle
    changes by 2.5e-5 between netchunk blockings) — not at the 1e-2 of a single-fp16 chain.  Also pins the
    single-fp16 fused kernel of round 1 (MOFA_B200_COARSE_FP16=1, a separate engine) on the same points."""
    import os
    from mofanerf_b200.engine import Engine
    meta, inp, _ = load_case("full_w1024")
    c, f, s = build_case_nets(meta)
    g = torch.Generator().manual_seed(17)
    n, S = 37, 64                       # 2368 points: not a multiple of the 256-row pair tile
    rays = O.make_ray_batch(inp["rays_o"][:n], inp["rays_d"][:n], 8.0, 26.0)
    z = torch.sort(8.0 + 18.0 * torch.rand(n, S, generator=g), -1)[0]
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[..., None]
    em = O.expression_mod(s, inp["shape"], inp["exp"])
    with torch.no_grad():
        ref = O.run_network(pts, rays[:, 8:11], c, inp["shape"], em, inp["tex"])
    scale = max(1.0, ref.abs().max().item())
    res = {}
    for mode in ("split", "fp16"):
        if mode == "fp16":
            os.environ["MOFA_B200_COARSE_FP16"] = "1"
        try:
            eng = Engine(DEV)
        finally:
            os.environ.pop("MOFA_B200_COARSE_FP16", None)
        eng.load_network(0, c.to(DEV))
        eng.set_latents(inp["shape"], em, inp["tex"])
        out = eng.run_network(0, pts.to(DEV), rays[:, None, 8:11].to(DEV)).cpu()
        torch.cuda.synchronize()
        eng.close()
        res[mode] = (out - ref).abs().max().item()
    print(f"[parity] coarse raw vs fp32 oracle: split {res['split']:.3e}, single fp16 {res['fp16']:.3e} (|ref|max {scale:.2f})")
    parity_log.record("coarse net raw[split vs fp16]", split_max=res["split"], fp16_max=res["fp16"], raw_refmax=scale)
    assert res["split"] <= 2e-4 * scale, f"split-precision coarse kernel: {res['split']:.3e}"
    assert res["fp16"] <= 2e-2 * scale, f"single-fp16 fused coarse kernel: {res['fp16']:.3e}"
    assert res["split"] < 0.1 * res["fp16"]


def test_training_mode_forward_equals_inference_forward():
    """With grad-requiring inputs the renderer switches to the activation-keeping forward (fitting).  It runs the
    per-layer single-fp16 kernels where inference runs the split-precision fused coarse kernel, so the maps agree to the
    fp16 chain's error, not bit for bit: coarse maps 2e-3, final maps 2e-2 (resampling feedback)."""
    from mofanerf_b200 import B200Renderer
    meta, inp, _ = load_case("small_w256")
    c, f, s = build_case_nets(meta)
    r = B200Renderer(expCodesLen=30).to(DEV)
    r.idSpecificMod.load_state_dict(s.state_dict())
    args = dict(rays=(inp["rays_o"][:16].to(DEV), inp["rays_d"][:16].to(DEV)), uvCodes=inp["tex"].to(DEV), expType=20,
                expCodes=inp["exp"].to(DEV), near=8., far=26., use_viewdirs=True, ndc=False, network_fn=c.to(DEV),
                network_fine=f.to(DEV), N_samples=64, N_importance=64)
    with torch.no_grad():
        a = r.render_fitting(4, 4, None, shapeCodes=inp["shape"].to(DEV), **args)
    b = r.render_fitting(4, 4, None, shapeCodes=inp["shape"].to(DEV).requires_grad_(True), **args)
    assert b[0].requires_grad and b[0].grad_fn is not None
    assert (a[3]["rgb0"] - b[3]["rgb0"].detach()).abs().max().item() <= 2e-3
    assert (a[0] - b[0].detach()).abs().max().item() <= 2e-2 and (a[2] - b[2].detach()).abs().max().item() <= 2e-2


def _frame_sample_render(name):
    meta, inp, gold = load_case(name)
    c, f, s = build_case_nets(meta)
    from mofanerf_b200 import B200Renderer
    r = B200Renderer(expCodesLen=30).to(DEV)
    r.idSpecificMod.load_state_dict(s.state_dict())
    with torch.no_grad():
        out = r.render_fitting(int(meta["H"]), int(meta["W"]), None, chunk=1 << 30,
                               rays=(inp["rays_o"].to(DEV), inp["rays_d"].to(DEV)), shapeCodes=inp["shape"].to(DEV),
                               uvCodes=inp["tex"].to(DEV), expType=20, expCodes=inp["exp"].to(DEV), retraw=True,
                               **_crop_kwargs(meta, c, f))
    return meta, inp, gold, _maps(*out)


def test_frame_sample_1024_dense_field_against_the_reference():
    """1024 rays spread over the 800x800 frame, rendered by the UNMODIFIED reference (tests/golden/frame_sample_1024.npz;
    the nets of the config #4 / #5 crops: a dense field, acc ~ 1): the per-ray bound of the fixtures holds on the whole
    sample."""
    meta, inp, gold, got = _frame_sample_render("frame_sample_1024")
    check_maps("frame_sample_1024", got, gold)


def test_frame_sample_1024_bench_frame_against_the_reference():
    """The same sample of bench.py's synthetic frame (seed-0 nets; tests/golden/frame_sample_bench_1024.npz, unmodified
    reference on the CPU).  The random-init fine net of this seed leaves the volume almost empty (median fine acc 0.03)
    and sigma hovers around 0 along whole rays: the reference's last-interval step (alpha_last = 1 - exp(-relu(sigma_last)
    * 1e10), models/render_class.py:449) then decides single rays on the SIGN of a sigma that is below the rounding of
    any 10-bit-mantissa arithmetic.  Asserted: the error distribution; that every opacity flip is such a ray according to
    the reference's own stored sigma_last; the per-ray bound on all other rays."""
    import bench
    meta, inp, gold, got = _frame_sample_render("frame_sample_bench_1024")
    n = gold["rgb_map"].shape[0]
    sig_last = got["raw"].reshape(n, -1, 4)[:, -1, 3]
    st = bench.parity_stats(got["rgb_map"], got["acc_map"], gold["rgb_map"], gold["acc_map"], sig_last, inp["sigma_last"])
    parity_log.record("frame_sample_bench_1024", rays=n, **st)
    print(f"[parity] frame_sample_bench_1024: {st}")
    assert (got["rgb0"] - gold["rgb0"]).abs().max().item() <= 5e-5            # the coarse pass is fp32-class here too
    assert st["err_p50"] <= 3e-4 and st["err_p99"] <= 1.5e-2, st
    assert st["frac_rays_within_3e-2"] >= 0.99, st
    # a flip = the sign of sigma at the last sample differs from the reference's: only where the reference's own value
    # is within the fp16 chain's rounding of zero
    assert st["opacity_gate_flips"] <= n // 100 and st["gate_flip_max_abs_sigma_last_of_reference"] < 2e-2, st
    assert st["max_abs_rgb_excluding_gate_flips"] <= 6e-2 and st["psnr_db_excluding_gate_flips"] >= 50.0, st


def test_frame_sample_error_distribution_with_the_references_own_yardsticks():
    """4096 rays spread over the whole 800x800 bench frame at the real widths (the fixtures above are 192-ray crops): the
    DISTRIBUTION of the engine's per-ray error against the fp32 reference algorithm, next to the reference algorithm's
    own deviation when its matmuls run in TF32 — the default of torch 1.9 (the reference's pinned version) on Ampere and
    later GPUs, and the same 10-bit operand mantissa as the engine's fp16 fine net.

    Why a distribution and not one maximum: raw2outputs gives the LAST sample of a ray an interval of 1e10
    (models/render_class.py:449), so alpha_last = 1 - exp(-relu(sigma_last) * 1e10) is a step function of sigma_last.  A
    ray whose last fine sample has |sigma| below the arithmetic's rounding (measured: fp16 chain p99 3e-3 on |sigma| ~ 0.5)
    renders with acc = 1 or acc << 1 depending on a sign no finite-precision implementation reproduces ("opacity gate
    flips": 3 of 8192 rays of this synthetic frame, for the engine AND for the TF32 reference, on the same rays).  The
    random-init fine net makes this frame almost empty (median fine acc 0.03), which is the worst case for it."""
    import bench
    n = 4096
    c, f, s = O.build_nets(0)
    shape, tex, exp, ro, rd = bench.synth_inputs(800, 800)
    idx = torch.linspace(0, ro.shape[0] - 1, n).long()
    cg, fg, sg = c.to(DEV), f.to(DEV), s.to(DEV)
    rays = O.make_ray_batch(ro[idx], rd[idx], 8.0, 26.0).to(DEV)
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        with torch.no_grad(), torch.device(DEV):
            em = O.expression_mod(sg, shape.to(DEV), exp.to(DEV))
            torch.backends.cuda.matmul.allow_tf32 = False
            ref = O.render_rays(rays, cg, fg, shape.to(DEV), em, tex.to(DEV), N_samples=64, N_importance=64,
                                netchunk=196608, retraw=True)
            torch.backends.cuda.matmul.allow_tf32 = True
            tf = O.render_rays(rays, cg, fg, shape.to(DEV), em, tex.to(DEV), N_samples=64, N_importance=64, netchunk=196608,
                               retraw=True)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    from mofanerf_b200 import B200Renderer
    r = B200Renderer(expCodesLen=30).to(DEV)
    r.idSpecificMod.load_state_dict(s.state_dict())
    with torch.no_grad():
        rgb, disp, acc, extras = r.render_fitting(
            1, n, None, chunk=1 << 30, rays=(ro[idx].to(DEV), rd[idx].to(DEV)), shapeCodes=shape.to(DEV),
            uvCodes=tex.to(DEV), expType=20, expCodes=exp.to(DEV), near=8.0, far=26.0, use_viewdirs=True, ndc=False,
            network_fn=cg, network_fine=fg, N_samples=64, N_importance=64, perturb=0.0, raw_noise_std=0.0, retraw=True)
    last = lambda raw: raw.reshape(n, -1, 4)[:, -1, 3]
    st = bench.parity_stats(rgb, acc, ref["rgb_map"], ref["acc_map"], last(extras["raw"]), last(ref["raw"]))
    yt = bench.parity_stats(tf["rgb_map"], tf["acc_map"], ref["rgb_map"], ref["acc_map"], last(tf["raw"]), last(ref["raw"]))
    for tag, d in (("engine_vs_fp32", st), ("reference_tf32_vs_fp32", yt)):
        parity_log.record("frame_sample_4096_" + tag, rays=n, **{k: v for k, v in d.items()})
    print(f"[parity] frame sample: engine {st}\n[parity] frame sample: TF32 reference {yt}")
    # coarse maps stay fp32-class on the large sample too
    assert (extras["rgb0"].reshape(-1, 3) - ref["rgb0"]).abs().max().item() <= 5e-5
    # the distribution (measured on 8192 rays: p50 1.1e-4, p99 7.5e-3, 99.87 % of rays within 3e-2, 57 dB without the flips)
    assert st["err_p50"] <= 3e-4 and st["err_p99"] <= 1.5e-2, st
    assert st["frac_rays_within_3e-2"] >= 0.995, st
    assert st["max_abs_rgb_excluding_gate_flips"] <= 6e-2 and st["psnr_db_excluding_gate_flips"] >= 50.0, st
    # every opacity-gate flip (sign of sigma at the last sample differs) is a ray whose REFERENCE sigma there is within
    # the fp16 chain's rounding of zero
    assert st["opacity_gate_flips"] <= n // 100 and st["gate_flip_max_abs_sigma_last_of_reference"] < 2e-2, st
    # the engine is at least as close to the fp32 reference as the reference's own TF32 mode (measured: 3x closer)
    assert st["mean_abs_rgb"] <= yt["mean_abs_rgb"] and st["err_p99"] <= yt["err_p99"], (st, yt)
