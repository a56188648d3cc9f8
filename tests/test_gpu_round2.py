"""Round-2 GPU tests: in-kernel ray generation, the fine-net chain kernel against one launch per layer,
padded 16-byte ray rows."""
import os

import numpy as np
import pytest
import torch

from oracle import mofa_oracle as O
from tests import parity_log
from tests.helpers import GOLDEN, build_case_nets, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ulps(a, b):
    def ordered(x):      # monotone integer image of the floats (negative values mirrored), so differences count ulps
        i = x.contiguous().view(torch.int32).long()
        return torch.where(i < 0, -(i & 0x7FFFFFFF), i)
    return (ordered(a) - ordered(b)).abs().max().item()


def test_generate_rays_matches_get_rays_bit_exact():
    """mofa_b200_generate_rays against tools/run_nerf_helpers.py:153-168 (fixture from the unmodified reference for a
    6x10 camera, and the oracle's restatement — itself pinned to that fixture — for the 800x800 frame of the bench):
    origins and directions bit for bit, row-major ray order, near / far columns, unit view directions; any sub-range
    (what a rank of a ray-sharded render generates) equals the same rows of the full batch."""
    from mofanerf_b200 import get_engine
    eng = get_engine(DEV)
    z = np.load(os.path.join(GOLDEN, "ops.npz"))
    K, c2w = z["rays_K"], torch.from_numpy(z["rays_c2w"])
    rays = eng.generate_rays(6, 10, K, c2w[:3, :4], 8.0, 26.0).cpu()
    assert rays.shape == (60, 12)
    assert torch.equal(rays[:, 0:3], torch.from_numpy(z["rays_o"]).reshape(-1, 3))
    assert torch.equal(rays[:, 3:6], torch.from_numpy(z["rays_d"]).reshape(-1, 3))
    H = W = 800
    focal = 1200.0 * H / 512.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    worst = 0
    for angle in (30.0, -110.0):
        c2w = O.pose_spherical(angle, 0.0, 16.0)
        ro, rd = O.get_rays(H, W, K, c2w[:3, :4])
        ref = O.make_ray_batch(ro, rd, 8.0, 26.0)
        got = eng.generate_rays(H, W, K, c2w[:3, :4].to(DEV), 8.0, 26.0).cpu()
        assert got.shape == (H * W, 12) and float(got[:, 11].abs().max()) == 0.0
        assert torch.equal(got[:, 0:8], ref[:, 0:8]), "origins / directions / near / far differ from get_rays"
        assert _ulps(got[:, 0:8], ref[:, 0:8]) == 0, "signed zeros differ from get_rays"
        u = _ulps(got[:, 8:11], ref[:, 8:11])
        worst = max(worst, u)
        assert u <= 1, f"view directions differ from rays_d / torch.norm(rays_d) by {u} ulp"
        sub = eng.generate_rays(H, W, K, c2w[:3, :4], 8.0, 26.0, first=123457, n=1001).cpu()
        assert torch.equal(sub, got[123457:123457 + 1001])
    parity_log.record("generate_rays 800x800", rays_od_max_ulp=0, viewdir_max_ulp=worst)
    with pytest.raises(RuntimeError, match="outside"):
        eng.generate_rays(4, 4, K, c2w[:3, :4], 8.0, 26.0, first=10, n=10)


def test_render_from_camera_equals_render_from_rays():
    """render_fitting(c2w=...) (rays generated on the device, 12-float rows) against render_fitting(rays=...) with the
    host-generated rays of the same camera, and 11-float rows against padded rows: same image."""
    from mofanerf_b200 import B200Renderer
    meta, inp, _ = load_case("full_w1024")
    c, f, s = build_case_nets(meta)
    r = B200Renderer(expCodesLen=30).to(DEV)
    r.idSpecificMod.load_state_dict(s.state_dict())
    H, W = 12, 10
    focal = 1200.0 * H / 512.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = O.pose_spherical(25.0, 0.0, 16.0)
    kw = dict(shapeCodes=inp["shape"].to(DEV), uvCodes=inp["tex"].to(DEV), expType=20, expCodes=inp["exp"].to(DEV), near=8.0,
              far=26.0, use_viewdirs=True, ndc=False, network_fn=c.to(DEV), network_fine=f.to(DEV), N_samples=64,
              N_importance=64, perturb=0.0, raw_noise_std=0.0)
    ro, rd = O.get_rays(H, W, K, c2w[:3, :4])
    with torch.no_grad():
        a = r.render_fitting(H, W, K, c2w=c2w[:3, :4].to(DEV), **kw)
        b = r.render_fitting(H, W, K, rays=(ro.to(DEV), rd.to(DEV)), **kw)
    assert a[0].shape == (H, W, 3) and b[0].shape == (H, W, 3)
    d = (a[0] - b[0]).abs().max().item()
    parity_log.record("render(c2w) vs render(rays)", rgb_map_max=d)
    assert d <= 2e-3, f"camera path vs ray path: {d:.2e}"      # view directions may differ by 1 ulp before fp16 rounding
    eng = r.engine(DEV)
    rays12 = eng.generate_rays(H, W, K, c2w[:3, :4], 8.0, 26.0)
    rays11 = rays12[:, :11].contiguous()
    o12 = eng.render_rays(rays12, 64, 64)
    o11 = eng.render_rays(rays11, 64, 64)
    assert torch.equal(o12["rgb_map"], o11["rgb_map"]) and torch.equal(o12["z_std"], o11["z_std"])


def test_fine_chain_kernel_equals_per_layer_launches():
    """The persistent chain kernel (all 25 dense layers of the W = 1024 net in one launch, L2-resident activation
    slabs) runs the same tiles with the same arithmetic as one launch per layer: every output bit for bit, for ray
    counts below, at and above a slab (74 rays at 128 samples) and not multiples of a tile."""
    from mofanerf_b200 import nets
    from mofanerf_b200.engine import Engine
    coarse, fine, _ = nets.build_nets(0, device=DEV)
    g = torch.Generator().manual_seed(3)
    shape, tex, em = torch.randn(50, generator=g) * 0.034, 0.14 + 0.26 * torch.randn(256, generator=g), torch.rand(30, generator=g)

    def rays_for(n):
        gg = torch.Generator().manual_seed(n)
        rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=gg) * 0.1 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
        ro = torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 16.0])
        return torch.cat([ro, rd, torch.full((n, 1), 8.0), torch.full((n, 1), 26.0), rd, torch.zeros(n, 1)], -1).to(DEV)

    outs = {}
    for mode in ("chain", "per_layer"):
        if mode == "per_layer":
            os.environ["MOFA_B200_FINE_PER_LAYER"] = "1"
        try:
            eng = Engine(DEV)
        finally:
            os.environ.pop("MOFA_B200_FINE_PER_LAYER", None)
        eng.load_network(0, coarse)
        eng.load_network(1, fine)
        eng.set_latents(shape, em, tex)
        for n, S, Ni in ((3, 64, 64), (74, 64, 64), (233, 64, 64), (41, 32, 16)):
            o = eng.render_rays(rays_for(n), S, Ni, retraw=True)
            torch.cuda.synchronize()
            outs[(mode, n)] = {k: v.clone() for k, v in o.items()}
        eng.close()
    for (mode, n), o in outs.items():
        if mode != "chain":
            continue
        ref = outs[("per_layer", n)]
        for k in o:
            assert torch.allclose(o[k], ref[k], rtol=0, atol=0, equal_nan=True), f"n={n} {k}: chain kernel differs from per-layer launches"


def test_packed_weight_cache_round_trip(tmp_path):
    """export the engine's layout of both networks, import it into a fresh engine (no fp32 tensors, no
    modules) and render: bit-identical to rendering from the modules; the on-disk cache hits on the second use and
    misses after a weight changes."""
    from mofanerf_b200 import nets
    from mofanerf_b200.checkpoint import PackedWeightCache
    from mofanerf_b200.engine import Engine
    coarse, fine, _ = nets.build_nets(3, device=DEV)
    g = torch.Generator().manual_seed(5)
    shape, tex, em = torch.randn(50, generator=g) * 0.034, 0.14 + 0.26 * torch.randn(256, generator=g), torch.rand(30, generator=g)
    n = 96
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    rays = torch.cat([torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 16.0]), rd, torch.full((n, 1), 8.0), torch.full((n, 1), 26.0),
                      rd, torch.zeros(n, 1)], -1).to(DEV)
    a = Engine(DEV)
    a.load_network(0, coarse)
    a.load_network(1, fine)
    a.set_latents(shape, em, tex)
    ref = a.render_rays(rays, 64, 64)
    blobs = [a.export_packed(0), a.export_packed(1)]
    assert blobs[1].nbytes > 2 * 27e6 and bytes(blobs[0][:8]) == b"MOFAPK02"
    b = Engine(DEV)
    b.import_packed(0, blobs[0])
    b.import_packed(1, blobs[1])
    b.set_latents(shape, em, tex)
    out = b.render_rays(rays, 64, 64)
    for k in ref:
        assert torch.allclose(out[k], ref[k], rtol=0, atol=0, equal_nan=True), k
    with pytest.raises(RuntimeError, match="blob"):
        b.import_packed(0, blobs[0][:-16])
    cache = PackedWeightCache(str(tmp_path))
    c = Engine(DEV)
    assert cache.load(c, 0, coarse) is False and cache.load(c, 1, fine) is False
    d = Engine(DEV)
    assert cache.load(d, 0, coarse) is True and cache.load(d, 1, fine) is True
    d.set_latents(shape, em, tex)
    assert torch.equal(d.render_rays(rays, 64, 64)["rgb_map"], ref["rgb_map"])
    d.load_network(1, fine)                      # found under the module's key: no repack
    with torch.no_grad():
        fine.rgb_linear.bias.add_(0.25)
    assert cache.load(d, 1, fine) is False       # new content -> new key
    assert not torch.equal(d.render_rays(rays, 64, 64)["rgb_map"], ref["rgb_map"])
    for e in (a, b, c, d):
        e.close()


def test_fine_chain_kernel_on_a_512_wide_net():
    """The chain kernel also takes W = 512 nets (two n-tiles per layer, one for the view layer).  Against one launch per
    layer the view layer's rgb head is summed in a different order (pair tile with two epilogue groups instead of a
    single-CTA tile), so the comparison is to fp32 rounding, not bit for bit."""
    from mofanerf_b200 import nets
    from mofanerf_b200.engine import Engine
    coarse, fine, _ = nets.build_nets(2, W_f=512, device=DEV)
    g = torch.Generator().manual_seed(8)
    shape, tex, em = torch.randn(50, generator=g) * 0.034, 0.14 + 0.26 * torch.randn(256, generator=g), torch.rand(30, generator=g)
    n = 150
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    rays = torch.cat([torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 16.0]), rd, torch.full((n, 1), 8.0), torch.full((n, 1), 26.0),
                      rd, torch.zeros(n, 1)], -1).to(DEV)
    outs = {}
    for mode in ("chain", "per_layer"):
        if mode == "per_layer":
            os.environ["MOFA_B200_FINE_PER_LAYER"] = "1"
        try:
            eng = Engine(DEV)
        finally:
            os.environ.pop("MOFA_B200_FINE_PER_LAYER", None)
        eng.load_network(0, coarse)
        eng.load_network(1, fine)
        eng.set_latents(shape, em, tex)
        outs[mode] = {k: v.clone() for k, v in eng.render_rays(rays, 64, 64, retraw=True).items()}
        torch.cuda.synchronize()
        eng.close()
    assert torch.equal(outs["chain"]["raw"][..., 3], outs["per_layer"]["raw"][..., 3])          # sigma: same tiles, same order
    assert (outs["chain"]["raw"] - outs["per_layer"]["raw"]).abs().max().item() <= 1e-4
    assert (outs["chain"]["rgb_map"] - outs["per_layer"]["rgb_map"]).abs().max().item() <= 1e-5


def test_fp8_variant_matches_its_emulation_and_is_not_the_default():
    """MOFA_B200_FP8 (opt-in measurement mode): the 19 plain 1024 -> 1024 fine layers run as
    tcgen05.mma.kind::f8f6f4 with e4m3 operands.  It must (a) compute what the emulation in tools/fp8_parity_study.py
    says such a variant computes (per-channel weight scales, activations x8, saturating round-to-nearest) and (b) stay
    OFF by default: the default engine on the same rays is the fp16 result of the fixtures."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fp8_parity_study as F8
    from mofanerf_b200.engine import Engine
    meta, inp, gold = load_case("cfg4_800_exp9")
    c, f, s = build_case_nets(meta)
    n = 64
    rays = O.make_ray_batch(inp["rays_o"][:n], inp["rays_d"][:n], 8.0, 26.0)
    em = O.expression_mod(s, inp["shape"], inp["exp"])
    with torch.no_grad():
        emu = O.render_rays(rays, c, f, inp["shape"], em, inp["tex"], forward_fn=F8.make_forward(19), retraw=True)
    res = {}
    for mode in ("fp8", "default"):
        if mode == "fp8":
            os.environ["MOFA_B200_FP8"] = "1"
        try:
            eng = Engine(DEV)
        finally:
            os.environ.pop("MOFA_B200_FP8", None)
        eng.load_network(0, c.to(DEV))
        eng.load_network(1, f.to(DEV))
        eng.set_latents(inp["shape"], em, inp["tex"])
        pad = torch.cat([rays, torch.zeros(n, 1)], 1).to(DEV)
        res[mode] = {k: v.cpu() for k, v in eng.render_rays(pad, 64, 64, retraw=True).items()}
        torch.cuda.synchronize()
        eng.close()
    c.cpu(); f.cpu()
    d_emu = (res["fp8"]["rgb_map"] - emu["rgb_map"]).abs()
    ps_emu = O.psnr(res["fp8"]["rgb_map"], emu["rgb_map"])
    ps_ref = O.psnr(res["fp8"]["rgb_map"], gold["rgb_map"][:n])
    ps_def = O.psnr(res["default"]["rgb_map"], gold["rgb_map"][:n])
    parity_log.record("fp8 variant (19 layers, opt-in)", vs_emulation_max=d_emu.max().item(), vs_emulation_psnr_db=ps_emu,
                      vs_reference_psnr_db=ps_ref, default_vs_reference_psnr_db=ps_def)
    print(f"[parity] fp8 engine vs emulation: max {d_emu.max().item():.2e} psnr {ps_emu:.1f} dB; vs reference {ps_ref:.1f} dB "
          f"(default engine {ps_def:.1f} dB)")
    assert ps_emu >= 45.0, f"the FP8 kernels do not compute the emulated FP8 arithmetic: {ps_emu:.1f} dB"
    assert 25.0 <= ps_ref <= 50.0          # a different product: well below the default path
    assert ps_def >= 60.0
