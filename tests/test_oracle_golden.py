"""Pin the oracle (oracle/mofa_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only.  Tolerance: the reference's own fp32 noise floor across
`netchunk` blocking is 2.5e-5 => atol 1e-4 on rendered maps; op-level cases that
do no GEMM are compared at 1e-6 or exactly."""
import numpy as np
import pytest
import torch

from oracle import mofa_oracle as O
from oracle import ref_loader
from tests.helpers import GOLDEN, assert_close_nan, load_case, oracle_render

OPS = np.load(f"{GOLDEN}/ops.npz")


def T(k):
    return torch.from_numpy(OPS[k])


def test_positional_encoding_exact():
    x = T("pe_x")
    assert torch.equal(O.embed(x, 10), T("pe_out10"))
    assert torch.equal(O.embed(x, 4), T("pe_out4"))
    assert O.embed_dim(10) == 63 and O.embed_dim(4) == 27


@pytest.mark.parametrize("wb", [0, 1])
def test_raw2outputs(wb):
    o = O.raw2outputs(T("r2o_raw"), T("r2o_z"), T("r2o_d"), None, bool(wb))
    for nm, v in zip(("rgb", "disp", "acc", "weights", "depth"), o):
        assert_close_nan(v, T(f"r2o_wb{wb}_{nm}"), 0.0, what=nm)      # identical op sequence: exact
    assert torch.isnan(o[1][:8]).all()                                 # zero density => NaN disparity


def test_sample_pdf():
    bins, w = T("pdf_bins"), T("pdf_w")
    assert torch.equal(O.sample_pdf(bins, w, 64, det=True), T("pdf_det"))
    assert torch.equal(O.sample_pdf(bins, w, 64, det=False, u=T("pdf_u")), T("pdf_rand_pytest"))


def test_nerf_forward():
    torch.manual_seed(3)
    net = O.NeRF(8, 256, 93, 27, 256, 50).eval()
    with torch.no_grad():
        out = net(T("net_in_pts"), T("net_in_shape"), T("net_in_views"), T("net_in_tex"))
    assert_close_nan(out, T("net_out"), 1e-5, what="NeRF.forward")


def test_get_rays():
    ro, rd = O.get_rays(6, 10, OPS["rays_K"], T("rays_c2w")[:3, :4])
    assert torch.equal(ro, T("rays_o")) and torch.equal(rd, T("rays_d"))


def test_flop_model():
    # 3 187 200 / 54 953 984 FLOP per point
    assert O.flops_per_ray(S_c=1, S_f=0, W_f=0) == 3187200
    assert O.flops_per_ray(S_c=0, S_f=1) == 54953984
    assert abs(O.flops_per_ray() - 7238.1e6) < 0.1e6


@pytest.mark.parametrize("name", ["cfg1_64x64_s32", "small_w256", "full_w1024", "perturb_pytest",
                                  "empty_white"])
def test_render_cases(name):
    meta, inp, gold = load_case(name)
    out, _, _ = oracle_render(meta, inp)
    for k, g in gold.items():
        if k == "raw":
            assert_close_nan(out["raw"], g, 2e-4, 1e-4, what=f"{name}:raw")
        else:
            # disparity = 1/depth can be large: relative tolerance there
            assert_close_nan(out[k], g, 1e-4, 1e-4, what=f"{name}:{k}")


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present on this box")
def test_live_reference_crosscheck():
    """Where /root/reference exists, also run the reference live on fresh inputs."""
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(11)
    raw = torch.randn(16, 32, 4, generator=g)
    z = torch.sort(torch.rand(16, 32, generator=g) * 10 + 2, -1)[0]
    d = torch.randn(16, 3, generator=g)
    a = ref.render_class.raw2outputs(raw, z, d, 0, False)
    b = O.raw2outputs(raw, z, d, None, False)
    for x, y in zip(a, b):
        assert_close_nan(y, x, 0.0)
