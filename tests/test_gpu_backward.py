"""Backward pass for fitting (run_fit.py:305-313) against torch.autograd through the oracle.

Both sides are evaluated at IDENTICAL sample depths (the engine's fine depths are fed to the oracle through
z_fine_override; the reference detaches the resampled depths, models/render_class.py:326), so the comparison
isolates the gradient arithmetic: fp16 operands / fp32 accumulation in the dense layers with a power-of-two
loss scale, fp32 everywhere else.  Stated tolerance: relative L2 error <= 3e-2 and cosine >= 0.999 per gradient
tensor (measured ~1e-3 .. 1e-2); compositing backward alone (fp32): 1e-4 relative.
"""
import pytest
import torch

from oracle import mofa_oracle as O
from tests.helpers import build_case_nets, case_randoms, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def test_composite_backward_matches_autograd():
    from mofanerf_b200 import get_engine, _lib
    import ctypes as C
    g = torch.Generator().manual_seed(0)
    n, S = 33, 96
    raw = (torch.randn(n, S, 4, generator=g) * 1.5).requires_grad_(True)
    z = torch.sort(torch.rand(n, S, generator=g) * 18 + 8, -1)[0]
    d = torch.randn(n, 3, generator=g).requires_grad_(True)
    noise = torch.rand(n, S, generator=g) * 0.5
    w_rgb, w_acc = torch.randn(n, 3, generator=g), torch.randn(n, generator=g)
    rgb, _, acc, _, _ = O.raw2outputs(raw, z, d, noise, True)
    ((rgb * w_rgb).sum() + (acc * w_acc).sum()).backward()
    eng = get_engine(DEV)
    rays = torch.zeros(n, 11)
    rays[:, 3:6] = d.detach()
    # drive the kernel through the library's internal launcher via the public backward entry point is not possible
    # without networks; use the op-level path exposed for tests
    out = eng.composite_bwd(raw.detach(), z, rays, noise, w_rgb, w_acc, white_bkgd=True)
    e, c = _rel(out[0].cpu(), raw.grad)
    assert e < 1e-4, f"d_raw rel {e:.2e}"
    e, c = _rel(out[1][:, 3:6].cpu(), d.grad)
    assert e < 1e-4, f"d_rays_d rel {e:.2e}"


@pytest.mark.parametrize("name,use0", [("small_w256", True), ("full_w1024", False), ("perturb_pytest", True)])
def test_fitting_gradients_match_oracle_autograd(name, use0):
    from mofanerf_b200 import B200Renderer
    meta, inp, _ = load_case(name)
    c, f, s = build_case_nets(meta)
    n = min(32, inp["rays_o"].shape[0])
    ro, rd = inp["rays_o"][:n].clone(), inp["rays_d"][:n].clone()
    rnd = {k: (None if v is None else v[:n]) for k, v in case_randoms(meta, inp["rays_o"].shape[0]).items()}
    g = torch.Generator().manual_seed(1)
    wts = dict(rgb=torch.randn(n, 3, generator=g), acc=torch.randn(n, generator=g) * 0.3,
               rgb0=torch.randn(n, 3, generator=g) * (1.0 if use0 else 0.0), acc0=torch.randn(n, generator=g) * 0.2 * use0)

    def loss_of(out):
        return ((out["rgb_map"] * wts["rgb"].to(out["rgb_map"].device)).sum() +
                (out["acc_map"] * wts["acc"].to(out["rgb_map"].device)).sum() +
                (out["rgb0"] * wts["rgb0"].to(out["rgb_map"].device)).sum() +
                (out["acc0"] * wts["acc0"].to(out["rgb_map"].device)).sum())

    # ---------------- engine (GPU), autograd through B200Renderer.render_fitting
    r = B200Renderer(expCodesLen=30).to(DEV)
    r.idSpecificMod.load_state_dict(s.state_dict())
    leaves = dict(ro=ro.to(DEV).requires_grad_(True), rd=rd.to(DEV).requires_grad_(True),
                  shape=inp["shape"].to(DEV).requires_grad_(True), tex=inp["tex"].to(DEV).requires_grad_(True),
                  exp=inp["exp"].to(DEV).requires_grad_(True))
    kw = dict(near=float(meta["near"]), far=float(meta["far"]), use_viewdirs=True, ndc=False, network_fn=c.to(DEV),
              network_fine=f.to(DEV), N_samples=int(meta["N_samples"]), N_importance=int(meta["N_importance"]),
              perturb=float(meta["perturb"]), raw_noise_std=float(meta["raw_noise_std"]),
              white_bkgd=bool(meta["white_bkgd"]), lindisp=bool(meta["lindisp"]), want_aux=True, **rnd)
    rgb, disp, acc, ex = r.render_fitting(1, n, None, rays=(leaves["ro"], leaves["rd"]), shapeCodes=leaves["shape"],
                                          uvCodes=leaves["tex"], expType=20, expCodes=leaves["exp"], **kw)
    assert rgb.requires_grad and not disp.requires_grad
    loss = loss_of(dict(rgb_map=rgb, acc_map=acc, rgb0=ex["rgb0"], acc0=ex["acc0"]))
    loss.backward()
    got = {k: v.grad.detach().cpu() for k, v in leaves.items()}
    z_fine = ex["z_vals"].detach().cpu()
    c.cpu(); f.cpu()

    # ---------------- oracle (CPU fp32), autograd, same fine depths
    ol = dict(ro=ro.clone().requires_grad_(True), rd=rd.clone().requires_grad_(True),
              shape=inp["shape"].clone().requires_grad_(True), tex=inp["tex"].clone().requires_grad_(True),
              exp=inp["exp"].clone().requires_grad_(True))
    rays = O.make_ray_batch(ol["ro"], ol["rd"], float(meta["near"]), float(meta["far"]))
    em = O.expression_mod(s, ol["shape"], ol["exp"])
    out = O.render_rays(rays, c, f, ol["shape"], em, ol["tex"], N_samples=int(meta["N_samples"]),
                        N_importance=int(meta["N_importance"]), perturb=float(meta["perturb"]),
                        white_bkgd=bool(meta["white_bkgd"]), lindisp=bool(meta["lindisp"]), z_fine_override=z_fine, **rnd)
    loss_of(out).backward()
    msgs = []
    for k in ("shape", "tex", "exp", "ro", "rd"):
        e, cs = _rel(got[k], ol[k].grad)
        msgs.append(f"{k}: rel {e:.2e} cos {cs:.5f}")
        assert e <= 3e-2 and cs >= 0.999, f"{name} grad {k}: rel {e:.3e} cos {cs:.5f} |ref| {ol[k].grad.norm().item():.3e}"
    print(f"[parity] {name} gradients: " + "; ".join(msgs))
