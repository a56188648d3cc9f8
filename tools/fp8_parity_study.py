"""Parity study for an FP8 (e4m3) variant of the fine net (SURVEY.md §8 row f4, round-1 verdict item 6): PSNR of the
rendered crop against the fp32 oracle as a function of HOW MANY of the 19 plain 1024 -> 1024 fine layers run with fp8
operands (per-output-channel weight scales, power-of-two activation scale, fp32 accumulation — what
tcgen05.mma.kind::f8f6f4 with an fp32 epilogue would compute), everything else as the engine runs it today (fp16 operands
in the remaining fine layers, exact coarse pass, fp32 heads / compositing).  CPU emulation through the oracle, run in the
build container:   python tools/fp8_parity_study.py   ->  profiles/r02_fp8_parity_study.json

This is measurement infrastructure (it imports oracle/); nothing here is on the product path."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mofa_oracle as O  # noqa: E402
from tests.helpers import build_case_nets, load_case  # noqa: E402

F8 = torch.float8_e4m3fn
ACT_SCALE = 8.0            # activations are multiplied by 8 before rounding (post-ReLU values of O(1): keeps them normal)


def q16(x):
    return x.half().float()


def q8_act(x):
    return (x * ACT_SCALE).clamp(-448.0, 448.0).to(F8).float() / ACT_SCALE


def q8_weight(W):
    s = W.abs().amax(dim=1, keepdim=True).clamp_min(1e-12) / 448.0
    return (W / s).to(F8).float() * s


class Plan:
    """Which plain layers use fp8: the first `n8` in network order."""

    def __init__(self, n8):
        self.n8, self.seen = n8, 0

    def plain(self, x, lin):
        use8 = self.seen < self.n8
        self.seen += 1
        if use8:
            return torch.relu(q8_act(x) @ q8_weight(lin.weight).t() + lin.bias)
        return torch.relu(q16(x) @ q16(lin.weight).t() + lin.bias)


def skip_mlp(mod, lat, x, n_lat, plan):
    l1 = [m for m in mod.linears1 if isinstance(m, torch.nn.Linear)]
    l2 = [m for m in mod.linears2 if isinstance(m, torch.nn.Linear)]
    W0 = l1[0].weight
    h = torch.relu(q16(x) @ q16(W0[:, n_lat:]).t() + l1[0].bias + lat @ W0[:, :n_lat].t())
    for m in l1[1:]:
        h = plan.plain(h, m)
    W, Wd = l2[0].weight, x.shape[1]
    h = torch.relu(q16(x) @ q16(W[:, n_lat:n_lat + Wd]).t() + q16(h) @ q16(W[:, n_lat + Wd:]).t() + l2[0].bias +
                   lat @ W[:, :n_lat].t())
    for m in l2[1:]:
        h = plan.plain(h, m)
    return h


def make_forward(n8):
    def fwd(net, emb, shp, emb_dirs, tex):
        if net.W != 1024:                    # the coarse net is fp32-class in the engine (split precision)
            return net(emb, shp, emb_dirs, tex)
        plan = Plan(n8)
        n_pe = emb.shape[1] - 30
        xl = [m for m in net.xyzEncode.linears1 if isinstance(m, torch.nn.Linear)]
        W0 = xl[0].weight
        h = torch.relu(q16(emb[:, :n_pe]) @ q16(W0[:, :n_pe]).t() + xl[0].bias + emb[:, n_pe:] @ W0[:, n_pe:].t())
        for m in xl[1:]:
            h = plan.plain(h, m)
        sigma = skip_mlp(net.linear_BiM_xyz, shp, h, shp.shape[1], plan)
        alpha = sigma @ net.alpha_linear[0].weight.t() + net.alpha_linear[0].bias
        rgbc = skip_mlp(net.linear_uv_xyzBiM, tex, sigma, tex.shape[1], plan)
        Wv, nv = net.linear_view_xyBMuv[0].weight, emb_dirs.shape[1]
        hv = torch.relu(q16(emb_dirs) @ q16(Wv[:, :nv]).t() + q16(rgbc) @ q16(Wv[:, nv:]).t() + net.linear_view_xyBMuv[0].bias)
        rgb = hv @ net.rgb_linear.weight.t() + net.rgb_linear.bias
        assert plan.seen == 19, plan.seen
        return torch.cat([rgb, alpha], -1)
    return fwd


def main():
    torch.set_num_threads(os.cpu_count())
    meta, inp, gold = load_case("cfg4_800_exp9")
    n = 96
    c, f, s = build_case_nets(meta)
    rays = O.make_ray_batch(inp["rays_o"][:n], inp["rays_d"][:n], 8.0, 26.0)
    em = O.expression_mod(s, inp["shape"], inp["exp"])
    rows = []
    with torch.no_grad():
        for n8 in (0, 1, 2, 4, 8, 12, 19):
            out = O.render_rays(rays, c, f, inp["shape"], em, inp["tex"], forward_fn=make_forward(n8))
            d = (out["rgb_map"] - gold["rgb_map"][:n]).abs()
            rows.append({"fp8_layers": n8, "rgb_max": d.max().item(), "rgb_mean": d.mean().item(),
                         "psnr_db": O.psnr(out["rgb_map"], gold["rgb_map"][:n])})
            print(rows[-1], flush=True)
    doc = {"what": "PSNR(rendered rgb, unmodified reference) vs number of the 19 plain 1024->1024 fine layers emulated with fp8 "
                   "e4m3 operands (per-channel weight scales, activations x8, fp32 accumulate); other fine layers fp16, coarse "
                   "pass exact; 96 rays of the cfg4_800_exp9 crop (800x800 frame, W_f = 1024); tools/fp8_parity_study.py",
           "stated_tolerance": "max 3e-2, mean 1e-3, PSNR >= 50 dB (tests/test_gpu_render.py)", "rows": rows}
    json.dump(doc, open(os.path.join(ROOT, "profiles", "r02_fp8_parity_study.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
