"""Fine-net chain kernel vs one launch per layer (run on the GPU box): bit-exactness and timing."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mofanerf_b200 import nets  # noqa: E402
from mofanerf_b200.engine import Engine  # noqa: E402

DEV = "cuda:0"
coarse, fine, style = nets.build_nets(0, device=DEV)
g = torch.Generator().manual_seed(3)
shape = torch.randn(50, generator=g) * 0.034
tex = 0.14 + 0.26 * torch.randn(256, generator=g)
em = torch.rand(30, generator=g)


def rays_for(n):
    gg = torch.Generator().manual_seed(n)
    ro = torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 16.0])
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=gg) * 0.1 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    return torch.cat([ro, rd, torch.full((n, 1), 8.0), torch.full((n, 1), 26.0), rd], -1).to(DEV)


res = {}
MODES = os.environ.get("MODES", "chain,per_layer").split(",")
for mode in MODES:
    if mode == "per_layer":
        os.environ["MOFA_B200_FINE_PER_LAYER"] = "1"
    eng = Engine(DEV)
    os.environ.pop("MOFA_B200_FINE_PER_LAYER", None)
    eng.load_network(0, coarse)
    eng.load_network(1, fine)
    eng.set_latents(shape, em, tex)
    for n in ((5, 75, 300, 1061) if len(MODES) == 2 else ()):
        out = eng.render_rays(rays_for(n), 64, 64, retraw=True)
        torch.cuda.synchronize()
        res[(mode, n)] = {k: v.clone() for k, v in out.items()}
        print(f"[{mode}] n={n} rgb mean {out['rgb_map'].mean().item():.6f} nan {int(torch.isnan(out['rgb_map']).sum())}", flush=True)
    n = int(os.environ.get("N", 8288))
    r = rays_for(n)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.render_rays(r, 64, 64)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"[{mode}] {n} rays: {dt * 1e3:.2f} ms -> {n / dt:.0f} rays/s", flush=True)
    eng.close()
for n in ((5, 75, 300, 1061) if len(MODES) == 2 else ()):
    a, b = res[("chain", n)], res[("per_layer", n)]
    same = {k: bool(torch.equal(a[k], b[k])) or bool(torch.allclose(a[k], b[k], equal_nan=True, atol=0, rtol=0)) for k in a}
    worst = max((a[k] - b[k]).abs().max().item() for k in ("rgb_map", "raw"))
    print(f"n={n}: bit-exact {same} worst |d| {worst:.3e}")
