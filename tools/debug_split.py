"""Diagnostics for the split-precision coarse kernel (run on the GPU box): error structure by row / column."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mofa_oracle as O  # noqa: E402
from mofanerf_b200.engine import Engine  # noqa: E402

DEV = "cuda:0"
c, f, s = O.build_nets(5, 256, 8, 0, 10)
g = torch.Generator().manual_seed(3)
shape = torch.randn(1, 50, generator=g) * 0.034
tex = 0.14 + 0.26 * torch.randn(256, generator=g)
exp = torch.rand(1, 30, generator=g)
em = O.expression_mod(s, shape, exp)
n, S = int(os.environ.get("N", 600)), 64
ro = torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 16.0])
rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
z = torch.sort(8.0 + 18.0 * torch.rand(n, S, generator=g), -1)[0]
pts = ro[:, None] + rd[:, None] * z[..., None]
with torch.no_grad():
    ref = O.run_network(pts, rd, c, shape, em, tex).reshape(-1, 4)
for mode in ("split", "fp16"):
    if mode == "fp16":
        os.environ["MOFA_B200_COARSE_FP16"] = "1"
    eng = Engine(DEV)
    os.environ.pop("MOFA_B200_COARSE_FP16", None)
    eng.load_network(0, c.to(DEV))
    eng.set_latents(shape, em, tex)
    out = eng.run_network(0, pts.to(DEV), rd[:, None].to(DEV)).cpu().reshape(-1, 4)
    torch.cuda.synchronize()
    d = (out - ref).abs()
    print(f"[{mode}] P={d.shape[0]} max {d.max().item():.3e} mean {d.mean().item():.3e} | per column max {[f'{x:.2e}' for x in d.max(0)[0].tolist()]}"
          f" | nan {int(torch.isnan(out).sum())}")
    rows = d.max(1)[0]
    blk = rows[: (rows.shape[0] // 128) * 128].reshape(-1, 128).max(1)[0]
    print(f"[{mode}] per-128-row block max (first 12): {[f'{x:.1e}' for x in blk[:12].tolist()]}  last rows max {rows[-64:].max().item():.2e}")
    import time
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            eng.run_network(0, pts.to(DEV), rd[:, None].to(DEV))
        torch.cuda.synchronize()
    print(f"[{mode}] {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per run_network of {n * S} points")
    eng.close()
