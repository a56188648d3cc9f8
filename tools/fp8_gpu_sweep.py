"""On the GPU box: PSNR of the engine's FP8 variant against the reference fixture as a function of the number of e4m3
layers (MOFA_B200_FP8=n), next to the emulated prediction of tools/fp8_parity_study.py.
    python tools/fp8_gpu_sweep.py > gpurun_out/fp8_gpu_sweep.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mofa_oracle as O  # noqa: E402
from tests.helpers import build_case_nets, load_case  # noqa: E402
from mofanerf_b200.engine import Engine  # noqa: E402

DEV = "cuda:0"
meta, inp, gold = load_case("cfg4_800_exp9")
c, f, s = build_case_nets(meta)
n = 192
rays = torch.cat([O.make_ray_batch(inp["rays_o"][:n], inp["rays_d"][:n], 8.0, 26.0), torch.zeros(n, 1)], 1).to(DEV)
em = O.expression_mod(s, inp["shape"], inp["exp"])
c.to(DEV); f.to(DEV)
rows = []
for n8 in (0, 1, 2, 4, 8, 12, 19):
    os.environ["MOFA_B200_FP8"] = f"layers={n8}"
    eng = Engine(DEV)
    os.environ.pop("MOFA_B200_FP8", None)
    eng.load_network(0, c)
    eng.load_network(1, f)
    eng.set_latents(inp["shape"], em, inp["tex"])
    out = eng.render_rays(rays, 64, 64)["rgb_map"].cpu()
    torch.cuda.synchronize()
    eng.close()
    d = (out - gold["rgb_map"][:n]).abs()
    rows.append({"fp8_layers": n8, "rgb_max": d.max().item(), "rgb_mean": d.mean().item(), "psnr_db": O.psnr(out, gold["rgb_map"][:n])})
print(json.dumps({"what": "engine (GPU) FP8 variant vs the unmodified reference, 192 rays of cfg4_800_exp9, MOFA_B200_FP8=n", "rows": rows}, indent=1))
