"""Diagnosis of bench.py's `parity` leg (GPU box): which rays of the sampled bench frame differ from the oracle, and who is
right.  Renders every k-th ray of the H x W bench frame with
  (ref)   the oracle port in fp32 on the GPU (what bench.py's torch_gpu_baseline leg renders),
  (cpu)   the oracle port on the CPU, for the rays that disagree only,
  (a, a2) the engine, default settings, twice (run-to-run determinism),
  (b)     the engine with short passes (chunk_rays = 448: four slabs of the chain kernel per pass),
  (c)     the engine's SIMT verification GEMMs (no tcgen05 / chain kernel) on the disagreeing rays,
and prints per-ray details for the rays whose rgb differs by more than --thr.
    python tools/diag_parity.py [--H 800 --W 800 --rays 8192 --thr 0.02] > gpurun_out/diag_parity.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=800)
    ap.add_argument("--W", type=int, default=800)
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--thr", type=float, default=0.02)
    ap.add_argument("--n-samples", type=int, default=64)
    ap.add_argument("--n-importance", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0, help="seed of the nets (0 = bench.py's; 5 = the fixtures' dense field)")
    ap.add_argument("--sigma-gain", type=float, default=1.0,
                    help="scale alpha_linear (weight and bias) of both nets: a sharper density field, sigma further from 0")
    args = ap.parse_args()
    import bench
    from mofanerf_b200 import B200Renderer
    from oracle import mofa_oracle as O

    dev = torch.device("cuda:0")
    c, f, s = O.build_nets(args.seed)
    shape, tex, exp, ro, rd = bench.synth_inputs(args.H, args.W)
    n = min(args.rays, ro.shape[0])
    idx = torch.linspace(0, ro.shape[0] - 1, n).long()
    ro_s, rd_s = ro[idx], rd[idx]
    rays_cpu = O.make_ray_batch(ro_s, rd_s, 8.0, 26.0)
    cg, fg, sg = O.build_nets(args.seed)
    if args.sigma_gain != 1.0:
        with torch.no_grad():
            for net in (c, f, cg, fg):
                net.alpha_linear[0].weight.mul_(args.sigma_gain)
                net.alpha_linear[0].bias.mul_(args.sigma_gain)
    cg, fg, sg = cg.to(dev), fg.to(dev), sg.to(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad(), torch.device(dev):
        em_g = O.expression_mod(sg, shape.to(dev), exp.to(dev))
        ref = O.render_rays(rays_cpu.to(dev), cg, fg, shape.to(dev), em_g, tex.to(dev), N_samples=args.n_samples,
                            N_importance=args.n_importance, netchunk=196608, retraw=True)
        ref_nc = O.render_rays(rays_cpu.to(dev), cg, fg, shape.to(dev), em_g, tex.to(dev), N_samples=args.n_samples,
                               N_importance=args.n_importance, netchunk=65536)
        torch.backends.cuda.matmul.allow_tf32 = True
        ref_tf = O.render_rays(rays_cpu.to(dev), cg, fg, shape.to(dev), em_g, tex.to(dev), N_samples=args.n_samples,
                               N_importance=args.n_importance, netchunk=196608)
        torch.backends.cuda.matmul.allow_tf32 = False
    ref = {k: v.float().cpu() for k, v in ref.items()}
    ref_nc = {k: v.float().cpu() for k, v in ref_nc.items()}
    ref_tf = {k: v.float().cpu() for k, v in ref_tf.items()}

    r = B200Renderer(expCodesLen=30).to(dev)
    r.idSpecificMod.load_state_dict(s.state_dict())
    kw = dict(near=8.0, far=26.0, use_viewdirs=True, ndc=False, network_fn=cg, network_fine=fg,
              N_samples=args.n_samples, N_importance=args.n_importance, perturb=0.0, raw_noise_std=0.0)

    def engine(sel=None, chunk_rays=None, **extra):
        o, d = (ro_s, rd_s) if sel is None else (ro_s[sel], rd_s[sel])
        eng = r.engine(dev)
        old = eng.chunk_rays
        if chunk_rays is not None:
            eng.chunk_rays = chunk_rays
        try:
            with torch.no_grad():
                rgb, disp, acc, ex = r.render_fitting(1, o.shape[0], None, chunk=1 << 30, rays=(o.to(dev), d.to(dev)),
                                                      shapeCodes=shape.to(dev), uvCodes=tex.to(dev), expType=20,
                                                      expCodes=exp.to(dev), **dict(kw, **extra))
        finally:
            eng.chunk_rays = old
        out = {"rgb_map": rgb, "acc_map": acc, "disp_map": disp}
        out.update(ex)
        return {k: v.float().cpu() for k, v in out.items() if torch.is_tensor(v)}

    a = engine(want_aux=True, retraw=True)
    a2 = engine()
    b = engine(chunk_rays=448)
    err = (a["rgb_map"].reshape(-1, 3) - ref["rgb_map"]).abs().max(dim=1).values
    bad = torch.nonzero(err > args.thr).reshape(-1)
    sig_last = ref["raw"].reshape(n, -1, 4)[:, -1, 3]                      # reference pre-activation sigma of the last sample
    sig_min = ref["raw"].reshape(n, -1, 4)[..., 3].abs().min(dim=1).values
    flips = (a["acc_map"].reshape(-1) - ref["acc_map"]).abs() > 0.5
    stats = {
        "engine_vs_fp32": bench.parity_stats(a["rgb_map"], a["acc_map"], ref["rgb_map"], ref["acc_map"],
                                             a["raw"].reshape(n, -1, 4)[:, -1, 3], sig_last),
        "tf32_vs_fp32": bench.parity_stats(ref_tf["rgb_map"], ref_tf["acc_map"], ref["rgb_map"], ref["acc_map"]),
        "fp32_netchunk65536_vs_fp32": bench.parity_stats(ref_nc["rgb_map"], ref_nc["acc_map"], ref["rgb_map"], ref["acc_map"]),
        "engine_vs_tf32": bench.parity_stats(a["rgb_map"], a["acc_map"], ref_tf["rgb_map"], ref_tf["acc_map"]),
        "sigma_gain": args.sigma_gain, "net_seed": args.seed,
        "ref_acc_quantiles": {str(q): float(torch.quantile(ref["acc_map"], q)) for q in (0.01, 0.05, 0.25, 0.5)},
        "frac_ref_acc_below_0.3": float((ref["acc_map"] < 0.3).float().mean()),
        "bad_rays_with_ref_acc_below_0.3": int((ref["acc_map"][bad] < 0.3).sum()),
        "gate_flip_abs_sigma_last_ref": [float(x) for x in sig_last[flips].abs().tolist()],
        "abs_sigma_last_quantiles_all_rays": {str(q): float(torch.quantile(sig_last.abs(), q)) for q in (0.001, 0.01, 0.1, 0.5)},
        "rays_with_abs_sigma_last_below_0.05": int((sig_last.abs() < 0.05).sum()),
    }
    dsig = (a["raw"].reshape(n, -1, 4)[..., 3] - ref["raw"].reshape(n, -1, 4)[..., 3]).abs().reshape(-1)
    stats["abs_raw_sigma_err_engine_vs_fp32_quantiles"] = {str(q): float(torch.quantile(dsig[:: max(1, dsig.numel() // 1000000)], q))
                                                           for q in (0.5, 0.99, 0.999)}
    stats["abs_raw_sigma_err_engine_vs_fp32_max"] = float(dsig.max())
    stats["abs_raw_sigma_ref_median"] = float(ref["raw"].reshape(n, -1, 4)[..., 3].abs().median())

    rep = {"rays": n, "frame": [args.H, args.W], "thr": args.thr, "n_bad": int(bad.numel()), "stats": stats,
           "max_err": float(err.max()), "mean_err": float((a["rgb_map"].reshape(-1, 3) - ref["rgb_map"]).abs().mean()),
           "a_equals_a2": bool(torch.equal(a["rgb_map"], a2["rgb_map"])),
           "a_equals_b_chunk448": bool(torch.equal(a["rgb_map"], b["rgb_map"])),
           "max_a_minus_b": float((a["rgb_map"] - b["rgb_map"]).abs().max()),
           "err_quantiles": {q: float(torch.quantile(err, q)) for q in (0.5, 0.9, 0.99, 0.999)},
           "bad": []}
    if bad.numel() > 0:
        sel = bad[:64]
        with torch.no_grad():
            em = O.expression_mod(s, shape, exp)
            cpu = O.render_rays(rays_cpu[sel], c, f, shape, em, tex, N_samples=args.n_samples,
                                N_importance=args.n_importance)
        simt = engine(sel=sel, gemm_simt=True, want_aux=True)
        alone = engine(sel=sel, want_aux=True)
        zf_ref = ref["z_vals_fine"].reshape(n, -1)
        za = a["z_vals"].reshape(n, -1)
        for j, i in enumerate(sel.tolist()):
            dz = (za[i] - zf_ref[i]).abs()
            rep["bad"].append({
                "sample_index": i, "frame_ray": int(idx[i]), "pass_of_4096": i // 4096, "row_in_pass": i % 4096,
                "rgb_engine": a["rgb_map"].reshape(-1, 3)[i].tolist(), "rgb_ref_gpu": ref["rgb_map"][i].tolist(),
                "rgb_ref_cpu": cpu["rgb_map"][j].tolist(), "rgb_engine_simt": simt["rgb_map"].reshape(-1, 3)[j].tolist(),
                "rgb_engine_alone": alone["rgb_map"].reshape(-1, 3)[j].tolist(),
                "acc_engine": float(a["acc_map"].reshape(-1)[i]), "acc_ref_gpu": float(ref["acc_map"][i]),
                "acc_ref_cpu": float(cpu["acc_map"][j]), "acc_engine_simt": float(simt["acc_map"].reshape(-1)[j]),
                "acc0_engine": float(a["acc0"].reshape(-1)[i]), "acc0_ref": float(ref["acc0"][i]),
                "rgb0_err": float((a["rgb0"].reshape(-1, 3)[i] - ref["rgb0"][i]).abs().max()),
                "z_std_engine": float(a["z_std"].reshape(-1)[i]), "z_std_ref": float(ref["z_std"][i]),
                "max_dz_fine": float(dz.max()), "n_dz_gt_1e-3": int((dz > 1e-3).sum()),
                "z_engine_first8": za[i][:8].tolist(), "z_ref_first8": zf_ref[i][:8].tolist(),
                "w_engine_max": float(a["weights"].reshape(n, -1)[i].max()), "w_ref_max": float(ref["weights"][i].max()),
                "cpu_vs_gpu_ref_rgb": float((cpu["rgb_map"][j] - ref["rgb_map"][i]).abs().max()),
                "simt_vs_cpu_rgb": float((simt["rgb_map"].reshape(-1, 3)[j] - cpu["rgb_map"][j]).abs().max()),
                "engine_vs_cpu_rgb": float((a["rgb_map"].reshape(-1, 3)[i] - cpu["rgb_map"][j]).abs().max()),
                "alone_vs_cpu_rgb": float((alone["rgb_map"].reshape(-1, 3)[j] - cpu["rgb_map"][j]).abs().max()),
            })
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
