"""torch.autograd bridge for fitting (run_fit.py:305-313).

Differentiable: the packed ray batch [N,11] (so pose / ray origins / directions / view directions, through
whatever PyTorch graph built it), the shape code, the modulated expression code and the texture code.
Network weights: constants for fitting (run_fit.py does not optimise them: the nets are in eval() mode there); when the
nets are in train() mode (run_train.py) their parameters are passed as extra inputs and receive gradients (f2).
Constants: sample depths (the reference detaches
z_samples, models/render_class.py:326).  disp_map / z_std are returned as non-differentiable.
"""
from __future__ import annotations

import math

import torch


class RenderRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays, shape, exp_mod, tex, engine, cfg, *params):
        # outputs the loss does not use arrive in backward as None instead of zero tensors: run_fit.py's loss is on
        # rgb_map only (:309), so the whole coarse-pass backward (rgb0 / acc0) is skipped there — the engine skips a pass
        # whose upstream gradients are both absent
        ctx.set_materialize_grads(False)
        engine.set_latents(shape, exp_mod, tex)
        out = engine.render_rays(rays.detach(), cfg["N_samples"], cfg["N_importance"], run_fine=cfg["run_fine"],
                                 fine_net=cfg["fine_net"], perturb=cfg["perturb"], raw_noise_std=cfg["raw_noise_std"],
                                 lindisp=cfg["lindisp"], white_bkgd=cfg["white_bkgd"], retraw=cfg["retraw"],
                                 seed=cfg["seed"], t_rand=cfg.get("t_rand"), u=cfg.get("u"), noise_c=cfg.get("noise_c"),
                                 noise_f=cfg.get("noise_f"), want_aux=cfg.get("want_aux", False), train=True)
        if cfg["raw_noise_std"] > 0 and cfg.get("noise_c") is None:
            raise NotImplementedError("training-mode render with in-kernel sigma noise: pass explicit noise tensors")
        ctx.engine, ctx.cfg = engine, cfg
        # backward re-uses the engine's packed weights and folded latents: remember which ones this forward used
        ctx.latents = tuple(t.detach().clone() for t in (shape, exp_mod, tex))
        ctx.net_keys = dict(engine._net_keys)
        ctx.param_shapes = [tuple(p.shape) for p in params]
        ctx.n_coarse = cfg.get("n_params_coarse", 0)
        ctx.saved = {k: out.pop(k) for k in ("_train_ws", "_rays", "_noise")}
        ctx.fine = "rgb0" in out
        keys = ["rgb_map", "acc_map", "disp_map"] + (["rgb0", "acc0", "disp0", "z_std"] if ctx.fine else [])
        extra = [k for k in out if k not in keys]
        ctx.keys = keys + extra
        ctx.mark_non_differentiable(*[out[k] for k in ctx.keys if k not in ("rgb_map", "acc_map", "rgb0", "acc0")])
        return tuple(out[k] for k in ctx.keys)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.saved is None:
            raise RuntimeError("mofanerf_b200: backward called twice on the same render (activations are released after "
                               "the first backward; retain_graph is not supported)")
        eng = ctx.engine
        for which in (0, 1):
            if ctx.net_keys[which] is not None and eng._net_keys[which] != ctx.net_keys[which]:
                raise RuntimeError("mofanerf_b200: network weights were changed or replaced between forward and backward of "
                                   "a render (load_network / optimizer step): the saved activations no longer match")
        eng.set_latents(*ctx.latents)     # another render may have folded other codes since the forward
        g = dict(zip(ctx.keys, grads))
        d_rgb, d_acc = g.get("rgb_map"), g.get("acc_map")
        d_rgb0, d_acc0 = g.get("rgb0"), g.get("acc0")
        # power-of-two loss scale: largest upstream gradient -> 2^13 in the fp16 inter-layer gradients.  Per-sample
        # gradients are ~1e-3 .. 1e-4 of the largest one (compositing weights), and fp16 keeps full precision only above
        # 6e-5; the epilogue clamps at +-65504, so an outlier saturates instead of overflowing.
        # Computed ON THE DEVICE ({scale, 1/scale} -> mofa_b200_bwd_args.loss_scale_dev): the host never waits for the GPU
        # here (round 1 did `float(tensor.max())` per upstream gradient: four synchronisations per backward).
        ups = [x.detach().abs().max() for x in (d_rgb, d_acc, d_rgb0, d_acc0) if x is not None]
        if ups:
            amax = torch.stack(ups).max().float()
            sc = torch.exp2(torch.round(torch.log2(8192.0 / amax.clamp_min(1e-30)))).clamp(2.0 ** -20, 2.0 ** 40)
            sc = torch.where((amax > 0) & torch.isfinite(amax), sc, torch.ones_like(sc))
            scale = torch.stack([sc, 1.0 / sc]).contiguous()
        else:
            scale = 1.0
        cfg = ctx.cfg
        grads_p, pg = [], None
        if ctx.param_shapes:      # training: weight gradients of the coarse (and fine) network, canonical order
            dev = ctx.saved["_rays"].device
            grads_p = [torch.zeros(s, dtype=torch.float32, device=dev) for s in ctx.param_shapes]
            nc = ctx.n_coarse
            pg = (grads_p[:nc], grads_p[nc:] if len(grads_p) > nc else None)
        d_rays, d_shape, d_exp, d_tex = ctx.engine.render_rays_bwd(
            ctx.saved, cfg["N_samples"], cfg["N_importance"], run_fine=cfg["run_fine"], fine_net=cfg["fine_net"],
            white_bkgd=cfg["white_bkgd"], lindisp=cfg["lindisp"], d_rgb=d_rgb, d_acc=d_acc,
            d_rgb0=d_rgb0 if ctx.fine else None, d_acc0=d_acc0 if ctx.fine else None, loss_scale=scale,
            param_grads=pg)
        n_cols = ctx.saved["_rays"].shape[1]
        if n_cols > 11:    # padded ray rows: the pad columns carry no gradient
            d_rays = torch.cat([d_rays, d_rays.new_zeros(d_rays.shape[0], n_cols - 11)], 1)
        ctx.saved = None   # release the activation workspace
        return (d_rays, d_shape, d_exp, d_tex, None, None) + tuple(grads_p)
