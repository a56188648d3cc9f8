"""A recording stand-in for mofanerf_b200.engine.Engine — TEST INFRASTRUCTURE ONLY (tests/test_scripts_cpu.py).

The product has no CPU path; the script-level plumbing test runs where the reference tree exists (a box without a GPU), so
it swaps the engine for this object to check everything AROUND the kernels: that the unchanged script constructs the
renderer through install(), hands it its networks and latents, calls render_fitting with arguments the drop-in accepts,
and gets back maps of the shapes it goes on to write to disk.  Outputs are a deterministic function of the rays."""
import torch


class FakeEngine:
    def __init__(self):
        self.calls = []
        self.chunk_rays = 0
        self._net_keys = {0: None, 1: None}
        self.launch_count = 0

    def load_network(self, which, net, force=False):
        mod = getattr(net, "module", net)
        self.calls.append(("load_network", which, type(net).__name__, sum(p.numel() for p in mod.parameters())))

    def set_latents(self, shape, exp_mod, tex):
        self.calls.append(("set_latents", int(shape.numel()), int(exp_mod.numel()), int(tex.numel())))

    class Stop(Exception):
        """Raised by the training-mode render once `stop_after_train_renders` of them have run (run_train.py's loop has
        600 001 iterations and no other way out)."""

    stop_after_train_renders = None

    def render_rays(self, rays, N_samples, N_importance=0, train=False, retraw=False, **kw):
        n = rays.shape[0]
        if train:
            done = sum(1 for c in self.calls if c[0] == "render_rays" and c[-1] == "train")
            if self.stop_after_train_renders is not None and done >= self.stop_after_train_renders:
                raise FakeEngine.Stop()
        self.calls.append(("render_rays", n, int(N_samples), int(N_importance), float(kw.get("perturb", 0.0)),
                           bool(kw.get("run_fine", True)), int(rays.shape[1])) + (("train",) if train else ()))
        rgb = 0.5 + 0.5 * rays[:, 8:11]
        out = {"rgb_map": rgb, "disp_map": rays[:, 8].abs(), "acc_map": torch.ones(n)}
        fine = N_importance > 0 and kw.get("run_fine", True)
        if fine:
            out.update(rgb0=rgb.clone(), disp0=out["disp_map"].clone(), acc0=out["acc_map"].clone(), z_std=torch.zeros(n))
        if retraw:
            out["raw"] = torch.zeros(n, int(N_samples) + (int(N_importance) if fine else 0), 4)
        if train:      # what mofanerf_b200.autograd.RenderRaysFn keeps for the backward pass
            out.update(_train_ws=torch.zeros(1), _rays=rays, _noise=(None, None))
        return out

    def render_rays_bwd(self, saved, N_samples, N_importance, *, run_fine, fine_net, white_bkgd, lindisp, d_rgb=None,
                        d_acc=None, d_rgb0=None, d_acc0=None, loss_scale=1.0, param_grads=None):
        n = saved["_rays"].shape[0]
        n_par = 0
        if param_grads is not None:          # (coarse list, fine list) of zero-initialised gradient tensors
            for lst in param_grads:
                for t in (lst or []):
                    t.add_(1e-3)
                    n_par += 1
        self.calls.append(("render_rays_bwd", n, d_rgb is not None, d_rgb0 is not None, n_par))
        g = 1e-3
        return torch.full((n, 11), g), torch.full((50,), g), torch.full((30,), g), torch.full((256,), g)
