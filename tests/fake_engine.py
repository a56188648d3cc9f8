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

    def render_rays(self, rays, N_samples, N_importance=0, **kw):
        n = rays.shape[0]
        self.calls.append(("render_rays", n, int(N_samples), int(N_importance), float(kw.get("perturb", 0.0)),
                           bool(kw.get("run_fine", True)), int(rays.shape[1])))
        rgb = 0.5 + 0.5 * rays[:, 8:11]
        out = {"rgb_map": rgb, "disp_map": rays[:, 8].abs(), "acc_map": torch.ones(n)}
        if N_importance > 0 and kw.get("run_fine", True):
            out.update(rgb0=rgb.clone(), disp0=out["disp_map"].clone(), acc0=out["acc_map"].clone(), z_std=torch.zeros(n))
        return out
