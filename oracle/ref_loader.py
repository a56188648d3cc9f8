"""Import the UNMODIFIED reference (/root/reference) on a CPU-only box.  TEST INFRASTRUCTURE ONLY.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the (skipped-when-absent)
live cross-check in ``tests/test_oracle_golden.py``.

Shim: the reference imports ``imageio`` at module level
(models/render_class.py:5) and hard-calls ``.cuda()`` (render_class.py:54,83,88,328,453-471);
stub the first, make the second the identity.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("MOFA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "render_class.py"))


_loaded = None


def load():
    """Returns a namespace with the reference modules: .model, .render_class, .helpers."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    import torch

    if "imageio" not in sys.modules:
        sys.modules["imageio"] = types.ModuleType("imageio")
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import models.model as ref_model
    import models.render_class as ref_render
    import tools.run_nerf_helpers as ref_helpers

    torch.autograd.set_detect_anomaly(False)  # reference turns it on globally (model.py:4)
    ns = types.SimpleNamespace(model=ref_model, render_class=ref_render, helpers=ref_helpers)
    _loaded = ns
    return ns


def build_reference(seed=0, W_c=256, D_c=8, W_f=1024, D_f=10):
    """Construct (coarse, fine, renderer) exactly as tools/create_model_condition.py:16-50 does
    for configs/exp_mofanerf.txt, with torch.manual_seed(seed) first."""
    import torch

    ref = load()
    torch.manual_seed(seed)
    embed_fn, input_ch = ref.model.get_embedder(10, 0)
    embeddirs_fn, input_ch_views = ref.model.get_embedder(4, 0)
    kw = dict(input_ch_shapeCodes=50, input_ch_textureCodes=256, input_ch=input_ch + 30, output_ch=5,
              skips=[4], input_ch_views=input_ch_views, use_viewdirs=True)
    coarse = ref.model.NeRF(D=D_c, W=W_c, **kw)
    fine = ref.model.NeRF(D=D_f, W=W_f, **kw) if W_f else None
    # renderer construction order: texEncoder (EnDeUVmap) then StyleModule then 20 exp codes
    # (render_class.py:47-58).  We need the StyleModule's RNG position to equal the oracle's
    # build_nets(), which builds coarse, fine, style back to back; so build the style module
    # *first* from a forked generator state, then the renderer, and swap it in.
    state = torch.random.get_rng_state()
    style = ref.model.StyleModule()
    torch.random.set_rng_state(state)
    renderer = ref.render_class.myRenderer(embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=65536,
                                           uvCodesLen=256, expCodesLen=30)
    renderer.idSpecificMod = style
    for m in (coarse, fine, renderer):
        if m is not None:
            m.eval()
    return coarse, fine, renderer
