"""install(): the engine replaces models.render_class.myRenderer in an importable reference tree, and
tools/create_model_condition.create_nerf() then hands out a B200Renderer with the reference's own NeRF modules
(SURVEY.md §8b).  Needs /root/reference (build container only)."""
import sys
import types

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present on this box")


def test_install_swaps_renderer_and_create_nerf_uses_it(tmp_path):
    ref_loader.load()
    import mofanerf_b200
    import models.render_class as rc
    orig = rc.myRenderer
    ref_r = orig(expCodesLen=30)
    ref_tex_keys, ref_style_keys = list(ref_r.texEncoder.state_dict()), list(ref_r.idSpecificMod.state_dict())
    try:
        mofanerf_b200.install()
        assert rc.myRenderer is mofanerf_b200.B200Renderer
        sys.modules.setdefault("configargparse", types.ModuleType("configargparse"))
        import tools.create_model_condition as cmc
        args = types.SimpleNamespace(
            multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=64, N_samples=64, netdepth=8,
            netwidth=256, netdepth_fine=10, netwidth_fine=256, input_ch_shapeCodes=50, input_ch_textureCodes=256,
            input_ch_expCodes=30, device="cpu", netchunk=65536, lrate=5e-5, basedir=str(tmp_path), expname="e",
            ft_path=None, no_reload=True, perturb=1.0, white_bkgd=False, raw_noise_std=0.0, dataset_type="blender",
            no_ndc=False, lindisp=False)
        (tmp_path / "e").mkdir()
        train_kw, test_kw, start, grad_vars, opt, logger, render = cmc.create_nerf(args)
        assert isinstance(render, mofanerf_b200.B200Renderer)
        assert test_kw["perturb"] is False and test_kw["network_query_fn"] == render.run_network
        # the reference's own NeRF modules are accepted by the weight packer
        from mofanerf_b200.nets import canonical_tensors
        t, W, D = canonical_tensors(test_kw["network_fine"])
        assert (len(t), W, D) == (54, 256, 10)
        # checkpoint keys of the pieces callers save (run_train.py:369-380) match the reference modules
        assert list(render.texEncoder.state_dict()) == ref_tex_keys
        assert list(render.idSpecificMod.state_dict()) == ref_style_keys
        assert len(render.expCodes_Sigma) == 20 and render.expCodes_Sigma[0].shape == (1, 30)
    finally:
        rc.myRenderer = orig
