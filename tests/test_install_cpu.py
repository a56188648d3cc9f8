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


def test_host_ray_helpers_match_the_reference_functions():
    """rays.get_rays / ndc_rays / pose_spherical (host path kept for NDC, explicit rays, c2w_staticcam) against the
    reference's own functions on fresh inputs (tools/run_nerf_helpers.py:153-199, tools/load_facescape.py:33-38)."""
    import numpy as np
    ref = ref_loader.load()
    from mofanerf_b200 import rays as R
    import tools.load_facescape as lf
    H, W, focal = 12, 9, 30.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    for angle in (-70.0, 15.0):
        c2w = lf.pose_spherical(angle, -10.0, 4.0)
        assert torch.equal(R.pose_spherical(angle, -10.0, 4.0), c2w)
        ro_r, rd_r = ref.helpers.get_rays(H, W, K, c2w[:3, :4])
        ro, rd = R.get_rays(H, W, K, c2w[:3, :4])
        assert torch.equal(ro, ro_r) and torch.equal(rd, rd_r)
        a = ref.helpers.ndc_rays(H, W, focal, 1.0, ro_r, rd_r)
        b = R.ndc_rays(H, W, focal, 1.0, ro, rd)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
