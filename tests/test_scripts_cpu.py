"""Script-level plumbing (BASELINE config #1): the UNCHANGED reference script
`run_fit.py --renderType rendering` runs under `python -m mofanerf_b200.launch --shim-missing` and renders its three views
through the drop-in renderer.  Needs /root/reference, so it runs in the build container only — which has no GPU: the CUDA
engine is replaced by tests/fake_engine.FakeEngine (the kernels themselves are covered by the `-m gpu` tests, which cannot
see the reference tree).  What is checked is the boundary: install(), create_nerf(), DataParallel-wrapped networks, the
keyword arguments run_fit.py passes to render_fitting, output shapes, the PNG files the script writes."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available() or torch.cuda.is_available(),
                                reason="needs the reference tree and a box without a GPU (the engine is faked)")


def test_launcher_shims_and_config_file(tmp_path, monkeypatch):
    from mofanerf_b200 import launch
    cfg = tmp_path / "c.txt"
    cfg.write_text("expname = abc #comment\nN_samples = 32\nuse_viewdirs = True\nno_batching = False\nlrate=5e-5\n")
    m = launch._configargparse_shim()
    p = m.ArgumentParser()
    p.add_argument("--config", is_config_file=True, default=str(cfg))
    p.add_argument("--expname", type=str, default="x")
    p.add_argument("--N_samples", type=int, default=64)
    p.add_argument("--use_viewdirs", action="store_true")
    p.add_argument("--no_batching", action="store_true")
    p.add_argument("--lrate", type=float, default=1.0)
    a = p.parse_args(["--N_samples", "16"])            # the command line overrides the file
    assert (a.expname, a.N_samples, a.use_viewdirs, a.no_batching, a.lrate) == ("abc", 16, True, False, 5e-5)
    io = launch._imageio_shim()
    img = (np.arange(48).reshape(4, 4, 3) * 5).astype(np.uint8)
    io.imwrite(str(tmp_path / "a.png"), img)
    assert np.array_equal(io.imread(str(tmp_path / "a.png")), img)


def test_run_fit_rendering_runs_unchanged_under_the_launcher(tmp_path, monkeypatch):
    from mofanerf_b200 import launch, renderer
    from tests.fake_engine import FakeEngine
    ref_loader.load()                                  # imageio stub + identity .cuda() on this GPU-less box
    fake = FakeEngine()
    monkeypatch.setattr(renderer, "get_engine", lambda device=None: fake)
    monkeypatch.setattr(renderer.B200Renderer, "_to_device", lambda self, rays: rays)
    monkeypatch.setattr(torch, "set_default_tensor_type", lambda t: None)      # 'torch.cuda.FloatTensor' needs a GPU
    # (ref_loader's imageio placeholder is an empty module that models.render_class already holds by reference: the
    # launcher fills such placeholders in place)
    data = tmp_path / "data" / "segRelRes"
    data.mkdir(parents=True)
    from PIL import Image
    Image.fromarray((np.random.RandomState(0).rand(512, 512, 3) * 255).astype(np.uint8)).save(data / "00001.png")
    np.save(tmp_path / "data" / "pose_00001.npy", {"pose": np.eye(4, dtype=np.float32), "kp": np.zeros((68, 2))},
            allow_pickle=True)
    cfg = tmp_path / "cfg.txt"
    cfg.write_text(f"""expname = plumb
basedir = {tmp_path}/logs
datadir = {tmp_path}/data
dataset_type = blender
person_num = 300
no_batching = True
netchunk = 4096
chunk = 4096
use_viewdirs = True
white_bkgd = False
N_samples = 32
N_importance = 0
N_rand = 64
netwidth = 256
netwidth_fine = 256
half_res = False
input_ch_shapeCodes = 50
input_ch_textureCodes = 256
input_ch_expCodes = 30
lrate = 5e-5
""")
    cwd = os.getcwd()
    argv = list(sys.argv)
    saved_modules = {k: sys.modules.get(k) for k in ("models.render_class",)}
    import models.render_class as rc
    orig_renderer = rc.myRenderer
    try:
        launch.main(["--shim-missing", os.path.join(ref_loader.REF_ROOT, "run_fit.py"), "--config", str(cfg),
                     "--filePath", str(data / "00001.png"), "--renderType", "rendering"])
    finally:
        os.chdir(cwd)
        sys.argv = argv
        rc.myRenderer = orig_renderer
    out = tmp_path / "data" / "fitting" / "segRelRes_00001" / "render"
    for angle in (-60, 0, 60):                          # run_fit.py:364-377
        f = out / f"fitRes_{angle}.png"
        assert f.exists(), f
        assert Image.open(f).size == (256, 256)
    kinds = [c[0] for c in fake.calls]
    assert kinds.count("render_rays") >= 3 and "load_network" in kinds and "set_latents" in kinds
    rr = [c for c in fake.calls if c[0] == "render_rays"]
    assert sum(c[1] for c in rr) == 3 * 256 * 256       # three 256x256 views (chunk = 4096 rays per render_rays call)
    assert all(c[2] == 32 and c[3] == 0 and c[4] == 0.0 and c[6] == 12 for c in rr)
    assert ("set_latents", 50, 30, 256) in fake.calls
    assert any(c[0] == "load_network" and c[2] == "DataParallel" for c in fake.calls)     # run_fit.py:166-167
