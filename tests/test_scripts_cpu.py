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


EXPRESSIONS = ["neutral", "smile", "mouth_stretch", "anger", "jaw_left", "jaw_right", "jaw_forward", "mouth_left",
               "mouth_right", "dimpler", "chin_raiser", "lip_puckerer", "lip_funneler", "sadness", "lip_roll", "grin",
               "cheek_blowing", "eye_closed", "brow_raiser", "brow_lower"]       # render_refine_trainSet.py:147-150


def _write_facescape_like_dataset(root, person="1", n_train=2000, hw=32, all_images=False):
    """A synthetic stand-in for the FaceScape multi-view set the scripts read (run_train.py:26-90,
    render_refine_trainSet.py:38-108): transforms_{train,val,test}_<id>.json with 100 views x 20 expressions for one
    person, ONE real image file (the loaders open only the first to learn H, W), and the directory entries
    getValidPerson() looks up by name (render_refine_trainSet.py:131-143)."""
    import json
    from PIL import Image
    ds = root / "ds"
    ds.mkdir(parents=True)
    for name in [person, "39", "52", "69", "295", "307", "413", "417", "587", "237", "353", "356", "440", "363"]:
        (ds / name).mkdir(exist_ok=True)
    (ds / person / "neutral").mkdir()
    Image.fromarray((np.random.RandomState(1).rand(hw, hw, 4) * 255).astype(np.uint8)).save(ds / person / "neutral" / "000.png")
    from oracle import mofa_oracle as O

    def frames(n):
        out = []
        for k in range(n):
            c2w = O.pose_spherical(-60.0 + 1.2 * (k % 100), 0.0, 16.0).numpy().tolist()
            out.append({"file_path": f"/{person}/{EXPRESSIONS[(k // 100) % 20]}/{k % 100:03d}", "transform_matrix": c2w,
                        "expression": (k // 100) % 20})
        return out
    for split, n in (("train", n_train), ("val", 1), ("test", 1)):
        with open(ds / f"transforms_{split}_{person}.json", "w") as fp:
            json.dump({"camera_angle_x": 0.45, "frames": frames(n)}, fp)
    if all_images:        # run_train.py reads the image of whichever training frame it draws (:278-281)
        for k, fr in enumerate(frames(n_train)):
            f = ds / (fr["file_path"][1:] + ".png")
            f.parent.mkdir(parents=True, exist_ok=True)
            if not f.exists():
                Image.fromarray((np.random.RandomState(10 + k).rand(hw, hw, 4) * 255).astype(np.uint8)).save(f)
    return ds


def _common_cfg(tmp_path, ds, extra=""):
    cfg = tmp_path / "cfg.txt"
    cfg.write_text(f"""expname = plumb
basedir = {tmp_path}/logs
datadir = {ds}
dataset_type = blender
personList = 1
no_batching = True
netchunk = 4096
chunk = 4096
use_viewdirs = True
white_bkgd = False
N_samples = 16
N_importance = 16
N_rand = 64
netwidth = 256
netwidth_fine = 256
netdepth_fine = 8
testskip = 1
input_ch_shapeCodes = 50
input_ch_textureCodes = 256
input_ch_expCodes = 30
lrate = 5e-5
{extra}""")
    return cfg


def test_render_refine_trainset_runs_unchanged_under_the_launcher(tmp_path, monkeypatch):
    """BASELINE config #5's script, UNCHANGED (render_refine_trainSet.py:146-311): one identity of a synthetic
    FaceScape-like set -> 10 expressions x 8 views through render_path -> render -> texture encoder -> (fake) engine ->
    80 PNG files in the directory layout the refine-net training set expects.  The script has no exit condition of its
    own other than running out of identities (it indexes images[i] for 300 persons): the synthetic set has one, so the
    run ends with the IndexError of person #3 — after person #1 is complete and person #2's slot was skipped as done."""
    from mofanerf_b200 import launch, renderer
    from tests.fake_engine import FakeEngine
    ref_loader.load()
    fake = FakeEngine()
    monkeypatch.setattr(renderer, "get_engine", lambda device=None: fake)
    monkeypatch.setattr(renderer.B200Renderer, "_to_device", lambda self, rays: rays)
    monkeypatch.setattr(torch, "set_default_tensor_type", lambda t: None)
    ds = _write_facescape_like_dataset(tmp_path)
    work = tmp_path / "work"
    (work / "data").mkdir(parents=True)
    np.save(work / "data" / "factors_id.npy", np.random.RandomState(2).randn(4, 50).astype(np.float32) * 0.03)   # load_bmData (:126)
    cfg = _common_cfg(tmp_path, ds)
    # the script reads every identity's UV map from an absolute path of its authors' machine (:287): serve it from memory
    launch.shim_missing_modules()
    import imageio
    real_imread = imageio.imread
    uv = (np.random.RandomState(3).rand(64, 64, 3) * 255).astype(np.uint8)
    monkeypatch.setattr(imageio, "imread", lambda p, *a, **k: uv if str(p).startswith("/data/myNerf/") else real_imread(p, *a, **k))
    cwd, argv = os.getcwd(), list(sys.argv)
    import models.render_class as rc
    orig_renderer = rc.myRenderer
    os.chdir(work)
    try:
        with pytest.raises(IndexError):
            launch.main(["--shim-missing", "--no-chdir", os.path.join(ref_loader.REF_ROOT, "render_refine_trainSet.py"),
                         "--config", str(cfg)])
    finally:
        os.chdir(cwd)
        sys.argv = argv
        rc.myRenderer = orig_renderer
    out = tmp_path / "logs" / "plumb_1" / "renderonly_path_000000" / "rf_trainSet" / "train" / "1"
    exps = sorted(p.name for p in out.iterdir())
    assert len(exps) == 10 and set(exps) <= set(EXPRESSIONS)                  # num_exp_type (:246)
    pngs = sorted(out.glob("*/*.png"))
    assert len(pngs) == 80                                                     # x num_images_per_exp (:247)
    from PIL import Image
    assert Image.open(pngs[0]).size == (16, 16)                                # half_res of the 32x32 set (:167)
    rr = [c for c in fake.calls if c[0] == "render_rays"]
    assert len(rr) == 80 and all(c[1] == 256 and c[2] == 16 and c[3] == 16 and c[4] == 0.0 for c in rr)
    assert sum(1 for c in fake.calls if c == ("set_latents", 50, 30, 256)) == 80   # a new identity / expression / texture per image
    lines = (tmp_path / "logs" / "plumb_1" / "renderonly_path_000000" / "renderImageList.txt").read_text().splitlines()
    assert len(lines) == 80


def test_run_train_iterations_run_unchanged_under_the_launcher(tmp_path, monkeypatch):
    """run_train.py UNCHANGED (:160-405) for two optimisation steps on a synthetic FaceScape-like set: data loading,
    create_nerf, DataParallel wrapping (:253-257), landmark-guided ray sampling, render(rays=..., uvMap=..., retraw=True,
    perturb=1) in TRAINING mode through the drop-in renderer's autograd node, the rgb + rgb0 loss, backward, Adam step.
    The engine is the recording fake (no GPU here): it hands back constant gradients, so what is checked is that the
    script's loss reaches the engine's backward with weight gradients requested for both networks, and that the
    optimiser then moves the NeRF weights, the texture encoder, the StyleModule and the expression code it used."""
    from mofanerf_b200 import launch, renderer
    from tests.fake_engine import FakeEngine
    ref_loader.load()
    fake = FakeEngine()
    fake.stop_after_train_renders = 2
    monkeypatch.setattr(renderer, "get_engine", lambda device=None: fake)
    monkeypatch.setattr(renderer.B200Renderer, "_to_device", lambda self, rays: rays)
    monkeypatch.setattr(torch, "set_default_tensor_type", lambda t: None)
    ds = _write_facescape_like_dataset(tmp_path, n_train=4, all_images=True)
    work = tmp_path / "work" / "repo"                      # the script reads ../data/... relative to its working directory
    work.mkdir(parents=True)
    data = tmp_path / "work" / "data"
    (data / "textureMap300" / "1").mkdir(parents=True)
    np.save(data / "factors_id.npy", np.random.RandomState(2).randn(4, 50).astype(np.float32) * 0.03)      # load_bmData (:120)
    np.save(data / "1_975_landmarks.npy", np.random.RandomState(4).randn(4, 20, 68, 3).astype(np.float32))    # LMModule (:126)
    from PIL import Image
    Image.fromarray((np.random.RandomState(3).rand(64, 64, 3) * 255).astype(np.uint8)).save(data / "textureMap300" / "1" / "1_neutral.jpg")
    cfg = _common_cfg(tmp_path, ds, extra="i_print = 1\n")
    captured = {}
    import tools.create_model_condition as cmc
    real_create = cmc.create_nerf

    def spy_create(args):
        res = real_create(args)
        captured["train_kwargs"], captured["render"], captured["optimizer"] = res[0], res[6], res[4]
        captured["before"] = {
            "coarse": res[0]["network_fn"].alpha_linear[0].weight.detach().clone(),
            "fine": res[0]["network_fine"].alpha_linear[0].weight.detach().clone(),
            "tex": next(res[6].texEncoder.parameters()).detach().clone(),
            "style": next(res[6].idSpecificMod.parameters()).detach().clone(),
            "exp": [c.detach().clone() for c in res[6].expCodes_Sigma],
        }
        return res
    monkeypatch.setattr(cmc, "create_nerf", spy_create)
    cwd, argv = os.getcwd(), list(sys.argv)
    import models.render_class as rc
    orig_renderer = rc.myRenderer
    os.chdir(work)
    try:
        with pytest.raises(FakeEngine.Stop):
            launch.main(["--shim-missing", "--no-chdir", os.path.join(ref_loader.REF_ROOT, "run_train.py"),
                         "--config", str(cfg)])
    finally:
        os.chdir(cwd)
        sys.argv = argv
        rc.myRenderer = orig_renderer
    fwd = [c for c in fake.calls if c[0] == "render_rays"]
    bwd = [c for c in fake.calls if c[0] == "render_rays_bwd"]
    assert len(fwd) == 2 and all(c[-1] == "train" and c[1] == 64 and c[2] == 16 and c[3] == 16 and c[4] == 1.0 for c in fwd)
    # both losses (rgb and rgb0, run_train.py:341-346) reach the backward; weight gradients of both nets are requested
    assert len(bwd) == 2 and all(c[2] and c[3] and c[4] > 90 for c in bwd), bwd
    kw, rend, before = captured["train_kwargs"], captured["render"], captured["before"]
    assert type(rend).__name__ == "B200Renderer"
    coarse = getattr(kw["network_fn"], "module", kw["network_fn"])
    fine = getattr(kw["network_fine"], "module", kw["network_fine"])
    assert not torch.equal(coarse.alpha_linear[0].weight, before["coarse"])
    assert not torch.equal(fine.alpha_linear[0].weight, before["fine"])
    assert not torch.equal(next(rend.texEncoder.parameters()), before["tex"])
    assert not torch.equal(next(rend.idSpecificMod.parameters()), before["style"])
    assert any(not torch.equal(a, b) for a, b in zip(rend.expCodes_Sigma, before["exp"]))
    assert (tmp_path / "logs" / "plumb_1" / "args.txt").exists()


def test_run_fit_fitting_iterations_run_unchanged_under_the_launcher(tmp_path, monkeypatch):
    """BASELINE config #3's script flow, UNCHANGED: `run_fit.py --renderType fitting --num_iterations 6` (:257-350; the
    script divides by num_iterations // 6) — seven iterations of landmark-guided ray sampling from a differentiable pose (get_rays_withGrad), render_fitting with
    requires_grad pose / shape / texture / expression codes through the drop-in renderer's autograd node, L1 loss, three Adam
    optimisers; plus the parameter file and the preview image the script writes at iteration 0.  The recording fake stands
    in for the engine (constant gradients): checked is that the nets stay in eval mode (no weight gradients requested,
    models/render_class.py:383-384), that every iteration reaches the engine's backward, and what lands on disk."""
    from mofanerf_b200 import launch, renderer
    from tests.fake_engine import FakeEngine
    ref_loader.load()
    fake = FakeEngine()
    monkeypatch.setattr(renderer, "get_engine", lambda device=None: fake)
    monkeypatch.setattr(renderer.B200Renderer, "_to_device", lambda self, rays: rays)
    monkeypatch.setattr(torch, "set_default_tensor_type", lambda t: None)
    data = tmp_path / "data" / "segRelRes"
    data.mkdir(parents=True)
    from PIL import Image
    Image.fromarray((np.random.RandomState(0).rand(512, 512, 3) * 200 + 30).astype(np.uint8)).save(data / "00002.png")
    rs = np.random.RandomState(5)
    kp = np.stack([200 + 110 * rs.rand(68), 200 + 110 * rs.rand(68)], 1).astype(np.float32)      # 68 landmarks inside the face box
    from oracle import mofa_oracle as O
    np.save(tmp_path / "data" / "pose_00002.npy", {"pose": O.pose_spherical(10.0, 0.0, 16.0).numpy(), "kp": kp}, allow_pickle=True)
    cfg = _common_cfg(tmp_path, tmp_path / "data", extra="person_num = 300\n")
    cwd, argv = os.getcwd(), list(sys.argv)
    import models.render_class as rc
    orig_renderer = rc.myRenderer
    try:
        launch.main(["--shim-missing", os.path.join(ref_loader.REF_ROOT, "run_fit.py"), "--config", str(cfg),
                     "--filePath", str(data / "00002.png"), "--renderType", "fitting", "--num_iterations", "6"])
    finally:
        os.chdir(cwd)
        sys.argv = argv
        rc.myRenderer = orig_renderer
    fwd = [c for c in fake.calls if c[0] == "render_rays"]
    bwd = [c for c in fake.calls if c[0] == "render_rays_bwd"]
    train_fwd = [c for c in fwd if c[-1] == "train"]
    assert len(train_fwd) == 7 and all(c[1] == 64 and c[2] == 16 and c[3] == 16 and c[4] == 0.0 for c in train_fwd)
    assert len(bwd) == 7 and all(c[2] and not c[3] and c[4] == 0 for c in bwd), bwd      # rgb loss only, weights are constants
    out = tmp_path / "data" / "fitting" / "segRelRes_00002"
    assert (out / "target.png").exists()
    prev = out / "segRelRes_00002_0.png"                     # preview at iteration 0 (:329-347): 128 x 128 at the first scale
    assert prev.exists() and Image.open(prev).size == (128, 128)
    assert sum(c[1] for c in fwd if c[-1] != "train") == 128 * 128
    ck = torch.load(out / "saving_Parameters.tar", weights_only=False)
    assert set(ck) >= {"saving_bm", "saving_uv", "saving_exp", "saving_pose", "saving_global_light", "iter"}
    assert not torch.equal(ck["saving_global_light"].cpu(), torch.ones(2))          # the first Adam step moved it
    assert ck["saving_bm"].shape[-1] == 50 and ck["saving_uv"].numel() == 256 and ck["saving_exp"].numel() == 30
