"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU, exports every
symbol include/mofa_b200.h declares, and the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mofa_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mofa_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from mofanerf_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mofa_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)
    assert _lib.load().mofa_b200_abi_version() == 2


def test_render_args_struct_matches_header_layout():
    from mofanerf_b200 import _lib
    # 2*u32, ptr, i64, 6*i32, 2*f32, u64, 14 pointers, ptr, size_t  (LP64)
    assert ctypes.sizeof(_lib.RenderArgs) == 8 + 8 + 8 + 24 + 8 + 8 + 14 * 8 + 8 + 8
    # 2*u32, ptr, i64, 6*i32, 6 pointers, 2*f32, 4 pointers, 2 pointers, 2*i32, ptr, size_t, ptr (loss_scale_dev)
    assert ctypes.sizeof(_lib.BwdArgs) == 8 + 8 + 8 + 24 + 6 * 8 + 8 + 4 * 8 + 2 * 8 + 8 + 8 + 8 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure():
    from mofanerf_b200 import B200Renderer, Engine, _lib
    with pytest.raises(RuntimeError, match="no CPU path"):
        Engine()
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.mofa_b200_create(ctypes.byref(h), 0) != 0
    assert b"no CUDA device" in lib.mofa_b200_last_error()
    r = B200Renderer(expCodesLen=30)
    from mofanerf_b200 import nets
    c, f, _ = nets.build_nets(0, 256, 8, 256, 10)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.render_fitting(2, 2, None, rays=(torch.zeros(4, 3), torch.ones(4, 3)), shapeCodes=torch.zeros(1, 50),
                         uvCodes=torch.zeros(256), expType=20, expCodes=torch.zeros(1, 30), near=8., far=26.,
                         use_viewdirs=True, ndc=False, network_fn=c, network_fine=f, N_samples=8, N_importance=8)


def test_param_containers_have_reference_state_dict_layout():
    from mofanerf_b200 import nets
    from oracle import mofa_oracle as O
    c, f, s = nets.build_nets(0, 256, 8, 256, 10)
    oc, of, os_ = O.build_nets(0, 256, 8, 256, 10)
    for a, b in ((c, oc), (f, of), (s, os_)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
    t, W, D = nets.canonical_tensors(torch.nn.DataParallel(f))
    assert (len(t), W, D) == (54, 256, 10)
    t, W, D = nets.canonical_tensors(c)
    assert (len(t), W, D) == (46, 256, 8)
    with pytest.raises(NotImplementedError):
        c(torch.zeros(1, 93), torch.zeros(1, 50), torch.zeros(1, 27), torch.zeros(1, 256))


def test_host_ray_generation_matches_fixture():
    import numpy as np
    from mofanerf_b200.rays import get_rays, pack_rays, pose_spherical
    z = np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))
    ro, rd = get_rays(6, 10, z["rays_K"], torch.from_numpy(z["rays_c2w"])[:3, :4])
    assert torch.equal(ro, torch.from_numpy(z["rays_o"])) and torch.equal(rd, torch.from_numpy(z["rays_d"]))
    assert torch.equal(pose_spherical(30.0, 0.0, 16.0), torch.from_numpy(z["rays_c2w"]))
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    rays = pack_rays(ro, rd, 8.0, 26.0, vd.reshape(-1, 3))
    # 11 reference columns + one zero pad column (16-byte rows for 128-bit loads)
    assert rays.shape == (60, 12) and float(rays[0, 6]) == 8.0 and float(rays[0, 7]) == 26.0 and float(rays[:, 11].abs().max()) == 0.0
