"""ctypes binding of the C ABI in include/mofa_b200.h (the only way Python reaches the CUDA engine).

There is deliberately no fallback: if the shared library is missing or no sm_100 GPU is present,
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_f32p = C.c_void_p  # device pointers are passed as integers


class RenderArgs(C.Structure):
    """mofa_b200_render_args (include/mofa_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("flags", C.c_uint32),
        ("rays", C.c_void_p), ("n_rays", C.c_int64),
        ("ray_stride", C.c_int32), ("n_samples", C.c_int32), ("n_importance", C.c_int32),
        ("run_fine", C.c_int32), ("fine_net", C.c_int32), ("chunk_rays", C.c_int32),
        ("perturb", C.c_float), ("raw_noise_std", C.c_float), ("seed", C.c_uint64),
        ("t_rand", C.c_void_p), ("u", C.c_void_p), ("noise_c", C.c_void_p), ("noise_f", C.c_void_p),
        ("rgb", C.c_void_p), ("disp", C.c_void_p), ("acc", C.c_void_p),
        ("rgb0", C.c_void_p), ("disp0", C.c_void_p), ("acc0", C.c_void_p), ("z_std", C.c_void_p),
        ("raw", C.c_void_p), ("weights", C.c_void_p), ("z_vals", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class BwdArgs(C.Structure):
    """mofa_b200_bwd_args (include/mofa_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("flags", C.c_uint32),
        ("rays", C.c_void_p), ("n_rays", C.c_int64),
        ("ray_stride", C.c_int32), ("n_samples", C.c_int32), ("n_importance", C.c_int32),
        ("run_fine", C.c_int32), ("fine_net", C.c_int32), ("reserved", C.c_int32),
        ("noise_c", C.c_void_p), ("noise_f", C.c_void_p),
        ("d_rgb", C.c_void_p), ("d_acc", C.c_void_p), ("d_rgb0", C.c_void_p), ("d_acc0", C.c_void_p),
        ("loss_scale", C.c_float), ("reserved_f", C.c_float),
        ("d_rays", C.c_void_p), ("d_shape", C.c_void_p), ("d_expmod", C.c_void_p), ("d_tex", C.c_void_p),
        ("d_params_coarse", C.c_void_p), ("d_params_fine", C.c_void_p),
        ("n_params_coarse", C.c_int32), ("n_params_fine", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("loss_scale_dev", C.c_void_p),
    ]


FLAG_LINDISP, FLAG_WHITE_BKGD, FLAG_GEMM_SIMT = 1, 2, 8
NET_COARSE, NET_FINE = 0, 1

# name -> (restype, argtypes); every symbol include/mofa_b200.h declares
SIGNATURES = {
    "mofa_b200_abi_version": (C.c_int, []),
    "mofa_b200_last_error": (C.c_char_p, []),
    "mofa_b200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mofa_b200_destroy": (C.c_int, [C.c_void_p]),
    "mofa_b200_load_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                         C.c_void_p]),
    "mofa_b200_packed_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "mofa_b200_export_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mofa_b200_import_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mofa_b200_set_latents": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mofa_b200_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "mofa_b200_render_rays_fwd": (C.c_int, [C.c_void_p, C.POINTER(RenderArgs), C.c_void_p]),
    "mofa_b200_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "mofa_b200_render_rays_train_fwd": (C.c_int, [C.c_void_p, C.POINTER(RenderArgs), C.c_void_p]),
    "mofa_b200_render_rays_bwd": (C.c_int, [C.c_void_p, C.POINTER(BwdArgs), C.c_void_p]),
    "mofa_b200_query_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "mofa_b200_run_network": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mofa_b200_generate_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                          C.c_float, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "mofa_b200_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "mofa_b200_raw2outputs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "mofa_b200_raw2outputs_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mofa_b200_sample_pdf_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mofa_b200_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_float, C.c_void_p,
                                  C.c_int, C.c_int, C.c_void_p]),
    "mofa_b200_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mofa_b200_launch_count": (C.c_int64, [C.c_void_p]),
    "mofa_b200_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mofa_b200_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """dlopen the in-tree library (built by mofanerf_b200.build / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `python -m mofanerf_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mofa_b200_abi_version() != 2:
        raise RuntimeError("libmofa_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("mofa_b200: " + load().mofa_b200_last_error().decode("utf-8", "replace"))
