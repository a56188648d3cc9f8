"""Reference checkpoint I/O: the `.tar` files written by run_train.py:369-380 and read by
tools/create_model_condition.py:62-89, and the fitting state of run_fit.py:320-331, map one-to-one onto this package's
modules (identical state_dict keys), so loading is plain `load_state_dict`; the packed tensor-core layout is rebuilt by
the engine the next time the networks are used (cached on parameter versions)."""
from __future__ import annotations

import torch

from .nets import NeRFParams

TRAIN_KEYS = ("global_step", "network_fn_state_dict", "network_fine_state_dict", "network_render_textureEncoder",
              "network_render_idSpecific", "optimizer_state_dict", "expression_latent_codes_sigma")


def nets_from_state_dicts(coarse_sd, fine_sd=None, device="cpu"):
    """Build parameter containers whose (D, W) are read off the state_dicts themselves."""
    def build(sd):
        W = sd["xyzEncode.linears1.Linear0.weight"].shape[0]
        n2 = sum(1 for k in sd if k.startswith("linear_BiM_xyz.linears2.") and k.endswith(".weight"))
        net = NeRFParams(D=n2 + 5, W=W, input_ch=sd["xyzEncode.linears1.Linear0.weight"].shape[1],
                         input_ch_views=sd["linear_view_xyBMuv.0.weight"].shape[1] - W,
                         input_ch_textureCodes=sd["linear_uv_xyzBiM.linears1.Linear0.weight"].shape[1] - W,
                         input_ch_shapeCodes=sd["linear_BiM_xyz.linears1.Linear0.weight"].shape[1] - W)
        net.load_state_dict(sd)
        return net.to(device)
    return build(coarse_sd), (build(fine_sd) if fine_sd is not None else None)


def load_train_checkpoint(path_or_dict, renderer, device="cpu"):
    """-> (coarse, fine, global_step).  Restores the renderer's texture encoder, StyleModule and expression codes in place
    exactly as create_model_condition.py:78-88 does."""
    ck = torch.load(path_or_dict, map_location="cpu") if isinstance(path_or_dict, str) else path_or_dict
    coarse, fine = nets_from_state_dicts(ck["network_fn_state_dict"], ck.get("network_fine_state_dict"), device)
    renderer.texEncoder.load_state_dict(ck["network_render_textureEncoder"])
    renderer.idSpecificMod.load_state_dict(ck["network_render_idSpecific"])
    for latent, saved in zip(renderer.expCodes_Sigma, ck["expression_latent_codes_sigma"]):
        latent.data[:] = saved[:].detach().clone().to(latent.device)
    return coarse, fine, int(ck.get("global_step", 0))


def save_train_checkpoint(path, global_step, coarse, fine, renderer, optimizer=None):
    """Writes the dict run_train.py:369-380 writes (so the reference can read it back)."""
    unwrap = lambda m: getattr(m, "module", m)
    torch.save({
        "global_step": global_step,
        "network_fn_state_dict": unwrap(coarse).state_dict(),
        "network_fine_state_dict": unwrap(fine).state_dict(),
        "network_render_textureEncoder": unwrap(renderer.texEncoder).state_dict(),
        "network_render_idSpecific": unwrap(renderer.idSpecificMod).state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {},
        "expression_latent_codes_sigma": renderer.expCodes_Sigma,
    }, path)
