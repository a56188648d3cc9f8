"""Reference checkpoint I/O: the `.tar` files written by run_train.py:369-380 and read by
tools/create_model_condition.py:62-89, and the fitting state of run_fit.py:320-331, map one-to-one onto this package's
modules (identical state_dict keys), so loading is plain `load_state_dict`.

PackedWeightCache is the second half of row f4: an on-disk cache of the ENGINE's layout of a network (fp16 K-major
images per concat segment, the split-precision low images of the coarse net, transposed copies, fp32 biases / latent
columns / heads — mofa_b200_export_packed), keyed by a hash of the checkpoint's tensors.  A process that finds the blob
uploads 2 bytes per weight straight into the kernels' buffers and never touches the fp32 parameters."""
from __future__ import annotations

import hashlib
import os

import numpy as np
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


class PackedWeightCache:
    """Directory of packed-weight blobs keyed by checkpoint content.

        cache = PackedWeightCache("~/.cache/mofanerf_b200")
        cache.load(engine, 0, coarse); cache.load(engine, 1, fine)      # first run: pack + write; later runs: read + upload

    The key is sha256 over every parameter's bytes in canonical order plus the library's layout version, so an edited or
    re-trained checkpoint can never pick up a stale blob; a blob written by another layout version is refused by the
    library itself (magic) and rebuilt."""

    LAYOUT = b"mofa_b200 packed layout 02"

    def __init__(self, directory: str):
        self.dir = os.path.expanduser(directory)
        os.makedirs(self.dir, exist_ok=True)
        self.hits = 0
        self.misses = 0

    @staticmethod
    def content_key(net) -> str:
        from .nets import canonical_tensors
        tensors, W, D = canonical_tensors(net)
        h = hashlib.sha256(PackedWeightCache.LAYOUT + f"|W={W}|D={D}|".encode())
        for t in tensors:
            h.update(np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy()).tobytes())
        return h.hexdigest()

    def path(self, key: str) -> str:
        return os.path.join(self.dir, key[:40] + ".mofapk")

    def load(self, engine, which: int, net) -> bool:
        """Make `net` network `which` of `engine`; returns True when the packed blob came from disk."""
        key = self.content_key(net)
        f = self.path(key)
        if os.path.exists(f):
            try:
                engine.import_packed(which, np.fromfile(f, dtype=np.uint8), key=engine._key(net))
                self.hits += 1
                return True
            except RuntimeError:
                os.remove(f)          # other layout version / truncated file: rebuild below
        engine.load_network(which, net, force=True)
        blob = engine.export_packed(which)
        tmp = f + f".tmp{os.getpid()}"
        blob.tofile(tmp)
        os.replace(tmp, f)
        self.misses += 1
        return False
