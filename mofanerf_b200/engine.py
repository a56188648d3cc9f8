"""Python handle on the sm_100a engine (ctypes over include/mofa_b200.h).  PyTorch is used only for
device memory, streams and tensors-as-buffers."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .nets import canonical_tensors


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """One context per CUDA device.  Not a fallback-capable object: construction fails without a B200."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("mofanerf_b200.Engine needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"mofanerf_b200.Engine: device must be CUDA, got {self.device}")
        self.lib = _lib.load()
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(self.lib.mofa_b200_create(C.byref(h), idx))
        self._h = h
        self._ws: Optional[torch.Tensor] = None
        self._net_keys = {0: None, 1: None}
        self._keep = {}   # fp32 staging tensors kept alive until the stream has consumed them
        self.chunk_rays = 0
        self.max_train_bytes = 64 << 30

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mofa_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    @property
    def launch_count(self) -> int:
        return int(self.lib.mofa_b200_launch_count(self._h))

    def profile_enable(self, on: bool = True) -> None:
        _lib.check(self.lib.mofa_b200_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        """{net: {ms, algo_flops, launches}} for the tensor-core dense launches since the last read."""
        buf = (C.c_double * 6)()
        _lib.check(self.lib.mofa_b200_profile_read(self._h, buf))
        return {n: dict(ms=buf[3 * n], algo_flops=buf[3 * n + 1], launches=int(buf[3 * n + 2])) for n in (0, 1)}

    # ------------------------------------------------------------------ weights / latents
    @staticmethod
    def _key(net) -> tuple:
        mod = getattr(net, "module", net)
        return (id(mod),) + tuple((p.data_ptr(), p._version) for p in mod.parameters())

    def load_network(self, which: int, net: torch.nn.Module, force: bool = False) -> None:
        """Upload + repack `net` (reference NeRF / NeRFParams, optionally DataParallel-wrapped).
        Cached on (module identity, parameter versions) so repeated render calls do not repack
        (SURVEY.md §7 'Weights behind DataParallel')."""
        key = self._key(net)
        if not force and self._net_keys[which] == key:
            return
        tensors, W, D = canonical_tensors(net)
        with torch.cuda.device(self.device):
            dev = [_f32c(t, self.device) for t in tensors]
            arr = (C.c_void_p * len(dev))(*[t.data_ptr() for t in dev])
            _lib.check(self.lib.mofa_b200_load_weights(self._h, which, W, D, arr, len(dev), self._stream()))
        self._keep[("net", which)] = dev
        self._net_keys[which] = key

    def export_packed(self, which: int):
        """The engine's own layout of network `which` as a host byte array (numpy uint8) — see mofa_b200_export_packed."""
        import numpy as np
        n = int(self.lib.mofa_b200_packed_bytes(self._h, which))
        if n == 0:
            raise RuntimeError(f"mofa_b200: network {which} is not loaded")
        buf = np.empty(n, dtype=np.uint8)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_export_packed(self._h, which, buf.ctypes.data, n, self._stream()))
        return buf

    def import_packed(self, which: int, blob, key=None) -> None:
        """Rebuild network `which` from an export_packed() blob: no fp32 tensors, no PyTorch modules involved.
        key: the cache key (Engine._key(net)) under which later load_network(net) calls should find it already loaded."""
        import numpy as np
        blob = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_import_packed(self._h, which, blob.ctypes.data, blob.size, self._stream()))
        self._keep.pop(("net", which), None)
        self._net_keys[which] = key

    def set_latents(self, shape: torch.Tensor, exp_mod: torch.Tensor, tex: torch.Tensor) -> None:
        s = _f32c(shape.reshape(-1)[:50], self.device)
        e = _f32c(exp_mod.reshape(-1), self.device)
        t = _f32c(tex.reshape(-1), self.device)
        if s.numel() != 50 or e.numel() != 30 or t.numel() != 256:
            raise ValueError(f"latent sizes must be 50/30/256, got {s.numel()}/{e.numel()}/{t.numel()}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_set_latents(self._h, s.data_ptr(), e.data_ptr(), t.data_ptr(), self._stream()))
        self._keep["lat"] = (s, e, t)

    # ------------------------------------------------------------------ hot path
    def render_rays(self, rays: torch.Tensor, N_samples: int, N_importance: int = 0, *, run_fine: bool = True,
                    fine_net: int = 1, perturb: float = 0.0, raw_noise_std: float = 0.0, lindisp: bool = False,
                    white_bkgd: bool = False, retraw: bool = False, seed: int = 0,
                    t_rand: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
                    noise_c: Optional[torch.Tensor] = None, noise_f: Optional[torch.Tensor] = None,
                    want_aux: bool = False, gemm_simt: bool = False,
                    chunk_rays: Optional[int] = None, train: bool = False) -> Dict[str, torch.Tensor]:
        """rays [N, >=11] fp32 on this device -> dict with the reference's render_rays keys
        (models/render_class.py:338-345)."""
        if rays.device != self.device:
            raise ValueError(f"rays on {rays.device}, engine on {self.device}")
        rays = rays.detach().to(torch.float32)
        if rays.stride(-1) != 1 or rays.dim() != 2 or rays.shape[1] < 11:
            rays = rays.reshape(-1, rays.shape[-1]).contiguous()
        n = rays.shape[0]
        fine = N_importance > 0 and run_fine
        S_last = N_samples + N_importance if fine else N_samples
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        out = {"rgb_map": torch.empty(n, 3, **f32), "disp_map": torch.empty(n, **f32),
               "acc_map": torch.empty(n, **f32)}
        if fine:
            out.update(rgb0=torch.empty(n, 3, **f32), disp0=torch.empty(n, **f32), acc0=torch.empty(n, **f32),
                       z_std=torch.empty(n, **f32))
        if retraw:
            out["raw"] = torch.empty(n, S_last, 4, **f32)
        if want_aux:
            out["weights"] = torch.empty(n, S_last, **f32)
            out["z_vals"] = torch.empty(n, S_last, **f32)
        if n == 0:
            return out
        chunk = self.chunk_rays if chunk_rays is None else int(chunk_rays)
        extra = [None if x is None else _f32c(x, dev) for x in (t_rand, u, noise_c, noise_f)]
        with torch.cuda.device(dev):
            if train:   # activations of every layer are kept: the workspace belongs to this call (returned)
                nbytes = self.lib.mofa_b200_train_workspace_bytes(self._h, n, N_samples, N_importance if fine else 0,
                                                                  int(fine_net))
                if nbytes == 0:
                    raise RuntimeError("mofa_b200: networks must be loaded before a training-mode render")
                if nbytes > self.max_train_bytes:
                    raise RuntimeError(f"training-mode render of {n} rays needs {nbytes / 2**30:.1f} GiB of activations "
                                       f"(limit {self.max_train_bytes / 2**30:.0f} GiB): use a smaller ray batch")
                ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            else:
                nbytes = self.lib.mofa_b200_workspace_bytes(self._h, n, N_samples, N_importance if fine else 0, chunk)
                ws = self._workspace(nbytes)
            a = _lib.RenderArgs()
            a.struct_size = C.sizeof(_lib.RenderArgs)
            a.flags = ((_lib.FLAG_LINDISP if lindisp else 0) | (_lib.FLAG_WHITE_BKGD if white_bkgd else 0) |
                       (_lib.FLAG_GEMM_SIMT if gemm_simt else 0))
            a.rays, a.n_rays, a.ray_stride = rays.data_ptr(), n, rays.stride(0)
            a.n_samples, a.n_importance = int(N_samples), int(N_importance)
            a.run_fine, a.fine_net, a.chunk_rays = int(bool(run_fine)), int(fine_net), chunk
            a.perturb, a.raw_noise_std, a.seed = float(perturb), float(raw_noise_std), int(seed) & (2 ** 64 - 1)
            a.t_rand, a.u, a.noise_c, a.noise_f = [_ptr(x) for x in extra]
            a.rgb, a.disp, a.acc = out["rgb_map"].data_ptr(), out["disp_map"].data_ptr(), out["acc_map"].data_ptr()
            if fine:
                a.rgb0, a.disp0, a.acc0 = out["rgb0"].data_ptr(), out["disp0"].data_ptr(), out["acc0"].data_ptr()
                a.z_std = out["z_std"].data_ptr()
            a.raw = _ptr(out.get("raw"))
            a.weights = _ptr(out.get("weights"))
            a.z_vals = _ptr(out.get("z_vals"))
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
            fn = self.lib.mofa_b200_render_rays_train_fwd if train else self.lib.mofa_b200_render_rays_fwd
            _lib.check(fn(self._h, C.byref(a), self._stream()))
        self._keep["extra"] = (rays, extra)
        if train:
            out["_train_ws"] = ws
            out["_rays"] = rays
            out["_noise"] = (extra[2], extra[3])
        return out

    def render_rays_bwd(self, saved: Dict[str, torch.Tensor], N_samples: int, N_importance: int, *, run_fine: bool,
                        fine_net: int, white_bkgd: bool, lindisp: bool, d_rgb=None, d_acc=None, d_rgb0=None,
                        d_acc0=None, loss_scale: float = 1.0, param_grads=None):
        """Backward of a training-mode render_rays: returns (d_rays [n,11], d_shape [50], d_expmod [30], d_tex [256]).
        loss_scale: a Python float, or a device tensor {scale, 1/scale} (no host synchronisation).
        param_grads: optional (coarse_list, fine_list) of zero-initialised fp32 tensors shaped like the canonical
        (weight, bias) parameter lists; weight gradients are accumulated into them (training)."""
        rays, ws = saved["_rays"], saved["_train_ws"]
        noise_c, noise_f = saved["_noise"]
        dev, n = self.device, rays.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        g = [None if x is None else _f32c(x, dev) for x in (d_rgb, d_acc, d_rgb0, d_acc0)]
        d_rays = torch.empty(n, 11, **f32)
        d_shape, d_exp, d_tex = torch.empty(50, **f32), torch.empty(30, **f32), torch.empty(256, **f32)
        with torch.cuda.device(dev):
            a = _lib.BwdArgs()
            a.struct_size = C.sizeof(_lib.BwdArgs)
            a.flags = (_lib.FLAG_LINDISP if lindisp else 0) | (_lib.FLAG_WHITE_BKGD if white_bkgd else 0)
            a.rays, a.n_rays, a.ray_stride = rays.data_ptr(), n, rays.stride(0)
            a.n_samples, a.n_importance = int(N_samples), int(N_importance)
            a.run_fine, a.fine_net = int(bool(run_fine)), int(fine_net)
            a.noise_c, a.noise_f = _ptr(noise_c), _ptr(noise_f)
            a.d_rgb, a.d_acc, a.d_rgb0, a.d_acc0 = [_ptr(x) for x in g]
            if torch.is_tensor(loss_scale):     # device-resident {scale, 1/scale}: no host round trip
                loss_scale = loss_scale.to(device=dev, dtype=torch.float32).contiguous()
                a.loss_scale, a.loss_scale_dev = 1.0, loss_scale.data_ptr()
            else:
                a.loss_scale = float(loss_scale)
            a.d_rays, a.d_shape, a.d_expmod, a.d_tex = d_rays.data_ptr(), d_shape.data_ptr(), d_exp.data_ptr(), d_tex.data_ptr()
            keep = []
            if param_grads is not None:
                for name, lst in (("coarse", param_grads[0]), ("fine", param_grads[1])):
                    if lst is None:
                        continue
                    arr = (C.c_void_p * len(lst))(*[t.data_ptr() for t in lst])
                    keep.append(arr)
                    setattr(a, f"d_params_{name}", C.cast(arr, C.c_void_p))
                    setattr(a, f"n_params_{name}", len(lst))
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
            _lib.check(self.lib.mofa_b200_render_rays_bwd(self._h, C.byref(a), self._stream()))
        self._keep["bwd"] = (g, loss_scale)
        return d_rays, d_shape, d_exp, d_tex

    def run_network(self, which: int, pts: torch.Tensor, viewdirs: torch.Tensor, gemm_simt: bool = False) -> torch.Tensor:
        """pts [..., 3], viewdirs broadcastable to pts -> raw [..., 4]  (models/render_class.py:69-94)."""
        shp = pts.shape[:-1]
        p = _f32c(pts, self.device).reshape(-1, 3)
        v = _f32c(viewdirs.expand(pts.shape), self.device).reshape(-1, 3)
        n = p.shape[0]
        out = torch.empty(n, 4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self._workspace(self.lib.mofa_b200_query_workspace_bytes(self._h, n))
            _lib.check(self.lib.mofa_b200_run_network(self._h, which, p.data_ptr(), v.data_ptr(), n, out.data_ptr(),
                                                      _lib.FLAG_GEMM_SIMT if gemm_simt else 0, ws.data_ptr(),
                                                      ws.numel(), self._stream()))
        self._keep["q"] = (p, v)
        return out.reshape(*shp, 4)

    def generate_rays(self, H: int, W: int, K, c2w, near: float, far: float, first: int = 0,
                      n: Optional[int] = None) -> torch.Tensor:
        """get_rays + ray packing on the device (tools/run_nerf_helpers.py:153-168, models/render_class.py:158-179):
        rows [first, first + n) of the row-major H x W ray batch as a [n, 12] tensor (o d near far viewdir pad).
        K (3x3) and c2w ([3,4] or [4,4]) are read on the host: 21 scalars go to the kernel by value."""
        total = int(H) * int(W)
        n = total - first if n is None else int(n)
        Kh = [float(K[a][b]) for a in range(3) for b in range(3)]
        c = c2w.detach().cpu() if torch.is_tensor(c2w) else c2w
        ch = [float(c[a][b]) for a in range(3) for b in range(4)]
        out = torch.empty(n, 12, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_generate_rays(self._h, int(H), int(W), (C.c_float * 9)(*Kh), (C.c_float * 12)(*ch),
                                                        float(near), float(far), int(first), n, out.data_ptr(), 12,
                                                        self._stream()))
        return out

    # ------------------------------------------------------------------ op-level entry points
    def embed(self, x: torch.Tensor, multires: int) -> torch.Tensor:
        x = _f32c(x, self.device).reshape(-1, 3)
        out = torch.empty(x.shape[0], 3 + 6 * multires, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_embed(self._h, x.data_ptr(), x.shape[0], multires, out.data_ptr(), self._stream()))
        return out

    def raw2outputs(self, raw, z_vals, rays_d, noise=None, white_bkgd=False):
        raw, z, d = _f32c(raw, self.device), _f32c(z_vals, self.device), _f32c(rays_d, self.device)
        nz = None if noise is None else _f32c(noise, self.device)
        n, S = z.shape
        f32 = dict(dtype=torch.float32, device=self.device)
        rgb, disp, acc = torch.empty(n, 3, **f32), torch.empty(n, **f32), torch.empty(n, **f32)
        w, depth = torch.empty(n, S, **f32), torch.empty(n, **f32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_raw2outputs(self._h, raw.data_ptr(), z.data_ptr(), d.data_ptr(), 3, _ptr(nz),
                                                      n, S, int(white_bkgd), rgb.data_ptr(), disp.data_ptr(),
                                                      acc.data_ptr(), w.data_ptr(), depth.data_ptr(), self._stream()))
        return rgb, disp, acc, w, depth

    def composite_bwd(self, raw, z_vals, rays, noise, d_rgb, d_acc, white_bkgd=False):
        """Adjoint of raw2outputs: returns (d_raw [n,S,4], d_rays [n,11] with the |rays_d| term in columns 3..5)."""
        raw, z, ry = _f32c(raw, self.device), _f32c(z_vals, self.device), _f32c(rays, self.device)
        nz = None if noise is None else _f32c(noise, self.device)
        gr = None if d_rgb is None else _f32c(d_rgb, self.device)
        ga = None if d_acc is None else _f32c(d_acc, self.device)
        n, S = z.shape
        d_raw = torch.empty(n, S, 4, dtype=torch.float32, device=self.device)
        d_rays = torch.empty(n, 11, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_raw2outputs_bwd(self._h, raw.data_ptr(), z.data_ptr(), ry.data_ptr(), ry.stride(0),
                                                          _ptr(nz), _ptr(gr), _ptr(ga), n, S, int(white_bkgd),
                                                          d_raw.data_ptr(), d_rays.data_ptr(), self._stream()))
        return d_raw, d_rays

    def sample_pdf_merge(self, z_vals, weights, N_importance, u=None):
        z, w = _f32c(z_vals, self.device), _f32c(weights, self.device)
        uu = None if u is None else _f32c(u, self.device)
        n, S = z.shape
        f32 = dict(dtype=torch.float32, device=self.device)
        zs, zm, sd = torch.empty(n, N_importance, **f32), torch.empty(n, S + N_importance, **f32), torch.empty(n, **f32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_sample_pdf_merge(self._h, z.data_ptr(), w.data_ptr(), _ptr(uu), n, S,
                                                           N_importance, zs.data_ptr(), zm.data_ptr(), sd.data_ptr(),
                                                           self._stream()))
        return zs, zm, sd

    def wgrad(self, dZ, X, n_valid=None, scale=1.0, simt=False):
        """C = scale * dZ^T · X  (fp16 operands [P, M'] and [P, K], fp32 result [M', n_valid]) — the weight-gradient GEMM."""
        dZ = dZ.to(self.device, torch.float16).contiguous()
        X = X.to(self.device, torch.float16).contiguous()
        P, Mp = dZ.shape
        Kb = X.shape[1]
        n_valid = Kb if n_valid is None else n_valid
        out = torch.zeros(Mp, n_valid, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_wgrad(self._h, dZ.data_ptr(), Mp, X.data_ptr(), Kb, n_valid, P, float(scale),
                                                out.data_ptr(), n_valid, int(simt), self._stream()))
        return out

    def dense(self, A0, B0, bias=None, A1=None, B1=None, relu=True, simt=False, mode=None):
        """C = act(A0·B0^T (+ A1·B1^T) + bias) with fp16 operands.
        mode: None/'default' (CTA-pair tcgen05 when N % 256 == 0), 'simt' (verification), '1cta' (single-CTA tcgen05)."""
        sel = {None: 1 if simt else 0, 'default': 0, 'simt': 1, '1cta': 2}[mode]
        A0 = A0.to(self.device, torch.float16).contiguous()
        B0 = B0.to(self.device, torch.float16).contiguous()
        M, K0 = A0.shape
        N = B0.shape[0]
        K1 = 0
        if A1 is not None:
            A1 = A1.to(self.device, torch.float16).contiguous()
            B1 = B1.to(self.device, torch.float16).contiguous()
            K1 = A1.shape[1]
        b = None if bias is None else _f32c(bias, self.device)
        out = torch.empty(M, N, dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mofa_b200_dense(self._h, A0.data_ptr(), B0.data_ptr(), K0, _ptr(A1), _ptr(B1), K1,
                                                _ptr(b), out.data_ptr(), M, N, int(relu), sel, self._stream()))
        return out


_engines: Dict[int, Engine] = {}


def get_engine(device=None) -> Engine:
    """Process-wide engine per CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError("mofanerf_b200 needs a CUDA device (sm_100a); there is no CPU path")
    idx = None if device is None else torch.device(device).index
    if idx is None:        # a bare "cuda" means the current device, not GPU 0
        idx = torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(torch.device("cuda", idx))
    return _engines[idx]
