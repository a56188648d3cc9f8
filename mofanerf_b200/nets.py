"""Host-side PyTorch modules that stay PyTorch and parameter containers for the
two MLPs whose forward pass runs in the CUDA engine.

The containers reproduce the reference's module tree (models/model.py:80-137,202-230) so that
checkpoints saved by the reference (`network_fn_state_dict`, tools/create_model_condition.py:75-78)
load with identical state_dict keys; their dense layers are executed by libmofa_b200, not here.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _xavier_relu(mod: nn.Module) -> None:
    for m in mod.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight.data, gain=nn.init.calculate_gain("relu"))


class SkipMLPParams(nn.Module):
    """Parameter container for models/model.py:202-230 `skipMLP`."""

    def __init__(self, D=8, W=256, input_ch=256, skip=None):
        super().__init__()
        self.skips = skip
        self.linears1 = nn.Sequential()
        self.linears2 = nn.Sequential()
        n1 = (skip if skip is not None else D) + 1
        for i in range(n1):
            self.linears1.add_module(f"Linear{i}", nn.Linear(input_ch if i == 0 else W, W))
            self.linears1.add_module(f"relu{i}", nn.ReLU())
            if skip is not None and i == 0:
                pass
        if skip is not None:
            # NB: the reference registers linears2.Linear0 *after* all of linears1 (init RNG order)
            for i in range(D - skip - 1):
                self.linears2.add_module(f"Linear{i}", nn.Linear(W + input_ch if i == 0 else W, W))
                self.linears2.add_module(f"relu{i}", nn.ReLU())
        _xavier_relu(self)


class NeRFParams(nn.Module):
    """Parameter container for models/model.py:80-137 `NeRF` (use_viewdirs=True)."""

    def __init__(self, D=8, W=256, input_ch=93, input_ch_views=27, input_ch_textureCodes=256,
                 input_ch_shapeCodes=50):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.input_ch_shapeCodes, self.input_ch_textureCodes = input_ch_shapeCodes, input_ch_textureCodes
        self.xyzEncode = SkipMLPParams(D=3, W=W, input_ch=input_ch, skip=None)
        self.linear_BiM_xyz = SkipMLPParams(D=D, W=W, input_ch=input_ch_shapeCodes + W, skip=4)
        self.linear_uv_xyzBiM = SkipMLPParams(D=D, W=W, input_ch=input_ch_textureCodes + W, skip=4)
        self.linear_view_xyBMuv = nn.Sequential(nn.Linear(input_ch_views + W, W // 2), nn.ReLU())
        self.alpha_linear = nn.Sequential(nn.Linear(W, 1))
        self.rgb_linear = nn.Linear(W // 2, 3)
        _xavier_relu(self)

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "NeRFParams holds parameters only; the dense layers run in the sm_100a engine "
            "(mofanerf_b200.B200Renderer / Engine.run_network).  There is no PyTorch fallback.")


class StyleModule(nn.Module):
    """models/model.py:174-199: shape code -> (expression scale, bias).  One row per call."""

    def __init__(self, D=4, W=256, input_ch_bm=50, out_ch=30):
        super().__init__()
        self.linears1 = nn.Sequential()
        for i in range(D):
            self.linears1.add_module(f"Linear{i}", nn.Linear(input_ch_bm if i == 0 else W, W))
            self.linears1.add_module(f"relu{i}", nn.ReLU())
        self.linears_scale = nn.Linear(W, out_ch)
        self.linears_bias = nn.Linear(W, out_ch)
        _xavier_relu(self)

    def forward(self, bmcodes):
        feature = self.linears1(bmcodes)
        return self.linears_scale(feature), self.linears_bias(feature)


def _xavier_std(m, gain):
    if isinstance(m, nn.Conv2d):
        k = m.kernel_size[0] * m.kernel_size[1]
        return gain * math.sqrt(2.0 / ((m.in_channels + m.out_channels) * k))
    if isinstance(m, nn.Linear):
        return gain * math.sqrt(2.0 / (m.in_features + m.out_features))
    return None


def _init_mod(m, gain=1.0):
    std = _xavier_std(m, gain)
    if std is not None:
        m.weight.data.uniform_(-std * math.sqrt(3.0), std * math.sqrt(3.0))
        if m.bias is not None:
            m.bias.data.zero_()


def _init_seq(seq):
    mods = list(seq)
    for a, b in zip(mods[:-1], mods[1:]):
        if isinstance(b, nn.LeakyReLU):
            _init_mod(a, nn.init.calculate_gain("leaky_relu", b.negative_slope))
        elif isinstance(b, nn.ReLU):
            _init_mod(a, nn.init.calculate_gain("relu"))
        else:
            _init_mod(a)
    _init_mod(mods[-1])


class TexEncoderCore(nn.Module):
    """models/tex_encoder_mod.py:22-100 `Encoder(ninputs=1)`: 512x512x3 UV texture -> 256-d code."""

    def __init__(self, ninputs=1, uvCodesLen=256):
        super().__init__()
        self.ninputs = ninputs
        self.down1 = nn.ModuleList([nn.Sequential(
            nn.Conv2d(3, 32, 4, 2, 1), nn.LeakyReLU(0.2),
            nn.Conv2d(32, 32, 4, 2, 1), nn.LeakyReLU(0.2),
            nn.Conv2d(32, 32, 4, 2, 1), nn.LeakyReLU(0.2),
            nn.Conv2d(32, 32, 4, 2, 1), nn.LeakyReLU(0.2),
            nn.Conv2d(32, 64, 4, 2, 1), nn.LeakyReLU(0.2),
            nn.Conv2d(64, 128, 4, 2, 1), nn.LeakyReLU(0.2),
            nn.Conv2d(128, 256, 4, 2, 1), nn.LeakyReLU(0.2)) for _ in range(ninputs)])
        self.down2 = nn.Sequential(nn.Linear(256 * ninputs * 4 * 4, 512), nn.LeakyReLU(0.2))
        h = w = 512
        ypad = ((h + 127) // 128) * 128 - h
        xpad = ((w + 127) // 128) * 128 - w
        self.pad = nn.ZeroPad2d((xpad // 2, xpad - xpad // 2, ypad // 2, ypad - ypad // 2))
        self.mu = nn.Linear(512, uvCodesLen)
        self.logstd = nn.Linear(512, uvCodesLen)   # unused by forward (dead in the reference too); kept for keys
        for i in range(ninputs):
            _init_seq(self.down1[i])
        _init_seq(self.down2)
        _init_mod(self.mu)
        _init_mod(self.logstd)
        self.decoding = nn.Sequential(
            nn.Linear(uvCodesLen, uvCodesLen), nn.LeakyReLU(0.1),
            nn.Linear(uvCodesLen, uvCodesLen), nn.LeakyReLU(0.1),
            nn.Linear(uvCodesLen, uvCodesLen), nn.LeakyReLU(0.1))
        _xavier_relu(self.decoding)

    def forward(self, x, losslist=()):
        x = self.pad(x)
        x = torch.cat([self.down1[i](x).view(-1, 256 * 4 * 4) for i in range(self.ninputs)], dim=1)
        x = self.down2(x)
        return self.decoding(self.mu(x)), {}


class TexEncoder(nn.Module):
    """models/tex_encoder_mod.py:7-19 `EnDeUVmap`."""

    def __init__(self, uvCodesLen=256):
        super().__init__()
        self.encoder = TexEncoderCore(1, uvCodesLen=uvCodesLen)

    def forward(self, uvMap, lossList=()):
        return self.encoder(uvMap, lossList)


def get_embedder(multires: int, i: int = 0):
    """models/model.py:48-63.  Returned for API compatibility (create_nerf passes embed_fn to the renderer);
    the engine computes the encoding in-kernel, this torch version serves host-side callers only."""
    if i == -1:
        return nn.Identity(), 3
    freqs = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)

    def embed(x):
        out = [x]
        for f in freqs:
            out.append(torch.sin(x * f.to(x.device)))
            out.append(torch.cos(x * f.to(x.device)))
        return torch.cat(out, -1)

    return embed, 3 + 6 * multires


def build_nets(seed: int = 0, W_c=256, D_c=8, W_f=1024, D_f=10, device="cpu"):
    """Seeded random-init (coarse, fine) parameter containers — synthetic weights for bench/smoke."""
    torch.manual_seed(seed)
    coarse = NeRFParams(D_c, W_c)
    fine = NeRFParams(D_f, W_f) if W_f else None
    style = StyleModule()
    return coarse.to(device), (fine.to(device) if fine is not None else None), style.to(device)


def canonical_tensors(net: nn.Module):
    """(weight, bias) tensors in the order mofa_b200_load_weights expects, plus (W, D).

    Accepts the reference's NeRF, the oracle's, NeRFParams, or any of them wrapped in DataParallel
    (run_fit.py:166-167)."""
    net = getattr(net, "module", net)
    sd = dict(net.named_parameters())

    def pair(prefix):
        return [sd[prefix + ".weight"], sd[prefix + ".bias"]]

    n2 = sum(1 for k in sd if k.startswith("linear_BiM_xyz.linears2.") and k.endswith(".weight"))
    D = n2 + 5
    W = sd["xyzEncode.linears1.Linear0.weight"].shape[0]
    out = []
    for i in range(4):
        out += pair(f"xyzEncode.linears1.Linear{i}")
    for blk in ("linear_BiM_xyz", "linear_uv_xyzBiM"):
        for i in range(5):
            out += pair(f"{blk}.linears1.Linear{i}")
        for i in range(n2):
            out += pair(f"{blk}.linears2.Linear{i}")
    out += pair("linear_view_xyBMuv.0")
    out += pair("alpha_linear.0")
    out += pair("rgb_linear")
    return out, W, D
