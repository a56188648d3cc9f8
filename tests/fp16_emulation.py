"""Differentiable CPU emulation of the engine's reduced-precision dense chain (test infrastructure).

Same network, same algebra as oracle.NeRF.forward, but every dense layer sees fp16-rounded activations and
weights (fp32 accumulation), latent columns are applied in fp32 (the engine folds them into the bias), heads read
the un-rounded fp32 activations (the engine fuses them into the epilogue).  Rounding uses a straight-through
estimator so autograd yields the gradient the engine's backward pass computes analytically; ReLU masks then agree
with the engine's except at exact ties, which is what makes a tight gradient comparison possible.
"""
import torch


def q(x):
    return x + (x.half().float() - x).detach()


def _lin(x_q, W, b, lat=None, W_lat=None):
    y = x_q @ q(W).t() + b
    if lat is not None:
        y = y + lat @ W_lat.t()
    return y


def _skip_mlp(mod, lat, x_q, n_lat):
    """skipMLP(D, skip=4) on cat[lat, x]: returns (fp32 output of the last layer, its fp16-rounded copy)."""
    l1 = [m for m in mod.linears1 if isinstance(m, torch.nn.Linear)]
    l2 = [m for m in mod.linears2 if isinstance(m, torch.nn.Linear)]
    W0 = l1[0].weight
    h = torch.relu(_lin(x_q, W0[:, n_lat:], l1[0].bias, lat, W0[:, :n_lat]))
    for m in l1[1:]:
        h = torch.relu(_lin(q(h), m.weight, m.bias))
    W = l2[0].weight
    Wd = x_q.shape[1]
    y = q(x_q) @ q(W[:, n_lat:n_lat + Wd]).t() + q(h) @ q(W[:, n_lat + Wd:]).t() + l2[0].bias + lat @ W[:, :n_lat].t()
    h = torch.relu(y)
    for m in l2[1:]:
        h = torch.relu(_lin(q(h), m.weight, m.bias))
    return h


def nerf_forward_fp16(net, emb, shp, emb_dirs, tex):
    n_pe = emb.shape[1] - 30
    xl = [m for m in net.xyzEncode.linears1 if isinstance(m, torch.nn.Linear)]
    W0 = xl[0].weight
    h = torch.relu(_lin(q(emb[:, :n_pe]), W0[:, :n_pe], xl[0].bias, emb[:, n_pe:], W0[:, n_pe:]))
    for m in xl[1:]:
        h = torch.relu(_lin(q(h), m.weight, m.bias))
    sigma = _skip_mlp(net.linear_BiM_xyz, shp, q(h), shp.shape[1])
    alpha = sigma @ net.alpha_linear[0].weight.t() + net.alpha_linear[0].bias
    rgbc = _skip_mlp(net.linear_uv_xyzBiM, tex, q(sigma), tex.shape[1])
    Wv = net.linear_view_xyBMuv[0].weight
    nv = emb_dirs.shape[1]
    hv = torch.relu(q(emb_dirs) @ q(Wv[:, :nv]).t() + q(rgbc) @ q(Wv[:, nv:]).t() + net.linear_view_xyBMuv[0].bias)
    rgb = hv @ net.rgb_linear.weight.t() + net.rgb_linear.bias
    return torch.cat([rgb, alpha], -1)
