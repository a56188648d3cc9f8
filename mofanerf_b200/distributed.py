"""Ray-range sharding across the GPUs of one box.

Rays are independent given (weights, latents): rank r of R renders the contiguous row-major range
[r*ceil(N/R), (r+1)*ceil(N/R)) and one all-gather of the per-rank RGB tile rebuilds the image —
the only collective on the rendering path (NCCL over NVLink on the GPU box; gloo in the CPU tests).

Training is the one place the path has a real exchange step: each rank draws its own N_rand rays
(run_train.py:318-331), and the weight gradients the backward kernels leave in `param.grad` are averaged across ranks
before the optimiser step — `allreduce_gradients` below, bucketed so that the ~116 MB of fp32 gradients (29 M NeRF
parameters) go out as a few large collectives instead of one per tensor.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate equally-partitioned row blocks (shard_range) from every rank, in rank order."""
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    pad = per - local.shape[0]
    if pad > 0:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], 0)
    local = local.contiguous()
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local, group=group)
    return out[:n_total]


def render_sharded(render_fn: Callable[[torch.Tensor], Dict[str, torch.Tensor]], rays: torch.Tensor,
                   group=None, max_floats_per_ray: int = 32) -> Dict[str, torch.Tensor]:
    """Run `render_fn` on this rank's ray range and rebuild every per-ray output on every rank with ONE all-gather:
    the keys render_rays returns (rgb_map, disp_map, acc_map, rgb0, disp0, acc0, z_std: 11 floats per ray) are packed
    side by side into a single [N/R, C] tile, gathered, and unpacked — the RGB tile BASELINE.json's north_star names plus
    the reference's extras in the same collective (28 MB per 800x800 frame).  Per-sample outputs (`raw`, weights) are
    refused here: gathering [N, S, 4] is not what the sharded mode is for."""
    if not (dist.is_available() and dist.is_initialized()):
        return render_fn(rays)
    n = rays.shape[0]
    lo, hi = shard_range(n, dist.get_rank(group), dist.get_world_size(group))
    return gather_ray_outputs(render_fn(rays[lo:hi]), n, lo, hi, group, max_floats_per_ray)


def gather_ray_outputs(local: Dict[str, torch.Tensor], n: int, lo: int, hi: int, group=None,
                       max_floats_per_ray: int = 32) -> Dict[str, torch.Tensor]:
    """The collective half of render_sharded: `local` holds this rank's outputs for rays [lo, hi) of n."""
    keys = sorted(local.keys())
    widths = []
    for k in keys:
        v = local[k]
        w = int(v[0].numel()) if v.shape[0] > 0 else int(torch.tensor(v.shape[1:]).prod().item()) if v.dim() > 1 else 1
        if v.requires_grad:
            raise RuntimeError("ray-sharded rendering is inference only (the all-gather is not differentiable)")
        widths.append(w)
    if sum(widths) > max_floats_per_ray:
        big = [k for k, w in zip(keys, widths) if w > 16]
        raise RuntimeError(f"ray-sharded rendering cannot return per-sample outputs {big}: render those without "
                           "MOFA_B200_SHARD or on one rank")
    packed = torch.cat([local[k].reshape(hi - lo, -1).float() for k in keys], 1)
    full = all_gather_rows(packed, n, group)
    out, c0 = {}, 0
    for k, w in zip(keys, widths):
        out[k] = full[:, c0:c0 + w].reshape((n,) + tuple(local[k].shape[1:]))
        c0 += w
    return out


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, bucket_bytes: int = 32 << 20,
                        average: bool = True) -> int:
    """Sum (or average) `p.grad` over the ranks of `group`, in place.  Gradients are packed into flat buckets of about
    `bucket_bytes` per (device, dtype) — NVSwitch makes the cost per collective launch-bound, not link-bound, so few
    large buckets beat one all-reduce per tensor.  Parameters whose grad is None on this rank contribute zeros (every
    rank must walk the same parameter list: the reference's optimiser groups, create_model_condition.py:55-60).
    Returns the number of collectives issued."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    params = [p for p in params if p.requires_grad]
    calls = 0
    i = 0
    while i < len(params):
        key = (params[i].device, params[i].dtype)
        bucket, nbytes = [], 0
        while i < len(params) and (params[i].device, params[i].dtype) == key and (not bucket or nbytes < bucket_bytes):
            bucket.append(params[i])
            nbytes += params[i].numel() * params[i].element_size()
            i += 1
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        calls += 1
    return calls
