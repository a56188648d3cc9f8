"""Ray-range sharding across the GPUs of one box.

Rays are independent given (weights, latents): rank r of R renders the contiguous row-major range
[r*ceil(N/R), (r+1)*ceil(N/R)) and one all-gather of the per-rank RGB tile rebuilds the image —
the only collective on the path (NCCL over NVLink on the GPU box; gloo in the CPU tests).
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
                   keys: Iterable[str] = ("rgb_map",), group=None) -> Dict[str, torch.Tensor]:
    """Run `render_fn` on this rank's ray range and all-gather `keys` (default: the RGB tile only, as
    BASELINE.json's north_star specifies).  Keys not gathered are returned for the local range only,
    under '<key>_local'."""
    if not (dist.is_available() and dist.is_initialized()):
        return render_fn(rays)
    n = rays.shape[0]
    lo, hi = shard_range(n, dist.get_rank(group), dist.get_world_size(group))
    local = render_fn(rays[lo:hi])
    out = {}
    keys = tuple(keys)
    for k, v in local.items():
        if k in keys:
            out[k] = all_gather_rows(v, n, group)
        else:
            out[k + "_local"] = v
    return out
