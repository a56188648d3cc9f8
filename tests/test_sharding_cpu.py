"""N>1 path on CPU: ray-range sharding + the single all-gather, world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mofanerf_b200.distributed import render_sharded, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 640000, 160001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= (n + world - 1) // world


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rays = torch.arange(n * 11, dtype=torch.float32).reshape(n, 11)

    calls = []

    def fake_render(r):   # stands in for the CUDA engine: rgb row i encodes ray i's identity
        calls.append(r.shape[0])
        return {"rgb_map": r[:, :3] * 2.0 + 1.0, "z_std": r[:, 0], "rgb0": r[:, 3:6] - 1.0, "acc_map": r[:, 7]}

    out = render_sharded(fake_render, rays)
    lo, hi = shard_range(n, rank, world)
    ok = calls == [hi - lo]                                     # this rank rendered only its own range
    ok = ok and torch.equal(out["rgb_map"], rays[:, :3] * 2.0 + 1.0)
    # the extras come back complete too (same single collective), in row-major ray order
    ok = ok and torch.equal(out["z_std"], rays[:, 0]) and torch.equal(out["rgb0"], rays[:, 3:6] - 1.0)
    ok = ok and out["acc_map"].shape == (n,) and out["rgb0"].shape == (n, 3)
    try:      # per-sample outputs are refused, loudly
        render_sharded(lambda r: {"rgb_map": r[:, :3], "raw": r[:, None, :4].expand(-1, 128, -1)}, rays)
        ok = False
    except RuntimeError:
        pass
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allgather_rebuilds_row_major_order():
    world, n = 2, 1001          # odd count: the last shard is shorter and padded for the gather
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mofanerf_b200.distributed import allreduce_gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.ReLU(), torch.nn.Linear(33, 5))   # same init on every rank
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
    x = torch.full((4, 7), float(rank + 1))
    net(x).sum().backward()
    net[2].bias.grad = None                       # a parameter one rank did not touch
    local = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    # tiny buckets: several collectives, bucket boundaries inside the parameter list
    calls = allreduce_gradients(list(net.parameters()) + [frozen], bucket_bytes=256)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    ok = calls >= 2 and frozen.grad is None
    for j, p in enumerate(net.parameters()):
        want = sum((g[j] if g[j] is not None else torch.zeros_like(p)) for g in gathered) / world
        ok = ok and p.grad is not None and torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_average_matches_manual_sum():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
