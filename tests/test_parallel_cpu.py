"""CPU, world_size 2, gloo: the image-sharding path (glare_b200/parallel.py) -- split, per-rank work, all_gather,
original order restored, uneven batches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glare_b200.parallel import shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 15, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glare_b200.parallel import enhance_sharded
    x = torch.arange(n * 6, dtype=torch.float32).reshape(n, 2, 3)
    calls = []

    def fn(t):
        calls.append(t.shape[0])
        return t * 2 + 1

    y = enhance_sharded(fn, x)
    q.put((rank, calls[0], bool(torch.equal(y, x * 2 + 1))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [15, 4, 1])
def test_enhance_sharded_gloo_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(ok for _, _, ok in res)
    assert res[0][1] + res[1][1] == n and res[0][1] - res[1][1] in (0, 1)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glare_b200.parallel import allreduce_gradients
    g = {"b.weight": torch.full((3, 2), float(rank + 1)), "a.bias": torch.arange(4, dtype=torch.float32) * (rank + 1)}
    out = allreduce_gradients(g)
    ok = torch.allclose(out["b.weight"], torch.full((3, 2), 1.5)) and torch.allclose(out["a.bias"], torch.arange(4, dtype=torch.float32) * 1.5)
    q.put((rank, bool(ok), tuple(out["b.weight"].shape)))
    dist.destroy_process_group()


def test_allreduce_gradients_gloo_world2():
    """stage-2 data parallelism: one flat all-reduce averages the per-rank gradient dictionaries"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(ok and shape == (3, 2) for _, ok, shape in res)


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glare_b200.parallel import AsyncGather
    ga = AsyncGather()
    ok = True
    for step in range(3):                                   # the two slots are reused from the third submit on
        local = torch.full((2, 4, 5, 3), rank * 10 + step, dtype=torch.uint8)
        out = ga.submit(local)
        ga.wait()
        want = torch.cat([torch.full((2, 4, 5, 3), r * 10 + step, dtype=torch.uint8) for r in range(world)])
        ok = ok and bool(torch.equal(out, want))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_async_gather_gloo_world2():
    """bench.py's result gather (uint8 batches, one all_gather per step); on CPU tensors it runs synchronously"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(ok for _, ok in res)
