"""The N>1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the sharding and
the final gather that bench.py / config 5 use over NCCL."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_units, q):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as entry
    entry.load_package()
    shard = importlib.import_module("tpose_b200.shard")
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_units, world, rank)
    # unit u produces a ragged "vertex buffer" that only depends on u (sharding must not change results)
    local = {u: np.arange(10 + 3 * u, dtype=np.float32) * (u + 1) for u in range(lo, hi)}
    got = shard.gather_units(local, n_units, dist)
    mx = shard.max_over_ranks(float(rank + 1), dist)
    sm = shard.sum_over_ranks(float(hi - lo), dist)
    if rank == 0:
        ok = sorted(got) == list(range(n_units)) and all(
            np.array_equal(got[u], np.arange(10 + 3 * u, dtype=np.float32) * (u + 1)) for u in range(n_units))
        q.put((ok, mx, sm))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_units", [(2, 64), (2, 5), (3, 7)])
def test_gather_over_gloo(world, n_units):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, mx, sm = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and mx == float(world) and sm == float(n_units)


def test_shard_ranges_partition_the_units(pkg):
    import importlib
    shard = importlib.import_module("tpose_b200.shard")
    for n in (1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            assert all(shard.owner_of(u, n, world) == r for r, (l, h) in enumerate(spans) for u in range(l, h))
    assert [shard.shard_range(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
