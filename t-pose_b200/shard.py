"""Multi-GPU plumbing: independent units (images / image pairs) are sharded across ranks with no
data-path collective; the only exchange is the final gather of the per-unit results to rank 0
(SURVEY.md section 8e).  Backend-agnostic torch.distributed (nccl on the GPUs, gloo in the CPU
tests); one process per GPU.
"""
import numpy as np


def shard_range(n_units, world, rank):
    """Contiguous block of units owned by `rank`: 64 pairs over 8 ranks -> 8 each; remainders go to
    the first ranks."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(unit, n_units, world):
    for r in range(world):
        lo, hi = shard_range(n_units, world, r)
        if lo <= unit < hi:
            return r
    raise IndexError(unit)


def gather_units(local, n_units, dist=None, device="cpu", dst=0, timing=None):
    """local: {unit_index: float32 array (any shape, flattened on the wire)} for the units this rank
    owns.  Returns, on rank `dst`, {unit_index: flat float32 array} for ALL units; None elsewhere.
    Ragged results are padded to the global maximum length (an all-reduce of one int) and gathered
    with one collective for the payload and one for the int64 lengths, then trimmed again.
    timing: optional dict; receives 'ms' (CUDA-event time of the two gathers when `device` is a GPU)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        if timing is not None:
            timing["ms"] = 0.0
        return {u: np.asarray(a, np.float32).ravel() for u, a in local.items()}
    import torch          # (only with a process group: a single rank never pays the import)
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(n_units, world, rank)
    assert sorted(local) == list(range(lo, hi)), "rank does not hold exactly its shard"
    per_rank = max(shard_range(n_units, world, r)[1] - shard_range(n_units, world, r)[0] for r in range(world))
    longest = torch.tensor([max([np.asarray(a).size for a in local.values()] + [0])], dtype=torch.int64, device=device)
    dist.all_reduce(longest, op=dist.ReduceOp.MAX)
    L = int(longest.item())
    host = np.zeros((per_rank, max(L, 1)), np.float32)
    lens = np.zeros(per_rank, np.int64)
    for k, u in enumerate(range(lo, hi)):
        a = np.asarray(local[u], np.float32).ravel()
        host[k, :a.size] = a
        lens[k] = a.size
    buf = torch.from_numpy(host).to(device)
    nbuf = torch.from_numpy(lens).to(device)
    bucket = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    nbucket = [torch.empty_like(nbuf) for _ in range(world)] if rank == dst else None
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.gather(buf, bucket, dst=dst)
    dist.gather(nbuf, nbucket, dst=dst)
    if on_gpu:
        e1.record()
        torch.cuda.synchronize()
        if timing is not None:
            timing["ms"] = float(e0.elapsed_time(e1))
    elif timing is not None:
        timing["ms"] = 0.0
    if rank != dst:
        return None
    out = {}
    for r in range(world):
        rlo, rhi = shard_range(n_units, world, r)
        rows = bucket[r].cpu().numpy()
        ns = nbucket[r].cpu().numpy()
        for k, u in enumerate(range(rlo, rhi)):
            out[u] = rows[k, :int(ns[k])].copy()
    return out


def max_over_ranks(x, dist=None, device="cpu"):
    """Timing rule: every multi-GPU number is the max over ranks."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, dist=None, device="cpu"):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
