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


def gather_units(local, n_units, dist=None, device="cpu", dst=0):
    """local: {unit_index: float32 array (any shape, flattened on the wire)} for the units this rank
    owns.  Returns, on rank `dst`, {unit_index: flat float32 array} for ALL units; None elsewhere.
    Ragged results are padded to the global maximum length (an all-reduce of one int), gathered
    with ONE collective, and trimmed again."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return {u: np.asarray(a, np.float32).ravel() for u, a in local.items()}
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(n_units, world, rank)
    assert sorted(local) == list(range(lo, hi)), "rank does not hold exactly its shard"
    per_rank = max(shard_range(n_units, world, r)[1] - shard_range(n_units, world, r)[0] for r in range(world))
    longest = torch.tensor([max([np.asarray(a).size for a in local.values()] + [0])], dtype=torch.int64, device=device)
    dist.all_reduce(longest, op=dist.ReduceOp.MAX)
    L = int(longest.item())
    # row layout: [length, payload..., padding]
    buf = torch.zeros((per_rank, L + 1), dtype=torch.float32, device=device)
    for k, u in enumerate(range(lo, hi)):
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(local[u], np.float32).ravel()))
        buf[k, 0] = float(a.numel())
        buf[k, 1:1 + a.numel()] = a.to(device)
    bucket = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bucket, dst=dst)
    if rank != dst:
        return None
    out = {}
    for r in range(world):
        rlo, rhi = shard_range(n_units, world, r)
        rows = bucket[r].cpu().numpy()
        for k, u in enumerate(range(rlo, rhi)):
            n = int(rows[k, 0])
            out[u] = rows[k, 1:1 + n].copy()
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
