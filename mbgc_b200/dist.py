"""Multi-GPU plumbing of the match stage: one process per GPU, torch.distributed for the control plane only.

The path shards by TARGET: within a batch every rank matches its own targets against the same
frozen matcher state (index + reference replicated per GPU), so the data path needs no collective.  What the
ranks do exchange is tiny and host-side: per-target match counts (to lay the lists out in target order, the
order MBGC_Encoder concatenates per-target streams in, mbgccoder/MBGC_Encoder.cpp:542-564) and timings
(max over ranks).  Works with the "nccl" backend on GPUs and with "gloo" on CPU (tests)."""
import numpy as np


def shard_range(n_items, rank, world):
    """contiguous, balanced split of range(n_items): the first n % world ranks get one item more"""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_in_target_order(local_lists, n_targets, device="cpu"):
    """local_lists: this rank's per-target (n_i, 3) uint64 match arrays for shard_range(n_targets, rank, world).
    Returns, on every rank, the list of all n_targets arrays in target order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(local_lists)
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(n_targets, dtype=torch.int64, device=device)
    b, e = shard_range(n_targets, rank, world)
    assert e - b == len(local_lists)
    for k, a in enumerate(local_lists):
        counts[b + k] = len(a)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    counts_h = counts.cpu().numpy()
    per_rank = [int(counts_h[slice(*shard_range(n_targets, r, world))].sum()) for r in range(world)]
    width = max(per_rank) if per_rank else 0
    mine = np.zeros((width, 3), dtype=np.int64)
    if local_lists and per_rank[rank]:
        mine[:per_rank[rank]] = np.concatenate(local_lists).astype(np.int64, copy=False)
    bufs = [torch.zeros((width, 3), dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(mine).to(device))
    out = []
    for r in range(world):
        rb, re_ = shard_range(n_targets, r, world)
        flat = bufs[r].cpu().numpy().astype(np.uint64)
        o = 0
        for t in range(rb, re_):
            out.append(flat[o:o + int(counts_h[t])])
            o += int(counts_h[t])
    return out
