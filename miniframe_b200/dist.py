"""Sharding of the hyper-parameter batch across the GPUs of one box.

The likelihood of one parameter set never talks to another, so the batch is cut
into contiguous shards, one per rank (one process per GPU), and the only
collective is an all-gather of the per-sample log-likelihoods (NCCL over NVLink
on GPUs; gloo in the CPU tests).  The reference has no counterpart: its only
concurrency is emcee's ``threads=4`` pool (examples/BIGgp_mcmc.py:88).
"""
import numpy as np


def shard_bounds(total, rank, world):
    """[lo, hi) of rank's contiguous shard; the first total % world ranks hold one more."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total, world):
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def all_gather_loglike(local, total, group=None):
    """local: this rank's (hi - lo,) float64 tensor.  Returns the (total,) tensor,
    identical on every rank.  Shards are padded to a common length so one
    all_gather_into_tensor call suffices."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    width = max(sizes)
    send = torch.full((width,), float("nan"), dtype=torch.float64, device=local.device)
    send[:local.numel()] = local
    recv = torch.empty((world * width,), dtype=torch.float64, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if all(s == width for s in sizes):
        return recv
    return torch.cat([recv[r * width:r * width + sizes[r]] for r in range(world)])


def log_likelihood_batch_sharded(gp, A, Bm=None, C=None, group=None, compute=None):
    """Every rank passes the same full batch; each evaluates its shard on its own
    GPU and all ranks return the full (B,) vector of log-likelihoods.

    ``compute(gp, A_shard, Bm_shard, C_shard) -> (b,) tensor`` defaults to the
    class's own ``log_likelihood_batch`` (the CUDA path)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    total = A.shape[0]
    lo, hi = shard_bounds(total, rank, world)
    cut = lambda x: None if x is None else x[lo:hi]
    if compute is None:
        def compute(gp_, a, bm, c):
            a = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
            args = [a.to("cuda")] + ([bm] if c is None else [bm, c])
            return gp_.log_likelihood_batch(*args)
    if hi > lo:
        local = compute(gp, cut(A), cut(Bm), cut(C))
    else:
        local = torch.empty((0,), dtype=torch.float64)
    if not torch.is_tensor(local):
        local = torch.as_tensor(np.asarray(local, dtype=np.float64))
    if dist.get_backend(group) == "nccl" and not local.is_cuda:
        local = local.cuda()
    return all_gather_loglike(local, total, group)
