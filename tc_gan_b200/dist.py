"""
Multi-GPU plumbing for the SSN path: one process per GPU, networks sharded by index,
no collective on the solve path, one all-reduce of the 12-element generator gradient
(dJ, dD, dS) per GAN step (the reference is single-process).

Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n, rank=None, world_size=None):
    """Global indices of the networks owned by `rank`: i = rank, rank + G, rank + 2G, ..."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, n, world_size))


def allreduce_generator_grads(dJ, dD, dS, group=None):
    """Sum (dJ, dD, dS) over ranks with ONE collective on a packed 12-vector."""
    packed = torch.cat([dJ.reshape(-1), dD.reshape(-1), dS.reshape(-1)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[0:4].reshape(2, 2), packed[4:8].reshape(2, 2), packed[8:12].reshape(2, 2)


def first_successes(ok_local, num, group=None):
    """
    Rejection sampling across ranks.  `ok_local[m]` says whether the m-th network of this
    rank (global index rank + m * G) converged for every stimulus.  Returns
    (keep_local, n_kept, rejections): a bool mask over the local networks selecting those
    among the first `num` successes IN GLOBAL GENERATOR ORDER (tc_gan/ssnode.py:489-495),
    how many were found in total (may be < num: draw more), and how many of the globally
    consumed networks were rejected.
    """
    rank, w = world()
    ok_local = ok_local.to(torch.int32)
    m = ok_local.numel()
    if w > 1:
        gathered = [torch.empty_like(ok_local) for _ in range(w)]
        dist.all_gather(gathered, ok_local, group=group)
        ok_all = torch.stack(gathered, dim=1).reshape(-1)        # global order: index = m * G + rank
    else:
        ok_all = ok_local
    csum = torch.cumsum(ok_all, 0)
    keep_all = (ok_all > 0) & (csum <= num)
    n_kept = int(keep_all.sum())
    if n_kept >= num:
        last = int(torch.nonzero(keep_all).max())
        consumed = last + 1
    else:
        consumed = ok_all.numel()
    rejections = int(consumed - int(ok_all[:consumed].sum()))
    keep_local = keep_all.reshape(m, w)[:, rank] if w > 1 else keep_all
    return keep_local.to(torch.bool), n_kept, rejections
