"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the box, gloo in
the CPU tests).  Three ways the path shards:

  runs        independent ``sampling`` runs round-robin over ranks; no data-path
              collective, one gather of the results at the end (src/mmsbm.py:182-185
              fans the same runs out to a process pool);
  folds       cv_fit folds x runs are the same thing one level up;
  ratings     ONE large run split by contiguous user range balanced by rating count:
              each rank owns its users' theta rows, eta and pr are replicated, and the
              unnormalised n_eta / n_pr are all-reduced every iteration before the
              normalisation epilogue (mmsbm_em_finalize).
"""
import numpy as np
import torch
import torch.distributed as dist


def dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_runs(n_runs, rank, world):
    """Run indices of ``rank``: round-robin, so every rank gets ceil or floor(S/world)."""
    return list(range(rank, n_runs, world))


def gather_runs(local, n_runs):
    """``local``: {run index: result dict} of this rank -> list of all S results, in run
    order, on every rank."""
    rank, world = dist_info()
    if world == 1:
        return [local[s] for s in range(n_runs)]
    boxes = [None] * world
    dist.all_gather_object(boxes, local)
    merged = {}
    for b in boxes:
        merged.update(b)
    return [merged[s] for s in range(n_runs)]


def user_partition(user_degree, world):
    """Contiguous user ranges [lo, hi) per rank with rating counts as equal as a prefix
    split allows.  Returns an int64 array of world+1 boundaries."""
    deg = np.asarray(user_degree, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(deg)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        cut = int(np.searchsorted(csum, target, side="left"))
        cut = max(cut, bounds[-1])
        bounds.append(min(cut, len(deg)))
    bounds.append(len(deg))
    return np.asarray(bounds, dtype=np.int64)


def allreduce_sum_(tensors):
    """In-place sum over ranks of a list of tensors (one flattened collective)."""
    rank, world = dist_info()
    if world == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
