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


def shard_rows_by_user(data, n_users, rank, world):
    """Rows of this rank when ratings are sharded by contiguous user range balanced by rating
    count.  Returns (local rows with user ids shifted to start at 0, lo, hi, bounds)."""
    data = np.asarray(data)
    deg = np.bincount(data[:, 0], minlength=n_users)
    bounds = user_partition(deg, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    keep = (data[:, 0] >= lo) & (data[:, 0] < hi)
    local = data[keep].copy()
    local[:, 0] -= lo
    return local, lo, hi, bounds


class RatingShardedEngine:
    """S runs with the ratings sharded by user range over the ranks of the default process
    group (the Netflix-shaped config).  Rank g owns theta[lo_g:hi_g] and
    all ratings of those users; eta and pr are replicated.  One exchange per iteration: the
    unnormalised n_eta and n_pr are summed over ranks (NCCL all-reduce over NVLink), then every
    rank applies the same normalisation epilogue (mmsbm_em_finalize)."""

    def __init__(self, data, n_users, n_items, n_levels, K, L, device=None):
        from . import _lib
        from .engine import Engine
        self._lib = _lib
        self.rank, self.world = dist_info()
        if self.world > n_users:
            raise ValueError("more ranks than users")
        self.U = int(n_users)
        local, self.lo, self.hi, self.bounds = shard_rows_by_user(data, n_users, self.rank, self.world)
        self.engine = Engine(local, self.hi - self.lo, n_items, n_levels, K, L, device=device)
        self.N = int(np.asarray(data).shape[0])
        # the degree that normalises eta is the GLOBAL item degree
        self.ideg = self.engine.ideg.clone()
        if self.world > 1:
            dist.all_reduce(self.ideg, op=dist.ReduceOp.SUM)

    def set_params(self, theta, eta, pr):
        theta = np.asarray(theta)
        if theta.ndim == 2:
            theta, eta, pr = theta[None], np.asarray(eta)[None], np.asarray(pr)[None]
        self.engine.set_params(theta[:, self.lo:self.hi], eta, pr)

    def run(self, iterations):
        e = self.engine
        for _ in range(int(iterations)):
            _, eta_raw, pr_raw = e.step_raw(self._lib.RAW_ETA_PR)   # theta' is final: users are owned
            if self.world > 1:      # n_eta and n_pr share one buffer: a single collective, in place
                dist.all_reduce(e._alt_flat, op=dist.ReduceOp.SUM)
            e.finalize(eta_raw, pr_raw, ideg=self.ideg)
            e.swap()

    def likelihood(self):
        lik = self.engine.likelihood_device().clone()
        if self.world > 1:
            dist.all_reduce(lik, op=dist.ReduceOp.SUM)
        return lik.cpu().numpy()

    def get_params(self):
        """Full (theta [S,U,K], eta, pr) on every rank."""
        th, et, pr = self.engine.get_params()
        if self.world == 1:
            return th, et, pr
        parts = [None] * self.world
        dist.all_gather_object(parts, th)
        return np.concatenate(parts, axis=1), et, pr
