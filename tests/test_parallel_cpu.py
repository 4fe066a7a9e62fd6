"""The N > 1 host logic on CPU: world_size-2 gloo process groups (no GPU, no kernels).
Run sharding + result gather, the cv_fit folds x runs jobs, the seed broadcast, the partitions,
and the algebra of the sharded iteration (own users' ratings -> theta rows, own items' ratings
-> eta rows, n_pr all-reduced) with the oracle standing in for the device compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import random_params, random_triples


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def _gather_case(rank, world):
    from mmsbm_b200.parallel import dist_info, gather_runs, shard_runs
    assert dist_info() == (rank, world)
    mine = shard_runs(5, rank, world)
    local = {s: {"likelihood": float(s), "theta": np.full((2, 2), s)} for s in mine}
    res = gather_runs(local, 5)
    return [float(r["likelihood"]) for r in res], [int(r["theta"][0, 0]) for r in res]


def test_runs_shard_and_gather_in_order():
    for liks, tags in _run(_gather_case):
        assert liks == [0.0, 1.0, 2.0, 3.0, 4.0] and tags == [0, 1, 2, 3, 4]


def _sharded_iteration_case(rank, world):
    """What one iteration of mmsbm_em_run_sharded computes on a rank, with the oracle as the
    compute: n_theta of the own users from THEIR ratings, n_eta of the own items from THEIRS
    (complete local sums, nothing to reduce), n_pr partial over the own items summed over the
    ranks, then the exchange of the new rows (here: an all-gather)."""
    from mmsbm_b200.parallel import shard_rows
    from oracle import mmsbm_oracle as orc
    N, U, I, K, L, R = 6000, 70, 40, 4, 3, 5
    data = random_triples(3, N, U, I, R, heavy_tail=True)
    theta, eta, pr = random_params(4, U, I, K, L, R)
    rows_u, rows_i, ub, ib = shard_rows(data, U, I, rank, world)
    assert ub[0] == 0 and ub[-1] == U and ib[0] == 0 and ib[-1] == I
    ulo, uhi, ilo, ihi = ub[rank], ub[rank + 1], ib[rank], ib[rank + 1]
    # by-user side: own users' ratings, item ids global -> full eta, own theta rows
    nt, _, _ = orc.em_sums(rows_u, theta[ulo:uhi], eta, pr)
    theta_new = nt / np.maximum(np.bincount(rows_u[:, 0], minlength=uhi - ulo), 1)[:, None]
    # by-item side: own items' ratings, user ids global -> full theta, own eta rows
    _, ne, npr = orc.em_sums(rows_i, theta, eta[ilo:ihi], pr)
    eta_new = ne / np.maximum(np.bincount(rows_i[:, 1], minlength=ihi - ilo), 1)[:, None]
    tp = torch.from_numpy(npr.copy())
    dist.all_reduce(tp)                                      # the one collective of the iteration
    pr_new = orc.normalize_pr(tp.numpy())
    th_parts, et_parts = [None] * world, [None] * world
    dist.all_gather_object(th_parts, theta_new)
    dist.all_gather_object(et_parts, eta_new)
    fu, fi = orc.degree_factors(data, K, L)
    want = orc.em_iteration(data, theta, eta, pr, fu, fi)
    return (float(np.abs(np.concatenate(th_parts) - want[0]).max()),
            float(np.abs(np.concatenate(et_parts) - want[1]).max()), float(np.abs(pr_new - want[2]).max()),
            int(len(rows_u)), int(len(rows_i)))


def test_sharded_iteration_equals_the_unsharded_one():
    res = _run(_sharded_iteration_case)
    assert sum(r[3] for r in res) == 6000 and sum(r[4] for r in res) == 6000
    for sizes in ([r[3] for r in res], [r[4] for r in res]):
        assert max(sizes) - min(sizes) < 0.2 * 6000          # balanced by rating count
    for dth, det, dpr, _, _ in res:
        assert dth < 1e-13 and det < 1e-13 and dpr < 1e-13


def _seed_case(rank, world):
    from mmsbm_b200.parallel import broadcast_seed
    assert broadcast_seed(7) == 7                            # an explicit seed is left alone
    s = broadcast_seed(None)
    kids = np.random.default_rng(s).bit_generator._seed_seq.spawn(3)
    return int(s), [int(np.random.default_rng(k).integers(1 << 62)) for k in kids]


def test_unseeded_model_uses_one_entropy_on_all_ranks():
    """seed=None under torch.distributed: every rank must build the same child seeds (else the
    'same' run starts from different parameters on different ranks and cv folds differ)."""
    a, b = _run(_seed_case)
    assert a == b


def _cv_jobs_case(rank, world):
    """cv_fit's folds x runs jobs sharded over the ranks: every job
    exactly once, in fold-major order, balanced to within one job."""
    from mmsbm_b200.parallel import shard_jobs
    mine = shard_jobs(5, 4, rank, world)
    boxes = [None] * world
    dist.all_gather_object(boxes, mine)
    return mine, boxes


def test_cv_jobs_shard_over_ranks():
    for world in (2, 3):
        res = _run(_cv_jobs_case, world=world)
        every = sorted(j for j in res[0][1] for j in j)
        assert every == [(f, s) for f in range(5) for s in range(4)]
        sizes = [len(r[0]) for r in res]
        assert max(sizes) - min(sizes) <= 1
    from mmsbm_b200.parallel import shard_jobs
    per_gpu = [len(shard_jobs(5, 4, r, 8)) for r in range(8)]
    assert per_gpu == [3, 3, 3, 3, 2, 2, 2, 2]               # the BASELINE cv config on 8 GPUs


def test_partition_never_leaves_a_rank_empty():
    """One id holding more than 1/world of the ratings used to produce empty ranges
    ([0,3,3,3,3]) -> a rank without ids -> a distributed hang."""
    from mmsbm_b200.parallel import balanced_partition
    b = balanced_partition([1, 1, 100, 1], 4)
    assert list(b) == [0, 1, 2, 3, 4]
    b = balanced_partition([100, 1, 1, 1, 1, 1], 3)
    assert b[0] == 0 and b[-1] == 6 and np.all(np.diff(b) >= 1)
    g = np.random.default_rng(0)
    for _ in range(50):
        n, w = int(g.integers(1, 40)), int(g.integers(1, 9))
        c = g.integers(0, 1000, n) * (g.random(n) < 0.5)
        if w > n:
            with pytest.raises(ValueError):
                balanced_partition(c, w)
            continue
        b = balanced_partition(c, w)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 1) and len(b) == w + 1
