"""The N > 1 host logic on CPU: world_size-2 gloo process groups (no GPU, no kernels).
Run sharding + result gather, the flattened all-reduce, and the algebra of the rating-sharded
iteration (shard by user range -> partial n_eta / n_pr -> all-reduce -> normalise) with the
oracle standing in for the device compute."""
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


def _allreduce_case(rank, world):
    from mmsbm_b200.parallel import allreduce_sum_
    a = torch.full((3, 4), float(rank + 1), dtype=torch.float64)
    b = torch.arange(5, dtype=torch.float64) * (rank + 1)
    allreduce_sum_([a, b])
    return a.numpy().copy(), b.numpy().copy()


def test_flattened_allreduce():
    for a, b in _run(_allreduce_case):
        np.testing.assert_array_equal(a, np.full((3, 4), 3.0))
        np.testing.assert_array_equal(b, np.arange(5) * 3.0)


def _sharded_iteration_case(rank, world):
    """What RatingShardedEngine.run does per iteration, with the oracle as the compute."""
    from mmsbm_b200.parallel import allreduce_sum_, shard_rows_by_user
    from oracle import mmsbm_oracle as orc
    N, U, I, K, L, R = 6000, 70, 40, 4, 3, 5
    data = random_triples(3, N, U, I, R, heavy_tail=True)
    theta, eta, pr = random_params(4, U, I, K, L, R)
    local, lo, hi, bounds = shard_rows_by_user(data, U, rank, world)
    assert bounds[0] == 0 and bounds[-1] == U
    # partial sums over the local ratings; items unseen locally contribute zero rows
    nt = np.zeros((hi - lo, K)); ne = np.zeros((I, L)); npr = np.zeros((K, L, R))
    if len(local):
        # oracle em_sums sizes its outputs from theta / eta, so hand it the owned theta rows
        nt, ne, npr = orc.em_sums(local, theta[lo:hi], eta, pr)
    deg_u = np.maximum(np.bincount(local[:, 0], minlength=hi - lo), 1)[:, None]
    theta_new = nt / deg_u                                   # owned users: final
    ideg = torch.from_numpy(np.bincount(local[:, 1], minlength=I).astype(np.int64))
    dist.all_reduce(ideg)                                    # global item degree
    te, tp = torch.from_numpy(ne.copy()), torch.from_numpy(npr.copy())
    allreduce_sum_([te, tp])
    eta_new = te.numpy() / np.maximum(ideg.numpy(), 1)[:, None]
    pr_new = orc.normalize_pr(tp.numpy())
    parts = [None] * world
    dist.all_gather_object(parts, theta_new)
    fu, fi = orc.degree_factors(data, K, L)
    want = orc.em_iteration(data, theta, eta, pr, fu, fi)
    return (float(np.abs(np.concatenate(parts) - want[0]).max()),
            float(np.abs(eta_new - want[1]).max()), float(np.abs(pr_new - want[2]).max()),
            int(len(local)))


def test_rating_sharded_iteration_equals_the_unsharded_one():
    res = _run(_sharded_iteration_case)
    assert sum(r[3] for r in res) == 6000
    sizes = [r[3] for r in res]
    assert max(sizes) - min(sizes) < 0.2 * 6000          # balanced by rating count
    for dth, det, dpr, _ in res:
        assert dth < 1e-13 and det < 1e-13 and dpr < 1e-13
