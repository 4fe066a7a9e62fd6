"""Multi-GPU checks, launched by tests/test_multi_gpu.py under torchrun (one rank per GPU):
  1. MMSBM.fit with runs sharded over ranks == the same fit in one process (bit-identical:
     runs are independent and every sum has a fixed order);
  2. MMSBM.fit(shard="ratings") (user-range x item-range shards, copy-engine exchange of the
     parameter rows, one NCCL all-reduce of n_pr per iteration: mmsbm_em_run_sharded) == the
     unsharded fit within 1e-9 per element after 12 iterations, likelihood within 1e-8;
  3. cv_fit with the folds x runs jobs sharded == serial folds;
  4. ShardedEngine at the BASELINE kernel shapes vs the one-GPU loop."""
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.chdir(os.environ.get("MMSBM_TMP", "/tmp"))
    from mmsbm_b200 import MMSBM
    from mmsbm_b200.engine import Engine
    from tests.util import rel_err

    g = np.random.default_rng(5)
    n = 30000
    df = pd.DataFrame({"users": g.integers(0, 400, n), "items": g.integers(0, 250, n),
                       "ratings": g.integers(1, 6, n)})
    S, K, L, it = 4, 6, 5, 12

    # reference for both checks: all runs in this process only (no process group involved)
    solo = MMSBM(K, L, iterations=it, sampling=S, seed=3)
    solo.data_handler = __import__("mmsbm_b200").DataHandler()
    train = solo.data_handler.format_train_data(df)
    solo._prepare_objects(train)
    want = solo._run_batch(solo._engine, list(solo.child_states), list(range(S)))

    a = MMSBM(K, L, iterations=it, sampling=S, seed=3)
    a.fit(df, silent=True)
    for s in range(S):
        for key in ("theta", "eta", "pr"):
            assert np.array_equal(a.results[s][key], want[s][key]), (s, key)
        assert a.results[s]["likelihood"] == want[s]["likelihood"]

    b = MMSBM(K, L, iterations=it, sampling=S, seed=3, shard="ratings")
    b.fit(df, silent=True)
    worst = 0.0
    for s in range(S):
        for key in ("theta", "eta", "pr"):
            worst = max(worst, rel_err(b.results[s][key], want[s][key]))
        assert abs(b.results[s]["likelihood"] - want[s]["likelihood"]) <= 1e-8 * abs(want[s]["likelihood"])
    assert worst < 1e-9, worst
    pa, pb = a.predict(df.iloc[:500]), b.predict(df.iloc[:500])
    assert rel_err(pb, pa) < 1e-9
    # 3. cv_fit: folds x runs jobs sharded over the ranks == serial folds in one process
    from tests.util import mock_data
    solo_cv = MMSBM(2, 2, iterations=10, sampling=3, seed=1)
    folds_pairs = solo_cv._make_folds(mock_data(1), 2)
    want_acc = []
    for tr, te in folds_pairs:                    # no process group use: _run_batch directly
        solo_cv.data_handler = __import__("mmsbm_b200").DataHandler()
        solo_cv._prepare_objects(solo_cv.data_handler.format_train_data(tr))
        r = solo_cv._run_batch(solo_cv._engine, list(solo_cv.child_states), [0, 1, 2])
        solo_cv.results = [r[0], r[1], r[2]]
        solo_cv.predict(te)
        want_acc.append(float(solo_cv.score(silent=True)["stats"]["accuracy"]))
    c = MMSBM(2, 2, iterations=10, sampling=3, seed=1)
    got_acc = [float(a) for a in c.cv_fit(mock_data(1), folds=2)]
    assert got_acc == want_acc, (got_acc, want_acc)
    # 4. the sharded loop itself at the shapes its kernels specialise on (six runs + a pair per warp
    #    at K=L=20; 8-lane rows, one run at K=L=32), heavy-tailed ids, against the one-GPU loop
    from mmsbm_b200.parallel import ShardedEngine
    from tests.util import random_params, random_triples
    worst2 = 0.0
    for (K2, L2, S2) in ((20, 20, 8), (32, 32, 1), (10, 10, 3)):
        N2, U2, I2, R2 = 120000, 3000, 700, 5
        data = random_triples(7, N2, U2, I2, R2, heavy_tail=True)
        th, et, pr = random_params(9, U2, I2, K2, L2, R2, S=S2)
        one = Engine(data, U2, I2, R2, K2, L2)
        one.set_params(th, et, pr)
        one.run(7)
        w_th, w_et, w_pr = one.get_params()
        w_lik = one.likelihood()
        sh = ShardedEngine(data, U2, I2, R2, K2, L2)
        sh.set_params(th, et, pr)
        sh.run(4)
        sh.run(3)
        g_th, g_et, g_pr = sh.get_params()
        g_lik = sh.likelihood()
        sh.close()
        worst2 = max(worst2, rel_err(g_th, w_th), rel_err(g_et, w_et), rel_err(g_pr, w_pr))
        assert np.max(np.abs(g_lik - w_lik) / np.abs(w_lik)) < 1e-8, (K2, g_lik, w_lik)
    assert worst2 < 1e-9, worst2
    dist.barrier()
    if rank == 0:
        print(f"multi-gpu ok: world={dist.get_world_size()} rating-sharded worst rel err {worst:.2e} / {worst2:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
