"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference package (eudald-seeslab/mmsbm v1.0.7) is imported read-only from
/root/reference/src; nothing of it is copied.  Outputs are small ``.npz`` /
``.json`` files committed next to this script; tests read only those.

What is pinned (file:line = the reference code that produced the numbers):
  toy_backends.npz   compute_omegas / prod_dist on the 3-row toy of
                     tests/test_backends.py:21-30,48-56 (numpy AND numba kernels)
  fixture.npz/.json  the 100-row fixture of tests/test_mmsbm.py:12-50
                     (mock_data(1) train, mock_data(2) test, MMSBM(2,2,
                     iterations=10, seed=1)): encoding, theta0/eta0/pr0,
                     state after 1 and 10 iterations, likelihood, prediction
                     matrix, stats, index lists of src/mmsbm.py:114-122
  sampling3.npz      same data with sampling=3: per-run results, accuracies,
                     chosen best run (src/mmsbm.py:474-478)
  medium.npz         N=4000, U=150, I=90, K=7, L=5, R=5 random problem:
                     update_coefficients, normalisations, likelihood, prod_dist
  wide.npz           N=1500, K=20, L=20 and K=3, L=32 one-iteration cases
  encoding.json      DataHandler on awkward ids ([1,10,2,100,11], float ratings,
                     unseen test ids)
  cvfit.json         cv_fit(mock_data(1), folds=2) accuracies and fold indices
"""
import io
import json
import logging
import os
import sys

import numpy as np
import pandas as pd

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def mock_data(seed, n=100):
    # the generator of the reference's test fixture (tests/test_mmsbm.py:12-22);
    # restated here because the GPU box has no /root/reference
    rng = np.random.default_rng(seed)
    return pd.DataFrame({
        "users": [f"user{rng.choice(list(range(5)))}" for _ in range(n)],
        "items": [f"item{rng.choice(list(range(10)))}" for _ in range(n)],
        "ratings": [rng.choice(list(range(1, 6))) for _ in range(n)],
    })


def main():
    assert os.path.isdir(REF), "reference not mounted; run this in the build container"
    sys.path.insert(0, REF)
    os.chdir("/tmp")  # the reference logger writes mmsbm.log into the cwd
    import kernels_numpy as kn
    import kernels_numba as kb
    from data_handler import DataHandler
    from expectation_maximization import ExpectationMaximization
    from mmsbm import MMSBM

    # ---------------------------------------------------------------- toy
    out = {}
    for tag, seed in (("omega", 0), ("prod", 1)):
        rng = np.random.default_rng(seed)
        data = np.array([[0, 0, 0], [1, 1, 1], [0, 1, 2]], dtype=np.int64)
        theta = rng.random((2, 2)); eta = rng.random((2, 2)); pr = rng.random((2, 2, 3))
        theta /= theta.sum(axis=1, keepdims=True)
        eta /= eta.sum(axis=1, keepdims=True)
        pr /= pr.sum(axis=2, keepdims=True)
        out[f"{tag}_data"] = data
        out[f"{tag}_theta"], out[f"{tag}_eta"], out[f"{tag}_pr"] = theta, eta, pr
        out[f"{tag}_omegas_numpy"] = kn.compute_omegas(data, theta, eta, pr)
        out[f"{tag}_omegas_numba"] = kb.compute_omegas(data, theta, eta, pr)
        out[f"{tag}_prod_numpy"] = kn.prod_dist(data, theta, eta, pr)
        nt, ne, npr = kn.update_coefficients(data, theta, eta, pr)
        out[f"{tag}_ntheta"], out[f"{tag}_neta"], out[f"{tag}_npr"] = nt, ne, npr
    np.savez_compressed(os.path.join(HERE, "toy_backends.npz"), **out)

    # ------------------------------------------------------------ fixture
    def prepared(sampling, iterations):
        mm = MMSBM(2, 2, iterations=iterations, sampling=sampling, seed=1, backend="numpy")
        mm.data_handler = DataHandler()
        train = mm.data_handler.format_train_data(mock_data(1))
        mm._prepare_objects(train)
        return mm, train

    mm, train = prepared(1, 10)
    seed0 = mm.child_states[0]
    rng = np.random.default_rng(seed0)
    theta0 = mm.em.normalize_with_d(rng.random((mm.p + 1, 2)), 'user')
    eta0 = mm.em.normalize_with_d(rng.random((mm.m + 1, 2)), 'item')
    pr0 = mm.em.normalize_with_self(rng.random((2, 2, mm._dims['n_ratings'])))
    nt, ne, npr = mm.em.update_coefficients(train, theta0, eta0, pr0)
    theta1 = mm.em.normalize_with_d(nt, 'user')
    eta1 = mm.em.normalize_with_d(ne, 'item')
    pr1 = mm.em.normalize_with_self(npr)
    mm.iterations = 10
    import tqdm.auto  # silence the progress bar
    res = mm.run_one_sampling(train, seed0, 0)
    mm.results = [res]
    pred = mm.predict(mock_data(2))
    score = mm.score(silent=True)
    fx = dict(
        train=train, test=mm.test,
        theta0=theta0, eta0=eta0, pr0=pr0,
        ntheta1=nt, neta1=ne, npr1=npr, theta1=theta1, eta1=eta1, pr1=pr1,
        theta10=res["theta"], eta10=res["eta"], pr10=res["pr"],
        likelihood10=np.float64(res["likelihood"]),
        likelihood1=np.float64(mm.em.compute_likelihood(train, theta1, eta1, pr1)),
        prediction=pred,
        norm_user=mm._normalization_factors['user'],
        norm_item=mm._normalization_factors['item'],
        user_index_concat=np.concatenate(mm._user_indices),
        user_index_len=np.array([len(a) for a in mm._user_indices]),
        item_index_concat=np.concatenate(mm._item_indices),
        item_index_len=np.array([len(a) for a in mm._item_indices]),
        rating_index_concat=np.concatenate(mm._rating_indices),
        rating_index_len=np.array([len(a) for a in mm._rating_indices]),
    )
    np.savez_compressed(os.path.join(HERE, "fixture.npz"), **fx)
    stats = {k: float(v) for k, v in score["stats"].items()}
    objs = score["objects"]
    meta = {
        "stats": stats,
        "obs_dict": mm.data_handler.obs_dict,
        "items_dict": mm.data_handler.items_dict,
        "ratings_dict": mm.data_handler.ratings_dict,
        "ratings": [int(a) for a in mm.ratings],
        "theta_index": [str(a) for a in objs["theta"].index],
        "eta_index": [str(a) for a in objs["eta"].index],
        "pr_keys": [str(a) for a in objs["pr"].keys()],
        "theta_col0_sum": float(objs["theta"].sum(axis=0)[0]),
        "eta_col0_sum": float(objs["eta"].sum(axis=0)[0]),
        "pr_sums": [float(a.sum().sum()) for a in objs["pr"].values()],
    }
    with open(os.path.join(HERE, "fixture.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)

    # ---------------------------------------------------------- sampling=3
    mm3, train3 = prepared(3, 10)
    runs = [mm3.run_one_sampling(train3, s, i) for i, s in enumerate(mm3.child_states)]
    mm3.results = runs
    pred3 = mm3.predict(mock_data(2))
    rats = [mm3.em.compute_prod_dist(mm3.test, a["theta"], a["eta"], a["pr"]) for a in runs]
    accs = [mm3._compute_stats(a)["accuracy"] for a in rats]
    sc3 = mm3.score(silent=True)
    np.savez_compressed(
        os.path.join(HERE, "sampling3.npz"),
        thetas=np.array([a["theta"] for a in runs]),
        etas=np.array([a["eta"] for a in runs]),
        prs=np.array([a["pr"] for a in runs]),
        likelihoods=np.array([a["likelihood"] for a in runs]),
        rats=np.array(rats), accuracies=np.array(accs),
        best=np.int64(mm3.choose_best_run(rats)), prediction=pred3,
        stats_keys=np.array(sorted(sc3["stats"].keys())),
        stats_vals=np.array([float(sc3["stats"][k]) for k in sorted(sc3["stats"].keys())]),
    )

    # -------------------------------------------------------------- medium
    def random_problem(seed, N, U, I, K, L, R):
        g = np.random.default_rng(seed)
        data = np.stack([g.integers(0, U, N), g.integers(0, I, N), g.integers(0, R, N)], axis=1)
        # every id present at least once so that max id + 1 == U / I / R
        data[:U, 0] = np.arange(U); data[:I, 1] = np.arange(I); data[:R, 2] = np.arange(R)
        theta = g.random((U, K)); eta = g.random((I, L)); pr = g.random((K, L, R))
        theta /= theta.sum(axis=1, keepdims=True)
        eta /= eta.sum(axis=1, keepdims=True)
        pr /= pr.sum(axis=2, keepdims=True)
        return data.astype(np.int64), theta, eta, pr

    def one_step(data, theta, eta, pr):
        K, L, R = pr.shape
        du = np.maximum(np.bincount(data[:, 0]), 1)
        di = np.maximum(np.bincount(data[:, 1]), 1)
        dims = {"n_samples": len(data), "n_user_groups": K, "n_item_groups": L, "n_ratings": R}
        em = ExpectationMaximization(
            dims, None, None, None,
            {"user": np.repeat(du[:, None], K, 1), "item": np.repeat(di[:, None], L, 1)},
            backend="numpy")
        nt, ne, npr = em.update_coefficients(data, theta, eta, pr)
        return dict(
            data=data, theta=theta, eta=eta, pr=pr, ntheta=nt, neta=ne, npr=npr,
            theta1=em.normalize_with_d(nt, "user"), eta1=em.normalize_with_d(ne, "item"),
            pr1=em.normalize_with_self(npr),
            likelihood=np.float64(em.compute_likelihood(data, theta, eta, pr)),
            prod=em.compute_prod_dist(data[:257], theta, eta, pr),
        )

    np.savez_compressed(os.path.join(HERE, "medium.npz"),
                        **one_step(*random_problem(7, 4000, 150, 90, 7, 5, 5)))
    wide = {}
    for tag, args in (("k20", (11, 1500, 60, 40, 20, 20, 5)), ("l32", (12, 1200, 50, 30, 3, 32, 4)),
                      ("odd", (13, 900, 40, 35, 9, 11, 7))):
        for k, v in one_step(*random_problem(*args)).items():
            wide[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "wide.npz"), **wide)

    # ------------------------------------------------------------ encoding
    enc = {}
    dh = DataHandler()
    df = pd.DataFrame({"users": [1, 10, 2, 100, 11, 10], "items": ["b", "a", "B", "10", "9", "a"],
                       "ratings": [5.0, 1.0, 3.5, 10.0, 2.0, 1.0]})
    enc["train_in"] = {c: [str(x) for x in df[c]] for c in df.columns}
    enc["train_out"] = dh.format_train_data(df).tolist()
    enc["dicts"] = [dh.obs_dict, dh.items_dict, dh.ratings_dict]
    tdf = pd.DataFrame({"users": [10, 7, 2, 100, 1], "items": ["a", "a", "zz", "9", "B"],
                        "ratings": [1.0, 5.0, 3.5, 4.0, 10.0]})
    enc["test_in"] = {c: [str(x) for x in tdf[c]] for c in tdf.columns}
    buf = io.StringIO()
    h = logging.StreamHandler(buf)
    logging.getLogger("MMSBM").addHandler(h)
    enc["test_out"] = dh.format_test_data(tdf).tolist()
    logging.getLogger("MMSBM").removeHandler(h)
    enc["test_warnings"] = buf.getvalue().strip().split("\n")
    with open(os.path.join(HERE, "encoding.json"), "w") as fh:
        json.dump(enc, fh, indent=1, sort_keys=True)

    # --------------------------------------------------------------- cvfit
    mmcv = MMSBM(2, 2, iterations=10, seed=1, backend="numpy")
    accs = mmcv.cv_fit(mock_data(1), folds=2)
    with open(os.path.join(HERE, "cvfit.json"), "w") as fh:
        json.dump({"accuracies": [float(a) for a in accs],
                   "final_test": np.asarray(mmcv.test).tolist()}, fh, indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
