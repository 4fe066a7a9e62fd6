"""The oracle (oracle/mmsbm_oracle.py) against the golden vectors produced by the
real reference (tests/golden/make_golden.py) and against the known answers of the
reference's own tests.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import mmsbm_oracle as orc

TIGHT = dict(rtol=1e-13, atol=1e-300)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_toy_backends(golden_dir):
    g = _load(golden_dir, "toy_backends.npz")
    for tag in ("omega", "prod"):
        args = [g[f"{tag}_{k}"] for k in ("data", "theta", "eta", "pr")]
        w = orc.omegas(*args)
        np.testing.assert_array_equal(w, g[f"{tag}_omegas_numpy"])
        # reference tests/test_backends.py:35,60 use atol=1e-8 across backends
        assert np.allclose(w, g[f"{tag}_omegas_numba"], atol=1e-8)
        np.testing.assert_allclose(orc.rating_distribution(*args), g[f"{tag}_prod_numpy"], **TIGHT)
        nt, ne, npr = orc.em_sums(*args)
        np.testing.assert_array_equal(nt, g[f"{tag}_ntheta"])
        np.testing.assert_array_equal(ne, g[f"{tag}_neta"])
        np.testing.assert_array_equal(npr, g[f"{tag}_npr"])


def test_fixture_one_and_ten_iterations(golden_dir):
    g = _load(golden_dir, "fixture.npz")
    train = g["train"]
    fu, fi = orc.degree_factors(train, 2, 2)
    np.testing.assert_array_equal(fu, g["norm_user"])
    np.testing.assert_array_equal(fi, g["norm_item"])
    _, kids = orc.child_seeds(1, 1)
    th0, et0, pr0 = orc.seeded_init(kids[0], fu.shape[0], fi.shape[0], 2, 2, 5, fu, fi)
    np.testing.assert_array_equal(th0, g["theta0"])
    np.testing.assert_array_equal(et0, g["eta0"])
    np.testing.assert_array_equal(pr0, g["pr0"])
    nt, ne, npr = orc.em_sums(train, th0, et0, pr0)
    np.testing.assert_array_equal(nt, g["ntheta1"])
    np.testing.assert_array_equal(ne, g["neta1"])
    np.testing.assert_array_equal(npr, g["npr1"])
    th1, et1, pr1 = orc.em_iteration(train, th0, et0, pr0, fu, fi)
    np.testing.assert_array_equal(th1, g["theta1"])
    np.testing.assert_array_equal(pr1, g["pr1"])
    assert orc.likelihood(train, th1, et1, pr1) == g["likelihood1"]
    res = orc.run_em(train, kids[0], 2, 2, 10)
    np.testing.assert_array_equal(res["theta"], g["theta10"])
    np.testing.assert_array_equal(res["eta"], g["eta10"])
    np.testing.assert_array_equal(res["pr"], g["pr10"])
    assert res["likelihood"] == g["likelihood10"]
    # chunked variant (used for large shapes) only moves the last bits
    resc = orc.run_em(train, kids[0], 2, 2, 10, chunk=17)
    np.testing.assert_allclose(resc["theta"], g["theta10"], rtol=1e-12)
    np.testing.assert_allclose(resc["likelihood"], g["likelihood10"], rtol=1e-12)


def test_fixture_known_answers(golden_dir):
    """The numbers asserted by the reference's tests/test_mmsbm.py:53-102."""
    g = _load(golden_dir, "fixture.npz")
    meta = json.load(open(os.path.join(golden_dir, "fixture.json")))
    test = g["test"]
    rat = orc.rating_distribution(test, g["theta10"], g["eta10"], g["pr10"])
    np.testing.assert_allclose(rat, g["prediction"], **TIGHT)
    assert rat.sum() == pytest.approx(100, 0.01)
    st = orc.prediction_stats(rat, test[:, 2], meta["ratings"])
    assert st["accuracy"] == pytest.approx(0.13, 0.01)
    assert st["one_off_accuracy"] == pytest.approx(0.55, 0.01)
    assert st["mae"] == pytest.approx(0.78, 0.01)
    assert st["s2"] == 153
    assert st["s2pond"] == pytest.approx(129.4766730930339, rel=1e-12)
    assert float(g["likelihood10"]) == pytest.approx(-13.773187406968459, rel=1e-12)
    for k in ("accuracy", "one_off_accuracy", "mae", "s2", "s2pond"):
        assert float(st[k]) == pytest.approx(meta["stats"][k], rel=1e-13)
    assert g["theta10"].sum(axis=0)[0] == pytest.approx(2.11, 0.1)
    assert g["eta10"].sum(axis=0)[0] == pytest.approx(5.93, 0.1)


def test_best_run_choice(golden_dir):
    g = _load(golden_dir, "sampling3.npz")
    fx = _load(golden_dir, "fixture.npz")
    train, test = fx["train"], fx["test"]
    _, kids = orc.child_seeds(1, 3)
    rats = []
    for s, kid in enumerate(kids):
        res = orc.run_em(train, kid, 2, 2, 10)
        np.testing.assert_array_equal(res["theta"], g["thetas"][s])
        np.testing.assert_array_equal(res["pr"], g["prs"][s])
        assert res["likelihood"] == g["likelihoods"][s]
        rats.append(orc.rating_distribution(test, res["theta"], res["eta"], res["pr"]))
    accs = [orc.prediction_stats(r, test[:, 2], [0, 1, 2, 3, 4])["accuracy"] for r in rats]
    np.testing.assert_array_equal(accs, g["accuracies"])
    assert orc.choose_best(rats, test[:, 2], [0, 1, 2, 3, 4]) == int(g["best"]) == 1
    # best by accuracy is NOT best by likelihood on this fixture
    assert int(np.argmax(g["likelihoods"])) == 0
    np.testing.assert_allclose(np.mean(rats, axis=0), g["prediction"], **TIGHT)


@pytest.mark.parametrize("name,tags", [("medium.npz", [""]), ("wide.npz", ["k20_", "l32_", "odd_"])])
def test_random_problems(golden_dir, name, tags):
    g = _load(golden_dir, name)
    for t in tags:
        data, theta, eta, pr = (g[t + k] for k in ("data", "theta", "eta", "pr"))
        nt, ne, npr = orc.em_sums(data, theta, eta, pr)
        np.testing.assert_array_equal(nt, g[t + "ntheta"])
        np.testing.assert_array_equal(ne, g[t + "neta"])
        np.testing.assert_array_equal(npr, g[t + "npr"])
        fu, fi = orc.degree_factors(data, theta.shape[1], eta.shape[1])
        th1, et1, pr1 = orc.em_iteration(data, theta, eta, pr, fu, fi)
        np.testing.assert_array_equal(th1, g[t + "theta1"])
        np.testing.assert_array_equal(et1, g[t + "eta1"])
        np.testing.assert_array_equal(pr1, g[t + "pr1"])
        assert orc.likelihood(data, theta, eta, pr) == g[t + "likelihood"]
        np.testing.assert_allclose(orc.rating_distribution(data[:257], theta, eta, pr),
                                   g[t + "prod"], **TIGHT)


def test_index_structure_matches_reference_lists(golden_dir):
    """bucket_order == the reference's np.where lists (src/mmsbm.py:114-122)."""
    g = _load(golden_dir, "fixture.npz")
    train = g["train"]
    R = int(train[:, 2].max()) + 1
    rlists = np.split(g["rating_index_concat"], np.cumsum(g["rating_index_len"])[:-1])
    for col, key in ((0, "user"), (1, "item")):
        n_ids = int(train[:, col].max()) + 1
        seg, perm = orc.bucket_order(train[:, col], train[:, 2], n_ids, R)
        assert seg.dtype == np.int32 and perm.dtype == np.int32
        lists = np.split(g[f"{key}_index_concat"], np.cumsum(g[f"{key}_index_len"])[:-1])
        for a in range(n_ids):
            whole = perm[seg[a * R]:seg[(a + 1) * R]]
            np.testing.assert_array_equal(np.sort(whole), lists[a])
            for r in range(R):
                np.testing.assert_array_equal(perm[seg[a * R + r]:seg[a * R + r + 1]],
                                              np.intersect1d(lists[a], rlists[r]))


def test_encoding(golden_dir):
    enc = json.load(open(os.path.join(golden_dir, "encoding.json")))
    tin = enc["train_in"]
    # the golden file stores str(cell); the oracle str()s again, which is idempotent
    out, dicts = orc.encode_train(tin["users"], tin["items"], tin["ratings"])
    assert out.tolist() == enc["train_out"]
    assert list(dicts) == enc["dicts"]
    assert dicts[0] == {"1": 0, "10": 1, "100": 2, "11": 3, "2": 4}   # lexicographic, not numeric
    te = enc["test_in"]
    tout, keep = orc.encode_test(te["users"], te["items"], te["ratings"], dicts)
    assert tout.tolist() == enc["test_out"]
    assert keep.tolist() == [True, False, False, False, True]


def test_empty_rating_level_keeps_zero_slab():
    """src/kernels_numpy.py:74-77 -- a rating level with no rows stays zero, and
    normalize_pr divides a zero row by one (expectation_maximization.py:154)."""
    data = np.array([[0, 0, 0], [1, 1, 2], [1, 0, 2]], dtype=np.int64)
    g = np.random.default_rng(3)
    theta, eta, pr = g.random((2, 3)), g.random((2, 2)), g.random((3, 2, 3))
    _, _, npr = orc.em_sums(data, theta, eta, pr)
    assert np.all(npr[:, :, 1] == 0) and np.all(npr[:, :, 0] > 0)
    z = orc.normalize_pr(np.zeros((2, 2, 3)))
    assert np.all(z == 0)


def test_numba_port_against_reference_numba_kernel_and_numpy_oracle(golden_dir):
    """oracle/mmsbm_oracle_numba.py (the CPU baseline bench.py times next to the numpy port):
    omega against the output of the reference's numba kernel on its own toy, the M-step sums and
    a full iteration against the numpy oracle (fastmath and eps-added-not-clamped: ~1e-12)."""
    onb = pytest.importorskip("oracle.mmsbm_oracle_numba")
    g = _load(golden_dir, "toy_backends.npz")
    for tag in ("omega", "prod"):
        args = [g[f"{tag}_{k}"] for k in ("data", "theta", "eta", "pr")]
        np.testing.assert_allclose(onb.omegas(*args), g[f"{tag}_omegas_numba"], rtol=1e-14)
        for got, want in zip(onb.em_sums(*args), (g[f"{tag}_ntheta"], g[f"{tag}_neta"], g[f"{tag}_npr"])):
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    m = _load(golden_dir, "medium.npz")
    fu, fi = orc.degree_factors(m["data"], m["theta"].shape[1], m["eta"].shape[1])
    want = orc.em_iteration(m["data"], m["theta"], m["eta"], m["pr"], fu, fi)
    for chunk in (None, 1000):
        got = onb.em_iteration(m["data"], m["theta"], m["eta"], m["pr"], fu, fi, chunk=chunk)
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-300)
