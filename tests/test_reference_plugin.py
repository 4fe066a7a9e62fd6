"""INTEGRATION.md level 1, for real: the UNMODIFIED reference (baseline/_ref, installed by
oracle/install_reference.py) runs ITS OWN ``MMSBM`` with ``backend="b200"``: its loader
(src/backend.py:16-22) imports this repo's ``kernels_b200`` plugin, its spawn pool
(src/mmsbm.py:182-185) pickles the model into worker processes, and its loop
(src/mmsbm.py:243-250) calls the plugin's ``update_coefficients`` every iteration.  The
reference's own known answers must come out (tests/test_mmsbm.py:53-81 of the reference).

Runs in a subprocess because the reference's flat module names (``mmsbm``, ``helpers``,
``logger`` ...) must not leak into this test session."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

SCRIPT = r'''
import json, os, sys
root, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, ref)            # the reference's flat modules: mmsbm, backend, ...
sys.path.insert(1, root)           # kernels_b200.py (the plugin) and the mmsbm_b200 package
import numpy as np, pandas as pd
os.chdir(sys.argv[3])

def mock_data(seed, n=100):        # the reference's fixture, tests/test_mmsbm.py:12-22
    rng = np.random.default_rng(seed)
    return pd.DataFrame({
        "users": [f"user{rng.choice(list(range(5)))}" for _ in range(n)],
        "items": [f"item{rng.choice(list(range(10)))}" for _ in range(n)],
        "ratings": [rng.choice(list(range(1, 6))) for _ in range(n)]})

if __name__ == "__main__":
    import mmsbm as ref_mmsbm
    assert os.path.dirname(os.path.abspath(ref_mmsbm.__file__)) == os.path.abspath(ref)
    from mmsbm_b200 import _lib
    out = {}
    for sampling in (1, 3):
        m = ref_mmsbm.MMSBM(2, 2, iterations=10, sampling=sampling, seed=1, backend="b200")
        m.fit(mock_data(1), silent=True)              # through the reference's spawn pool
        assert m.em._backend == "b200"
        m.predict(mock_data(2))
        s = m.score(silent=True)["stats"]
        out[str(sampling)] = {k: float(v) for k, v in s.items()}
        out[str(sampling)]["theta0"] = float(m.theta.sum(axis=0).iloc[0])
        out[str(sampling)]["psum"] = float(m.prediction_matrix.sum())
    # the reference's own loop in THIS process: the plugin's index cache must hit after call one
    m = ref_mmsbm.MMSBM(2, 2, iterations=10, seed=1, backend="b200")
    m.data_handler = __import__("data_handler").DataHandler()
    train = m.data_handler.format_train_data(mock_data(1))
    m._prepare_objects(train)
    import ctypes
    lib = _lib.load()
    h0, m0 = ctypes.c_int64(), ctypes.c_int64()
    lib.mmsbm_index_cache_stats(ctypes.byref(h0), ctypes.byref(m0))
    res = m.run_one_sampling(train, m.child_states[0], 0)
    h1, m1 = ctypes.c_int64(), ctypes.c_int64()
    lib.mmsbm_index_cache_stats(ctypes.byref(h1), ctypes.byref(m1))
    out["cache"] = {"hits": h1.value - h0.value, "misses": m1.value - m0.value}
    out["inproc_likelihood"] = float(res["likelihood"])
    print("RESULT " + json.dumps(out))
'''


def test_reference_mmsbm_runs_on_the_b200_plugin(tmp_path):
    if not os.path.exists(os.path.join(REF, "mmsbm.py")):
        pytest.skip("baseline/_ref not installed (python oracle/install_reference.py needs /root/reference)")
    script = tmp_path / "drive_reference.py"
    script.write_text(SCRIPT)
    res = subprocess.run([sys.executable, str(script), ROOT, REF, str(tmp_path)], capture_output=True, text=True,
                         timeout=900, cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    one = out["1"]
    # the reference's known answers for its fixture (measured with its numpy backend)
    assert one["accuracy"] == pytest.approx(0.13, abs=1e-12)
    assert one["one_off_accuracy"] == pytest.approx(0.55, abs=1e-12)
    assert one["mae"] == pytest.approx(0.78, abs=1e-12)
    assert one["s2"] == 153
    assert one["s2pond"] == pytest.approx(129.4766730930339, rel=1e-9)
    assert one["likelihood"] == pytest.approx(-13.773187406968459, rel=1e-8)
    assert one["theta0"] == pytest.approx(2.1112326760042786, rel=1e-9)
    assert one["psum"] == pytest.approx(100.0, rel=1e-9)
    # sampling=3 through the pool: three worker processes; the reference picks run 1 (likelihood
    # -17.739...) and scores the mean prediction matrix (tests/golden/sampling3.npz, made by the
    # reference's numpy backend)
    three = out["3"]
    assert three["accuracy"] == pytest.approx(0.10, abs=1e-12) and three["s2"] == 164
    assert three["likelihood"] == pytest.approx(-17.73915051, rel=1e-8)
    assert three["s2pond"] == pytest.approx(126.599259, rel=1e-7)
    # 10 update_coefficients calls on one array: one index build, then nine hits (the reference's
    # compute_likelihood goes through compute_omegas, which needs no index)
    assert out["cache"]["misses"] == 1 and out["cache"]["hits"] == 9
    assert out["inproc_likelihood"] == pytest.approx(-13.773187406968459, rel=1e-8)
