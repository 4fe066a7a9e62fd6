"""The reference's loop through the plugin (INTEGRATION.md level 1): ML-1M-shaped data, K = L = 10,
one run, 50 x (kernels_b200.update_coefficients + the three host normalisations, exactly the body
of src/mmsbm.py:244-250), with and without the library's index cache.

    python profiles/scripts/plugin_cache.py [iterations]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
U, I, N, K, L, S = bench.WORKLOADS["ml1m"]
data = bench.synth_triples(U, I, N, seed=0)
th0, et0, pr0 = (a[0] for a in bench.seeded_inits(data, U, I, K, L, bench.child_seeds(1)))
du = np.maximum(np.bincount(data[:, 0], minlength=U), 1)[:, None]
di = np.maximum(np.bincount(data[:, 1], minlength=I), 1)[:, None]
out = {}
for mode in ("0", "1"):
    os.environ["MMSBM_INDEX_CACHE"] = mode
    from mmsbm_b200 import _lib, kernels_b200 as kb
    lib = _lib.load()
    lib.mmsbm_index_cache_clear()
    th, et, pr = th0, et0, pr0
    kb.update_coefficients(data, th, et, pr)            # warm-up: library pool, first index build
    t0 = time.perf_counter()
    for _ in range(T):
        nt, ne, npr = kb.update_coefficients(data, th, et, pr)
        th, et = nt / du, ne / di
        tot = npr.sum(axis=2)
        pr = npr / np.where(tot == 0, 1, tot)[:, :, None]
    dt = time.perf_counter() - t0
    h, m = ctypes.c_int64(), ctypes.c_int64()
    lib.mmsbm_index_cache_stats(ctypes.byref(h), ctypes.byref(m))
    out["cache_on" if mode == "1" else "cache_off"] = {
        "ms_per_iteration": dt / T * 1e3, "rating_updates_per_s": N * T / dt, "hits": h.value, "misses": m.value,
        "checksum": float(th.sum())}
print(json.dumps({"workload": "ml1m, one run, %d iterations through kernels_b200.update_coefficients" % T, **out}))
