"""ms per EM iteration of ONE run of a BASELINE shape on one GPU (what a rank does when the runs of
ML-1M are sharded one per GPU), with the cooperative small-problem kernel and without.

    python profiles/scripts/one_run.py <workload> [iterations]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmsbm_b200.engine import Engine  # noqa: E402

w = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
U, I, N, K, L, _ = bench.WORKLOADS[w]
data = bench.synth_triples(U, I, N, seed=0)
th0, et0, pr0 = bench.seeded_inits(data, U, I, K, L, bench.child_seeds(1))
out = {"workload": w, "runs": 1, "iterations": T}
for coop in ("1", "0"):
    os.environ["MMSBM_COOP"] = coop
    e = Engine(data, U, I, bench.R, K, L)
    e.set_params(th0, et0, pr0)
    e.run(T)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); e.run(T); e.run(T); b.record(); torch.cuda.synchronize()
    out["coop" + coop] = {"ms_per_iteration": a.elapsed_time(b) / (2 * T), "likelihood": float(e.likelihood()[0])}
print(json.dumps(out))
