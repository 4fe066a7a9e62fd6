"""One rank's share of a sharded iteration, alone on one GPU (no peers, no NVLink traffic):
what the kernels of rank r of W cost when nothing else is going on.  Compare with the per-rank
stage times bench.py reports from a real W-GPU run (modes.sharded.stage_ms_when_profiled).

    python profiles/scripts/rank_compute.py <workload> <world> [rank] [iterations]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmsbm_b200.parallel import ShardedEngine  # noqa: E402

w, world = sys.argv[1], int(sys.argv[2])
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 40
U, I, N, K, L, S = bench.WORKLOADS[w]
data = bench.synth_triples(U, I, N, seed=0, ids=os.environ.get("IDS", "uniform"))
th0, et0, pr0 = bench.seeded_inits(data, U, I, K, L, bench.child_seeds(S))
torch.cuda.set_device(0)
sh = ShardedEngine(data, U, I, bench.R, K, L, pretend=(rank, world))
sh.set_params(th0, et0, pr0)
sh.run(4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sh.run(iters); e1.record(); torch.cuda.synchronize()
prof = sh.run(10, prof=True)
names = ("iteration", "wait_n_pr", "p_tables_w", "pass_1", "n_publish_pr_1", "pass_2", "n_publish_2", "host_issue")
print(json.dumps({"workload": w, "rank": rank, "of": world, "own_users": sh.Uo, "own_items": sh.Io,
                  "ratings_of_own_users": sh.Nu, "ratings_of_own_items": sh.Ni,
                  "ms_per_iteration": e0.elapsed_time(e1) / iters,
                  "stage_ms_when_profiled": dict(zip(names, [round(x, 4) for x in prof]))}))
