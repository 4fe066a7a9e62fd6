"""Per-kernel time of ONE rank's share when ratings are sharded by user over `world` ranks
(no collective): what a rank computes per iteration, on one GPU."""
import ctypes, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from mmsbm_b200 import _lib
from mmsbm_b200.engine import Engine
from mmsbm_b200.parallel import shard_rows_by_user
wl, world = sys.argv[1], int(sys.argv[2])
U, I, N, K, L, S = bench.WORKLOADS[wl]
R = bench.R
data = bench.synth_triples(U, I, N, seed=0)
local, lo, hi, _ = shard_rows_by_user(data, U, 0, world)
seeds = np.random.default_rng(1).bit_generator._seed_seq.spawn(S)
th0, et0, pr0 = bench.seeded_inits(data, U, I, K, L, seeds)
eng = Engine(local, hi - lo, I, R, K, L)
eng.set_params(th0[:, lo:hi], et0, pr0)
lib = eng.lib
ms4 = (ctypes.c_float * 7)(); acc = np.zeros(7)
for rep in range(8):
    b = eng._alt
    _lib.check(lib.mmsbm_em_step_profiled(*eng._graph_args(), eng.N, hi - lo, I, R, K, L, S, eng.theta.data_ptr(),
               eng.eta.data_ptr(), eng.pr.data_ptr(), b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), 0,
               eng._ws.data_ptr(), eng._ws_bytes, eng._stream(), ctypes.addressof(ms4)), "prof")
    if rep >= 3: acc += np.array(list(ms4))
    eng.swap()
names = ["p_tables_w", "by_user", "n_users", "by_item", "n_items", "pr_accumulate", "pr_finalize"]
k = acc / 5
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.run(10); e0.record(); eng.run(40); e1.record(); torch.cuda.synchronize()
print(wl, "world", world, "local N", eng.N, "sum %.3f ms" % k.sum(), "run %.3f ms/iter" % (e0.elapsed_time(e1) / 40), {n: round(float(v), 3) for n, v in zip(names, k)})
