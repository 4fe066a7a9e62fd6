"""How much of the by-user pass is per-level overhead?  Same ML-20M problem, ratings all at one level."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from mmsbm_b200 import _lib
from mmsbm_b200.engine import Engine
U, I, N, K, L, S = bench.WORKLOADS["ml20m"]
R = bench.R
data = bench.synth_triples(U, I, N, seed=0)
seeds = np.random.default_rng(1).bit_generator._seed_seq.spawn(S)
th0, et0, pr0 = bench.seeded_inits(data, U, I, K, L, seeds)
for mode in ("five levels", "one level", "two levels"):
    d = data.copy()
    if mode == "one level": d[:, 2] = 0
    if mode == "two levels": d[:, 2] = d[:, 2] % 2
    eng = Engine(d, U, I, R, K, L)
    eng.set_params(th0, et0, pr0)
    ms4 = (ctypes.c_float * 7)(); acc = np.zeros(7)
    for rep in range(8):
        b = eng._alt
        _lib.check(eng.lib.mmsbm_em_step_profiled(*eng._graph_args(), eng.N, U, I, R, K, L, S, eng.theta.data_ptr(),
                   eng.eta.data_ptr(), eng.pr.data_ptr(), b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), 0,
                   eng._ws.data_ptr(), eng._ws_bytes, eng._stream(), ctypes.addressof(ms4)), "prof")
        if rep >= 3: acc += np.array(list(ms4))
    k = acc / 5
    print(mode, "by_user %.3f by_item %.3f" % (k[1], k[3]))
    del eng
