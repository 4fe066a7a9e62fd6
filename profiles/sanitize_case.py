"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): index build with a long
segment (several pieces), EM steps through every launch strategy, likelihood, prod_dist, stats."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mmsbm_b200.engine import Engine, predict_stats
from tests.util import random_params, random_triples

data = random_triples(5, 9000, 60, 45, 5, heavy_tail=True)
data[200:6000, 0] = 3                         # one user with ~5800 rows: 3 pieces
theta, eta, pr = random_params(7, 60, 45, 20, 20, 5, S=2)
for env in ({"MMSBM_NO_GRAPH": "1", "MMSBM_NO_OVERLAP": "1"}, {"MMSBM_FORCE_OVERLAP": "1"}):
    for k in ("MMSBM_NO_GRAPH", "MMSBM_NO_OVERLAP", "MMSBM_FORCE_OVERLAP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    e = Engine(data, 60, 45, 5, 20, 20)
    e.set_params(theta, eta, pr)
    e.run(3)
    print(env, e.likelihood())
rat = e.prod_dist_device(data[:500])
print(predict_stats(rat, data[:500, 2])[0])
theta, eta, pr = random_params(9, 60, 45, 7, 33, 5, S=1)     # odd sizes: padded rows, CH=2 path
e = Engine(data, 60, 45, 5, 7, 33)
e.set_params(theta, eta, pr)
e.run(2)
print(e.likelihood())
