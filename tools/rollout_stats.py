#!/usr/bin/env python
"""High-power comparison of the two device rollout implementations (permutation vs ply by ply): ~1M rollouts each
from the empty 15x15 board and from SURVEY 8(d) mid-game positions; prints means with standard errors."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from alphapig_b200.engine import Engine  # noqa: E402

G = 16384
eng = Engine(width=15, height=15, n_in_row=5, n_games=G, c_puct=5, n_playout=1, node_capacity=4)
for name in ("empty", "midgame"):
    if name == "midgame":
        bench.synthetic_positions(eng, G)
    for impl in (0, 2):
        vs, ps = [], []
        for seed in range(64):
            v, p = eng.rollout_eval(seed=seed * 7 + impl, impl=impl)
            vs.append(v), ps.append(p)
        v = np.concatenate(vs).astype(np.float64)
        p = np.concatenate(ps).astype(np.float64)
        n = len(v)
        print("%-8s impl %d: n=%d  value %+.5f +- %.5f  win %.5f loss %.5f tie %.5f  plies %.3f +- %.3f  sd %.3f"
              % (name, impl, n, v.mean(), v.std() / np.sqrt(n), (v == 1).mean(), (v == -1).mean(), (v == 0).mean(),
                 p.mean(), p.std() / np.sqrt(n), p.std()))
