#!/usr/bin/env python
"""A/B of the mcts_pure kernel pieces on one GPU: rollouts alone (permutation vs ply by ply) and the fused
search with random rollouts (modes 0 / 2) and with the hashed leaf value (mode 1 = tree work only)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from alphapig_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=8192)
ap.add_argument("--playouts", type=int, default=1000)
a = ap.parse_args()
eng = Engine(width=15, height=15, n_in_row=5, n_games=a.games, c_puct=5, n_playout=a.playouts,
             node_capacity=a.playouts * 225 + 2)
bench.synthetic_positions(eng, a.games)
for impl in (0, 2):
    eng.rollout_eval(seed=0, impl=impl)
    t0 = time.perf_counter()
    reps = 20
    for i in range(reps):
        v, p = eng.rollout_eval(seed=i, impl=impl)
    dt = time.perf_counter() - t0
    print("rollout_eval impl %d: %.1f M rollouts/s (wall, incl. D2H), mean plies %.1f, mean value %.4f"
          % (impl, reps * a.games / dt / 1e6, p.mean(), v.mean()))
for mode in (0, 2, 1):
    eng.pure_run(a.playouts, seed=1, rollout_mode=mode)
    eng.search_stats()
    eng.pure_run(a.playouts, seed=2, rollout_mode=mode)
    ms = eng.search_timing()[0]
    st = eng.search_stats()
    print("pure_run mode %d: %.2f ms, %.1f M playouts/s, scanned/playout %.1f, path %.2f, plies/playout %.1f"
          % (mode, ms, a.games * a.playouts / ms / 1e3, st["children_scanned"] / st["playouts"],
             st["path_nodes"] / st["playouts"], st["rollout_plies"] / st["playouts"]))
