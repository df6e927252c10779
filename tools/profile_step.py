#!/usr/bin/env python
"""Short run of the hot path for ncu: G games, a few lock-steps of ap_search_run.
    ncu ... python tools/profile_step.py [--games 4096] [--playouts 6] [--arch simple]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from alphapig_b200.engine import Engine  # noqa: E402
from alphapig_b200.params import init_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--playouts", type=int, default=6)
ap.add_argument("--arch", default="simple")
ap.add_argument("--blocks", type=int, default=10)
ap.add_argument("--precision", default="auto")
ap.add_argument("--pure", type=int, default=-1, help="profile ap_pure_run with this rollout mode instead (0 / 2)")
a = ap.parse_args()
if a.pure >= 0:
    eng = Engine(width=15, height=15, n_in_row=5, n_games=a.games, c_puct=5, n_playout=a.playouts,
                 node_capacity=a.playouts * 225 + 2)
    bench.synthetic_positions(eng, a.games)
    eng.pure_run(a.playouts, seed=1, rollout_mode=a.pure)
    print("pure total ms", eng.search_timing()[0], eng.search_stats())
    sys.exit(0)
arg, aux = init_params(a.arch, 15, 15, n_blocks=a.blocks, seed=0, synthetic_stats=True)
merged = dict(arg)
merged.update(aux)
eng = Engine(width=15, height=15, n_in_row=5, n_games=a.games, c_puct=5, n_playout=a.playouts,
             node_capacity=a.playouts * 225 + 2)
eng.net_load(a.arch, merged, n_blocks=a.blocks, precision=a.precision)
bench.synthetic_positions(eng, a.games)
eng.search_profile(True)
eng.search_run(a.playouts)
print("total ms", eng.search_timing()[0], "phase ms", np.round(eng.search_profile(True), 3))
