#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from alphapig_b200.engine import Engine  # noqa: E402
from alphapig_b200.params import init_params  # noqa: E402

G = 24
eng = Engine(width=15, height=15, n_in_row=5, n_games=G, c_puct=5, n_playout=40, node_capacity=40 * 225 + 2)
bench.synthetic_positions(eng, G)
for mode in (0, 1, 2):
    mv = eng.pure_run(40, seed=3, rollout_mode=mode)
    assert mv.min() >= 0
v, p = eng.rollout_eval(seed=1, impl=0)
v, p = eng.rollout_eval(seed=1, impl=2)
keys = np.random.RandomState(0).randint(0, 1 << 24, size=(G, 256)).astype(np.uint32)
eng.rollout_eval_keys(keys)
for arch, nb in (("simple", 0), ("resnet", 2), ("inception", 1)):
    arg, aux = init_params(arch, 15, 15, n_blocks=max(nb, 1), seed=0, synthetic_stats=True)
    merged = dict(arg)
    merged.update(aux)
    eng.net_load(arch, merged, n_blocks=nb)
    eng.search_advance(-1)
    eng.search_run(6)
    c, a, vis, _, rn = eng.search_root()
    assert int(rn.min()) == 6
eng.close()
print("sanitize_small ok")
