#!/usr/bin/env python
"""Cost of the per-phase CUDA events inside ap_search_run: the bench workload with ap_search_profile on / off."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from alphapig_b200.engine import Engine  # noqa: E402
from alphapig_b200.params import init_params  # noqa: E402

arg, aux = init_params("simple", 15, 15, seed=0, synthetic_stats=True)
merged = dict(arg)
merged.update(aux)
G, NP = 4096, 400
eng = Engine(width=15, height=15, n_in_row=5, n_games=G, c_puct=5, n_playout=NP, node_capacity=NP * 225 + 2)
eng.net_load("simple", merged)
bench.synthetic_positions(eng, G)
for prof in (True, False, True, False):
    eng.search_profile(prof)
    ms = []
    for _ in range(3):
        eng.search_advance(-1)
        eng.search_run(NP)
        ms.append(eng.search_timing()[0])
    print("profile events %s: ms per move search %s -> %.3f M playouts/s" % (prof, np.round(ms, 1), G * NP / min(ms) / 1e3))
