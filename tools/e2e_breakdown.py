#!/usr/bin/env python
"""Where the e2e (host-buffer) step of bench.py spends its wall clock: per-call timing of the four ABI calls."""
import os
import sys
import time

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
cells, meta = bench.synthetic_positions(eng, G)
for rep in range(3):
    t = [time.perf_counter()]
    eng.boards_import(cells, meta); t.append(time.perf_counter())
    eng.search_advance(-1); t.append(time.perf_counter())
    eng.search_run(NP); t.append(time.perf_counter())
    eng.search_root(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("import %.2f ms, advance %.2f ms, run %.2f ms (device %.2f ms), root %.2f ms" % (d[0], d[1], d[2], eng.search_timing()[0], d[3]))
