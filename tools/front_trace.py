#!/usr/bin/env python
"""Development tool: per-tile clock64 timeline of the fused front kernel (build with AP_NVCC_EXTRA=-DAP_FRONT_TRACE).
    AP_NVCC_EXTRA=-DAP_FRONT_TRACE python -m alphapig_b200.build --force
    AP_FRONT_TRACE_FILE=gpurun_out/front_trace.txt python tools/front_trace.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from alphapig_b200.engine import Engine  # noqa: E402
from alphapig_b200.params import init_params  # noqa: E402

G = 4096
eng = Engine(width=15, height=15, n_in_row=5, n_games=G, n_playout=4, node_capacity=2000)
arg, aux = init_params("simple", 15, 15, seed=0, synthetic_stats=True)
m = dict(arg)
m.update(aux)
eng.net_load("simple", m)
bench.synthetic_positions(eng, G)
eng.search_select(want_path=False)
for _ in range(3):
    eng.net_forward_leaves(fetch=False)
path = os.environ.get("AP_FRONT_TRACE_FILE")
rows = [list(map(int, ln.split())) for ln in open(path)]
mma = {r[1]: r[2:] for r in rows if r[0] == 0}
epi = {r[1]: r[2:] for r in rows if r[0] == 1}
tma = {r[1]: r[2:] for r in rows if r[0] == 2}
print("tile | MMA: c1 wait_a1e wait_f issue | c2 wait_a2e wait_s2 issue | period || EPI: e1 wait_s2e wait_a1 work | e2 wait_a2 ld work")
prev = None
for t in sorted(mma):
    a = mma[t]
    e = epi.get(t, [-1] * 8)
    period = (a[7] - prev) if prev is not None else 0
    prev = a[7]
    print("%3d | %5d %5d %5d | %5d %5d %5d | %6d || %5d %5d %5d | %5d %5d %5d   @%d" % (
        t, a[1] - a[0], a[2] - a[1], a[3] - a[2], a[5] - a[4], a[6] - a[5], a[7] - a[6], period,
        e[1] - e[0], e[2] - e[1], e[3] - e[2], e[5] - e[4], e[6] - e[5], e[7] - e[6], a[0]))
