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
# round-2 kernels: multi-leaf search with virtual loss, device-side trajectories + outbox, packed ring push, pool growth
# (back on the 6-conv net: the multi-leaf search needs the compacted-batch path the inception variant does not have)
arg, aux = init_params("simple", 15, 15, seed=0, synthetic_stats=True)
merged = dict(arg)
merged.update(aux)
eng.net_load("simple", merged)
eng.search_advance(-1)
eng.search_run_vl(8, 4)
eng.traj_create(outbox_records=4096)
eng.replay_create(8 * 512)
moves, pi = eng.selfplay_pick(1.0, 0.25, 0.3, seed=5, ply=0)
eng.search_advance(moves)
eng.boards_do_move(moves)
ids = np.arange(0, G, 3, dtype=np.int32)
legal = eng.boards_legal(ids)
forced = np.array([int(np.nonzero(r)[0][0]) for r in legal], np.int32)
eng.traj_append_forced(forced, ids)
eng.boards_do_move(forced, ids)
eng.traj_finish(ids, np.where(np.arange(len(ids)) % 3 == 0, -1, 1 + np.arange(len(ids)) % 2).astype(np.int8))
ptr, n, rw = eng.traj_outbox()
assert n == 2 * len(ids)
eng.replay_push_packed(None, n=n, device_ptr=ptr)
eng.traj_outbox_clear()
st, p2, z = eng.replay_gather(np.arange(8 * n))
assert abs(float(p2.sum()) - 8 * n) < 0.1
grow = Engine(width=8, height=8, n_in_row=5, n_games=4, c_puct=5, n_playout=2)  # library-chosen capacity: 2*2*64+66 nodes
arg8, aux8 = init_params("simple", 8, 8, seed=0, synthetic_stats=True)
m8 = dict(arg8)
m8.update(aux8)
grow.net_load("simple", m8)
c0 = grow.node_capacity()
grow.search_run(40)  # needs 40 * 64 free nodes per game: the pools grow first
assert grow.node_capacity() > c0
grow.close()
eng.close()
print("sanitize_small ok")
