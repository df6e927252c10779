#!/usr/bin/env python
"""Golden vectors for the TrainPipeline data path, written FROM THE UNMODIFIED REFERENCE
(train_mxnet.TrainPipeline.get_equi_data, deque(maxlen) + random.sample, Game.start_self_play on an SGF record).
Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_pipeline.py
Output: tests/golden/pipeline_cases.npz -- pins oracle/pipeline.py and the device replay ring on boxes where
the reference does not exist."""
import os
import random
import sys
import types
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import refimport  # noqa: E402

ref = refimport.load()


def play_data(W, n, seed):
    """(state, pi, z) of n pseudo positions: real Board.current_state() planes from the reference Board"""
    rs = np.random.RandomState(seed)
    b = ref.Board(width=W, height=W, n_in_row=4 if W < 8 else 5)
    b.init_board(0)
    out = []
    for _ in range(n):
        m = int(b.availables[rs.randint(len(b.availables))])
        pi = rs.dirichlet(np.ones(W * W)).astype(np.float32).astype(np.float64)
        out.append((np.ascontiguousarray(b.current_state()), pi, float(rs.choice([-1.0, 0.0, 1.0]))))
        b.do_move(m)
    return out


def main():
    arrs = {}
    for ci, (W, maxlen, lens) in enumerate([(6, 150, (5, 7, 9, 4)), (15, 400, (12, 20, 25))]):
        me = types.SimpleNamespace(board_height=W, board_width=W)
        dq = deque(maxlen=maxlen)
        games = [play_data(W, n, 1000 * ci + g) for g, n in enumerate(lens)]
        for g in games:
            dq.extend(ref.TrainPipeline.get_equi_data(me, g))
        arrs["c%d_meta" % ci] = np.array([W, maxlen, len(lens)])
        arrs["c%d_lens" % ci] = np.array(lens)
        arrs["c%d_in_states" % ci] = np.concatenate([np.stack([np.packbits(s.astype(np.uint8).ravel()) for s, _, _ in g]) for g in games])
        arrs["c%d_in_pi" % ci] = np.concatenate([np.stack([p for _, p, _ in g]) for g in games])
        arrs["c%d_in_z" % ci] = np.concatenate([np.array([z for _, _, z in g]) for g in games])
        arrs["c%d_dq_states" % ci] = np.stack([np.packbits(np.asarray(s).astype(np.uint8).ravel()) for s, _, _ in dq])
        arrs["c%d_dq_pi" % ci] = np.stack([p for _, p, _ in dq])
        arrs["c%d_dq_z" % ci] = np.array([z for _, _, z in dq])
        random.seed(7 + ci)
        mini = random.sample(dq, 16)
        # position of every sampled tuple in the deque (identity comparison: tuples are unique objects)
        pos = {id(t): i for i, t in enumerate(dq)}
        arrs["c%d_sample_idx" % ci] = np.array([pos[id(t)] for t in mini])
    # SGF replay (game.py:233-304) of one record
    rec = {"winner": 2, "seq_num_list": [112, 113, 97, 98, 127, 128, 82, 83, 67]}
    ref.sgf_records["g.sgf"] = rec

    class P(object):
        def reset_player(self):
            pass
    warn, winner, data = ref.Game(ref.Board(width=15, height=15, n_in_row=5)).start_self_play(P(), sgf_home=".", file_name="g.sgf")
    data = list(data)
    arrs["sgf_moves"] = np.array(rec["seq_num_list"])
    arrs["sgf_winner"] = np.array([winner, warn])
    arrs["sgf_states"] = np.stack([np.packbits(np.asarray(s).astype(np.uint8).ravel()) for s, _, _ in data])
    arrs["sgf_pi"] = np.stack([p for _, p, _ in data])
    arrs["sgf_z"] = np.array([z for _, _, z in data])
    np.savez_compressed(os.path.join(HERE, "pipeline_cases.npz"), **arrs)
    print("wrote pipeline_cases.npz", {k: v.shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
