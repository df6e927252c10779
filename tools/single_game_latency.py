#!/usr/bin/env python
"""BASELINE configs[0] shape on the GPU: ONE 8x8 game, policy_value_net_mxnet_simple, n_playout=400, through the
reference-named shims (MCTSPlayer.get_action) - the latency a human-play / evaluation client sees per move."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from alphapig_b200.game import Board  # noqa: E402
from alphapig_b200.mcts_alphaZero import MCTSPlayer  # noqa: E402
from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet  # noqa: E402

for W in (8, 15):
    net = PolicyValueNet(W, W, batch_size=128, seed=0)
    player = MCTSPlayer(net.policy_value_fn, c_puct=5, n_playout=400, is_selfplay=1)
    b = Board(width=W, height=W, n_in_row=5)
    b.init_board(0)
    np.random.seed(0)
    ts = []
    for ply in range(8):
        t0 = time.perf_counter()
        mv = player.get_action(b, temp=1.0)
        ts.append(time.perf_counter() - t0)
        b.do_move(mv)
    ts = np.array(ts[2:])
    print("%dx%d, 1 game, n_playout=400: %.1f ms per move (%.0f playouts/s), min %.1f ms"
          % (W, W, 1e3 * ts.mean(), 400 / ts.mean(), 1e3 * ts.min()))
