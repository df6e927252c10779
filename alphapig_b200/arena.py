"""Batched arena: ``Game.start_play(alphazero_player, pure_mcts_player, start_player=i % 2)`` (game.py:204-230,
train_mxnet.py:239-263) for many games at once on the device.

Games are split into two groups by who moves first, so that inside a group every live game is in the same
phase: the AlphaZero group-move is one ``ap_search_run`` (n_playout lock-steps, tree reset first: the arena
player has ``is_selfplay=0`` and calls ``update_with_move(-1)`` after every move, mcts_alphaZero.py:203-209),
the pure-MCTS group-move is one ``ap_pure_run`` (fused rollout search, mcts_pure.py:159-169).  The AlphaZero
move is drawn from ``softmax(1/temp * log(visits))`` with the default ``temp=1e-3`` (mcts_alphaZero.py:187),
i.e. the most visited move up to exact ties.
"""
import numpy as np

from .engine import Engine
from .selfplay import visit_softmax


def batched_arena(net, n_games, n_playout=400, c_puct=5, pure_n_playout=1000, n_in_row=5, seed=0, temp=1e-3):
    """-> winners int[n_games] in {1, 2, -1}; player 1 = AlphaZero search with ``net``, player 2 = pure MCTS;
    game i is started by player ``i % 2 + 1``... precisely ``start_player = i % 2`` (0: player 1 first)."""
    rs = np.random.RandomState(seed)
    W, H = net.board_width, net.board_height
    winners = np.zeros(n_games, np.int32)
    for first in (0, 1):
        ids = np.arange(first, n_games, 2)
        G = len(ids)
        if not G:
            continue
        az = net.search_engine(n_in_row=n_in_row, c_puct=c_puct, n_playout=n_playout, n_games=G,
                               node_capacity=n_playout * W * H + 2)
        pure = Engine(width=W, height=H, n_in_row=n_in_row, n_games=G, c_puct=5, n_playout=pure_n_playout,
                      node_capacity=pure_n_playout * W * H + 2, device=net._device)
        az.boards_reset(start_player=first)
        live = np.ones(G, bool)
        win = np.full(G, -1, np.int32)
        to_move = 1 if first == 0 else 2
        S = W * H
        while live.any():
            # finished games cost nothing: both search kernels skip them (ap_search_set_active)
            if to_move == 1:
                az.search_set_active(live)
                az.search_advance(-1)
                az.search_run(n_playout)
                count, acts, visits, _, _ = az.search_root()
                probs = visit_softmax(visits, np.maximum(count, 1), temp)
                cdf = np.cumsum(probs, axis=1)
                u = rs.random_sample(G) * cdf[:, -1]
                idx = np.minimum((cdf <= u[:, None]).sum(axis=1), np.maximum(count, 1) - 1)
                moves = acts[np.arange(G), idx].astype(np.int32)
            else:
                cells, meta = az.boards_export()
                pure.boards_import(cells, meta)
                pure.search_set_active(live)
                moves = pure.pure_run(pure_n_playout, seed=int(rs.randint(0, 2 ** 31 - 1)), rollout_mode=0).astype(np.int32)
            lid = np.nonzero(live)[0].astype(np.int32)
            az.boards_do_move(moves[lid], lid)
            end, w = az.boards_status()
            done = live & end
            win[done] = w[done]
            live &= ~end
            to_move = 3 - to_move
        winners[ids] = win
        az.search_set_active(None)
        pure.close()
    return winners
