#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ FROM THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference ships no tests or fixtures for this path, so these
files -- outputs of the reference's own Board / MCTS / Game_AI classes, driven
with seeded inputs and the deterministic evaluators in oracle/evaluators.py --
are the pins for both the oracle and the CUDA engine.
"""
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import refimport  # noqa: E402
from oracle.evaluators import EVALUATORS, position_key  # noqa: E402

ref = refimport.load()


def random_game(W, H, n, seed, start_player=0):
    """Play uniformly random legal moves on the reference Board until game_end."""
    rs = np.random.RandomState(seed)
    b = ref.Board(width=W, height=H, n_in_row=n)
    b.init_board(start_player)
    moves, ends, winners, feats = [], [], [], []
    while True:
        m = int(b.availables[rs.randint(len(b.availables))])
        feats.append(np.packbits(np.ascontiguousarray(b.current_state()).astype(np.uint8).ravel()))
        b.do_move(m)
        moves.append(m)
        end, winner = b.game_end()
        ends.append(int(end))
        winners.append(int(winner))
        if end:
            return moves, ends, winners, feats


def gen_boards(W, H, n, n_games, seed0):
    out = {"meta": np.array([W, H, n, n_games])}
    L = W * H
    moves = -np.ones((n_games, L), np.int16)
    ends = np.zeros((n_games, L), np.int8)
    wins = np.zeros((n_games, L), np.int8)
    nb = (9 * W * H + 7) // 8
    feats = np.zeros((n_games, L, nb), np.uint8)
    start = np.zeros(n_games, np.int8)
    for g in range(n_games):
        sp = g % 2
        start[g] = sp
        mv, en, wi, ft = random_game(W, H, n, seed0 + g, sp)
        k = len(mv)
        moves[g, :k], ends[g, :k], wins[g, :k] = mv, en, wi
        feats[g, :k] = np.stack(ft)
    out.update(moves=moves, ends=ends, winners=wins, feats=feats, start_player=start)
    return out


def synth_position(W, H, n, seed, max_pairs=31):
    """SURVEY 8(d) synthetic positions: k = 2*randint(0,max_pairs) random legal
    plies from the empty board, redrawn if the game ends."""
    rs = np.random.RandomState(seed)
    while True:
        k = 2 * rs.randint(0, max_pairs)
        b = ref.Board(width=W, height=H, n_in_row=n)
        b.init_board(0)
        mv, ok = [], True
        for _ in range(k):
            m = int(b.availables[rs.randint(len(b.availables))])
            b.do_move(m)
            mv.append(m)
            if b.game_end()[0]:
                ok = False
                break
        if ok:
            return mv


def gen_tree_case(W, H, n, evaluator, n_playout, c_puct, start_moves, n_plies, seed, selfplay, temp):
    """Reference MCTSPlayer over n_plies consecutive moves (tree reuse iff selfplay)."""
    b = ref.Board(width=W, height=H, n_in_row=n)
    b.init_board(0)
    for m in start_moves:
        b.do_move(m)
    player = ref.mcts_alphaZero.MCTSPlayer(EVALUATORS[evaluator], c_puct=c_puct,
                                           n_playout=n_playout, is_selfplay=selfplay)
    np.random.seed(seed)
    S = W * H
    visits = np.zeros((n_plies, S), np.int32)
    qs = np.zeros((n_plies, S), np.float64)
    pis = np.zeros((n_plies, S), np.float64)
    rootn = np.zeros(n_plies, np.int32)
    chosen = -np.ones(n_plies, np.int32)
    done = 0
    for ply in range(n_plies):
        if b.game_end()[0]:
            break
        # run the search exactly as get_action does, but peek at the root first
        mcts = player.mcts
        acts, probs = mcts.get_move_probs(b, temp)
        root = mcts._root
        for a, node in root._children.items():
            visits[ply, a] = node._n_visits
            qs[ply, a] = float(node._Q)
        rootn[ply] = root._n_visits
        pis[ply, list(acts)] = probs
        if selfplay:
            move = np.random.choice(acts, p=0.75 * probs + 0.25 * np.random.dirichlet(0.3 * np.ones(len(probs))))
            mcts.update_with_move(move)
        else:
            move = np.random.choice(acts, p=probs)
            mcts.update_with_move(-1)
        chosen[ply] = move
        b.do_move(int(move))
        done += 1
    return dict(W=W, H=H, n=n, evaluator=evaluator, n_playout=n_playout, c_puct=c_puct,
                start_moves=[int(x) for x in start_moves], n_plies=done, seed=seed,
                selfplay=int(selfplay), temp=temp), dict(
        visits=visits[:done], q=qs[:done], pi=pis[:done], root_n=rootn[:done], chosen=chosen[:done])


def hash_rollout(state):
    """Injected rollout result for mcts_pure bookkeeping parity: the terminal
    handling of mcts_pure.py:143-157 with the random play replaced by a hash."""
    player = state.get_current_player()
    end, winner = state.game_end()
    if not end:
        return int(position_key(state) % 3) - 1
    if winner == -1:
        return 0
    return 1 if winner == player else -1


def gen_pure_case(W, H, n, n_playout, c_puct, start_moves):
    b = ref.Board(width=W, height=H, n_in_row=n)
    b.init_board(0)
    for m in start_moves:
        b.do_move(m)
    mcts = ref.mcts_pure.MCTS(ref.mcts_pure.policy_value_fn, c_puct, n_playout)
    mcts._evaluate_rollout = lambda state, limit=1000: hash_rollout(state)
    move = mcts.get_move(b)
    S = W * H
    visits = np.zeros(S, np.int32)
    qs = np.zeros(S, np.float64)
    for a, node in mcts._root._children.items():
        visits[a] = node._n_visits
        qs[a] = float(node._Q)
    return dict(W=W, H=H, n=n, n_playout=n_playout, c_puct=c_puct,
                start_moves=[int(x) for x in start_moves], move=int(move),
                root_n=int(mcts._root._n_visits)), dict(visits=visits, q=qs)


def gen_selfplay(W, H, n, evaluator, n_playout, seed, temp):
    b = ref.Board(width=W, height=H, n_in_row=n)
    g = ref.Game_AI(b)
    player = ref.mcts_alphaZero.MCTSPlayer(EVALUATORS[evaluator], c_puct=5, n_playout=n_playout, is_selfplay=1)
    np.random.seed(seed)
    random.seed(seed)
    orig = ref.game_ai.random.random
    if W < 15:
        ref.game_ai.random.random = lambda: 0.5  # skip the 15-wide opening (game_ai.py:77-111)
    try:
        winner, data = g.start_self_play(player, temp=temp)
    finally:
        ref.game_ai.random.random = orig
    data = list(data)
    states = np.stack([np.packbits(np.ascontiguousarray(s).astype(np.uint8).ravel()) for s, _, _ in data])
    pis = np.stack([p for _, p, _ in data])
    zs = np.array([z for _, _, z in data])
    moves = np.array([m for m, _ in b.history], np.int32)
    return dict(W=W, H=H, n=n, evaluator=evaluator, n_playout=n_playout, seed=seed, temp=temp,
                winner=int(winner)), dict(states=states, pi=pis, z=zs, moves=moves)


def main():
    np.savez_compressed(os.path.join(HERE, "boards_8x8.npz"), **gen_boards(8, 8, 5, 24, 100))
    np.savez_compressed(os.path.join(HERE, "boards_15x15.npz"), **gen_boards(15, 15, 5, 12, 200))
    np.savez_compressed(os.path.join(HERE, "boards_6x6_4.npz"), **gen_boards(6, 6, 4, 16, 300))

    # full-board tie reachable quickly: 5x5 board with 5-in-row mostly ties
    np.savez_compressed(os.path.join(HERE, "boards_5x5.npz"), **gen_boards(5, 5, 5, 16, 400))

    tree_meta, tree_arrays = [], {}
    cases = [
        # W, H, n, eval, n_playout, c_puct, start, plies, seed, selfplay, temp
        (8, 8, 5, "e1", 400, 5, [], 3, 1, 0, 1e-3),
        (8, 8, 5, "e2", 400, 5, synth_position(8, 8, 5, 11, 8), 6, 2, 1, 1.0),
        (8, 8, 5, "e3", 300, 5, synth_position(8, 8, 5, 12, 8), 6, 3, 1, 1.0),
        (15, 15, 5, "e1", 400, 5, [], 2, 4, 1, 1.0),
        (15, 15, 5, "e2", 400, 5, synth_position(15, 15, 5, 1234), 4, 5, 1, 1.0),
        (15, 15, 5, "e2", 400, 3, synth_position(15, 15, 5, 1235), 4, 6, 1, 1e-3),
        (15, 15, 5, "e3", 200, 5, synth_position(15, 15, 5, 1236), 4, 7, 1, 1.0),
        (6, 6, 4, "e2", 500, 5, synth_position(6, 6, 4, 21, 6), 12, 8, 1, 1.0),
        (5, 5, 5, "e2", 200, 5, synth_position(5, 5, 5, 22, 9), 10, 9, 1, 1.0),
    ]
    for i, c in enumerate(cases):
        meta, arrs = gen_tree_case(*c)
        tree_meta.append(meta)
        for k, v in arrs.items():
            tree_arrays["c%d_%s" % (i, k)] = v
        print("tree case", i, meta["evaluator"], meta["W"], "plies", meta["n_plies"], flush=True)
    np.savez_compressed(os.path.join(HERE, "tree_cases.npz"), **tree_arrays)
    json.dump(tree_meta, open(os.path.join(HERE, "tree_cases.json"), "w"), indent=1)

    pure_meta, pure_arrays = [], {}
    pcases = [
        (8, 8, 5, 1000, 5, synth_position(8, 8, 5, 31, 8)),
        (15, 15, 5, 1000, 5, synth_position(15, 15, 5, 1300)),
        (6, 6, 4, 800, 5, synth_position(6, 6, 4, 32, 8)),
    ]
    for i, c in enumerate(pcases):
        meta, arrs = gen_pure_case(*c)
        pure_meta.append(meta)
        for k, v in arrs.items():
            pure_arrays["p%d_%s" % (i, k)] = v
        print("pure case", i, meta["move"], flush=True)
    np.savez_compressed(os.path.join(HERE, "pure_cases.npz"), **pure_arrays)
    json.dump(pure_meta, open(os.path.join(HERE, "pure_cases.json"), "w"), indent=1)

    sp_meta, sp_arrays = [], {}
    for i, c in enumerate([(8, 8, 5, "e2", 120, 5, 1.0), (6, 6, 4, "e3", 100, 6, 1.0),
                           (8, 8, 5, "e1", 60, 7, 1e-3)]):
        meta, arrs = gen_selfplay(*c)
        sp_meta.append(meta)
        for k, v in arrs.items():
            sp_arrays["s%d_%s" % (i, k)] = v
        print("selfplay case", i, "winner", meta["winner"], "plies", len(arrs["moves"]), flush=True)
    np.savez_compressed(os.path.join(HERE, "selfplay_cases.npz"), **sp_arrays)
    json.dump(sp_meta, open(os.path.join(HERE, "selfplay_cases.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
