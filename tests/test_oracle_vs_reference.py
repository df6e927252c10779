"""CPU, build container only: the oracle restatement vs the UNMODIFIED reference imported live from
/root/reference (skipped where the reference does not exist, e.g. on the GPU box)."""
import random

import numpy as np
import pytest

from oracle import refimport
from oracle.board import OBoard
from oracle.evaluators import EVALUATORS
from oracle.mcts import OMCTSPlayer, OPureMCTSPlayer

pytestmark = pytest.mark.skipif(not refimport.available(), reason="/root/reference not present")


@pytest.mark.parametrize("W,H,n,seed", [(8, 8, 5, 1), (15, 15, 5, 2), (6, 7, 4, 3), (9, 6, 4, 4)])
def test_board_random_games_live(W, H, n, seed):
    ref = refimport.load()
    rs = np.random.RandomState(seed)
    for game in range(6):
        a = ref.Board(width=W, height=H, n_in_row=n)
        b = OBoard(W, H, n)
        sp = game % 2
        a.init_board(sp)
        b.init_board(sp)
        while True:
            assert a.availables == b.availables and a.current_player == b.current_player
            if W == H:  # the reference's current_state indexes [m // width, m % height]; only square boards are coherent
                assert np.array_equal(np.asarray(a.current_state()), np.asarray(b.current_state()))
            m = int(a.availables[rs.randint(len(a.availables))])
            a.do_move(m)
            b.do_move(m)
            ea, eb = a.game_end(), b.game_end()
            assert ea == eb and a.has_a_winner() == b.has_a_winner()
            assert a.states == b.states and a.last_move == b.last_move
            if ea[0]:
                break


@pytest.mark.parametrize("ev,selfplay,temp", [("e2", 1, 1.0), ("e3", 0, 1e-3), ("e1", 1, 0.5)])
def test_mcts_player_live(ev, selfplay, temp):
    ref = refimport.load()
    a = ref.Board(width=7, height=7, n_in_row=4)
    b = OBoard(7, 7, 4)
    a.init_board()
    b.init_board()
    pa = ref.mcts_alphaZero.MCTSPlayer(EVALUATORS[ev], c_puct=5, n_playout=120, is_selfplay=selfplay)
    pb = OMCTSPlayer(EVALUATORS[ev], c_puct=5, n_playout=120, is_selfplay=selfplay)
    for ply in range(8):
        np.random.seed(100 + ply)
        ma, pia = pa.get_action(a, temp=temp, return_prob=1)
        np.random.seed(100 + ply)
        mb, pib = pb.get_action(b, temp=temp, return_prob=1)
        assert ma == mb and np.array_equal(pia, pib)
        a.do_move(ma)
        b.do_move(mb)
        if a.game_end()[0]:
            break


def test_pure_player_live():
    ref = refimport.load()
    a = ref.Board(width=6, height=6, n_in_row=4)
    b = OBoard(6, 6, 4)
    a.init_board()
    b.init_board()
    for m in (14, 15, 20, 21):
        a.do_move(m)
        b.do_move(m)
    np.random.seed(5)
    ma = ref.mcts_pure.MCTSPlayer(c_puct=5, n_playout=150).get_action(a)
    np.random.seed(5)
    mb = OPureMCTSPlayer(c_puct=5, n_playout=150).get_action(b)
    assert ma == mb


def test_self_play_live():
    from oracle import selfplay as osp
    ref = refimport.load()
    a = ref.Board(width=6, height=6, n_in_row=4)
    b = OBoard(6, 6, 4)
    pa = ref.mcts_alphaZero.MCTSPlayer(EVALUATORS["e2"], c_puct=5, n_playout=60, is_selfplay=1)
    pb = OMCTSPlayer(EVALUATORS["e2"], c_puct=5, n_playout=60, is_selfplay=1)
    orig = ref.game_ai.random.random
    ref.game_ai.random.random = lambda: 0.5
    o2 = osp.random.random
    osp.random.random = lambda: 0.5
    try:
        np.random.seed(9)
        random.seed(9)
        wa, da = ref.Game_AI(a).start_self_play(pa, temp=1.0)
        np.random.seed(9)
        random.seed(9)
        wb, db = osp.start_self_play(b, pb, temp=1.0)
    finally:
        ref.game_ai.random.random = orig
        osp.random.random = o2
    da = list(da)
    assert wa == wb and len(da) == len(db)
    for (sa, pa_, za), (sb, pb_, zb) in zip(da, db):
        assert np.array_equal(np.asarray(sa), np.asarray(sb)) and np.array_equal(pa_, pb_) and za == zb


def test_sgf_replay_live():
    """Game.start_self_play (SGF replay, game.py:233-304) vs the oracle restatement, via a fake record."""
    from oracle import selfplay as osp
    ref = refimport.load()
    rec = {"winner": 2, "seq_num_list": [112, 113, 97, 98, 127, 128, 82, 83]}
    ref.sgf_records["fake.sgf"] = rec

    class P:
        def reset_player(self):
            pass
    a = ref.Board(width=15, height=15, n_in_row=5)
    warn, winner, data = ref.Game(a).start_self_play(P(), sgf_home=".", file_name="fake.sgf")
    b = OBoard(15, 15, 5)
    w2, winner2, data2 = osp.sgf_self_play(b, P(), rec)
    data = list(data)
    assert (warn, winner) == (w2, winner2) and len(data) == len(data2)
    for (sa, pa_, za), (sb, pb_, zb) in zip(data, data2):
        assert np.array_equal(np.asarray(sa), np.asarray(sb)) and np.array_equal(pa_, pb_) and za == zb
    # illegal move -> (1, None, None)
    ref.sgf_records["bad.sgf"] = {"winner": 1, "seq_num_list": [3, 3]}
    assert ref.Game(a).start_self_play(P(), sgf_home=".", file_name="bad.sgf") == (1, None, None)
    assert osp.sgf_self_play(b, P(), {"winner": 1, "seq_num_list": [3, 3]}) == (1, None, None)
