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


def test_rollout_restatements_live(monkeypatch):
    """oracle/rollout.py against the reference's own MCTS._evaluate_rollout (mcts_pure.py:138-157): the reference's
    np.random.rand(A) draws are replaced by the uniforms of a given move order (so its arg-max plays exactly that
    order); value and game length must equal rollout_by_play / rollout_by_descent / rollout_sample_numpy for the
    same order."""
    from oracle.rollout import rollout_by_descent, rollout_by_play, rollout_sample_numpy
    ref = refimport.load()
    rs = np.random.RandomState(9)
    for W, H, n, pre in ((15, 15, 5, 0), (15, 15, 5, 40), (8, 8, 5, 6), (6, 6, 4, 0)):
        for trial in range(12):
            while True:
                a = ref.Board(width=W, height=H, n_in_row=n)
                b = OBoard(W, H, n)
                a.init_board()
                b.init_board()
                for _ in range(pre):
                    m = int(a.availables[rs.randint(len(a.availables))])
                    a.do_move(m)
                    b.do_move(m)
                    if a.game_end()[0]:
                        break
                if not a.game_end()[0]:
                    break
            keys = rs.random_sample(len(b.availables))         # one uniform per empty cell
            order = np.array(b.availables)[np.argsort(keys)]    # ascending keys = the order the cells are played in
            key_of = {int(m): 1.0 - float(k) for m, k in zip(b.availables, keys)}  # arg-max picks the smallest key left
            monkeypatch.setattr(ref.mcts_pure.np.random, "rand", lambda A, s=a: np.array([key_of[m] for m in s.availables]))
            stones0 = len(a.states)
            value = ref.mcts_pure.MCTS(ref.mcts_pure.policy_value_fn)._evaluate_rollout(a)
            plies = len(a.states) - stones0
            assert rollout_by_play(b, order) == (value, plies)
            assert rollout_by_descent(b, order) == (value, plies)

            class Fixed(object):
                def random_sample(self, shape, k=keys):
                    return np.asarray(k)[None, :]

            v, p = rollout_sample_numpy(b, 1, Fixed())
            assert (int(v[0]), int(p[0])) == (value, plies)
    monkeypatch.undo()


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


# ---- TrainPipeline data path: augmentation, replay sampling, policy_update ----
class _FakeNet(object):
    """Deterministic stand-in for PolicyValueNet: probabilities drift with every train_step so the KL
    early-stop / lr-multiplier branches are all reachable."""

    def __init__(self, S, drift):
        self.S, self.drift, self.steps, self.lrs = S, drift, 0, []

    def policy_value(self, state_batch):
        B = len(state_batch)
        base = np.stack([np.asarray(s, dtype=np.float64).reshape(9, -1).sum(0) for s in state_batch])
        logits = base + self.drift * self.steps * np.linspace(-1, 1, self.S)[None, :]
        e = np.exp(logits - logits.max(1, keepdims=True))
        return e / e.sum(1, keepdims=True), np.tanh(base.mean(1, keepdims=True) - 0.3 + 0.1 * self.steps)

    def train_step(self, state_batch, mcts_probs, winner_batch, lr):
        self.steps += 1
        self.lrs.append(lr)
        return np.array([1.0 / self.steps]), np.array([2.0 + self.steps])


def _fake_play_data(H, W, n, seed):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        st = (rs.rand(9, H, W) < 0.3).astype(np.float64)
        pi = rs.dirichlet(np.ones(H * W))
        out.append((st, pi, float(rs.choice([-1.0, 0.0, 1.0]))))
    return out


def test_equi_data_live():
    from oracle import pipeline as opl
    ref = refimport.load()
    import types
    for H in (6, 15):
        data = _fake_play_data(H, H, 3, H)
        a = ref.TrainPipeline.get_equi_data(types.SimpleNamespace(board_height=H, board_width=H), data)
        b = opl.equi_data(data, H, H)
        assert len(a) == len(b) == 24
        for (sa, pa_, za), (sb, pb_, zb) in zip(a, b):
            assert np.array_equal(sa, sb) and np.array_equal(pa_, pb_) and za == zb


@pytest.mark.parametrize("drift,mult0", [(0.0, 1.0), (0.05, 1.0), (0.6, 1.0), (3.0, 0.04), (0.0, 25.0)])
def test_policy_update_live(drift, mult0):
    """Reference TrainPipeline.policy_update driven with a fake net vs the oracle restatement: same
    minibatch (random.sample), same number of epochs, same lr multiplier."""
    from collections import deque
    import types
    from oracle import pipeline as opl
    ref = refimport.load()
    H = 6
    games = [_fake_play_data(H, H, 5, 100 + g) for g in range(4)]
    rep = opl.ReplayDeque(150, H, H)
    dq = deque(maxlen=150)
    for g in games:
        rep.extend_game(g)
        dq.extend(ref.TrainPipeline.get_equi_data(types.SimpleNamespace(board_height=H, board_width=H), g))
    assert len(rep) == len(dq) == 150
    na, nb = _FakeNet(H * H, drift), _FakeNet(H * H, drift)
    me = types.SimpleNamespace(data_buffer=dq, batch_size=16, policy_value_net=na, learn_rate=2e-3, lr_multiplier=mult0,
                               epochs=5, kl_targ=0.02)
    random.seed(3)
    la, ea = ref.TrainPipeline.policy_update(me)
    random.seed(3)
    lb, eb, mult, kl, ran, _, _ = opl.policy_update(nb, rep, 16, 2e-3, mult0, 5, 0.02)
    assert np.array_equal(la, lb) and np.array_equal(ea, eb)
    assert me.lr_multiplier == mult and na.steps == nb.steps == ran and na.lrs == nb.lrs


def test_policy_evaluate_live():
    import types
    from oracle import pipeline as opl
    ref = refimport.load()
    seq = [1, 2, -1, 1, 1, 2, 1, -1, 1, 1]
    calls = []

    class G(object):
        def start_play(self, p1, p2, start_player=0, is_shown=1):
            calls.append(start_player)
            return seq[len(calls) - 1]
    me = types.SimpleNamespace(policy_value_net=types.SimpleNamespace(policy_value_fn=None), c_puct=5, n_playout=10,
                               pure_mcts_playout_num=20, game=G())
    ra = ref.TrainPipeline.policy_evaluate(me, n_games=10)
    it = iter(seq)
    rb, cnt = opl.policy_evaluate(lambda a, b, sp: next(it), None, None, 10)
    assert ra == rb == (6 + 0.5 * 2) / 10 and calls == [0, 1] * 5


def test_symbol_fixture_is_current_live():
    """tests/golden/res10_symbol_ops.json is exactly what tests/golden/make_graph_golden.py extracts from the
    reference's committed policy_value_loss.json today."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_graph_golden", os.path.join(here, "golden", "make_graph_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    live = mod.extract(os.path.join(refimport.REF, "policy_value_loss.json"))
    assert live == json.load(open(os.path.join(here, "golden", "res10_symbol_ops.json")))


def test_train_config_defaults_live():
    """The TrainPipeline shim's DEFAULT_CONF carries exactly the keys and values of the reference's own
    conf/train_config.yaml (everything except the logging section), so `TrainPipeline({})` trains with the
    reference's hyper-parameters (train_mxnet.py:37-75 reads the same keys)."""
    import os
    import yaml
    from alphapig_b200.train_mxnet import DEFAULT_CONF
    ref = yaml.safe_load(open(os.path.join(refimport.REF, "conf", "train_config.yaml"), encoding="utf-8"))
    ref.pop("train_logging")
    assert ref == DEFAULT_CONF
