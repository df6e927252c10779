"""GPU: the reference-named Python shims (Board, Game, Game_AI, MCTSPlayer, mcts_pure.MCTSPlayer,
PolicyValueNet) driven exactly like the reference's own scripts drive them, checked against the
golden vectors written from the reference and against the oracle."""
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import net as onet
from oracle import selfplay as osp
from oracle.board import OBoard
from oracle.evaluators import EVALUATORS
from oracle.mcts import OMCTSPlayer

pytestmark = pytest.mark.gpu


def _packed(state):
    return np.packbits(np.ascontiguousarray(state).astype(np.uint8).ravel())


@pytest.mark.parametrize("name", ["boards_8x8", "boards_5x5", "boards_15x15"])
def test_board_shim_golden(name):
    from alphapig_b200.game import Board
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    W, H, n, G = [int(x) for x in z["meta"]]
    for g in range(min(G, 6)):
        b = Board(width=W, height=H, n_in_row=n)
        b.init_board(int(z["start_player"][g]))
        for k in range(W * H):
            m = int(z["moves"][g, k])
            if m < 0:
                break
            st = b.current_state()
            assert st.dtype == np.float64 and st.shape == (9, W, H)
            assert np.array_equal(_packed(st), z["feats"][g, k][: (9 * W * H + 7) // 8])
            b.do_move(m)
            end, winner = b.game_end()
            assert (int(end), int(winner)) == (int(z["ends"][g, k]), int(z["winners"][g, k]))
            win, who = b.has_a_winner()
            assert win == (end and winner != -1) and (who == winner if win else who == -1)


def test_mcts_player_shim_golden():
    from alphapig_b200.game import Board
    from alphapig_b200.mcts_alphaZero import MCTSPlayer
    metas = json.load(open(os.path.join(GOLDEN, "tree_cases.json")))
    z = np.load(os.path.join(GOLDEN, "tree_cases.npz"))
    for i, meta in enumerate(metas):
        if meta["W"] == 15 and meta["n_playout"] > 200:
            continue  # covered through the C ABI in test_gpu_tree.py; the per-playout host callback is slow
        b = Board(width=meta["W"], height=meta["H"], n_in_row=meta["n"])
        b.init_board(0)
        for m in meta["start_moves"]:
            b.do_move(m)
        player = MCTSPlayer(EVALUATORS[meta["evaluator"]], c_puct=meta["c_puct"], n_playout=meta["n_playout"],
                            is_selfplay=meta["selfplay"])
        np.random.seed(meta["seed"])
        for ply in range(meta["n_plies"]):
            move, pi = player.get_action(b, temp=meta["temp"], return_prob=1)
            assert np.array_equal(pi, z["c%d_pi" % i][ply]), "case %d ply %d" % (i, ply)
            assert int(move) == int(z["c%d_chosen" % i][ply])
            b.do_move(int(move))


def test_game_ai_self_play_shim_golden():
    from alphapig_b200 import game_ai
    from alphapig_b200.game import Board
    from alphapig_b200.mcts_alphaZero import MCTSPlayer
    metas = json.load(open(os.path.join(GOLDEN, "selfplay_cases.json")))
    z = np.load(os.path.join(GOLDEN, "selfplay_cases.npz"))
    for i, meta in enumerate(metas):
        b = Board(width=meta["W"], height=meta["H"], n_in_row=meta["n"])
        g = game_ai.Game_AI(b)
        player = MCTSPlayer(EVALUATORS[meta["evaluator"]], c_puct=5, n_playout=meta["n_playout"], is_selfplay=1)
        np.random.seed(meta["seed"])
        random.seed(meta["seed"])
        orig = game_ai.random.random
        game_ai.random.random = lambda: 0.5
        try:
            winner, data = g.start_self_play(player, temp=meta["temp"])
        finally:
            game_ai.random.random = orig
        data = list(data)
        assert winner == meta["winner"]
        assert [m for m, _ in b.history] == list(z["s%d_moves" % i])
        assert np.array_equal(np.stack([p for _, p, _ in data]), z["s%d_pi" % i])
        assert np.array_equal(np.array([zz for _, _, zz in data]), z["s%d_z" % i])
        assert np.array_equal(np.stack([_packed(s) for s, _, _ in data]), z["s%d_states" % i])


def test_forced_opening_branch_15x15():
    """game_ai.py:77-111: with probability 0.09 a random two-ply opening is recorded with one-hot-ish pi."""
    from alphapig_b200 import game_ai
    from alphapig_b200.game import Board
    from alphapig_b200.mcts_alphaZero import MCTSPlayer
    b = Board(width=15, height=15, n_in_row=5)
    ob = OBoard(15, 15, 5)
    for shim in (True, False):
        np.random.seed(3)
        random.seed(3)
        if shim:
            orig = game_ai.random.random
            game_ai.random.random = lambda: 0.01
            try:
                w1, d1 = game_ai.Game_AI(b).start_self_play(
                    MCTSPlayer(EVALUATORS["e3"], c_puct=5, n_playout=12, is_selfplay=1), temp=1.0)
            finally:
                game_ai.random.random = orig
            d1 = list(d1)
        else:
            orig = osp.random.random
            osp.random.random = lambda: 0.01
            try:
                w2, d2 = osp.start_self_play(ob, OMCTSPlayer(EVALUATORS["e3"], c_puct=5, n_playout=12, is_selfplay=1),
                                             temp=1.0)
            finally:
                osp.random.random = orig
    assert w1 == w2 and len(d1) == len(d2)
    assert d1[0][1].max() == 0.99999 and np.isclose(d1[0][1].min(), 1e-6)
    for (sa, pa, za), (sb, pb, zb) in zip(d1, d2):
        assert np.array_equal(sa, sb) and np.array_equal(pa, pb) and za == zb


def test_start_play_and_sgf_replay_shims():
    from alphapig_b200.game import Board, Game
    from alphapig_b200.mcts_alphaZero import MCTSPlayer
    b = Board(width=6, height=6, n_in_row=4)
    game = Game(b)
    with pytest.raises(Exception):
        game.start_play(None, None, start_player=2)
    np.random.seed(11)
    w1 = game.start_play(MCTSPlayer(EVALUATORS["e2"], 5, 40), MCTSPlayer(EVALUATORS["e3"], 5, 40), start_player=1,
                         is_shown=0)
    ob = OBoard(6, 6, 4)
    np.random.seed(11)
    w2 = osp.start_play(ob, OMCTSPlayer(EVALUATORS["e2"], 5, 40), OMCTSPlayer(EVALUATORS["e3"], 5, 40), start_player=1)
    assert w1 == w2 and b.history == ob.history
    # SGF replay (game.py:233-304)
    rec = {"winner": 1, "seq_num_list": [14, 15, 20, 21, 8, 9, 26]}

    class P:
        resets = 0

        def reset_player(self):
            self.resets += 1
    p = P()
    g2 = Game(Board(width=6, height=6, n_in_row=4), sgf_loader=lambda f, h: rec)
    warn, winner, data = g2.start_self_play(p, sgf_home=".", file_name="x.sgf")
    w3, winner3, data3 = osp.sgf_self_play(OBoard(6, 6, 4), P(), rec)
    data = list(data)
    assert (warn, winner) == (w3, winner3) == (0, 1) and p.resets == 1
    for (sa, pa, za), (sb, pb, zb) in zip(data, data3):
        assert np.array_equal(sa, sb) and np.array_equal(pa, pb) and za == zb
    bad = Game(Board(width=6, height=6, n_in_row=4), sgf_loader=lambda f, h: {"winner": 1, "seq_num_list": [3, 3]})
    assert bad.start_self_play(P(), file_name="bad") == (1, None, None)


def test_pure_player_shim():
    from alphapig_b200.game import Board
    from alphapig_b200.mcts_pure import MCTSPlayer
    b = Board(width=8, height=8, n_in_row=5)
    b.init_board()
    # player 1 has four in a row at 10..13 (open at 9 and 14); player 2 stones scattered
    for m in (10, 40, 11, 48, 12, 56, 13, 63):
        b.do_move(m)
    np.random.seed(0)
    player = MCTSPlayer(c_puct=5, n_playout=1000)
    player.set_player_ind(1)
    mv = player.get_action(b)
    assert mv in (9, 14), mv
    # a full board prints the warning and returns None (mcts_pure.py:202-203)
    full = Board(width=5, height=5, n_in_row=5)
    full.init_board()
    full.availables = []
    assert player.get_action(full) is None


def test_policy_value_net_shim_inference_and_device_search():
    from alphapig_b200.game import Board
    from alphapig_b200.mcts_alphaZero import MCTS
    from alphapig_b200.params import init_params
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    W = 8
    arg, aux = init_params("simple", W, W, seed=4, synthetic_stats=True)
    net = PolicyValueNet(W, W, batch_size=16, model_params=(arg, aux))
    b = Board(width=W, height=W, n_in_row=5)
    b.init_board()
    for m in (27, 28, 35, 36, 20):
        b.do_move(m)
    ap, v = net.policy_value_fn(b)
    ap = list(ap)
    assert [a for a, _ in ap] == b.availables and v.shape == (1,)
    st = np.ascontiguousarray(b.current_state(), dtype=np.float32)[None]
    rp, rv = onet.forward(arg, aux, st, "simple")
    assert np.abs(np.log([p for _, p in ap]) - np.log(rp[0][b.availables])).max() <= 1e-3
    assert abs(float(v[0]) - float(rv[0, 0])) <= 1e-3
    probs, vals = net.policy_value(np.repeat(st, 5, axis=0))
    assert probs.shape == (5, W * W) and vals.shape == (5, 1)
    # device-net search == host-callback search fed by the same net (fp32 outputs widen exactly)
    dev = MCTS(net.policy_value_fn, c_puct=5, n_playout=80)
    host = MCTS(lambda board: net.policy_value_fn(board), c_puct=5, n_playout=80)
    assert dev._net is net and host._net is None
    a1, p1 = dev.get_move_probs(b, temp=1.0)
    a2, p2 = host.get_move_probs(b, temp=1.0)
    assert a1 == a2 and np.array_equal(p1, p2)
    # get_policy_param round trip
    garg, gaux = net.get_policy_param()
    assert list(garg.keys()) == list(arg.keys()) and all(np.array_equal(garg[k], arg[k]) for k in arg)
    assert all(np.array_equal(gaux[k], aux[k]) for k in aux)


def test_train_step_gpu_updates_all_replicas(tmp_path):
    import pickle
    from alphapig_b200.policy_value_net_mxnet import PolicyValueNet
    W = 6
    S = W * W
    net = PolicyValueNet(W, W, batch_size=32, n_blocks=2, n_filter=128, seed=0)
    rep = net.search_engine(n_in_row=4, c_puct=5, n_playout=8, n_games=4)
    rs = np.random.RandomState(0)
    x = (rs.rand(32, 9, W, W) > 0.7).astype(np.float32)
    pi = rs.dirichlet(np.ones(S), size=32)
    z = rs.choice([-1.0, 1.0], size=32)
    p0, v0 = net.policy_value(x[:4])
    losses = []
    for _ in range(5):
        loss, ent = net.train_step(x, pi, z, 2e-3)
        assert loss.shape == (1,) and ent.shape == (1,)
        losses.append(float(loss[0]))
    assert losses[-1] < losses[0]
    p1, v1 = net.policy_value(x[:4])
    assert not np.allclose(p0, p1)
    pr, vr = rep.net_forward(x[:4])           # the search replica got the new weights too
    assert np.array_equal(pr, p1) and np.array_equal(vr, v1)
    arg, aux = net.get_policy_param()
    rp, rv = onet.forward(arg, aux, x[:4], "resnet", n_blocks=2)
    assert np.abs(np.log(p1) - np.log(rp)).max() <= 1e-3 and np.abs(v1 - rv).max() <= 1e-3
    f = tmp_path / "m.model"
    net.save_model(str(f))
    a2, x2 = pickle.load(open(str(f), "rb"))
    net2 = PolicyValueNet(W, W, batch_size=32, n_blocks=2, n_filter=128, model_params=(a2, x2))
    p2, v2 = net2.policy_value(x[:4])
    assert np.array_equal(p2, p1) and np.array_equal(v2, v1)


def test_batched_self_play_records():
    from alphapig_b200.params import init_params
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    from alphapig_b200.selfplay import BatchedSelfPlay
    W = 6
    S = W * W
    arg, aux = init_params("simple", W, W, seed=2, synthetic_stats=True)
    net = PolicyValueNet(W, W, batch_size=16, model_params=(arg, aux), n_in_row=4)
    sp = BatchedSelfPlay(net, n_games=48, n_playout=24, c_puct=5, temp=1.0, n_in_row=4, seed=1)
    recs = []
    for _ in range(40):
        recs += sp.step()
        if len(recs) >= 30:
            break
    assert len(recs) >= 30 and sp.finished_games == len(recs)
    for winner, states, pis, z in recs:
        n = len(z)
        assert states.shape == (n, (9 * S + 7) // 8) and pis.shape == (n, S)
        assert np.allclose(pis.sum(1), 1.0) and set(np.unique(z)) <= {-1.0, 0.0, 1.0}
        planes = np.unpackbits(states, axis=1)[:, :9 * S].reshape(n, 9, W, W)
        # replaying: stones only ever accumulate, colour plane alternates, z alternates with the mover
        stones = planes[:, 6].sum((1, 2)) + planes[:, 7].sum((1, 2))
        assert list(stones) == list(range(n))
        assert list(planes[:, 8, 0, 0]) == [1 - (i % 2) for i in range(n)]
        if winner == -1:
            assert not z.any()
        else:
            assert z[-1] == 1.0 and all(z[i] == -z[i + 1] for i in range(n - 1))


def test_pipelined_self_play_groups():
    """Two staggered groups on their own engines / host threads produce well-formed game records and keep
    counting plies per group."""
    from alphapig_b200.params import init_params
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    from alphapig_b200.selfplay import PipelinedSelfPlay
    W = 6
    S = W * W
    arg, aux = init_params("simple", W, W, seed=2, synthetic_stats=True)
    net = PolicyValueNet(W, W, batch_size=16, model_params=(arg, aux), n_in_row=4)
    sp = PipelinedSelfPlay(net, n_games=50, n_groups=2, n_playout=16, c_puct=5, temp=1.0, n_in_row=4, seed=1)
    assert [g.G for g in sp.groups] == [25, 25] and sp.groups[0].eng is not sp.groups[1].eng
    recs, moves = [], 0
    for _ in range(90):
        recs += sp.step()
        moves += sp.last_moves
        if len(recs) >= 20:
            break
    recs += sp.drain()
    assert len(recs) >= 20 and sp.finished_games == len(recs)
    for winner, states, pis, z in recs:
        n = len(z)
        assert states.shape == (n, (9 * S + 7) // 8) and np.allclose(pis.sum(1), 1.0)
        planes = np.unpackbits(states, axis=1)[:, :9 * S].reshape(n, 9, W, W)
        assert list(planes[:, 6].sum((1, 2)) + planes[:, 7].sum((1, 2))) == list(range(n))


def test_selfplay_pick_distribution():
    """ap_selfplay_pick (MCTSPlayer.get_action in self-play, mcts_alphaZero.py:187-215, sampled on the device):
    pi equals softmax(1/temp * log(visits + 1e-10)) scattered by move; with eps = 0 the chosen moves follow pi, with
    eps = 1 they follow the Dirichlet sample whose marginals are Beta(alpha, (A - 1) alpha); full boards give -1."""
    from scipy import stats
    from alphapig_b200.engine import Engine
    from helpers import export_oboard, oboard_from
    W, G = 6, 4096
    eng = Engine(width=W, height=W, n_in_row=4, n_games=G, c_puct=5, n_playout=48)
    b = oboard_from(W, W, 4, [14, 15, 20])
    c, m = export_oboard(b)
    eng.boards_import(np.repeat(c[None], G, 0), np.repeat(m[None], G, 0))
    # identical trees in every game, grown with the uniform evaluator through the dense device path
    for _ in range(48):
        term, depth, path = eng.search_select()
        eng.search_expand_backup_dense(np.full((G, W * W), 1.0 / (W * W), np.float32), np.zeros(G, np.float32))
    count, acts, visits, _, _ = eng.search_root()
    assert (count == count[0]).all() and (visits == visits[0]).all()
    A = int(count[0])
    temp = 1.0
    x = np.log(visits[0, :A].astype(np.float64) + 1e-10) / temp
    pi_ref = np.exp(x - x.max())
    pi_ref /= pi_ref.sum()
    dense = np.zeros(W * W)
    dense[acts[0, :A]] = pi_ref
    # eps = 0: moves ~ pi
    mv, pi = eng.selfplay_pick(temp=temp, eps=0.0, alpha=0.3, seed=5, ply=0)
    assert np.allclose(pi, dense[None], atol=1e-7)
    cnt = np.bincount(mv, minlength=W * W)[acts[0, :A]]
    keep = pi_ref * G >= 5
    f_obs = np.append(cnt[keep], cnt[~keep].sum())
    f_exp = np.append(pi_ref[keep], pi_ref[~keep].sum()) * G
    if f_exp[-1] == 0:
        f_obs, f_exp = f_obs[:-1], f_exp[:-1]
    assert stats.chisquare(f_obs, f_exp)[1] > 1e-6
    # eps = 1: moves ~ Dirichlet sample; the sample's marginals are Beta(alpha, (A-1) alpha), rows sum to 1
    mv, pi, nz = eng.selfplay_pick(temp=temp, eps=1.0, alpha=0.3, seed=6, ply=3, want_noise=True)
    d = nz[:, acts[0, :A]]
    assert np.allclose(d.sum(1), 1.0, atol=1e-9) and d.min() >= 0
    for k in (0, A // 2, A - 1):
        assert stats.kstest(d[:, k], "beta", args=(0.3, (A - 1) * 0.3))[1] > 1e-6, k
    cnt = np.bincount(mv, minlength=W * W)[acts[0, :A]]
    assert stats.chisquare(cnt, np.full(A, G / A))[1] > 1e-6   # E[Dirichlet] is uniform
    # reproducible per (seed, ply); different plies differ
    mv2, _ = eng.selfplay_pick(temp=temp, eps=1.0, alpha=0.3, seed=6, ply=3)
    mv3, _ = eng.selfplay_pick(temp=temp, eps=1.0, alpha=0.3, seed=6, ply=4)
    assert np.array_equal(mv, mv2) and not np.array_equal(mv, mv3)
    eng.close()


def test_batched_selfplay_device_pick_records():
    """BatchedSelfPlay(device_pick=True): same record contract as the host-sampling path (game_ai.py:113-139):
    one (state, pi, z) row per ply of a finished game, z = +1 / -1 by the winner from the mover's side, pi rows sum
    to 1 over legal moves only; the next ply's search runs in the background while records are assembled."""
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    from alphapig_b200.selfplay import BatchedSelfPlay
    W = 6
    net = PolicyValueNet(W, W, batch_size=32, seed=0)
    sp = BatchedSelfPlay(net, n_games=64, n_playout=24, n_in_row=4, seed=3, device_pick=True)
    games = []
    for _ in range(40):
        games.extend(sp.step())
    sp.drain()
    assert len(games) >= 64 and sp.finished_games == len(games)
    for winner, states, pis, z in games:
        n = len(z)
        assert states.shape == (n, (9 * W * W + 7) // 8) and pis.shape == (n, W * W)
        assert np.allclose(pis.sum(1), 1.0, atol=1e-5)
        planes = np.unpackbits(states, axis=1)[:, :9 * W * W].reshape(n, 9, W, W)
        occupied = (planes[:, 6] + planes[:, 7]) > 0           # stones of both sides before the move
        assert (pis.reshape(n, W, W)[:, ::-1][occupied] == 0).all()  # no mass on occupied cells (planes are row-flipped)
        if winner == -1:
            assert (z == 0).all()
        else:
            assert set(np.unique(z)) <= {-1.0, 1.0} and z[-1] == 1.0   # the last mover made the line
            assert (z[::-1][::2] == 1.0).all() and (z[::-1][1::2] == -1.0).all()
