"""CPU: the oracle restatement reproduces the golden vectors that
tests/golden/make_golden.py wrote from the unmodified reference."""
import json
import os
import random

import numpy as np
import pytest

from oracle.board import OBoard
from oracle.evaluators import EVALUATORS, position_key
from oracle.mcts import OMCTSPlayer, OPureMCTS, pure_policy_value_fn
from oracle import selfplay as osp

from conftest import GOLDEN


def _packed_state(b):
    return np.packbits(np.ascontiguousarray(b.current_state()).astype(np.uint8).ravel())


@pytest.mark.parametrize("name", ["boards_8x8", "boards_15x15", "boards_6x6_4", "boards_5x5"])
def test_board_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    W, H, n, G = [int(x) for x in z["meta"]]
    ties = 0
    for g in range(G):
        b = OBoard(W, H, n)
        b.init_board(int(z["start_player"][g]))
        for k in range(W * H):
            m = int(z["moves"][g, k])
            if m < 0:
                break
            assert np.array_equal(_packed_state(b), z["feats"][g, k][: (9 * W * H + 7) // 8])
            b.do_move(m)
            end, winner = b.game_end()
            assert int(end) == int(z["ends"][g, k])
            assert int(winner) == int(z["winners"][g, k])
            ties += int(end and winner == -1)
    if name == "boards_5x5":
        assert ties > 0  # the tie branch (game.py:164-165) is covered


def test_board_errors():
    b = OBoard(4, 8, 5)
    with pytest.raises(Exception):
        b.init_board()
    b = OBoard(8, 8, 5)
    b.init_board()
    b.do_move(3)
    with pytest.raises(ValueError):
        b.do_move(3)


def _run_tree_case(meta, arrs, Board=OBoard, Player=OMCTSPlayer):
    W, H, n = meta["W"], meta["H"], meta["n"]
    b = Board(width=W, height=H, n_in_row=n)
    b.init_board(0)
    for m in meta["start_moves"]:
        b.do_move(m)
    player = Player(EVALUATORS[meta["evaluator"]], c_puct=meta["c_puct"],
                    n_playout=meta["n_playout"], is_selfplay=meta["selfplay"])
    np.random.seed(meta["seed"])
    for ply in range(meta["n_plies"]):
        move, pi = player.get_action(b, temp=meta["temp"], return_prob=1)
        assert np.array_equal(pi, arrs["pi"][ply]), "pi differs at ply %d" % ply
        assert int(move) == int(arrs["chosen"][ply])
        b.do_move(int(move))


def test_tree_golden():
    metas = json.load(open(os.path.join(GOLDEN, "tree_cases.json")))
    z = np.load(os.path.join(GOLDEN, "tree_cases.npz"))
    for i, meta in enumerate(metas):
        arrs = {k: z["c%d_%s" % (i, k)] for k in ("visits", "q", "pi", "root_n", "chosen")}
        _run_tree_case(meta, arrs)


def hash_rollout(state):
    player = state.get_current_player()
    end, winner = state.game_end()
    if not end:
        return int(position_key(state) % 3) - 1
    if winner == -1:
        return 0
    return 1 if winner == player else -1


def test_pure_golden():
    metas = json.load(open(os.path.join(GOLDEN, "pure_cases.json")))
    z = np.load(os.path.join(GOLDEN, "pure_cases.npz"))
    for i, meta in enumerate(metas):
        b = OBoard(meta["W"], meta["H"], meta["n"])
        b.init_board(0)
        for m in meta["start_moves"]:
            b.do_move(m)
        mcts = OPureMCTS(pure_policy_value_fn, meta["c_puct"], meta["n_playout"], rollout_fn=hash_rollout)
        move = mcts.get_move(b)
        assert move == meta["move"]
        assert mcts.root.N == meta["root_n"]
        S = meta["W"] * meta["H"]
        visits = np.zeros(S, np.int32)
        qs = np.zeros(S)
        for a, node in mcts.root.children.items():
            visits[a], qs[a] = node.N, node.Q
        assert np.array_equal(visits, z["p%d_visits" % i])
        assert np.array_equal(qs, z["p%d_q" % i])


def test_selfplay_golden():
    metas = json.load(open(os.path.join(GOLDEN, "selfplay_cases.json")))
    z = np.load(os.path.join(GOLDEN, "selfplay_cases.npz"))
    for i, meta in enumerate(metas):
        b = OBoard(meta["W"], meta["H"], meta["n"])
        player = OMCTSPlayer(EVALUATORS[meta["evaluator"]], c_puct=5, n_playout=meta["n_playout"], is_selfplay=1)
        np.random.seed(meta["seed"])
        random.seed(meta["seed"])
        orig = osp.random.random
        osp.random.random = lambda: 0.5
        try:
            winner, data = osp.start_self_play(b, player, temp=meta["temp"])
        finally:
            osp.random.random = orig
        assert winner == meta["winner"]
        assert [m for m, _ in b.history] == list(z["s%d_moves" % i])
        assert np.array_equal(np.stack([p for _, p, _ in data]), z["s%d_pi" % i])
        assert np.array_equal(np.array([zz for _, _, zz in data]), z["s%d_z" % i])
        st = np.stack([np.packbits(np.ascontiguousarray(s).astype(np.uint8).ravel()) for s, _, _ in data])
        assert np.array_equal(st, z["s%d_states" % i])


# ---- TrainPipeline data path: oracle/pipeline.py vs vectors written from the reference (make_golden_pipeline.py) ----
def _pipeline_case(z, ci):
    W, maxlen, ng = [int(x) for x in z["c%d_meta" % ci]]
    S = W * W
    st = np.unpackbits(z["c%d_in_states" % ci], axis=1)[:, :9 * S].reshape(-1, 9, W, W).astype(np.float64)
    pi, zz = z["c%d_in_pi" % ci], z["c%d_in_z" % ci]
    games, k = [], 0
    for n in z["c%d_lens" % ci]:
        games.append([(st[i], pi[i], float(zz[i])) for i in range(k, k + int(n))])
        k += int(n)
    return W, maxlen, games


@pytest.mark.parametrize("ci", [0, 1])
def test_pipeline_golden(ci):
    from oracle import pipeline as opl
    z = np.load(os.path.join(GOLDEN, "pipeline_cases.npz"))
    W, maxlen, games = _pipeline_case(z, ci)
    rep = opl.ReplayDeque(maxlen, W, W)
    for g in games:
        rep.extend_game(g)
    assert len(rep) == z["c%d_dq_z" % ci].shape[0]
    for j, (s, p, zz) in enumerate(rep.buf):
        assert np.array_equal(np.packbits(np.asarray(s).astype(np.uint8).ravel()), z["c%d_dq_states" % ci][j])
        assert np.array_equal(p, z["c%d_dq_pi" % ci][j]) and zz == z["c%d_dq_z" % ci][j]
    random.seed(7 + ci)
    idx = random.sample(range(len(rep)), 16)
    assert idx == [int(i) for i in z["c%d_sample_idx" % ci]]


def test_sgf_golden():
    z = np.load(os.path.join(GOLDEN, "pipeline_cases.npz"))

    class P(object):
        def reset_player(self):
            pass
    warn, winner, data = osp.sgf_self_play(OBoard(15, 15, 5), P(), {"winner": int(z["sgf_winner"][0]),
                                                                   "seq_num_list": [int(m) for m in z["sgf_moves"]]})
    assert (winner, warn) == tuple(int(x) for x in z["sgf_winner"])
    data = list(data)
    assert len(data) == len(z["sgf_z"])
    for j, (s, p, zz) in enumerate(data):
        assert np.array_equal(_packed_state_arr(s), z["sgf_states"][j])
        assert np.array_equal(p, z["sgf_pi"][j]) and zz == z["sgf_z"][j]


def _packed_state_arr(s):
    return np.packbits(np.ascontiguousarray(s).astype(np.uint8).ravel())


def test_resnet_symbol_matches_reference_graph():
    """Structural pin of oracle/net.py: the op list it derives from its own trunk / head spec equals, node for node
    (op, MXNet name, attrs, input names), the reference's committed train symbol of the 10-block residual net
    (policy_value_loss.json, fixture written by tests/golden/make_graph_golden.py): layer order, kernels / pads /
    filters, which BatchNorms train gamma, the residual adds, Dropout(0.5), tanh / SoftmaxActivation, the loss
    mean((z - v)^2) + mean(-sum(pi * log p)) and the BlockGrad entropy output.  Parameter names and the declared
    input shape too."""
    from oracle import net as onet
    gold = json.load(open(os.path.join(GOLDEN, "res10_symbol_ops.json")))
    mine = onet.symbol_ops("resnet", n_blocks=10, n_filter=128, width=15, height=15)
    assert len(mine) == len(gold["ops"]) == 104
    for a, b in zip(mine, gold["ops"]):
        assert a == b, (a, b)
    assert gold["heads"] == ["makeloss0", "makeloss1"]
    nulls = {n["name"]: n["attrs"] for n in gold["nulls"]}
    arg, aux = onet.param_shapes("resnet", 15, 15, 10, 128)
    assert set(arg) | set(aux) == set(nulls) - {"input_states", "input_labels", "mcts_probs"}
    assert nulls["input_states"]["__shape__"] == "(128, 9, 15, 15)"  # (batch, 9 planes, H, W): train_mxnet.py batch 128
    for name in aux:  # moving statistics start at mean 0 / var 1 (oracle init_params without synthetic stats)
        want = '["zero", {}]' if name.endswith("mean") else '["one", {}]'
        assert nulls[name]["__init__"] == want, name
    for name, shp in arg.items():  # the 3x3 trunk weights are (128, cin, 3, 3), the heads 1x1
        if name.endswith("_weight") and len(shp) == 4:
            k = 1 if name.startswith("conv3_") else 3
            assert shp[2:] == (k, k), name


def test_mxnet_pin():
    """Arithmetic pin of the nets at the MXNet boundary - runs once tests/golden/mxnet_pin.npz exists (written by
    tools/mxnet_pin.py on a machine with MXNet; MXNet is not installable here, so until then the oracle's net arithmetic
    stays "parity unpinned" and this test is skipped, not faked)."""
    import pytest
    path = os.path.join(GOLDEN, "mxnet_pin.npz")
    if not os.path.exists(path):
        pytest.skip("no MXNet-written fixture (python tools/mxnet_pin.py --reference <AlphaPig checkout>)")
    from alphapig_b200.params import init_params
    from oracle import net as onet
    z = np.load(path)
    st = z["states"]
    for tag, arch, nb in (("simple", "simple", 0), ("res3", "resnet", 3)):
        arg, aux = init_params(arch, 15, 15, n_blocks=nb, seed=0, synthetic_stats=True)
        p, v = onet.forward(arg, aux, st, arch, n_blocks=nb)
        assert np.abs(np.log(p) - np.log(z[tag + "_probs"])).max() < 2e-5, tag
        assert np.abs(v - z[tag + "_values"]).max() < 2e-5, tag
