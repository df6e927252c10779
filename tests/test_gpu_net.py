"""GPU: policy/value net forward through the C ABI vs the oracle fp32 restatement
(oracle/net.py; parity unpinned at the MXNet boundary, see its docstring).
Tolerance (north_star): |d log p| <= 1e-3 and |d v| <= 1e-3 absolute, identical weights."""
import numpy as np
import pytest

from helpers import assert_root_equals_oracle, engine_net_evaluator, export_oboard, oboard_from, synth_position
from oracle import net as onet

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _engine(**kw):
    from alphapig_b200.engine import Engine
    return Engine(**kw)


def _states(W, n_states, seed0, max_pairs):
    boards = [oboard_from(W, W, 5, synth_position(W, W, 5, seed0 + g, max_pairs)) for g in range(n_states)]
    return boards, np.stack([np.ascontiguousarray(b.current_state()) for b in boards]).astype(np.float32)


def _merged(arg, aux):
    d = dict(arg)
    d.update(aux)
    return d


@pytest.mark.parametrize("W,arch,nblk", [(15, "simple", 0), (8, "simple", 0), (15, "resnet", 3)])
def test_net_forward_vs_oracle(W, arch, nblk):
    arg, aux = onet.init_params(arch, W, W, seed=0, n_blocks=nblk)
    boards, st = _states(W, 40, 1234, 31 if W == 15 else 10)
    ref_p, ref_v = onet.forward(arg, aux, st, arch, n_blocks=nblk)
    eng = _engine(width=W, height=W, n_in_row=5, n_games=8)
    eng.net_load(arch, _merged(arg, aux), n_blocks=nblk)
    # independent fp32 CUDA-core path: fp32 round-off only
    pp, pv = eng.net_forward_precise(st)
    assert np.abs(np.log(pp) - np.log(ref_p)).max() < 2e-5
    assert np.abs(pv - ref_v).max() < 2e-5
    # tensor-core path (fp16 operands, fp32 accumulate in TMEM)
    p, v = eng.net_forward(st)
    dlp = np.abs(np.log(p) - np.log(ref_p)).max()
    dv = np.abs(v - ref_v).max()
    print("arch %s W %d: max|dlogp| %.3e max|dv| %.3e" % (arch, W, dlp, dv))
    assert np.allclose(p.sum(1), 1.0, atol=1e-5)
    assert dlp <= TOL and dv <= TOL
    eng.close()


@pytest.mark.parametrize("mode", ["0", "1", "2"])
@pytest.mark.parametrize("W,arch,nblk,nst", [(15, "simple", 0, 301), (8, "simple", 0, 131), (15, "resnet", 2, 140),
                                             (6, "simple", 0, 20)])
def test_net_head_modes(monkeypatch, mode, W, arch, nblk, nst):
    """The three head implementations (fp32 CUDA-core FC; split-fp16 tensor-core FC fed by k_head_conv; head convs
    fused into the last trunk epilogue) against the oracle, on batches that are not a multiple of the 128-board
    FC tile and span several tiles."""
    monkeypatch.setenv("AP_HEAD_MODE", mode)
    n_row = 5 if W >= 8 else 4
    arg, aux = onet.init_params(arch, W, W, seed=2, n_blocks=nblk)
    boards = [oboard_from(W, W, n_row, synth_position(W, W, n_row, 500 + g, 31 if W == 15 else 6)) for g in range(nst)]
    st = np.stack([np.ascontiguousarray(b.current_state()) for b in boards]).astype(np.float32)
    ref_p, ref_v = onet.forward(arg, aux, st, arch, n_blocks=nblk)
    eng = _engine(width=W, height=W, n_in_row=n_row, n_games=4)
    eng.net_load(arch, _merged(arg, aux), n_blocks=nblk)
    for rep in range(2):  # second pass: stale operand rows of the first must not leak
        sub = st if rep == 0 else st[: nst // 3]
        p, v = eng.net_forward(sub)
        dlp = np.abs(np.log(p) - np.log(ref_p[: len(sub)])).max()
        dv = np.abs(v - ref_v[: len(sub)]).max()
        print("mode %s arch %s W %d: max|dlogp| %.3e max|dv| %.3e" % (mode, arch, W, dlp, dv))
        assert np.allclose(p.sum(1), 1.0, atol=1e-5)
        assert dlp <= TOL and dv <= TOL
    eng.close()


@pytest.mark.parametrize("nblk,precision,tol", [(10, "split", 5e-5), (10, "split_act", 8e-4), (10, "auto", 8e-4),
                                                (3, "split", 5e-5), (3, "split_act", 8e-4), (10, "fp16", 4e-3)])
def test_resnet_split_precision(nblk, precision, tol):
    """The as-trained residual net (10 blocks x 128, train_mxnet.py:79-91): plain fp16 operands drift to ~1.8e-3,
    the split-precision kernels (hi + lo fp16 pairs, three tensor-core products) stay at fp32 round-off, and
    hi + lo activations x error-diffusion-rounded fp16 weights (two products, the "auto" choice) stay under 8e-4."""
    W = 15
    arg, aux = onet.init_params("resnet", W, W, seed=0, n_blocks=nblk)
    boards, st = _states(W, 150, 1234, 31)
    ref_p, ref_v = onet.forward(arg, aux, st, "resnet", n_blocks=nblk)
    eng = _engine(width=W, height=W, n_in_row=5, n_games=4)
    eng.net_load("resnet", _merged(arg, aux), n_blocks=nblk, precision=precision)
    assert eng.net_precision == {"auto": "split_act"}.get(precision, precision)
    p, v = eng.net_forward(st)
    dlp = np.abs(np.log(p) - np.log(ref_p)).max()
    dv = np.abs(v - ref_v).max()
    print("resnet-%d %s: max|dlogp| %.3e max|dv| %.3e" % (nblk, precision, dlp, dv))
    assert dlp <= tol and dv <= tol
    assert dlp <= TOL or precision == "fp16"
    eng.close()


def test_net_on_leaf_boards_and_policy_value_fn_order():
    """Features emitted on device from bitboards feed the net: same result as the host-state path,
    and probabilities line up with move indices (policy_value_net_mxnet_simple.py:207-226)."""
    W = 15
    G = 24
    arg, aux = onet.init_params("simple", W, W, seed=3)
    eng = _engine(width=W, height=W, n_in_row=5, n_games=G)
    eng.net_load("simple", _merged(arg, aux))
    boards, st = _states(W, G, 77, 31)
    cm = [export_oboard(b) for b in boards]
    eng.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))
    term, depth, _ = eng.search_select()  # fresh trees: every leaf is its root
    assert not term.any() and not depth.any()
    assert np.array_equal(eng.search_leaf_features(), st)
    p_leaf, v_leaf = eng.net_forward_leaves(precise=False)
    p_host, v_host = eng.net_forward(st)
    assert np.array_equal(p_leaf, p_host) and np.array_equal(v_leaf, v_host[:, 0])
    pp, pv = eng.net_forward_leaves(precise=True)
    ref_p, ref_v = onet.forward(arg, aux, st, "simple")
    assert np.abs(np.log(pp) - np.log(ref_p)).max() < 2e-5
    assert np.abs(np.log(p_leaf) - np.log(ref_p)).max() <= TOL
    assert np.abs(v_leaf - ref_v[:, 0]).max() <= TOL
    eng.close()


def test_net_weight_refresh_and_layout():
    W = 8
    arg, aux = onet.init_params("simple", W, W, seed=1)
    eng = _engine(width=W, height=W, n_in_row=5, n_games=2)
    eng.net_load("simple", _merged(arg, aux))
    lay = eng.net_layout()
    assert [n for n, _, _ in lay] == list(arg.keys()) + list(aux.keys())
    assert all(sz == int(np.prod(_merged(arg, aux)[n].shape)) for n, _, sz in lay)
    ptr, numel = eng.net_weights()
    assert ptr and numel >= sum(sz for _, _, sz in lay)
    eng.close()


def test_search_run_device_net_matches_host_evaluator_path():
    """ap_search_run (select -> features -> net -> expand/backup on device) must build exactly the
    tree that the host path builds when fed the same fp32 priors/values."""
    W = 8
    G, n_playout = 6, 60
    arg, aux = onet.init_params("simple", W, W, seed=5)
    roots = [oboard_from(W, W, 5, synth_position(W, W, 5, 40 + g, 8)) for g in range(G)]
    cm = [export_oboard(b) for b in roots]
    a = _engine(width=W, height=W, n_in_row=5, n_games=G, n_playout=n_playout)
    b = _engine(width=W, height=W, n_in_row=5, n_games=G, n_playout=n_playout)
    for e in (a, b):
        e.net_load("simple", _merged(arg, aux))
        e.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))
    a.search_run(n_playout)
    for _ in range(n_playout):
        b.search_select(want_path=False)
        p, v = b.net_forward_leaves(precise=False)
        b.search_expand_backup_dense(p, v)
    ca, aa, va, qa, ra = a.search_root(want_q=True)
    cb, ab, vb, qb, rb = b.search_root(want_q=True)
    assert np.array_equal(ca, cb) and np.array_equal(aa, ab) and np.array_equal(va, vb)
    assert np.array_equal(qa, qb) and np.array_equal(ra, rb)
    assert ra.min() == n_playout
    a.close()
    b.close()


def test_search_run_compacts_terminal_leaves():
    """Near-won positions: the search reaches terminal leaves within a few playouts.  The device path evaluates only
    the non-terminal leaves (compacted batch + slot map); the tree must equal the host-evaluator path's, which
    evaluates every leaf and discards the terminal answers as the reference does (mcts_alphaZero.py:124-136)."""
    W = 8
    G, n_playout = 12, 80
    arg, aux = onet.init_params("simple", W, W, seed=7)
    roots = []
    for g in range(G):
        # black has an open four on row g % 3 + 2 (white scattered): black to move wins at once, many lines end fast
        r = g % 3 + 2
        mv = []
        whites = [0, 7, 56, 63][: 4]
        for i in range(4):
            mv += [r * W + 2 + i, whites[i]]
        roots.append(oboard_from(W, W, 5, mv))
    cm = [export_oboard(b) for b in roots]
    a = _engine(width=W, height=W, n_in_row=5, n_games=G, n_playout=n_playout)
    b = _engine(width=W, height=W, n_in_row=5, n_games=G, n_playout=n_playout)
    for e in (a, b):
        e.net_load("simple", _merged(arg, aux))
        e.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))
    a.search_stats()
    a.search_run(n_playout)
    st = a.search_stats()
    assert st["playouts"] == G * n_playout and st["terminal_leaves"] > G  # terminal leaves did occur
    for _ in range(n_playout):
        b.search_select(want_path=False)
        p, v = b.net_forward_leaves(precise=False)
        b.search_expand_backup_dense(p, v)
    ca, aa, va, qa, ra = a.search_root(want_q=True)
    cb, ab, vb, qb, rb = b.search_root(want_q=True)
    assert np.array_equal(ca, cb) and np.array_equal(aa, ab) and np.array_equal(va, vb)
    assert np.array_equal(qa, qb) and np.array_equal(ra, rb)
    a.close()
    b.close()


@pytest.mark.parametrize("W,nblk,nst", [(15, 2, 70), (15, 10, 40), (8, 3, 131)])
def test_inception_variant_vs_oracle(W, nblk, nst):
    """Builder-defined Inception-ResNet variant (configs[3]; no reference semantics): the device graph -
    tower stems merged into one 1x1 layer, slice-reading towers as zero-block 3x3 layers writing channel slices in
    place, 1x1 up-projection with the 0.17 residual scale folded in - against the repo's own fp32 restatement."""
    arg, aux = onet.init_params("inception", W, W, seed=4, n_blocks=nblk)
    boards, st = _states(W, nst, 4321, 31 if W == 15 else 10)
    ref_p, ref_v = onet.forward(arg, aux, st, "inception", n_blocks=nblk)
    eng = _engine(width=W, height=W, n_in_row=5, n_games=4)
    eng.net_load("inception", _merged(arg, aux), n_blocks=nblk)
    p, v = eng.net_forward(st)
    dlp = np.abs(np.log(p) - np.log(ref_p)).max()
    dv = np.abs(v - ref_v).max()
    print("inception-%d W %d: max|dlogp| %.3e max|dv| %.3e" % (nblk, W, dlp, dv))
    assert np.allclose(p.sum(1), 1.0, atol=1e-5)
    assert dlp <= TOL and dv <= TOL
    # the search runs on it (uncompacted batch: no fused-head instantiation for the 1x1 up-projection)
    cm = [export_oboard(b) for b in boards[:4]]
    eng.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))
    eng.search_run(20)
    _, _, visits, _, rootn = eng.search_root()
    assert (rootn == 20).all() and (visits.sum(1) == 19).all()
    eng.close()


def test_inception_shim_train_step():
    from alphapig_b200.policy_value_net_inception import PolicyValueNet
    W = 8
    net = PolicyValueNet(W, W, batch_size=8, n_blocks=2, seed=0)
    rs = np.random.RandomState(0)
    st = (rs.rand(8, 9, W, W) < 0.3).astype(np.float32)
    pi = rs.dirichlet(np.ones(W * W), size=8)
    z = rs.choice([-1.0, 1.0], size=8)
    p0, _ = net.policy_value(st)
    loss, ent = net.train_step(st, pi, z, 2e-3)
    p1, _ = net.policy_value(st)
    assert np.isfinite(loss).all() and not np.array_equal(p0, p1)
    arg, aux = net.get_policy_param()
    rp, rv = onet.forward(arg, aux, st, "inception", n_blocks=2)
    assert np.abs(np.log(p1) - np.log(rp)).max() <= TOL


def _peaky_simple_net(W, seed, scale):
    """the reference net with its policy FC weights scaled: soft-max outputs become very peaky and position
    dependent - what a trained net looks like to the tree (Xavier weights give near-uniform priors)"""
    arg, aux = onet.init_params("simple", W, W, seed=seed)
    arg = dict(arg)
    arg["fc_3_1_1_weight"] = arg["fc_3_1_1_weight"] * np.float32(scale)
    return arg, aux


def test_device_search_deep_trees_peaky_net_vs_oracle():
    """ap_search_run (all-device loop) at 15x15, n_playout 400, 6 plies with tree reuse on a peaky net: mean select
    depth >= 5, root N grows past 1000, the retained subtrees outgrow the default node capacity (pools grow between
    plies).  The oracle MCTS, fed the engine's own fp32 priors / values through a second handle, must rebuild every
    root bit for bit (visits, Q, root N) on every ply."""
    from oracle.mcts import OMCTS
    W = 15
    G, n_playout, plies = 4, 400, 6
    arg, aux = _peaky_simple_net(W, 0, 70.0)
    a = _engine(width=W, height=W, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout)
    b = _engine(width=W, height=W, n_in_row=5, n_games=1, c_puct=5, n_playout=1, node_capacity=8)
    for e in (a, b):
        e.net_load("simple", _merged(arg, aux))
    cap0 = a.node_capacity()
    roots = [oboard_from(W, W, 5, synth_position(W, W, 5, 1234 + g)) for g in range(G)]
    cm = [export_oboard(r) for r in roots]
    a.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))
    ev = engine_net_evaluator(b)
    oracles = [OMCTS(ev, 5, n_playout) for _ in range(G)]
    depth_sum = playouts = 0
    max_root_n = 0
    for ply in range(plies):
        a.search_stats()
        a.search_run(n_playout)
        st = a.search_stats()
        depth_sum += st["path_nodes"] - st["playouts"]
        playouts += st["playouts"]
        moves = np.zeros(G, np.int32)
        for g in range(G):
            o_acts, _ = oracles[g].get_move_probs(roots[g], 1.0)
            assert_root_equals_oracle(a, g, oracles[g], "ply %d" % ply)
            max_root_n = max(max_root_n, oracles[g].root.N)
            vis = [nd.N for nd in oracles[g].root.children.values()]
            moves[g] = o_acts[int(np.argmax(vis))]
        a.search_advance(moves)
        a.boards_do_move(moves)
        for g in range(G):
            oracles[g].update_with_move(int(moves[g]))
            roots[g].do_move(int(moves[g]))
        if any(r.game_end()[0] for r in roots):
            break
    print("peaky net: mean depth %.2f, max root N %d, capacity %d -> %d" % (depth_sum / playouts, max_root_n, cap0,
                                                                        a.node_capacity()))
    assert depth_sum / playouts >= 5.0
    assert max_root_n > 2 * n_playout
    assert a.node_capacity() > cap0
    a.close()
    b.close()


def test_single_game_15x15_graph_path_vs_oracle():
    """ONE 15x15 game, n_playout 400 (what human_play_mxnet.py / evaluate/ChessClient.py run): batches of <= 256 games
    replay the lock-steps as a CUDA graph.  Two plies with tree reuse, then a reset, against the oracle MCTS fed the
    engine's own priors / values; the second search replays the graph captured by the first."""
    from oracle.mcts import OMCTS
    W = 15
    n_playout = 400
    arg, aux = onet.init_params("simple", W, W, seed=9)
    a = _engine(width=W, height=W, n_in_row=5, n_games=1, c_puct=5, n_playout=n_playout)
    b = _engine(width=W, height=W, n_in_row=5, n_games=1, c_puct=5, n_playout=1, node_capacity=8)
    for e in (a, b):
        e.net_load("simple", _merged(arg, aux))
    root = oboard_from(W, W, 5, synth_position(W, W, 5, 4242))
    c, m = export_oboard(root)
    a.boards_import(c[None], m[None])
    o = OMCTS(engine_net_evaluator(b), 5, n_playout)
    launches = []
    for ply, reuse in enumerate((True, True, False)):
        l0 = a.launch_count()
        a.search_run(n_playout)
        launches.append(a.launch_count() - l0)
        o_acts, _ = o.get_move_probs(root, 1.0)
        assert_root_equals_oracle(a, 0, o, "ply %d" % ply)
        vis = [nd.N for nd in o.root.children.values()]
        mv = int(o_acts[int(np.argmax(vis))])
        a.search_advance([mv if reuse else -1])
        o.update_with_move(mv if reuse else -1)
        a.boards_do_move([mv])
        root.do_move(mv)
    assert launches[0] == launches[1] == launches[2] and launches[0] >= 8 * n_playout
    a.close()
    b.close()


def test_virtual_loss_mode_k1_is_parity_mode_and_k8_keeps_the_playout_budget():
    """ap_search_run_vl (opt-in multi-leaf search, virtual loss): k = 1 must build ap_search_run's tree bit for bit
    (two plies, tree reuse); k = 8 gives every game exactly n_playout playouts (root N), root children summing to
    n_playout - 1 on a fresh tree, legal ascending children - in ~1/6 of the lock-steps."""
    W = 15
    G, n_playout = 3, 200
    arg, aux = onet.init_params("simple", W, W, seed=5)
    roots = [oboard_from(W, W, 5, synth_position(W, W, 5, 600 + g)) for g in range(G)]
    cm = [export_oboard(b) for b in roots]
    engs = [_engine(width=W, height=W, n_in_row=5, n_games=G, n_playout=n_playout) for _ in range(3)]
    for e in engs:
        e.net_load("simple", _merged(arg, aux))
        e.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))
    a, b, c = engs
    for ply in range(2):
        a.search_run(n_playout)
        b.search_run_vl(n_playout, 1)
        ra, rb = a.search_root(want_q=True), b.search_root(want_q=True)
        assert all(np.array_equal(x, y) for x, y in zip(ra, rb)), "k = 1, ply %d" % ply
        count, acts, visits = ra[0], ra[1], ra[2]
        mv = np.array([acts[g, int(np.argmax(visits[g, :count[g]]))] for g in range(G)], np.int32)
        for e in (a, b):
            e.search_advance(mv)
            e.boards_do_move(mv)
    # k = 8 on fresh trees
    c.search_stats()
    l0 = c.launch_count()
    c.search_run_vl(n_playout, 8)
    launches = c.launch_count() - l0
    st = c.search_stats()
    count, acts, visits, q, rootn = c.search_root(want_q=True)
    legal = c.boards_legal()
    assert st["playouts"] == G * n_playout and (rootn == n_playout).all()
    for g in range(G):
        n = int(count[g])
        assert n == int(legal[g].sum()) and list(acts[g, :n]) == list(np.nonzero(legal[g])[0])
        assert int(visits[g, :n].sum()) == n_playout - 1
        assert (np.abs(q[g, :n]) <= 1.0).all()
    assert launches <= 9 * (1 + (n_playout - 1 + 7) // 8) + 4  # 26 lock-steps (of 8 - 9 launches) instead of 200
    # reuse + a second multi-leaf search: root N accumulates as in the sequential search
    mv = np.array([acts[g, int(np.argmax(visits[g, :count[g]]))] for g in range(G)], np.int32)
    kept = np.array([visits[g, int(np.argmax(visits[g, :count[g]]))] for g in range(G)])
    c.search_advance(mv)
    c.boards_do_move(mv)
    c.search_run_vl(n_playout, 8)
    _, _, _, _, rootn2 = c.search_root()
    assert (rootn2 == kept + n_playout).all()
    for e in engs:
        e.close()


def test_virtual_loss_shim_and_limits():
    from alphapig_b200._lib import EngineError
    from alphapig_b200.game import Board
    from alphapig_b200.mcts_alphaZero import MCTSPlayer
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    W = 8
    net = PolicyValueNet(W, W, batch_size=16, seed=0)
    player = MCTSPlayer(net.policy_value_fn, c_puct=5, n_playout=100, is_selfplay=1, leaves_per_step=8)
    b = Board(width=W, height=W, n_in_row=5)
    b.init_board()
    np.random.seed(1)
    for _ in range(4):
        mv, pi = player.get_action(b, temp=1.0, return_prob=1)
        assert mv in b.availables and abs(pi.sum() - 1.0) < 1e-9 and (pi[list(b.states)] == 0).all()
        b.do_move(mv)
    eng = net.search_engine(n_games=200, n_playout=10, tag="vl-limit")
    with pytest.raises(EngineError):
        eng.search_run_vl(10, 4)  # 200 games x 4 leaves do not fit the 256-board batch of a small engine
    net.close()


@pytest.mark.parametrize("W,nst", [(15, 611), (8, 131), (15, 3)])
def test_fused_front_kernel_is_bit_identical_to_separate_layers(monkeypatch, W, nst):
    """front_tc.cu (conv1 + conv2 of the 6-conv net in one kernel, conv1's output handed to conv2 through shared
    memory) against the two separate conv launches (AP_FRONT_FUSED=0): identical probabilities and values, bit for bit,
    on batches of several tiles per CTA, one tile per CTA and fewer tiles than CTAs - and both within 1e-3 of the
    oracle."""
    arg, aux = onet.init_params("simple", W, W, seed=11)
    boards, st = _states(W, nst, 9000, 31 if W == 15 else 10)
    outs = []
    for fused in ("0", "1"):
        monkeypatch.setenv("AP_FRONT_FUSED", fused)
        eng = _engine(width=W, height=W, n_in_row=5, n_games=4)
        eng.net_load("simple", _merged(arg, aux))
        outs.append(eng.net_forward(st))
        # second pass on the same handle: the kernel's barriers / TMEM are set up from scratch every launch
        p2, v2 = eng.net_forward(st[::-1].copy())
        assert np.array_equal(p2[::-1], outs[-1][0]) and np.array_equal(v2[::-1], outs[-1][1])
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    ref_p, ref_v = onet.forward(arg, aux, st, "simple")
    assert np.abs(np.log(outs[1][0]) - np.log(ref_p)).max() <= TOL and np.abs(outs[1][1] - ref_v).max() <= TOL
