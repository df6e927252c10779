"""GPU: MCTS kernels (select / expand / backup / re-root) through the C ABI.
Visit counts, Q, pi and chosen moves must be bit-exact vs the golden vectors written from the
reference MCTS and vs the oracle, under injected deterministic evaluators, Dirichlet noise off."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import assert_root_equals_oracle, export_oboard, host_playouts, oboard_from, synth_position
from oracle.evaluators import EVALUATORS, position_key
from oracle.mcts import OMCTS, OPureMCTS, pure_policy_value_fn, softmax

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from alphapig_b200.engine import Engine
    return Engine(**kw)


def _root_arrays(eng, g=0):
    count, acts, visits, q, rootn = eng.search_root(want_q=True)
    S = eng.S
    v = np.zeros(S, np.int32)
    qq = np.zeros(S, np.float64)
    a = acts[g, :count[g]]
    v[a] = visits[g, :count[g]]
    qq[a] = q[g, :count[g]]
    return a, visits[g, :count[g]], v, qq, int(rootn[g])


def test_tree_golden_cases():
    metas = json.load(open(os.path.join(GOLDEN, "tree_cases.json")))
    z = np.load(os.path.join(GOLDEN, "tree_cases.npz"))
    for i, meta in enumerate(metas):
        W, H, n = meta["W"], meta["H"], meta["n"]
        eng = _engine(width=W, height=H, n_in_row=n, n_games=1, c_puct=meta["c_puct"], n_playout=meta["n_playout"])
        root = oboard_from(W, H, n, meta["start_moves"])
        c, m = export_oboard(root)
        eng.boards_import(c[None], m[None])
        for ply in range(meta["n_plies"]):
            host_playouts(eng, [root], EVALUATORS[meta["evaluator"]], meta["n_playout"])
            acts, vis_list, visits, q, rootn = _root_arrays(eng)
            assert np.array_equal(visits, z["c%d_visits" % i][ply]), "case %d ply %d visits" % (i, ply)
            assert np.array_equal(q, z["c%d_q" % i][ply]), "case %d ply %d Q" % (i, ply)
            assert rootn == z["c%d_root_n" % i][ply]
            # acts are the legal moves ascending = child insertion order
            assert list(acts) == list(root.availables)
            pi = np.zeros(W * H)
            pi[acts] = softmax(1.0 / meta["temp"] * np.log(np.array(vis_list) + 1e-10))
            assert np.array_equal(pi, z["c%d_pi" % i][ply])
            dev_pi = eng.search_root_probs(meta["temp"])[0]
            assert np.allclose(dev_pi, pi, rtol=0, atol=1e-12)
            move = int(z["c%d_chosen" % i][ply])
            eng.search_advance([move if meta["selfplay"] else -1])
            eng.boards_do_move([move])
            root.do_move(move)
        eng.close()


def test_tree_lockstep_many_games_vs_oracle():
    """32 different 8x8 positions searched concurrently, 3 plies with tree reuse, vs the oracle MCTS."""
    W = H = 8
    G, n_playout, plies = 32, 150, 3
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout)
    roots, oracles = [], []
    for g in range(G):
        roots.append(oboard_from(W, H, 5, synth_position(W, H, 5, 900 + g, 10)))
        oracles.append(OMCTS(EVALUATORS["e2" if g % 2 else "e3"], 5, n_playout))
    cm = [export_oboard(b) for b in roots]
    eng.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))
    for ply in range(plies):
        # drive the engine with a per-game evaluator
        for _ in range(n_playout):
            term, depth, path = eng.search_select()
            counts = np.zeros(G, np.int32)
            acts = np.zeros((G, W * H), np.int16)
            pri = np.zeros((G, W * H))
            vals = np.zeros(G)
            for g in range(G):
                leaf = roots[g].clone()
                for m in path[g, :depth[g]]:
                    leaf.do_move(int(m))
                ap, v = EVALUATORS["e2" if g % 2 else "e3"](leaf)
                ap = list(ap)
                counts[g] = len(ap)
                acts[g, :len(ap)] = [a for a, _ in ap]
                pri[g, :len(ap)] = [p for _, p in ap]
                vals[g] = v
            eng.search_expand_backup(counts, acts, pri, vals)
        count, acts, visits, q, rootn = eng.search_root(want_q=True)
        moves = np.zeros(G, np.int32)
        for g in range(G):
            o_acts, o_probs = oracles[g].get_move_probs(roots[g], 1.0)
            o_vis = [nd.N for nd in oracles[g].root.children.values()]
            o_q = [float(nd.Q) for nd in oracles[g].root.children.values()]
            assert list(acts[g, :count[g]]) == list(o_acts)
            assert list(visits[g, :count[g]]) == o_vis, "game %d ply %d" % (g, ply)
            assert list(q[g, :count[g]]) == o_q
            assert rootn[g] == oracles[g].root.N
            # most-visited child (first max) as the move; every 4th game resets its tree instead of reusing it
            moves[g] = o_acts[int(np.argmax(o_vis))]
        adv = np.where(np.arange(G) % 4 == 3, -1, moves).astype(np.int32)
        eng.search_advance(adv)
        eng.boards_do_move(moves)
        for g in range(G):
            oracles[g].update_with_move(int(adv[g]))
            roots[g].do_move(int(moves[g]))
    eng.close()


def test_pure_bookkeeping_golden():
    """mcts_pure tree bookkeeping (uniform priors, injected rollout result) vs the golden vectors
    written from the reference mcts_pure.MCTS."""
    metas = json.load(open(os.path.join(GOLDEN, "pure_cases.json")))
    z = np.load(os.path.join(GOLDEN, "pure_cases.npz"))

    def policy(leaf):
        ap, _ = pure_policy_value_fn(leaf)
        # value the reference would back up: _evaluate_rollout with the random play replaced by a hash
        return ap, float(int(position_key(leaf) % 3) - 1)

    for i, meta in enumerate(metas):
        W, H, n = meta["W"], meta["H"], meta["n"]
        eng = _engine(width=W, height=H, n_in_row=n, n_games=1, c_puct=meta["c_puct"], n_playout=meta["n_playout"])
        root = oboard_from(W, H, n, meta["start_moves"])
        c, m = export_oboard(root)
        eng.boards_import(c[None], m[None])
        host_playouts(eng, [root], policy, meta["n_playout"])
        acts, vis_list, visits, q, rootn = _root_arrays(eng)
        assert np.array_equal(visits, z["p%d_visits" % i])
        assert np.array_equal(q, z["p%d_q" % i])
        assert rootn == meta["root_n"]
        assert int(acts[int(np.argmax(vis_list))]) == meta["move"]
        eng.close()


def test_pure_fused_kernel_vs_oracle():
    """The fused one-CTA-per-game mcts_pure kernel with the deterministic rollout hash must give the
    oracle's visit counts and move exactly."""
    from alphapig_b200.engine import rollout_hash_host
    W = H = 8
    G, n_playout = 16, 600
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout)
    roots = [oboard_from(W, H, 5, synth_position(W, H, 5, 500 + g, 10)) for g in range(G)]
    cm = [export_oboard(b) for b in roots]
    eng.boards_import(np.stack([c for c, _ in cm]), np.stack([m for _, m in cm]))

    def rollout(state):
        player = state.get_current_player()
        end, winner = state.game_end()
        if not end:
            c, _ = export_oboard(state)
            return rollout_hash_host(c, state.current_player, W, H)
        if winner == -1:
            return 0
        return 1 if winner == player else -1

    # the device hash agrees with its host twin on the roots
    dev = eng.rollout_hash()
    assert list(dev) == [rollout(b) for b in roots]
    moves = eng.pure_run(n_playout, seed=1, rollout_mode=1)
    count, acts, visits, q, rootn = eng.search_root(want_q=True)
    for g in range(G):
        o = OPureMCTS(pure_policy_value_fn, 5, n_playout, rollout_fn=rollout)
        mv = o.get_move(roots[g])
        assert mv == moves[g]
        assert list(visits[g, :count[g]]) == [nd.N for nd in o.root.children.values()]
        assert list(q[g, :count[g]]) == [float(nd.Q) for nd in o.root.children.values()]
    eng.close()


def test_rollout_distribution():
    """Random rollouts (uniform legal moves) from the empty 15x15 board: game length and result
    distribution vs the oracle's rollout (mcts_pure.py:138-157) on >= 10^4 rollouts."""
    from oracle.board import OBoard
    W = H = 15
    G = 16384
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, n_playout=1, node_capacity=4)
    v, plies = eng.rollout_eval(seed=123)
    assert plies.min() >= 9 and plies.max() <= 225
    # oracle sample (slow): 200 rollouts
    rs = np.random.RandomState(5)
    o_plies, o_v = [], []
    for _ in range(200):
        b = OBoard(W, H, 5)
        b.init_board(0)
        while not b.game_end()[0]:
            b.do_move(int(b.availables[rs.randint(len(b.availables))]))
        o_plies.append(len(b.states))
        w = b.game_end()[1]
        o_v.append(0 if w == -1 else (1 if w == 1 else -1))
    # means agree within 4 standard errors of the small oracle sample
    se = np.std(o_plies) / np.sqrt(len(o_plies))
    assert abs(plies.mean() - np.mean(o_plies)) < 4 * se + 0.5
    se_v = np.std(o_v) / np.sqrt(len(o_v))
    assert abs(v.mean() - np.mean(o_v)) < 4 * se_v + 0.02
    # first player (to move at the root) wins slightly more often than the second; ties are rare
    assert (v == 0).mean() < 0.02
    # different seeds give different streams, same seed is reproducible
    v2, p2 = eng.rollout_eval(seed=123)
    assert np.array_equal(plies, p2) and np.array_equal(v, v2)
    v3, p3 = eng.rollout_eval(seed=124)
    assert not np.array_equal(plies, p3)
    eng.close()


def _two_sample_chi2_p(a, b, bins):
    """p-value of the chi-square homogeneity test of two samples over the given bin edges"""
    from scipy.stats import chi2_contingency
    ha, _ = np.histogram(a, bins)
    hb, _ = np.histogram(b, bins)
    keep = (ha + hb) >= 20
    tab = np.stack([np.append(ha[keep], ha[~keep].sum()), np.append(hb[keep], hb[~keep].sum())])
    tab = tab[:, tab.sum(0) > 0]
    return chi2_contingency(tab)[1]


def test_rollout_permutation_exact_with_injected_draws():
    """ap_rollout_eval_keys: with the per-cell draws supplied by the host (NumPy MT19937) the device permutation
    rollout must return exactly the (value, plies) of the oracle playing the empty cells in ascending
    (draw, cell) order move by move (mcts_pure.py:138-157) - 15x15 empty / mid-game / late positions, 8x8."""
    from oracle.rollout import rollout_by_play
    rs = np.random.RandomState(21)
    for W, H, n, G in ((15, 15, 5, 384), (8, 8, 5, 64), (11, 9, 4, 64)):
        eng = _engine(width=W, height=H, n_in_row=n, n_games=G, n_playout=1, node_capacity=4)
        boards = []
        for g in range(G):
            while True:  # random legal play of random length that leaves the game open
                b = oboard_from(W, H, n, [])
                for _ in range(rs.randint(0, W * H - 1) if g % 3 else rs.randint(0, 8)):
                    b.do_move(int(b.availables[rs.randint(len(b.availables))]))
                    if b.game_end()[0]:
                        break
                if not b.game_end()[0]:
                    break
            boards.append(b)
        ex = [export_oboard(b) for b in boards]
        eng.boards_import(np.stack([c for c, _ in ex]), np.stack([m for _, m in ex]))
        keys = rs.randint(0, 1 << 24, size=(G, 256)).astype(np.uint32)
        keys[::5] &= 0xFF  # every fifth game: 8-bit draws, many ties -> the cell index decides
        v, p = eng.rollout_eval_keys(keys)
        for g, b in enumerate(boards):
            # draw of cell (h, w) = keys[g, h*16 + w]; equal draws are ordered by the device's slot index
            slot = lambda w: ((w & 3) << 2) | (w >> 2)
            order = [m for _, _, m in sorted((int(min(keys[g, (m // W) * 16 + m % W], 0xFFFFFE)), (m // W) * 16 + slot(m % W), m)
                                             for m in b.availables)]
            assert (int(v[g]), int(p[g])) == rollout_by_play(b, order), (W, H, n, g)
        eng.close()


def _moments_agree(dev_v, dev_p, ref_v, ref_p, what, k=4.5):
    """mean value, mean plies and the spread of plies of two samples agree within k standard errors"""
    dv, dp, rv, rp = (np.asarray(x, np.float64) for x in (dev_v, dev_p, ref_v, ref_p))
    se = lambda a, b_: np.sqrt(a.var() / len(a) + b_.var() / len(b_))
    assert abs(dv.mean() - rv.mean()) < k * se(dv, rv), (what, "value", dv.mean(), rv.mean())
    assert abs(dp.mean() - rp.mean()) < k * se(dp, rp), (what, "plies", dp.mean(), rp.mean())
    # standard error of a standard deviation ~ sd * sqrt((kurt - 1) / 4n); kurtosis of these lengths is < 4
    se_sd = np.sqrt(dp.var() * 0.75 / len(dp) + rp.var() * 0.75 / len(rp))
    assert abs(dp.std() - rp.std()) < k * se_sd, (what, "sd(plies)", dp.std(), rp.std())


def test_rollout_distribution_pinned_to_numpy_reference_sample():
    """ap_rollout_eval2, both implementations (0 = permutation + bit descent, 2 = ply by ply), against
    oracle/rollout.py:rollout_sample_numpy (mcts_pure.py:138-157 with NumPy's own generator): ~1M device
    rollouts vs 300k reference rollouts from the empty board and from two SURVEY 8(d) mid-game positions.
    Mean value, mean length, spread of the length (4.5 standard errors) and the length histogram."""
    from oracle.rollout import rollout_sample_numpy
    W = H = 15
    G = 16384
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, n_playout=1, node_capacity=4)
    for name, moves in (("empty", []), ("mid1240", synth_position(W, H, 5, 1240)), ("mid1251", synth_position(W, H, 5, 1251))):
        b = oboard_from(W, H, 5, moves)
        c, m = export_oboard(b)
        eng.boards_import(np.repeat(c[None], G, 0), np.repeat(m[None], G, 0))
        ref_v, ref_p = rollout_sample_numpy(b, 300000, np.random.RandomState(77))
        for impl, n_seeds in ((0, 64), (2, 16)):
            out = [eng.rollout_eval(seed=1000 * impl + sd, impl=impl) for sd in range(n_seeds)]
            v = np.concatenate([o[0] for o in out])
            p = np.concatenate([o[1] for o in out])
            what = (name, impl)
            assert p.min() >= 1 and p.max() <= len(b.availables)
            _moments_agree(v, p, ref_v, ref_p, what)
            assert _two_sample_chi2_p(p, ref_p, np.arange(0, 232, 3)) > 1e-5, what
            for val in (-1, 1):
                assert _two_sample_chi2_p(p[v == val], ref_p[ref_v == val], np.arange(0, 232, 3)) > 1e-5, (what, val)
            # parity law of the alternating game: the side to move at the leaf wins on odd ply counts only
            assert np.all(p[v == 1] % 2 == 1) and np.all(p[v == -1] % 2 == 0)
    eng.close()


def test_rollout_permutation_exact_small_position():
    """A position with 6 empty cells: the exact distribution of (value, plies) over all 720 move orders
    (oracle/rollout.py on the oracle board) vs 16384 device permutation rollouts, and the edge cases
    E = 1 and an already finished game."""
    import itertools
    from collections import Counter
    from scipy.stats import chisquare
    from oracle.rollout import rollout_by_play
    W = H = 6
    n = 4
    rs = np.random.RandomState(3)
    while True:  # random legal play down to 6 empty cells without finishing the game
        b = oboard_from(W, H, n, [])
        ok = True
        for _ in range(W * H - 6):
            b.do_move(int(b.availables[rs.randint(len(b.availables))]))
            if b.game_end()[0]:
                ok = False
                break
        if ok and {1, -1} <= {rollout_by_play(b, o)[0] for o in itertools.islice(itertools.permutations(b.availables), 200)}:
            break
    exact = Counter(rollout_by_play(b, o) for o in itertools.permutations(b.availables))
    G = 16384
    eng = _engine(width=W, height=H, n_in_row=n, n_games=G, n_playout=1, node_capacity=4)
    c, m = export_oboard(b)
    eng.boards_import(np.repeat(c[None], G, 0), np.repeat(m[None], G, 0))
    v, p = eng.rollout_eval(seed=9, impl=0)
    got = Counter(zip(v.tolist(), p.tolist()))
    assert set(got) <= set(exact), (got, exact)
    keys = sorted(exact)
    f_exp = np.array([exact[k] for k in keys], float) / 720.0 * G
    f_obs = np.array([got.get(k, 0) for k in keys], float)
    big = f_exp >= 5
    if (~big).any():
        f_exp = np.append(f_exp[big], f_exp[~big].sum())
        f_obs = np.append(f_obs[big], f_obs[~big].sum())
    assert chisquare(f_obs, f_exp)[1] > 1e-6, (keys, f_obs, f_exp)
    # E = 1: one forced ply; finished game: 0 plies and the reference's terminal value
    while len(b.availables) > 1 and not b.game_end()[0]:
        b.do_move(b.availables[0])
    c, m = export_oboard(b)
    eng.boards_import(np.repeat(c[None], G, 0), np.repeat(m[None], G, 0))
    v, p = eng.rollout_eval(seed=1, impl=0)
    want = rollout_by_play(b, list(b.availables))
    assert set(zip(v.tolist(), p.tolist())) == {want}
    eng.close()


def test_pool_exhaustion_is_reported():
    from alphapig_b200._lib import EngineError
    eng = _engine(width=8, height=8, n_in_row=5, n_games=1, node_capacity=100)
    root = oboard_from(8, 8, 5, [])
    with pytest.raises(EngineError) as ei:
        host_playouts(eng, [root], EVALUATORS["e1"], 5)
    assert ei.value.code == -3
    eng.close()


def test_rollout_eval_matches_host_model_exactly():
    """ap_rollout_eval with the device's OWN random numbers: oracle/rollout.py:device_perm_rollout restates the
    Philox4x32-10 draws, the slot mapping and the (draw, slot) ordering of csrc/rollout.cu and plays that order move
    by move on the oracle board (mcts_pure.py:138-157) - value and game length of every game must be identical."""
    import bench
    from oracle.rollout import device_perm_rollout
    from oracle.board import OBoard
    W = H = 15
    G = 256
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, n_playout=1, node_capacity=4)
    cells, meta = bench.synthetic_positions(eng, G)
    for seed in (19, (7 << 32) + 3):
        v, p = eng.rollout_eval(seed=seed, impl=0)
        for g in range(G):
            b = OBoard(W, H, 5)
            b.init_board(0)
            b.states = {int(m): int(cells[g, m]) for m in np.nonzero(cells[g])[0]}
            b.availables = [m for m in range(W * H) if m not in b.states]
            b.current_player, b.last_move = int(meta[g, 0]), int(meta[g, 1])
            assert (int(v[g]), int(p[g])) == device_perm_rollout(b, seed, g), (seed, g)
    eng.close()


def test_pure_run_random_rollouts_match_host_model_exactly():
    """ap_pure_run in rollout_mode 0 (the real mcts_pure configuration): with the device's rollouts predicted
    exactly on the host (oracle/rollout.py:device_perm_rollout, draws keyed by (seed, game, playout)), the oracle's
    OPureMCTS must reproduce the device's root visit counts, Q and chosen moves bit for bit."""
    from oracle.rollout import device_perm_rollout
    W = H = 15
    G, n_playout, seed = 6, 300, 11
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout)
    roots = [oboard_from(W, H, 5, synth_position(W, H, 5, 1234 + i)) for i in range(G)]
    ex = [export_oboard(b) for b in roots]
    eng.boards_import(np.stack([c for c, _ in ex]), np.stack([m for _, m in ex]))
    moves = eng.pure_run(n_playout, seed=seed, rollout_mode=0)
    count, acts, visits, q, rootn = eng.search_root(want_q=True)
    for g in range(G):
        it = [0]

        def rollout(state, g=g, it=it):
            v, _ = device_perm_rollout(state, seed, g, playout=it[0])
            it[0] += 1
            return v

        o = OPureMCTS(pure_policy_value_fn, 5, n_playout, rollout_fn=rollout)
        assert o.get_move(roots[g]) == moves[g]
        assert list(visits[g, :count[g]]) == [nd.N for nd in o.root.children.values()]
        assert list(q[g, :count[g]]) == [float(nd.Q) for nd in o.root.children.values()]
    eng.close()


def test_deep_trees_peaky_evaluator_with_reuse_and_pool_growth():
    """15x15, n_playout 400, 6 plies with tree reuse under a very peaky injected evaluator (oracle/evaluators.py:e4_peaky):
    the search runs ~7-10 plies deep, the re-rooted subtree keeps most of its visits (root N grows past 1000) and the
    retained nodes outgrow the library-chosen capacity, so the pools must GROW mid-game (the reference's trees are
    unbounded) - visits, Q and root N stay bit-exact vs the oracle throughout.  Exercises k_advance's BFS compaction on
    deep subtrees and the multi-level backup."""
    W = H = 15
    G, n_playout, plies = 4, 400, 6
    ev = EVALUATORS["e4"]
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout)  # capacity: library default
    cap0 = eng.node_capacity()
    roots = [oboard_from(W, H, 5, synth_position(W, H, 5, 1234 + g)) for g in range(G)]
    oracles = [OMCTS(ev, 5, n_playout) for _ in range(G)]
    ex = [export_oboard(b) for b in roots]
    eng.boards_import(np.stack([c for c, _ in ex]), np.stack([m for _, m in ex]))
    depth_sum = playouts = 0
    max_root_n = 0
    for ply in range(plies):
        eng.search_stats()
        host_playouts(eng, roots, ev, n_playout)
        st = eng.search_stats()
        depth_sum += st["path_nodes"] - st["playouts"]
        playouts += st["playouts"]
        moves = np.zeros(G, np.int32)
        for g in range(G):
            o_acts, _ = oracles[g].get_move_probs(roots[g], 1.0)
            assert_root_equals_oracle(eng, g, oracles[g], "ply %d" % ply)
            max_root_n = max(max_root_n, oracles[g].root.N)
            vis = [nd.N for nd in oracles[g].root.children.values()]
            moves[g] = o_acts[int(np.argmax(vis))]
        eng.search_advance(moves)
        eng.boards_do_move(moves)
        for g in range(G):
            oracles[g].update_with_move(int(moves[g]))
            roots[g].do_move(int(moves[g]))
            assert not roots[g].game_end()[0]
    assert depth_sum / playouts >= 5.0, depth_sum / playouts
    assert max_root_n > 2 * n_playout
    assert eng.node_capacity() > cap0, "the retained subtrees were expected to outgrow the default capacity"
    eng.close()


def test_active_mask_skips_games():
    """ap_search_set_active: skipped games keep their trees untouched and cost no playouts; the others are unaffected."""
    W = H = 8
    G, n_playout = 8, 40
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout)
    roots = [oboard_from(W, H, 5, synth_position(W, H, 5, 300 + g, 8)) for g in range(G)]
    ex = [export_oboard(b) for b in roots]
    eng.boards_import(np.stack([c for c, _ in ex]), np.stack([m for _, m in ex]))
    active = np.arange(G) % 3 != 1
    eng.search_set_active(active)
    eng.search_stats()
    host_playouts(eng, roots, EVALUATORS["e2"], n_playout, active=active)
    assert eng.search_stats()["playouts"] == int(active.sum()) * n_playout
    _, _, _, _, rootn = eng.search_root()
    assert list(rootn) == [n_playout if a else 0 for a in active]
    for g in np.nonzero(active)[0]:
        o = OMCTS(EVALUATORS["e2"], 5, n_playout)
        o.get_move_probs(roots[g], 1.0)
        assert_root_equals_oracle(eng, int(g), o)
    mv = eng.pure_run(50, seed=2, rollout_mode=1)
    assert (mv[~active] == -1).all() and (mv[active] >= 0).all()
    eng.search_set_active(None)
    mv = eng.pure_run(50, seed=2, rollout_mode=1)
    assert (mv >= 0).all()
    eng.close()
