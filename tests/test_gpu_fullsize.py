"""GPU: BASELINE.json's full sizes through size-independent properties (the oracle cannot run 4096 x 400 playouts
in seconds): visit-count conservation, legality and ordering of root actions, probability normalisation,
determinism of the whole search, and - for a sample of 24 games of the 4096 x 400 run - bit-exact agreement with
the oracle MCTS fed the engine's own priors/values (test_c2_full_size_search_matches_oracle_on_sampled_games)."""
import numpy as np
import pytest

import bench
from oracle.board import OBoard

pytestmark = pytest.mark.gpu


def _net_engine(G, n_playout, node_capacity=None):
    from alphapig_b200.engine import Engine
    from alphapig_b200.params import init_params
    arg, aux = init_params("simple", 15, 15, seed=0, synthetic_stats=True)
    merged = dict(arg)
    merged.update(aux)
    eng = Engine(width=15, height=15, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout,
                 node_capacity=n_playout * 225 + 2 if node_capacity is None else node_capacity)
    eng.net_load("simple", merged)
    return eng


def test_c2_full_size_search_properties():
    G, n_playout = 4096, 400
    eng = _net_engine(G, n_playout)
    cells, meta = bench.synthetic_positions(eng, G)
    eng.search_advance(-1)
    eng.search_run(n_playout)
    count, acts, visits, q, rootn = eng.search_root(want_q=True)
    legal = eng.boards_legal()
    S = 225
    assert (rootn == n_playout).all()
    mask = np.arange(S)[None, :] < count[:, None]
    assert (np.where(mask, visits, 0).sum(1) == n_playout - 1).all()      # first playout expands the root only
    assert (count == legal.sum(1)).all()                                  # one child per legal move
    a = np.where(mask, acts, 10 ** 6)
    assert (np.diff(a, axis=1)[mask[:, 1:]] > 0).all()                     # children in ascending move order
    assert legal[np.nonzero(mask)[0], acts[mask]].all()
    assert (np.abs(np.where(mask, q, 0.0)) <= 1.0).all()
    probs = eng.search_root_probs(1.0)
    assert np.allclose(probs.sum(1), 1.0, atol=1e-12) and (probs[~legal] == 0).all()
    # determinism: the same search again builds the same tree
    eng.boards_import(cells, meta)
    eng.search_advance(-1)
    eng.search_run(n_playout)
    c2, a2, v2, q2, r2 = eng.search_root(want_q=True)
    assert np.array_equal(count, c2) and np.array_equal(visits, v2) and np.array_equal(q, q2)
    eng.close()


def _oboard_from_cells(cells, meta):
    """C-ABI board export (cells, meta) -> oracle board"""
    b = OBoard(15, 15, 5)
    b.init_board(0)
    b.states = {int(m): int(cells[m]) for m in np.nonzero(cells)[0]}
    b.availables = [m for m in range(225) if m not in b.states]
    b.current_player, b.last_move = int(meta[0]), int(meta[1])
    hist = [int(h) for h in meta[3:7] if h >= 0]
    # history (most recent first in meta) feeds current_state(): older stones have no order that matters
    older = [m for m in b.states if m not in hist]
    b.history = [(m, b.states[m]) for m in older] + [(m, b.states[m]) for m in reversed(hist)]
    return b


def test_c2_full_size_search_matches_oracle_on_sampled_games():
    """The BENCHMARKED path - 4096 games x 400 playouts on 15x15 through ap_search_run: plain stream launches (no
    CUDA graph above 256 games), leaf compaction by atomic ticket, feature emission inside k_select, split-K FC with
    the soft-max finished inside k_expand_backup - against the oracle MCTS for 24 sampled games: the oracle is fed
    the engine's own fp32 priors / values (second handle, ap_net_forward, batch 1) and must reproduce visit counts,
    Q and root N bit for bit; then ONE more ply with tree reuse on the same 4096-game batch (k_advance at full size)
    and the same comparison."""
    from helpers import assert_root_equals_oracle, engine_net_evaluator
    from oracle.mcts import OMCTS
    G, n_playout = 4096, 400
    eng = _net_engine(G, n_playout, node_capacity=0)  # library default: room for a re-rooted subtree, grows on demand
    probe = _net_engine(1, 1)
    cells, meta = bench.synthetic_positions(eng, G)
    sample = [0, 1, 2, 3, 255, 256, 257, 1000, 1023, 1024, 2047, 2048, 2049, 3000, 3333, 4000, 4093, 4094, 4095,
              77, 513, 1500, 2500, 3500]
    ev = engine_net_evaluator(probe)
    roots = {g: _oboard_from_cells(cells[g], meta[g]) for g in sample}
    # the oracle board rebuilt from the export must present the same net input as the engine's own features
    f = eng.boards_features(sample)
    for k, g in enumerate(sample):
        assert np.array_equal(f[k], np.ascontiguousarray(roots[g].current_state(), dtype=np.float32)), g
    oracles = {g: OMCTS(ev, 5, n_playout) for g in sample}
    eng.search_advance(-1)
    l0 = eng.launch_count()
    eng.search_run(n_playout)
    # the fused lock-step, launched directly: select, conv1+conv2, conv3, conv4, conv5, conv_final+heads, FC,
    # expand/backup = 8 launches (9 with AP_FRONT_FUSED=0), + the pool-capacity check
    assert 8 * n_playout <= eng.launch_count() - l0 <= 9 * n_playout + 2
    count, acts, visits, _, _ = eng.search_root()
    moves = np.zeros(G, np.int32)
    for g in range(G):
        moves[g] = acts[g, int(np.argmax(visits[g, :count[g]]))]
    for g in sample:
        oracles[g].get_move_probs(roots[g], 1.0)
        assert_root_equals_oracle(eng, g, oracles[g], "ply 0")
        vis = [nd.N for nd in oracles[g].root.children.values()]
        assert int(moves[g]) == list(oracles[g].root.children.keys())[int(np.argmax(vis))]
    # second ply with reuse for all 4096 games
    eng.search_advance(moves)
    eng.boards_do_move(moves)
    end, _ = eng.boards_status()
    eng.search_run(n_playout)
    for g in sample:
        oracles[g].update_with_move(int(moves[g]))
        roots[g].do_move(int(moves[g]))
        assert bool(end[g]) == roots[g].game_end()[0]
        if end[g]:
            continue
        oracles[g].get_move_probs(roots[g], 1.0)
        assert_root_equals_oracle(eng, g, oracles[g], "ply 1")
    eng.close()
    probe.close()


def test_c3_full_size_pure_properties():
    from alphapig_b200.engine import Engine
    G, n_playout = 8192, 1000
    eng = Engine(width=15, height=15, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout,
                 node_capacity=n_playout * 225 + 2)
    bench.synthetic_positions(eng, G)
    eng.search_stats()
    mv = eng.pure_run(n_playout, seed=3)
    legal = eng.boards_legal()
    assert legal[np.arange(G), mv].all()
    st = eng.search_stats()
    assert st["playouts"] == G * n_playout
    # every rollout ends within the board: plies per rollout bounded by the empty cells
    assert 0 < st["rollout_plies"] <= G * n_playout * 225
    # hashed-rollout mode is a pure function of the position: two runs agree move for move
    a = eng.pure_run(200, seed=1, rollout_mode=1)
    b = eng.pure_run(200, seed=2, rollout_mode=1)
    assert np.array_equal(a, b)
    eng.close()
