"""GPU: board kernels through the C ABI vs golden vectors written from the reference Board
and vs the oracle on random play (bit-exact)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import export_oboard, oboard_from

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from alphapig_b200.engine import Engine
    return Engine(**kw)


@pytest.mark.parametrize("name", ["boards_8x8", "boards_15x15", "boards_6x6_4", "boards_5x5"])
def test_boards_golden_lockstep(name):
    """All golden games advance in lock-step, one engine game each."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    W, H, n, G = [int(x) for x in z["meta"]]
    eng = _engine(width=W, height=H, n_in_row=n, n_games=G)
    eng.boards_reset(start_player=z["start_player"].astype(np.int32))
    nbytes = (9 * W * H + 7) // 8
    alive = np.ones(G, bool)
    for k in range(W * H):
        ids = np.nonzero(alive & (z["moves"][:, k] >= 0))[0].astype(np.int32)
        if len(ids) == 0:
            break
        feats = eng.boards_features(ids)
        packed = np.packbits(feats.astype(np.uint8).reshape(len(ids), -1), axis=1)
        assert np.array_equal(packed, z["feats"][ids, k, :nbytes]), "current_state differs at ply %d" % k
        legal = eng.boards_legal(ids)
        st = eng.boards_do_move(z["moves"][ids, k].astype(np.int32), ids)
        assert np.all(st == 0)
        assert np.all(legal[np.arange(len(ids)), z["moves"][ids, k]])
        end, win = eng.boards_status(ids)
        assert np.array_equal(end.astype(np.int8), z["ends"][ids, k])
        assert np.array_equal(win.astype(np.int8), z["winners"][ids, k])
        alive[ids[end]] = False
    eng.close()


def test_boards_errors_and_roundtrip():
    from alphapig_b200.engine import Engine
    with pytest.raises(Exception):
        Engine(width=4, height=8, n_in_row=5)
    eng = _engine(width=8, height=8, n_in_row=5, n_games=3)
    with pytest.raises(Exception):
        eng.boards_reset(start_player=[0, 2, 1])
    eng.boards_reset(start_player=[0, 1, 0])
    eng.boards_do_move([3, 3, 3])
    with pytest.raises(ValueError):
        eng.boards_do_move([3], [1])           # occupied
    with pytest.raises(ValueError):
        eng.boards_do_move([64], [0])          # off board
    with pytest.raises(ValueError):
        eng.boards_do_move([-1], [0])
    cells, meta = eng.boards_export()
    assert cells[0, 3] == 1 and cells[1, 3] == 2 and meta[0, 0] == 2 and meta[1, 0] == 1
    assert list(meta[0, 1:7]) == [3, 1, 3, -1, -1, -1]
    # export -> import into other slots -> identical
    eng.boards_import(cells[[1, 0, 2]], meta[[1, 0, 2]], [0, 1, 2])
    c2, m2 = eng.boards_export()
    assert np.array_equal(c2, cells[[1, 0, 2]]) and np.array_equal(m2, meta[[1, 0, 2]])
    legal = eng.boards_legal()
    assert legal.sum() == 3 * 63 and not legal[:, 3].any()
    eng.close()


def test_boards_random_vs_oracle_full_size():
    """4096 concurrent 15x15 games of random legal play, status/legality checked against the oracle
    on a sample and through size-independent properties on all of them."""
    W = H = 15
    G = 4096
    eng = _engine(width=W, height=H, n_in_row=5, n_games=G)
    rs = np.random.RandomState(7)
    alive = np.ones(G, bool)
    hist = [[] for _ in range(G)]
    sample = list(range(0, G, 128))
    while alive.any():
        ids = np.nonzero(alive)[0].astype(np.int32)
        legal = eng.boards_legal(ids)
        cnt = legal.sum(1)
        assert np.all(cnt == W * H - np.array([len(hist[g]) for g in ids]))
        pick = (rs.random_sample(len(ids)) * cnt).astype(np.int64)
        order = np.argsort(~legal, axis=1, kind="stable")  # legal moves first, ascending
        moves = order[np.arange(len(ids)), pick].astype(np.int32)
        eng.boards_do_move(moves, ids)
        for g, m in zip(ids, moves):
            hist[g].append(int(m))
        end, win = eng.boards_status(ids)
        for j, g in enumerate(ids):
            if g in sample:
                ob = oboard_from(W, H, 5, hist[g])
                assert ob.game_end() == (bool(end[j]), int(win[j]))
        alive[ids[end]] = False
    # every finished game: winner's last stone completes a line (oracle check on the sample), and
    # the exported cells equal the replayed history for all games
    cells, meta = eng.boards_export()
    for g in range(0, G, 16):
        ob = oboard_from(W, H, 5, hist[g])
        c, m = export_oboard(ob)
        assert np.array_equal(cells[g], c) and np.array_equal(meta[g, :7], m[:7])
    eng.close()


def test_features_packed_equals_packbits():
    """ap_boards_features_packed == np.packbits(ap_boards_features), incl. a non-square board"""
    from alphapig_b200.engine import Engine
    for (W, H, n) in ((15, 15, 5), (8, 8, 5), (7, 5, 4)):
        G = 33
        eng = Engine(width=W, height=H, n_in_row=n, n_games=G)
        rs = np.random.RandomState(W)
        for ply in range(6):
            legal = eng.boards_legal()
            mv = np.array([rs.choice(np.nonzero(legal[g])[0]) for g in range(G)], np.int32)
            eng.boards_do_move(mv)
            f = eng.boards_features()
            assert np.array_equal(eng.boards_features_packed(), np.packbits(f.reshape(G, -1).astype(np.uint8), axis=1))
        sub = np.array([5, 0, 17], np.int32)
        assert np.array_equal(eng.boards_features_packed(sub), np.packbits(eng.boards_features(sub).reshape(3, -1).astype(np.uint8), axis=1))
        eng.close()
