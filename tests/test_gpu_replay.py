"""GPU: device replay ring + symmetry gather + SGF bootstrap vs the oracle restatement of the reference's
deque / get_equi_data / random.sample / Game.start_self_play(SGF) (oracle/pipeline.py, oracle/selfplay.py;
both pinned live against the reference in tests/test_oracle_vs_reference.py).  Bit-exact."""
import random

import numpy as np
import pytest

from helpers import oboard_from, synth_position
from oracle import pipeline as opl
from oracle import selfplay as osp
from oracle.board import OBoard

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from alphapig_b200.engine import Engine
    return Engine(**kw)


def _game_data(W, seed, n_pos):
    """play_data of one pseudo game: real Board.current_state() planes, random pi / z"""
    rs = np.random.RandomState(seed)
    mv = synth_position(W, W, 5 if W >= 8 else 4, seed, 12 if W >= 8 else 5)
    mv = mv if len(mv) >= 2 else [0, 1]
    out = []
    for k in range(n_pos):
        b = oboard_from(W, W, 5 if W >= 8 else 4, mv[: 1 + (k % len(mv))])
        out.append((np.ascontiguousarray(b.current_state()), rs.dirichlet(np.ones(W * W)).astype(np.float32).astype(np.float64),
                    float(rs.choice([-1.0, 0.0, 1.0]))))
    return out


@pytest.mark.parametrize("W,maxlen", [(6, 150), (15, 1000), (8, 64)])
def test_ring_matches_reference_deque(W, maxlen):
    from alphapig_b200.replay import ReplayBuffer
    eng = _engine(width=W, height=W, n_in_row=5 if W >= 8 else 4, n_games=1)
    ring = ReplayBuffer(eng, maxlen)
    ref = opl.ReplayDeque(maxlen, W, W)
    for g in range(7):
        data = _game_data(W, 10 * W + g, 3 + 2 * g)
        ring.extend(data)
        ref.extend_game(data)
        assert len(ring) == len(ref)
        n = len(ref)
        st, pi, z = eng.replay_gather(np.arange(n))
        for j in range(n):
            s_ref, p_ref, z_ref = ref.buf[j]
            assert np.array_equal(st[j], s_ref.astype(np.float32)), (g, j)
            assert np.array_equal(pi[j], p_ref.astype(np.float32)), (g, j)
            assert z[j] == z_ref
    # same minibatch for the same Python random state
    random.seed(5)
    a = ring.sample(16)
    random.seed(5)
    b = ref.sample(16)
    assert np.array_equal(a[0], np.asarray(b[0], dtype=np.float32))
    assert np.array_equal(a[1], np.asarray(b[1], dtype=np.float32))
    assert np.array_equal(a[2], np.asarray(b[2], dtype=np.float32))
    s0, p0, z0 = ring[0]
    assert np.array_equal(s0, ref.buf[0][0].astype(np.float32)) and z0 == ref.buf[0][2]
    eng.close()


def test_ring_gather_to_device_tensors():
    import torch
    from alphapig_b200.replay import ReplayBuffer
    W = 15
    eng = _engine(width=W, height=W, n_in_row=5, n_games=1)
    ring = ReplayBuffer(eng, 4000)
    for g in range(5):
        ring.extend(_game_data(W, 300 + g, 20))
    random.seed(1)
    host = ring.sample(128)
    random.seed(1)
    dev = ring.sample_torch(128, "cuda:0")
    for h, d in zip(host, dev):
        assert d.is_cuda and np.array_equal(h, d.cpu().numpy())
    eng.close()


def test_sgf_bootstrap_into_ring():
    """Batched SGF replay on device == reference Game.start_self_play per record + get_equi_data + deque."""
    from alphapig_b200.replay import ReplayBuffer
    W = 15
    rs = np.random.RandomState(0)
    seqs, winners = [], []
    for g in range(9):
        mv = synth_position(W, W, 5, 900 + g, 25)
        if len(mv) < 4:
            mv = [112, 113, 97, 98]
        seqs.append([int(m) for m in mv])
        winners.append(int(rs.choice([1, 2])))
    seqs[4] = seqs[4][:3] + [seqs[4][1]] + seqs[4][3:]  # illegal: repeats an occupied cell
    seqs[7] = [3, 400]                                  # illegal: off the board
    eng = _engine(width=W, height=W, n_in_row=5, n_games=1)
    maxlen = 8 * 60
    ring = ReplayBuffer(eng, maxlen)
    ref = opl.ReplayDeque(maxlen, W, W)

    class P(object):
        def reset_player(self):
            pass
    warn = eng.replay_push_sgf(seqs, winners)
    for g, (q, w) in enumerate(zip(seqs, winners)):
        try:
            wr, _, data = osp.sgf_self_play(OBoard(W, W, 5), P(), {"winner": w, "seq_num_list": q})
        except Exception:
            wr, data = 1, None  # an index beyond the board: the reference raises IndexError on probs[move]
        assert bool(warn[g]) == bool(wr), g
        if not wr:
            ref.extend_game(data)
    assert list(warn) == [0, 0, 0, 0, 1, 0, 0, 1, 0]
    assert len(ring) == len(ref) == maxlen
    st, pi, z = eng.replay_gather(np.arange(maxlen))
    for j in range(maxlen):
        s_ref, p_ref, z_ref = ref.buf[j]
        assert np.array_equal(st[j], s_ref.astype(np.float32)), j
        assert np.array_equal(pi[j], p_ref.astype(np.float32)), j
        assert z[j] == z_ref
    eng.close()


@pytest.mark.parametrize("ci", [0, 1])
def test_ring_matches_reference_golden(ci, golden_dir):
    """Device ring vs the deque contents the REFERENCE produced (tests/golden/pipeline_cases.npz)."""
    import os
    from alphapig_b200.replay import ReplayBuffer
    z = np.load(os.path.join(golden_dir, "pipeline_cases.npz"))
    W, maxlen, ng = [int(x) for x in z["c%d_meta" % ci]]
    eng = _engine(width=W, height=W, n_in_row=4 if W < 8 else 5, n_games=1)
    ring = ReplayBuffer(eng, maxlen)
    k = 0
    for n in z["c%d_lens" % ci]:
        n = int(n)
        ring.extend_positions(z["c%d_in_states" % ci][k:k + n], z["c%d_in_pi" % ci][k:k + n], z["c%d_in_z" % ci][k:k + n])
        k += n
    L = z["c%d_dq_z" % ci].shape[0]
    assert len(ring) == L
    st, pi, zz = eng.replay_gather(np.arange(L))
    S = W * W
    want = np.unpackbits(z["c%d_dq_states" % ci], axis=1)[:, :9 * S].reshape(L, 9, W, W).astype(np.float32)
    assert np.array_equal(st, want)
    assert np.array_equal(pi, z["c%d_dq_pi" % ci].astype(np.float32))
    assert np.array_equal(zz, z["c%d_dq_z" % ci].astype(np.float32))
    random.seed(7 + ci)
    assert ring.sample_indices(16) == [int(i) for i in z["c%d_sample_idx" % ci]]
    eng.close()


def test_sgf_push_matches_reference_golden(golden_dir):
    import os
    from alphapig_b200.replay import ReplayBuffer
    z = np.load(os.path.join(golden_dir, "pipeline_cases.npz"))
    eng = _engine(width=15, height=15, n_in_row=5, n_games=1)
    n = len(z["sgf_z"])
    ring = ReplayBuffer(eng, 8 * n)
    warn = eng.replay_push_sgf([[int(m) for m in z["sgf_moves"]]], [int(z["sgf_winner"][0])])
    assert list(warn) == [0] and len(ring) == 8 * n
    # symmetry 7 of every record (rot90^4 + fliplr) then fliplr back is the identity: check the un-augmented
    # records through symmetry index 6 = rot90^4 = identity
    st, pi, zz = eng.replay_gather(np.arange(n) * 8 + 6)
    want = np.unpackbits(z["sgf_states"], axis=1)[:, :9 * 225].reshape(n, 9, 15, 15).astype(np.float32)
    assert np.array_equal(st, want)
    assert np.array_equal(pi, z["sgf_pi"].astype(np.float32)) and np.array_equal(zz, z["sgf_z"].astype(np.float32))
    eng.close()


def test_replay_error_paths():
    """The reference raises on misuse (random.sample larger than the deque: ValueError; deque index: IndexError);
    the ABI reports AP_ERR_BAD_ARG with a message instead of touching memory."""
    from alphapig_b200._lib import EngineError
    from alphapig_b200.replay import ReplayBuffer
    eng = _engine(width=6, height=6, n_in_row=4, n_games=1)
    with pytest.raises(EngineError):
        eng.replay_size()                       # no ring yet
    with pytest.raises(EngineError):
        eng.replay_create(0)
    ring = ReplayBuffer(eng, 40)
    assert len(ring) == 0
    with pytest.raises(ValueError):
        ring.sample_indices(1)                  # random.sample on an empty population
    ring.extend(_game_data(6, 1, 3))
    assert len(ring) == 24
    with pytest.raises(EngineError):
        eng.replay_gather([24])                 # one past the end
    with pytest.raises(EngineError):
        eng.replay_gather([-1])
    with pytest.raises(IndexError):
        ring[24]
    assert ring[-1][2] == ring[23][2]
    ring.extend([])                             # no-op
    assert len(ring) == 24
    ring.extend(_game_data(6, 2, 4))            # 24 + 32 > 40: oldest 16 samples evicted
    assert len(ring) == 40 and eng.replay_size() == (40, 56)
    # a non-square board cannot be augmented by rot90 (train_mxnet.py:122-126 would mis-shape): rejected at create
    e2 = _engine(width=7, height=5, n_in_row=4, n_games=1)
    with pytest.raises(EngineError):
        e2.replay_create(16)
    e2.close()
    eng.close()
