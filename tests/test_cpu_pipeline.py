"""CPU: host logic of the TrainPipeline shim that needs no device - the KL / lr-multiplier rule against the
reference-pinned oracle, the SGF reader, the checkpoint formats, get_equi_data."""
import os
import pickle
import types

import numpy as np
import pytest

from oracle import pipeline as opl


def test_kl_rule_matches_oracle():
    from alphapig_b200.train_mxnet import kl_and_lr_rule
    rs = np.random.RandomState(0)
    for trial in range(50):
        old = rs.dirichlet(np.ones(36), size=8)
        new = old if trial % 5 == 0 else rs.dirichlet(np.ones(36) * (1 + trial), size=8) * 0.2 + old * 0.8
        mult = [0.04, 1.0, 25.0, 3.0][trial % 4]

        class N(object):
            calls = 0

            def policy_value(self, sb):
                N.calls += 1
                return (old if N.calls == 1 else new), np.zeros((8, 1))

            def train_step(self, *a):
                return np.zeros(1), np.zeros(1)
        rep = types.SimpleNamespace(sample=lambda b: ([0] * 8, [0] * 8, list(rs.choice([-1.0, 1.0], 8))))
        _, _, want, kl_want, _, _, _ = opl.policy_update(N(), rep, 8, 1e-3, mult, 1, 0.02)
        kl, got = kl_and_lr_rule(old, new, 0.02, mult)
        assert got == want and kl == kl_want


def test_get_equi_data_matches_oracle():
    from alphapig_b200.train_mxnet import TrainPipeline
    rs = np.random.RandomState(1)
    data = [((rs.rand(9, 8, 8) < 0.3).astype(np.float64), rs.dirichlet(np.ones(64)), 1.0) for _ in range(3)]
    a = TrainPipeline.get_equi_data(types.SimpleNamespace(board_height=8, board_width=8), data)
    b = opl.equi_data(data, 8, 8)
    assert len(a) == len(b) == 24
    for (sa, pa, za), (sb, pb, zb) in zip(a, b):
        assert np.array_equal(sa, sb) and np.array_equal(pa, pb) and za == zb


def test_checkpoint_formats(tmp_path):
    from alphapig_b200 import checkpoint
    from alphapig_b200.params import init_params
    arg, aux = init_params("simple", 8, 8, seed=0)
    p = str(tmp_path / "m.model")
    with open(p, "wb") as f:
        pickle.dump((dict(arg), dict(aux)), f, protocol=2)  # what save_model writes (policy_value_net_mxnet.py:305-309)
    a2, x2 = checkpoint.load_model(p)
    assert set(a2) == set(arg) and all(np.array_equal(a2[k], arg[k]) for k in arg)
    checkpoint.save_npz(str(tmp_path / "m.npz"), (arg, aux))
    a3, x3 = checkpoint.load_npz(str(tmp_path / "m.npz"))
    assert list(a3) == list(arg) and list(x3) == list(aux)
    assert all(np.array_equal(x3[k], aux[k]) for k in aux)

    class FakeND(object):  # an MXNet NDArray stand-in: only .asnumpy() is used
        def __init__(self, a):
            self.a = a

        def asnumpy(self):
            return self.a
    checkpoint.save_npz(str(tmp_path / "n.npz"), ({k: FakeND(v) for k, v in arg.items()}, aux))
    a4, _ = checkpoint.load_npz(str(tmp_path / "n.npz"))
    assert all(np.array_equal(a4[k], arg[k]) for k in arg)


def test_sgf_reader(tmp_path):
    from alphapig_b200.utils import sgf_dataIter
    text = "(;GM[4]FF[4]SZ[15]\nB[hh];W[ii];B[hi];W[gg])\n\n\n"
    with open(os.path.join(str(tmp_path), "12_blank_a_.sgf"), "w") as f:
        f.write(text)
    r = sgf_dataIter.get_data_from_files("12_blank_a_.sgf", str(tmp_path))
    assert r["seq_list"] == ["hh", "ii", "hi", "gg"] and r["seq_num_list"] == [112, 128, 113, 96] and r["winner"] == 1
    with pytest.raises(ValueError):
        sgf_dataIter.winner_from_name("12_draw__a_.sgf")
