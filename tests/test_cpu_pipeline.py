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


RUN_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
from alphapig_b200.train_mxnet import TrainPipeline

W = 6
S = W * W


class FakeRingEngine(object):
    """the replay-ring slice of the Engine API (ap_replay_*), host lists instead of HBM"""
    width = height = W
    S = S

    def replay_create(self, maxlen):
        self.maxlen, self.rows = maxlen, []

    def replay_push(self, bits, pis, zs):
        for b, p, z in zip(np.asarray(bits), np.asarray(pis), np.asarray(zs).reshape(-1)):
            self.rows.append((b.copy(), p.copy(), float(z)))

    def replay_size(self):
        total = 8 * len(self.rows)
        return min(total, self.maxlen), total

    def replay_gather(self, idx):
        n = len(idx)
        rec = [self.rows[(8 * len(self.rows) - self.replay_size()[0] + int(j)) // 8] for j in idx]
        st = np.stack([np.unpackbits(r[0])[:9 * S].reshape(9, W, W) for r in rec]).astype(np.float32)
        return st, np.stack([r[1] for r in rec]).astype(np.float32), np.array([r[2] for r in rec], np.float32)


class FakeNet(object):
    """the PolicyValueNet surface TrainPipeline touches; the weights are one flat CPU tensor"""
    board_width = board_height = W
    _device = 0

    def __init__(self):
        self._eng = FakeRingEngine()
        self.flat = torch.zeros(16)
        self.steps = self.synced = 0

    def policy_value_fn(self, board):
        raise AssertionError("the test injects its own self-play data")

    def policy_value(self, states):
        n = len(states)
        return np.full((n, S), 1.0 / S, np.float32), np.zeros((n, 1), np.float32)

    def train_step(self, states, pis, zs, lr):
        self.steps += 1
        self.flat += 1.0
        return np.array([1.0]), np.array([2.0])

    def _views(self):
        return self.flat, {}

    def sync_replicas(self):
        self.synced += 1

    def save_model(self, path):
        pass


conf = dict(board_width=W, board_height=W, n_in_row=4, n_playout=4, batch_size=8, epochs=2, buffer_size=10000,
            sgf_dir="/nonexistent", game_batch_num=3, check_freq=1000, play_batch_size=1)
net = FakeNet()
tp = TrainPipeline(conf, net=net)
made = []


def fake_collect(n_games=1, training_index=None):
    # rank-specific synthetic game: 3 + rank + training_index positions, z = rank marker
    n = 3 + rank + training_index
    rs = np.random.RandomState(100 * rank + training_index)
    states = (rs.rand(n, 9, W, W) > 0.5).astype(np.float64)
    pis = rs.dirichlet(np.ones(S), size=n)
    zs = np.full(n, float(rank) * 2 - 1)
    tp.episode_len = n
    made.append(n)
    tp._store(states, pis, zs)


tp.collect_selfplay_data_ai = fake_collect
tp.run(model_dir=%(tmp)r)
mine = sum(made)
both = torch.tensor([mine], dtype=torch.int64)
dist.all_reduce(both)
if rank == 0:
    # every rank's positions reached the trainer's buffer (8 augmented samples each), in rank order per iteration
    assert len(tp.data_buffer) == 8 * int(both), (len(tp.data_buffer), int(both))
    zs = [r[2] for r in net._eng.rows]
    assert zs[:3] == [-1.0] * 3 and zs[3:7] == [1.0] * 4, zs[:8]       # iteration 0: rank 0's 3 rows, then rank 1's 4
    assert net.steps >= 2                                              # policy_update ran on rank 0 only
else:
    assert len(tp.data_buffer) == 0 and net.steps == 0                 # nothing trains or accumulates off the trainer rank
# the weights every rank ends up with are rank 0's (broadcast after each iteration)
w = [torch.zeros(16) for _ in range(2)]
dist.all_gather(w, net.flat)
assert torch.equal(w[0], w[1]) and float(w[0][0]) >= 2.0 and net.synced == 3
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_train_pipeline_run_two_ranks_gathers_records_to_the_trainer(tmp_path):
    """TrainPipeline.run under torch.distributed (world size 2, gloo, CPU): the self-play records of EVERY rank reach
    rank 0's buffer before policy_update, only rank 0 trains, and every rank holds rank 0's weights afterwards
    (the non-zero ranks' data used to be discarded).  The net and the ring are host fakes; the gather /
    broadcast plumbing is the product's (alphapig_b200/dist.py)."""
    import socket
    import subprocess
    import sys
    ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(RUN_WORKER % {"root": ROOT, "port": port, "tmp": str(tmp_path / "models")})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o
