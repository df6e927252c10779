"""CPU (-m "not gpu"): C-ABI surface, host-side logic, training step, multi-process plumbing (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/alphapig_b200.h declares."""
    from alphapig_b200 import _lib, build
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "alphapig_b200.h")).read()
    declared = set(re.findall(r"\b(ap_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ap_engine", "ap_config", "ap_status", "ap_tensor"}
    assert declared == set(_lib.SIGNATURES.keys()), declared ^ set(_lib.SIGNATURES.keys())
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.ap_version()


def test_no_gpu_fails_loudly():
    """No CPU fallback: creating an engine without a CUDA device raises."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alphapig_b200._lib import EngineError
    from alphapig_b200.engine import Engine
    with pytest.raises(EngineError):
        Engine(width=8, height=8, n_in_row=5, n_games=1)
    # bad geometry is the reference's plain Exception (game.py:36-38), checked before touching the device
    from alphapig_b200.game import Board
    b = Board(width=4, height=4, n_in_row=5)
    with pytest.raises(Exception):
        b.init_board()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "alphapig_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_board_shim_host_mirror():
    from alphapig_b200.game import Board, export_board_state
    b = Board(width=8, height=8, n_in_row=5)
    b.init_board(1)
    assert b.current_player == 2 and b.availables == list(range(64)) and b.last_move == -1
    b.do_move(10)
    b.do_move(11)
    assert b.states == {10: 2, 11: 1} and b.history == [(10, 2), (11, 1)] and 10 not in b.availables
    with pytest.raises(ValueError):
        b.do_move(10)
    assert b.move_to_location(11) == [1, 3] and b.location_to_move([1, 3]) == 11
    assert b.location_to_move([9, 9]) == -1 and b.location_to_move([1]) == -1
    import copy
    c = copy.deepcopy(b)
    c.do_move(12)
    assert 12 in b.availables and 12 not in c.availables
    cells, meta = export_board_state(c)
    # note: the failed do_move(10) above already overwrote states[10] (same as the reference, game.py:118)
    assert cells[12] == c.states[12] and list(meta[:3]) == [c.current_player, 12, len(c.states)]


def test_params_and_flops_match_survey():
    from alphapig_b200.params import flop_per_leaf, init_params, param_shapes
    assert flop_per_leaf("simple", 15, 15) == 517682700
    assert flop_per_leaf("simple", 8, 8) == 147169536
    assert flop_per_leaf("resnet", 15, 15, n_blocks=10) == 1332521100
    arg, aux = param_shapes("simple", 15, 15)
    assert arg["conv1_weight"] == (64, 9, 3, 3) and arg["fc_3_1_1_weight"] == (225, 900)
    assert arg["fc_3_2_1_weight"] == (1, 450) and aux["conv_final_var"] == (256,)
    a, x = init_params("resnet", 15, 15, n_blocks=2, seed=1)
    assert "convB2_weight" in a and "bnA1_moving_var" in x and a["bnA1_gamma"].dtype == np.float32
    # product parameter table == oracle parameter table (names, order, shapes)
    from oracle import net as onet
    oa, ox = onet.param_shapes("resnet", 15, 15, n_blocks=2)
    pa, px = param_shapes("resnet", 15, 15, n_blocks=2)
    assert list(oa.items()) == list(pa.items()) and list(ox.items()) == list(px.items())


def test_visit_softmax_matches_reference_formula():
    from alphapig_b200.selfplay import visit_softmax
    from oracle.mcts import softmax
    rs = np.random.RandomState(0)
    visits = rs.randint(0, 50, size=(5, 64)).astype(np.int32)
    counts = np.array([64, 10, 1, 33, 64], np.int32)
    for temp in (1.0, 1e-3, 0.5):
        p = visit_softmax(visits, counts, temp)
        for g in range(5):
            ref = softmax(1.0 / temp * np.log(visits[g, :counts[g]] + 1e-10))
            assert np.allclose(p[g, :counts[g]], ref, rtol=1e-12, atol=0)
            assert np.all(p[g, counts[g]:] == 0)


def test_train_step_cpu_fp32_vs_fp64_and_learns():
    """The PyTorch training step (device-agnostic) decreases the loss and its fp32 run tracks fp64."""
    from alphapig_b200 import train as T
    from alphapig_b200.params import init_params
    W = 6
    S = W * W
    rs = np.random.RandomState(0)
    arg0, aux0 = init_params("simple", W, W, seed=0)
    x = (rs.rand(16, 9, W, W) > 0.7).astype(np.float32)
    pi = rs.dirichlet(np.ones(S), size=16).astype(np.float32)
    z = rs.choice([-1.0, 1.0], size=16).astype(np.float32)
    losses = {}
    for dt in (torch.float32, torch.float64):
        arg = {k: torch.tensor(v, dtype=dt) for k, v in arg0.items()}
        aux = {k: torch.tensor(v, dtype=dt) for k, v in aux0.items()}
        opt = T.AdamState()
        ls = []
        for _ in range(6):
            l, ent = T.train_step(arg, aux, opt, torch.tensor(x, dtype=dt), torch.tensor(pi, dtype=dt),
                                  torch.tensor(z, dtype=dt), 2e-3, "simple", dropout=False)
            ls.append(float(l))
        losses[dt] = ls
        assert ls[-1] < ls[0]
        assert float(ent) > 0
        # fix_gamma BNs keep gamma == 1; moving stats moved away from their init
        assert torch.all(arg["conv1_gamma"] == 1)
        assert not torch.allclose(aux["conv1_var"], torch.ones_like(aux["conv1_var"]))
    assert np.allclose(losses[torch.float32], losses[torch.float64], rtol=2e-3)


def test_train_step_matches_autograd_adam_reference():
    """One step against an independent formulation: torch.optim.Adam on grad/B + wd*w (MXNet semantics)."""
    from alphapig_b200 import train as T
    from alphapig_b200.params import init_params
    W = 5
    S = W * W
    rs = np.random.RandomState(1)
    arg0, aux0 = init_params("resnet", W, W, n_blocks=1, seed=2)
    x = torch.tensor((rs.rand(8, 9, W, W) > 0.6).astype(np.float64))
    pi = torch.tensor(rs.dirichlet(np.ones(S), size=8))
    z = torch.tensor(rs.choice([-1.0, 1.0], size=8))
    arg = {k: torch.tensor(v, dtype=torch.float64) for k, v in arg0.items()}
    aux = {k: torch.tensor(v, dtype=torch.float64) for k, v in aux0.items()}
    # independent: autograd + torch Adam (same bias-corrected update as MXNet's)
    P = {k: v.clone().requires_grad_(True) for k, v in arg.items()}
    P.update({k: v.clone() for k, v in aux.items()})
    probs, value, _ = T.forward_train(P, x, "resnet", 1, dropout=False)
    loss = ((z.reshape(-1, 1) - value) ** 2).mean() + (-(torch.log(probs) * pi).sum(1)).mean()
    names = [k for k in arg if not (k.endswith("_gamma") and not k.startswith("bn"))]
    grads = torch.autograd.grad(loss, [P[k] for k in names])
    expect = {}
    for k, g in zip(names, grads):
        g = g / 8 + (1e-4 * arg[k] if k.endswith(("_weight", "_gamma")) else 0)
        m = 0.1 * g
        v = 0.001 * g * g
        lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
        expect[k] = arg[k] - lr_t * m / (v.sqrt() + 1e-8)
    l, _ = T.train_step(arg, aux, T.AdamState(), x, pi, z, 1e-3, "resnet", n_blocks=1, dropout=False)
    assert abs(float(l) - float(loss)) < 1e-12
    for k in names:
        assert torch.allclose(arg[k], expect[k], rtol=1e-9, atol=1e-12), k


def test_replay_pack_roundtrip_and_sharding():
    from alphapig_b200 import dist as D
    S = 225
    rs = np.random.RandomState(0)
    n = 7
    st = (rs.rand(n, 9, 15, 15) > 0.5)
    bits = np.packbits(st.reshape(n, -1).astype(np.uint8), axis=1)
    pis = rs.dirichlet(np.ones(S), size=n)
    zs = rs.choice([-1.0, 0.0, 1.0], size=n)
    packed = D.pack_records(bits, pis, zs, S)
    assert packed.shape == (n, D.record_width(S)) and D.record_width(S) == 256 + 900 + 4
    s2, p2, z2 = D.unpack_records(packed, S, 15, 15)
    assert np.array_equal(s2, st.astype(np.float32)) and np.array_equal(p2, pis.astype(np.float32))
    assert np.array_equal(z2, zs.astype(np.float32))
    # contiguous game shards cover everything exactly once
    for total, world in ((4096, 8), (10, 3), (5, 8)):
        spans = [D.shard_games(total, r, world) for r in range(world)]
        assert sum(c for _, c in spans) == total
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from alphapig_b200 import dist as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
# (1) replay gather with a different number of records per rank
S = 36
n = 3 + 2 * rank
rs = np.random.RandomState(rank)
bits = np.packbits((rs.rand(n, 9 * S) > 0.5).astype(np.uint8), axis=1)
packed = D.pack_records(bits, rs.dirichlet(np.ones(S), size=n), np.full(n, float(rank)), S)
allp = D.gather_replay(packed, device=torch.device("cpu"))
assert allp.shape == (8, D.record_width(S))
_, _, z = D.unpack_records(allp, S, 6, 6)
assert list(z) == [0.0] * 3 + [1.0] * 5
assert np.array_equal(allp[3 * rank: 3 * rank + n] if rank == 0 else allp[3:], packed)
# (1b) the device-tensor form (CPU tensors under gloo): to every rank, and to the trainer rank only
t = torch.from_numpy(packed)
parts = D.gather_records_device(t)
assert [p.shape[0] for p in parts] == [3, 5] and torch.equal(parts[rank], t)
assert np.array_equal(torch.cat(parts).numpy(), allp)
parts0 = D.gather_records_device(t, dst=0)
if rank == 0:
    assert np.array_equal(torch.cat(parts0).numpy(), allp)
else:
    assert parts0 == []
cg = dist.new_group(backend="gloo")  # the host-side rendezvous group of the loop
parts1 = D.gather_records_device(t, dst=1, cpu_group=cg)
assert (np.array_equal(torch.cat(parts1).numpy(), allp) if rank == 1 else parts1 == [])
bits2, pis2, z2 = D.split_records(allp, S)
st_u, pi_u, z_u = D.unpack_records(allp, S, 6, 6)
assert np.array_equal(np.unpackbits(bits2, axis=1)[:, :9 * S].reshape(-1, 9, 6, 6).astype(np.float32), st_u)
assert np.array_equal(pis2, pi_u) and np.array_equal(z2, z_u)
# (2) weight broadcast through the same code path the GPU build uses (flat buffer + sync hook)
class FakeNet:
    def __init__(self):
        self.flat = torch.full((1000,), float(rank + 1))
        self.synced = 0
    def _views(self):
        return self.flat, {}
    def sync_replicas(self):
        self.synced += 1
net = FakeNet()
D.broadcast_weights(net, src=0)
assert net.synced == 1 and torch.all(net.flat == 1.0)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_plumbing(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_bench_synthetic_positions_shape():
    sys.path.insert(0, ROOT)
    import bench
    rs = np.random.RandomState(3)
    for _ in range(20):
        cells, meta = bench.draw_position(rs)
        k = int(meta[2])
        assert k % 2 == 0 and (cells == 1).sum() == k // 2 == (cells == 2).sum()
        assert meta[0] == 1 and (meta[1] == -1) == (k == 0)
        if k >= 4:
            assert all(cells[m] != 0 for m in meta[3:7])
            assert cells[meta[3]] == 2  # last stone was white's


def test_permutation_rollout_restatement_equals_play():
    """oracle/rollout.py: the bit-descent over a random permutation of the empty cells (what the device
    rollout computes) gives the same (value, plies) as playing that permutation move by move
    (mcts_pure.py:138-157), on empty, mid-game, nearly full and small boards."""
    from helpers import oboard_from, synth_position
    from oracle.rollout import rollout_by_descent, rollout_by_play
    rs = np.random.RandomState(11)
    cases = [(15, 15, 5, []), (8, 8, 5, []), (6, 6, 4, []), (15, 15, 5, synth_position(15, 15, 5, 1240)),
             (15, 15, 5, synth_position(15, 15, 5, 1251)), (8, 8, 5, synth_position(8, 8, 5, 7, max_pairs=10)),
             (5, 5, 5, []), (6, 5, 3, [0, 7])]
    ties = 0
    for W, H, n, moves in cases:
        b = oboard_from(W, H, n, moves)
        for _ in range(40):
            order = rs.permutation(b.availables)
            got = rollout_by_descent(b, order)
            want = rollout_by_play(b, order)
            assert got == want, (W, H, n, moves, list(order))
            ties += want[0] == 0
    assert ties > 0  # the 5x5x5 board mostly ties: the t >= E branch is exercised


def test_vectorised_rollout_sampler_equals_play():
    """oracle/rollout.py:rollout_sample_numpy (the statistical pin of the device rollouts) against playing the
    same move orders one by one on the oracle board (mcts_pure.py:138-157)."""
    from helpers import oboard_from, synth_position
    from oracle.rollout import rollout_by_play, rollout_sample_numpy

    class Replay(object):  # RandomState stand-in that records the uniforms it hands out
        def __init__(self, seed):
            self.rs = np.random.RandomState(seed)
            self.keys = None

        def random_sample(self, shape):
            self.keys = self.rs.random_sample(shape)
            return self.keys

    for moves in ([], synth_position(15, 15, 5, 1240), synth_position(15, 15, 5, 1251)):
        b = oboard_from(15, 15, 5, moves)
        rp = Replay(5)
        v, p = rollout_sample_numpy(b, 60, rp)
        empties = np.array(b.availables)
        for i in range(60):
            order = empties[np.argsort(rp.keys[i])]
            assert (int(v[i]), int(p[i])) == rollout_by_play(b, order)


def test_philox_known_answers():
    """The Philox4x32-10 round structure the device rollouts and the self-play move sampling use (restated in
    oracle/rollout.py with the same constants and key schedule as csrc/rollout.cu) reproduces the Random123
    known-answer vectors."""
    from oracle.rollout import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert philox4x32_10(ctr, key) == want


def test_device_rollout_model_reproduces_logged_device_run():
    """The exact host model of the device rollouts (oracle/rollout.py: Philox draws + slot order) against a LOGGED
    B200 run: profiles/r1_pure_ab.txt records `ap_rollout_eval(seed=19)` over the 8192 SURVEY 8(d) bench positions as
    "mean plies 79.4, mean value 0.0439"; the model must print the same (a value mean of 0.0439 over 8192 games pins
    the sum of the values to exactly 360).  The per-game equality is the GPU test
    test_rollout_eval_matches_host_model_exactly."""
    import re
    import bench
    from oracle.board import OBoard
    from oracle.rollout import outcomes_from_ranks_numpy, philox4x32_10
    line = [ln for ln in open(os.path.join(ROOT, "profiles", "r1_pure_ab.txt")) if ln.startswith("rollout_eval impl 0")][0]
    want_plies, want_value = re.search(r"mean plies ([\d.]+), mean value ([-\d.]+)", line).groups()
    W = H = 15
    G, seed = 8192, 19
    full = np.zeros((G, W * H), np.int16)
    n_empty = np.zeros(G, np.int64)
    for g in range(G):
        rs = np.random.RandomState(1234 + g)
        while True:  # bench.synthetic_positions: redraw positions whose random play already ended the game
            cells, meta = bench.draw_position(rs)
            b = OBoard(W, H, 5)
            b.init_board(0)
            b.states = {int(m): int(cells[m]) for m in np.nonzero(cells)[0]}
            b.availables = [m for m in range(W * H) if m not in b.states]
            if not b.game_end()[0]:
                break
        key = (seed & 0xFFFFFFFF, seed >> 32)
        draw = np.zeros(W * H, np.int64)
        for lane in range(30):  # rows 0..14 = lanes 0..29
            words = philox4x32_10((0, 0, g, lane), key) + philox4x32_10((0, 1, g, lane), key)
            for r in range(8):
                q = (lane & 1) * 8 + r
                col = ((q & 3) << 2) | (q >> 2)
                if col < W:
                    draw[(lane >> 1) * W + col] = (min(words[r] >> 8, 0xFFFFFE) << 8) | (lane * 8 + r)
        empties = np.array(b.availables)
        full[g] = np.where(cells == 1, -2, -1)  # side to move is always player 1 in these positions
        full[g, empties[np.argsort(draw[empties])]] = np.arange(len(empties))
        n_empty[g] = len(empties)
    v, p = outcomes_from_ranks_numpy(full, n_empty, W, H, 5)
    assert "%.1f" % p.mean() == want_plies and "%.4f" % v.mean() == want_value, (p.mean(), v.mean(), line)
    assert int(v.sum()) == 360


def test_reference_arm_positions_are_the_gpu_arms():
    """bench.py --impl reference searches from the GPU arm's own synthetic positions (RandomState(1234 + g),
    0 - 60 stones, redrawn while the random play already ended the game): the oracle board rebuilt from
    draw_position's (cells, meta) carries the same stones, side to move, last move and last-four history."""
    sys.path.insert(0, ROOT)
    import bench
    for g in (0, 1, 5, 77, 4095):
        b = bench.synthetic_position_cpu(g)
        cells, meta = bench.draw_position(np.random.RandomState(1234 + g))  # first draw of the same stream
        if not bench.oracle_board_from_position(cells, meta).game_end()[0]:   # (kept unless it had to be redrawn)
            assert {m: int(cells[m]) for m in np.nonzero(cells)[0]} == b.states and int(meta[1]) == b.last_move
        assert not b.game_end()[0]
        assert len(b.states) % 2 == 0 and b.current_player == 1 and len(b.states) <= 60
        assert sorted(b.availables) == [m for m in range(225) if m not in b.states]
        if b.states:
            assert b.last_move == b.history[-1][0] and b.states[b.last_move] == 2
            assert [m for m, _ in b.history[-4:]] == [m for m, _ in b.history][-4:]
        st = b.current_state()
        assert st.shape == (9, 15, 15) and st[8].all()           # even stone count: colour plane all ones
        assert st[6].sum() + st[7].sum() == len(b.states)         # planes 6 / 7: all stones of both sides
