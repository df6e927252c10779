"""GPU: the TrainPipeline shim (train_mxnet.py) end to end on a small board: SGF bootstrap into the device ring,
batched self-play, policy_update with the KL / lr rule, batched arena, checkpoint round trip."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_sgf(dirname, name, moves):
    body = ";".join("%s[%s%s]" % ("BW"[i % 2], "abcdefghijklmno"[m // 15], "abcdefghijklmno"[m % 15])
                    for i, m in enumerate(moves))
    with open(os.path.join(dirname, name), "w") as f:
        f.write("(;GM[4]FF[4]SZ[15]\n" + body + ")\n\n\n")


def test_sgf_reader_matches_reference_slicing(tmp_path):
    """utils/sgf_dataIter.py:45-66: body = text[index('SZ[15]')+7 : -4], move = 2 letters after 'B[' / 'W['."""
    from alphapig_b200.utils import sgf_dataIter
    moves = [112, 113, 97, 98, 127]
    _write_sgf(str(tmp_path), "0001_Blank_x_.sgf", moves)
    _write_sgf(str(tmp_path), "0002_white_x_.sgf", moves[:4])
    a = sgf_dataIter.get_data_from_files("0001_Blank_x_.sgf", str(tmp_path))
    b = sgf_dataIter.get_data_from_files("0002_white_x_.sgf", str(tmp_path))
    assert a["seq_num_list"] == moves and a["winner"] == 1 and a["seq_list"][0] == "hh"
    assert b["seq_num_list"] == moves[:4] and b["winner"] == 2
    assert sorted(sgf_dataIter.get_files_as_list(str(tmp_path))) == ["0001_Blank_x_.sgf", "0002_white_x_.sgf"]
    assert sgf_dataIter.num2char(112) == "HH"


def test_train_pipeline_small(tmp_path):
    from alphapig_b200 import checkpoint
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    from alphapig_b200.train_mxnet import TrainPipeline
    W = 15
    sgf = tmp_path / "sgf"
    sgf.mkdir()
    rs = np.random.RandomState(0)
    for g in range(6):
        mv = [int(m) for m in rs.permutation(W * W)[:20 + g]]
        _write_sgf(str(sgf), "%04d_%s_x_.sgf" % (g, "Blank" if g % 2 else "White"), mv)
    net = PolicyValueNet(W, W, batch_size=32, seed=0)
    conf = dict(board_width=W, board_height=W, n_in_row=5, n_playout=16, batch_size=32, epochs=3, buffer_size=4000,
                sgf_dir=str(sgf), pure_mcts_playout_num=30, selfplay_games=16, check_freq=1000, game_batch_num=2)
    random.seed(0)
    np.random.seed(0)
    tp = TrainPipeline(conf, net=net)
    # SGF bootstrap: 6 records in one launch
    tp.collect_selfplay_data(6, training_index=0)
    assert len(tp.data_buffer) == 8 * sum(20 + g for g in range(6))
    st, pi, z = tp.data_buffer.sample(8)
    assert st.shape == (8, 9, W, W) and np.allclose(pi.sum(1), 0.99999 + 224e-6, atol=1e-5) and set(np.abs(z)) == {1.0}
    # learning: loss is finite, weights move, KL rule updates the multiplier
    before = net.get_policy_param()[0]["conv1_weight"].copy()
    loss, entropy = tp.policy_update()
    assert np.isfinite(loss).all() and np.isfinite(entropy).all()
    assert not np.array_equal(before, net.get_policy_param()[0]["conv1_weight"])
    assert tp.lr_multiplier in (1.0, 1.5, 1.0 / 1.5)
    # batched self-play + arena on a small board (games end within a few dozen plies)
    net6 = PolicyValueNet(6, 6, batch_size=32, seed=1)
    tp6 = TrainPipeline(dict(board_width=6, board_height=6, n_in_row=4, n_playout=16, batch_size=32, epochs=2,
                             buffer_size=2000, sgf_dir=str(tmp_path / "none"), pure_mcts_playout_num=30,
                             selfplay_games=16), net=net6)
    fin = 0
    for _ in range(40):
        fin += tp6.collect_selfplay_data_batched(1)
        if fin >= 4:
            break
    assert fin >= 4 and len(tp6.data_buffer) >= 8 * 4 * 7  # a 4-in-row game lasts at least 7 plies
    st6, pi6, z6 = tp6.data_buffer.sample(16)
    assert np.allclose(pi6.sum(1), 1.0, atol=1e-5) and set(np.unique(z6)) <= {-1.0, 0.0, 1.0}
    loss6, _ = tp6.policy_update()
    assert np.isfinite(loss6).all()
    # batched arena vs pure MCTS: every game ends with a legal result
    ratio = tp6.policy_evaluate_batched(n_games=6)
    assert 0.0 <= ratio <= 1.0 and (ratio * 12) == int(ratio * 12)
    # checkpoint round trip: reference-style pickle and npz
    model = str(tmp_path / "cur.model")
    net.save_model(model)
    arg, aux = checkpoint.load_model(model)
    checkpoint.save_npz(str(tmp_path / "cur.npz"), (arg, aux))
    arg2, aux2 = checkpoint.load_npz(str(tmp_path / "cur.npz"))
    assert list(arg) == list(arg2) and all(np.array_equal(arg[k], arg2[k]) for k in arg)
    net2 = PolicyValueNet(W, W, batch_size=32, model_params=(arg2, aux2))
    x = st[:4]
    p1, v1 = net.policy_value(x)
    p2, v2 = net2.policy_value(x)
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2)


def _play_and_collect(device_records, W, n_in_row, G, n_playout, seed, max_plies, opening_prob):
    """finished-game records of BatchedSelfPlay(device_pick=True), host-assembled or from the device outbox, as one
    (bits, pi, z) triple in emission order"""
    from alphapig_b200 import dist as apdist
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    from alphapig_b200.selfplay import BatchedSelfPlay
    S = W * W
    net = PolicyValueNet(W, W, batch_size=32, seed=0, n_in_row=n_in_row)
    sp = BatchedSelfPlay(net, n_games=G, n_playout=n_playout, n_in_row=n_in_row, seed=seed, device_pick=True,
                         device_records=device_records, forced_opening_prob=opening_prob)
    bits, pis, zs, winners, boxes = [], [], [], [], []
    if device_records:
        sp.boundary_hook = lambda s: boxes.append(s.take_outbox())
    for _ in range(max_plies):
        for winner, st, pi, z in sp.step():
            winners.append(winner)
            if not device_records:
                bits.append(st)
                pis.append(pi.astype(np.float32))
                zs.append(z.astype(np.float32))
    sp.drain()
    forced = sp.forced_openings
    if device_records:
        rec = np.concatenate([b.cpu().numpy() for b in boxes], axis=0)
        assert rec.shape[1] == apdist.record_width(S)
        b_, p_, z_ = apdist.split_records(rec, S)
        out = (b_, p_, z_)
    else:
        out = (np.concatenate(bits), np.concatenate(pis), np.concatenate(zs))
    net.close()
    return out, winners, forced


@pytest.mark.parametrize("W,n_in_row,G,n_playout,plies,opening", [(6, 4, 48, 20, 30, 0.0), (15, 5, 12, 12, 150, 1.0)])
def test_device_trajectories_equal_host_assembled_records(W, n_in_row, G, n_playout, plies, opening):
    """Device-side trajectories + outbox (csrc/traj.cu: every pick appends its ply's packed record in HBM, finished games
    move to the outbox with z filled in) against the host-assembled records of the same self-play run (same Philox
    seeds => same moves): bit-identical state planes, pi and z, in the same order - including, on 15x15, the forced
    random two-ply opening records of game_ai.py:78-111 (probability forced to 1 here)."""
    host, w_host, f_host = _play_and_collect(False, W, n_in_row, G, n_playout, 5, plies, opening)
    devc, w_dev, f_dev = _play_and_collect(True, W, n_in_row, G, n_playout, 5, plies, opening)
    assert w_host == w_dev and len(w_host) >= (20 if W == 6 else 3) and f_host == f_dev
    if opening:
        assert f_host >= G
    assert host[0].shape == devc[0].shape and host[0].shape[0] > 0
    assert np.array_equal(host[0], devc[0]), "state bits"
    assert np.array_equal(host[1], devc[1]), "pi"
    assert np.array_equal(host[2], devc[2]), "z"
    if opening:  # the first two records of a game are the forced plies: pi = 0.99999 at the move, 1e-6 elsewhere
        pi0 = devc[1][0]
        assert np.isclose(pi0.max(), 0.99999) and np.isclose(np.sort(pi0)[-2], 1e-6)


def test_ring_push_packed_equals_push():
    """ap_replay_push_packed (host pointer and device pointer) fills the ring exactly as ap_replay_push does."""
    import torch
    from alphapig_b200 import dist as apdist
    from alphapig_b200.engine import Engine
    W = 8
    S = W * W
    rs = np.random.RandomState(3)
    n = 37
    bits = np.packbits((rs.rand(n, 9 * S) > 0.6).astype(np.uint8), axis=1)
    pis = rs.dirichlet(np.ones(S), size=n).astype(np.float32)
    zs = rs.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32)
    packed = apdist.pack_records(bits, pis, zs, S)
    engs = [Engine(width=W, height=W, n_games=1) for _ in range(3)]
    for e in engs:
        e.replay_create(8 * 20)  # smaller than n records: the ring wraps
    engs[0].replay_push(bits, pis, zs)
    engs[1].replay_push_packed(packed)
    t = torch.from_numpy(packed).cuda()
    engs[2].replay_push_packed(None, n=n, device_ptr=t.data_ptr())
    ref = engs[0].replay_gather(np.arange(8 * 20))
    for e in engs[1:]:
        assert e.replay_size() == engs[0].replay_size()
        got = e.replay_gather(np.arange(8 * 20))
        assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    for e in engs:
        e.close()


@pytest.mark.parametrize("overlap", [True, False])
def test_selfplay_train_loop_single_gpu(overlap):
    """configs[4] on one GPU, small board: games finish, their records reach the ring (device outbox -> packed push in the
    overlapped loop), policy_update runs (in the trainer thread, overlapped with search), new weights are swapped into
    the search engine at ply boundaries and the searches keep working on them."""
    from alphapig_b200.loop import selfplay_train_loop
    from alphapig_b200.policy_value_net_mxnet_simple import PolicyValueNet
    W = 6
    net = PolicyValueNet(W, W, batch_size=32, seed=0, n_in_row=4)
    w0 = net.get_policy_param()[0]["conv1_weight"].copy()
    res = selfplay_train_loop(net, 64, 4, plies_per_iter=8, n_playout=16, batch_size=32, epochs=3, buffer_size=20000,
                              n_in_row=4, seed=1, overlap=overlap, warmup_iters=1 if overlap else 0)
    assert res["overlap"] == overlap and res["games"] > 20 and res["records"] >= 7 * res["games"] * (0 if overlap else 1)
    assert res["train_steps"] >= 1 and np.isfinite(res["losses"]).all()  # (KL early stop may end a policy_update after one step)
    assert res["playouts"] == 4 * 8 * 64 * 16 and res["t_total"] > 0
    if overlap:
        # nobody waits for the trainer: a swap happens whenever a policy_update had finished at an iteration boundary
        assert 1 <= res["weight_swaps"] <= 5 and res["ring_records"] >= res["records"] > 0
        assert res["lr_multiplier"] in (1.0, 1.5, 2.25, 3.375, 1 / 1.5, 1 / 2.25, 1 / 3.375, 5.0625, 1 / 5.0625)
    assert not np.array_equal(w0, net.get_policy_param()[0]["conv1_weight"])
    net.close()
