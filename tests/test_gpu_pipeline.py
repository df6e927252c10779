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
