"""Oracle restatement of the data path either side of self-play in the reference ``TrainPipeline``
(SURVEY 8(f) rows 1-2).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pinned live against the imported reference
``train_mxnet.TrainPipeline`` methods in ``tests/test_oracle_vs_reference.py`` and by the golden
vectors ``tests/golden/pipeline_cases.npz``.

* ``equi_data``      : ``train_mxnet.py:115-135`` (8-fold rotation / flip augmentation)
* ``ReplayDeque``    : ``train_mxnet.py:57,153,180,196`` (``deque(maxlen)`` + ``random.sample``)
* ``policy_update``  : ``train_mxnet.py:194-237`` (KL early stop, adaptive lr multiplier)
* ``policy_evaluate``: ``train_mxnet.py:239-263`` (arena vs pure MCTS, win ratio)
"""
import random
from collections import defaultdict, deque

import numpy as np


def equi_data(play_data, board_height, board_width):
    """[(state (C,H,W), pi (S,), z)] -> 8x as many, in the reference's order: for i in 1..4:
    (rot90^i), (rot90^i then fliplr)."""
    out = []
    for state, pi, z in play_data:
        for i in (1, 2, 3, 4):
            es = np.array([np.rot90(s, i) for s in state])
            ep = np.rot90(np.flipud(pi.reshape(board_height, board_width)), i)
            out.append((es, np.flipud(ep).flatten(), z))
            es = np.array([np.fliplr(s) for s in es])
            ep = np.fliplr(ep)
            out.append((es, np.flipud(ep).flatten(), z))
    return out


class ReplayDeque(object):
    """``deque(maxlen=buffer_size)`` of augmented samples, sampled with ``random.sample``."""

    def __init__(self, buffer_size, board_height, board_width):
        self.buf = deque(maxlen=buffer_size)
        self.h, self.w = board_height, board_width

    def extend_game(self, play_data):
        self.buf.extend(equi_data(list(play_data), self.h, self.w))

    def __len__(self):
        return len(self.buf)

    def sample(self, batch_size):
        mini = random.sample(self.buf, batch_size)
        return [d[0] for d in mini], [d[1] for d in mini], [d[2] for d in mini]


def policy_update(net, replay, batch_size, learn_rate, lr_multiplier, epochs, kl_targ):
    """-> (loss, entropy, new lr_multiplier, kl, epochs_run, explained_var_old, explained_var_new)"""
    state_batch, mcts_probs_batch, winner_batch = replay.sample(batch_size)
    old_probs, old_v = net.policy_value(state_batch)
    lr = learn_rate * lr_multiplier
    ran = 0
    for i in range(epochs):
        loss, entropy = net.train_step(state_batch, mcts_probs_batch, winner_batch, lr)
        new_probs, new_v = net.policy_value(state_batch)
        kl = np.mean(np.sum(old_probs * (np.log(old_probs + 1e-10) - np.log(new_probs + 1e-10)), axis=1))
        ran += 1
        if kl > kl_targ * 4:
            break
    if kl > kl_targ * 2 and lr_multiplier > 0.05:
        lr_multiplier /= 1.5
    elif kl < kl_targ / 2 and lr_multiplier < 20:
        lr_multiplier *= 1.5
    ev_old = 1 - np.var(np.array(winner_batch) - old_v.flatten()) / np.var(np.array(winner_batch))
    ev_new = 1 - np.var(np.array(winner_batch) - new_v.flatten()) / np.var(np.array(winner_batch))
    return loss, entropy, lr_multiplier, kl, ran, ev_old, ev_new


def policy_evaluate(start_play, current_player, pure_player, n_games=10):
    """``start_play(p1, p2, start_player) -> winner``; win ratio of player 1 (ties count half)."""
    win_cnt = defaultdict(int)
    for i in range(n_games):
        win_cnt[start_play(current_player, pure_player, i % 2)] += 1
    return 1.0 * (win_cnt[1] + 0.5 * win_cnt[-1]) / n_games, dict(win_cnt)
