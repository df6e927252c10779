"""Drop-in for the reference ``train_mxnet.TrainPipeline`` (train_mxnet.py:37-283): same constructor
(``conf`` dict with the keys of conf/train_config.yaml, optional ``init_model``), same methods
(``get_equi_data``, ``collect_selfplay_data`` (SGF), ``collect_selfplay_data_ai``, ``policy_update``,
``policy_evaluate``, ``run``), same KL early-stop / adaptive learning-rate rule.

What moved to the GPU behind that face:

* ``data_buffer`` is the device replay ring (``alphapig_b200.replay.ReplayBuffer``): 8-fold augmentation is
  applied by the gather kernel, minibatches reach ``train_step`` as device tensors, and
  ``random.sample`` picks the same indices as it would on the reference's deque.
* ``collect_selfplay_data(n_games)`` replays n SGF records in one kernel launch straight into the ring.
* ``collect_selfplay_data_ai`` keeps the reference's one-game-at-a-time ``Game_AI.start_self_play``;
  ``collect_selfplay_data_batched`` plays ``conf['selfplay_games']`` concurrent games per GPU
  (``alphapig_b200.selfplay.BatchedSelfPlay``) - what ``bench.py`` measures.
* ``policy_evaluate`` keeps the reference's sequential arena; ``policy_evaluate_batched`` plays all arena
  games concurrently on the device (AlphaZero search via ``ap_search_run``, pure MCTS via ``ap_pure_run``).
* multi-GPU (``run`` under ``torch.distributed``): every rank plays its own self-play games on its own GPU; the
  records of each iteration are gathered to rank 0 (``alphapig_b200.dist.gather_replay``) and pushed into ITS
  ring, rank 0 runs ``policy_update`` and the weights are broadcast (``dist.broadcast_weights``).  The SGF
  bootstrap phase is data loading, not search: rank 0 does it alone.

Logging config, e-mail and the YAML loader of the reference are out of scope: ``conf`` is a plain dict.
"""
from __future__ import print_function

import logging
import os
import random
import time
from collections import defaultdict

import numpy as np

from .game import Board, Game
from .game_ai import Game_AI
from .mcts_alphaZero import MCTSPlayer
from .mcts_pure import MCTSPlayer as MCTS_Pure
from .replay import ReplayBuffer
from .utils import sgf_dataIter

_logger = logging.getLogger(__name__)

DEFAULT_CONF = {
    'board_width': 15, 'board_height': 15, 'n_in_row': 5, 'learn_rate': 0.0004, 'lr_multiplier': 1.0, 'temp': 1.0,
    'n_playout': 400, 'c_puct': 5, 'buffer_size': 2198800, 'batch_size': 128, 'epochs': 8, 'play_batch_size': 1,
    'kl_targ': 0.02, 'check_freq': 1000, 'pure_mcts_playout_num': 1000, 'game_batch_num': 240000,
    'sgf_dir': './sgf_data', 'ai_data_dir': './pickle_ai_data',
}


def kl_and_lr_rule(old_probs, new_probs, kl_targ, lr_multiplier):
    """KL(old || new) averaged over the batch and the lr-multiplier update (train_mxnet.py:209-221)."""
    kl = np.mean(np.sum(old_probs * (np.log(old_probs + 1e-10) - np.log(new_probs + 1e-10)), axis=1))
    if kl > kl_targ * 2 and lr_multiplier > 0.05:
        lr_multiplier /= 1.5
    elif kl < kl_targ / 2 and lr_multiplier < 20:
        lr_multiplier *= 1.5
    return kl, lr_multiplier


class TrainPipeline(object):
    def __init__(self, conf, init_model=None, net=None, device=0, sgf_bootstrap_batches=4000):
        c = dict(DEFAULT_CONF)
        c.update(conf)
        self.conf = c
        self.board_width, self.board_height, self.n_in_row = c['board_width'], c['board_height'], c['n_in_row']
        self.board = Board(width=self.board_width, height=self.board_height, n_in_row=self.n_in_row)
        self.game = Game(self.board, sgf_loader=sgf_dataIter.get_data_from_files)
        self.game_ai = Game_AI(self.board)
        self.learn_rate = c['learn_rate']
        self.lr_multiplier = c['lr_multiplier']
        self.temp = c['temp']
        self.n_playout = c['n_playout']
        self.c_puct = c['c_puct']
        self.buffer_size = c['buffer_size']
        self.batch_size = c['batch_size']
        self.play_batch_size = c['play_batch_size']
        self.epochs = c['epochs']
        self.kl_targ = c['kl_targ']
        self.check_freq = c['check_freq']
        self.game_batch_num = c['game_batch_num']
        self.best_win_ratio = 0.0
        self.pure_mcts_playout_num = c['pure_mcts_playout_num']
        self._sgf_bootstrap_batches = sgf_bootstrap_batches  # train_mxnet.py:270: SGF data for the first 4000 batches
        self._sgf_home = c['sgf_dir']
        self._training_data = []
        if os.path.isdir(self._sgf_home):
            self._load_training_data(self._sgf_home)
        self._length_train_data = len(self._training_data)
        if net is not None:
            self.policy_value_net = net
        else:
            # the reference hard-codes the 10 x 128 residual net (train_mxnet.py:79-91); conf may override
            from .policy_value_net_mxnet import PolicyValueNet
            self.policy_value_net = PolicyValueNet(self.board_width, self.board_height, self.batch_size,
                                                   n_blocks=c.get('n_blocks', 10), n_filter=128,
                                                   model_params=init_model, device=device)
        self.mcts_player = MCTSPlayer(self.policy_value_net.policy_value_fn, c_puct=self.c_puct,
                                      n_playout=self.n_playout, is_selfplay=1)
        self.data_buffer = ReplayBuffer(self.policy_value_net._eng, self.buffer_size)
        self.episode_len = 0
        self._batched = None
        self._pending = None  # multi-rank run(): packed records of this iteration, gathered to rank 0 afterwards

    # -- data ---------------------------------------------------------------------------------
    def _load_training_data(self, data_dir):
        self._training_data = sgf_dataIter.get_files_as_list(data_dir)
        random.shuffle(self._training_data)
        self._length_train_data = len(self._training_data)

    def get_equi_data(self, play_data):
        """8-fold rotation / flip augmentation on the host (train_mxnet.py:115-135); kept for callers that
        want the tuples - the pipeline itself stores un-augmented records and augments in the gather kernel."""
        extend_data = []
        for state, mcts_porb, winner in play_data:
            for i in [1, 2, 3, 4]:
                equi_state = np.array([np.rot90(s, i) for s in state])
                equi_mcts_prob = np.rot90(np.flipud(mcts_porb.reshape(self.board_height, self.board_width)), i)
                extend_data.append((equi_state, np.flipud(equi_mcts_prob).flatten(), winner))
                equi_state = np.array([np.fliplr(s) for s in equi_state])
                equi_mcts_prob = np.fliplr(equi_mcts_prob)
                extend_data.append((equi_state, np.flipud(equi_mcts_prob).flatten(), winner))
        return extend_data

    def collect_selfplay_data(self, n_games=1, training_index=None):
        """SGF records -> training data (train_mxnet.py:137-154); all n_games records go through ONE device
        replay launch.  Records with an illegal move are skipped with an error log, as in the reference."""
        if not self._length_train_data:
            raise RuntimeError("no SGF records under %r" % self._sgf_home)
        data_index = training_index % self._length_train_data
        if data_index == 0:
            random.shuffle(self._training_data)
        names = [self._training_data[(data_index + i) % self._length_train_data] for i in range(n_games)]
        recs = [sgf_dataIter.get_data_from_files(nm, self._sgf_home) for nm in names]
        warn = self.policy_value_net._eng.replay_push_sgf([r['seq_num_list'] for r in recs], [r['winner'] for r in recs])
        for nm, r, w in zip(names, recs, warn):
            if w:
                _logger.error('WARNING training_index: %s, data_index: %s, file: %s', training_index, data_index, nm)
            else:
                _logger.info('winner: %s, file: %s ', r['winner'], nm)
                self.episode_len = len(r['seq_num_list'])
        _logger.info('game_batch_index: %s, length of data_buffer: %s', training_index, len(self.data_buffer))

    def _store(self, states, pis, zs):
        """``data_buffer.extend(get_equi_data(play_data))`` for one game's un-augmented positions - into the local
        ring, or (multi-rank ``run``) into the packed records that go to the trainer rank."""
        states = np.asarray(states)
        if states.dtype != np.uint8 or states.ndim != 2:
            states = np.packbits(states.reshape(states.shape[0], -1).astype(np.uint8), axis=1)
        if self._pending is None:
            self.data_buffer.extend_positions(states, pis, zs)
        else:
            from . import dist as apdist
            self._pending.append(apdist.pack_records(states, pis, zs, self.board_width * self.board_height))

    def collect_selfplay_data_ai(self, n_games=1, training_index=None):
        """One reference-style self-play game at a time (train_mxnet.py:171-180)."""
        for _ in range(n_games):
            winner, play_data = self.game_ai.start_self_play(self.mcts_player, temp=self.temp)
            _logger.info('traing_index: %s,   winner is: %s', training_index, winner)
            play_data = list(play_data)[:]
            self.episode_len = len(play_data)
            if play_data:
                self._store(np.stack([np.asarray(st) for st, _, _ in play_data]),
                            np.stack([np.asarray(pi) for _, pi, _ in play_data]),
                            np.array([z for _, _, z in play_data], dtype=np.float32))

    def collect_selfplay_data_batched(self, n_plies=1, n_games=None, seed=0):
        """``n_games`` concurrent self-play games advance ``n_plies`` plies; finished games go into the ring."""
        from .selfplay import BatchedSelfPlay
        if self._batched is None:
            self._batched = BatchedSelfPlay(self.policy_value_net, n_games or self.conf.get('selfplay_games', 1024),
                                            n_playout=self.n_playout, c_puct=self.c_puct, temp=self.temp,
                                            n_in_row=self.n_in_row, seed=seed)
        finished = 0
        for _ in range(n_plies):
            for winner, states, pis, zs in self._batched.step():
                if states is not None:
                    self._store(states, pis, zs)
                    self.episode_len = len(zs)
                    finished += 1
        return finished

    # -- learning -----------------------------------------------------------------------------
    def policy_update(self):
        """update the policy-value net (train_mxnet.py:194-237)"""
        net = self.policy_value_net
        state_batch, mcts_probs_batch, winner_batch = self.data_buffer.sample(self.batch_size)
        old_probs, old_v = net.policy_value(state_batch)
        learn_rate = self.learn_rate * self.lr_multiplier
        for i in range(self.epochs):
            loss, entropy = net.train_step(state_batch, mcts_probs_batch, winner_batch, learn_rate)
            new_probs, new_v = net.policy_value(state_batch)
            kl = np.mean(np.sum(old_probs * (np.log(old_probs + 1e-10) - np.log(new_probs + 1e-10)), axis=1))
            if kl > self.kl_targ * 4:  # early stopping if D_KL diverges badly
                _logger.info('early stopping. i:%s.   epochs: %s', i, self.epochs)
                break
        _, self.lr_multiplier = kl_and_lr_rule(old_probs, new_probs, self.kl_targ, self.lr_multiplier)
        wb = np.array(winner_batch)
        explained_var_old = 1 - np.var(wb - old_v.flatten()) / np.var(wb)
        explained_var_new = 1 - np.var(wb - new_v.flatten()) / np.var(wb)
        _logger.info("kl:%.4f,lr:%.1e,loss:%s,entropy:%s,explained_var_old:%.3f,explained_var_new:%.3f",
                     kl, learn_rate, loss, entropy, explained_var_old, explained_var_new)
        self.last_kl = kl
        return loss, entropy

    def policy_evaluate(self, n_games=10):
        """Sequential arena vs pure MCTS (train_mxnet.py:239-263); returns the win ratio."""
        current_mcts_player = MCTSPlayer(self.policy_value_net.policy_value_fn, c_puct=self.c_puct,
                                         n_playout=self.n_playout)
        pure_mcts_player = MCTS_Pure(c_puct=5, n_playout=self.pure_mcts_playout_num)
        win_cnt = defaultdict(int)
        for i in range(n_games):
            winner = self.game.start_play(current_mcts_player, pure_mcts_player, start_player=i % 2, is_shown=0)
            win_cnt[winner] += 1
        win_ratio = 1.0 * (win_cnt[1] + 0.5 * win_cnt[-1]) / n_games
        _logger.info("num_playouts:%s, win: %s, lose: %s, tie:%s", self.pure_mcts_playout_num, win_cnt[1], win_cnt[2],
                     win_cnt[-1])
        return win_ratio

    def policy_evaluate_batched(self, n_games=10, seed=0):
        """The same arena with all games played concurrently on the device.  Game i starts with player
        ``i % 2`` as in the reference; player 1 is the AlphaZero searcher in every game."""
        from .arena import batched_arena
        winners = batched_arena(self.policy_value_net, n_games, n_playout=self.n_playout, c_puct=self.c_puct,
                                pure_n_playout=self.pure_mcts_playout_num, n_in_row=self.n_in_row, seed=seed)
        win_cnt = defaultdict(int)
        for w in winners:
            win_cnt[int(w)] += 1
        return 1.0 * (win_cnt[1] + 0.5 * win_cnt[-1]) / n_games

    def run(self, model_dir='./logs', batched=False):
        """run the training pipeline (train_mxnet.py:265-283).  Under torch.distributed every rank collects
        self-play games on its own GPU, the iteration's records are gathered to rank 0, rank 0 trains and the
        weights are broadcast after each update."""
        import torch.distributed as dist
        from . import dist as apdist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        rank = dist.get_rank() if multi else 0
        S = self.board_width * self.board_height
        try:
            for i in range(self.game_batch_num):
                t0 = time.time()
                sgf_phase = i < self._sgf_bootstrap_batches and self._length_train_data
                if sgf_phase:
                    if rank == 0:
                        self.collect_selfplay_data(self.play_batch_size, training_index=i)
                else:
                    self._pending = [] if multi else None
                    if batched:
                        self.collect_selfplay_data_batched(1)
                    else:
                        self.collect_selfplay_data_ai(self.play_batch_size, training_index=i)
                    if multi:
                        mine = (np.concatenate(self._pending, axis=0) if self._pending
                                else np.zeros((0, apdist.record_width(S)), np.uint8))
                        self._pending = None
                        allrec = apdist.gather_replay(mine)
                        if rank == 0 and allrec.shape[0]:
                            bits, pis, zs = apdist.split_records(allrec, S)
                            self.data_buffer.extend_positions(bits, pis, zs)
                _logger.info('collection cost time: %d ', time.time() - t0)
                _logger.info("batch i:%s, episode_len:%s, buffer_len:%s", i + 1, self.episode_len, len(self.data_buffer))
                if rank == 0 and len(self.data_buffer) > self.batch_size:
                    self.policy_update()
                if multi:
                    apdist.broadcast_weights(self.policy_value_net, src=0)
                if rank != 0:
                    continue
                if (i + 1) % 50 == 0:
                    os.makedirs(model_dir, exist_ok=True)
                    self.policy_value_net.save_model(os.path.join(model_dir, 'current_policy.model'))
                if (i + 1) % self.check_freq == 0:
                    win_ratio = self.policy_evaluate_batched() if batched else self.policy_evaluate()
                    if win_ratio > self.best_win_ratio:
                        _logger.info("New best policy!!!!!!!!")
                        self.best_win_ratio = win_ratio
                        os.makedirs(model_dir, exist_ok=True)
                        self.policy_value_net.save_model(os.path.join(model_dir, 'best_policy_%s.model' % i))
                        if self.best_win_ratio >= 0.98 and self.pure_mcts_playout_num < 8000:
                            self.pure_mcts_playout_num += 1000
                            self.best_win_ratio = 0.0
        except KeyboardInterrupt:
            _logger.info('\n\rquit')
