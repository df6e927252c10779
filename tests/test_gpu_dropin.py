"""GPU: the reference callers' call sequences through the ALIASED module names (``alphapig_b200.install()``).

The GPU box has no /root/reference, so the sequences are restated here call for call; on the build box
tests/test_cpu_dropin.py executes the unmodified reference files on the same aliases."""
import os
import pickle
import random
import sys
from collections import deque

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def aliases():
    import alphapig_b200
    done = alphapig_b200.install(force=True)
    yield done
    alphapig_b200.uninstall()


def test_train_pipeline_sequence_through_aliases(aliases):
    """train_mxnet.py:18-29 imports; TrainPipeline.__init__ (:44-48, 79-95) -> collect_selfplay_data_ai(1) (:171-180)
    -> policy_update (:194-237) with the reference's own deque / random.sample / get_equi_data code shape."""
    from game import Board, Game
    from game_ai import Game_AI
    from mcts_pure import MCTSPlayer as MCTS_Pure
    from mcts_alphaZero import MCTSPlayer
    from utils import config_loader, send_email  # noqa: F401
    from policy_value_net_mxnet import PolicyValueNet
    W = H = 8
    batch_size = 32
    random.seed(3)
    np.random.seed(3)
    import torch
    torch.manual_seed(3)  # Dropout masks of train_step
    board = Board(width=W, height=H, n_in_row=5)
    game = Game(board)
    game_ai = Game_AI(board)  # train_mxnet.py:47-48: Game and Game_AI share ONE board
    assert game.board is game_ai.board
    data_buffer = deque(maxlen=10000)
    net = PolicyValueNet(W, H, batch_size, n_blocks=2, n_filter=128)
    player = MCTSPlayer(net.policy_value_fn, c_puct=5, n_playout=40, is_selfplay=1)
    # collect_selfplay_data_ai(1); random.random patched as SURVEY G2 prescribes for boards narrower than 15
    import game_ai as game_ai_module
    real_random = game_ai_module.random.random
    game_ai_module.random.random = lambda: 0.5
    try:
        for _ in range(2):
            winner, play_data = game_ai.start_self_play(player, temp=1.0)
            play_data = list(play_data)[:]
            assert winner in (1, 2, -1) and len(play_data) >= 9
            # get_equi_data (:115-135)
            for state, mcts_porb, winner_z in play_data:
                assert state.shape == (9, W, H) and state.dtype == np.float64
                assert mcts_porb.shape == (W * H,) and abs(mcts_porb.sum() - 1.0) < 1e-9 and winner_z in (1.0, -1.0, 0.0)
                for i in [1, 2, 3, 4]:
                    equi_state = np.array([np.rot90(s, i) for s in state])
                    equi_mcts_prob = np.rot90(np.flipud(mcts_porb.reshape(H, W)), i)
                    data_buffer.append((equi_state, np.flipud(equi_mcts_prob).flatten(), winner_z))
                    equi_state = np.array([np.fliplr(s) for s in equi_state])
                    equi_mcts_prob = np.fliplr(equi_mcts_prob)
                    data_buffer.append((equi_state, np.flipud(equi_mcts_prob).flatten(), winner_z))
    finally:
        game_ai_module.random.random = real_random
    assert len(data_buffer) > batch_size
    # policy_update
    mini_batch = random.sample(list(data_buffer), batch_size)
    state_batch = [d[0] for d in mini_batch]
    mcts_probs_batch = [d[1] for d in mini_batch]
    winner_batch = [d[2] for d in mini_batch]
    old_probs, old_v = net.policy_value(state_batch)
    assert old_probs.shape == (batch_size, W * H) and old_v.shape == (batch_size, 1)
    losses = []
    for _ in range(5):
        loss, entropy = net.train_step(state_batch, mcts_probs_batch, winner_batch, 2e-3)
        assert loss.shape == (1,) and entropy.shape == (1,)
        losses.append(float(loss[0]))
    new_probs, new_v = net.policy_value(state_batch)
    kl = np.mean(np.sum(old_probs * (np.log(old_probs + 1e-10) - np.log(new_probs + 1e-10)), axis=1))
    assert kl > 0 and losses[-1] < losses[0]
    # the self-play player's search engine follows the trained weights (train -> predict copy, :295-297)
    b2 = Board(width=W, height=H, n_in_row=5)
    b2.init_board()
    acts, probs = player.mcts.get_move_probs(b2, temp=1.0)
    assert len(acts) == W * H and abs(probs.sum() - 1) < 1e-9
    # policy_evaluate's two players (:244-248) and one arena game (:250-253)
    current_mcts_player = MCTSPlayer(net.policy_value_fn, c_puct=5, n_playout=30)
    pure_mcts_player = MCTS_Pure(c_puct=5, n_playout=200)
    assert current_mcts_player.mcts._engine(b2) is not player.mcts._engine(b2)  # every MCTS owns its tree
    w = game.start_play(current_mcts_player, pure_mcts_player, start_player=1, is_shown=0)
    assert w in (1, 2, -1)
    # save_model / pickle.load round trip as train_mxnet.py:286-293 reads it
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        net.save_model(os.path.join(d, 'current_policy.model'))
        policy_param = pickle.load(open(os.path.join(d, 'current_policy.model'), 'rb'))
        net2 = PolicyValueNet(W, H, batch_size, n_blocks=2, n_filter=128, model_params=policy_param)
        p2, v2 = net2.policy_value(state_batch)
        assert np.array_equal(p2, new_probs) and np.array_equal(v2, new_v)


def test_human_play_and_chess_client_wiring(aliases):
    """human_play_mxnet.py:51-81 (Board, Game, PolicyValueNet from a pickled model, MCTSPlayer, start_play with a scripted
    'human') and evaluate/ChessClient.py:189-235 (``import AlphaPig as gomoku_zm``: simple net, c_puct 3, n_playout 80,
    board.states / do_move / get_action(board))."""
    from game import Board, Game
    from mcts_alphaZero import MCTSPlayer
    from policy_value_net_mxnet import PolicyValueNet
    import AlphaPig as gomoku_zm
    from alphapig_b200.params import init_params
    n, width, height = 5, 15, 15
    arg, aux = init_params('resnet', width, height, 2, 128, seed=1)
    blob = pickle.dumps((dict(arg), dict(aux)), protocol=2)
    policy_param = pickle.loads(blob)
    board = Board(width=width, height=height, n_in_row=n)
    game = Game(board)
    best_policy = PolicyValueNet(board_width=width, board_height=height, batch_size=512, n_blocks=2, model_params=policy_param)
    mcts_player = MCTSPlayer(best_policy.policy_value_fn, c_puct=5, n_playout=60)

    class Scripted(object):  # human_play_mxnet.py:17-44 with the keyboard replaced by "first free cell of row 7, then any"
        def set_player_ind(self, p):
            self.player = p

        def get_action(self, board):
            for mv in list(range(7 * 15, 8 * 15)) + list(range(225)):
                if mv in board.availables:
                    return board.location_to_move(board.move_to_location(mv))

    winner = game.start_play(Scripted(), mcts_player, start_player=1, is_shown=0)
    assert winner in (1, 2, -1) and len(board.states) >= 9
    # ChessClient.__init__ (complex_ == 's') + play_one_piece
    simple_arg, simple_aux = init_params('simple', width, height, seed=2)
    pvn = gomoku_zm.policy_value_net_mxnet_simple.PolicyValueNet(height, width, batch_size=16,
                                                                 model_params=(dict(simple_arg), dict(simple_aux)))
    player = gomoku_zm.mcts_alphaZero.MCTSPlayer(pvn.policy_value_fn, c_puct=3, n_playout=80)
    cb = gomoku_zm.game.Board(width=width, height=height, n_in_row=5)
    cb.init_board(0)
    g2 = gomoku_zm.game.Game(cb)
    p1, p2 = cb.players
    player.set_player_ind(p1)
    for opp in (112, 113, 98):
        if opp not in cb.states:           # ChessClient.py:217 (`has_key`)
            cb.do_move(opp)
        move = player.get_action(cb)
        assert move in cb.availables
        cb.do_move(move)
    assert len(cb.states) == 6 and g2.board is cb
    with pytest.raises(ValueError):        # list.remove semantics of an occupied cell
        cb.do_move(112)
    # a board of the wrong size for the net is refused instead of silently mis-evaluated
    small = gomoku_zm.game.Board(width=8, height=8, n_in_row=5)
    small.init_board()
    with pytest.raises(ValueError):
        player.get_action(small)
