"""Oracle restatement of the reference game drivers.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

* ``start_play``       : reference ``game.py:204-230`` (= ``game_ai.py:42-68``)
* ``start_self_play``  : reference ``game_ai.py:70-139`` (real MCTS self-play)
* ``sgf_self_play``    : reference ``game.py:233-304`` (SGF replay, no search)
"""
import random

import numpy as np

# game_ai.py:76-77: the hard-coded 15-wide opening tables
_BLANK_MOVES = [r * 15 + c for r in range(7) for c in range(9)]
_WHITE_MOVES = range(0, 103)


def start_play(board, player1, player2, start_player=0):
    if start_player not in (0, 1):
        raise Exception('start_player should be either 0 (player1 first) '
                        'or 1 (player2 first)')
    board.init_board(start_player)
    p1, p2 = board.players
    player1.set_player_ind(p1)
    player2.set_player_ind(p2)
    players = {p1: player1, p2: player2}
    while True:
        move = players[board.get_current_player()].get_action(board)
        board.do_move(move)
        end, winner = board.game_end()
        if end:
            return winner


def _one_hot_pi(size, move):
    # game_ai.py:86-88 / game.py:249-251
    probs = [0.000001 for _ in range(size)]
    probs[move] = 0.99999
    return np.asarray(probs)


def _finish(current_players, winner):
    # game_ai.py:127-131
    z = np.zeros(len(current_players))
    if winner != -1:
        z[np.array(current_players) == winner] = 1.0
        z[np.array(current_players) != winner] = -1.0
    return z


def start_self_play(board, player, temp=1e-3):
    """Returns (winner, [(state, pi, z), ...]) -- game_ai.py:70-139."""
    size = board.width * board.height
    board.init_board()
    states, pis, cur = [], [], []
    if random.random() < 0.09:
        while True:
            mb = random.choice(_BLANK_MOVES)
            mw = random.choice(_WHITE_MOVES)
            if mb != mw:
                break
        for m in (mb, mw):
            states.append(board.current_state())
            pis.append(_one_hot_pi(size, m))
            cur.append(board.current_player)
            board.do_move(m)
    while True:
        move, move_probs = player.get_action(board, temp=temp, return_prob=1)
        states.append(board.current_state())
        pis.append(move_probs)
        cur.append(board.current_player)
        board.do_move(move)
        end, winner = board.game_end()
        if end:
            z = _finish(cur, winner)
            player.reset_player()
            return winner, list(zip(states, pis, z))


def sgf_self_play(board, player, record):
    """``record`` = {'winner': w, 'seq_num_list': [...]} as returned by the
    reference ``utils/sgf_dataIter.get_data_from_files`` (``:45-66``).
    Returns (warning, winner, data) -- game.py:233-304."""
    size = board.width * board.height
    seq = record['seq_num_list']
    board.init_board()
    states, pis, cur = [], [], []
    for idx, move in enumerate(seq):
        states.append(board.current_state())
        pis.append(_one_hot_pi(size, move))
        cur.append(board.current_player)
        try:
            board.do_move(move)
        except Exception:
            return 1, None, None
        if idx + 1 == len(seq):
            winner = record['winner']
            z = _finish(cur, winner)
            player.reset_player()
            return 0, winner, list(zip(states, pis, z))
