"""Drop-in for the reference ``game`` module: ``Board`` and ``Game``.

Same class names, attributes, signatures and error behaviour as reference
``game.py:21-304``; the hot functions (``has_a_winner`` / ``game_end`` /
``current_state``) run as sm_100a kernels through the C ABI.  The Python object
keeps the reference's observable attributes (``states``, ``availables``,
``history``, ``current_player``, ``last_move``) as a host mirror so callers
such as ``evaluate/ChessClient.py:189-235`` that read them keep working, and so
the object stays ``copy.deepcopy``-able (``mcts_alphaZero.py:148``).
"""
from __future__ import print_function

import numpy as np

from .engine import Engine
from .utils import sgf_dataIter

_SERVICES = {}


def _service(width, height, n_in_row):
    """One single-game engine per board geometry, shared by every Board object."""
    key = (width, height, n_in_row)
    eng = _SERVICES.get(key)
    if eng is None:
        eng = Engine(width=width, height=height, n_in_row=n_in_row, n_games=1, n_playout=1, node_capacity=4)
        _SERVICES[key] = eng
    return eng


def export_board_state(board):
    """Any object with the reference Board protocol -> (cells int8[S], meta int32[8]) in the C-ABI format."""
    S = board.width * board.height
    cells = np.zeros(S, np.int8)
    for m, p in board.states.items():
        if 0 <= m < S:
            cells[m] = p
    hist = [m for m, _ in board.history[-1:-5:-1]]
    hist += [-1] * (4 - len(hist))
    meta = np.array([board.current_player, board.last_move, len(board.states)] + hist +
                    [getattr(board, '_start', 0)], np.int32)
    return cells, meta


class Board(object):
    """board for the game (reference game.py:21-170)"""

    def __init__(self, **kwargs):
        self.width = int(kwargs.get('width', 8))
        self.height = int(kwargs.get('height', 8))
        self.states = {}
        self.n_in_row = int(kwargs.get('n_in_row', 5))
        self.players = [1, 2]

    def init_board(self, start_player=0):
        if self.width < self.n_in_row or self.height < self.n_in_row:
            raise Exception('board width and height can not be '
                            'less than {}'.format(self.n_in_row))
        self.current_player = self.players[start_player]
        self.availables = list(range(self.width * self.height))
        self.states = {}
        self.history = []
        self.last_move = -1
        self._start = start_player

    def __deepcopy__(self, memo):
        b = Board(width=self.width, height=self.height, n_in_row=self.n_in_row)
        if hasattr(self, 'availables'):
            b.current_player = self.current_player
            b.availables = list(self.availables)
            b.states = dict(self.states)
            b.history = list(self.history)
            b.last_move = self.last_move
            b._start = getattr(self, '_start', 0)
        return b

    def move_to_location(self, move):
        return list(divmod(move, self.width))  # [h, w], row 0 at the bottom (game.py:46-56)

    def location_to_move(self, location):
        if len(location) != 2:
            return -1
        move = location[0] * self.width + location[1]
        return move if move in range(self.width * self.height) else -1

    def do_move(self, move):
        # same statement order as game.py:117-125 (ValueError from list.remove on an illegal move)
        me = self.current_player
        self.states[move] = me
        self.history.append((move, me))
        self.availables.remove(move)
        self.current_player = self.players[1] if me == self.players[0] else self.players[0]
        self.last_move = move

    # ---- device side ------------------------------------------------------
    def export_state(self):
        """(cells int8[S], meta int32[8]) in the C-ABI board format."""
        return export_board_state(self)

    def _upload(self):
        eng = _service(self.width, self.height, self.n_in_row)
        cells, meta = self.export_state()
        eng.boards_import(cells[None], meta[None])
        return eng

    def current_state(self):
        """(9, width, height) float64, as game.py:68-94 (axis-1 flip included)."""
        eng = self._upload()
        return eng.boards_features()[0].astype(np.float64)

    def has_a_winner(self):
        eng = self._upload()
        end, winner = eng.boards_status()
        if end[0] and winner[0] != -1:
            return True, int(winner[0])
        return False, -1

    def game_end(self):
        """Check whether the game is ended or not (game.py:160-167)"""
        eng = self._upload()
        end, winner = eng.boards_status()
        return bool(end[0]), int(winner[0])

    def get_current_player(self):
        return self.current_player


class Game(object):
    """game server (reference game.py:173-304)"""

    def __init__(self, board, **kwargs):
        self.board = board
        self._boardSize = board.width * board.height
        # callable(file_name, sgf_home) -> {'winner', 'seq_num_list'}; the reference calls utils.sgf_dataIter (game.py:240)
        self.sgf_loader = kwargs.get('sgf_loader') or sgf_dataIter.get_data_from_files

    def graphic(self, board, player1, player2):
        """ASCII rendering, row 0 at the bottom (game.py:180-202)."""
        W, H = board.width, board.height
        glyph = {player1: 'X', player2: 'O'}
        out = ["Player %s %s" % (player1, "with X".rjust(3)), "Player %s %s" % (player2, "with O".rjust(3)), "",
               "".join("{0:8}".format(x) for x in range(W)) + "\r\n"]
        for i in reversed(range(H)):
            cells = (glyph.get(board.states.get(i * W + j, -1), '_').center(8) for j in range(W))
            out.append("{0:4d}".format(i) + "".join(cells) + "\r\n\r\n")
        print("\n".join(out))

    def start_play(self, player1, player2, start_player=0, is_shown=1):
        """Two-player game loop (game.py:204-230); returns the winner (1, 2 or -1)."""
        if start_player not in (0, 1):
            raise Exception('start_player should be either 0 (player1 first) '
                            'or 1 (player2 first)')
        board = self.board
        board.init_board(start_player)
        seats = dict(zip(board.players, (player1, player2)))
        for ind, pl in seats.items():
            pl.set_player_ind(ind)
        show = (lambda: self.graphic(board, player1.player, player2.player)) if is_shown else (lambda: None)
        show()
        end, winner = False, -1
        while not end:
            board.do_move(seats[board.get_current_player()].get_action(board))
            show()
            end, winner = board.game_end()
        if is_shown:
            print("Game end. Winner is %s" % seats[winner] if winner != -1 else "Game end. Tie")
        return winner

    def start_self_play(self, player, is_shown=0, temp=1e-3, sgf_home=None, file_name=None):
        """SGF replay for the supervised bootstrap (game.py:233-304): every recorded ply becomes a training
        sample whose pi is 0.99999 at the human move and 1e-6 elsewhere, z comes from the winner in the file
        name; returns (warning, winner, zip(states, mcts_probs, winners_z)), or (1, None, None) as soon as a
        recorded move is illegal.  MCTS is never consulted; ``player`` is only reset at the end."""
        record = self.sgf_loader(file_name, sgf_home)
        board = self.board
        board.init_board()
        samples = []  # (state, pi, player to move) before each recorded move
        for move in record['seq_num_list']:
            pi = np.full(self._boardSize, 0.000001)
            pi[move] = 0.99999
            samples.append((board.current_state(), pi, board.current_player))
            try:
                board.do_move(move)
            except Exception:
                return 1, None, None
            if is_shown:
                self.graphic(board, *board.players)
        if not samples:
            return None  # the reference's loop body never runs on an empty record and falls off the end
        winner = record['winner']
        movers = np.array([p for _, _, p in samples])
        z = np.zeros(len(samples))
        if winner != -1:
            z = np.where(movers == winner, 1.0, -1.0)
        player.reset_player()
        if is_shown:
            print("Game end. Winner is player: %s" % winner if winner != -1 else "Game end. Tie")
        return 0, winner, zip([s for s, _, _ in samples], [p for _, p, _ in samples], z)
