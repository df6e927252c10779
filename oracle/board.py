"""Oracle restatement of the reference ``Board`` (reference ``game.py:21-170``).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Pure Python + numpy.
"""
import numpy as np


class OBoard(object):
    """Same observable state as reference ``game.Board``.

    Attributes mirror ``game.py:24-44``: ``width``, ``height``, ``n_in_row``,
    ``players``, ``states`` (move -> player), ``availables`` (ascending list),
    ``history`` ([(move, player)]), ``current_player``, ``last_move``.
    """

    def __init__(self, width=8, height=8, n_in_row=5):
        self.width = int(width)
        self.height = int(height)
        self.n_in_row = int(n_in_row)
        self.players = [1, 2]
        self.states = {}

    # game.py:35-44
    def init_board(self, start_player=0):
        if self.width < self.n_in_row or self.height < self.n_in_row:
            raise Exception('board width and height can not be '
                            'less than {}'.format(self.n_in_row))
        self.current_player = self.players[start_player]
        self.availables = list(range(self.width * self.height))
        self.states = {}
        self.history = []
        self.last_move = -1

    def clone(self):
        b = OBoard(self.width, self.height, self.n_in_row)
        b.current_player = self.current_player
        b.availables = list(self.availables)
        b.states = dict(self.states)
        b.history = list(self.history)
        b.last_move = self.last_move
        return b

    def __deepcopy__(self, memo):
        return self.clone()

    # game.py:46-66
    def move_to_location(self, move):
        return [move // self.width, move % self.width]

    def location_to_move(self, location):
        if len(location) != 2:
            return -1
        move = location[0] * self.width + location[1]
        if move not in range(self.width * self.height):
            return -1
        return move

    # game.py:117-125 -- list.remove raises ValueError on an illegal move,
    # *after* states/history were already written (kept: observable state).
    def do_move(self, move):
        self.states[move] = self.current_player
        self.history.append((move, self.current_player))
        self.availables.remove(move)
        self.current_player = (self.players[0]
                               if self.current_player == self.players[1]
                               else self.players[1])
        self.last_move = move

    # game.py:127-158.  The reference iterates occupied cells in CPython set
    # order; under legal play only one colour can own a line so the order is
    # unobservable.  We iterate ascending.
    def has_a_winner(self):
        W, H, n, st = self.width, self.height, self.n_in_row, self.states
        if len(st) < n + 2:
            return False, -1
        for m in sorted(st):
            h, w = divmod(m, W)
            p = st[m]
            right = w <= W - n
            up = h <= H - n
            if right and all(st.get(m + k, -1) == p for k in range(n)):
                return True, p
            if up and all(st.get(m + k * W, -1) == p for k in range(n)):
                return True, p
            if right and up and all(st.get(m + k * (W + 1), -1) == p for k in range(n)):
                return True, p
            if w >= n - 1 and up and all(st.get(m + k * (W - 1), -1) == p for k in range(n)):
                return True, p
        return False, -1

    # game.py:160-167
    def game_end(self):
        win, winner = self.has_a_winner()
        if win:
            return True, winner
        if not self.availables:
            return True, -1
        return False, -1

    def get_current_player(self):
        return self.current_player

    # game.py:68-94.  Planes 6-2i / 7-2i hold own / opponent stones with the
    # last i plies dropped (i=0..3); plane 8 is all-ones iff the stone count is
    # even; indices are [m // width, m % height]; the result is flipped on
    # axis 1.  float64, shape (9, width, height).
    def current_state(self):
        W, H = self.width, self.height
        sq = np.zeros((9, W, H))
        L = len(self.history)
        if L:
            for i in range(4):
                for (m, p) in self.history[:L - i]:
                    plane = (6 - 2 * i) if p == self.current_player else (7 - 2 * i)
                    sq[plane, m // W, m % H] = 1.0
                if L - i == 0:
                    break
        if len(self.states) % 2 == 0:
            sq[8, :, :] = 1.0
        return sq[:, ::-1, :]
