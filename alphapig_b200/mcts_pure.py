"""Drop-in for the reference ``mcts_pure`` module (mcts_pure.py:1-206).

``MCTSPlayer(c_puct, n_playout).get_action(board)`` runs the whole pure-MCTS move search --
n_playout x (select, uniform-prior expand, uniformly random rollout to the end of the game,
backup) -- in ONE kernel launch, one CTA per game (csrc/rollout.cu), and returns the most visited
root action (first maximum, mcts_pure.py:168-169).  The rollouts draw from device Philox streams seeded
from ``numpy.random`` (the reference draws ``np.random.rand`` per ply, mcts_pure.py:16; its MT19937
sequence is not reproduced): each rollout is one random permutation of the empty cells and a bit descent to
the first completed line -- the same distribution of results and game lengths as playing ply by ply, pinned
exactly with injected draws and statistically against a NumPy reference sample.
"""
import numpy as np

from .engine import Engine
from .game import export_board_state


def policy_value_fn(board):
    """uniform probabilities and 0 score (mcts_pure.py:20-25); kept for callers that import it."""
    n = len(board.availables)
    return zip(board.availables, np.ones(n) / n), 0


class MCTS(object):
    def __init__(self, policy_value_fn=policy_value_fn, c_puct=5, n_playout=10000):
        self._policy = policy_value_fn
        self._c_puct = c_puct
        self._n_playout = n_playout
        self._eng = None
        self._geom = None

    def _engine(self, board):
        geom = (board.width, board.height, board.n_in_row)
        if self._eng is None or self._geom != geom:
            self._eng = Engine(width=board.width, height=board.height, n_in_row=board.n_in_row, n_games=1,
                               c_puct=self._c_puct, n_playout=self._n_playout,
                               node_capacity=self._n_playout * board.width * board.height + 2)
            self._geom = geom
        return self._eng

    def get_move(self, state):
        """Runs all playouts and returns the most visited action (mcts_pure.py:159-169)."""
        eng = self._engine(state)
        cells, meta = export_board_state(state)
        eng.boards_import(cells[None], meta[None])
        seed = int(np.random.randint(0, 2 ** 31 - 1))
        return int(eng.pure_run(self._n_playout, seed=seed, rollout_mode=0)[0])

    def update_with_move(self, last_move):
        # the fused kernel always starts from a fresh root, which is all the reference player ever
        # asks for (get_action calls update_with_move(-1) after every move, mcts_pure.py:199-200)
        pass

    def __str__(self):
        return "MCTS"


class MCTSPlayer(object):
    """AI player based on pure MCTS (mcts_pure.py:185-206)"""

    def __init__(self, c_puct=5, n_playout=2000):
        self.mcts = MCTS(policy_value_fn, c_puct, n_playout)

    def set_player_ind(self, p):
        self.player = p

    def reset_player(self):
        self.mcts.update_with_move(-1)

    def get_action(self, board):
        if len(board.availables) > 0:
            move = self.mcts.get_move(board)
            self.mcts.update_with_move(-1)
            return move
        print("WARNING: the board is full")

    def __str__(self):
        return "MCTS {}".format(self.player)
