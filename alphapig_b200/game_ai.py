"""Drop-in for the reference ``game_ai`` module: ``Game_AI`` (game_ai.py:11-139), the real MCTS
self-play driver.  Same signatures and return shapes; see ``alphapig_b200.selfplay`` for the
batched (thousands of games per GPU) form of the same loop."""
from __future__ import print_function

import random

import numpy as np

from .game import Board, Game  # noqa: F401  (reference module exposes Board too)

# the hard-coded 15-wide opening tables of game_ai.py:76-77
_BLACK_OPENINGS = [r * 15 + c for r in range(7) for c in range(9)]
_WHITE_OPENINGS = range(0, 103)


class Game_AI(Game):
    """game server; ``start_play`` and ``graphic`` are shared with ``Game`` (identical in the reference)."""

    def _record(self, states, pis, players, pi, move, is_shown):
        b = self.board
        states.append(b.current_state())
        pis.append(pi)
        players.append(b.current_player)
        b.do_move(move)
        if is_shown:
            self.graphic(b, *b.players)

    def start_self_play(self, player, is_shown=0, temp=1e-3):
        """One self-play game with tree reuse; returns (winner, zip(states, mcts_probs, winners_z))
        exactly as game_ai.py:70-139, including the 9% forced random two-ply opening (:77-111)."""
        b = self.board
        b.init_board()
        states, pis, players = [], [], []
        if random.random() < 0.09:
            while True:
                first = random.choice(_BLACK_OPENINGS)
                second = random.choice(_WHITE_OPENINGS)
                if first != second:
                    break
            for mv in (first, second):
                pi = np.full(self._boardSize, 0.000001)
                pi[mv] = 0.99999
                self._record(states, pis, players, pi, mv, is_shown)
        while True:
            move, move_probs = player.get_action(b, temp=temp, return_prob=1)
            self._record(states, pis, players, move_probs, move, is_shown)
            end, winner = b.game_end()
            if not end:
                continue
            z = np.zeros(len(players))
            if winner != -1:
                mine = np.array(players) == winner
                z[mine] = 1.0
                z[~mine] = -1.0
            player.reset_player()
            if is_shown:
                print("Game end. Winner is player: %s" % winner if winner != -1 else "Game end. Tie")
            return winner, zip(states, pis, z)
