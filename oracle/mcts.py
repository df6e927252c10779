"""Oracle restatement of the reference tree searches.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

* AlphaZero search: reference ``mcts_alphaZero.py:19-221``.
* Pure random-rollout search: reference ``mcts_pure.py:13-206``.

Numerics follow the reference exactly: Q/u in Python floats (fp64), the PUCT
term evaluated left to right ``((c*P)*sqrt(Np))/(1+N)`` (``mcts_alphaZero.py:78-79``),
the running mean as ``Q += 1.0*(v-Q)/N`` (``:59``), children kept in insertion
order (= ascending legal move) and ``max`` returning the first maximum (``:48``).
"""
import copy
import math

import numpy as np


def softmax(x):
    # mcts_alphaZero.py:13-16
    p = np.exp(x - np.max(x))
    p /= np.sum(p)
    return p


class ONode(object):
    __slots__ = ("parent", "children", "N", "Q", "P")

    def __init__(self, parent, prior):
        self.parent = parent
        self.children = {}
        self.N = 0
        self.Q = 0
        self.P = prior

    # mcts_alphaZero.py:34-41
    def expand(self, action_priors):
        ch = self.children
        for a, p in action_priors:
            if a not in ch:
                ch[a] = ONode(self, p)

    # mcts_alphaZero.py:69-80 (np.sqrt(int) is fp64 sqrt; math.sqrt is the same IEEE op)
    def value(self, c_puct):
        u = (c_puct * self.P * math.sqrt(self.parent.N) / (1 + self.N))
        return self.Q + u

    # mcts_alphaZero.py:43-49: first maximum in insertion order
    def select(self, c_puct):
        best_a, best_n, best_v = None, None, None
        for a, n in self.children.items():
            v = n.value(c_puct)
            if best_v is None or v > best_v:
                best_a, best_n, best_v = a, n, v
        return best_a, best_n

    # mcts_alphaZero.py:51-59
    def update(self, v):
        self.N += 1
        self.Q += 1.0 * (v - self.Q) / self.N

    # mcts_alphaZero.py:61-67 (iterative form; every node's update is independent)
    def update_recursive(self, v):
        node = self
        while node is not None:
            node.update(v)
            v = -v
            node = node.parent

    def is_leaf(self):
        return not self.children


class OMCTS(object):
    """AlphaZero-style search (mcts_alphaZero.py:90-170)."""

    def __init__(self, policy_value_fn, c_puct=5, n_playout=10000):
        self.root = ONode(None, 1.0)
        self.policy = policy_value_fn
        self.c_puct = c_puct
        self.n_playout = n_playout

    # mcts_alphaZero.py:108-139
    def playout(self, state):
        node = self.root
        while not node.is_leaf():
            a, node = node.select(self.c_puct)
            state.do_move(a)
        action_probs, leaf_value = self.policy(state)
        end, winner = state.game_end()
        if not end:
            node.expand(action_probs)
        else:
            if winner == -1:
                leaf_value = 0.0
            else:
                leaf_value = 1.0 if winner == state.get_current_player() else -1.0
        node.update_recursive(-leaf_value)

    # mcts_alphaZero.py:141-157
    def get_move_probs(self, state, temp=1e-3):
        for _ in range(self.n_playout):
            self.playout(copy.deepcopy(state))
        acts = tuple(self.root.children.keys())
        visits = tuple(n.N for n in self.root.children.values())
        probs = softmax(1.0 / temp * np.log(np.array(visits) + 1e-10))
        return acts, probs

    # mcts_alphaZero.py:159-167
    def update_with_move(self, last_move):
        if last_move in self.root.children:
            self.root = self.root.children[last_move]
            self.root.parent = None
        else:
            self.root = ONode(None, 1.0)


class OMCTSPlayer(object):
    """mcts_alphaZero.py:173-221.  RNG = numpy legacy global, as the reference."""

    def __init__(self, policy_value_function, c_puct=5, n_playout=2000, is_selfplay=0):
        self.mcts = OMCTS(policy_value_function, c_puct, n_playout)
        self._is_selfplay = is_selfplay

    def set_player_ind(self, p):
        self.player = p

    def reset_player(self):
        self.mcts.update_with_move(-1)

    def get_action(self, board, temp=1e-3, return_prob=0):
        move_probs = np.zeros(board.width * board.height)
        if len(board.availables) > 0:
            acts, probs = self.mcts.get_move_probs(board, temp)
            move_probs[list(acts)] = probs
            if self._is_selfplay:
                move = np.random.choice(
                    acts, p=0.75 * probs + 0.25 * np.random.dirichlet(0.3 * np.ones(len(probs))))
                self.mcts.update_with_move(move)
            else:
                move = np.random.choice(acts, p=probs)
                self.mcts.update_with_move(-1)
            if return_prob:
                return move, move_probs
            return move
        print("WARNING: the board is full")


# ----------------------------------------------------------------------------
# mcts_pure
# ----------------------------------------------------------------------------

def pure_policy_value_fn(board):
    # mcts_pure.py:20-25
    n = len(board.availables)
    return zip(board.availables, np.ones(n) / n), 0


def rollout_policy_fn(board):
    # mcts_pure.py:13-17
    return zip(board.availables, np.random.rand(len(board.availables)))


class OPureMCTS(object):
    """mcts_pure.py:96-182.  ``rollout_fn(state) -> leaf value`` can be injected
    (tests pin the tree bookkeeping with a deterministic rollout result)."""

    def __init__(self, policy_value_fn=pure_policy_value_fn, c_puct=5, n_playout=10000,
                 rollout_fn=None):
        self.root = ONode(None, 1.0)
        self.policy = policy_value_fn
        self.c_puct = c_puct
        self.n_playout = n_playout
        self.rollout_fn = rollout_fn or self.evaluate_rollout

    # mcts_pure.py:114-136
    def playout(self, state):
        node = self.root
        while not node.is_leaf():
            a, node = node.select(self.c_puct)
            state.do_move(a)
        action_probs, _ = self.policy(state)
        end, _winner = state.game_end()
        if not end:
            node.expand(action_probs)
        leaf_value = self.rollout_fn(state)
        node.update_recursive(-leaf_value)

    # mcts_pure.py:138-157
    @staticmethod
    def evaluate_rollout(state, limit=1000):
        player = state.get_current_player()
        for _ in range(limit):
            end, winner = state.game_end()
            if end:
                break
            best, best_p = None, None
            for a, p in rollout_policy_fn(state):
                if best_p is None or p > best_p:
                    best, best_p = a, p
            state.do_move(best)
        else:
            print("WARNING: rollout reached move limit")
        if winner == -1:
            return 0
        return 1 if winner == player else -1

    # mcts_pure.py:159-169: first max by visit count
    def get_move(self, state):
        for _ in range(self.n_playout):
            self.playout(copy.deepcopy(state))
        best_a, best_n = None, None
        for a, n in self.root.children.items():
            if best_n is None or n.N > best_n:
                best_a, best_n = a, n.N
        return best_a

    def update_with_move(self, last_move):
        if last_move in self.root.children:
            self.root = self.root.children[last_move]
            self.root.parent = None
        else:
            self.root = ONode(None, 1.0)


class OPureMCTSPlayer(object):
    """mcts_pure.py:185-206."""

    def __init__(self, c_puct=5, n_playout=2000):
        self.mcts = OPureMCTS(pure_policy_value_fn, c_puct, n_playout)

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
