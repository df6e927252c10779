"""Drop-in for the reference ``mcts_alphaZero`` module (mcts_alphaZero.py:1-221).

``MCTS`` / ``MCTSPlayer`` keep the reference's constructor signatures and methods; the tree lives in
a flat SoA node pool in HBM and select / expand / backup / re-root are sm_100a kernels
(csrc/tree.cu).  Two evaluator paths:

* any Python ``policy_value_fn(board) -> (iterable[(action, prior)], value)`` -- the reference's
  plugin point (mcts_alphaZero.py:93-106,124).  Select runs on device, the callback is invoked on
  the host with a ``Board`` of the leaf, expand/backup run on device (fp64 priors and values).
* ``PolicyValueNet.policy_value_fn`` of the shim net: the whole search (select -> features -> net ->
  expand/backup, n_playout times) runs on device with no host round trip (``ap_search_run``).
"""
import copy

import numpy as np

from .engine import Engine
from .game import export_board_state


def softmax(x):
    probs = np.exp(x - np.max(x))
    probs /= np.sum(probs)
    return probs


class MCTS(object):
    """Monte Carlo Tree Search over a device-resident tree (one game)."""

    def __init__(self, policy_value_fn, c_puct=5, n_playout=10000, leaves_per_step=1):
        # leaves_per_step > 1 (device-net evaluator only): opt-in multi-leaf search with virtual loss
        # (ap_search_run_vl) - several playouts in flight per lock-step, ~k times lower move latency for ONE game,
        # visit counts no longer those of the reference's sequential search.  1 = the reference's search exactly.
        self._leaves_per_step = int(leaves_per_step)
        self._policy = policy_value_fn
        self._c_puct = c_puct
        self._n_playout = n_playout
        self._eng = None
        self._geom = None
        self._net = getattr(policy_value_fn, "__self__", None)
        if not getattr(self._net, "_is_alphapig_b200_net", False):
            self._net = None

    # -- engine management --------------------------------------------------
    def _engine(self, board):
        geom = (board.width, board.height, board.n_in_row)
        if self._eng is not None and self._geom == geom:
            return self._eng
        if self._eng is not None:
            self._release()
        if self._net is not None:
            if (board.width, board.height) != (self._net.board_width, self._net.board_height):
                raise ValueError("board is %dx%d but the policy-value net was built for %dx%d"
                                 % (board.width, board.height, self._net.board_width, self._net.board_height))
            # every MCTS object owns its tree, as in the reference (mcts_alphaZero.py:93-106): two players on the
            # same net with equal c_puct / n_playout must not share a device tree
            eng = self._net.search_engine(board.n_in_row, self._c_puct, self._n_playout, tag=("mcts", id(self)))
        else:
            eng = Engine(width=board.width, height=board.height, n_in_row=board.n_in_row, n_games=1,
                         c_puct=self._c_puct, n_playout=self._n_playout)
        self._eng, self._geom = eng, geom
        return eng

    def _release(self):
        eng, self._eng = self._eng, None
        if eng is None:
            return
        if self._net is not None:
            self._net.release_engine(eng)
        else:
            eng.close()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _load_root(self, eng, state):
        cells, meta = export_board_state(state)
        eng.boards_import(cells[None], meta[None])

    def _playout_host(self, eng, state):
        """One MCTS._playout (mcts_alphaZero.py:108-139) with the evaluator on the host."""
        term, depth, path = eng.search_select()
        leaf = copy.deepcopy(state)
        for m in path[0, :depth[0]]:
            leaf.do_move(int(m))
        action_probs, leaf_value = self._policy(leaf)
        S = eng.S
        acts = np.zeros((1, S), np.int16)
        pri = np.zeros((1, S), np.float64)
        k = 0
        seen = set()
        if not term[0]:
            for a, p in action_probs:
                a = int(a)
                if a in seen:  # TreeNode.expand skips actions already present (:39-41)
                    continue
                seen.add(a)
                acts[0, k] = a
                pri[0, k] = p
                k += 1
        val = np.asarray(leaf_value, dtype=np.float64).reshape(-1)[:1]  # MXNet-style shape-(1,) arrays allowed
        eng.search_expand_backup(np.array([k], np.int32), acts, pri, val)

    def get_move_probs(self, state, temp=1e-3):
        """Run all playouts and return the available actions and their probabilities (:141-157)."""
        eng = self._engine(state)
        self._load_root(eng, state)
        if self._net is not None:
            if self._leaves_per_step > 1:
                eng.search_run_vl(self._n_playout, self._leaves_per_step)
            else:
                eng.search_run(self._n_playout)
        else:
            for _ in range(self._n_playout):
                self._playout_host(eng, state)
        count, acts, visits, _, _ = eng.search_root()
        n = int(count[0])
        acts = tuple(int(a) for a in acts[0, :n])
        visits = tuple(int(v) for v in visits[0, :n])
        act_probs = softmax(1.0 / temp * np.log(np.array(visits) + 1e-10))
        return acts, act_probs

    def update_with_move(self, last_move):
        """Step forward in the tree keeping the subtree, or start a fresh root (:159-167)."""
        if self._eng is not None:
            self._eng.search_advance([int(last_move)])

    def __str__(self):
        return "MCTS"


class MCTSPlayer(object):
    """AI player based on MCTS (mcts_alphaZero.py:173-221)"""

    def __init__(self, policy_value_function, c_puct=5, n_playout=2000, is_selfplay=0, leaves_per_step=1):
        self.mcts = MCTS(policy_value_function, c_puct, n_playout, leaves_per_step=leaves_per_step)
        self._is_selfplay = is_selfplay

    def set_player_ind(self, p):
        self.player = p

    def reset_player(self):
        self.mcts.update_with_move(-1)

    def get_action(self, board, temp=1e-3, return_prob=0):
        move_probs = np.zeros(board.width * board.height)
        if len(board.availables) == 0:
            print("WARNING: the board is full")
            return None
        acts, probs = self.mcts.get_move_probs(board, temp)
        move_probs[list(acts)] = probs
        if self._is_selfplay:
            # Dirichlet noise perturbs only the sampling distribution (:198-201); pi stays un-noised
            noisy = 0.75 * probs + 0.25 * np.random.dirichlet(0.3 * np.ones(len(probs)))
            move = np.random.choice(acts, p=noisy)
            self.mcts.update_with_move(move)
        else:
            move = np.random.choice(acts, p=probs)
            self.mcts.update_with_move(-1)
        return (move, move_probs) if return_prob else move

    def __str__(self):
        return "MCTS {}".format(self.player)
