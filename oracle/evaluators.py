"""Deterministic injected evaluators for tree-parity tests.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Both return fp64 priors and
a Python-float value so NumPy-1.14 vs NumPy-2 promotion differences in
``TreeNode.get_value`` (``mcts_alphaZero.py:78``) cannot arise.

They accept anything with the reference ``Board`` protocol (``states``,
``availables``, ``current_player``): the reference ``game.Board``, ``OBoard``
or the product shim.
"""
import zlib

import numpy as np


def e1_uniform(board):
    """= ``mcts_pure.policy_value_fn`` (``mcts_pure.py:20-25``) with value 0.0."""
    n = len(board.availables)
    return zip(board.availables, np.ones(n) / n), 0.0


def position_key(board):
    items = sorted(board.states.items())
    buf = np.asarray([board.current_player] + [v for kv in items for v in kv], dtype=np.int32)
    return zlib.crc32(buf.tobytes())


def e2_hash(board):
    """Peaky pseudo-random priors (Dirichlet 0.3) and a value in (-1, 1), both a
    pure function of (stones, side to move)."""
    rs = np.random.RandomState(position_key(board))
    n = len(board.availables)
    pri = rs.dirichlet(0.3 * np.ones(n)) if n > 1 else np.ones(n)
    val = float(rs.uniform(-0.9, 0.9))
    return zip(board.availables, pri), val


def e3_quantised(board):
    """Priors are multiples of 1/64 and values multiples of 1/4 so many PUCT
    scores tie exactly -- exercises the first-max tie-break."""
    rs = np.random.RandomState(position_key(board) ^ 0x5bd1e995)
    n = len(board.availables)
    pri = rs.randint(0, 4, size=n).astype(np.float64) / 64.0
    val = float(rs.randint(-3, 4)) / 4.0
    return zip(board.availables, pri), val


EVALUATORS = {"e1": e1_uniform, "e2": e2_hash, "e3": e3_quantised}


def e4_peaky(board):
    """Very peaky priors (Dirichlet with concentration 0.003 per move, i.e. total concentration < 1 over ~200 legal
    moves: nearly all mass on one move; 0.03 still spreads over ~6 moves and searches only ~3 plies deep) and small
    values, a pure function of the position: the search follows one line 7-10 plies into the tree at 400 playouts and
    most of a re-rooted subtree survives every ply (root N grows past 1000) - the regime of a trained net, which flat
    random-weight priors never reach (stresses the re-root compaction, deep backups and pool growth)."""
    rs = np.random.RandomState(position_key(board) ^ 0x1b873593)
    n = len(board.availables)
    pri = rs.dirichlet(0.003 * np.ones(n)) if n > 1 else np.ones(n)
    val = float(rs.uniform(-0.3, 0.3))
    return zip(board.availables, pri), val


EVALUATORS["e4"] = e4_peaky
