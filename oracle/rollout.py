"""Oracle restatement of the *permutation* form of the reference rollout
(reference ``mcts_pure.py:13-17`` rollout_policy_fn + ``:138-157`` _evaluate_rollout).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pure Python.

The reference plays ``argmax(np.random.rand(len(availables)))`` every ply: a uniformly random legal move,
i.e. sampling the empty cells without replacement = one uniformly random permutation ``order`` of the
cells that are empty at the leaf, ply i taking the colour of the player to move at ply i.
``rollout_by_play`` is that loop verbatim on an ``OBoard`` for a given ``order``;
``rollout_by_descent`` is what ``csrc/rollout.cu:rollout_eval_perm`` computes from the same order: the
largest t such that the first t plies complete no line, found by descending the bits of t (the predicate
is monotone in t), with the reference's ``len(moved) < n_in_row + 2`` early-out (``game.py:134``) kept.
Both return ``(value for the player to move at the leaf, plies played)``.
"""


def rollout_by_play(board, order):
    """mcts_pure.py:138-157 with the random choices replaced by the given permutation of the empty cells."""
    state = board.clone()
    player = state.get_current_player()
    plies = 0
    order = list(order)
    while True:
        end, winner = state.game_end()
        if end:
            break
        state.do_move(int(order[plies]))
        plies += 1
    if winner == -1:
        return 0, plies
    return (1 if winner == player else -1), plies


def _has_line(stones, W, H, n):
    """any n-in-a-row (right / up / up-right / up-left, game.py:141-156) in {move: player}"""
    for m, p in stones.items():
        h, w = divmod(m, W)
        if w <= W - n and all(stones.get(m + k) == p for k in range(n)):
            return True
        if h <= H - n and all(stones.get(m + k * W) == p for k in range(n)):
            return True
        if w <= W - n and h <= H - n and all(stones.get(m + k * (W + 1)) == p for k in range(n)):
            return True
        if w >= n - 1 and h <= H - n and all(stones.get(m + k * (W - 1)) == p for k in range(n)):
            return True
    return False


def rollout_by_descent(board, order):
    W, H, n = board.width, board.height, board.n_in_row
    player = board.get_current_player()
    end, winner = board.game_end()
    if end:
        return (0 if winner == -1 else (1 if winner == player else -1)), 0
    order = [int(m) for m in order]
    E = len(order)
    assert sorted(order) == sorted(board.availables)
    rank = {m: i for i, m in enumerate(order)}
    need = n + 2 - len(board.states)
    colour = (player, 3 - player)

    def line_after(t):
        if min(t, E) < need:
            return False
        stones = dict(board.states)
        for m, r in rank.items():
            if r < t:
                stones[m] = colour[r & 1]
        return _has_line(stones, W, H, n)

    t = 0
    for bit in range(7, -1, -1):
        tt = t | (1 << bit)
        if not line_after(tt):
            t = tt
    if t >= E:
        return 0, E
    return (-1 if (t & 1) else 1), t + 1
