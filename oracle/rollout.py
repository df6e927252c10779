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


def line_windows(W, H, n):
    """all n-cell windows (right / up / up-right / up-left, game.py:141-156) as an (n_windows, n) index array"""
    import numpy as np
    wins = []
    for h in range(H):
        for w in range(W):
            m = h * W + w
            if w <= W - n:
                wins.append([m + k for k in range(n)])
            if h <= H - n:
                wins.append([m + k * W for k in range(n)])
            if w <= W - n and h <= H - n:
                wins.append([m + k * (W + 1) for k in range(n)])
            if w >= n - 1 and h <= H - n:
                wins.append([m + k * (W - 1) for k in range(n)])
    return np.array(wins)


def rollout_sample_numpy(board, n_samples, rs, chunk=20000):
    """``n_samples`` independent reference rollouts (mcts_pure.py:138-157) from a non-terminal ``board`` whose
    stone count already passes the ``n_in_row + 2`` early-out of game.py:134 whenever a line can exist,
    vectorised: the move order of a rollout is a uniformly random permutation of the empty cells (argsort of
    iid uniforms from ``rs``, a ``numpy.random.RandomState``); a window is completed by the side whose
    plies all have the window's parity, at its largest ply; the game ends at the earliest completion.
    Returns (values for the player to move at ``board``, plies played) -- the statistical pin of the device
    rollouts (tests/test_gpu_tree.py), itself checked against ``rollout_by_play`` in the CPU suite."""
    import numpy as np
    W, H, n = board.width, board.height, board.n_in_row
    S = W * H
    wins = line_windows(W, H, n)
    player = board.get_current_player()
    empties = np.array(board.availables)
    E = len(empties)
    base = np.zeros(S, np.int16)
    for m, p in board.states.items():
        base[m] = -2 if p == player else -1  # parity 0 = the side to move, parity 1 = its opponent
    vals, plies = [], []
    for c0 in range(0, n_samples, chunk):
        c = min(chunk, n_samples - c0)
        order = np.argsort(rs.random_sample((c, E)), axis=1)  # order[i, t] = index of the empty cell played at ply t
        rank = np.empty((c, E), np.int16)
        np.put_along_axis(rank, order, np.broadcast_to(np.arange(E, dtype=np.int16), (c, E)), axis=1)
        full = np.broadcast_to(base, (c, S)).copy()
        full[:, empties] = rank
        r = full[:, wins]                                      # (c, n_windows, n)
        par = r & 1
        same = (par == par[:, :, :1]).all(axis=2)
        t = np.where(same, r.max(axis=2), 30000).min(axis=1)   # ply index (0-based) that completes the first line
        tie = t >= 30000
        plies.append(np.where(tie, E, t + 1))
        vals.append(np.where(tie, 0, np.where(t % 2 == 0, 1, -1)))
    return np.concatenate(vals), np.concatenate(plies)


def philox4x32_10(counter, key, rounds=10):
    """Philox4x32 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) exactly as
    ``csrc/rollout.cu:philox4x32_10`` / ``csrc/tree.cu:philox_pick`` code it (multipliers, Weyl key bumps after the
    round); ``tests/test_cpu_host.py`` checks it against the Random123 known-answer vectors."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(rounds):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & 0xFFFFFFFF, lo1, (hi0 ^ c3 ^ k1) & 0xFFFFFFFF, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def device_move_order(board, seed, g, playout=0):
    """The move order the device's permutation rollout plays for game ``g`` from ``board`` (a 15-wide-or-narrower
    position): ``csrc/rollout.cu:perm_draws`` gives board slot (lane, r) the 24 high bits of word r of two
    Philox4x32-10 blocks (counter (playout, block, g, lane), key = the 64-bit seed); slot (lane, r) is board row
    lane // 2, column rank_slot((lane % 2) * 8 + r); the empty cells are played in ascending (draw, slot) order."""
    W = board.width
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    draw = {}
    for lane in range(32):
        words = philox4x32_10((playout, 0, g, lane), key) + philox4x32_10((playout, 1, g, lane), key)
        for r in range(8):
            q = (lane & 1) * 8 + r
            col = ((q & 3) << 2) | (q >> 2)
            draw[(lane >> 1, col)] = (min(words[r] >> 8, 0xFFFFFE), lane * 8 + r)
    return sorted(board.availables, key=lambda m: draw[(m // W, m % W)])


def device_perm_rollout(board, seed, g, playout=0):
    """Exact host model of ``ap_rollout_eval`` (impl 0) for game ``g``: (value, plies), RNG included."""
    end, winner = board.game_end()
    if end:
        return (0 if winner == -1 else (1 if winner == board.get_current_player() else -1)), 0
    return rollout_by_play(board, device_move_order(board, seed, g, playout))


def outcomes_from_ranks_numpy(full, n_empty, W, H, n):
    """Vectorised ``rollout_by_play`` for many positions at once.  full: int16 [G][S], -2 / -1 for the stones of the
    side to move / its opponent, else the ply index (rank) at which the empty cell is played.  Returns (values for the
    side to move, plies).  Same window rule as ``rollout_sample_numpy``."""
    import numpy as np
    wins = line_windows(W, H, n)
    r = full[:, wins]
    par = r & 1
    same = (par == par[:, :, :1]).all(axis=2)
    t = np.where(same, r.max(axis=2), 30000).min(axis=1)
    tie = t >= 30000
    return np.where(tie, 0, np.where(t % 2 == 0, 1, -1)), np.where(tie, n_empty, t + 1)
