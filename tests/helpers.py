"""Shared helpers for the GPU parity tests (oracle side + engine driving)."""
import numpy as np

from oracle.board import OBoard


def oboard_from(W, H, n, moves, start_player=0):
    b = OBoard(W, H, n)
    b.init_board(start_player)
    for m in moves:
        b.do_move(int(m))
    return b


def export_oboard(b):
    """OBoard -> (cells, meta) in the C-ABI board format."""
    S = b.width * b.height
    cells = np.zeros(S, np.int8)
    for m, p in b.states.items():
        cells[m] = p
    hist = [m for m, _ in b.history[-1:-5:-1]]
    hist += [-1] * (4 - len(hist))
    meta = np.array([b.current_player, b.last_move, len(b.states)] + hist + [0], np.int32)
    return cells, meta


def synth_position(W, H, n, seed, max_pairs=31):
    """SURVEY 8(d) synthetic positions (same recipe as tests/golden/make_golden.py, on the oracle board)."""
    rs = np.random.RandomState(seed)
    while True:
        k = 2 * rs.randint(0, max_pairs)
        b = OBoard(W, H, n)
        b.init_board(0)
        mv, ok = [], True
        for _ in range(k):
            m = int(b.availables[rs.randint(len(b.availables))])
            b.do_move(m)
            mv.append(m)
            if b.game_end()[0]:
                ok = False
                break
        if ok:
            return mv


def host_playouts(eng, roots, policy_fn, n_playout, active=None):
    """n_playout lock-steps of the host-evaluator path: select on device, evaluator on host
    (called with an OBoard of the leaf), expand/backup on device.  active: games the engine searches
    (ap_search_set_active); the others are left alone."""
    G, S = eng.G, eng.S
    for _ in range(n_playout):
        term, depth, path = eng.search_select()
        counts = np.zeros(G, np.int32)
        acts = np.zeros((G, S), np.int16)
        pri = np.zeros((G, S), np.float64)
        vals = np.zeros(G, np.float64)
        for g in range(G):
            if active is not None and not active[g]:
                continue
            leaf = roots[g].clone()
            for m in path[g, :depth[g]]:
                leaf.do_move(int(m))
            ap, v = policy_fn(leaf)
            ap = list(ap)
            counts[g] = len(ap)
            for k, (a, p) in enumerate(ap):
                acts[g, k] = a
                pri[g, k] = p
            vals[g] = v
            assert bool(term[g]) == leaf.game_end()[0]
        eng.search_expand_backup(counts, acts, pri, vals)


def engine_net_evaluator(eng):
    """``policy_value_fn`` for the oracle MCTS that returns the ENGINE's own fp32 priors / value of a position
    (``ap_net_forward`` on a second handle, batch 1), widened exactly to fp64 - SURVEY T4's contract for how evaluator
    output enters the tree.  With it the oracle must rebuild, bit for bit, the tree the all-device search built."""
    def fn(board):
        st = np.ascontiguousarray(board.current_state(), dtype=np.float32)[None]
        p, v = eng.net_forward(st)
        av = list(board.availables)
        return zip(av, p[0][av].astype(np.float64)), float(v[0, 0])
    return fn


def assert_root_equals_oracle(eng, g, omcts, what=""):
    """root children (acts, visits, Q) and the root's own N of game g == the oracle tree's, exactly"""
    count, acts, visits, q, rootn = eng.search_root(game_ids=[g], want_q=True)
    n = int(count[0])
    kids = omcts.root.children
    assert list(acts[0, :n]) == list(kids.keys()), (what, g, "acts")
    assert list(visits[0, :n]) == [nd.N for nd in kids.values()], (what, g, "visits")
    assert list(q[0, :n]) == [float(nd.Q) for nd in kids.values()], (what, g, "Q")
    assert int(rootn[0]) == omcts.root.N, (what, g, "root N")
