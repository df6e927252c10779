"""Batched self-play: the loop of reference ``Game_AI.start_self_play`` (game_ai.py:113-139) +
``MCTSPlayer.get_action`` (mcts_alphaZero.py:187-218) for thousands of concurrent games on one GPU.

Per ply: ``ap_search_run`` (n_playout lock-steps of select -> net -> expand/backup, all on device),
one D2H of the root visit counts, host-side pi / Dirichlet / sampling with numpy exactly as the
reference does per game, then re-root + do_move + status on device.  Finished games are emitted as
(states, pis, z) records and their slots immediately start a new game.
"""
import numpy as np


def visit_softmax(visits, counts, temp):
    """Row-wise softmax(1/temp * log(visits + 1e-10)) over the first counts[g] entries (fp64)."""
    G, S = visits.shape
    mask = np.arange(S)[None, :] < counts[:, None]
    x = (1.0 / temp) * np.log(visits.astype(np.float64) + 1e-10)
    x = np.where(mask, x, -np.inf)
    x = x - x.max(axis=1, keepdims=True)
    p = np.where(mask, np.exp(x), 0.0)
    return p / p.sum(axis=1, keepdims=True)


class BatchedSelfPlay(object):
    def __init__(self, net, n_games, n_playout=400, c_puct=5, temp=1.0, n_in_row=5, seed=0,
                 noise_eps=0.25, dirichlet_alpha=0.3, node_capacity=0, record_states=True, tag=0, device_pick=False,
                 forced_opening_prob=0.09, device_records=False):
        """device_pick: sample the moves on the device (``ap_selfplay_pick``: pi, Dirichlet noise and the inverse-cdf
        draw per game from Philox streams instead of NumPy's global generator) and run the NEXT ply's search in a
        background thread while this ply's records are assembled on the host - one group of games then keeps the GPU
        busy (``PipelinedSelfPlay`` needs two half-size groups for that).

        forced_opening_prob: the reference starts 9 % of its self-play games with a forced random two-ply opening drawn
        from hard-coded 15-wide tables and records both plies with pi = 0.99999 at the move / 1e-6 elsewhere
        (game_ai.py:76-111); done here per restarting slot on 15-wide boards (the tables index a 15-wide board: the
        reference itself crashes with them on 8x8).  0 disables it.

        device_records (needs device_pick): the (state, pi, z) records never visit the host - every pick appends its
        ply to the game's trajectory in HBM (``ap_traj_create``), finished games move to the device OUTBOX with z
        filled in, and ``take_outbox()`` hands the packed records over as a device tensor (for ``ReplayBuffer.
        extend_packed`` / an NCCL all-gather).  ``step()`` then returns ``(winner, None, None, None)`` per finished game.
        ``boundary_hook(self)``, if set, runs once per ply at the point where no search is in flight (after the moves
        were played, before the next search starts): the place to drain the outbox or swap in new weights."""
        self.net = net
        self.device_pick = device_pick
        self._seed = int(seed)
        self._pool = None
        self._fut = None           # device_pick: the search of the next ply, running in the background thread
        self._search_done = False  # ... or already joined by drain()
        self.G = n_games
        self.n_playout = n_playout
        self.temp = temp
        self.eps = noise_eps
        self.alpha = dirichlet_alpha
        self.record_states = record_states
        self.rs = np.random.RandomState(seed)
        self.eng = net.search_engine(n_in_row=n_in_row, c_puct=c_puct, n_playout=n_playout, n_games=n_games,
                                     node_capacity=node_capacity, tag=tag)
        self.S = self.eng.S
        self.device_records = bool(device_records)
        self.boundary_hook = None
        if self.device_records:
            if not device_pick:
                raise ValueError("device_records needs device_pick=True (the pick kernel writes the records)")
            self.record_states = False
            self.eng.traj_create()
        self.opening_prob = float(forced_opening_prob) if self.eng.width == 15 and self.S >= 103 else 0.0
        self.forced_openings = 0
        self._reset_history()
        self._restart(np.arange(self.G, dtype=np.int32))
        self.finished_games = 0
        self.plies = 0
        self.host_seconds = 0.0  # time spent outside ap_search_run (sampling, recording, re-rooting)

    def _reset_history(self):
        # per-ply records for ALL games (bit-packed state, pi, player to move); a finished game gathers its own rows
        self._ply = 0
        self._rec = {}                                  # ply -> (feats uint8[G][sb], pi float32[G][S], players int8[G])
        self._start = np.zeros(self.G, np.int64)        # first ply of the current game of every slot
        self._prefix = {}                               # slot -> [(feat, pi, player)] forced-opening records of its game

    # the hard-coded opening tables of game_ai.py:76-77 (15-wide board indices)
    _BLACK_OPENINGS = np.array([r * 15 + c for r in range(7) for c in range(9)], np.int32)
    _WHITE_OPENINGS = np.arange(0, 103, dtype=np.int32)

    def _restart(self, ids):
        """``init_board()`` + fresh root for the slots ``ids`` (a new game each), with the reference's forced random
        two-ply opening for a ``forced_opening_prob`` share of them (game_ai.py:78-111)."""
        eng = self.eng
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if not len(ids):
            return
        eng.boards_reset(ids)
        for g in ids:
            self._prefix.pop(int(g), None)
        if self.opening_prob > 0.0:
            forced = ids[self.rs.random_sample(len(ids)) < self.opening_prob]
            if len(forced):
                first = self.rs.choice(self._BLACK_OPENINGS, size=len(forced))
                second = self.rs.choice(self._WHITE_OPENINGS, size=len(forced))
                clash = first == second
                while clash.any():  # the reference redraws both moves until they differ (:79-84)
                    first[clash] = self.rs.choice(self._BLACK_OPENINGS, size=int(clash.sum()))
                    second[clash] = self.rs.choice(self._WHITE_OPENINGS, size=int(clash.sum()))
                    clash = first == second
                recs = [[] for _ in forced]
                for mv in (first, second):
                    if self.record_states:
                        feats = eng.boards_features_packed(forced)
                        _, meta = eng.boards_export(forced)
                        for k in range(len(forced)):
                            pi = np.full(self.S, 0.000001, np.float32)
                            pi[mv[k]] = 0.99999
                            recs[k].append((feats[k], pi, np.int8(meta[k, 0])))
                    if self.device_records:
                        eng.traj_append_forced(mv, forced)
                    eng.boards_do_move(mv.astype(np.int32), forced)
                if self.record_states:
                    for k, g in enumerate(forced):
                        self._prefix[int(g)] = recs[k]
                self.forced_openings += len(forced)
        eng.search_advance(np.full(len(ids), -1, np.int32), ids)

    def load_positions(self, cells, meta):
        """Start every slot from a given position (benchmark's synthetic positions)."""
        self.drain()
        self._search_done = False
        self._reset_history()
        if self.device_records:
            self.eng.traj_discard()
        self.eng.boards_import(cells, meta)
        self.eng.search_advance(-1)

    def _finish(self, done, winner):
        """records of the games that just ended (rows of the per-ply history), slots restarted by the caller"""
        out = []
        for g in done:
            t_first = int(self._start[g])
            if self.record_states and self._ply >= t_first:
                rows = list(self._prefix.get(int(g), ())) + [(self._rec[t][0][g], self._rec[t][1][g], self._rec[t][2][g])
                                                             for t in range(t_first, self._ply + 1)]
                players = np.array([r[2] for r in rows])
                z = np.zeros(len(players))
                if winner[g] != -1:
                    z[players == winner[g]] = 1.0
                    z[players != winner[g]] = -1.0
                out.append((int(winner[g]), np.stack([r[0] for r in rows]),
                            np.stack([r[1] for r in rows]).astype(np.float64), z))
            else:
                out.append((int(winner[g]), None, None, None))
            self._start[g] = self._ply + 1
        return out

    def _search(self):
        """One move search for every game.  A library-chosen node capacity grows on demand; if the device cannot hold
        the grown pools (AP_ERR_POOL_EXHAUSTED before any playout ran) the retained subtrees are dropped for this ply
        - the searches then start from fresh roots, as after ``reset_player`` - instead of aborting the batch."""
        from ._lib import AP_ERR_POOL_EXHAUSTED, EngineError
        try:
            self.eng.search_run(self.n_playout)
        except EngineError as err:
            if err.code != AP_ERR_POOL_EXHAUSTED:
                raise
            self.tree_drops = getattr(self, "tree_drops", 0) + 1
            self.eng.search_advance(-1)
            self.eng.search_run(self.n_playout)

    def _step_device_pick(self):
        import time
        from concurrent.futures import ThreadPoolExecutor
        eng = self.eng
        if self._fut is not None:
            self._fut.result()          # the search of this ply, started at the end of the previous step
            self._fut = None
        elif not self._search_done:
            self._search()
        self._search_done = False
        t0 = time.perf_counter()
        moves, pi = eng.selfplay_pick(self.temp, self.eps, self.alpha, seed=self._seed, ply=self._ply)
        if self.record_states:
            feats = eng.boards_features_packed()
            _, meta = eng.boards_export()
        eng.search_advance(moves)
        eng.boards_do_move(moves)
        end, winner = eng.boards_status()
        done = np.nonzero(end)[0].astype(np.int32)
        if self.record_states:
            self._rec[self._ply] = (feats, pi, meta[:, 0].astype(np.int8))
        out = self._finish(done, winner)
        if self.device_records and len(done):
            eng.traj_finish(done, winner[done])
        self._restart(done)
        if self.boundary_hook is not None:
            self.boundary_hook(self)
        # the device is ready for the next ply: search it while the host assembles the records of this one
        if self._pool is None:
            self._pool = ThreadPoolExecutor(1)
        self._fut = self._pool.submit(self._search)
        self.plies += self.G
        self.finished_games += len(done)
        self._ply += 1
        if self.record_states and self._rec:
            oldest = int(self._start.min())
            for t in [t for t in self._rec if t < oldest]:
                del self._rec[t]
        self.host_seconds += time.perf_counter() - t0
        return out

    def take_outbox(self):
        """device_records: the finished games' packed records as a uint8 CUDA tensor [n][record width] (a copy; the
        outbox is emptied).  Only when no search is in flight: inside ``boundary_hook`` or after ``drain()``."""
        import torch
        from .nets import _DevView
        ptr, n, rw = self.eng.traj_outbox()
        dev = torch.device("cuda", self.net._device)
        if n == 0:
            return torch.empty((0, rw), dtype=torch.uint8, device=dev)
        view = torch.as_tensor(_DevView(ptr, n * rw, "|u1"), device=dev).view(n, rw)
        out = view.clone()
        torch.cuda.current_stream(dev).synchronize()
        self.eng.traj_outbox_clear()
        return out

    def drain(self):
        """Wait for the search in flight (device_pick) before touching the engine from outside."""
        if self._fut is not None:
            self._fut.result()
            self._fut = None
            self._search_done = True

    def step(self):
        """One ply for every game.  Returns the list of finished-game records
        [(winner, states uint8[n][ceil(9S/8)] bit-packed, pis float64[n][S], z float64[n])]."""
        import time
        if self.device_pick:
            return self._step_device_pick()
        eng, G, S = self.eng, self.G, self.S
        self._search()
        t0 = time.perf_counter()
        count, acts, visits, _, _ = eng.search_root()
        probs = visit_softmax(visits, count, self.temp)  # child order
        # sampling distribution: 0.75 p + 0.25 Dir(0.3)   (mcts_alphaZero.py:198-201)
        noise = self.rs.gamma(self.alpha, size=(G, S))
        mask = np.arange(S)[None, :] < count[:, None]
        noise = np.where(mask, noise, 0.0)
        noise /= noise.sum(axis=1, keepdims=True)
        samp = (1 - self.eps) * probs + self.eps * noise
        cdf = np.cumsum(samp, axis=1)
        u = self.rs.random_sample(G) * cdf[:, -1]
        idx = np.minimum((cdf <= u[:, None]).sum(axis=1), count - 1)
        moves = acts[np.arange(G), idx].astype(np.int32)
        if self.record_states:
            feats = eng.boards_features_packed()
            _, meta = eng.boards_export()
            pi = np.zeros((G, S), np.float32)
            rows = np.nonzero(mask)
            pi[rows[0], acts[rows]] = probs[rows]
            self._rec[self._ply] = (feats, pi, meta[:, 0].astype(np.int8))
        eng.search_advance(moves)
        eng.boards_do_move(moves)
        end, winner = eng.boards_status()
        self.plies += G
        done = np.nonzero(end)[0].astype(np.int32)
        out = self._finish(done, winner)
        self._restart(done)
        self.finished_games += len(done)
        self._ply += 1
        # plies older than every live game's start are no longer needed
        if self.record_states and self._rec:
            oldest = int(self._start.min())
            for t in [t for t in self._rec if t < oldest]:
                del self._rec[t]
        self.host_seconds += time.perf_counter() - t0
        return out


class PipelinedSelfPlay(object):
    """``n_groups`` independent ``BatchedSelfPlay`` groups (own engine, own CUDA stream), each stepped by its own
    host thread and started half a period apart: while one group is in its host phase (visit-count softmax,
    Dirichlet sampling, recording, re-rooting: ~70 ms per 4096 games) the GPU runs the search kernels of the other.
    ``ctypes`` releases the GIL inside ``ap_search_run``, so plain threads are enough.  Games stay independent.

    ``step()`` returns the finished games of ONE group's next ply (``last_moves`` = plies it covered)."""

    def __init__(self, net, n_games, n_groups=2, seed=0, **kw):
        from concurrent.futures import ThreadPoolExecutor
        per = n_games // n_groups
        self.groups = [BatchedSelfPlay(net, per + (1 if i < n_games - per * n_groups else 0), seed=seed + i, tag=("pipe", i), **kw)
                       for i in range(n_groups)]
        self.G = sum(g.G for g in self.groups)
        self.pool = ThreadPoolExecutor(n_groups)
        self._fut = None
        self._next = 0
        self._backlog = []
        self.last_moves = 0

    def load_positions(self, cells, meta):
        assert self._fut is None, "load_positions before the first step"
        k = 0
        for g in self.groups:
            g.load_positions(cells[k:k + g.G], meta[k:k + g.G])
            k += g.G

    def _start(self):
        import time
        t0 = time.perf_counter()
        self._backlog = self.groups[0].step()           # one ply of group 0 alone: measures the period
        period = time.perf_counter() - t0
        self._fut = [self.pool.submit(self.groups[0].step)]
        for g in self.groups[1:]:                        # the others start staggered by period / n_groups
            time.sleep(period / len(self.groups))
            self._fut.append(self.pool.submit(g.step))

    def step(self):
        if self._fut is None:
            self._start()
        i = self._next
        out = self._backlog + self._fut[i].result()
        self._backlog = []
        self._fut[i] = self.pool.submit(self.groups[i].step)
        self._next = (i + 1) % len(self.groups)
        self.last_moves = self.groups[i].G
        return out

    def drain(self):
        """Wait for the plies in flight (call before reading engine state or shutting down)."""
        out = []
        if self._fut is not None:
            for f in self._fut:
                out.extend(f.result())
            self._fut = None
        return out

    @property
    def finished_games(self):
        return sum(g.finished_games for g in self.groups)

    @property
    def host_seconds(self):
        return sum(g.host_seconds for g in self.groups)
