"""numpy face of the C ABI: one ``Engine`` = one handle = G games on one GPU."""
import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import EngineError


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ids(game_ids):
    if game_ids is None:
        return None
    return np.ascontiguousarray(game_ids, dtype=np.int32)


class Engine(object):
    """G concurrent Gomoku games + their MCTS trees, resident in HBM.

    Mirrors include/alphapig_b200.h one to one; arrays are numpy, errors raise
    ``EngineError`` (illegal moves raise ``ValueError`` as the reference's
    ``list.remove`` does, game.py:120; bad sizes raise ``Exception`` as
    game.py:36-38)."""

    def __init__(self, width=8, height=8, n_in_row=5, n_games=1, c_puct=5.0, n_playout=400,
                 node_capacity=0, device=0, high_priority=False):
        self.lib = L.load()
        self.width, self.height, self.n_in_row = int(width), int(height), int(n_in_row)
        self.S = self.width * self.height
        self.G = int(n_games)
        self.c_puct = float(c_puct)
        cfg = L.ApConfig(self.width, self.height, self.n_in_row, self.G, int(node_capacity), int(n_playout),
                         int(device), L.AP_FLAG_HIGH_PRIORITY_STREAM if high_priority else 0, self.c_puct)
        h = C.c_void_p()
        rc = self.lib.ap_engine_create(C.byref(cfg), C.byref(h))
        if rc == L.AP_ERR_BAD_ARG:
            if self.width < self.n_in_row or self.height < self.n_in_row:
                raise Exception('board width and height can not be less than {}'.format(self.n_in_row))
            raise EngineError(rc, "bad engine configuration")
        if rc != L.AP_OK:
            raise EngineError(rc, "ap_engine_create failed (no CUDA device / not sm_100a / out of memory); "
                                  "this library has no CPU fallback")
        self.h = h
        self.net_names = None

    # -- plumbing -----------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.ap_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == L.AP_OK:
            return
        msg = (self.lib.ap_last_error(self.h) or b"").decode()
        if rc == L.AP_ERR_ILLEGAL_MOVE:
            raise ValueError(msg)
        raise EngineError(rc, msg)

    def _n(self, ids):
        return self.G if ids is None else len(ids)

    def memory_bytes(self):
        v = C.c_uint64()
        self._check(self.lib.ap_engine_memory(self.h, C.byref(v)))
        return v.value

    def node_capacity(self):
        """tree nodes per game the pools currently hold (a library-chosen capacity grows on demand)"""
        v = C.c_int32()
        self._check(self.lib.ap_engine_node_capacity(self.h, C.byref(v)))
        return v.value

    def launch_count(self):
        v = C.c_uint64()
        self._check(self.lib.ap_launch_count(self.h, C.byref(v)))
        return v.value

    # -- boards -------------------------------------------------------------
    def boards_reset(self, game_ids=None, start_player=None):
        ids = _ids(game_ids)
        n = self._n(ids)
        sp = None
        if start_player is not None:
            sp = np.ascontiguousarray(np.broadcast_to(np.asarray(start_player, dtype=np.int32), (n,)))
            if np.any((sp != 0) & (sp != 1)):
                raise Exception('start_player should be either 0 (player1 first) or 1 (player2 first)')
        self._check(self.lib.ap_boards_reset(self.h, _ptr(ids), n, _ptr(sp)))

    def boards_do_move(self, moves, game_ids=None):
        ids = _ids(game_ids)
        mv = np.ascontiguousarray(moves, dtype=np.int32)
        n = self._n(ids)
        assert mv.shape == (n,)
        st = np.zeros(n, np.int32)
        self._check(self.lib.ap_boards_do_move(self.h, _ptr(ids), _ptr(mv), n, _ptr(st)))
        return st

    def boards_status(self, game_ids=None):
        ids = _ids(game_ids)
        n = self._n(ids)
        end = np.zeros(n, np.uint8)
        win = np.zeros(n, np.int8)
        self._check(self.lib.ap_boards_status(self.h, _ptr(ids), n, _ptr(end), _ptr(win)))
        return end.astype(bool), win.astype(np.int32)

    def boards_legal(self, game_ids=None):
        """-> bool [n][S]"""
        ids = _ids(game_ids)
        n = self._n(ids)
        m = np.zeros((n, 8), np.uint32)
        self._check(self.lib.ap_boards_legal(self.h, _ptr(ids), n, _ptr(m)))
        bits = np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")
        return bits[:, :self.S].astype(bool)

    def boards_features(self, game_ids=None):
        ids = _ids(game_ids)
        n = self._n(ids)
        out = np.zeros((n, 9, self.width, self.height), np.float32)
        self._check(self.lib.ap_boards_features(self.h, _ptr(ids), n, _ptr(out)))
        return out

    def boards_features_packed(self, game_ids=None):
        """np.packbits(Board.current_state()) per game: uint8 [n][ceil(9S/8)]"""
        ids = _ids(game_ids)
        n = self._n(ids)
        out = np.zeros((n, (9 * self.S + 7) // 8), np.uint8)
        self._check(self.lib.ap_boards_features_packed(self.h, _ptr(ids), n, _ptr(out)))
        return out

    def boards_export(self, game_ids=None):
        ids = _ids(game_ids)
        n = self._n(ids)
        cells = np.zeros((n, self.S), np.int8)
        meta = np.zeros((n, L.AP_META_INTS), np.int32)
        self._check(self.lib.ap_boards_export(self.h, _ptr(ids), n, _ptr(cells), _ptr(meta)))
        return cells, meta

    def boards_import(self, cells, meta, game_ids=None):
        ids = _ids(game_ids)
        n = self._n(ids)
        cells = np.ascontiguousarray(cells, dtype=np.int8).reshape(n, self.S)
        meta = np.ascontiguousarray(meta, dtype=np.int32).reshape(n, L.AP_META_INTS)
        self._check(self.lib.ap_boards_import(self.h, _ptr(ids), n, _ptr(cells), _ptr(meta)))

    # -- search -------------------------------------------------------------
    def search_select(self, want_path=True):
        term = np.zeros(self.G, np.uint8)
        depth = np.zeros(self.G, np.int32)
        path = np.zeros((self.G, self.S), np.int16) if want_path else None
        self._check(self.lib.ap_search_select(self.h, _ptr(term), _ptr(depth), _ptr(path)))
        return term.astype(bool), depth, path

    def search_leaf_export(self):
        cells = np.zeros((self.G, self.S), np.int8)
        meta = np.zeros((self.G, L.AP_META_INTS), np.int32)
        self._check(self.lib.ap_search_leaf_export(self.h, _ptr(cells), _ptr(meta)))
        return cells, meta

    def search_leaf_features(self):
        out = np.zeros((self.G, 9, self.width, self.height), np.float32)
        self._check(self.lib.ap_search_leaf_features(self.h, _ptr(out)))
        return out

    def search_expand_backup(self, counts, acts, priors, values):
        counts = np.ascontiguousarray(counts, dtype=np.int32).reshape(self.G)
        acts = np.ascontiguousarray(acts, dtype=np.int16).reshape(self.G, self.S)
        priors = np.ascontiguousarray(priors, dtype=np.float64).reshape(self.G, self.S)
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(self.G)
        self._check(self.lib.ap_search_expand_backup(self.h, _ptr(counts), _ptr(acts), _ptr(priors), _ptr(values)))

    def search_expand_backup_dense(self, priors, values):
        priors = np.ascontiguousarray(priors, dtype=np.float32).reshape(self.G, self.S)
        values = np.ascontiguousarray(values, dtype=np.float32).reshape(self.G)
        self._check(self.lib.ap_search_expand_backup_dense(self.h, _ptr(priors), _ptr(values)))

    def search_run(self, n_playout):
        self._check(self.lib.ap_search_run(self.h, int(n_playout)))

    def search_run_vl(self, n_playout, k):
        """Opt-in multi-leaf search: up to k playouts per game in flight per lock-step (virtual loss); k = 1 is
        ``search_run``'s tree bit for bit."""
        self._check(self.lib.ap_search_run_vl(self.h, int(n_playout), int(k)))

    def search_timing(self):
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.ap_search_timing(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def search_profile(self, enable=True):
        """Toggle per-phase timing; returns the per-phase ms of the last search_run
        (select, features, trunk convs..., heads, expand/backup)."""
        out = np.zeros(64, np.float32)
        n = self.lib.ap_search_profile(self.h, int(bool(enable)), _ptr(out), 64)
        return out[:max(n, 0)].copy()

    def search_root(self, game_ids=None, want_q=False):
        ids = _ids(game_ids)
        n = self._n(ids)
        count = np.zeros(n, np.int32)
        acts = np.zeros((n, self.S), np.int16)
        visits = np.zeros((n, self.S), np.int32)
        q = np.zeros((n, self.S), np.float64) if want_q else None
        rootn = np.zeros(n, np.int32)
        self._check(self.lib.ap_search_root(self.h, _ptr(ids), n, _ptr(count), _ptr(acts), _ptr(visits), _ptr(q),
                                            _ptr(rootn)))
        return count, acts, visits, q, rootn

    def search_root_probs(self, temp):
        out = np.zeros((self.G, self.S), np.float64)
        self._check(self.lib.ap_search_root_probs(self.h, float(temp), _ptr(out)))
        return out

    def selfplay_pick(self, temp=1.0, eps=0.25, alpha=0.3, seed=0, ply=0, want_noise=False):
        """Device-side MCTSPlayer.get_action(temp, return_prob=1) in self-play for every game:
        returns (moves int32 [G], pi float32 [G][S]) (+ the Dirichlet sample when want_noise)."""
        mv = np.zeros(self.G, np.int32)
        pi = np.zeros((self.G, self.S), np.float32)
        nz = np.zeros((self.G, self.S), np.float64) if want_noise else None
        self._check(self.lib.ap_selfplay_pick(self.h, float(temp), float(eps), float(alpha), int(seed), int(ply),
                                              _ptr(mv), _ptr(pi), _ptr(nz) if want_noise else None))
        return (mv, pi, nz) if want_noise else (mv, pi)

    def search_advance(self, moves, game_ids=None):
        ids = _ids(game_ids)
        n = self._n(ids)
        mv = np.ascontiguousarray(np.broadcast_to(np.asarray(moves, dtype=np.int32), (n,)))
        self._check(self.lib.ap_search_advance(self.h, _ptr(ids), n, _ptr(mv)))

    def search_set_active(self, active=None):
        """Searches skip the games whose ``active`` entry is false (None = all games search again)."""
        a = None if active is None else np.ascontiguousarray(np.asarray(active).astype(np.uint8)).reshape(self.G)
        self._check(self.lib.ap_search_set_active(self.h, _ptr(a)))

    def search_stats(self):
        out = np.zeros(8, np.uint64)
        self._check(self.lib.ap_search_stats(self.h, _ptr(out)))
        return dict(playouts=int(out[0]), children_scanned=int(out[1]), children_written=int(out[2]),
                    path_nodes=int(out[3]), terminal_leaves=int(out[4]), rollout_plies=int(out[5]))

    # -- mcts_pure ------------------------------------------------------------
    def pure_run(self, n_playout, seed=0, rollout_mode=0):
        mv = np.zeros(self.G, np.int32)
        self._check(self.lib.ap_pure_run(self.h, int(n_playout), int(seed), int(rollout_mode), _ptr(mv)))
        return mv

    def rollout_eval(self, seed=0, impl=0):
        """impl 0 = permutation rollout (default), 2 = ply-by-ply rollout (cross-check)."""
        v = np.zeros(self.G, np.int8)
        p = np.zeros(self.G, np.int16)
        self._check(self.lib.ap_rollout_eval2(self.h, int(seed), int(impl), _ptr(v), _ptr(p)))
        return v.astype(np.int32), p.astype(np.int32)

    def rollout_eval_keys(self, keys):
        """Permutation rollout with injected draws: keys uint32 [G][256], slot = row*16 + column (low 24 bits)."""
        keys = np.ascontiguousarray(keys, np.uint32)
        assert keys.shape == (self.G, 256)
        v = np.zeros(self.G, np.int8)
        p = np.zeros(self.G, np.int16)
        self._check(self.lib.ap_rollout_eval_keys(self.h, _ptr(keys), _ptr(v), _ptr(p)))
        return v.astype(np.int32), p.astype(np.int32)

    def rollout_hash(self):
        v = np.zeros(self.G, np.int8)
        self._check(self.lib.ap_rollout_hash(self.h, _ptr(v)))
        return v.astype(np.int32)

    # -- net ----------------------------------------------------------------
    def net_load(self, arch, params, n_blocks=0, n_filter=128, precision="auto"):
        """params: dict name -> float32 ndarray with the reference's names/shapes
        (arg and aux params merged; policy_value_net_mxnet.py:125-138).

        precision: "fp16" = fp16 tensor-core operands (fp32 accumulate); "split" = hi + lo fp16 pairs for
        activations and weights, three products per K step (residual net only, near-fp32); "split_act" = hi + lo
        activations x fp16 weights rounded by error diffusion along K, two products per K step (residual net
        only); "auto" = split_act for a residual net deeper than 3 blocks (plain fp16 measures 8.9e-4 at 3 blocks
        and 1.8e-3 at 10, against the 1e-3 parity budget; split_act 4.7e-4 at 10), fp16 otherwise."""
        if precision not in ("auto", "fp16", "split", "split_act"):
            raise ValueError("precision must be 'auto', 'fp16', 'split' or 'split_act'")
        if precision == "auto":
            precision = "split_act" if (arch == "resnet" and int(n_blocks) > 3) else "fp16"
        self.net_precision = precision
        names = list(params.keys())
        keep = [np.ascontiguousarray(params[k], dtype=np.float32) for k in names]
        arr = (L.ApTensor * len(names))()
        for i, (k, a) in enumerate(zip(names, keep)):
            arr[i].name = k.encode()
            arr[i].data = a.ctypes.data
            arr[i].numel = a.size
        code = {"simple": L.AP_ARCH_SIMPLE, "resnet": L.AP_ARCH_RESNET, "inception": L.AP_ARCH_INCEPTION}[arch]
        code |= {"fp16": 0, "split": L.AP_NET_SPLIT, "split_act": L.AP_NET_SPLIT_ACT}[precision]
        self._check(self.lib.ap_net_load(self.h, code, int(n_blocks), int(n_filter), arr, len(names)))
        self.net_names = names

    # -- replay ring (train_mxnet.py:57,115-135,196-199) ------------------------
    def replay_create(self, maxlen):
        self._check(self.lib.ap_replay_create(self.h, int(maxlen)))

    def replay_push(self, state_bits, pis, zs):
        """state_bits uint8 [n][ceil(9S/8)] (np.packbits of the 9xHxW planes), pis [n][S], zs [n]"""
        zs = np.ascontiguousarray(zs, dtype=np.float32).reshape(-1)
        n = zs.shape[0]
        sb = (9 * self.S + 7) // 8
        bits = np.ascontiguousarray(state_bits, dtype=np.uint8).reshape(n, sb)
        pis = np.ascontiguousarray(pis, dtype=np.float32).reshape(n, self.S)
        self._check(self.lib.ap_replay_push(self.h, _ptr(bits), _ptr(pis), _ptr(zs), n))

    def replay_push_sgf(self, seqs, winners):
        """SGF bootstrap (game.py:233-304) for a batch of recorded games: seqs = list of move-index lists
        (``seq_num_list``), winners = 1 / 2 / -1 per game.  Returns the per-game warning flags (1 = illegal move,
        game skipped)."""
        n = len(seqs)
        max_len = max(1, max(len(q) for q in seqs))
        mv = np.full((n, max_len), -1, np.int16)
        for g, q in enumerate(seqs):
            mv[g, :len(q)] = q
        ln = np.array([len(q) for q in seqs], np.int32)
        wn = np.ascontiguousarray(winners, dtype=np.int8).reshape(n)
        warn = np.zeros(n, np.uint8)
        self._check(self.lib.ap_replay_push_sgf(self.h, _ptr(mv), max_len, _ptr(ln), _ptr(wn), n, _ptr(warn)))
        return warn

    def replay_size(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.ap_replay_size(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def replay_gather(self, idx):
        """-> (states float32 [B][9][H][W], pis [B][S], zs [B]) as numpy (host copies)"""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        B = idx.shape[0]
        st = np.zeros((B, 9, self.height, self.width), np.float32)
        pi = np.zeros((B, self.S), np.float32)
        z = np.zeros(B, np.float32)
        self._check(self.lib.ap_replay_gather(self.h, _ptr(idx), B, _ptr(st), _ptr(pi), _ptr(z), 0))
        return st, pi, z

    def replay_gather_device(self, idx, states_ptr, pi_ptr, z_ptr):
        """Same, written straight into device buffers (raw pointers, e.g. torch ``data_ptr()``)."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        self._check(self.lib.ap_replay_gather(self.h, _ptr(idx), idx.shape[0], C.c_void_p(states_ptr), C.c_void_p(pi_ptr),
                                              C.c_void_p(z_ptr), 1))

    def replay_push_packed(self, records, n=None, device_ptr=None):
        """Append packed records ([state bytes padded to 4][S fp32 pi][fp32 z] each): a uint8 ndarray [n][record
        width] on the host, or ``device_ptr`` + ``n`` for records already on this GPU (outbox / all-gather result)."""
        if device_ptr is not None:
            self._check(self.lib.ap_replay_push_packed(self.h, C.c_void_p(int(device_ptr)), int(n), 1))
            return
        rec = np.ascontiguousarray(records, dtype=np.uint8)
        self._check(self.lib.ap_replay_push_packed(self.h, _ptr(rec), rec.shape[0], 0))

    # -- device-side self-play trajectories (game_ai.py:75,113-131) -------------
    def traj_create(self, max_plies=0, outbox_records=0):
        """From now on every ``selfplay_pick`` records its ply on the device; ``traj_finish`` moves finished games'
        records (z filled in) to the outbox."""
        if outbox_records <= 0:
            outbox_records = 2 * self.G * 64
        self._check(self.lib.ap_traj_create(self.h, int(max_plies), int(outbox_records)))

    def traj_append_forced(self, moves, game_ids):
        ids = _ids(game_ids)
        mv = np.ascontiguousarray(moves, dtype=np.int32)
        self._check(self.lib.ap_traj_append_forced(self.h, _ptr(ids), len(ids), _ptr(mv)))

    def traj_finish(self, game_ids, winners):
        ids = _ids(game_ids)
        w = np.ascontiguousarray(winners, dtype=np.int8)
        self._check(self.lib.ap_traj_finish(self.h, _ptr(ids), len(ids), _ptr(w)))

    def traj_discard(self, game_ids=None):
        ids = _ids(game_ids)
        self._check(self.lib.ap_traj_discard(self.h, _ptr(ids), self._n(ids)))

    def traj_outbox(self):
        """-> (device pointer, number of records, bytes per record)"""
        p, n, w = C.c_void_p(), C.c_int64(), C.c_int32()
        self._check(self.lib.ap_traj_outbox(self.h, C.byref(p), C.byref(n), C.byref(w)))
        return p.value, n.value, w.value

    def traj_outbox_clear(self):
        self._check(self.lib.ap_traj_outbox_clear(self.h))

    def _fwd(self, fn, states):
        st = np.ascontiguousarray(states, dtype=np.float32).reshape(-1, 9, self.height, self.width)
        B = st.shape[0]
        probs = np.zeros((B, self.S), np.float32)
        vals = np.zeros(B, np.float32)
        self._check(fn(self.h, _ptr(st), B, _ptr(probs), _ptr(vals)))
        return probs, vals.reshape(B, 1)

    def net_forward(self, states):
        return self._fwd(self.lib.ap_net_forward, states)

    def net_forward_precise(self, states):
        return self._fwd(self.lib.ap_net_forward_precise, states)

    def net_forward_leaves(self, precise=False, fetch=True):
        probs = np.zeros((self.G, self.S), np.float32) if fetch else None
        vals = np.zeros(self.G, np.float32) if fetch else None
        self._check(self.lib.ap_net_forward_leaves(self.h, int(bool(precise)), _ptr(probs), _ptr(vals)))
        return probs, vals

    def net_weights(self):
        """-> (device pointer, numel) of the flat fp32 master weights."""
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.ap_net_weights_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def net_layout(self):
        cap = 512
        names = (C.c_char_p * cap)()
        offs = np.zeros(cap, np.int64)
        nums = np.zeros(cap, np.int64)
        n = self.lib.ap_net_layout(self.h, cap, names, _ptr(offs), _ptr(nums))
        if n < 0:
            self._check(n)
        return [(names[i].decode(), int(offs[i]), int(nums[i])) for i in range(n)]

    def net_refresh(self):
        self._check(self.lib.ap_net_refresh(self.h))


def rollout_hash_host(cells, cur, width, height):
    """Host twin of board_hash()/hash_eval() in csrc/rollout.cu for a NON-terminal
    position: FNV-1a over (cur, rows) -> value in {-1,0,1}."""
    h = 2166136261
    h = ((h ^ int(cur)) * 16777619) & 0xffffffff
    c = np.asarray(cells).reshape(height, width)
    for r in range(16):
        p1 = p2 = 0
        if r < height:
            for w in range(width):
                if c[r, w] == 1:
                    p1 |= 1 << w
                elif c[r, w] == 2:
                    p2 |= 1 << w
        h = ((h ^ p1) * 16777619) & 0xffffffff
        h = ((h ^ p2) * 16777619) & 0xffffffff
    return int(h % 3) - 1
