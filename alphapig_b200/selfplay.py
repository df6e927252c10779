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
                 noise_eps=0.25, dirichlet_alpha=0.3, node_capacity=0, record_states=True):
        self.net = net
        self.G = n_games
        self.n_playout = n_playout
        self.temp = temp
        self.eps = noise_eps
        self.alpha = dirichlet_alpha
        self.record_states = record_states
        self.rs = np.random.RandomState(seed)
        self.eng = net.search_engine(n_in_row=n_in_row, c_puct=c_puct, n_playout=n_playout, n_games=n_games,
                                     node_capacity=node_capacity)
        self.S = self.eng.S
        self.eng.boards_reset()
        self.eng.search_advance(-1)
        self._hist = [[] for _ in range(n_games)]  # per game: (state bits, pi, player)
        self.finished_games = 0
        self.plies = 0

    def load_positions(self, cells, meta):
        """Start every slot from a given position (benchmark's synthetic positions)."""
        self.eng.boards_import(cells, meta)
        self.eng.search_advance(-1)
        self._hist = [[] for _ in range(self.G)]

    def step(self):
        """One ply for every game.  Returns the list of finished-game records
        [(winner, states uint8[n][9*S/8 packed], pis float64[n][S], z float64[n])]."""
        eng, G, S = self.eng, self.G, self.S
        eng.search_run(self.n_playout)
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
            feats = np.packbits(eng.boards_features().astype(np.uint8).reshape(G, -1), axis=1)
            _, meta = eng.boards_export()
            pi = np.zeros((G, S))
            rows = np.nonzero(mask)
            pi[rows[0], acts[rows]] = probs[rows]
            for g in range(G):
                self._hist[g].append((feats[g], pi[g], int(meta[g, 0])))
        eng.search_advance(moves)
        eng.boards_do_move(moves)
        end, winner = eng.boards_status()
        self.plies += G
        done = np.nonzero(end)[0].astype(np.int32)
        out = []
        if len(done):
            for g in done:
                h = self._hist[g]
                if h:
                    players = np.array([p for _, _, p in h])
                    z = np.zeros(len(h))
                    if winner[g] != -1:
                        z[players == winner[g]] = 1.0
                        z[players != winner[g]] = -1.0
                    out.append((int(winner[g]), np.stack([s for s, _, _ in h]), np.stack([p for _, p, _ in h]), z))
                else:
                    out.append((int(winner[g]), None, None, None))
                self._hist[g] = []
            eng.boards_reset(done)
            eng.search_advance(np.full(len(done), -1, np.int32), done)
            self.finished_games += len(done)
        return out
