// mcts_pure on device: whole searches (n_playout playouts incl. random rollouts) in one launch,
// one CTA (= one warp) per game.  Replaces reference mcts_pure.MCTS (mcts_pure.py:96-182).
#include "kernels.h"
#include "tree.cuh"

struct Pcg {
  unsigned long long s, inc;
};
__device__ __forceinline__ unsigned pcg_next(Pcg& r) {
  unsigned long long o = r.s;
  r.s = o * 6364136223846793005ull + r.inc;
  unsigned x = (unsigned)(((o >> 18u) ^ o) >> 27u);
  unsigned rot = (unsigned)(o >> 59u);
  return (x >> rot) | (x << ((32u - rot) & 31u));
}
__device__ __forceinline__ Pcg pcg_seed(unsigned long long seed, unsigned long long seq) {
  Pcg r;
  r.s = 0;
  r.inc = (seq << 1) | 1ull;
  pcg_next(r);
  r.s += seed;
  pcg_next(r);
  return r;
}

// FNV-1a over (cur, rows[0..16)); host twin: alphapig_b200.engine.rollout_hash_host
__device__ __forceinline__ unsigned board_hash(const WBoard& b) {
  unsigned h = 2166136261u;
  h = (h ^ (unsigned)b.cur) * 16777619u;
  for (int i = 0; i < AP_ROWS; ++i) {
    unsigned r = __shfl_sync(AP_FULL, b.row, i);
    h = (h ^ (r & 0xffffu)) * 16777619u;
    h = (h ^ (r >> 16)) * 16777619u;
  }
  return h;
}

// one uniformly random legal move: rollout_policy_fn + arg-max of iid uniforms (mcts_pure.py:13-17,148-150)
__device__ __forceinline__ void rollout_step(WBoard& b, Pcg& rng, int W, int H, int lane) {
  uint32_t e = wb_empty_row(b, W, H, lane);
  int c = __popc(e);
  int pre = c;
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    int t = __shfl_up_sync(AP_FULL, pre, d);
    if (lane >= d) pre += t;
  }
  int total = __shfl_sync(AP_FULL, pre, 15);
  int r = (int)__umulhi(pcg_next(rng), (unsigned)total);
  bool mine = (lane < 16) && (r >= pre - c) && (r < pre);
  unsigned who = __ballot_sync(AP_FULL, mine);
  int h = __ffs(who) - 1;
  int k = r - (pre - c);
  int w = 0;
  if (mine) {
    for (int j = 0; j < k; ++j) e &= e - 1;
    w = __ffs(e) - 1;
    b.row |= (1u << w) << ((b.cur == 2) ? 16 : 0);
  }
  w = __shfl_sync(AP_FULL, w, h);
  int mv = h * W + w;
  b.hist = (b.hist << 16) | (unsigned long long)(uint16_t)mv;
  b.nst += 1;
  b.last = mv;
  b.cur = 3 - b.cur;
}

// MCTS._evaluate_rollout (mcts_pure.py:138-157): value from the perspective of the player to move at `b`.
__device__ __forceinline__ int rollout_eval(WBoard b, Pcg& rng, const Geo& geo, int lane, int& plies) {
  const int player = b.cur;
  int winner;
  bool end = wb_game_end(b, geo.n_in_row, geo.S, winner);
  plies = 0;
  while (!end && plies < 1000) {
    rollout_step(b, rng, geo.W, geo.H, lane);
    ++plies;
    // only the side that just moved can have completed a line
    int mover = 3 - b.cur;
    uint32_t x = (mover == 1) ? (b.row & 0xffffu) : (b.row >> 16);
    bool win = (b.nst >= geo.n_in_row + 2) && wb_colour_wins(x, geo.n_in_row);
    winner = win ? mover : -1;
    end = win || (b.nst >= geo.S);
  }
  if (winner == -1) return 0;
  return (winner == player) ? 1 : -1;
}

// ---- permutation rollout ---------------------------------------------------------------------------
// A game of uniformly random legal moves from a position with E empty cells is a uniformly random
// permutation of those cells, stone i of the permutation taking the colour of ply i (the reference draws
// np.random.rand(A) and plays the arg-max every ply, mcts_pure.py:13-17,148-150 - sampling without
// replacement, i.e. exactly that permutation).  "Some line of n exists after the first t plies" is monotone
// in t, so instead of ~100 sequential plies the warp
//   1. gives every empty cell an iid random key and sorts the 256 board slots once (register bitonic
//      network, 8 slots per lane) - slot order = ply index ("rank") of every cell,
//   2. keeps the 8 rank bit-planes of its board row, and
//   3. finds the largest t with no line by descending the rank bits (8 win checks on both colours at once):
//      the game ends at ply t+1, won by the colour of ply t, or is a tie when t reaches E.
// Same distribution of (result, plies) as the ply-by-ply loop at ~1/10 of the instructions.
// Keys are 24 random bits + the cell index (unique); two cells tie in the random part with probability
// < 2e-3 per rollout and are then ordered by cell index - far below anything the parity statistics resolve.
// Requires W <= 15 (bit 15 of each colour half stays clear and isolates the two halves in the packed check).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// ascending bitonic sort of 256 keys, element i = lane*8 + r.  Mirror formulation: every block size k starts
// with the flip step (partner i ^ (k-1)) and continues with half-cleaners (partner i ^ j, j = k/4 .. 1), so the
// lower index always keeps the minimum and every in-lane exchange has a compile-time direction.
__device__ __forceinline__ void warp_sort256(uint32_t (&v)[8], int lane) {
#pragma unroll
  for (int k = 2; k <= 256; k <<= 1) {
    if (k <= 8) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int r2 = r ^ (k - 1);
        if (r < r2) {
          const uint32_t a = min(v[r], v[r2]), b = max(v[r], v[r2]);
          v[r] = a;
          v[r2] = b;
        }
      }
    } else {
      const int m = (k >> 3) - 1;                       // partner lane = lane ^ m, partner register = 7 - r
      const bool keep_min = (lane & (k >> 4)) == 0;     // this lane is the lower one of the pair
      uint32_t o[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) o[r] = __shfl_xor_sync(AP_FULL, v[7 - r], m);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = keep_min ? min(v[r], o[r]) : max(v[r], o[r]);
    }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
      if (j >= 8) {
        const int lj = j >> 3;
        const bool keep_min = (lane & lj) == 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t o = __shfl_xor_sync(AP_FULL, v[r], lj);
          v[r] = keep_min ? min(v[r], o) : max(v[r], o);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if ((r & j) == 0) {
            const int r2 = r | j;
            const uint32_t a = min(v[r], v[r2]), b = max(v[r], v[r2]);
            v[r] = a;
            v[r2] = b;
          }
        }
      }
    }
  }
}

// both colours at once: x = player-1 row | player-2 row << 16 with bits 15 and 31 clear (W <= 15), so every
// window that would cross from one half into the other contains a clear bit
__device__ __forceinline__ bool wb_packed_wins(uint32_t x, int n) {
  uint32_t th = x, tv = x, td = x, ta = x;
  for (int k = 1; k < n; ++k) {
    const uint32_t up = __shfl_down_sync(AP_FULL, x, k);
    th &= x >> k;
    tv &= up;
    td &= up >> k;
    ta &= up << k;
  }
  return __any_sync(AP_FULL, (th | tv | td | ta) != 0u);
}

// bit `bit` of the four bytes of w -> 4-bit nibble (byte 0 -> bit 0)
__device__ __forceinline__ uint32_t byte_bits(uint32_t w, int bit) {
  return (((w >> bit) & 0x01010101u) * 0x10204080u) >> 28;
}

// MCTS._evaluate_rollout (mcts_pure.py:138-157) by permutation; s_rank: 256 bytes of shared memory of this warp
__device__ __forceinline__ int rollout_eval_perm(const WBoard& b, Pcg& rng, const Geo& geo, int lane, int& plies,
                                                 uint8_t* s_rank) {
  const int player = b.cur;
  plies = 0;
  {
    int winner;
    if (wb_game_end(b, geo.n_in_row, geo.S, winner)) return (winner == -1) ? 0 : ((winner == player) ? 1 : -1);
  }
  const uint32_t e = wb_empty_row(b, geo.W, geo.H, lane);  // lanes >= H: 0
  const int E = geo.S - b.nst;
  const uint32_t k0 = pcg_next(rng), k1 = pcg_next(rng);
  // slot (lane, r) starts as cell16 = lane*8 + r = row (lane >> 1), column (lane & 1) * 8 + r
  const uint32_t er = __shfl_sync(AP_FULL, e, lane >> 1) >> ((lane & 1) * 8);
  uint32_t v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t cell = (uint32_t)(lane * 8 + r);
    const uint32_t h = min((mix32(k0 ^ (cell * 0x9E3779B9u)) + k1) >> 8, 0xFFFFFEu);
    v[r] = ((er >> r) & 1u) ? ((h << 8) | cell) : 0xFFFFFFFFu;
  }
  warp_sort256(v, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (v[r] != 0xFFFFFFFFu) s_rank[v[r] & 0xFFu] = (uint8_t)(lane * 8 + r);
  __syncwarp();
  uint32_t pl[8];
  {
    uint4 rr = make_uint4(0u, 0u, 0u, 0u);
    if (lane < AP_ROWS) rr = *reinterpret_cast<const uint4*>(s_rank + lane * 16);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit)
      pl[bit] = byte_bits(rr.x, bit) | (byte_bits(rr.y, bit) << 4) | (byte_bits(rr.z, bit) << 8) |
                (byte_bits(rr.w, bit) << 12);
  }
  // largest t such that the first t plies complete no line
  uint32_t eq = e, lt = 0u;
  int t = 0;
  const int need = geo.n_in_row + 2 - b.nst;  // has_a_winner looks only at boards with >= n+2 stones (game.py:134)
#pragma unroll
  for (int bit = 7; bit >= 0; --bit) {
    const uint32_t placed = lt | (eq & ~pl[bit]);          // cells with rank < t | 1 << bit
    const uint32_t first = placed & ~pl[0], second = placed & pl[0];  // even ranks: the side to move at the leaf
    const uint32_t x = b.row | ((player == 1) ? (first | (second << 16)) : (second | (first << 16)));
    const int tt = t | (1 << bit);
    const bool line = (min(tt, E) >= need) && wb_packed_wins(x, geo.n_in_row);
    if (!line) {
      t = tt;
      lt = placed;
      eq &= pl[bit];
    } else {
      eq &= ~pl[bit];
    }
  }
  if (t >= E) {
    plies = E;
    return 0;  // board full, no line: tie
  }
  plies = t + 1;
  return (t & 1) ? -1 : 1;  // ply t (0-based) completes the line; even plies belong to `player`
}

__device__ __forceinline__ int hash_eval(const WBoard& b, const Geo& geo) {
  int winner;
  bool end = wb_game_end(b, geo.n_in_row, geo.S, winner);
  if (!end) return (int)(board_hash(b) % 3u) - 1;
  if (winner == -1) return 0;
  return (winner == b.cur) ? 1 : -1;
}

// MCTS.get_move (mcts_pure.py:159-169): tree reset, n_playout x _playout (:114-136), arg-max visits.
__global__ void __launch_bounds__(32)
k_pure_run(Geo geo, const uint32_t* __restrict__ rows, const BoardMeta* __restrict__ meta, Pools pl, int n_playout,
           unsigned long long seed, int mode, int32_t* out_move, int32_t* errflag, unsigned long long* stats) {
  __shared__ int16_t s_list[AP_MAX_S];
  __shared__ __align__(16) uint8_t s_rank[256];
  const int lane = threadIdx.x;
  const int g = blockIdx.x;
  const bool perm = (mode == 0) && geo.W <= 15;  // mode 2 (and 16-wide boards): ply-by-ply rollout
  const size_t base = (size_t)g * geo.cap;
  const WBoard root = wb_load(rows, meta, g, lane);
  Pcg rng = pcg_seed(seed, (unsigned long long)g);
  if (lane == 0) tree_write_root(pl, base, g);
  __syncwarp();
  unsigned long long scanned = 0, written = 0, pathn = 0, plies_total = 0;
  for (int it = 0; it < n_playout; ++it) {
    WBoard b = root;
    int node = 0;
    while (true) {
      int cs = pl.child_start[base + node];
      if (cs < 0) break;
      int cc = pl.child_count[base + node];
      int np = pl.N[base + node];
      int bi = tree_select_child(pl, base, cs, cc, np, geo.c_puct, lane);
      int mv = pl.move[base + cs + bi];
      wb_do_move(b, mv, geo.W, lane);
      node = cs + bi;
      scanned += cc;
    }
    int winner;
    bool end = wb_game_end(b, geo.n_in_row, geo.S, winner);
    if (!end) {
      int A = wb_legal_list(b, geo.W, geo.H, lane, s_list);
      double p = __ddiv_rn(1.0, (double)A);  // np.ones(A)/A  (mcts_pure.py:24)
      bool ok = tree_expand(pl, base, g, geo.cap, node, A, s_list, [&](int, int) { return p; }, lane);
      if (!ok) {
        if (lane == 0) errflag[g] = AP_ERR_POOL_EXHAUSTED;
        break;
      }
      written += A;
    }
    int plies = 0;
    int v = (mode == 1) ? hash_eval(b, geo)
                        : (perm ? rollout_eval_perm(b, rng, geo, lane, plies, s_rank) : rollout_eval(b, rng, geo, lane, plies));
    plies_total += plies;
    if (lane == 0) pathn += tree_backup(pl, base, node, -(double)v);
    __syncwarp();
  }
  // first max by visit count over root children
  int cs = pl.child_start[base];
  int cc = (cs >= 0) ? pl.child_count[base] : 0;
  int bn = -1, bi = INT_MAX;
  for (int k = lane; k < cc; k += 32) {
    int n = pl.N[base + cs + k];
    if (n > bn) {
      bn = n;
      bi = k;
    }
  }
  for (int d = 16; d >= 1; d >>= 1) {
    int on = __shfl_xor_sync(AP_FULL, bn, d);
    int oi = __shfl_xor_sync(AP_FULL, bi, d);
    if (on > bn || (on == bn && oi < bi)) {
      bn = on;
      bi = oi;
    }
  }
  if (lane == 0) {
    out_move[g] = (bi != INT_MAX) ? (int)pl.move[base + cs + bi] : -1;
    atomicAdd(&stats[0], (unsigned long long)n_playout);
    atomicAdd(&stats[1], scanned);
    atomicAdd(&stats[2], written);
    atomicAdd(&stats[3], pathn);
    atomicAdd(&stats[5], plies_total);
  }
}

__global__ void k_rollout_eval(Geo geo, const uint32_t* rows, const BoardMeta* meta, unsigned long long seed, int impl,
                               int8_t* out_value, int16_t* out_plies) {
  __shared__ __align__(16) uint8_t s_rank[4][256];
  int lane = threadIdx.x & 31;
  int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  WBoard b = wb_load(rows, meta, g, lane);
  Pcg rng = pcg_seed(seed, (unsigned long long)g);
  int plies;
  int v = (impl == 0 && geo.W <= 15) ? rollout_eval_perm(b, rng, geo, lane, plies, s_rank[threadIdx.x >> 5])
                                     : rollout_eval(b, rng, geo, lane, plies);
  if (lane == 0) {
    out_value[g] = (int8_t)v;
    out_plies[g] = (int16_t)plies;
  }
}

__global__ void k_rollout_hash(Geo geo, const uint32_t* rows, const BoardMeta* meta, int8_t* out_value) {
  int lane = threadIdx.x & 31;
  int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  WBoard b = wb_load(rows, meta, g, lane);
  int v = hash_eval(b, geo);
  if (lane == 0) out_value[g] = (int8_t)v;
}

void launch_pure_run(ap_engine* e, int n_playout, uint64_t seed, int mode, int32_t* d_move) {
  k_pure_run<<<e->geo.G, 32, 0, e->stream>>>(e->geo, e->rows, e->meta, e->pools, n_playout, seed, mode, d_move,
                                            e->errflag, e->stats);
}
void launch_rollout_eval(ap_engine* e, uint64_t seed, int impl, int8_t* d_value, int16_t* d_plies) {
  k_rollout_eval<<<(e->geo.G + 3) / 4, 128, 0, e->stream>>>(e->geo, e->rows, e->meta, seed, impl, d_value, d_plies);
}
void launch_rollout_hash(ap_engine* e, int8_t* d_value) {
  k_rollout_hash<<<(e->geo.G + 3) / 4, 128, 0, e->stream>>>(e->geo, e->rows, e->meta, d_value);
}
