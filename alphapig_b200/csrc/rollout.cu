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
  const int lane = threadIdx.x;
  const int g = blockIdx.x;
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
    int v = (mode == 1) ? hash_eval(b, geo) : rollout_eval(b, rng, geo, lane, plies);
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

__global__ void k_rollout_eval(Geo geo, const uint32_t* rows, const BoardMeta* meta, unsigned long long seed,
                               int8_t* out_value, int16_t* out_plies) {
  int lane = threadIdx.x & 31;
  int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  WBoard b = wb_load(rows, meta, g, lane);
  Pcg rng = pcg_seed(seed, (unsigned long long)g);
  int plies;
  int v = rollout_eval(b, rng, geo, lane, plies);
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
void launch_rollout_eval(ap_engine* e, uint64_t seed, int8_t* d_value, int16_t* d_plies) {
  k_rollout_eval<<<(e->geo.G + 3) / 4, 128, 0, e->stream>>>(e->geo, e->rows, e->meta, seed, d_value, d_plies);
}
void launch_rollout_hash(ap_engine* e, int8_t* d_value) {
  k_rollout_hash<<<(e->geo.G + 3) / 4, 128, 0, e->stream>>>(e->geo, e->rows, e->meta, d_value);
}
