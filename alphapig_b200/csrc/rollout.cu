// mcts_pure on device: whole searches (n_playout playouts incl. random rollouts) in one launch,
// one CTA (= one warp) per game.  Replaces reference mcts_pure.MCTS (mcts_pure.py:96-182).
// Random rollouts are drawn as one random permutation of the empty cells + a bit descent to the first completed line
// (rollout_eval_perm), the tree keeps children lazy (pure_select_child); DESIGN.md 3.5 has the measurements.
#include "kernels.h"
#include "tree.cuh"

// resident CTAs (= games) per SM the pure kernel is compiled for (register budget 65536 / (32 * N)).  Measured on
// B200, configs[2]: 24 (80 registers, no spills) is the best of 32 / 28 / 24 / 20; the kernel is issue-bound
// (73 % issue-active under ncu), so more resident warps at the price of spills do not pay.  A variant that kept the root's children
// in shared memory executed 16 % more instructions and was 13 % slower - the children blocks are L2-resident anyway.
#ifndef AP_PURE_MINBLK
#define AP_PURE_MINBLK 24
#endif

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

// Philox4x32-10 (Salmon et al., SC'11): counter-based, every (key, counter) an independent stream - the random
// bits of the permutation rollouts, addressed by (playout index, draw, game, lane) under the 64-bit seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll 1
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
// 8 x 24 random bits for the 8 board slots of this lane
__device__ __forceinline__ void perm_draws(unsigned long long seed, uint32_t g, uint32_t lane, uint32_t playout,
                                           uint32_t (&h)[8]) {
  uint32_t a[4], b[4];
  philox4x32_10(playout, 0u, g, lane, (uint32_t)seed, (uint32_t)(seed >> 32), a);
  philox4x32_10(playout, 1u, g, lane, (uint32_t)seed, (uint32_t)(seed >> 32), b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = a[i] >> 8;
    h[4 + i] = b[i] >> 8;
  }
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
// The random bits are Philox4x32-10 outputs (one independent stream per game, lane and playout).  Cheaper
// sources were tried and REJECTED by the 10^6-rollout comparison with the NumPy reference sample
// (oracle/rollout.py): a two-multiply hash of (rollout, cell) put a 1 % excess into the spread of game lengths,
// and PCG32 streams that differ only in their increment are correlated enough across the lanes of one rollout
// to lengthen the mean game by 0.3 plies.  The algorithm itself is pinned exactly with injected draws
// (ap_rollout_eval_keys vs the oracle playing the same order move by move).
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
// lower index always keeps the minimum and every in-lane exchange has a compile-time direction.  The block
// sizes that cross lanes (k = 16 .. 256) run as a ROLLED loop: the kernel's hot loop has to stay inside the
// 32 KB instruction cache shared by warps that sit at different points of it (first version, fully unrolled:
// "no instruction" was the top stall reason under ncu).
__device__ __forceinline__ void cmpx(uint32_t& a, uint32_t& b) {
  const uint32_t lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}
__device__ __forceinline__ void sort_tail421(uint32_t (&v)[8]) {
  cmpx(v[0], v[4]); cmpx(v[1], v[5]); cmpx(v[2], v[6]); cmpx(v[3], v[7]);
  cmpx(v[0], v[2]); cmpx(v[1], v[3]); cmpx(v[4], v[6]); cmpx(v[5], v[7]);
  cmpx(v[0], v[1]); cmpx(v[2], v[3]); cmpx(v[4], v[5]); cmpx(v[6], v[7]);
}
__device__ __forceinline__ void warp_sort256(uint32_t (&v)[8], int lane) {
  // k = 2, 4, 8: inside the lane
  cmpx(v[0], v[1]); cmpx(v[2], v[3]); cmpx(v[4], v[5]); cmpx(v[6], v[7]);
  cmpx(v[0], v[3]); cmpx(v[1], v[2]); cmpx(v[4], v[7]); cmpx(v[5], v[6]);
  cmpx(v[0], v[1]); cmpx(v[2], v[3]); cmpx(v[4], v[5]); cmpx(v[6], v[7]);
  cmpx(v[0], v[7]); cmpx(v[1], v[6]); cmpx(v[2], v[5]); cmpx(v[3], v[4]);
  cmpx(v[0], v[2]); cmpx(v[1], v[3]); cmpx(v[4], v[6]); cmpx(v[5], v[7]);
  cmpx(v[0], v[1]); cmpx(v[2], v[3]); cmpx(v[4], v[5]); cmpx(v[6], v[7]);
#pragma unroll 1
  for (int kk = 1; kk <= 16; kk <<= 1) {  // kk = k / 16
    {
      const int m = 2 * kk - 1;                  // flip: partner lane = lane ^ m, partner register = 7 - r
      const bool keep_min = (lane & kk) == 0;    // this lane is the lower one of the pair
      uint32_t o[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) o[r] = __shfl_xor_sync(AP_FULL, v[7 - r], m);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = keep_min ? min(v[r], o[r]) : max(v[r], o[r]);
    }
#pragma unroll 1
    for (int lj = kk >> 1; lj > 0; lj >>= 1) {   // half-cleaners across lanes: j = 8 * lj
      const bool keep_min = (lane & lj) == 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t o = __shfl_xor_sync(AP_FULL, v[r], lj);
        v[r] = keep_min ? min(v[r], o) : max(v[r], o);
      }
    }
    sort_tail421(v);                             // j = 4, 2, 1
  }
}

// both colours at once: x = player-1 row | player-2 row << 16 with bits 15 and 31 clear (W <= 15), so every
// window that would cross from one half into the other contains a clear bit
template <int N>
__device__ __forceinline__ bool wb_packed_wins_n(uint32_t x) {
  uint32_t th = x, tv = x, td = x, ta = x;
#pragma unroll
  for (int k = 1; k < N; ++k) {
    const uint32_t up = __shfl_down_sync(AP_FULL, x, k);
    th &= x >> k;
    tv &= up;
    td &= up >> k;
    ta &= up << k;
  }
  return __any_sync(AP_FULL, (th | tv | td | ta) != 0u);
}
__device__ __forceinline__ bool wb_packed_wins(uint32_t x, int n) {
  if (n == 5) return wb_packed_wins_n<5>(x);  // five-in-a-row: immediate shifts, no loop
  uint32_t th = x, tv = x, td = x, ta = x;
#pragma unroll 1
  for (int k = 1; k < n; ++k) {
    const uint32_t up = __shfl_down_sync(AP_FULL, x, k);
    th &= x >> k;
    tv &= up;
    td &= up >> k;
    ta &= up << k;
  }
  return __any_sync(AP_FULL, (th | tv | td | ta) != 0u);
}

// s_rank holds the rank of column c of board row h at byte h*16 + rank_slot(c); with that order bit `bit` of the
// 16 ranks of a row gathers into a 16-bit column mask with a handful of shifts.  rank_slot is an involution.
// The sort payload of a board slot is this byte offset, so the ranks scatter with one store each.
__host__ __device__ __forceinline__ uint32_t rank_slot(uint32_t c) { return ((c & 3u) << 2) | (c >> 2); }
__device__ __forceinline__ uint32_t rank_plane(const uint4& rr, int bit) {
  const uint32_t M = 0x01010101u;
  const uint32_t a = (rr.x >> bit) & M, b = (rr.y >> bit) & M, c = (rr.z >> bit) & M, d = (rr.w >> bit) & M;
  const uint32_t s = (a + 2u * b) + 4u * (c + 2u * d);  // byte i: columns 4i .. 4i+3 in bits 0..3
  const uint32_t t = s | (s >> 4);
  return __byte_perm(t, 0u, 0x4420);
}

// MCTS._evaluate_rollout (mcts_pure.py:138-157) by permutation from a NON-terminal position (the caller has
// done the game_end() of :143); s_rank: 256 bytes of shared memory of this warp
// h: this lane's 8 random draws (24 bits each); element r of lane L is the board slot with payload L*8 + r =
// row (L >> 1), column rank_slot((L & 1) * 8 + r); ties in the draw are ordered by payload
__device__ __forceinline__ int rollout_eval_perm(const WBoard& b, const uint32_t (&h)[8], const Geo& geo, int lane,
                                                 int& plies, uint8_t* s_rank) {
  const int player = b.cur;
  const uint32_t e = wb_empty_row(b, geo.W, geo.H, lane);  // lanes >= H: 0
  const int E = geo.S - b.nst;
  // column of element r: rank_slot((lane & 1) * 8 + r) = (r & 3) * 4 + (r >> 2) + 2 * (lane & 1)
  const uint32_t er = __shfl_sync(AP_FULL, e, lane >> 1) >> ((lane & 1) * 2);
  uint32_t v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t payload = (uint32_t)(lane * 8 + r);
    v[r] = ((er >> ((r & 3) * 4 + (r >> 2))) & 1u) ? ((min(h[r], 0xFFFFFEu) << 8) | payload) : 0xFFFFFFFFu;
  }
  warp_sort256(v, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (v[r] != 0xFFFFFFFFu) s_rank[v[r] & 0xFFu] = (uint8_t)(lane * 8 + r);
  __syncwarp();
  uint4 rr = make_uint4(0u, 0u, 0u, 0u);
  if (lane < AP_ROWS) rr = *reinterpret_cast<const uint4*>(s_rank + lane * 16);
  const uint32_t odd = rank_plane(rr, 0);  // odd ranks: the opponent of the side to move at the leaf
  const uint32_t mine_sh = (player == 1) ? 0u : 16u, other_sh = 16u - mine_sh;
  // largest t such that the first t plies complete no line
  uint32_t eq = e, lt = 0u;
  int t = 0;
  const int need = geo.n_in_row + 2 - b.nst;  // has_a_winner looks only at boards with >= n+2 stones (game.py:134)
#pragma unroll 1
  for (int bit = 7; bit >= 0; --bit) {
    const uint32_t plb = rank_plane(rr, bit);
    const uint32_t placed = lt | (eq & ~plb);  // cells with rank < t | 1 << bit
    const uint32_t x = b.row | ((placed & ~odd) << mine_sh) | ((placed & odd) << other_sh);
    const int tt = t | (1 << bit);
    const bool line = (min(tt, E) >= need) && wb_packed_wins(x, geo.n_in_row);
    if (!line) {
      t = tt;
      lt = placed;
      eq &= plb;
    } else {
      eq &= ~plb;
    }
  }
  if (t >= E) {
    plies = E;
    return 0;  // board full, no line: tie
  }
  plies = t + 1;
  return (t & 1) ? -1 : 1;  // ply t (0-based) completes the line; even plies belong to `player`
}

// Board.game_end (game.py:160-167) with one packed line check for both colours (W <= 15); the colour is only
// resolved when a line exists
__device__ __forceinline__ bool wb_game_end_packed(const WBoard& b, int n, int S, int& winner) {
  const bool line = (b.nst >= n + 2) && wb_packed_wins(b.row, n);
  winner = -1;
  if (line) winner = wb_colour_wins(b.row & 0xffffu, n) ? 1 : 2;
  return line || b.nst >= S;
}

__device__ __noinline__ double ddiv_slow(double x, int d) { return __ddiv_rn(x, (double)d); }

__device__ __forceinline__ int hash_eval(const WBoard& b, const Geo& geo) {
  int winner;
  bool end = wb_game_end(b, geo.n_in_row, geo.S, winner);
  if (!end) return (int)(board_hash(b) % 3u) - 1;
  if (winner == -1) return 0;
  return (winner == b.cur) ? 1 : -1;
}

// TreeNode.select (mcts_pure.py:43-49,78-80) for mcts_pure trees.  Every child of a node carries the same prior
// 1/A (policy_value_fn, mcts_pure.py:20-25; A = child count), so x = (c_puct*P)*sqrt(Np) is shared and
// u = x/(1+N) depends on the visit count alone: lane L computes the correctly rounded quotient for N = L once
// and the children fetch theirs by shuffle (own division only for N >= 32) - the same fp64 operations on the
// same operands as tree_select_child, ~6 instead of ~35 instructions per child and no load of P.
//
// LAZY CHILDREN.  All unvisited children of a node score exactly 0.0 + x/(1+0) and Python's max() takes the first
// maximum, so an unvisited child is only ever chosen as the LOWEST-index unvisited one: the visited children of a
// node are always the prefix [0, nv) of its (ascending-move) child list.  The kernel therefore reserves the child
// block at expansion but writes nothing; select scans the nv visited records plus ONE virtual candidate (index nv,
// score 0.0 + u[0]), and a record is materialised (move = nv-th legal move of the node's position, which the
// descent holds in registers) only when that candidate wins.  Same visit counts, Q and moves as the eager tree
// (golden and oracle tests in rollout_mode 1), ~200x fewer child records written, no legal-move list per playout.
// nv lives in the node's `parent` field (the path is kept in shared memory, parents are not needed during the
// search); the root's block is completed and the parents restored before the kernel returns.
// s_prior[A] = 1.0 / A (np.ones(A)/A), filled once per kernel
__device__ __forceinline__ int pure_select_child(const Pools& pl, size_t base, int cs, int cc, int nv, int np,
                                                 double c_puct, const double* s_prior, int lane, int& best_move) {
  constexpr int PER = AP_MAX_S / 32;
  double q[PER];
  int n[PER], m[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lane + 32 * j;
    const size_t c = base + cs + i;
    q[j] = 0.0, n[j] = 0, m[j] = -1;
    if (i < nv) {
      q[j] = pl.Q[c];
      n[j] = pl.N[c];
      m[j] = pl.move[c];
    }
  }
  const double x = __dmul_rn(__dmul_rn(c_puct, s_prior[cc]), __dsqrt_rn((double)np));
  const double ut = __ddiv_rn(x, (double)(1 + lane));
  double bv = -CUDART_INF;
  int bi = INT_MAX, bm = -1;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    if (32 * j < nv) {  // warp-uniform
      const int i = lane + 32 * j;
      double u = __shfl_sync(AP_FULL, ut, n[j] & 31);
      if (n[j] >= 32) u = ddiv_slow(x, 1 + n[j]);
      const double v = __dadd_rn(q[j], u);
      if (i < nv && (v > bv || bi == INT_MAX)) {  // first element always taken, later only if strictly greater
        bv = v;
        bi = i;
        bm = m[j];
      }
    }
  }
  {
    // the first unvisited child (Q = 0, N = 0): index nv is larger than every index this lane has seen
    const double v0 = __dadd_rn(0.0, __shfl_sync(AP_FULL, ut, 0));
    if (nv < cc && lane == (nv & 31) && (v0 > bv || bi == INT_MAX)) {
      bv = v0;
      bi = nv;
      bm = -1;
    }
  }
  // python max(): the largest score, the earliest child among equals.  Scores map to order-preserving 64-bit
  // integers (-0.0 folded into +0.0 first; no NaNs: Q and u are finite), then three warp reductions.
  const unsigned long long bits = (unsigned long long)__double_as_longlong(__dadd_rn(bv, 0.0));
  const unsigned long long key = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
  const bool have = bi != INT_MAX;
  const uint32_t khi = have ? (uint32_t)(key >> 32) : 0u, klo = (uint32_t)key;
  const uint32_t mhi = __reduce_max_sync(AP_FULL, khi);
  const bool c1 = have && khi == mhi;
  const uint32_t mlo = __reduce_max_sync(AP_FULL, c1 ? klo : 0u);
  const bool c2 = c1 && klo == mlo;
  const uint32_t best = __reduce_min_sync(AP_FULL, c2 ? (uint32_t)bi : 0xFFFFFFFFu);
  best_move = __shfl_sync(AP_FULL, bm, (int)(best & 31u));
  return (int)best;
}

// k-th (0-based) legal move in ascending order of position b (k < number of empty cells)
__device__ __forceinline__ int wb_kth_legal(const WBoard& b, int k, int W, int H, int lane) {
  const uint32_t e = wb_empty_row(b, W, H, lane);
  const int c = __popc(e);
  int pre = c;
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const int t = __shfl_up_sync(AP_FULL, pre, d);
    if (lane >= d) pre += t;
  }
  const bool mine = (lane < AP_ROWS) && (k >= pre - c) && (k < pre);
  const int h = __ffs(__ballot_sync(AP_FULL, mine)) - 1;
  int w = 0;
  if (mine) w = (int)__fns(e, 0, k - (pre - c) + 1);
  w = __shfl_sync(AP_FULL, w, h);
  return h * W + w;
}

// MCTS.get_move (mcts_pure.py:159-169): tree reset, n_playout x _playout (:114-136), arg-max visits.
// P of the children is not stored (implicit 1/child_count) and children are materialised lazily (see
// pure_select_child); the path of a playout is kept in shared memory so that update_recursive (:61-67) touches all
// its nodes in one memory round trip, one lane each.
template <int MODE>  // 0 = permutation rollouts (W <= 15), 1 = position hash, 2 = ply-by-ply rollouts
__global__ void __launch_bounds__(32, AP_PURE_MINBLK)
k_pure_run(Geo geo, const uint32_t* __restrict__ rows, const BoardMeta* __restrict__ meta, Pools pl, int n_playout,
           unsigned long long seed, int32_t* out_move, int32_t* errflag, unsigned long long* stats,
           const uint8_t* __restrict__ active) {
  if (!active[blockIdx.x]) {  // skipped game (ap_search_set_active): no search, no move
    if (threadIdx.x == 0) out_move[blockIdx.x] = -1;
    return;
  }
  __shared__ int16_t s_list[AP_MAX_S];
  __shared__ __align__(16) uint8_t s_rank[256];
  __shared__ int s_path[AP_MAX_S + 1];
  __shared__ double s_prior[AP_MAX_S + 1];
  const int lane = threadIdx.x;
  const int g = blockIdx.x;
  const size_t base = (size_t)g * geo.cap;
  const WBoard root = wb_load(rows, meta, g, lane);
  const uint32_t inv_w = 65536u / (uint32_t)geo.W + 1u;  // (mv * inv_w) >> 16 == mv / W for mv < 256, W <= 16
  Pcg rng = pcg_seed(seed, (unsigned long long)g);  // MODE 2 only
  for (int a = lane; a <= AP_MAX_S; a += 32) s_prior[a] = __ddiv_rn(1.0, (double)(a > 0 ? a : 1));
  if (lane == 0) tree_write_root(pl, base, g);
  int a_next = 1;  // allocation cursor of this game's pool (node 0 = root)
  __syncwarp();
  // counters keep SURVEY 8(d)'s algorithmic meaning (the reference scans / creates all A children of a node)
  unsigned int scanned = 0, written = 0, pathn = 0, plies_total = 0;  // < 2^32 per game for n_playout <= 10^6
  bool failed = false;
#pragma unroll 1
  for (int it = 0; it < n_playout; ++it) {
    WBoard b = root;
    int node = 0, depth = 0;
    auto play = [&](int mv) {
      const int h = (int)(((uint32_t)mv * inv_w) >> 16), w = mv - h * geo.W;
      if (lane == h) b.row |= (1u << w) << ((b.cur == 2) ? 16 : 0);
      b.nst += 1;
      b.last = mv;
      b.cur = 3 - b.cur;
    };
#pragma unroll 1
    while (true) {
      const int cs = pl.child_start[base + node];
      const int cc = pl.child_count[base + node];
      const int np = pl.N[base + node];
      const int nv = pl.parent[base + node];  // visited (= materialised) children of an expanded node
      if (lane == 0) s_path[depth] = node;
      if (cs < 0) break;
      int mv;
      const int bi = pure_select_child(pl, base, cs, cc, nv, np, geo.c_puct, s_prior, lane, mv);
      scanned += cc;
      if (bi == nv) {
        // first visit of child nv: materialise its record; it is the (unexpanded) leaf of this playout
        mv = wb_kth_legal(b, nv, geo.W, geo.H, lane);
        if (lane == 0) {
          const size_t c = base + cs + nv;
          pl.Q[c] = 0.0;
          pl.N[c] = 0;
          pl.child_start[c] = -1;
          pl.child_count[c] = 0;
          pl.parent[c] = 0;
          pl.move[c] = (int16_t)mv;
          pl.parent[base + node] = nv + 1;
          s_path[depth + 1] = cs + nv;
        }
        play(mv);
        node = cs + nv;
        ++depth;
        break;
      }
      play(mv);
      node = cs + bi;
      ++depth;
    }
    int winner;
    const bool end = (MODE == 0) ? wb_game_end_packed(b, geo.n_in_row, geo.S, winner)
                                 : wb_game_end(b, geo.n_in_row, geo.S, winner);
    if (!end) {
      // TreeNode.expand (mcts_pure.py:34-41): reserve the block of the A = |availables| children
      const int A = geo.S - b.nst;
      if (a_next + A > geo.cap) {
        if (lane == 0) errflag[g] = AP_ERR_POOL_EXHAUSTED;
        failed = true;
        break;
      }
      if (lane == 0) {
        pl.child_start[base + node] = a_next;
        pl.child_count[base + node] = (uint16_t)A;
        pl.parent[base + node] = 0;
      }
      a_next += A;
      written += A;
    }
    int plies = 0;
    int v;
    if constexpr (MODE == 1) {
      v = hash_eval(b, geo);
    } else if constexpr (MODE == 2) {
      v = rollout_eval(b, rng, geo, lane, plies);
    } else {
      if (end) {
        v = (winner == -1) ? 0 : ((winner == b.cur) ? 1 : -1);
      } else {
        uint32_t h[8];
        perm_draws(seed, (uint32_t)g, (uint32_t)lane, (uint32_t)it, h);
        v = rollout_eval_perm(b, h, geo, lane, plies, s_rank);
      }
    }
    plies_total += plies;
    // update_recursive(-leaf_value): the leaf takes -v, its parent +v, ... one lane per path node
    __syncwarp();
#pragma unroll 1
    for (int d0 = 0; d0 <= depth; d0 += 32) {
      const int d = d0 + lane;
      if (d <= depth) {
        const size_t c = base + s_path[d];
        const double x = ((depth - d) & 1) ? (double)v : -(double)v;
        const int n = pl.N[c] + 1;
        const double q = pl.Q[c];
        pl.N[c] = n;
        pl.Q[c] = __dadd_rn(q, __ddiv_rn(__dmul_rn(1.0, __dsub_rn(x, q)), (double)n));
      }
    }
    pathn += depth + 1;
    __syncwarp();
  }
  // complete the root's child block (the unvisited children as fresh records), restore the parent fields the search
  // used as visited counters, publish the allocation cursor
  const int cs = pl.child_start[base];
  const int cc = (cs >= 0) ? pl.child_count[base] : 0;
  const int nv_root = (cs >= 0) ? pl.parent[base] : 0;
  __syncwarp();
  if (cc > 0) {
    wb_legal_list(root, geo.W, geo.H, lane, s_list);
    for (int k = lane; k < cc; k += 32) {
      const size_t c = base + cs + k;
      if (k >= nv_root) {
        pl.Q[c] = 0.0;
        pl.N[c] = 0;
        pl.child_start[c] = -1;
        pl.child_count[c] = 0;
        pl.move[c] = s_list[k];
      }
      pl.parent[c] = 0;
    }
  }
  if (lane == 0) {
    pl.parent[base] = -1;
    pl.alloc[g] = a_next;
  }
  __syncwarp();
  // first max by visit count over root children
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
  (void)failed;
  if (lane == 0) {
    out_move[g] = (bi != INT_MAX) ? (int)pl.move[base + cs + bi] : -1;
    atomicAdd(&stats[0], (unsigned long long)n_playout);
    atomicAdd(&stats[1], (unsigned long long)scanned);
    atomicAdd(&stats[2], (unsigned long long)written);
    atomicAdd(&stats[3], (unsigned long long)pathn);
    atomicAdd(&stats[5], (unsigned long long)plies_total);
  }
}

// keys != nullptr: the 24-bit draw of every board slot comes from the caller ([G][256], slot = row*16 + column)
__global__ void k_rollout_eval(Geo geo, const uint32_t* rows, const BoardMeta* meta, unsigned long long seed, int impl,
                               const uint32_t* __restrict__ keys, int8_t* out_value, int16_t* out_plies) {
  __shared__ __align__(16) uint8_t s_rank[4][256];
  int lane = threadIdx.x & 31;
  int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  WBoard b = wb_load(rows, meta, g, lane);
  Pcg rng = pcg_seed(seed, (unsigned long long)g);
  int plies;
  int v;
  if (impl == 0 && geo.W <= 15) {
    int winner;
    plies = 0;
    if (wb_game_end_packed(b, geo.n_in_row, geo.S, winner)) {
      v = (winner == -1) ? 0 : ((winner == b.cur) ? 1 : -1);
    } else {
      uint32_t h[8];
      if (keys) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          h[r] = keys[(size_t)g * 256 + (lane >> 1) * 16 + rank_slot((uint32_t)((lane & 1) * 8 + r))] & 0xFFFFFFu;
      } else {
        perm_draws(seed, (uint32_t)g, (uint32_t)lane, 0u, h);
      }
      v = rollout_eval_perm(b, h, geo, lane, plies, s_rank[threadIdx.x >> 5]);
    }
  } else {
    v = rollout_eval(b, rng, geo, lane, plies);
  }
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
  if (mode == 0 && e->geo.W > 15) mode = 2;  // the packed two-colour line check needs a spare column bit
  auto k = (mode == 0) ? k_pure_run<0> : (mode == 1) ? k_pure_run<1> : k_pure_run<2>;
  k<<<e->geo.G, 32, 0, e->stream>>>(e->geo, e->rows, e->meta, e->pools, n_playout, seed, d_move, e->errflag, e->stats,
                                  e->leaves.active);
}
void launch_rollout_eval(ap_engine* e, uint64_t seed, int impl, const uint32_t* d_keys, int8_t* d_value, int16_t* d_plies) {
  k_rollout_eval<<<(e->geo.G + 3) / 4, 128, 0, e->stream>>>(e->geo, e->rows, e->meta, seed, impl, d_keys, d_value,
                                                          d_plies);
}
void launch_rollout_hash(ap_engine* e, int8_t* d_value) {
  k_rollout_hash<<<(e->geo.G + 3) / 4, 128, 0, e->stream>>>(e->geo, e->rows, e->meta, d_value);
}
