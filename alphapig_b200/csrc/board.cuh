// Warp-level bitboard for one game: lane h (< H) holds board row h.
//   row bits [0,16)  : stones of player 1, bit w = column w
//   row bits [16,32) : stones of player 2
// Replaces the dict/list state of reference game.Board (game.py:21-170).
#pragma once
#include "ap_common.cuh"

struct WBoard {
  uint32_t row;             // this lane's row (0 for lanes >= H)
  int cur;                  // player to move (warp-uniform)
  int nst;                  // stones on board (warp-uniform)
  int last;                 // last move (warp-uniform)
  unsigned long long hist;  // 4 x int16, most recent in bits [0,16) (warp-uniform)
};

__device__ __forceinline__ unsigned long long pack_hist(const BoardMeta& m) {
  return (unsigned long long)(uint16_t)m.hist[0] | ((unsigned long long)(uint16_t)m.hist[1] << 16) |
         ((unsigned long long)(uint16_t)m.hist[2] << 32) | ((unsigned long long)(uint16_t)m.hist[3] << 48);
}

__device__ __forceinline__ WBoard wb_load(const uint32_t* __restrict__ rows, const BoardMeta* __restrict__ meta, int g,
                                          int lane) {
  WBoard b;
  b.row = (lane < AP_ROWS) ? rows[(size_t)g * AP_ROWS + lane] : 0u;
  BoardMeta m = meta[g];
  b.cur = m.cur;
  b.nst = m.n_stones;
  b.last = m.last_move;
  b.hist = pack_hist(m);
  return b;
}

__device__ __forceinline__ void wb_store(const WBoard& b, uint32_t* rows, BoardMeta* meta, int g, int lane, int start) {
  if (lane < AP_ROWS) rows[(size_t)g * AP_ROWS + lane] = b.row;
  if (lane == 0) {
    BoardMeta m;
    m.hist[0] = (int16_t)(b.hist & 0xffff);
    m.hist[1] = (int16_t)((b.hist >> 16) & 0xffff);
    m.hist[2] = (int16_t)((b.hist >> 32) & 0xffff);
    m.hist[3] = (int16_t)((b.hist >> 48) & 0xffff);
    m.n_stones = (int16_t)b.nst;
    m.last_move = (int16_t)b.last;
    m.cur = (int8_t)b.cur;
    m.start = (int8_t)start;
    m.pad = 0;
    meta[g] = m;
  }
}

// Board.do_move (game.py:117-125); caller guarantees legality.
__device__ __forceinline__ void wb_do_move(WBoard& b, int move, int W, int lane) {
  int h = move / W;
  int w = move - h * W;
  if (lane == h) b.row |= (1u << w) << ((b.cur == 2) ? 16 : 0);
  b.hist = (b.hist << 16) | (unsigned long long)(uint16_t)move;
  b.nst += 1;
  b.last = move;
  b.cur = 3 - b.cur;
}

// empty cells of this lane's row
__device__ __forceinline__ uint32_t wb_empty_row(const WBoard& b, int W, int H, int lane) {
  uint32_t occ = (b.row | (b.row >> 16)) & 0xffffu;
  return (lane < H) ? (~occ & ((1u << W) - 1u)) : 0u;
}

__device__ __forceinline__ bool wb_is_legal(const WBoard& b, int move, int W, int H, int lane) {
  if (move < 0 || move >= W * H) return false;
  int h = move / W, w = move - h * W;
  uint32_t e = wb_empty_row(b, W, H, lane);
  uint32_t eh = __shfl_sync(AP_FULL, e, h);
  return (eh >> w) & 1u;
}

// n-in-a-row lines in one colour's rows x (16 bits per lane): any line of n
// starting at (h,w) going right / up / up-right / up-left (game.py:141-156).
__device__ __forceinline__ bool wb_colour_wins(uint32_t x, int n) {
  uint32_t th = x, tv = x, td = x, ta = x;
  for (int k = 1; k < n; ++k) {
    uint32_t up = __shfl_down_sync(AP_FULL, x, k);
    // lanes whose source lane+k >= 32 get their own value; those lanes are >= 16 and hold 0.
    th &= x >> k;
    tv &= up;
    td &= up >> k;
    ta &= (up << k) & 0xffffu;
  }
  return __any_sync(AP_FULL, (th | tv | td | ta) != 0u);
}

// Board.has_a_winner (game.py:127-158): 0 = none, else winner 1/2.
__device__ __forceinline__ int wb_winner(const WBoard& b, int n) {
  if (b.nst < n + 2) return 0;  // game.py:134-135
  bool w1 = wb_colour_wins(b.row & 0xffffu, n);
  bool w2 = wb_colour_wins(b.row >> 16, n);
  return w1 ? 1 : (w2 ? 2 : 0);
}

// Board.game_end (game.py:160-167): returns end flag, winner in {1,2,-1}.
__device__ __forceinline__ bool wb_game_end(const WBoard& b, int n, int S, int& winner) {
  int w = wb_winner(b, n);
  if (w) {
    winner = w;
    return true;
  }
  winner = -1;
  return b.nst >= S;
}

// Ascending legal-move list into smem (list must hold S entries); returns count.
__device__ __forceinline__ int wb_legal_list(const WBoard& b, int W, int H, int lane, int16_t* list) {
  uint32_t e = wb_empty_row(b, W, H, lane);
  int c = __popc(e);
  int pre = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(AP_FULL, pre, d);
    if (lane >= d) pre += t;
  }
  int total = __shfl_sync(AP_FULL, pre, 31);
  int off = pre - c;
  int base = lane * W;
  while (e) {
    int w = __ffs(e) - 1;
    e &= e - 1;
    list[off++] = (int16_t)(base + w);
  }
  __syncwarp();
  return total;
}

// Stones of (player) with the last `drop` plies removed; returns this lane's 16-bit row.
__device__ __forceinline__ uint32_t wb_rows_dropped(const WBoard& b, int player, int drop, int W, int lane) {
  uint32_t r = b.row;
  for (int i = 0; i < drop; ++i) {
    int16_t mv = (int16_t)((b.hist >> (16 * i)) & 0xffff);
    if (mv >= 0) {
      int h = mv / W, w = mv - h * W;
      if (lane == h) r &= ~((1u << w) | (1u << (w + 16)));
    }
  }
  return (player == 1) ? (r & 0xffffu) : (r >> 16);
}
