// Board kernels: one warp per game, lane = board row (see board.cuh).
// Replaces reference game.Board (game.py:21-170) for G games at once.
#include "board.cuh"
#include "kernels.h"

#define WARPS_PER_BLOCK 4

__device__ __forceinline__ int warp_entry(int& lane) {
  lane = threadIdx.x & 31;
  return blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
}

// Board.init_board (game.py:35-44)
__global__ void k_boards_reset(Geo geo, uint32_t* rows, BoardMeta* meta, const int32_t* ids, int n,
                               const int32_t* start_player) {
  int lane, i = warp_entry(lane);
  if (i >= n) return;
  int g = ids[i];
  int sp = start_player ? start_player[i] : 0;
  WBoard b;
  b.row = 0;
  b.cur = sp ? 2 : 1;
  b.nst = 0;
  b.last = -1;
  b.hist = ~0ull;
  wb_store(b, rows, meta, g, lane, sp);
}

// Board.do_move (game.py:117-125) with the legality check that list.remove implies.
__global__ void k_boards_do_move(Geo geo, uint32_t* rows, BoardMeta* meta, const int32_t* ids, const int32_t* moves,
                                 int n, int32_t* status) {
  int lane, i = warp_entry(lane);
  if (i >= n) return;
  int g = ids[i];
  WBoard b = wb_load(rows, meta, g, lane);
  int mv = moves[i];
  bool ok = wb_is_legal(b, mv, geo.W, geo.H, lane);
  if (ok) {
    int start = meta[g].start;
    wb_do_move(b, mv, geo.W, lane);
    wb_store(b, rows, meta, g, lane, start);
  }
  if (lane == 0) status[i] = ok ? AP_OK : AP_ERR_ILLEGAL_MOVE;
}

// Board.game_end (game.py:160-167)
__global__ void k_boards_status(Geo geo, const uint32_t* rows, const BoardMeta* meta, const int32_t* ids, int n,
                                uint8_t* out_end, int8_t* out_winner) {
  int lane, i = warp_entry(lane);
  if (i >= n) return;
  int g = ids ? ids[i] : i;
  WBoard b = wb_load(rows, meta, g, lane);
  int winner;
  bool end = wb_game_end(b, geo.n_in_row, geo.S, winner);
  if (lane == 0) {
    out_end[i] = end ? 1 : 0;
    out_winner[i] = (int8_t)winner;
  }
}

// Board.availables as a 256-bit mask over move indices.
__global__ void k_boards_legal(Geo geo, const uint32_t* rows, const BoardMeta* meta, const int32_t* ids, int n,
                               uint32_t* out_mask) {
  __shared__ int16_t list[WARPS_PER_BLOCK][AP_MAX_S];
  int lane, i = warp_entry(lane);
  if (i >= n) return;
  int g = ids ? ids[i] : i;
  int w = threadIdx.x >> 5;
  WBoard b = wb_load(rows, meta, g, lane);
  int A = wb_legal_list(b, geo.W, geo.H, lane, list[w]);
  uint32_t word = 0;
  if (lane < 8) {
    for (int k = 0; k < A; ++k) {
      int m = list[w][k];
      if ((m >> 5) == lane) word |= 1u << (m & 31);
    }
    out_mask[(size_t)i * 8 + lane] = word;
  }
}

// Board.current_state (game.py:68-94) as float32 [9][W][H], axis-1 flip included.
// rows/meta may be the root boards or the leaf boards of the last select.
__global__ void k_boards_features(Geo geo, const uint32_t* rows, const BoardMeta* meta, const int32_t* ids, int n,
                                  float* out) {
  int lane, i = warp_entry(lane);
  if (i >= n) return;
  int g = ids ? ids[i] : i;
  WBoard b = wb_load(rows, meta, g, lane);
  const int W = geo.W, H = geo.H;
  float* o = out + (size_t)i * 9 * W * H;
  // plane 8: colour to play (all ones iff stone count even)
  float p8 = (b.nst % 2 == 0) ? 1.f : 0.f;
  for (int k = lane; k < W * H; k += 32) o[8 * W * H + k] = p8;
  for (int d = 0; d < 4; ++d) {
    uint32_t own = wb_rows_dropped(b, b.cur, d, W, lane);
    uint32_t opp = wb_rows_dropped(b, 3 - b.cur, d, W, lane);
    // a = m // W = lane (board row), b2 = m % H; output row index W-1-a
    if (lane < H && lane < W) {
      float* po = o + (size_t)(6 - 2 * d) * W * H + (size_t)(W - 1 - lane) * H;
      float* pp = o + (size_t)(7 - 2 * d) * W * H + (size_t)(W - 1 - lane) * H;
      for (int w = 0; w < W; ++w) {
        int b2 = (lane * W + w) % H;
        // several w may alias the same b2 only when W != H; '=' of 1.0 wins as in the reference
        if ((own >> w) & 1u) po[b2] = 1.f;
        if ((opp >> w) & 1u) pp[b2] = 1.f;
      }
    }
  }
}

// np.packbits (MSB first) of the 0/1 feature planes: one thread per output byte
__global__ void k_pack_bits(const float* __restrict__ f, int n, int nbits, int sb, uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * sb) return;
  const int g = (int)(i / sb), b = (int)(i % sb);
  const float* src = f + (size_t)g * nbits + (size_t)b * 8;
  unsigned v = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (b * 8 + k < nbits && src[k] != 0.f) v |= 0x80u >> k;
  out[i] = (uint8_t)v;
}

__global__ void k_fill_f32(float* p, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// Board.states / history / current_player / last_move
__global__ void k_boards_export(Geo geo, const uint32_t* rows, const BoardMeta* meta, const int32_t* ids, int n,
                                int8_t* cells, int32_t* ometa) {
  int lane, i = warp_entry(lane);
  if (i >= n) return;
  int g = ids ? ids[i] : i;
  WBoard b = wb_load(rows, meta, g, lane);
  for (int h = 0; h < geo.H; ++h) {
    uint32_t r = __shfl_sync(AP_FULL, b.row, h);
    if (lane < geo.W) {
      int v = ((r >> lane) & 1u) ? 1 : (((r >> (lane + 16)) & 1u) ? 2 : 0);
      cells[(size_t)i * geo.S + h * geo.W + lane] = (int8_t)v;
    }
  }
  if (lane == 0) {
    BoardMeta m = meta[g];
    int32_t* o = ometa + (size_t)i * AP_META_INTS;
    o[0] = m.cur;
    o[1] = m.last_move;
    o[2] = m.n_stones;
    o[3] = m.hist[0];
    o[4] = m.hist[1];
    o[5] = m.hist[2];
    o[6] = m.hist[3];
    o[7] = m.start;
  }
}

__global__ void k_boards_import(Geo geo, uint32_t* rows, BoardMeta* meta, const int32_t* ids, int n,
                                const int8_t* cells, const int32_t* imeta) {
  int lane, i = warp_entry(lane);
  if (i >= n) return;
  int g = ids[i];
  uint32_t r = 0;
  if (lane < geo.H) {
    for (int w = 0; w < geo.W; ++w) {
      int v = cells[(size_t)i * geo.S + lane * geo.W + w];
      if (v == 1) r |= 1u << w;
      if (v == 2) r |= 1u << (w + 16);
    }
  }
  if (lane < AP_ROWS) rows[(size_t)g * AP_ROWS + lane] = r;
  if (lane == 0) {
    const int32_t* o = imeta + (size_t)i * AP_META_INTS;
    BoardMeta m;
    m.cur = (int8_t)o[0];
    m.last_move = (int16_t)o[1];
    m.n_stones = (int16_t)o[2];
    m.hist[0] = (int16_t)o[3];
    m.hist[1] = (int16_t)o[4];
    m.hist[2] = (int16_t)o[5];
    m.hist[3] = (int16_t)o[6];
    m.start = (int8_t)o[7];
    m.pad = 0;
    meta[g] = m;
  }
}

static inline dim3 warp_grid(int n) { return dim3((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK); }

void launch_boards_reset(ap_engine* e, int n, const int32_t* d_start) {
  k_boards_reset<<<warp_grid(n), 32 * WARPS_PER_BLOCK, 0, e->stream>>>(e->geo, e->rows, e->meta, e->d_ids, n, d_start);
}
void launch_boards_do_move(ap_engine* e, int n, const int32_t* d_moves, int32_t* d_status) {
  k_boards_do_move<<<warp_grid(n), 32 * WARPS_PER_BLOCK, 0, e->stream>>>(e->geo, e->rows, e->meta, e->d_ids, d_moves, n,
                                                                        d_status);
}
void launch_boards_status(ap_engine* e, const uint32_t* rows, const BoardMeta* meta, const int32_t* d_ids, int n,
                          uint8_t* d_end, int8_t* d_winner) {
  k_boards_status<<<warp_grid(n), 32 * WARPS_PER_BLOCK, 0, e->stream>>>(e->geo, rows, meta, d_ids, n, d_end, d_winner);
}
void launch_boards_legal(ap_engine* e, const int32_t* d_ids, int n, uint32_t* d_mask) {
  k_boards_legal<<<warp_grid(n), 32 * WARPS_PER_BLOCK, 0, e->stream>>>(e->geo, e->rows, e->meta, d_ids, n, d_mask);
}
void launch_boards_features(ap_engine* e, const uint32_t* rows, const BoardMeta* meta, const int32_t* d_ids, int n,
                            float* d_out) {
  size_t tot = (size_t)n * 9 * e->geo.S;
  k_fill_f32<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(d_out, tot, 0.f);
  k_boards_features<<<warp_grid(n), 32 * WARPS_PER_BLOCK, 0, e->stream>>>(e->geo, rows, meta, d_ids, n, d_out);
}
void launch_pack_bits(ap_engine* e, const float* d_f, int n, int nbits, uint8_t* d_out) {
  const int sb = (nbits + 7) / 8;
  const size_t tot = (size_t)n * sb;
  k_pack_bits<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(d_f, n, nbits, sb, d_out);
}
void launch_boards_export(ap_engine* e, const uint32_t* rows, const BoardMeta* meta, const int32_t* d_ids, int n,
                          int8_t* d_cells, int32_t* d_meta) {
  k_boards_export<<<warp_grid(n), 32 * WARPS_PER_BLOCK, 0, e->stream>>>(e->geo, rows, meta, d_ids, n, d_cells, d_meta);
}
void launch_boards_import(ap_engine* e, int n, const int8_t* d_cells, const int32_t* d_meta) {
  k_boards_import<<<warp_grid(n), 32 * WARPS_PER_BLOCK, 0, e->stream>>>(e->geo, e->rows, e->meta, e->d_ids, n, d_cells,
                                                                       d_meta);
}
