// Board.current_state (game.py:68-94) of one board as fp16 planes of the net's input layout, one warp per board.
// Shared by k_emit_features (net.cu) and k_select (tree.cu: the leaf board is still in registers there).
#pragma once
#include <cuda_fp16.h>

#include "board.cuh"
#include "net.h"

// Lane h holds the 8 stone planes of board row h (bitboards with the last i plies dropped); the 16-byte pixel
// records are then written a tensor row at a time: lanes 0-15 = the 16 pixels of the row in the channel-0..7
// plane, lanes 16-31 = the same pixels in the channel-8..15 plane (two contiguous 256-byte runs per store
// instruction instead of 16 runs of 16 bytes).  feat: [2][mpad][8] fp16; tile = net tile of this board.
__device__ __forceinline__ void emit_features_warp(const WBoard& wb, int W, int H, int tile, __half* feat, long long mpad,
                                                   int lane) {
  // planes 2c, 2c+1 packed as (lo16, hi16): pk[3 - d] = (mine, theirs) with the last d plies dropped
  uint32_t pk[4];
#pragma unroll
  for (int d = 0; d < 4; ++d)
    pk[3 - d] = wb_rows_dropped(wb, wb.cur, d, W, lane) | (wb_rows_dropped(wb, 3 - wb.cur, d, W, lane) << 16);
  const uint32_t p8 = (wb.nst % 2 == 0) ? 0x3C00u : 0u;  // fp16 1.0 in channel 8
  const int x = lane & 15, grp = lane >> 4;
  __half* dst = feat + ((long long)grp * mpad + NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS + x) * 8;
  for (int y = 0; y < H; ++y) {
    const int src = H - 1 - y;  // axis-1 flip
    uint4 v;
    uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t r = __shfl_sync(AP_FULL, pk[c], src) >> x;
      vv[c] = ((r & 1u) ? 0x3C00u : 0u) | ((r & 0x10000u) ? 0x3C000000u : 0u);
    }
    if (grp) v = make_uint4(p8, 0u, 0u, 0u);
    if (x < W) *reinterpret_cast<uint4*>(dst + (long long)y * 16 * 8) = v;
  }
}
