// Board.current_state() (game.py:68-94) of a warp-resident board as np.packbits bytes: the record format of the
// replay ring (replay.cu) and of the self-play trajectories (traj.cu).  Square boards (the ring's rot90 augmentation).
#pragma once
#include "board.cuh"

// wbuf: (sb + 3) / 4 shared-memory words of this warp; on return they hold the packed planes (byte i of the
// np.packbits stream = byte i of wbuf, MSB first).
__device__ __forceinline__ void wb_pack_state(const WBoard& b, int W, int H, int S, int lane, uint32_t* wbuf, int nw) {
  for (int i = lane; i < nw; i += 32) wbuf[i] = 0u;
  __syncwarp();
  auto setbit = [&](int f) { atomicOr(&wbuf[f >> 5], 1u << ((((f >> 3) & 3) << 3) + 7 - (f & 7))); };
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const uint32_t own = wb_rows_dropped(b, b.cur, d, W, lane);
    const uint32_t opp = wb_rows_dropped(b, 3 - b.cur, d, W, lane);
    if (lane < H) {
      const int r = W - 1 - lane;  // axis-1 flip of current_state (game.py:94)
      for (int w = 0; w < W; ++w) {
        if ((own >> w) & 1u) setbit((6 - 2 * d) * S + r * W + w);
        if ((opp >> w) & 1u) setbit((7 - 2 * d) * S + r * W + w);
      }
    }
  }
  if (b.nst % 2 == 0)
    for (int k = lane; k < S; k += 32) setbit(8 * S + k);
  __syncwarp();
}
