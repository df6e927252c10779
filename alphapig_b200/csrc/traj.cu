// Self-play trajectories on the device.
// Replaces the per-game Python lists of Game_AI.start_self_play (game_ai.py:75,113-131: states / mcts_probs /
// current_players appended before every move, winners_z computed when the game ends) for thousands of concurrent
// games: every ply appends one packed record (np.packbits of Board.current_state() | pi fp32[S] | z) to the game's
// own trajectory in HBM, and when a game ends its records - z filled in from the winner - move to the OUTBOX, a flat
// device array of fixed-size records that goes to the trainer rank as is (NCCL all-gather of a device tensor) and
// from there into the replay ring (ap_replay_push_packed) without touching the host.
//
// record = [ceil(9S/8) state bytes, zero padded to a multiple of 4][S x fp32 pi][fp32 z]   (1160 B on 15x15)
#include "kernels.h"
#include "state_bits.cuh"

struct TrajState {
  int max_plies = 0;
  int sb = 0, rw = 0, off_pi = 0;
  uint8_t* rec = nullptr;     // [G][max_plies][rw]
  int8_t* player = nullptr;   // [G][max_plies]  player to move when the record was taken
  int32_t* len = nullptr;     // [G]
  uint8_t* outbox = nullptr;  // [out_cap][rw]
  int64_t out_cap = 0, out_n = 0;
  int32_t* d_offs = nullptr;  // [G + 1] exclusive offsets of the finishing games, total in [n]
};

int traj_record_width(int S) { return ((9 * S + 7) / 8 + 3) / 4 * 4 + 4 * S + 4; }

// one warp per game: append (state of the CURRENT board, pi, player to move).  pi == nullptr: forced[i] is the move
// of a forced opening ply (pi = 0.99999 at the move, 1e-6 elsewhere, game_ai.py:87-89).
__global__ void __launch_bounds__(128)
k_traj_append(Geo geo, const uint32_t* __restrict__ rows, const BoardMeta* __restrict__ meta, const int32_t* __restrict__ ids,
              int n, const float* __restrict__ pi, const int32_t* __restrict__ forced, uint8_t* rec, int8_t* player,
              int32_t* len, int max_plies, int sb, int rw, int off_pi, int32_t* errflag) {
  extern __shared__ uint32_t s_words[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + wib;
  if (i >= n) return;
  const int g = ids ? ids[i] : i;
  const int nw = (sb + 3) / 4;
  uint32_t* wbuf = s_words + wib * nw;
  const WBoard b = wb_load(rows, meta, g, lane);
  const int t = len[g];
  if (t >= max_plies) {
    if (lane == 0) errflag[g] = AP_ERR_BAD_ARG;
    return;
  }
  wb_pack_state(b, geo.W, geo.H, geo.S, lane, wbuf, nw);
  uint8_t* out = rec + ((size_t)g * max_plies + t) * rw;
  uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
  for (int k = lane; k < off_pi / 4; k += 32) o32[k] = (k < nw) ? wbuf[k] : 0u;  // bits past 9S are zero in wbuf
  float* op = reinterpret_cast<float*>(out + off_pi);
  if (pi) {
    const float* src = pi + (size_t)g * geo.S;
    for (int k = lane; k < geo.S; k += 32) op[k] = src[k];
  } else {
    const int mv = forced[i];
    for (int k = lane; k < geo.S; k += 32) op[k] = (k == mv) ? 0.99999f : 0.000001f;
  }
  if (lane == 0) {
    op[geo.S] = 0.f;
    player[(size_t)g * max_plies + t] = (int8_t)b.cur;
    len[g] = t + 1;
  }
}

// exclusive scan of the finishing games' lengths (one CTA; n <= G, a few thousand)
__global__ void __launch_bounds__(1024) k_traj_offsets(const int32_t* __restrict__ ids, int n, const int32_t* __restrict__ len,
                                                       int32_t* offs) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const int v = (i < n) ? len[ids[i]] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(AP_FULL, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    int off = s_base;
    for (int k = 0; k < w; ++k) off += s_warp[k];
    if (i < n) offs[i] = off + incl - v;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int k = 0; k < 32; ++k) t += s_warp[k];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0) offs[n] = s_base;
}

// one CTA per finishing game: its records move to the outbox with z = +1 / -1 for the plies of the winner / loser,
// 0 for a tie (game_ai.py:124-128); the trajectory restarts empty.
__global__ void __launch_bounds__(256)
k_traj_flush(const int32_t* __restrict__ ids, const int8_t* __restrict__ winners, const int32_t* __restrict__ offs,
             const uint8_t* __restrict__ rec, const int8_t* __restrict__ player, int32_t* len, int max_plies, int rw,
             uint8_t* outbox, long long out_base) {
  const int i = blockIdx.x;
  const int g = ids[i];
  const int n = len[g];
  const int win = winners[i];
  uint8_t* dst_b = outbox + (size_t)(out_base + offs[i]) * rw;
  // rw is a multiple of 4 only (332 B on 8x8): copy words
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(rec + (size_t)g * max_plies * rw);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(dst_b);
  const int words = n * (rw / 4);
  for (int k = threadIdx.x; k < words; k += blockDim.x) d32[k] = s32[k];
  __syncthreads();
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int p = player[(size_t)g * max_plies + t];
    const float z = (win == -1) ? 0.f : (p == win ? 1.f : -1.f);
    *reinterpret_cast<float*>(dst_b + (size_t)t * rw + rw - 4) = z;
  }
  __syncthreads();
  if (threadIdx.x == 0) len[g] = 0;
}

__global__ void k_traj_discard(const int32_t* __restrict__ ids, int n, int32_t* len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[ids ? ids[i] : i] = 0;
}

void traj_destroy(ap_engine* e) {
  TrajState* t = e->traj;
  if (!t) return;
  cudaFree(t->rec);
  cudaFree(t->player);
  cudaFree(t->len);
  cudaFree(t->outbox);
  cudaFree(t->d_offs);
  delete t;
  e->traj = nullptr;
}

// called by ap_selfplay_pick after the pick kernel: pi is the device buffer the pick just wrote
int traj_append_pick(ap_engine* e, const float* d_pi) {
  TrajState* t = e->traj;
  if (!t) return AP_OK;
  const int wpb = 4, nw = (t->sb + 3) / 4;
  k_traj_append<<<(e->geo.G + wpb - 1) / wpb, 32 * wpb, (size_t)wpb * nw * 4, e->stream>>>(
      e->geo, e->rows, e->meta, nullptr, e->geo.G, d_pi, nullptr, t->rec, t->player, t->len, t->max_plies, t->sb, t->rw,
      t->off_pi, e->errflag);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}

extern "C" {

int ap_traj_create(ap_engine* e, int32_t max_plies, int64_t outbox_records) {
  AP_ENTER(e);
  if (e->geo.W != e->geo.H) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_create: square boards only (record format of the replay ring)");
  traj_destroy(e);
  const int S = e->geo.S;
  if (max_plies <= 0) max_plies = S;
  if (max_plies > S || outbox_records < 1) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_create: bad argument");
  TrajState* t = new TrajState();
  e->traj = t;
  t->max_plies = max_plies;
  t->sb = (9 * S + 7) / 8;
  t->rw = traj_record_width(S);
  t->off_pi = t->rw - 4 * S - 4;
  t->out_cap = outbox_records;
  const size_t G = e->geo.G;
  AP_CUDA(e, cudaMalloc(&t->rec, G * max_plies * t->rw));
  AP_CUDA(e, cudaMalloc(&t->player, G * max_plies));
  AP_CUDA(e, cudaMalloc(&t->len, G * 4));
  AP_CUDA(e, cudaMalloc(&t->outbox, (size_t)outbox_records * t->rw));
  AP_CUDA(e, cudaMalloc(&t->d_offs, (G + 1) * 4));
  AP_CUDA(e, cudaMemsetAsync(t->len, 0, G * 4, e->stream));
  e->bytes += G * max_plies * (t->rw + 1) + (size_t)outbox_records * t->rw;
  return ap_sync(e);
}

int ap_traj_append_forced(ap_engine* e, const int32_t* game_ids, int32_t n, const int32_t* moves) {
  AP_ENTER(e);
  TrajState* t = e->traj;
  if (!t) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_append_forced: no trajectories (ap_traj_create)");
  if (n <= 0) return AP_OK;
  if (!moves) return ap_fail(e, AP_ERR_BAD_ARG, "null argument");
  for (int i = 0; i < n; ++i)
    if (moves[i] < 0 || moves[i] >= e->geo.S) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_append_forced: move out of range");
  int rc = ap_ids(e, game_ids, n);
  if (rc != AP_OK) return rc;
  rc = ap_stage(e, (size_t)n * 4, 0);
  if (rc != AP_OK) return rc;
  AP_CUDA(e, cudaMemcpyAsync(e->d_stage, moves, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
  const int wpb = 4, nw = (t->sb + 3) / 4;
  k_traj_append<<<(n + wpb - 1) / wpb, 32 * wpb, (size_t)wpb * nw * 4, e->stream>>>(
      e->geo, e->rows, e->meta, e->d_ids, n, nullptr, (const int32_t*)e->d_stage, t->rec, t->player, t->len, t->max_plies,
      t->sb, t->rw, t->off_pi, e->errflag);
  AP_LAUNCH_CHECK(e);
  return ap_sync(e);
}

int ap_traj_finish(ap_engine* e, const int32_t* game_ids, int32_t n, const int8_t* winners) {
  AP_ENTER(e);
  TrajState* t = e->traj;
  if (!t) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_finish: no trajectories (ap_traj_create)");
  if (n <= 0) return AP_OK;
  if (!winners) return ap_fail(e, AP_ERR_BAD_ARG, "null argument");
  int rc = ap_ids(e, game_ids, n);
  if (rc != AP_OK) return rc;
  rc = ap_stage(e, (size_t)n, 0);
  if (rc != AP_OK) return rc;
  AP_CUDA(e, cudaMemcpyAsync(e->d_stage, winners, (size_t)n, cudaMemcpyHostToDevice, e->stream));
  k_traj_offsets<<<1, 1024, 0, e->stream>>>(e->d_ids, n, t->len, t->d_offs);
  AP_LAUNCH_CHECK(e);
  int32_t total = 0;
  AP_CUDA(e, cudaMemcpyAsync(&total, t->d_offs + n, 4, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  if (t->out_n + total > t->out_cap)
    return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_finish: outbox full (" + std::to_string(t->out_n) + " + " + std::to_string(total) +
                                          " > " + std::to_string(t->out_cap) + " records): drain it with ap_traj_outbox / _clear");
  k_traj_flush<<<n, 256, 0, e->stream>>>(e->d_ids, (const int8_t*)e->d_stage, t->d_offs, t->rec, t->player, t->len,
                                         t->max_plies, t->rw, t->outbox, (long long)t->out_n);
  AP_LAUNCH_CHECK(e);
  t->out_n += total;
  return ap_sync(e);
}

int ap_traj_discard(ap_engine* e, const int32_t* game_ids, int32_t n) {
  AP_ENTER(e);
  TrajState* t = e->traj;
  if (!t) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_discard: no trajectories (ap_traj_create)");
  if (n <= 0) return AP_OK;
  int rc = ap_ids(e, game_ids, n);
  if (rc != AP_OK) return rc;
  k_traj_discard<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_ids, n, t->len);
  AP_LAUNCH_CHECK(e);
  return ap_sync(e);
}

int ap_traj_outbox(ap_engine* e, void** out_dev_ptr, int64_t* out_records, int32_t* out_record_bytes) {
  AP_ENTER(e);
  TrajState* t = e->traj;
  if (!t) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_outbox: no trajectories (ap_traj_create)");
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  if (out_dev_ptr) *out_dev_ptr = t->outbox;
  if (out_records) *out_records = t->out_n;
  if (out_record_bytes) *out_record_bytes = t->rw;
  return AP_OK;
}

int ap_traj_outbox_clear(ap_engine* e) {
  AP_ENTER(e);
  TrajState* t = e->traj;
  if (!t) return ap_fail(e, AP_ERR_BAD_ARG, "ap_traj_outbox_clear: no trajectories (ap_traj_create)");
  t->out_n = 0;
  return AP_OK;
}

}  // extern "C"
