// Device-resident replay ring with on-the-fly 8-fold symmetry augmentation.
// Replaces TrainPipeline.data_buffer = deque(maxlen) + get_equi_data + random.sample of the reference
// (train_mxnet.py:57,115-135,153,180,196-199): the reference stores all 8 rotated/flipped copies of
// every (state, pi, z) sample in a Python deque; here ONE packed record per position lives in HBM
// (ceil(9S/8) state bytes + S fp32 pi + z) and the symmetry is applied by the gather kernel when a
// minibatch is drawn - identical samples, 1/8 of the memory, no host copies on the way to train_step.
//
// Deque bookkeeping: logical sample a (counted since creation) = record a/8 under symmetry a%8, in the
// reference's order (for i in 1..4: rot90^i, rot90^i + fliplr).  The deque holds the last
// min(total, maxlen) logical samples; index j of the deque is logical sample total - len + j.
#include "kernels.h"
#include "state_bits.cuh"

struct ReplayState {
  int64_t maxlen = 0;      // deque maxlen, in augmented samples
  int64_t total = 0;       // augmented samples ever appended
  int cap = 0;             // records in the ring
  int sb = 0;              // state bytes per record
  uint8_t* bits = nullptr; // [cap][sb]   np.packbits order (MSB first) of the (9,H,W) planes
  float* pi = nullptr;     // [cap][S]
  float* z = nullptr;      // [cap]
  int16_t* perm = nullptr; // [2][8][S]: source index of every output cell, states then pi
  int64_t* d_idx = nullptr;
  int idx_cap = 0;
};

__global__ void k_replay_gather(const uint8_t* __restrict__ bits, const float* __restrict__ pi, const float* __restrict__ z,
                                const int16_t* __restrict__ perm, const int64_t* __restrict__ logical, int S, int sb,
                                int cap, float* __restrict__ out_states, float* __restrict__ out_pi,
                                float* __restrict__ out_z) {
  const int b = blockIdx.x;
  const int64_t a = logical[b];
  const int slot = (int)((a >> 3) % cap), sym = (int)(a & 7);
  const uint8_t* rb = bits + (size_t)slot * sb;
  const int16_t* ps = perm + sym * S;
  const int16_t* pp = perm + (8 + sym) * S;
  for (int i = threadIdx.x; i < 9 * S; i += blockDim.x) {
    const int c = i / S, cell = i - c * S;
    const int f = c * S + ps[cell];
    out_states[(size_t)b * 9 * S + i] = (float)((rb[f >> 3] >> (7 - (f & 7))) & 1);
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) out_pi[(size_t)b * S + i] = pi[(size_t)slot * S + pp[i]];
  if (threadIdx.x == 0) out_z[b] = z[slot];
}

// SGF bootstrap (Game.start_self_play, game.py:233-304) for a batch of recorded games: one warp replays one game
// on a register bitboard and writes one record per ply - Board.current_state() bit-packed, pi = 0.99999 at the
// recorded move / 1e-6 elsewhere, z = +-1 from the recorded winner - straight into the ring.
// emit == 0: validation pass only (status[g] = 1 if a move is illegal: the reference returns warning=1, no data).
__global__ void k_replay_sgf(Geo geo, const int16_t* __restrict__ moves, int max_len, const int32_t* __restrict__ lengths,
                             const int8_t* __restrict__ winners, const int64_t* __restrict__ rec_base, int n_games, int emit,
                             uint8_t* __restrict__ status, uint8_t* bits, float* pi, float* z, int cap, int sb) {
  extern __shared__ uint32_t s_words[];  // per warp: (sb + 3) / 4 words
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  if (g >= n_games) return;
  const int W = geo.W, H = geo.H, S = geo.S;
  const int nw = (sb + 3) / 4;
  uint32_t* wbuf = s_words + wib * nw;
  WBoard b;
  b.row = 0u;
  b.cur = 1;  // init_board(): players[0] moves first
  b.nst = 0;
  b.last = -1;
  b.hist = ~0ull;
  const int len = lengths[g];
  const int win = winners[g];
  if (emit && status[g]) return;
  for (int ply = 0; ply < len; ++ply) {
    const int move = moves[(size_t)g * max_len + ply];
    if (!wb_is_legal(b, move, W, H, lane)) {
      if (lane == 0) status[g] = 1;
      return;
    }
    if (emit) {
      const int slot = (int)((rec_base[g] + ply) % cap);
      wb_pack_state(b, W, H, S, lane, wbuf, nw);
      uint8_t* ob = bits + (size_t)slot * sb;
      const uint8_t* wb8 = reinterpret_cast<const uint8_t*>(wbuf);
      for (int i = lane; i < sb; i += 32) ob[i] = wb8[i];
      float* op = pi + (size_t)slot * S;
      for (int k = lane; k < S; k += 32) op[k] = (k == move) ? 0.99999f : 0.000001f;
      if (lane == 0) z[slot] = (win == -1) ? 0.f : (b.cur == win ? 1.f : -1.f);
      __syncwarp();
    }
    wb_do_move(b, move, W, lane);
  }
}

static void replay_free(ap_engine* e) {
  ReplayState* r = e->replay;
  if (!r) return;
  cudaFree(r->bits);
  cudaFree(r->pi);
  cudaFree(r->z);
  cudaFree(r->perm);
  cudaFree(r->d_idx);
  delete r;
  e->replay = nullptr;
}

void replay_destroy(ap_engine* e) { replay_free(e); }

// rot90 (counterclockwise, numpy) applied k times then optional fliplr: source coordinates of out[r][c]
static void sym_src(int n, int k, int flip, int r, int c, int* sr, int* sc) {
  if (flip) c = n - 1 - c;           // out = fliplr(R): out[r][c] = R[r][n-1-c]
  for (int t = 0; t < k; ++t) {      // R = rot90(M): R[r][c] = M[c][n-1-r]
    const int nr = c, nc = n - 1 - r;
    r = nr;
    c = nc;
  }
  *sr = r;
  *sc = c;
}

extern "C" int ap_replay_create(ap_engine* e, int64_t maxlen) {
  AP_ENTER(e);
  if (maxlen < 1) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_create: maxlen must be >= 1");
  if (e->geo.W != e->geo.H)
    return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_create: the rot90 augmentation needs a square board (train_mxnet.py:122-126)");
  replay_free(e);
  ReplayState* r = new ReplayState();
  e->replay = r;
  const int S = e->geo.S, n = e->geo.W;
  r->maxlen = maxlen;
  r->cap = (int)((maxlen + 7) / 8 + 1);
  r->sb = (9 * S + 7) / 8;
  AP_CUDA(e, cudaMalloc(&r->bits, (size_t)r->cap * r->sb));
  AP_CUDA(e, cudaMalloc(&r->pi, (size_t)r->cap * S * 4));
  AP_CUDA(e, cudaMalloc(&r->z, (size_t)r->cap * 4));
  AP_CUDA(e, cudaMalloc(&r->perm, (size_t)16 * S * 2));
  e->bytes += (size_t)r->cap * (r->sb + S * 4 + 4);
  std::vector<int16_t> perm((size_t)16 * S);
  for (int sym = 0; sym < 8; ++sym) {
    const int k = sym / 2 + 1, flip = sym & 1;
    for (int rr = 0; rr < n; ++rr)
      for (int cc = 0; cc < n; ++cc) {
        int sr, sc;
        // states: equi_state = [fliplr](rot90(s, k))
        sym_src(n, k, flip, rr, cc, &sr, &sc);
        perm[(size_t)sym * S + rr * n + cc] = (int16_t)(sr * n + sc);
        // pi: out = flipud(T(flipud(P)))  =>  out[r][c] = T(F)[n-1-r][c],  F[a][b] = P[n-1-a][b]
        sym_src(n, k, flip, n - 1 - rr, cc, &sr, &sc);
        perm[(size_t)(8 + sym) * S + rr * n + cc] = (int16_t)((n - 1 - sr) * n + sc);
      }
  }
  AP_CUDA(e, cudaMemcpyAsync(r->perm, perm.data(), perm.size() * 2, cudaMemcpyHostToDevice, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

extern "C" int ap_replay_push(ap_engine* e, const uint8_t* state_bits, const float* pi, const float* z, int32_t n) {
  AP_ENTER(e);
  ReplayState* r = e->replay;
  if (!r) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push: no replay ring (ap_replay_create)");
  if (n < 0 || (n && (!state_bits || !pi || !z))) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push: bad argument");
  const int S = e->geo.S;
  for (int i = 0; i < n;) {
    const int slot = (int)((r->total / 8) % r->cap);
    int run = n - i;
    if (run > r->cap - slot) run = r->cap - slot;  // contiguous run up to the ring's end
    AP_CUDA(e, cudaMemcpyAsync(r->bits + (size_t)slot * r->sb, state_bits + (size_t)i * r->sb, (size_t)run * r->sb,
                               cudaMemcpyHostToDevice, e->stream));
    AP_CUDA(e, cudaMemcpyAsync(r->pi + (size_t)slot * S, pi + (size_t)i * S, (size_t)run * S * 4, cudaMemcpyHostToDevice,
                               e->stream));
    AP_CUDA(e, cudaMemcpyAsync(r->z + slot, z + i, (size_t)run * 4, cudaMemcpyHostToDevice, e->stream));
    r->total += 8ll * run;
    i += run;
  }
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

// packed records ([state bytes padded to 4][S x fp32 pi][fp32 z], the outbox / exchange format of traj.cu and
// alphapig_b200/dist.py) -> ring slots
__global__ void k_replay_unpack(const uint8_t* __restrict__ recs, int rw, int off_pi, int S, int sb, int cap, long long slot0,
                                uint8_t* bits, float* pi, float* z) {
  const int i = blockIdx.x;
  const int slot = (int)((slot0 + i) % cap);
  const uint8_t* r = recs + (size_t)i * rw;
  for (int k = threadIdx.x; k < sb; k += blockDim.x) bits[(size_t)slot * sb + k] = r[k];
  const float* rp = reinterpret_cast<const float*>(r + off_pi);
  for (int k = threadIdx.x; k < S; k += blockDim.x) pi[(size_t)slot * S + k] = rp[k];
  if (threadIdx.x == 0) z[slot] = rp[S];
}

extern "C" int ap_replay_push_packed(ap_engine* e, const void* records, int64_t n, int32_t on_device) {
  AP_ENTER(e);
  ReplayState* r = e->replay;
  if (!r) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push_packed: no replay ring (ap_replay_create)");
  if (n < 0 || (n && !records)) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push_packed: bad argument");
  const int S = e->geo.S, rw = traj_record_width(S), off_pi = rw - 4 * S - 4;
  const uint8_t* src = (const uint8_t*)records;
  for (int64_t i = 0; i < n;) {
    int64_t run = n - i;
    if (run > r->cap - 1) run = r->cap - 1;  // slots of one launch must not alias
    const uint8_t* d = src + (size_t)i * rw;
    if (!on_device) {
      int rc = ap_stage(e, (size_t)run * rw, 0);
      if (rc != AP_OK) return rc;
      AP_CUDA(e, cudaMemcpyAsync(e->d_stage, d, (size_t)run * rw, cudaMemcpyHostToDevice, e->stream));
      d = (const uint8_t*)e->d_stage;
    }
    k_replay_unpack<<<(unsigned)run, 128, 0, e->stream>>>(d, rw, off_pi, S, r->sb, r->cap, (long long)(r->total / 8), r->bits,
                                                         r->pi, r->z);
    AP_LAUNCH_CHECK(e);
    AP_CUDA(e, cudaStreamSynchronize(e->stream));
    r->total += 8ll * run;
    i += run;
  }
  return AP_OK;
}

extern "C" int ap_replay_size(ap_engine* e, int64_t* out_len, int64_t* out_total) {
  AP_ENTER(e);
  ReplayState* r = e->replay;
  if (!r) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_size: no replay ring (ap_replay_create)");
  if (out_len) *out_len = r->total < r->maxlen ? r->total : r->maxlen;
  if (out_total) *out_total = r->total;
  return AP_OK;
}

extern "C" int ap_replay_gather(ap_engine* e, const int64_t* idx, int32_t B, float* out_states, float* out_pi, float* out_z,
                                int32_t out_on_device) {
  AP_ENTER(e);
  ReplayState* r = e->replay;
  if (!r) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_gather: no replay ring (ap_replay_create)");
  if (B <= 0 || !idx || !out_states || !out_pi || !out_z) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_gather: bad argument");
  const int S = e->geo.S;
  const int64_t len = r->total < r->maxlen ? r->total : r->maxlen;
  std::vector<int64_t> logical(B);
  for (int i = 0; i < B; ++i) {
    if (idx[i] < 0 || idx[i] >= len) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_gather: index out of range");
    logical[i] = r->total - len + idx[i];
  }
  if (B > r->idx_cap) {
    cudaFree(r->d_idx);
    r->d_idx = nullptr;
    AP_CUDA(e, cudaMalloc(&r->d_idx, (size_t)B * 8));
    r->idx_cap = B;
  }
  AP_CUDA(e, cudaMemcpyAsync(r->d_idx, logical.data(), (size_t)B * 8, cudaMemcpyHostToDevice, e->stream));
  float *ds = out_states, *dp = out_pi, *dz = out_z;
  const size_t ns = (size_t)B * 9 * S, np_ = (size_t)B * S;
  if (!out_on_device) {
    int rc = ap_stage(e, (ns + np_ + B) * 4, 0);
    if (rc != AP_OK) return rc;
    ds = (float*)e->d_stage;
    dp = ds + ns;
    dz = dp + np_;
  }
  k_replay_gather<<<B, 256, 0, e->stream>>>(r->bits, r->pi, r->z, r->perm, r->d_idx, S, r->sb, r->cap, ds, dp, dz);
  AP_LAUNCH_CHECK(e);
  if (!out_on_device) {
    AP_CUDA(e, cudaMemcpyAsync(out_states, ds, ns * 4, cudaMemcpyDeviceToHost, e->stream));
    AP_CUDA(e, cudaMemcpyAsync(out_pi, dp, np_ * 4, cudaMemcpyDeviceToHost, e->stream));
    AP_CUDA(e, cudaMemcpyAsync(out_z, dz, (size_t)B * 4, cudaMemcpyDeviceToHost, e->stream));
  }
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

extern "C" int ap_replay_push_sgf(ap_engine* e, const int16_t* moves, int32_t max_len, const int32_t* lengths,
                                  const int8_t* winners, int32_t n_games, uint8_t* out_warning) {
  AP_ENTER(e);
  ReplayState* r = e->replay;
  if (!r) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push_sgf: no replay ring (ap_replay_create)");
  if (n_games <= 0 || max_len <= 0 || !moves || !lengths || !winners)
    return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push_sgf: bad argument");
  for (int g = 0; g < n_games; ++g)
    if (lengths[g] < 0 || lengths[g] > max_len) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push_sgf: bad length");
  const size_t mb = (size_t)n_games * max_len * 2, lb = (size_t)n_games * 4, wbytes = (size_t)n_games, bb = (size_t)n_games * 8;
  const size_t o_len = (mb + 15) & ~(size_t)15, o_win = o_len + ((lb + 15) & ~(size_t)15), o_st = o_win + ((wbytes + 15) & ~(size_t)15),
               o_base = o_st + ((wbytes + 15) & ~(size_t)15);
  int rc = ap_stage(e, o_base + bb, 0);
  if (rc != AP_OK) return rc;
  char* d = (char*)e->d_stage;
  int16_t* d_moves = (int16_t*)d;
  int32_t* d_len = (int32_t*)(d + o_len);
  int8_t* d_win = (int8_t*)(d + o_win);
  uint8_t* d_status = (uint8_t*)(d + o_st);
  int64_t* d_base = (int64_t*)(d + o_base);
  AP_CUDA(e, cudaMemcpyAsync(d_moves, moves, mb, cudaMemcpyHostToDevice, e->stream));
  AP_CUDA(e, cudaMemcpyAsync(d_len, lengths, lb, cudaMemcpyHostToDevice, e->stream));
  AP_CUDA(e, cudaMemcpyAsync(d_win, winners, wbytes, cudaMemcpyHostToDevice, e->stream));
  AP_CUDA(e, cudaMemsetAsync(d_status, 0, wbytes, e->stream));
  const int wpb = 4, nw = (r->sb + 3) / 4;
  const size_t smem = (size_t)wpb * nw * 4;
  const int grid = (n_games + wpb - 1) / wpb;
  // pass 1: legality of every recorded move
  k_replay_sgf<<<grid, 32 * wpb, smem, e->stream>>>(e->geo, d_moves, max_len, d_len, d_win, d_base, n_games, 0, d_status,
                                                   r->bits, r->pi, r->z, r->cap, r->sb);
  AP_LAUNCH_CHECK(e);
  std::vector<uint8_t> status(n_games);
  AP_CUDA(e, cudaMemcpyAsync(status.data(), d_status, wbytes, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  // pass 2: emit the valid games in order, never more than the ring holds per launch (slots must not alias)
  std::vector<int64_t> base(n_games, 0);
  int g0 = 0;
  while (g0 < n_games) {
    int64_t recs = 0;
    int g1 = g0;
    std::vector<uint8_t> st2(status);
    while (g1 < n_games) {
      const int64_t add = status[g1] ? 0 : lengths[g1];
      if (recs + add > r->cap - 1 && g1 > g0) break;
      if (add > r->cap - 1) return ap_fail(e, AP_ERR_BAD_ARG, "ap_replay_push_sgf: a single game exceeds the ring");
      base[g1] = r->total / 8 + recs;
      recs += add;
      ++g1;
    }
    // games outside [g0, g1) are masked out of this launch
    for (int g = 0; g < n_games; ++g) st2[g] = (g < g0 || g >= g1) ? 1 : status[g];
    AP_CUDA(e, cudaMemcpyAsync(d_status, st2.data(), wbytes, cudaMemcpyHostToDevice, e->stream));
    AP_CUDA(e, cudaMemcpyAsync(d_base, base.data(), bb, cudaMemcpyHostToDevice, e->stream));
    k_replay_sgf<<<grid, 32 * wpb, smem, e->stream>>>(e->geo, d_moves, max_len, d_len, d_win, d_base, n_games, 1, d_status,
                                                     r->bits, r->pi, r->z, r->cap, r->sb);
    AP_LAUNCH_CHECK(e);
    AP_CUDA(e, cudaStreamSynchronize(e->stream));
    r->total += 8 * recs;
    g0 = g1;
  }
  if (out_warning)
    for (int g = 0; g < n_games; ++g) out_warning[g] = status[g];
  return AP_OK;
}
