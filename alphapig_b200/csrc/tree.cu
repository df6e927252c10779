// Lock-step MCTS kernels for G concurrent games (one warp per game; re-root: one CTA per game).
// Replaces reference MCTS._playout / get_move_probs / update_with_move (mcts_alphaZero.py:108-167).
#include "fc_finish.cuh"
#include "features.cuh"
#include "kernels.h"
#include "tree.cuh"

#define SEL_WARPS 4

__global__ void k_tree_reset_all(Geo geo, Pools pl) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= geo.G) return;
  tree_write_root(pl, (size_t)g * geo.cap, g);
}

// MCTS._playout lines 113-121 + the game_end() of :126 for every game.
// 7 CTAs x 4 warps per SM: 4096 games are resident at once on 148 SMs (one wave of dependent-load chains)
// compact != 0: the non-terminal leaves also take a net tile each (slot / game_of_slot / n_eval, the batch the
// net evaluates); tiles are handed out by an atomic ticket - the order is arbitrary, every board's result is
// independent of the tile it sits in.  n_eval is zeroed by k_expand_backup (and before the first lock-step).
#ifndef AP_SEL_MINBLK
#define AP_SEL_MINBLK 7
#endif
__global__ void __launch_bounds__(32 * SEL_WARPS, AP_SEL_MINBLK)
k_select(Geo geo, const uint32_t* __restrict__ rows, const BoardMeta* __restrict__ meta, Pools pl, Leaves lv,
         unsigned long long* stats, int compact, __half* feat, long long mpad) {
  int lane = threadIdx.x & 31;
  int g = blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  if (!lv.active[g]) {  // skipped game: no leaf, no net tile, nothing for k_expand_backup to do
    if (lane == 0) {
      lv.node[g] = -1;
      lv.terminal[g] = 1;
      lv.winner[g] = -1;
      lv.depth[g] = 0;
      lv.slot[g] = -1;
    }
    return;
  }
  size_t base = (size_t)g * geo.cap;
  WBoard b = wb_load(rows, meta, g, lane);
  int node = 0, depth = 0;
  unsigned long long scanned = 0;
  while (true) {
    // the node fields travel together (one round trip): child block, visit count and the move that led here
    const int cs = pl.child_start[base + node];
    const int cc = pl.child_count[base + node];
    const int np = pl.N[base + node];
    const int mv = pl.move[base + node];
    if (depth > 0) {
      if (lane == 0) lv.path[(size_t)g * geo.S + depth - 1] = (int16_t)mv;
      wb_do_move(b, mv, geo.W, lane);
    }
    if (cs < 0) break;
    node = cs + tree_select_child(pl, base, cs, cc, np, geo.c_puct, lane);
    ++depth;
    scanned += cc;
  }
  int winner;
  bool end = wb_game_end(b, geo.n_in_row, geo.S, winner);
  wb_store(b, lv.rows, lv.meta, g, lane, 0);
  if (lane == 0) {
    lv.node[g] = node;
    lv.terminal[g] = end ? 1 : 0;
    lv.winner[g] = (int8_t)winner;
    lv.depth[g] = depth;
    atomicAdd(&stats[0], 1ull);
    atomicAdd(&stats[1], scanned);
    atomicAdd(&stats[3], (unsigned long long)(depth + 1));
    if (end) atomicAdd(&stats[4], 1ull);
  }
  if (compact) {
    int sl = -1;
    if (lane == 0) {
      if (!end) {
        sl = atomicAdd(lv.n_eval, 1);
        lv.game_of_slot[sl] = g;
      }
      lv.slot[g] = sl;
    }
    // the leaf board is still in registers: write its net input planes straight into its tile (no features kernel)
    sl = __shfl_sync(AP_FULL, sl, 0);
    if (feat && sl >= 0) emit_features_warp(b, geo.W, geo.H, sl, feat, mpad, lane);
  }
}

// MCTS._playout lines 126-139: expand (unless terminal), terminal value override, update_recursive(-v).
// Children either from an explicit (acts, priors) list or, dense mode, = legal moves ascending.
__global__ void __launch_bounds__(32 * SEL_WARPS)
k_expand_backup(Geo geo, Pools pl, Leaves lv, const int32_t* __restrict__ counts, const int16_t* __restrict__ acts,
                const double* __restrict__ pri64, const double* __restrict__ val64, const float* __restrict__ pri32,
                const float* __restrict__ val32, const int32_t* __restrict__ slot, int32_t* errflag,
                unsigned long long* stats, FcFinish fin) {
  __shared__ int16_t s_list[SEL_WARPS][AP_MAX_S];
  __shared__ float s_prob[SEL_WARPS][AP_MAX_S];
  int lane = threadIdx.x & 31;
  int w = threadIdx.x >> 5;
  int g = blockIdx.x * SEL_WARPS + w;
  if (g >= geo.G) return;
  size_t base = (size_t)g * geo.cap;
  int leaf = lv.node[g];
  if (leaf < 0) {  // skipped game (ap_search_set_active)
    if (slot && g == 0 && lane == 0) *lv.n_eval = 0;
    return;
  }
  double v;
  float fused_v = 0.f;
  if (!lv.terminal[g]) {
    int A;
    const int src = slot ? slot[g] : g;  // compacted net batch: row of this game's priors / value
    const size_t row = (size_t)src * geo.S;
    bool ok;
    if (acts) {
      A = counts[g];
      for (int k = lane; k < A; k += 32) s_list[w][k] = acts[row + k];
      __syncwarp();
      ok = tree_expand(pl, base, g, geo.cap, leaf, A, s_list[w],
                       [&](int k, int mv) { return pri64[row + k]; }, lane);
    } else {
      WBoard b = wb_load(lv.rows, lv.meta, g, lane);
      A = wb_legal_list(b, geo.W, geo.H, lane, s_list[w]);
      if (fin.partial) {
        // fused lock-step: softmax / tanh of the split-K FC partial logits of this game's net tile, priors from
        // shared memory (same arithmetic as k_head_fc_finish, no probability matrix in HBM)
        float pr[8];
        fused_v = fc_finish_warp(fin, src, geo.S, lane, pr);
#pragma unroll
        for (int j = 0; j < 8; ++j) s_prob[w][lane + 32 * j] = pr[j];
        __syncwarp();
        ok = tree_expand(pl, base, g, geo.cap, leaf, A, s_list[w],
                         [&](int k, int mv) { return (double)s_prob[w][mv]; }, lane);
      } else if (pri64)
        ok = tree_expand(pl, base, g, geo.cap, leaf, A, s_list[w],
                         [&](int k, int mv) { return pri64[row + mv]; }, lane);
      else
        ok = tree_expand(pl, base, g, geo.cap, leaf, A, s_list[w],
                         [&](int k, int mv) { return (double)pri32[row + mv]; }, lane);
    }
    if (!ok && lane == 0) errflag[g] = AP_ERR_POOL_EXHAUSTED;
    if (ok && lane == 0) atomicAdd(&stats[2], (unsigned long long)A);
    v = (fin.partial && !acts) ? (double)fused_v : (val64 ? val64[src] : (double)val32[src]);
  } else {
    int winner = lv.winner[g];
    int cur = lv.meta[g].cur;
    v = (winner == -1) ? 0.0 : ((winner == cur) ? 1.0 : -1.0);  // mcts_alphaZero.py:131-136
  }
  if (lane == 0) tree_backup(pl, base, leaf, -v);
  if (slot && g == 0 && lane == 0) *lv.n_eval = 0;  // next lock-step's ticket counter (nobody reads it in this kernel)
}

// ---------------------------------------------------------------------------------------------
// Opt-in multi-leaf mode (ap_search_run_vl): up to k playouts of a game in flight per lock-step, kept apart by
// VIRTUAL LOSS.  The reference runs its playouts strictly one after the other (mcts_alphaZero.py:147-149), so a single
// interactive game (human_play_mxnet.py, evaluate/ChessClient.py) is latency bound on a GPU: 400 lock-steps of ~9
// dependent kernels.  Here a game's warp selects up to k leaves back to back; every node on a selected path carries a
// virtual visit (vn) until its backup, and a child with vn > 0 in-flight visits is scored as if each of them had
// already returned a loss:  N' = N + vn,  Q' = (N Q - vn) / N',  u = c P sqrt(Np + vn_parent) / (1 + N').
// vn == 0 leaves the reference's arithmetic untouched, so k = 1 builds the parity-mode tree bit for bit.  All k
// leaves are evaluated as ONE net batch, then expanded / backed up in selection order.  Visit counts differ from the
// sequential search for k > 1 (that is what virtual loss does); every playout still passes through exactly one root
// child, so the root's children still sum to n_playout - 1 for a fresh tree.
// A game whose root is still a leaf issues ONE playout in that lock-step (k traversals would all stop at the root).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tree_select_child_vl(const Pools& pl, const int32_t* __restrict__ vn, size_t base, int cs,
                                                    int cc, int np_eff, double c_puct, int lane) {
  constexpr int PER = AP_MAX_S / 32;
  double p[PER], q[PER];
  int n[PER], v[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lane + 32 * j;
    const size_t c = base + cs + i;
    p[j] = 0.0, q[j] = 0.0, n[j] = 0, v[j] = 0;
    if (i < cc) {
      p[j] = pl.P[c];
      q[j] = pl.Q[c];
      n[j] = pl.N[c];
      v[j] = vn[c];
    }
  }
  const double sq = __dsqrt_rn((double)np_eff);
  double bv = -CUDART_INF;
  int bi = INT_MAX;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lane + 32 * j;
    if (i < cc) {
      double qq = q[j];
      int nn = n[j];
      if (v[j] > 0) {  // in-flight visits count as losses
        nn = n[j] + v[j];
        qq = __ddiv_rn(__dsub_rn(__dmul_rn((double)n[j], q[j]), (double)v[j]), (double)nn);
      }
      const double u = __ddiv_rn(__dmul_rn(__dmul_rn(c_puct, p[j]), sq), (double)(1 + nn));
      const double val = __dadd_rn(qq, u);
      if (val > bv || bi == INT_MAX) {
        bv = val;
        bi = i;
      }
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    double ov = __shfl_xor_sync(AP_FULL, bv, d);
    int oi = __shfl_xor_sync(AP_FULL, bi, d);
    bool take = (oi != INT_MAX) && (bi == INT_MAX || (oi < bi ? !(bv > ov) : (ov > bv)));
    if (take) {
      bv = ov;
      bi = oi;
    }
  }
  return bi;
}

// one warp per game; leaf record j = g * kstride + i.  remain[g] = playouts this game still owes in this search.
__global__ void __launch_bounds__(32 * SEL_WARPS)
k_select_vl(Geo geo, const uint32_t* __restrict__ rows, const BoardMeta* __restrict__ meta, Pools pl, int32_t* vn,
            Leaves lv, int k, int kstride, int32_t* remain, int32_t* issued, unsigned long long* stats, __half* feat,
            long long mpad) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  const size_t base = (size_t)g * geo.cap;
  int want = min(k, remain[g]);
  if (!lv.active[g]) want = 0;
  if (want > 1 && pl.child_start[base] < 0) want = 1;  // fresh root: k traversals would all end at the root
  const WBoard root = wb_load(rows, meta, g, lane);
  for (int i = 0; i < want; ++i) {
    const int j = g * kstride + i;
    WBoard b = root;
    int node = 0, depth = 0;
    unsigned long long scanned = 0;
    while (true) {
      const int cs = pl.child_start[base + node];
      const int cc = pl.child_count[base + node];
      const int np = pl.N[base + node];
      const int mv = pl.move[base + node];
      const int vp = vn[base + node];  // in-flight visits through this node BEFORE this traversal
      if (depth > 0) {
        if (lane == 0) lv.path[(size_t)j * geo.S + depth - 1] = (int16_t)mv;
        wb_do_move(b, mv, geo.W, lane);
      }
      __syncwarp();
      if (lane == 0) vn[base + node] = vp + 1;
      __syncwarp();
      if (cs < 0) break;
      node = cs + tree_select_child_vl(pl, vn, base, cs, cc, np + vp, geo.c_puct, lane);
      ++depth;
      scanned += cc;
    }
    int winner;
    const bool end = wb_game_end(b, geo.n_in_row, geo.S, winner);
    wb_store(b, lv.rows, lv.meta, j, lane, 0);
    int sl = -1;
    if (lane == 0) {
      lv.node[j] = node;
      lv.terminal[j] = end ? 1 : 0;
      lv.winner[j] = (int8_t)winner;
      lv.depth[j] = depth;
      atomicAdd(&stats[0], 1ull);
      atomicAdd(&stats[1], scanned);
      atomicAdd(&stats[3], (unsigned long long)(depth + 1));
      if (end) atomicAdd(&stats[4], 1ull);
      if (!end) {
        sl = atomicAdd(lv.n_eval, 1);
        lv.game_of_slot[sl] = j;
      }
      lv.slot[j] = sl;
    }
    sl = __shfl_sync(AP_FULL, sl, 0);
    if (sl >= 0) emit_features_warp(b, geo.W, geo.H, sl, feat, mpad, lane);
    __syncwarp();
  }
  if (lane == 0) {
    issued[g] = want;
    remain[g] -= want;
  }
}

// the issued[g] leaves of game g in selection order: expand (unless terminal or already expanded by an earlier leaf of
// this lock-step), back the value up, take the virtual visits off the path
__global__ void __launch_bounds__(32 * SEL_WARPS)
k_expand_backup_vl(Geo geo, Pools pl, int32_t* vn, Leaves lv, int kstride, const int32_t* __restrict__ issued,
                   int32_t* errflag, unsigned long long* stats, FcFinish fin) {
  __shared__ int16_t s_list[SEL_WARPS][AP_MAX_S];
  __shared__ float s_prob[SEL_WARPS][AP_MAX_S];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int g = blockIdx.x * SEL_WARPS + w;
  if (g >= geo.G) return;
  const size_t base = (size_t)g * geo.cap;
  const int n = issued[g];
  for (int i = 0; i < n; ++i) {
    const int j = g * kstride + i;
    const int leaf = lv.node[j];
    double v;
    if (!lv.terminal[j]) {
      const int src = lv.slot[j];
      float pr[8];
      const float fv = fc_finish_warp(fin, src, geo.S, lane, pr);
      v = (double)fv;
      if (pl.child_start[base + leaf] < 0) {  // not yet expanded by a duplicate leaf of this lock-step
        WBoard b = wb_load(lv.rows, lv.meta, j, lane);
        const int A = wb_legal_list(b, geo.W, geo.H, lane, s_list[w]);
#pragma unroll
        for (int q = 0; q < 8; ++q) s_prob[w][lane + 32 * q] = pr[q];
        __syncwarp();
        const bool ok = tree_expand(pl, base, g, geo.cap, leaf, A, s_list[w],
                                    [&](int kk, int mv) { return (double)s_prob[w][mv]; }, lane);
        if (!ok && lane == 0) errflag[g] = AP_ERR_POOL_EXHAUSTED;
        if (ok && lane == 0) atomicAdd(&stats[2], (unsigned long long)A);
      }
    } else {
      const int winner = lv.winner[j];
      const int cur = lv.meta[j].cur;
      v = (winner == -1) ? 0.0 : ((winner == cur) ? 1.0 : -1.0);
    }
    __syncwarp();
    if (lane == 0) {
      tree_backup(pl, base, leaf, -v);
      for (int nd = leaf; nd >= 0; nd = pl.parent[base + nd]) vn[base + nd] -= 1;
    }
    __syncwarp();
  }
  if (g == 0 && lane == 0) *lv.n_eval = 0;
}

void launch_select_vl(ap_engine* e, const Leaves& lv, int32_t* vn, int k, int kstride, int32_t* remain, int32_t* issued) {
  __half* feat = nullptr;
  long long mpad = 0;
  net_feature_planes(e, &feat, &mpad);
  k_select_vl<<<(e->geo.G + SEL_WARPS - 1) / SEL_WARPS, 32 * SEL_WARPS, 0, e->stream>>>(
      e->geo, e->rows, e->meta, e->pools, vn, lv, k, kstride, remain, issued, e->stats, feat, mpad);
}
void launch_expand_backup_vl(ap_engine* e, const Leaves& lv, int32_t* vn, int kstride, const int32_t* issued) {
  FcFinish fin{nullptr, nullptr, 0, 0, 0};
  net_fc_finish_args(e, &fin.partial, &fin.bias, &fin.rows, &fin.np, &fin.ksplit);
  k_expand_backup_vl<<<(e->geo.G + SEL_WARPS - 1) / SEL_WARPS, 32 * SEL_WARPS, 0, e->stream>>>(
      e->geo, e->pools, vn, lv, kstride, issued, e->errflag, e->stats, fin);
}

// ---------------------------------------------------------------------------------------------
// MCTS.update_with_move (mcts_alphaZero.py:159-167): re-root on a child keeping its subtree,
// compacted breadth-first (children blocks stay contiguous and ordered) through a per-CTA
// scratch slot, or a fresh TreeNode(None, 1.0).
// ---------------------------------------------------------------------------------------------
#define ADV_THREADS 256

struct Scratch {
  double* P;
  double* Q;
  int32_t* N;
  int32_t* child_start;
  int32_t* parent;
  int32_t* src;
  uint16_t* child_count;
  int16_t* move;
};

__host__ __device__ inline size_t slot_bytes(int cap) { return (size_t)cap * (8 + 8 + 4 + 4 + 4 + 4 + 2 + 2); }
size_t scratch_bytes_per_slot(int cap) { return slot_bytes(cap); }

__device__ __forceinline__ Scratch scratch_slot(void* p, int slot, int cap) {
  char* b = (char*)p + (size_t)slot * slot_bytes(cap);
  Scratch s;
  s.P = (double*)b;            b += (size_t)cap * 8;
  s.Q = (double*)b;            b += (size_t)cap * 8;
  s.N = (int32_t*)b;           b += (size_t)cap * 4;
  s.child_start = (int32_t*)b; b += (size_t)cap * 4;
  s.parent = (int32_t*)b;      b += (size_t)cap * 4;
  s.src = (int32_t*)b;         b += (size_t)cap * 4;
  s.child_count = (uint16_t*)b; b += (size_t)cap * 2;
  s.move = (int16_t*)b;
  return s;
}

__global__ void __launch_bounds__(ADV_THREADS)
k_advance(Geo geo, Pools pl, const int32_t* __restrict__ ids, const int32_t* __restrict__ moves, int n, void* scratch) {
  __shared__ int s_scan[ADV_THREADS];
  __shared__ int s_e_old[ADV_THREADS], s_e_off[ADV_THREADS], s_e_par[ADV_THREADS];
  __shared__ int s_found, s_count, s_ne, s_total;
  const int tid = threadIdx.x;
  Scratch sc = scratch_slot(scratch, blockIdx.x, geo.cap);
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int g = ids[i];
    const int mv = moves[i];
    const size_t base = (size_t)g * geo.cap;
    if (tid == 0) s_found = -1;
    __syncthreads();
    const int rcs = pl.child_start[base];
    const int rcc = (rcs >= 0) ? pl.child_count[base] : 0;
    if (mv >= 0)
      for (int k = tid; k < rcc; k += ADV_THREADS)
        if (pl.move[base + rcs + k] == mv) s_found = rcs + k;
    __syncthreads();
    const int r = s_found;
    if (r < 0) {
      if (tid == 0) tree_write_root(pl, base, g);
      __syncthreads();
      continue;
    }
    if (tid == 0) {
      sc.P[0] = pl.P[base + r];
      sc.Q[0] = pl.Q[base + r];
      sc.N[0] = pl.N[base + r];
      sc.child_start[0] = pl.child_start[base + r];
      sc.child_count[0] = pl.child_count[base + r];
      sc.move[0] = pl.move[base + r];
      sc.parent[0] = -1;  // _root._parent = None
      sc.src[0] = r;
      s_count = 1;
    }
    __syncthreads();
    int head = 0;
    while (true) {
      const int count = s_count;
      if (head >= count) break;
      const int chunk_end = min(head + ADV_THREADS, count);
      const int me = head + tid;
      int my_cc = 0, my_ocs = -1;
      if (me < chunk_end) {
        int o = sc.src[me];
        my_ocs = pl.child_start[base + o];
        my_cc = (my_ocs >= 0) ? (int)pl.child_count[base + o] : 0;
      }
      // inclusive scans of (children, has-children) packed in one int: cc < 2^9, count of entries < 2^9
      int val = (my_cc << 10) | (my_cc > 0 ? 1 : 0);
      s_scan[tid] = val;
      __syncthreads();
      for (int d = 1; d < ADV_THREADS; d <<= 1) {
        int t = (tid >= d) ? s_scan[tid - d] : 0;
        __syncthreads();
        s_scan[tid] += t;
        __syncthreads();
      }
      const int incl = s_scan[tid];
      const int off = (incl >> 10) - my_cc;
      const int rank = (incl & 1023) - (my_cc > 0 ? 1 : 0);
      if (my_cc > 0) {
        sc.child_start[me] = count + off;
        s_e_old[rank] = my_ocs;
        s_e_off[rank] = off;
        s_e_par[rank] = me;
      }
      if (tid == ADV_THREADS - 1) {
        s_total = incl >> 10;
        s_ne = incl & 1023;
      }
      __syncthreads();
      const int total = s_total, ne = s_ne;
      for (int f = tid; f < total; f += ADV_THREADS) {
        int lo = 0, hi = ne - 1;  // last entry with off <= f
        while (lo < hi) {
          int mid = (lo + hi + 1) >> 1;
          if (s_e_off[mid] <= f) lo = mid; else hi = mid - 1;
        }
        const int o = s_e_old[lo] + (f - s_e_off[lo]);
        const int d = count + f;
        sc.P[d] = pl.P[base + o];
        sc.Q[d] = pl.Q[base + o];
        sc.N[d] = pl.N[base + o];
        sc.child_start[d] = pl.child_start[base + o];  // fixed up when node d is processed
        sc.child_count[d] = pl.child_count[base + o];
        sc.move[d] = pl.move[base + o];
        sc.parent[d] = s_e_par[lo];
        sc.src[d] = o;
      }
      __syncthreads();
      if (tid == 0) s_count = count + total;
      head = chunk_end;
      __syncthreads();
    }
    const int count = s_count;
    for (int k = tid; k < count; k += ADV_THREADS) {
      pl.P[base + k] = sc.P[k];
      pl.Q[base + k] = sc.Q[k];
      pl.N[base + k] = sc.N[k];
      pl.child_start[base + k] = sc.child_start[k];
      pl.child_count[base + k] = sc.child_count[k];
      pl.move[base + k] = sc.move[k];
      pl.parent[base + k] = sc.parent[k];
    }
    if (tid == 0) pl.alloc[g] = count;
    __syncthreads();
  }
}

// acts / visits / Q of the root's children in insertion order (mcts_alphaZero.py:152-154)
__global__ void k_root(Geo geo, Pools pl, const int32_t* ids, int n, int32_t* out_count, int16_t* out_acts,
                       int32_t* out_visits, double* out_q, int32_t* out_rootn) {
  int lane = threadIdx.x & 31;
  int i = blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
  if (i >= n) return;
  int g = ids ? ids[i] : i;
  size_t base = (size_t)g * geo.cap;
  int cs = pl.child_start[base];
  int cc = (cs >= 0) ? pl.child_count[base] : 0;
  for (int k = lane; k < geo.S; k += 32) {
    bool in = k < cc;
    out_acts[(size_t)i * geo.S + k] = in ? pl.move[base + cs + k] : (int16_t)-1;
    out_visits[(size_t)i * geo.S + k] = in ? pl.N[base + cs + k] : 0;
    if (out_q) out_q[(size_t)i * geo.S + k] = in ? pl.Q[base + cs + k] : 0.0;
  }
  if (lane == 0) {
    out_count[i] = cc;
    if (out_rootn) out_rootn[i] = pl.N[base];
  }
}

// softmax(1/temp * log(visits + 1e-10)) with max subtraction (mcts_alphaZero.py:13-16,155),
// scattered by move index.  fp64; transcendental ulps may differ from numpy's (the Python shim
// therefore recomputes pi from the exact visit counts when bit-equality with numpy is wanted).
__global__ void k_root_probs(Geo geo, Pools pl, double temp, double* out) {
  int lane = threadIdx.x & 31;
  int g = blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  size_t base = (size_t)g * geo.cap;
  int cs = pl.child_start[base];
  int cc = (cs >= 0) ? pl.child_count[base] : 0;
  for (int k = lane; k < geo.S; k += 32) out[(size_t)g * geo.S + k] = 0.0;
  __syncwarp();
  double inv = 1.0 / temp;
  double mx = -CUDART_INF;
  for (int k = lane; k < cc; k += 32) mx = fmax(mx, inv * log((double)pl.N[base + cs + k] + 1e-10));
  for (int d = 16; d >= 1; d >>= 1) mx = fmax(mx, __shfl_xor_sync(AP_FULL, mx, d));
  double sum = 0.0;
  for (int k = lane; k < cc; k += 32) sum += exp(inv * log((double)pl.N[base + cs + k] + 1e-10) - mx);
  for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(AP_FULL, sum, d);
  for (int k = lane; k < cc; k += 32) {
    double p = exp(inv * log((double)pl.N[base + cs + k] + 1e-10) - mx) / sum;
    out[(size_t)g * geo.S + pl.move[base + cs + k]] = p;
  }
}

// ---------------------------------------------------------------------------------------------
// MCTSPlayer.get_action in self-play (mcts_alphaZero.py:187-215) for every game, on the device:
//   pi = softmax(1/temp * log(visits + 1e-10)); move ~ (1 - eps) * pi + eps * Dirichlet(alpha * ones(A)); the
//   returned pi is the un-noised one, scattered by move index (fp32, the training record).
// The reference draws from NumPy's global generator; here every (game, ply) owns a Philox4x32-10 stream, the
// Dirichlet sample is normalised Gamma(alpha) variates (Marsaglia-Tsang, boosted for alpha < 1) and the move is the
// first child whose running sum exceeds u * total (np.random.choice's inverse-cdf rule).  Distributional parity
// (tests/test_gpu_shims.py): move frequencies vs pi and vs the uniform mean of the noise, the noise marginals vs
// Beta(alpha, (A - 1) alpha).  One warp per game.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_pick(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
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
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {  // 53-bit uniform in (0, 1)
  const unsigned long long b = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}
// Gamma(alpha, 1): Marsaglia & Tsang (2000) for alpha + 1 >= 1, times U^(1/alpha) when alpha < 1
__device__ double gamma_variate(double alpha, unsigned long long seed, uint32_t g, uint32_t child, uint32_t ply) {
  const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  uint32_t r[4];
  double out = 0.0;
  for (uint32_t trial = 0; trial < 64; ++trial) {
    philox_pick(ply, (child << 8) | trial, g, 0x5e1fu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);  // Box-Muller
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    philox_pick(ply, (child << 8) | trial, g, 0xacce97u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = u01(r[0], r[1]);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
      out = d * v;
      if (alpha < 1.0) out *= pow(u01(r[2], r[3]), 1.0 / alpha);
      break;
    }
  }
  return out;
}

__global__ void __launch_bounds__(32 * SEL_WARPS)
k_selfplay_pick(Geo geo, Pools pl, double temp, double eps, double alpha, unsigned long long seed, uint32_t ply,
                int32_t* __restrict__ out_move, float* __restrict__ out_pi, double* __restrict__ out_noise) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
  if (g >= geo.G) return;
  const size_t base = (size_t)g * geo.cap;
  const int cs = pl.child_start[base];
  const int cc = (cs >= 0) ? pl.child_count[base] : 0;
  for (int k = lane; k < geo.S; k += 32) {
    out_pi[(size_t)g * geo.S + k] = 0.f;
    if (out_noise) out_noise[(size_t)g * geo.S + k] = 0.0;
  }
  __syncwarp();
  if (cc == 0) {  // "WARNING: the board is full" (mcts_alphaZero.py:217-218)
    if (lane == 0) out_move[g] = -1;
    return;
  }
  constexpr int PER = AP_MAX_S / 32;
  double p[PER], nz[PER];
  const double inv = 1.0 / temp;
  double mx = -CUDART_INF;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int k = lane + 32 * j;
    p[j] = -CUDART_INF;
    nz[j] = 0.0;
    if (k < cc) {
      p[j] = inv * log((double)pl.N[base + cs + k] + 1e-10);
      mx = fmax(mx, p[j]);
      if (eps > 0.0) nz[j] = gamma_variate(alpha, seed, (uint32_t)g, (uint32_t)k, ply);
    }
  }
  for (int d = 16; d >= 1; d >>= 1) mx = fmax(mx, __shfl_xor_sync(AP_FULL, mx, d));
  double sum = 0.0, nsum = 0.0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    p[j] = (lane + 32 * j < cc) ? exp(p[j] - mx) : 0.0;
    sum += p[j];
    nsum += nz[j];
  }
  for (int d = 16; d >= 1; d >>= 1) {
    sum += __shfl_xor_sync(AP_FULL, sum, d);
    nsum += __shfl_xor_sync(AP_FULL, nsum, d);
  }
  if (!(nsum > 0.0)) nsum = 1.0;
  // sampling weights in child order; the running sum walks j-major (k = lane + 32 j), so accumulate per j
  uint32_t r[4];
  philox_pick(ply, 0xffffffffu, (uint32_t)g, 0x9a3eu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u = u01(r[0], r[1]);
  double w[PER], total = 0.0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    p[j] /= sum;
    nz[j] /= nsum;
    w[j] = (1.0 - eps) * p[j] + eps * nz[j];
    double t = w[j];
    for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(AP_FULL, t, d);
    total += t;
  }
  const double target = u * total;
  // first child (in child order) whose inclusive prefix sum exceeds the target
  double before = 0.0;
  int chosen = cc - 1;
  bool found = false;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    double incl = w[j];
    for (int d = 1; d < 32; d <<= 1) {
      const double t = __shfl_up_sync(AP_FULL, incl, d);
      if (lane >= d) incl += t;
    }
    incl += before;
    const unsigned hit = __ballot_sync(AP_FULL, (lane + 32 * j < cc) && (incl > target));
    if (!found && hit) {
      chosen = 32 * j + (__ffs(hit) - 1);
      found = true;
    }
    before = __shfl_sync(AP_FULL, incl, 31);
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int k = lane + 32 * j;
    if (k < cc) {
      const int mv = pl.move[base + cs + k];
      out_pi[(size_t)g * geo.S + mv] = (float)p[j];
      if (out_noise) out_noise[(size_t)g * geo.S + mv] = nz[j];
    }
  }
  if (lane == 0) out_move[g] = pl.move[base + cs + chosen];
}

// largest number of nodes in use over all games (pool growth check)
__global__ void k_max_alloc(int G, const int32_t* __restrict__ alloc, int32_t* out) {
  int m = 0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) m = max(m, alloc[g]);
  for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(AP_FULL, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}
// move every game's nodes [0, alloc[g]) into pools of a larger per-game stride; node indices are relative to the
// game's base, so no field changes
__global__ void k_pool_copy(int G, Pools from, int from_cap, Pools to, int to_cap) {
  const int g = blockIdx.x;
  const int n = from.alloc[g];
  const size_t a = (size_t)g * from_cap, b = (size_t)g * to_cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    to.P[b + i] = from.P[a + i];
    to.Q[b + i] = from.Q[a + i];
    to.N[b + i] = from.N[a + i];
    to.child_start[b + i] = from.child_start[a + i];
    to.parent[b + i] = from.parent[a + i];
    to.child_count[b + i] = from.child_count[a + i];
    to.move[b + i] = from.move[a + i];
  }
  if (threadIdx.x == 0) to.alloc[g] = n;
}
void launch_max_alloc(ap_engine* e, int32_t* d_out) {
  cudaMemsetAsync(d_out, 0, 4, e->stream);
  int blocks = (e->geo.G + 255) / 256;
  k_max_alloc<<<blocks < 64 ? blocks : 64, 256, 0, e->stream>>>(e->geo.G, e->pools.alloc, d_out);
}
void launch_pool_copy(ap_engine* e, const Pools& from, int from_cap, const Pools& to, int to_cap) {
  k_pool_copy<<<e->geo.G, 256, 0, e->stream>>>(e->geo.G, from, from_cap, to, to_cap);
}

static inline dim3 sel_grid(int n) { return dim3((n + SEL_WARPS - 1) / SEL_WARPS); }

void launch_tree_reset_all(ap_engine* e) {
  k_tree_reset_all<<<(e->geo.G + 127) / 128, 128, 0, e->stream>>>(e->geo, e->pools);
}
void launch_select(ap_engine* e, bool compact) {
  __half* feat = nullptr;
  long long mpad = 0;
  if (compact) net_feature_planes(e, &feat, &mpad);
  k_select<<<sel_grid(e->geo.G), 32 * SEL_WARPS, 0, e->stream>>>(e->geo, e->rows, e->meta, e->pools, e->leaves, e->stats,
                                                                compact ? 1 : 0, feat, mpad);
}
void launch_expand_backup(ap_engine* e, const int32_t* d_counts, const int16_t* d_acts, const double* d_pri64,
                          const double* d_val64, const float* d_pri32, const float* d_val32, const int32_t* d_slot,
                          bool fuse_fc_finish) {
  FcFinish fin{nullptr, nullptr, 0, 0, 0};
  if (fuse_fc_finish) net_fc_finish_args(e, &fin.partial, &fin.bias, &fin.rows, &fin.np, &fin.ksplit);
  k_expand_backup<<<sel_grid(e->geo.G), 32 * SEL_WARPS, 0, e->stream>>>(e->geo, e->pools, e->leaves, d_counts, d_acts,
                                                                       d_pri64, d_val64, d_pri32, d_val32, d_slot,
                                                                       e->errflag, e->stats, fin);
}

// Order-preserving compaction of the non-terminal leaves: slot[g] = number of non-terminal leaves before game g
// (or -1), game_of_slot its inverse, n_eval the count.  One CTA; G is a few thousand.
__global__ void __launch_bounds__(1024) k_compact_leaves(int G, const int8_t* __restrict__ terminal, int32_t* slot,
                                                         int32_t* game_of_slot, int32_t* n_eval) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int g0 = 0; g0 < G; g0 += 1024) {
    const int g = g0 + tid;
    const int live = (g < G && !terminal[g]) ? 1 : 0;
    const unsigned m = __ballot_sync(AP_FULL, live);
    const int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[w] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int i = 0; i < w; ++i) off += s_warp[i];
    if (g < G) {
      slot[g] = live ? off + pre : -1;
      if (live) game_of_slot[off + pre] = g;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int i = 0; i < 32; ++i) t += s_warp[i];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0) *n_eval = s_base;
}

void launch_compact_leaves(ap_engine* e) {
  k_compact_leaves<<<1, 1024, 0, e->stream>>>(e->geo.G, e->leaves.terminal, e->leaves.slot, e->leaves.game_of_slot,
                                             e->leaves.n_eval);
}
void launch_advance(ap_engine* e, int n, const int32_t* d_moves) {
  int grid = n < e->scratch_slots ? n : e->scratch_slots;
  k_advance<<<grid, ADV_THREADS, 0, e->stream>>>(e->geo, e->pools, e->d_ids, d_moves, n, e->scratch);
}
void launch_root(ap_engine* e, const int32_t* d_ids, int n, int32_t* d_count, int16_t* d_acts, int32_t* d_visits,
                 double* d_q, int32_t* d_rootn) {
  k_root<<<sel_grid(n), 32 * SEL_WARPS, 0, e->stream>>>(e->geo, e->pools, d_ids, n, d_count, d_acts, d_visits, d_q,
                                                       d_rootn);
}
void launch_selfplay_pick(ap_engine* e, double temp, double eps, double alpha, uint64_t seed, uint32_t ply,
                          int32_t* d_move, float* d_pi, double* d_noise) {
  k_selfplay_pick<<<sel_grid(e->geo.G), 32 * SEL_WARPS, 0, e->stream>>>(e->geo, e->pools, temp, eps, alpha, seed, ply,
                                                                       d_move, d_pi, d_noise);
}
void launch_root_probs(ap_engine* e, double temp, double* d_out) {
  k_root_probs<<<sel_grid(e->geo.G), 32 * SEL_WARPS, 0, e->stream>>>(e->geo, e->pools, temp, d_out);
}
