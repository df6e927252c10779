// Warp-level MCTS primitives on the flat SoA node pool of one game.
// Replaces reference TreeNode (mcts_alphaZero.py:19-87; identical copy mcts_pure.py:28-94).
// All tree arithmetic is fp64 with explicitly rounded ops (no FMA contraction) in the
// reference's evaluation order, so visit counts and Q are bit-exact.
#pragma once
#include <limits.h>
#include <math_constants.h>

#include "board.cuh"

// TreeNode.select (mcts_alphaZero.py:43-49) over the children block [cs, cs+cc) of a node
// whose visit count is np:  score = Q + ((c_puct*P)*sqrt(Np))/(1+N)   (:78-80),
// first maximum in child (insertion) order.  Returns the child offset within the block.
// Latency: every lane issues the loads of all its (<= AP_MAX_S/32) children before the first fp64 op, so a
// level costs one memory round trip instead of one per 32 children (the fp64 division's slow-path branch
// otherwise keeps the compiler from hoisting the next iteration's loads).  The caller reads the winner's move
// together with its node fields (one more round trip for both).
__device__ __forceinline__ int tree_select_child(const Pools& pl, size_t base, int cs, int cc, int np, double c_puct,
                                                 int lane) {
  constexpr int PER = AP_MAX_S / 32;
  double p[PER], q[PER];
  int n[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lane + 32 * j;
    const size_t c = base + cs + i;
    p[j] = 0.0, q[j] = 0.0, n[j] = 0;
    if (i < cc) {
      p[j] = pl.P[c];
      q[j] = pl.Q[c];
      n[j] = pl.N[c];
    }
  }
  const double sq = __dsqrt_rn((double)np);
  double bv = -CUDART_INF;
  int bi = INT_MAX;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lane + 32 * j;
    if (i < cc) {
      const double u = __ddiv_rn(__dmul_rn(__dmul_rn(c_puct, p[j]), sq), (double)(1 + n[j]));
      const double v = __dadd_rn(q[j], u);
      if (v > bv || bi == INT_MAX) {  // first element always taken, later only if strictly greater
        bv = v;
        bi = i;
      }
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    double ov = __shfl_xor_sync(AP_FULL, bv, d);
    int oi = __shfl_xor_sync(AP_FULL, bi, d);
    // python max(): keep the earliest index unless a later one is strictly greater
    bool take = (oi != INT_MAX) && (bi == INT_MAX || (oi < bi ? !(bv > ov) : (ov > bv)));
    if (take) {
      bv = ov;
      bi = oi;
    }
  }
  return bi;
}

// TreeNode.update (mcts_alphaZero.py:51-59) along parent links = update_recursive (:61-67).
// x is the value for `node` itself (sign flips per ply).  Single lane.
__device__ __forceinline__ int tree_backup(const Pools& pl, size_t base, int node, double x) {
  int len = 0;
  while (node >= 0) {
    size_t c = base + node;
    int n = pl.N[c] + 1;
    double q = pl.Q[c];
    pl.N[c] = n;
    pl.Q[c] = __dadd_rn(q, __ddiv_rn(__dmul_rn(1.0, __dsub_rn(x, q)), (double)n));
    x = -x;
    node = pl.parent[c];
    ++len;
  }
  return len;
}

__device__ __forceinline__ void tree_write_root(const Pools& pl, size_t base, int g) {
  pl.P[base] = 1.0;  // TreeNode(None, 1.0)  mcts_alphaZero.py:102,167
  pl.Q[base] = 0.0;
  pl.N[base] = 0;
  pl.child_start[base] = -1;
  pl.child_count[base] = 0;
  pl.parent[base] = -1;
  pl.move[base] = -1;
  pl.alloc[g] = 1;
}

// TreeNode.expand (mcts_alphaZero.py:34-41): children in list order.  prior(k) supplies P.
// Returns false when the pool is exhausted (nothing written).
template <class PriorFn>
__device__ __forceinline__ bool tree_expand(const Pools& pl, size_t base, int g, int cap, int leaf, int A,
                                            const int16_t* list, PriorFn prior, int lane) {
  if (A <= 0) return true;  // node stays a leaf: is_leaf() is `_children == {}`  (:82-84)
  int a0 = pl.alloc[g];
  if (a0 + A > cap) return false;
  for (int k = lane; k < A; k += 32) {
    size_t c = base + a0 + k;
    int mv = list[k];
    pl.P[c] = prior(k, mv);
    pl.Q[c] = 0.0;
    pl.N[c] = 0;
    pl.child_start[c] = -1;
    pl.child_count[c] = 0;
    pl.parent[c] = leaf;
    pl.move[c] = (int16_t)mv;
  }
  __syncwarp();
  if (lane == 0) {
    pl.child_start[base + leaf] = a0;
    pl.child_count[base + leaf] = (uint16_t)A;
    pl.alloc[g] = a0 + A;
  }
  __syncwarp();
  return true;
}
