// Second half of the split-K FC heads (heads_tc.cu): one warp per board adds the partial logits in fixed order,
// the bias, and does the softmax over the S policy logits and the tanh of the value logit
// (SoftmaxActivation / tanh, policy_value_net_mxnet_simple.py:84,90).  Shared by k_head_fc_finish and by
// k_expand_backup, which consumes the probabilities as priors without a round trip through HBM.
#pragma once
#include <math.h>

#include "ap_common.cuh"

struct FcFinish {
  const float* partial;  // [ksplit][rows][np] raw partial logits; nullptr = not fused
  const float* bias;     // [np]
  long long rows;
  int np, ksplit;
};

// lane holds column n = lane + 32 j of board b in p[j] (probability for n < S); returns tanh(value logit) in all lanes
__device__ __forceinline__ float fc_finish_warp(const FcFinish& f, int b, int S, int lane, float (&p)[8]) {
  float mx = -INFINITY, vlogit = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    p[j] = 0.f;
    if (n < f.np) {
      float a = 0.f;
      for (int s = 0; s < f.ksplit; ++s) a += f.partial[((size_t)s * f.rows + b) * f.np + n];
      p[j] = a + f.bias[n];
      if (n < S) mx = fmaxf(mx, p[j]);
      if (n == S) vlogit = p[j];
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(AP_FULL, mx, d));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    if (n < S) {
      p[j] = expf(p[j] - mx);
      sum += p[j];
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(AP_FULL, sum, d);
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] *= inv;
  return tanhf(__shfl_sync(AP_FULL, vlogit, S & 31));
}
