// Policy/value net: parameter table with the reference's names, BN folding / operand images for the
// tensor-core trunk (conv_tc.cu), feature emission from bitboards, the fused heads kernel, and an
// independent fp32 CUDA-core path used as an on-device cross-check.
// Replaces PolicyValueNet inference (policy_value_net_mxnet_simple.py:68-92,178-226;
// policy_value_net_mxnet.py:70-102,232-280).
#include <stdlib.h>
#include <string.h>

#include "board.cuh"
#include "features.cuh"
#include "kernels.h"
#include "net.h"
#include "ptx.cuh"

#define BN_EPS 1e-3f

#define AP_TRY(x)               \
  do {                          \
    int _r = (x);               \
    if (_r != AP_OK) return _r; \
  } while (0)

// ------------------------------------------------------------------------------------------
// prep kernels (run at load and after every weight refresh)
// ------------------------------------------------------------------------------------------
__global__ void k_prep_conv(const float* __restrict__ master, long long w, long long b, long long gamma, long long beta,
                            long long mean, long long var, int fix_gamma, int cin, int cin_pad, int cout, int kc, int ntaps,
                            float post, __half* wimg, __half* wimg2, __half* wimg_lo, __half* wimg4, float* scale,
                            float* shift) {
  const long long total = (long long)ntaps * cin_pad * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // image order: [kcI][tap][j][n][e]
    int e = (int)(i & 7);
    long long r = i >> 3;
    int n = (int)(r % cout);
    r /= cout;
    int j = (int)(r % (kc >> 3));
    r /= (kc >> 3);
    int tap = (int)(r % ntaps);
    int kcI = (int)(r / ntaps);
    int k = kcI * kc + j * 8 + e;
    float sc = post / sqrtf(master[var + n] + BN_EPS);
    if (!fix_gamma) sc *= master[gamma + n];
    float v = (k < cin) ? master[w + ((long long)n * cin + k) * ntaps + tap] * sc : 0.f;  // BN scale folded in fp32
    const __half vh = __float2half_rn(v);
    wimg[i] = vh;
    if (wimg_lo) wimg_lo[i] = __float2half_rn(v - __half2float(vh));
    // CTA-pair image: [r][kcI][tap][j][n'][e], n = r*cout/2 + n'
    const int nh = cout >> 1, rr = n / nh, np = n - rr * nh;
    const long long i2 = ((((long long)rr * (cin_pad / kc) + kcI) * ntaps + tap) * (kc >> 3) + j) * nh * 8 + (long long)np * 8 + e;
    wimg2[i2] = __float2half_rn(v);
    if (wimg4) {  // [nh][r][kcI][tap][j][n''][e], n = nh*128 + r*64 + n''
      const int nh4 = n >> 7, r4 = (n >> 6) & 1, n4 = n & 63;
      const long long i4 = (((((long long)nh4 * 2 + r4) * (cin_pad / kc) + kcI) * ntaps + tap) * (kc >> 3) + j) * 64 * 8 +
                           (long long)n4 * 8 + e;
      wimg4[i4] = vh;
    }
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < cout; n += gridDim.x * blockDim.x) {
    float s = 1.f / sqrtf(master[var + n] + BN_EPS);
    if (!fix_gamma) s *= master[gamma + n];
    scale[n] = s * post;
    shift[n] = ((master[b + n] - master[mean + n]) * s + master[beta + n]) * post;
  }
}

// AP_NET_SPLIT_ACT: the fp16 weight image by ERROR DIFFUSION along K instead of round-to-nearest.  One thread per
// output channel walks its (cin, tap) weights in order, carries the accumulated rounding error and picks the fp16
// neighbour (below / above) that keeps the running sum of errors closest to zero.  Post-ReLU activations are
// non-negative with a mean comparable to their spread, so the output error sum_k e_k a_k of round-to-nearest is
// dominated by mean(a) * sum_k e_k (a random walk of K ulps); diffusion pins sum_k e_k below one ulp.  With hi + lo
// activations this brings the 10-block residual net from 1.2e-3 to 4.7e-4 (max |d log p|, fp32 emulation and B200)
// at two tensor-core products per K step instead of three.
__device__ __forceinline__ __half half_neighbour(__half r, bool up) {
  unsigned short b = __half_as_ushort(r);
  if ((b & 0x7fffu) == 0) return __ushort_as_half(up ? 0x0001u : 0x8001u);  // +-0: smallest subnormals
  const bool neg = (b & 0x8000u) != 0;
  b = (unsigned short)((up != neg) ? b + 1 : b - 1);  // away from / towards zero in sign-magnitude
  return __ushort_as_half(b);
}
__global__ void k_prep_conv_diffuse(const float* __restrict__ master, long long w, long long gamma, long long var,
                                    int fix_gamma, int cin, int cout, int kc, int ntaps, float post, __half* wimg) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= cout) return;
  float sc = post / sqrtf(master[var + n] + BN_EPS);
  if (!fix_gamma) sc *= master[gamma + n];
  double acc = 0.0;
  for (int k = 0; k < cin; ++k) {
    const int kcI = k / kc, j = (k % kc) >> 3, e = k & 7;
    for (int tap = 0; tap < ntaps; ++tap) {
      const float v = master[w + ((long long)n * cin + k) * ntaps + tap] * sc;  // same fp32 value as k_prep_conv
      const __half r = __float2half_rn(v);
      const float rf = __half2float(r);
      const __half lo = (rf <= v) ? r : half_neighbour(r, false);
      const __half hi = (rf <= v) ? half_neighbour(r, true) : r;
      const double elo = (double)__half2float(lo) - (double)v, ehi = (double)__half2float(hi) - (double)v;
      const bool pick_hi = fabs(acc + ehi) < fabs(acc + elo);
      acc += pick_hi ? ehi : elo;
      const long long i = ((((long long)kcI * ntaps + tap) * (kc >> 3) + j) * cout + n) * 8 + e;
      wimg[i] = pick_hi ? hi : lo;
    }
  }
}

// Inception-ResNet variant: one named conv (weight [cout][cin][taps], bias, beta, mean, var) copied into the
// (oo, io) block of a wider zero-initialised "virtual" conv (its var initialised to 1): several towers that read
// the same input become one layer, a tower reading a channel slice becomes a layer over the whole buffer.
__global__ void k_build_virtual(const float* __restrict__ master, float* vm, NetState::VCopy c) {
  const long long nw = (long long)c.cout * c.cin * c.taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % c.taps);
    const int ci = (int)((i / c.taps) % c.cin);
    const int o = (int)(i / ((long long)c.taps * c.cin));
    vm[c.dst_w + ((long long)(c.oo + o) * c.cin_v + (c.io + ci)) * c.taps + t] = master[c.src_w + i];
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < c.cout; o += gridDim.x * blockDim.x) {
    vm[c.dst_b + c.oo + o] = master[c.src_b + o];
    vm[c.dst_beta + c.oo + o] = master[c.src_beta + o];
    vm[c.dst_mean + c.oo + o] = master[c.src_mean + o];
    vm[c.dst_var + c.oo + o] = master[c.src_var + o];
  }
}
__global__ void k_fill_range(float* p, long long off, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[off + i] = v;
}

__global__ void k_prep_heads(const float* __restrict__ master, HeadParams h, int S) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < 6 * h.cfin; i += nth) {
    int o = i / h.cfin, c = i % h.cfin;
    bool pol = o < 4;
    int oo = pol ? o : o - 4;
    float s = 1.f / sqrtf(master[(pol ? h.pvar : h.vvar) + oo] + BN_EPS);  // fix_gamma=True (conv_act)
    h.w1x1[i] = master[(pol ? h.pw : h.vw) + (long long)oo * h.cfin + c] * s;
  }
  for (int o = tid; o < 6; o += nth) {
    bool pol = o < 4;
    int oo = pol ? o : o - 4;
    float s = 1.f / sqrtf(master[(pol ? h.pvar : h.vvar) + oo] + BN_EPS);
    h.b1x1[o] = (master[(pol ? h.pb : h.vb) + oo] - master[(pol ? h.pmean : h.vmean) + oo]) * s +
                master[(pol ? h.pbeta : h.vbeta) + oo];
  }
  for (long long i = tid; i < (long long)4 * S * 232; i += nth) {  // [4S][FC_NPAD], zero padded columns
    int s = (int)(i % 232);
    long long k = i / 232;
    h.fcpT[i] = (s < S) ? master[h.fcp_w + (long long)s * 4 * S + k] : 0.f;
  }
  for (int i = tid; i < S; i += nth) h.fcp_bias[i] = master[h.fcp_b + i];
  for (int i = tid; i < 2 * S; i += nth) h.fcv[i] = master[h.fcv_w + i];
  if (tid == 0) h.fcv_bias[0] = master[h.fcv_b];
}

// ------------------------------------------------------------------------------------------
// features: leaf bitboards -> fp16 planes [2][mpad][8]  (Board.current_state, game.py:68-94)
// ------------------------------------------------------------------------------------------
__global__ void k_emit_features(Geo geo, const uint32_t* __restrict__ rows, const BoardMeta* __restrict__ meta, int nb,
                                const int32_t* __restrict__ game_of_slot, const int32_t* __restrict__ nb_dev, __half* feat,
                                long long mpad) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // net tile
  if (b >= (nb_dev ? *nb_dev : nb)) return;
  const WBoard wb = wb_load(rows, meta, game_of_slot ? game_of_slot[b] : b, lane);
  emit_features_warp(wb, geo.W, geo.H, b, feat, mpad, lane);
}

// host fp32 states [B][9][H][W] (already on device) -> fp16 planes
__global__ void k_pack_states(const float* __restrict__ st, int nb, int W, int H, __half* feat, long long mpad) {
  const int S = W * H;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * S) return;
  int b = i / S, p = i % S, y = p / W, x = p % W;
  long long row = NET_PAD_ROWS + (long long)b * NET_TILE_ROWS + y * 16 + x;
  uint4 g0, g1;
  __half* h0 = reinterpret_cast<__half*>(&g0);
  __half* h1 = reinterpret_cast<__half*>(&g1);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    h0[c] = __float2half_rn(st[((size_t)b * 9 + c) * S + p]);
    h1[c] = __float2half_rn(0.f);
  }
  h1[0] = __float2half_rn(st[((size_t)b * 9 + 8) * S + p]);
  *reinterpret_cast<uint4*>(feat + row * 8) = g0;
  *reinterpret_cast<uint4*>(feat + (mpad + row) * 8) = g1;
}

// ------------------------------------------------------------------------------------------
// heads: 1x1 convs (+BN+ReLU) -> FC(4S->S)+softmax, FC(2S->1)+tanh   (..._simple.py:78-90)
// HB boards per CTA so every FC weight read from L2 feeds HB FMAs.
// ------------------------------------------------------------------------------------------
#define HEAD_THREADS 256
#define FC_HB 16         // boards per CTA of the FC kernel
#define FC_THREADS 128   // 4 board groups (4 boards each) x 29 output groups (8 outputs each) = 116 active
#define FC_KC 36         // K chunk of the policy FC staged in shared memory (4S = 900 = 25 * 36 on 15x15)
#define FC_NPAD 232      // S <= 225 outputs padded to 29 groups of 8

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int d = 16; d >= 1; d >>= 1) {
    float o = __shfl_xor_sync(AP_FULL, v, d);
    v = is_max ? fmaxf(v, o) : v + o;
  }
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < HEAD_THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// (A) the two 1x1 convs (+ folded BN + ReLU): one CTA per board, thread = padded pixel row.
// HBM-bound: reads the trunk output once (coalesced 512 B per warp per channel group).
// hbuf [nb][6][S] fp32, pixel index p = y*W + x in tensor coordinates (Flatten order C,H,W).
__global__ void __launch_bounds__(256)
k_head_conv(const __half* __restrict__ act, long long mpad, int cfin, int W, int H, HeadParams h,
            float* __restrict__ hbuf, __half* __restrict__ fc_a, long long fc_rows, int fc_kg) {
  extern __shared__ __align__(16) float sh[];
  float* s_w = sh;  // [6][cfin]
  const int tid = threadIdx.x, b = blockIdx.x, S = W * H;
  for (int i = tid; i < 6 * cfin; i += 256) s_w[i] = h.w1x1[i];
  __syncthreads();
  const int y = tid >> 4, x = tid & 15;
  if (x >= W || y >= H) return;
  const long long row = NET_PAD_ROWS + (long long)b * NET_TILE_ROWS + tid;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int ncg = cfin >> 3;
  for (int cg0 = 0; cg0 < ncg; cg0 += 8) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = *reinterpret_cast<const uint4*>(act + ((long long)(cg0 + u) * mpad + row) * 8);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const __half2* hv = reinterpret_cast<const __half2*>(&v[u]);
      float a[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 t = __half22float2(hv[k]);
        a[2 * k] = t.x;
        a[2 * k + 1] = t.y;
      }
#pragma unroll
      for (int o = 0; o < 6; ++o) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + o * cfin + (cg0 + u) * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(s_w + o * cfin + (cg0 + u) * 8 + 4);
        acc[o] = fmaf(a[0], w0.x, fmaf(a[1], w0.y, fmaf(a[2], w0.z, fmaf(a[3], w0.w, acc[o]))));
        acc[o] = fmaf(a[4], w1.x, fmaf(a[5], w1.y, fmaf(a[6], w1.z, fmaf(a[7], w1.w, acc[o]))));
      }
    }
  }
  const int p = y * W + x;
#pragma unroll
  for (int o = 0; o < 6; ++o) {
    const float v = fmaxf(acc[o] + h.b1x1[o], 0.f);
    if (fc_a) {  // split-fp16 A operand of the tensor-core FC (heads_tc.cu)
      const __half hi = __float2half_rn(v);
      const int k = o * S + p;
      const long long at = ((long long)(k >> 3) * fc_rows + b) * 8 + (k & 7);
      fc_a[at] = hi;
      fc_a[(long long)fc_kg * fc_rows * 8 + at] = __float2half_rn(v - __half2float(hi));
    } else {
      hbuf[((size_t)b * 6 + o) * S + p] = v;
    }
  }
}

// (B) FC(4S->S)+softmax and FC(2S->1)+tanh for FC_HB boards per CTA.  The policy FC is a
// register-tiled fp32 GEMM [16 x 4S] x [4S x S] (thread = 4 boards x 8 outputs); the weight matrix
// streams through shared memory in K chunks by TMA bulk copies, double buffered on two mbarriers.
// smem: s_h [FC_HB][6S] | s_wc [2][FC_KC][FC_NPAD] | s_lg [FC_HB][FC_NPAD] | 2 mbarriers
__global__ void __launch_bounds__(FC_THREADS)
k_head_fc(const float* __restrict__ hbuf, int S, int nb, HeadParams h, float* __restrict__ probs,
          float* __restrict__ values) {
  extern __shared__ __align__(128) float sh[];
  const int K = 4 * S;
  float* s_wc = sh;                                   // 2 * FC_KC * FC_NPAD (16-byte aligned for TMA)
  float* s_h = s_wc + 2 * FC_KC * FC_NPAD;            // FC_HB * 6S
  float* s_lg = s_h + FC_HB * 6 * S;                  // FC_HB * FC_NPAD
  uint64_t* bars = (uint64_t*)(s_lg + FC_HB * FC_NPAD);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b0 = blockIdx.x * FC_HB;
  const int nchunks = (K + FC_KC - 1) / FC_KC;
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int c) {  // one thread: TMA bulk copy of weight chunk c into buffer c&1
    const int kn = min(FC_KC, K - c * FC_KC);
    const uint32_t bytes = (uint32_t)kn * FC_NPAD * 4;
    const uint32_t bar = smem_u32(&bars[c & 1]);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(s_wc + (c & 1) * FC_KC * FC_NPAD), h.fcpT + (size_t)c * FC_KC * FC_NPAD, bytes, bar);
  };
  if (tid == 0) {
    issue(0);
    if (nchunks > 1) issue(1);
  }
  for (int i = tid; i < FC_HB * 6 * S; i += FC_THREADS) {
    const int bi = i / (6 * S);
    s_h[i] = (b0 + bi < nb) ? hbuf[(size_t)b0 * 6 * S + i] : 0.f;
  }
  __syncthreads();
  const int og = tid % (FC_NPAD / 8), bg = tid / (FC_NPAD / 8);
  const bool fc_thread = bg < FC_HB / 4;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    const int kn = min(FC_KC, K - c * FC_KC);
    // parity of the (c/2)-th completion of barrier c&1
    while (!mbar_try(smem_u32(&bars[c & 1]), (c >> 1) & 1)) {
    }
    if (fc_thread) {
      const float* wc = s_wc + (c & 1) * FC_KC * FC_NPAD + og * 8;
      const float* hp = s_h + (4 * bg) * 6 * S + c * FC_KC;
#pragma unroll 4
      for (int kk = 0; kk < kn; ++kk) {
        const float4 w0 = *reinterpret_cast<const float4*>(wc + kk * FC_NPAD);
        const float4 w1 = *reinterpret_cast<const float4*>(wc + kk * FC_NPAD + 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xv = hp[i * 6 * S + kk];
          acc[i][0] = fmaf(xv, w0.x, acc[i][0]); acc[i][1] = fmaf(xv, w0.y, acc[i][1]);
          acc[i][2] = fmaf(xv, w0.z, acc[i][2]); acc[i][3] = fmaf(xv, w0.w, acc[i][3]);
          acc[i][4] = fmaf(xv, w1.x, acc[i][4]); acc[i][5] = fmaf(xv, w1.y, acc[i][5]);
          acc[i][6] = fmaf(xv, w1.z, acc[i][6]); acc[i][7] = fmaf(xv, w1.w, acc[i][7]);
        }
      }
    }
    __syncthreads();  // buffer c&1 fully consumed
    if (tid == 0 && c + 2 < nchunks) issue(c + 2);
  }
  if (fc_thread) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sidx = og * 8 + j;
        s_lg[(4 * bg + i) * FC_NPAD + sidx] = (sidx < S) ? acc[i][j] + h.fcp_bias[sidx] : -INFINITY;
      }
  }
  __syncthreads();
  // softmax (SoftmaxActivation over the S logits) and the value head: one warp per board
  for (int bi = warp; bi < FC_HB; bi += FC_THREADS / 32) {
    const int b = b0 + bi;
    if (b >= nb) continue;
    const float* lg = s_lg + bi * FC_NPAD;
    float mx = -INFINITY;
    for (int i = lane; i < S; i += 32) mx = fmaxf(mx, lg[i]);
    for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(AP_FULL, mx, d));
    float sum = 0.f;
    for (int i = lane; i < S; i += 32) sum += expf(lg[i] - mx);
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(AP_FULL, sum, d);
    const float inv = 1.f / sum;
    for (int i = lane; i < S; i += 32) probs[(size_t)b * S + i] = expf(lg[i] - mx) * inv;
    float part = 0.f;
    const float* hv = s_h + (bi * 6 + 4) * S;
    for (int k = lane; k < 2 * S; k += 32) part = fmaf(h.fcv[k], hv[k], part);
    for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(AP_FULL, part, d);
    if (lane == 0) values[b] = tanhf(part + h.fcv_bias[0]);
  }
}

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core reference path (independent of the tensor-core path; NCHW dense, unfolded BN)
// ------------------------------------------------------------------------------------------
__global__ void k_ref_conv(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ resid,
                           const float* __restrict__ master, long long w, long long b, long long gamma, long long beta,
                           long long mean, long long var, int fix_gamma, int relu, int nb, int cin, int cout, int W, int H,
                           int ksz) {
  const int S = W * H;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nb * cout * S) return;
  int p = (int)(i % S);
  int o = (int)((i / S) % cout);
  int bb = (int)(i / ((long long)S * cout));
  int y = p / W, x = p % W;
  const int r = ksz / 2;
  float acc = 0.f;
  for (int c = 0; c < cin; ++c) {
    const float* ip = in + ((size_t)bb * cin + c) * S;
    const float* wp = master + w + ((long long)o * cin + c) * ksz * ksz;
    for (int ty = 0; ty < ksz; ++ty) {
      int yy = y + ty - r;
      if (yy < 0 || yy >= H) continue;
      for (int tx = 0; tx < ksz; ++tx) {
        int xx = x + tx - r;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(ip[yy * W + xx], wp[ty * ksz + tx], acc);
      }
    }
  }
  acc += master[b + o];
  float v = (acc - master[mean + o]) * (1.f / sqrtf(master[var + o] + BN_EPS));
  if (!fix_gamma) v *= master[gamma + o];
  v += master[beta + o];
  if (resid) v += resid[i];
  if (relu) v = fmaxf(v, 0.f);
  out[i] = v;
}

// one block per board: FCs + softmax/tanh on hp [nb][4][S], hv [nb][2][S]
__global__ void k_ref_heads(const float* __restrict__ hp, const float* __restrict__ hv, const float* __restrict__ master,
                            long long fcp_w, long long fcp_b, long long fcv_w, long long fcv_b, int S,
                            float* __restrict__ probs, float* __restrict__ values) {
  __shared__ float s_red[HEAD_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  float logit = -INFINITY;
  if (tid < S) {
    float acc = 0.f;
    for (int k = 0; k < 4 * S; ++k) acc = fmaf(hp[(size_t)b * 4 * S + k], master[fcp_w + (long long)tid * 4 * S + k], acc);
    logit = acc + master[fcp_b + tid];
  }
  float mx = block_reduce(logit, s_red, true);
  float ex = tid < S ? expf(logit - mx) : 0.f;
  float sum = block_reduce(ex, s_red, false);
  if (tid < S) probs[(size_t)b * S + tid] = ex / sum;
  float part = 0.f;
  for (int k = tid; k < 2 * S; k += HEAD_THREADS) part = fmaf(hv[(size_t)b * 2 * S + k], master[fcv_w + k], part);
  float tot = block_reduce(part, s_red, false);
  if (tid == 0) values[b] = tanhf(tot + master[fcv_b]);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int nalloc(ap_engine* e, NetState* n, void** p, size_t bytes) {
  AP_CUDA(e, cudaMalloc(p, bytes));
  AP_CUDA(e, cudaMemsetAsync(*p, 0, bytes, e->stream));
  n->allocs.push_back(*p);
  e->bytes += bytes;
  return AP_OK;
}

static long long off_of(NetState* n, const std::string& name, long long want_numel, std::string* err) {
  auto it = n->index.find(name);
  if (it == n->index.end()) {
    *err = "missing parameter " + name;
    return -1;
  }
  if (want_numel >= 0 && n->numels[it->second] != want_numel) {
    *err = "parameter " + name + " has " + std::to_string(n->numels[it->second]) + " elements, expected " +
           std::to_string(want_numel);
    return -1;
  }
  return n->offsets[it->second];
}

int net_destroy(ap_engine* e) {
  if (!e->net) return AP_OK;
  for (void* p : e->net->allocs) cudaFree(p);
  delete e->net;
  e->net = nullptr;
  return AP_OK;
}

static int net_prep(ap_engine* e) {
  NetState* n = e->net;
  e->net_generation++;  // captured lock-step graphs hold copies of the prepared head weights
  if (n->vmaster) {
    AP_CUDA(e, cudaMemsetAsync(n->vmaster, 0, (size_t)n->vmaster_numel * 4, e->stream));
    for (size_t i = 0; i + 1 < n->vvar_ranges.size(); i += 2) {
      k_fill_range<<<8, 256, 0, e->stream>>>(n->vmaster, n->vvar_ranges[i], n->vvar_ranges[i + 1], 1.f);
      AP_LAUNCH_CHECK(e);
    }
    for (auto& c : n->vcopies) {
      k_build_virtual<<<64, 256, 0, e->stream>>>(n->master, n->vmaster, c);
      AP_LAUNCH_CHECK(e);
    }
  }
  for (auto& L : n->trunk) {
    const int kc = conv_tc_kc(L, n->split);
    k_prep_conv<<<256, 256, 0, e->stream>>>(L.virt ? n->vmaster : n->master, L.w, L.b, L.gamma, L.beta, L.mean, L.var,
                                            L.fix_gamma, L.cin, L.cin_pad, L.cout, kc, L.ksz * L.ksz, L.post_scale, L.wimg,
                                            L.wimg2, L.wimg_lo, L.wimg4, L.scale, L.shift);
    AP_LAUNCH_CHECK(e);
    if (n->split == 2) {
      k_prep_conv_diffuse<<<(L.cout + 63) / 64, 64, 0, e->stream>>>(L.virt ? n->vmaster : n->master, L.w, L.gamma, L.var,
                                                                   L.fix_gamma, L.cin, L.cout, kc, L.ksz * L.ksz,
                                                                   L.post_scale, L.wimg);
      AP_LAUNCH_CHECK(e);
    }
  }
  k_prep_heads<<<256, 256, 0, e->stream>>>(n->master, n->head, n->S);
  AP_LAUNCH_CHECK(e);
  AP_TRY(fc_tc_prep(e, n));
  // host copy of the folded 1x1 weights for the fused-head conv kernels (kernel-parameter constants)
  memset(&n->head_w, 0, sizeof(n->head_w));
  AP_CUDA(e, cudaMemcpyAsync(n->head_w.w, n->head.w1x1, (size_t)6 * n->head.cfin * 4, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaMemcpyAsync(n->head_w.b, n->head.b1x1, 6 * 4, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

static int conv_alloc(ap_engine* e, NetState* n, ConvLayer& L);
static int conv_bind(ap_engine* e, NetState* n, ConvLayer& L, const std::string& cname, const std::string& bnname,
                     bool conv_act_style, std::string* err) {
  const long long wn = (long long)L.cout * L.cin * 9;
  L.w = off_of(n, cname + "_weight", wn, err);
  L.b = off_of(n, cname + "_bias", L.cout, err);
  if (conv_act_style) {
    L.gamma = off_of(n, cname + "_gamma", L.cout, err);
    L.beta = off_of(n, cname + "_beta", L.cout, err);
    L.mean = off_of(n, cname + "_mean", L.cout, err);
    L.var = off_of(n, cname + "_var", L.cout, err);
    L.fix_gamma = 1;
  } else {
    L.gamma = off_of(n, bnname + "_gamma", L.cout, err);
    L.beta = off_of(n, bnname + "_beta", L.cout, err);
    L.mean = off_of(n, bnname + "_moving_mean", L.cout, err);
    L.var = off_of(n, bnname + "_moving_var", L.cout, err);
    L.fix_gamma = 0;
  }
  if (L.w < 0 || L.b < 0 || L.gamma < 0 || L.beta < 0 || L.mean < 0 || L.var < 0) return AP_ERR_BAD_ARG;
  return conv_alloc(e, n, L);
}

static int conv_alloc(ap_engine* e, NetState* n, ConvLayer& L) {
  L.cin_pad = (L.cin + 15) & ~15;
  AP_TRY(nalloc(e, n, (void**)&L.wimg, (size_t)9 * L.cin_pad * L.cout * 2));
  AP_TRY(nalloc(e, n, (void**)&L.wimg2, (size_t)9 * L.cin_pad * L.cout * 2));
  if (n->split == 1) AP_TRY(nalloc(e, n, (void**)&L.wimg_lo, (size_t)9 * L.cin_pad * L.cout * 2));
  if ((L.cout == 256 || L.cout == 128) && L.cin_pad % 64 == 0 && !n->split) AP_TRY(nalloc(e, n, (void**)&L.wimg4, (size_t)9 * L.cin_pad * L.cout * 2));
  AP_TRY(nalloc(e, n, (void**)&L.scale, (size_t)L.cout * 4));
  AP_TRY(nalloc(e, n, (void**)&L.shift, (size_t)L.cout * 4));
  return AP_OK;
}

extern "C" int ap_net_load(ap_engine* e, int32_t arch, int32_t n_blocks, int32_t n_filter, const ap_tensor* tensors,
                           int32_t n_tensors) {
  AP_ENTER(e);
  if (!tensors || n_tensors <= 0) return ap_fail(e, AP_ERR_BAD_ARG, "no tensors");
  if (e->geo.W != e->geo.H || e->geo.W > 15)
    return ap_fail(e, AP_ERR_BAD_ARG, "the net path needs a square board of width <= 15");
  if ((arch & AP_NET_SPLIT) && (arch & AP_NET_SPLIT_ACT))
    return ap_fail(e, AP_ERR_BAD_ARG, "AP_NET_SPLIT and AP_NET_SPLIT_ACT are exclusive");
  const int split = (arch & AP_NET_SPLIT) ? 1 : ((arch & AP_NET_SPLIT_ACT) ? 2 : 0);
  arch &= ~(AP_NET_SPLIT | AP_NET_SPLIT_ACT);
  if (arch != AP_ARCH_SIMPLE && arch != AP_ARCH_RESNET && arch != AP_ARCH_INCEPTION)
    return ap_fail(e, AP_ERR_BAD_ARG, "unknown arch");
  if (split && arch != AP_ARCH_RESNET)
    return ap_fail(e, AP_ERR_BAD_ARG, "AP_NET_SPLIT / AP_NET_SPLIT_ACT are implemented for the residual net only (the 6-conv net meets 1e-3 in fp16)");
  net_destroy(e);
  NetState* n = new NetState();
  e->net = n;
  n->arch = arch;
  n->split = split;
  n->n_blocks = n_blocks;
  n->n_filter = n_filter;
  n->W = e->geo.W;
  n->H = e->geo.H;
  n->S = e->geo.S;
  cudaDeviceGetAttribute(&n->sm_count, cudaDevAttrMultiProcessorCount, e->cfg.device);
  if (const char* m = getenv("AP_CONV_MODE")) n->conv_mode = (m[0] == '1') ? 1 : (m[0] == '2') ? 2 : 0;
  if (const char* m = getenv("AP_CONV4")) n->conv4 = m[0] - '0';
  if (const char* m = getenv("AP_CONV4_128")) n->conv4_128 = m[0] - '0';
  if (const char* m = getenv("AP_HEAD_PAIR")) n->head_pair = m[0] != '0';
  if (const char* m = getenv("AP_FRONT_FUSED")) n->front_fused = m[0] != '0';
  if (const char* m = getenv("AP_PAIR_CIN64")) n->pair_cin64 = m[0] != '0';
  if (const char* m = getenv("AP_HEAD_MODE")) n->head_mode = (m[0] == '0') ? 0 : (m[0] == '1') ? 1 : 2;
  long long total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    n->names.push_back(tensors[i].name);
    n->offsets.push_back(total);
    n->numels.push_back(tensors[i].numel);
    n->index[tensors[i].name] = i;
    total += (tensors[i].numel + 3) & ~3ll;  // keep every tensor 16-byte aligned
  }
  n->master_numel = total;
  int rc;
  if ((rc = nalloc(e, n, (void**)&n->master, (size_t)total * 4)) != AP_OK) return rc;
  for (int i = 0; i < n_tensors; ++i)
    AP_CUDA(e, cudaMemcpyAsync(n->master + n->offsets[i], tensors[i].data, (size_t)tensors[i].numel * 4,
                               cudaMemcpyHostToDevice, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));

  std::string err;
  auto bad = [&](int code) {
    std::string m = err.empty() ? e->err : err;
    net_destroy(e);
    return ap_fail(e, code, m);
  };
  if (arch == AP_ARCH_SIMPLE) {
    const char* nm[6] = {"conv1", "conv2", "conv3", "conv4", "conv5", "conv_final"};
    const int co[6] = {64, 64, 128, 128, 256, 256};
    int cin = 9, buf = -1;
    for (int i = 0; i < 6; ++i) {
      ConvLayer L{};
      L.cin = cin;
      L.cout = co[i];
      L.relu = 1;
      L.in_buf = buf;
      L.out_buf = (buf == 0) ? 1 : 0;
      L.resid_buf = -1;
      if ((rc = conv_bind(e, n, L, nm[i], "", true, &err)) != AP_OK) return bad(rc);
      n->trunk.push_back(L);
      cin = co[i];
      buf = L.out_buf;
    }
    n->final_buf = buf;
    n->head.cfin = cin;
  } else if (arch == AP_ARCH_INCEPTION) {
    // builder-defined board-sized Inception-ResNet: 3x3 stem + n_blocks x block35 (inception-resnet-v2.py:41-58), see
    // alphapig_b200/params.py.  Per block, on buffers X (block input), T0, T1, Y:
    //   mix1 : 1x1  X(128) -> T0 = [t0 32 | t1a 32 | t2a 32 | 0 32]      the three tower stems in one layer
    //   t1b  : 3x3  T0 (input channels 32..63) -> T0[32:64]               in place: a tile reads its board before it writes
    //   t2b  : 3x3  T0 (input channels 64..95) -> T1 (48 of 64)
    //   t2c  : 3x3  T1 -> T0[64:128]      T0 is now the concat [t0 | t1b | t2c]
    //   up   : 1x1  T0(128) -> Y = relu(X + 0.17 * BN(conv))
    if (n_blocks < 1 || n_filter != 128) return bad(ap_fail(e, AP_ERR_BAD_ARG, "inception: n_filter must be 128, n_blocks >= 1"));
    ConvLayer L{};
    L.cin = 9; L.cout = 128; L.relu = 1; L.in_buf = -1; L.out_buf = 0; L.resid_buf = -1; L.force_single = 1;
    if ((rc = conv_bind(e, n, L, "incep_conv1", "", true, &err)) != AP_OK) return bad(rc);
    n->trunk.push_back(L);
    long long vtot = 0;
    auto valloc = [&](long long cnt) { long long o = vtot; vtot += (cnt + 3) & ~3ll; return o; };
    struct Piece { const char* name; int cout, cin, oo, io; };
    auto vlayer = [&](int blk, int cin_v, int cout_v, int ksz, std::vector<Piece> pieces, int in_buf, int out_buf,
                      int out_coff, int cout_store, int resid_buf, int relu, float post, int in_coff = 0) -> int {
      ConvLayer V{};
      V.in_coff = in_coff;
      V.cin = cin_v; V.cout = cout_v; V.ksz = ksz; V.relu = relu; V.in_buf = in_buf; V.out_buf = out_buf;
      V.resid_buf = resid_buf; V.out_coff = out_coff; V.cout_store = cout_store; V.force_single = 1; V.virt = 1;
      V.fix_gamma = 1; V.post_scale = post;
      const int taps = ksz * ksz;
      V.w = valloc((long long)cout_v * cin_v * taps);
      V.b = valloc(cout_v); V.beta = valloc(cout_v); V.mean = valloc(cout_v); V.var = valloc(cout_v);
      V.gamma = V.beta;  // unused (fix_gamma)
      n->vvar_ranges.push_back(V.var);
      n->vvar_ranges.push_back(cout_v);
      for (auto& pc : pieces) {
        const std::string nm = "b35_" + std::to_string(blk) + "_" + pc.name;
        NetState::VCopy c{};
        c.dst_w = V.w; c.dst_b = V.b; c.dst_beta = V.beta; c.dst_mean = V.mean; c.dst_var = V.var;
        c.src_w = off_of(n, nm + "_weight", (long long)pc.cout * pc.cin * taps, &err);
        c.src_b = off_of(n, nm + "_bias", pc.cout, &err);
        c.src_beta = off_of(n, nm + "_beta", pc.cout, &err);
        c.src_mean = off_of(n, nm + "_mean", pc.cout, &err);
        c.src_var = off_of(n, nm + "_var", pc.cout, &err);
        off_of(n, nm + "_gamma", pc.cout, &err);  // present in the table (fix_gamma: value unused)
        if (!err.empty()) return AP_ERR_BAD_ARG;
        c.cout = pc.cout; c.cin = pc.cin; c.cin_v = cin_v; c.taps = taps; c.oo = pc.oo; c.io = pc.io;
        n->vcopies.push_back(c);
      }
      int r2 = conv_alloc(e, n, V);
      if (r2 != AP_OK) return r2;
      n->trunk.push_back(V);
      return AP_OK;
    };
    int x = 0;  // buffers: 0 / 1 = block input / output (ping-pong), 2 = T0, 3 = T1
    for (int i = 1; i <= n_blocks; ++i) {
      const int y = 1 - x;
      if ((rc = vlayer(i, 128, 128, 1, {{"t0", 32, 128, 0, 0}, {"t1a", 32, 128, 32, 0}, {"t2a", 32, 128, 64, 0}}, x, 2, 0, 96, -1, 1, 1.f)) != AP_OK) return bad(rc);
      // tower layers read only their own channel slice (in_coff) - K = 9 * 32 / 9 * 48 instead of 9 * 128
      if ((rc = vlayer(i, 32, 64, 3, {{"t1b", 32, 32, 0, 0}}, 2, 2, 32, 32, -1, 1, 1.f, 32)) != AP_OK) return bad(rc);
      if ((rc = vlayer(i, 32, 64, 3, {{"t2b", 48, 32, 0, 0}}, 2, 3, 0, 48, -1, 1, 1.f, 64)) != AP_OK) return bad(rc);
      if ((rc = vlayer(i, 48, 64, 3, {{"t2c", 64, 48, 0, 0}}, 3, 2, 64, 64, -1, 1, 1.f)) != AP_OK) return bad(rc);
      if ((rc = vlayer(i, 128, 128, 1, {{"up", 128, 128, 0, 0}}, 2, y, 0, 128, x, 1, 0.17f)) != AP_OK) return bad(rc);
      x = y;
    }
    n->vmaster_numel = vtot;
    if ((rc = nalloc(e, n, (void**)&n->vmaster, (size_t)vtot * 4)) != AP_OK) return bad(rc);
    n->final_buf = x;
    n->head.cfin = 128;
    n->head_mode = n->head_mode == 0 ? 0 : 1;  // no fused-head instantiation for the 1x1 up-projection
  } else {
    if (n_blocks < 0 || n_filter != 128)
      return bad(ap_fail(e, AP_ERR_BAD_ARG, "resnet: n_filter must be 128 (stem is hard-coded 128, policy_value_net_mxnet.py:73)"));
    ConvLayer L{};
    L.cin = 9;
    L.cout = 128;
    L.relu = 1;
    L.in_buf = -1;
    L.out_buf = 0;
    L.resid_buf = -1;
    if ((rc = conv_bind(e, n, L, "res_conv1", "", true, &err)) != AP_OK) return bad(rc);
    n->trunk.push_back(L);
    int x = 0;
    for (int i = 1; i <= n_blocks; ++i) {
      int t = (x + 1) % 3, y = (x + 2) % 3;
      ConvLayer A{};
      A.cin = 128; A.cout = n_filter; A.relu = 1; A.in_buf = x; A.out_buf = t; A.resid_buf = -1;
      if ((rc = conv_bind(e, n, A, "convA" + std::to_string(i), "bnA" + std::to_string(i), false, &err)) != AP_OK)
        return bad(rc);
      n->trunk.push_back(A);
      ConvLayer B{};
      B.cin = n_filter; B.cout = n_filter; B.relu = 1; B.in_buf = t; B.out_buf = y; B.resid_buf = x;
      if ((rc = conv_bind(e, n, B, "convB" + std::to_string(i), "bnB" + std::to_string(i), false, &err)) != AP_OK)
        return bad(rc);
      n->trunk.push_back(B);
      x = y;
    }
    n->final_buf = x;
    n->head.cfin = n_filter;
  }
  HeadParams& h = n->head;
  const int S = n->S;
  h.pw = off_of(n, "conv3_1_1_weight", 4ll * h.cfin, &err);
  h.pb = off_of(n, "conv3_1_1_bias", 4, &err);
  h.pgamma = off_of(n, "conv3_1_1_gamma", 4, &err);
  h.pbeta = off_of(n, "conv3_1_1_beta", 4, &err);
  h.pmean = off_of(n, "conv3_1_1_mean", 4, &err);
  h.pvar = off_of(n, "conv3_1_1_var", 4, &err);
  h.vw = off_of(n, "conv3_2_1_weight", 2ll * h.cfin, &err);
  h.vb = off_of(n, "conv3_2_1_bias", 2, &err);
  h.vgamma = off_of(n, "conv3_2_1_gamma", 2, &err);
  h.vbeta = off_of(n, "conv3_2_1_beta", 2, &err);
  h.vmean = off_of(n, "conv3_2_1_mean", 2, &err);
  h.vvar = off_of(n, "conv3_2_1_var", 2, &err);
  h.fcp_w = off_of(n, "fc_3_1_1_weight", 4ll * S * S, &err);
  h.fcp_b = off_of(n, "fc_3_1_1_bias", S, &err);
  h.fcv_w = off_of(n, "fc_3_2_1_weight", 2ll * S, &err);
  h.fcv_b = off_of(n, "fc_3_2_1_bias", 1, &err);
  if (!err.empty()) return bad(AP_ERR_BAD_ARG);
  if ((rc = nalloc(e, n, (void**)&h.w1x1, 6ull * h.cfin * 4)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&h.b1x1, 6 * 4)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&h.fcpT, 4ull * S * 232 * 4)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&h.fcp_bias, (size_t)S * 4)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&h.fcv, 2ull * S * 4)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&h.fcv_bias, 4)) != AP_OK) return bad(rc);

  // activation planes for max(G, 256) boards
  n->bcap = e->geo.G > 256 ? e->geo.G : 256;
  n->mpad = 2ll * NET_PAD_ROWS + (long long)n->bcap * NET_TILE_ROWS;
  int maxc = 0;
  for (auto& L : n->trunk) maxc = L.cout > maxc ? L.cout : maxc;
  const int nbuf = (arch == AP_ARCH_INCEPTION) ? 4 : (arch == AP_ARCH_RESNET) ? 3 : 2;
  if ((rc = nalloc(e, n, (void**)&n->feat, (size_t)2 * n->mpad * 16)) != AP_OK) return bad(rc);
  for (int i = 0; i < nbuf; ++i)
    if ((rc = nalloc(e, n, (void**)&n->act[i], (size_t)(maxc / 8) * n->mpad * 16)) != AP_OK) return bad(rc);
  if (n->split) {
    if ((rc = nalloc(e, n, (void**)&n->feat_lo, (size_t)2 * n->mpad * 16)) != AP_OK) return bad(rc);
    for (int i = 0; i < nbuf; ++i)
      if ((rc = nalloc(e, n, (void**)&n->act_lo[i], (size_t)(maxc / 8) * n->mpad * 16)) != AP_OK) return bad(rc);
    for (auto& L : n->trunk)
      if (!conv_tc_split_supported(L)) return bad(ap_fail(e, AP_ERR_BAD_ARG, "split precision: unsupported layer shape"));
  }
  n->bcap_ref = 128;
  const size_t refb = (size_t)n->bcap_ref * 256 * S * 4;
  if ((rc = nalloc(e, n, (void**)&n->ref_a, refb)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&n->ref_b, refb)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&n->ref_c, refb)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&n->ref_in, (size_t)e->geo.G * 9 * S * 4 > (size_t)n->bcap_ref * 9 * S * 4
                                                   ? (size_t)e->geo.G * 9 * S * 4
                                                   : (size_t)n->bcap_ref * 9 * S * 4)) != AP_OK)
    return bad(rc);
  if ((rc = nalloc(e, n, (void**)&n->d_err, 4)) != AP_OK) return bad(rc);
  if ((rc = conv_tc_configure(e)) != AP_OK) return bad(rc);
  if ((rc = front_tc_configure(e)) != AP_OK) return bad(rc);
  if (cudaFuncSetAttribute(k_head_fc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
    return bad(ap_fail(e, AP_ERR_CUDA, "cudaFuncSetAttribute(k_head_fc)"));
  if ((rc = nalloc(e, n, (void**)&n->hbuf, (size_t)n->bcap * 6 * S * 4)) != AP_OK) return bad(rc);
  fc_tc_dims(S, &n->fc_kp, &n->fc_np);
  n->fc_rows = ((long long)n->bcap + 127) / 128 * 128;
  if ((rc = nalloc(e, n, (void**)&n->fc_a, (size_t)2 * (n->fc_kp / 8) * n->fc_rows * 16)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&n->fc_w, (size_t)2 * (n->fc_kp / 8) * n->fc_np * 16)) != AP_OK) return bad(rc);
  if ((rc = nalloc(e, n, (void**)&n->fc_bias, (size_t)n->fc_np * 4)) != AP_OK) return bad(rc);
  {
    const char* ks = getenv("AP_FC_KSPLIT");  // 1 = single-pass FC with the softmax in the GEMM epilogue (A/B)
    n->fc_ksplit = ks ? atoi(ks) : 4;
    if (n->fc_ksplit < 1 || n->fc_ksplit > 8) n->fc_ksplit = 4;
    if (n->fc_ksplit > n->fc_kp / 32) n->fc_ksplit = n->fc_kp / 32;  // at least one K stage (32) per split
  }
  if (n->fc_ksplit > 1 &&
      (rc = nalloc(e, n, (void**)&n->fc_partial, (size_t)n->fc_ksplit * n->fc_rows * n->fc_np * 4)) != AP_OK)
    return bad(rc);
  if ((rc = fc_tc_configure(e, n)) != AP_OK) return bad(rc);
  if (n->head_mode == 2 && !conv_tc_head_supported(n->trunk.back())) n->head_mode = 1;
  if (n->split && n->head_mode != 2)
    return bad(ap_fail(e, AP_ERR_BAD_ARG, "split precision needs the fused head epilogue (n_blocks >= 1, AP_HEAD_MODE=2)"));
  if ((rc = net_prep(e)) != AP_OK) return bad(rc);
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

extern "C" int ap_net_refresh(ap_engine* e) {
  AP_ENTER(e);
  if (!e->net) return ap_fail(e, AP_ERR_NO_NET, "no net loaded");
  AP_TRY(net_prep(e));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return AP_OK;
}

extern "C" int ap_net_weights_ptr(ap_engine* e, void** out_dev_ptr, int64_t* out_numel) {
  AP_ENTER(e);
  if (!e->net) return ap_fail(e, AP_ERR_NO_NET, "no net loaded");
  *out_dev_ptr = e->net->master;
  *out_numel = e->net->master_numel;
  return AP_OK;
}

extern "C" int ap_net_layout(ap_engine* e, int32_t cap, const char** out_names, int64_t* out_offsets, int64_t* out_numels) {
  AP_ENTER(e);
  if (!e->net) return ap_fail(e, AP_ERR_NO_NET, "no net loaded");
  int n = (int)e->net->names.size();
  for (int i = 0; i < n && i < cap; ++i) {
    out_names[i] = e->net->names[i].c_str();
    out_offsets[i] = e->net->offsets[i];
    out_numels[i] = e->net->numels[i];
  }
  return n;
}

// trunk + heads on the first nb boards of the feature planes -> probs/values (device)
// compacted leaf batches need the fused head path (device-side tile counts in every kernel of the forward)
bool net_can_compact(ap_engine* e) { return e->net && e->net->head_mode == 2; }
// partial logits of the split-K FC for a consumer that finishes them itself (partial == nullptr: not available)
void net_fc_finish_args(ap_engine* e, const float** partial, const float** bias, long long* rows, int* np, int* ksplit) {
  NetState* n = e->net;
  const bool on = n && n->head_mode == 2 && n->fc_ksplit > 1;
  *partial = on ? n->fc_partial : nullptr;
  *bias = on ? n->fc_bias : nullptr;
  *rows = on ? n->fc_rows : 0;
  *np = on ? n->fc_np : 0;
  *ksplit = on ? n->fc_ksplit : 0;
}
void net_feature_planes(ap_engine* e, __half** feat, long long* mpad) {
  *feat = e->net->feat;
  *mpad = e->net->mpad;
}
int net_phase_count(ap_engine* e) { return e->net ? (int)e->net->trunk.size() + 2 : 0; }

// nb_dev != nullptr: the number of boards is read on the device (compacted leaf batch, at most nb)
static int run_fast(ap_engine* e, int nb, float* d_probs, float* d_values, const int32_t* nb_dev = nullptr,
                    bool skip_finish = false) {
  NetState* n = e->net;
  const size_t last = n->trunk.size() - 1;
  if (nb_dev && n->head_mode != 2) return ap_fail(e, AP_ERR_BAD_ARG, "compacted batches need the fused head path");
  size_t first = 0;
  if (n->front_fused && front_tc_supported(n)) {
    // conv1 + conv2 in one kernel: conv1's output goes TMEM -> shared memory -> conv2's MMAs, never to HBM
    AP_TRY(front_tc_launch(e, n, nb, nb_dev));
    prof_mark(e);  // phase "conv1" (empty: the fused kernel is accounted to conv2)
    prof_mark(e);
    first = 2;
  }
  for (size_t i = first; i < n->trunk.size(); ++i) {
    // head_mode 2: the last trunk layer also computes the two 1x1 head convs and never stores its own output
    AP_TRY(conv_tc_launch(e, n, n->trunk[i], nb, nb_dev, n->head_mode == 2 && i == last));
    prof_mark(e);
  }
  const int S = n->S;
  if (n->head_mode != 2) {
    k_head_conv<<<nb, 256, (size_t)6 * n->head.cfin * 4, e->stream>>>(n->act[n->final_buf], n->mpad, n->head.cfin, n->W,
                                                                    n->H, n->head, n->hbuf, n->head_mode ? n->fc_a : nullptr,
                                                                    n->fc_rows, n->fc_kp / 8);
    AP_LAUNCH_CHECK(e);
  }
  if (n->head_mode == 0) {
    size_t smem = (size_t)(2 * FC_KC * FC_NPAD + FC_HB * 6 * S + FC_HB * FC_NPAD) * 4 + 16;
    k_head_fc<<<(nb + FC_HB - 1) / FC_HB, FC_THREADS, smem, e->stream>>>(n->hbuf, S, nb, n->head, d_probs, d_values);
    AP_LAUNCH_CHECK(e);
    return AP_OK;
  }
  return fc_tc_launch(e, n, nb, d_probs, d_values, nb_dev, skip_finish);
}

int net_board_capacity(ap_engine* e) { return e->net ? e->net->bcap : 0; }
// the tensor-core path on a compacted batch whose feature planes are already written (at most nb_max boards, the
// count read on the device); the split-K FC partial logits are left for the consumer (net_fc_finish_args)
int net_run_compacted(ap_engine* e, int nb_max, const int32_t* nb_dev) {
  NetState* n = e->net;
  if (!n || n->head_mode != 2 || n->fc_ksplit <= 1) return ap_fail(e, AP_ERR_BAD_ARG, "needs the fused-head split-K path");
  return run_fast(e, nb_max, nullptr, nullptr, nb_dev, true);
}

// fp32 path on dense NCHW states (device) for nb <= bcap_ref boards
static int run_ref(ap_engine* e, const float* d_states, int nb, float* d_probs, float* d_values) {
  NetState* n = e->net;
  if (n->arch == AP_ARCH_INCEPTION)
    return ap_fail(e, AP_ERR_BAD_ARG, "the fp32 CUDA-core cross-check path is not built for the inception variant");
  const int S = n->S, W = n->W, H = n->H;
  float* bufs[3] = {n->ref_a, n->ref_b, n->ref_c};
  auto conv = [&](const float* in, float* out, const float* resid, long long w, long long b, long long gamma,
                  long long beta, long long mean, long long var, int fix_gamma, int relu, int cin, int cout, int ksz) {
    long long tot = (long long)nb * cout * S;
    k_ref_conv<<<(unsigned)((tot + 255) / 256), 256, 0, e->stream>>>(in, out, resid, n->master, w, b, gamma, beta, mean,
                                                                    var, fix_gamma, relu, nb, cin, cout, W, H, ksz);
    e->launches++;
  };
  for (auto& L : n->trunk) {
    const float* in = (L.in_buf < 0) ? d_states : bufs[L.in_buf];
    conv(in, bufs[L.out_buf], L.resid_buf >= 0 ? bufs[L.resid_buf] : nullptr, L.w, L.b, L.gamma, L.beta, L.mean, L.var,
         L.fix_gamma, L.relu, L.cin, L.cout, 3);
  }
  const float* fin = bufs[n->final_buf];
  float* hp = bufs[(n->final_buf + 1) % 3];
  float* hv = hp + (size_t)nb * 4 * S;
  const HeadParams& h = n->head;
  conv(fin, hp, nullptr, h.pw, h.pb, h.pgamma, h.pbeta, h.pmean, h.pvar, 1, 1, h.cfin, 4, 1);
  conv(fin, hv, nullptr, h.vw, h.vb, h.vgamma, h.vbeta, h.vmean, h.vvar, 1, 1, h.cfin, 2, 1);
  k_ref_heads<<<nb, HEAD_THREADS, 0, e->stream>>>(hp, hv, n->master, h.fcp_w, h.fcp_b, h.fcv_w, h.fcv_b, S, d_probs,
                                                 d_values);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}

int net_check_err(ap_engine* e) {
  int h = 0;
  AP_CUDA(e, cudaMemcpyAsync(&h, e->net->d_err, 4, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  if (h) {
    cudaMemsetAsync(e->net->d_err, 0, 4, e->stream);
    return ap_fail(e, AP_ERR_CUDA, "conv kernel: mbarrier wait timed out (pipeline protocol error)");
  }
  return AP_OK;
}

int net_emit_features_launch(ap_engine* e, bool compact) {
  NetState* n = e->net;
  k_emit_features<<<(e->geo.G + 3) / 4, 128, 0, e->stream>>>(e->geo, e->leaves.rows, e->leaves.meta, e->geo.G,
                                                            compact ? e->leaves.game_of_slot : nullptr,
                                                            compact ? e->leaves.n_eval : nullptr, n->feat, n->mpad);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}

// leaves of the last select -> e->d_probs / e->d_values
int net_forward_leaves(ap_engine* e, int precise, bool compact, bool compacted_by_select) {
  if (!e->net) return ap_fail(e, AP_ERR_NO_NET, "no net loaded");
  NetState* n = e->net;
  const int G = e->geo.G;
  if (!precise) {
    compact = compact && n->head_mode == 2;
    if (compact && !compacted_by_select) {
      launch_compact_leaves(e);
      AP_LAUNCH_CHECK(e);
    }
    // compacted_by_select: k_select also wrote the feature planes of its leaves (the board was still in registers)
    if (!(compact && compacted_by_select)) AP_TRY(net_emit_features_launch(e, compact));
    prof_mark(e);
    // compacted_by_select = the fused lock-step of ap_search_run: k_expand_backup finishes the split-K FC itself
    AP_TRY(run_fast(e, G, e->d_probs, e->d_values, compact ? e->leaves.n_eval : nullptr,
                    compact && compacted_by_select && n->fc_ksplit > 1));
    prof_mark(e);
    return AP_OK;
  }
  launch_boards_features(e, e->leaves.rows, e->leaves.meta, nullptr, G, n->ref_in);
  e->launches += 2;
  for (int b0 = 0; b0 < G; b0 += n->bcap_ref) {
    int nb = (G - b0 < n->bcap_ref) ? G - b0 : n->bcap_ref;
    AP_TRY(run_ref(e, n->ref_in + (size_t)b0 * 9 * n->S, nb, e->d_probs + (size_t)b0 * n->S, e->d_values + b0));
  }
  return AP_OK;
}

extern "C" int ap_net_forward_leaves(ap_engine* e, int32_t precise, float* out_probs, float* out_values) {
  AP_ENTER(e);
  AP_TRY(net_forward_leaves(e, precise));
  const size_t G = e->geo.G, S = e->geo.S;
  if (out_probs) AP_CUDA(e, cudaMemcpyAsync(out_probs, e->d_probs, G * S * 4, cudaMemcpyDeviceToHost, e->stream));
  if (out_values) AP_CUDA(e, cudaMemcpyAsync(out_values, e->d_values, G * 4, cudaMemcpyDeviceToHost, e->stream));
  AP_CUDA(e, cudaStreamSynchronize(e->stream));
  return net_check_err(e);
}

// PolicyValueNet.policy_value(state_batch): host fp32 in, host fp32 out.
// flags via B sign is avoided: precise path is selected with ap_net_forward_precise.
static int forward_host(ap_engine* e, const float* states, int32_t B, float* out_probs, float* out_values, int precise) {
  if (!e->net) return ap_fail(e, AP_ERR_NO_NET, "no net loaded");
  if (!states || B <= 0 || !out_probs || !out_values) return ap_fail(e, AP_ERR_BAD_ARG, "bad argument");
  NetState* n = e->net;
  const int S = n->S;
  const int chunk = precise ? n->bcap_ref : n->bcap;
  const size_t per = (size_t)9 * S * 4;
  AP_TRY(ap_stage(e, (size_t)chunk * per + (size_t)chunk * (S + 1) * 4, 0));
  float* d_st = (float*)e->d_stage;
  float* d_pr = (float*)((char*)e->d_stage + (size_t)chunk * per);
  float* d_va = d_pr + (size_t)chunk * S;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    int nb = (B - b0 < chunk) ? B - b0 : chunk;
    AP_CUDA(e, cudaMemcpyAsync(d_st, states + (size_t)b0 * 9 * S, (size_t)nb * per, cudaMemcpyHostToDevice, e->stream));
    if (precise) {
      AP_TRY(run_ref(e, d_st, nb, d_pr, d_va));
    } else {
      k_pack_states<<<(nb * S + 255) / 256, 256, 0, e->stream>>>(d_st, nb, n->W, n->H, n->feat, n->mpad);
      AP_LAUNCH_CHECK(e);
      AP_TRY(run_fast(e, nb, d_pr, d_va));
    }
    AP_CUDA(e, cudaMemcpyAsync(out_probs + (size_t)b0 * S, d_pr, (size_t)nb * S * 4, cudaMemcpyDeviceToHost, e->stream));
    AP_CUDA(e, cudaMemcpyAsync(out_values + b0, d_va, (size_t)nb * 4, cudaMemcpyDeviceToHost, e->stream));
    AP_CUDA(e, cudaStreamSynchronize(e->stream));
  }
  return net_check_err(e);
}

extern "C" int ap_net_forward(ap_engine* e, const float* states, int32_t B, float* out_probs, float* out_values) {
  AP_ENTER(e);
  return forward_host(e, states, B, out_probs, out_values, 0);
}

extern "C" int ap_net_forward_precise(ap_engine* e, const float* states, int32_t B, float* out_probs, float* out_values) {
  AP_ENTER(e);
  return forward_host(e, states, B, out_probs, out_values, 1);
}
