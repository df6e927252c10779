// Policy + value fully-connected heads as ONE tcgen05 GEMM with near-fp32 accuracy.
// Replaces FullyConnected(4S -> S) + SoftmaxActivation and FullyConnected(2S -> 1) + tanh of the
// reference heads (policy_value_net_mxnet_simple.py:78-90; policy_value_net_mxnet.py:85-97).
//
//   logits[b][n] = sum_k X[b][k] * Wbig[k][n],  k = o*S + pixel over the 6 head-conv channels (K = 6S),
//   Wbig[k][n] = fc_3_1_1_weight[n][k]        n <  S, k <  4S   (policy)
//              = fc_3_2_1_weight[0][k - 4S]   n == S, k >= 4S   (value)
//              = 0                            elsewhere
//
// fp16 operands alone would cost ~4e-4 on a logit (K = 900 products, 11-bit significands) - half of the
// 1e-3 parity budget - so BOTH operands are split x = hi + lo (two fp16 values, 22 significant bits)
// and the GEMM runs three tensor-core products per K step: hi*hi + lo*hi + hi*lo (the dropped lo*lo term
// is 2^-22 relative).  It is still only 3 * 2*128*240*1376 = 254 MFLOP per 128 boards.
//
// A operand (written by the conv epilogue / k_head_conv): [2 (hi,lo)][Kp/8][rows][8] fp16 - the
//   K-major no-swizzle core-matrix image, so a K chunk of 128 boards is KC/8 contiguous 2 KB pieces.
// B operand (k_prep_fc): [2][Kp/8][Np][8] fp16, one contiguous piece per K chunk.
// One CTA = 128 boards: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (thread = board row:
// the softmax over the S logits of a board is thread-local in TMEM).
// Split K (default 4, AP_FC_KSPLIT): 4096 boards are only 32 tiles, a quarter of the SMs, and every CTA streams
// the whole 2 MB operand set through one SM's L2 port.  With gridDim.y = 4 each CTA takes a quarter of the K
// chunks and stores its raw fp32 partial logits ([split][board][np], L2-resident); k_head_fc_finish (one warp
// per board) adds the partials in fixed order, the bias, and does the softmax / tanh.
#include "fc_finish.cuh"
#include "kernels.h"
#include "net.h"
#include "ptx.cuh"

namespace {

constexpr int kFcKC = 32;        // K per pipeline stage
constexpr int kFcStages = 4;
constexpr int kFcThreads = 192;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct FcParams {
  const __half* a;   // [2][kg][rows][8]
  const __half* w;   // [2][kg][np][8]
  const float* bias; // [np]: fc_3_1_1_bias | fc_3_2_1_bias | 0
  float* probs;      // [nb][S]
  float* values;     // [nb]
  long long rows;
  int kg, np, S, nb;
  const int* nb_dev;  // when non-null the number of boards is read from device memory (compacted leaf batches)
  int* errflag;
  float* partial;    // ksplit > 1: [ksplit][rows][np] raw partial logits
  int ksplit;
};

__global__ void __launch_bounds__(kFcThreads, 1) k_head_fc_tc(FcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(AP_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t a_bytes = (kFcKC / 8) * 128 * 16;            // one of hi / lo
  const uint32_t b_bytes = (uint32_t)(kFcKC / 8) * p.np * 16;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* full = (uint64_t*)(smem + (size_t)kFcStages * stage_bytes);
  uint64_t* empty = full + kFcStages;
  uint64_t* acc_full = empty + kFcStages;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  float* s_bias = (float*)(tmem_slot + 2);
  const int nchunks_all = p.kg / (kFcKC / 8);
  const int c_begin = (int)((long long)nchunks_all * blockIdx.y / p.ksplit);
  const int c_end = (int)((long long)nchunks_all * (blockIdx.y + 1) / p.ksplit);
  const int nchunks = c_end - c_begin;
  const int b0 = blockIdx.x * 128;
  const int nb = p.nb_dev ? *p.nb_dev : p.nb;
  if (b0 >= nb) return;  // whole CTA, before any barrier / TMEM allocation

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFcStages; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.np; i += kFcThreads) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    int st = 0, ph = 0;
    bool ok = true;
    for (int c = 0; c < nchunks && ok; ++c) {
      ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&empty[st]), ph ^ 1, p.errflag));
      if (!ok) break;
      const uint32_t fb = smem_u32(&full[st]);
      uint8_t* sb = smem + (size_t)st * stage_bytes;
      if (elect_one()) {
        mbar_expect_tx(fb, stage_bytes);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
          for (int j = 0; j < kFcKC / 8; ++j)
            bulk_g2s(smem_u32(sb + part * a_bytes + j * 2048),
                     p.a + (((long long)part * p.kg + ((c_begin + c) * (kFcKC / 8) + j)) * p.rows + b0) * 8, 2048, fb);
          bulk_g2s(smem_u32(sb + 2 * a_bytes + part * b_bytes),
                   p.w + ((long long)part * p.kg + (c_begin + c) * (kFcKC / 8)) * p.np * 8, b_bytes, fb);
        }
      }
      __syncwarp();
      if (++st == kFcStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.np >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t a_lbo = 2048u >> 4, b_lbo = (uint32_t)p.np;  // bytes >> 4 between K groups
    int st = 0, ph = 0;
    bool ok = true;
    for (int c = 0; c < nchunks && ok; ++c) {
      ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&full[st]), ph, p.errflag));
      if (!ok) break;
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
      const uint32_t ahi = (sa >> 4) | (a_lbo << 16), alo = ((sa + a_bytes) >> 4) | (a_lbo << 16);
      const uint32_t bhi = ((sa + 2 * a_bytes) >> 4) | (b_lbo << 16), blo = ((sa + 2 * a_bytes + b_bytes) >> 4) | (b_lbo << 16);
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < kFcKC / 16; ++j) {
          const uint32_t ao = 2 * j * a_lbo, bo = 2 * j * b_lbo;
          tc_mma_f16(tmem_base, desc_hi | (uint64_t)(ahi + ao), desc_hi | (uint64_t)(bhi + bo), idesc, (c | j) != 0);
          tc_mma_f16(tmem_base, desc_hi | (uint64_t)(alo + ao), desc_hi | (uint64_t)(bhi + bo), idesc, 1);
          tc_mma_f16(tmem_base, desc_hi | (uint64_t)(ahi + ao), desc_hi | (uint64_t)(blo + bo), idesc, 1);
        }
        tc_commit(smem_u32(&empty[st]));
        if (c == nchunks - 1) tc_commit(smem_u32(acc_full));
      }
      __syncwarp();
      if (++st == kFcStages) { st = 0; ph ^= 1; }
    }
  } else {
    // ===== epilogue: thread = board; softmax over its S logits, tanh of the value logit =====
    const int q = warp & 3;
    bool ok = mbar_wait(smem_u32(acc_full), 0, p.errflag);
    ok = __all_sync(AP_FULL, ok);
    if (ok) {
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16);
      const int b = b0 + q * 32 + lane;
      const int S = p.S, nch = p.np >> 4;
      uint32_t v[16];
      if (p.ksplit > 1) {
        // raw partial logits of this K range; rows beyond nb are padding of the last tile and are not read back
        float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.y * p.rows + b) * p.np);
        for (int ch = 0; ch < nch; ++ch) {
          tmem_ld16(acc + ch * 16, v);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[ch * 4 + i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                          __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      } else {
      float mx = -INFINITY;
      for (int ch = 0; ch < nch; ++ch) {
        tmem_ld16(acc + ch * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = ch * 16 + i;
          if (n < S) mx = fmaxf(mx, __uint_as_float(v[i]) + s_bias[n]);
        }
      }
      float sum = 0.f, vlogit = 0.f;
      for (int ch = 0; ch < nch; ++ch) {
        tmem_ld16(acc + ch * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = ch * 16 + i;
          const float l = __uint_as_float(v[i]) + s_bias[n];
          if (n < S) sum += expf(l - mx);
          if (n == S) vlogit = l;
        }
      }
      const float inv = 1.f / sum;
      if (b < nb) p.values[b] = tanhf(vlogit);
      for (int ch = 0; ch < nch; ++ch) {
        tmem_ld16(acc + ch * 16, v);
        if (b < nb) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = ch * 16 + i;
            if (n < S) p.probs[(size_t)b * S + n] = expf(__uint_as_float(v[i]) + s_bias[n] - mx) * inv;
          }
        }
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// split-K second half: one warp per board; logits = sum of the partials (fixed order) + bias, softmax over the S
// policy logits, tanh of the value logit (SoftmaxActivation / tanh of ..._simple.py:84,90)
__global__ void __launch_bounds__(128) k_head_fc_finish(FcParams p) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int nb = p.nb_dev ? *p.nb_dev : p.nb;
  if (b >= nb) return;
  const FcFinish f{p.partial, p.bias, p.rows, p.np, p.ksplit};
  float pr[8];
  const float v = fc_finish_warp(f, b, p.S, lane, pr);
  if (lane == 0) p.values[b] = v;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = lane + 32 * j;
    if (n < p.S) p.probs[(size_t)b * p.S + n] = pr[j];
  }
}

// fp32 master weights -> split-fp16 B image + bias row (run at load and after every weight refresh)
__global__ void k_prep_fc(const float* __restrict__ master, long long fcp_w, long long fcp_b, long long fcv_w,
                          long long fcv_b, int S, int kg, int np, __half* w, float* bias) {
  const long long total = (long long)kg * np * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    const int n = (int)((i >> 3) % np);
    const int k = (int)((i >> 3) / np) * 8 + e;
    float x = 0.f;
    if (n < S && k < 4 * S) x = master[fcp_w + (long long)n * 4 * S + k];
    else if (n == S && k >= 4 * S && k < 6 * S) x = master[fcv_w + (k - 4 * S)];
    const __half hi = __float2half_rn(x);
    w[i] = hi;
    w[total + i] = __float2half_rn(x - __half2float(hi));
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < np; n += gridDim.x * blockDim.x)
    bias[n] = n < S ? master[fcp_b + n] : (n == S ? master[fcv_b] : 0.f);
}

}  // namespace

int fc_tc_smem_bytes(int np) {
  return kFcStages * (2 * (kFcKC / 8) * 128 * 16 + 2 * (kFcKC / 8) * np * 16) + (2 * kFcStages + 1) * 8 + 8 + np * 4 + 128;
}

void fc_tc_dims(int S, int* kp, int* np) {
  *kp = (6 * S + kFcKC - 1) / kFcKC * kFcKC;
  *np = (S + 1 + 15) / 16 * 16;
}

int fc_tc_configure(ap_engine* e, NetState* n) {
  if (n->fc_np > 256 || fc_tc_smem_bytes(n->fc_np) > 227 * 1024) return ap_fail(e, AP_ERR_BAD_ARG, "fc_tc: board too large");
  AP_CUDA(e, cudaFuncSetAttribute(k_head_fc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, fc_tc_smem_bytes(n->fc_np)));
  return AP_OK;
}

int fc_tc_prep(ap_engine* e, NetState* n) {
  const HeadParams& h = n->head;
  k_prep_fc<<<256, 256, 0, e->stream>>>(n->master, h.fcp_w, h.fcp_b, h.fcv_w, h.fcv_b, n->S, n->fc_kp / 8, n->fc_np,
                                        n->fc_w, n->fc_bias);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}

// skip_finish: the caller consumes the partial logits itself (k_expand_backup, see fc_finish.cuh / net_fc_finish_args)
int fc_tc_launch(ap_engine* e, NetState* n, int nb, float* d_probs, float* d_values, const int* nb_dev, bool skip_finish) {
  FcParams p;
  p.a = n->fc_a;
  p.w = n->fc_w;
  p.bias = n->fc_bias;
  p.probs = d_probs;
  p.values = d_values;
  p.rows = n->fc_rows;
  p.kg = n->fc_kp / 8;
  p.np = n->fc_np;
  p.S = n->S;
  p.nb = nb;
  p.nb_dev = nb_dev;
  p.errflag = n->d_err;
  p.partial = n->fc_partial;
  p.ksplit = n->fc_ksplit;
  k_head_fc_tc<<<dim3((nb + 127) / 128, p.ksplit), kFcThreads, fc_tc_smem_bytes(n->fc_np), e->stream>>>(p);
  AP_LAUNCH_CHECK(e);
  if (p.ksplit > 1 && !skip_finish) {
    k_head_fc_finish<<<(nb + 3) / 4, 128, 0, e->stream>>>(p);
    AP_LAUNCH_CHECK(e);
  }
  return AP_OK;
}
