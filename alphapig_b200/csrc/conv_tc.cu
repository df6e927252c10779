// 3x3 convolution + folded BatchNorm + ReLU (+ residual) as an implicit GEMM on the sm_100a
// tensor cores: tcgen05.mma (fp16 x fp16 -> fp32 in TMEM), operands staged in shared memory by
// the TMA engine (cp.async.bulk + mbarrier), warp-specialised persistent kernel.
// Replaces the Convolution->BatchNorm->Activation triples of reference conv_act
// (policy_value_net_mxnet_simple.py:39-58) and the residual blocks (policy_value_net_mxnet.py:74-83).
//
// Layouts (DESIGN.md "net"):
//   activations  [C/8][Mpad][8] fp16; board b pixel (y,x) is row PAD + b*256 + y*16 + x; the
//                16th row/column of every board and the PAD rows are zero, so the 3x3 halo of any
//                real pixel is a plain row offset dy*16+dx into the same plane.
//   one tile     = one board = 256 rows = two UMMA M=128 accumulators (TMEM columns [0,N) and [N,2N)).
//   A operand    = a 290-row slab of KC<=64 channels ([KC/8][290][8]: K-major, no swizzle, 8x16B core
//                matrices; any start row is a legal descriptor base) loaded ONCE per K-chunk and reused
//                by all 9 taps through descriptor start offsets.
//   B operand    = per (K-chunk, tap) image [KC/8][Cout][8] prepared by net.cu.
#include "kernels.h"
#include "net.h"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSlabRows = NET_SLAB_ROWS;
constexpr int kSlabGroupBytes = kSlabRows * 16;  // one 8-channel group of the slab

struct ConvParams {
  const __half* in;
  __half* out;
  const __half* resid;
  const __half* wimg;
  const float* bias;  // folded BN shift; the BN scale is folded into the fp16 weights
  long long mpad;
  int cout, kc, nkc, relu;
  int n_tiles, W, H;
  int nb;         // B-tile ring depth
  int tmem_cols;  // power of two >= acc_stages*2*cout
  int acc_stages; // 2 when two tiles' accumulators fit in TMEM (cout <= 128): epilogue overlaps the next tile
  int* errflag;
};

__global__ void __launch_bounds__(kThreads, 1) k_conv3x3_tc(ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kg = p.kc >> 3;                         // 8-channel groups per K-chunk
  const uint32_t slab_bytes = kg * kSlabGroupBytes;  // multiple of 16
  const uint32_t slab_stride = (slab_bytes + 127u) & ~127u;
  const uint32_t btile_bytes = (uint32_t)p.kc * p.cout * 2;
  uint8_t* slab0 = smem;
  uint8_t* btile0 = smem + 2 * slab_stride;
  uint64_t* bars = (uint64_t*)(btile0 + (size_t)p.nb * btile_bytes);
  // barrier map: [0,2) slab_full, [2,4) slab_empty, [4,4+nb) b_full, [4+nb,4+2nb) b_empty, then tmem_full, tmem_empty
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = bars + 2;
  uint64_t* b_full = bars + 4;
  uint64_t* b_empty = bars + 4 + p.nb;
  uint64_t* tmem_full = bars + 4 + 2 * p.nb;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* s_bias = (float*)(tmem_slot + 4);     // [cout]

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&slab_full[i]), 1);
      mbar_init(smem_u32(&slab_empty[i]), 1);
    }
    for (int i = 0; i < p.nb; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full[i]), 1);
      mbar_init(smem_u32(&tmem_empty[i]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.cout; i += kThreads) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int sl = 0, slph = 0, bs = 0, bph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
      const long long row0 = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS - 17;
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = mbar_wait(smem_u32(&slab_empty[sl]), slph ^ 1, p.errflag);
        if (!ok) break;
        const uint32_t fb = smem_u32(&slab_full[sl]);
        mbar_expect_tx(fb, slab_bytes);
        for (int j = 0; j < kg; ++j)
          bulk_g2s(smem_u32(slab0 + sl * slab_stride + j * kSlabGroupBytes),
                   p.in + ((long long)(kc * kg + j) * p.mpad + row0) * 8, kSlabGroupBytes, fb);
        for (int tap = 0; tap < 9; ++tap) {
          ok = mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1, p.errflag);
          if (!ok) break;
          const uint32_t bb = smem_u32(&b_full[bs]);
          mbar_expect_tx(bb, btile_bytes);
          bulk_g2s(smem_u32(btile0 + (size_t)bs * btile_bytes),
                   p.wimg + (size_t)(kc * 9 + tap) * ((size_t)p.kc * p.cout), btile_bytes, bb);
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
        if (++sl == 2) { sl = 0; slph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.cout >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t b_lbo = (uint32_t)p.cout * 16;
    int sl = 0, slph = 0, bs = 0, bph = 0, as = 0, aph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
      ok = mbar_wait(smem_u32(&tmem_empty[as]), aph ^ 1, p.errflag);
      if (!ok) break;
      tc_fence_after();
      const uint32_t acc_base = tmem_base + (uint32_t)(as * 2 * p.cout);
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = mbar_wait(smem_u32(&slab_full[sl]), slph, p.errflag);
        if (!ok) break;
        const uint32_t sbase = smem_u32(slab0 + sl * slab_stride);
        for (int tap = 0; tap < 9; ++tap) {
          ok = mbar_wait(smem_u32(&b_full[bs]), bph, p.errflag);
          if (!ok) break;
          tc_fence_after();
          const int off = (tap / 3 - 1) * 16 + (tap % 3 - 1);
          const uint32_t bbase = smem_u32(btile0 + (size_t)bs * btile_bytes);
          for (int half = 0; half < 2; ++half) {
            const uint32_t arow = sbase + (uint32_t)(17 + off + half * 128) * 16;
            for (int j = 0; j < (p.kc >> 4); ++j) {
              const uint64_t ad = make_desc(arow + (uint32_t)(2 * j) * kSlabGroupBytes, kSlabGroupBytes, 128);
              const uint64_t bd = make_desc(bbase + (uint32_t)(2 * j) * b_lbo, b_lbo, 128);
              tc_mma_f16(acc_base + (uint32_t)(half * p.cout), ad, bd, idesc, (kc | tap | j) != 0);
            }
          }
          tc_commit(smem_u32(&b_empty[bs]));
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
        tc_commit(smem_u32(&slab_empty[sl]));
        if (++sl == 2) { sl = 0; slph ^= 1; }
      }
      tc_commit(smem_u32(&tmem_full[as]));
      if (++as == p.acc_stages) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> regs -> scale/shift (+resid) -> ReLU -> fp16 -> global =====
    const int q = warp & 3;
    int as = 0, aph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
      ok = mbar_wait(smem_u32(&tmem_full[as]), aph, p.errflag);
      ok = __all_sync(AP_FULL, ok);
      if (!ok) break;
      tc_fence_after();
      const uint32_t acc_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * p.cout);
      for (int half = 0; half < 2; ++half) {
        const int r = half * 128 + q * 32 + lane;
        const bool valid = ((r & 15) < p.W) && ((r >> 4) < p.H);
        const long long grow = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS + r;
        for (int c0 = 0; c0 < p.cout; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(acc_base + (uint32_t)(half * p.cout + c0), v);
          tmem_ld_wait();
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            const int c = c0 + gi * 8;
            const long long idx = ((long long)(c >> 3) * p.mpad + grow) * 8;
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c + 4);
            float f[8];
            f[0] = __uint_as_float(v[gi * 8 + 0]) + b0.x;
            f[1] = __uint_as_float(v[gi * 8 + 1]) + b0.y;
            f[2] = __uint_as_float(v[gi * 8 + 2]) + b0.z;
            f[3] = __uint_as_float(v[gi * 8 + 3]) + b0.w;
            f[4] = __uint_as_float(v[gi * 8 + 4]) + b1.x;
            f[5] = __uint_as_float(v[gi * 8 + 5]) + b1.y;
            f[6] = __uint_as_float(v[gi * 8 + 6]) + b1.z;
            f[7] = __uint_as_float(v[gi * 8 + 7]) + b1.w;
            if (p.resid) {
              uint4 rv = *reinterpret_cast<const uint4*>(p.resid + idx);
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float2 t = __half22float2(rh[k]);
                f[2 * k] += t.x;
                f[2 * k + 1] += t.y;
              }
            }
            uint4 ov;
            __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float a = f[2 * k], b = f[2 * k + 1];
              if (p.relu) {
                a = fmaxf(a, 0.f);
                b = fmaxf(b, 0.f);
              }
              if (!valid) a = b = 0.f;
              oh[k] = __floats2half2_rn(a, b);
            }
            *reinterpret_cast<uint4*>(p.out + idx) = ov;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[as]));
      if (++as == p.acc_stages) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

}  // namespace

static int smem_layout(const ConvLayer& L, int* nb_out) {
  const int kc = L.cin_pad < 64 ? L.cin_pad : 64;
  const int slab = (((kc >> 3) * kSlabGroupBytes) + 127) & ~127;
  const int btile = kc * L.cout * 2;
  const int budget = 220 * 1024;
  int nb = (budget - 2 * slab - 2048) / btile;
  if (nb > 9) nb = 9;
  if (nb < 2) nb = 2;
  *nb_out = nb;
  return 2 * slab + nb * btile + (4 + 2 * nb + 4) * 8 + 16 + L.cout * 4;
}

int conv_tc_smem_bytes(const ConvLayer& L, int* out_nb) { return smem_layout(L, out_nb); }

// per-device opt-in to > 48 KB dynamic shared memory (called from ap_net_load)
int conv_tc_configure(ap_engine* e) {
  AP_CUDA(e, cudaFuncSetAttribute(k_conv3x3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return AP_OK;
}

int conv_tc_launch(ap_engine* e, NetState* n, const ConvLayer& L, int n_boards) {
  ConvParams p;
  p.in = (L.in_buf < 0) ? n->feat : n->act[L.in_buf];
  p.out = n->act[L.out_buf];
  p.resid = (L.resid_buf >= 0) ? n->act[L.resid_buf] : nullptr;
  p.wimg = L.wimg;
  p.bias = L.shift;
  p.mpad = n->mpad;
  p.cout = L.cout;
  p.kc = L.cin_pad < 64 ? L.cin_pad : 64;
  p.nkc = L.cin_pad / p.kc;
  p.relu = L.relu;
  p.n_tiles = n_boards;
  p.W = n->W;
  p.H = n->H;
  p.acc_stages = (4 * L.cout <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < p.acc_stages * 2 * L.cout) cols <<= 1;
  p.tmem_cols = cols;
  p.errflag = n->d_err;
  int nb;
  int smem = smem_layout(L, &nb);
  p.nb = nb;
  int grid = n_boards < n->sm_count ? n_boards : n->sm_count;
  k_conv3x3_tc<<<grid, kThreads, smem, e->stream>>>(p);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}
