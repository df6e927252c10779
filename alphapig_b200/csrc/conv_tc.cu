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

namespace {

constexpr int kThreads = 256;
constexpr int kSlabRows = NET_SLAB_ROWS;
constexpr int kSlabGroupBytes = kSlabRows * 16;  // one 8-channel group of the slab

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box.  ~2 s at 2 GHz.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* errflag) {
  if (mbar_try(bar, parity)) return true;
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      atomicExch(errflag, 1);
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no swizzle: rows 16 B apart inside an 8-row core matrix, SBO between 8-row groups,
// LBO between the two 16-byte K halves of one K=16 MMA.  version=1 (Blackwell), layout_type=0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct ConvParams {
  const __half* in;
  __half* out;
  const __half* resid;
  const __half* wimg;
  const float* scale;
  const float* shift;
  long long mpad;
  int cout, kc, nkc, relu;
  int n_tiles, W, H;
  int nb;         // B-tile ring depth
  int tmem_cols;  // power of two >= 2*cout
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
  uint64_t* tmem_full = bars + 4 + 2 * p.nb;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&slab_full[i]), 1);
      mbar_init(smem_u32(&slab_empty[i]), 1);
    }
    for (int i = 0; i < p.nb; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    mbar_init(smem_u32(tmem_empty), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
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
    int sl = 0, slph = 0, bs = 0, bph = 0, tph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
      ok = mbar_wait(smem_u32(tmem_empty), tph ^ 1, p.errflag);
      if (!ok) break;
      tc_fence_after();
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
              tc_mma_f16(tmem_base + (uint32_t)(half * p.cout), ad, bd, idesc, (kc | tap | j) != 0);
            }
          }
          tc_commit(smem_u32(&b_empty[bs]));
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
        tc_commit(smem_u32(&slab_empty[sl]));
        if (++sl == 2) { sl = 0; slph ^= 1; }
      }
      tc_commit(smem_u32(tmem_full));
      tph ^= 1;
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> regs -> scale/shift (+resid) -> ReLU -> fp16 -> global =====
    const int q = warp & 3;
    int tph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.n_tiles && ok; tile += gridDim.x) {
      ok = mbar_wait(smem_u32(tmem_full), tph, p.errflag);
      ok = __all_sync(AP_FULL, ok);
      if (!ok) break;
      tc_fence_after();
      for (int half = 0; half < 2; ++half) {
        const int r = half * 128 + q * 32 + lane;
        const bool valid = ((r & 15) < p.W) && ((r >> 4) < p.H);
        const long long grow = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS + r;
        for (int c0 = 0; c0 < p.cout; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * p.cout + c0), v);
          tmem_ld_wait();
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            const int c = c0 + gi * 8;
            const long long idx = ((long long)(c >> 3) * p.mpad + grow) * 8;
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaf(__uint_as_float(v[gi * 8 + k]), p.scale[c + k], p.shift[c + k]);
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
      mbar_arrive(smem_u32(tmem_empty));
      tph ^= 1;
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
  int nb = (budget - 2 * slab - 512) / btile;
  if (nb > 9) nb = 9;
  if (nb < 2) nb = 2;
  *nb_out = nb;
  return 2 * slab + nb * btile + (4 + 2 * nb + 2) * 8 + 16;
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
  p.scale = L.scale;
  p.shift = L.shift;
  p.mpad = n->mpad;
  p.cout = L.cout;
  p.kc = L.cin_pad < 64 ? L.cin_pad : 64;
  p.nkc = L.cin_pad / p.kc;
  p.relu = L.relu;
  p.n_tiles = n_boards;
  p.W = n->W;
  p.H = n->H;
  int cols = 32;
  while (cols < 2 * L.cout) cols <<= 1;
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
