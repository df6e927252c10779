// Fused front of the 6-conv net: conv1 (9 -> 64) + BN + ReLU + conv2 (64 -> 64) + BN + ReLU in ONE kernel
// (policy_value_net_mxnet_simple.py:68-69, two conv_act triples).
//
// Why: the two 64-column layers are the furthest below the tensor roofline (N = 64 MMAs are bound by the 128 B/clk
// shared-memory operand read, conv1 on top of that by its 125 MB activation write).  A tile is a whole board, so
// conv1's output never needs a halo from another tile: its fp32 accumulator goes TMEM -> registers -> +shift, ReLU,
// fp16 -> straight into the shared-memory slab conv2's MMAs read as their A operand (same K-major no-swizzle
// core-matrix layout the TMA would have produced: [8 channel groups][290 rows][8]), and the 64-channel activation
// plane between the two layers (2 x 125 MB per 3840 boards) is neither written to nor read from HBM.
// The arithmetic is unchanged - same MMA order per tile, same fp16 rounding of conv1's output - so the results are
// bit-identical to the two separate launches (tests/test_gpu_net.py).
//
// One CTA per SM, persistent over boards.  Warp roles: 0 TMA producer (both weight tensors once: 18 + 72 KB resident;
// then the 9 KB feature slab of every board, 2 stages), 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue.
// TMEM: conv1 accumulator 2 stages x 128 columns, conv2 accumulator 2 stages x 128 columns (512).
// Issue order  c1(0) c1(1) c2(0) c1(2) c2(1) ...  and epilogue order  e1(0) e1(1) e2(0) e1(2) e2(1) ...  : while the
// epilogue warps turn conv1(t+1)'s accumulator into conv2's operand, the tensor pipe runs conv2(t).
#include "kernels.h"
#include "net.h"
#include "ptx.cuh"

namespace {

constexpr int kFThreads = 32 * 12;
constexpr int kRows = NET_SLAB_ROWS;       // 290
constexpr int kGroupBytes = kRows * 16;    // one 8-channel group of a slab
constexpr int C1 = 64, C2 = 64, KC1 = 16;
constexpr uint32_t W1_BYTES = 9u * KC1 * C1 * 2;   // 18432
constexpr uint32_t W2_BYTES = 9u * C1 * C2 * 2;    // 73728
constexpr uint32_t SLABF_BYTES = (KC1 / 8) * kGroupBytes;                 // 9280
constexpr uint32_t SLABF_STRIDE = (SLABF_BYTES + 127u) & ~127u;           // 9344
constexpr uint32_t SLAB2_BYTES = (C1 / 8) * kGroupBytes;                  // 37120 (multiple of 128)
constexpr uint32_t OFF_W1 = 0, OFF_W2 = OFF_W1 + W1_BYTES, OFF_SF = OFF_W2 + W2_BYTES, OFF_S2 = OFF_SF + 2 * SLABF_STRIDE;
constexpr uint32_t OFF_BAR = OFF_S2 + 2 * SLAB2_BYTES;
constexpr int kNumBars = 1 + 2 * 8;
constexpr uint32_t OFF_BIAS = OFF_BAR + kNumBars * 8 + 16;
constexpr uint32_t FRONT_SMEM = OFF_BIAS + (C1 + C2) * 4 + 128;

struct FrontParams {
  const __half* feat;   // [2][mpad][8] input planes (9 channels padded to 16)
  __half* out;          // [8][mpad][8] conv2 output planes
  const __half* w1;     // conv1 image [9][2][64][8]
  const __half* w2;     // conv2 image [9][8][64][8]
  const float* shift1;  // folded BN shifts
  const float* shift2;
  long long mpad;
  int n_tiles, W, H;
  const int* n_tiles_dev;
  int* errflag;
#ifdef AP_FRONT_TRACE
  long long* trace;  // development build: [role][tile][slot] clock64 stamps of CTA 0
#endif
};

#ifdef AP_FRONT_TRACE
// roles: 0 MMA issuer, 1 epilogue warp 4, 2 TMA producer; 64 tiles x 8 slots each
#define FTRACE(role, t, slot)                                                                   \
  do {                                                                                          \
    if (p.trace && blockIdx.x == 0 && (t) < 64 && (threadIdx.x & 31) == 0)                      \
      p.trace[(((role) * 64) + (t)) * 8 + (slot)] = clock64();                                  \
  } while (0)
#else
#define FTRACE(role, t, slot) \
  do {                        \
  } while (0)
#endif

__device__ __forceinline__ void tmem_ld_wait_regs32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// 8 accumulator columns -> +shift -> ReLU -> 8 fp16 (one 16-byte pixel record); zero for the pad column / row
__device__ __forceinline__ uint4 finish8(const uint32_t* v, const float* bias, bool valid) {
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
  const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __half2 h = __floats2half2_rn(__uint_as_float(v[2 * k]) + bias[2 * k], __uint_as_float(v[2 * k + 1]) + bias[2 * k + 1]);
    oh[k] = __hmax2(h, zero2);
  }
  if (!valid) o = make_uint4(0u, 0u, 0u, 0u);
  return o;
}

__global__ void __launch_bounds__(kFThreads, 1) k_front_tc(const __grid_constant__ FrontParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(AP_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t* w_full = bars;            // [1]
  uint64_t* f_full = bars + 1;        // [2] feature slab landed (TMA)
  uint64_t* f_empty = bars + 3;       // [2] conv1 MMAs done reading it
  uint64_t* a1_full = bars + 5;       // [2] conv1 accumulator complete
  uint64_t* a1_empty = bars + 7;      // [2] drained by the epilogue
  uint64_t* s2_full = bars + 9;       // [2] conv2 operand slab written by the epilogue
  uint64_t* s2_empty = bars + 11;     // [2] conv2 MMAs done reading it
  uint64_t* a2_full = bars + 13;      // [2]
  uint64_t* a2_empty = bars + 15;     // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + kNumBars);
  float* s_b1 = (float*)(smem + OFF_BIAS);
  float* s_b2 = s_b1 + C1;
  const int n_tiles = p.n_tiles_dev ? *p.n_tiles_dev : p.n_tiles;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(w_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&f_full[i]), 1);
      mbar_init(smem_u32(&f_empty[i]), 1);
      mbar_init(smem_u32(&a1_full[i]), 1);
      mbar_init(smem_u32(&a1_empty[i]), 256);
      mbar_init(smem_u32(&s2_full[i]), 256);
      mbar_init(smem_u32(&s2_empty[i]), 1);
      mbar_init(smem_u32(&a2_full[i]), 1);
      mbar_init(smem_u32(&a2_empty[i]), 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C1; i += kFThreads) {
    s_b1[i] = p.shift1[i];
    s_b2[i] = p.shift2[i];
  }
  // the halo rows of conv2's operand slabs (17 before and after the board) are zero for the whole kernel: the
  // epilogue only ever writes the 256 board rows
  for (int i = threadIdx.x; i < 2 * (C1 / 8) * 34; i += kFThreads) {
    const int buf = i / ((C1 / 8) * 34), r = i % ((C1 / 8) * 34);
    const int g = r / 34, h = r % 34;
    const int row = h < 17 ? h : 256 + h;  // 0..16 and 273..289
    *reinterpret_cast<uint4*>(smem + OFF_S2 + (size_t)buf * SLAB2_BYTES + (size_t)g * kGroupBytes + (size_t)row * 16) =
        make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ===== TMA producer =====
    if (my_tiles > 0 && elect_one()) {
      mbar_expect_tx(smem_u32(w_full), W1_BYTES + W2_BYTES);
      bulk_g2s(smem_u32(smem + OFF_W1), p.w1, W1_BYTES, smem_u32(w_full));
      bulk_g2s(smem_u32(smem + OFF_W2), p.w2, W2_BYTES, smem_u32(w_full));
    }
    __syncwarp();
    bool ok = true;
    for (int t = 0; t < my_tiles && ok; ++t) {
      const int s = t & 1, ph = (t >> 1) & 1;
      FTRACE(2, t, 0);
      ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&f_empty[s]), ph ^ 1, p.errflag));
      if (!ok) break;
      FTRACE(2, t, 1);
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const long long row0 = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS - 17;
      if (elect_one()) {
        const uint32_t fb = smem_u32(&f_full[s]);
        mbar_expect_tx(fb, SLABF_BYTES);
#pragma unroll
        for (int j = 0; j < KC1 / 8; ++j)
          bulk_g2s(smem_u32(smem + OFF_SF + s * SLABF_STRIDE + j * kGroupBytes), p.feat + ((long long)j * p.mpad + row0) * 8,
                   kGroupBytes, fb);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);  // M 128, N 64, f16 -> f32
    constexpr uint64_t DESC_HI = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    constexpr uint32_t A_LBO = (uint32_t)kGroupBytes >> 4;  // 290
    constexpr uint32_t B_LBO = 64;                          // 64 columns x 16 B between the two K halves
    bool ok = my_tiles == 0 || __all_sync(AP_FULL, mbar_wait(smem_u32(w_full), 0, p.errflag));
    tc_fence_after();
    const uint32_t w1_lo = (smem_u32(smem + OFF_W1) >> 4) | (B_LBO << 16);
    const uint32_t w2_lo = (smem_u32(smem + OFF_W2) >> 4) | (B_LBO << 16);
    auto conv1 = [&](int t) -> bool {
      const int s = t & 1, ph = (t >> 1) & 1;
      FTRACE(0, t, 0);
      if (!__all_sync(AP_FULL, mbar_wait(smem_u32(&a1_empty[s]), ph ^ 1, p.errflag))) return false;
      FTRACE(0, t, 1);
      if (!__all_sync(AP_FULL, mbar_wait(smem_u32(&f_full[s]), ph, p.errflag))) return false;
      FTRACE(0, t, 2);
      tc_fence_after();
      const uint32_t a_lo = (smem_u32(smem + OFF_SF + s * SLABF_STRIDE) >> 4) | (A_LBO << 16);
      const uint32_t acc = tmem_base + (uint32_t)(s * 128);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int off = 17 + (tap / 3 - 1) * 16 + (tap % 3 - 1);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (uint32_t)(off + half * 128));
            const uint64_t bd = DESC_HI | (uint64_t)(w1_lo + (uint32_t)(tap * (int)((KC1 * C1 * 2) >> 4)));
            tc_mma_f16(acc + (uint32_t)(half * 64), ad, bd, IDESC, tap != 0);
          }
        }
        tc_commit(smem_u32(&f_empty[s]));
        tc_commit(smem_u32(&a1_full[s]));
      }
      __syncwarp();
      FTRACE(0, t, 3);
      return true;
    };
    auto conv2 = [&](int t) -> bool {
      const int s = t & 1, ph = (t >> 1) & 1;
      FTRACE(0, t, 4);
      if (!__all_sync(AP_FULL, mbar_wait(smem_u32(&a2_empty[s]), ph ^ 1, p.errflag))) return false;
      FTRACE(0, t, 5);
      if (!__all_sync(AP_FULL, mbar_wait(smem_u32(&s2_full[s]), ph, p.errflag))) return false;
      FTRACE(0, t, 6);
      tc_fence_after();
      const uint32_t a_lo = (smem_u32(smem + OFF_S2 + s * SLAB2_BYTES) >> 4) | (A_LBO << 16);
      const uint32_t acc = tmem_base + 256u + (uint32_t)(s * 128);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int off = 17 + (tap / 3 - 1) * 16 + (tap % 3 - 1);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int j = 0; j < C1 / 16; ++j) {
              const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (uint32_t)(off + half * 128 + 2 * j * (int)A_LBO));
              const uint64_t bd = DESC_HI | (uint64_t)(w2_lo + (uint32_t)(tap * (int)((C1 * C2 * 2) >> 4) + 2 * j * (int)B_LBO));
              tc_mma_f16(acc + (uint32_t)(half * 64), ad, bd, IDESC, (tap | j) != 0);
            }
          }
        }
        tc_commit(smem_u32(&s2_empty[s]));
        tc_commit(smem_u32(&a2_full[s]));
      }
      __syncwarp();
      FTRACE(0, t, 7);
      return true;
    };
    if (ok && my_tiles > 0) ok = conv1(0);
    for (int t = 0; t < my_tiles && ok; ++t) {
      if (t + 1 < my_tiles) ok = conv1(t + 1);
      if (ok) ok = conv2(t);
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp = (TMEM lane quarter q, column half cg) =====
    const int q = warp & 3;
    const int cg = (warp - 4) >> 2;
    const int c0 = cg * 32;  // first of this warp's 32 columns
    bool ok = true;
    // e1: conv1 accumulator -> conv2 operand slab
    auto epi1 = [&](int t) -> bool {
      const int s = t & 1, ph = (t >> 1) & 1;
      if (warp == 4) FTRACE(1, t, 0);
      if (!__all_sync(AP_FULL, mbar_wait(smem_u32(&s2_empty[s]), ph ^ 1, p.errflag))) return false;
      if (warp == 4) FTRACE(1, t, 1);
      if (!__all_sync(AP_FULL, mbar_wait(smem_u32(&a1_full[s]), ph, p.errflag))) return false;
      if (warp == 4) FTRACE(1, t, 2);
      tc_fence_after();
      uint32_t v[2][32];
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128 + c0);
      tmem_ld32(acc, v[0]);
      tmem_ld32(acc + 64u, v[1]);
      tmem_ld_wait_regs32(v[0]);
      tmem_ld_wait_regs32(v[1]);
      tc_fence_before();
      mbar_arrive(smem_u32(&a1_empty[s]));  // both halves are in registers
      uint8_t* slab = smem + OFF_S2 + (size_t)s * SLAB2_BYTES;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = half * 128 + q * 32 + lane;
        const bool valid = ((r & 15) < p.W) && ((r >> 4) < p.H);
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const uint4 o = finish8(&v[half][gi * 8], s_b1 + c0 + gi * 8, valid);
          *reinterpret_cast<uint4*>(slab + (size_t)((c0 >> 3) + gi) * kGroupBytes + (size_t)(17 + r) * 16) = o;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA's reads
      mbar_arrive(smem_u32(&s2_full[s]));
      if (warp == 4) FTRACE(1, t, 3);
      return true;
    };
    // e2: conv2 accumulator -> global activation planes
    auto epi2 = [&](int t) -> bool {
      const int s = t & 1, ph = (t >> 1) & 1;
      if (warp == 4) FTRACE(1, t, 4);
      if (!__all_sync(AP_FULL, mbar_wait(smem_u32(&a2_full[s]), ph, p.errflag))) return false;
      if (warp == 4) FTRACE(1, t, 5);
      tc_fence_after();
      uint32_t v[2][32];
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(s * 128 + c0);
      tmem_ld32(acc, v[0]);
      tmem_ld32(acc + 64u, v[1]);
      tmem_ld_wait_regs32(v[0]);
      tmem_ld_wait_regs32(v[1]);
      tc_fence_before();
      mbar_arrive(smem_u32(&a2_empty[s]));
      if (warp == 4) FTRACE(1, t, 6);
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = half * 128 + q * 32 + lane;
        const bool valid = ((r & 15) < p.W) && ((r >> 4) < p.H);
        const long long grow = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS + r;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const uint4 o = finish8(&v[half][gi * 8], s_b2 + c0 + gi * 8, valid);
          *reinterpret_cast<uint4*>(p.out + ((long long)((c0 >> 3) + gi) * p.mpad + grow) * 8) = o;
        }
      }
      if (warp == 4) FTRACE(1, t, 7);
      return true;
    };
    if (my_tiles > 0) ok = epi1(0);
    for (int t = 0; t < my_tiles && ok; ++t) {
      if (t + 1 < my_tiles) ok = epi1(t + 1);
      if (ok) ok = epi2(t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

bool front_tc_supported(const NetState* n) {
  if (n->arch != AP_ARCH_SIMPLE || n->split || n->trunk.size() < 3) return false;
  const ConvLayer& a = n->trunk[0];
  const ConvLayer& b = n->trunk[1];
  return a.ksz == 3 && b.ksz == 3 && a.in_buf < 0 && a.cin_pad == KC1 && a.cout == C1 && b.cin_pad == C1 && b.cout == C2 &&
         a.relu && b.relu && a.resid_buf < 0 && b.resid_buf < 0 && b.in_buf == a.out_buf && !a.out_coff && !b.out_coff &&
         !a.in_coff && !b.in_coff;
}

int front_tc_configure(ap_engine* e) {
  AP_CUDA(e, cudaFuncSetAttribute(k_front_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FRONT_SMEM));
  return AP_OK;
}

// conv1 + conv2 of the 6-conv net on n_boards tiles (count optionally read on the device)
int front_tc_launch(ap_engine* e, NetState* n, int n_boards, const int* n_boards_dev) {
  const ConvLayer& a = n->trunk[0];
  const ConvLayer& b = n->trunk[1];
  FrontParams p;
  p.feat = n->feat;
  p.out = n->act[b.out_buf];
  p.w1 = a.wimg;
  p.w2 = b.wimg;
  p.shift1 = a.shift;
  p.shift2 = b.shift;
  p.mpad = n->mpad;
  p.n_tiles = n_boards;
  p.n_tiles_dev = n_boards_dev;
  p.W = n->W;
  p.H = n->H;
  p.errflag = n->d_err;
  const int grid = n_boards < n->sm_count ? n_boards : n->sm_count;
#ifdef AP_FRONT_TRACE
  static long long* d_trace = nullptr;
  if (!d_trace) cudaMalloc(&d_trace, 3 * 64 * 8 * 8);
  cudaMemsetAsync(d_trace, 0, 3 * 64 * 8 * 8, e->stream);
  p.trace = d_trace;
#endif
  k_front_tc<<<grid, kFThreads, FRONT_SMEM, e->stream>>>(p);
  AP_LAUNCH_CHECK(e);
#ifdef AP_FRONT_TRACE
  if (const char* path = getenv("AP_FRONT_TRACE_FILE")) {
    static long long h[3 * 64 * 8];
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(path, "w")) {
      long long t0 = h[0];
      for (int i = 0; i < 3 * 64 * 8; ++i)
        if (h[i] && h[i] < t0) t0 = h[i];
      for (int role = 0; role < 3; ++role)
        for (int t = 0; t < 64; ++t) {
          if (!h[(role * 64 + t) * 8] && !h[(role * 64 + t) * 8 + 1]) continue;
          fprintf(f, "%d %d", role, t);
          for (int sl = 0; sl < 8; ++sl) fprintf(f, " %lld", h[(role * 64 + t) * 8 + sl] ? h[(role * 64 + t) * 8 + sl] - t0 : -1ll);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
#endif
  return AP_OK;
}
