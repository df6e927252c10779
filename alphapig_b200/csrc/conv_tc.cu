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
//   B operand    = per (K-chunk, tap) image [KC/8][Cout][8] prepared by net.cu; TPS consecutive taps
//                travel as one bulk copy / one ring stage.  When the whole layer fits in the ring
//                (conv1..conv3, the residual stem) it is loaded once per CTA and stays resident.
//
// Issue-rate notes (tools/mma_probe.cu, measured on B200): an M=128 MMA executes in 128/64/48 cycles
// for N=256/128/64 (N=64 is bound by the 128 B/cycle shared-memory operand read), and rebuilding both
// descriptors per instruction costs ~64-70 issue cycles - more than the N<=128 instruction itself.
// The issuer below therefore keeps COUT/KC compile-time, unrolls all taps and only adds constants to
// pre-built descriptor words.
#include "kernels.h"
#include "net.h"
#include "ptx.cuh"

namespace {

constexpr int kCtrlWarps = 4;                       // TMA producer, MMA issuer, TMEM allocator, spare
constexpr int kEpiWarps = 8;                        // two per TMEM lane quarter, each owning half of the columns
constexpr int kThreads = 32 * (kCtrlWarps + kEpiWarps);
constexpr int kSlabRows = NET_SLAB_ROWS;
constexpr int kSlabGroupBytes = kSlabRows * 16;     // one 8-channel group of the slab
constexpr int kMaxSlabs = 4, kMaxStages = 9;

struct ConvParams {
  const __half* in;
  __half* out;
  const __half* resid;
  const __half* wimg;
  // SPLIT instantiations (near-fp32 "hi + lo" fp16 pairs, three tensor-core products per K step):
  const __half* in_lo;
  __half* out_lo;
  const __half* resid_lo;
  const __half* wimg_lo;
  const float* bias;  // folded BN shift; the BN scale is folded into the fp16 weights
  long long mpad;
  int in_goff;      // first 8-channel group of the input planes this layer reads (single-CTA kernel only)
  int out_coff;     // the COUT output channels land at channels [out_coff, out_coff + cout_store) of the output planes
  int cout_store;   // channels actually stored (a multiple of 8; the rest are zero-weight padding)
  int nkc, relu;
  int n_tiles, W, H;
  const int* n_tiles_dev;  // when non-null the tile count is read from device memory (compacted leaf batches)
  int ns;                  // slab ring depth
  int nb;                  // B ring depth (stages of TPS taps)
  int* errflag;
  // HEAD instantiations only: the two 1x1 head convs (+folded BN+ReLU) run in the epilogue and their outputs
  // go straight into the split-fp16 A operand of the FC GEMM (heads_tc.cu); the trunk output is never stored.
  // The 1x1 weights travel as a second kernel parameter (HeadArg): they sit in the constant bank and feed the
  // FFMAs as immediate c[0x0][..] operands - no shared-memory traffic next to the tensor-core operand reads
  __half* ha;        // [2 (hi,lo)][ha_kg][ha_rows][8]
  long long ha_rows; // board rows of the A operand (multiple of 128)
  int ha_kg;         // K groups (of 8) of the A operand
#ifdef AP_CONV_TRACE
  long long* trace;
#endif
};

#ifdef AP_CONV_TRACE
// development build only (tools/conv_bench.cu): per-role clock64 stamps of the first tiles of CTAs 0 and 1
static long long* g_conv_trace = nullptr;
#define CTRACE(role, it, slot)                                                                      \
  do {                                                                                              \
    if (p.trace && blockIdx.x < 2 && (it) < 64 && (threadIdx.x & 31) == 0)                          \
      p.trace[((blockIdx.x * 4 + (role)) * 64 + (it)) * 8 + (slot)] = clock64();                    \
  } while (0)
#else
#define CTRACE(role, it, slot) \
  do {                         \
  } while (0)
#endif

template <int COUT>
struct ConvCfg {
  static constexpr int TPS = (COUT == 256) ? 1 : 3;          // taps per B stage
  static constexpr int ACC_STAGES = (COUT <= 128) ? 2 : 1;   // accumulator double buffering when TMEM allows
  static constexpr int TMEM_COLS = (ACC_STAGES * 2 * COUT <= 256) ? 256 : 512;
};

// wait for every outstanding tcgen05.ld and tie the destination registers to the wait
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}


// ---- epilogue arithmetic shared by both kernels ------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}

// second kernel parameter of the HEAD instantiations (empty otherwise)
template <bool HEAD>
struct HeadArg {};
template <>
struct HeadArg<true> {
  NetHeadW h;
};

// residual operand of one chunk, fetched one chunk ahead of its use (the first one before the accumulator is even
// complete) so that its HBM latency hides under the TMEM drain instead of serialising the epilogue
template <int NG, bool SPLIT>
struct ResidRegs {
  uint4 hi[NG];
  uint4 lo[SPLIT ? NG : 1];
};
template <bool RESID, int NG, bool SPLIT>
__device__ __forceinline__ void resid_load(const ConvParams& p, int c0, long long grow, ResidRegs<NG, SPLIT>& r) {
  if constexpr (RESID) {
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      const long long idx = ((long long)((c0 + gi * 8) >> 3) * p.mpad + grow) * 8;
      r.hi[gi] = __ldg(reinterpret_cast<const uint4*>(p.resid + idx));
      if constexpr (SPLIT) r.lo[gi] = __ldg(reinterpret_cast<const uint4*>(p.resid_lo + idx));
    }
  }
}

// one chunk (NG groups of 8 columns) of one accumulator row: +shift (+residual) -> ReLU, then either fp16 store of the
// activation row (trunk layers) or 6 running dot products with the 1x1 head weights (HEAD; c0 is a compile-time
// constant after inlining + unrolling, so every weight is an immediate constant-bank operand)
template <int COUT, bool RESID, bool HEAD, int NG, bool SPLIT = false>
__device__ __forceinline__ void epi_chunk(const ConvParams& p, const HeadArg<HEAD>& hw, const uint32_t (&v)[8 * NG], int c0,
                                          long long grow, bool valid, uint32_t s_bias, float (&hacc)[6],
                                          const ResidRegs<NG, SPLIT>& rr) {
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    const int c = c0 + gi * 8;
    const long long idx = ((long long)(c >> 3) * p.mpad + grow) * 8;                    // residual operand (same channel)
    const long long oidx = ((long long)((c + p.out_coff) >> 3) * p.mpad + grow) * 8;     // output
    const float4 b0 = lds128(s_bias + c * 4);
    const float4 b1 = lds128(s_bias + c * 4 + 16);
    float f[8];
    f[0] = __uint_as_float(v[gi * 8 + 0]) + b0.x;
    f[1] = __uint_as_float(v[gi * 8 + 1]) + b0.y;
    f[2] = __uint_as_float(v[gi * 8 + 2]) + b0.z;
    f[3] = __uint_as_float(v[gi * 8 + 3]) + b0.w;
    f[4] = __uint_as_float(v[gi * 8 + 4]) + b1.x;
    f[5] = __uint_as_float(v[gi * 8 + 5]) + b1.y;
    f[6] = __uint_as_float(v[gi * 8 + 6]) + b1.z;
    f[7] = __uint_as_float(v[gi * 8 + 7]) + b1.w;
    if (RESID) {
      const __half2* rh = reinterpret_cast<const __half2*>(&rr.hi[gi]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 t = __half22float2(rh[k]);
        f[2 * k] += t.x;
        f[2 * k + 1] += t.y;
      }
      if constexpr (SPLIT) {
        const __half2* rlh = reinterpret_cast<const __half2*>(&rr.lo[gi]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 t = __half22float2(rlh[k]);
          f[2 * k] += t.x;
          f[2 * k + 1] += t.y;
        }
      }
    }
    if constexpr (HEAD) {
      if (p.relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
      }
#pragma unroll
      for (int o = 0; o < 6; ++o) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hacc[o] = fmaf(f[k], hw.h.w[o * COUT + c + k], hacc[o]);
      }
    } else if (c >= p.cout_store) {
      // zero-weight padding channels of a narrower layer: nothing to store
    } else if constexpr (SPLIT) {
      uint4 ov, ol;
      __half2* oh = reinterpret_cast<__half2*>(&ov);
      __half2* olh = reinterpret_cast<__half2*>(&ol);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a = f[2 * k], b = f[2 * k + 1];
        if (p.relu) {
          a = fmaxf(a, 0.f);
          b = fmaxf(b, 0.f);
        }
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        oh[k] = h;
        olh[k] = __floats2half2_rn(a - hf.x, b - hf.y);
      }
      if (!valid) {
        ov = make_uint4(0u, 0u, 0u, 0u);
        ol = ov;
      }
      *reinterpret_cast<uint4*>(p.out + oidx) = ov;
      *reinterpret_cast<uint4*>(p.out_lo + oidx) = ol;
    } else {
      uint4 ov;
      __half2* oh = reinterpret_cast<__half2*>(&ov);
      const __half2 zero2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __half2 h = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
        if (p.relu) h = __hmax2(h, zero2);
        oh[k] = h;
      }
      if (!valid) ov = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(p.out + oidx) = ov;
    }
  }
}

// HEAD: the NCG warps that share a TMEM lane quarter (column groups 0..NCG-1) combine their partial dot
// products through shared memory; the group-0 warp finishes (shift, ReLU) and writes hi/lo fp16 into the
// FC operand.  NR = accumulator rows per thread (2 in the single-CTA kernel, 1 in the pair kernel).
template <int NR, int NCG>
__device__ __forceinline__ void head_finish(const ConvParams& p, const HeadArg<true>& hw, float (&hacc)[NR][6], uint32_t s_hx,
                                            int bar_id, int q, int lane, int cg, int tile, const int (&rows)[NR],
                                            bool live = true) {
  // live == false (a pair tile's second board past the end of the batch): no data, but the same two barrier
  // instructions - every warp of a named barrier always meets its peer at the same bar.sync
  if (live && cg != 0) {
    const uint32_t slot = s_hx + (uint32_t)(((cg - 1) * 128 + q * 32 + lane) * (NR * 6) * 4);
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int o = 0; o < 6; ++o)
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot + (uint32_t)((r * 6 + o) * 4)), "f"(hacc[r][o]) : "memory");
  }
  __syncwarp();  // bar.sync is the .aligned form: the warp must arrive converged (the predicated blocks above may have split it)
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NCG) : "memory");
  if (live && cg == 0) {
    const int S = p.W * p.H;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int x = rows[r] & 15, y = rows[r] >> 4;
      if (x < p.W && y < p.H) {
        const int pix = y * p.W + x;
#pragma unroll
        for (int o = 0; o < 6; ++o) {
          float a = hacc[r][o];
#pragma unroll
          for (int g = 1; g < NCG; ++g) {
            float t;
            asm volatile("ld.shared.f32 %0, [%1];"
                         : "=f"(t)
                         : "r"(s_hx + (uint32_t)((((g - 1) * 128 + q * 32 + lane) * (NR * 6) + r * 6 + o) * 4))
                         : "memory");
            a += t;
          }
          const float h = fmaxf(a + hw.h.b[o], 0.f);
          const __half hi = __float2half_rn(h);
          const __half lo = __float2half_rn(h - __half2float(hi));
          const int k = o * S + pix;
          const long long at = ((long long)(k >> 3) * p.ha_rows + tile) * 8 + (k & 7);
          p.ha[at] = hi;
          p.ha[(long long)p.ha_kg * p.ha_rows * 8 + at] = lo;
        }
      }
    }
  }
  __syncwarp();  // bar.sync is the .aligned form: the warp must arrive converged (the predicated blocks above may have split it)
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NCG) : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_regs16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// HEAD instantiations run 16 epilogue warps (four per TMEM lane quarter, 16-column chunks): the accumulator of
// a 256-channel layer cannot be double buffered (512 TMEM columns), so the extra dot-product work of the fused
// head convs must drain the accumulator as fast as the plain epilogue does
template <bool HEAD>
struct EpiCfg {
  static constexpr int WARPS = HEAD ? 16 : kEpiWarps;
  static constexpr int THREADS = 32 * (kCtrlWarps + WARPS);
};

// TAPS = 9: 3x3 conv; TAPS = 1: 1x1 conv (only the centre tap of the same slab)
// SPLITM: 0 = fp16 operands; 1 = hi + lo pairs for activations AND weights (three products per K step);
//         2 = hi + lo activations x fp16 weights rounded by error diffusion along K (two products, net.cu)
template <int COUT, int KC, bool RESID, bool HEAD, int SPLITM = 0, int TAPS = 9>
__global__ void __launch_bounds__(EpiCfg<HEAD>::THREADS, 1)
k_conv3x3_tc(const __grid_constant__ ConvParams p, const __grid_constant__ HeadArg<HEAD> hw) {
  using Cfg = ConvCfg<COUT>;
  constexpr bool SPLIT = SPLITM != 0;    // activations travel as hi + lo
  constexpr bool SPLITW = SPLITM == 1;   // weights too
  constexpr int EPI = EpiCfg<HEAD>::WARPS;
  constexpr int NTHR = EpiCfg<HEAD>::THREADS;
  constexpr int TPS = TAPS == 1 ? 1 : Cfg::TPS;
  constexpr int STAGES_PER_KC = TAPS / TPS;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr int KG = KC / 8;                              // 8-channel groups per K-chunk
  constexpr uint32_t SLAB_HALF = KG * kSlabGroupBytes;                 // one of hi / lo; multiple of 16
  constexpr uint32_t SLAB_BYTES = (SPLIT ? 2 : 1) * SLAB_HALF;         // SPLIT: [hi groups][lo groups]
  constexpr uint32_t SLAB_STRIDE = (SLAB_BYTES + 127u) & ~127u;
  constexpr uint32_t TAP_BYTES = (uint32_t)KC * COUT * 2;
  constexpr uint32_t STAGE_HALF = TPS * TAP_BYTES;
  constexpr uint32_t STAGE_BYTES = (SPLITW ? 2 : 1) * STAGE_HALF;      // SPLITW: [hi taps][lo taps]

  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(AP_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint8_t* slab0 = smem;
  uint8_t* bstage0 = smem + (size_t)p.ns * SLAB_STRIDE;
  uint64_t* bars = (uint64_t*)(bstage0 + (size_t)p.nb * STAGE_BYTES);
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = slab_full + kMaxSlabs;
  uint64_t* b_full = slab_empty + kMaxSlabs;
  uint64_t* b_empty = b_full + kMaxStages;
  uint64_t* tmem_full = b_empty + kMaxStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* s_bias = (float*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);  // [COUT], float4 reads
  float* s_hx = s_bias + COUT;          // HEAD: [NCG-1][128][NR*6] partial exchange between the warps of a lane quarter

  const int n_tiles = p.n_tiles_dev ? *p.n_tiles_dev : p.n_tiles;
  // the whole layer's weights fit in the ring: load them once, never release
  const bool resident = p.nkc * STAGES_PER_KC <= p.nb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxSlabs; ++i) {
      mbar_init(smem_u32(&slab_full[i]), 1);
      mbar_init(smem_u32(&slab_empty[i]), 1);
    }
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full[i]), 1);
      mbar_init(smem_u32(&tmem_empty[i]), 32 * EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < COUT; i += NTHR) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (whole warp loops, one elected lane issues) =====
    int sl = 0, slph = 0, bs = 0, bph = 0, titer = 0;
    bool ok = true, first = true;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
      const int it = titer++;
      const long long row0 = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS - 17;
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        if (kc == 0) CTRACE(0, it, 0);
        ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&slab_empty[sl]), slph ^ 1, p.errflag));
        if (!ok) break;
        if (kc == 0) CTRACE(0, it, 1);
        const uint32_t fb = smem_u32(&slab_full[sl]);
        if (elect_one()) {
          mbar_expect_tx(fb, SLAB_BYTES);
#pragma unroll
          for (int j = 0; j < KG; ++j)
            bulk_g2s(smem_u32(slab0 + sl * SLAB_STRIDE + j * kSlabGroupBytes),
                     p.in + ((long long)(p.in_goff + kc * KG + j) * p.mpad + row0) * 8, kSlabGroupBytes, fb);
          if constexpr (SPLIT) {
#pragma unroll
            for (int j = 0; j < KG; ++j)
              bulk_g2s(smem_u32(slab0 + sl * SLAB_STRIDE + SLAB_HALF + j * kSlabGroupBytes),
                       p.in_lo + ((long long)(p.in_goff + kc * KG + j) * p.mpad + row0) * 8, kSlabGroupBytes, fb);
          }
        }
        __syncwarp();
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
        if (resident && !first) continue;
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          if (!resident) {
            ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1, p.errflag));
            if (!ok) break;
          }
          const uint32_t bb = smem_u32(&b_full[bs]);
          if (elect_one()) {
            mbar_expect_tx(bb, STAGE_BYTES);
            bulk_g2s(smem_u32(bstage0 + (size_t)bs * STAGE_BYTES),
                     p.wimg + (size_t)(kc * TAPS + ts * TPS) * ((size_t)KC * COUT), STAGE_HALF, bb);
            if constexpr (SPLITW)
              bulk_g2s(smem_u32(bstage0 + (size_t)bs * STAGE_BYTES + STAGE_HALF),
                       p.wimg_lo + (size_t)(kc * TAPS + ts * TPS) * ((size_t)KC * COUT), STAGE_HALF, bb);
          }
          __syncwarp();
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
      }
      first = false;
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((128u >> 4) << 24);
    // descriptor words: lo = addr>>4 | LBO>>4 << 16, hi = SBO>>4 | version 1 (bit 46); advancing a
    // K-major no-swizzle operand by rows / K-groups only adds to the 14-bit address field
    constexpr uint64_t DESC_HI = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    constexpr uint32_t A_LBO = (uint32_t)kSlabGroupBytes >> 4;   // 290
    constexpr uint32_t B_LBO = (uint32_t)COUT;                   // COUT*16 >> 4
    int sl = 0, slph = 0, bs = 0, bph = 0, as = 0, aph = 0, titer = 0;
    bool ok = true, first = true;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
      const int it = titer++;
      CTRACE(1, it, 0);
      ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&tmem_empty[as]), aph ^ 1, p.errflag));
      if (!ok) break;
      CTRACE(1, it, 1);
      tc_fence_after();
      const uint32_t acc_base = tmem_base + (uint32_t)(as * 2 * COUT);
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&slab_full[sl]), slph, p.errflag));
        if (!ok) break;
        if (kc == 0) CTRACE(1, it, 2);
        const uint32_t a_lo = (smem_u32(slab0 + sl * SLAB_STRIDE) >> 4) | (A_LBO << 16);
#pragma unroll
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          const int stage = resident ? kc * STAGES_PER_KC + ts : bs;
          if (!resident || first) {
            ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&b_full[stage]), resident ? 0 : bph, p.errflag));
            if (!ok) break;
          }
          tc_fence_after();
          const uint32_t b_lo = (smem_u32(bstage0 + (size_t)stage * STAGE_BYTES) >> 4) | (B_LBO << 16);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
              const int tap = ts * TPS + t;
              const int off = TAPS == 1 ? 17 : 17 + (tap / 3 - 1) * 16 + (tap % 3 - 1);
#pragma unroll
              for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < KC / 16; ++j) {
                  const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (uint32_t)(off + half * 128 + 2 * j * (int)A_LBO));
                  const uint64_t bd = DESC_HI | (uint64_t)(b_lo + (uint32_t)(t * (int)(TAP_BYTES >> 4) + 2 * j * (int)B_LBO));
                  tc_mma_f16(acc_base + (uint32_t)(half * COUT), ad, bd, IDESC, (kc | tap | j) != 0);
                  if constexpr (SPLIT)   // + a_lo * b_hi
                    tc_mma_f16(acc_base + (uint32_t)(half * COUT), ad + (SLAB_HALF >> 4), bd, IDESC, 1);
                  if constexpr (SPLITW)  // + a_hi * b_lo (lo * lo is below fp32 round-off)
                    tc_mma_f16(acc_base + (uint32_t)(half * COUT), ad, bd + (STAGE_HALF >> 4), IDESC, 1);
                }
              }
            }
            if (!resident) tc_commit(smem_u32(&b_empty[bs]));
            if (ts == STAGES_PER_KC - 1) {
              tc_commit(smem_u32(&slab_empty[sl]));
              if (kc == p.nkc - 1) tc_commit(smem_u32(&tmem_full[as]));
            }
          }
          __syncwarp();
          if (!resident && ++bs == p.nb) { bs = 0; bph ^= 1; }
        }
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
      }
      CTRACE(1, it, 3);
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
      first = false;
    }
  } else if (warp >= kCtrlWarps) {
    // ===== epilogue: TMEM -> regs -> +shift (+resid) -> ReLU -> fp16 -> global (or the fused head convs) =====
    // warp w reads TMEM lanes 32*(w&3)..+31 (hardware restriction) and column group (w-4)>>2 of EPI/4
    constexpr int NCG = EPI / 4;                 // warps per lane quarter
    constexpr int COLS_PER_WARP = COUT / NCG;
    constexpr int CHUNKS = COLS_PER_WARP / 32;   // 32-column chunks per accumulator half (plain epilogue)
    constexpr int NCH = 2 * CHUNKS;              // chunks per tile and warp
    const int q = warp & 3;
    const int cg = (warp - kCtrlWarps) >> 2;
    const int cbase = cg * COLS_PER_WARP;
    int as = 0, aph = 0, titer = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
      const int it = titer++;
      if (warp == kCtrlWarps) CTRACE(2, it, 0);
      ok = mbar_wait(smem_u32(&tmem_full[as]), aph, p.errflag);
      ok = __all_sync(AP_FULL, ok);
      if (!ok) break;
      if (warp == kCtrlWarps) CTRACE(2, it, 1);
      tc_fence_after();
      const uint32_t acc_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * COUT);
      const long long grow0 = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS + q * 32 + lane;
      const uint32_t sb = smem_u32(s_bias), sx = smem_u32(s_hx), eb = smem_u32(&tmem_empty[as]);
      if constexpr (HEAD) {
        // four warps per lane quarter: warp (hh, cc) owns accumulator half hh (rows 128*hh..) and column half cc, so
        // every head weight is used exactly once per warp and is an immediate constant-bank operand
        static_assert(NCG == 4, "HEAD epilogue: four warps per lane quarter");
        const int hh = cg >> 1, cc = cg & 1;
        const int r = hh * 128 + q * 32 + lane;
        float hacc[1][6];
#pragma unroll
        for (int o = 0; o < 6; ++o) hacc[0][o] = 0.f;
        auto body = [&](const int cb) {
          constexpr int NCHH = COUT / 2 / 16;
          const uint32_t a0 = acc_base + (uint32_t)(hh * COUT + cb);
          ResidRegs<2, SPLIT> rrh[2];
          resid_load<RESID, 2, SPLIT>(p, cb, grow0 + hh * 128, rrh[0]);
          uint32_t v[2][16];
          tmem_ld16_nowait(a0, v[0]);
#pragma unroll
          for (int i = 0; i < NCHH; ++i) {
            tmem_ld_wait_regs16(v[i & 1]);
            if (i + 1 < NCHH) {
              tmem_ld16_nowait(a0 + (uint32_t)((i + 1) * 16), v[(i + 1) & 1]);
              resid_load<RESID, 2, SPLIT>(p, cb + (i + 1) * 16, grow0 + hh * 128, rrh[(i + 1) & 1]);
            } else {
              tc_fence_before();
              mbar_arrive(eb);
              if (warp == kCtrlWarps) CTRACE(2, it, 2);
            }
            epi_chunk<COUT, RESID, true, 2, SPLIT>(p, hw, v[i & 1], cb + i * 16, grow0 + hh * 128, true, sb, hacc[0], rrh[i & 1]);
          }
        };
        // the column half is a compile-time constant inside body (head weights as immediate constant-bank operands);
        // the two warps of a named barrier meet again HERE, at one bar.sync instruction, not inside the two copies
        if (cc == 0) body(0); else body(COUT / 2);
        const int rows[1] = {r};
        head_finish<1, 2>(p, hw, hacc, sx + (uint32_t)(hh * 128 * 6 * 4), 1 + q * 2 + hh, q, lane, cc, tile, rows);
      } else {
        uint32_t v[2][32];
        float hacc[6];
        ResidRegs<4, SPLIT> rr[2];
        resid_load<RESID, 4, SPLIT>(p, cbase, grow0, rr[0]);
        tmem_ld32(acc_base + cbase, v[0]);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int half = i / CHUNKS, c0 = cbase + (i % CHUNKS) * 32;
          tmem_ld_wait_regs(v[i & 1]);
          if (i + 1 < NCH) {
            tmem_ld32(acc_base + (uint32_t)(cbase + ((i + 1) / CHUNKS) * COUT + ((i + 1) % CHUNKS) * 32), v[(i + 1) & 1]);
            resid_load<RESID, 4, SPLIT>(p, cbase + ((i + 1) % CHUNKS) * 32, grow0 + ((i + 1) / CHUNKS) * 128, rr[(i + 1) & 1]);
          } else {
            // the accumulator stage is fully in registers: hand it back before the arithmetic and the stores
            tc_fence_before();
            mbar_arrive(eb);
            if (warp == kCtrlWarps) CTRACE(2, it, 2);
          }
          const int r = half * 128 + q * 32 + lane;
          const bool valid = ((r & 15) < p.W) && ((r >> 4) < p.H);
          epi_chunk<COUT, RESID, false, 4, SPLIT>(p, hw, v[i & 1], c0, grow0 + half * 128, valid, sb, hacc, rr[i & 1]);
        }
      }
      if (warp == kCtrlWarps) CTRACE(2, it, 3);
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair version (cta_group::2): one board per cluster of two CTAs, UMMA M=256.  CTA r owns board
// rows [128r, 128r+128): it stages a 162-row slab of them (A) and HALF of the weight tile (B columns
// [r*COUT/2, (r+1)*COUT/2)); the tensor cores of both SMs read both halves.  Per SM this halves the
// shared-memory operand traffic and the TMEM footprint, so the accumulator is double buffered for
// every COUT (the epilogue of tile i hides under the MMAs of tile i+1) and conv4 / the residual
// blocks keep their whole weight tensor resident in shared memory.
//   barriers (same offsets in both CTAs):
//     slab_full / b_full    local TMA completion (tx bytes)
//     slab_ready / b_ready  used in the leader only, count 2: one relay arrive per CTA after its
//                           local *_full completed (plain bulk copies cannot signal a peer barrier)
//     slab_empty / b_empty / tmem_full   tcgen05.commit multicast to both CTAs
//     tmem_empty            leader only, count 2*kEpiWarps: one arrive per epilogue warp of the pair
// ---------------------------------------------------------------------------------------------
constexpr int kSlab2Rows = 128 + 2 * 17;
constexpr int kSlab2GroupBytes = kSlab2Rows * 16;
constexpr int kTps2 = 3;  // taps per B stage

template <int COUT, int KC, bool RESID, bool HEAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_conv3x3_tc2(const __grid_constant__ ConvParams p, const __grid_constant__ HeadArg<HEAD> hw) {
  constexpr int TPS = kTps2;
  constexpr int STAGES_PER_KC = 9 / TPS;
  constexpr int ACC_STAGES = 2;
  constexpr int TMEM_COLS = (2 * COUT < 32) ? 32 : 2 * COUT;   // 128 / 256 / 512
  constexpr int KG = KC / 8;
  constexpr int NH = COUT / 2;                                 // B columns held by one CTA
  constexpr uint32_t SLAB_BYTES = KG * kSlab2GroupBytes;
  constexpr uint32_t SLAB_STRIDE = (SLAB_BYTES + 127u) & ~127u;
  constexpr uint32_t TAP_BYTES = (uint32_t)KC * NH * 2;
  constexpr uint32_t STAGE_BYTES = TPS * TAP_BYTES;

  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(AP_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* slab0 = smem;
  uint8_t* bstage0 = smem + (size_t)p.ns * SLAB_STRIDE;
  uint64_t* bars = (uint64_t*)(bstage0 + (size_t)p.nb * STAGE_BYTES);
  uint64_t* slab_full = bars;
  uint64_t* slab_ready = slab_full + kMaxSlabs;
  uint64_t* slab_empty = slab_ready + kMaxSlabs;
  uint64_t* b_full = slab_empty + kMaxSlabs;
  uint64_t* b_ready = b_full + kMaxStages;
  uint64_t* b_empty = b_ready + kMaxStages;
  uint64_t* tmem_full = b_empty + kMaxStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* s_bias = (float*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);  // [COUT], float4 reads
  float* s_hx = s_bias + COUT;          // HEAD: [NCG-1][128][NR*6] partial exchange between the warps of a lane quarter

  const int n_tiles = p.n_tiles_dev ? *p.n_tiles_dev : p.n_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const bool resident = p.nkc * STAGES_PER_KC <= p.nb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxSlabs; ++i) {
      mbar_init(smem_u32(&slab_full[i]), 1);
      mbar_init(smem_u32(&slab_ready[i]), 2);
      mbar_init(smem_u32(&slab_empty[i]), 1);
    }
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_ready[i]), 2);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full[i]), 1);
      mbar_init(smem_u32(&tmem_empty[i]), 2 * kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < COUT; i += kThreads) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: own slab rows + own half of the weight tile =====
    const __half* wsrc = p.wimg + (size_t)rank * ((size_t)p.nkc * 9 * KC * NH);
    int sl = 0, slph = 0, bs = 0, bph = 0, titer = 0;
    bool ok = true, first = true;
    for (int tile = cluster_id; tile < n_tiles && ok; tile += n_clusters) {
      const int it = titer++;
      const long long row0 = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS + 128 * (int)rank - 17;
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        if (kc == 0) CTRACE(0, it, 0);
        ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&slab_empty[sl]), slph ^ 1, p.errflag));
        if (!ok) break;
        if (kc == 0) CTRACE(0, it, 1);
        const uint32_t fb = smem_u32(&slab_full[sl]);
        if (elect_one()) {
          mbar_expect_tx(fb, SLAB_BYTES);
#pragma unroll
          for (int j = 0; j < KG; ++j)
            bulk_g2s(smem_u32(slab0 + sl * SLAB_STRIDE + j * kSlab2GroupBytes),
                     p.in + ((long long)(kc * KG + j) * p.mpad + row0) * 8, kSlab2GroupBytes, fb);
        }
        __syncwarp();
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
        if (resident && !first) continue;
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          if (!resident) {
            ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1, p.errflag));
            if (!ok) break;
          }
          const uint32_t bb = smem_u32(&b_full[bs]);
          if (elect_one()) {
            mbar_expect_tx(bb, STAGE_BYTES);
            bulk_g2s(smem_u32(bstage0 + (size_t)bs * STAGE_BYTES), wsrc + (size_t)(kc * 9 + ts * TPS) * ((size_t)KC * NH),
                     STAGE_BYTES, bb);
          }
          __syncwarp();
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
      }
      first = false;
    }
  } else if (warp == 3 && lane == 0) {
    // ===== relay: local TMA completion -> arrive on the leader's *_ready barrier =====
    int sl = 0, slph = 0, bs = 0, bph = 0, titer = 0;
    bool ok = true, first = true;
    for (int tile = cluster_id; tile < n_tiles && ok; tile += n_clusters) {
      const int it = titer++;
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = mbar_wait(smem_u32(&slab_full[sl]), slph, p.errflag);
        if (!ok) break;
        if (kc == 0) CTRACE(3, it, 0);
        mbar_arrive_cluster(mapa_u32(smem_u32(&slab_ready[sl]), 0));
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
        if (resident && !first) continue;
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          ok = mbar_wait(smem_u32(&b_full[bs]), bph, p.errflag);
          if (!ok) break;
          mbar_arrive_cluster(mapa_u32(smem_u32(&b_ready[bs]), 0));
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
      }
      first = false;
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA only): the whole warp runs the loop, one elected lane issues =====
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((256u >> 4) << 24);
    constexpr uint64_t DESC_HI = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    constexpr uint32_t A_LBO = (uint32_t)kSlab2GroupBytes >> 4;  // 162
    constexpr uint32_t B_LBO = (uint32_t)NH;                     // NH*16 >> 4
    int sl = 0, slph = 0, bs = 0, bph = 0, as = 0, aph = 0, titer = 0;
    bool ok = true, first = true;
    for (int tile = cluster_id; tile < n_tiles && ok; tile += n_clusters) {
      const int it = titer++;
      CTRACE(1, it, 0);
      ok = __all_sync(AP_FULL, mbar_wait_cl(smem_u32(&tmem_empty[as]), aph ^ 1, p.errflag));
      if (!ok) break;
      CTRACE(1, it, 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(as * COUT);
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = __all_sync(AP_FULL, mbar_wait_cl(smem_u32(&slab_ready[sl]), slph, p.errflag));
        if (!ok) break;
        if (kc == 0) CTRACE(1, it, 2);
        const uint32_t a_lo = (smem_u32(slab0 + sl * SLAB_STRIDE) >> 4) | (A_LBO << 16);
#pragma unroll
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          const int stage = resident ? kc * STAGES_PER_KC + ts : bs;
          if (!resident || first) {
            ok = __all_sync(AP_FULL, mbar_wait_cl(smem_u32(&b_ready[stage]), resident ? 0 : bph, p.errflag));
            if (!ok) break;
          }
          tc_fence_after();
          const uint32_t b_lo = (smem_u32(bstage0 + (size_t)stage * STAGE_BYTES) >> 4) | (B_LBO << 16);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
              const int tap = ts * TPS + t;
              const int off = 17 + (tap / 3 - 1) * 16 + (tap % 3 - 1);
#pragma unroll
              for (int j = 0; j < KC / 16; ++j) {
                const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (uint32_t)(off + 2 * j * (int)A_LBO));
                const uint64_t bd = DESC_HI | (uint64_t)(b_lo + (uint32_t)(t * (int)(TAP_BYTES >> 4) + 2 * j * (int)B_LBO));
                tc_mma_f16_2cta(acc, ad, bd, IDESC, (kc | tap | j) != 0);
              }
            }
            if (!resident) tc_commit_2cta(smem_u32(&b_empty[bs]));
            if (ts == STAGES_PER_KC - 1) {
              tc_commit_2cta(smem_u32(&slab_empty[sl]));
              if (kc == p.nkc - 1) tc_commit_2cta(smem_u32(&tmem_full[as]));
            }
          }
          __syncwarp();
          if (!resident && ++bs == p.nb) { bs = 0; bph ^= 1; }
        }
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
      }
      CTRACE(1, it, 3);
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
      first = false;
    }
  } else if (warp >= kCtrlWarps) {
    // ===== epilogue: this CTA's 128 rows; warp w reads TMEM lanes 32*(w&3)..+31, column half (w-4)>>2 =====
    constexpr int NCH = NH / 32;  // 32-column chunks per warp and tile: 1 / 2 / 4
    const int q = warp & 3;
    const int cbase = ((warp - kCtrlWarps) >> 2) * NH;
    const uint32_t empty_remote = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    int as = 0, aph = 0, titer = 0;
    bool ok = true;
    for (int tile = cluster_id; tile < n_tiles && ok; tile += n_clusters) {
      const int it = titer++;
      if (warp == kCtrlWarps) CTRACE(2, it, 0);
      ok = mbar_wait(smem_u32(&tmem_full[as]), aph, p.errflag);
      ok = __all_sync(AP_FULL, ok);
      if (!ok) break;
      if (warp == kCtrlWarps) CTRACE(2, it, 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * COUT);
      const int r = (int)rank * 128 + q * 32 + lane;
      const bool valid = ((r & 15) < p.W) && ((r >> 4) < p.H);
      const long long grow = NET_PAD_ROWS + (long long)tile * NET_TILE_ROWS + r;
      const uint32_t sb = smem_u32(s_bias), sx = smem_u32(s_hx);
      float hacc[1][6];
      if (HEAD) {
#pragma unroll
        for (int o = 0; o < 6; ++o) hacc[0][o] = 0.f;
      }
      auto body = [&](const int cb) {
        uint32_t v[2][32];
        ResidRegs<4, false> rr[2];
        resid_load<RESID, 4, false>(p, cb, grow, rr[0]);
        tmem_ld32(acc + cb, v[0]);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c0 = cb + i * 32;
          tmem_ld_wait_regs(v[i & 1]);
          if (i + 1 < NCH) {
            tmem_ld32(acc + (uint32_t)(cb + (i + 1) * 32), v[(i + 1) & 1]);
            resid_load<RESID, 4, false>(p, cb + (i + 1) * 32, grow, rr[(i + 1) & 1]);
          } else {
            // the accumulator stage is fully in registers: hand it back before the arithmetic and the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(empty_remote + (uint32_t)(as * 8));
            if (warp == kCtrlWarps) CTRACE(2, it, 2);
          }
          epi_chunk<COUT, RESID, HEAD, 4>(p, hw, v[i & 1], c0, grow, valid, sb, hacc[0], rr[i & 1]);
        }
      };
      if constexpr (HEAD) {
        if (cbase == 0) body(0); else body(NH);
        // one bar.sync instruction for both warps of the named barrier (not one per inlined copy of body)
        const int rows[1] = {r};
        head_finish<1, 2>(p, hw, hacc, sx, 1 + q, q, lane, cbase != 0 ? 1 : 0, tile, rows);
      } else {
        body(cbase);
      }
      if (warp == kCtrlWarps) CTRACE(2, it, 3);
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody exits while the peer may still arrive on its barriers / read its smem
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// "Two boards per CTA pair" version for the 256-channel layers (conv5, conv_final of the simple net).
// A cluster tile is 2 boards x 128 output channels: CTA r stages the full 290-row slab of ITS board and 64
// of the 128 weight columns; per (tap, K step) the leader issues two cta_group::2 MMAs (M=256 = rows
// [128h,128h+128) of both boards, N=128).  Per FLOP this moves 35 % fewer bytes from L2 into the SMs than
// the one-board pair tile (weights are shared by twice as many rows) and the accumulator (2 x 128 columns)
// is still double buffered.  The one-board pair form needs ~50 B/clk/SM on conv5 (32 weights + 4.5 slab +
// 14 output) against a fabric that sustains ~42 B/clk/SM: ncu shows 82 % tensor-pipe activity there.
// Tiles are ordered column half 0 for every board pair, then column half 1, so a layer whose half weight
// tensor fits the ring (cin 128: 6 stages) only re-reads it once per launch.
// ---------------------------------------------------------------------------------------------
constexpr int kNT4 = 128;  // tile columns

// iteration -> (board pair, column half) of one cluster.  Plain layers run column half 0 for all their board pairs,
// then column half 1 (weights stay in the ring); the fused-head layer alternates the halves per board pair so that
// one epilogue warp sees all 256 channels of its rows in two consecutive tiles and keeps the head sums in registers.
template <bool HEAD, int NHALVES>
struct Tc4Iter {
  int n_my, cluster_id, n_clusters;
  __device__ Tc4Iter(int n_pairs, int cid, int ncl) : cluster_id(cid), n_clusters(ncl) {
    n_my = cid < n_pairs ? (n_pairs - cid + ncl - 1) / ncl : 0;
  }
  __device__ int count() const { return NHALVES * n_my; }
  __device__ int nh(int it) const { return NHALVES == 1 ? 0 : (HEAD ? (it & 1) : (it >= n_my ? 1 : 0)); }
  __device__ int bp(int it) const {
    return cluster_id + (NHALVES == 1 ? it : (HEAD ? (it >> 1) : (it >= n_my ? it - n_my : it))) * n_clusters;
  }
  __device__ bool first_of_half(int it) const { return HEAD ? true : (it == 0 || (NHALVES == 2 && it == n_my)); }
};

template <int KC, bool RESID, bool HEAD, int COUT = 256>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_conv3x3_tc4(const __grid_constant__ ConvParams p, const __grid_constant__ HeadArg<HEAD> hw) {
  static_assert(COUT == 256 || (COUT == 128 && !HEAD), "column tiles of 128: 128- or 256-channel layers");
  constexpr int TPS = kTps2;
  constexpr int STAGES_PER_KC = 9 / TPS;
  constexpr int ACC_STAGES = 2;
  constexpr int TMEM_COLS = 512;                      // 2 stages x 2 row halves x 128 columns
  constexpr int KG = KC / 8;
  constexpr int NHC = kNT4 / 2;                       // B columns staged by one CTA
  constexpr uint32_t SLAB_BYTES = KG * kSlabGroupBytes;
  constexpr uint32_t SLAB_STRIDE = (SLAB_BYTES + 127u) & ~127u;
  constexpr uint32_t TAP_BYTES = (uint32_t)KC * NHC * 2;
  constexpr uint32_t STAGE_BYTES = TPS * TAP_BYTES;

  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(AP_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* slab0 = smem;
  uint8_t* bstage0 = smem + (size_t)p.ns * SLAB_STRIDE;
  uint64_t* bars = (uint64_t*)(bstage0 + (size_t)p.nb * STAGE_BYTES);
  uint64_t* slab_full = bars;
  uint64_t* slab_ready = slab_full + kMaxSlabs;
  uint64_t* slab_empty = slab_ready + kMaxSlabs;
  uint64_t* b_full = slab_empty + kMaxSlabs;
  uint64_t* b_ready = b_full + kMaxStages;
  uint64_t* b_empty = b_ready + kMaxStages;
  uint64_t* tmem_full = b_empty + kMaxStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* s_bias = (float*)(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);  // [COUT]
  float* s_hx = s_bias + COUT;                  // HEAD: [128][12] partial exchange between the two column-half warps

  const int n_boards = p.n_tiles_dev ? *p.n_tiles_dev : p.n_tiles;
  const Tc4Iter<HEAD, COUT / kNT4> iter((n_boards + 1) >> 1, blockIdx.x >> 1, gridDim.x >> 1);
  const int n_iter = iter.count();
  // the ring holds exactly one column half of the layer: a stage keeps its content from tile to tile
  const bool keeps = !HEAD && p.nkc * STAGES_PER_KC == p.nb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxSlabs; ++i) {
      mbar_init(smem_u32(&slab_full[i]), 1);
      mbar_init(smem_u32(&slab_ready[i]), 2);
      mbar_init(smem_u32(&slab_empty[i]), 1);
    }
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_ready[i]), 2);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full[i]), 1);
      mbar_init(smem_u32(&tmem_empty[i]), 2 * kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < COUT; i += kThreads) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the slab of this CTA's board + this CTA's 64 weight columns =====
    int sl = 0, slph = 0, bs = 0, bph = 0;
    bool ok = true;
    for (int it = 0; it < n_iter && ok; ++it) {
      const int nh = iter.nh(it), bp = iter.bp(it);
      const bool first = iter.first_of_half(it);
      // image [nh][rank][kc][tap][KC/8][64][8]
      const __half* wsrc = p.wimg + ((size_t)nh * 2 + rank) * ((size_t)p.nkc * 9 * KC * NHC);
      int board = 2 * bp + (int)rank;
      if (board >= n_boards) board = n_boards - 1;  // odd batch: the peer re-reads the last board, stores nothing
      const long long row0 = NET_PAD_ROWS + (long long)board * NET_TILE_ROWS - 17;
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&slab_empty[sl]), slph ^ 1, p.errflag));
        if (!ok) break;
        const uint32_t fb = smem_u32(&slab_full[sl]);
        if (elect_one()) {
          mbar_expect_tx(fb, SLAB_BYTES);
#pragma unroll
          for (int j = 0; j < KG; ++j)
            bulk_g2s(smem_u32(slab0 + sl * SLAB_STRIDE + j * kSlabGroupBytes),
                     p.in + ((long long)(kc * KG + j) * p.mpad + row0) * 8, kSlabGroupBytes, fb);
        }
        __syncwarp();
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          ok = __all_sync(AP_FULL, mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1, p.errflag));
          if (!ok) break;
          const uint32_t bb = smem_u32(&b_full[bs]);
          if (elect_one()) {
            if (keeps && !first) {
              mbar_arrive(bb);  // the stage still holds (nh, kc, ts): complete the phase without moving data
            } else {
              mbar_expect_tx(bb, STAGE_BYTES);
              bulk_g2s(smem_u32(bstage0 + (size_t)bs * STAGE_BYTES), wsrc + (size_t)(kc * 9 + ts * TPS) * ((size_t)KC * NHC),
                       STAGE_BYTES, bb);
            }
          }
          __syncwarp();
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 3 && lane == 0) {
    // ===== relay: local TMA completion -> arrive on the leader's *_ready barrier =====
    int sl = 0, slph = 0, bs = 0, bph = 0;
    bool ok = true;
    for (int it = 0; it < n_iter && ok; ++it)
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = mbar_wait(smem_u32(&slab_full[sl]), slph, p.errflag);
        if (!ok) break;
        mbar_arrive_cluster(mapa_u32(smem_u32(&slab_ready[sl]), 0));
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          ok = mbar_wait(smem_u32(&b_full[bs]), bph, p.errflag);
          if (!ok) break;
          mbar_arrive_cluster(mapa_u32(smem_u32(&b_ready[bs]), 0));
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
      }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA only) =====
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(kNT4 >> 3) << 17) | ((256u >> 4) << 24);
    constexpr uint64_t DESC_HI = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    constexpr uint32_t A_LBO = (uint32_t)kSlabGroupBytes >> 4;  // 290
    constexpr uint32_t B_LBO = (uint32_t)NHC;                   // 64 columns * 16 B >> 4
    int sl = 0, slph = 0, bs = 0, bph = 0, as = 0, aph = 0;
    bool ok = true;
    for (int it = 0; it < n_iter && ok; ++it) {
      ok = __all_sync(AP_FULL, mbar_wait_cl(smem_u32(&tmem_empty[as]), aph ^ 1, p.errflag));
      if (!ok) break;
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(as * 2 * kNT4);
      for (int kc = 0; kc < p.nkc && ok; ++kc) {
        ok = __all_sync(AP_FULL, mbar_wait_cl(smem_u32(&slab_ready[sl]), slph, p.errflag));
        if (!ok) break;
        const uint32_t a_lo = (smem_u32(slab0 + sl * SLAB_STRIDE) >> 4) | (A_LBO << 16);
#pragma unroll
        for (int ts = 0; ts < STAGES_PER_KC; ++ts) {
          ok = __all_sync(AP_FULL, mbar_wait_cl(smem_u32(&b_ready[bs]), bph, p.errflag));
          if (!ok) break;
          tc_fence_after();
          const uint32_t b_lo = (smem_u32(bstage0 + (size_t)bs * STAGE_BYTES) >> 4) | (B_LBO << 16);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < TPS; ++t) {
              const int tap = ts * TPS + t;
              const int off = 17 + (tap / 3 - 1) * 16 + (tap % 3 - 1);
#pragma unroll
              for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < KC / 16; ++j) {
                  const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (uint32_t)(off + half * 128 + 2 * j * (int)A_LBO));
                  const uint64_t bd = DESC_HI | (uint64_t)(b_lo + (uint32_t)(t * (int)(TAP_BYTES >> 4) + 2 * j * (int)B_LBO));
                  tc_mma_f16_2cta(acc + (uint32_t)(half * kNT4), ad, bd, IDESC, (kc | tap | j) != 0);
                }
              }
            }
            tc_commit_2cta(smem_u32(&b_empty[bs]));
            if (ts == STAGES_PER_KC - 1) {
              tc_commit_2cta(smem_u32(&slab_empty[sl]));
              if (kc == p.nkc - 1) tc_commit_2cta(smem_u32(&tmem_full[as]));
            }
          }
          __syncwarp();
          if (++bs == p.nb) { bs = 0; bph ^= 1; }
        }
        if (++sl == p.ns) { sl = 0; slph ^= 1; }
      }
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
  } else if (warp >= kCtrlWarps) {
    // ===== epilogue: this CTA's board (256 rows) x 128 columns; warp = lane quarter x 64-column half =====
    const int q = warp & 3;
    const int ch = (warp - kCtrlWarps) >> 2;
    const int cw = ch * 64;
    const uint32_t empty_remote = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const uint32_t sb = smem_u32(s_bias), sx = smem_u32(s_hx);
    int as = 0, aph = 0;
    bool ok = true;
    float hsum[2][6];  // HEAD: head-conv partial sums of this thread's two rows over its 2 x 64 channels
    for (int it = 0; it < n_iter && ok; ++it) {
      const int nh = iter.nh(it), bp = iter.bp(it);
      ok = mbar_wait(smem_u32(&tmem_full[as]), aph, p.errflag);
      ok = __all_sync(AP_FULL, ok);
      if (!ok) break;
      tc_fence_after();
      const int board = 2 * bp + (int)rank;
      const bool live = board < n_boards;
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * kNT4 + cw);
      const long long grow0 = NET_PAD_ROWS + (long long)(live ? board : 0) * NET_TILE_ROWS + q * 32 + lane;
      if (HEAD && nh == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int o = 0; o < 6; ++o) hsum[r][o] = 0.f;
      }
      auto body = [&](const int cabs) {  // cabs = first output channel of this warp (compile-time under HEAD)
        uint32_t v[2][32];
        float dummy[6];
        ResidRegs<4, false> rr[2];
        if (live) resid_load<RESID, 4, false>(p, cabs, grow0, rr[0]);
        tmem_ld32(acc, v[0]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // (row half, 32-column chunk)
          const int half = i >> 1, c0 = cabs + (i & 1) * 32;
          tmem_ld_wait_regs(v[i & 1]);
          if (i + 1 < 4) {
            tmem_ld32(acc + (uint32_t)(((i + 1) >> 1) * kNT4 + ((i + 1) & 1) * 32), v[(i + 1) & 1]);
            if (live) resid_load<RESID, 4, false>(p, cabs + ((i + 1) & 1) * 32, grow0 + ((i + 1) >> 1) * 128, rr[(i + 1) & 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(empty_remote + (uint32_t)(as * 8));
          }
          const int r = half * 128 + q * 32 + lane;
          const bool valid = ((r & 15) < p.W) && ((r >> 4) < p.H);
          if (live) {
            if constexpr (HEAD)
              epi_chunk<COUT, RESID, true, 4>(p, hw, v[i & 1], c0, grow0 + half * 128, valid, sb, hsum[half], rr[i & 1]);
            else
              epi_chunk<COUT, RESID, false, 4>(p, hw, v[i & 1], c0, grow0 + half * 128, valid, sb, dummy, rr[i & 1]);
          }
        }
      };
      if constexpr (HEAD) {
        switch (nh * 2 + ch) {
          case 0: body(0); break;
          case 1: body(64); break;
          case 2: body(128); break;
          default: body(192); break;
        }
        if (nh == 1) {
          // both column halves of this board are in hsum: combine the two 64-column warps, finish, write the FC operand
          const int rows[2] = {q * 32 + lane, 128 + q * 32 + lane};
          head_finish<2, 2>(p, hw, hsum, sx, 1 + q, q, lane, ch, board, rows, live);
        }
      } else {
        body(nh * kNT4 + cw);
      }
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

struct SmemPlan {
  int ns, nb, bytes;
};

// HEAD instantiations also hold the partial-exchange slots [2][128][6]
constexpr int head_smem_bytes(int) { return 2 * 128 * 6 * 4; }

// slabs + B ring + barriers + TMEM slot + bias inside the 227 KB opt-in limit
SmemPlan plan_smem(int cout, int kc, int nkc, bool head, int split = 0, int taps = 9) {
  const int tps = (cout == 256 || taps == 1) ? 1 : 3;
  const int mul = split ? 2 : 1, mulw = split == 1 ? 2 : 1;  // split 2: hi + lo activations, fp16 weights
  const int slab = ((mul * (kc >> 3) * kSlabGroupBytes) + 127) & ~127;
  const int stage = mulw * tps * kc * cout * 2;
  const int fixed = (2 * kMaxSlabs + 2 * kMaxStages + 4) * 8 + 16 + cout * 4 + 128 + (head ? head_smem_bytes(cout) : 0);
  const int budget = 227 * 1024 - fixed;
  const int all = nkc * (taps / tps);  // stages that hold the whole layer
  SmemPlan s;
  if (all <= kMaxStages && 2 * slab + all * stage <= budget) {
    s.nb = all;  // resident weights; spend what is left on a deeper slab ring
    s.ns = (budget - all * stage) / slab;
    if (s.ns > kMaxSlabs) s.ns = kMaxSlabs;
  } else {
    s.ns = 2;
    s.nb = (budget - 2 * slab) / stage;
    if (s.nb > kMaxStages) s.nb = kMaxStages;
    if (s.nb >= all) s.nb = all - 1;  // never "resident" by accident with a ring smaller than planned
  }
  s.bytes = s.ns * slab + s.nb * stage + fixed;
  return s;
}

SmemPlan plan_smem2(int cout, int kc, int nkc, bool head) {
  const int slab = (((kc >> 3) * kSlab2GroupBytes) + 127) & ~127;
  const int stage = kTps2 * kc * (cout / 2) * 2;
  const int fixed = (3 * kMaxSlabs + 3 * kMaxStages + 4) * 8 + 16 + cout * 4 + 128 + (head ? head_smem_bytes(cout) : 0);
  const int budget = 227 * 1024 - fixed;
  const int all = nkc * (9 / kTps2);
  SmemPlan s;
  if (all <= kMaxStages && 2 * slab + all * stage <= budget) {
    s.nb = all;
    s.ns = (budget - all * stage) / slab;
    if (s.ns > kMaxSlabs) s.ns = kMaxSlabs;
  } else {
    s.ns = 3;
    s.nb = (budget - s.ns * slab) / stage;
    if (s.nb > kMaxStages) s.nb = kMaxStages;
    if (s.nb >= all) s.nb = all - 1;
  }
  s.bytes = s.ns * slab + s.nb * stage + fixed;
  return s;
}

SmemPlan plan_smem4(int kc, int nkc, bool head) {
  const int slab = (((kc >> 3) * kSlabGroupBytes) + 127) & ~127;
  const int stage = kTps2 * kc * (kNT4 / 2) * 2;
  const int fixed = (3 * kMaxSlabs + 3 * kMaxStages + 4) * 8 + 16 + 256 * 4 + 128 + (head ? 128 * 12 * 4 : 0);
  const int budget = 227 * 1024 - fixed;
  const int all = nkc * (9 / kTps2);
  SmemPlan s;
  s.ns = 2;
  s.nb = (budget - s.ns * slab) / stage;
  if (s.nb > kMaxStages) s.nb = kMaxStages;
  if (s.nb > all) s.nb = all;  // exactly one column half of the layer: stages keep their content between tiles
  if (s.nb == all && budget - all * stage >= 3 * slab) s.ns = 3;
  s.bytes = s.ns * slab + s.nb * stage + fixed;
  return s;
}

template <int KC, bool RESID, bool HEAD, int COUT = 256>
int launch4(ap_engine* e, const ConvParams& p, const HeadArg<HEAD>& hw, int grid, int smem) {
  k_conv3x3_tc4<KC, RESID, HEAD, COUT><<<grid, kThreads, smem, e->stream>>>(p, hw);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}

template <int COUT, int KC, bool RESID, bool HEAD>
int launch1(ap_engine* e, const ConvParams& p, const HeadArg<HEAD>& hw, int grid, int smem) {
  k_conv3x3_tc<COUT, KC, RESID, HEAD><<<grid, EpiCfg<HEAD>::THREADS, smem, e->stream>>>(p, hw);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}
template <int COUT, int KC, bool RESID, bool HEAD>
int launch2(ap_engine* e, const ConvParams& p, const HeadArg<HEAD>& hw, int grid, int smem) {
  k_conv3x3_tc2<COUT, KC, RESID, HEAD><<<grid, kThreads, smem, e->stream>>>(p, hw);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}

template <int COUT, int KC>
int launch_t(ap_engine* e, const ConvParams& p, bool resid, bool pair, int grid, int smem) {
  const HeadArg<false> none{};
  if (pair)
    return resid ? launch2<COUT, KC, true, false>(e, p, none, grid, smem) : launch2<COUT, KC, false, false>(e, p, none, grid, smem);
  return resid ? launch1<COUT, KC, true, false>(e, p, none, grid, smem) : launch1<COUT, KC, false, false>(e, p, none, grid, smem);
}

template <int COUT, int KC>
cudaError_t optin_t() {
  const cudaFuncAttribute a = cudaFuncAttributeMaxDynamicSharedMemorySize;
  const int lim = 227 * 1024;
  cudaError_t r = cudaFuncSetAttribute(k_conv3x3_tc<COUT, KC, true, false>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc<COUT, KC, false, false>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc2<COUT, KC, true, false>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc2<COUT, KC, false, false>, a, lim);
  return r;
}

// SPLIT instantiations (residual net at near-fp32 accuracy): stem, convA, convB (+residual), last convB with heads
template <int M>
cudaError_t optin_split_m() {
  const cudaFuncAttribute a = cudaFuncAttributeMaxDynamicSharedMemorySize;
  const int lim = 227 * 1024;
  cudaError_t r = cudaFuncSetAttribute(k_conv3x3_tc<128, 16, false, false, M>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc<128, 32, false, false, M>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc<128, 32, true, false, M>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc<128, 32, true, true, M>, a, lim);
  return r;
}
cudaError_t optin_split() {
  cudaError_t r = optin_split_m<1>();
  return r == cudaSuccess ? optin_split_m<2>() : r;
}

template <int COUT, int KC, bool RESID, bool HEAD, int M>
int launch1s(ap_engine* e, const ConvParams& p, const HeadArg<HEAD>& hw, int grid, int smem) {
  k_conv3x3_tc<COUT, KC, RESID, HEAD, M><<<grid, EpiCfg<HEAD>::THREADS, smem, e->stream>>>(p, hw);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}
template <int M>
int launch_split(ap_engine* e, const ConvParams& p, const NetHeadW* head_w, bool resid, int kc, int grid, int smem) {
  const HeadArg<false> none{};
  if (head_w) {
    HeadArg<true> hw;
    hw.h = *head_w;
    return launch1s<128, 32, true, true, M>(e, p, hw, grid, smem);
  }
  if (kc == 16) return launch1s<128, 16, false, false, M>(e, p, none, grid, smem);
  return resid ? launch1s<128, 32, true, false, M>(e, p, none, grid, smem)
               : launch1s<128, 32, false, false, M>(e, p, none, grid, smem);
}

// 1x1 instantiations (Inception-ResNet variant: tower stems and the up-projection, 128 -> 128)
template <bool RESID>
int launch1x1(ap_engine* e, const ConvParams& p, int grid, int smem) {
  const HeadArg<false> none{};
  k_conv3x3_tc<128, 64, RESID, false, false, 1><<<grid, kThreads, smem, e->stream>>>(p, none);
  AP_LAUNCH_CHECK(e);
  return AP_OK;
}

// HEAD instantiations: the last trunk layer of the simple net (256 -> 256, no residual) and of the
// residual net (128 -> 128 with residual)
cudaError_t optin_head() {
  const cudaFuncAttribute a = cudaFuncAttributeMaxDynamicSharedMemorySize;
  const int lim = 227 * 1024;
  cudaError_t r = cudaFuncSetAttribute(k_conv3x3_tc<256, 64, false, true>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc2<256, 64, false, true>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc<128, 64, true, true>, a, lim);
  if (r == cudaSuccess) r = cudaFuncSetAttribute(k_conv3x3_tc2<128, 64, true, true>, a, lim);
  return r;
}

}  // namespace

int conv_tc_smem_bytes(const ConvLayer& L, int* out_nb) {
  const int kc = conv_tc_kc(L, 0);
  SmemPlan s = plan_smem(L.cout, kc, L.cin_pad / kc, false);
  if (out_nb) *out_nb = s.nb;
  return s.bytes;
}

// K chunk: 64 channels where the layer has them, else 32 (narrow tower layers of the inception variant, 64 outputs
// only) or 16
static int kc_of(int cin_pad) { return cin_pad % 64 == 0 ? 64 : (cin_pad % 32 == 0 ? 32 : 16); }
bool conv_tc_supported(int cin_pad, int cout) {
  const int kc = kc_of(cin_pad);
  return (cout == 64 || cout == 128 || cout == 256) && cin_pad % kc == 0 && (kc != 32 || cout == 64);
}

// can this layer run the fused head epilogue?
// K chunk of a layer: the split-precision kernels stage hi + lo of both operands, so they use half the chunk
int conv_tc_kc(const ConvLayer& L, int split) {
  const int kc = kc_of(L.cin_pad);
  return (split && kc == 64) ? 32 : kc;
}

bool conv_tc_split_supported(const ConvLayer& L) { return L.cout == 128 && (L.cin_pad == 16 || L.cin_pad % 32 == 0); }

bool conv_tc_head_supported(const ConvLayer& L) {
  if (L.cin_pad < 64 || L.cin_pad % 64) return false;
  return (L.cout == 256 && L.resid_buf < 0) || (L.cout == 128 && L.resid_buf >= 0);
}

// per-device opt-in to > 48 KB dynamic shared memory for every instantiation (called from ap_net_load)
int conv_tc_configure(ap_engine* e) {
  AP_CUDA(e, (optin_t<64, 16>()));
  AP_CUDA(e, (optin_t<64, 64>()));
  AP_CUDA(e, (optin_t<128, 16>()));
  AP_CUDA(e, (optin_t<128, 64>()));
  AP_CUDA(e, (optin_t<256, 16>()));
  AP_CUDA(e, (optin_t<256, 64>()));
  AP_CUDA(e, optin_head());
  AP_CUDA(e, optin_split());
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc<128, 64, false, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc<128, 64, true, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc<64, 32, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc4<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc4<64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc4<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc4<64, false, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  AP_CUDA(e, (cudaFuncSetAttribute(k_conv3x3_tc4<64, true, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
  return AP_OK;
}

int conv_tc_launch(ap_engine* e, NetState* n, const ConvLayer& L, int n_boards, const int* n_boards_dev, bool head) {
  const int split = n->split;  // 0 fp16, 1 split both operands, 2 split activations only
  ConvParams p;
  p.in = (L.in_buf < 0) ? n->feat : n->act[L.in_buf];
  p.out = n->act[L.out_buf];
  p.resid = (L.resid_buf >= 0) ? n->act[L.resid_buf] : nullptr;
  p.wimg = L.wimg;
  p.in_lo = (L.in_buf < 0) ? n->feat_lo : n->act_lo[L.in_buf];
  p.out_lo = n->act_lo[L.out_buf];
  p.resid_lo = (L.resid_buf >= 0) ? n->act_lo[L.resid_buf] : nullptr;
  p.wimg_lo = L.wimg_lo;
  p.bias = L.shift;
  p.mpad = n->mpad;
  p.in_goff = L.in_coff / 8;
  p.out_coff = L.out_coff;
  p.cout_store = L.cout_store > 0 ? L.cout_store : L.cout;
  const int kc = conv_tc_kc(L, split);
  p.nkc = L.cin_pad / kc;
  p.relu = L.relu;
  p.n_tiles = n_boards;
  p.n_tiles_dev = n_boards_dev;
  p.W = n->W;
  p.H = n->H;
  p.errflag = n->d_err;
  p.ha = n->fc_a;
  p.ha_rows = n->fc_rows;
  p.ha_kg = n->fc_kp / 8;
#ifdef AP_CONV_TRACE
  p.trace = g_conv_trace;
#endif
  if (!conv_tc_supported(L.cin_pad, L.cout)) return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: unsupported channel counts");
  if (head && !conv_tc_head_supported(L)) return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: no fused-head instantiation for this layer");
  const bool resid = p.resid != nullptr;
  if (L.in_coff && (!L.force_single || (L.in_coff & 7)))
    return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: input channel slices need the single-CTA kernel and a multiple of 8");
  if (L.ksz == 1) {
    // 1x1 conv: single-CTA kernel, centre tap only
    if (split || head || L.cout != 128 || kc != 64) return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: 1x1 convs are built for 128 output channels");
    const SmemPlan s1 = plan_smem(L.cout, kc, p.nkc, false, false, 1);
    p.ns = s1.ns;
    p.nb = s1.nb;
    const int grid1 = n_boards < n->sm_count ? n_boards : n->sm_count;
    return resid ? launch1x1<true>(e, p, grid1, s1.bytes) : launch1x1<false>(e, p, grid1, s1.bytes);
  }
  if (split) {
    // near-fp32 path of the residual net: single-CTA kernel, three products per K step
    if (!conv_tc_split_supported(L)) return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: no split-precision instantiation for this layer");
    const SmemPlan s = plan_smem(L.cout, kc, p.nkc, head, split);
    p.ns = s.ns;
    p.nb = s.nb;
    const int grid = n_boards < n->sm_count ? n_boards : n->sm_count;
    if (head && (!resid || kc != 32))
      return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: split fused head needs a residual 128-channel layer");
    return split == 1 ? launch_split<1>(e, p, head ? &n->head_w : nullptr, resid, kc, grid, s.bytes)
                      : launch_split<2>(e, p, head ? &n->head_w : nullptr, resid, kc, grid, s.bytes);
  }
  // 256-channel layers without a fused head: two boards per CTA pair, 128-column tiles (AP_CONV4=0 disables)
  if (n->conv4_128 && !L.force_single && L.cout == 128 && kc == 64 && n->conv_mode == 0 && !head && (n->conv4_128 > 1 || L.cin_pad == 64)) {
    p.wimg = L.wimg4;
    const SmemPlan s4 = plan_smem4(kc, p.nkc, false);
    p.ns = s4.ns;
    p.nb = s4.nb;
    const int pairs = n->sm_count / 2, bpairs = (n_boards + 1) / 2;
    const int grid4 = 2 * (bpairs < pairs ? bpairs : pairs);
    const HeadArg<false> none{};
    return resid ? launch4<64, true, false, 128>(e, p, none, grid4, s4.bytes) : launch4<64, false, false, 128>(e, p, none, grid4, s4.bytes);
  }
  if (n->conv4 && L.cout == 256 && kc == 64 && n->conv_mode == 0 && !(head && (resid || n->conv4 < 2))) {
    p.wimg = L.wimg4;
    const SmemPlan s4 = plan_smem4(kc, p.nkc, head);
    p.ns = s4.ns;
    p.nb = s4.nb;
    const int pairs = n->sm_count / 2, bpairs = (n_boards + 1) / 2;
    const int grid4 = 2 * (bpairs < pairs ? bpairs : pairs);
    const HeadArg<false> none{};
    if (head) {
      HeadArg<true> hw;
      hw.h = n->head_w;
      return launch4<64, false, true>(e, p, hw, grid4, s4.bytes);
    }
    return resid ? launch4<64, true, false>(e, p, none, grid4, s4.bytes) : launch4<64, false, false>(e, p, none, grid4, s4.bytes);
  }
  // auto: CTA pairs where they measured faster on B200 (K = 9*128: conv4, conv5, the residual blocks, the fused-head
  // layer); the memory-bound small layers run the single-CTA kernel
  const bool pair = !L.force_single &&
                    (n->conv_mode == 2 || (n->conv_mode == 0 && (L.cin_pad == 128 || (head && n->head_pair) ||
                                                                 (n->pair_cin64 && L.cin_pad == 64 && L.cout == 128))));
  SmemPlan s;
  int grid;
  if (pair) {
    // CTA pairs: one board per cluster of two
    p.wimg = L.wimg2;
    s = plan_smem2(L.cout, kc, p.nkc, head);
    const int pairs = n->sm_count / 2;
    grid = 2 * (n_boards < pairs ? n_boards : pairs);
  } else {
    s = plan_smem(L.cout, kc, p.nkc, head);
    grid = n_boards < n->sm_count ? n_boards : n->sm_count;
  }
  p.ns = s.ns;
  p.nb = s.nb;
  if (head) {
    HeadArg<true> hw;
    hw.h = n->head_w;  // host copy of the folded 1x1 weights (refreshed by net_prep), passed in the constant bank
    if (L.cout == 256)
      return pair ? launch2<256, 64, false, true>(e, p, hw, grid, s.bytes) : launch1<256, 64, false, true>(e, p, hw, grid, s.bytes);
    return pair ? launch2<128, 64, true, true>(e, p, hw, grid, s.bytes) : launch1<128, 64, true, true>(e, p, hw, grid, s.bytes);
  }
  if (kc == 32) {  // narrow tower layers (inception variant): single-CTA kernel, 64 output columns
    const HeadArg<false> none{};
    if (pair || resid || L.cout != 64) return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: K chunk 32 is built for plain 64-output layers");
    return launch1<64, 32, false, false>(e, p, none, grid, s.bytes);
  }
  switch (L.cout * 100 + kc) {
    case 64 * 100 + 16: return launch_t<64, 16>(e, p, resid, pair, grid, s.bytes);
    case 64 * 100 + 64: return launch_t<64, 64>(e, p, resid, pair, grid, s.bytes);
    case 128 * 100 + 16: return launch_t<128, 16>(e, p, resid, pair, grid, s.bytes);
    case 128 * 100 + 64: return launch_t<128, 64>(e, p, resid, pair, grid, s.bytes);
    case 256 * 100 + 16: return launch_t<256, 16>(e, p, resid, pair, grid, s.bytes);
    case 256 * 100 + 64: return launch_t<256, 64>(e, p, resid, pair, grid, s.bytes);
  }
  return ap_fail(e, AP_ERR_BAD_ARG, "conv_tc: unsupported channel counts");
}
