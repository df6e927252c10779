// Micro-benchmark (development tool, not product): raw tcgen05.mma issue/execute rate on sm_100a for
// the operand layouts conv_tc.cu uses (K-major, no swizzle, 8x16B core matrices).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_probe tools/mma_probe.cu
// Prints cycles per MMA for N in {32,64,128,256}, A start aligned / misaligned by one 16-byte row,
// descriptors hoisted vs recomputed per instruction, and with a concurrent TMA stream into smem.
#include <cstdio>
#include <cuda_fp16.h>
#include "../alphapig_b200/csrc/ptx.cuh"

struct ProbeParams {
  int n;          // UMMA N
  int a_off;      // A start offset in 16-byte rows
  int recompute;  // 1: rebuild descriptors per MMA like conv_tc.cu
  int iters;      // MMAs = iters * 8
  int tma;        // 1: a second thread streams bulk copies into a spare smem region meanwhile
  int b_shared;   // unused
  const __half* gsrc;
  long long* out_cycles;
  int* errflag;
};

__global__ void __launch_bounds__(128, 1) k_probe(ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_tma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A region: 64 channels x 290 rows (like a slab), B region: 64 x 256
  uint8_t* a_reg = smem;
  uint8_t* b_reg = smem + 8 * 290 * 16 + 128;
  uint8_t* t_reg = b_reg + 64 * 256 * 2;
  for (int i = threadIdx.x; i < (8 * 290 * 16 + 128 + 64 * 256 * 2) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    mbar_init(smem_u32(&bar_tma), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t group = 290 * 16;
    const uint32_t b_lbo = (uint32_t)p.n * 16;
    const uint32_t abase = smem_u32(a_reg) + (uint32_t)(17 + p.a_off) * 16;
    const uint32_t bbase = smem_u32(b_reg);
    long long t0 = clock64();
    if (!p.recompute) {
      uint64_t ad[4], bd[4];
      for (int j = 0; j < 4; ++j) {
        ad[j] = make_desc(abase + 2 * j * group, group, 128);
        bd[j] = make_desc(bbase + 2 * j * b_lbo, b_lbo, 128);
      }
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) tc_mma_f16(tmem_base + h * p.n, ad[j], bd[j], idesc, 1);
      }
    } else {
      for (int it = 0; it < p.iters; ++it) {
        const int tap = it % 9;
        const int off = (tap / 3 - 1) * 16 + (p.a_off ? (tap % 3 - 1) : 0);
        for (int half = 0; half < 2; ++half) {
          const uint32_t arow = smem_u32(a_reg) + (uint32_t)(17 + off + half * 128) * 16;
          for (int j = 0; j < 4; ++j) {
            const uint64_t a = make_desc(arow + (uint32_t)(2 * j) * group, group, 128);
            const uint64_t b = make_desc(bbase + (uint32_t)(2 * j) * b_lbo, b_lbo, 128);
            tc_mma_f16(tmem_base + (uint32_t)(half * p.n), a, b, idesc, (it | j) != 0);
          }
        }
      }
    }
    long long t1 = clock64();
    tc_commit(smem_u32(&bar_done));
    mbar_wait(smem_u32(&bar_done), 0, p.errflag);
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      p.out_cycles[0] = t1 - t0;
      p.out_cycles[1] = t2 - t0;
    }
  } else if (warp == 2 && lane == 0 && p.tma) {
    // stream 16 KB bulk copies into a scratch region at the rate the conv kernel's B ring does
    uint32_t ph = 0;
    const int copies = p.iters * 8 * (p.n / 2) / 512;  // ~ one 16 KB tile per 512 tensor cycles
    for (int c = 0; c < copies; ++c) {
      mbar_expect_tx(smem_u32(&bar_tma), 16384);
      bulk_g2s(smem_u32(t_reg), p.gsrc + ((size_t)(blockIdx.x * 64 + (c & 63)) * 8192), 16384, smem_u32(&bar_tma));
      if (!mbar_wait(smem_u32(&bar_tma), ph, p.errflag)) break;
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int main() {
  long long* d_out;
  int* d_err;
  __half* d_src;
  cudaMalloc(&d_out, 16);
  cudaMalloc(&d_err, 4);
  cudaMemset(d_err, 0, 4);
  cudaMalloc(&d_src, (size_t)148 * 64 * 16384);
  cudaMemset(d_src, 0, (size_t)148 * 64 * 16384);
  const int smem = 8 * 290 * 16 + 128 + 64 * 256 * 2 + 16384 + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("N a_off recompute tma  issue_cyc/mma  total_cyc/mma  ms\n");
  for (int tma = 0; tma < 2; ++tma)
    for (int rec = 0; rec < 2; ++rec)
      for (int aoff = 0; aoff < 2; ++aoff)
        for (int n = 32; n <= 256; n *= 2) {
          ProbeParams p{n, aoff, rec, 2048, tma, 0, d_src, d_out, d_err};
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          k_probe<<<148, 128, smem>>>(p);  // warm
          cudaEventRecord(e0);
          k_probe<<<148, 128, smem>>>(p);
          cudaEventRecord(e1);
          cudaError_t st = cudaDeviceSynchronize();
          if (st != cudaSuccess) {
            printf("CUDA error: %s\n", cudaGetErrorString(st));
            return 1;
          }
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          long long h[2];
          cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
          printf("%3d %d %d %d  %8.1f  %8.1f  %.3f\n", n, aoff, rec, tma, h[0] / (2048.0 * 8), h[1] / (2048.0 * 8), ms);
        }
  int herr;
  cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost);
  printf("errflag %d\n", herr);
  return 0;
}
