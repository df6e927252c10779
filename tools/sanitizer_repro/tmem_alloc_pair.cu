// Minimal repro for the one racecheck report class left on the conv kernels (tools/gpu_sanitize.sh):
//   "Race reported between Write access at <kernel>+0xfffffffffffffe80 and Read access at <kernel>+0x.. [N hazards]"
// once per launch of every CTA-pair kernel (k_conv3x3_tc2 / k_conv3x3_tc4), never for the single-CTA kernels.
// Both flagged READ instructions are SYNCS.PHASECHK.TRYWAIT / SYNCS.ARRIVE on a word in the driver-reserved shared
// memory window - the allocation-permit hand-shake between the two CTAs that ptxas emits for the ONE PTX instruction
// tcgen05.alloc.cta_group::2 (cuobjdump -sass: UTCATOMSWS.2CTA.FIND_AND_SET followed by that mbarrier sequence); the
// WRITE has no instruction in the kernel (PC before its first byte): it is the peer CTA's arrival on that word.  No
// user-visible memory is involved and no user-level synchronisation can order it.
// This kernel contains nothing but the documented allocation hand-off (alloc, relinquish, fence, __syncthreads,
// cluster barrier, read of the returned address, dealloc).  Measured on B200 (profiles/r2_sanitize_repro_*.log):
//   cta_group::1, any shape ............................. 0 reports
//   cta_group::2,   2 CTAs x 128 threads, 1 launch ....... 0 reports (the two arrivals never overlap)
//   cta_group::2, 148 CTAs x 384 threads, 3 launches ..... 3 reports, the same signature as on the conv kernels
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tmem_alloc_pair tmem_alloc_pair.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <bool PAIR>
__global__ void __cluster_dims__(2, 1, 1) k_alloc(unsigned* out, int alloc_warp, unsigned cols) {
  __shared__ unsigned slot;
  if ((int)(threadIdx.x >> 5) == alloc_warp) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)), "r"(cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)), "r"(cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = slot;  // <- the flagged read
  if (threadIdx.x == 0) out[blockIdx.x] = base;
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if ((int)(threadIdx.x >> 5) == alloc_warp) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
  }
}
int main(int argc, char** argv) {  // tmem_alloc_pair <p|s> [threads] [alloc warp] [CTAs] [TMEM columns] [launches]
  const bool pair = argc > 1 && argv[1][0] == 'p';
  const int threads = argc > 2 ? atoi(argv[2]) : 128, warp = argc > 3 ? atoi(argv[3]) : 0, grid = argc > 4 ? atoi(argv[4]) : 2;
  const unsigned cols = argc > 5 ? atoi(argv[5]) : 64;
  const int launches = argc > 6 ? atoi(argv[6]) : 1;
  unsigned* d; cudaMalloc(&d, 4 * grid); unsigned h[2] = {9, 9};
  for (int i = 0; i < launches; ++i)
    if (pair) k_alloc<true><<<grid, threads>>>(d, warp, cols); else k_alloc<false><<<grid, threads>>>(d, warp, cols);
  cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s x%d launches, %d CTAs x %d threads, warp %d allocates %u columns: %s, tmem base of CTA 0/1: %u %u\n",
         pair ? "cta_group::2" : "cta_group::1", launches, grid, threads, warp, cols, cudaGetErrorString(e), h[0], h[1]);
  return e != cudaSuccess;
}
