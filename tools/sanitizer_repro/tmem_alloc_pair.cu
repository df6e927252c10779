// Minimal repro for the one racecheck report class left on the conv kernels (tools/gpu_sanitize.sh):
//   "Race reported between Write access at <kernel>+0xffff...fe80 and Read access at ... (*tmem_slot)".
// The write is the tcgen05.alloc result landing in shared memory; the read happens after
// tcgen05.fence::before_thread_sync + __syncthreads() + barrier.cluster arrive.release / wait.acquire +
// tcgen05.fence::after_thread_sync - the allocation hand-off the PTX ISA prescribes.  PAIR = true (cta_group::2 in a
// 2-CTA cluster) is the sequence of k_conv3x3_tc2 / k_conv3x3_tc4 (conv_tc.cu); PAIR = false (cta_group::1) is
// k_conv3x3_tc's, which racecheck does not flag.  Expected: the tool reports hazards for PAIR = true only, with a
// write PC outside the kernel's code (negative offset): it attributes the peer-visible alloc write to no instruction
// and does not order it by the cluster barrier.
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
