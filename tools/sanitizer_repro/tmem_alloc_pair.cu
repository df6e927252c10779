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
#include <cuda_runtime.h>
template <bool PAIR>
__global__ void __cluster_dims__(2, 1, 1) k_alloc(unsigned* out) {
  __shared__ unsigned slot;
  if (threadIdx.x < 32) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = slot;  // <- the flagged read
  if (threadIdx.x == 64) out[blockIdx.x] = base;
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
  }
}
int main(int argc, char** argv) {
  unsigned* d; cudaMalloc(&d, 64); unsigned h[4] = {9, 9, 9, 9};
  const bool pair = argc > 1 && argv[1][0] == 'p';
  if (pair) k_alloc<true><<<2, 128>>>(d); else k_alloc<false><<<2, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s: %s, tmem base per CTA: %u %u\n", pair ? "cta_group::2" : "cta_group::1", cudaGetErrorString(e), h[0], h[1]);
  return e != cudaSuccess;
}
