// Minimal reproducer for the only report compute-sanitizer --tool racecheck makes on libpodb200 (profiles/r1e_sanitizer_racecheck.txt):
// a "race" between the shared-memory write of tcgen05.alloc.cta_group::2 and ... tcgen05.alloc.cta_group::2 itself.
// This kernel contains NOTHING but the canonical allocation handshake of a CTA pair
//     warp 1 (both CTAs, converged):  tcgen05.alloc.cta_group::2 [smem], 512 ; tcgen05.relinquish_alloc_permit.cta_group::2
//     all threads:                    tcgen05.fence::before_thread_sync ; __syncthreads ; barrier.cluster ; tcgen05.fence::after_thread_sync
//     all threads:                    read the TMEM base address from smem
//     ...                             tcgen05.dealloc.cta_group::2
// If racecheck reports the same hazard here, the report is about the instruction (its 32 converged lanes name the same
// destination word; the pair's allocator handshake is invisible to the tool), not about conv_tc.cu's use of it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/_racecheck_tmem_alloc2.bin tools/racecheck_tmem_alloc2.cu
//   compute-sanitizer --tool racecheck ./tools/_racecheck_tmem_alloc2.bin
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_alloc2(uint32_t* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s;
  if (threadIdx.x == 0) out[blockIdx.x] = base;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
  }
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 8 * sizeof(uint32_t));
  k_alloc2<<<4, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  uint32_t h[4] = {9, 9, 9, 9};
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s; TMEM base per CTA: %u %u %u %u\n", cudaGetErrorString(e), h[0], h[1], h[2], h[3]);
  return e == cudaSuccess ? 0 : 1;
}
