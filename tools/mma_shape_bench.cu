// Micro-benchmark behind DESIGN.md's Winograd decision: how many cycles does one tcgen05.mma
// (kind::f16, K = 16, fp32 accumulate) of shape M x N cost when its operands come from shared memory, for the shapes a
// Winograd F(2x2,3x3) tower layer would need?  F(2x2,3x3) keeps SIXTEEN accumulators alive per output tile (one per
// transformed position); TMEM has 512 columns, so each accumulator is at most 512/16 = 32 columns wide -> N = 32 MMAs.
// The direct convolution runs N = 256 (one accumulator, double-buffered).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_mma_shape_bench.bin tools/mma_shape_bench.cu
//   ./tools/_mma_shape_bench.bin          (one line per shape: cycles per MMA, % of the N/2-cycle tensor-pipe floor)
//
// One CTA (or CTA pair) per SM, one elected thread issues ITER x 4 MMAs over a 3-stage ring of operand tiles in the
// SWIZZLE_128B K-major layout (random fp16 contents), rotating over `accs` accumulators; a final tcgen05.commit ->
// mbarrier wait closes the timed region (clock64 on the issuing SM).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {           // K-major, 128-byte rows, SWIZZLE_128B
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.b32 %0, 1, 0, q;\n\t}" : "=r"(p));
  return p;
}

template <int CG>
__global__ void __launch_bounds__(128, 1) k_bench(int N, int accs, int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // fill the operand ring with pseudo-random fp16 values in [-1, 1)
  uint16_t* sm = reinterpret_cast<uint16_t*>(smem_dyn + (base - smem_u32(smem_dyn)));
  const int halfs = 3 * (16384 + 32768) / 2;
  uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = threadIdx.x; i < halfs; i += 128) {
    s = s * 1664525u + 1013904223u;
    sm[i] = __half_as_ushort(__float2half(((int)(s >> 16) - 32768) / 32768.0f));
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy fills -> async-proxy (MMA) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int M = CG == 2 ? 256 : 128;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (warp == 1 && rank == 0) {
    const uint32_t leader = elect_one();
    const int cols = N;                                   // accumulator width in TMEM columns
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t st = (uint32_t)(it % 3);
      const uint64_t a0 = make_desc(base + st * 49152u), b0 = make_desc(base + st * 49152u + 16384u);
      const uint32_t d = tmem + (uint32_t)((it % accs) * cols);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ko = (uint64_t)(k * 32) >> 4;
        const uint32_t acc = (it >= accs || k > 0) ? 1u : 0u;
        if (CG == 1)
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                       "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(a0 + ko), "l"(b0 + ko), "r"(idesc), "r"(acc), "r"(leader) : "memory");
        else
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                       "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(a0 + ko), "l"(b0 + ko), "r"(idesc), "r"(acc), "r"(leader) : "memory");
      }
    }
    if (CG == 1)
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                   ::"r"(smem_u32(&bar)), "r"(leader) : "memory");
    else
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                   ::"r"(smem_u32(&bar)), "r"(leader) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles_out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int CG>
static void run(int N, int accs, int iters, int sms) {
  long long* d;
  CK(cudaMalloc(&d, sizeof(long long) * sms));
  CK(cudaMemset(d, 0, sizeof(long long) * sms));
  const int smem = 3 * 49152 + 1024;
  CK(cudaFuncSetAttribute(k_bench<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms / CG * CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, k_bench<CG>, N, accs, iters, d));          // warm-up
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, k_bench<CG>, N, accs, iters, d));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(sms);
  CK(cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double cyc = 0; int n = 0;
  for (int i = 0; i < sms; i += CG) { cyc += (double)h[i]; ++n; }
  cyc /= n;
  const double per_mma = cyc / (iters * 4.0);
  const int M = CG == 2 ? 256 : 128;
  const double floor_cyc = (double)M * N * 16 / 4096.0 / CG;             // 4096 MAC/clk/SM dense fp16
  const double flops = 2.0 * M * N * 16 * iters * 4.0 * n;
  printf("cta_group::%d  M=%3d N=%3d accumulators=%2d : %7.1f cycles/MMA  (tensor-pipe floor %5.1f -> %5.1f %% of peak rate), %7.1f TFLOP/s chip-wide, kernel %.3f ms\n",
         CG, M, N, accs, per_mma, floor_cyc, 100.0 * floor_cyc / per_mma, flops / (ms * 1e-3) / 1e12, ms);
  CK(cudaFree(d));
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int iters = 20000;
  printf("# tcgen05.mma kind::f16 K=16, operands from shared memory (3-stage ring), %d SMs, %d x 4 MMAs per SM\n", sms, iters);
  for (int N : {256, 128, 64, 32, 16}) run<1>(N, N >= 256 ? 2 : (512 / N > 16 ? 16 : 512 / N), iters, sms);
  for (int N : {256, 128, 64, 32}) run<2>(N, N >= 256 ? 2 : (512 / N > 16 ? 16 : 512 / N), iters, sms);
  return 0;
}
