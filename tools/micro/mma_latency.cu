// Micro-benchmark: latency from issuing n tcgen05.mma (M=128, N=256, K=16, bf16, both operands in shared memory)
// + tcgen05.commit until the mbarrier completes, seen by the issuing thread.  Design aid; not part of the product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
constexpr uint32_t instr_desc(uint32_t n, uint32_t m) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) k_lat(int n_mma, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mem;
  __shared__ uint32_t tmem_ptr;
  const uint32_t bar = smem_u32(&bar_mem);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem) + 32768);
    const uint32_t idesc = instr_desc(256, 128);
    long long total = 0, total_issue = 0;
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                     "l"(a + 2 * (i & 3)), "l"(b + 2 * (i & 3)), "r"(idesc), "r"(i) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      const long long t1 = clock64();
      while (!try_wait(bar, parity)) {}
      const long long t2 = clock64();
      parity ^= 1;
      if (r > 0) { total += t2 - t0; total_issue += t1 - t0; }
    }
    out[0] = total / (reps - 1);
    out[1] = total_issue / (reps - 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  printf("n_mma | issue->barrier cycles | issue loop cycles\n");
  for (int n : {0, 1, 2, 4, 8, 16, 32, 64}) {
    k_lat<<<1, 128, 96 * 1024>>>(n, 20, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%5d | %8lld | %8lld\n", n, h[0], h[1]);
  }
  return 0;
}
