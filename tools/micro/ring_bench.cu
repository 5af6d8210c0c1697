// Micro-benchmark: how fast can every SM stream the same L2-resident buffer into a shared-memory ring with
// cp.async.bulk, as a function of chunk size, ring depth and consumer hold time?  (Design aid for the weight ring of
// k_render_tc; not part of the product.)   nvcc -arch=sm_100a -O3 -o ring_bench ring_bench.cu && ./ring_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// warps 0..3 lane 0: producers (chunk i is handled by producer i % 4); warps 4..7 lane 0: consumers (likewise), so
// that the per-chunk handshake overhead of a single thread (~300 cycles of dependent mbarrier / clock latencies) does
// not hide the memory system.  A consumer holds its chunk `hold` cycles, then frees the stage.
__global__ void __launch_bounds__(256, 1) k_ring(const uint8_t* src, uint32_t stream_bytes, uint32_t chunk_bytes,
                                                int stages, int n_chunks, int hold, int lag, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 32);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8 * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty0 + 8 * s));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_begin = clock64();
  const uint32_t n_stream_chunks = stream_bytes / chunk_bytes;
  const uint32_t first = (uint32_t)((blockIdx.x * (unsigned long long)lag) / chunk_bytes);
  if (warp < 4 && lane == 0) {
    for (int i = warp; i < n_chunks; i += 4) {
      const int s = i % stages;
      const uint32_t par = ((i / stages) & 1) ^ 1;
      while (!try_wait(empty0 + 8 * s, par)) {}
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * s), "r"(chunk_bytes) : "memory");
      const uint32_t off = ((first + i) % n_stream_chunks) * chunk_bytes;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem) + s * chunk_bytes), "l"(src + off), "r"(chunk_bytes), "r"(full0 + 8 * s) : "memory");
    }
  } else if (warp >= 4 && lane == 0) {
    for (int i = warp - 4; i < n_chunks; i += 4) {
      const int s = i % stages;
      const uint32_t par = (i / stages) & 1;
      while (!try_wait(full0 + 8 * s, par)) {}
      if (hold > 0) { const long long t = clock64(); while (clock64() - t < hold) {} }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
    }
    out[8 * blockIdx.x + warp - 4] = clock64() - t_begin;
  }
}

int main() {
  const uint32_t stream_bytes = 2 * 1294336 / 16384 * 16384;
  uint8_t* src;
  cudaMalloc(&src, stream_bytes);
  cudaMemset(src, 1, stream_bytes);
  unsigned long long* out;
  cudaMalloc(&out, 148 * 8 * 8);
  cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  const int n_chunks = 4000;
  printf("chunk_B stages hold(x4 consumers) lag grid | cycles/chunk  B/clk/SM  chip_B/clk\n");
  for (int grid : {148, 74, 16}) {
    for (int lag : {0, 16384 * 7}) {
      for (uint32_t chunk : {4096u, 8192u, 16384u}) {
        for (int stages : {4, 8, 12, 16, 24}) {
          if ((size_t)chunk * stages > 200 * 1024) continue;
          for (int hold : {0, 1024}) {
            k_ring<<<grid, 256, 201 * 1024>>>(src, stream_bytes, chunk, stages, n_chunks, hold, lag, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<unsigned long long> h(grid * 8);
            cudaMemcpy(h.data(), out, grid * 8 * 8, cudaMemcpyDeviceToHost);
            double cyc = 0;
            for (int b = 0; b < grid; ++b) cyc += (double)h[8 * b];
            cyc /= (double)grid * n_chunks;
            printf("%6u %4d %4d %7d %4d | %8.1f %8.1f %9.0f\n", chunk, stages, hold, lag, grid, cyc, chunk / cyc, grid * chunk / cyc);
          }
        }
      }
    }
  }
  return 0;
}
