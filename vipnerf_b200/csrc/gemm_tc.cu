// The two tensor-core kernels of the training step (row f1), each in two arithmetic modes with the same byte geometry:
// fp32 arrays read as tf32 (tcgen05 kind::tf32; VIPNERF_FLAG_TRAIN_TF32) or fp16 arrays (kind::f16; VIPNERF_FLAG_TRAIN_F16,
// half the bytes of every stream).
//
// k_gemm_tn_tc<half> - parameter gradients:  C[m][n] = sum_p A[p][m] * B[p][n]  (dW = dY^T X).
// Both operands are the point-major arrays the chain kernels already write (A = pre-activation gradients [P][M],
// B = layer inputs [P][256]); the reduction index p is the OUTER index of both, i.e. both are "MN-major" operands for
// the MMA.  tcgen05.mma reads them from shared memory directly, so no conversion or transposition pass exists: the TMA
// engine drops [32 points] x [128 bytes of columns] boxes into a ring (fp32: 128-byte swizzle with 32-byte atoms - the
// only layout tcgen05 reads MN-major 32-bit operands from; fp16: the plain 128-byte swizzle), one elected thread issues
// M=128, N=256 MMAs (K = 8 / 16 points each, one per 128-row half of the output) that accumulate the CTA's whole
// 256 x 256 (or 128 x 256) fp32 tile in TMEM (all 512 columns), and after the CTA's point range is exhausted four warps
// drain TMEM into a partial tile that k_reduce_partials (train_kernels.cu) sums over the CTAs in a fixed order.
// One CTA per SM, 148 point ranges.
//
//   warp 0      TMA producer (one lane)
//   warp 1      TMEM allocation, MMA issue (one lane), commits that free ring stages
//   warps 2..5  epilogue: column sums of A (the bias gradient) from the staged boxes while the main loop runs, then
//               TMEM -> registers -> global partial tile (warp w owns TMEM lanes 32*(w%4)..)
//
// Roofline: HBM.  Each CTA reads every A and B row of its point range once: (M + 256) elements per point and layer,
// i.e. 2 KiB/point (tf32) or 1 KiB/point (fp16) for the 256 x 256 layers.  Measured (tf32): 256 us for 786,432 points
// (0.97 of the HBM peak); the FFMA kernel it replaces needs 1.7 ms.
//
// Numerics: tf32 operands are rounded (10-bit mantissa) by the TMA copy (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32); fp16
// operands were rounded to the same 10-bit mantissa when they were stored; products accumulate in fp32.  The tensor-core
// training modes are opt-in (configs['model']['train_precision'] = 'tf32' | 'fp16'); the default path stays fp32 FFMA.
// The second kernel of this file, k_linear_tc (below), runs the forward and backward-data chains of both modes.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "kernels.h"

namespace vipnerf {
namespace {

constexpr int kTcRows = 32;                        // points per ring stage
constexpr int kBoxBytes = kTcRows * 128;           // one [32 points][128 bytes = 32 fp32 / 64 fp16] box
// Operand geometry of the two arithmetic modes.  The BYTE geometry is the same (128-byte box rows, 4 KiB boxes); fp16
// operands pack twice the columns into a box and twice the points into an MMA (K = 16), so every stream is half as long.
template <bool kHalf> struct TcGeom {
  static constexpr int kStages = kHalf ? 6 : 3;          // ring stages of k_gemm_tn (192 KiB in flight either way)
  static constexpr int kColsPerBox = kHalf ? 64 : 32;
  static constexpr int kKPerMma = kHalf ? 16 : 8;        // reduction elements per MMA (32 bytes of a K-major row)
};
constexpr int kTcThreads = 192;
constexpr long long kTcTimeoutCycles = 4000000000ll;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // a protocol bug traps instead of hanging
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kTcTimeoutCycles) {
      printf("vipnerf gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool kHalf>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kHalf) umma_f16(d_tmem, a_desc, b_desc, idesc, accumulate);
  else umma_tf32(d_tmem, a_desc, b_desc, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// MN-major shared-memory matrix descriptors (cute::UMMA::SmemDescriptor): start >> 4 in [0,14), LBO >> 4 in [16,30),
// SBO >> 4 in [32,46), version 1 in [46,48), layout type in [61,64).  Both operands of dW = dY^T X have the reduction
// index (the point) as their OUTER index, so the 128-byte box rows run along M / N:
//  * 32-bit operands (tf32) can only be read in the SWIZZLE_128B_BASE32B layout (layout type 1): atoms of [4 k] x
//    [32 fp32 = 128 B] whose 32-byte chunks are XOR-ed with (k % 4) - what the TMA writes in the
//    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B mode; an MMA of K = 8 reads two atoms along K (SBO = 512 B).
//  * 16-bit operands (fp16) use the plain SWIZZLE_128B layout (layout type 2): atoms of [8 k] x [64 fp16 = 128 B] with the
//    16-byte chunks XOR-ed with (k % 8) (CU_TENSOR_MAP_SWIZZLE_128B); an MMA of K = 16 reads two atoms along K
//    (SBO = 1 KiB).
// Consecutive atoms along M / N are one TMA box (4 KiB) apart = the leading-dimension byte offset.
template <bool kHalf>
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kBoxBytes >> 4) << 16) |
         ((uint64_t)((kHalf ? 1024 : 512) >> 4) << 32) | (1ull << 46) | ((kHalf ? 2ull : 1ull) << 61);
}
// Instruction descriptor: D = F32 (c_format 1 at [4,6)), A / B format at [7,10) / [10,13) (TF32 = 2 for kind::tf32,
// F16 = 0 for kind::f16), both operands MN-major (a_major bit 15, b_major bit 16), N >> 3 at [17,23), M >> 4 at [24,29).
template <bool kHalf>
constexpr uint32_t instr_desc_mn(uint32_t n, uint32_t m) {
  return (1u << 4) | ((kHalf ? 0u : 2u) << 7) | ((kHalf ? 0u : 2u) << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// Up to kGemmGroupMax products of the same shape over the same point count share a launch (the eight 256 x 256
// products of a sample set: trunk layers + feature_linear): CTA b works on problem b / splits, point range b % splits.
// With 148 / 8 = 18 CTAs per problem every SM still streams its share of the rows, but a problem leaves 18 partial tiles
// instead of 148 - the partial-tile writes at the end of a launch and the reductions behind it shrink eightfold.
struct GemmTcParams {
  CUtensorMap map_a[kGemmGroupMax], map_b[kGemmGroupMax];
  int n_problems, splits;   // grid = n_problems * splits
  int M;                    // 128 or 256
  int N;                    // tf32: 32, 64 or 256; fp16: 64 or 256 (columns of B)
  int64_t n_rows, rows_per_split;
  float* partial;           // [gridDim.x][M][N]
  float* colsum_partial;    // [gridDim.x][M] column sums of A over the CTA's point range (bias gradient)
  uint32_t colsum_mask;     // bit g: problem g wants them
};

template <bool kHalf>
__global__ void __launch_bounds__(kTcThreads, 1) k_gemm_tn_tc(const __grid_constant__ GemmTcParams p) {
  using G = TcGeom<kHalf>;
  constexpr int kStages = G::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];   // ring stages, then the mbarriers and the TMEM base slot
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int halves = p.M / 128;
  const int a_boxes = p.M / G::kColsPerBox;
  const int b_boxes = p.N / G::kColsPerBox;
  const uint32_t stage_bytes = (uint32_t)(a_boxes + b_boxes) * kBoxBytes;
  uint8_t* tail = smem + kStages * stage_bytes;
  uint32_t* tmem_ptr_slot = reinterpret_cast<uint32_t*>(tail + 8 * (2 * kStages + 1));
  const uint32_t bar0 = smem_u32(tail);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * kStages);

  const int prob = (int)blockIdx.x / p.splits;
  const CUtensorMap* map_a = &p.map_a[prob];
  const CUtensorMap* map_b = &p.map_b[prob];
  const bool want_colsum = (p.colsum_mask >> prob) & 1u;
  const int64_t r_begin = (int64_t)((int)blockIdx.x % p.splits) * p.rows_per_split;
  const int64_t r_end = min(p.n_rows, r_begin + p.rows_per_split);
  const int n_steps = r_end > r_begin ? (int)((r_end - r_begin + kTcRows - 1) / kTcRows) : 0;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) { printf("vipnerf gemm_tc: shared memory base not 1 KiB aligned\n"); __trap(); }
    // a stage is free when its MMAs have retired (one tcgen05.commit arrival) and, with column sums, when the four
    // epilogue warps have read its A boxes (one arrival each)
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), want_colsum ? 5 : 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_slot);

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < n_steps; ++s) {
        const int st = s % kStages;
        if (s >= kStages) mbar_wait(empty_bar(st), ((s / kStages) - 1) & 1);
        mbar_expect_tx(full_bar(st), stage_bytes);
        const uint32_t dst = smem_u32(smem) + (uint32_t)st * stage_bytes;
        const int row = (int)(r_begin + (int64_t)s * kTcRows);   // rows past the end of the arrays are zero-filled by the TMA
        for (int j = 0; j < a_boxes; ++j) tma_load_2d(dst + j * kBoxBytes, map_a, j * G::kColsPerBox, row, full_bar(st));
        for (int j = 0; j < b_boxes; ++j) tma_load_2d(dst + (a_boxes + j) * kBoxBytes, map_b, j * G::kColsPerBox, row, full_bar(st));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc_mn<kHalf>((uint32_t)p.N, 128);
      constexpr int kMmaBytes = G::kKPerMma * 128;            // K points of every box = K 128-byte rows
      for (int s = 0; s < n_steps; ++s) {
        const int st = s % kStages;
        mbar_wait(full_bar(st), (s / kStages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(smem) + (uint32_t)st * stage_bytes;
        const uint32_t b0 = a0 + (uint32_t)a_boxes * kBoxBytes;
#pragma unroll
        for (int k = 0; k < kTcRows / G::kKPerMma; ++k) {     // one MMA per K points and 128-row half of the output
          const uint64_t b_desc = make_desc_mn<kHalf>(b0 + k * kMmaBytes);
          for (int h = 0; h < halves; ++h) {
            const uint64_t a_desc = make_desc_mn<kHalf>(a0 + h * (128 / G::kColsPerBox) * kBoxBytes + k * kMmaBytes);
            umma<kHalf>(tmem_base + h * 256, a_desc, b_desc, idesc, (s > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(st));     // frees the stage when its MMAs have retired
      }
      umma_commit(done_bar);            // all MMAs of this CTA retired -> TMEM holds the tile
    }
  } else {
    // epilogue warps: warp w may touch TMEM lanes [32 * (w % 4), +32)
    float* out = p.partial + (size_t)blockIdx.x * p.M * p.N;
    const int quarter = warp & 3;
    if (want_colsum) {
      // While the main loop runs these 128 threads are idle: they add up the columns of the A boxes of every stage
      // straight from shared memory (db = sum_p dY[p][m], the bias gradient) - the separate column-sum pass over the
      // same array (k_colsum, 1 KiB per point and layer from HBM once more) is gone.
      const int t = (warp - 2) * 32 + lane;
      float acc0 = 0.f, acc1 = 0.f;
      for (int s = 0; s < n_steps; ++s) {
        const int st = s % kStages;
        mbar_wait(full_bar(st), (s / kStages) & 1);
        const uint8_t* a0 = smem + (size_t)st * stage_bytes;
        if constexpr (kHalf) {
          // thread t owns columns 2t, 2t + 1 (one half2); box layout = SWIZZLE_128B: point k at k * 128 B, 16-byte chunk
          // (c / 8) ^ (k % 8).  M = 128: threads 0..63 only.
          if (2 * t < p.M) {
            const int m = 2 * t, c = m & 63;
            const uint8_t* box = a0 + (size_t)(m >> 6) * kBoxBytes + (c & 7) * 2;
#pragma unroll
            for (int k = 0; k < kTcRows; ++k) {
              const float2 v = __half22float2(*reinterpret_cast<const __half2*>(box + k * 128 + (((c >> 3) ^ (k & 7)) << 4)));
              acc0 += v.x; acc1 += v.y;
            }
          }
        } else {
          // thread t owns columns t and t + 128; box layout = SWIZZLE_128B_ATOM_32B: point k at k * 128 B, 32-byte chunk
          // (c / 8) ^ (k % 4).
          for (int h = 0; h < halves; ++h) {
            const int m = t + h * 128, c = m & 31;
            const uint8_t* box = a0 + (size_t)(m >> 5) * kBoxBytes + (c & 7) * 4;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < kTcRows; ++k)
              a += *reinterpret_cast<const float*>(box + k * 128 + ((((c >> 3) ^ (k & 3))) << 5));
            if (h == 0) acc0 += a; else acc1 += a;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(empty_bar(st));
      }
      if constexpr (kHalf) {
        if (2 * t < p.M) {
          p.colsum_partial[(size_t)blockIdx.x * p.M + 2 * t] = acc0;
          p.colsum_partial[(size_t)blockIdx.x * p.M + 2 * t + 1] = acc1;
        }
      } else {
        p.colsum_partial[(size_t)blockIdx.x * p.M + t] = acc0;
        if (halves == 2) p.colsum_partial[(size_t)blockIdx.x * p.M + t + 128] = acc1;
      }
    }
    if (n_steps > 0) {
      mbar_wait(done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    for (int h = 0; h < halves; ++h) {
      const int m = h * 128 + quarter * 32 + lane;
      for (int c = 0; c < p.N / 32; ++c) {
        uint32_t v[32];
        if (n_steps > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + h * 256 + c * 32, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        float4* dst = reinterpret_cast<float4*>(out + (size_t)m * p.N + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                               __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// [n_rows][ld] row-major (fp32 or fp16), first `cols` columns used; box = 128 bytes of columns x 32 rows.
// fp32: 128-byte swizzle with 32-byte atoms, tf32 rounding by the copy; fp16: plain 128-byte swizzle.
cudaError_t encode_rows_map(CUtensorMap* map, const void* base, int ld, int cols, int64_t n_rows, bool half) {
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) return cudaErrorNotSupported;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * (half ? 2 : 4)};
  const cuuint32_t box[2] = {half ? 64u : 32u, (cuuint32_t)kTcRows};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2,
                            const_cast<void*>(base), dims, strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            half ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------------------
// One linear layer of the training chains on the tensor cores:  Y[p][n] = epilogue( sum_k X[p][k] * W[n][k] ).
// Both operands are K-major (k contiguous): X = a point-major activation / gradient array, W = an [N][K] weight matrix
// (nn.Linear's own orientation for the forward, the transposed image for the backward-data chain).  Two arithmetic
// modes with the same byte geometry (TcGeom): fp32 arrays read as tf32 (kind::tf32, 32 k per 128-byte box row) or fp16
// arrays (kind::f16, 64 k per row - half the bytes per element of every stream).
// The TMA engine loads [128 points x 128 B] and [N x 128 B] boxes (plain 128-byte swizzle) into a 3-stage
// ring, one thread issues four MMAs per box pair (M=128 points, N columns, accumulators in N TMEM columns), and
// four epilogue warps drain TMEM row by row: + bias, + a rank-1 term (the density head's contribution to dL/dh8),
// ReLU, ReLU-mask from a saved activation; outputs and masks move as [32 rows x 128 B] boxes through shared memory and
// the TMA engine (bulk tensor stores / loads).  Up to two (X, W) pairs accumulate into the same tile
// (the skip layer's cat([encoding, h4]) input).  Persistent, one CTA per SM: the ring (144 KiB) keeps loads in
// flight across tile boundaries and the two 256-column TMEM accumulators alternate, so the epilogue of one 128-point
// tile overlaps the loads and MMAs of the next.
// fp16 gradient arrays carry a power-of-two scale (grad_scale_from_amax, kernels.h): the epilogue re-centres its output
// on the measured maximum of its INPUT array and records the maximum of what it writes for the next layer.
// Roofline: HBM - K elements read and N written (+ N for a mask) per point; the weights come from L2.
constexpr int kLinStages = 3;
constexpr int kLinRows = 128;
constexpr int kLinStageBytes = 4 * 16384;   // epilogue staging: per epilogue warp 2 output boxes + 2 mask boxes of 4 KiB

struct LinearTcParams {
  CUtensorMap map_a[2], map_b[2];
  CUtensorMap map_out, map_mask;   // [128 B of columns x 32 rows] boxes of the output / the mask array
  int chunks[2];
  int N;
  int64_t n_rows;
  const float* bias;
  const float* rank1_row;
  const float* rank1_col;
  const void* mask;
  const uint32_t* relu_bits;     // fp16 backward: ReLU masks as bits, [rows][8] words (bit c % 32 of word c / 32 = unit c was active)
  uint32_t* relu_bits_out;       // fp16 forward: where kEpiFwdRelu leaves them
  int relu;
  const float* dot_vec;
  float* dot_out;
  const float* dot_bias;
  const float* dot_noise;
  const float* head_wout;
  const float* head_bout;
  float* head_rgb;
  float* head_vis;
  int head_vis_stride;
  const uint32_t* scale_in;
  const uint32_t* scale_out;
  uint32_t* amax_out;
};

// K-major SWIZZLE_128B descriptor: 128-byte rows, 8-row atoms 1 KiB apart (SBO), LBO unused (1), layout type 2
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <bool kHalf>
constexpr uint32_t instr_desc_k(uint32_t n, uint32_t m) {
  return (1u << 4) | ((kHalf ? 0u : 2u) << 7) | ((kHalf ? 0u : 2u) << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// two fp32 -> one packed fp16 pair (lo = a, hi = b), round to nearest, saturating at +-65504 instead of overflowing
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// What the fp16-output epilogue does per element is a compile-time choice (the fp16 step is epilogue-bound as soon as
// its per-tile work is more than ~5 k cycles: every runtime flag removed is instructions and branches saved 256 times
// per row): forward layers add a bias and apply ReLU inside the fp16 conversion; backward layers re-scale, add the
// density head's rank-1 term, apply the ReLU mask on the packed halves and track the maximum on the packed halves; the
// views layer (kEpiFwdHead) also evaluates views_output_linear (128 -> 4) and the sigmoids on the row it holds.
enum : int { kEpiGeneric = 0, kEpiFwdRelu, kEpiFwdLinear, kEpiBwdPlain, kEpiBwdMask, kEpiBwdMaskRank1, kEpiFwdHead };
constexpr int kVecBytes = 3 * 1024;   // bias / rank-1 column / dot vector staged in shared memory (256 floats each)

// two fp32 -> one packed fp16 pair with the ReLU inside the conversion (F2FP.RELU)
__device__ __forceinline__ uint32_t pack_half2_relu_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// kDot: the epilogue also reduces every output row against p.dot_vec (a separate instantiation: the plain layers must not
// pay registers or predicated loads for it).  kHalf: fp16 operands (and masks); kOutHalf: fp16 output, with the
// per-element work selected by kEpi.
template <bool kDot, bool kHalf, bool kOutHalf, int kEpi>
__global__ void __launch_bounds__(kTcThreads, 1) k_linear_tc(const __grid_constant__ LinearTcParams p) {
  using G = TcGeom<kHalf>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = kLinRows * 128, b_bytes = (uint32_t)p.N * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* stage_base = smem + kLinStages * stage_bytes;      // 1 KiB aligned: stage_bytes is a multiple of 16 KiB
  uint8_t* tail = stage_base + kLinStageBytes;
  uint32_t* tmem_ptr_slot = reinterpret_cast<uint32_t*>(tail + 8 * (2 * kLinStages + 4 + 8));
  float* vec_s = reinterpret_cast<float*>(tail + 256);        // [3][256]: bias, rank-1 column, dot vector (fp16-output epilogue)
  const uint32_t bar0 = smem_u32(tail);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kLinStages + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * kLinStages + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * kLinStages + 2 + b); };
  auto mask_full = [&](int w, int b) { return bar0 + 8u * (2 * kLinStages + 4 + 2 * w + b); };
  const int n_chunks = p.chunks[0] + p.chunks[1];
  const int n_tiles = (int)((p.n_rows + kLinRows - 1) / kLinRows);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) { printf("vipnerf linear_tc: shared memory base not 1 KiB aligned\n"); __trap(); }
    for (int s = 0; s < kLinStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 4); }   // one arrival per epilogue warp
    for (int w = 0; w < 4; ++w) { mbar_init(mask_full(w, 0), 1); mbar_init(mask_full(w, 1), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_slot);

  // Persistent: CTA b owns tiles b, b + gridDim.x, ...  The ring runs continuously across tiles (step counter g); the
  // two TMEM accumulators alternate, so the epilogue of tile i overlaps the loads and MMAs of tile i + 1.
  if (warp == 0) {
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kLinRows;
        for (int c = 0; c < n_chunks; ++c, ++g) {
          const int st = g % kLinStages;
          if (g >= kLinStages) mbar_wait(empty_bar(st), ((g / kLinStages) - 1) & 1);
          mbar_expect_tx(full_bar(st), stage_bytes);
          const uint32_t dst = smem_u32(smem) + (uint32_t)st * stage_bytes;
          const int pair = c < p.chunks[0] ? 0 : 1;
          const int kc = (pair ? c - p.chunks[0] : c) * G::kColsPerBox;
          tma_load_2d(dst, &p.map_a[pair], kc, row0, full_bar(st));           // rows past the end are zero-filled
          tma_load_2d(dst + a_bytes, &p.map_b[pair], kc, 0, full_bar(st));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc_k<kHalf>((uint32_t)p.N, 128);
      int g = 0, i = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
        const int buf = i & 1;
        if (i >= 2) {   // the epilogue has drained this accumulator's previous tile
          mbar_wait(acc_empty(buf), ((i >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int c = 0; c < n_chunks; ++c, ++g) {
          const int st = g % kLinStages;
          mbar_wait(full_bar(st), (g / kLinStages) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = smem_u32(smem) + (uint32_t)st * stage_bytes;
          const uint32_t b0 = a0 + a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)   // one MMA per 32 bytes (8 fp32 / 16 fp16) along the 128-byte swizzled rows
            umma<kHalf>(tmem_base + buf * 256, make_desc_k(a0 + k * 32), make_desc_k(b0 + k * 32), idesc, (c > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(st));
        }
        umma_commit(acc_full(buf));
      }
    }
  } else {
    const int quarter = warp & 3;
    int i = 0;
    uint32_t mask_n = 0;    // mask boxes consumed so far by this warp (buffer = mask_n & 1, phase = (mask_n >> 1) & 1)
    // fp16 gradient chain: scale of the input array, scale of the output array (both powers of two)
    const float s_in = p.scale_in ? grad_scale_from_amax(*p.scale_in) : 1.f;
    const float s_out = p.scale_out ? grad_scale_from_amax(*p.scale_out) : 1.f;
    const float fac = s_out / s_in;
    uint32_t amax16 = 0u;   // fp16-output backward epilogues: largest |stored value| of this lane, as two packed fp16 bit patterns
    if constexpr (kOutHalf) {   // the per-column vectors of the epilogue: one copy in shared memory (broadcast reads)
      const int t = (warp - 2) * 32 + lane;
      for (int c = t; c < p.N; c += 128) {
        vec_s[c] = p.bias ? p.bias[c] : 0.f;
        if constexpr (kEpi != kEpiFwdHead) {
          vec_s[256 + c] = p.rank1_col ? p.rank1_col[c] : 0.f;
          vec_s[512 + c] = p.dot_vec ? p.dot_vec[c] : 0.f;
        }
      }
      if constexpr (kEpi == kEpiFwdHead)     // views_output_linear transposed: [128 hidden units][4 outputs]
        for (int c = t; c < 512; c += 128) vec_s[256 + c] = p.head_wout[c];
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
      const int buf = i & 1;
      const int64_t pg = (int64_t)tile * kLinRows + quarter * 32 + lane;
      const bool valid = pg < p.n_rows;
      // fp16 backward: this row's 256 ReLU-mask bits (32 bytes), requested before the wait for the accumulator so that
      // the loads fly while the tile's MMAs finish
      uint4 mw0 = make_uint4(0u, 0u, 0u, 0u), mw1 = mw0;
      if constexpr (kOutHalf && (kEpi == kEpiBwdMask || kEpi == kEpiBwdMaskRank1)) {
        if (valid) {
          mw0 = *reinterpret_cast<const uint4*>(p.relu_bits + pg * 8);
          mw1 = *reinterpret_cast<const uint4*>(p.relu_bits + pg * 8 + 4);
        }
      }
      mbar_wait(acc_full(buf), (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // Epilogue through shared memory and the TMA engine.  A lane owns one accumulator row, so direct global accesses
      // touch 32 different 128-byte lines per instruction (the r01 kernel ran at 0.45 of the HBM roofline because of it).
      // Instead every warp stages its [32 rows x 128 B] chunk in a swizzled 4 KiB box and ONE bulk tensor store
      // writes it (full lines, asynchronous, rows past the end clipped by the tensor map); the tf32 backward chain's ReLU
      // mask arrives the same way (bulk tensor load of the saved activation's box, one chunk ahead; the fp16 chains carry
      // their masks as bits instead).
      uint8_t* wstage = stage_base + (warp - 2) * 16384;             // [2] output boxes, then [2] mask boxes (tf32)
      const uint32_t out_s = smem_u32(wstage), msk_s = out_s + 8192;
      const uint32_t mfull0 = mask_full(warp - 2, 0);
      const int row0 = tile * kLinRows + quarter * 32;
      const bool has_mask = !kOutHalf && p.mask != nullptr;
      if (has_mask && lane == 0) {
        mbar_expect_tx(mfull0 + 8u * (mask_n & 1), 4096);
        tma_load_2d(msk_s + 4096u * (mask_n & 1), &p.map_mask, 0, row0, mfull0 + 8u * (mask_n & 1));
      }
      float dot = 0.f;
      if constexpr (kOutHalf) {
        const int n_sc = p.N / 64;                                    // one box = 64 fp16 columns = two TMEM loads
        constexpr bool kMask = kEpi == kEpiBwdMask || kEpi == kEpiBwdMaskRank1;
        constexpr bool kBwd = kEpi == kEpiBwdPlain || kMask;
        float r1 = 0.f;
        if constexpr (kEpi == kEpiBwdMaskRank1) r1 = valid ? p.rank1_row[pg] * s_out : 0.f;
        float hd[4] = {0.f, 0.f, 0.f, 0.f};     // kEpiFwdHead: the four views_output_linear logits of this row
        const uint32_t tm0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256;
        // eight columns: v[8 qq ..] -> one 16-byte chunk (index q) of the output box
        uint32_t wb0 = 0u, wb1 = 0u;    // mask bits of the current box: columns [0,32) and [32,64) (read backward, written forward)
        auto eight = [&](const uint32_t (&v)[32], int qq, int q, int sc, uint32_t orow) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[8 * qq + j]);
          const int col = sc * 64 + q * 8;
          if constexpr (kEpi == kEpiFwdRelu || kEpi == kEpiFwdLinear || kEpi == kEpiFwdHead) {
            const float4 b0 = *reinterpret_cast<const float4*>(vec_s + col), b1 = *reinterpret_cast<const float4*>(vec_s + col + 4);
            o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
          }
          if constexpr (kBwd) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] *= fac;                 // exact: a power of two
          }
          if constexpr (kEpi == kEpiBwdMaskRank1) {
            const float4 u0 = *reinterpret_cast<const float4*>(vec_s + 256 + col), u1 = *reinterpret_cast<const float4*>(vec_s + 256 + col + 4);
            o[0] = fmaf(r1, u0.x, o[0]); o[1] = fmaf(r1, u0.y, o[1]); o[2] = fmaf(r1, u0.z, o[2]); o[3] = fmaf(r1, u0.w, o[3]);
            o[4] = fmaf(r1, u1.x, o[4]); o[5] = fmaf(r1, u1.y, o[5]); o[6] = fmaf(r1, u1.z, o[6]); o[7] = fmaf(r1, u1.w, o[7]);
          }
          if constexpr (kEpi == kEpiFwdRelu) {   // ReLU mask for the backward chain: one bit per unit
            uint32_t byte = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) byte |= (o[j] > 0.f ? 1u : 0u) << j;
            if (q < 4) wb0 |= byte << (8 * (q & 3)); else wb1 |= byte << (8 * (q & 3));
          }
          if constexpr (kMask) {                 // the unit was active in the forward iff its bit is set
            const uint32_t byte = (q < 4 ? wb0 : wb1) >> (8 * (q & 3));
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (!((byte >> j) & 1u)) o[j] = 0.f;
          }
          if constexpr (kDot) {   // the density head reads the fp32 post-ReLU values
            const float4 d0 = *reinterpret_cast<const float4*>(vec_s + 512 + col), d1 = *reinterpret_cast<const float4*>(vec_s + 512 + col + 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
            dot = fmaf(o[0], d0.x, dot); dot = fmaf(o[1], d0.y, dot); dot = fmaf(o[2], d0.z, dot); dot = fmaf(o[3], d0.w, dot);
            dot = fmaf(o[4], d1.x, dot); dot = fmaf(o[5], d1.y, dot); dot = fmaf(o[6], d1.z, dot); dot = fmaf(o[7], d1.w, dot);
          }
          if constexpr (kEpi == kEpiFwdHead) {   // views_output_linear reads the fp32 post-ReLU values (VipNeRF01.py:580-582)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o[j] = fmaxf(o[j], 0.f);
              const float4 w = *reinterpret_cast<const float4*>(vec_s + 256 + 4 * (col + j));
              hd[0] = fmaf(o[j], w.x, hd[0]); hd[1] = fmaf(o[j], w.y, hd[1]); hd[2] = fmaf(o[j], w.z, hd[2]); hd[3] = fmaf(o[j], w.w, hd[3]);
            }
          }
          uint32_t h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            h[j] = kEpi == kEpiFwdRelu ? pack_half2_relu_sat(o[2 * j], o[2 * j + 1]) : pack_half2_sat(o[2 * j], o[2 * j + 1]);
          const uint32_t sw = (uint32_t)((q ^ (lane & 7)) << 4);      // 128-byte swizzle: 16-byte chunk index ^ (row % 8)
          if constexpr (kBwd) {    // maximum of |stored value|: fp16 bit patterns of non-negative values order like integers
            amax16 = __vimax3_u16x2(amax16, h[0] & 0x7fff7fffu, h[1] & 0x7fff7fffu);
            amax16 = __vimax3_u16x2(amax16, h[2] & 0x7fff7fffu, h[3] & 0x7fff7fffu);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + sw), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        };
        // TMEM loads run one 32-column chunk ahead of the arithmetic (tcgen05.wait::ld waits for ALL of a thread's loads,
        // so the next load is issued right after the wait and lands while the current chunk is processed)
        uint32_t va[32], vb[32];
        tmem_ld32(tm0, va);
#pragma unroll 1
        for (int sc = 0; sc < n_sc; ++sc) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tmem_ld32(tm0 + sc * 64 + 32, vb);
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store issued two boxes ago has read its box
          __syncwarp();
          if constexpr (kMask) {          // words 2 sc, 2 sc + 1 of the row's mask
            wb0 = sc == 0 ? mw0.x : (sc == 1 ? mw0.z : (sc == 2 ? mw1.x : mw1.z));
            wb1 = sc == 0 ? mw0.y : (sc == 1 ? mw0.w : (sc == 2 ? mw1.y : mw1.w));
          }
          if constexpr (kEpi == kEpiFwdRelu) wb0 = wb1 = 0u;
          const uint32_t orow = out_s + 4096u * (sc & 1) + lane * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) eight(va, q, q, sc, orow);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (sc + 1 < n_sc) tmem_ld32(tm0 + (sc + 1) * 64, va);
#pragma unroll
          for (int q = 0; q < 4; ++q) eight(vb, q, q + 4, sc, orow);
          if constexpr (kEpi == kEpiFwdRelu) {
            if (valid && p.relu_bits_out != nullptr) *reinterpret_cast<uint2*>(p.relu_bits_out + pg * 8 + 2 * sc) = make_uint2(wb0, wb1);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&p.map_out), "r"(sc * 64), "r"(row0), "r"(out_s + 4096u * (sc & 1)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if constexpr (kEpi == kEpiFwdHead) {
          if (valid) {
            const float4 bo = *reinterpret_cast<const float4*>(p.head_bout);
            if (p.head_rgb != nullptr) {
              p.head_rgb[3 * pg + 0] = 1.f / (1.f + expf(-(hd[0] + bo.x)));
              p.head_rgb[3 * pg + 1] = 1.f / (1.f + expf(-(hd[1] + bo.y)));
              p.head_rgb[3 * pg + 2] = 1.f / (1.f + expf(-(hd[2] + bo.z)));
            }
            p.head_vis[pg * p.head_vis_stride] = 1.f / (1.f + expf(-(hd[3] + bo.w)));
          }
        }
      } else {
        const int n_ch = p.N / 32;
        const float r1 = (p.rank1_row != nullptr && valid) ? p.rank1_row[pg] : 0.f;
        for (int c = 0; c < n_ch; ++c, ++mask_n) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256 + c * 32, v);
          if (has_mask && lane == 0 && c + 1 < n_ch) {                  // next chunk's mask box (its buffer was read at c - 1)
            mbar_expect_tx(mfull0 + 8u * ((mask_n + 1) & 1), 4096);
            tma_load_2d(msk_s + 4096u * ((mask_n + 1) & 1), &p.map_mask, (c + 1) * 32, row0, mfull0 + 8u * ((mask_n + 1) & 1));
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store issued two chunks ago has read its box
          __syncwarp();
          if (has_mask) mbar_wait(mfull0 + 8u * (mask_n & 1), (mask_n >> 1) & 1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const uint32_t orow = out_s + 4096u * (c & 1) + lane * 128, mrow = msk_s + 4096u * (mask_n & 1) + lane * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                   __uint_as_float(v[4 * q + 3]));
            if (p.bias != nullptr) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + c * 32 + 4 * q);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (p.rank1_row != nullptr) {
              const float4 u = *reinterpret_cast<const float4*>(p.rank1_col + c * 32 + 4 * q);
              o.x = fmaf(r1, u.x, o.x); o.y = fmaf(r1, u.y, o.y); o.z = fmaf(r1, u.z, o.z); o.w = fmaf(r1, u.w, o.w);
            }
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (kDot) {
              const float4 dv = *reinterpret_cast<const float4*>(p.dot_vec + c * 32 + 4 * q);
              dot = fmaf(o.x, dv.x, dot); dot = fmaf(o.y, dv.y, dot); dot = fmaf(o.z, dv.z, dot); dot = fmaf(o.w, dv.w, dot);
            }
            const uint32_t sw = (uint32_t)((q ^ (lane & 7)) << 4);    // 128-byte swizzle: 16-byte chunk index ^ (row % 8)
            if (has_mask) {
              float4 m;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(m.x), "=f"(m.y), "=f"(m.z), "=f"(m.w) : "r"(mrow + sw));
              o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f; o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(orow + sw), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&p.map_out), "r"(c * 32), "r"(row0), "r"(out_s + 4096u * (c & 1)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (kDot && valid) {
        if (p.dot_bias != nullptr) {   // the density head finished here: + bias, + noise, ReLU (VipNeRF01.py:546-553)
          float pre = dot + *p.dot_bias;
          if (p.dot_noise != nullptr) pre = pre + p.dot_noise[pg];
          dot = fmaxf(pre, 0.f);
        }
        p.dot_out[pg] = dot;
      }
      // this warp's TMEM reads of the accumulator are complete: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
    }
    if (kOutHalf && p.amax_out != nullptr) {   // largest |value| this warp wrote, un-scaled; float bits of non-negative values order like integers
      const uint32_t m16 = __reduce_max_sync(0xffffffffu, max(amax16 & 0xffffu, amax16 >> 16));
      const float mv = __half2float(__ushort_as_half((unsigned short)m16)) / s_out;
      if (lane == 0 && m16 != 0u) atomicMax(p.amax_out, __float_as_uint(mv));
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory outlives the last stores
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// [rows][ld] row-major, first `cols` columns; box = 128 bytes of columns x box_rows rows, plain 128-byte swizzle.
// kind 0: fp32 read as tf32 (rounded by the copy), 1: exact fp32 (the epilogue's output / mask boxes), 2: fp16
cudaError_t encode_kmajor_map(CUtensorMap* map, const void* base, int ld, int cols, int64_t rows, int box_rows, int kind) {
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) return cudaErrorNotSupported;
  const bool half = kind == 2;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * (half ? 2 : 4)};
  const cuuint32_t box[2] = {half ? 64u : 32u, (cuuint32_t)box_rows};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUtensorMapDataType dt = half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                      : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32);
  const CUresult r = encode(map, dt, 2, const_cast<void*>(base), dims, strides, box,
                            elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_linear_tc(const LinearTcArgs& a, cudaStream_t s) {
  if (a.n_rows < 1) return cudaSuccess;
  const int kq = a.half_in ? 64 : 32;     // reduction elements per 128-byte box row
  if ((a.N != 128 && a.N != 256) || a.k[0] < kq || (a.k[0] % kq) || (a.k[1] % kq) || a.k[1] < 0) return cudaErrorInvalidValue;
  if (a.half_out && !a.half_in) return cudaErrorInvalidValue;       // fp16 outputs exist in the fp16 mode only
  if (a.mask != nullptr && a.half_in) return cudaErrorInvalidValue;    // fp16 chains take their ReLU masks as bits (relu_bits)
  const int es_in = a.half_in ? 2 : 4, es_out = a.half_out ? 2 : 4;
  LinearTcParams p{};
  cudaError_t e;
  for (int i = 0; i < 2; ++i) {
    p.chunks[i] = a.k[i] / kq;
    if (a.k[i] == 0) continue;
    if ((reinterpret_cast<uintptr_t>(a.x[i]) & 15u) || (reinterpret_cast<uintptr_t>(a.w[i]) & 15u) || ((a.ldx[i] * es_in) & 15) || ((a.ldw[i] * es_in) & 15))
      return cudaErrorInvalidValue;
    if ((e = encode_kmajor_map(&p.map_a[i], a.x[i], a.ldx[i], a.k[i], a.n_rows, kLinRows, a.half_in ? 2 : 0)) != cudaSuccess) return e;
    if ((e = encode_kmajor_map(&p.map_b[i], a.w[i], a.ldw[i], a.k[i], a.N, a.N, a.half_in ? 2 : 0)) != cudaSuccess) return e;
  }
  if (a.k[1] == 0) { p.map_a[1] = p.map_a[0]; p.map_b[1] = p.map_b[0]; }
  p.N = a.N; p.n_rows = a.n_rows; p.bias = a.bias; p.rank1_row = a.rank1_row; p.rank1_col = a.rank1_col;
  p.mask = a.mask; p.relu = a.relu ? 1 : 0;
  p.relu_bits = a.relu_bits; p.relu_bits_out = a.relu_bits_out;
  if ((reinterpret_cast<uintptr_t>(a.relu_bits) & 15u) || (reinterpret_cast<uintptr_t>(a.relu_bits_out) & 7u)) return cudaErrorInvalidValue;
  if ((a.relu_bits || a.relu_bits_out) && (!a.half_out || a.N != 256)) return cudaErrorInvalidValue;
  p.dot_vec = a.dot_vec; p.dot_out = a.dot_vec ? a.dot_out : nullptr;
  p.dot_bias = a.dot_bias; p.dot_noise = a.dot_noise;
  p.head_wout = a.head_wout; p.head_bout = a.head_bout; p.head_rgb = a.head_rgb; p.head_vis = a.head_vis;
  p.head_vis_stride = a.head_vis_stride;
  p.scale_in = a.scale_in; p.scale_out = a.scale_out; p.amax_out = a.amax_out;
  if ((reinterpret_cast<uintptr_t>(a.out) & 15u) || ((a.ld_out * es_out) & 15) ||
      (a.mask && ((reinterpret_cast<uintptr_t>(a.mask) & 15u) || ((a.ld_mask * es_out) & 15))))
    return cudaErrorInvalidValue;
  if ((e = encode_kmajor_map(&p.map_out, a.out, a.ld_out, a.N, a.n_rows, 32, a.half_out ? 2 : 1)) != cudaSuccess) return e;
  if (a.mask != nullptr) {
    if ((e = encode_kmajor_map(&p.map_mask, a.mask, a.ld_mask, a.N, a.n_rows, 32, a.half_out ? 2 : 1)) != cudaSuccess) return e;
  } else {
    p.map_mask = p.map_out;
  }
  const size_t smem = (size_t)kLinStages * (kLinRows * 128 + a.N * 128) + kLinStageBytes + 256 + kVecBytes;
  const bool dot = a.dot_vec != nullptr && a.dot_out != nullptr;
  int dev = 0, sms = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  const int64_t n_tiles = (a.n_rows + kLinRows - 1) / kLinRows;
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
#define VIPNERF_LAUNCH_LINEAR(DOT, HALF, OUTHALF, EPI)                                                                    \
  do {                                                                                                                  \
    if ((e = cudaFuncSetAttribute(k_linear_tc<DOT, HALF, OUTHALF, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e; \
    k_linear_tc<DOT, HALF, OUTHALF, EPI><<<grid, kTcThreads, smem, s>>>(p);                                             \
  } while (0)
  if (!a.half_in) {
    if (dot) VIPNERF_LAUNCH_LINEAR(true, false, false, kEpiGeneric); else VIPNERF_LAUNCH_LINEAR(false, false, false, kEpiGeneric);
  } else if (!a.half_out) {
    if (dot) return cudaErrorInvalidValue;
    VIPNERF_LAUNCH_LINEAR(false, true, false, kEpiGeneric);
  } else {
    // the fp16-output epilogue is specialised at compile time: the combinations the training chains use
    const bool scaled = a.scale_in != nullptr || a.scale_out != nullptr;
    const bool rank1 = a.rank1_row != nullptr && a.rank1_col != nullptr;
    const bool head = a.head_wout != nullptr;
    if (head && (a.N != 128 || !a.head_bout || !a.head_vis || dot || !a.relu || (reinterpret_cast<uintptr_t>(a.head_bout) & 15u)))
      return cudaErrorInvalidValue;
    if (!scaled && a.bias != nullptr && !rank1 && a.mask == nullptr) {            // forward layers
      if (head) VIPNERF_LAUNCH_LINEAR(false, true, true, kEpiFwdHead);
      else if (a.relu && dot) VIPNERF_LAUNCH_LINEAR(true, true, true, kEpiFwdRelu);
      else if (a.relu) VIPNERF_LAUNCH_LINEAR(false, true, true, kEpiFwdRelu);
      else if (!dot) VIPNERF_LAUNCH_LINEAR(false, true, true, kEpiFwdLinear);
      else return cudaErrorInvalidValue;
    } else if (scaled && a.bias == nullptr && !a.relu && !dot) {                  // backward-data layers
      if (a.relu_bits != nullptr && rank1) VIPNERF_LAUNCH_LINEAR(false, true, true, kEpiBwdMaskRank1);
      else if (a.relu_bits != nullptr) VIPNERF_LAUNCH_LINEAR(false, true, true, kEpiBwdMask);
      else if (!rank1) VIPNERF_LAUNCH_LINEAR(false, true, true, kEpiBwdPlain);
      else return cudaErrorInvalidValue;
    } else {
      return cudaErrorInvalidValue;
    }
  }
#undef VIPNERF_LAUNCH_LINEAR
  return cudaGetLastError();
}

size_t gemm_tn_tc_partial_floats(int sms) { return (size_t)sms * 256 * 256; }

cudaError_t launch_gemm_tn_tc_group(const GemmProblem* problems, int n_problems, int M, int N, int64_t n_rows,
                                    float* partial, float* colsum_scratch, bool half, cudaStream_t s) {
  if (n_problems < 1 || n_problems > kGemmGroupMax || (M != 128 && M != 256) || n_rows < 1) return cudaErrorInvalidValue;
  if (half ? (N != 64 && N != 256) : (N != 32 && N != 64 && N != 256)) return cudaErrorInvalidValue;
  const int es = half ? 2 : 4;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  // the split-reduction scratch is sized for kGemmTcMaxSplits partial tiles (api.cu: gemm_tn_tc_partial_floats(160))
  int64_t n_split = (sms < kGemmTcMaxSplits ? sms : kGemmTcMaxSplits) / n_problems;      // point ranges per problem
  const int64_t max_by_rows = (n_rows + 4 * kTcRows - 1) / (4 * kTcRows);
  if (n_split > max_by_rows) n_split = max_by_rows;
  if (n_split < 1) n_split = 1;
  int64_t rows_per_split = (n_rows + n_split - 1) / n_split;
  rows_per_split = (rows_per_split + kTcRows - 1) / kTcRows * kTcRows;
  n_split = (n_rows + rows_per_split - 1) / rows_per_split;

  GemmTcParams p{};
  for (int g = 0; g < n_problems; ++g) {
    const GemmProblem& q = problems[g];
    if ((reinterpret_cast<uintptr_t>(q.A) & 15u) || (reinterpret_cast<uintptr_t>(q.B) & 15u) || ((q.lda * es) & 15) || ((q.ldb * es) & 15))
      return cudaErrorInvalidValue;
    if (q.bias_dst != nullptr && colsum_scratch == nullptr) return cudaErrorInvalidValue;
    if ((e = encode_rows_map(&p.map_a[g], q.A, q.lda, M, n_rows, half)) != cudaSuccess) return e;
    if ((e = encode_rows_map(&p.map_b[g], q.B, q.ldb, N, n_rows, half)) != cudaSuccess) return e;
    if (q.bias_dst != nullptr) p.colsum_mask |= 1u << g;
  }
  p.n_problems = n_problems; p.splits = (int)n_split;
  p.M = M; p.N = N; p.n_rows = n_rows; p.rows_per_split = rows_per_split; p.partial = partial;
  p.colsum_partial = colsum_scratch;
  const int cpb = half ? 64 : 32;
  const size_t smem = (size_t)(half ? TcGeom<true>::kStages : TcGeom<false>::kStages) * (M / cpb + N / cpb) * kBoxBytes + 128;
  const unsigned grid = (unsigned)(n_problems * n_split);
  if (half) {
    if ((e = cudaFuncSetAttribute(k_gemm_tn_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k_gemm_tn_tc<true><<<grid, kTcThreads, smem, s>>>(p);
  } else {
    if ((e = cudaFuncSetAttribute(k_gemm_tn_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k_gemm_tn_tc<false><<<grid, kTcThreads, smem, s>>>(p);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ReduceJob jobs[2 * kGemmGroupMax];      // weights and bias reductions of all problems: one launch
  int n_jobs = 0;
  for (int g = 0; g < n_problems; ++g) {
    const GemmProblem& q = problems[g];
    jobs[n_jobs++] = ReduceJob{partial + (size_t)g * n_split * M * N, (int)n_split, M, N, q.dst, q.ldc, q.n_valid, q.scale_def};
    if (q.bias_dst != nullptr)
      jobs[n_jobs++] = ReduceJob{colsum_scratch + (size_t)g * n_split * M, (int)n_split, M, 1, q.bias_dst, 1, 1, q.scale_def};
  }
  return launch_reduce_jobs(jobs, n_jobs, s);
}

cudaError_t launch_gemm_tn_tc(const void* A, int lda, int M, const void* B, int ldb, int N, int64_t n_rows, float* dst,
                              int ldc, int n_valid, float* partial, cudaStream_t s, float* bias_dst,
                              float* colsum_scratch, bool half, const uint32_t* scale_def) {
  const GemmProblem q{A, lda, B, ldb, dst, ldc, n_valid, bias_dst, scale_def};
  return launch_gemm_tn_tc_group(&q, 1, M, N, n_rows, partial, colsum_scratch, half, s);
}

}  // namespace vipnerf
