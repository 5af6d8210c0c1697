// CUDA-core (fp32 FFMA) evaluation of the radiance/visibility MLP on sample points.
// This is the VIPNERF_PRECISION_FP32 path: it follows the reference's arithmetic as closely as a GPU can
// (fp32 operands, fp32 accumulate, accurate sincosf/expf) and is the strict-parity arm of the test-suite and
// the only path that evaluates the per-sample secondary-view visibility (visibility2).  The throughput arm
// is mlp_tc.cu.
//
// Reference: MLP.forward VipNeRF01.py:509-535, get_view_independent_outputs :537-566,
// get_view_dependent_outputs :568-596, PositionalEncoder :416-448, run_network :264-293.
//
// One block = 64 consecutive sample points, 128 threads.  Activations live in shared memory k-major
// (act[k][point]) so that a thread's 8 points are two 128-bit loads; the transposed weights stream through a
// double-buffered 32-row shared window with cp.async; each thread owns an 8-point x 16-column register tile.
#include <cuda_runtime.h>

#include "kernels.h"
#include "layout.cuh"

namespace vipnerf {
namespace {

constexpr int kPts = 64;
constexpr int kThreads = 128;
constexpr int kKC = 32;  // weight rows per shared window

constexpr int kActFloats = 320 * kPts;
constexpr int kWbufFloats = 2 * kKC * 256;
constexpr int kPevFloats = kEncView * kPts;
constexpr int kDirFloats = 3 * kPts;
constexpr size_t kSmemBytes = (kActFloats + kWbufFloats + kPevFloats + kDirFloats) * sizeof(float);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[8][4*NJ] += act[a0 + k][8 points] (x) Wt[k][16 or 8 columns], k over K rows streamed through wbuf.
template <int NJ>
__device__ __forceinline__ void matmul_layer(const float* __restrict__ wt, int K, const float* act_rows, float* wbuf,
                                             int tp, int tn, float (&acc)[8][4 * NJ]) {
  constexpr int N = 64 * NJ;
  constexpr int kVecPerThread = (kKC * N / 4) / kThreads;
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * NJ; ++j) acc[i][j] = 0.f;
  const int n_chunks = K / kKC;
  auto prefetch = [&](int c) {
    const float* src = wt + (size_t)c * kKC * N;
    float* dst = wbuf + (c & 1) * (kKC * 256);
#pragma unroll
    for (int v = 0; v < kVecPerThread; ++v) {
      const int e = (v * kThreads + tid) * 4;
      cp_async16(dst + e, src + e);
    }
    cp_async_commit();
  };
  prefetch(0);
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) { prefetch(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * (kKC * 256);
    const float* ar = act_rows + (size_t)c * kKC * kPts + tp * 8;
#pragma unroll 4
    for (int kk = 0; kk < kKC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(ar + kk * kPts);
      const float4 a1 = *reinterpret_cast<const float4*>(ar + kk * kPts + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(wb + kk * N + j * 64 + tn * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][4 * j + 0] = fmaf(a[i], w.x, acc[i][4 * j + 0]);
          acc[i][4 * j + 1] = fmaf(a[i], w.y, acc[i][4 * j + 1]);
          acc[i][4 * j + 2] = fmaf(a[i], w.z, acc[i][4 * j + 2]);
          acc[i][4 * j + 3] = fmaf(a[i], w.w, acc[i][4 * j + 3]);
        }
      }
    }
    __syncthreads();  // window (c&1) is free for chunk c+2; on the last chunk: all reads of act are done
  }
}

__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// views_linears.0 ReLU + views_output_linear for this thread's 8 points x 8 columns, reduced over the 16
// threads that share the points: o[i][0..2] = rgb logits, o[i][3] = visibility logit (bias not yet added).
__device__ __forceinline__ void view_head(const float (&acc9)[8][8], const float* __restrict__ bv,
                                          const float* __restrict__ wvd, const float* __restrict__ wout,
                                          const float* pev, int tp, int tn, float (&o)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) o[i][k] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int n0 = j * 64 + tn * 4;
    float pre[8][4];
    const float4 b4 = *reinterpret_cast<const float4*>(bv + n0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pre[i][0] = acc9[i][4 * j + 0] + b4.x;
      pre[i][1] = acc9[i][4 * j + 1] + b4.y;
      pre[i][2] = acc9[i][4 * j + 2] + b4.z;
      pre[i][3] = acc9[i][4 * j + 3] + b4.w;
    }
#pragma unroll 1
    for (int e = 0; e < kEncView; ++e) {
      const float4 w = *reinterpret_cast<const float4*>(wvd + e * 128 + n0);
      const float4 p0 = *reinterpret_cast<const float4*>(pev + e * kPts + tp * 8);
      const float4 p1 = *reinterpret_cast<const float4*>(pev + e * kPts + tp * 8 + 4);
      const float pe[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        pre[i][0] = fmaf(pe[i], w.x, pre[i][0]);
        pre[i][1] = fmaf(pe[i], w.y, pre[i][1]);
        pre[i][2] = fmaf(pe[i], w.z, pre[i][2]);
        pre[i][3] = fmaf(pe[i], w.w, pre[i][3]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 wo = *reinterpret_cast<const float4*>(wout + (n0 + q) * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = fmaxf(pre[i][q], 0.f);
        o[i][0] = fmaf(h, wo.x, o[i][0]);
        o[i][1] = fmaf(h, wo.y, o[i][1]);
        o[i][2] = fmaf(h, wo.z, o[i][2]);
        o[i][3] = fmaf(h, wo.w, o[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) o[i][k] = half_warp_sum(o[i][k]);
}

__global__ void __launch_bounds__(kThreads, 1)
k_mlp_fp32(RayPtrs rp, RenderFlags fl, int64_t n_points, int S, const float* __restrict__ z,
           const float* __restrict__ small, const float* __restrict__ big, float* __restrict__ out_sigma,
           float* __restrict__ out_rgb, float* __restrict__ out_vis, float* __restrict__ out_vis2) {
  extern __shared__ __align__(16) float smem[];
  float* act = smem;                    // [320][64]: rows 0..63 encoding (row 63 = 0), rows 64..319 hidden
  float* wbuf = act + kActFloats;       // [2][32][256]
  float* pev = wbuf + kWbufFloats;      // [27][64] view-direction encoding per point
  float* dir2 = pev + kPevFloats;       // [3][64]
  const int tid = threadIdx.x;
  const int tp = tid >> 4, tn = tid & 15;
  const int64_t p0 = (int64_t)blockIdx.x * kPts;

  // ---- sample points and encodings: 64 points x 3 axes = 192 tasks
  for (int t = tid; t < 3 * kPts; t += kThreads) {
    const int p = t % kPts, axis = t / kPts;
    const int64_t pg = min(p0 + p, n_points - 1);
    const int64_t ray = pg / S;
    const float x = fadd(rp.pts_o[3 * ray + axis], fmul(rp.pts_d[3 * ray + axis], z[pg]));  // :105-107
    encode_axis<kLPts>(x, axis, [&](int col, float v) { act[col * kPts + p] = v; });
    encode_axis<kLView>(rp.view_dirs[3 * ray + axis], axis, [&](int col, float v) { pev[col * kPts + p] = v; });
  }
  if (tid < kPts) act[63 * kPts + tid] = 0.f;
  __syncthreads();

  float sigma_partial[8];
  // ---- M0 .. M8: 256-wide layers
  {
    float acc[8][16];
    for (int l = 0; l < 9; ++l) {
      const int K = layer_k(l);
      const float* rows = act + ((l == 0 || l == 5) ? 0 : 64 * kPts);
      matmul_layer<4>(big + fp32_layer_offset(l), K, rows, wbuf, tp, tn, acc);
      const float* bias = small + kOffBias + l * 256;
      if (l == 7) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sigma_partial[i] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = j * 64 + tn * 4 + q;
          const float b = bias[n];
          float h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            h[i] = acc[i][4 * j + q] + b;
            if (l < 8) h[i] = fmaxf(h[i], 0.f);
          }
          if (l == 7) {
            const float ws = small[kOffWSigma + n];
#pragma unroll
            for (int i = 0; i < 8; ++i) sigma_partial[i] = fmaf(h[i], ws, sigma_partial[i]);
          }
          float* dst = act + (64 + n) * kPts + tp * 8;
          *reinterpret_cast<float4*>(dst) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(h[4], h[5], h[6], h[7]);
        }
      }
      if (l == 7) {  // density head: relu(w . h7 + b), VipNeRF01.py:546-553 (eval: no noise)
        const float bs = small[kOffBSigma];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s = half_warp_sum(sigma_partial[i]);
          const int64_t pg = p0 + tp * 8 + i;
          if (tn == 0 && pg < n_points) out_sigma[pg] = fmaxf(s + bs, 0.f);
        }
      }
      __syncthreads();
    }
  }

  // ---- M9 + heads: views_linears.0 (feature columns by matmul, view columns added per point) and
  //      views_output_linear, VipNeRF01.py:568-596
  float acc9[8][8];
  matmul_layer<2>(big + fp32_layer_offset(9), 256, act + 64 * kPts, wbuf, tp, tn, acc9);
  const float* bv = small + kOffBiasViews;
  const float* wvd = small + kOffWViewDir;
  const float* wout = small + kOffWOut;
  float o[8][4];
  view_head(acc9, bv, wvd, wout, pev, tp, tn, o);
  const float* bo = small + kOffBOut;
  if (tn == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t pg = p0 + tp * 8 + i;
      if (pg < n_points) {
        out_rgb[3 * pg + 0] = sigmoidf(o[i][0] + bo[0]);
        out_rgb[3 * pg + 1] = sigmoidf(o[i][1] + bo[1]);
        out_rgb[3 * pg + 2] = sigmoidf(o[i][2] + bo[2]);
        out_vis[pg] = sigmoidf(o[i][3] + bo[3]);
      }
    }
  }
  // ---- secondary views: same head with the direction from each other camera (:527-530, :218-226)
  for (int v = 0; v < fl.n_sec_views; ++v) {
    __syncthreads();
    if (tid < kPts) {
      const int64_t pg = min(p0 + tid, n_points - 1);
      const int64_t ray = pg / S;
      const float o3[3] = {rp.rays_o[3 * ray], rp.rays_o[3 * ray + 1], rp.rays_o[3 * ray + 2]};
      const float d3[3] = {rp.rays_d[3 * ray], rp.rays_d[3 * ray + 1], rp.rays_d[3 * ray + 2]};
      const float* c2 = rp.rays_o2 + (ray * fl.n_sec_views + v) * 3;
      const float o2[3] = {c2[0], c2[1], c2[2]};
      float zz = z[pg];
      if (fl.ndc) zz = depth_from_ndc_secondary(zz, o3[2], d3[2]);
      float dd[3];
      secondary_view_dir(o3, d3, zz, o2, dd);
      dir2[0 * kPts + tid] = dd[0];
      dir2[1 * kPts + tid] = dd[1];
      dir2[2 * kPts + tid] = dd[2];
    }
    __syncthreads();
    for (int t = tid; t < 3 * kPts; t += kThreads) {
      const int p = t % kPts, axis = t / kPts;
      encode_axis<kLView>(dir2[axis * kPts + p], axis, [&](int col, float val) { pev[col * kPts + p] = val; });
    }
    __syncthreads();
    view_head(acc9, bv, wvd, wout, pev, tp, tn, o);
    if (tn == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t pg = p0 + tp * 8 + i;
        if (pg < n_points) out_vis2[pg * fl.n_sec_views + v] = sigmoidf(o[i][3] + bo[3]);
      }
    }
  }
}

}  // namespace

cudaError_t launch_mlp_fp32(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                            const void* packed, float* sigma, float* rgb, float* vis, float* vis2, cudaStream_t s) {
  const int64_t n_points = n_rays * S;
  if (n_points == 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(k_mlp_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return e;
  const float* small = reinterpret_cast<const float*>(packed);
  const float* big = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + kSmallBytes);
  RenderFlags f = fl;
  if (vis2 == nullptr) f.n_sec_views = 0;
  k_mlp_fp32<<<(unsigned)((n_points + kPts - 1) / kPts), kThreads, kSmemBytes, s>>>(rp, f, n_points, S, z, small, big,
                                                                                   sigma, rgb, vis, vis2);
  return cudaGetLastError();
}

}  // namespace vipnerf
