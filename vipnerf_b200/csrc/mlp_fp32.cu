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
// three-stage ring of 8-row shared windows with cp.async; each thread owns an 8-point x 16-column register tile.
//
// Training (reference Trainer01.py:93-102 = model(batch) + loss.backward()):
//  * k_mlp_fp32<true> is the same forward with the density noise of VipNeRF01.py:549-552 added and every
//    activation the backward needs written point-major to HBM (MlpSave) - with 180 GB per GPU saving is cheaper
//    than re-computing;
//  * k_mlp_bwd_fp32 is the backward-data chain of the same tile: head gradients -> views_linears.0 -> feature_linear
//    -> pts_linears.7..1, each a [64 x 256] x [256 x 256] product against the nn.Linear weight in its ORIGINAL
//    [out][in] orientation, masked by the saved ReLU outputs; it stores the pre-activation gradient of every layer,
//    from which train_kernels.cu forms the parameter gradients as split reductions over all points.
#include <cuda_runtime.h>

#include "kernels.h"
#include "layout.cuh"

namespace vipnerf {
namespace {

constexpr int kPts = 64;
constexpr int kThreads = 128;
constexpr int kKC = 8;   // weight rows per shared window: 16 KiB for both buffers, so that TWO blocks fit an SM (2 warps per scheduler hide each other's barriers and load latencies)

constexpr int kActFloats = 320 * kPts;
constexpr int kStages = 3;  // weight ring depth (prefetch distance 2 chunks)
constexpr int kWbufFloats = kStages * kKC * 256;
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
    float* dst = wbuf + (c % kStages) * (kKC * 256);
#pragma unroll
    for (int v = 0; v < kVecPerThread; ++v) {
      const int e = (v * kThreads + tid) * 4;
      cp_async16(dst + e, src + e);
    }
    cp_async_commit();
  };
  prefetch(0);
  if (n_chunks > 1) prefetch(1);
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 2 < n_chunks) { prefetch(c + 2); cp_async_wait<2>(); }
    else if (c + 1 < n_chunks) { cp_async_wait<1>(); }
    else { cp_async_wait<0>(); }
    __syncthreads();
    const float* wb = wbuf + (c % kStages) * (kKC * 256);
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
    __syncthreads();  // stage c % kStages is free for chunk c+3; on the last chunk: all reads of act are done
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
//   hv_out != null (training): relu(pre) of view `view` is stored at hv_out[(point * nviews + view) * 128 + n].
__device__ __forceinline__ void view_head(const float (&acc9)[8][8], const float* __restrict__ bv,
                                          const float* __restrict__ wvd, const float* __restrict__ wout,
                                          const float* pev, int tp, int tn, float (&o)[8][4],
                                          float* __restrict__ hv_out = nullptr, int64_t p0 = 0, int64_t n_points = 0,
                                          int nviews = 1, int view = 0) {
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
#pragma unroll 3
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
    if (hv_out != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t pg = p0 + tp * 8 + i;
        if (pg < n_points)
          *reinterpret_cast<float4*>(hv_out + (pg * nviews + view) * 128 + n0) =
              make_float4(fmaxf(pre[i][0], 0.f), fmaxf(pre[i][1], 0.f), fmaxf(pre[i][2], 0.f), fmaxf(pre[i][3], 0.f));
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

// [27][64] shared view-direction encodings -> pev[(point * nviews + view)][32] (columns 27..31 zero)
__device__ __forceinline__ void save_view_encoding(const float* pev_s, float* __restrict__ dst, int64_t p0,
                                                   int64_t n_points, int nviews, int view) {
  for (int t = threadIdx.x; t < kPts * 32; t += kThreads) {
    const int p = t >> 5, e = t & 31;
    const int64_t pg = p0 + p;
    if (pg < n_points) dst[(pg * nviews + view) * 32 + e] = e < kEncView ? pev_s[e * kPts + p] : 0.f;
  }
}

template <bool kSave>
__global__ void __launch_bounds__(kThreads, 2)
k_mlp_fp32(RayPtrs rp, RenderFlags fl, int64_t n_points, int S, const float* __restrict__ z,
           const float* __restrict__ small, const float* __restrict__ big, float* __restrict__ out_sigma,
           float* __restrict__ out_rgb, float* __restrict__ out_vis, float* __restrict__ out_vis2, MlpSave sv) {
  extern __shared__ __align__(16) float smem[];
  float* act = smem;                    // [320][64]: rows 0..63 encoding (row 63 = 0), rows 64..319 hidden
  float* wbuf = act + kActFloats;       // [kStages][kKC][256]
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
  const int nviews = 1 + fl.n_sec_views;
  if (kSave) {  // encodings, point-major
    for (int t = tid; t < kPts * 64; t += kThreads) {
      const int p = t >> 6, k = t & 63;
      if (p0 + p < n_points) sv.enc[(p0 + p) * 64 + k] = act[k * kPts + p];
    }
    save_view_encoding(pev, sv.pev, p0, n_points, nviews, 0);
  }

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
      float4 bias4[4];   // all 16 bias values of this thread's columns in one round trip
#pragma unroll
      for (int j = 0; j < 4; ++j) bias4[j] = *reinterpret_cast<const float4*>(bias + j * 64 + tn * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float bq[4] = {bias4[j].x, bias4[j].y, bias4[j].z, bias4[j].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = j * 64 + tn * 4 + q;
          const float b = bq[q];
          float h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            h[i] = acc[i][4 * j + q] + b;
            if (l < 8) h[i] = fmaxf(h[i], 0.f);
            if (kSave) acc[i][4 * j + q] = h[i];
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
        if (kSave) {  // h_{l+1} (l < 8) or the feature vector (l == 8), point-major: 16 threads cover 256 B of a row
          float* base = l < 8 ? sv.h + (size_t)l * n_points * 256 : sv.feat;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t pg = p0 + tp * 8 + i;
            if (pg < n_points)
              *reinterpret_cast<float4*>(base + pg * 256 + j * 64 + tn * 4) =
                  make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
          }
        }
      }
      if (l == 7) {  // density head: relu(w . h7 + b), VipNeRF01.py:546-553 (eval: no noise)
        const float bs = small[kOffBSigma];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s = half_warp_sum(sigma_partial[i]);
          const int64_t pg = p0 + tp * 8 + i;
          if (tn == 0 && pg < n_points) {
            float pre = s + bs;
            if (kSave && sv.noise != nullptr) pre = pre + sv.noise[pg];  // raw_noise_std * randn, :549-552
            out_sigma[pg] = fmaxf(pre, 0.f);
          }
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
  view_head(acc9, bv, wvd, wout, pev, tp, tn, o, kSave ? sv.hv : nullptr, p0, n_points, nviews, 0);
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
    if (kSave) save_view_encoding(pev, sv.pev, p0, n_points, nviews, 1 + v);
    view_head(acc9, bv, wvd, wout, pev, tp, tn, o, kSave ? sv.hv : nullptr, p0, n_points, nviews, 1 + v);
    if (tn == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t pg = p0 + tp * 8 + i;
        if (pg < n_points) out_vis2[pg * fl.n_sec_views + v] = sigmoidf(o[i][3] + bo[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward-data chain of one 64-point tile (training).  g = shared [256][64] gradient rows, k-major like `act`.
constexpr int kMaxViews = 17;  // primary + up to 16 secondary views (17 KiB of logit gradients per tile)
constexpr size_t kBwdSmemBytes = (256 * kPts + kWbufFloats + kPts * kMaxViews * 4 + kPts) * sizeof(float);

// Pulls the 64 x 256 fp32 rows a later epilogue will read (ReLU masks) from HBM into L2 while the product before it
// runs: 512 lines of 128 B per tile, four prefetches per thread.
__device__ __forceinline__ void prefetch_rows_l2(const float* __restrict__ base, int64_t p0, int64_t n_points) {
  for (int t = threadIdx.x; t < kPts * 8; t += kThreads) {
    const int64_t pg = p0 + (t >> 3);
    if (pg < n_points) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + pg * 256 + (t & 7) * 32));
  }
}

// Epilogue of one backward product: acc = gradient w.r.t. the OUTPUT of the layer below (its post-ReLU h, or the
// feature vector).  Adds the density head's contribution, applies the ReLU mask from the saved activations,
// stores the result point-major (it is that layer's pre-activation gradient) and k-major into `g` for the next product.
__device__ __forceinline__ void bwd_epilogue(float (&acc)[8][16], float* g, int tp, int tn, int64_t p0, int64_t n_points,
                                             const float* __restrict__ mask_h, float* __restrict__ dst,
                                             const float* dsig_s, const float* __restrict__ wsigma) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n0 = j * 64 + tn * 4;
    float4 ws = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wsigma != nullptr) ws = *reinterpret_cast<const float4*>(wsigma + n0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t pg = p0 + tp * 8 + i;
      const bool valid = pg < n_points;
      float4 v = make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]);
      if (wsigma != nullptr) {
        const float ds = dsig_s[tp * 8 + i];
        v.x = fmaf(ds, ws.x, v.x); v.y = fmaf(ds, ws.y, v.y); v.z = fmaf(ds, ws.z, v.z); v.w = fmaf(ds, ws.w, v.w);
      }
      if (mask_h != nullptr) {
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) m = *reinterpret_cast<const float4*>(mask_h + pg * 256 + n0);
        v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
      }
      if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) *reinterpret_cast<float4*>(dst + pg * 256 + n0) = v;
      acc[i][4 * j] = v.x; acc[i][4 * j + 1] = v.y; acc[i][4 * j + 2] = v.z; acc[i][4 * j + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float* d = g + (n0 + q) * kPts + tp * 8;
      *reinterpret_cast<float4*>(d) = make_float4(acc[0][4 * j + q], acc[1][4 * j + q], acc[2][4 * j + q], acc[3][4 * j + q]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(acc[4][4 * j + q], acc[5][4 * j + q], acc[6][4 * j + q], acc[7][4 * j + q]);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 2)
k_mlp_bwd_fp32(MlpBwdArgs a, const float* __restrict__ small, const float* __restrict__ wb) {
  extern __shared__ __align__(16) float smem[];
  float* g = smem;                         // [256][64]
  float* wbuf = g + 256 * kPts;            // [kStages][kKC][256]
  float* dl_s = wbuf + kWbufFloats;        // [64][nviews][4] head-logit gradients
  float* dsig_s = dl_s + kPts * kMaxViews * 4;  // [64]
  const int tid = threadIdx.x;
  const int tp = tid >> 4, tn = tid & 15;
  const int64_t p0 = (int64_t)blockIdx.x * kPts;
  const int64_t P = a.n_points;
  const int nv = a.nviews;

  for (int t = tid; t < kPts * nv * 4; t += kThreads) {
    const int64_t e = p0 * nv * 4 + t;
    dl_s[t] = e < P * nv * 4 ? a.dlogit[e] : 0.f;
  }
  if (tid < kPts) dsig_s[tid] = p0 + tid < P ? a.dsig[p0 + tid] : 0.f;
  __syncthreads();

  // ---- heads (VipNeRF01.py:579-594 backwards): thread = hidden unit n of views_linears.0, all 64 points.
  //      g_pre[p][view][n] = relu'(hv) * sum_k dlogit[p][view][k] * W_out[k][n];  g_acc9[p][n] = sum over views.
  {
    const int n = tid;
    const float4 wo = *reinterpret_cast<const float4*>(small + kOffWOut + n * 4);
    float accp[kPts];
#pragma unroll
    for (int p = 0; p < kPts; ++p) accp[p] = 0.f;
    const int n_valid = (int)min((int64_t)kPts, P - p0);
    for (int v = 0; v < nv; ++v) {
      // all 64 loads of a view are issued before the first use (clamped address instead of a branch per point):
      // one DRAM latency per view instead of one per point
      float hvv[kPts];
#pragma unroll
      for (int p = 0; p < kPts; ++p) hvv[p] = a.hv[((p0 + min(p, n_valid - 1)) * nv + v) * 128 + n];
#pragma unroll
      for (int p = 0; p < kPts; ++p) {
        const float4 d4 = *reinterpret_cast<const float4*>(dl_s + (p * nv + v) * 4);   // zero for rows past the end
        const float gsum = fmaf(d4.x, wo.x, fmaf(d4.y, wo.y, fmaf(d4.z, wo.z, d4.w * wo.w)));
        hvv[p] = hvv[p] > 0.f ? gsum : 0.f;
        accp[p] += hvv[p];
      }
      if (n_valid == kPts) {
#pragma unroll
        for (int p = 0; p < kPts; ++p) a.dhv[((p0 + p) * nv + v) * 128 + n] = hvv[p];
      } else {
#pragma unroll
        for (int p = 0; p < kPts; ++p)
          if (p < n_valid) a.dhv[((p0 + p) * nv + v) * 128 + n] = hvv[p];
      }
    }
#pragma unroll
    for (int p = 0; p < kPts; ++p) {
      g[n * kPts + p] = accp[p];
      if (p0 + p < P) a.dacc9[(p0 + p) * 128 + n] = accp[p];
    }
  }
  __syncthreads();

  float acc[8][16];
  // views_linears.0 feature columns: g_feature = g_acc9 . W_v[:, :256]
  matmul_layer<4>(wb + kBwdOffViews, 128, g, wbuf, tp, tn, acc);
  bwd_epilogue(acc, g, tp, tn, p0, P, nullptr, a.dfeat, nullptr, nullptr);
  // feature_linear and the density head meet at h8: g_h8 = g_feature . W_f + g_sigma_pre * w_sigma, masked by h8 > 0
  prefetch_rows_l2(a.h + (size_t)7 * P * 256, p0, P);
  matmul_layer<4>(wb + kBwdOffFeature, 256, g, wbuf, tp, tn, acc);
  bwd_epilogue(acc, g, tp, tn, p0, P, a.h + (size_t)7 * P * 256, a.dpre + (size_t)7 * P * 256, dsig_s, small + kOffWSigma);
  // pts_linears.7 .. 1: g_h_l = g_pre_l . W_l (hidden columns), masked by h_l > 0 -> g_pre_{l-1}
  for (int l = 7; l >= 1; --l) {
    prefetch_rows_l2(a.h + (size_t)(l - 1) * P * 256, p0, P);
    matmul_layer<4>(wb + kBwdOffTrunk + (size_t)(7 - l) * 65536, 256, g, wbuf, tp, tn, acc);
    bwd_epilogue(acc, g, tp, tn, p0, P, a.h + (size_t)(l - 1) * P * 256, a.dpre + (size_t)(l - 1) * P * 256, nullptr, nullptr);
  }
}

}  // namespace

cudaError_t launch_mlp_fp32(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                            const void* packed, float* sigma, float* rgb, float* vis, float* vis2, cudaStream_t s,
                            const MlpSave* save) {
  const int64_t n_points = n_rays * S;
  if (n_points == 0) return cudaSuccess;
  const float* small = reinterpret_cast<const float*>(packed);
  const float* big = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + kSmallBytes);
  RenderFlags f = fl;
  if (vis2 == nullptr) f.n_sec_views = 0;
  const unsigned grid = (unsigned)((n_points + kPts - 1) / kPts);
  cudaError_t e;
  // two blocks per SM need the maximum shared-memory carve-out (2 x 112.5 KiB)
  if (save == nullptr) {
    e = cudaFuncSetAttribute(k_mlp_fp32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fp32<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    k_mlp_fp32<false><<<grid, kThreads, kSmemBytes, s>>>(rp, f, n_points, S, z, small, big, sigma, rgb, vis, vis2, MlpSave{});
  } else {
    e = cudaFuncSetAttribute(k_mlp_fp32<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fp32<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    k_mlp_fp32<true><<<grid, kThreads, kSmemBytes, s>>>(rp, f, n_points, S, z, small, big, sigma, rgb, vis, vis2, *save);
  }
  return cudaGetLastError();
}

cudaError_t launch_mlp_bwd_fp32(const MlpBwdArgs& a, const void* packed, cudaStream_t s) {
  if (a.n_points == 0) return cudaSuccess;
  if (a.nviews < 1 || a.nviews > kMaxViews) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_mlp_bwd_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_bwd_fp32, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  const float* small = reinterpret_cast<const float*>(packed);
  const float* wb = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + kSmallBytes) + kFp32BigFloats;
  k_mlp_bwd_fp32<<<(unsigned)((a.n_points + kPts - 1) / kPts), kThreads, kBwdSmemBytes, s>>>(a, small, wb);
  return cudaGetLastError();
}

}  // namespace vipnerf
