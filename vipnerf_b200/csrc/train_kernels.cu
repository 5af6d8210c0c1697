// Training backward of the render path: the gradient of volume_rendering and the
// parameter-gradient reductions.  The reference gets these from torch.autograd over its op graph
// (Trainer01.py:93-102: model(batch) -> LossComputer -> loss.backward()); here each is one explicit kernel.
//
//  k_composite_bwd  one warp per ray, the mirror image of composite_ray (stages.cuh): re-computes alpha /
//                   transmittance / weights from (z, sigma), turns the upstream gradients of every output of
//                   volume_rendering (VipNeRF01.py:331-384) into gradients of the per-sample network outputs with
//                   warp-shuffle scans (the cumprod's backward is an exclusive SUFFIX sum along the ray), and folds in
//                   the ReLU / sigmoid derivatives of the heads, so that what leaves the kernel are logit gradients.
//  k_gemm_tn        dW = dY^T X: C[m][n] = sum_p A[p][m] B[p][n], the reduction running over ALL sample points
//                   (about 10^6 per batch).  fp32 FFMA with 128 x BN tiles fed by a 4-stage cp.async ring; the point
//                   range is split over the grid,
//                   partial tiles go to a scratch buffer and k_reduce_partials adds them in a fixed order, so
//                   gradients are bit-reproducible run to run (no atomics).  Bias gradients (column sums of dY) ride
//                   along in the threads that already hold the dY values.
//  k_small_tn       the same for the 1- and 4-row heads (pts_output_linear, views_output_linear): HBM-bound streaming.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "kernels.h"
#include "layout.cuh"

namespace vipnerf {
namespace {

// ---------------------------------------------------------------------------------------------------------------
template <int SPL>
__global__ void __launch_bounds__(128)
k_composite_bwd(RayPtrs rp, RenderFlags fl, int64_t n_rays, int S, const float* __restrict__ z,
                const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ vis,
                const float* __restrict__ vis2, PassGradPtrs G, LossGrad L, float* __restrict__ dsig, float* __restrict__ dlogit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + warp;
  if (ray >= n_rays) return;
  const int V = fl.n_sec_views, nviews = 1 + V;
  const bool ndc = fl.ndc;
  const float dnorm = vec3_norm(rp.pts_d[3 * ray], rp.pts_d[3 * ray + 1], rp.pts_d[3 * ray + 2]);
  const float oz = rp.rays_o[3 * ray + 2], dz = rp.rays_d[3 * ray + 2];
  const int base = lane * SPL;
  const float* zr = z + ray * S;
  const float* sr = sigma + ray * S;

  // ---- forward quantities, exactly as composite_ray computes them
  float zz[SPL], al[SPL], om[SPL], ex[SPL], dl[SPL], tr[SPL], w[SPL], sg[SPL];
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = base + j;
    const bool valid = i < S;
    zz[j] = valid ? zr[i] : 0.f;
    const float z_next = (i + 1 < S) ? zr[i + 1] : (ndc ? 1.f : 1e10f);
    dl[j] = fmul(fsub(z_next, zz[j]), dnorm);
    sg[j] = valid ? sr[i] : 0.f;
    ex[j] = expf(fmul(-sg[j], dl[j]));
    al[j] = valid ? fsub(1.f, ex[j]) : 0.f;
    om[j] = fadd(fsub(1.f, al[j]), 1e-10f);
  }
  float incl = 1.f;
#pragma unroll
  for (int j = 0; j < SPL; ++j) incl *= om[j];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.f;
  tr[0] = excl;
#pragma unroll
  for (int j = 1; j < SPL; ++j) tr[j] = tr[j - 1] * om[j - 1];
  float acc = 0.f, wz = 0.f;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    w[j] = base + j < S ? fmul(al[j], tr[j]) : 0.f;
    acc += w[j];
    wz += w[j] * zz[j];
  }
  acc = warp_sum(acc);
  wz = warp_sum(wz);
  const float denom = fadd(acc, 1e-6f);
  const float inv = 1.f / denom;
  const float d_nat = wz * inv;            // depth in the sampling space (NDC z in NDC mode)
  float zw[SPL], d_wld = 0.f, e_nat = 0.f, e_wld = 0.f;
  {
    float s1 = 0.f;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      zw[j] = ndc ? depth_from_ndc(zz[j], oz, dz) : zz[j];
      s1 += w[j] * zw[j];
    }
    d_wld = warp_sum(s1) * inv;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      e_nat += w[j] * (zz[j] - d_nat);
      e_wld += w[j] * (zw[j] - d_wld);
    }
    e_nat = warp_sum(e_nat);
    e_wld = warp_sum(e_wld);
  }

  // ---- upstream gradients of the per-ray maps
  float g_rgb[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) g_rgb[k] = G.rgb ? G.rgb[3 * ray + k] : 0.f;
  float g_acc = G.acc ? G.acc[ray] : 0.f;
  // ---- fused losses: dTotalLoss/d(map) from the forward values re-computed above (LossGrad, kernels.h)
  float lg_scale = 0.f, lg_vis = 0.f, lg_prior = 0.f, lg_depth = 0.f;
  if (L.enabled) {
    lg_scale = L.upstream ? *L.upstream : 1.f;
    const bool m_nerf = L.mask_nerf ? L.mask_nerf[ray] != 0 : true;
    const float n_nerf = L.stats[5], n_depth = L.stats[6];
    if (m_nerf && n_nerf > 0.f && L.w_mse != 0.f) {     // MSE01: mean over rays of mean_c (rgb - target)^2
      float cm[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < SPL; ++j)
        if (base + j < S) {
          const float* c = rgb + (ray * S + base + j) * 3;
          cm[0] += w[j] * c[0]; cm[1] += w[j] * c[1]; cm[2] += w[j] * c[2];
        }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float m = warp_sum(cm[k]);
        if (fl.white_bkgd) m += 1.f - acc;
        g_rgb[k] += lg_scale * L.w_mse * 2.f * (m - L.target_rgb[3 * ray + k]) / (3.f * n_nerf);
      }
    }
    // VisibilityLoss01: mean_{r,s} |pred - sg(T)| + |sg(pred) - T|  ->  +-sign / (R S) on the head output and on T
    lg_vis = lg_scale * L.w_vis / ((float)n_rays * (float)S);
    // VisibilityPriorLoss01: mean over nerf rays of sum_v prior (1 - visibility2)
    lg_prior = (m_nerf && n_nerf > 0.f) ? -lg_scale * L.w_prior / n_nerf : 0.f;
    // SparseDepthMSE01: mean over sparse-depth rays of (depth - gt)^2
    if (L.depth_here && L.mask_depth && L.mask_depth[ray] != 0 && n_depth > 0.f)
      lg_depth = lg_scale * L.w_depth * 2.f * (d_wld - L.sparse_depth[ray]) / n_depth;
  }
  if (fl.white_bkgd) g_acc -= g_rgb[0] + g_rgb[1] + g_rgb[2];            // rgb += 1 - acc, :363-364
  // depth / depth_var are the world-space statistics; in NDC mode depth_ndc / depth_var_ndc are the native ones (:356-361)
  const float gd_w = (G.depth ? G.depth[ray] : 0.f) + lg_depth;
  const float gv_w = G.depth_var ? G.depth_var[ray] : 0.f;
  const float gd_n = (ndc && G.depth_ndc) ? G.depth_ndc[ray] : 0.f;
  const float gv_n = (ndc && G.depth_var_ndc) ? G.depth_var_ndc[ray] : 0.f;

  // ---- dL/dw_i
  float gw[SPL];
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = base + j;
    float t = 0.f;
    if (i < S) {
      const float* c = rgb + (ray * S + i) * 3;
      t = g_rgb[0] * c[0] + g_rgb[1] * c[1] + g_rgb[2] * c[2] + g_acc;
      const float aw = zw[j] - d_wld;
      t += gd_w * aw * inv + gv_w * (aw * aw - 2.f * aw * inv * e_wld);
      const float an = zz[j] - d_nat;
      t += gd_n * an * inv + gv_n * (an * an - 2.f * an * inv * e_nat);
      if (G.weights) t += G.weights[ray * S + i];
    }
    gw[j] = t;
  }
  const bool has_v2 = V > 0 && vis2 != nullptr;
  if (has_v2 && (G.visibility2 != nullptr || lg_prior != 0.f)) {
    for (int v = 0; v < V; ++v) {
      const float g2 = (G.visibility2 ? G.visibility2[ray * V + v] : 0.f) + lg_prior * (L.prior ? L.prior[ray * V + v] : 1.f);
      float m = 0.f;
#pragma unroll
      for (int j = 0; j < SPL; ++j)
        if (base + j < S) m += w[j] * vis2[(ray * S + base + j) * V + v];
      const float vis2_map = warp_sum(m) * inv;
#pragma unroll
      for (int j = 0; j < SPL; ++j)
        if (base + j < S) gw[j] += g2 * (vis2[(ray * S + base + j) * V + v] - vis2_map) * inv;
    }
  }

  // ---- w = alpha * T;  T = exclusive cumprod of om;  om = 1 - alpha + 1e-10
  float ga[SPL], x[SPL];
  float loc = 0.f;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = base + j;
    const bool valid = i < S;
    ga[j] = gw[j] * tr[j] + ((valid && G.alpha) ? G.alpha[ray * S + i] : 0.f);
    float gt = gw[j] * al[j] + ((valid && G.visibility) ? G.visibility[ray * S + i] : 0.f);
    if (valid && lg_vis != 0.f) {   // d|sg(pred) - T| / dT
      const float d = tr[j] - vis[ray * S + i];
      gt += d > 0.f ? lg_vis : (d < 0.f ? -lg_vis : 0.f);
    }
    x[j] = valid ? gt * tr[j] : 0.f;
    loc += x[j];
  }
  float sfx_incl = loc;   // inclusive suffix sum over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_down_sync(0xffffffffu, sfx_incl, o);
    if (lane + o < 32) sfx_incl += t;
  }
  float sfx = __shfl_down_sync(0xffffffffu, sfx_incl, 1);  // sum over the samples of all later lanes
  if (lane == 31) sfx = 0.f;
#pragma unroll
  for (int j = SPL - 1; j >= 0; --j) {
    const int i = base + j;
    // sfx = sum_{i' > i} gT_i' T_i'  ->  dL/d om_i = sfx / om_i  ->  alpha_i receives the negative of it
    const float g_alpha = ga[j] - sfx / om[j];
    sfx += x[j];
    if (i < S) {
      float gs = g_alpha * dl[j] * ex[j];                              // alpha = 1 - exp(-sigma * delta)
      if (G.raw_sigma) gs += G.raw_sigma[ray * S + i];
      dsig[ray * S + i] = sg[j] > 0.f ? gs : 0.f;                      // sigma = relu(logit (+ noise))
    }
  }

  // ---- head logits: rgb / visibility (primary view), visibility2 (secondary views); sigmoid' = y (1 - y)
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = base + j;
    if (i >= S) continue;
    const int64_t p = ray * S + i;
    float4 o;
    float* of = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float c = rgb[p * 3 + k];
      float gc = w[j] * g_rgb[k];
      if (G.raw_rgb) gc += G.raw_rgb[p * 3 + k];
      of[k] = gc * c * (1.f - c);
    }
    const float sv = vis[p];
    float g_sv = G.raw_visibility ? G.raw_visibility[p] : 0.f;
    if (lg_vis != 0.f) {            // d|pred - sg(T)| / dpred
      const float d = sv - tr[j];
      g_sv += d > 0.f ? lg_vis : (d < 0.f ? -lg_vis : 0.f);
    }
    of[3] = g_sv * sv * (1.f - sv);
    *reinterpret_cast<float4*>(dlogit + p * nviews * 4) = o;
    for (int v = 0; v < V; ++v) {
      float gl = 0.f;
      if (has_v2) {
        const float s2 = vis2[p * V + v];
        float gv = (G.visibility2 || lg_prior != 0.f)
                       ? w[j] * ((G.visibility2 ? G.visibility2[ray * V + v] : 0.f) + lg_prior * (L.prior ? L.prior[ray * V + v] : 1.f)) * inv
                       : 0.f;
        if (G.raw_visibility2) gv += G.raw_visibility2[p * V + v];
        gl = gv * s2 * (1.f - s2);
      }
      *reinterpret_cast<float4*>(dlogit + (p * nviews + 1 + v) * 4) = make_float4(0.f, 0.f, 0.f, gl);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int kGemmBM = 128;
constexpr int kGemmBK = 16;
constexpr int kGemmTargetCtas = 296;        // two CTAs per SM
constexpr int kGemmMaxMainFloats = 80 * 65536;
constexpr int kGemmMaxBiasFloats = kGemmTargetCtas * 256;

constexpr int kGemmStages = 4;   // cp.async ring depth: loads run three k-steps ahead of the FMAs
template <int BN> constexpr size_t gemm_smem_bytes() { return (size_t)kGemmStages * kGemmBK * (kGemmBM + BN) * sizeof(float); }

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

template <int BN>
__global__ void __launch_bounds__(256, 2)
k_gemm_tn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int64_t n_rows,
          int64_t rows_per_split, int M, int N, float* __restrict__ partial, float* __restrict__ bias_partial) {
  constexpr int TN = BN / 16;
  extern __shared__ __align__(16) float gemm_smem[];
  float* As = gemm_smem;                                        // [stage][16][128]
  float* Bs = gemm_smem + kGemmStages * kGemmBK * kGemmBM;      // [stage][16][BN]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * BN;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(n_rows, r_begin + rows_per_split);
  const int n_steps = r_end > r_begin ? (int)((r_end - r_begin + kGemmBK - 1) / kGemmBK) : 0;
  const bool do_bias = bias_partial != nullptr && blockIdx.y == 0 && tx == 0;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bsum[i] = 0.f;

  // loader mapping: A tile = 16 rows x 32 float4 (two per thread); B tile = 16 rows x BN/4 float4.  Rows past the
  // end of this CTA's point range are zero-filled (clamped source address, src-size 0).
  constexpr int kBVec = kGemmBK * BN / 4;           // 512 / 256 / 128
  constexpr int kBPer = (kBVec + 255) / 256;        // 2 / 1 / 1
  const int rows_total = (int)max((int64_t)0, r_end - r_begin);
  const float* a_src[2];
  const float* b_src[kBPer];
  int a_row[2], b_row[kBPer], a_dst[2], b_dst[kBPer];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int idx = tid + t * 256, c4 = idx & 31;
    a_row[t] = idx >> 5;
    a_dst[t] = a_row[t] * kGemmBM + c4 * 4;
    a_src[t] = A + (r_begin + a_row[t]) * lda + m0 + c4 * 4;
  }
#pragma unroll
  for (int t = 0; t < kBPer; ++t) {
    const int idx = tid + t * 256, c4 = idx % (BN / 4);
    b_row[t] = idx / (BN / 4);
    b_dst[t] = b_row[t] * BN + c4 * 4;
    b_src[t] = B + (r_begin + b_row[t]) * ldb + n0 + c4 * 4;
  }
  auto issue_tile = [&](int step) {
    if (step < n_steps) {
      const int left = rows_total - step * kGemmBK;     // rows of this CTA's range at or after the tile's first row
      float* as = As + (step % kGemmStages) * kGemmBK * kGemmBM;
      float* bs = Bs + (step % kGemmStages) * kGemmBK * BN;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const bool valid = a_row[t] < left;
        cp_async16_zfill(as + a_dst[t], valid ? a_src[t] : A, valid);
        a_src[t] += (size_t)kGemmBK * lda;
      }
#pragma unroll
      for (int t = 0; t < kBPer; ++t) {
        if (kBVec >= 256 * (t + 1) || tid + t * 256 < kBVec) {
          const bool valid = b_row[t] < left;
          cp_async16_zfill(bs + b_dst[t], valid ? b_src[t] : B, valid);
          b_src[t] += (size_t)kGemmBK * ldb;
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);   // one group per step, empty ones included, keeps the count uniform
  };

#pragma unroll
  for (int st = 0; st < kGemmStages - 1; ++st) issue_tile(st);
  for (int step = 0; step < n_steps; ++step) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kGemmStages - 2));   // this step's tile has landed
    __syncthreads();                                                     // ... for every thread; step-1's buffer is free
    issue_tile(step + kGemmStages - 1);
    const float* as = As + (step % kGemmStages) * kGemmBK * kGemmBM;
    const float* bs = Bs + (step % kGemmStages) * kGemmBK * BN;
#pragma unroll
    for (int k = 0; k < kGemmBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(as + k * kGemmBM + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(as + k * kGemmBM + 64 + ty * 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(bs + k * BN + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(bs + k * BN + BN / 2 + tx * 4);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      } else if constexpr (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(bs + k * BN + tx * 4);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      } else {
        const float2 b0 = *reinterpret_cast<const float2*>(bs + k * BN + tx * 2);
        b[0] = b0.x; b[1] = b0.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) bsum[i] += a[i];
      }
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);

  // partial[z][m][n]
  float* out = partial + (size_t)blockIdx.z * M * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    float* row = out + (size_t)m * N + n0;
    if constexpr (TN == 8) {
      *reinterpret_cast<float4*>(row + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(row + BN / 2 + tx * 4) = make_float4(acc[i][TN - 4], acc[i][TN - 3], acc[i][TN - 2], acc[i][TN - 1]);
    } else if constexpr (TN == 4) {
      *reinterpret_cast<float4*>(row + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
      *reinterpret_cast<float2*>(row + tx * 2) = make_float2(acc[i][0], acc[i][1]);
    }
    if (do_bias) bias_partial[(size_t)blockIdx.z * M + m] = bsum[i];
  }
}

// dst[m * ldc + n] = sum_s partial[s][m][n], n < n_valid  (fixed summation order: bit-reproducible).
// A block owns 64 output elements; its four warp pairs each add every fourth partial tile (eight independent loads in
// flight per thread, 128-byte rows per warp), and the four slice sums are combined in a fixed order through shared memory
// - 4 x the parallelism of one thread per element, which left the 38 MB of a 256 x 256 product's partials latency-bound.
__device__ __forceinline__ void reduce_partials_block(const float* __restrict__ partial, int n_split, int M, int N,
                                                      float* __restrict__ dst, int ldc, int n_valid,
                                                      const uint32_t* __restrict__ scale_def, int block) {
  __shared__ float slice_sum[4][64];
  const int e = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int i = block * 64 + e;
  const bool valid = i < M * n_valid;
  const int m = valid ? i / n_valid : 0, n = valid ? i % n_valid : 0;
  float s = 0.f;
  const size_t stride = (size_t)M * N;
  const float* src = partial + (size_t)m * N + n;
  if (valid) {
    int k = slice;
    for (; k + 28 < n_split; k += 32) {   // eight independent loads in flight, summed in index order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(k + 4 * u) * stride];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; k < n_split; k += 4) s += src[(size_t)k * stride];
  }
  slice_sum[slice][e] = s;
  __syncthreads();
  if (slice == 0 && valid) {
    // fp16 training mode: the partial sums carry the power-of-two scale of their gradient operand (exact to undo)
    const float inv_scale = scale_def ? 1.f / grad_scale_from_amax(*scale_def) : 1.f;
    dst[(size_t)m * ldc + n] = ((slice_sum[0][e] + slice_sum[1][e]) + (slice_sum[2][e] + slice_sum[3][e])) * inv_scale;
  }
}

__global__ void __launch_bounds__(256)
k_reduce_partials(const float* __restrict__ partial, int n_split, int M, int N, float* __restrict__ dst,
                  int ldc, int n_valid, const uint32_t* __restrict__ scale_def) {
  reduce_partials_block(partial, n_split, M, N, dst, ldc, n_valid, scale_def, blockIdx.x);
}

// the reductions behind a grouped product launch (weights and bias of up to kGemmGroupMax problems) as ONE launch:
// blockIdx.y = job
struct ReduceJobs { ReduceJob job[2 * kGemmGroupMax]; };
__global__ void __launch_bounds__(256) k_reduce_partials_jobs(const __grid_constant__ ReduceJobs jobs) {
  const ReduceJob& j = jobs.job[blockIdx.y];
  if ((int)blockIdx.x * 64 >= j.M * j.n_valid) return;
  reduce_partials_block(j.partial, j.n_split, j.M, j.N, j.dst, j.ldc, j.n_valid, j.scale_def, blockIdx.x);
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
}

// out[m][n] = sum_r G[r][m] * H[r][n], gsum[m] = sum_r G[r][m] for the 1- and 4-row heads: an HBM stream over H.
// A thread owns EIGHT adjacent columns (one 16-byte load of an fp16 row, two of an fp32 row), so N / 8 threads cover a row
// and a 256-thread block works on 2048 / N rows at a time.  The wide loads are what keeps enough bytes in flight: with
// column pairs per thread the compiler kept only two 4-byte loads per thread outstanding and the stream ran at 0.3 of
// the HBM roofline (the r01 kernel - one row per block iteration, half of its threads idle at N = 128 - at a tenth).
// The row groups' sums are combined in a fixed order through shared memory.
template <int M, typename T>
__global__ void __launch_bounds__(256)
k_small_tn(const float* __restrict__ G, const T* __restrict__ H, int N, int64_t n_rows, int64_t rows_per_split,
           float* __restrict__ partial, float* __restrict__ gsum_partial) {
  constexpr int kC = 8;                          // columns per thread
  constexpr int kStride = kC * M + 1;            // floats a thread publishes (+ its share of the column sums of G)
  __shared__ float red[256 * kStride];
  const int tpr = N / kC;                        // threads per row
  const int rpp = 256 / tpr;                     // rows per pass of the block
  const int rg = threadIdx.x / tpr, c = threadIdx.x % tpr;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_split;
  const int64_t r_end = min(n_rows, r_begin + rows_per_split);
  float acc[M][kC];
  float gs[M];      // column sums of G, kept by every thread (the row's G values are in its registers anyway); thread
                    // c < M of a row group publishes column c
#pragma unroll
  for (int m = 0; m < M; ++m) {
    gs[m] = 0.f;
#pragma unroll
    for (int j = 0; j < kC; ++j) acc[m][j] = 0.f;
  }
  auto load_g = [&](int64_t rr, float (&g)[M]) {   // one row of G: a single 16-byte load for the 4-row head
    if constexpr (M == 4) {
      const float4 t = *reinterpret_cast<const float4*>(G + rr * 4);
      g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
    } else {
#pragma unroll
      for (int m = 0; m < M; ++m) g[m] = G[rr * M + m];
    }
  };
  constexpr int kU = 4;
  int64_t r = r_begin + rg;
  for (; r + (kU - 1) * rpp < r_end; r += kU * rpp) {
    float h[kU][kC];
    float g[kU][M];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t rr = r + u * rpp;
      load8(H + rr * N + kC * c, h[u]);
      load_g(rr, g[u]);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
#pragma unroll
        for (int j = 0; j < kC; ++j) acc[m][j] = fmaf(g[u][m], h[u][j], acc[m][j]);
        gs[m] += g[u][m];
      }
    }
  }
  for (; r < r_end; r += rpp) {
    float h[kC], g[M];
    load8(H + r * N + kC * c, h);
    load_g(r, g);
#pragma unroll
    for (int m = 0; m < M; ++m) {
#pragma unroll
      for (int j = 0; j < kC; ++j) acc[m][j] = fmaf(g[m], h[j], acc[m][j]);
      gs[m] += g[m];
    }
  }
  float gsel = 0.f;
#pragma unroll
  for (int m = 0; m < M; ++m)
    if (c == m) gsel = gs[m];
  // combine the row groups in index order
  float* mine = red + threadIdx.x * kStride;
#pragma unroll
  for (int m = 0; m < M; ++m)
#pragma unroll
    for (int j = 0; j < kC; ++j) mine[kC * m + j] = acc[m][j];
  mine[kC * M] = gsel;
  __syncthreads();
  if (rg == 0) {
    float tot[kStride];
#pragma unroll
    for (int j = 0; j < kStride; ++j) tot[j] = 0.f;
    for (int q = 0; q < rpp; ++q) {
      const float* other = red + (q * tpr + c) * kStride;
#pragma unroll
      for (int j = 0; j < kStride; ++j) tot[j] += other[j];
    }
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
      for (int j = 0; j < kC; ++j) partial[((size_t)blockIdx.x * M + m) * N + kC * c + j] = tot[kC * m + j];
    if (c < M) gsum_partial[(size_t)blockIdx.x * M + c] = tot[kC * M];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core training path (VIPNERF_FLAG_TRAIN_TF32): the pieces of the MLP that are not matrix products.
//
// k_encode_points: sample points o + d z and their 63-d encoding (column 63 = 0), plus the 27-d view-direction encoding
// of the primary view and of every secondary view (compute_other_view_dirs, VipNeRF01.py:218-226).  One thread per point.
// T = float: enc [P][64], pev [P][nviews][32] fp32.  T = __half (fp16 training mode): enc [P][64], pev [P][nviews][64]
// fp16 - 128-byte rows either way (what the TMA boxes of the parameter-gradient product want).
__device__ __forceinline__ void store_as(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_as(__half* p, float v) { *p = __float2half_rn(v); }
// gradients: fp16 stores saturate at +-65504 instead of overflowing to inf
__device__ __forceinline__ void store_sat(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_sat(__half* p, float v) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  *reinterpret_cast<unsigned short*>(p) = r;
}

template <typename T>
__global__ void __launch_bounds__(128)
k_encode_points(RayPtrs rp, RenderFlags fl, int64_t n_points, int S, const float* __restrict__ z,
                T* __restrict__ enc, T* __restrict__ pev) {
  constexpr int kPevCols = sizeof(T) == 2 ? 64 : 32;
  // A thread computes one point's row, but writing 64 (32) floats of its own row touches 32 different lines per
  // instruction; the block's rows are contiguous in memory, so they are staged in shared memory (row stride 65 / 33:
  // conflict-free for row-per-thread writes) and copied out with full-line stores.
  __shared__ float tile[128 * 65];
  const int64_t p0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t pg = p0 + threadIdx.x;
  const bool valid = pg < n_points;
  const int n_here = (int)min((int64_t)blockDim.x, n_points - p0);
  const int64_t pc = valid ? pg : n_points - 1;
  const int64_t ray = pc / S;
  const int nviews = 1 + fl.n_sec_views;
  const float zz = z[pc];
  float* e = tile + threadIdx.x * 65;
#pragma unroll
  for (int axis = 0; axis < 3; ++axis) {
    const float x = fadd(rp.pts_o[3 * ray + axis], fmul(rp.pts_d[3 * ray + axis], zz));   // :105-107
    encode_axis<kLPts>(x, axis, [&](int col, float v) { e[col] = v; });
  }
  e[63] = 0.f;
  __syncthreads();
  if constexpr (sizeof(T) == 2) {   // fp16: eight columns = one 16-byte store per item
    for (int i = threadIdx.x; i < n_here * 8; i += blockDim.x) {
      const float* src = tile + (i >> 3) * 65 + (i & 7) * 8;
      const __half2 a = __floats2half2_rn(src[0], src[1]), b = __floats2half2_rn(src[2], src[3]),
                    c2 = __floats2half2_rn(src[4], src[5]), d = __floats2half2_rn(src[6], src[7]);
      *reinterpret_cast<uint4*>(enc + p0 * 64 + (int64_t)i * 8) =
          make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                     *reinterpret_cast<const uint32_t*>(&c2), *reinterpret_cast<const uint32_t*>(&d));
    }
  } else {
    for (int i = threadIdx.x; i < n_here * 64; i += blockDim.x) store_as(enc + p0 * 64 + i, tile[(i >> 6) * 65 + (i & 63)]);
  }
  for (int v = 0; v < nviews; ++v) {
    float dir[3];
    if (v == 0) {
      dir[0] = rp.view_dirs[3 * ray]; dir[1] = rp.view_dirs[3 * ray + 1]; dir[2] = rp.view_dirs[3 * ray + 2];
    } else {
      const float o3[3] = {rp.rays_o[3 * ray], rp.rays_o[3 * ray + 1], rp.rays_o[3 * ray + 2]};
      const float d3[3] = {rp.rays_d[3 * ray], rp.rays_d[3 * ray + 1], rp.rays_d[3 * ray + 2]};
      const float* c2 = rp.rays_o2 + (ray * fl.n_sec_views + (v - 1)) * 3;
      const float o2[3] = {c2[0], c2[1], c2[2]};
      const float zw = fl.ndc ? depth_from_ndc_secondary(zz, o3[2], d3[2]) : zz;
      secondary_view_dir(o3, d3, zw, o2, dir);
    }
    __syncthreads();     // the previous copy-out has read the tile
    float* pe = tile + threadIdx.x * 33;
#pragma unroll
    for (int axis = 0; axis < 3; ++axis) encode_axis<kLView>(dir[axis], axis, [&](int col, float val) { pe[col] = val; });
#pragma unroll
    for (int c = kEncView; c < 32; ++c) pe[c] = 0.f;
    __syncthreads();
    // row r of this view lives at pev[((p0 + r) * nviews + v) * kPevCols ...]: one full 128-byte line per row
    if constexpr (sizeof(T) == 2) {   // fp16 rows: 8 x 16 bytes, the upper four all zero (columns 32..63)
      for (int i = threadIdx.x; i < n_here * 8; i += blockDim.x) {
        const int r = i >> 3, q = i & 7;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (q < 4) {
          const float* src = tile + r * 33 + q * 8;
          const __half2 a = __floats2half2_rn(src[0], src[1]), b = __floats2half2_rn(src[2], src[3]),
                        c2 = __floats2half2_rn(src[4], src[5]), d = __floats2half2_rn(src[6], src[7]);
          w = make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                         *reinterpret_cast<const uint32_t*>(&c2), *reinterpret_cast<const uint32_t*>(&d));
        }
        *reinterpret_cast<uint4*>(pev + ((p0 + r) * nviews + v) * kPevCols + q * 8) = w;
      }
    } else {
      for (int i = threadIdx.x; i < n_here * kPevCols; i += blockDim.x) {
        const int r = i / kPevCols, c = i % kPevCols;
        store_as(pev + ((p0 + r) * nviews + v) * kPevCols + c, c < 32 ? tile[r * 33 + c] : 0.f);
      }
    }
  }
}

// k_heads_fwd: everything behind the products - density head (+ noise, ReLU; VipNeRF01.py:546-553) from h7 and, per
// view, views_linears.0's direction columns + bias on top of the feature product (acc9), ReLU, views_output_linear and
// the sigmoids (:576-594).  One thread per point; the 16 KiB of head weights sit in shared memory (every lane reads the
// same word: broadcasts).  Stores the views layer's output per view for the backward.
template <typename T>
__global__ void __launch_bounds__(128)
k_heads_fwd(int64_t n_points, int nviews, const float* __restrict__ small, const float* __restrict__ h7,
            const float* __restrict__ acc9, const T* __restrict__ pev, const float* __restrict__ noise,
            float* __restrict__ out_sigma, float* __restrict__ out_rgb, float* __restrict__ out_vis,
            float* __restrict__ out_vis2, T* __restrict__ hv) {
  constexpr bool kHalf = sizeof(T) == 2;
  constexpr int kPevCols = kHalf ? 64 : 32;
  __shared__ __align__(16) float s_wvd[kEncView * 128];
  __shared__ __align__(16) float s_wout[128 * 4];
  __shared__ float s_bv[128];
  __shared__ float s_ws[256];
  for (int i = threadIdx.x; i < kEncView * 128; i += blockDim.x) s_wvd[i] = small[kOffWViewDir + i];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s_wout[i] = small[kOffWOut + i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_bv[i] = small[kOffBiasViews + i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ws[i] = small[kOffWSigma + i];
  __syncthreads();
  const int64_t pg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pg >= n_points) return;
  {
    float s = 0.f;
    if (h7 == nullptr) {
      s = out_sigma[pg];     // the producer of h7 (k_linear_tc's epilogue, LinearTcArgs::dot_vec) left w_sigma . h7 here
    } else {
      const float4* h = reinterpret_cast<const float4*>(h7 + pg * 256);
#pragma unroll 8
      for (int i = 0; i < 64; ++i) {
        const float4 a = h[i];
        s = fmaf(a.x, s_ws[4 * i], s); s = fmaf(a.y, s_ws[4 * i + 1], s); s = fmaf(a.z, s_ws[4 * i + 2], s); s = fmaf(a.w, s_ws[4 * i + 3], s);
      }
    }
    float pre = s + small[kOffBSigma];
    if (noise != nullptr) pre = pre + noise[pg];
    out_sigma[pg] = fmaxf(pre, 0.f);
  }
  const float4* a9 = reinterpret_cast<const float4*>(acc9 + pg * 128);
  const float bo[4] = {small[kOffBOut], small[kOffBOut + 1], small[kOffBOut + 2], small[kOffBOut + 3]};
  for (int v = 0; v < nviews; ++v) {
    float pe[kEncView];
    const T* pp = pev + (pg * nviews + v) * kPevCols;
#pragma unroll
    for (int e = 0; e < kEncView; ++e) pe[e] = to_f32(pp[e]);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    T* hv_row = hv + (pg * nviews + v) * 128;
    for (int n4 = 0; n4 < 32; ++n4) {
      const float4 a = a9[n4];
      float pre[4] = {a.x + s_bv[4 * n4], a.y + s_bv[4 * n4 + 1], a.z + s_bv[4 * n4 + 2], a.w + s_bv[4 * n4 + 3]};
#pragma unroll
      for (int e = 0; e < kEncView; ++e) {
        const float4 w = *reinterpret_cast<const float4*>(s_wvd + e * 128 + 4 * n4);
        pre[0] = fmaf(pe[e], w.x, pre[0]); pre[1] = fmaf(pe[e], w.y, pre[1]);
        pre[2] = fmaf(pe[e], w.z, pre[2]); pre[3] = fmaf(pe[e], w.w, pre[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pre[q] = fmaxf(pre[q], 0.f);
        const float4 wo = *reinterpret_cast<const float4*>(s_wout + (4 * n4 + q) * 4);
        o[0] = fmaf(pre[q], wo.x, o[0]); o[1] = fmaf(pre[q], wo.y, o[1]);
        o[2] = fmaf(pre[q], wo.z, o[2]); o[3] = fmaf(pre[q], wo.w, o[3]);
      }
      if constexpr (kHalf) {   // saved in fp16: the parameter-gradient product reads 11 significand bits either way
        const __half2 lo = __floats2half2_rn(pre[0], pre[1]), hi = __floats2half2_rn(pre[2], pre[3]);
        *reinterpret_cast<uint2*>(hv_row + 4 * n4) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      } else {
        *reinterpret_cast<float4*>(hv_row + 4 * n4) = make_float4(pre[0], pre[1], pre[2], pre[3]);
      }
    }
    if (v == 0) {
      out_rgb[3 * pg + 0] = 1.f / (1.f + expf(-(o[0] + bo[0])));
      out_rgb[3 * pg + 1] = 1.f / (1.f + expf(-(o[1] + bo[1])));
      out_rgb[3 * pg + 2] = 1.f / (1.f + expf(-(o[2] + bo[2])));
      out_vis[pg] = 1.f / (1.f + expf(-(o[3] + bo[3])));
    } else {
      out_vis2[pg * (nviews - 1) + (v - 1)] = 1.f / (1.f + expf(-(o[3] + bo[3])));
    }
  }
}

// k_heads_bwd: the head part of the backward (the first phase of k_mlp_bwd_fp32 as a kernel of its own): per view
// g_pre[p][view][n] = relu'(hv) * sum_k dlogit[p][view][k] * W_out[k][n]; their sum over views is the gradient of the
// feature product.
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const __half* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4_sat(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4_sat(__half* p, const float (&v)[4]) {   // saturating at +-65504
  uint32_t lo, hi;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v[1]), "f"(v[0]));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v[3]), "f"(v[2]));
  *reinterpret_cast<uint2*>(p) = make_uint2(lo, hi);
}

// One warp per (point, view) row: a lane owns four adjacent hidden units, so every load and store of a warp is one full
// 256- / 512-byte row (the r01 kernel moved 2 - 4 bytes per lane and instruction and ran at a third of the HBM
// roofline).  A warp walks point pairs with all 2 * NV rows of a pair loaded before the arithmetic; NV = 0 = any count.
template <typename T, int NV>
__global__ void __launch_bounds__(256, 3)
k_heads_bwd(int64_t n_points, int nviews_rt, const float* __restrict__ small, const float* __restrict__ dlogit,
            const T* __restrict__ hv, T* __restrict__ dhv, T* __restrict__ dacc9, const uint32_t* __restrict__ scale_def,
            uint32_t* __restrict__ amax_out) {
  constexpr int kPts = 2;
  constexpr int kMaxV = NV > 0 ? NV : 1;
  const int nviews = NV > 0 ? NV : nviews_rt;
  const int lane = threadIdx.x & 31;
  // fp16 mode: both outputs carry the power-of-two scale defined by max |dlogit| (grad_scale_from_amax)
  const float scale = scale_def ? grad_scale_from_amax(*scale_def) : 1.f;
  float wo[4][4];          // views_output_linear.weight[k][4 lane + j] (small: transposed [128][4])
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 w = *reinterpret_cast<const float4*>(small + kOffWOut + (4 * lane + j) * 4);
    wo[j][0] = w.x * scale; wo[j][1] = w.y * scale; wo[j][2] = w.z * scale; wo[j][3] = w.w * scale;   // exact: a power of two
  }
  float amax = 0.f;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t p0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kPts; p0 < n_points; p0 += n_warps * kPts) {
    float acc[kPts][4];
#pragma unroll
    for (int i = 0; i < kPts; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    auto one_view = [&](const float (&h)[4], const float4& d, int i, int v) {
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gsum = fmaf(d.x, wo[j][0], fmaf(d.y, wo[j][1], fmaf(d.z, wo[j][2], d.w * wo[j][3])));
        g[j] = h[j] > 0.f ? gsum : 0.f;
        acc[i][j] += g[j];
      }
      if (p0 + i < n_points) store4_sat(dhv + ((p0 + i) * nviews + v) * 128 + 4 * lane, g);
    };
    if constexpr (NV > 0) {
      float h[kPts][kMaxV][4];
      float4 d4[kPts][kMaxV];
#pragma unroll
      for (int i = 0; i < kPts; ++i)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int64_t row = min(p0 + i, n_points - 1) * NV + v;
          load4(hv + row * 128 + 4 * lane, h[i][v]);
          d4[i][v] = *reinterpret_cast<const float4*>(dlogit + row * 4);
        }
#pragma unroll
      for (int i = 0; i < kPts; ++i)
#pragma unroll
        for (int v = 0; v < NV; ++v) one_view(h[i][v], d4[i][v], i, v);
    } else {
      for (int v = 0; v < nviews; ++v)
#pragma unroll
        for (int i = 0; i < kPts; ++i) {
          const int64_t row = min(p0 + i, n_points - 1) * nviews + v;
          float h[4];
          load4(hv + row * 128 + 4 * lane, h);
          one_view(h, *reinterpret_cast<const float4*>(dlogit + row * 4), i, v);
        }
    }
#pragma unroll
    for (int i = 0; i < kPts; ++i)
      if (p0 + i < n_points) {
        store4_sat(dacc9 + (p0 + i) * 128 + 4 * lane, acc[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) amax = fmaxf(amax, fabsf(acc[i][j]));
      }
  }
  if (amax_out != nullptr) {   // un-scaled maximum of what feeds the chain's first product
    const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(amax / scale));
    if (lane == 0 && m != 0u) atomicMax(amax_out, m);
  }
}

// fp16 training mode: the two maxima that anchor the gradient scales of the backward chain (kernels.h: launch_grad_amax)
__global__ void __launch_bounds__(256)
k_grad_amax(int64_t n_logit, int64_t n_points, const float* __restrict__ dlogit, const float* __restrict__ dsig,
            const float* __restrict__ small, uint32_t* __restrict__ amax_logit, uint32_t* __restrict__ amax_sigma) {
  float ml = 0.f, ms = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_logit; i += stride) ml = fmaxf(ml, fabsf(dlogit[i]));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += stride) ms = fmaxf(ms, fabsf(dsig[i]));
  ms *= small[kOffBSigma + 1];      // max |pts_output_linear.weight| (k_pack_small)
  const uint32_t bl = __reduce_max_sync(0xffffffffu, __float_as_uint(ml));
  const uint32_t bs = __reduce_max_sync(0xffffffffu, __float_as_uint(ms));
  if ((threadIdx.x & 31) == 0) {
    if (bl != 0u && bl < 0x7f800000u) atomicMax(amax_logit, bl);
    if (bs != 0u && bs < 0x7f800000u) atomicMax(amax_sigma, bs);
  }
}

}  // namespace

cudaError_t launch_encode_points(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                                 void* enc, void* pev, cudaStream_t s, bool half) {
  const int64_t P = n_rays * S;
  if (P == 0) return cudaSuccess;
  if (half) k_encode_points<__half><<<(unsigned)((P + 127) / 128), 128, 0, s>>>(rp, fl, P, S, z, static_cast<__half*>(enc), static_cast<__half*>(pev));
  else k_encode_points<float><<<(unsigned)((P + 127) / 128), 128, 0, s>>>(rp, fl, P, S, z, static_cast<float*>(enc), static_cast<float*>(pev));
  return cudaGetLastError();
}

cudaError_t launch_heads_fwd(int64_t n_points, int nviews, const void* packed, const float* h7, const float* acc9,
                             const void* pev, const float* noise, float* sigma, float* rgb, float* vis, float* vis2,
                             void* hv, cudaStream_t s, bool half) {
  if (n_points == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n_points + 127) / 128);
  const float* small = reinterpret_cast<const float*>(packed);
  if (half) k_heads_fwd<__half><<<grid, 128, 0, s>>>(n_points, nviews, small, h7, acc9, static_cast<const __half*>(pev), noise, sigma, rgb, vis, vis2, static_cast<__half*>(hv));
  else k_heads_fwd<float><<<grid, 128, 0, s>>>(n_points, nviews, small, h7, acc9, static_cast<const float*>(pev), noise, sigma, rgb, vis, vis2, static_cast<float*>(hv));
  return cudaGetLastError();
}

template <typename T>
void launch_heads_bwd_t(int64_t n_points, int nviews, const float* small, const float* dlogit, const void* hv, void* dhv,
                        void* dacc9, const uint32_t* scale_def, uint32_t* amax_out, cudaStream_t s) {
  // 8 warps per block, a point pair per warp and pass; the grid is capped and the warps loop
  int64_t blocks = (n_points + 15) / 16;
  if (blocks > 148 * 12) blocks = 148 * 12;
  const T* h = static_cast<const T*>(hv);
  T* dh = static_cast<T*>(dhv);
  T* da = static_cast<T*>(dacc9);
#define VIPNERF_HEADS_BWD(NV) k_heads_bwd<T, NV><<<(unsigned)blocks, 256, 0, s>>>(n_points, nviews, small, dlogit, h, dh, da, scale_def, amax_out)
  switch (nviews) {
    case 1: VIPNERF_HEADS_BWD(1); break;
    case 2: VIPNERF_HEADS_BWD(2); break;
    case 3: VIPNERF_HEADS_BWD(3); break;
    case 4: VIPNERF_HEADS_BWD(4); break;
    default: VIPNERF_HEADS_BWD(0); break;
  }
#undef VIPNERF_HEADS_BWD
}

cudaError_t launch_heads_bwd(int64_t n_points, int nviews, const void* packed, const float* dlogit, const void* hv,
                             void* dhv, void* dacc9, cudaStream_t s, bool half, const uint32_t* scale_def, uint32_t* amax_out) {
  if (n_points == 0) return cudaSuccess;
  const float* small = reinterpret_cast<const float*>(packed);
  if (half) launch_heads_bwd_t<__half>(n_points, nviews, small, dlogit, hv, dhv, dacc9, scale_def, amax_out, s);
  else launch_heads_bwd_t<float>(n_points, nviews, small, dlogit, hv, dhv, dacc9, scale_def, amax_out, s);
  return cudaGetLastError();
}

cudaError_t launch_grad_amax(int64_t n_points, int nviews, const void* packed, const float* dlogit, const float* dsig,
                             uint32_t* amax_logit, uint32_t* amax_sigma, cudaStream_t s) {
  if (n_points == 0) return cudaSuccess;
  k_grad_amax<<<4 * 148, 256, 0, s>>>(n_points * nviews * 4, n_points, dlogit, dsig, reinterpret_cast<const float*>(packed), amax_logit, amax_sigma);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Values of the four training losses from the forward outputs (LossGrad / LossFwdArgs, kernels.h): one warp per ray
// accumulates the ray's terms, a block writes its nine partial sums, k_fused_losses_final adds the blocks in a fixed
// order (bit-reproducible) and applies the means and the weights of LossComputer01.compute_losses (:33-51).
__global__ void __launch_bounds__(128) k_fused_losses_partial(LossFwdArgs a, LossGrad L, int64_t n_rays, float* __restrict__ partial) {
  __shared__ float sh[4][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + warp;
  float t[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) t[i] = 0.f;
  if (ray < n_rays) {
    const bool m_nerf = L.mask_nerf ? L.mask_nerf[ray] != 0 : true;
    const bool m_depth = L.mask_depth ? L.mask_depth[ray] != 0 : false;
    if (lane < 3 && m_nerf) {            // MSE01.py:52-58: sum_c (rgb - target)^2
      const float tg = L.target_rgb[3 * ray + lane];
      if (a.rgb_c) { const float d = a.rgb_c[3 * ray + lane] - tg; t[0] = d * d; }
      if (a.rgb_f) { const float d = a.rgb_f[3 * ray + lane] - tg; t[1] = d * d; }
    }
    if (a.pred_c) for (int i = lane; i < a.Sc; i += 32) t[2] += fabsf(a.pred_c[ray * a.Sc + i] - a.trans_c[ray * a.Sc + i]);   // VisibilityLoss01.py:70-74
    if (a.pred_f) for (int i = lane; i < a.Sf; i += 32) t[3] += fabsf(a.pred_f[ray * a.Sf + i] - a.trans_f[ray * a.Sf + i]);
    if (m_nerf && lane < a.V) {          // VisibilityPriorLoss01.py:76-80: sum_v prior (1 - visibility2)
      const float pr = L.prior ? L.prior[ray * a.V + lane] : 1.f;
      if (a.vis2_c) t[4] = pr * (1.f - a.vis2_c[ray * a.V + lane]);
      if (a.vis2_f) t[5] = pr * (1.f - a.vis2_f[ray * a.V + lane]);
    }
    if (lane == 0) {
      if (m_depth && a.depth) { const float d = a.depth[ray] - L.sparse_depth[ray]; t[6] = d * d; }   // SparseDepthMSE01.py:60-63
      t[7] = m_nerf ? 1.f : 0.f;
      t[8] = m_depth ? 1.f : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float v = warp_sum(t[i]);
    if (lane == 0) sh[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) partial[(int64_t)blockIdx.x * 16 + threadIdx.x] = (sh[0][threadIdx.x] + sh[1][threadIdx.x]) + (sh[2][threadIdx.x] + sh[3][threadIdx.x]);
}
__global__ void k_fused_losses_final(const float* __restrict__ partial, int n_blocks, LossFwdArgs a, LossGrad L, int64_t n_rays,
                                     float* __restrict__ out) {
  __shared__ double tot[9];
  if (threadIdx.x < 9) {
    double acc = 0.0;
    for (int b = 0; b < n_blocks; ++b) acc += (double)partial[(int64_t)b * 16 + threadIdx.x];
    tot[threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n_nerf = tot[7], n_depth = tot[8];
    const double mse = n_nerf > 0 ? (tot[0] + tot[1]) / (3.0 * n_nerf) : 0.0;
    double vis = 0.0;
    // mean|pred - sg(T)| + mean|sg(pred) - T|: the same value twice (the detaches only split the gradient, :57-58)
    if (a.pred_c) vis += 2.0 * tot[2] / ((double)n_rays * a.Sc);
    if (a.pred_f) vis += 2.0 * tot[3] / ((double)n_rays * a.Sf);
    const double prior = (n_nerf > 0 && a.V > 0) ? (tot[4] + tot[5]) / n_nerf : 0.0;
    const double depth = n_depth > 0 ? tot[6] / n_depth : 0.0;
    out[0] = (float)mse; out[1] = (float)vis; out[2] = (float)prior; out[3] = (float)depth;
    out[4] = (float)(L.w_mse * mse + L.w_vis * vis + L.w_prior * prior + L.w_depth * depth);
    out[5] = (float)n_nerf; out[6] = (float)n_depth; out[7] = 0.f;
  }
}

cudaError_t launch_fused_losses(const LossFwdArgs& a, const LossGrad& lg, int64_t n_rays, float* losses_dev, float* partial,
                                cudaStream_t s) {
  const int n_blocks = (int)((n_rays + 3) / 4);
  if (n_blocks > 0) k_fused_losses_partial<<<n_blocks, 128, 0, s>>>(a, lg, n_rays, partial);
  k_fused_losses_final<<<1, 32, 0, s>>>(partial, n_blocks, a, lg, n_rays, losses_dev);
  return cudaGetLastError();
}

cudaError_t launch_composite_bwd(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                                 const float* sigma, const float* rgb, const float* vis, const float* vis2,
                                 const PassGradPtrs& g, const LossGrad& lg, float* dsig, float* dlogit, cudaStream_t s) {
  if (n_rays == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n_rays + 3) / 4);
  const int spl = (S + 31) / 32;
#define VIPNERF_LAUNCH_CBWD(SPL) \
  k_composite_bwd<SPL><<<grid, 128, 0, s>>>(rp, fl, n_rays, S, z, sigma, rgb, vis, vis2, g, lg, dsig, dlogit)
  if (spl <= 2) VIPNERF_LAUNCH_CBWD(2);
  else if (spl <= 6) VIPNERF_LAUNCH_CBWD(6);
  else VIPNERF_LAUNCH_CBWD(8);
#undef VIPNERF_LAUNCH_CBWD
  return cudaGetLastError();
}

size_t gemm_tn_partial_floats() { return (size_t)kGemmMaxMainFloats + kGemmMaxBiasFloats; }

cudaError_t launch_gemm_tn(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t n_rows, float* dst,
                           int ldc, int n_valid, float* bias_dst, float* partial, cudaStream_t s) {
  if ((M != 128 && M != 256) || (N != 32 && N != 64 && N != 128 && N != 256)) return cudaErrorInvalidValue;
  const int BN = N >= 128 ? 128 : N;
  const int tiles = (M / kGemmBM) * (N / BN);
  int64_t n_split = kGemmTargetCtas / tiles;
  const int64_t max_by_rows = (n_rows + 255) / 256;
  if (n_split > max_by_rows) n_split = max_by_rows;
  if (n_split < 1) n_split = 1;
  while (n_split * M * N > kGemmMaxMainFloats) --n_split;
  int64_t rows_per_split = (n_rows + n_split - 1) / n_split;
  rows_per_split = (rows_per_split + kGemmBK - 1) / kGemmBK * kGemmBK;
  if (rows_per_split < kGemmBK) rows_per_split = kGemmBK;
  float* bias_partial = bias_dst ? partial + kGemmMaxMainFloats : nullptr;
  const dim3 grid(M / kGemmBM, N / BN, (unsigned)n_split);
  cudaError_t e;
#define VIPNERF_LAUNCH_GEMM(BNV)                                                                                     \
  do {                                                                                                               \
    e = cudaFuncSetAttribute(k_gemm_tn<BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<BNV>()); \
    if (e != cudaSuccess) return e;                                                                                  \
    k_gemm_tn<BNV><<<grid, 256, gemm_smem_bytes<BNV>(), s>>>(A, lda, B, ldb, n_rows, rows_per_split, M, N, partial,  \
                                                             bias_partial);                                          \
  } while (0)
  if (BN == 128) VIPNERF_LAUNCH_GEMM(128);
  else if (BN == 64) VIPNERF_LAUNCH_GEMM(64);
  else VIPNERF_LAUNCH_GEMM(32);
#undef VIPNERF_LAUNCH_GEMM
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_reduce_partials<<<(M * n_valid + 63) / 64, 256, 0, s>>>(partial, (int)n_split, M, N, dst, ldc, n_valid, nullptr);
  if (bias_dst) k_reduce_partials<<<(M + 63) / 64, 256, 0, s>>>(bias_partial, (int)n_split, M, 1, bias_dst, 1, 1, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_reduce_jobs(const ReduceJob* jobs, int n_jobs, cudaStream_t s) {
  if (n_jobs < 1 || n_jobs > 2 * kGemmGroupMax) return cudaErrorInvalidValue;
  ReduceJobs all{};
  int max_blocks = 1;
  for (int i = 0; i < n_jobs; ++i) {
    all.job[i] = jobs[i];
    const int blocks = (jobs[i].M * jobs[i].n_valid + 63) / 64;
    if (blocks > max_blocks) max_blocks = blocks;
  }
  k_reduce_partials_jobs<<<dim3((unsigned)max_blocks, (unsigned)n_jobs), 256, 0, s>>>(all);
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const float* partial, int n_split, int M, int N, float* dst, int ldc, int n_valid,
                                   cudaStream_t s, const uint32_t* scale_def) {
  k_reduce_partials<<<(M * n_valid + 63) / 64, 256, 0, s>>>(partial, n_split, M, N, dst, ldc, n_valid, scale_def);
  return cudaGetLastError();
}

// bias_dst[m] = sum_p A[p][m] alone (the tensor-core GEMM does not produce the column sums): A viewed as G with M
// columns against a one-column H of ones is what k_small_tn's gsum path computes - here a dedicated streaming kernel.
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ A, int lda, int M, int64_t n_rows,
                                                int64_t rows_per_split, float* __restrict__ partial) {
  const int m = threadIdx.x;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_split;
  const int64_t r_end = min(n_rows, r_begin + rows_per_split);
  float acc = 0.f;
  if (m < M) {
#pragma unroll 8
    for (int64_t r = r_begin; r < r_end; ++r) acc += A[r * lda + m];
    partial[(size_t)blockIdx.x * M + m] = acc;
  }
}

cudaError_t launch_colsum(const float* A, int lda, int M, int64_t n_rows, float* dst, float* partial, cudaStream_t s) {
  if (M < 1 || M > 256) return cudaErrorInvalidValue;
  int64_t n_split = 4 * 148;
  const int64_t max_by_rows = (n_rows + 63) / 64;
  if (n_split > max_by_rows) n_split = max_by_rows;
  if (n_split < 1) n_split = 1;
  const int64_t rows_per_split = (n_rows + n_split - 1) / n_split;
  k_colsum<<<(unsigned)n_split, 256, 0, s>>>(A, lda, M, n_rows, rows_per_split, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_reduce_partials(partial, (int)n_split, M, 1, dst, 1, 1, s);
}

cudaError_t launch_small_tn(const float* G, int M, const void* H, int N, int64_t n_rows, float* dst, float* gsum_dst,
                            float* partial, cudaStream_t s, bool half_h) {
  if ((M != 1 && M != 4) || (N != 128 && N != 256)) return cudaErrorInvalidValue;   // a row = 64 or 128 column pairs
  int64_t n_split = 6 * 148;       // 6 blocks of 256 threads per SM (33 KiB of shared memory each)
  const int64_t max_by_rows = (n_rows + 63) / 64;
  if (n_split > max_by_rows) n_split = max_by_rows;
  if (n_split < 1) n_split = 1;
  const int64_t rows_per_split = (n_rows + n_split - 1) / n_split;
  float* gsum_partial = partial + kGemmMaxMainFloats;
  const float* Hf = static_cast<const float*>(H);
  const __half* Hh = static_cast<const __half*>(H);
  if (M == 1 && !half_h) k_small_tn<1, float><<<(unsigned)n_split, 256, 0, s>>>(G, Hf, N, n_rows, rows_per_split, partial, gsum_partial);
  else if (M == 1) k_small_tn<1, __half><<<(unsigned)n_split, 256, 0, s>>>(G, Hh, N, n_rows, rows_per_split, partial, gsum_partial);
  else if (!half_h) k_small_tn<4, float><<<(unsigned)n_split, 256, 0, s>>>(G, Hf, N, n_rows, rows_per_split, partial, gsum_partial);
  else k_small_tn<4, __half><<<(unsigned)n_split, 256, 0, s>>>(G, Hh, N, n_rows, rows_per_split, partial, gsum_partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_reduce_partials<<<(M * N + 63) / 64, 256, 0, s>>>(partial, (int)n_split, M, N, dst, N, N, nullptr);
  if (gsum_dst) k_reduce_partials<<<1, 256, 0, s>>>(gsum_partial, (int)n_split, M, 1, gsum_dst, 1, 1, nullptr);
  return cudaGetLastError();
}

}  // namespace vipnerf
