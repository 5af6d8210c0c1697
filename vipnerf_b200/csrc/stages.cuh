// Device functions for the non-matmul stages of the render path: sample placement, positional encoding,
// alpha compositing and hierarchical re-sampling.  One warp owns one ray; the S samples of the ray are
// spread over the 32 lanes (SPL consecutive samples per lane) and every scan / reduction along the ray is
// a warp-shuffle, so the per-sample intermediates stay in registers.
//
// Arithmetic that the reference evaluates as separate fp32 torch ops is written with the non-contracting
// intrinsics (__fmul_rn / __fadd_rn / ...) so that nvcc cannot fuse it into FMAs: the results then agree with
// the reference bit-for-bit up to transcendental and summation-order differences.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace vipnerf {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Linear (or inverse-linear) sample placement between near and far.
// Reference: get_z_vals_coarse, VipNeRF01.py:186-190.
__device__ __forceinline__ float lerp_depth(float near, float far, float t, bool lindisp) {
  const float omt = fsub(1.f, t);
  if (!lindisp) return fadd(fmul(near, omt), fmul(far, t));
  return fdiv(1.f, fadd(fmul(fdiv(1.f, near), omt), fmul(fdiv(1.f, far), t)));
}

// z of coarse sample i of a ray, including the stratified jitter of training mode (VipNeRF01.py:194-202).
__device__ __forceinline__ float coarse_z_at(float near, float far, const float* __restrict__ t_vals, int i, int n,
                                             bool lindisp, const float* __restrict__ t_rand_row) {
  const float z = lerp_depth(near, far, t_vals[i], lindisp);
  if (t_rand_row == nullptr) return z;
  const float z_prev = i > 0 ? lerp_depth(near, far, t_vals[i - 1], lindisp) : z;
  const float z_next = i < n - 1 ? lerp_depth(near, far, t_vals[i + 1], lindisp) : z;
  const float upper = i < n - 1 ? fmul(.5f, fadd(z_next, z)) : z;
  const float lower = i > 0 ? fmul(.5f, fadd(z, z_prev)) : z;
  return fadd(lower, fmul(fsub(upper, lower), t_rand_row[i]));
}

// NDC depth -> metric depth.  Reference: convert_depth_from_ndc, VipNeRF01.py:386-403.
__device__ __forceinline__ float depth_from_ndc(float z_ndc, float oz, float dz) {
  const float tn = fdiv(-fadd(1.f, oz), dz);
  const float c = z_ndc == 1.f ? 1e-3f : 0.f;
  const float a = fdiv(fadd(oz, fmul(tn, dz)), dz);
  const float b = fsub(fdiv(1.f, fadd(fsub(1.f, z_ndc), c)), 1.f);
  return fadd(fmul(a, b), tn);
}

// NDC depth -> metric depth as used for the secondary-view directions (note: +1e-6, different rule).
// Reference: compute_other_view_dirs, VipNeRF01.py:219-222.
__device__ __forceinline__ float depth_from_ndc_secondary(float z_ndc, float oz, float dz) {
  const float tn = fdiv(-fadd(1.f, oz), dz);
  const float num = fadd(oz, fmul(tn, dz));
  return fdiv(fsub(fdiv(num, fadd(fsub(1.f, z_ndc), 1e-6f)), oz), dz);
}

// Unit vector from secondary camera centre o2 to the sample at depth z along (o, d).
// Reference: compute_other_view_dirs, VipNeRF01.py:223-225.
__device__ __forceinline__ void secondary_view_dir(const float o[3], const float d[3], float z, const float o2[3],
                                                   float out[3]) {
  float v[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) v[a] = fsub(fadd(o[a], fmul(z, d[a])), o2[a]);
  const float n = sqrtf(fadd(fadd(fmul(v[0], v[0]), fmul(v[1], v[1])), fmul(v[2], v[2])));
#pragma unroll
  for (int a = 0; a < 3; ++a) out[a] = fdiv(v[a], n);
}

__device__ __forceinline__ float vec3_norm(float x, float y, float z) {
  return sqrtf(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
}

// Sinusoidal encoding of one scalar: writes x, then (sin(2^k x), cos(2^k x)) for k < L at out[0], out[stride*(1+2k)],
// out[stride*(2+2k)] - the caller interleaves the three axes (layout of PositionalEncoder.encode,
// VipNeRF01.py:439-448: [x(3), sin(f0 x)(3), cos(f0 x)(3), ...]).  Multiplying by a power of two is exact,
// so sincosf(x * 2^k) sees the same argument as torch.sin(x * freq).
template <int L, typename Store>
__device__ __forceinline__ void encode_axis(float x, int axis, Store&& store) {
  store(axis, x);
#pragma unroll
  for (int k = 0; k < L; ++k) {
    float s, c;
    sincosf(x * (float)(1 << k), &s, &c);
    store(3 + 6 * k + axis, s);
    store(6 + 6 * k + axis, c);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-ray constants for compositing
struct RayConsts {
  float dnorm;   // ||rays_d|| (world) or ||rays_d_ndc|| (NDC): scales the sample spacing (:337, :342)
  float oz, dz;  // z components of the world ray (NDC depth conversion only)
};

struct PassOutPtrs {  // device mirror of vipnerf_pass_out
  float *rgb, *acc, *depth, *depth_var, *depth_ndc, *depth_var_ndc, *visibility2;
  float *alpha, *z_vals, *visibility, *weights, *raw_sigma, *raw_rgb, *raw_visibility, *raw_visibility2;
};

// Alpha-composites one ray with one warp.  Reference: volume_rendering, VipNeRF01.py:331-384.
//   z, sigma [S]; rgb [S*3]; vis2 [S*V] (may be null) - any address space;  w_out[SPL]: this lane's weights.
template <int SPL>
__device__ __forceinline__ void composite_ray(int lane, int S, const float* z_in, const float* sigma_in,
                                              const float* rgb_in, const float* vis2_in, int V, bool ndc,
                                              bool white_bkgd, const RayConsts& rc, const PassOutPtrs& out,
                                              int64_t ray, float z_reg[SPL], float w_reg[SPL]) {
  float alpha[SPL], om[SPL], trans[SPL];
  const int base = lane * SPL;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = base + j;
    const bool valid = i < S;
    const float z = valid ? z_in[i] : 0.f;
    const float z_next = (i + 1 < S) ? z_in[i + 1] : (ndc ? 1.f : 1e10f);
    const float delta = fmul(fsub(z_next, z), rc.dnorm);
    const float sg = valid ? sigma_in[i] : 0.f;
    const float a = valid ? fsub(1.f, expf(fmul(-sg, delta))) : 0.f;
    z_reg[j] = z;
    alpha[j] = a;
    om[j] = fadd(fsub(1.f, a), 1e-10f);
  }
  // exclusive product scan along the ray: lane-local products, then a shuffle scan of the lane totals
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
  trans[0] = excl;
#pragma unroll
  for (int j = 1; j < SPL; ++j) trans[j] = trans[j - 1] * om[j - 1];

  float acc = 0.f, r = 0.f, g = 0.f, b = 0.f, wz = 0.f;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = base + j;
    const float w = fmul(alpha[j], trans[j]);
    w_reg[j] = w;
    if (i < S) {
      acc += w;
      r += fmul(w, rgb_in[3 * i + 0]);
      g += fmul(w, rgb_in[3 * i + 1]);
      b += fmul(w, rgb_in[3 * i + 2]);
      wz += fmul(w, z_reg[j]);
    }
  }
  acc = warp_sum(acc);
  r = warp_sum(r);
  g = warp_sum(g);
  b = warp_sum(b);
  wz = warp_sum(wz);
  const float denom = fadd(acc, 1e-6f);
  const float depth_native = fdiv(wz, denom);  // NDC depth in NDC mode, metric depth otherwise
  float var_native = 0.f;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const float dlt = fsub(z_reg[j], depth_native);
    if (base + j < S) var_native += fmul(w_reg[j], fmul(dlt, dlt));
  }
  var_native = warp_sum(var_native);

  float depth_world = depth_native, var_world = var_native;
  if (ndc) {
    float zw[SPL], s1 = 0.f;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      zw[j] = depth_from_ndc(z_reg[j], rc.oz, rc.dz);
      if (base + j < S) s1 += fmul(w_reg[j], zw[j]);
    }
    depth_world = fdiv(warp_sum(s1), denom);
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const float dlt = fsub(zw[j], depth_world);
      if (base + j < S) s2 += fmul(w_reg[j], fmul(dlt, dlt));
    }
    var_world = warp_sum(s2);
  }
  if (white_bkgd) {
    const float bg = fsub(1.f, acc);
    r = fadd(r, bg);
    g = fadd(g, bg);
    b = fadd(b, bg);
  }
  if (lane == 0) {
    if (out.rgb) {
      out.rgb[3 * ray + 0] = r;
      out.rgb[3 * ray + 1] = g;
      out.rgb[3 * ray + 2] = b;
    }
    if (out.acc) out.acc[ray] = acc;
    if (out.depth) out.depth[ray] = depth_world;
    if (out.depth_var) out.depth_var[ray] = var_world;
    if (ndc && out.depth_ndc) out.depth_ndc[ray] = depth_native;
    if (ndc && out.depth_var_ndc) out.depth_var_ndc[ray] = var_native;
  }
  if (vis2_in != nullptr && out.visibility2 != nullptr) {
    for (int v = 0; v < V; ++v) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < SPL; ++j) {
        const int i = base + j;
        if (i < S) s += fmul(w_reg[j], vis2_in[(int64_t)i * V + v]);
      }
      s = warp_sum(s);
      if (lane == 0) out.visibility2[ray * V + v] = fdiv(s, denom);
    }
  }
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = base + j;
    if (i < S) {
      const int64_t o = ray * S + i;
      if (out.alpha) out.alpha[o] = alpha[j];
      if (out.visibility) out.visibility[o] = trans[j];
      if (out.weights) out.weights[o] = w_reg[j];
      if (out.z_vals) out.z_vals[o] = z_reg[j];
    }
  }
}

// number of entries of sorted arr[0..n) that are <= x (torch.searchsorted(..., right=True))
__device__ __forceinline__ int count_le(const float* arr, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (arr[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// number of entries of sorted arr[0..n) that are < x
__device__ __forceinline__ int count_lt(const float* arr, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (arr[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// In-place ascending bitonic sort of buf[0..n) (n a power of two) by one warp.
__device__ __forceinline__ void warp_bitonic_sort(float* buf, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int p = i ^ j;
        if (p > i) {
          const float a = buf[i], b = buf[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { buf[i] = b; buf[p] = a; }
        }
      }
      __syncwarp();
    }
  }
}

// Scratch one warp needs for resample_ray (floats): z_c[Nc] | w[Nc] | cdf[Nc-1] | bins[Nc-1] | samples[pow2(Nf)]
__host__ __device__ constexpr int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
__host__ __device__ constexpr int resample_scratch_floats(int nc, int nf) { return 4 * nc + next_pow2(nf); }

// Hierarchical re-sampling of one ray by one warp: n_fine inverse-CDF samples from the coarse weights, merged
// with the coarse positions into a sorted list of Nc + n_fine depths.
// Reference: get_z_vals_fine VipNeRF01.py:205-216, sample_pdf :229-262.
//   z_reg / w_reg: this lane's SPL coarse depths / weights (from composite_ray);  u: [n_fine] cdf positions
//   (the host's linspace table, or this ray's random row);  sorted_u: u is non-decreasing (deterministic mode).
template <int SPL>
__device__ __forceinline__ void resample_ray(int lane, int Nc, int n_fine, const float z_reg[SPL],
                                             const float w_reg[SPL], const float* __restrict__ u, bool sorted_u,
                                             float* scratch, float* z_fine_out) {
  float* zc = scratch;
  float* wc = zc + Nc;
  float* cdf = wc + Nc;
  float* bins = cdf + (Nc - 1);
  float* samples = bins + (Nc - 1);
  const int B = Nc - 1;   // bins (mid-points) == cdf entries
  const int NW = Nc - 2;  // interior weights
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int i = lane * SPL + j;
    if (i < Nc) { zc[i] = z_reg[j]; wc[i] = w_reg[j]; }
  }
  __syncwarp();
  // pdf over the interior weights (+1e-5), blocked over lanes so the cumulative sum is a lane-local
  // running sum plus a shuffle scan of lane totals
  const int CH = (NW + 31) / 32;
  float tot = 0.f;
  for (int j = 0; j < CH; ++j) {
    const int i = lane * CH + j;
    if (i < NW) tot += fadd(wc[i + 1], 1e-5f);
  }
  const float total = warp_sum(tot);
  float run = 0.f;
  for (int j = 0; j < CH; ++j) {
    const int i = lane * CH + j;
    if (i < NW) run += fdiv(fadd(wc[i + 1], 1e-5f), total);
  }
  float incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  float c = incl - run;  // exclusive prefix of this lane's block
  if (lane == 0) cdf[0] = 0.f;
  for (int j = 0; j < CH; ++j) {
    const int i = lane * CH + j;
    if (i < NW) {
      c += fdiv(fadd(wc[i + 1], 1e-5f), total);
      cdf[i + 1] = c;
    }
  }
  for (int i = lane; i < B; i += 32) bins[i] = fmul(.5f, fadd(zc[i + 1], zc[i]));
  __syncwarp();
  // invert the cdf
  for (int k = lane; k < n_fine; k += 32) {
    const float uk = u[k];
    const int idx = count_le(cdf, B, uk);
    const int below = max(idx - 1, 0);
    const int above = min(idx, B - 1);
    const float cb = cdf[below], ca = cdf[above];
    float denom = fsub(ca, cb);
    if (denom < 1e-5f) denom = 1.f;
    const float t = fdiv(fsub(uk, cb), denom);
    const float bb = bins[below], ba = bins[above];
    samples[k] = fadd(bb, fmul(t, fsub(ba, bb)));
  }
  __syncwarp();
  if (!sorted_u) {
    const int np2 = next_pow2(n_fine);
    for (int k = n_fine + lane; k < np2; k += 32) samples[k] = INFINITY;
    __syncwarp();
    warp_bitonic_sort(samples, np2, lane);
  }
  // merge the two sorted runs by rank (ties: coarse positions first) == values of torch.sort(cat(...))
  for (int i = lane; i < Nc; i += 32) {
    const float v = zc[i];
    z_fine_out[i + count_lt(samples, n_fine, v)] = v;
  }
  for (int k = lane; k < n_fine; k += 32) {
    const float v = samples[k];
    z_fine_out[k + count_le(zc, Nc, v)] = v;
  }
  __syncwarp();
}

}  // namespace vipnerf
