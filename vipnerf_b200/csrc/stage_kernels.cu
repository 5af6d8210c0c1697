// Weight packing, coarse sample placement and the stand-alone compositing / re-sampling kernel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"
#include "layout.cuh"

namespace vipnerf {

struct ParamPtrs { const float* p[24]; };

// ---------------------------------------------------------------------------------------------------
// pack: reference nn.Linear tensors (VipNeRF01.py:472-491, [out,in] row-major fp32) -> kernel layout
// feature_linear folded into views_linears.0 (layout.cuh): bias and weight of the fused 256 -> 128 layer, formed in
// double precision from the fp32 masters (p[16] = views_linears.0.weight [128][283], p[17] its bias,
// p[20] = feature_linear.weight [256][256], p[21] its bias)
__device__ __forceinline__ float fused_views_bias(const ParamPtrs& pp, int n) {
  double acc = (double)pp.p[17][n];
  for (int j = 0; j < kWidth; ++j) acc += (double)pp.p[16][n * (kWidth + kEncView) + j] * (double)pp.p[21][j];
  return (float)acc;
}
__device__ __forceinline__ float fused_views_weight(const ParamPtrs& pp, int n, int k) {
  double acc = 0.0;
  for (int j = 0; j < kWidth; ++j) acc += (double)pp.p[16][n * (kWidth + kEncView) + j] * (double)pp.p[20][j * kWidth + k];
  return (float)acc;
}

// with_fused: the folded views bias (128 double-precision dot products) is read by the tensor-core eval kernels only
__global__ void k_pack_small(ParamPtrs pp, float* __restrict__ small, bool with_fused) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kSmallFloats) return;
  float v = 0.f;
  if (i < kOffBiasViews) {
    const int l = i / 256, n = i % 256;
    v = (l < 8 ? pp.p[2 * l + 1] : pp.p[21])[n];
  } else if (i < kOffWSigma) {
    v = pp.p[17][i - kOffBiasViews];
  } else if (i < kOffBSigma) {
    v = pp.p[18][i - kOffWSigma];
  } else if (i < kOffWViewDir) {
    if (i == kOffBSigma) {
      v = pp.p[19][0];
    } else if (i == kOffBSigma + 1) {   // max |pts_output_linear.weight|: anchors a gradient scale of the fp16 training mode
      for (int k = 0; k < kWidth; ++k) v = fmaxf(v, fabsf(pp.p[18][k]));
    }
  } else if (i < kOffWOut) {
    const int j = (i - kOffWViewDir) / 128, c = (i - kOffWViewDir) % 128;
    v = pp.p[16][c * (kWidth + kEncView) + kWidth + j];
  } else if (i < kOffBOut) {
    const int c = (i - kOffWOut) / 4, k = (i - kOffWOut) % 4;
    v = pp.p[22][k * 128 + c];
  } else if (i < kOffBiasViewsFused) {
    v = pp.p[23][i - kOffBOut];
  } else {
    v = with_fused ? fused_views_bias(pp, i - kOffBiasViewsFused) : 0.f;
  }
  small[i] = v;
}

__device__ __forceinline__ float source_weight(const ParamPtrs& pp, int l, int n, int k) {
  const int sc = source_col(l, k);
  return sc < 0 ? 0.f : pp.p[source_param(l)][(int64_t)n * source_in_features(l) + sc];
}

__global__ void k_pack_fp32(ParamPtrs pp, float* __restrict__ big) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kFp32BigFloats) return;
  int l = 0, off = 0;
  while (l < kNumMatLayers - 1 && i >= off + layer_k(l) * layer_n(l)) { off += layer_k(l) * layer_n(l); ++l; }
  const int N = layer_n(l);
  const int k = (i - off) / N, n = (i - off) % N;
  big[i] = source_weight(pp, l, n, k);
}

// backward-data images of the training path (layout.cuh): original [out][in] rows, hidden input columns only
__global__ void k_pack_fp32_bwd(ParamPtrs pp, float* __restrict__ bwd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kFp32BwdFloats) return;
  float v;
  if (i < kBwdOffFeature) {
    const int o = i / 256, c = i % 256;
    v = pp.p[16][o * (kWidth + kEncView) + c];
  } else if (i < kBwdOffTrunk) {
    v = pp.p[20][i - kBwdOffFeature];
  } else if (i < kBwdOffEnc0) {
    const int e = i - kBwdOffTrunk;
    const int l = 7 - e / 65536, o = (e % 65536) / 256, c = e % 256;
    v = l == 5 ? pp.p[10][o * (kWidth + kEncPts) + kEncPts + c] : pp.p[2 * l][o * 256 + c];
  } else if (i < kBwdOffEnc5) {
    const int o = (i - kBwdOffEnc0) / 64, c = (i - kBwdOffEnc0) % 64;
    v = c < kEncPts ? pp.p[0][o * kEncPts + c] : 0.f;
  } else {
    const int o = (i - kBwdOffEnc5) / 64, c = (i - kBwdOffEnc5) % 64;
    v = c < kEncPts ? pp.p[10][o * (kWidth + kEncPts) + c] : 0.f;
  }
  bwd[i] = v;
}

// fp16 mirror of the two fp32 regions (layout.cuh): weights rounded to nearest, saturating
__global__ void k_pack_f16_mirror(const float* __restrict__ big, __half* __restrict__ mirror) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kF16MirrorHalves) return;
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(big[i]));
  reinterpret_cast<unsigned short*>(mirror)[i] = r;
}

__global__ void k_pack_f16_viewdir(ParamPtrs pp, __half* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kF16ViewDirHalves) return;
  const int n = i / 64, e = i % 64;
  const float w = e < kEncView ? pp.p[16][n * (kWidth + kEncView) + kWidth + e] : 0.f;
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(w));
  reinterpret_cast<unsigned short*>(dst)[i] = r;
}

template <bool kSplit3, bool kHalf = false>
__global__ void k_pack_tc(ParamPtrs pp, uint8_t* __restrict__ big) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one bf16 element of the hi image set
  if (i >= kTcBigBytes / 2) return;
  int l = 0;
  while (l < kNumTcLayers - 1 && 2 * i >= tc_layer_byte_offset(l + 1)) ++l;
  const int N = layer_n(l);
  const int e = i - tc_layer_byte_offset(l) / 2;          // element index inside the layer
  const int chunk = e / (N * kChunkK), r = e % (N * kChunkK);
  const int n = r / kChunkK, k_local = r % kChunkK;
  float w;
  if (l == kViewChunkLayer) {                             // view-direction columns of views_linears.0, bias in column 31
    w = k_local < kEncView ? pp.p[16][n * (kWidth + kEncView) + kWidth + k_local]
                           : (k_local == kChunkK - 1 ? fused_views_bias(pp, n) : 0.f);
  } else if (chunk == layer_chunks(l)) {                         // the layer's bias chunk: column 31 <-> encoding column 63
    w = k_local == kChunkK - 1 ? pp.p[bias_param(l)][n] : 0.f;
  } else {
    const int k = chunk * kChunkK + k_local;
    const bool bias_col = (l == 0 || l == 5) && k == 63;  // the encoding block's constant-one column
    w = bias_col ? pp.p[bias_param(l)][n] : (l == 9 ? fused_views_weight(pp, n, k) : source_weight(pp, l, n, k));
  }
  const uint32_t byte = n * 64 + ((((k_local >> 3) ^ ((n >> 1) & 3))) << 4) + (k_local & 7) * 2;
  const size_t chunk_bytes = (size_t)layer_chunk_bytes(l);
  const size_t base = (size_t)tc_layer_byte_offset(l) * (kSplit3 ? 2 : 1);
  if (kHalf) {   // fp16 image (VIPNERF_PRECISION_FP16): saturate instead of overflowing to inf
    *reinterpret_cast<__half*>(big + base + chunk * chunk_bytes + byte) = __float2half_rn(fminf(fmaxf(w, -65504.f), 65504.f));
    return;
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  if (!kSplit3) {
    *reinterpret_cast<__nv_bfloat16*>(big + base + chunk * chunk_bytes + byte) = hi;
  } else {
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(big + base + (2 * chunk) * chunk_bytes + byte) = hi;
    *reinterpret_cast<__nv_bfloat16*>(big + base + (2 * chunk + 1) * chunk_bytes + byte) = lo;
  }
}

cudaError_t launch_pack_weights(int precision, const float* const params_dev[24], void* packed, cudaStream_t s) {
  ParamPtrs pp;
  for (int i = 0; i < 24; ++i) pp.p[i] = params_dev[i];
  float* small = reinterpret_cast<float*>(packed);
  uint8_t* big = reinterpret_cast<uint8_t*>(packed) + kSmallBytes;
  k_pack_small<<<(kSmallFloats + 255) / 256, 256, 0, s>>>(pp, small, precision != VIPNERF_PRECISION_FP32);
  if (precision == VIPNERF_PRECISION_FP32) {
    k_pack_fp32<<<(kFp32BigFloats + 255) / 256, 256, 0, s>>>(pp, reinterpret_cast<float*>(big));
    k_pack_fp32_bwd<<<(kFp32BwdFloats + 255) / 256, 256, 0, s>>>(pp, reinterpret_cast<float*>(big) + kFp32BigFloats);
    __half* mirror = reinterpret_cast<__half*>(reinterpret_cast<float*>(big) + kFp32BigFloats + kFp32BwdFloats);
    k_pack_f16_mirror<<<(kF16MirrorHalves + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float*>(big), mirror);
    k_pack_f16_viewdir<<<(kF16ViewDirHalves + 255) / 256, 256, 0, s>>>(pp, mirror + kF16MirrorHalves);
  } else {
    const int n = kTcBigBytes / 2;
    if (precision == VIPNERF_PRECISION_BF16X3) k_pack_tc<true><<<(n + 255) / 256, 256, 0, s>>>(pp, big);
    else if (precision == VIPNERF_PRECISION_FP16) k_pack_tc<false, true><<<(n + 255) / 256, 256, 0, s>>>(pp, big);
    else k_pack_tc<false><<<(n + 255) / 256, 256, 0, s>>>(pp, big);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// get_z_vals_coarse (VipNeRF01.py:173-203)
__global__ void k_coarse_z(const float* __restrict__ near, const float* __restrict__ far,
                           const float* __restrict__ t_vals, const float* __restrict__ t_rand, int64_t n_rays,
                           int n, bool lindisp, float* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * n) return;
  const int64_t ray = i / n;
  const int s = (int)(i % n);
  z[i] = coarse_z_at(near[ray], far[ray], t_vals, s, n, lindisp, t_rand ? t_rand + ray * n : nullptr);
}

cudaError_t launch_coarse_z(const RayPtrs& rp, int64_t n_rays, int n_coarse, bool lindisp, float* z, cudaStream_t s) {
  const int64_t total = n_rays * n_coarse;
  if (total == 0) return cudaSuccess;
  k_coarse_z<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(rp.near, rp.far, rp.t_vals, rp.t_rand, n_rays, n_coarse,
                                                             lindisp, z);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// volume_rendering (+ get_z_vals_fine): one warp per ray, four rays per block
template <int SPL>
__global__ void __launch_bounds__(128) k_composite(RayPtrs rp, RenderFlags fl, int64_t n_rays, int S,
                                                   const float* __restrict__ z, const float* __restrict__ sigma,
                                                   const float* __restrict__ rgb, const float* __restrict__ vis2,
                                                   PassOutPtrs out, int n_fine, float* __restrict__ z_fine_out) {
  extern __shared__ float scratch_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + warp;
  if (ray >= n_rays) return;
  RayConsts rc;
  rc.dnorm = vec3_norm(rp.pts_d[3 * ray], rp.pts_d[3 * ray + 1], rp.pts_d[3 * ray + 2]);
  rc.oz = rp.rays_o[3 * ray + 2];
  rc.dz = rp.rays_d[3 * ray + 2];
  float z_reg[SPL], w_reg[SPL];
  composite_ray<SPL>(lane, S, z + ray * S, sigma + ray * S, rgb + ray * S * 3,
                     vis2 ? vis2 + ray * S * fl.n_sec_views : nullptr, fl.n_sec_views, fl.ndc, fl.white_bkgd, rc, out,
                     ray, z_reg, w_reg);
  if (z_fine_out != nullptr) {
    float* scratch = scratch_all + warp * resample_scratch_floats(S, n_fine);
    const float* u = rp.u_rand ? rp.u_rand + ray * n_fine : rp.u_vals;
    resample_ray<SPL>(lane, S, n_fine, z_reg, w_reg, u, rp.u_rand == nullptr, scratch,
                      z_fine_out + ray * (S + n_fine));
  }
}

cudaError_t launch_composite(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                             const float* sigma, const float* rgb, const float* vis2, const PassOutPtrs& out,
                             int n_fine, float* z_fine_out, cudaStream_t s) {
  if (n_rays == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n_rays + 3) / 4);
  const size_t smem = z_fine_out ? 4 * resample_scratch_floats(S, n_fine) * sizeof(float) : 0;
  const int spl = (S + 31) / 32;
#define VIPNERF_LAUNCH_COMPOSITE(SPL)                                                                          \
  k_composite<SPL><<<grid, 128, smem, s>>>(rp, fl, n_rays, S, z, sigma, rgb, vis2, out, n_fine, z_fine_out)
  if (spl <= 2) VIPNERF_LAUNCH_COMPOSITE(2);
  else if (spl <= 6) VIPNERF_LAUNCH_COMPOSITE(6);
  else VIPNERF_LAUNCH_COMPOSITE(8);
#undef VIPNERF_LAUNCH_COMPOSITE
  return cudaGetLastError();
}

}  // namespace vipnerf
