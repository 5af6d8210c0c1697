// Internal launch interface between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/vipnerf.h"
#include "stages.cuh"

namespace vipnerf {

// What every kernel needs to know about the rays of a launch (device pointers).
struct RayPtrs {
  const float* rays_o;
  const float* rays_d;
  const float* view_dirs;
  const float* pts_o;     // origins used for sample points: rays_o (world) or rays_o_ndc
  const float* pts_d;     // directions used for sample points and for delta scaling
  const float* near;      // the pair matching pts_*: near/far or near_ndc/far_ndc
  const float* far;
  const float* rays_o2;   // [R,V,3] or null
  const float* t_vals;
  const float* u_vals;
  const float* t_rand;
  const float* u_rand;
};

struct RenderFlags {
  bool ndc, white_bkgd, lindisp;
  int n_sec_views;
};

// stage_kernels.cu
cudaError_t launch_pack_weights(int precision, const float* const params_dev[24], void* packed, cudaStream_t s);
cudaError_t launch_coarse_z(const RayPtrs& rp, int64_t n_rays, int n_coarse, bool lindisp, float* z, cudaStream_t s);
cudaError_t launch_composite(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                             const float* sigma, const float* rgb, const float* vis2, const PassOutPtrs& out,
                             int n_fine, float* z_fine_out, cudaStream_t s);

// Training: what the forward of one sample set keeps for its backward (point-major fp32, P = R*S points,
// nviews = 1 + V view directions per point).
struct MlpSave {
  const float* noise;  // [P] raw_noise_std * randn added to the density logit (VipNeRF01.py:549-552), or null
  float* enc;          // [P][64]   positional encoding (column 63 = 0)
  float* h;            // [8][P][256] outputs of pts_linears.0..7 (post-ReLU)
  float* feat;         // [P][256]  feature_linear output
  float* hv;           // [P][nviews][128] views_linears.0 output (post-ReLU) per view
  float* pev;          // [P][nviews][32]  view-direction encodings (27 used)
};
struct MlpBwdArgs {
  int64_t n_points;
  int nviews;
  const float* dsig;    // [P]            gradient w.r.t. the density logit (after the ReLU mask)
  const float* dlogit;  // [P][nviews][4] gradient w.r.t. the views_output_linear logits
  const float* h;       // saved activations
  const float* hv;
  float* dpre;          // [8][P][256] pre-activation gradients of pts_linears.0..7
  float* dfeat;         // [P][256]    gradient w.r.t. the feature vector
  float* dacc9;         // [P][128]    pre-activation gradient of views_linears.0 summed over views
  float* dhv;           // [P][nviews][128] the same per view
};

// mlp_fp32.cu : CUDA-core evaluation of the MLP on R*S sample points (pts = pts_o + pts_d * z); `save` != null =
// training forward.  launch_mlp_bwd_fp32 = the backward-data chain of the same points.
cudaError_t launch_mlp_fp32(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                            const void* packed, float* sigma, float* rgb, float* vis, float* vis2, cudaStream_t s,
                            const MlpSave* save = nullptr);
cudaError_t launch_mlp_bwd_fp32(const MlpBwdArgs& a, const void* packed, cudaStream_t s);

// train_kernels.cu : backward of volume_rendering and the parameter-gradient reductions
struct PassGradPtrs {  // upstream gradients of one sample set's outputs (null = zero)
  const float *rgb, *acc, *depth, *depth_var, *depth_ndc, *depth_var_ndc, *visibility2;
  const float *alpha, *visibility, *weights, *raw_sigma, *raw_rgb, *raw_visibility, *raw_visibility2;
};
// The reference's four training losses fused into the compositing backward (loss_functions/MSE01.py:25-67,
// VisibilityLoss01.py:26-74 with its mutual detach, VisibilityPriorLoss01.py:25-89, SparseDepthMSE01.py:26-71,
// weighted as LossComputer01.py:33-69 does): k_composite_bwd forms dTotalLoss/d(output) analytically from the forward
// values it re-computes anyway, so no [R,S] gradient tensor is written to or read from HBM.  enabled = 0: off.
struct LossGrad {
  const float* target_rgb;      // [R,3]
  const uint8_t* mask_nerf;     // [R] bool, null = every ray
  const uint8_t* mask_depth;    // [R] bool, null = no sparse-depth term
  const float* sparse_depth;    // [R]
  const float* prior;           // [R,V] visibility prior masks / weights, null = ones
  const float* stats;           // device [8]: [5] = number of nerf rays, [6] = number of sparse-depth rays (launch_fused_losses)
  const float* upstream;        // device scalar dL/dTotalLoss, null = 1
  float w_mse, w_vis, w_prior, w_depth;
  int depth_here;               // the sparse-depth loss reads THIS pass's depth map
  int enabled;
};
cudaError_t launch_composite_bwd(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                                 const float* sigma, const float* rgb, const float* vis, const float* vis2,
                                 const PassGradPtrs& g, const LossGrad& lg, float* dsig, float* dlogit, cudaStream_t s);
// frame_kernels.cu : training batches from the per-pixel caches (vipnerf_gather_train_batch)
constexpr int kMaxGatherColumns = 24;
cudaError_t launch_gather_train_batch(const int64_t* indices, const uint8_t* row_class, int64_t n_rows,
                                      const vipnerf_gather_column* columns, int n_columns, cudaStream_t s);
// Loss values of one training batch from the forward outputs: losses_dev[0..4] = MSE, Visibility, VisibilityPrior,
// SparseDepth, TotalLoss (weighted); [5], [6] = the mask counts the means divide by.  `partial`: 8 floats per 4 rays.
struct LossFwdArgs {
  const float *rgb_c, *rgb_f, *pred_c, *pred_f, *trans_c, *trans_f, *vis2_c, *vis2_f, *depth;   // depth: the pass SparseDepthMSE reads
  int Sc, Sf, V;
};
cudaError_t launch_fused_losses(const LossFwdArgs& a, const LossGrad& lg, int64_t n_rays, float* losses_dev, float* partial,
                                cudaStream_t s);
// C[m][n] = sum_p A[p][m] * B[p][n] over n_rows rows (A: lda floats per row, M in {128, 256}; B: ldb floats per row,
// N in {32, 64, 128, 256}), written to dst[m * ldc + n] for n < n_valid; bias_dst[m] = sum_p A[p][m] when non-null.
// `partial` must hold gemm_tn_partial_floats() floats.
size_t gemm_tn_partial_floats();
cudaError_t launch_gemm_tn(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t n_rows, float* dst,
                           int ldc, int n_valid, float* bias_dst, float* partial, cudaStream_t s);
constexpr int kGemmGroupMax = 8;   // products per grouped launch (launch_gemm_tn_tc_group)
// several reductions in one launch (the weights and bias reductions behind a grouped product launch)
struct ReduceJob {
  const float* partial; int n_split, M, N;
  float* dst; int ldc, n_valid;
  const uint32_t* scale_def;
};
cudaError_t launch_reduce_jobs(const ReduceJob* jobs, int n_jobs, cudaStream_t s);
// scale_def (fp16 training mode): amax slot that defines the power-of-two scale the partial sums carry; dst = sum / scale
cudaError_t launch_reduce_partials(const float* partial, int n_split, int M, int N, float* dst, int ldc, int n_valid,
                                   cudaStream_t s, const uint32_t* scale_def = nullptr);
cudaError_t launch_colsum(const float* A, int lda, int M, int64_t n_rows, float* dst, float* partial, cudaStream_t s);
// gemm_tc.cu : the same product on the tensor cores - tcgen05 kind::tf32 reading fp32 arrays through TMA (N in {32, 64,
// 256}), or kind::f16 reading fp16 arrays (`half`, N in {64, 256}; lda / ldb in elements) - M in {128, 256}.
// `partial`: gemm_tn_tc_partial_floats(#SMs) floats.  scale_def: see launch_reduce_partials.
constexpr int kGemmTcMaxSplits = 160;   // point ranges (one CTA each) the scratch of the tensor-core product is sized for
size_t gemm_tn_tc_partial_floats(int sms);
// bias_dst (optional): column sums of A (db), accumulated by the kernel's epilogue warps from the A boxes in shared
// memory while the main loop runs; colsum_scratch: kGemmTcMaxSplits * M floats.
cudaError_t launch_gemm_tn_tc(const void* A, int lda, int M, const void* B, int ldb, int N, int64_t n_rows, float* dst,
                              int ldc, int n_valid, float* partial, cudaStream_t s, float* bias_dst = nullptr,
                              float* colsum_scratch = nullptr, bool half = false, const uint32_t* scale_def = nullptr);
// Up to kGemmGroupMax products of the same M x N over the same n_rows in ONE launch (CTAs split between the problems):
// the eight 256 x 256 products of a sample set leave 18 instead of 148 partial tiles each.
struct GemmProblem {
  const void* A; int lda;
  const void* B; int ldb;
  float* dst; int ldc; int n_valid;
  float* bias_dst;              // column sums of A, or null
  const uint32_t* scale_def;    // see launch_reduce_partials
};
cudaError_t launch_gemm_tn_tc_group(const GemmProblem* problems, int n_problems, int M, int N, int64_t n_rows,
                                    float* partial, float* colsum_scratch, bool half, cudaStream_t s);
// train_kernels.cu : the non-product pieces of the tensor-core training path (encodings, heads forward / backward).
// `half` = the fp16 mode (VIPNERF_FLAG_TRAIN_F16): enc / pev / hv / dhv / dacc9 are fp16 arrays and a row of pev has 64
// columns (128 bytes, 27 used) instead of 32 floats; gradients carry the power-of-two scale of their amax slot.
cudaError_t launch_encode_points(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                                 void* enc, void* pev, cudaStream_t s, bool half = false);
cudaError_t launch_heads_fwd(int64_t n_points, int nviews, const void* packed, const float* h7, const float* acc9,
                             const void* pev, const float* noise, float* sigma, float* rgb, float* vis, float* vis2,
                             void* hv, cudaStream_t s, bool half = false);
// scale_def: amax slot defining the scale of dhv / dacc9 (fp16 mode); amax_out: records max |dacc9|
cudaError_t launch_heads_bwd(int64_t n_points, int nviews, const void* packed, const float* dlogit, const void* hv,
                             void* dhv, void* dacc9, cudaStream_t s, bool half = false, const uint32_t* scale_def = nullptr,
                             uint32_t* amax_out = nullptr);
// fp16 mode: amax[slot_logit] = max |dlogit|, amax[slot_sigma] = max |dsig| * max |w_sigma| (the rank-1 term that joins
// the chain at pts_linears.7); both slots must have been zeroed
cudaError_t launch_grad_amax(int64_t n_points, int nviews, const void* packed, const float* dlogit, const float* dsig,
                             uint32_t* amax_logit, uint32_t* amax_sigma, cudaStream_t s);
// gemm_tc.cu : one linear layer of the training chains on the tensor cores (tcgen05 kind::tf32 on fp32 arrays, or
// kind::f16 on fp16 arrays; K-major operands):
// out[p][n] = epilogue(sum_k x0[p][k] w0[n][k] (+ sum_k x1[p][k] w1[n][k])), epilogue = + bias[n], + rank1_row[p] *
// rank1_col[n], ReLU, ReLU-mask (mask[p][n] > 0), each optional.  k[i] = reduction length of pair i (multiple of 32, of
// 64 for fp16 operands; k[1] may be 0), N in {128, 256}; rows are points; leading dimensions in elements.
struct LinearTcArgs {
  const void* x[2]; int ldx[2];
  const void* w[2]; int ldw[2];
  int k[2];
  int N;
  int64_t n_rows;
  const float* bias;
  const float* rank1_row;
  const float* rank1_col;
  const void* mask; int ld_mask;
  // fp16 chains: the ReLU masks travel as bits, [rows][8] words for a 256-wide layer (32 bytes per row instead of the 512
  // bytes of the saved activation): written by a forward layer (bias + ReLU), read by the backward layer through it
  const uint32_t* relu_bits; uint32_t* relu_bits_out;
  bool relu;
  void* out; int ld_out;
  bool half_in;     // x, w (and mask) are fp16
  bool half_out;    // out is fp16 (fp16 operands only)
  // fp16 gradient chain: amax slots that define the power-of-two scale of x / of out (null = unscaled), and where to
  // record max |out| (un-scaled, float bits; atomicMax)
  const uint32_t* scale_in;
  const uint32_t* scale_out;
  uint32_t* amax_out;
  // optional rank-1 reduction of the OUTPUT rows: dot_out[p] = sum_n Y[p][n] * dot_vec[n] (the density head on h7,
  // VipNeRF01.py:546: it rides along in the epilogue that holds the row instead of re-reading 1 KiB per point)
  const float* dot_vec; float* dot_out;
  // fp16 forward: the density head is finished where the row is held - dot_out[p] = relu(dot + *dot_bias + dot_noise[p])
  // (VipNeRF01.py:546-553; dot_bias null = the raw dot product)
  const float* dot_bias; const float* dot_noise;
  // fp16 forward, views layer (N = 128, bias + ReLU): views_output_linear and the sigmoids ride along (:582-594) -
  // logits[k] = sum_n out[p][n] * head_wout[n * 4 + k] + head_bout[k]; head_rgb[p][3] (null for a secondary view) and
  // head_vis[p * head_vis_stride] receive the sigmoids
  const float* head_wout; const float* head_bout;
  float* head_rgb; float* head_vis; int head_vis_stride;
};
cudaError_t launch_linear_tc(const LinearTcArgs& a, cudaStream_t s);
// out[m][n] = sum_p G[p][m] * H[p][n] for M <= 4 (G: M floats per row), N <= 256; gsum_dst[m] = sum_p G[p][m].
// half_h: H is an fp16 array.
cudaError_t launch_small_tn(const float* G, int M, const void* H, int N, int64_t n_rows, float* dst, float* gsum_dst,
                            float* partial, cudaStream_t s, bool half_h = false);

#ifdef __CUDACC__
// fp16 training mode: a gradient array is stored as value * 2^k with k chosen so that `amax` (float bits of a non-negative
// maximum measured on the device) lands in [16, 32): 11 binades of head room to fp16's 65504 (conversions saturate), and
// everything down to 2^-28 of amax stays representable.  0 / inf / nan: unscaled.  (Host-callable for the C ABI's
// vipnerf_grad_scale, which the CPU tests hold to this specification.)
__host__ __device__ __forceinline__ float grad_scale_from_amax(uint32_t amax_bits) {
  const uint32_t e = (amax_bits >> 23) & 0xffu;
  float one = 1.f;
  if (amax_bits == 0u || e == 0xffu || (amax_bits >> 31)) return one;
  int k = 4 - ((int)(e == 0u ? 1u : e) - 127);
  k = k < -120 ? -120 : (k > 120 ? 120 : k);
  const uint32_t out = (uint32_t)(k + 127) << 23;
#ifdef __CUDA_ARCH__
  return __uint_as_float(out);
#else
  float f;
  memcpy(&f, &out, sizeof(f));
  return f;
#endif
}
#endif

// mlp_tc.cu : tcgen05 evaluation (precision = BF16 or BF16X3)
cudaError_t launch_mlp_tc(int precision, const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S,
                          const float* z, const void* packed, float* sigma, float* rgb, float* vis, float* vis2,
                          cudaStream_t s);
// mlp_tc.cu : the fused coarse+fine render of a ray batch in one launch
struct FusedArgs {
  RayPtrs rp;
  RenderFlags fl;
  int64_t n_rays;
  int n_coarse, n_fine;
  const void* packed_coarse;
  const void* packed_fine;
  PassOutPtrs out_coarse, out_fine;
  float* ws_z_coarse;   // [R,Nc]      workspace
  float* ws_z_fine;     // [R,Nc+Nf]
  float* ws_sigma;      // [R,Nc+Nf]   network outputs of the pass in flight (reused coarse -> fine)
  float* ws_rgb;        // [R,Nc+Nf,3]
  float* ws_vis;        // [R,Nc+Nf]
  float* ws_sigma_c;    // [R,Nc]      network outputs of the coarse pass (read by the ray warps while fine tiles run)
  float* ws_rgb_c;      // [R,Nc,3]
  float* ws_vis_c;      // [R,Nc]
  float* ws_vis2;       // [R,Nc+Nf,V] secondary-view visibilities per sample (V > 0)
  float* ws_vis2_c;     // [R,Nc,V]
};

cudaError_t launch_render_fused_tc(int precision, const FusedArgs& a, cudaStream_t s);
// frame_kernels.cu : per-pixel ray generation and frame post-processing
cudaError_t launch_generate_rays(const vipnerf_camera& camera, int64_t first_pixel, int64_t n_rays,
                                 const vipnerf_ray_buffers& out, cudaStream_t s);
cudaError_t launch_postprocess_frame(int64_t n_rays, int n_sec_views, const float* rgb, uint8_t* image, int n_depth,
                                     const float* const* depth_in, float* const* depth_out, const float* vis2,
                                     float* vis2_out, cudaStream_t s);
// prior_kernels.cu : plane-sweep-volume visibility weights (the visibility prior generator)
cudaError_t launch_visibility_weights(int h, int w, const uint8_t* frame1, const uint8_t* frame2, const double k1inv[9],
                                      const double t[16], const double k2[9], const double* planes_host, int n_planes,
                                      double temperature, double* weights, uint8_t* mask, cudaStream_t s);
// debug: 64 x u64 device buffer that CTA 0 of the next tensor-core launches fills with cycle counters (null = off)
void set_tc_profile_buffer(void* dev_ptr);

}  // namespace vipnerf
