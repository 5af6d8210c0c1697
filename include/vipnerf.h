/*
 * vipnerf.h - C ABI of the B200-native ViP-NeRF volumetric render path (libvipnerf_b200.so).
 *
 * The reference (NagabhushanSN95/ViP-NeRF) is pure Python; it has no FFI today.  The boundary a maintainer
 * binds is its model plugin API (src/models/ModelFactory.py:10-22 -> VipNeRF.forward,
 * src/models/VipNeRF01.py:34-41).  This header is what a ctypes stub for that path calls
 * (see INTEGRATION.md); every entry point cites the reference function(s) it replaces.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  All array pointers are DEVICE pointers on the
 *    current CUDA device unless a name ends in _host; arrays are fp32, row-major, contiguous, 16-byte aligned.
 *  - ownership: the caller owns every buffer.  The library never allocates device memory, never
 *    synchronises the device and launches only on the `stream` it is given (a cudaStream_t passed as
 *    void*; NULL = legacy default stream).
 *  - errors: 0 on success, a negative vipnerf_status otherwise; vipnerf_last_error() returns a
 *    thread-local message.  Nothing throws or aborts.
 *  - threading: re-entrant; no global mutable state besides a mutex-guarded per-device attribute cache.
 *    (Under torch.nn.DataParallel, src/Tester01.py:42, one Python thread per GPU calls concurrently.)
 */
#ifndef VIPNERF_H_
#define VIPNERF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIPNERF_ABI_VERSION 1

typedef enum vipnerf_status {
  VIPNERF_OK = 0,
  VIPNERF_EINVAL = -1,        /* NULL / misaligned pointer, bad size                       */
  VIPNERF_EUNSUPPORTED = -2,  /* configuration outside the shapes the kernels are built for */
  VIPNERF_ECUDA = -3,         /* a CUDA runtime call failed (message has the CUDA error)    */
  VIPNERF_EWORKSPACE = -4,    /* workspace too small                                        */
  VIPNERF_EABI = -5           /* cfg->abi != VIPNERF_ABI_VERSION                            */
} vipnerf_status;

/* cfg.flags */
#define VIPNERF_FLAG_NDC         (1u << 0) /* configs['data_loader']['ndc']   (VipNeRF01.py:16)          */
#define VIPNERF_FLAG_WHITE_BKGD  (1u << 1) /* configs['model']['white_bkgd']  (VipNeRF01.py:363-364)     */
#define VIPNERF_FLAG_LINDISP     (1u << 2) /* configs['model']['lindisp']     (VipNeRF01.py:183-190)     */
#define VIPNERF_FLAG_TRAIN_TF32  (1u << 3) /* vipnerf_train_forward / _backward: every 256-wide product of the step
                                              (forward chain, backward-data chain, parameter gradients dW = dY^T X) runs
                                              on the tensor cores (tcgen05 kind::tf32: operands rounded to tf32, fp32
                                              accumulate) instead of fp32 CUDA cores                               */
#define VIPNERF_FLAG_TRAIN_F16   (1u << 4) /* the same three product families on tcgen05 kind::f16 with every saved
                                              activation and chain gradient stored as fp16 (11-bit significand like tf32,
                                              half the HBM bytes of every stream; gradients carry per-array power-of-two
                                              scales measured on the device, fp32 accumulate).  Exclusive with TRAIN_TF32 */

/* cfg.precision: arithmetic of the 256-wide matmuls (trunk layers, feature_linear, feature columns of
 * views_linears.0).  Everything else (encodings, heads, compositing, sampling) is always fp32. */
#define VIPNERF_PRECISION_FP32    0 /* CUDA-core FFMA, staged kernels; the bit-for-bit-closest path      */
#define VIPNERF_PRECISION_BF16    1 /* tcgen05.mma kind::f16, bf16 operands, fp32 accumulate in TMEM     */
#define VIPNERF_PRECISION_BF16X3  2 /* tcgen05.mma, hi/lo bf16 split of both operands (3 MMAs / product) */
#define VIPNERF_PRECISION_FP16    3 /* tcgen05.mma kind::f16, fp16 operands (11-bit significand, saturating
                                       at +-65504), fp32 accumulate: the bf16 rate at ~8x smaller rounding   */

typedef struct vipnerf_cfg {
  int32_t abi;          /* VIPNERF_ABI_VERSION */
  int32_t n_coarse;     /* coarse_mlp.num_samples (64)                                  */
  int32_t n_fine;       /* fine_mlp.num_samples (128); 0 = no fine pass                 */
  int32_t l_pts;        /* points_positional_encoding_degree (10)                       */
  int32_t l_view;       /* views_positional_encoding_degree (4)                         */
  int32_t depth;        /* netdepth (8)                                                 */
  int32_t width;        /* netwidth (256)                                               */
  int32_t skip;         /* skip layer index (4, VipNeRF01.py:466)                       */
  int32_t n_sec_views;  /* V = secondary views evaluated per sample (0 = sec_views_vis off) */
  uint32_t flags;       /* VIPNERF_FLAG_*                                               */
  int32_t precision;    /* VIPNERF_PRECISION_*                                          */
  int32_t reserved;
} vipnerf_cfg;

/* Inputs of VipNeRF.render_rays (VipNeRF01.py:74-98; key names of the reference's input dict). */
typedef struct vipnerf_rays {
  const float* rays_o;      /* [R,3] world origins                                            */
  const float* rays_d;      /* [R,3] world directions (not normalised)                        */
  const float* view_dirs;   /* [R,3] unit view directions fed to the view encoder             */
  const float* near;        /* [R,1]  (world path)                                            */
  const float* far;         /* [R,1]                                                          */
  const float* rays_o_ndc;  /* [R,3]  NDC only                                                */
  const float* rays_d_ndc;  /* [R,3]  NDC only                                                */
  const float* near_ndc;    /* [R,1]  NDC only                                                */
  const float* far_ndc;     /* [R,1]  NDC only                                                */
  const float* rays_o2;     /* [R,V,3] secondary camera centres, n_sec_views > 0 only         */
  const float* t_vals;      /* [n_coarse] torch.linspace(0,1,n_coarse) made by the host (:186) */
  const float* u_vals;      /* [n_fine]   torch.linspace(0,1,n_fine) made by the host   (:239) */
  const float* t_rand;      /* [R,n_coarse] stratified jitter (:200) or NULL = deterministic  */
  const float* u_rand;      /* [R,n_fine]   random cdf samples (:242) or NULL = deterministic */
} vipnerf_rays;

/* Outputs of one sample set (coarse or fine) - the keys VipNeRF.volume_rendering returns
 * (VipNeRF01.py:366-383) plus the raw network outputs (:128-133).  NULL = not wanted. */
typedef struct vipnerf_pass_out {
  float* rgb;            /* [R,3]            */
  float* acc;            /* [R]              */
  float* depth;          /* [R]              */
  float* depth_var;      /* [R]              */
  float* depth_ndc;      /* [R]   NDC only   */
  float* depth_var_ndc;  /* [R]   NDC only   */
  float* visibility2;    /* [R,V]            */
  float* alpha;          /* [R,S]            */
  float* z_vals;         /* [R,S]            */
  float* visibility;     /* [R,S] transmittance                       */
  float* weights;        /* [R,S]            */
  float* raw_sigma;      /* [R,S]   (reference shape [R,S,1])         */
  float* raw_rgb;        /* [R,S,3]          */
  float* raw_visibility; /* [R,S]   (reference shape [R,S,1])         */
  float* raw_visibility2;/* [R,S,V] (reference shape [R,S,V,1])       */
} vipnerf_pass_out;

typedef struct vipnerf_out {
  vipnerf_pass_out coarse;  /* S = n_coarse          */
  vipnerf_pass_out fine;    /* S = n_coarse + n_fine */
} vipnerf_out;

int vipnerf_abi_version(void);
const char* vipnerf_last_error(void);

/* 0 if this build can run cfg (else VIPNERF_EUNSUPPORTED with the reason in vipnerf_last_error). */
int vipnerf_check_config(const vipnerf_cfg* cfg);

/* --- weights: replaces MLP.__init__/load_state_dict's fp32 nn.Linear storage (VipNeRF01.py:472-491) with the
 * kernel's packed layout.  params[24] = the tensors of ONE MLP in the reference's state_dict order:
 * pts_linears.{0..7}.{weight,bias}, views_linears.0.{weight,bias}, pts_output_linear.{weight,bias},
 * feature_linear.{weight,bias}, views_output_linear.{weight,bias} (device pointers, fp32). */
size_t vipnerf_packed_weight_bytes(const vipnerf_cfg* cfg);
int vipnerf_pack_weights(const vipnerf_cfg* cfg, const float* const params[24], void* packed, void* stream);

/* --- the whole path: replaces VipNeRF.render_rays (VipNeRF01.py:74-171) for n_rays rays (any count; the
 * reference's chunk / netchunk loops :47-72, :295-329 are subsumed).  packed_fine may be NULL iff n_fine == 0. */
size_t vipnerf_workspace_bytes(const vipnerf_cfg* cfg, int64_t n_rays);
int vipnerf_render_forward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays,
                           const void* packed_coarse, const void* packed_fine, const vipnerf_out* out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* --- stage entry points (used by the teacher-forced parity tests; each is also a piece of the path) --- */

/* MLP.forward on sample points (VipNeRF01.py:509-535 through run_network :264-293):
 * pts = pts_o + pts_d * z (:105-107).  z [R,S]; outputs raw_* of `out` only. */
int vipnerf_mlp_forward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, int32_t n_samples,
                        const float* z_vals, const void* packed, const vipnerf_pass_out* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* get_z_vals_coarse (VipNeRF01.py:173-203): near/far (+t_rand) -> z [R,n_coarse]. */
int vipnerf_coarse_z(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, float* z_vals, void* stream);

/* volume_rendering (VipNeRF01.py:331-384) on given network outputs: sigma [R,S], rgb [R,S,3],
 * vis2 [R,S,V] or NULL, z [R,S]; if z_fine_out != NULL additionally get_z_vals_fine (:205-216 =
 * sample_pdf :229-262 + sort) -> [R, S + n_fine]. */
int vipnerf_composite(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, int32_t n_samples,
                      const float* z_vals, const float* sigma, const float* rgb, const float* vis2,
                      const vipnerf_pass_out* out, float* z_fine_out, void* stream);

/* --- the steps either side of the path, on the device -----------------------------
 *
 * vipnerf_generate_rays replaces the per-pixel part of DataPreprocessor.create_test_data
 * (src/data_preprocessors/DataPreprocessor01.py:776-864): get_rays :335-352 (pinhole direction K^-1 [x, y, 1],
 * y/z flipped, rotated by the camera-to-world pose), get_view_dirs :376-378, get_ndc_rays :355-373, the near/far
 * fills and the broadcast of the secondary camera centres (rays_o2, :851-854) for pixels
 * [first_pixel, first_pixel + n_rays) of the frame in row-major order.  The 4x4 pose algebra
 * (preprocess_poses :906-945) stays on the host: `pose` is the processed 3x4 camera-to-world matrix. */
typedef struct vipnerf_camera {
  int32_t height, width;
  int32_t ndc;             /* also fill rays_o_ndc / rays_d_ndc / near_ndc / far_ndc                    */
  int32_t n_sec_views;     /* V <= 8 secondary camera centres                                           */
  int32_t has_view_pose;   /* view_dirs from view_kinv / view_pose instead of the render camera (:801-814) */
  float kinv[9];           /* numpy.linalg.inv(intrinsic), row-major, fp32 (:345)                        */
  float pose[12];          /* processed pose[:3, :4], row-major                                          */
  float view_kinv[9];
  float view_pose[12];
  float near, far;         /* model_configs['near'], ['far']                                             */
  float near_ndc, far_ndc; /* model_configs['near_ndc'], ['far_ndc']                                     */
  float sx, sy;            /* -1 / (w / (2 fx)),  -1 / (h / (2 fy))   (:364-369)                         */
  float sec_origins[24];   /* [V][3] processed secondary poses' translation column                       */
} vipnerf_camera;

/* Device buffers vipnerf_generate_rays fills (same keys as vipnerf_rays; NULL = not wanted). */
typedef struct vipnerf_ray_buffers {
  float* rays_o;      /* [R,3] */
  float* rays_d;      /* [R,3] */
  float* view_dirs;   /* [R,3] */
  float* near;        /* [R,1] */
  float* far;         /* [R,1] */
  float* rays_o_ndc;  /* [R,3] */
  float* rays_d_ndc;  /* [R,3] */
  float* near_ndc;    /* [R,1] */
  float* far_ndc;     /* [R,1] */
  float* rays_o2;     /* [R,V,3] */
} vipnerf_ray_buffers;

int vipnerf_generate_rays(const vipnerf_camera* camera, int64_t first_pixel, int64_t n_rays,
                          const vipnerf_ray_buffers* out, void* stream);

/* vipnerf_postprocess_frame replaces DataPreprocessor.retrieve_inference_outputs' per-pixel work
 * (DataPreprocessor01.py:866-894): post_process_image :1074-1078 (clip to [0,1], round(x * 255) half-to-even,
 * uint8), post_process_depth :1080-1083 (clip to [0, inf)) for up to 4 depth-like maps, and the
 * (h*w, V) -> (V, h*w) transpose of visibility2 (:887-890).  Any pointer may be NULL (skipped). */
int vipnerf_postprocess_frame(int64_t n_rays, int32_t n_sec_views, const float* rgb, uint8_t* image_u8,
                              int32_t n_depth_maps, const float* const* depth_in, float* const* depth_out,
                              const float* visibility2, float* visibility2_out, void* stream);

/* --- training batches from the per-pixel caches (the caller side of the training step) --------------------------
 * Replaces the data assembly of DataPreprocessor.load_cached_next_batch (src/data_preprocessors/DataPreprocessor01.py
 * :498-530 = load_nerf_cached_batch :571-615, load_sparse_depth_cached_batch :635-683, load_visibility_prior_cached_batch
 * :699-724): ~40 boolean-mask gathers `out = -1; out[mask] = table[indices[mask]]`, each a device sync + several
 * launches in the reference, as ONE launch.  Row r of every output column takes table[indices[r]] when the row's class
 * (1 = indices_mask_nerf ray, 2 = indices_mask_sparse_depth ray) is in the column's `row_classes` bit set (bit 0 = class
 * 1, bit 1 = class 2), else the reference's fill value -1.  Tables and outputs are 4-byte elements, row-major. */
typedef struct vipnerf_gather_column {
  const void* table;    /* [N, width] fp32 or int32 per-pixel cache                          */
  void* out;            /* [R, width]                                                        */
  int32_t width;        /* elements per row (1, 3, V ...)                                    */
  int32_t row_classes;  /* bit 0: class-1 rows gather, bit 1: class-2 rows gather            */
  int32_t fill_is_int;  /* -1 is written as int32 (pixel_id) instead of fp32                 */
  int32_t reserved;
} vipnerf_gather_column;
int vipnerf_gather_train_batch(const int64_t* indices, const uint8_t* row_class, int64_t n_rows,
                               const vipnerf_gather_column* columns_host, int32_t n_columns, void* stream);

/* --- the visibility prior generator -----------------------------------------------
 * Replaces VisibilityWeightsComputer.compute_weights
 * (src/prior_generators/visibility/VisibilityMask02_NeRF_LLFF.py:27-35 = create_psv :41-47,
 * compute_transformed_coordinates :49-82, bilinear_interpolation :84-162) for one ordered frame pair:
 * weights[y, x] = exp(-min_d mean_c |warp_d(frame2)[y, x, c] - frame1[y, x, c]| / temperature) over the given depth
 * planes, fp64 like the reference; mask = weights > 0.5 (start_generation :276-277; may be NULL).
 * frame1 / frame2: DEVICE uint8 [h, w, 3]; k1inv = inv(intrinsic1), t = extrinsic2 @ inv(extrinsic1), k2 = intrinsic2
 * and depth_planes_host[n_planes <= 256] are small HOST arrays (row-major fp64, made with the reference's own numpy
 * expressions); weights: DEVICE fp64 [h, w]; mask: DEVICE uint8 [h, w]. */
int vipnerf_visibility_prior(int32_t height, int32_t width, const uint8_t* frame1, const uint8_t* frame2,
                             const double* k1inv_host, const double* t_host, const double* k2_host,
                             const double* depth_planes_host, int32_t n_planes, double temperature,
                             double* weights, uint8_t* mask, void* stream);

/* --- training: forward with saved activations + backward --------------------------
 * Replaces what torch.autograd does for the reference's training step (src/Trainer01.py:93-102:
 * `model(batch)` in train mode -> LossComputer.compute_losses -> `TotalLoss.backward()`): the forward of
 * VipNeRF.render_rays with perturb / raw_noise_std active (VipNeRF01.py:194-202, :242, :549-552; retraw and
 * sec_views_vis are forced on, :40) and the gradient of every parameter given the gradients of the outputs.
 * The losses themselves stay the caller's (loss_functions/*.py operate on the returned tensors).
 * Arithmetic: cfg->precision must be VIPNERF_PRECISION_FP32 (fp32 packed weights); fp32 CUDA-core kernels like the
 * reference's training arithmetic by default, tensor cores with VIPNERF_FLAG_TRAIN_TF32 / VIPNERF_FLAG_TRAIN_F16 (the same
 * flag must be set for the forward, the backward and the two size queries).  Random numbers are the caller's: rays->t_rand [R,Nc], rays->u_rand [R,Nf] (torch.rand) and
 * sigma_noise_* = raw_noise_std * torch.randn, drawn exactly where the reference draws them; NULL = that source off.
 *
 * `saved` (vipnerf_train_saved_bytes, about 11 KB per sample point; 5.6 KB with VIPNERF_FLAG_TRAIN_F16: fp16 arrays + the
 * ReLU masks as bits) receives the activations of both MLPs; the caller
 * keeps it, together with the forward outputs z_vals / raw_sigma / raw_rgb / raw_visibility (/ raw_visibility2) of
 * both sample sets, until vipnerf_train_backward.  `grad_out` has the layout of vipnerf_out and holds the upstream
 * gradient of each output (NULL = zero; z_vals carries no gradient: sample positions are detached, :213).
 * param_grads_*[24] = one fp32 device buffer per parameter tensor, same order and shapes as vipnerf_pack_weights'
 * params; every element is overwritten.  packed_* must be VIPNERF_PRECISION_FP32 packs of the current weights. */
size_t vipnerf_train_saved_bytes(const vipnerf_cfg* cfg, int64_t n_rays);
size_t vipnerf_train_workspace_bytes(const vipnerf_cfg* cfg, int64_t n_rays);
int vipnerf_train_forward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays,
                          const float* sigma_noise_coarse, const float* sigma_noise_fine, const void* packed_coarse,
                          const void* packed_fine, const vipnerf_out* out, void* saved, size_t saved_bytes,
                          void* workspace, size_t workspace_bytes, void* stream);
int vipnerf_train_backward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, const void* packed_coarse,
                           const void* packed_fine, const vipnerf_out* fwd_out, const vipnerf_out* grad_out,
                           const void* saved, size_t saved_bytes, float* const param_grads_coarse[24],
                           float* const param_grads_fine[24], void* workspace, size_t workspace_bytes, void* stream);

/* --- the four training losses fused into the step (second half of row f1) ---------------------------------------
 * Replaces loss_functions/LossComputer01.py:33-51 with MSE01.py:25-67 (rgb_coarse + rgb_fine vs target_rgb on
 * indices_mask_nerf), VisibilityLoss01.py:26-74 (MAE between raw_visibility[..., 0] and the transmittance `visibility`,
 * each side detached in turn, :57-58), VisibilityPriorLoss01.py:25-89 (prior-masked 1 - visibility2 on
 * indices_mask_nerf) and SparseDepthMSE01.py:26-71 (depth_fine - depth_coarse for a coarse-only model - vs
 * sparse_depth_values[:, 0] on indices_mask_sparse_depth).  A weight of 0 switches a loss off.
 * vipnerf_fused_losses reads the forward outputs of vipnerf_train_forward and writes losses_dev[8] (device):
 * [0..3] the four loss values, [4] TotalLoss = sum of weight * value, [5] / [6] the mask counts the means divide by.
 * vipnerf_train_backward_fused is vipnerf_train_backward with dTotalLoss/d(output) formed INSIDE the compositing
 * backward from `spec` (and scaled by the device scalar *upstream_dev, NULL = 1): no per-sample gradient tensor exists.
 * grad_out may still carry gradients of other consumers of the outputs (added on top; NULL = none).
 * workspace of vipnerf_fused_losses: 64 * ceil(n_rays / 4) bytes, 16-byte aligned. */
typedef struct vipnerf_loss_spec {
  const float* target_rgb;            /* [R,3]                                                         */
  const uint8_t* mask_nerf;           /* [R] bool (indices_mask_nerf), NULL = every ray                */
  const uint8_t* mask_sparse_depth;   /* [R] bool (indices_mask_sparse_depth), NULL = loss off         */
  const float* sparse_depth;          /* [R] (sparse_depth_values[:, 0])                               */
  const float* prior;                 /* [R,V] visibility_prior_masks / _weights, NULL = ones          */
  const float* losses_dev;            /* backward: the losses_dev[8] vipnerf_fused_losses wrote        */
  float w_mse, w_visibility, w_prior, w_sparse_depth;
} vipnerf_loss_spec;
int vipnerf_fused_losses(const vipnerf_cfg* cfg, int64_t n_rays, const vipnerf_out* fwd_out, const vipnerf_loss_spec* spec,
                         float* losses_dev, void* workspace, size_t workspace_bytes, void* stream);
int vipnerf_train_backward_fused(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, const void* packed_coarse,
                                 const void* packed_fine, const vipnerf_out* fwd_out, const vipnerf_out* grad_out,
                                 const vipnerf_loss_spec* spec, const float* upstream_dev,
                                 const void* saved, size_t saved_bytes, float* const param_grads_coarse[24],
                                 float* const param_grads_fine[24], void* workspace, size_t workspace_bytes, void* stream);

/* Backward of volume_rendering alone (stage entry point of the parity tests; VipNeRF01.py:331-384 differentiated):
 * given the network outputs of one sample set (sigma [R,S] after its ReLU, rgb [R,S,3] / vis [R,S] / vis2 [R,S,V] after
 * their sigmoids) and the upstream gradients `grad_out` of the outputs, writes d_sigma_logit [R,S] (gradient w.r.t. the
 * density logit, i.e. through the ReLU) and d_head_logits [R,S,1+V,4] (gradient w.r.t. the views_output_linear logits
 * of the primary view [rgb, visibility] and of each secondary view [0, 0, 0, visibility2]). */
int vipnerf_composite_backward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, int32_t n_samples,
                               const float* z_vals, const float* sigma, const float* rgb, const float* vis,
                               const float* vis2, const vipnerf_pass_out* grad_out, float* d_sigma_logit,
                               float* d_head_logits, void* stream);

/* The parameter-gradient product of the training backward alone (stage entry point of the parity tests):
 * dw[m * ld_dw + n] = sum_p dy[p * ld_dy + m] * x[p * ld_x + n] for n < n_valid (what autograd computes for
 * nn.Linear.weight: grad_output^T @ input), db[m] = sum_p dy[p * ld_dy + m] (NULL = skip).
 * m in {128, 256}; n in {32, 64, 128, 256}.  mode 0 = fp32 CUDA-core kernel (k_gemm_tn; what vipnerf_train_backward
 * uses by default), mode 1 = tcgen05 kind::tf32 kernel (k_gemm_tn_tc<false>; n in {32, 64, 256}; operands rounded to tf32 by the
 * TMA copy; what VIPNERF_FLAG_TRAIN_TF32 selects), mode 2 = tcgen05 kind::f16 kernel (k_gemm_tn_tc<true>; dy and x are
 * fp16 arrays, leading dimensions in elements, n in {64, 256}; what VIPNERF_FLAG_TRAIN_F16 selects).
 * workspace: vipnerf_param_gradient_gemm_workspace_bytes() bytes, 256-byte aligned. */
size_t vipnerf_param_gradient_gemm_workspace_bytes(void);
int vipnerf_param_gradient_gemm(const void* dy, int32_t ld_dy, int32_t m, const void* x, int32_t ld_x, int32_t n,
                                int64_t n_rows, float* dw, int32_t ld_dw, int32_t n_valid, float* db, int32_t mode,
                                void* workspace, size_t workspace_bytes, void* stream);

/* The gradient-scale rule of VIPNERF_FLAG_TRAIN_F16 (host-side evaluation of the device function, for tests and
 * documentation): every fp16 gradient array of the backward-data chain is stored times the power of two that moves the
 * measured maximum `amax` of its defining array into [16, 32); 0, negative, inf and nan give 1. */
float vipnerf_grad_scale(float amax);

/* --- profiling aid (not part of the reference-facing path): a device buffer of 64 uint64 that CTA 0 of every
 * subsequent tensor-core launch fills with cycle counters of its warp roles (see tools/tc_cycle_breakdown.py);
 * NULL switches it off.  Process-global. */
int vipnerf_debug_set_profile_buffer(void* dev_u64x64);

#ifdef __cplusplus
}
#endif
#endif /* VIPNERF_H_ */
