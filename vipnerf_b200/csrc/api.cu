// C ABI of libvipnerf_b200.so (declared in include/vipnerf.h): argument checking, workspace carving and the
// launch sequence of the render path.  No device allocation, no device synchronisation, launches only on the
// caller's stream.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "kernels.h"
#include "layout.cuh"

using namespace vipnerf;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int fail_cuda(cudaError_t e, const char* what) {
  return fail(VIPNERF_ECUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

bool is_tc(int precision) {
  return precision == VIPNERF_PRECISION_BF16 || precision == VIPNERF_PRECISION_BF16X3 || precision == VIPNERF_PRECISION_FP16;
}

int check_cfg(const vipnerf_cfg* cfg) {
  if (cfg == nullptr) return fail(VIPNERF_EINVAL, "cfg is NULL");
  if (cfg->abi != VIPNERF_ABI_VERSION) return fail(VIPNERF_EABI, "cfg.abi=%d, library is %d", cfg->abi, VIPNERF_ABI_VERSION);
  if (cfg->depth != 8 || cfg->width != kWidth || cfg->skip != 4 || cfg->l_pts != kLPts || cfg->l_view != kLView)
    return fail(VIPNERF_EUNSUPPORTED,
                "network shape D=%d W=%d skip=%d L_pts=%d L_view=%d: kernels are built for D=8 W=256 skip=4 L_pts=10 L_view=4",
                cfg->depth, cfg->width, cfg->skip, cfg->l_pts, cfg->l_view);
  if (cfg->n_coarse < 3 || cfg->n_coarse > 256) return fail(VIPNERF_EUNSUPPORTED, "n_coarse=%d outside [3,256]", cfg->n_coarse);
  if (cfg->n_fine < 0 || cfg->n_coarse + cfg->n_fine > 256)
    return fail(VIPNERF_EUNSUPPORTED, "n_coarse+n_fine=%d outside [n_coarse,256]", cfg->n_coarse + cfg->n_fine);
  if (cfg->n_sec_views < 0 || cfg->n_sec_views > 16) return fail(VIPNERF_EUNSUPPORTED, "n_sec_views=%d outside [0,16]", cfg->n_sec_views);
  if (cfg->precision != VIPNERF_PRECISION_FP32 && !is_tc(cfg->precision))
    return fail(VIPNERF_EUNSUPPORTED, "precision=%d unknown", cfg->precision);
  if (is_tc(cfg->precision)) {
    if (cfg->n_sec_views > 8)
      return fail(VIPNERF_EUNSUPPORTED, "tensor-core path holds at most 8 secondary views per tile (n_sec_views=%d)", cfg->n_sec_views);
    if (cfg->n_coarse != 64 || (cfg->n_fine != 0 && cfg->n_fine != 128))
      return fail(VIPNERF_EUNSUPPORTED, "tensor-core path is built for 64 coarse + 128 fine samples (got %d + %d)", cfg->n_coarse, cfg->n_fine);
  }
  return VIPNERF_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  size_t off_z_coarse, off_z_fine, off_sigma, off_rgb, off_vis, off_vis2, off_sigma_c, off_rgb_c, off_vis_c, off_vis2_c, total;
};

Workspace carve(const vipnerf_cfg* cfg, int64_t n_rays) {
  Workspace w{};
  const size_t R = (size_t)(n_rays > 0 ? n_rays : 0);
  const size_t sf = (size_t)cfg->n_coarse + cfg->n_fine;
  size_t off = 0;
  auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats * sizeof(float), 256); return o; };
  w.off_z_coarse = take(R * cfg->n_coarse);
  w.off_z_fine = take(R * sf);
  w.off_sigma = take(R * sf);
  w.off_rgb = take(R * sf * 3);
  w.off_vis = take(R * sf);
  w.off_vis2 = take(R * sf * (size_t)cfg->n_sec_views);
  w.off_sigma_c = take(R * cfg->n_coarse);
  w.off_rgb_c = take(R * cfg->n_coarse * 3);
  w.off_vis_c = take(R * cfg->n_coarse);
  w.off_vis2_c = take(R * cfg->n_coarse * (size_t)cfg->n_sec_views);
  w.total = off + 256;
  return w;
}

bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

int make_ray_ptrs(const vipnerf_cfg* cfg, const vipnerf_rays* r, RayPtrs* rp, bool need_near_far, bool need_u,
                  bool need_o2 = true) {
  if (r == nullptr) return fail(VIPNERF_EINVAL, "rays is NULL");
  const bool ndc = cfg->flags & VIPNERF_FLAG_NDC;
  if (!r->rays_o || !r->rays_d || !r->view_dirs) return fail(VIPNERF_EINVAL, "rays_o / rays_d / view_dirs must be non-NULL");
  if (ndc && (!r->rays_o_ndc || !r->rays_d_ndc)) return fail(VIPNERF_EINVAL, "NDC flag set but rays_o_ndc / rays_d_ndc is NULL");
  rp->rays_o = r->rays_o;
  rp->rays_d = r->rays_d;
  rp->view_dirs = r->view_dirs;
  rp->pts_o = ndc ? r->rays_o_ndc : r->rays_o;
  rp->pts_d = ndc ? r->rays_d_ndc : r->rays_d;
  rp->near = ndc ? r->near_ndc : r->near;
  rp->far = ndc ? r->far_ndc : r->far;
  if (need_near_far && (!rp->near || !rp->far || !r->t_vals))
    return fail(VIPNERF_EINVAL, "near / far (%s) and t_vals must be non-NULL", ndc ? "NDC" : "world");
  rp->rays_o2 = r->rays_o2;
  if (need_o2 && cfg->n_sec_views > 0 && !r->rays_o2) return fail(VIPNERF_EINVAL, "n_sec_views=%d but rays_o2 is NULL", cfg->n_sec_views);
  rp->t_vals = r->t_vals;
  rp->u_vals = r->u_vals;
  rp->t_rand = r->t_rand;
  rp->u_rand = r->u_rand;
  if (need_u && cfg->n_fine > 0 && !r->u_vals && !r->u_rand) return fail(VIPNERF_EINVAL, "u_vals (or u_rand) must be non-NULL when n_fine > 0");
  const void* ptrs[] = {r->rays_o, r->rays_d, r->view_dirs, r->near, r->far, r->rays_o_ndc, r->rays_d_ndc, r->near_ndc,
                        r->far_ndc, r->rays_o2, r->t_vals, r->u_vals, r->t_rand, r->u_rand};
  for (const void* p : ptrs)
    if (misaligned(p)) return fail(VIPNERF_EINVAL, "input pointer %p is not 16-byte aligned", p);
  return VIPNERF_OK;
}

RenderFlags make_flags(const vipnerf_cfg* cfg) {
  RenderFlags f;
  f.ndc = cfg->flags & VIPNERF_FLAG_NDC;
  f.white_bkgd = cfg->flags & VIPNERF_FLAG_WHITE_BKGD;
  f.lindisp = cfg->flags & VIPNERF_FLAG_LINDISP;
  f.n_sec_views = cfg->n_sec_views;
  return f;
}

PassOutPtrs to_dev(const vipnerf_pass_out* o) {
  PassOutPtrs p{};
  if (o == nullptr) return p;
  p.rgb = o->rgb; p.acc = o->acc; p.depth = o->depth; p.depth_var = o->depth_var;
  p.depth_ndc = o->depth_ndc; p.depth_var_ndc = o->depth_var_ndc; p.visibility2 = o->visibility2;
  p.alpha = o->alpha; p.z_vals = o->z_vals; p.visibility = o->visibility; p.weights = o->weights;
  p.raw_sigma = o->raw_sigma; p.raw_rgb = o->raw_rgb; p.raw_visibility = o->raw_visibility;
  p.raw_visibility2 = o->raw_visibility2;
  return p;
}

PassGradPtrs to_grads(const vipnerf_pass_out* o) {
  PassGradPtrs g{};
  if (o == nullptr) return g;
  g.rgb = o->rgb; g.acc = o->acc; g.depth = o->depth; g.depth_var = o->depth_var; g.depth_ndc = o->depth_ndc;
  g.depth_var_ndc = o->depth_var_ndc; g.visibility2 = o->visibility2; g.alpha = o->alpha; g.visibility = o->visibility;
  g.weights = o->weights; g.raw_sigma = o->raw_sigma; g.raw_rgb = o->raw_rgb; g.raw_visibility = o->raw_visibility;
  g.raw_visibility2 = o->raw_visibility2;
  return g;
}

// Activations the training forward keeps per sample set (floats; P points, nv = 1 + V views): kernels.h MlpSave
struct SavedPass { size_t enc, h, feat, hv, pev, bits, end; };
struct SavedLayout { SavedPass coarse, fine; size_t total; };

bool train_f16(const vipnerf_cfg* cfg) { return (cfg->flags & VIPNERF_FLAG_TRAIN_F16) != 0; }

// fp32 arrays, or fp16 arrays in the fp16 mode (VIPNERF_FLAG_TRAIN_F16; a pev row is then 64 columns = 128 bytes as well)
SavedLayout carve_saved(const vipnerf_cfg* cfg, int64_t n_rays) {
  SavedLayout L{};
  const size_t R = (size_t)(n_rays > 0 ? n_rays : 0), nv = 1 + (size_t)cfg->n_sec_views;
  const size_t es = train_f16(cfg) ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t elems) { size_t o = off; off = align_up(off + elems * es, 256); return o; };
  auto pass = [&](size_t P) {
    SavedPass s{};
    s.enc = take(P * 64); s.h = take(P * 256 * 8); s.feat = take(P * 256); s.hv = take(P * nv * 128); s.pev = take(P * nv * (es == 2 ? 64 : 32));
    s.bits = es == 2 ? take(P * 8 * 8 * 2) : off;   // fp16 mode: ReLU masks of pts_linears.0..7 as bits, [8][P][8] words
    s.end = off;
    return s;
  };
  L.coarse = pass(R * cfg->n_coarse);
  if (cfg->n_fine > 0) L.fine = pass(R * ((size_t)cfg->n_coarse + cfg->n_fine));
  L.total = off + 256;
  return L;
}

// Split-reduction scratch of the parameter-gradient products: the larger of the FFMA and the tensor-core kernels' partial
// tiles (the latter: one 256 x 256 tile per SM, up to 160 SMs), followed by the column-sum partials.
constexpr size_t kColsumPartialFloats = 4 * 148 * 256;
size_t gemm_partial_floats() {
  const size_t a = gemm_tn_partial_floats(), b = gemm_tn_tc_partial_floats(kGemmTcMaxSplits);
  return a > b ? a : b;
}

// Scratch of the backward: sized for the larger sample set, reused by both
struct BwdLayout { size_t dsig, dlogit, dpre, dfeat, dacc9, dhv, partial, amax, total; };
constexpr int kAmaxSlots = 16;   // fp16 mode: measured maxima that define the gradient scales (one set per sample set)

BwdLayout carve_bwd(const vipnerf_cfg* cfg, int64_t n_rays) {
  BwdLayout L{};
  const size_t R = (size_t)(n_rays > 0 ? n_rays : 0), nv = 1 + (size_t)cfg->n_sec_views;
  const size_t P = R * ((size_t)cfg->n_coarse + cfg->n_fine);
  const size_t es = train_f16(cfg) ? 2 : 4;     // the chain's gradient arrays are fp16 in the fp16 mode
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.dsig = take(P * 4); L.dlogit = take(P * nv * 4 * 4); L.dpre = take(P * 256 * 8 * es); L.dfeat = take(P * 256 * es);
  L.dacc9 = take(P * 128 * es); L.dhv = take(P * nv * 128 * es);
  L.partial = take((gemm_partial_floats() + kColsumPartialFloats) * sizeof(float));
  L.amax = take(2 * kAmaxSlots * sizeof(uint32_t));
  L.total = off + 256;
  return L;
}

// ---- tensor-core training chains (VIPNERF_FLAG_TRAIN_TF32 / _F16): every 256-wide product is one k_linear_tc launch
const float* packed_small(const void* packed) { return reinterpret_cast<const float*>(packed); }
const float* packed_big(const void* packed) {   // forward images Wt[in][out] (layout.cuh)
  return reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + kSmallBytes);
}
const float* packed_out_in(const void* packed) { return packed_big(packed) + kFp32BigFloats; }   // [out][in] images

LinearTcArgs linear_args(const void* x, int ldx, int k, const void* w, int ldw, int N, int64_t P, void* out, int ld_out,
                         bool half) {
  LinearTcArgs a{};
  a.x[0] = x; a.ldx[0] = ldx; a.k[0] = k; a.w[0] = w; a.ldw[0] = ldw;
  a.N = N; a.n_rows = P; a.out = out; a.ld_out = ld_out;
  a.half_in = half; a.half_out = half;
  return a;
}

// byte pointers into the saved-activation / gradient arrays (element size es) and the two weight regions
struct ElemPtr {
  const uint8_t* base; size_t es;
  const void* at(size_t elem) const { return base + elem * es; }
};
ElemPtr weights_out_in(const void* packed, bool half) {   // [out][in] images
  if (!half) return {reinterpret_cast<const uint8_t*>(packed_out_in(packed)), 4};
  return {reinterpret_cast<const uint8_t*>(packed_big(packed) + kFp32BigFloats + kFp32BwdFloats) + (size_t)kFp32BigFloats * 2, 2};
}
ElemPtr weights_in_out(const void* packed, bool half) {   // forward images Wt[in][out]
  if (!half) return {reinterpret_cast<const uint8_t*>(packed_big(packed)), 4};
  return {reinterpret_cast<const uint8_t*>(packed_big(packed) + kFp32BigFloats + kFp32BwdFloats), 2};
}

// what the forward keeps (kernels.h MlpSave) with the element type of the mode
struct SavePtrs { const float* noise; uint8_t *enc, *h, *feat, *hv, *pev; uint32_t* bits; };

// MLP.forward of one sample set on the tensor cores: encodings -> ten products -> heads; fills the saved activations.
// half = fp16 arrays and kind::f16 products, else fp32 arrays read as tf32.
cudaError_t mlp_forward_tc(const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S, const float* z,
                           const void* packed, const SavePtrs& sv, float* acc9, float* sigma, float* rgb, float* vis,
                           float* vis2, cudaStream_t s, bool half) {
  const int64_t P = n_rays * S;
  const size_t es = half ? 2 : 4;
  const size_t PL = (size_t)P * 256 * es;      // bytes of one [P][256] array
  const float* small = packed_small(packed);
  const ElemPtr woi = weights_out_in(packed, half);
  cudaError_t e;
  if ((e = launch_encode_points(rp, fl, n_rays, S, z, sv.enc, sv.pev, s, half)) != cudaSuccess) return e;
  for (int l = 0; l < 8; ++l) {
    LinearTcArgs a;
    if (l == 0) {
      a = linear_args(sv.enc, 64, 64, woi.at(kBwdOffEnc0), 64, 256, P, sv.h, 256, half);
    } else {
      a = linear_args(sv.h + (l - 1) * PL, 256, 256, woi.at(kBwdOffTrunk + (size_t)(7 - l) * 65536), 256, 256, P, sv.h + l * PL, 256, half);
      if (l == 5) {   // cat([encoding, h4]) (:543-544): a second operand pair accumulates into the same tile
        a.x[1] = sv.enc; a.ldx[1] = 64; a.k[1] = 64; a.w[1] = woi.at(kBwdOffEnc5); a.ldw[1] = 64;
      }
    }
    a.bias = small + kOffBias + l * 256;
    a.relu = true;
    if (half) a.relu_bits_out = sv.bits + (size_t)l * P * 8;
    if (l == 7) {   // the density head rides along: the raw dot product (finished by k_heads_fwd), or all of it in the fp16 mode
      a.dot_vec = small + kOffWSigma; a.dot_out = sigma;
      if (half) { a.dot_bias = small + kOffBSigma; a.dot_noise = sv.noise; }
    }
    if ((e = launch_linear_tc(a, s)) != cudaSuccess) return e;
  }
  LinearTcArgs f = linear_args(sv.h + 7 * PL, 256, 256, woi.at(kBwdOffFeature), 256, 256, P, sv.feat, 256, half);
  f.bias = small + kOffBias + 8 * 256;
  if ((e = launch_linear_tc(f, s)) != cudaSuccess) return e;
  const int nv = 1 + fl.n_sec_views;
  if (half) {
    // fp16 mode: the whole views branch of a view direction is ONE product launch - [feature | direction encoding] against
    // [views_linears.0 feature columns | direction columns] (two operand pairs into the same accumulator), + bias, ReLU,
    // and views_output_linear + the sigmoids in the epilogue that holds the row (VipNeRF01.py:576-594).  hv / pev are
    // point-major [P][nv][.]: view v is a strided 2-D view for the TMA engine.
    const uint8_t* wvd = reinterpret_cast<const uint8_t*>(packed_big(packed) + kFp32BigFloats + kFp32BwdFloats) + (size_t)kF16MirrorHalves * 2;
    for (int view = 0; view < nv; ++view) {
      LinearTcArgs v = linear_args(sv.feat, 256, 256, woi.at(kBwdOffViews), 256, 128, P, sv.hv + (size_t)view * 128 * es, nv * 128, true);
      v.x[1] = sv.pev + (size_t)view * 64 * es; v.ldx[1] = nv * 64; v.k[1] = 64; v.w[1] = wvd; v.ldw[1] = 64;
      v.bias = small + kOffBiasViews; v.relu = true;
      v.head_wout = small + kOffWOut; v.head_bout = small + kOffBOut;
      if (view == 0) { v.head_rgb = rgb; v.head_vis = vis; v.head_vis_stride = 1; }
      else { v.head_vis = vis2 + (view - 1); v.head_vis_stride = nv - 1; }
      if ((e = launch_linear_tc(v, s)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  LinearTcArgs v = linear_args(sv.feat, 256, 256, woi.at(kBwdOffViews), 256, 128, P, acc9, 128, half);
  if ((e = launch_linear_tc(v, s)) != cudaSuccess) return e;
  return launch_heads_fwd(P, nv, packed, nullptr, acc9, sv.pev, sv.noise, sigma, rgb, vis, vis2, sv.hv, s, half);
}

// Gradient scales of the fp16 mode: every fp16 gradient array is stored times a power of two derived from a maximum
// measured on the device (grad_scale_from_amax, kernels.h).  Slot that DEFINES the scale of each array, and what the
// producing kernel records for the next one:
//   slot 0  max |dlogit|                       (k_grad_amax)        defines dhv, dacc9
//   slot 1  max |dacc9|                        (k_heads_bwd)        defines dfeat
//   slot 2  max(max |dfeat|, max |dsig| max |w_sigma|)  (k_linear_tc / k_grad_amax)  defines dpre[7]
//   slot 10 - l  max |dpre[l + 1]|             (k_linear_tc)        defines dpre[l], l = 6 .. 0
// A layer's output is re-centred on the maximum of its INPUT: one layer moves the magnitude by far less than the 11
// binades of head room (and the conversions saturate), so no array needs a second pass.
constexpr int kSlotLogit = 0, kSlotAcc9 = 1, kSlotFeat = 2;
constexpr int dpre_slot(int l) { return l == 7 ? kSlotFeat : 10 - l; }

// backward-data chain of one sample set on the tensor cores (same outputs as launch_mlp_bwd_fp32); byte pointers
struct BwdPtrs {
  int64_t n_points; int nviews;
  const float *dsig, *dlogit;
  const uint8_t *h, *hv;
  const uint32_t* bits;   // fp16 mode: ReLU masks as bits [8][P][8]
  uint8_t *dpre, *dfeat, *dacc9, *dhv;
  uint32_t* amax;     // kAmaxSlots slots of this sample set (fp16 mode), zeroed here
};
cudaError_t mlp_backward_tc(const BwdPtrs& a, const void* packed, cudaStream_t s, bool half) {
  const int64_t P = a.n_points;
  const size_t es = half ? 2 : 4;
  const size_t PL = (size_t)P * 256 * es;
  const float* small = packed_small(packed);
  const ElemPtr wio = weights_in_out(packed, half);
  uint32_t* am = half ? a.amax : nullptr;
  cudaError_t e;
  if (half) {
    if ((e = cudaMemsetAsync(am, 0, kAmaxSlots * sizeof(uint32_t), s)) != cudaSuccess) return e;
    if ((e = launch_grad_amax(P, a.nviews, packed, a.dlogit, a.dsig, am + kSlotLogit, am + kSlotFeat, s)) != cudaSuccess) return e;
  }
  if ((e = launch_heads_bwd(P, a.nviews, packed, a.dlogit, a.hv, a.dhv, a.dacc9, s, half, half ? am + kSlotLogit : nullptr,
                            half ? am + kSlotAcc9 : nullptr)) != cudaSuccess) return e;
  LinearTcArgs g = linear_args(a.dacc9, 128, 128, wio.at(fp32_layer_offset(9)), 128, 256, P, a.dfeat, 256, half);
  if (half) { g.scale_in = am + kSlotLogit; g.scale_out = am + kSlotAcc9; g.amax_out = am + kSlotFeat; }
  if ((e = launch_linear_tc(g, s)) != cudaSuccess) return e;
  g = linear_args(a.dfeat, 256, 256, wio.at(fp32_layer_offset(8)), 256, 256, P, a.dpre + 7 * PL, 256, half);
  g.rank1_row = a.dsig; g.rank1_col = small + kOffWSigma;
  if (half) g.relu_bits = a.bits + (size_t)7 * P * 8; else { g.mask = a.h + 7 * PL; g.ld_mask = 256; }
  if (half) { g.scale_in = am + kSlotAcc9; g.scale_out = am + dpre_slot(7); g.amax_out = am + dpre_slot(6); }
  if ((e = launch_linear_tc(g, s)) != cudaSuccess) return e;
  for (int l = 7; l >= 1; --l) {
    g = linear_args(a.dpre + l * PL, 256, 256, wio.at(fp32_layer_offset(l) + (l == 5 ? 64 * 256 : 0)), 256, 256, P,
                    a.dpre + (l - 1) * PL, 256, half);
    if (half) g.relu_bits = a.bits + (size_t)(l - 1) * P * 8; else { g.mask = a.h + (l - 1) * PL; g.ld_mask = 256; }
    if (half) { g.scale_in = am + dpre_slot(l); g.scale_out = am + dpre_slot(l - 1); g.amax_out = l >= 2 ? am + dpre_slot(l - 2) : nullptr; }
    if ((e = launch_linear_tc(g, s)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

int check_train_cfg(const vipnerf_cfg* cfg) {
  if (int rc = check_cfg(cfg)) return rc;
  if (cfg->precision != VIPNERF_PRECISION_FP32)
    return fail(VIPNERF_EUNSUPPORTED, "the training path computes in fp32 (cfg.precision=%d)", cfg->precision);
  if ((cfg->flags & VIPNERF_FLAG_TRAIN_TF32) && (cfg->flags & VIPNERF_FLAG_TRAIN_F16))
    return fail(VIPNERF_EINVAL, "VIPNERF_FLAG_TRAIN_TF32 and VIPNERF_FLAG_TRAIN_F16 are exclusive");
  return VIPNERF_OK;
}

}  // namespace

extern "C" {

int vipnerf_abi_version(void) { return VIPNERF_ABI_VERSION; }

const char* vipnerf_last_error(void) { return g_last_error.c_str(); }

int vipnerf_check_config(const vipnerf_cfg* cfg) { return check_cfg(cfg); }

int vipnerf_generate_rays(const vipnerf_camera* camera, int64_t first_pixel, int64_t n_rays,
                          const vipnerf_ray_buffers* out, void* stream) {
  if (camera == nullptr || out == nullptr) return fail(VIPNERF_EINVAL, "camera / out is NULL");
  if (camera->height < 1 || camera->width < 1) return fail(VIPNERF_EINVAL, "resolution %d x %d", camera->height, camera->width);
  if (camera->n_sec_views < 0 || camera->n_sec_views > 8) return fail(VIPNERF_EUNSUPPORTED, "n_sec_views=%d outside [0,8]", camera->n_sec_views);
  if (n_rays < 0 || first_pixel < 0 || first_pixel + n_rays > (int64_t)camera->height * camera->width)
    return fail(VIPNERF_EINVAL, "pixels [%lld, %lld) outside the %d x %d frame", (long long)first_pixel,
                (long long)(first_pixel + n_rays), camera->height, camera->width);
  if (n_rays == 0) return VIPNERF_OK;
  if (camera->n_sec_views > 0 && out->rays_o2 == nullptr) return fail(VIPNERF_EINVAL, "n_sec_views=%d but rays_o2 is NULL", camera->n_sec_views);
  cudaError_t e = launch_generate_rays(*camera, first_pixel, n_rays, *out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "generate_rays");
  return VIPNERF_OK;
}

int vipnerf_gather_train_batch(const int64_t* indices, const uint8_t* row_class, int64_t n_rows,
                               const vipnerf_gather_column* columns_host, int32_t n_columns, void* stream) {
  if (n_rows < 0 || n_columns < 0 || n_columns > kMaxGatherColumns)
    return fail(VIPNERF_EINVAL, "n_rows=%lld n_columns=%d (at most %d columns)", (long long)n_rows, n_columns, kMaxGatherColumns);
  if (n_rows == 0 || n_columns == 0) return VIPNERF_OK;
  if (!indices || !row_class || !columns_host) return fail(VIPNERF_EINVAL, "indices / row_class / columns is NULL");
  for (int i = 0; i < n_columns; ++i) {
    const vipnerf_gather_column& c = columns_host[i];
    if (!c.table || !c.out || c.width < 1 || c.width > 64)
      return fail(VIPNERF_EINVAL, "column %d: table / out is NULL or width %d outside [1,64]", i, c.width);
  }
  const cudaError_t e = launch_gather_train_batch(indices, row_class, n_rows, columns_host, n_columns, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "gather_train_batch");
  return VIPNERF_OK;
}

int vipnerf_postprocess_frame(int64_t n_rays, int32_t n_sec_views, const float* rgb, uint8_t* image_u8,
                              int32_t n_depth_maps, const float* const* depth_in, float* const* depth_out,
                              const float* visibility2, float* visibility2_out, void* stream) {
  if (n_rays < 0 || n_depth_maps < 0 || n_depth_maps > 4 || n_sec_views < 0)
    return fail(VIPNERF_EINVAL, "n_rays=%lld n_depth_maps=%d n_sec_views=%d", (long long)n_rays, n_depth_maps, n_sec_views);
  if (n_depth_maps > 0 && (depth_in == nullptr || depth_out == nullptr)) return fail(VIPNERF_EINVAL, "depth map lists are NULL");
  for (int i = 0; i < n_depth_maps; ++i)
    if (depth_in[i] == nullptr || depth_out[i] == nullptr) return fail(VIPNERF_EINVAL, "depth map %d is NULL", i);
  if (n_rays == 0) return VIPNERF_OK;
  cudaError_t e = launch_postprocess_frame(n_rays, n_sec_views, rgb, image_u8, n_depth_maps, depth_in, depth_out,
                                           visibility2, visibility2_out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "postprocess_frame");
  return VIPNERF_OK;
}

int vipnerf_visibility_prior(int32_t height, int32_t width, const uint8_t* frame1, const uint8_t* frame2,
                             const double* k1inv_host, const double* t_host, const double* k2_host,
                             const double* depth_planes_host, int32_t n_planes, double temperature,
                             double* weights, uint8_t* mask, void* stream) {
  if (height < 1 || width < 1) return fail(VIPNERF_EINVAL, "resolution %d x %d", height, width);
  if (n_planes < 1 || n_planes > 256) return fail(VIPNERF_EUNSUPPORTED, "n_planes=%d outside [1,256]", n_planes);
  if (!frame1 || !frame2 || !k1inv_host || !t_host || !k2_host || !depth_planes_host || !weights)
    return fail(VIPNERF_EINVAL, "frame / matrix / planes / weights pointer is NULL");
  if (!(temperature > 0.0)) return fail(VIPNERF_EINVAL, "temperature must be positive");
  cudaError_t e = launch_visibility_weights(height, width, frame1, frame2, k1inv_host, t_host, k2_host,
                                            depth_planes_host, n_planes, temperature, weights, mask,
                                            static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "visibility_prior");
  return VIPNERF_OK;
}

int vipnerf_debug_set_profile_buffer(void* dev_u64x64) {
  set_tc_profile_buffer(dev_u64x64);
  return VIPNERF_OK;
}

size_t vipnerf_packed_weight_bytes(const vipnerf_cfg* cfg) {
  if (check_cfg(cfg) != VIPNERF_OK) return 0;
  switch (cfg->precision) {
    case VIPNERF_PRECISION_FP32: return kFp32PackBytes;
    case VIPNERF_PRECISION_BF16:
    case VIPNERF_PRECISION_FP16: return kSmallBytes + (size_t)kTcBigBytes;
    default: return kSmallBytes + (size_t)2 * kTcBigBytes;
  }
}

int vipnerf_pack_weights(const vipnerf_cfg* cfg, const float* const params[24], void* packed, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (params == nullptr || packed == nullptr) return fail(VIPNERF_EINVAL, "params / packed is NULL");
  for (int i = 0; i < 24; ++i)
    if (params[i] == nullptr) return fail(VIPNERF_EINVAL, "params[%d] is NULL", i);
  if (reinterpret_cast<uintptr_t>(packed) & 1023u) return fail(VIPNERF_EINVAL, "packed buffer must be 1024-byte aligned");
  cudaError_t e = launch_pack_weights(cfg->precision, params, packed, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "pack_weights");
  return VIPNERF_OK;
}

size_t vipnerf_workspace_bytes(const vipnerf_cfg* cfg, int64_t n_rays) {
  if (check_cfg(cfg) != VIPNERF_OK || n_rays < 0) return 0;
  return carve(cfg, n_rays).total;
}

int vipnerf_coarse_z(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, float* z_vals, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (n_rays == 0) return VIPNERF_OK;  // empty batch: nothing to validate (empty tensors have NULL pointers)
  RayPtrs rp{};
  if (int rc = make_ray_ptrs(cfg, rays, &rp, true, false, false)) return rc;
  if (n_rays < 0 || z_vals == nullptr) return fail(VIPNERF_EINVAL, "n_rays < 0 or z_vals NULL");
  cudaError_t e = launch_coarse_z(rp, n_rays, cfg->n_coarse, cfg->flags & VIPNERF_FLAG_LINDISP, z_vals,
                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "coarse_z");
  return VIPNERF_OK;
}

int vipnerf_mlp_forward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, int32_t n_samples,
                        const float* z_vals, const void* packed, const vipnerf_pass_out* out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (int rc = check_cfg(cfg)) return rc;
  if (n_rays == 0) return VIPNERF_OK;
  RayPtrs rp{};
  if (int rc = make_ray_ptrs(cfg, rays, &rp, false, false)) return rc;
  if (n_rays < 0 || n_samples < 1 || n_samples > 256) return fail(VIPNERF_EINVAL, "n_rays=%lld n_samples=%d", (long long)n_rays, n_samples);
  if (!z_vals || !packed || !out) return fail(VIPNERF_EINVAL, "z_vals / packed / out is NULL");
  if (!out->raw_sigma || !out->raw_rgb || !out->raw_visibility) return fail(VIPNERF_EINVAL, "raw_sigma / raw_rgb / raw_visibility outputs are required");
  if (cfg->n_sec_views > 0 && !out->raw_visibility2) return fail(VIPNERF_EINVAL, "n_sec_views > 0 needs raw_visibility2");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const RenderFlags fl = make_flags(cfg);
  cudaError_t e;
  if (cfg->precision == VIPNERF_PRECISION_FP32) {
    e = launch_mlp_fp32(rp, fl, n_rays, n_samples, z_vals, packed, out->raw_sigma, out->raw_rgb, out->raw_visibility,
                        out->raw_visibility2, s);
  } else {
    if (n_samples != 64 && n_samples != 192) return fail(VIPNERF_EUNSUPPORTED, "tensor-core MLP takes 64 or 192 samples per ray (got %d)", n_samples);
    e = launch_mlp_tc(cfg->precision, rp, fl, n_rays, n_samples, z_vals, packed, out->raw_sigma, out->raw_rgb,
                      out->raw_visibility, out->raw_visibility2, s);
  }
  if (e != cudaSuccess) return fail_cuda(e, "mlp_forward");
  return VIPNERF_OK;
}

int vipnerf_composite(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, int32_t n_samples,
                      const float* z_vals, const float* sigma, const float* rgb, const float* vis2,
                      const vipnerf_pass_out* out, float* z_fine_out, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (n_rays == 0) return VIPNERF_OK;
  RayPtrs rp{};
  if (int rc = make_ray_ptrs(cfg, rays, &rp, false, z_fine_out != nullptr, false)) return rc;
  if (n_rays < 0 || n_samples < 3 || n_samples > 256) return fail(VIPNERF_EINVAL, "n_rays=%lld n_samples=%d", (long long)n_rays, n_samples);
  if (!z_vals || !sigma || !rgb || !out) return fail(VIPNERF_EINVAL, "z_vals / sigma / rgb / out is NULL");
  if (z_fine_out && (cfg->n_fine < 1 || n_samples + cfg->n_fine > 256)) return fail(VIPNERF_EINVAL, "z_fine_out given but n_fine=%d", cfg->n_fine);
  cudaError_t e = launch_composite(rp, make_flags(cfg), n_rays, n_samples, z_vals, sigma, rgb, vis2, to_dev(out),
                                   cfg->n_fine, z_fine_out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "composite");
  return VIPNERF_OK;
}

int vipnerf_render_forward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, const void* packed_coarse,
                           const void* packed_fine, const vipnerf_out* out, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (n_rays == 0) return VIPNERF_OK;
  RayPtrs rp{};
  if (int rc = make_ray_ptrs(cfg, rays, &rp, true, true)) return rc;
  if (n_rays < 0) return fail(VIPNERF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if (!packed_coarse || !out) return fail(VIPNERF_EINVAL, "packed_coarse / out is NULL");
  if (cfg->n_fine > 0 && !packed_fine) return fail(VIPNERF_EINVAL, "n_fine=%d but packed_fine is NULL", cfg->n_fine);
  const Workspace w = carve(cfg, n_rays);
  if (n_rays > 0 && (!workspace || workspace_bytes < w.total))
    return fail(VIPNERF_EWORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, w.total);
  if (n_rays == 0) return VIPNERF_OK;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(VIPNERF_EINVAL, "workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto at = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  const RenderFlags fl = make_flags(cfg);
  const int Nc = cfg->n_coarse, Sf = cfg->n_coarse + cfg->n_fine;
  const bool has_fine = cfg->n_fine > 0;
  const PassOutPtrs oc = to_dev(&out->coarse), of = to_dev(&out->fine);
  cudaError_t e;

  if (is_tc(cfg->precision)) {
    FusedArgs a{};
    a.rp = rp; a.fl = fl; a.n_rays = n_rays; a.n_coarse = Nc; a.n_fine = cfg->n_fine;
    a.packed_coarse = packed_coarse; a.packed_fine = packed_fine;
    a.out_coarse = oc; a.out_fine = of;
    a.ws_z_coarse = at(w.off_z_coarse); a.ws_z_fine = at(w.off_z_fine);
    a.ws_sigma = at(w.off_sigma); a.ws_rgb = at(w.off_rgb); a.ws_vis = at(w.off_vis);
    a.ws_sigma_c = at(w.off_sigma_c); a.ws_rgb_c = at(w.off_rgb_c); a.ws_vis_c = at(w.off_vis_c);
    a.ws_vis2 = at(w.off_vis2); a.ws_vis2_c = at(w.off_vis2_c);
    e = launch_render_fused_tc(cfg->precision, a, s);
    if (e != cudaSuccess) return fail_cuda(e, "render_fused_tc");
    return VIPNERF_OK;
  }

  // staged fp32 path: z_coarse -> MLP -> composite(+resample) -> MLP -> composite
  float* z_c = oc.z_vals ? oc.z_vals : at(w.off_z_coarse);
  float* z_f = of.z_vals ? of.z_vals : at(w.off_z_fine);
  if ((e = launch_coarse_z(rp, n_rays, Nc, fl.lindisp, z_c, s)) != cudaSuccess) return fail_cuda(e, "coarse_z");
  for (int pass = 0; pass < (has_fine ? 2 : 1); ++pass) {
    const PassOutPtrs& o = pass ? of : oc;
    const int S = pass ? Sf : Nc;
    const float* z = pass ? z_f : z_c;
    float* sig = o.raw_sigma ? o.raw_sigma : at(w.off_sigma);
    float* rgb = o.raw_rgb ? o.raw_rgb : at(w.off_rgb);
    float* vis = o.raw_visibility ? o.raw_visibility : at(w.off_vis);
    float* vis2 = fl.n_sec_views ? (o.raw_visibility2 ? o.raw_visibility2 : at(w.off_vis2)) : nullptr;
    e = launch_mlp_fp32(rp, fl, n_rays, S, z, pass ? packed_fine : packed_coarse, sig, rgb, vis, vis2, s);
    if (e != cudaSuccess) return fail_cuda(e, "mlp_fp32");
    PassOutPtrs oo = o;
    oo.z_vals = nullptr;  // z already sits in its final place
    e = launch_composite(rp, fl, n_rays, S, z, sig, rgb, vis2, oo, cfg->n_fine, (pass == 0 && has_fine) ? z_f : nullptr, s);
    if (e != cudaSuccess) return fail_cuda(e, "composite");
  }
  return VIPNERF_OK;
}

size_t vipnerf_train_saved_bytes(const vipnerf_cfg* cfg, int64_t n_rays) {
  if (check_train_cfg(cfg) != VIPNERF_OK || n_rays < 0) return 0;
  return carve_saved(cfg, n_rays).total;
}

size_t vipnerf_train_workspace_bytes(const vipnerf_cfg* cfg, int64_t n_rays) {
  if (check_train_cfg(cfg) != VIPNERF_OK || n_rays < 0) return 0;
  const size_t fwd = carve(cfg, n_rays).total, bwd = carve_bwd(cfg, n_rays).total;
  return fwd > bwd ? fwd : bwd;
}

int vipnerf_train_forward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays,
                          const float* sigma_noise_coarse, const float* sigma_noise_fine, const void* packed_coarse,
                          const void* packed_fine, const vipnerf_out* out, void* saved, size_t saved_bytes,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_train_cfg(cfg)) return rc;
  if (n_rays == 0) return VIPNERF_OK;
  RayPtrs rp{};
  if (int rc = make_ray_ptrs(cfg, rays, &rp, true, true)) return rc;
  if (n_rays < 0) return fail(VIPNERF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if (!packed_coarse || !out || !saved) return fail(VIPNERF_EINVAL, "packed_coarse / out / saved is NULL");
  const bool has_fine = cfg->n_fine > 0;
  if (has_fine && !packed_fine) return fail(VIPNERF_EINVAL, "n_fine=%d but packed_fine is NULL", cfg->n_fine);
  if (misaligned(sigma_noise_coarse) || misaligned(sigma_noise_fine)) return fail(VIPNERF_EINVAL, "sigma_noise pointers must be 16-byte aligned");
  const size_t ws_need = vipnerf_train_workspace_bytes(cfg, n_rays);
  const SavedLayout L = carve_saved(cfg, n_rays);
  if (!workspace || workspace_bytes < ws_need) return fail(VIPNERF_EWORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, ws_need);
  if (saved_bytes < L.total) return fail(VIPNERF_EWORKSPACE, "saved buffer %zu bytes < required %zu", saved_bytes, L.total);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) || (reinterpret_cast<uintptr_t>(saved) & 255u))
    return fail(VIPNERF_EINVAL, "workspace / saved must be 256-byte aligned");
  const PassOutPtrs oc = to_dev(&out->coarse), of = to_dev(&out->fine);
  for (int pass = 0; pass < (has_fine ? 2 : 1); ++pass) {
    const PassOutPtrs& o = pass ? of : oc;
    if (!o.z_vals || !o.raw_sigma || !o.raw_rgb || !o.raw_visibility || (cfg->n_sec_views > 0 && !o.raw_visibility2))
      return fail(VIPNERF_EINVAL, "training forward needs the z_vals / raw_sigma / raw_rgb / raw_visibility (/ raw_visibility2) outputs of both sample sets (the backward reads them)");
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* sv = static_cast<uint8_t*>(saved);
  const RenderFlags fl = make_flags(cfg);
  const int Nc = cfg->n_coarse, Sf = cfg->n_coarse + cfg->n_fine;
  cudaError_t e;
  if ((e = launch_coarse_z(rp, n_rays, Nc, fl.lindisp, oc.z_vals, s)) != cudaSuccess) return fail_cuda(e, "coarse_z");
  for (int pass = 0; pass < (has_fine ? 2 : 1); ++pass) {
    const PassOutPtrs& o = pass ? of : oc;
    const SavedPass& sp = pass ? L.fine : L.coarse;
    const int S = pass ? Sf : Nc;
    MlpSave ms{};
    ms.noise = pass ? sigma_noise_fine : sigma_noise_coarse;
    ms.enc = reinterpret_cast<float*>(sv + sp.enc); ms.h = reinterpret_cast<float*>(sv + sp.h);
    ms.feat = reinterpret_cast<float*>(sv + sp.feat); ms.hv = reinterpret_cast<float*>(sv + sp.hv);
    ms.pev = reinterpret_cast<float*>(sv + sp.pev);
    float* vis2 = fl.n_sec_views ? o.raw_visibility2 : nullptr;
    if (cfg->flags & (VIPNERF_FLAG_TRAIN_TF32 | VIPNERF_FLAG_TRAIN_F16)) {
      // the feature product of the views layer (fp32 [P][128]) lives in the backward's (idle) scratch until the heads have
      // consumed it
      float* acc9 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + carve_bwd(cfg, n_rays).dpre);
      const SavePtrs sp8{ms.noise, sv + sp.enc, sv + sp.h, sv + sp.feat, sv + sp.hv, sv + sp.pev,
                         reinterpret_cast<uint32_t*>(sv + sp.bits)};
      e = mlp_forward_tc(rp, fl, n_rays, S, o.z_vals, pass ? packed_fine : packed_coarse, sp8, acc9, o.raw_sigma, o.raw_rgb,
                         o.raw_visibility, vis2, s, train_f16(cfg));
    } else {
      e = launch_mlp_fp32(rp, fl, n_rays, S, o.z_vals, pass ? packed_fine : packed_coarse, o.raw_sigma, o.raw_rgb,
                          o.raw_visibility, vis2, s, &ms);
    }
    if (e != cudaSuccess) return fail_cuda(e, "mlp forward (training)");
    PassOutPtrs oo = o;
    oo.z_vals = nullptr;
    e = launch_composite(rp, fl, n_rays, S, o.z_vals, o.raw_sigma, o.raw_rgb, vis2, oo, cfg->n_fine,
                         (pass == 0 && has_fine) ? of.z_vals : nullptr, s);
    if (e != cudaSuccess) return fail_cuda(e, "composite");
  }
  return VIPNERF_OK;
}

namespace {
LossGrad to_loss_grad(const vipnerf_loss_spec* spec, const float* upstream) {
  LossGrad lg{};
  if (spec == nullptr) return lg;
  lg.target_rgb = spec->target_rgb; lg.mask_nerf = spec->mask_nerf; lg.mask_depth = spec->mask_sparse_depth;
  lg.sparse_depth = spec->sparse_depth; lg.prior = spec->prior; lg.stats = spec->losses_dev; lg.upstream = upstream;
  lg.w_mse = spec->w_mse; lg.w_vis = spec->w_visibility; lg.w_prior = spec->w_prior; lg.w_depth = spec->w_sparse_depth;
  lg.enabled = 1;
  return lg;
}
}  // namespace

int vipnerf_fused_losses(const vipnerf_cfg* cfg, int64_t n_rays, const vipnerf_out* fwd_out, const vipnerf_loss_spec* spec,
                         float* losses_dev, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_train_cfg(cfg)) return rc;
  if (n_rays < 0) return fail(VIPNERF_EINVAL, "n_rays=%lld", (long long)n_rays);
  if (!fwd_out || !spec || !losses_dev) return fail(VIPNERF_EINVAL, "fwd_out / spec / losses_dev is NULL");
  if (spec->w_mse != 0.f && !spec->target_rgb) return fail(VIPNERF_EINVAL, "spec.target_rgb is NULL");
  if (spec->mask_sparse_depth && !spec->sparse_depth) return fail(VIPNERF_EINVAL, "spec.sparse_depth is NULL");
  const size_t need = (size_t)((n_rays + 3) / 4) * 16 * sizeof(float);
  if (need > 0 && (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15u)))
    return fail(VIPNERF_EWORKSPACE, "workspace %zu bytes (need %zu, 16-byte aligned)", workspace_bytes, need);
  const bool has_fine = cfg->n_fine > 0;
  const vipnerf_pass_out &c = fwd_out->coarse, &f = fwd_out->fine;
  LossFwdArgs a{};
  a.Sc = cfg->n_coarse; a.Sf = cfg->n_coarse + cfg->n_fine; a.V = cfg->n_sec_views;
  a.rgb_c = c.rgb; a.pred_c = c.raw_visibility; a.trans_c = c.visibility; a.vis2_c = a.V > 0 ? c.visibility2 : nullptr;
  if (has_fine) { a.rgb_f = f.rgb; a.pred_f = f.raw_visibility; a.trans_f = f.visibility; a.vis2_f = a.V > 0 ? f.visibility2 : nullptr; }
  a.depth = has_fine ? f.depth : c.depth;
  if (!a.rgb_c || !a.pred_c || !a.trans_c || (has_fine && (!a.rgb_f || !a.pred_f || !a.trans_f)) || !a.depth || (a.V > 0 && (!a.vis2_c || (has_fine && !a.vis2_f))))
    return fail(VIPNERF_EINVAL, "fwd_out must hold rgb / raw_visibility / visibility / depth (/ visibility2) of every sample set");
  const cudaError_t e = launch_fused_losses(a, to_loss_grad(spec, nullptr), n_rays, losses_dev, static_cast<float*>(workspace),
                                            static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "fused_losses");
  return VIPNERF_OK;
}

int vipnerf_train_backward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, const void* packed_coarse,
                           const void* packed_fine, const vipnerf_out* fwd_out, const vipnerf_out* grad_out,
                           const void* saved, size_t saved_bytes, float* const param_grads_coarse[24],
                           float* const param_grads_fine[24], void* workspace, size_t workspace_bytes, void* stream) {
  return vipnerf_train_backward_fused(cfg, rays, n_rays, packed_coarse, packed_fine, fwd_out, grad_out, nullptr, nullptr, saved,
                                      saved_bytes, param_grads_coarse, param_grads_fine, workspace, workspace_bytes, stream);
}

int vipnerf_train_backward_fused(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, const void* packed_coarse,
                                 const void* packed_fine, const vipnerf_out* fwd_out, const vipnerf_out* grad_out,
                                 const vipnerf_loss_spec* spec, const float* upstream_dev,
                                 const void* saved, size_t saved_bytes, float* const param_grads_coarse[24],
                                 float* const param_grads_fine[24], void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_train_cfg(cfg)) return rc;
  if (spec != nullptr && spec->losses_dev == nullptr) return fail(VIPNERF_EINVAL, "spec.losses_dev is NULL (run vipnerf_fused_losses first)");
  static const vipnerf_out kNoGrads{};
  if (grad_out == nullptr && spec != nullptr) grad_out = &kNoGrads;
  RayPtrs rp{};
  if (n_rays > 0)
    if (int rc = make_ray_ptrs(cfg, rays, &rp, false, false)) return rc;
  if (n_rays < 0) return fail(VIPNERF_EINVAL, "n_rays=%lld", (long long)n_rays);
  const bool has_fine = cfg->n_fine > 0;
  if (!packed_coarse || !fwd_out || !grad_out || !param_grads_coarse) return fail(VIPNERF_EINVAL, "packed_coarse / fwd_out / grad_out / param_grads_coarse is NULL");
  if (has_fine && (!packed_fine || !param_grads_fine)) return fail(VIPNERF_EINVAL, "n_fine=%d but packed_fine / param_grads_fine is NULL", cfg->n_fine);
  for (int i = 0; i < 24; ++i)
    if (!param_grads_coarse[i] || (has_fine && !param_grads_fine[i])) return fail(VIPNERF_EINVAL, "param_grads[%d] is NULL", i);
  const SavedLayout L = carve_saved(cfg, n_rays);
  const BwdLayout B = carve_bwd(cfg, n_rays);
  if (n_rays > 0 && (!saved || saved_bytes < L.total)) return fail(VIPNERF_EWORKSPACE, "saved buffer %zu bytes < required %zu", saved_bytes, L.total);
  if (!workspace || workspace_bytes < B.total) return fail(VIPNERF_EWORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, B.total);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) || (reinterpret_cast<uintptr_t>(saved) & 255u))
    return fail(VIPNERF_EINVAL, "workspace / saved must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint8_t* sv = static_cast<const uint8_t*>(saved);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto wf = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  const RenderFlags fl = make_flags(cfg);
  const int nv = 1 + cfg->n_sec_views;
  const int Nc = cfg->n_coarse, Sf = cfg->n_coarse + cfg->n_fine;
  cudaError_t e;
  // shapes of the 24 parameter tensors (vipnerf_pack_weights order), for the empty-batch zero fill
  static const int kParamFloats[24] = {256 * 63, 256, 65536, 256, 65536, 256, 65536, 256, 65536, 256, 256 * 319, 256,
                                       65536, 256, 65536, 256, 128 * 283, 128, 256, 1, 65536, 256, 512, 4};
  for (int pass = 0; pass < (has_fine ? 2 : 1); ++pass) {
    float* const* pg = pass ? param_grads_fine : param_grads_coarse;
    if (n_rays == 0) {  // no rays: every gradient is zero
      for (int i = 0; i < 24; ++i)
        if ((e = cudaMemsetAsync(pg[i], 0, kParamFloats[i] * sizeof(float), s)) != cudaSuccess) return fail_cuda(e, "memset");
      continue;
    }
    const vipnerf_pass_out& fo = pass ? fwd_out->fine : fwd_out->coarse;
    const SavedPass& sp = pass ? L.fine : L.coarse;
    const int S = pass ? Sf : Nc;
    const int64_t P = n_rays * S;
    if (!fo.z_vals || !fo.raw_sigma || !fo.raw_rgb || !fo.raw_visibility || (cfg->n_sec_views > 0 && !fo.raw_visibility2))
      return fail(VIPNERF_EINVAL, "fwd_out must hold z_vals / raw_sigma / raw_rgb / raw_visibility (/ raw_visibility2) of both sample sets");
    const bool tf32 = (cfg->flags & VIPNERF_FLAG_TRAIN_TF32) != 0, f16 = train_f16(cfg);
    const size_t es = f16 ? 2 : 4;                       // element size of the saved activations / chain gradients
    const uint8_t* enc = sv + sp.enc; const uint8_t* h = sv + sp.h; const uint8_t* feat = sv + sp.feat;
    const uint8_t* hv = sv + sp.hv; const uint8_t* pev = sv + sp.pev;
    float* dsig = wf(B.dsig); float* dlogit = wf(B.dlogit); float* partial = wf(B.partial);
    uint8_t* dpre = ws + B.dpre; uint8_t* dfeat = ws + B.dfeat; uint8_t* dacc9 = ws + B.dacc9; uint8_t* dhv = ws + B.dhv;
    uint32_t* amax = reinterpret_cast<uint32_t*>(ws + B.amax) + pass * kAmaxSlots;
    auto as_f = [](const uint8_t* p) { return reinterpret_cast<const float*>(p); };

    // 1. volume_rendering backwards (+ the fused loss gradients) -> logit gradients per sample
    LossGrad lgp = to_loss_grad(spec, upstream_dev);
    lgp.depth_here = (has_fine ? pass == 1 : pass == 0) ? 1 : 0;
    e = launch_composite_bwd(rp, fl, n_rays, S, fo.z_vals, fo.raw_sigma, fo.raw_rgb, fo.raw_visibility,
                             cfg->n_sec_views ? fo.raw_visibility2 : nullptr,
                             to_grads(pass ? &grad_out->fine : &grad_out->coarse), lgp, dsig, dlogit, s);
    if (e != cudaSuccess) return fail_cuda(e, "composite_bwd");
    // 2. backward-data chain through the MLP
    if (tf32 || f16) {
      BwdPtrs a{};
      a.n_points = P; a.nviews = nv; a.dsig = dsig; a.dlogit = dlogit; a.h = h; a.hv = hv;
      a.bits = reinterpret_cast<const uint32_t*>(sv + sp.bits);
      a.dpre = dpre; a.dfeat = dfeat; a.dacc9 = dacc9; a.dhv = dhv; a.amax = amax;
      e = mlp_backward_tc(a, pass ? packed_fine : packed_coarse, s, f16);
    } else {
      MlpBwdArgs a{};
      a.n_points = P; a.nviews = nv; a.dsig = dsig; a.dlogit = dlogit; a.h = as_f(h); a.hv = as_f(hv);
      a.dpre = reinterpret_cast<float*>(dpre); a.dfeat = reinterpret_cast<float*>(dfeat);
      a.dacc9 = reinterpret_cast<float*>(dacc9); a.dhv = reinterpret_cast<float*>(dhv);
      e = launch_mlp_bwd_fp32(a, pass ? packed_fine : packed_coarse, s);
    }
    if (e != cudaSuccess) return fail_cuda(e, "mlp backward-data chain");
    // 3. parameter gradients: dW = dY^T X over all points, db = column sums of dY
    const size_t PL = (size_t)P * 256 * es;              // bytes of one [P][256] array
    float* colsum = partial + gemm_partial_floats();
    // one product: fp32 CUDA cores, or the tensor cores (the bias gradient = column sums of dY rides along in both);
    // fp16 mode: the partial sums carry the scale of dY (slot), undone by the final reduction
    auto gemm = [&](const uint8_t* dy, int M, const uint8_t* x, int N, int64_t rows, float* dw, int ldc, int n_valid,
                    float* db, int slot) -> cudaError_t {
      if (!tf32 && !f16) return launch_gemm_tn(as_f(dy), M, M, as_f(x), N, N, rows, dw, ldc, n_valid, db, partial, s);
      return launch_gemm_tn_tc(dy, M, M, x, N, N, rows, dw, ldc, n_valid, partial, s, db, colsum, f16, f16 ? amax + slot : nullptr);
    };
    // the eight 256 x 256 products of the sample set (pts_linears.1-7 incl. the hidden columns of the skip layer,
    // feature_linear): ONE launch on the tensor cores, the CTAs split between the problems
    GemmProblem wide[kGemmGroupMax], narrow[2];
    int n_wide = 0, n_narrow = 0;
    const bool tc = tf32 || f16;
    for (int l = 0; l < 8 && e == cudaSuccess; ++l) {
      const uint8_t* dy = dpre + l * PL;
      float* dw = pg[2 * l];
      float* db = pg[2 * l + 1];
      const int slot = dpre_slot(l);
      const uint32_t* sd = f16 ? amax + slot : nullptr;
      if (l == 0) {          // the 64 encoding columns (+ the layer's bias gradient): a narrow product
        if (tc) narrow[n_narrow++] = GemmProblem{dy, 256, enc, 64, dw, kEncPts, kEncPts, db, sd};
        else e = gemm(dy, 256, enc, 64, P, dw, kEncPts, kEncPts, db, slot);
      } else if (l == 5) {   // input = cat([encoding, h4]) (:543-544)
        if (tc) {
          narrow[n_narrow++] = GemmProblem{dy, 256, enc, 64, dw, kWidth + kEncPts, kEncPts, db, sd};
          wide[n_wide++] = GemmProblem{dy, 256, h + 4 * PL, 256, dw + kEncPts, kWidth + kEncPts, 256, nullptr, sd};
        } else {
          e = gemm(dy, 256, enc, 64, P, dw, kWidth + kEncPts, kEncPts, db, slot);
          if (e == cudaSuccess) e = gemm(dy, 256, h + 4 * PL, 256, P, dw + kEncPts, kWidth + kEncPts, 256, nullptr, slot);
        }
      } else if (tc) {
        wide[n_wide++] = GemmProblem{dy, 256, h + (l - 1) * PL, 256, dw, 256, 256, db, sd};
      } else {
        e = gemm(dy, 256, h + (l - 1) * PL, 256, P, dw, 256, 256, db, slot);
      }
    }
    if (e == cudaSuccess && tc) e = launch_gemm_tn_tc_group(narrow, n_narrow, 256, 64, P, partial, colsum, f16, s);   // both encoding-column products
    if (e != cudaSuccess) return fail_cuda(e, "gemm_tn (pts_linears)");
    // feature_linear (input h7 = output of pts_linears.7)
    if (tf32 || f16) {
      wide[n_wide++] = GemmProblem{dfeat, 256, h + 7 * PL, 256, pg[20], 256, 256, pg[21], f16 ? amax + kSlotAcc9 : nullptr};
      if ((e = launch_gemm_tn_tc_group(wide, n_wide, 256, 256, P, partial, colsum, f16, s)) != cudaSuccess)
        return fail_cuda(e, "gemm_tn (grouped 256 x 256 products)");
    } else if ((e = gemm(dfeat, 256, h + 7 * PL, 256, P, pg[20], 256, 256, pg[21], kSlotAcc9)) != cudaSuccess) {
      return fail_cuda(e, "gemm_tn (feature_linear)");
    }
    // views_linears.0: feature columns over points, direction columns and bias over (point, view) rows
    if ((e = gemm(dacc9, 128, feat, 256, P, pg[16], kWidth + kEncView, 256, nullptr, kSlotLogit)) != cudaSuccess)
      return fail_cuda(e, "gemm_tn (views_linears feature columns)");
    if ((e = gemm(dhv, 128, pev, f16 ? 64 : 32, P * nv, pg[16] + kWidth, kWidth + kEncView, kEncView, pg[17], kSlotLogit)) != cudaSuccess)
      return fail_cuda(e, "gemm_tn (views_linears direction columns)");
    // views_output_linear [4][128] and pts_output_linear [1][256]
    if ((e = launch_small_tn(dlogit, 4, hv, 128, P * nv, pg[22], pg[23], partial, s, f16)) != cudaSuccess)
      return fail_cuda(e, "small_tn (views_output_linear)");
    if ((e = launch_small_tn(dsig, 1, h + 7 * PL, 256, P, pg[18], pg[19], partial, s, f16)) != cudaSuccess)
      return fail_cuda(e, "small_tn (pts_output_linear)");
  }
  return VIPNERF_OK;
}

int vipnerf_composite_backward(const vipnerf_cfg* cfg, const vipnerf_rays* rays, int64_t n_rays, int32_t n_samples,
                               const float* z_vals, const float* sigma, const float* rgb, const float* vis,
                               const float* vis2, const vipnerf_pass_out* grad_out, float* d_sigma_logit,
                               float* d_head_logits, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (n_rays == 0) return VIPNERF_OK;
  RayPtrs rp{};
  if (int rc = make_ray_ptrs(cfg, rays, &rp, false, false, false)) return rc;
  if (n_rays < 0 || n_samples < 3 || n_samples > 256) return fail(VIPNERF_EINVAL, "n_rays=%lld n_samples=%d", (long long)n_rays, n_samples);
  if (!z_vals || !sigma || !rgb || !vis || !grad_out || !d_sigma_logit || !d_head_logits)
    return fail(VIPNERF_EINVAL, "z_vals / sigma / rgb / vis / grad_out / outputs is NULL");
  if (cfg->n_sec_views > 0 && !vis2) return fail(VIPNERF_EINVAL, "n_sec_views=%d but vis2 is NULL", cfg->n_sec_views);
  cudaError_t e = launch_composite_bwd(rp, make_flags(cfg), n_rays, n_samples, z_vals, sigma, rgb, vis, vis2,
                                       to_grads(grad_out), LossGrad{}, d_sigma_logit, d_head_logits, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "composite_bwd");
  return VIPNERF_OK;
}

float vipnerf_grad_scale(float amax) {
  uint32_t bits;
  memcpy(&bits, &amax, sizeof(bits));
  return grad_scale_from_amax(bits);
}

size_t vipnerf_param_gradient_gemm_workspace_bytes(void) { return (gemm_partial_floats() + kColsumPartialFloats) * sizeof(float) + 256; }

int vipnerf_param_gradient_gemm(const void* dy, int32_t ld_dy, int32_t m, const void* x, int32_t ld_x, int32_t n,
                                int64_t n_rows, float* dw, int32_t ld_dw, int32_t n_valid, float* db, int32_t mode,
                                void* workspace, size_t workspace_bytes, void* stream) {
  if (!dy || !x || !dw || !workspace) return fail(VIPNERF_EINVAL, "dy / x / dw / workspace is NULL");
  if ((m != 128 && m != 256) || (n != 32 && n != 64 && n != 128 && n != 256))
    return fail(VIPNERF_EUNSUPPORTED, "m=%d n=%d: m in {128, 256}, n in {32, 64, 128, 256}", m, n);
  if (n_rows < 1 || ld_dy < m || ld_x < n || n_valid < 1 || n_valid > n || ld_dw < n_valid)
    return fail(VIPNERF_EINVAL, "n_rows=%lld ld_dy=%d ld_x=%d n_valid=%d ld_dw=%d", (long long)n_rows, ld_dy, ld_x, n_valid, ld_dw);
  const int es = mode == 2 ? 2 : 4;
  if (misaligned(dy) || misaligned(x) || ((ld_dy * es) & 15) || ((ld_x * es) & 15))
    return fail(VIPNERF_EINVAL, "dy / x must be 16-byte aligned with row strides that are multiples of 16 bytes");
  if (workspace_bytes < vipnerf_param_gradient_gemm_workspace_bytes() || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return fail(VIPNERF_EWORKSPACE, "workspace %zu bytes (need %zu, 256-byte aligned)", workspace_bytes, vipnerf_param_gradient_gemm_workspace_bytes());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  cudaError_t e;
  if (mode < 0 || mode > 2) return fail(VIPNERF_EINVAL, "mode=%d", mode);
  if (mode == 1) {
    if (n != 256 && n != 64 && n != 32) return fail(VIPNERF_EUNSUPPORTED, "the tf32 tensor-core product is built for n in {32, 64, 256} (got %d)", n);
    e = launch_gemm_tn_tc(dy, ld_dy, m, x, ld_x, n, n_rows, dw, ld_dw, n_valid, partial, s, db, partial + gemm_partial_floats());
  } else if (mode == 2) {
    if (n != 256 && n != 64) return fail(VIPNERF_EUNSUPPORTED, "the fp16 tensor-core product is built for n in {64, 256} (got %d)", n);
    e = launch_gemm_tn_tc(dy, ld_dy, m, x, ld_x, n, n_rows, dw, ld_dw, n_valid, partial, s, db, partial + gemm_partial_floats(), true);
  } else {
    e = launch_gemm_tn(static_cast<const float*>(dy), ld_dy, m, static_cast<const float*>(x), ld_x, n, n_rows, dw, ld_dw, n_valid, db, partial, s);
  }
  if (e != cudaSuccess) return fail_cuda(e, "param_gradient_gemm");
  return VIPNERF_OK;
}

}  // extern "C"
