// Visibility-prior generator: plane-sweep-volume visibility weights of frame 1 w.r.t.
// frame 2.  Reference: src/prior_generators/visibility/VisibilityMask02_NeRF_LLFF.py:27-171 (compute_weights,
// create_psv, compute_transformed_coordinates, bilinear_interpolation).
//
// One thread per pixel of frame 1; for each inverse-depth plane: back-project, transform into camera 2, project,
// bilinear-sample frame 2 (zero border), mean absolute colour error; keep the minimum over the planes;
// weight = exp(-min_error / temperature).  All arithmetic in fp64 like the reference (its float64 camera matrices
// promote the numpy pipeline), so the weights agree to rounding and the `> 0.5` mask exactly.
// HBM/L2-bound gather: 3 B (frame 1) + 4 x 3 B x planes of cached gathers from frame 2 read, 8 + 1 B written per pixel;
// the reference materialises [h, w, planes, .] float64 arrays (several hundred bytes per plane sample) instead.
#include "kernels.h"

namespace vipnerf {
namespace {

constexpr int kMaxPlanes = 256;

struct PriorParams {
  int h, w, n_planes;
  double k1inv[9];      // inv(intrinsic1)
  double t[16];         // extrinsic2 @ inv(extrinsic1)
  double k2[9];         // intrinsic2
  double temperature;
  double planes[kMaxPlanes];
};

__global__ void k_visibility_weights(const __grid_constant__ PriorParams p, const uint8_t* __restrict__ frame1,
                                     const uint8_t* __restrict__ frame2, double* __restrict__ weights,
                                     uint8_t* __restrict__ mask) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.h * p.w) return;
  const int x = (int)(idx % p.w), y = (int)(idx / p.w);
  // unnormalized_pos = inv(K1) @ [x, y, 1]   (:71)
  double u[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) u[i] = __dadd_rn(__dadd_rn(__dmul_rn(p.k1inv[3 * i], (double)x), __dmul_rn(p.k1inv[3 * i + 1], (double)y)), p.k1inv[3 * i + 2]);
  const double c1[3] = {(double)frame1[3 * idx], (double)frame1[3 * idx + 1], (double)frame1[3 * idx + 2]};
  double min_err = 1e300;
  for (int d = 0; d < p.n_planes; ++d) {
    const double depth = p.planes[d];
    const double wp[3] = {__dmul_rn(depth, u[0]), __dmul_rn(depth, u[1]), __dmul_rn(depth, u[2])};   // :72
    double tw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)   // :74 (homogeneous 1 last)
      tw[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(p.t[4 * i], wp[0]), __dmul_rn(p.t[4 * i + 1], wp[1])), __dmul_rn(p.t[4 * i + 2], wp[2])), p.t[4 * i + 3]);
    double n[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)   // :76
      n[i] = __dadd_rn(__dadd_rn(__dmul_rn(p.k2[3 * i], tw[0]), __dmul_rn(p.k2[3 * i + 1], tw[1])), __dmul_rn(p.k2[3 * i + 2], tw[2]));
    // :77, then create_psv's flow (coords - grid) and bilinear_interpolation's (flow + grid), then the +1 border offset
    const double px = __dadd_rn(__dadd_rn(__dsub_rn(__ddiv_rn(n[0], n[2]), (double)x), (double)x), 1.0);
    const double py = __dadd_rn(__dadd_rn(__dsub_rn(__ddiv_rn(n[1], n[2]), (double)y), (double)y), 1.0);
    // numpy casts floor()/ceil() to int64; NaN / inf become INT64_MIN there, which the clip turns into 0
    const double wmax = (double)(p.w + 1), hmax = (double)(p.h + 1);
    auto clipi = [](double v, double hi) -> int {
      if (!(v == v) || v > 9.2e18 || v < -9.2e18) return 0;
      return (int)(v < 0.0 ? 0.0 : (v > hi ? hi : v));
    };
    auto clipd = [](double v, double hi) -> double { return v < 0.0 ? 0.0 : (v > hi ? hi : v); };   // NaN passes through
    const int fx = clipi(floor(px), wmax), cx = clipi(ceil(px), wmax);
    const int fy = clipi(floor(py), hmax), cy = clipi(ceil(py), hmax);
    const double ox = clipd(px, wmax), oy = clipd(py, hmax);
    const double ax = __dsub_rn(1.0, __dsub_rn(ox, (double)fx)), bx = __dsub_rn(1.0, __dsub_rn((double)cx, ox));
    const double ay = __dsub_rn(1.0, __dsub_rn(oy, (double)fy)), by = __dsub_rn(1.0, __dsub_rn((double)cy, oy));
    const double wt[4] = {__dmul_rn(ay, ax), __dmul_rn(by, ax), __dmul_rn(ay, bx), __dmul_rn(by, bx)};   // nw, sw, ne, se
    const int ys[4] = {fy, cy, fy, cy}, xs[4] = {fx, fx, cx, cx};
    double nr[3] = {0.0, 0.0, 0.0}, dr = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool inside = ys[k] >= 1 && ys[k] <= p.h && xs[k] >= 1 && xs[k] <= p.w;   // the zero border of mask2
      const double m = inside ? 1.0 : 0.0;
      const uint8_t* px2 = frame2 + ((int64_t)(inside ? ys[k] - 1 : 0) * p.w + (inside ? xs[k] - 1 : 0)) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double f = inside ? (double)px2[c] : 0.0;
        nr[c] = __dadd_rn(nr[c], __dmul_rn(__dmul_rn(wt[k], f), m));
      }
      dr = __dadd_rn(dr, __dmul_rn(wt[k], m));
    }
    double e = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = dr > 0.0 ? __ddiv_rn(nr[c], dr) : 0.0;     // numpy.where(dr > 0, nr / dr, 0); NaN dr -> 0
      e = __dadd_rn(e, fabs(__dsub_rn(v, c1[c])));
    }
    e = __ddiv_rn(e, 3.0);                                        // numpy.mean over the channels
    if (e < min_err || !(e == e)) min_err = e;                    // numpy.min propagates NaN
  }
  const double wv = exp(__ddiv_rn(-min_err, p.temperature));
  weights[idx] = wv;
  if (mask != nullptr) mask[idx] = wv > 0.5 ? 1 : 0;
}

}  // namespace

cudaError_t launch_visibility_weights(int h, int w, const uint8_t* frame1, const uint8_t* frame2, const double k1inv[9],
                                      const double t[16], const double k2[9], const double* planes_host, int n_planes,
                                      double temperature, double* weights, uint8_t* mask, cudaStream_t s) {
  PriorParams p;
  p.h = h; p.w = w; p.n_planes = n_planes; p.temperature = temperature;
  for (int i = 0; i < 9; ++i) { p.k1inv[i] = k1inv[i]; p.k2[i] = k2[i]; }
  for (int i = 0; i < 16; ++i) p.t[i] = t[i];
  for (int i = 0; i < n_planes; ++i) p.planes[i] = planes_host[i];
  const int64_t n = (int64_t)h * w;
  const int threads = 128;
  k_visibility_weights<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(p, frame1, frame2, weights, mask);
  return cudaGetLastError();
}

}  // namespace vipnerf
