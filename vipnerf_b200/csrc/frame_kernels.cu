// The steps either side of the render path, per pixel on the device:
// ray generation (DataPreprocessor.create_test_data, src/data_preprocessors/DataPreprocessor01.py:776-864) and
// output post-processing (retrieve_inference_outputs :866-894).  HBM-bound: 76 B/ray written (NDC; 44 B world,
// + 12 B per secondary view), 3 B/ray + 4 B per depth map for the frame outputs.
#include "kernels.h"

namespace vipnerf {
namespace {

struct CameraDev {
  vipnerf_camera c;
};

// K^-1 [x, y, 1] (numpy matmul, :345), y and z negated (:346), rotated by pose[:3,:3] (:348: sum over the last axis
// of dirs * pose rows, left to right).  Non-fused multiplies / adds like numpy.
__device__ __forceinline__ void pixel_direction(const float* kinv, const float* pose, float x, float y, float d[3]) {
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = fadd(fadd(fmul(kinv[3 * i], x), fmul(kinv[3 * i + 1], y)), kinv[3 * i + 2]);
  c[1] = -c[1];
  c[2] = -c[2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    d[i] = fadd(fadd(fmul(c[0], pose[4 * i]), fmul(c[1], pose[4 * i + 1])), fmul(c[2], pose[4 * i + 2]));
}

__global__ void k_generate_rays(const __grid_constant__ CameraDev cam, int64_t first_pixel, int64_t n_rays,
                                vipnerf_ray_buffers out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const vipnerf_camera& c = cam.c;
  const int64_t pix = first_pixel + r;
  const float x = (float)(pix % c.width), y = (float)(pix / c.width);
  float d[3];
  pixel_direction(c.kinv, c.pose, x, y, d);
  const float o[3] = {c.pose[3], c.pose[7], c.pose[11]};
  if (out.rays_o) { out.rays_o[3 * r] = o[0]; out.rays_o[3 * r + 1] = o[1]; out.rays_o[3 * r + 2] = o[2]; }
  if (out.rays_d) { out.rays_d[3 * r] = d[0]; out.rays_d[3 * r + 1] = d[1]; out.rays_d[3 * r + 2] = d[2]; }
  if (out.view_dirs) {   // get_view_dirs :376-378 (of the view camera's rays when one is given, :801-814)
    float v[3] = {d[0], d[1], d[2]};
    if (c.has_view_pose) pixel_direction(c.view_kinv, c.view_pose, x, y, v);
    const float n = sqrtf(fadd(fadd(fmul(v[0], v[0]), fmul(v[1], v[1])), fmul(v[2], v[2])));
    out.view_dirs[3 * r] = fdiv(v[0], n);
    out.view_dirs[3 * r + 1] = fdiv(v[1], n);
    out.view_dirs[3 * r + 2] = fdiv(v[2], n);
  }
  if (out.near) out.near[r] = c.near;
  if (out.far) out.far[r] = c.far;
  if (c.ndc) {   // get_ndc_rays :355-373
    const float t = fdiv(-fadd(c.near, o[2]), d[2]);
    const float p0 = fadd(o[0], fmul(t, d[0])), p1 = fadd(o[1], fmul(t, d[1])), p2 = fadd(o[2], fmul(t, d[2]));
    if (out.rays_o_ndc) {
      out.rays_o_ndc[3 * r] = fdiv(fmul(c.sx, p0), p2);
      out.rays_o_ndc[3 * r + 1] = fdiv(fmul(c.sy, p1), p2);
      out.rays_o_ndc[3 * r + 2] = fadd(1.f, fdiv(fmul(2.f, c.near), p2));
    }
    if (out.rays_d_ndc) {
      out.rays_d_ndc[3 * r] = fmul(c.sx, fsub(fdiv(d[0], d[2]), fdiv(p0, p2)));
      out.rays_d_ndc[3 * r + 1] = fmul(c.sy, fsub(fdiv(d[1], d[2]), fdiv(p1, p2)));
      out.rays_d_ndc[3 * r + 2] = fdiv(fmul(-2.f, c.near), p2);
    }
    if (out.near_ndc) out.near_ndc[r] = c.near_ndc;
    if (out.far_ndc) out.far_ndc[r] = c.far_ndc;
  }
  if (out.rays_o2) {
    for (int v = 0; v < c.n_sec_views; ++v) {
#pragma unroll
      for (int a = 0; a < 3; ++a) out.rays_o2[(r * c.n_sec_views + v) * 3 + a] = c.sec_origins[3 * v + a];
    }
  }
}

struct DepthMaps {
  const float* in[4];
  float* out[4];
  int n;
};

__global__ void k_postprocess_frame(int64_t n_rays, int V, const float* __restrict__ rgb, uint8_t* __restrict__ image,
                                    DepthMaps dm, const float* __restrict__ vis2, float* __restrict__ vis2_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  if (rgb != nullptr && image != nullptr) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float v = rgb[3 * r + ch];
      const float cl = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);     // numpy.clip (NaN passes through, as in numpy)
      image[3 * r + ch] = (uint8_t)__float2int_rn(fmul(cl, 255.f));   // numpy.round: half to even
    }
  }
  for (int m = 0; m < dm.n; ++m) {
    const float v = dm.in[m][r];
    dm.out[m][r] = v < 0.f ? 0.f : v;
  }
  if (vis2 != nullptr && vis2_out != nullptr) {
    for (int v = 0; v < V; ++v) vis2_out[(int64_t)v * n_rays + r] = vis2[r * V + v];
  }
}

// One thread per batch row: every column of the row is a gather from its per-pixel table or the fill value -1
// (DataPreprocessor01.py:571-724).  HBM-bound and tiny (about 150 bytes per ray); what it buys is one launch instead of
// ~40 masked gathers with a device synchronisation each.
struct GatherColumns {
  vipnerf_gather_column c[kMaxGatherColumns];
  int n;
};
__global__ void k_gather_train_batch(const int64_t* __restrict__ indices, const uint8_t* __restrict__ row_class,
                                     int64_t n_rows, GatherColumns cols) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int64_t idx = indices[r];
  const int cls = row_class[r];
  for (int i = 0; i < cols.n; ++i) {
    const vipnerf_gather_column& c = cols.c[i];
    const bool take = ((cls == 1) && (c.row_classes & 1)) || ((cls == 2) && (c.row_classes & 2));
    const uint32_t* tab = static_cast<const uint32_t*>(c.table) + idx * c.width;
    uint32_t* out = static_cast<uint32_t*>(c.out) + r * c.width;
    const uint32_t fill = c.fill_is_int ? 0xFFFFFFFFu : __float_as_uint(-1.f);
    for (int k = 0; k < c.width; ++k) out[k] = take ? tab[k] : fill;
  }
}

}  // namespace

cudaError_t launch_gather_train_batch(const int64_t* indices, const uint8_t* row_class, int64_t n_rows,
                                      const vipnerf_gather_column* columns, int n_columns, cudaStream_t s) {
  if (n_rows <= 0 || n_columns <= 0) return cudaSuccess;
  GatherColumns cols{};
  cols.n = n_columns;
  for (int i = 0; i < n_columns; ++i) cols.c[i] = columns[i];
  const int threads = 128;
  k_gather_train_batch<<<(unsigned)((n_rows + threads - 1) / threads), threads, 0, s>>>(indices, row_class, n_rows, cols);
  return cudaGetLastError();
}

cudaError_t launch_generate_rays(const vipnerf_camera& camera, int64_t first_pixel, int64_t n_rays,
                                 const vipnerf_ray_buffers& out, cudaStream_t s) {
  if (n_rays <= 0) return cudaSuccess;
  CameraDev cam;
  cam.c = camera;
  const int threads = 256;
  k_generate_rays<<<(unsigned)((n_rays + threads - 1) / threads), threads, 0, s>>>(cam, first_pixel, n_rays, out);
  return cudaGetLastError();
}

cudaError_t launch_postprocess_frame(int64_t n_rays, int n_sec_views, const float* rgb, uint8_t* image, int n_depth,
                                     const float* const* depth_in, float* const* depth_out, const float* vis2,
                                     float* vis2_out, cudaStream_t s) {
  if (n_rays <= 0) return cudaSuccess;
  DepthMaps dm{};
  dm.n = n_depth;
  for (int i = 0; i < n_depth; ++i) { dm.in[i] = depth_in[i]; dm.out[i] = depth_out[i]; }
  const int threads = 256;
  k_postprocess_frame<<<(unsigned)((n_rays + threads - 1) / threads), threads, 0, s>>>(n_rays, n_sec_views, rgb, image,
                                                                                      dm, vis2, vis2_out);
  return cudaGetLastError();
}

}  // namespace vipnerf
