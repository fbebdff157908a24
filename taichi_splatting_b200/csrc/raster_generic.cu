// raster_generic.cu -- R8/R9 fallback kernels: any tile size (8/16/32), any feature count <= 16,
// antialias / quantile mode, fp32 and fp64 (the reference instantiates f64 for gradcheck,
// tests/test_rasterizer.py:30-90).  One thread per pixel, splats read straight from global memory,
// warp tree reduction + global atomics.  The fp32 / tile 16 / plain-pdf hot configurations are served
// by the tuned kernels in raster_fwd.cu / raster_bwd.cu instead.
//
// Semantics: rasterizer/forward.py:39-135 and rasterizer/backward.py:73-225 with every overlap
// composited exactly once and quantile mode stopping per pixel.
#include "raster_common.cuh"

namespace gs {

constexpr int kMaxF = 16;

template <typename real>
__device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// __launch_bounds__(1024): tile_size 32 launches 1024-thread blocks; without the bound the fp64 backward took 72
// registers per thread (72 x 1024 > the 64 K-register block limit) and failed at launch.
template <typename real>
__global__ void __launch_bounds__(1024) raster_fwd_generic_kernel(const real *__restrict__ points, const real *__restrict__ features,
                                          const int32_t *__restrict__ ranges,
                                          const int32_t *__restrict__ overlap_to_point, RasterParams<real> P,
                                          real *__restrict__ image, real *__restrict__ image_alpha,
                                          real *__restrict__ visibility) {
  const int F = P.num_features, ts = P.tile_size;
  int tile = blockIdx.x;
  int u, v;
  tile_pixel(threadIdx.x, ts, u, v);
  int px = (tile % P.tiles_wide) * ts + u, py = (tile / P.tiles_wide) * ts + v;
  bool in_bounds = px < P.width && py < P.height;
  real fx = (real)px + real(0.5), fy = (real)py + real(0.5);
  real accum[kMaxF];
  for (int c = 0; c < F; ++c) accum[c] = 0;
  real total_weight = in_bounds ? real(0) : real(1);
  bool done = !in_bounds;
  const real sat_lim = real(1) - P.sat;
  int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  for (int s = start; s < end; ++s) {
    if (__all_sync(0xffffffffu, done)) break;
    int id = overlap_to_point[s];
    const real *g = points + 7 * (int64_t)id;
    real weight = 0;
    bool hit = false;
    if (!done) {
      real ga = P.antialias ? pdf_aa(fx, fy, g) : pdf_plain(fx, fy, g);
      real alpha = math<real>::min(g[6] * ga, P.clamp_max);
      if (alpha > P.thr) {
        hit = true;
        weight = alpha * (real(1) - total_weight);
        total_weight += weight;
        const real *f = features + (int64_t)F * id;
        if (P.blend) {
          for (int c = 0; c < F; ++c) accum[c] += f[c] * weight;
          if (P.fwd_eps > real(0) && real(1) - total_weight <= P.fwd_eps) done = true;
        } else if (total_weight >= sat_lim) {
          for (int c = 0; c < F; ++c) accum[c] = f[c];
          done = true;
        }
      }
    }
    if (P.vis && __any_sync(0xffffffffu, hit)) {
      real w = warp_sum(weight);
      if ((threadIdx.x & 31) == 0) atomicAdd(visibility + id, w);
    }
  }
  if (in_bounds) {
    real *out = image + ((int64_t)py * P.width + px) * F;
    for (int c = 0; c < F; ++c) out[c] = accum[c];
    image_alpha[(int64_t)py * P.width + px] = P.blend ? total_weight : (total_weight > real(0) ? real(1) : real(0));
  }
}

template <typename real>
__global__ void __launch_bounds__(1024) raster_bwd_generic_kernel(const real *__restrict__ points, const real *__restrict__ features,
                                          const int32_t *__restrict__ ranges,
                                          const int32_t *__restrict__ overlap_to_point,
                                          const real *__restrict__ image, const real *__restrict__ grad_image,
                                          RasterParams<real> P, real *__restrict__ grad_points,
                                          real *__restrict__ grad_features, real *__restrict__ heuristic) {
  const int F = P.num_features, ts = P.tile_size;
  int tile = blockIdx.x;
  int u, v;
  tile_pixel(threadIdx.x, ts, u, v);
  int px = (tile % P.tiles_wide) * ts + u, py = (tile / P.tiles_wide) * ts + v;
  bool in_bounds = px < P.width && py < P.height;
  real fx = (real)px + real(0.5), fy = (real)py + real(0.5);
  real remaining[kMaxF], gpix[kMaxF];
  for (int c = 0; c < F; ++c) { remaining[c] = 0; gpix[c] = 0; }
  real total_weight = real(1);
  if (in_bounds) {
    const real *img = image + ((int64_t)py * P.width + px) * F;
    const real *gi = grad_image + ((int64_t)py * P.width + px) * F;
    for (int c = 0; c < F; ++c) { remaining[c] = img[c]; gpix[c] = gi[c]; }
    total_weight = 0;
  }
  int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  const bool lane0 = (threadIdx.x & 31) == 0;
  for (int s = start; s < end; ++s) {
    if (__all_sync(0xffffffffu, total_weight >= P.sat)) break;
    int id = overlap_to_point[s];
    const real *g = points + 7 * (int64_t)id;
    real gp[7] = {0, 0, 0, 0, 0, 0, 0}, gf[kMaxF], h[2] = {0, 0};
    for (int c = 0; c < F; ++c) gf[c] = 0;
    bool has_grad = false;
    if (total_weight < P.sat) {
      real dmean[2], daxis[2], dsigma[2];
      real ga = P.antialias ? pdf_aa_grad(fx, fy, g, dmean, daxis, dsigma)
                            : pdf_plain_grad(fx, fy, g, dmean, daxis, dsigma);
      real point_alpha = g[6];
      real alpha = point_alpha * ga;
      if (alpha > P.thr) {
        has_grad = true;
        alpha = math<real>::min(alpha, P.clamp_max);
        const real *f = features + (int64_t)F * id;
        real T_i = real(1) - total_weight;
        real weight = alpha * T_i;
        total_weight += weight;
        real alpha_grad = 0;
        for (int c = 0; c < F; ++c) {
          remaining[c] -= f[c] * weight;
          real diff = f[c] * T_i - remaining[c] / (real(1) - alpha);
          alpha_grad += diff * gpix[c];
          gf[c] = weight * gpix[c];
        }
        real aag = point_alpha * alpha_grad;
        gp[0] = aag * dmean[0]; gp[1] = aag * dmean[1];
        gp[2] = aag * daxis[0]; gp[3] = aag * daxis[1];
        gp[4] = aag * dsigma[0]; gp[5] = aag * dsigma[1];
        gp[6] = ga * alpha_grad;
        h[0] = aag * aag;
        h[1] = math<real>::abs(gp[0]) + math<real>::abs(gp[1]);
      }
    }
    if (__any_sync(0xffffffffu, has_grad)) {
      if (grad_points) {
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          real r = warp_sum(gp[c]);
          if (lane0) atomicAdd(grad_points + 7 * (int64_t)id + c, r);
        }
      }
      if (grad_features)
        for (int c = 0; c < F; ++c) {
          real r = warp_sum(gf[c]);
          if (lane0) atomicAdd(grad_features + (int64_t)F * id + c, r);
        }
      if (heuristic) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          real r = warp_sum(h[c]);
          if (lane0) atomicAdd(heuristic + 2 * (int64_t)id + c, r);
        }
      }
    }
  }
}

template <typename real>
int raster_fwd_generic(const real *points, const real *features, const int32_t *ranges, const int32_t *o2p,
                       int width, int height, int F, const gs_raster_config *cfg, real *image, real *image_alpha,
                       real *visibility, cudaStream_t stream) {
  GS_CHECK_ARG(F >= 1 && F <= kMaxF, "raster: num_features=%d unsupported (1..%d)", F, kMaxF);
  int ts = cfg->tile_size;
  GS_CHECK_ARG(ts == 8 || ts == 16 || ts == 32, "raster: tile_size=%d unsupported (8, 16, 32)", ts);
  RasterParams<real> P = make_params<real>(cfg, width, height, F);
  int tiles = P.tiles_wide * ((height + ts - 1) / ts);
  if (tiles == 0) return GS_OK;
  raster_fwd_generic_kernel<real><<<tiles, ts * ts, 0, stream>>>(points, features, ranges, o2p, P, image,
                                                                 image_alpha, P.vis ? visibility : nullptr);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

template <typename real>
int raster_bwd_generic(const real *points, const real *features, const int32_t *ranges, const int32_t *o2p,
                       const real *image, const real *grad_image, int width, int height, int F,
                       const gs_raster_config *cfg, real *grad_points, real *grad_features, real *heuristic,
                       cudaStream_t stream) {
  GS_CHECK_ARG(F >= 1 && F <= kMaxF, "raster: num_features=%d unsupported (1..%d)", F, kMaxF);
  int ts = cfg->tile_size;
  GS_CHECK_ARG(ts == 8 || ts == 16 || ts == 32, "raster: tile_size=%d unsupported (8, 16, 32)", ts);
  RasterParams<real> P = make_params<real>(cfg, width, height, F);
  int tiles = P.tiles_wide * ((height + ts - 1) / ts);
  if (tiles == 0) return GS_OK;
  raster_bwd_generic_kernel<real><<<tiles, ts * ts, 0, stream>>>(points, features, ranges, o2p, image, grad_image,
                                                                 P, grad_points, grad_features,
                                                                 P.heur ? heuristic : nullptr);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

template int raster_fwd_generic<float>(const float *, const float *, const int32_t *, const int32_t *, int, int, int,
                                       const gs_raster_config *, float *, float *, float *, cudaStream_t);
template int raster_fwd_generic<double>(const double *, const double *, const int32_t *, const int32_t *, int, int,
                                        int, const gs_raster_config *, double *, double *, double *, cudaStream_t);
template int raster_bwd_generic<float>(const float *, const float *, const int32_t *, const int32_t *, const float *,
                                       const float *, int, int, int, const gs_raster_config *, float *, float *,
                                       float *, cudaStream_t);
template int raster_bwd_generic<double>(const double *, const double *, const int32_t *, const int32_t *,
                                        const double *, const double *, int, int, int, const gs_raster_config *,
                                        double *, double *, double *, cudaStream_t);

}  // namespace gs
