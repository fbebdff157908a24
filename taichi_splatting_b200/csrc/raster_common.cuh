// raster_common.cuh -- per-pixel Gaussian evaluation shared by the rasteriser kernels.
// Semantics: taichi_lib/generic.py:306-404 (gaussian_pdf[_with_grad], antialias variants).
#pragma once
#include "common.cuh"

namespace gs {

template <typename real>
struct RasterParams {
  int width, height, tiles_wide, num_features, tile_size;
  int antialias, blend, vis, heur;
  real clamp_max, thr, sat, fwd_eps;
  int64_t gs_y, gs_x, gs_c;   // element strides of dL/dimage (tuned backward kernel); contiguous (H,W,F) by default
};

template <typename real>
inline RasterParams<real> make_params(const gs_raster_config *c, int width, int height, int F) {
  RasterParams<real> p;
  p.width = width; p.height = height; p.num_features = F; p.tile_size = c->tile_size;
  p.tiles_wide = (width + c->tile_size - 1) / c->tile_size;
  p.antialias = c->antialias; p.blend = c->use_alpha_blending;
  p.vis = c->compute_visibility; p.heur = c->compute_point_heuristic;
  p.clamp_max = (real)c->clamp_max_alpha; p.thr = (real)c->alpha_threshold;
  p.sat = (real)c->saturate_threshold; p.fwd_eps = (real)c->forward_saturate_eps;
  p.gs_y = (int64_t)width * F; p.gs_x = F; p.gs_c = 1;
  return p;
}

template <typename real>
__device__ __forceinline__ real pdf_plain(real px, real py, const real *g) {
  real dx = px - g[0], dy = py - g[1];
  real tx = (dx * g[2] + dy * g[3]) / g[4];
  real ty = (dy * g[2] - dx * g[3]) / g[5];
  return math<real>::exp(real(-0.5) * (tx * tx + ty * ty));
}

template <typename real>
__device__ __forceinline__ real pdf_plain_grad(real px, real py, const real *g, real *dmean, real *daxis,
                                               real *dsigma) {
  real dx = px - g[0], dy = py - g[1];
  real ax = g[2], ay = g[3], sx = g[4], sy = g[5];
  real tx = (dx * ax + dy * ay) / sx;
  real ty = (dy * ax - dx * ay) / sy;
  real tx2 = tx * tx, ty2 = ty * ty;
  real p = math<real>::exp(real(-0.5) * (tx2 + ty2));
  dsigma[0] = tx2 * p / sx; dsigma[1] = ty2 * p / sy;
  real tx_s = tx / sx, ty_s = ty / sy;
  daxis[0] = p * (-tx_s * dx - ty_s * dy);
  daxis[1] = p * (-tx_s * dy + ty_s * dx);
  dmean[0] = p * (tx_s * ax - ty_s * ay);
  dmean[1] = p * (tx_s * ay + ty_s * ax);
  return p;
}

// Exact culling of an 8x8 pixel block against a splat's support (tuned kernels).  In t = U d space (U = [u; w]) the
// support { alpha pdf > threshold } is the disc |t| <= rcs and the block -- pixel centres c + (a, b), |a|, |b| <= 3.5
// -- is the parallelogram t0 + a e1 + b e2 with e1 = (ux, wx), e2 = (uy, wy).  The block can hold a contributing
// pixel only if the parallelogram comes within rcs of the origin: minimise the convex quadratic |t0 + a e1 + b e2|^2
// over the box (zero if the unconstrained minimiser is inside, else the best of the four edges).  Conservative: rcs
// carries the evaluation-error margin, the continuous box contains the pixel centres, r2 has slack for rounding.
struct SupportMetric {
  float G11, G12, G22, i11, i22, idet, r2;
};

__device__ __forceinline__ SupportMetric support_metric(float ux, float wx, float uy, float wy, float rcs) {
  SupportMetric m;
  m.G11 = fmaf(ux, ux, wx * wx); m.G22 = fmaf(uy, uy, wy * wy); m.G12 = fmaf(ux, uy, wx * wy);
  const float det_u = ux * wy - uy * wx;
  // approximate reciprocals (1 ulp-level error): they only place the minimiser, and r2 carries a 2e-4 relative margin
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(m.idet) : "f"(fmaxf(det_u * det_u, 1e-30f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(m.i11) : "f"(fmaxf(m.G11, 1e-30f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(m.i22) : "f"(fmaxf(m.G22, 1e-30f)));
  m.r2 = rcs * rcs * 1.0002f + 1e-5f;
  return m;
}

__device__ __forceinline__ bool block_reaches_support(const SupportMetric &m, float t0x, float t0y, float ux, float wx,
                                                      float uy, float wy) {
  const float h = 3.5f;
  const float g1 = fmaf(t0x, ux, t0y * wx), g2 = fmaf(t0x, uy, t0y * wy), q0 = fmaf(t0x, t0x, t0y * t0y);
  const float a0 = (g2 * m.G12 - g1 * m.G22) * m.idet, b0 = (g1 * m.G12 - g2 * m.G11) * m.idet;
  if (fabsf(a0) <= h && fabsf(b0) <= h) return true;      // the origin of t-space lies inside the parallelogram
  float qmin = 3.0e38f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const float sh = s ? h : -h;
    {   // edge a = sh, free b
      const float c0 = fmaf(h * h, m.G11, fmaf(2.0f * sh, g1, q0)), c1 = fmaf(sh, m.G12, g2);
      const float b = fminf(fmaxf(-c1 * m.i22, -h), h);
      qmin = fminf(qmin, fmaf(b, fmaf(b, m.G22, 2.0f * c1), c0));
    }
    {   // edge b = sh, free a
      const float c0 = fmaf(h * h, m.G22, fmaf(2.0f * sh, g2, q0)), c1 = fmaf(sh, m.G12, g1);
      const float a = fminf(fmaxf(-c1 * m.i11, -h), h);
      qmin = fminf(qmin, fmaf(a, fmaf(a, m.G11, 2.0f * c1), c0));
    }
  }
  return qmin <= m.r2;
}

template <typename real>
__device__ __forceinline__ real s_sig(real x, real sigma) {
  real z = x / sigma;
  return real(1) / (real(1) + math<real>::exp(real(-1.6) * z - real(0.07) * z * z * z));
}

template <typename real>
__device__ __forceinline__ real pdf_aa(real px, real py, const real *g) {
  real dx = px - g[0], dy = py - g[1];
  real sx = g[4], sy = g[5];
  real tx = dx * g[2] + dy * g[3];
  real ty = dy * g[2] - dx * g[3];
  real Sx1 = s_sig(tx + real(0.5), sx), Sx2 = s_sig(tx - real(0.5), sx);
  real Sy1 = s_sig(ty + real(0.5), sy), Sy2 = s_sig(ty - real(0.5), sy);
  return real(6.283185307179586) * sx * (Sx1 - Sx2) * sy * (Sy1 - Sy2);
}

template <typename real>
__device__ __forceinline__ void s_sig_grad(real x, real sigma, real &s, real &ds_dx, real &ds_dsig) {
  real z = x / sigma;
  s = real(1) / (real(1) + math<real>::exp(real(-1.6) * z - real(0.07) * z * z * z));
  real d = (real(1.6) + real(0.21) * z * z) * s * (real(1) - s);
  ds_dx = d / sigma;
  ds_dsig = ds_dx * -z;
}

template <typename real>
__device__ __forceinline__ real pdf_aa_grad(real px, real py, const real *g, real *dmean, real *daxis,
                                            real *dsigma) {
  real dx = px - g[0], dy = py - g[1];
  real ax = g[2], ay = g[3], sx = g[4], sy = g[5];
  real tx = dx * ax + dy * ay;
  real ty = dy * ax - dx * ay;
  real Sx1, dSx1, dSx1s, Sx2, dSx2, dSx2s, Sy1, dSy1, dSy1s, Sy2, dSy2, dSy2s;
  s_sig_grad(tx + real(0.5), sx, Sx1, dSx1, dSx1s);
  s_sig_grad(tx - real(0.5), sx, Sx2, dSx2, dSx2s);
  s_sig_grad(ty + real(0.5), sy, Sy1, dSy1, dSy1s);
  s_sig_grad(ty - real(0.5), sy, Sy2, dSy2, dSy2s);
  real ix = sx * (Sx1 - Sx2), iy = sy * (Sy1 - Sy2);
  const real tau = real(6.283185307179586);
  real dSx = iy * sx * (dSx1 - dSx2);
  real dSy = ix * sy * (dSy1 - dSy2);
  dmean[0] = tau * (-dSx * ax + dSy * ay);
  dmean[1] = tau * (-dSx * ay - dSy * ax);
  dsigma[0] = tau * iy * (Sx1 - Sx2 + (dSx1s - dSx2s) * sx);
  dsigma[1] = tau * ix * (Sy1 - Sy2 + (dSy1s - dSy2s) * sy);
  daxis[0] = tau * (dSx * dx + dSy * dy);
  daxis[1] = tau * (dSx * dy - dSy * dx);
  return tau * ix * iy;
}

}  // namespace gs
