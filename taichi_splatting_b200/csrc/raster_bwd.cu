// raster_bwd.cu -- R9: backward gradient sweep, tuned fp32 / 16x16-tile kernel + C ABI dispatch.
//
// Semantics: _backward_kernel, rasterizer/backward.py:50-225 and gaussian_pdf_with_grad,
// taichi_lib/generic.py:320-336: re-walk front to back with (total_weight, remaining = image - sum f w),
// stop a pixel at total_weight >= saturate_threshold, clamp passes gradient through (D4).
//
// B200 design (not the reference's):
//   * same CTA / warp-rectangle / staged-record / per-warp hit-list structure as raster_fwd.cu;
//   * per (pixel, splat) only seven splat-independent moments are formed
//       {G p tx, G p ty, G p tx dx, G p tx dy, G p ty dx, G p ty dy, p dL/dalpha}
//     (G = alpha * dL/dalpha); the seven parameter gradients are linear in their sums, so the
//     multiplication by axis / 1/sigma happens once per (splat, tile) at flush time;
//   * the warp reduction is a transposed butterfly: 16 values x 32 lanes are reduced with
//     8+4+2+1+1 = 16 shuffles (instead of 16 x 5), leaving one finished sum in every second lane,
//     which then issues ONE conflict-free shared-memory atomic instruction per warp;
//   * one global float atomic per (splat, tile, component) at the end of each 256-splat batch.
#include "raster_common.cuh"

namespace gs {

template <typename real>
int raster_bwd_generic(const real *points, const real *features, const int32_t *ranges, const int32_t *o2p,
                       const real *image, const real *grad_image, int width, int height, int F,
                       const gs_raster_config *cfg, real *grad_points, real *grad_features, real *heuristic,
                       cudaStream_t stream);

constexpr int kTileB = 16;
constexpr int kBatchB = 256;
constexpr int kAccStride = 17;  // 16 slots + 1 pad: conflict-free both for the warp atomics and the flush
constexpr float kExpScaleB = 0.84932180028801904f;

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct BwdSmem {
  float4 a[kBatchB];  // mean.x, mean.y, (axis/sx)*k
  float4 b[kBatchB];  // (perp/sy)*k, alpha, unused
  float4 f[kBatchB];
  float acc[kBatchB * kAccStride];
  unsigned char mask[kBatchB];
  unsigned char list[8][kBatchB];
  int warp_done[8];
};

// 16 values per lane -> lane l (even) ends with the warp-wide sum of value (l >> 1) in v[0].
__device__ __forceinline__ void warp_transpose_reduce16(float (&v)[16], int lane) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, off);
    }
  }
  v[0] += __shfl_xor_sync(full, v[0], 1);
}

// slots: 0..6 moments, 7..7+F-1 feature grads, 14,15 heuristics
template <int F, bool GP, bool GF, bool HEUR>
__global__ void __launch_bounds__(kBatchB)
raster_bwd_kernel(const float *__restrict__ points, const float *__restrict__ features,
                  const int32_t *__restrict__ ranges, const int32_t *__restrict__ overlap_to_point,
                  const float *__restrict__ image, const float *__restrict__ grad_image, RasterParams<float> P,
                  float *__restrict__ grad_points, float *__restrict__ grad_features,
                  float *__restrict__ heuristic) {
  __shared__ BwdSmem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % P.tiles_wide) * kTileB, tile_y0 = (tile / P.tiles_wide) * kTileB;
  const int px = tile_x0 + (warp & 1) * 8 + (lane & 7), py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
  const bool in_bounds = px < P.width && py < P.height;
  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;

  float remaining[F], gpix[F];
#pragma unroll
  for (int c = 0; c < F; ++c) { remaining[c] = 0.f; gpix[c] = 0.f; }
  float total_weight = 1.0f;
  if (in_bounds) {
    const float *img = image + ((int64_t)py * P.width + px) * F;
    const float *gi = grad_image + ((int64_t)py * P.width + px) * F;
#pragma unroll
    for (int c = 0; c < F; ++c) { remaining[c] = img[c]; gpix[c] = gi[c]; }
    total_weight = 0.f;
  }

  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  if (lane == 0) sm.warp_done[warp] = 0;

  for (int base = start; base < end; base += kBatchB) {
    const int nb = min(kBatchB, end - base);
    __syncthreads();
    {
      int all_done = 1;
#pragma unroll
      for (int w = 0; w < 8; ++w) all_done &= sm.warp_done[w];
      if (all_done) break;
    }
    // ---- stage (thread j owns splat j of the batch, and flushes it at the end) ----
    int my_id = -1;
    float s_ax = 0.f, s_ay = 0.f, s_isx = 0.f, s_isy = 0.f;
    if (tid < nb) {
      my_id = overlap_to_point[base + tid];
      const float *g = points + 7 * (int64_t)my_id;
      float mx = g[0], my = g[1], ax = g[2], ay = g[3], sx = g[4], sy = g[5], alpha = g[6];
      float isx = 1.0f / sx, isy = 1.0f / sy;
      s_ax = ax; s_ay = ay; s_isx = isx; s_isy = isy;
      float ux = ax * isx * kExpScaleB, uy = ay * isx * kExpScaleB;
      float wx = -ay * isy * kExpScaleB, wy = ax * isy * kExpScaleB;
      sm.a[tid] = make_float4(mx, my, ux, uy);
      sm.b[tid] = make_float4(wx, wy, alpha, 0.f);
      unsigned mask = 0;
      if (alpha > P.thr) {
        float rc = sqrtf(2.0f * __logf(alpha / P.thr)) * 1.001f + 0.01f;
        float rcs = rc * kExpScaleB;
        float e1x = ax * sx, e1y = ay * sx, e2x = ay * sy, e2y = ax * sy;
        float ex = rc * sqrtf(e1x * e1x + e2x * e2x), ey = rc * sqrtf(e1y * e1y + e2y * e2y);
        float hu = fabsf(ux) * 3.5f + fabsf(uy) * 1.5f + rcs;
        float hw = fabsf(wx) * 3.5f + fabsf(wy) * 1.5f + rcs;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          float dcx = (float)(tile_x0 + (w & 1) * 8) + 4.0f - mx;
          float dcy = (float)(tile_y0 + (w >> 1) * 4) + 2.0f - my;
          bool hit = (fabsf(dcx) - 3.5f <= ex) && (fabsf(dcy) - 1.5f <= ey) &&
                     (fabsf(ux * dcx + uy * dcy) <= hu) && (fabsf(wx * dcx + wy * dcy) <= hw);
          mask |= hit ? (1u << w) : 0u;
        }
      }
      sm.mask[tid] = (unsigned char)mask;
      float4 fv = make_float4(0.f, 0.f, 0.f, 0.f);
      const float *fp = features + (int64_t)F * my_id;
      fv.x = fp[0];
      if (F > 1) fv.y = fp[1];
      if (F > 2) fv.z = fp[2];
      if (F > 3) fv.w = fp[3];
      sm.f[tid] = fv;
#pragma unroll
      for (int c = 0; c < 16; ++c) sm.acc[tid * kAccStride + c] = 0.f;
    }
    __syncthreads();

    // ---- per-warp ordered hit list ----
    int nhit = 0;
    if (!__all_sync(0xffffffffu, total_weight >= P.sat)) {
      for (int c = 0; c < nb; c += 32) {
        int j = c + lane;
        bool hit = j < nb && ((sm.mask[j] >> warp) & 1);
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) sm.list[warp][nhit + __popc(bal & ((1u << lane) - 1))] = (unsigned char)j;
        nhit += __popc(bal);
      }
      __syncwarp();
    }

    // ---- gradient sweep ----
    for (int h = 0; h < nhit; ++h) {
      const int j = sm.list[warp][h];
      const float4 A = sm.a[j], B = sm.b[j];
      float dx = fx - A.x, dy = fy - A.y;
      float tx = dx * A.z + dy * A.w, ty = dx * B.x + dy * B.y;
      float ga = ex2_approx_b(-(tx * tx + ty * ty));
      float alpha = B.z * ga;
      const bool has_grad = alpha > P.thr && total_weight < P.sat;
      float v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = 0.f;
      if (has_grad) {
        alpha = fminf(alpha, P.clamp_max);
        const float4 fv = sm.f[j];
        const float feat[4] = {fv.x, fv.y, fv.z, fv.w};
        float T_i = 1.0f - total_weight;
        float weight = alpha * T_i;
        total_weight += weight;
        float inv_1ma = __fdividef(1.0f, 1.0f - alpha);
        float alpha_grad = 0.f;
#pragma unroll
        for (int c = 0; c < F; ++c) {
          remaining[c] -= feat[c] * weight;
          float diff = feat[c] * T_i - remaining[c] * inv_1ma;
          alpha_grad += diff * gpix[c];
          if (GF) v[7 + c] = weight * gpix[c];
        }
        float G = B.z * alpha_grad;
        if (GP || HEUR) {
          float Gp = G * ga;
          float a1 = Gp * tx, a2 = Gp * ty;  // scaled by k (tx, ty carry the exp scale)
          v[0] = a1; v[1] = a2;
          v[2] = a1 * dx; v[3] = a1 * dy; v[4] = a2 * dx; v[5] = a2 * dy;
          v[6] = ga * alpha_grad;
          if (HEUR) {
            // |G dp/dmean|_1 with dp/dmean = p (tx u + ty w); A.zw, B.xy and a1, a2 each carry one k
            const float inv_k2 = 1.0f / (kExpScaleB * kExpScaleB);
            v[14] = G * G;
            v[15] = (fabsf(a1 * A.z + a2 * B.x) + fabsf(a1 * A.w + a2 * B.y)) * inv_k2;
          }
        }
      }
      if (__any_sync(0xffffffffu, has_grad)) {
        warp_transpose_reduce16(v, lane);
        if ((lane & 1) == 0 && v[0] != 0.f) atomicAdd(&sm.acc[j * kAccStride + (lane >> 1)], v[0]);
      }
      if (__all_sync(0xffffffffu, total_weight >= P.sat)) break;
    }
    if (__all_sync(0xffffffffu, total_weight >= P.sat) && lane == 0) sm.warp_done[warp] = 1;

    // ---- flush: one thread per staged splat ----
    __syncthreads();
    if (tid < nb) {
      float S[16];
      bool any = false;
#pragma unroll
      for (int c = 0; c < 16; ++c) { S[c] = sm.acc[tid * kAccStride + c]; any |= (S[c] != 0.f); }
      if (any) {
        if (GP) {
          const float inv_k = 1.0f / kExpScaleB;
          float S1 = S[0] * inv_k, S2 = S[1] * inv_k, S3 = S[2] * inv_k, S4 = S[3] * inv_k, S5 = S[4] * inv_k,
                S6 = S[5] * inv_k;
          float ux = s_ax * s_isx, uy = s_ay * s_isx, wx = -s_ay * s_isy, wy = s_ax * s_isy;
          float *gp = grad_points + 7 * (int64_t)my_id;
          atomicAdd(gp + 0, S1 * ux + S2 * wx);
          atomicAdd(gp + 1, S1 * uy + S2 * wy);
          atomicAdd(gp + 2, -s_isx * S3 - s_isy * S6);
          atomicAdd(gp + 3, -s_isx * S4 + s_isy * S5);
          atomicAdd(gp + 4, s_isx * (ux * S3 + uy * S4));
          atomicAdd(gp + 5, s_isy * (wx * S5 + wy * S6));
          atomicAdd(gp + 6, S[6]);
        }
        if (GF) {
          float *gf = grad_features + (int64_t)F * my_id;
#pragma unroll
          for (int c = 0; c < F; ++c) atomicAdd(gf + c, S[7 + c]);
        }
        if (HEUR) {
          atomicAdd(heuristic + 2 * (int64_t)my_id, S[14]);
          atomicAdd(heuristic + 2 * (int64_t)my_id + 1, S[15]);
        }
      }
    }
  }
}

template <int F>
static int launch_bwd(const float *points, const float *features, const int32_t *ranges, const int32_t *o2p,
                      const float *image, const float *grad_image, const RasterParams<float> &P, int tiles,
                      float *grad_points, float *grad_features, float *heuristic, cudaStream_t stream) {
  const bool gp = grad_points != nullptr, gf = grad_features != nullptr, he = P.heur && heuristic != nullptr;
#define GS_BWD(GP_, GF_, HE_)                                                                              \
  raster_bwd_kernel<F, GP_, GF_, HE_><<<tiles, kBatchB, 0, stream>>>(points, features, ranges, o2p, image, \
                                                                     grad_image, P, grad_points,           \
                                                                     grad_features, heuristic)
  if (gp && gf && he) GS_BWD(true, true, true);
  else if (gp && gf) GS_BWD(true, true, false);
  else if (gp && he) GS_BWD(true, false, true);
  else if (gp) GS_BWD(true, false, false);
  else if (gf && he) GS_BWD(false, true, true);
  else if (gf) GS_BWD(false, true, false);
  else if (he) GS_BWD(false, false, true);
  else return GS_OK;
#undef GS_BWD
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // namespace gs

extern "C" int gs_raster_bwd_f32(const float *points, const float *features, const int32_t *tile_ranges,
                                 const int32_t *overlap_to_point, const float *image, const float *grad_image,
                                 int64_t v, int64_t k, int32_t width, int32_t height, int32_t F,
                                 const gs_raster_config *cfg, float *grad_points, float *grad_features,
                                 float *point_heuristic, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(cfg != nullptr, "raster_bwd: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_bwd: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_point_heuristic || point_heuristic != nullptr, "raster_bwd: compute_point_heuristic needs a buffer");
  (void)v; (void)k;
  if (cfg->tile_size == gs::kTileB && !cfg->antialias && F >= 1 && F <= 4) {
    gs::RasterParams<float> P = gs::make_params<float>(cfg, width, height, F);
    int tiles = P.tiles_wide * ((height + gs::kTileB - 1) / gs::kTileB);
    switch (F) {
      case 1: return gs::launch_bwd<1>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
      case 2: return gs::launch_bwd<2>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
      case 3: return gs::launch_bwd<3>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
      default: return gs::launch_bwd<4>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
    }
  }
  return gs::raster_bwd_generic<float>(points, features, tile_ranges, overlap_to_point, image, grad_image, width,
                                       height, F, cfg, grad_points, grad_features, point_heuristic, stream);
}

extern "C" int gs_raster_bwd_f64(const double *points, const double *features, const int32_t *tile_ranges,
                                 const int32_t *overlap_to_point, const double *image, const double *grad_image,
                                 int64_t v, int64_t k, int32_t width, int32_t height, int32_t F,
                                 const gs_raster_config *cfg, double *grad_points, double *grad_features,
                                 double *point_heuristic, void *stream_) {
  GS_CHECK_ARG(cfg != nullptr, "raster_bwd: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_bwd: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_point_heuristic || point_heuristic != nullptr, "raster_bwd: compute_point_heuristic needs a buffer");
  (void)v; (void)k;
  return gs::raster_bwd_generic<double>(points, features, tile_ranges, overlap_to_point, image, grad_image, width,
                                        height, F, cfg, grad_points, grad_features, point_heuristic,
                                        (cudaStream_t)stream_);
}
