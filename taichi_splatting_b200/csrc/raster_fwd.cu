// raster_fwd.cu -- R8: front-to-back alpha compositing, tuned fp32 / 16x16-tile kernel + C ABI dispatch.
//
// Semantics: _forward_kernel, rasterizer/forward.py:22-135 (each overlap composited exactly once, D1;
// forward early-out only below config.forward_saturate_eps, D2).  The optional fused median-depth output
// reproduces the reference's second raster pass (renderer.py:77-82: use_alpha_blending=False,
// saturate_threshold=median_threshold, features=depths) from the same walk.
//
// B200 design (not the reference's):
//   * one CTA per 16x16 tile, 8 warps, warp w owns an 8x4 pixel rectangle;
//   * splats of the tile are staged 256 at a time into shared memory as pre-digested 16-byte records
//     {mean, axis/sigma} {perp/sigma, alpha, depth} {features}: per pixel 2 FADD + 6 FMUL/FFMA +
//     1 MUFU.EX2 + blend, read with broadcast LDS.128;
//   * the staging thread classifies its splat against the eight warp rectangles (AABB + oriented-box
//     test, conservative) and each warp compacts its own ordered hit list, so a warp only iterates over
//     splats that can exceed the alpha threshold somewhere in its 32 pixels.  Skipped splats contribute
//     exactly zero in the reference too (alpha <= threshold), so results are unchanged;
//   * the inner loop is branch-free and unrolled by 16; per-splat visibility (sum of blend weights over
//     pixels) is reduced 16 splats at a time with one transposed butterfly (16 shuffles per 16 splats
//     instead of 5 per splat) and one shared-memory atomic instruction per 16 splats.
#include "packed_f32.cuh"
#include "raster_common.cuh"

#ifndef GS_PACKED
#define GS_PACKED 1   // tile-centred affine form of (tx, ty) evaluated with FFMA2 (2 issue slots instead of 6)
#endif

namespace gs {

template <typename real>
int raster_fwd_generic(const real *points, const real *features, const int32_t *ranges, const int32_t *o2p,
                       int width, int height, int F, const gs_raster_config *cfg, real *image, real *image_alpha,
                       real *visibility, cudaStream_t stream);

constexpr int kTile = 16;
constexpr int kBatch = 256;
#ifndef GS_FWD_UNROLL
#define GS_FWD_UNROLL 16
#endif
constexpr int kUnroll = GS_FWD_UNROLL;
constexpr float kExpScale = 0.84932180028801904f;  // sqrt(0.5 * log2(e)):  exp(-0.5 r^2) = 2^-(k r)^2

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct FwdSmem {
  float4 a[kBatch + 1];    // mean.x, mean.y, (axis/sx)*k          (+1: null record for list padding)
  float4 b[kBatch + 1];    // (perp/sy)*k, alpha, depth
  float4 f[kBatch + 1];    // features (F <= 4)
  int id[kBatch];
  float vis[kBatch + 1];
  unsigned char mask[kBatch];
  unsigned short list[8][kBatch + kUnroll];
  int warp_done[8];
};

// Stage one splat: returns the 8-bit mask of warp rectangles it can touch.
__device__ __forceinline__ unsigned stage_splat(const float *__restrict__ g, float thr, float tile_x0,
                                                float tile_y0, float4 &A, float4 &B) {
  float mx = g[0], my = g[1], ax = g[2], ay = g[3], sx = g[4], sy = g[5], alpha = g[6];
  float isx = 1.0f / sx, isy = 1.0f / sy;
  float ux = ax * isx * kExpScale, uy = ay * isx * kExpScale;
  float wx = -ay * isy * kExpScale, wy = ax * isy * kExpScale;
#if GS_PACKED
  {
    // (tx, ty) = X (ux, wx) + Y (uy, wy) + (tx0, ty0) with (X, Y) the pixel centre relative to the tile centre
    const float ddx = mx - (tile_x0 + 8.0f), ddy = my - (tile_y0 + 8.0f);
    A = make_float4(-fmaf(ux, ddx, uy * ddy), -fmaf(wx, ddx, wy * ddy), ux, wx);
    B = make_float4(uy, wy, alpha, 0.f);
  }
#else
  A = make_float4(mx, my, ux, uy);
  B = make_float4(wx, wy, alpha, 0.f);
#endif
  if (!(alpha > thr)) return 0u;
  // conservative support radius in sigma units (margin covers fp32 / ex2.approx evaluation error)
  float rc = sqrtf(2.0f * __logf(alpha / thr)) * 1.001f + 0.01f;
  float rcs = rc * kExpScale;
  float e1x = ax * sx, e1y = ay * sx, e2x = ay * sy, e2y = ax * sy;
  float ex = rc * sqrtf(e1x * e1x + e2x * e2x), ey = rc * sqrtf(e1y * e1y + e2y * e2y);
  float hu = fabsf(ux) * 3.5f + fabsf(uy) * 1.5f + rcs;
  float hw = fabsf(wx) * 3.5f + fabsf(wy) * 1.5f + rcs;
  unsigned mask = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    float dcx = tile_x0 + (float)((w & 1) * 8) + 4.0f - mx;
    float dcy = tile_y0 + (float)((w >> 1) * 4) + 2.0f - my;
    bool hit = (fabsf(dcx) - 3.5f <= ex) && (fabsf(dcy) - 1.5f <= ey) &&
               (fabsf(ux * dcx + uy * dcy) <= hu) && (fabsf(wx * dcx + wy * dcy) <= hw);
    mask |= hit ? (1u << w) : 0u;
  }
  return mask;
}

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

#ifndef GS_FWD_MIN_BLOCKS
#define GS_FWD_MIN_BLOCKS 3
#endif
template <int F, bool VIS, bool BLEND, bool MEDIAN>
__global__ void __launch_bounds__(kBatch, GS_FWD_MIN_BLOCKS)
raster_fwd_kernel(const float *__restrict__ points, const float *__restrict__ features,
                  const float *__restrict__ depths, const int32_t *__restrict__ ranges,
                  const int32_t *__restrict__ overlap_to_point, RasterParams<float> P, float median_lim,
                  float *__restrict__ image, float *__restrict__ image_alpha, float *__restrict__ visibility,
                  float *__restrict__ median_image) {
  __shared__ FwdSmem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % P.tiles_wide) * kTile, tile_y0 = (tile / P.tiles_wide) * kTile;
  const int px = tile_x0 + (warp & 1) * 8 + (lane & 7), py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
  const bool in_bounds = px < P.width && py < P.height;
#if GS_PACKED
  const float lx = (float)((warp & 1) * 8 + (lane & 7)) - 7.5f, ly = (float)((warp >> 1) * 4 + (lane >> 3)) - 7.5f;
  const f32x2 lx2 = pk(lx, lx), ly2 = pk(ly, ly);
#else
  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
#endif
  const float clamp_max = P.clamp_max, thr = P.thr, eps = P.fwd_eps;

  float accum[F];
#pragma unroll
  for (int c = 0; c < F; ++c) accum[c] = 0.f;
  float total_weight = in_bounds ? 0.f : 1.f;
  bool done = !in_bounds;
  float median = 0.f;
  const float sat_lim = 1.0f - P.sat;

  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  if (lane == 0) sm.warp_done[warp] = 0;
  if (tid == 0) {  // null record: alpha = 0 never passes the threshold
    sm.a[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.b[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.f[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int base = start; base < end; base += kBatch) {
    const int nb = min(kBatch, end - base);
    __syncthreads();  // previous batch fully consumed (and warp_done visible)
    {
      int all_done = 1;
#pragma unroll
      for (int w = 0; w < 8; ++w) all_done &= sm.warp_done[w];
      if (all_done) break;
    }
    if (tid < nb) {
      int id = overlap_to_point[base + tid];
      float4 A, B;
      unsigned m = stage_splat(points + 7 * (int64_t)id, thr, (float)tile_x0, (float)tile_y0, A, B);
      if (MEDIAN) B.w = depths[id];
      sm.a[tid] = A; sm.b[tid] = B;
      float4 fv = make_float4(0.f, 0.f, 0.f, 0.f);
      const float *fp = features + (int64_t)F * id;
      fv.x = fp[0];
      if (F > 1) fv.y = fp[1];
      if (F > 2) fv.z = fp[2];
      if (F > 3) fv.w = fp[3];
      sm.f[tid] = fv;
      sm.mask[tid] = (unsigned char)m;
      if (VIS) { sm.id[tid] = id; sm.vis[tid] = 0.f; }
    }
    __syncthreads();

    // per-warp ordered compaction of the splats that can touch this warp's 8x4 pixels
    int nhit = 0;
    if (!__all_sync(0xffffffffu, done)) {
      for (int c = 0; c < nb; c += 32) {
        int j = c + lane;
        bool hit = j < nb && ((sm.mask[j] >> warp) & 1);
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) sm.list[warp][nhit + __popc(bal & ((1u << lane) - 1))] = (unsigned short)j;
        nhit += __popc(bal);
      }
      if (lane < kUnroll) sm.list[warp][nhit + lane] = (unsigned short)kBatch;  // pad with the null record
      __syncwarp();
    }

    for (int h0 = 0; h0 < nhit; h0 += kUnroll) {
      float wv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int j = sm.list[warp][h0 + u];
        const float4 A = sm.a[j], B = sm.b[j];
        const float4 fv = sm.f[j];
#if GS_PACKED
        float tx, ty;
        upk(fma2(lx2, pk(A.z, A.w), fma2(ly2, pk(B.x, B.y), pk(A.x, A.y))), tx, ty);
#else
        float dx = fx - A.x, dy = fy - A.y;
        float tx = dx * A.z + dy * A.w, ty = dx * B.x + dy * B.y;
#endif
        float ga = ex2_approx(-(tx * tx + ty * ty));
        float alpha = fminf(B.z * ga, clamp_max);
        const bool hit = alpha > thr && !done;
        float weight = alpha * (1.0f - total_weight);
        weight = hit ? weight : 0.f;
        if (MEDIAN) median = (hit && total_weight < median_lim) ? B.w : median;   // last splat entered below the limit
        total_weight += weight;
        if (BLEND) {
#if GS_PACKED
          if (F == 1) accum[0] = fmaf(fv.x, weight, accum[0]);
          if (F >= 2) {
            const f32x2 w2 = pk(weight, weight);
            upk(fma2(pk(fv.x, fv.y), w2, pk(accum[0], accum[F > 1 ? 1 : 0])), accum[0], accum[F > 1 ? 1 : 0]);
            if (F == 3) accum[F > 2 ? 2 : 0] = fmaf(fv.z, weight, accum[F > 2 ? 2 : 0]);
            if (F == 4)
              upk(fma2(pk(fv.z, fv.w), w2, pk(accum[F > 2 ? 2 : 0], accum[F > 3 ? 3 : 0])), accum[F > 2 ? 2 : 0],
                  accum[F > 3 ? 3 : 0]);
          }
#else
          accum[0] = fmaf(fv.x, weight, accum[0]);
          if (F > 1) accum[1] = fmaf(fv.y, weight, accum[1]);
          if (F > 2) accum[2] = fmaf(fv.z, weight, accum[2]);
          if (F > 3) accum[3] = fmaf(fv.w, weight, accum[3]);
#endif
          done = done || (1.0f - total_weight <= eps);   // eps == 0: never (total_weight < 1 in bounds)
        } else {
          const bool trig = hit && total_weight >= sat_lim;
          accum[0] = trig ? fv.x : accum[0];
          if (F > 1) accum[1] = trig ? fv.y : accum[1];
          if (F > 2) accum[2] = trig ? fv.z : accum[2];
          if (F > 3) accum[3] = trig ? fv.w : accum[3];
          done = done || trig;
        }
        wv[u] = weight;
      }
      if (VIS) {
        warp_transpose_reduce16(wv, lane);
        const int h = h0 + (lane >> 1);
        if ((lane & 1) == 0 && h < nhit && wv[0] != 0.f) atomicAdd(&sm.vis[sm.list[warp][h]], wv[0]);
      }
      if (__all_sync(0xffffffffu, done)) break;
    }
    if (__all_sync(0xffffffffu, done) && lane == 0) sm.warp_done[warp] = 1;

    if (VIS) {
      __syncthreads();
      if (tid < nb) {
        float vsum = sm.vis[tid];
        if (vsum != 0.f) atomicAdd(visibility + sm.id[tid], vsum);
      }
    }
  }

  if (in_bounds) {
    float *out = image + ((int64_t)py * P.width + px) * F;
#pragma unroll
    for (int c = 0; c < F; ++c) out[c] = accum[c];
    image_alpha[(int64_t)py * P.width + px] = BLEND ? total_weight : (total_weight > 0.f ? 1.f : 0.f);
    // the splat that crossed the limit is the last one entered below it -- if the limit was crossed at all
    if (MEDIAN) median_image[(int64_t)py * P.width + px] = total_weight >= median_lim ? median : 0.f;
  }
}

template <int F>
static int launch_fwd(const float *points, const float *features, const float *depths, const int32_t *ranges,
                      const int32_t *o2p, const RasterParams<float> &P, float median_lim, int tiles, float *image,
                      float *image_alpha, float *visibility, float *median_image, cudaStream_t stream) {
  const bool vis = P.vis && visibility != nullptr;
  const bool med = median_image != nullptr;
#define GS_FWD(VIS, BLEND, MED)                                                                              \
  raster_fwd_kernel<F, VIS, BLEND, MED><<<tiles, kBatch, 0, stream>>>(points, features, depths, ranges, o2p, \
                                                                      P, median_lim, image, image_alpha,    \
                                                                      visibility, median_image)
  if (P.blend) {
    if (med) { if (vis) GS_FWD(true, true, true); else GS_FWD(false, true, true); }
    else     { if (vis) GS_FWD(true, true, false); else GS_FWD(false, true, false); }
  } else {
    if (vis) GS_FWD(true, false, false); else GS_FWD(false, false, false);
  }
#undef GS_FWD
  GS_LAUNCH_CHECK();
  return GS_OK;
}

static int raster_fwd_f32_impl(const float *points, const float *features, const float *depths,
                               const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t v, int32_t width,
                               int32_t height, int32_t F, const gs_raster_config *cfg, double median_threshold,
                               float *image, float *image_alpha, float *visibility, float *median_image,
                               cudaStream_t stream) {
  GS_CHECK_ARG(cfg != nullptr, "raster_fwd: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_fwd: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_visibility || visibility != nullptr || v == 0, "raster_fwd: compute_visibility needs a visibility buffer");
  const bool fast = cfg->tile_size == kTile && !cfg->antialias && F >= 1 && F <= 4;
  if (median_image != nullptr) {
    if (!fast || !cfg->use_alpha_blending) {
      set_error("raster_fwd_median: needs tile_size 16, no antialias, 1..4 features, alpha blending and depths");
      return GS_ERR_UNSUPPORTED;
    }
  }
  if (fast) {
    RasterParams<float> P = make_params<float>(cfg, width, height, F);
    int tiles = P.tiles_wide * ((height + kTile - 1) / kTile);
    float median_lim = (float)(1.0 - median_threshold);
    switch (F) {
      case 1: return launch_fwd<1>(points, features, depths, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
      case 2: return launch_fwd<2>(points, features, depths, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
      case 3: return launch_fwd<3>(points, features, depths, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
      default: return launch_fwd<4>(points, features, depths, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
    }
  }
  return raster_fwd_generic<float>(points, features, tile_ranges, overlap_to_point, width, height, F, cfg, image,
                                   image_alpha, visibility, stream);
}

}  // namespace gs

extern "C" int gs_raster_fwd_f32(const float *points, const float *features, const int32_t *tile_ranges,
                                 const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width,
                                 int32_t height, int32_t F, const gs_raster_config *cfg, float *image,
                                 float *image_alpha, float *visibility, void *stream_) {
  (void)k;
  return gs::raster_fwd_f32_impl(points, features, nullptr, tile_ranges, overlap_to_point, v, width, height, F, cfg, 0.0,
                                 image, image_alpha, visibility, nullptr, (cudaStream_t)stream_);
}

extern "C" int gs_raster_fwd_median_f32(const float *points, const float *features, const float *depths,
                                        const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t v,
                                        int64_t k, int32_t width, int32_t height, int32_t F,
                                        const gs_raster_config *cfg, double median_threshold, float *image,
                                        float *image_alpha, float *visibility, float *median_image, void *stream_) {
  (void)k;
  GS_CHECK_ARG(median_image != nullptr && (depths != nullptr || v == 0), "raster_fwd_median: depths / median_image is NULL");
  return gs::raster_fwd_f32_impl(points, features, depths, tile_ranges, overlap_to_point, v, width, height, F, cfg,
                                 median_threshold, image, image_alpha, visibility, median_image, (cudaStream_t)stream_);
}

extern "C" int gs_raster_fwd_f64(const double *points, const double *features, const int32_t *tile_ranges,
                                 const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width,
                                 int32_t height, int32_t F, const gs_raster_config *cfg, double *image,
                                 double *image_alpha, double *visibility, void *stream_) {
  GS_CHECK_ARG(cfg != nullptr, "raster_fwd: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_fwd: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_visibility || visibility != nullptr || v == 0, "raster_fwd: compute_visibility needs a visibility buffer");
  (void)k;
  return gs::raster_fwd_generic<double>(points, features, tile_ranges, overlap_to_point, width, height, F, cfg, image,
                                        image_alpha, visibility, (cudaStream_t)stream_);
}
