// raster_bwd.cu -- R9: backward gradient sweep, tuned fp32 / 16x16-tile kernel + C ABI dispatch.
//
// Semantics: _backward_kernel, rasterizer/backward.py:50-225 and gaussian_pdf_with_grad,
// taichi_lib/generic.py:320-336: re-walk front to back with (total_weight, remaining = image - sum f w),
// stop a pixel at total_weight >= saturate_threshold, clamp passes gradient through (D4).
//
// B200 design (not the reference's):
//   * same CTA / warp-rectangle / staged-record / per-warp hit-list structure as raster_fwd.cu;
//   * per (pixel, splat) the only splat-independent quantities formed are the six tile-local moments of
//     Gp = alpha_point * dL/dalpha * pdf  over the pixel offset l from the tile centre,
//         {1, lx, ly, lx^2, lx ly, ly^2} * Gp,
//     plus weight * dL/dimage (feature gradient) and the two densification heuristics.  All seven
//     parameter gradients are linear in the moment sums (the pdf is exp of a quadratic form in the pixel
//     position), so axis / sigma / mean enter once per (splat, tile) at flush time, and
//     dL/dalpha_point = M0 / alpha_point;
//   * because the moment multipliers {1, lx, ...} are per-lane constants, each lane keeps them in a
//     lane-dependent (XOR-permuted) register order, which makes the transposed-butterfly warp reduction
//     select-free: v[r] += shfl_xor(v[r + half], off); the three value types (8 moment slots, 4 feature slots,
//     2 heuristics) then share one 3-shuffle tail -- 15 shuffles for 11 sums, against 5 per value for a tree;
//   * the inner loop is branch-free; one shared-memory atomic instruction per (warp, splat); one global
//     float atomic per (splat, tile, component) at the end of each 256-splat batch.
#include <stdlib.h>
#include <string.h>

#include "raster_common.cuh"

namespace gs {

template <typename real>
int raster_bwd_generic(const real *points, const real *features, const int32_t *ranges, const int32_t *o2p,
                       const real *image, const real *grad_image, int width, int height, int F,
                       const gs_raster_config *cfg, real *grad_points, real *grad_features, real *heuristic,
                       cudaStream_t stream);

#ifndef GS_BWD_UNROLL
#define GS_BWD_UNROLL 1
#endif
constexpr int kTileB = 16;
constexpr int kBatchB = 256;
constexpr float kExpScaleB = 0.84932180028801904f;

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// accumulator slots per staged splat: [M0, Lx, Ly, Lxx, Lxy, Lyy | f0..f(F-1) | h0, h1], odd stride
template <int F> struct AccLayout {
  static constexpr int kFeat = 6, kHeur = 6 + F, kUsed = 8 + F, kStride = kUsed | 1;
};

template <int F>
struct BwdSmem {
  float4 a[kBatchB];  // mean.x, mean.y, (axis/sx)*k
  float4 b[kBatchB];  // (perp/sy)*k, alpha, unused
  float4 f[kBatchB];
  float acc[kBatchB * AccLayout<F>::kStride];
  unsigned char mask[kBatchB];
  unsigned short list[8][kBatchB];
  int warp_done[8];
};

#ifndef GS_BWD_MIN_BLOCKS
#define GS_BWD_MIN_BLOCKS 4
#endif
template <int F, bool GP, bool GF, bool HEUR>
__global__ void __launch_bounds__(kBatchB, GS_BWD_MIN_BLOCKS)
raster_bwd_kernel(const float *__restrict__ points, const float *__restrict__ features,
                  const int32_t *__restrict__ ranges, const int32_t *__restrict__ overlap_to_point,
                  const float *__restrict__ image, const float *__restrict__ grad_image, RasterParams<float> P,
                  float *__restrict__ grad_points, float *__restrict__ grad_features,
                  float *__restrict__ heuristic) {
  using L = AccLayout<F>;
  __shared__ BwdSmem<F> sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % P.tiles_wide) * kTileB, tile_y0 = (tile / P.tiles_wide) * kTileB;
  const int lxi = (warp & 1) * 8 + (lane & 7), lyi = (warp >> 1) * 4 + (lane >> 3);
  const int px = tile_x0 + lxi, py = tile_y0 + lyi;
  const bool in_bounds = px < P.width && py < P.height;
  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
  const float clamp_max = P.clamp_max, thr = P.thr, sat = P.sat;

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

  // ---- lane-constant, XOR-permuted multipliers (see header) ----
  const int b16 = (lane >> 4) & 1, b8 = (lane >> 3) & 1, b4 = (lane >> 2) & 1;
  const int mA = (b16 << 2) | (b8 << 1) | b4;   // this lane ends up owning moment index mA
  const int mB = (b16 << 1) | b8;               // ... and feature slot mB
  // coefA[p] = c[p ^ mA], gpixB[p] = gpix[p ^ mB]: XOR by a bit = conditional swap of register pairs
  float coefA[8];
  {
    const float lx = (float)lxi - 7.5f, ly = (float)lyi - 7.5f;   // pixel centre relative to the tile centre
    coefA[0] = 1.f; coefA[1] = lx; coefA[2] = ly; coefA[3] = lx * lx; coefA[4] = lx * ly; coefA[5] = ly * ly;
    coefA[6] = 0.f; coefA[7] = 0.f;
#pragma unroll
    for (int bit = 1; bit <= 4; bit <<= 1) {
      const bool sw = (mA & bit) != 0;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if ((p & bit) == 0) {
          float lo = coefA[p], hi = coefA[p | bit];
          coefA[p] = sw ? hi : lo;
          coefA[p | bit] = sw ? lo : hi;
        }
    }
  }
  float gpixB[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) gpixB[p] = p < F ? gpix[p < F ? p : 0] : 0.f;
#pragma unroll
  for (int bit = 1; bit <= 2; bit <<= 1) {
    const bool sw = (mB & bit) != 0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if ((p & bit) == 0) {
        float lo = gpixB[p], hi = gpixB[p | bit];
        gpixB[p] = sw ? hi : lo;
        gpixB[p | bit] = sw ? lo : hi;
      }
  }
  // which accumulator slot this lane adds after the reductions (-1: none)
  const int b2 = (lane >> 1) & 1;
  int my_slot = -1;
  if ((lane & 1) == 0) {
    if (b2 == 0) my_slot = GP && mA < 6 ? mA : -1;                          // moments
    else if (b4 == 0) my_slot = GF && mB < F ? L::kFeat + mB : -1;          // feature gradients
    else if (b8 == 0) my_slot = HEUR ? L::kHeur + b16 : -1;                 // heuristics
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
    float s_mx = 0.f, s_my = 0.f, s_ax = 0.f, s_ay = 0.f, s_isx = 0.f, s_isy = 0.f, s_alpha = 1.f;
    if (tid < nb) {
      my_id = overlap_to_point[base + tid];
      const float *g = points + 7 * (int64_t)my_id;
      float mx = g[0], my = g[1], ax = g[2], ay = g[3], sx = g[4], sy = g[5], alpha = g[6];
      float isx = 1.0f / sx, isy = 1.0f / sy;
      s_mx = mx; s_my = my; s_ax = ax; s_ay = ay; s_isx = isx; s_isy = isy; s_alpha = alpha;
      float ux = ax * isx * kExpScaleB, uy = ay * isx * kExpScaleB;
      float wx = -ay * isy * kExpScaleB, wy = ax * isy * kExpScaleB;
      sm.a[tid] = make_float4(mx, my, ux, uy);
      sm.b[tid] = make_float4(wx, wy, alpha, 0.f);
      unsigned mask = 0;
      if (alpha > thr) {
        float rc = sqrtf(2.0f * __logf(alpha / thr)) * 1.001f + 0.01f;
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
      for (int c = 0; c < L::kUsed; ++c) sm.acc[tid * L::kStride + c] = 0.f;
    }
    __syncthreads();

    // ---- per-warp ordered hit list ----
    int nhit = 0;
    if (!__all_sync(0xffffffffu, total_weight >= sat)) {
      for (int c = 0; c < nb; c += 32) {
        int j = c + lane;
        bool hit = j < nb && ((sm.mask[j] >> warp) & 1);
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) sm.list[warp][nhit + __popc(bal & ((1u << lane) - 1))] = (unsigned short)j;
        nhit += __popc(bal);
      }
      __syncwarp();
    }

    // ---- gradient sweep (branch-free body) ----
    const unsigned full = 0xffffffffu;
    constexpr int kUnrollB = GS_BWD_UNROLL;
#pragma unroll kUnrollB
    for (int h = 0; h < nhit; ++h) {
      const int j = sm.list[warp][h];
      const float4 A = sm.a[j], B = sm.b[j];
      const float4 fv = sm.f[j];
      const float feat[4] = {fv.x, fv.y, fv.z, fv.w};
      float dx = fx - A.x, dy = fy - A.y;
      float tx = dx * A.z + dy * A.w, ty = dx * B.x + dy * B.y;
      float ga = ex2_approx_b(-(tx * tx + ty * ty));
      float alpha = B.z * ga;
      const bool has_grad = alpha > thr && total_weight < sat;
      alpha = fminf(alpha, clamp_max);
      float T_i = 1.0f - total_weight;
      float weight = has_grad ? alpha * T_i : 0.f;
      total_weight += weight;
      float inv_1ma = rcp_approx(1.0f - alpha);
      float alpha_grad = 0.f;
#pragma unroll
      for (int c = 0; c < F; ++c) {
        remaining[c] = fmaf(-feat[c], weight, remaining[c]);
        float diff = fmaf(-remaining[c], inv_1ma, feat[c] * T_i);
        alpha_grad = fmaf(diff, gpix[c], alpha_grad);
      }
      float G = has_grad ? B.z * alpha_grad : 0.f;
      float Gp = G * ga;

      // type-specific halving stages (select-free thanks to the XOR-permuted multipliers) ...
      float ra = 0.f, rb = 0.f, rc = 0.f;
      if (GP) {   // six moments (two spare slots): 4 + 2 + 1 shuffles -> sum over 8 lanes of moment mA
        float v0 = Gp * coefA[0], v1 = Gp * coefA[1], v2 = Gp * coefA[2], v3 = Gp * coefA[3];
        float v4 = Gp * coefA[4], v5 = Gp * coefA[5], v6 = Gp * coefA[6], v7 = Gp * coefA[7];
        v0 += __shfl_xor_sync(full, v4, 16); v1 += __shfl_xor_sync(full, v5, 16);
        v2 += __shfl_xor_sync(full, v6, 16); v3 += __shfl_xor_sync(full, v7, 16);
        v0 += __shfl_xor_sync(full, v2, 8); v1 += __shfl_xor_sync(full, v3, 8);
        ra = v0 + __shfl_xor_sync(full, v1, 4);
      }
      if (GF) {   // weight * dL/dimage: 2 + 1 shuffles -> sum over 4 lanes of feature slot mB
        float v0 = weight * gpixB[0], v1 = weight * gpixB[1], v2 = weight * gpixB[2], v3 = weight * gpixB[3];
        v0 += __shfl_xor_sync(full, v2, 16); v1 += __shfl_xor_sync(full, v3, 16);
        rb = v0 + __shfl_xor_sync(full, v1, 8);
      }
      if (HEUR) {  // [(alpha dL/dalpha)^2, |alpha dL/dalpha dpdf/dmean|_1]: 1 + 1 shuffles -> sum over 4 lanes
        const float inv_k2 = 1.0f / (kExpScaleB * kExpScaleB);
        float a1 = Gp * tx, a2 = Gp * ty;   // each carries one exp-scale factor k, as do A.zw / B.xy
        float h0 = G * G;
        float h1 = (fabsf(a1 * A.z + a2 * B.x) + fabsf(a1 * A.w + a2 * B.y)) * inv_k2;
        float v0 = b16 ? h1 : h0, v1 = b16 ? h0 : h1;
        v0 += __shfl_xor_sync(full, v1, 16);
        rc = v0 + __shfl_xor_sync(full, v0, 8);
      }
      // ... then ONE shared tail for the three partial results: at each remaining lane bit two value types are
      // exchanged transposed (keep one, send the other), so 3 shuffles finish all of them.
      float x = (b4 ? rc : rb) + __shfl_xor_sync(full, b4 ? rb : rc, 4);   // b4=0 lanes: features, b4=1: heuristics
      float y = (b2 ? x : ra) + __shfl_xor_sync(full, b2 ? ra : x, 2);      // b2=0 lanes: moments,  b2=1: x
      const float add_val = y + __shfl_xor_sync(full, y, 1);
      if (my_slot >= 0 && add_val != 0.f) atomicAdd(&sm.acc[j * L::kStride + my_slot], add_val);
      if (__all_sync(full, total_weight >= sat)) break;
    }
    if (__all_sync(full, total_weight >= sat) && lane == 0) sm.warp_done[warp] = 1;

    // ---- flush: one thread per staged splat ----
    __syncthreads();
    if (tid < nb) {
      float S[L::kUsed];
      bool any = false;
#pragma unroll
      for (int c = 0; c < L::kUsed; ++c) { S[c] = sm.acc[tid * L::kStride + c]; any |= (S[c] != 0.f); }
      if (any) {
        if (GP) {
          // shift the tile-centred moments to the splat mean: d = l + c
          const float cx = (float)tile_x0 + 8.0f - s_mx, cy = (float)tile_y0 + 8.0f - s_my;
          const float M0 = S[0], Lx = S[1], Ly = S[2], Lxx = S[3], Lxy = S[4], Lyy = S[5];
          const float Mx = fmaf(cx, M0, Lx), My = fmaf(cy, M0, Ly);
          const float Mxx = Lxx + cx * (2.0f * Lx + cx * M0);
          const float Myy = Lyy + cy * (2.0f * Ly + cy * M0);
          const float Mxy = Lxy + cx * Ly + cy * Lx + cx * cy * M0;
          const float ux = s_ax * s_isx, uy = s_ay * s_isx, wx = -s_ay * s_isy, wy = s_ax * s_isy;
          const float S1 = ux * Mx + uy * My, S2 = wx * Mx + wy * My;             // sum Gp tx, sum Gp ty
          const float S3 = ux * Mxx + uy * Mxy, S4 = ux * Mxy + uy * Myy;         // sum Gp tx dx, sum Gp tx dy
          const float S5 = wx * Mxx + wy * Mxy, S6 = wx * Mxy + wy * Myy;         // sum Gp ty dx, sum Gp ty dy
          float *gp = grad_points + 7 * (int64_t)my_id;
          atomicAdd(gp + 0, S1 * ux + S2 * wx);
          atomicAdd(gp + 1, S1 * uy + S2 * wy);
          atomicAdd(gp + 2, -s_isx * S3 - s_isy * S6);
          atomicAdd(gp + 3, -s_isx * S4 + s_isy * S5);
          atomicAdd(gp + 4, s_isx * (ux * S3 + uy * S4));
          atomicAdd(gp + 5, s_isy * (wx * S5 + wy * S6));
          atomicAdd(gp + 6, M0 / s_alpha);
        }
        if (GF) {
          float *gf = grad_features + (int64_t)F * my_id;
#pragma unroll
          for (int c = 0; c < F; ++c) atomicAdd(gf + c, S[L::kFeat + c]);
        }
        if (HEUR) {
          atomicAdd(heuristic + 2 * (int64_t)my_id, S[L::kHeur]);
          atomicAdd(heuristic + 2 * (int64_t)my_id + 1, S[L::kHeur + 1]);
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

// raster_bwd_t.cu: same contract, shared-memory transpose instead of per-splat shuffle reduction
template <int F>
int launch_bwd_transpose(const float *points, const float *features, const int32_t *ranges, const int32_t *o2p,
                         const float *image, const float *grad_image, const RasterParams<float> &P, int tiles,
                         float *grad_points, float *grad_features, float *heuristic, cudaStream_t stream);

static bool use_transpose_kernel() {
  static int choice = -1;
  if (choice < 0) {
    const char *e = getenv("GS_BWD_KERNEL");   // "shuffle" | "transpose" (A/B switch for profiling)
    choice = (e != nullptr && strcmp(e, "shuffle") == 0) ? 0 : 1;
  }
  return choice == 1;
}

template <int F>
static int launch_bwd_any(const float *points, const float *features, const int32_t *ranges, const int32_t *o2p,
                          const float *image, const float *grad_image, const RasterParams<float> &P, int tiles,
                          float *grad_points, float *grad_features, float *heuristic, cudaStream_t stream) {
  if (use_transpose_kernel())
    return launch_bwd_transpose<F>(points, features, ranges, o2p, image, grad_image, P, tiles, grad_points,
                                   grad_features, heuristic, stream);
  return launch_bwd<F>(points, features, ranges, o2p, image, grad_image, P, tiles, grad_points, grad_features,
                       heuristic, stream);
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
  GS_CHECK_ARG(!cfg->compute_point_heuristic || point_heuristic != nullptr || v == 0, "raster_bwd: compute_point_heuristic needs a buffer");
  (void)k;
  if (cfg->tile_size == gs::kTileB && !cfg->antialias && F >= 1 && F <= 4) {
    gs::RasterParams<float> P = gs::make_params<float>(cfg, width, height, F);
    int tiles = P.tiles_wide * ((height + gs::kTileB - 1) / gs::kTileB);
    switch (F) {
      case 1: return gs::launch_bwd_any<1>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
      case 2: return gs::launch_bwd_any<2>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
      case 3: return gs::launch_bwd_any<3>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
      default: return gs::launch_bwd_any<4>(points, features, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
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
  GS_CHECK_ARG(!cfg->compute_point_heuristic || point_heuristic != nullptr || v == 0, "raster_bwd: compute_point_heuristic needs a buffer");
  (void)k;
  return gs::raster_bwd_generic<double>(points, features, tile_ranges, overlap_to_point, image, grad_image, width,
                                        height, F, cfg, grad_points, grad_features, point_heuristic,
                                        (cudaStream_t)stream_);
}
