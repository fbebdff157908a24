// raster_fwd.cu -- R8: front-to-back alpha compositing, tuned fp32 / 16x16-tile kernel + C ABI dispatch.
//
// Semantics: _forward_kernel, rasterizer/forward.py:22-135 (each overlap composited exactly once, D1;
// forward early-out only below config.forward_saturate_eps, D2).  The optional fused median-depth output
// reproduces the reference's second raster pass (renderer.py:77-82: use_alpha_blending=False,
// saturate_threshold=median_threshold, features=depths) from the same walk.
//
// B200 design (not the reference's; measurements and the reasoning behind each point: DESIGN.md section 4):
//   * the kernel is bound by the L1/shared-memory data pipe -- a broadcast LDS.128 of a splat record costs 2.69
//     pipe cycles -- so one CTA per 16x16 tile runs 4 warps, each owning an 8x8 pixel block with TWO pixels per lane
//     (column lane & 7, rows lane >> 3 and + 4): one record load feeds 64 pixel evaluations;
//   * splats are gathered from the 64-byte raster digest (raster_digest.cu, four LDG.128) and staged 256 at a time
//     into shared memory as tile-centred 16-byte records {tx0, ty0, ux, wx} {uy, wy, alpha, depth} {features}:
//     (tx, ty) = X (ux, wx) + Y (uy, wy) + (tx0, ty0) is three packed FFMA2 for both pixels of a lane, and the
//     rest of the per-pixel arithmetic runs as FMUL2 / FFMA2 / FADD2 on (pixel 0, pixel 1) register pairs;
//   * the staging thread classifies its splat against the four 8x8 blocks (separating-axis test against the
//     oriented box of the support ellipse, conservative) and each warp compacts its own ordered hit list, so a warp
//     only iterates over splats that can exceed the alpha threshold somewhere in its 64 pixels.  Skipped splats
//     contribute exactly zero in the reference too (alpha <= threshold), so results are unchanged;
//   * per-pixel state is the transmittance T (w = alpha T, T -= w; T = 0 outside the image), so the inner loop has no
//     bounds / done predicate; it is branch-free and unrolled by 16, with the median-depth bookkeeping compiled out
//     once every pixel of the warp has crossed the median limit; per-splat visibility (sum of blend weights over
//     pixels) is reduced 16 splats at a time with one transposed butterfly (16 shuffles per 16 splats instead of 5
//     per splat) and one shared-memory atomic instruction per 16 splats.
#include <stdlib.h>

#include <type_traits>

#include "packed_f32.cuh"
#include "raster_common.cuh"


namespace gs {

template <typename real>
int raster_fwd_generic(const real *points, const real *features, const int32_t *ranges, const int32_t *o2p,
                       int width, int height, int F, const gs_raster_config *cfg, real *image, real *image_alpha,
                       real *visibility, cudaStream_t stream);

constexpr int kTile = 16;
constexpr int kBatch = 256;
constexpr int kWarps = 4;             // one warp per 8x8 pixel block of the tile
constexpr int kThreads = kWarps * 32;
#ifndef GS_EXACT_CULL
#define GS_EXACT_CULL 0   // exact block-vs-ellipse test after the box tests: 7.72 -> 7.33 hit entries per Gaussian, but the
#endif                    // forward loses more in staging than it saves in the sweep (0.416 -> 0.439 ms); the backward uses it
#ifndef GS_FWD_UNROLL
#define GS_FWD_UNROLL 16
#endif
constexpr int kUnroll = GS_FWD_UNROLL;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct FwdSmem {
  float4 a[kBatch + 1];    // tx0, ty0, ux, wx   (t = (tx, ty) at the tile centre; u = axis/sx*k, w = perp/sy*k)
  float4 b[kBatch + 1];    // uy, wy, alpha, depth                     (+1: null record for list padding)
  float4 f[kBatch + 1];    // features (F <= 4)
  int id[kBatch];
  float vis[kBatch + 1];
  unsigned char mask[kBatch];
  alignas(16) unsigned list[kWarps][kBatch + kUnroll];   // byte offsets (16 j) of the records a warp must visit
};

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Stage one splat from its digest (raster_digest.cu): tile-centred records and the 4-bit mask of 8x8 pixel blocks
// it can touch.  (tx, ty) = X (ux, wx) + Y (uy, wy) + (tx0, ty0) with (X, Y) the pixel centre relative to the
// tile centre.  The block test is the separating-axis test of the block against the oriented box of the support
// ellipse (conservative: tile axes via the ellipse's bounding box, then the two ellipse axes).
__device__ __forceinline__ unsigned stage_splat(const float4 *__restrict__ rec, float tile_cx, float tile_cy,
                                                float4 &A, float4 &B, float4 &fv) {
  const float4 R0 = __ldg(rec), R1 = __ldg(rec + 1), R2 = __ldg(rec + 2), R3 = __ldg(rec + 3);
  const float ux = R0.z, wx = R0.w, uy = R1.x, wy = R1.y, rcs = R3.x;
  const float ddx = R0.x - tile_cx, ddy = R0.y - tile_cy;
  const float tx0 = -fmaf(ux, ddx, uy * ddy), ty0 = -fmaf(wx, ddx, wy * ddy);
  A = make_float4(tx0, ty0, ux, wx);
  B = R1;
  fv = R2;
  if (!(rcs > 0.f)) return 0u;
  // half extents of the support ellipse { |U d| <= rcs }, U = [u; w]:  rcs sqrt(uy^2 + wy^2) / |det U| in x
  const float s = rcs * rcp_approx(fabsf(ux * wy - uy * wx)) * 1.0001f;
  const float ex = s * sqrtf(fmaf(uy, uy, wy * wy)), ey = s * sqrtf(fmaf(ux, ux, wx * wx));
  const float hu = (fabsf(ux) + fabsf(uy)) * 3.5f + rcs;     // pixel centres of a block span +-3.5 around its centre
  const float hw = (fabsf(wx) + fabsf(wy)) * 3.5f + rcs;
#if GS_EXACT_CULL
  const SupportMetric metric = support_metric(ux, wx, uy, wy, rcs);
#endif
  unsigned mask = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const float ox = (w & 1) ? 4.0f : -4.0f, oy = (w >> 1) ? 4.0f : -4.0f;   // block centre - tile centre
    const float t0x = fmaf(ux, ox, fmaf(uy, oy, tx0)), t0y = fmaf(wx, ox, fmaf(wy, oy, ty0));
    bool hit = (fabsf(ox - ddx) - 3.5f <= ex) && (fabsf(oy - ddy) - 3.5f <= ey) && (fabsf(t0x) <= hu) && (fabsf(t0y) <= hw);
#if GS_EXACT_CULL
    if (hit) hit = block_reaches_support(metric, t0x, t0y, ux, wx, uy, wy);   // the box tests leave ~6 % corner cases
#endif
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
#define GS_FWD_MIN_BLOCKS 8
#endif
// Every lane owns TWO pixels of its warp's 8x8 block -- column (lane & 7), rows (lane >> 3) and (lane >> 3) + 4 --
// so one broadcast LDS.128 of a splat record feeds 64 pixel evaluations (the kernel is bound by the shared-memory
// data pipe, ~2.7 cycles per broadcast LDS.128), and the two pixels share packed FFMA2 / FMUL2 / FADD2 issue slots.
template <int F, bool VIS, bool BLEND, bool MEDIAN>
__global__ void __launch_bounds__(kThreads, GS_FWD_MIN_BLOCKS)
raster_fwd_kernel(const float4 *__restrict__ digest, const int32_t *__restrict__ ranges,
                  const int32_t *__restrict__ overlap_to_point, RasterParams<float> P, float median_lim,
                  float *__restrict__ image, float *__restrict__ image_alpha, float *__restrict__ visibility,
                  float *__restrict__ median_image) {
  __shared__ FwdSmem sm;
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % P.tiles_wide) * kTile, tile_y0 = (tile / P.tiles_wide) * kTile;
  const int bx = (warp & 1) * 8 + (lane & 7), by = (warp >> 1) * 8 + (lane >> 3);
  const int px = tile_x0 + bx, py[2] = {tile_y0 + by, tile_y0 + by + 4};
  const bool in_bounds[2] = {px < P.width && py[0] < P.height, px < P.width && py[1] < P.height};
  const float lx = (float)bx - 7.5f, ly0 = (float)by - 7.5f;
  const f32x2 lx2 = pk(lx, lx), ly2[2] = {pk(ly0, ly0), pk(ly0 + 4.0f, ly0 + 4.0f)};
  const float clamp_max = P.clamp_max, thr = P.thr, eps = P.fwd_eps;
  const float median_trans = 1.0f - median_lim;   // sum of weights < lim  <=>  transmittance > 1 - lim
  const float sat_trans = P.sat;                  // non-blend mode: sum of weights >= 1 - sat <=> transmittance <= sat

  float accum[2][F];
#pragma unroll
  for (int c = 0; c < F; ++c) accum[0][c] = accum[1][c] = 0.f;
  // transmittance 1 - sum of weights; 0 outside the image, where nothing can contribute (forward.py:49-52)
  float trans[2] = {in_bounds[0] ? 1.f : 0.f, in_bounds[1] ? 1.f : 0.f};
  bool done[2] = {!in_bounds[0], !in_bounds[1]};   // non-blend mode only: pixel frozen after its trigger
  float median[2] = {0.f, 0.f};

  const unsigned char *rec_a = reinterpret_cast<const unsigned char *>(sm.a);
  const unsigned char *rec_b = reinterpret_cast<const unsigned char *>(sm.b);
  const unsigned char *rec_f = reinterpret_cast<const unsigned char *>(sm.f);
  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  int warp_is_done = 0;   // nothing can change any more in this warp's block (uniform over the warp)
  if (tid == 0) {  // null record: alpha = 0 never passes the threshold
    sm.a[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.b[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.f[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int base = start; base < end; base += kBatch) {
    const int nb = min(kBatch, end - base);
    // previous batch fully consumed; the barrier also decides, identically for every thread, whether the tile is done
    // (a shared flag per warp re-read after the barrier raced with fast warps setting theirs again: racecheck)
    if (__syncthreads_and(warp_is_done)) break;
    for (int j = tid; j < nb; j += kThreads) {
      int id = overlap_to_point[base + j];
      float4 A, B, fv;
      unsigned m = stage_splat(digest + 4 * (int64_t)id, (float)tile_x0 + 8.0f, (float)tile_y0 + 8.0f, A, B, fv);
      sm.a[j] = A; sm.b[j] = B;
      sm.f[j] = fv;
      sm.mask[j] = (unsigned char)m;
      if (VIS) { sm.id[j] = id; sm.vis[j] = 0.f; }
    }
    __syncthreads();

    // a warp is finished when neither of its pixels can change any more
    const bool lane_done = BLEND ? (trans[0] <= eps && trans[1] <= eps) : (done[0] && done[1]);

    // per-warp ordered compaction of the splats that can touch this warp's 8x8 pixels
    int nhit = 0;
    if (!__all_sync(full, lane_done)) {
      for (int c = 0; c < nb; c += 32) {
        int j = c + lane;
        bool hit = j < nb && ((sm.mask[j] >> warp) & 1);
        unsigned bal = __ballot_sync(full, hit);
        if (hit) sm.list[warp][nhit + __popc(bal & ((1u << lane) - 1))] = 16u * (unsigned)j;
        nhit += __popc(bal);
      }
      if (lane < kUnroll) sm.list[warp][nhit + lane] = 16u * (unsigned)kBatch;  // pad with the null record
      __syncwarp();
    }

    for (int h0 = 0; h0 < nhit; h0 += kUnroll) {
      float wv[kUnroll];
      unsigned offs[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; u += 4) {
        const uint4 o = *reinterpret_cast<const uint4 *>(&sm.list[warp][h0 + u]);
        offs[u] = o.x; offs[u + 1] = o.y; offs[u + 2] = o.z; offs[u + 3] = o.w;
      }
      // one unrolled chunk of the sweep; the median bookkeeping is compiled out once every pixel of the warp has
      // crossed the median limit (it happens within the first few splats of a pixel)
      auto sweep = [&](auto median_tag) {
      constexpr bool kMedian = decltype(median_tag)::value;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const unsigned off = offs[u];
        const float4 A = *reinterpret_cast<const float4 *>(rec_a + off);
        const float4 B = *reinterpret_cast<const float4 *>(rec_b + off);
        const float4 fv = *reinterpret_cast<const float4 *>(rec_f + off);
        const float feat[4] = {fv.x, fv.y, fv.z, fv.w};
        const f32x2 tbase = fma2(lx2, pk(A.z, A.w), pk(A.x, A.y)), uw_y = pk(B.x, B.y);
        float t0x, t0y, t1x, t1y;
        upk(fma2(ly2[0], uw_y, tbase), t0x, t0y);
        upk(fma2(ly2[1], uw_y, tbase), t1x, t1y);
        const float g0 = ex2_approx(-fmaf(t0x, t0x, t0y * t0y)), g1 = ex2_approx(-fmaf(t1x, t1x, t1y * t1y));
        float alpha[2], weight[2];
        upk(mul2(pk(g0, g1), pk(B.z, B.z)), alpha[0], alpha[1]);
        alpha[0] = fminf(alpha[0], clamp_max);
        alpha[1] = fminf(alpha[1], clamp_max);
        if (BLEND) {
          // no per-splat early-out test: a pixel below eps keeps compositing (as the reference does, D2) until
          // the whole warp is below eps; pixels outside the image have trans == 0 and so weight == 0.
          const bool hit[2] = {alpha[0] > thr, alpha[1] > thr};
          upk(mul2(pk(alpha[0], alpha[1]), pk(trans[0], trans[1])), weight[0], weight[1]);
          weight[0] = hit[0] ? weight[0] : 0.f;
          weight[1] = hit[1] ? weight[1] : 0.f;
          if (kMedian) {   // the splat that crosses the limit is the last one entered below it
            median[0] = (hit[0] && trans[0] > median_trans) ? B.w : median[0];
            median[1] = (hit[1] && trans[1] > median_trans) ? B.w : median[1];
          }
          upk(sub2(pk(trans[0], trans[1]), pk(weight[0], weight[1])), trans[0], trans[1]);
          const f32x2 w2 = pk(weight[0], weight[1]);
#pragma unroll
          for (int c = 0; c < F; ++c)
            upk(fma2(pk(feat[c], feat[c]), w2, pk(accum[0][c], accum[1][c])), accum[0][c], accum[1][c]);
        } else {
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const bool hit = alpha[p] > thr && !done[p];
            weight[p] = hit ? alpha[p] * trans[p] : 0.f;
            trans[p] -= weight[p];
            const bool trig = hit && trans[p] <= sat_trans;
#pragma unroll
            for (int c = 0; c < F; ++c) accum[p][c] = trig ? feat[c] : accum[p][c];
            done[p] = done[p] || trig;
          }
        }
        wv[u] = weight[0] + weight[1];
      }
      };
      if (MEDIAN && !__all_sync(full, trans[0] <= median_trans && trans[1] <= median_trans))
        sweep(std::true_type{});
      else
        sweep(std::false_type{});
      if (VIS) {
        warp_transpose_reduce16(wv, lane);
        const int h = h0 + (lane >> 1);
        if ((lane & 1) == 0 && h < nhit && wv[0] != 0.f) atomicAdd(&sm.vis[sm.list[warp][h] >> 4], wv[0]);
      }
      const bool now_done = BLEND ? (trans[0] <= eps && trans[1] <= eps) : (done[0] && done[1]);
      if (__all_sync(full, now_done)) break;
    }
    {
      const bool now_done = BLEND ? (trans[0] <= eps && trans[1] <= eps) : (done[0] && done[1]);
      warp_is_done = __all_sync(full, now_done);
    }

    if (VIS) {
      __syncthreads();
      for (int j = tid; j < nb; j += kThreads) {
        float vsum = sm.vis[j];
        if (vsum != 0.f) atomicAdd(visibility + sm.id[j], vsum);
      }
    }
  }

#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (!in_bounds[p]) continue;
    const int64_t pix = (int64_t)py[p] * P.width + px;
    float *out = image + pix * F;
#pragma unroll
    for (int c = 0; c < F; ++c) out[c] = accum[p][c];
    const float total_weight = 1.0f - trans[p];
    image_alpha[pix] = BLEND ? total_weight : (total_weight > 0.f ? 1.f : 0.f);
    if (MEDIAN) median_image[pix] = trans[p] <= median_trans ? median[p] : 0.f;
  }
}

template <int F>
static int launch_fwd(const float4 *digest, const int32_t *ranges, const int32_t *o2p, const RasterParams<float> &P,
                      float median_lim, int tiles, float *image, float *image_alpha, float *visibility,
                      float *median_image, cudaStream_t stream) {
  const bool vis = P.vis && visibility != nullptr;
  const bool med = median_image != nullptr;
#define GS_FWD(VIS, BLEND, MED)                                                                                  \
  raster_fwd_kernel<F, VIS, BLEND, MED><<<tiles, kThreads, 0, stream>>>(digest, ranges, o2p, P, median_lim, image, \
                                                                        image_alpha, visibility, median_image)
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

int raster_digest_f32(const float *points, const float *features, const float *depths, int64_t v, int F,
                      double alpha_threshold, void *digest, cudaStream_t stream);   // raster_digest.cu
int raster_pack_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t k,
                    int32_t width, int32_t height, int32_t F, void *records, void *flush, cudaStream_t stream,
                    const uint32_t *sorted_tiles = nullptr, int32_t *ranges_out = nullptr);   // raster_pack.cu

// GS_RASTER_STAGING=gather keeps the in-kernel digest gather of this file for alpha blending too (A/B switch);
// default: packed records + bulk-copy staging (raster_pack.cu, raster_fwd_bulk.cu).
static bool use_bulk_staging() {
  static int choice = -1;
  if (choice < 0) {
    const char *e = getenv("GS_RASTER_STAGING");
    choice = (e != nullptr && e[0] == 'g') ? 0 : 1;
  }
  return choice == 1;
}

static bool fwd_tuned(const gs_raster_config *cfg, int F) {
  return cfg->tile_size == kTile && !cfg->antialias && F >= 1 && F <= 4;
}

// Tuned kernel on a ready digest.
static int raster_fwd_digest_impl(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point,
                                  int64_t v, int64_t k, int32_t width, int32_t height, int32_t F,
                                  const gs_raster_config *cfg, double median_threshold, float *image,
                                  float *image_alpha, float *visibility, float *median_image, cudaStream_t stream,
                                  size_t prefix_bytes = 0) {
  GS_CHECK_ARG(cfg != nullptr, "raster_fwd: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_fwd: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_visibility || visibility != nullptr || v == 0, "raster_fwd: compute_visibility needs a visibility buffer");
  GS_CHECK_ARG(digest != nullptr || v == 0, "raster_fwd: digest is NULL");
  if (!fwd_tuned(cfg, F) || (median_image != nullptr && !cfg->use_alpha_blending)) {
    set_error("raster_fwd (digest): needs tile_size 16, no antialias, 1..4 features%s",
              median_image != nullptr ? ", alpha blending for the fused median" : "");
    return GS_ERR_UNSUPPORTED;
  }
  if (cfg->use_alpha_blending && use_bulk_staging()) {
    // packed per-overlap records into library scratch (after `prefix_bytes` the caller already uses), bulk-copy kernel
    void *records = nullptr;
    if (k > 0) {
      unsigned char *ws = (unsigned char *)stream_workspace(stream, prefix_bytes + (size_t)k * (F <= 3 ? 48 : 64));
      if (ws == nullptr) return GS_ERR_CUDA;
      records = ws + prefix_bytes;
      int rc = raster_pack_f32(digest, tile_ranges, overlap_to_point, k, width, height, F, records, nullptr, stream);
      if (rc != GS_OK) return rc;
    }
    return gs_raster_fwd_packed_f32(records, tile_ranges, overlap_to_point, v, k, width, height, F, cfg,
                                    median_threshold, image, image_alpha, visibility, median_image, stream);
  }
  RasterParams<float> P = make_params<float>(cfg, width, height, F);
  const int tiles = P.tiles_wide * ((height + kTile - 1) / kTile);
  const float median_lim = (float)(1.0 - median_threshold);
  const float4 *d = reinterpret_cast<const float4 *>(digest);
  switch (F) {
    case 1: return launch_fwd<1>(d, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
    case 2: return launch_fwd<2>(d, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
    case 3: return launch_fwd<3>(d, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
    default: return launch_fwd<4>(d, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
  }
}

// Reference-shaped entry (raw (V,7) points + features): digest into library scratch, then the tuned kernel;
// other configurations go to the generic kernel.
static int raster_fwd_f32_impl(const float *points, const float *features, const float *depths,
                               const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width,
                               int32_t height, int32_t F, const gs_raster_config *cfg, double median_threshold,
                               float *image, float *image_alpha, float *visibility, float *median_image,
                               cudaStream_t stream) {
  GS_CHECK_ARG(cfg != nullptr, "raster_fwd: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_fwd: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_visibility || visibility != nullptr || v == 0, "raster_fwd: compute_visibility needs a visibility buffer");
  const bool fast = fwd_tuned(cfg, F);
  if (median_image != nullptr && (!fast || !cfg->use_alpha_blending)) {
    set_error("raster_fwd_median: needs tile_size 16, no antialias, 1..4 features, alpha blending and depths");
    return GS_ERR_UNSUPPORTED;
  }
  if (fast) {
    void *digest = nullptr;
    const size_t digest_bytes = align_up((size_t)v * 64, 256);
    if (v > 0) {
      // digest first, the packed records follow it in the same scratch
      digest = stream_workspace(stream, digest_bytes + (size_t)k * (F <= 3 ? 48 : 64));
      if (digest == nullptr) return GS_ERR_CUDA;
      int rc = raster_digest_f32(points, features, depths, v, F, cfg->alpha_threshold, digest, stream);
      if (rc != GS_OK) return rc;
    }
    return raster_fwd_digest_impl(digest, tile_ranges, overlap_to_point, v, k, width, height, F, cfg, median_threshold,
                                  image, image_alpha, visibility, median_image, stream, digest_bytes);
  }
  return raster_fwd_generic<float>(points, features, tile_ranges, overlap_to_point, width, height, F, cfg, image,
                                   image_alpha, visibility, stream);
}

}  // namespace gs

extern "C" int gs_raster_fwd_f32(const float *points, const float *features, const int32_t *tile_ranges,
                                 const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width,
                                 int32_t height, int32_t F, const gs_raster_config *cfg, float *image,
                                 float *image_alpha, float *visibility, void *stream_) {
  return gs::raster_fwd_f32_impl(points, features, nullptr, tile_ranges, overlap_to_point, v, k, width, height, F, cfg, 0.0,
                                 image, image_alpha, visibility, nullptr, (cudaStream_t)stream_);
}

extern "C" int gs_raster_fwd_median_f32(const float *points, const float *features, const float *depths,
                                        const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t v,
                                        int64_t k, int32_t width, int32_t height, int32_t F,
                                        const gs_raster_config *cfg, double median_threshold, float *image,
                                        float *image_alpha, float *visibility, float *median_image, void *stream_) {
  GS_CHECK_ARG(median_image != nullptr && (depths != nullptr || v == 0), "raster_fwd_median: depths / median_image is NULL");
  return gs::raster_fwd_f32_impl(points, features, depths, tile_ranges, overlap_to_point, v, k, width, height, F, cfg,
                                 median_threshold, image, image_alpha, visibility, median_image, (cudaStream_t)stream_);
}

extern "C" int gs_raster_fwd_digest_f32(const void *digest, const int32_t *tile_ranges,
                                        const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width,
                                        int32_t height, int32_t F, const gs_raster_config *cfg,
                                        double median_threshold, float *image, float *image_alpha, float *visibility,
                                        float *median_image, void *stream_) {
  return gs::raster_fwd_digest_impl(digest, tile_ranges, overlap_to_point, v, k, width, height, F, cfg, median_threshold,
                                    image, image_alpha, visibility, median_image, (cudaStream_t)stream_);
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
