/*
 * gs_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's Taichi tile mapper and rasteriser
 * (uc-vision/taichi-splatting v0.32.0).  The reference ships no CPU version of
 * these stages, so this file follows the Taichi sources statement by statement:
 *
 *   tile mapper   taichi_lib/grid_query.py:9-93, mapper/tile_mapper.py:20-146,171-198
 *   raster fwd    rasterizer/forward.py:39-135, rasterizer/tiling.py:34-70,
 *                 taichi_lib/generic.py:306-357
 *   raster bwd    rasterizer/backward.py:73-225, taichi_lib/generic.py:320-404
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library.  The product path never does.
 *
 * Parity status: the reference has no golden vectors for these stages and its
 * Taichi runtime is not installable here, so the mapper/raster part of the
 * oracle is "parity unpinned" against reference *outputs*; it is pinned
 * (tests/test_oracle.py) against (1) finite differences (bwd == d fwd),
 * (2) the reference's own test property visibility == d(sum image)/d feature
 * (tests/test_visibility.py:34-64) and (3) brute-force geometric checks.
 *
 * Arithmetic rules that make the integer stages bit-reproducible on GPU:
 *   - compiled with -ffp-contract=off (no FMA contraction),
 *   - log() in the OBB query is the correctly rounded fp32 log, obtained as
 *     (float)log((double)x) on both CPU and GPU,
 *   - sqrtf and '/' are IEEE on both sides.
 *
 * Build: see oracle/Makefile (two shared objects: REAL=float and REAL=double).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#ifdef REAL_IS_DOUBLE
#define FN(name) CAT(name, _f64)
#define R_EXP exp
#define R_FABS fabs
#else
#define FN(name) CAT(name, _f32)
#define R_EXP expf
#define R_FABS fabsf
#endif

/* ------------------------------------------------------------------------- */
/* Tile mapper (fp32 only, like the reference: tile_mapper.py:14)             */
/* ------------------------------------------------------------------------- */
#ifndef REAL_IS_DOUBLE

typedef struct {
  float inv00, inv01, inv10, inv11; /* rows: axis1/scale.x, axis2/scale.y   */
  float relx, rely;                 /* min_tile*ts - mean                   */
  int minx, miny, spanx, spany;
} obb_query_t;

static inline float log_rn(float x) { return (float)log((double)x); }

/* grid_query.py:61-87 obb_grid_query + :9-27 tile_ranges */
static obb_query_t obb_grid_query(const float *g, int w_pad, int h_pad, int ts,
                                  float alpha_threshold) {
  obb_query_t q;
  memset(&q, 0, sizeof q);
  float mx = g[0], my = g[1], ax = g[2], ay = g[3], sx = g[4], sy = g[5],
        alpha = g[6];
  if (!(alpha > alpha_threshold)) return q; /* D18: NaN radius -> no tiles */

  float gaussian_scale = sqrtf(2.0f * log_rn(alpha / alpha_threshold));
  float scx = sx * gaussian_scale, scy = sy * gaussian_scale;
  float a2x = -ay, a2y = ax;
  /* ellipse_bounds(mean, axis1*scale.x, axis2*scale.y): generic.py:235-237 */
  float v1x = ax * scx, v1y = ay * scx, v2x = a2x * scy, v2y = a2y * scy;
  float ex = sqrtf(v1x * v1x + v2x * v2x), ey = sqrtf(v1y * v1y + v2y * v2y);
  float lox = mx - ex, loy = my - ey, hix = mx + ex, hiy = my + ey;

  q.inv00 = ax / scx; q.inv01 = ay / scx;
  q.inv10 = a2x / scy; q.inv11 = a2y / scy;

  float fts = (float)ts;
  int max_tx = (w_pad - 1) / ts, max_ty = (h_pad - 1) / ts;
  float flx = floorf(lox / fts), fly = floorf(loy / fts);
  float chx = ceilf(hix / fts), chy = ceilf(hiy / fts);
  /* guard the float->int casts (Taichi would wrap/UB on huge values) */
  const float BIG = 1.0e9f;
  if (!(flx > -BIG && flx < BIG && fly > -BIG && fly < BIG && chx > -BIG &&
        chx < BIG && chy > -BIG && chy < BIG))
    return q;
  int min_tx = (int)flx, min_ty = (int)fly;
  if (min_tx < 0) min_tx = 0;
  if (min_ty < 0) min_ty = 0;
  int max_bx = (int)chx, max_by = (int)chy;
  if (max_bx < min_tx + 1) max_bx = min_tx + 1;
  if (max_by < min_ty + 1) max_by = min_ty + 1;
  if (max_bx > max_tx + 1) max_bx = max_tx + 1;
  if (max_by > max_ty + 1) max_by = max_ty + 1;

  q.minx = min_tx; q.miny = min_ty;
  q.spanx = max_bx - min_tx; q.spany = max_by - min_ty;
  q.relx = (float)(min_tx * ts) - mx;
  q.rely = (float)(min_ty * ts) - my;
  return q;
}

/* grid_query.py:29-43 separates_bbox + :53-56 test_tile */
static inline int test_tile(const obb_query_t *q, int u, int v, int ts) {
  float lx = q->relx + (float)(u * ts), ly = q->rely + (float)(v * ts);
  float ux = lx + (float)ts, uy = ly + (float)ts;
  float px[4] = {lx, ux, ux, lx}, py[4] = {ly, ly, uy, uy};
  for (int i = 0; i < 2; ++i) {
    float r0 = i == 0 ? q->inv00 : q->inv10, r1 = i == 0 ? q->inv01 : q->inv11;
    float mn = INFINITY, mxv = -INFINITY;
    for (int j = 0; j < 4; ++j) {
      float a = r0 * px[j];
      float b = r1 * py[j];
      float l = a + b;
      if (l < mn) mn = l;
      if (l > mxv) mxv = l;
    }
    if (mn > 1.0f || mxv < -1.0f) return 0;
  }
  return 1;
}

/* tile_mapper.py:75-86 tile_overlaps_kernel */
void orc_tile_counts(const float *gaussians, int64_t V, int w_pad, int h_pad,
                     int ts, float alpha_threshold, int32_t *counts) {
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < V; ++i) {
    obb_query_t q = obb_grid_query(gaussians + 7 * i, w_pad, h_pad, ts,
                                   alpha_threshold);
    int c = 0;
    for (int u = 0; u < q.spanx; ++u)
      for (int v = 0; v < q.spany; ++v) c += test_tile(&q, u, v, ts);
    counts[i] = c;
  }
}

/* cuda_lib/full_cumsum.cu:16-47: exclusive scan into V+1 entries, returns total */
int64_t orc_full_cumsum(const int32_t *counts, int64_t V, int32_t *out) {
  int64_t acc = 0;
  for (int64_t i = 0; i < V; ++i) { out[i] = (int32_t)acc; acc += counts[i]; }
  out[V] = (int32_t)acc;
  return acc;
}

/* tile_mapper.py:35-66 make_sort_key, :114-146 generate_sort_keys_kernel */
void orc_tile_keys(const float *gaussians, const float *depths,
                   const int32_t *cum, int64_t V, int w_pad, int h_pad, int ts,
                   float alpha_threshold, int use_depth16, uint64_t *keys,
                   int32_t *overlap_to_point) {
  int tiles_wide = w_pad / ts;
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < V; ++i) {
    obb_query_t q = obb_grid_query(gaussians + 7 * i, w_pad, h_pad, ts,
                                   alpha_threshold);
    int64_t k = cum[i];
    float depth = depths[i];
    uint32_t dbits;
    if (use_depth16) {
      float c = depth < 0.0f ? 0.0f : (depth > 1.0f ? 1.0f : depth);
      dbits = (uint32_t)(c * 65535.0f);
    } else {
      memcpy(&dbits, &depth, 4);
    }
    for (int u = 0; u < q.spanx; ++u)
      for (int v = 0; v < q.spany; ++v)
        if (test_tile(&q, u, v, ts)) {
          uint64_t tile_id = (uint64_t)((q.minx + u) + (q.miny + v) * tiles_wide);
          keys[k] = use_depth16 ? ((tile_id << 16) | dbits)
                                : ((tile_id << 32) | dbits);
          overlap_to_point[k] = (int32_t)i;
          ++k;
        }
  }
}

/* cuda_lib/radix_sort_pairs.cu:7-29: stable LSD radix sort on bits [0,end_bit) */
void orc_sort_pairs(const uint64_t *keys_in, const int32_t *vals_in, int64_t K,
                    int end_bit, uint64_t *keys_out, int32_t *vals_out) {
  if (K == 0) return;
  uint64_t *ka = (uint64_t *)malloc(sizeof(uint64_t) * K);
  uint64_t *kb = (uint64_t *)malloc(sizeof(uint64_t) * K);
  int32_t *va = (int32_t *)malloc(sizeof(int32_t) * K);
  int32_t *vb = (int32_t *)malloc(sizeof(int32_t) * K);
  memcpy(ka, keys_in, sizeof(uint64_t) * K);
  memcpy(va, vals_in, sizeof(int32_t) * K);
  int64_t *hist = (int64_t *)malloc(sizeof(int64_t) * 65537);
  for (int bit = 0; bit < end_bit; bit += 16) {
    int nb = end_bit - bit < 16 ? end_bit - bit : 16;
    uint64_t mask = ((uint64_t)1 << nb) - 1;
    memset(hist, 0, sizeof(int64_t) * 65537);
    for (int64_t i = 0; i < K; ++i) hist[((ka[i] >> bit) & mask) + 1]++;
    for (int d = 0; d < 65536; ++d) hist[d + 1] += hist[d];
    for (int64_t i = 0; i < K; ++i) {
      int64_t p = hist[(ka[i] >> bit) & mask]++;
      kb[p] = ka[i]; vb[p] = va[i];
    }
    uint64_t *tk = ka; ka = kb; kb = tk;
    int32_t *tv = va; va = vb; vb = tv;
  }
  memcpy(keys_out, ka, sizeof(uint64_t) * K);
  memcpy(vals_out, va, sizeof(int32_t) * K);
  free(ka); free(kb); free(va); free(vb); free(hist);
}

/* tile_mapper.py:92-112 find_ranges_kernel; ranges must be zero-initialised */
void orc_tile_ranges(const uint64_t *sorted_keys, int64_t K, int use_depth16,
                     int32_t *ranges /* (T,2) zero-initialised */) {
  const int max_tile = 65535;
  int shift = use_depth16 ? 16 : 32;
  for (int64_t idx = 0; idx < K; ++idx) {
    int tile_id = (int)(sorted_keys[idx] >> shift);
    int next_tile_id = max_tile;
    if (idx + 1 < K) next_tile_id = (int)(sorted_keys[idx + 1] >> shift);
    if (tile_id != next_tile_id) {
      ranges[2 * tile_id + 1] = (int32_t)(idx + 1);
      if (next_tile_id < max_tile) ranges[2 * next_tile_id] = (int32_t)(idx + 1);
    }
  }
}
#endif /* !REAL_IS_DOUBLE */

/* ------------------------------------------------------------------------- */
/* Rasteriser                                                                 */
/* ------------------------------------------------------------------------- */

typedef struct {
  int tile_size;
  int pixel_stride_x, pixel_stride_y;
  int antialias;
  int use_alpha_blending;
  int compute_visibility;
  int compute_point_heuristic;
  int emulate_stale_group; /* D1: reproduce the reference's stale last group */
  double clamp_max_alpha;
  double alpha_threshold;
  double saturate_threshold;
} orc_raster_config_t;

/* generic.py:310-317 gaussian_pdf */
static inline REAL pdf_plain(REAL px, REAL py, const REAL *g) {
  REAL dx = px - g[0], dy = py - g[1];
  REAL tx = (dx * g[2] + dy * g[3]) / g[4];
  REAL ty = (dx * -g[3] + dy * g[2]) / g[5];
  return R_EXP((REAL)-0.5 * (tx * tx + ty * ty));
}

/* generic.py:320-336 gaussian_pdf_with_grad */
static inline REAL pdf_plain_grad(REAL px, REAL py, const REAL *g, REAL *dmean,
                                  REAL *daxis, REAL *dsigma) {
  REAL dx = px - g[0], dy = py - g[1];
  REAL ax = g[2], ay = g[3], sx = g[4], sy = g[5];
  REAL tx = (dx * ax + dy * ay) / sx;
  REAL ty = (dx * -ay + dy * ax) / sy;
  REAL tx2 = tx * tx, ty2 = ty * ty;
  REAL p = R_EXP((REAL)-0.5 * (tx2 + ty2));
  dsigma[0] = tx2 * p / sx; dsigma[1] = ty2 * p / sy;
  REAL tx_s = tx / sx, ty_s = ty / sy;
  /* perp(v) = (-v.y, v.x) */
  daxis[0] = p * (tx_s * -dx + ty_s * -dy);
  daxis[1] = p * (tx_s * -dy + ty_s * dx);
  dmean[0] = p * (tx_s * ax + ty_s * -ay);
  dmean[1] = p * (tx_s * ay + ty_s * ax);
  return p;
}

/* generic.py:340-345 S_sig */
static inline REAL s_sig(REAL x, REAL sigma) {
  REAL z = x / sigma;
  return (REAL)1 / ((REAL)1 + R_EXP((REAL)-1.6 * z - (REAL)0.07 * z * z * z));
}

/* generic.py:347-357 gaussian_pdf_antialias */
static inline REAL pdf_aa(REAL px, REAL py, const REAL *g) {
  REAL dx = px - g[0], dy = py - g[1];
  REAL sx = g[4], sy = g[5];
  REAL tx = dx * g[2] + dy * g[3];
  REAL ty = dx * -g[3] + dy * g[2];
  REAL Sx1 = s_sig(tx + (REAL)0.5, sx), Sx2 = s_sig(tx - (REAL)0.5, sx);
  REAL Sy1 = s_sig(ty + (REAL)0.5, sy), Sy2 = s_sig(ty - (REAL)0.5, sy);
  return (REAL)2 * (REAL)M_PI * sx * (Sx1 - Sx2) * sy * (Sy1 - Sy2);
}

/* generic.py:359-369 S_sig_grad */
static inline void s_sig_grad(REAL x, REAL sigma, REAL *s, REAL *ds_dx,
                              REAL *ds_dsig) {
  REAL z = x / sigma;
  REAL sv = (REAL)1 / ((REAL)1 + R_EXP((REAL)-1.6 * z - (REAL)0.07 * z * z * z));
  REAL d = ((REAL)1.6 + (REAL)0.21 * z * z) * sv * ((REAL)1 - sv);
  REAL dx = d / sigma;
  *s = sv; *ds_dx = dx; *ds_dsig = dx * -z;
}

/* generic.py:371-404 gaussian_pdf_antialias_with_grad */
static inline REAL pdf_aa_grad(REAL px, REAL py, const REAL *g, REAL *dmean,
                               REAL *daxis, REAL *dsigma) {
  REAL dx = px - g[0], dy = py - g[1];
  REAL ax = g[2], ay = g[3], sx = g[4], sy = g[5];
  REAL tx = dx * ax + dy * ay;
  REAL ty = dx * -ay + dy * ax;
  REAL Sx1, dSx1, dSx1s, Sx2, dSx2, dSx2s, Sy1, dSy1, dSy1s, Sy2, dSy2, dSy2s;
  s_sig_grad(tx + (REAL)0.5, sx, &Sx1, &dSx1, &dSx1s);
  s_sig_grad(tx - (REAL)0.5, sx, &Sx2, &dSx2, &dSx2s);
  s_sig_grad(ty + (REAL)0.5, sy, &Sy1, &dSy1, &dSy1s);
  s_sig_grad(ty - (REAL)0.5, sy, &Sy2, &dSy2, &dSy2s);
  REAL ix = sx * (Sx1 - Sx2), iy = sy * (Sy1 - Sy2);
  REAL tau = (REAL)2 * (REAL)M_PI;
  REAL i2d = tau * ix * iy;
  REAL dSx = iy * sx * (dSx1 - dSx2);
  REAL dSy = ix * sy * (dSy1 - dSy2);
  /* di_dmean = tau * (dSx * -axis + dSy * -perp(axis)) */
  dmean[0] = tau * (dSx * -ax + dSy * ay);
  dmean[1] = tau * (dSx * -ay + dSy * -ax);
  dsigma[0] = tau * iy * (Sx1 - Sx2 + (dSx1s - dSx2s) * sx);
  dsigma[1] = tau * ix * (Sy1 - Sy2 + (dSy1s - dSy2s) * sy);
  /* di_daxis = tau * (dSx * d + dSy * -perp(d)) ; perp(d) = (-dy, dx) */
  daxis[0] = tau * (dSx * dx + dSy * dy);
  daxis[1] = tau * (dSx * dy + dSy * -dx);
  return i2d;
}

/*
 * Builds the sequence of (splat slot, fresh?) a tile processes.
 * Intended semantics: every overlap once, all fresh.
 * emulate_stale_group (D1, forward.py:86-89 / backward.py:138-141): groups of
 * `group` slots; the inner loop bound is min(group, count - group_id), so the
 * last group re-processes stale slots of the previous group.
 */
static int64_t build_sequence(int start, int end, int group, int emulate,
                              int32_t **seq_out, uint8_t **fresh_out) {
  int count = end - start;
  if (count <= 0) { *seq_out = NULL; *fresh_out = NULL; return 0; }
  int ngroups = (count + group - 1) / group;
  int64_t cap = (int64_t)ngroups * group;
  int32_t *seq = (int32_t *)malloc(sizeof(int32_t) * cap);
  uint8_t *fresh = (uint8_t *)malloc(cap);
  int64_t n = 0;
  if (!emulate) {
    for (int i = 0; i < count; ++i) { seq[n] = start + i; fresh[n] = 1; ++n; }
  } else {
    int32_t *slots = (int32_t *)malloc(sizeof(int32_t) * group);
    for (int gid = 0; gid < ngroups; ++gid) {
      int gstart = start + gid * group;
      int nfresh = end - gstart < group ? end - gstart : group;
      for (int s = 0; s < nfresh; ++s) slots[s] = gstart + s;
      int remaining = count - gid;
      int bound = remaining < group ? remaining : group;
      for (int s = 0; s < bound; ++s) {
        seq[n] = slots[s]; fresh[n] = (uint8_t)(s < nfresh); ++n;
      }
    }
    free(slots);
  }
  *seq_out = seq; *fresh_out = fresh;
  return n;
}

/* rasterizer/forward.py:39-135 */
void FN(orc_raster_fwd)(const REAL *points /* (V,7) */,
                        const REAL *features /* (V,F) */,
                        const int32_t *ranges /* (T,2) */,
                        const int32_t *overlap_to_point /* (K) */, int W, int H,
                        int F, const orc_raster_config_t *cfg,
                        REAL *image /* (H,W,F) */, REAL *image_alpha /* (H,W) */,
                        REAL *visibility /* (V) pre-zeroed, or NULL */) {
  int ts = cfg->tile_size;
  int tiles_wide = (W + ts - 1) / ts, tiles_high = (H + ts - 1) / ts;
  REAL clamp_max = (REAL)cfg->clamp_max_alpha, thr = (REAL)cfg->alpha_threshold;
  REAL sat_lim = (REAL)(1.0 - cfg->saturate_threshold);
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < tiles_wide * tiles_high; ++tile) {
    int x0 = (tile % tiles_wide) * ts, y0 = (tile / tiles_wide) * ts;
    int start = ranges[2 * tile], end = ranges[2 * tile + 1];
    int32_t *seq; uint8_t *fresh;
    int64_t n = build_sequence(start, end, ts * ts, cfg->emulate_stale_group,
                               &seq, &fresh);
    REAL *vis_local = NULL;
    if (cfg->compute_visibility && visibility && n > 0)
      vis_local = (REAL *)calloc((size_t)(end - start), sizeof(REAL));
    REAL accum[16];
    for (int py = y0; py < y0 + ts; ++py)
      for (int px = x0; px < x0 + ts; ++px) {
        if (py >= H || px >= W) continue; /* out-of-image threads write nothing */
        REAL fx = (REAL)px + (REAL)0.5, fy = (REAL)py + (REAL)0.5;
        for (int c = 0; c < F; ++c) accum[c] = 0;
        REAL total_weight = 0;
        int saturated = 0;
        for (int64_t s = 0; s < n; ++s) {
          if (saturated) break; /* only reachable in non-blending mode */
          int32_t id = overlap_to_point[seq[s]];
          const REAL *g = points + 7 * (int64_t)id;
          REAL ga = cfg->antialias ? pdf_aa(fx, fy, g) : pdf_plain(fx, fy, g);
          REAL alpha = g[6] * ga;
          if (alpha > clamp_max) alpha = clamp_max;
          if (alpha > thr) {
            REAL weight = alpha * ((REAL)1 - total_weight);
            total_weight += weight;
            const REAL *f = features + (int64_t)F * id;
            if (cfg->use_alpha_blending) {
              for (int c = 0; c < F; ++c) accum[c] += f[c] * weight;
            } else {
              if (total_weight >= sat_lim && !saturated)
                for (int c = 0; c < F; ++c) accum[c] = f[c];
              saturated = total_weight >= sat_lim;
            }
            if (vis_local && fresh[s]) vis_local[seq[s] - start] += weight;
          }
        }
        REAL *out = image + ((int64_t)py * W + px) * F;
        for (int c = 0; c < F; ++c) out[c] = accum[c];
        image_alpha[(int64_t)py * W + px] =
            cfg->use_alpha_blending ? total_weight
                                    : (REAL)(total_weight > 0 ? 1 : 0);
      }
    if (vis_local) {
      for (int i = 0; i < end - start; ++i) {
        REAL v = vis_local[i];
        if (v != 0) {
#pragma omp atomic
          visibility[overlap_to_point[start + i]] += v;
        }
      }
      free(vis_local);
    }
    free(seq); free(fresh);
  }
}

/* rasterizer/backward.py:73-225 (blend mode; use_alpha_blending is ignored by
 * the reference backward, D3) */
void FN(orc_raster_bwd)(const REAL *points, const REAL *features,
                        const int32_t *ranges, const int32_t *overlap_to_point,
                        const REAL *image /* (H,W,F) forward output */,
                        const REAL *grad_image /* (H,W,F) */, int W, int H,
                        int F, const orc_raster_config_t *cfg,
                        REAL *grad_points /* (V,7) pre-zeroed or NULL */,
                        REAL *grad_features /* (V,F) pre-zeroed or NULL */,
                        REAL *point_heuristic /* (V,2) accumulates, or NULL */) {
  int ts = cfg->tile_size;
  int tiles_wide = (W + ts - 1) / ts, tiles_high = (H + ts - 1) / ts;
  int group = (ts * ts) / (cfg->pixel_stride_x * cfg->pixel_stride_y);
  REAL clamp_max = (REAL)cfg->clamp_max_alpha, thr = (REAL)cfg->alpha_threshold;
  REAL sat = (REAL)cfg->saturate_threshold;
  int do_heur = cfg->compute_point_heuristic && point_heuristic;
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < tiles_wide * tiles_high; ++tile) {
    int x0 = (tile % tiles_wide) * ts, y0 = (tile / tiles_wide) * ts;
    int start = ranges[2 * tile], end = ranges[2 * tile + 1];
    int32_t *seq; uint8_t *fresh;
    int64_t n = build_sequence(start, end, group, cfg->emulate_stale_group,
                               &seq, &fresh);
    if (n == 0) continue;
    int cnt = end - start;
    int stride = 7 + F + 2;
    double *acc = (double *)calloc((size_t)cnt * stride, sizeof(double));
    REAL remaining[16], gpix[16];
    for (int py = y0; py < y0 + ts; ++py)
      for (int px = x0; px < x0 + ts; ++px) {
        if (py >= H || px >= W) continue; /* total_weight = 1 => saturated */
        REAL fx = (REAL)px + (REAL)0.5, fy = (REAL)py + (REAL)0.5;
        const REAL *img = image + ((int64_t)py * W + px) * F;
        const REAL *gi = grad_image + ((int64_t)py * W + px) * F;
        for (int c = 0; c < F; ++c) { remaining[c] = img[c]; gpix[c] = gi[c]; }
        REAL total_weight = 0;
        for (int64_t s = 0; s < n; ++s) {
          if (total_weight >= sat) break;
          int32_t id = overlap_to_point[seq[s]];
          const REAL *g = points + 7 * (int64_t)id;
          REAL dmean[2], daxis[2], dsigma[2];
          REAL ga = cfg->antialias
                        ? pdf_aa_grad(fx, fy, g, dmean, daxis, dsigma)
                        : pdf_plain_grad(fx, fy, g, dmean, daxis, dsigma);
          REAL point_alpha = g[6];
          REAL alpha = point_alpha * ga;
          if (!(alpha > thr)) continue;
          if (alpha > clamp_max) alpha = clamp_max;
          const REAL *f = features + (int64_t)F * id;
          REAL T_i = (REAL)1 - total_weight;
          REAL weight = alpha * T_i;
          total_weight += weight;
          REAL alpha_grad = 0;
          for (int c = 0; c < F; ++c) {
            remaining[c] -= f[c] * weight;
            REAL diff = f[c] * T_i - remaining[c] / ((REAL)1 - alpha);
            alpha_grad += diff * gpix[c];
          }
          if (!fresh[s]) continue; /* stale duplicates: grads are dropped */
          REAL aag = point_alpha * alpha_grad;
          double *a = acc + (size_t)(seq[s] - start) * stride;
          a[0] += aag * dmean[0]; a[1] += aag * dmean[1];
          a[2] += aag * daxis[0]; a[3] += aag * daxis[1];
          a[4] += aag * dsigma[0]; a[5] += aag * dsigma[1];
          a[6] += ga * alpha_grad;
          for (int c = 0; c < F; ++c) a[7 + c] += weight * gpix[c];
          a[7 + F] += aag * aag;
          a[8 + F] += R_FABS(aag * dmean[0]) + R_FABS(aag * dmean[1]);
        }
      }
    for (int i = 0; i < cnt; ++i) {
      const double *a = acc + (size_t)i * stride;
      int64_t id = overlap_to_point[start + i];
      if (grad_points)
        for (int c = 0; c < 7; ++c) {
#pragma omp atomic
          grad_points[7 * id + c] += (REAL)a[c];
        }
      if (grad_features)
        for (int c = 0; c < F; ++c) {
#pragma omp atomic
          grad_features[F * id + c] += (REAL)a[7 + c];
        }
      if (do_heur)
        for (int c = 0; c < 2; ++c) {
#pragma omp atomic
          point_heuristic[2 * id + c] += (REAL)a[7 + F + c];
        }
    }
    free(acc); free(seq); free(fresh);
  }
}
