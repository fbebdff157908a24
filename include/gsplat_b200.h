/*
 * gsplat_b200.h -- C ABI of libgsplat_b200.so: the B200-native (sm_100a) replacement for the
 * Taichi/CUB kernels on taichi-splatting's render hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces one device-code unit
 * the reference's Python operators launch; the reference-side binding is in INTEGRATION.md.
 * Paths below are relative to the reference tree (uc-vision/taichi-splatting v0.32.0).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host` (pinned host memory);
 *  - tensors are dense row-major with the reference's layouts; the library never allocates or
 *    frees caller-visible memory: outputs and workspaces are caller-owned (CUB-style size query);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *    the device; the only host-visible results (V, K) are written asynchronously into pinned
 *    host words the caller reads after synchronising that stream (the whole-frame drivers
 *    gs_render_* return V and K themselves and wait for them internally);
 *  - lifetime: every buffer passed to a call must stay allocated until the work enqueued by that
 *    call has run -- the calls return before the device has touched anything.  A binding must not
 *    take the address of a temporary (in Python: `ptr(x.contiguous())` frees the copy as soon as
 *    `ptr` returns); buffers may be released in stream order (a stream-ordered allocator such as
 *    torch's caching allocator on that stream is sufficient);
 *  - every function returns 0 on success or a negative GS_ERR_* code; gs_last_error_string()
 *    gives the thread-local message (reference convention: Python assert / RuntimeError,
 *    mapper/tile_mapper.py:177-178, cuda_lib/radix_sort_pairs.cu:66);
 *  - `_f32` / `_f64`: the reference instantiates every operator for both (taichi_lib/__init__.py:8-14);
 *    the tile mapper is f32 only (mapper/tile_mapper.py:14).
 */
#ifndef GSPLAT_B200_H
#define GSPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_OK 0
#define GS_ERR_INVALID_ARGUMENT (-1)
#define GS_ERR_UNSUPPORTED (-2)
#define GS_ERR_CUDA (-3)
#define GS_ERR_WORKSPACE_TOO_SMALL (-4)

/* Mirrors RasterConfig (data_types.py:16-46): compile-time constants of the Taichi kernels. */
typedef struct gs_raster_config {
  int32_t tile_size;               /* 8, 16 or 32 (fast path: 16) */
  int32_t pixel_stride_x;          /* accepted for API parity; thread mapping is the library's own */
  int32_t pixel_stride_y;
  int32_t antialias;               /* generic.py:340-404 */
  int32_t use_alpha_blending;      /* 0: quantile (median-depth) mode, forward.py:107-112 */
  int32_t compute_visibility;      /* forward.py:114-126 */
  int32_t compute_point_heuristic; /* backward.py:190-194 */
  int32_t reserved;
  double clamp_max_alpha;          /* 0.99 */
  double alpha_threshold;          /* 1/255 */
  double saturate_threshold;       /* 0.9999 (backward stop; forward quantile threshold) */
  double forward_saturate_eps;     /* forward early-out when 1-total_weight <= eps; 0 = never
                                      (the reference forward never stops early: SURVEY D2) */
} gs_raster_config;

int gs_version(void);
const char *gs_last_error_string(void);

/* ---- R1: projection + cull + order-preserving compaction --------------------------------------
 * replaces project_kernel + torch.nonzero + 2 gathers (perspective/projection.py:32-81,125-163).
 * Two calls around the one unavoidable host read of V (the reference syncs in torch.nonzero):
 *   gs_project_cull_*   : per-Gaussian in-view flag -> exclusive scan -> *num_visible_host
 *   gs_project_write_*  : recompute and write compacted points (V,7), depth (V,1), indexes (V) i64,
 *                         and optionally ndc depth (V,1) (torch_lib/projection.py:120-123, R11)
 * T_camera_world: 16 values row-major (4,4); projection: [fx,fy,cx,cy].                        */
int gs_project_workspace_bytes(int64_t n, size_t *bytes);
int gs_project_cull_f32(const float *position, const float *log_scaling, const float *rotation,
                        const float *alpha_logit, const float *T_camera_world, const float *projection,
                        int64_t n, int32_t width, int32_t height, double near_plane, double far_plane,
                        double blur_cov, double clamp_margin, double alpha_threshold,
                        void *workspace, size_t workspace_bytes, int32_t *num_visible_host, void *stream);
/* Single-pass alternative to gs_project_cull_* + gs_project_write_*: project, cull and compact IN ORDER in one kernel
 * (the projection is the load of a cub::DeviceSelect stream compaction), each Gaussian projected once.  Outputs need
 * capacity n rows (V rows are written; V arrives in *num_visible_host after the stream is synchronised); ndc_depth may
 * be NULL.  Same results as the two-kernel form.  Workspace: gs_project_workspace_bytes(n) is enough. */
int gs_project_compact_workspace_bytes(int64_t n, int32_t fp64, size_t *bytes);
int gs_project_compact_f32(const float *position, const float *log_scaling, const float *rotation,
                           const float *alpha_logit, const float *T_camera_world, const float *projection, int64_t n,
                           int32_t width, int32_t height, double near_plane, double far_plane, double blur_cov,
                           double clamp_margin, double alpha_threshold, void *workspace, size_t workspace_bytes,
                           float *points, float *depth, int64_t *indexes, float *ndc_depth, int32_t *num_visible_host,
                           void *stream);
int gs_project_compact_f64(const double *position, const double *log_scaling, const double *rotation,
                           const double *alpha_logit, const double *T_camera_world, const double *projection, int64_t n,
                           int32_t width, int32_t height, double near_plane, double far_plane, double blur_cov,
                           double clamp_margin, double alpha_threshold, void *workspace, size_t workspace_bytes,
                           double *points, double *depth, int64_t *indexes, double *ndc_depth,
                           int32_t *num_visible_host, void *stream);
int gs_project_write_f32(const float *position, const float *log_scaling, const float *rotation,
                         const float *alpha_logit, const float *T_camera_world, const float *projection,
                         int64_t n, int32_t width, int32_t height, double near_plane, double far_plane,
                         double blur_cov, double clamp_margin, const void *workspace,
                         float *points, float *depth, int64_t *indexes, float *ndc_depth /* may be NULL */,
                         void *stream);
int gs_project_cull_f64(const double *position, const double *log_scaling, const double *rotation,
                        const double *alpha_logit, const double *T_camera_world, const double *projection,
                        int64_t n, int32_t width, int32_t height, double near_plane, double far_plane,
                        double blur_cov, double clamp_margin, double alpha_threshold,
                        void *workspace, size_t workspace_bytes, int32_t *num_visible_host, void *stream);
int gs_project_write_f64(const double *position, const double *log_scaling, const double *rotation,
                         const double *alpha_logit, const double *T_camera_world, const double *projection,
                         int64_t n, int32_t width, int32_t height, double near_plane, double far_plane,
                         double blur_cov, double clamp_margin, const void *workspace,
                         double *points, double *depth, int64_t *indexes, double *ndc_depth,
                         void *stream);

/* camera position in world space, -A^-1 t for T_camera_world = [A t; 0 0 0 1]: replaces
 * torch.inverse(T_camera_world)[0:3, 3] (perspective/params.py:78-80) on the SH path (renderer.py:53). */
int gs_camera_position_f32(const float *T_camera_world, float *camera_pos, void *stream);
int gs_camera_position_f64(const double *T_camera_world, double *camera_pos, void *stream);

/* ---- R1b: projection backward ------------------------------------------------------------------
 * replaces indexed_project_kernel.grad (Taichi autodiff; perspective/projection.py:84-119,165-188).
 * Hand-derived reverse chain.  All grad outputs must be zero-initialised by
 * the caller; d_T_camera_world is (4,4) (rows 0-2 written), d_projection is (4).               */
int gs_project_bwd_f32(const float *position, const float *log_scaling, const float *rotation,
                       const float *alpha_logit, const float *T_camera_world, const float *projection,
                       const int64_t *indexes, int64_t v, int32_t width, int32_t height,
                       double blur_cov, double clamp_margin, const float *d_points, const float *d_depth,
                       float *d_position, float *d_log_scaling, float *d_rotation, float *d_alpha_logit,
                       float *d_T_camera_world, float *d_projection, void *stream);
int gs_project_bwd_f64(const double *position, const double *log_scaling, const double *rotation,
                       const double *alpha_logit, const double *T_camera_world, const double *projection,
                       const int64_t *indexes, int64_t v, int32_t width, int32_t height,
                       double blur_cov, double clamp_margin, const double *d_points, const double *d_depth,
                       double *d_position, double *d_log_scaling, double *d_rotation, double *d_alpha_logit,
                       double *d_T_camera_world, double *d_projection, void *stream);

/* ---- R2: spherical harmonics at gathered indexes ----------------------------------------------
 * replaces evaluate_sh_at_kernel (+ .grad) (indexed_spherical_harmonics.py:118-134,152-160).
 * params (M,C,D), D=(degree+1)^2, degree 0..3; out (V,C) = clamp(Y(dir).params + 0.5, 0, 1).
 * Backward accumulates (atomic) into zero-initialised d_params (M,C,D), d_positions (M,3),
 * d_camera_pos (3); any of the three may be NULL.  unique_indexes!=0 promises no index repeats
 * (true for the renderer's visible set) and enables plain vector stores into d_params.  `out` (the
 * forward result, may be NULL) lets the backward take the clamp mask from it instead of re-reading
 * the coefficient rows.                                                                          */
int gs_sh_fwd_f32(const float *params, const float *positions, const int64_t *indexes,
                  const float *camera_pos, int64_t v, int32_t channels, int32_t degree, float *out,
                  void *stream);
int gs_sh_bwd_f32(const float *params, const float *positions, const int64_t *indexes,
                  const float *camera_pos, const float *d_out, const float *out /* forward output or NULL */,
                  int64_t v, int32_t channels, int32_t degree,
                  int32_t unique_indexes, float *d_params, float *d_positions, float *d_camera_pos,
                  void *stream);
int gs_sh_fwd_f64(const double *params, const double *positions, const int64_t *indexes,
                  const double *camera_pos, int64_t v, int32_t channels, int32_t degree, double *out,
                  void *stream);
int gs_sh_bwd_f64(const double *params, const double *positions, const int64_t *indexes,
                  const double *camera_pos, const double *d_out, const double *out, int64_t v, int32_t channels,
                  int32_t degree,
                  int32_t unique_indexes, double *d_params, double *d_positions, double *d_camera_pos,
                  void *stream);

/* View-parallel exchange of the SH gradient (multi-GPU, no reference counterpart).  The per-view SH coefficient
 * gradient is rank-1, d_params[i,c,:] = Y(dir_view(i)) * g_view[i,c], so W views exchange only their dense
 * (N, C) colour gradients g_w (all-gathered into g_all, `view_stride` elements apart, zero where culled or
 * clamped) and camera centres (W,3); this kernel rebuilds sum_w Y_w(i) * g_w[i,c] into d_params (N,C,D).      */
int gs_sh_bwd_views_f32(const float *positions, const float *cam_positions, const float *g_all, int64_t n,
                        int32_t views, int32_t channels, int64_t view_stride, int32_t degree, float *d_params,
                        void *stream);
/* This rank's contribution to that exchange in one launch: out (n * channels + 3) = dense (n, channels) colour
 * gradients, zero where the colour is clamped (colour <= 0 or >= 1) or the Gaussian was culled, followed by the camera
 * centre.  d_colours / colours (v, channels), indexes (v). */
int gs_sh_pack_factors_f32(const float *colours, const float *d_colours, const int64_t *indexes,
                           const float *camera_pos, int64_t v, int32_t channels, int64_t n, float *out, void *stream);
/* The same pack FUSED with the all-gather, over peer memory: every value is stored straight into this rank's slot of
 * the gathered buffer of every rank.  peer_bases_host[w] = device pointer (valid in THIS process: symmetric-memory /
 * CUDA IPC mapping) of rank w's gathered buffer; values land at element slot_offset + (the gs_sh_pack_factors_f32
 * layout), so the caller passes slot_offset = buffer_slot * world * view_stride + rank * view_stride.  The stores to
 * remote buffers travel over NVLink while the kernel runs; the caller orders them before the readers with a
 * symmetric-memory barrier on the same stream. */
/* In-place all-reduce (sum) of `count` floats (multiple of 4) over peer memory: peer_bases_host[w] = rank w's copy of the
 * buffer as mapped in this process (16-byte aligned).  Rank r sums its 1/world slice from every copy and stores the sum
 * into every copy; the caller puts a symmetric-memory barrier before (all copies written) and after (all sums landed). */
int gs_allreduce_peers_f32(const uint64_t *peer_bases_host, int32_t world, int32_t rank, int64_t count, void *stream);
int gs_sh_pack_factors_peers_f32(const float *colours, const float *d_colours, const int64_t *indexes,
                                 const float *camera_pos, int64_t v, int32_t channels, int64_t n,
                                 const uint64_t *peer_bases_host, int32_t world, int64_t slot_offset, void *stream);

/* ---- R3-R7: tile mapper -------------------------------------------------------------------------
 * gs_tile_count      replaces tile_overlaps_kernel (mapper/tile_mapper.py:75-86, grid_query.py:46-93)
 * gs_tile_scan       replaces cuda_lib.full_cumsum (cuda_lib/full_cumsum.cu:16-47): cum (V+1) i32,
 *                    total K written asynchronously to *total_host (no device-wide sync, D11)
 * gs_tile_emit_keys  replaces generate_sort_keys_kernel (mapper/tile_mapper.py:35-66,114-146)
 * gs_sort_pairs      replaces cuda_lib.radix_sort_pairs (cuda_lib/radix_sort_pairs.cu:7-70):
 *                    stable LSD radix sort of (key, i32 value) on bits [begin_bit,end_bit)
 * gs_tile_ranges     replaces find_ranges_kernel (mapper/tile_mapper.py:92-112); zero-fills ranges
 * width_padded/height_padded: image size rounded up to the tile size (pad_to_tile, :20-24).    */
int gs_tile_count(const float *gaussians, int64_t v, int32_t width_padded, int32_t height_padded,
                  int32_t tile_size, double alpha_threshold, int32_t *counts, void *stream);
int gs_tile_scan_workspace_bytes(int64_t v, size_t *bytes);
int gs_tile_scan(const int32_t *counts, int64_t v, int32_t *cum /* (v+1) */, void *workspace,
                 size_t workspace_bytes, int32_t *total_host, void *stream);
int gs_tile_emit_keys(const float *gaussians, const float *depths, const int32_t *cum, int64_t v,
                      int32_t width_padded, int32_t height_padded, int32_t tile_size,
                      double alpha_threshold, int32_t use_depth16, void *keys /* u64 | u32 */,
                      int32_t *overlap_to_point, void *stream);
int gs_sort_pairs_workspace_bytes(int64_t k, int32_t key_bytes, size_t *bytes);
int gs_sort_pairs(const void *keys_in, const int32_t *values_in, void *keys_out, int32_t *values_out,
                  int64_t k, int32_t key_bytes /* 4 | 8 */, int32_t begin_bit, int32_t end_bit,
                  void *workspace, size_t workspace_bytes, void *stream);
int gs_tile_ranges(const void *sorted_keys, int64_t k, int32_t key_bytes, int32_t *tile_ranges /* (T,2) */,
                   int64_t num_tiles, void *stream);

/* Two-level ordering: identical final order to gs_tile_emit_keys + gs_sort_pairs(48 bits), with the depth passes
 * run on the V Gaussians instead of the K overlaps (an LSD sort over tile|depth == stable sort by depth, then
 * stable sort by tile; every overlap of a Gaussian shares its depth):
 *   gs_depth_order           order (V) i32 = Gaussian indexes sorted by (depth bits, index)
 *   gs_tile_count_ordered    counts[r] for Gaussian order[r]
 *   gs_tile_emit_ordered     tile_keys (K) u32 = tile id, overlap_to_point (K), emitted in depth order
 *   gs_sort_pairs(tile_keys, 4 bytes, bits [0, ceil(log2 T)))  -- stable
 *   gs_tile_ranges_from_tiles                                                                          */
int gs_depth_order_workspace_bytes(int64_t v, size_t *bytes);
int gs_depth_order(const float *depths, int64_t v, int32_t use_depth16, int32_t *order, void *workspace,
                   size_t workspace_bytes, void *stream);
int gs_tile_count_ordered(const float *gaussians, const int32_t *order, int64_t v, int32_t width_padded,
                          int32_t height_padded, int32_t tile_size, double alpha_threshold, int32_t *counts,
                          void *stream);
int gs_tile_emit_ordered(const float *gaussians, const int32_t *order, const int32_t *cum, int64_t v,
                         int32_t width_padded, int32_t height_padded, int32_t tile_size, double alpha_threshold,
                         uint32_t *tile_keys, int32_t *overlap_to_point, void *stream);
/* Same two kernels sharing ONE grid query: the count kernel also leaves a 16-byte hit record per Gaussian (tile span
 * as 4 x u16 + a 64-bit mask of the accepted tiles in enumeration order; spans over 64 tiles are marked and
 * re-queried), and the emit kernel walks the set bits instead of repeating the query.  hits: (v) x 16 bytes.
 * tile_lo / tile_hi: a tile-sharded multi-GPU rank keeps only the overlaps of its own contiguous tile-id range, so its
 * sort, pack and raster kernels see K / world overlaps.  gs_tile_count_ordered_hits with order == NULL counts
 * Gaussian r itself (counts / hits at the Gaussians' own indices); the whole-frame driver runs that form beside the
 * depth sort and scans / emits through the order afterwards. */
int gs_tile_count_ordered_hits(const float *gaussians, const int32_t *order, int64_t v, int32_t width_padded,
                               int32_t height_padded, int32_t tile_size, double alpha_threshold,
                               int32_t tile_lo, int32_t tile_hi /* keep tiles [lo, hi) only; (0, 0): all */,
                               int32_t *counts, void *hits, void *stream);
int gs_tile_emit_hits(const float *gaussians, const int32_t *order, const int32_t *cum, const void *hits, int64_t v,
                      int32_t width_padded, int32_t height_padded, int32_t tile_size, double alpha_threshold,
                      int32_t tile_lo, int32_t tile_hi, uint32_t *tile_keys, int32_t *overlap_to_point, void *stream);
int gs_tile_ranges_from_tiles(const uint32_t *sorted_tiles, int64_t k, int32_t *tile_ranges, int64_t num_tiles,
                              void *stream);

/* Binned ordering: identical final order once more, without any global sort.  Overlaps are counted per TILE
 * (atomics), the counts become the tile ranges, every overlap takes a slot inside its tile's segment in arrival
 * order, and each segment is sorted on (depth bits << 32 | Gaussian index) in shared memory (keys are distinct, so
 * the result is independent of the arrival order).  Segments longer than gs_tile_bin_max_per_tile() are not
 * supported (GS_ERR_UNSUPPORTED): the caller reads the largest population with K and falls back to the two-level
 * ordering.
 *   gs_tile_bin_count    tile_counts (T) i32 (zeroed here)
 *   gs_tile_bin_offsets  tile_ranges (T,2), cursor (T) = segment starts, totals {K, max per tile} -> device + pinned host
 *   gs_tile_bin_emit     keys (K) u64 in segment slots (advances cursor)
 *   gs_tile_bin_sort     overlap_to_point (K) i32                                                            */
int gs_tile_bin_count(const float *gaussians, int64_t v, int32_t width_padded, int32_t height_padded,
                      int32_t tile_size, double alpha_threshold, int32_t *tile_counts, void *stream);
int gs_tile_bin_offsets(const int32_t *tile_counts, int64_t num_tiles, int32_t *tile_ranges, int32_t *cursor,
                        int32_t *totals_dev /* 2 */, int32_t *totals_host /* 2, pinned */, void *stream);
int gs_tile_bin_emit(const float *gaussians, const float *depths, int64_t v, int32_t width_padded,
                     int32_t height_padded, int32_t tile_size, double alpha_threshold, int32_t use_depth16,
                     int32_t *cursor, uint64_t *keys, void *stream);
int gs_tile_bin_max_per_tile(void);
int gs_tile_bin_sort(const uint64_t *keys, const int32_t *tile_ranges, int64_t num_tiles, int32_t max_per_tile,
                     int32_t *overlap_to_point, void *stream);

/* ---- R8: rasteriser forward ---------------------------------------------------------------------
 * replaces _forward_kernel (rasterizer/forward.py:22-135).  image (H,W,F), image_alpha (H,W),
 * visibility (V) zero-initialised by the caller (NULL unless compute_visibility).            */
int gs_raster_fwd_f32(const float *points, const float *features, const int32_t *tile_ranges,
                      const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width, int32_t height,
                      int32_t num_features, const gs_raster_config *config, float *image,
                      float *image_alpha, float *visibility, void *stream);
/* Fused variant: also writes median_image (H,W): depth of the first splat at which the accumulated weight
 * reaches 1 - median_threshold, i.e. the output of the reference's second, non-blending raster pass
 * over features=depths (renderer.py:77-82, forward.py:107-112) without re-walking the tile lists.
 * fp32, tile_size 16, no antialias, 1..4 features only (GS_ERR_UNSUPPORTED otherwise).          */
int gs_raster_fwd_median_f32(const float *points, const float *features, const float *depths,
                             const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t v, int64_t k,
                             int32_t width, int32_t height, int32_t num_features, const gs_raster_config *config,
                             double median_threshold, float *image, float *image_alpha, float *visibility,
                             float *median_image, void *stream);
int gs_raster_fwd_f64(const double *points, const double *features, const int32_t *tile_ranges,
                      const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width, int32_t height,
                      int32_t num_features, const gs_raster_config *config, double *image,
                      double *image_alpha, double *visibility, void *stream);

/* ---- R9: rasteriser backward --------------------------------------------------------------------
 * replaces _backward_kernel (rasterizer/backward.py:50-225).  grad_points (V,7) / grad_features
 * (V,F) zero-initialised by the caller, NULL when the input does not require grad
 * (points_requires_grad / features_requires_grad, backward.py:12-17); point_heuristic (V,2)
 * accumulates in place (rasterizer/function.py:52-55,92), NULL unless compute_point_heuristic. */
int gs_raster_bwd_f32(const float *points, const float *features, const int32_t *tile_ranges,
                      const int32_t *overlap_to_point, const float *image, const float *grad_image,
                      int64_t v, int64_t k, int32_t width, int32_t height, int32_t num_features,
                      const gs_raster_config *config, float *grad_points, float *grad_features,
                      float *point_heuristic, void *stream);
int gs_raster_bwd_f64(const double *points, const double *features, const int32_t *tile_ranges,
                      const int32_t *overlap_to_point, const double *image, const double *grad_image,
                      int64_t v, int64_t k, int32_t width, int32_t height, int32_t num_features,
                      const gs_raster_config *config, double *grad_points, double *grad_features,
                      double *point_heuristic, void *stream);

/* ---- Raster digest: per-Gaussian records shared by the tuned forward and backward kernels ----------------------
 * The reference's rasteriser reads the packed (V,7) Gaussians and (V,F) features once per (tile, splat) overlap
 * (rasterizer/forward.py:82-97 load_point / backward.py:95-112).  The tuned fp32 kernels (tile_size 16, no
 * antialias, 1..4 features) instead gather a 64-byte, 64-byte-aligned record per visible Gaussian holding the
 * exp-scaled inverse-sigma basis, alpha, depth, features, support radius and 1/sigma -- written once per frame by
 * gs_raster_digest_f32 and passed to gs_raster_fwd_digest_f32 / gs_raster_bwd_digest_f32 (same arguments as
 * gs_raster_fwd_median_f32 / gs_raster_bwd_f32 with the three input tensors replaced by the digest; median_image
 * may be NULL).  gs_raster_fwd_f32 / gs_raster_fwd_median_f32 / gs_raster_bwd_f32 keep the reference-shaped
 * argument lists and build the digest into library-owned scratch (one grow-only buffer per device and stream).
 * `depths` may be NULL when no median depth is wanted.  Returns GS_ERR_UNSUPPORTED for other configurations.   */
int gs_raster_digest_bytes(int64_t v, size_t *bytes);
int gs_raster_digest_f32(const float *points, const float *features, const float *depths, int64_t v,
                         int32_t num_features, const gs_raster_config *config, void *digest, void *stream);
int gs_raster_fwd_digest_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point,
                             int64_t v, int64_t k, int32_t width, int32_t height, int32_t num_features,
                             const gs_raster_config *config, double median_threshold, float *image,
                             float *image_alpha, float *visibility, float *median_image, void *stream);
int gs_raster_bwd_digest_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point,
                             const float *image, const float *grad_image, int64_t v, int64_t k, int32_t width,
                             int32_t height, int32_t num_features, const gs_raster_config *config,
                             float *grad_points, float *grad_features, float *point_heuristic, void *stream);
/* As above with dL/dimage addressed through three element strides (y, x, channel; host array): the gradient torch
 * hands a backward is often not contiguous -- an expanded scalar for sum()/mean() losses (strides 0,0,0), a permuted
 * (C,H,W) tensor -- and the reference pays a full-image copy for it (rasterizer/function.py:91 .contiguous()). */
int gs_raster_bwd_digest_strided_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point,
                                     const float *image, const float *grad_image,
                                     const int64_t *grad_image_strides_host, int64_t v, int64_t k, int32_t width,
                                     int32_t height, int32_t num_features, const gs_raster_config *config,
                                     float *grad_points, float *grad_features, float *point_heuristic, void *stream);

/* ---- Packed per-overlap records + bulk-copy staged raster kernels -----------------------------------------------
 * Replaces the synchronous cooperative gather in front of every batch of the reference's raster kernels
 * (rasterizer/forward.py:67-83, rasterizer/backward.py:100-118).  gs_raster_pack_f32 runs once per frame after the
 * sort: for every sorted overlap k it resolves overlap_to_point[k], recentres the splat on its tile and classifies it
 * against the tile's four 8x8 pixel blocks (exact test), writing
 *   records[k]        48 bytes (1..3 features) | 64 bytes (4 features): {tx0,ty0,ux,wx | uy,wy,alpha,depth | f.., mask}
 *   flush_records[k]  16 bytes (backward only; may be NULL): {mean - tile centre, 1/sigma.x, 1/sigma.y}
 * so that a tile's batch is one contiguous range, which gs_raster_fwd_packed_f32 / gs_raster_bwd_packed_f32 fetch with
 * cp.async.bulk (TMA engine) + mbarrier into double-buffered shared memory while the previous batch is swept.
 * Results are those of gs_raster_fwd_digest_f32 / gs_raster_bwd_digest_f32 (which pack into library scratch and call
 * these).  Alpha blending, tile_size 16, no antialias, 1..4 features; buffers 16-byte aligned. */
int gs_raster_pack_bytes(int64_t k, int32_t num_features, size_t *record_bytes, size_t *flush_bytes);
int gs_raster_pack_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t k,
                       int32_t width, int32_t height, int32_t num_features, void *records, void *flush_records,
                       void *stream);
/* Same records from the sorted tile-id array (one thread per overlap instead of one CTA per tile; the whole-frame
 * driver has the array from its tile sort).  tile_ranges_out (tiles, 2) or NULL: the same pass also writes the tile
 * ranges (= gs_tile_ranges_from_tiles; replaces find_ranges_kernel, mapper/tile_mapper.py:92-112). */
int gs_raster_pack_sorted_f32(const void *digest, const uint32_t *sorted_tiles, const int32_t *overlap_to_point,
                              int64_t k, int32_t width, int32_t height, int32_t num_features, void *records,
                              void *flush_records, int32_t *tile_ranges_out, void *stream);
int gs_raster_fwd_packed_f32(const void *records, const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t v,
                             int64_t k, int32_t width, int32_t height, int32_t num_features,
                             const gs_raster_config *config, double median_threshold, float *image, float *image_alpha,
                             float *visibility, float *median_image /* or NULL */, void *stream);
int gs_raster_bwd_packed_f32(const void *records, const void *flush_records, const int32_t *tile_ranges,
                             const int32_t *overlap_to_point, const float *image, const float *grad_image,
                             const int64_t *grad_image_strides_host /* (y, x, channel) element strides, or NULL */,
                             int64_t v, int64_t k, int32_t width, int32_t height, int32_t num_features,
                             const gs_raster_config *config, float *grad_points, float *grad_features,
                             float *point_heuristic, void *stream);

/* ---- Whole-frame host drivers (renderer.py:22-108 in one call per phase) --------------------------------------
 * The per-stage entry points above mirror the reference's operators one to one, and a Python caller that chains
 * them spends 20-30 us of interpreter time per launch: at the bench workload the front end (13 short kernels,
 * two host reads) is then bound by the HOST, the GPU idling between launches.  These three drivers enqueue the
 * same kernels, in the same order, with the same arguments, from C -- nothing else changes, and the results are
 * bit-identical.  fp32, tuned raster configuration (tile_size 16, no antialias, 1..4 features, alpha blending).
 *
 * Buffers are still caller-owned.  Because V (visible Gaussians) is only known inside stage A, every per-visible
 * buffer is passed with CAPACITY n rows and the caller narrows it to V rows afterwards; the K-sized buffers are
 * allocated by the caller between stage A (which returns K) and stage B.
 *
 *   stage A : project+cull -> [host read V] -> compacted write (+ndc) -> depth order -> tile counts -> scan ->
 *             [host read K];  on the library's auxiliary stream, beside the mapper chain: SH evaluation (or the
 *             feature gather), zero fills of visibility / heuristic, the raster digest.
 *   stage B : ordered key emit -> stable tile sort -> tile ranges -> join the auxiliary stream -> raster pack ->
 *             raster forward (bulk-copy staged).
 *             (ordering = GS_ORDERING_BINNED: per-tile counts / slot emission / per-tile shared-memory sort
 *             instead, falling back to the above when a tile is too crowded for it.)
 *   backward: zero fills (auxiliary stream, beside the raster backward) -> raster backward -> projection backward
 *             on the caller's stream beside the SH backward (or feature scatter) on the auxiliary stream -> join.
 * The auxiliary stream is owned by the library (one per device) and fenced against `stream` with events.
 * ev_* : optional cudaEvent_t handles recorded around the raster launch on `stream` (NULL: not recorded).      */
#define GS_ORDERING_TWO_LEVEL 0
#define GS_ORDERING_BINNED 1

typedef struct gs_render_args {
  const float *position, *log_scaling, *rotation, *alpha_logit;   /* (n,3) (n,3) (n,4) (n,1) */
  const float *feature;            /* use_sh: SH coefficients (n,channels,(degree+1)^2); else features (n,channels) */
  const float *T_camera_world, *projection;
  int64_t n;
  int32_t width, height;
  double near_plane, far_plane, blur_cov, clamp_margin, median_threshold;
  int32_t use_sh, sh_degree, channels, use_depth16, want_median;
  int32_t ordering;                /* GS_ORDERING_TWO_LEVEL (default) | GS_ORDERING_BINNED; same result */
  gs_raster_config config;
  /* outputs with capacity n rows (V rows are written) */
  float *points;                   /* (n,7) */
  float *depths, *ndc;             /* (n,1) */
  int64_t *indexes;                /* (n) */
  float *features;                 /* (n,channels) */
  void *digest;                    /* 64 n bytes, 64-byte aligned */
  float *visibility;               /* (n) or NULL */
  float *heuristic;                /* (n,2) or NULL */
  float *camera_pos;               /* (3) */
  int32_t *order, *counts, *cum;   /* (n) (n) (n+1) */
  /* workspaces (gs_project_workspace_bytes / gs_depth_order_workspace_bytes / gs_tile_scan_workspace_bytes of n) */
  void *ws_project; size_t ws_project_bytes;
  void *ws_order; size_t ws_order_bytes;
  void *ws_scan; size_t ws_scan_bytes;
  /* image-sized outputs */
  float *image, *image_alpha, *median_image /* NULL unless want_median */;
  int32_t *tile_ranges;            /* (tiles,2) */
  void *ev_raster_start, *ev_raster_end;
  int32_t *tile_counts, *tile_cursor;   /* (tiles) each: binned ordering */
  int32_t *tile_totals;                 /* (2) */
  /* K-sized like tiles / overlap_to_point (same capacity), set before stage B: packed per-overlap raster records
   * (gs_raster_pack_bytes) written by stage B and kept for the backward.  NULL: stage B packs into library scratch. */
  void *records, *flush_records;
  void *hits;                           /* (n) x 16 bytes: hit records shared by tile count and emit, or NULL */
  int32_t tile_lo, tile_hi;             /* tile-sharded rank: bin / rasterise tiles [lo, hi) only (needs hits); (0, 0): all */
} gs_render_args;

int gs_render_stage_a_f32(const gs_render_args *args, int64_t *v_out, int64_t *k_out, int64_t *max_per_tile_out,
                          void *stream);
int gs_render_stage_b_f32(const gs_render_args *args, int64_t v, int64_t k, int64_t max_per_tile,
                          int64_t k_stride /* >= k */,
                          uint32_t *tiles /* (2,k_stride) */,
                          int32_t *overlap_to_point /* (2,k_stride): sorted result in row 1 */, void *ws_sort,
                          size_t ws_sort_bytes, void *stream);
/* Stage A, then stage B straight away when the caller's K-sized buffers (capacity k_capacity, e.g. sized from the
 * previous frame) are large enough: *stage_b_done = 1.  Otherwise *stage_b_done = 0 and the caller allocates and
 * calls gs_render_stage_b_f32 itself. */
int gs_render_forward_f32(const gs_render_args *args, int64_t k_capacity, uint32_t *tiles, int32_t *overlap_to_point,
                          void *ws_sort, size_t ws_sort_bytes, int64_t *v_out, int64_t *k_out,
                          int64_t *max_per_tile_out, int32_t *stage_b_done, void *stream);

typedef struct gs_render_bwd_args {
  const float *position, *log_scaling, *rotation, *alpha_logit, *feature, *T_camera_world, *projection;
  int64_t n, v, k;
  int32_t width, height;
  double blur_cov, clamp_margin;
  int32_t use_sh, sh_degree, channels;
  int32_t d_image_strided;         /* != 0: d_image is addressed through d_image_strides (below) */
  gs_raster_config config;
  /* saved by the forward */
  const int64_t *indexes;
  const float *features, *image, *camera_pos;
  const void *digest;
  const int32_t *overlap_to_point, *tile_ranges;
  const void *records, *flush_records;   /* stage B's packed records, or NULL (re-packed from the digest) */
  /* incoming gradients */
  const float *d_image;            /* (H,W,channels): contiguous, or any element strides with d_image_strided */
  const float *d_depths;           /* (v,1) or NULL (zero) */
  /* scratch / accumulated: grad_points (v,7) and grad_features (v,channels) are zero-filled here unless
   * grad_*_preset != 0 (the caller already stored incoming gradients in them) */
  float *grad_points, *grad_features;
  int32_t grad_points_preset, grad_features_preset;
  float *heuristic;                /* (v,2) accumulated in place, or NULL */
  /* outputs (zero-filled here; any may be NULL) */
  float *d_position, *d_log_scaling, *d_rotation, *d_alpha_logit, *d_T_camera_world, *d_projection;
  float *d_feature;                /* same shape as feature */
  void *ev_raster_start, *ev_raster_end;
  int64_t d_image_strides[3];      /* element strides (y, x, channel) of d_image when d_image_strided */
  /* use_sh: gradient of the loss w.r.t. the camera centre (3) through the SH view directions, or NULL.  The reference
   * gets it from autograd through camera_params.camera_position = inverse(T_camera_world)[:3,3]
   * (perspective/params.py:78-80); the caller chains it into d_T_camera_world (zero-filled here). */
  float *d_camera_pos;
  /* which parts of the backward to enqueue (0 = all).  A view-parallel caller runs GS_BWD_RASTER (zero fills + raster
   * backward), launches its exchange of the colour gradients, then GS_BWD_PROJECT, and replaces GS_BWD_FEATURE (SH
   * backward / feature scatter) by the exchange's own kernel. */
  int32_t phases;
} gs_render_bwd_args;
#define GS_BWD_RASTER 1
#define GS_BWD_FEATURE 2
#define GS_BWD_PROJECT 4

int gs_render_backward_f32(const gs_render_bwd_args *args, void *stream);

/* ---- N4: Morton ordering of a point cloud (misc/morton_sort.py:95-130) ------------------------------------------
 * gs_morton_codes64 replaces code_points64_kernel: codes[i] = 63-bit Morton code of the grid cell of points[i]
 * (cell = clamp((p - lower) / inc, 0, grid_size - 1) per axis, 21 bits each, x lowest), ids[i] = i.  lower / inc are
 * HOST arrays of 3 floats.  Follow with gs_sort_pairs(codes, ids, 8-byte keys, bits [0, 63)) = cuda_lib.radix_argsort. */
int gs_morton_codes64(const float *points, int64_t n, const float *lower_host, const float *inc_host,
                      int64_t grid_size, uint64_t *codes, int32_t *ids, void *stream);

/* ---- N3: sparse, visibility-weighted optimiser step (the step right after backward) ---------------------------
 * gs_optim_step_f32 replaces the four Taichi kernels of optim/fractional_adam.py:7-86 and
 * optim/fractional_laprop.py:6-76 (algorithm x scalar | vector second moment) and, when `param` is not NULL, the torch
 * passes that follow them in optim/fractional.py:131-147,184-199: clip to +-lr*clip (clip <= 0: none), * mask_lr[j],
 * * point_lr[idx], non-finite -> 0, param[idx] -= step * (1 - exp(-2 w)).  Rows: indexes (M) i64 into the (N, D)
 * tensors; weight (M); m_state (N,D); v_state (N,D) scalar kinds | (N) vector kinds; total_weight (N) already
 * advanced by the caller.  grad_scale (M, may be NULL): gradients are divided by (grad_scale[i] + grad_smooth)
 * first (optim/visibility_aware.py:97-99).  lr_step (M,D) receives the reference kernels' output (before clip /
 * masks) and may be NULL when param is given.
 * gs_optim_update_visibility_f32 replaces update_visibility + the step-count update
 * (optim/visibility_aware.py:37-48,90-91): running_vis[idx] = lerp(beta, vis^4, running_vis[idx]^4)^(1/4),
 * weight_out[i] = vis / max(running_vis[idx], eps), total_weight[idx] += weight_out[i].                       */
#define GS_OPTIM_ADAM 0
#define GS_OPTIM_LAPROP 1
int gs_optim_step_f32(int32_t algorithm, int32_t vector, int32_t bias_correction, const int64_t *indexes,
                      const float *weight, const float *grad_scale, double grad_smooth, int64_t m_rows, int32_t d,
                      float *m_state, float *v_state, const float *total_weight, const float *grad, double lr,
                      double beta1, double beta2, double eps, float *lr_step, float *param, double clip,
                      const float *mask_lr, const float *point_lr, void *stream);
int gs_optim_update_visibility_f32(float *running_vis, const float *visibility, const int64_t *indexes,
                                   float *total_weight, double beta, double eps, int64_t m_rows, float *weight_out,
                                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GSPLAT_B200_H */
