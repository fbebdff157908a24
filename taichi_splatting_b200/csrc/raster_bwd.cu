// raster_bwd.cu -- R9: C ABI of the backward gradient sweep and dispatch between the tuned kernel
// (raster_bwd_t.cu: fp32, 16x16 tiles, plain pdf, 1..4 features) and the generic one (raster_generic.cu).
//
// Semantics: _backward_kernel, rasterizer/backward.py:50-225 and gaussian_pdf_with_grad,
// taichi_lib/generic.py:320-336: re-walk front to back with (total_weight, remaining = image - sum f w),
// stop a pixel at total_weight >= saturate_threshold, clamp passes gradient through (D4).
#include "raster_common.cuh"

namespace gs {

template <typename real>
int raster_bwd_generic(const real *points, const real *features, const int32_t *ranges, const int32_t *o2p,
                       const real *image, const real *grad_image, int width, int height, int F,
                       const gs_raster_config *cfg, real *grad_points, real *grad_features, real *heuristic,
                       cudaStream_t stream);

// raster_bwd_t.cu
template <int F>
int launch_bwd_transpose(const float4 *records, const float4 *flush_records, const int32_t *ranges, const int32_t *o2p,
                         const float *image, const float *grad_image, const RasterParams<float> &P, int tiles,
                         float *grad_points, float *grad_features, float *heuristic, cudaStream_t stream);

int raster_digest_f32(const float *points, const float *features, const float *depths, int64_t v, int F,
                      double alpha_threshold, void *digest, cudaStream_t stream);   // raster_digest.cu
int raster_pack_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t k,
                    int32_t width, int32_t height, int32_t F, void *records, void *flush, cudaStream_t stream,
                    const uint32_t *sorted_tiles = nullptr, int32_t *ranges_out = nullptr);   // raster_pack.cu

constexpr int kTileB = 16;

static bool bwd_tuned(const gs_raster_config *cfg, int F) {
  return cfg->tile_size == kTileB && !cfg->antialias && F >= 1 && F <= 4;
}

// Tuned kernel on packed per-overlap records (raster_pack.cu).
static int raster_bwd_packed_impl(const void *records, const void *flush_records, const int32_t *tile_ranges,
                                  const int32_t *overlap_to_point, const float *image, const float *grad_image,
                                  int64_t k, int32_t width, int32_t height, int32_t F, const gs_raster_config *cfg,
                                  float *grad_points, float *grad_features, float *point_heuristic, cudaStream_t stream,
                                  const int64_t *grad_image_strides = nullptr) {
  GS_CHECK_ARG(cfg != nullptr, "raster_bwd: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_bwd: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_point_heuristic || point_heuristic != nullptr || k == 0, "raster_bwd: compute_point_heuristic needs a buffer");
  GS_CHECK_ARG((records != nullptr && flush_records != nullptr) || k == 0, "raster_bwd: packed records are NULL");
  if (!bwd_tuned(cfg, F)) {
    set_error("raster_bwd (packed): needs tile_size 16, no antialias, 1..4 features");
    return GS_ERR_UNSUPPORTED;
  }
  if (k == 0) return GS_OK;
  RasterParams<float> P = make_params<float>(cfg, width, height, F);
  if (grad_image_strides != nullptr) {
    P.gs_y = grad_image_strides[0]; P.gs_x = grad_image_strides[1]; P.gs_c = grad_image_strides[2];
  }
  const int tiles = P.tiles_wide * ((height + kTileB - 1) / kTileB);
  const float4 *r = reinterpret_cast<const float4 *>(records), *fl = reinterpret_cast<const float4 *>(flush_records);
  switch (F) {
    case 1: return launch_bwd_transpose<1>(r, fl, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
    case 2: return launch_bwd_transpose<2>(r, fl, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
    case 3: return launch_bwd_transpose<3>(r, fl, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
    default: return launch_bwd_transpose<4>(r, fl, tile_ranges, overlap_to_point, image, grad_image, P, tiles, grad_points, grad_features, point_heuristic, stream);
  }
}

// Digest given: pack the per-overlap records into library scratch (after `prefix_bytes` the caller already uses),
// then the tuned kernel.
static int raster_bwd_digest_impl(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point,
                                  const float *image, const float *grad_image, int64_t v, int64_t k, int32_t width,
                                  int32_t height, int32_t F, const gs_raster_config *cfg, float *grad_points,
                                  float *grad_features, float *point_heuristic, cudaStream_t stream,
                                  const int64_t *grad_image_strides = nullptr, size_t prefix_bytes = 0) {
  GS_CHECK_ARG(cfg != nullptr, "raster_bwd: config is NULL");
  GS_CHECK_ARG(digest != nullptr || v == 0, "raster_bwd: digest is NULL");
  if (!bwd_tuned(cfg, F)) {
    set_error("raster_bwd (digest): needs tile_size 16, no antialias, 1..4 features");
    return GS_ERR_UNSUPPORTED;
  }
  if (v == 0 || k == 0) return GS_OK;
  const size_t rec_bytes = (size_t)k * (F <= 3 ? 48 : 64), flush_bytes = (size_t)k * 16;
  unsigned char *ws = (unsigned char *)stream_workspace(stream, prefix_bytes + rec_bytes + flush_bytes);
  if (ws == nullptr) return GS_ERR_CUDA;
  void *records = ws + prefix_bytes, *flush = ws + prefix_bytes + rec_bytes;
  int rc = raster_pack_f32(digest, tile_ranges, overlap_to_point, k, width, height, F, records, flush, stream);
  if (rc != GS_OK) return rc;
  return raster_bwd_packed_impl(records, flush, tile_ranges, overlap_to_point, image, grad_image, k, width, height, F,
                                cfg, grad_points, grad_features, point_heuristic, stream, grad_image_strides);
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
  if (gs::bwd_tuned(cfg, F)) {
    // reference-shaped entry: digest the raw (V,7) points + features into library scratch first, the packed records
    // follow it in the same scratch
    if (v == 0 || k == 0) return GS_OK;
    const size_t digest_bytes = gs::align_up((size_t)v * 64, 256);
    const size_t total = digest_bytes + (size_t)k * (F <= 3 ? 48 : 64) + (size_t)k * 16;
    void *digest = gs::stream_workspace(stream, total);
    if (digest == nullptr) return GS_ERR_CUDA;
    int rc = gs::raster_digest_f32(points, features, nullptr, v, F, cfg->alpha_threshold, digest, stream);
    if (rc != GS_OK) return rc;
    return gs::raster_bwd_digest_impl(digest, tile_ranges, overlap_to_point, image, grad_image, v, k, width, height, F,
                                      cfg, grad_points, grad_features, point_heuristic, stream, nullptr, digest_bytes);
  }
  return gs::raster_bwd_generic<float>(points, features, tile_ranges, overlap_to_point, image, grad_image, width,
                                       height, F, cfg, grad_points, grad_features, point_heuristic, stream);
}

extern "C" int gs_raster_bwd_digest_f32(const void *digest, const int32_t *tile_ranges,
                                        const int32_t *overlap_to_point, const float *image, const float *grad_image,
                                        int64_t v, int64_t k, int32_t width, int32_t height, int32_t F,
                                        const gs_raster_config *cfg, float *grad_points, float *grad_features,
                                        float *point_heuristic, void *stream_) {
  return gs::raster_bwd_digest_impl(digest, tile_ranges, overlap_to_point, image, grad_image, v, k, width, height, F,
                                    cfg, grad_points, grad_features, point_heuristic, (cudaStream_t)stream_);
}

extern "C" int gs_raster_bwd_digest_strided_f32(const void *digest, const int32_t *tile_ranges,
                                                const int32_t *overlap_to_point, const float *image,
                                                const float *grad_image, const int64_t *grad_image_strides_host,
                                                int64_t v, int64_t k, int32_t width, int32_t height, int32_t F,
                                                const gs_raster_config *cfg, float *grad_points, float *grad_features,
                                                float *point_heuristic, void *stream_) {
  GS_CHECK_ARG(grad_image_strides_host != nullptr, "raster_bwd (strided): strides is NULL");
  return gs::raster_bwd_digest_impl(digest, tile_ranges, overlap_to_point, image, grad_image, v, k, width, height, F,
                                    cfg, grad_points, grad_features, point_heuristic, (cudaStream_t)stream_,
                                    grad_image_strides_host);
}

extern "C" int gs_raster_bwd_packed_f32(const void *records, const void *flush_records, const int32_t *tile_ranges,
                                        const int32_t *overlap_to_point, const float *image, const float *grad_image,
                                        const int64_t *grad_image_strides_host, int64_t v, int64_t k, int32_t width,
                                        int32_t height, int32_t F, const gs_raster_config *cfg, float *grad_points,
                                        float *grad_features, float *point_heuristic, void *stream_) {
  (void)v;
  return gs::raster_bwd_packed_impl(records, flush_records, tile_ranges, overlap_to_point, image, grad_image, k, width,
                                    height, F, cfg, grad_points, grad_features, point_heuristic,
                                    (cudaStream_t)stream_, grad_image_strides_host);
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
