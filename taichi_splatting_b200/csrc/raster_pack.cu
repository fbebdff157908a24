// raster_pack.cu -- per-overlap splat records in SORTED order, the staging source of the bulk-copy raster kernels.
//
// The reference's raster kernels stage a tile's splats with a cooperative, synchronous gather through
// overlap_to_point (rasterizer/forward.py:67-83, backward.py:100-118): two dependent global latencies (index, then
// the scattered rows) in front of every batch.  Here ONE pass over the K sorted overlaps, right after the sort,
// resolves the indirection and writes everything a tile needs as a contiguous run of fixed-size records:
//
//   records[k]  (48 bytes, 1..3 features | 64 bytes, 4 features), tile-centred, ready for the sweep:
//       Q0 = { tx0, ty0, ux, wx }      (tx, ty) = X (ux, wx) + Y (uy, wy) + (tx0, ty0), (X, Y) = pixel - tile centre
//       Q1 = { uy, wy, alpha, f0 }
//       Q2 = { f1, f2, depth, mask }   mask = bit w set: the splat can reach 8x8 pixel block w of its tile
//      [Q2 = { f1, f2, f3, depth }, Q3 = { mask, 0, 0, 0 }   with 4 features]
//     ordered so that a sweep which does not need the depth (every sweep but the forward's first few splats of a
//     pixel, which track the median depth) reads Q0, Q1 and only HALF of Q2: a warp-wide broadcast LDS.64 costs 1.63
//     cycles of the shared-memory data pipe against 2.69 for an LDS.128 (profiles/r01p_micro.txt)
//   flush[k]    (16 bytes, backward only) = { mean - tile centre, 1/sigma.x, 1/sigma.y }
//
// so the raster kernels fetch a batch with a single `cp.async.bulk` (bulk_copy.cuh) while they sweep the previous
// one, and spend no instruction on staging arithmetic.  Because this pass runs with full parallelism over K, it can
// afford the EXACT block-vs-ellipse test (raster_common.cuh: block_reaches_support, ~60 flops per block) for the
// forward too, which the in-kernel staging could not (7.72 -> 7.33 swept block entries per Gaussian).
// Extra traffic: 48 K written once + read twice (fwd, bwd) instead of 2 x 64 K gathered -- about the same bytes, but
// streamed.
#include "raster_common.cuh"

namespace gs {

constexpr int kPackThreads = 128;
#ifndef GS_PACK_EXACT
#define GS_PACK_EXACT 1   // exact block-vs-ellipse test after the separating-axis tests (7.72 -> 7.33 block entries per Gaussian)
#endif

__device__ __forceinline__ float rcp_approx_pack(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float sqrt_approx_pack(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int RECW, bool FLUSH>
__device__ __forceinline__ void pack_one(const float4 *__restrict__ digest, const int32_t *__restrict__ overlap_to_point,
                                         int64_t k, float tile_cx, float tile_cy, float4 *__restrict__ records,
                                         float4 *__restrict__ flush) {
  const float4 *rec = digest + 4 * (int64_t)overlap_to_point[k];
  const float4 R0 = __ldg(rec), R1 = __ldg(rec + 1), R2 = __ldg(rec + 2), R3 = __ldg(rec + 3);
  const float ux = R0.z, wx = R0.w, uy = R1.x, wy = R1.y, rcs = R3.x;
  const float ddx = R0.x - tile_cx, ddy = R0.y - tile_cy;
  const float tx0 = -fmaf(ux, ddx, uy * ddy), ty0 = -fmaf(wx, ddx, wy * ddy);
  unsigned mask = 0;
  if (rcs > 0.f) {
    // separating-axis tests against the support ellipse's bounding box and oriented box, then the exact test.  The
    // approximate rcp / sqrt (1 ulp-level error) sit under the 1e-4 relative margin of `sc`.
    const float sc = rcs * rcp_approx_pack(fabsf(ux * wy - uy * wx)) * 1.0001f;
    const float ex = sc * sqrt_approx_pack(fmaf(uy, uy, wy * wy)), ey = sc * sqrt_approx_pack(fmaf(ux, ux, wx * wx));
    const float hu = (fabsf(ux) + fabsf(uy)) * 3.5f + rcs;   // pixel centres of a block span +-3.5 around its centre
    const float hw = (fabsf(wx) + fabsf(wy)) * 3.5f + rcs;
#if GS_PACK_EXACT
    const SupportMetric metric = support_metric(ux, wx, uy, wy, rcs);
#endif
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float ox = (w & 1) ? 4.0f : -4.0f, oy = (w >> 1) ? 4.0f : -4.0f;   // block centre - tile centre
      const float t0x = fmaf(ux, ox, fmaf(uy, oy, tx0)), t0y = fmaf(wx, ox, fmaf(wy, oy, ty0));
      bool hit = (fabsf(ox - ddx) - 3.5f <= ex) && (fabsf(oy - ddy) - 3.5f <= ey) && (fabsf(t0x) <= hu) && (fabsf(t0y) <= hw);
#if GS_PACK_EXACT
      if (hit) hit = block_reaches_support(metric, t0x, t0y, ux, wx, uy, wy);
#endif
      mask |= hit ? (1u << w) : 0u;
    }
  }
  float4 *out = records + (int64_t)RECW * k;
  out[0] = make_float4(tx0, ty0, ux, wx);
  out[1] = make_float4(R1.x, R1.y, R1.z, R2.x);
  if (RECW == 3) {
    out[2] = make_float4(R2.y, R2.z, R1.w, __uint_as_float(mask));
  } else {
    out[2] = make_float4(R2.y, R2.z, R2.w, R1.w);
    out[3] = make_float4(__uint_as_float(mask), 0.f, 0.f, 0.f);
  }
  if (FLUSH) flush[k] = make_float4(ddx, ddy, R3.y, R3.z);
}

// one CTA per tile, walking the tile's range (callers that only have tile_ranges: the operator API)
template <int RECW, bool FLUSH>
__global__ void __launch_bounds__(kPackThreads)
raster_pack_kernel(const float4 *__restrict__ digest, const int32_t *__restrict__ ranges,
                   const int32_t *__restrict__ overlap_to_point, int tiles_wide, float4 *__restrict__ records,
                   float4 *__restrict__ flush) {
  const int tile = blockIdx.x;
  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  const float tile_cx = (float)((tile % tiles_wide) * 16) + 8.0f, tile_cy = (float)((tile / tiles_wide) * 16) + 8.0f;
  for (int k = start + (int)threadIdx.x; k < end; k += kPackThreads)
    pack_one<RECW, FLUSH>(digest, overlap_to_point, k, tile_cx, tile_cy, records, flush);
}

// one thread per overlap, tile id read from the sorted key array (the whole-frame driver has it): perfectly balanced
template <int RECW, bool FLUSH>
__global__ void __launch_bounds__(kPackThreads)
raster_pack_flat_kernel(const float4 *__restrict__ digest, const uint32_t *__restrict__ sorted_tiles,
                        const int32_t *__restrict__ overlap_to_point, int64_t k_total, int tiles_wide,
                        float4 *__restrict__ records, float4 *__restrict__ flush, int32_t *__restrict__ ranges_out) {
  const int64_t k = (int64_t)blockIdx.x * kPackThreads + threadIdx.x;
  if (k >= k_total) return;
  const int tile = (int)sorted_tiles[k];
  if (ranges_out != nullptr) {
    // the tile ranges (find_ranges_kernel, mapper/tile_mapper.py:92-112) fall out of the same pass over the sorted
    // tile ids: a range ends where the next overlap belongs to another tile (ranges_out is zero-filled beforehand)
    const int next = k + 1 < k_total ? (int)sorted_tiles[k + 1] : -1;
    if (next != tile) {
      ranges_out[2 * tile + 1] = (int32_t)(k + 1);
      if (next >= 0) ranges_out[2 * next] = (int32_t)(k + 1);
    }
  }
  const float tile_cx = (float)((tile % tiles_wide) * 16) + 8.0f, tile_cy = (float)((tile / tiles_wide) * 16) + 8.0f;
  pack_one<RECW, FLUSH>(digest, overlap_to_point, k, tile_cx, tile_cy, records, flush);
}

int raster_pack_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point, int64_t k,
                    int32_t width, int32_t height, int32_t F, void *records, void *flush, cudaStream_t stream,
                    const uint32_t *sorted_tiles, int32_t *ranges_out) {
  GS_CHECK_ARG(F >= 1 && F <= 4, "raster_pack: 1..4 features, got %d", F);
  GS_CHECK_ARG(width > 0 && height > 0, "raster_pack: bad image size %dx%d", width, height);
  if (k == 0) return GS_OK;
  GS_CHECK_ARG(digest != nullptr && (tile_ranges != nullptr || sorted_tiles != nullptr) && overlap_to_point != nullptr &&
                   records != nullptr, "raster_pack: NULL buffer");
  GS_CHECK_ARG((reinterpret_cast<uintptr_t>(records) & 15) == 0 && (reinterpret_cast<uintptr_t>(flush) & 15) == 0,
               "raster_pack: record buffers must be 16-byte aligned");
  const int tiles_wide = (width + 15) / 16, tiles = tiles_wide * ((height + 15) / 16);
  const float4 *d = reinterpret_cast<const float4 *>(digest);
  float4 *r = reinterpret_cast<float4 *>(records), *f = reinterpret_cast<float4 *>(flush);
  const unsigned flat_grid = (unsigned)ceil_div(k, kPackThreads);
#define GS_PACK(RECW_, FLUSH_)                                                                                        \
  do {                                                                                                                \
    if (sorted_tiles != nullptr)                                                                                      \
      raster_pack_flat_kernel<RECW_, FLUSH_><<<flat_grid, kPackThreads, 0, stream>>>(d, sorted_tiles, overlap_to_point, k, \
                                                                                    tiles_wide, r, f, ranges_out);    \
    else                                                                                                              \
      raster_pack_kernel<RECW_, FLUSH_><<<tiles, kPackThreads, 0, stream>>>(d, tile_ranges, overlap_to_point, tiles_wide, r, f); \
  } while (0)
  if (F <= 3) { if (flush) GS_PACK(3, true); else GS_PACK(3, false); }
  else        { if (flush) GS_PACK(4, true); else GS_PACK(4, false); }
#undef GS_PACK
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // namespace gs

extern "C" int gs_raster_pack_bytes(int64_t k, int32_t num_features, size_t *record_bytes, size_t *flush_bytes) {
  GS_CHECK_ARG(k >= 0 && num_features >= 1 && num_features <= 4, "raster_pack_bytes: bad arguments");
  if (record_bytes) *record_bytes = (size_t)k * (num_features <= 3 ? 48 : 64);
  if (flush_bytes) *flush_bytes = (size_t)k * 16;
  return GS_OK;
}

extern "C" int gs_raster_pack_f32(const void *digest, const int32_t *tile_ranges, const int32_t *overlap_to_point,
                                  int64_t k, int32_t width, int32_t height, int32_t num_features, void *records,
                                  void *flush_records, void *stream) {
  return gs::raster_pack_f32(digest, tile_ranges, overlap_to_point, k, width, height, num_features, records,
                             flush_records, (cudaStream_t)stream, nullptr, nullptr);
}

extern "C" int gs_raster_pack_sorted_f32(const void *digest, const uint32_t *sorted_tiles,
                                         const int32_t *overlap_to_point, int64_t k, int32_t width, int32_t height,
                                         int32_t num_features, void *records, void *flush_records,
                                         int32_t *tile_ranges_out, void *stream_) {
  GS_CHECK_ARG(sorted_tiles != nullptr || k == 0, "raster_pack_sorted: sorted_tiles is NULL");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (tile_ranges_out != nullptr) {   // untouched tiles stay (0, 0) (tile_mapper.py:186-188)
    const int64_t tiles = (int64_t)((width + 15) / 16) * ((height + 15) / 16);
    GS_CUDA(cudaMemsetAsync(tile_ranges_out, 0, sizeof(int32_t) * 2 * tiles, stream));
  }
  return gs::raster_pack_f32(digest, nullptr, overlap_to_point, k, width, height, num_features, records, flush_records,
                             stream, sorted_tiles, tile_ranges_out);
}
