// raster_fwd_bulk.cu -- R8: front-to-back alpha compositing, tuned fp32 / 16x16-tile kernel with BULK-COPY staging.
//
// Semantics: _forward_kernel, rasterizer/forward.py:22-135 in alpha-blending mode (each overlap composited exactly
// once, D1; early-out only below config.forward_saturate_eps, D2), plus the fused median-depth output of the
// reference's second pass (renderer.py:77-82).  Same sweep as raster_fwd.cu (8x8 pixel block per warp, two pixels per
// lane, packed f32x2 arithmetic, per-warp hit lists); what differs is where the splat records come from:
//
//   reference        cooperative synchronous gather through overlap_to_point, then a block barrier (forward.py:67-83)
//   raster_fwd.cu    the same gather from the 64-byte digest + per-overlap staging arithmetic in the kernel
//   this kernel      the tile's records were written in sorted order by raster_pack.cu, so a batch is ONE contiguous
//                    range: thread 0 issues `cp.async.bulk` (TMA engine, SASS UBLKCP) into one of two shared-memory
//                    buffers and arms an mbarrier with the byte count; the copy of batch b+1 is in flight while batch
//                    b is swept, and batch b+2 is issued the moment batch b's buffer is free.  No staging arithmetic,
//                    no index indirection, no LSU traffic for staging, one block barrier per batch.
//
// Measured and rejected (profiles/r02/r02e_barrier_free.md): a batch loop WITHOUT block barriers -- warps count themselves
// out of a buffer and the last one flushes / refills it, so warps run ahead of each other.  The four warps of a tile
// carry different loads (their 8x8 blocks hold different numbers of splats); letting the light warps run ahead and exit
// leaves the heavy warp alone at the end of every tile while the CTA still holds its shared memory, and a lone warp
// hides none of its own latency: resident warps fell from 24.5 % to 17.8 % of peak and the backward went from 1.02 to
// 1.65 ms (forward 0.39 -> 0.43 ms).  The barrier keeps the light warps' work spread over the tile's lifetime, where
// it overlaps the heavy warp's stalls.
#include <type_traits>

#include "bulk_copy.cuh"
#include "packed_f32.cuh"
#include "raster_common.cuh"

namespace gs {
namespace fwdb {

constexpr int kTile = 16;
constexpr int kWarps = 4;             // one warp per 8x8 pixel block of the tile
constexpr int kThreads = kWarps * 32;
#ifndef GS_FWDB_BATCH
#define GS_FWDB_BATCH 128
#endif
constexpr int kBatch = GS_FWDB_BATCH;   // records per bulk copy
static_assert(kBatch % kThreads == 0, "one flush slot per thread and batch slice");
#ifndef GS_FWDB_UNROLL
#define GS_FWDB_UNROLL 8   // 8: half the list padding of 16 (measured 0.02 ms faster at the bench workload)
#endif
constexpr int kUnroll = GS_FWDB_UNROLL;
static_assert(kUnroll == 8 || kUnroll == 16, "sweep chunk: 8 or 16 splats");
#ifndef GS_FWDB_MIN_BLOCKS
#define GS_FWDB_MIN_BLOCKS 8
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Third quad of a raster_pack record: { f1, f2, depth, mask } (1..3 features) | { f1, f2, f3, depth } (4 features); f0
// travels in Q1.w.  Only as much of it is read as the sweep needs.
template <int F, bool DEPTH>
__device__ __forceinline__ void load_tail(const unsigned char *q2, float f0, float (&feat)[4], float &depth) {
  feat[0] = f0; feat[1] = feat[2] = feat[3] = 0.f; depth = 0.f;
  if (F == 4 || DEPTH) {
    const float4 t = *reinterpret_cast<const float4 *>(q2);
    feat[1] = t.x; feat[2] = t.y;
    if (F == 4) { feat[3] = t.z; depth = t.w; } else { depth = t.z; }
  } else if (F == 3) {
    const float2 t = *reinterpret_cast<const float2 *>(q2);
    feat[1] = t.x; feat[2] = t.y;
  } else if (F == 2) {
    feat[1] = *reinterpret_cast<const float *>(q2);
  }
}

template <int RECW>
struct Smem {
  float4 rec[2][(kBatch + 1) * RECW];   // two landing buffers of raster_pack records (+1: null record for list padding)
  float vis[2][kBatch];                 // per-splat visibility of the batch (double-buffered like the records)
  alignas(16) unsigned list[kWarps][kBatch + kUnroll];   // byte offsets of the records a warp must visit
  alignas(8) uint64_t full[2];          // mbarriers: "buffer b holds its batch"
};

// N values per lane -> every lane of group g = lane / (32 / N) ends with the warp-wide sum of value g in v[0]
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N], int lane) {
  const unsigned full = 0xffffffffu;
  int off = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, off);
    }
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(full, v[0], off);
}

template <int F, bool VIS, bool MEDIAN, int RECW>
__global__ void __launch_bounds__(kThreads, GS_FWDB_MIN_BLOCKS)
raster_fwd_bulk_kernel(const float4 *__restrict__ records, const int32_t *__restrict__ ranges,
                       const int32_t *__restrict__ overlap_to_point, RasterParams<float> P, float median_lim,
                       float *__restrict__ image, float *__restrict__ image_alpha, float *__restrict__ visibility,
                       float *__restrict__ median_image) {
  __shared__ Smem<RECW> sm;
  constexpr unsigned kRecBytes = 16u * RECW;
  constexpr int kMaskWord = RECW == 3 ? 11 : 12;
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % P.tiles_wide) * kTile, tile_y0 = (tile / P.tiles_wide) * kTile;
  const int bx = (warp & 1) * 8 + (lane & 7), by = (warp >> 1) * 8 + (lane >> 3);
  const int px = tile_x0 + bx, py[2] = {tile_y0 + by, tile_y0 + by + 4};
  const bool in_bounds[2] = {px < P.width && py[0] < P.height, px < P.width && py[1] < P.height};
  const float lx = (float)bx - 7.5f, ly0 = (float)by - 7.5f;
  const f32x2 lx2 = pk(lx, lx), lyp = pk(ly0, ly0 + 4.0f);   // the lane's column; the rows of its two pixels
  const float clamp_max = P.clamp_max, thr = P.thr, eps = P.fwd_eps;
  const float median_trans = 1.0f - median_lim;   // sum of weights < lim  <=>  transmittance > 1 - lim

  float accum[2][F];
#pragma unroll
  for (int c = 0; c < F; ++c) accum[0][c] = accum[1][c] = 0.f;
  // transmittance 1 - sum of weights; 0 outside the image, where nothing can contribute (forward.py:49-52)
  float trans[2] = {in_bounds[0] ? 1.f : 0.f, in_bounds[1] ? 1.f : 0.f};
  float median[2] = {0.f, 0.f};

  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  const int nbatches = (end - start + kBatch - 1) / kBatch;

  auto issue = [&](int b) {   // thread 0 only: arm the barrier with the byte count, hand the range to the copy engine
    const int base = start + b * kBatch;
    const uint32_t bytes = (uint32_t)min(kBatch, end - base) * kRecBytes;
    mbar_arrive_expect_tx(&sm.full[b & 1], bytes);
    bulk_copy_g2s(sm.rec[b & 1], records + (int64_t)RECW * base, bytes, &sm.full[b & 1]);
  };

  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_fence_init();
#pragma unroll
    for (int b = 0; b < 2; ++b)   // null record: alpha = 0 never passes the threshold
#pragma unroll
      for (int q = 0; q < RECW; ++q) sm.rec[b][kBatch * RECW + q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nbatches > 0) issue(0);
    if (nbatches > 1) issue(1);
  }
  if (VIS) {
#pragma unroll
    for (int j = tid; j < kBatch; j += kThreads) { sm.vis[0][j] = 0.f; sm.vis[1][j] = 0.f; }
  }
  __syncthreads();

  for (int b = 0; b < nbatches; ++b) {
    const int buf = b & 1;
    const int base = start + b * kBatch, nb = min(kBatch, end - base);
    // ids of the splats this thread flushes after the sweep: loaded now, used then
    int my_id[kBatch / kThreads];
    if (VIS) {
#pragma unroll
      for (int s = 0; s < kBatch / kThreads; ++s) {
        const int j = tid + s * kThreads;
        my_id[s] = j < nb ? overlap_to_point[base + j] : 0;
      }
    }
    mbar_wait(&sm.full[buf], (uint32_t)(b >> 1) & 1u);   // the batch has landed
    const unsigned char *rb = reinterpret_cast<const unsigned char *>(sm.rec[buf]);
    const unsigned *words = reinterpret_cast<const unsigned *>(sm.rec[buf]);

    // per-warp ordered compaction of the splats that can touch this warp's 8x8 pixels
    int nhit = 0;
    if (!__all_sync(full, trans[0] <= eps && trans[1] <= eps)) {
      for (int c = 0; c < nb; c += 32) {
        const int j = c + lane;
        const bool hit = j < nb && ((words[j * (4 * RECW) + kMaskWord] >> warp) & 1u);
        const unsigned bal = __ballot_sync(full, hit);
        if (hit) sm.list[warp][nhit + __popc(bal & ((1u << lane) - 1))] = kRecBytes * (unsigned)j;
        nhit += __popc(bal);
      }
      if (lane < kUnroll) sm.list[warp][nhit + lane] = kRecBytes * (unsigned)kBatch;  // pad with the null record
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
          const float4 A = *reinterpret_cast<const float4 *>(rb + off);
          const float4 B = *reinterpret_cast<const float4 *>(rb + off + 16);
          float feat[4], depth;
          load_tail<F, kMedian>(rb + off + 32, B.w, feat, depth);
          // t = lx (ux, wx) + ly (uy, wy) + t0, kept component-major over the pixel pair -- tx2 = (tx of pixel 0, of
          // pixel 1) -- so that |t|^2 of both pixels is one FMUL2 + one FFMA2 (scalar operands broadcast for free)
          float cx, cy;
          upk(fma2(lx2, pk(A.z, A.w), pk(A.x, A.y)), cx, cy);
          const f32x2 tx2 = fma2(pk(B.x, B.x), lyp, pk(cx, cx)), ty2 = fma2(pk(B.y, B.y), lyp, pk(cy, cy));
          float q0, q1;
          upk(fma2(tx2, tx2, mul2(ty2, ty2)), q0, q1);
          const float g0 = ex2_approx(-q0), g1 = ex2_approx(-q1);
          float alpha[2], weight[2];
          upk(mul2(pk(g0, g1), pk(B.z, B.z)), alpha[0], alpha[1]);
          alpha[0] = fminf(alpha[0], clamp_max);
          alpha[1] = fminf(alpha[1], clamp_max);
          // no per-splat early-out test: a pixel below eps keeps compositing (as the reference does, D2) until the
          // whole warp is below eps; pixels outside the image have trans == 0 and so weight == 0.
          const bool hit[2] = {alpha[0] > thr, alpha[1] > thr};
          upk(mul2(pk(alpha[0], alpha[1]), pk(trans[0], trans[1])), weight[0], weight[1]);
          weight[0] = hit[0] ? weight[0] : 0.f;
          weight[1] = hit[1] ? weight[1] : 0.f;
          if (kMedian) {   // the splat that crosses the limit is the last one entered below it
            median[0] = (hit[0] && trans[0] > median_trans) ? depth : median[0];
            median[1] = (hit[1] && trans[1] > median_trans) ? depth : median[1];
          }
          upk(sub2(pk(trans[0], trans[1]), pk(weight[0], weight[1])), trans[0], trans[1]);
          const f32x2 w2 = pk(weight[0], weight[1]);
#pragma unroll
          for (int c = 0; c < F; ++c)
            upk(fma2(pk(feat[c], feat[c]), w2, pk(accum[0][c], accum[1][c])), accum[0][c], accum[1][c]);
          wv[u] = weight[0] + weight[1];
        }
      };
      if (MEDIAN && !__all_sync(full, trans[0] <= median_trans && trans[1] <= median_trans))
        sweep(std::true_type{});
      else
        sweep(std::false_type{});
      if (VIS) {
        warp_transpose_reduce<kUnroll>(wv, lane);
        constexpr int kGroup = 32 / kUnroll;
        const int h = h0 + lane / kGroup;
        if ((lane % kGroup) == 0 && h < nhit && wv[0] != 0.f) atomicAdd(&sm.vis[buf][sm.list[warp][h] / kRecBytes], wv[0]);
      }
      if (__all_sync(full, trans[0] <= eps && trans[1] <= eps)) break;
    }
    // every warp is through with this buffer (records and visibility sums); the barrier also decides, identically for
    // every thread, whether all pixels of the tile are done
    const int all_done = __syncthreads_and(__all_sync(full, trans[0] <= eps && trans[1] <= eps));
    if (tid == 0 && b + 2 < nbatches && !all_done) issue(b + 2);   // refill the buffer just released
    if (VIS) {
#pragma unroll
      for (int s = 0; s < kBatch / kThreads; ++s) {
        const int j = tid + s * kThreads;
        if (j < nb) {
          const float vsum = sm.vis[buf][j];
          if (vsum != 0.f) { atomicAdd(visibility + my_id[s], vsum); sm.vis[buf][j] = 0.f; }
        }
      }
    }
    if (all_done) {
      // batch b + 1 was handed to the copy engine earlier: it must land before this CTA's shared memory is released
      if (b + 1 < nbatches) mbar_wait(&sm.full[buf ^ 1], (uint32_t)((b + 1) >> 1) & 1u);
      break;
    }
  }

#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (!in_bounds[p]) continue;
    const int64_t pix = (int64_t)py[p] * P.width + px;
    float *out = image + pix * F;
#pragma unroll
    for (int c = 0; c < F; ++c) out[c] = accum[p][c];
    image_alpha[pix] = 1.0f - trans[p];
    if (MEDIAN) median_image[pix] = trans[p] <= median_trans ? median[p] : 0.f;
  }
}

template <int F>
static int launch(const float4 *records, const int32_t *ranges, const int32_t *o2p, const RasterParams<float> &P,
                  float median_lim, int tiles, float *image, float *image_alpha, float *visibility,
                  float *median_image, cudaStream_t stream) {
  constexpr int RECW = F <= 3 ? 3 : 4;
  const bool vis = P.vis && visibility != nullptr;
  const bool med = median_image != nullptr;
#define GS_FWDB(VIS, MED)                                                                                        \
  raster_fwd_bulk_kernel<F, VIS, MED, RECW><<<tiles, kThreads, 0, stream>>>(records, ranges, o2p, P, median_lim, \
                                                                            image, image_alpha, visibility, median_image)
  if (med) { if (vis) GS_FWDB(true, true); else GS_FWDB(false, true); }
  else     { if (vis) GS_FWDB(true, false); else GS_FWDB(false, false); }
#undef GS_FWDB
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // namespace fwdb
}  // namespace gs

extern "C" int gs_raster_fwd_packed_f32(const void *records, const int32_t *tile_ranges,
                                        const int32_t *overlap_to_point, int64_t v, int64_t k, int32_t width,
                                        int32_t height, int32_t F, const gs_raster_config *cfg,
                                        double median_threshold, float *image, float *image_alpha, float *visibility,
                                        float *median_image, void *stream_) {
  using namespace gs;
  (void)v;
  GS_CHECK_ARG(cfg != nullptr, "raster_fwd_packed: config is NULL");
  GS_CHECK_ARG(width > 0 && height > 0, "raster_fwd_packed: bad image size %dx%d", width, height);
  GS_CHECK_ARG(!cfg->compute_visibility || visibility != nullptr || k == 0, "raster_fwd_packed: compute_visibility needs a visibility buffer");
  GS_CHECK_ARG(records != nullptr || k == 0, "raster_fwd_packed: records is NULL");
  GS_CHECK_ARG((reinterpret_cast<uintptr_t>(records) & 15) == 0, "raster_fwd_packed: records must be 16-byte aligned");
  if (cfg->tile_size != fwdb::kTile || cfg->antialias || !cfg->use_alpha_blending || F < 1 || F > 4) {
    set_error("raster_fwd_packed: needs tile_size 16, no antialias, alpha blending, 1..4 features");
    return GS_ERR_UNSUPPORTED;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  RasterParams<float> P = make_params<float>(cfg, width, height, F);
  const int tiles = P.tiles_wide * ((height + fwdb::kTile - 1) / fwdb::kTile);
  const float median_lim = (float)(1.0 - median_threshold);
  const float4 *r = reinterpret_cast<const float4 *>(records);
  switch (F) {
    case 1: return fwdb::launch<1>(r, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
    case 2: return fwdb::launch<2>(r, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
    case 3: return fwdb::launch<3>(r, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
    default: return fwdb::launch<4>(r, tile_ranges, overlap_to_point, P, median_lim, tiles, image, image_alpha, visibility, median_image, stream);
  }
}
