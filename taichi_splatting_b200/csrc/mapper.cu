// mapper.cu -- R3..R7: per-tile overlap binning, scan, (tile|depth) keys, radix sort, tile ranges.
//
// Semantics: taichi_lib/grid_query.py:9-93 (OBB-vs-tile query), mapper/tile_mapper.py:35-146 (keys, ranges),
// cuda_lib/full_cumsum.cu, cuda_lib/radix_sort_pairs.cu (CUB scan / onesweep radix sort).
// Integer outputs must be bit-exact against the CPU oracle, so the OBB query uses explicitly rounded
// fp32 ops (__fmul_rn/__fadd_rn: no FMA contraction), IEEE sqrt/div, and a correctly rounded fp32 log
// obtained through fp64 -- the same recipe as oracle/gs_oracle.c.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"

namespace gs {

struct ObbQuery {
  float inv00, inv01, inv10, inv11;
  float relx, rely;
  int minx, miny, spanx, spany;
};

__device__ __forceinline__ ObbQuery obb_grid_query(const float *__restrict__ g, int w_pad, int h_pad, int ts,
                                                   float thr) {
  ObbQuery q;
  q.spanx = q.spany = 0; q.minx = q.miny = 0;
  q.inv00 = q.inv01 = q.inv10 = q.inv11 = q.relx = q.rely = 0.f;
  float mx = g[0], my = g[1], ax = g[2], ay = g[3], sx = g[4], sy = g[5], alpha = g[6];
  if (!(alpha > thr)) return q;  // D18
  float gscale = __fsqrt_rn(__fmul_rn(2.0f, (float)log((double)__fdiv_rn(alpha, thr))));
  float scx = __fmul_rn(sx, gscale), scy = __fmul_rn(sy, gscale);
  float a2x = -ay, a2y = ax;
  float v1x = __fmul_rn(ax, scx), v1y = __fmul_rn(ay, scx), v2x = __fmul_rn(a2x, scy), v2y = __fmul_rn(a2y, scy);
  float ex = __fsqrt_rn(__fadd_rn(__fmul_rn(v1x, v1x), __fmul_rn(v2x, v2x)));
  float ey = __fsqrt_rn(__fadd_rn(__fmul_rn(v1y, v1y), __fmul_rn(v2y, v2y)));
  float lox = __fsub_rn(mx, ex), loy = __fsub_rn(my, ey), hix = __fadd_rn(mx, ex), hiy = __fadd_rn(my, ey);
  q.inv00 = __fdiv_rn(ax, scx); q.inv01 = __fdiv_rn(ay, scx);
  q.inv10 = __fdiv_rn(a2x, scy); q.inv11 = __fdiv_rn(a2y, scy);
  float fts = (float)ts;
  int max_tx = (w_pad - 1) / ts, max_ty = (h_pad - 1) / ts;
  float flx = floorf(__fdiv_rn(lox, fts)), fly = floorf(__fdiv_rn(loy, fts));
  float chx = ceilf(__fdiv_rn(hix, fts)), chy = ceilf(__fdiv_rn(hiy, fts));
  const float BIG = 1.0e9f;
  if (!(flx > -BIG && flx < BIG && fly > -BIG && fly < BIG && chx > -BIG && chx < BIG && chy > -BIG && chy < BIG))
    return q;
  int min_tx = max((int)flx, 0), min_ty = max((int)fly, 0);
  int max_bx = min(max((int)chx, min_tx + 1), max_tx + 1);
  int max_by = min(max((int)chy, min_ty + 1), max_ty + 1);
  q.minx = min_tx; q.miny = min_ty;
  q.spanx = max_bx - min_tx; q.spany = max_by - min_ty;
  q.relx = __fsub_rn((float)(min_tx * ts), mx);
  q.rely = __fsub_rn((float)(min_ty * ts), my);
  return q;
}

__device__ __forceinline__ bool test_tile(const ObbQuery &q, int u, int v, int ts) {
  float lx = __fadd_rn(q.relx, (float)(u * ts)), ly = __fadd_rn(q.rely, (float)(v * ts));
  float ux = __fadd_rn(lx, (float)ts), uy = __fadd_rn(ly, (float)ts);
  // axis 0
  float a0 = __fmul_rn(q.inv00, lx), a1 = __fmul_rn(q.inv00, ux);
  float b0 = __fmul_rn(q.inv01, ly), b1 = __fmul_rn(q.inv01, uy);
  float p0 = __fadd_rn(a0, b0), p1 = __fadd_rn(a1, b0), p2 = __fadd_rn(a1, b1), p3 = __fadd_rn(a0, b1);
  float mn = fminf(fminf(p0, p1), fminf(p2, p3)), mxv = fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
  if (mn > 1.0f || mxv < -1.0f) return false;
  // axis 1
  a0 = __fmul_rn(q.inv10, lx); a1 = __fmul_rn(q.inv10, ux);
  b0 = __fmul_rn(q.inv11, ly); b1 = __fmul_rn(q.inv11, uy);
  p0 = __fadd_rn(a0, b0); p1 = __fadd_rn(a1, b0); p2 = __fadd_rn(a1, b1); p3 = __fadd_rn(a0, b1);
  mn = fminf(fminf(p0, p1), fminf(p2, p3)); mxv = fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
  return !(mn > 1.0f || mxv < -1.0f);
}

__global__ void __launch_bounds__(128)
tile_count_kernel(const float *__restrict__ gaussians, int64_t v, int w_pad, int h_pad, int ts, float thr,
                  int32_t *__restrict__ counts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= v) return;
  ObbQuery q = obb_grid_query(gaussians + 7 * i, w_pad, h_pad, ts, thr);
  int c = 0;
  for (int u = 0; u < q.spanx; ++u)
    for (int w = 0; w < q.spany; ++w) c += test_tile(q, u, w, ts) ? 1 : 0;
  counts[i] = c;
}

template <typename key_t, bool DEPTH16>
__global__ void __launch_bounds__(128)
tile_emit_kernel(const float *__restrict__ gaussians, const float *__restrict__ depths,
                 const int32_t *__restrict__ cum, int64_t v, int w_pad, int h_pad, int ts, float thr,
                 key_t *__restrict__ keys, int32_t *__restrict__ overlap_to_point) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= v) return;
  ObbQuery q = obb_grid_query(gaussians + 7 * i, w_pad, h_pad, ts, thr);
  int tiles_wide = w_pad / ts;
  int64_t k = cum[i];
  float depth = depths[i];
  uint32_t dbits;
  if (DEPTH16) {
    float c = fminf(fmaxf(depth, 0.0f), 1.0f);
    dbits = (uint32_t)__fmul_rn(c, 65535.0f);
  } else {
    dbits = __float_as_uint(depth);
  }
  for (int u = 0; u < q.spanx; ++u)
    for (int w = 0; w < q.spany; ++w)
      if (test_tile(q, u, w, ts)) {
        key_t tile_id = (key_t)((q.minx + u) + (q.miny + w) * tiles_wide);
        keys[k] = DEPTH16 ? (key_t)((tile_id << 16) | dbits) : (key_t)((tile_id << (sizeof(key_t) * 4)) | dbits);
        overlap_to_point[k] = (int32_t)i;
        ++k;
      }
}

// ---- two-level ordering (same final order as the reference's 48-bit LSD sort, far less sort traffic) -----------
// An LSD radix sort over (tile | depth) is a stable sort by depth followed by a stable sort by tile.  All overlaps
// of one Gaussian share its depth, so the depth passes can run on the V Gaussians BEFORE the expansion to K
// overlaps: sort (depth bits, index) once, emit the overlaps in that order with the tile id as the only key, then
// a stable sort on ceil(log2 T) bits finishes it.  Ties keep ascending Gaussian index in both schemes.
template <bool DEPTH16>
__global__ void depth_key_kernel(const float *__restrict__ depths, int64_t v, uint32_t *__restrict__ keys,
                                 int32_t *__restrict__ ids) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= v) return;
  float d = depths[i];
  keys[i] = DEPTH16 ? (uint32_t)__fmul_rn(fminf(fmaxf(d, 0.0f), 1.0f), 65535.0f) : __float_as_uint(d);
  ids[i] = (int32_t)i;
}

__global__ void __launch_bounds__(128)
tile_count_ordered_kernel(const float *__restrict__ gaussians, const int32_t *__restrict__ order, int64_t v, int w_pad,
                          int h_pad, int ts, float thr, int32_t *__restrict__ counts) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= v) return;
  ObbQuery q = obb_grid_query(gaussians + 7 * (int64_t)order[r], w_pad, h_pad, ts, thr);
  int c = 0;
  for (int u = 0; u < q.spanx; ++u)
    for (int w = 0; w < q.spany; ++w) c += test_tile(q, u, w, ts) ? 1 : 0;
  counts[r] = c;
}

__global__ void __launch_bounds__(128)
tile_emit_ordered_kernel(const float *__restrict__ gaussians, const int32_t *__restrict__ order,
                         const int32_t *__restrict__ cum, int64_t v, int w_pad, int h_pad, int ts, float thr,
                         uint32_t *__restrict__ tile_keys, int32_t *__restrict__ overlap_to_point) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= v) return;
  int32_t i = order[r];
  ObbQuery q = obb_grid_query(gaussians + 7 * (int64_t)i, w_pad, h_pad, ts, thr);
  int tiles_wide = w_pad / ts;
  int64_t k = cum[r];
  for (int u = 0; u < q.spanx; ++u)
    for (int w = 0; w < q.spany; ++w)
      if (test_tile(q, u, w, ts)) {
        tile_keys[k] = (uint32_t)((q.minx + u) + (q.miny + w) * tiles_wide);
        overlap_to_point[k] = i;
        ++k;
      }
}

// ---- count and emit sharing ONE grid query ----------------------------------------------------------------------
// The count kernel already runs the OBB-vs-tile test for every tile of the Gaussian's span; instead of repeating the
// whole query (7 scattered loads, an fp64 log, 4 divisions, up to 16 separating-axis tests) in the emit kernel, it
// leaves a 16-byte "hit record" per Gaussian -- {minx, miny, spanx, spany} as 4 x u16 and a 64-bit mask of the accepted
// tiles in enumeration order (x outer, y inner) -- and the emit kernel only walks the set bits.  Spans of more than 64
// tiles (huge splats) are marked and re-queried.
__global__ void __launch_bounds__(128)
tile_count_hits_kernel(const float *__restrict__ gaussians, const int32_t *__restrict__ order, int64_t v, int w_pad,
                       int h_pad, int ts, float thr, int tile_lo, int tile_hi, int32_t *__restrict__ counts,
                       ulonglong2 *__restrict__ hits) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= v) return;
  // order == NULL: Gaussian r itself (counts / hits then sit at the Gaussian's own index, not at its depth rank)
  ObbQuery q = obb_grid_query(gaussians + 7 * (int64_t)(order != nullptr ? order[r] : (int32_t)r), w_pad, h_pad, ts, thr);
  int c = 0;
  unsigned long long mask = 0ull;
  const bool small = q.spanx * q.spany <= 64;
  const int tiles_wide = w_pad / ts;
  int bit = 0;
  for (int u = 0; u < q.spanx; ++u)
    for (int w = 0; w < q.spany; ++w, ++bit) {
      const int tile = (q.minx + u) + (q.miny + w) * tiles_wide;   // tile-sharded runs keep only [tile_lo, tile_hi)
      if (tile >= tile_lo && tile < tile_hi && test_tile(q, u, w, ts)) {
        ++c;
        if (small) mask |= 1ull << bit;
      }
    }
  counts[r] = c;
  const unsigned long long box = (unsigned long long)(unsigned)q.minx | ((unsigned long long)(unsigned)q.miny << 16) |
                                 ((unsigned long long)(unsigned)(small ? q.spanx : 0xffff) << 32) |
                                 ((unsigned long long)(unsigned)q.spany << 48);
  hits[r] = make_ulonglong2(box, mask);
}

template <bool HITS_BY_POINT>   // hit records indexed by Gaussian (count ran before / beside the depth sort) | by depth rank
__global__ void __launch_bounds__(128)
tile_emit_hits_kernel(const float *__restrict__ gaussians, const int32_t *__restrict__ order,
                      const int32_t *__restrict__ cum, const ulonglong2 *__restrict__ hits, int64_t v, int w_pad,
                      int h_pad, int ts, float thr, int tile_lo, int tile_hi, uint32_t *__restrict__ tile_keys,
                      int32_t *__restrict__ overlap_to_point) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= v) return;
  const int32_t i = order[r];
  const ulonglong2 h = hits[HITS_BY_POINT ? (int64_t)i : r];
  const int tiles_wide = w_pad / ts;
  int64_t k = cum[r];
  const int minx = (int)(h.x & 0xffff), miny = (int)((h.x >> 16) & 0xffff);
  const int spanx = (int)((h.x >> 32) & 0xffff), spany = (int)((h.x >> 48) & 0xffff);
  if (spanx == 0xffff) {   // span too large for the mask: run the query again
    ObbQuery q = obb_grid_query(gaussians + 7 * (int64_t)i, w_pad, h_pad, ts, thr);
    for (int u = 0; u < q.spanx; ++u)
      for (int w = 0; w < q.spany; ++w) {
        const int tile = (q.minx + u) + (q.miny + w) * tiles_wide;
        if (tile >= tile_lo && tile < tile_hi && test_tile(q, u, w, ts)) {
          tile_keys[k] = (uint32_t)tile;
          overlap_to_point[k] = i;
          ++k;
        }
      }
    return;
  }
  unsigned long long mask = h.y;
  while (mask) {
    const int bit = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const int u = bit / spany, w = bit - u * spany;
    tile_keys[k] = (uint32_t)((minx + u) + (miny + w) * tiles_wide);
    overlap_to_point[k] = i;
    ++k;
  }
}

// ---- binned ordering (same final order again, no global sort at all) --------------------------------------------
// The final order is: by tile, then by depth bits, ties by ascending Gaussian index.  Count overlaps per TILE
// (atomics), turn the counts into tile offsets (= the tile ranges), let every overlap grab a slot inside its
// tile's segment in arrival order (atomic cursor), then sort each tile's segment on the 64-bit key
// (depth bits << 32 | Gaussian index) in shared memory.  The keys of a segment are distinct, so the result does not
// depend on the arrival order.  ~234 keys per tile at the bench workload: one small bitonic network per CTA instead
// of six onesweep passes over V + K pairs.
__global__ void __launch_bounds__(128)
tile_bin_count_kernel(const float *__restrict__ gaussians, int64_t v, int w_pad, int h_pad, int ts, float thr,
                      int32_t *__restrict__ tile_counts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= v) return;
  ObbQuery q = obb_grid_query(gaussians + 7 * i, w_pad, h_pad, ts, thr);
  const int tiles_wide = w_pad / ts;
  for (int u = 0; u < q.spanx; ++u)
    for (int w = 0; w < q.spany; ++w)
      if (test_tile(q, u, w, ts)) atomicAdd(tile_counts + (q.minx + u) + (q.miny + w) * tiles_wide, 1);
}

// one CTA: exclusive scan of the tile counts -> tile ranges (empty tiles stay (0,0), tile_mapper.py:188), slot
// cursors, total K and the largest tile population (both to pinned host words)
__global__ void __launch_bounds__(1024)
tile_bin_offsets_kernel(const int32_t *__restrict__ tile_counts, int num_tiles, int32_t *__restrict__ ranges,
                        int32_t *__restrict__ cursor, int32_t *__restrict__ totals) {
  __shared__ int32_t warp_sum[32], warp_max[32];
  __shared__ int32_t part[1024];
  __shared__ int32_t pmax[1024];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (num_tiles + 1023) / 1024;
  const int lo = tid * per, hi = min(lo + per, num_tiles);
  int32_t sum = 0, mx = 0;
  for (int t = lo; t < hi; ++t) { int32_t c = tile_counts[t]; sum += c; mx = max(mx, c); }
  // inclusive scan of the per-thread sums: warp shuffles, then the 32 warp totals
  int32_t incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int32_t a = __shfl_up_sync(0xffffffffu, incl, off);
    int32_t b = __shfl_xor_sync(0xffffffffu, mx, off);
    if (lane >= off) incl += a;
    mx = max(mx, b);
  }
  if (lane == 31) warp_sum[warp] = incl;
  if (lane == 0) warp_max[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_sum[lane], m = warp_max[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int32_t a = __shfl_up_sync(0xffffffffu, w, off);
      int32_t b = __shfl_xor_sync(0xffffffffu, m, off);
      if (lane >= off) w += a;
      m = max(m, b);
    }
    warp_sum[lane] = w;    // inclusive over warps
    warp_max[lane] = m;    // global maximum in every lane
  }
  __syncthreads();
  part[tid] = incl + (warp > 0 ? warp_sum[warp - 1] : 0);
  pmax[tid] = warp_max[0];
  __syncthreads();
  int32_t run = part[tid] - sum;
  for (int t = lo; t < hi; ++t) {
    const int32_t c = tile_counts[t];
    ranges[2 * t] = c > 0 ? run : 0;
    ranges[2 * t + 1] = c > 0 ? run + c : 0;
    cursor[t] = run;
    run += c;
  }
  if (tid == 1023) { totals[0] = part[1023]; totals[1] = pmax[1023]; }
}

template <bool DEPTH16>
__global__ void __launch_bounds__(128)
tile_bin_emit_kernel(const float *__restrict__ gaussians, const float *__restrict__ depths, int64_t v, int w_pad,
                     int h_pad, int ts, float thr, int32_t *__restrict__ cursor, uint64_t *__restrict__ keys) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= v) return;
  ObbQuery q = obb_grid_query(gaussians + 7 * i, w_pad, h_pad, ts, thr);
  const int tiles_wide = w_pad / ts;
  const float depth = depths[i];
  const uint32_t dbits = DEPTH16 ? (uint32_t)__fmul_rn(fminf(fmaxf(depth, 0.0f), 1.0f), 65535.0f) : __float_as_uint(depth);
  const uint64_t key = ((uint64_t)dbits << 32) | (uint32_t)i;
  for (int u = 0; u < q.spanx; ++u)
    for (int w = 0; w < q.spany; ++w)
      if (test_tile(q, u, w, ts)) {
        const int32_t slot = atomicAdd(cursor + (q.minx + u) + (q.miny + w) * tiles_wide, 1);
        keys[slot] = key;
      }
}

// one CTA per tile: bitonic sort of the tile's (depth bits | index) keys in shared memory
template <int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS)
tile_bin_sort_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ ranges,
                     int32_t *__restrict__ overlap_to_point) {
  __shared__ uint64_t sk[CAP];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int start = ranges[2 * tile], n = ranges[2 * tile + 1] - start;
  if (n <= 0) return;
  int p = 2;
  while (p < n) p <<= 1;
  for (int i = tid; i < p; i += THREADS) sk[i] = i < n ? keys[start + i] : ~0ull;
  __syncthreads();
  for (int k = 2; k <= p; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < (p >> 1); i += THREADS) {
        const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1)), b = a | j;
        const uint64_t x = sk[a], y = sk[b];
        const bool up = (a & k) == 0;
        if ((x > y) == up) { sk[a] = y; sk[b] = x; }
      }
      // With one pair per thread, pairs at distance j <= 32 stay inside the 64 elements a warp owns, so a run of
      // such steps only needs warp-level ordering; a CTA barrier separates it from any longer-distance step.
      const int next_j = j > 1 ? (j >> 1) : k;
      if (j > 32 || next_j > 32 || p > 2 * THREADS) __syncthreads(); else __syncwarp();
    }
  __syncthreads();
  for (int i = tid; i < n; i += THREADS) overlap_to_point[start + i] = (int32_t)(uint32_t)sk[i];
}

// counts[order[i]]: the scan runs over the depth order while the counts sit at the Gaussians' own indices
struct GatherCount {
  const int32_t *counts, *order;
  __host__ __device__ __forceinline__ int32_t operator()(int i) const { return counts[order[i]]; }
};
using GatherCountIt = cub::TransformInputIterator<int32_t, GatherCount, cub::CountingInputIterator<int>>;

__global__ void finish_scan_kernel(const int32_t *__restrict__ counts, const int32_t *__restrict__ order,
                                   int32_t *__restrict__ cum, int64_t v, int32_t *__restrict__ total_dev,
                                   volatile int32_t *total_mapped) {
  // cum[0..v-1] holds the exclusive scan; complete entry v (cuda_lib/full_cumsum.cu:6-10)
  int32_t t = cum[v - 1] + counts[order != nullptr ? order[v - 1] : v - 1];
  cum[v] = t;
  *total_dev = t;
  if (total_mapped != nullptr) {   // mapped pinned host word the host polls (common.cuh: kWordPending)
    *total_mapped = t;
    __threadfence_system();
  }
}

template <typename key_t>
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const key_t *__restrict__ keys, int64_t k, int shift, int32_t *__restrict__ ranges) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= k) return;
  const int max_tile = 65535;
  int tile_id = (int)(keys[idx] >> shift);
  int next_tile_id = max_tile;
  if (idx + 1 < k) next_tile_id = (int)(keys[idx + 1] >> shift);
  if (tile_id != next_tile_id) {
    ranges[2 * tile_id + 1] = (int32_t)(idx + 1);
    if (next_tile_id < max_tile) ranges[2 * next_tile_id] = (int32_t)(idx + 1);
  }
}

}  // namespace gs

extern "C" int gs_tile_count(const float *gaussians, int64_t v, int32_t w_pad, int32_t h_pad, int32_t ts,
                             double alpha_threshold, int32_t *counts, void *stream) {
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_count: image %dx%d not padded to tile %d", w_pad, h_pad, ts);
  GS_CHECK_ARG((int64_t)(w_pad / ts) * (h_pad / ts) < 65535, "tile dimensions (%d, %d) exceed maximum tile count (16 bit id), try increasing tile_size", h_pad / ts, w_pad / ts);
  if (v == 0) return GS_OK;
  gs::tile_count_kernel<<<(unsigned)gs::ceil_div(v, 128), 128, 0, (cudaStream_t)stream>>>(
      gaussians, v, w_pad, h_pad, ts, (float)alpha_threshold, counts);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_tile_scan_workspace_bytes(int64_t v, size_t *bytes) {
  size_t temp = 0;
  if (v > 0) cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)v);
  *bytes = gs::align_up(temp, 256) + 256;
  return GS_OK;
}

namespace gs {
static int tile_scan_impl(const int32_t *counts, int64_t v, int32_t *cum, void *workspace, size_t workspace_bytes,
                          int32_t *total_host, cudaStream_t stream, bool mapped, const int32_t *order = nullptr) {
  GS_CHECK_ARG(total_host != nullptr, "tile_scan: total_host is NULL");
  GS_CHECK_ARG(v >= 0 && v < (int64_t(1) << 31), "tile_scan: v out of range");
  if (v == 0) {
    GS_CUDA(cudaMemsetAsync(cum, 0, sizeof(int32_t), stream));
    *total_host = 0;
    return GS_OK;
  }
  size_t need = 0;
  gs_tile_scan_workspace_bytes(v, &need);
  if (workspace_bytes < need) {
    gs::set_error("tile_scan: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GS_ERR_WORKSPACE_TOO_SMALL;
  }
  int32_t *total_dev = (int32_t *)workspace;
  void *temp = (char *)workspace + 256;
  size_t temp_bytes = workspace_bytes - 256;
  if (order != nullptr) {
    GatherCountIt in(cub::CountingInputIterator<int>(0), GatherCount{counts, order});
    size_t gather_need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, gather_need, in, cum, (int)v);
    if (gather_need > temp_bytes) {
      gs::set_error("tile_scan: workspace too small for the gathered scan (%zu < %zu)", temp_bytes, gather_need);
      return GS_ERR_WORKSPACE_TOO_SMALL;
    }
    GS_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, cum, (int)v, stream));
  } else {
    GS_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, cum, (int)v, stream));
  }
  gs::finish_scan_kernel<<<1, 1, 0, stream>>>(counts, order, cum, v, total_dev, mapped ? total_host : nullptr);
  GS_LAUNCH_CHECK();
  if (!mapped) GS_CUDA(cudaMemcpyAsync(total_host, total_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  return GS_OK;
}

int tile_scan_word(const int32_t *counts, int64_t v, int32_t *cum, void *workspace, size_t workspace_bytes,
                   int32_t *word, bool word_is_mapped, const int32_t *order, cudaStream_t stream) {
  return tile_scan_impl(counts, v, cum, workspace, workspace_bytes, word, stream, word_is_mapped, order);
}

int tile_emit_hits_by_point(const float *gaussians, const int32_t *order, const int32_t *cum, const void *hits,
                            int64_t v, int32_t w_pad, int32_t h_pad, int32_t ts, double alpha_threshold,
                            int32_t tile_lo, int32_t tile_hi, uint32_t *tile_keys, int32_t *overlap_to_point,
                            cudaStream_t stream) {
  if (tile_lo == 0 && tile_hi == 0) tile_hi = 0x7fffffff;
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_emit: image not padded to tile size");
  if (v == 0) return GS_OK;
  tile_emit_hits_kernel<true><<<(unsigned)ceil_div(v, 128), 128, 0, stream>>>(
      gaussians, order, cum, reinterpret_cast<const ulonglong2 *>(hits), v, w_pad, h_pad, ts, (float)alpha_threshold,
      tile_lo, tile_hi, tile_keys, overlap_to_point);
  GS_LAUNCH_CHECK();
  return GS_OK;
}
}  // namespace gs

extern "C" int gs_tile_scan(const int32_t *counts, int64_t v, int32_t *cum, void *workspace, size_t workspace_bytes,
                            int32_t *total_host, void *stream_) {
  return gs::tile_scan_impl(counts, v, cum, workspace, workspace_bytes, total_host, (cudaStream_t)stream_, false);
}

extern "C" int gs_tile_emit_keys(const float *gaussians, const float *depths, const int32_t *cum, int64_t v,
                                 int32_t w_pad, int32_t h_pad, int32_t ts, double alpha_threshold,
                                 int32_t use_depth16, void *keys, int32_t *overlap_to_point, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_emit_keys: image not padded to tile size");
  if (v == 0) return GS_OK;
  unsigned grid = (unsigned)gs::ceil_div(v, 128);
  if (use_depth16)
    gs::tile_emit_kernel<uint32_t, true><<<grid, 128, 0, stream>>>(gaussians, depths, cum, v, w_pad, h_pad, ts,
                                                                   (float)alpha_threshold, (uint32_t *)keys, overlap_to_point);
  else
    gs::tile_emit_kernel<uint64_t, false><<<grid, 128, 0, stream>>>(gaussians, depths, cum, v, w_pad, h_pad, ts,
                                                                    (float)alpha_threshold, (uint64_t *)keys, overlap_to_point);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_sort_pairs_workspace_bytes(int64_t k, int32_t key_bytes, size_t *bytes) {
  size_t temp = 0;
  if (k > 0) {
    if (key_bytes == 8)
      cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                      (const int32_t *)nullptr, (int32_t *)nullptr, (int)k);
    else
      cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                      (const int32_t *)nullptr, (int32_t *)nullptr, (int)k);
  }
  *bytes = gs::align_up(temp, 256) + 256;
  return GS_OK;
}

extern "C" int gs_sort_pairs(const void *keys_in, const int32_t *values_in, void *keys_out, int32_t *values_out,
                             int64_t k, int32_t key_bytes, int32_t begin_bit, int32_t end_bit, void *workspace,
                             size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(key_bytes == 4 || key_bytes == 8, "sort_pairs: key_bytes must be 4 or 8 (got %d)", key_bytes);
  GS_CHECK_ARG(k >= 0 && k < (int64_t(1) << 31), "sort_pairs: k out of range");
  if (end_bit <= 0) end_bit = key_bytes * 8;  // radix_sort_pairs.cu:14
  GS_CHECK_ARG(begin_bit >= 0 && begin_bit < end_bit && end_bit <= key_bytes * 8, "sort_pairs: bad bit range");
  if (k == 0) return GS_OK;
  size_t temp_bytes = workspace_bytes;
  if (key_bytes == 8)
    GS_CUDA(cub::DeviceRadixSort::SortPairs(workspace, temp_bytes, (const uint64_t *)keys_in, (uint64_t *)keys_out,
                                            values_in, values_out, (int)k, begin_bit, end_bit, stream));
  else
    GS_CUDA(cub::DeviceRadixSort::SortPairs(workspace, temp_bytes, (const uint32_t *)keys_in, (uint32_t *)keys_out,
                                            values_in, values_out, (int)k, begin_bit, end_bit, stream));
  return GS_OK;
}

extern "C" int gs_tile_ranges(const void *sorted_keys, int64_t k, int32_t key_bytes, int32_t *tile_ranges,
                              int64_t num_tiles, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(key_bytes == 4 || key_bytes == 8, "tile_ranges: key_bytes must be 4 or 8");
  GS_CUDA(cudaMemsetAsync(tile_ranges, 0, sizeof(int32_t) * 2 * num_tiles, stream));
  if (k == 0) return GS_OK;
  unsigned grid = (unsigned)gs::ceil_div(k, 256);
  if (key_bytes == 8)
    gs::tile_ranges_kernel<uint64_t><<<grid, 256, 0, stream>>>((const uint64_t *)sorted_keys, k, 32, tile_ranges);
  else
    gs::tile_ranges_kernel<uint32_t><<<grid, 256, 0, stream>>>((const uint32_t *)sorted_keys, k, 16, tile_ranges);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_depth_order_workspace_bytes(int64_t v, size_t *bytes) {
  size_t temp = 0;
  if (v > 0)
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int)v);
  size_t n = gs::align_up((size_t)(v > 0 ? v : 0) * 4, 256);
  *bytes = 3 * n + gs::align_up(temp, 256) + 256;   // keys in/out, ids in, cub temp
  return GS_OK;
}

extern "C" int gs_depth_order(const float *depths, int64_t v, int32_t use_depth16, int32_t *order, void *workspace,
                              size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(v >= 0 && v < (int64_t(1) << 31), "depth_order: v out of range");
  if (v == 0) return GS_OK;
  size_t need = 0;
  gs_depth_order_workspace_bytes(v, &need);
  if (workspace_bytes < need) {
    gs::set_error("depth_order: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GS_ERR_WORKSPACE_TOO_SMALL;
  }
  size_t n = gs::align_up((size_t)v * 4, 256);
  uint32_t *keys_in = (uint32_t *)workspace, *keys_out = (uint32_t *)((char *)workspace + n);
  int32_t *ids_in = (int32_t *)((char *)workspace + 2 * n);
  void *temp = (char *)workspace + 3 * n;
  size_t temp_bytes = workspace_bytes - 3 * n;
  unsigned grid = (unsigned)gs::ceil_div(v, 256);
  if (use_depth16) gs::depth_key_kernel<true><<<grid, 256, 0, stream>>>(depths, v, keys_in, ids_in);
  else gs::depth_key_kernel<false><<<grid, 256, 0, stream>>>(depths, v, keys_in, ids_in);
  GS_LAUNCH_CHECK();
  GS_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, ids_in, order, (int)v, 0,
                                          use_depth16 ? 16 : 32, stream));
  return GS_OK;
}

extern "C" int gs_tile_count_ordered(const float *gaussians, const int32_t *order, int64_t v, int32_t w_pad,
                                     int32_t h_pad, int32_t ts, double alpha_threshold, int32_t *counts, void *stream) {
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_count: image %dx%d not padded to tile %d", w_pad, h_pad, ts);
  GS_CHECK_ARG((int64_t)(w_pad / ts) * (h_pad / ts) < 65535, "tile dimensions (%d, %d) exceed maximum tile count (16 bit id), try increasing tile_size", h_pad / ts, w_pad / ts);
  if (v == 0) return GS_OK;
  gs::tile_count_ordered_kernel<<<(unsigned)gs::ceil_div(v, 128), 128, 0, (cudaStream_t)stream>>>(
      gaussians, order, v, w_pad, h_pad, ts, (float)alpha_threshold, counts);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_tile_emit_ordered(const float *gaussians, const int32_t *order, const int32_t *cum, int64_t v,
                                    int32_t w_pad, int32_t h_pad, int32_t ts, double alpha_threshold,
                                    uint32_t *tile_keys, int32_t *overlap_to_point, void *stream) {
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_emit: image not padded to tile size");
  if (v == 0) return GS_OK;
  gs::tile_emit_ordered_kernel<<<(unsigned)gs::ceil_div(v, 128), 128, 0, (cudaStream_t)stream>>>(
      gaussians, order, cum, v, w_pad, h_pad, ts, (float)alpha_threshold, tile_keys, overlap_to_point);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_tile_count_ordered_hits(const float *gaussians, const int32_t *order, int64_t v, int32_t w_pad,
                                          int32_t h_pad, int32_t ts, double alpha_threshold, int32_t tile_lo,
                                          int32_t tile_hi, int32_t *counts, void *hits, void *stream) {
  if (tile_lo == 0 && tile_hi == 0) tile_hi = 0x7fffffff;   // (0, 0): every tile
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_count: image %dx%d not padded to tile %d", w_pad, h_pad, ts);
  GS_CHECK_ARG((int64_t)(w_pad / ts) * (h_pad / ts) < 65535, "tile dimensions (%d, %d) exceed maximum tile count (16 bit id), try increasing tile_size", h_pad / ts, w_pad / ts);
  GS_CHECK_ARG(hits != nullptr || v == 0, "tile_count_hits: hits is NULL");
  GS_CHECK_ARG((reinterpret_cast<uintptr_t>(hits) & 15) == 0, "tile_count_hits: hits must be 16-byte aligned");
  if (v == 0) return GS_OK;
  gs::tile_count_hits_kernel<<<(unsigned)gs::ceil_div(v, 128), 128, 0, (cudaStream_t)stream>>>(
      gaussians, order, v, w_pad, h_pad, ts, (float)alpha_threshold, tile_lo, tile_hi, counts,
      reinterpret_cast<ulonglong2 *>(hits));
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_tile_emit_hits(const float *gaussians, const int32_t *order, const int32_t *cum, const void *hits,
                                 int64_t v, int32_t w_pad, int32_t h_pad, int32_t ts, double alpha_threshold,
                                 int32_t tile_lo, int32_t tile_hi, uint32_t *tile_keys, int32_t *overlap_to_point,
                                 void *stream) {
  if (tile_lo == 0 && tile_hi == 0) tile_hi = 0x7fffffff;
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_emit: image not padded to tile size");
  if (v == 0) return GS_OK;
  gs::tile_emit_hits_kernel<false><<<(unsigned)gs::ceil_div(v, 128), 128, 0, (cudaStream_t)stream>>>(
      gaussians, order, cum, reinterpret_cast<const ulonglong2 *>(hits), v, w_pad, h_pad, ts, (float)alpha_threshold,
      tile_lo, tile_hi, tile_keys, overlap_to_point);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_tile_ranges_from_tiles(const uint32_t *sorted_tiles, int64_t k, int32_t *tile_ranges,
                                         int64_t num_tiles, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CUDA(cudaMemsetAsync(tile_ranges, 0, sizeof(int32_t) * 2 * num_tiles, stream));
  if (k == 0) return GS_OK;
  gs::tile_ranges_kernel<uint32_t><<<(unsigned)gs::ceil_div(k, 256), 256, 0, stream>>>(sorted_tiles, k, 0, tile_ranges);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

// ---- binned ordering --------------------------------------------------------------------------------------------
extern "C" int gs_tile_bin_count(const float *gaussians, int64_t v, int32_t w_pad, int32_t h_pad, int32_t ts,
                                 double alpha_threshold, int32_t *tile_counts, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_bin_count: image %dx%d not padded to tile %d", w_pad, h_pad, ts);
  const int64_t num_tiles = (int64_t)(w_pad / ts) * (h_pad / ts);
  GS_CHECK_ARG(num_tiles < 65535, "tile dimensions (%d, %d) exceed maximum tile count (16 bit id), try increasing tile_size", h_pad / ts, w_pad / ts);
  GS_CUDA(cudaMemsetAsync(tile_counts, 0, sizeof(int32_t) * num_tiles, stream));
  if (v == 0) return GS_OK;
  gs::tile_bin_count_kernel<<<(unsigned)gs::ceil_div(v, 128), 128, 0, stream>>>(gaussians, v, w_pad, h_pad, ts,
                                                                                 (float)alpha_threshold, tile_counts);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_tile_bin_offsets(const int32_t *tile_counts, int64_t num_tiles, int32_t *tile_ranges,
                                   int32_t *cursor, int32_t *totals_dev, int32_t *totals_host, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(num_tiles > 0 && num_tiles < 65535, "tile_bin_offsets: bad tile count %lld", (long long)num_tiles);
  GS_CHECK_ARG(totals_dev != nullptr && totals_host != nullptr, "tile_bin_offsets: totals is NULL");
  gs::tile_bin_offsets_kernel<<<1, 1024, 0, stream>>>(tile_counts, (int)num_tiles, tile_ranges, cursor, totals_dev);
  GS_LAUNCH_CHECK();
  GS_CUDA(cudaMemcpyAsync(totals_host, totals_dev, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  return GS_OK;
}

extern "C" int gs_tile_bin_emit(const float *gaussians, const float *depths, int64_t v, int32_t w_pad, int32_t h_pad,
                                int32_t ts, double alpha_threshold, int32_t use_depth16, int32_t *cursor,
                                uint64_t *keys, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(ts > 0 && w_pad % ts == 0 && h_pad % ts == 0, "tile_bin_emit: image not padded to tile size");
  if (v == 0) return GS_OK;
  const unsigned grid = (unsigned)gs::ceil_div(v, 128);
  if (use_depth16)
    gs::tile_bin_emit_kernel<true><<<grid, 128, 0, stream>>>(gaussians, depths, v, w_pad, h_pad, ts, (float)alpha_threshold, cursor, keys);
  else
    gs::tile_bin_emit_kernel<false><<<grid, 128, 0, stream>>>(gaussians, depths, v, w_pad, h_pad, ts, (float)alpha_threshold, cursor, keys);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_tile_bin_max_per_tile(void) { return 4096; }

extern "C" int gs_tile_bin_sort(const uint64_t *keys, const int32_t *tile_ranges, int64_t num_tiles,
                                int32_t max_per_tile, int32_t *overlap_to_point, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (max_per_tile > gs_tile_bin_max_per_tile()) {
    gs::set_error("tile_bin_sort: %d overlaps in one tile exceed the shared-memory sort capacity (%d); use the "
                  "two-level ordering", max_per_tile, gs_tile_bin_max_per_tile());
    return GS_ERR_UNSUPPORTED;
  }
  if (num_tiles == 0 || max_per_tile <= 0) return GS_OK;
  const unsigned grid = (unsigned)num_tiles;
  if (max_per_tile <= 512) gs::tile_bin_sort_kernel<512, 128><<<grid, 128, 0, stream>>>(keys, tile_ranges, overlap_to_point);
  else if (max_per_tile <= 1024) gs::tile_bin_sort_kernel<1024, 256><<<grid, 256, 0, stream>>>(keys, tile_ranges, overlap_to_point);
  else if (max_per_tile <= 2048) gs::tile_bin_sort_kernel<2048, 512><<<grid, 512, 0, stream>>>(keys, tile_ranges, overlap_to_point);
  else gs::tile_bin_sort_kernel<4096, 1024><<<grid, 1024, 0, stream>>>(keys, tile_ranges, overlap_to_point);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

// ---- N4: Morton ordering of a point cloud (misc/morton_sort.py:13-130) --------------------------------------------
namespace gs {
__device__ __forceinline__ uint64_t spread_bits64(uint64_t x) {   // morton_sort.py:23-31
  x &= 0x1fffffull;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(256)
morton_codes64_kernel(const float *__restrict__ points, int64_t n, float lx, float ly, float lz, float ix, float iy,
                      float iz, float max_cell, uint64_t *__restrict__ codes, int32_t *__restrict__ ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // Grid.grid_cell (morton_sort.py:49-53): cast(clamp((p - lower) / inc, 0, size - 1), u32)
  const float vx = __fdiv_rn(__fsub_rn(points[3 * i + 0], lx), ix);
  const float vy = __fdiv_rn(__fsub_rn(points[3 * i + 1], ly), iy);
  const float vz = __fdiv_rn(__fsub_rn(points[3 * i + 2], lz), iz);
  const uint64_t cx = (uint64_t)fminf(fmaxf(vx, 0.0f), max_cell), cy = (uint64_t)fminf(fmaxf(vy, 0.0f), max_cell);
  const uint64_t cz = (uint64_t)fminf(fmaxf(vz, 0.0f), max_cell);
  codes[i] = spread_bits64(cx) | (spread_bits64(cy) << 1) | (spread_bits64(cz) << 2);
  ids[i] = (int32_t)i;
}
}  // namespace gs

extern "C" int gs_morton_codes64(const float *points, int64_t n, const float *lower_host, const float *inc_host,
                                 int64_t grid_size, uint64_t *codes, int32_t *ids, void *stream) {
  GS_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "morton_codes64: n out of range");
  GS_CHECK_ARG(grid_size >= 1 && grid_size <= (int64_t(1) << 21), "morton_codes64: grid size must be in [1, 2^21]");
  GS_CHECK_ARG(lower_host != nullptr && inc_host != nullptr, "morton_codes64: lower / inc is NULL");
  if (n == 0) return GS_OK;
  gs::morton_codes64_kernel<<<(unsigned)gs::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      points, n, lower_host[0], lower_host[1], lower_host[2], inc_host[0], inc_host[1], inc_host[2],
      (float)(grid_size - 1), codes, ids);
  GS_LAUNCH_CHECK();
  return GS_OK;
}
