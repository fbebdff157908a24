// render.cu -- whole-frame host drivers: the render path of renderer.py:22-108 enqueued from C.
//
// Nothing here computes: each driver calls the per-stage entry points of this library (same kernels, same
// arguments, same order as the Python operator chain), so results are bit-identical to chaining the operators.
// What changes is who pays the launch cost.  Chained from Python, every launch costs 20-30 us of interpreter /
// allocator / ctypes time; the front end is 13 short kernels with two host reads in it, so the GPU idled between
// launches for ~0.2-0.3 ms per frame (torch.profiler timeline, profiles/).  From C a launch costs 2-3 us, and
// the work that does not feed the tile mapper -- SH evaluation, raster digest, zero fills, SH backward -- is put on
// an auxiliary stream so that it really runs beside the latency-bound mapper chain / raster backward.
#include <stdlib.h>

#include <atomic>
#include <map>
#include <mutex>

#include "common.cuh"

namespace gs {

static std::atomic<int> g_devices_seen{0};   // devices this process has rendered on (see use_mapped_words)

struct DeviceAux {
  cudaStream_t side_stream = nullptr;
  cudaEvent_t fence = nullptr, side_done = nullptr, raster_done = nullptr, fills_done = nullptr, bwd_join = nullptr;
  cudaEvent_t count_done = nullptr;
  int32_t *host_words = nullptr;   // pinned: [0] V, [1] K, [2] largest tile population (binned ordering), [4] K (fallback)
  // The events, the side stream and the pinned words are shared by every frame on this device, so the whole-frame
  // drivers serialise on this lock: two host threads (or two streams) rendering on one device take turns instead of
  // overwriting each other's V / K read-backs or re-recording an event the other frame still waits on.
  std::recursive_mutex frame_mu;
};

// Joins the auxiliary stream back into the caller's stream when a driver returns -- on the error paths too, so that
// torch-owned buffers already handed to work on the auxiliary stream are never freed / reused while it still runs.
struct SideJoin {
  cudaStream_t stream, side;
  cudaEvent_t ev;
  SideJoin(cudaStream_t stream_, cudaStream_t side_, cudaEvent_t ev_) : stream(stream_), side(side_), ev(ev_) {}
  ~SideJoin() {
    if (side == stream) return;
    if (cudaEventRecord(ev, side) == cudaSuccess) cudaStreamWaitEvent(stream, ev, 0);
  }
};

// One auxiliary state per (device, caller stream): frames enqueued on different streams of a device do not share
// events, side stream or read-back words.
static DeviceAux *device_aux(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, DeviceAux> table;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  DeviceAux &a = table[{dev, stream}];
  if (a.side_stream == nullptr) {
    bool new_device = true;
    for (const auto &entry : table) new_device = new_device && (entry.first.first != dev || &entry.second == &a);
    if (new_device) g_devices_seen.fetch_add(1, std::memory_order_relaxed);
    if (cudaStreamCreateWithFlags(&a.side_stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&a.fence, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&a.side_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&a.raster_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&a.fills_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&a.bwd_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&a.count_done, cudaEventDisableTiming);
    if (cudaHostAlloc((void **)&a.host_words, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
  }
  return &a;
}

// GS_SIDE_STREAM=0: enqueue the auxiliary work on the caller's stream as well (A/B switch)
static bool use_side_stream() {
  static int choice = -1;
  if (choice < 0) {
    const char *e = getenv("GS_SIDE_STREAM");
    choice = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return choice == 1;
}

static bool render_supported(const gs_raster_config &c, int channels) {
  return c.tile_size == 16 && !c.antialias && c.use_alpha_blending && channels >= 1 && channels <= 4;
}

__global__ void gather_rows_kernel(const float *__restrict__ src, const int64_t *__restrict__ indexes, int64_t v,
                                   int c, float *__restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v * c) return;
  dst[i] = src[indexes[i / c] * c + i % c];
}

__global__ void scatter_rows_kernel(const float *__restrict__ src, const int64_t *__restrict__ indexes, int64_t v,
                                    int c, float *__restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v * c) return;
  dst[indexes[i / c] * c + i % c] = src[i];
}

#define GS_TRY(expr)           \
  do {                         \
    int _rc = (expr);          \
    if (_rc != GS_OK) return _rc; \
  } while (0)

static inline int pad_to(int x, int ts) { return (x + ts - 1) / ts * ts; }

static inline int tile_bits(int64_t num_tiles) {
  int bits = 1;
  while ((int64_t(1) << bits) < num_tiles) ++bits;
  return bits;
}

}  // namespace gs

namespace gs {
static int stage_a_impl(const gs_render_args *a, int64_t *v_out, int64_t *k_out, int64_t *max_per_tile_out,
                        cudaStream_t stream, DeviceAux *aux, cudaStream_t side);
static int stage_b_impl(const gs_render_args *a, int64_t v, int64_t k, int64_t max_per_tile, int64_t k_stride,
                        uint32_t *tiles, int32_t *o2p, void *ws_sort, size_t ws_sort_bytes, cudaStream_t stream,
                        DeviceAux *aux);
}  // namespace gs

extern "C" int gs_render_stage_a_f32(const gs_render_args *a, int64_t *v_out, int64_t *k_out,
                                     int64_t *max_per_tile_out, void *stream_) {
  using namespace gs;
  GS_CHECK_ARG(a != nullptr && v_out != nullptr && k_out != nullptr && max_per_tile_out != nullptr,
               "render_stage_a: NULL argument");
  GS_CHECK_ARG(render_supported(a->config, a->channels), "render_stage_a: needs tile_size 16, no antialias, alpha blending, 1..4 features");
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceAux *aux = device_aux(stream);
  if (aux == nullptr) { set_error("render_stage_a: cannot create the auxiliary stream"); return GS_ERR_CUDA; }
  std::lock_guard<std::recursive_mutex> frame_lock(aux->frame_mu);
  GS_NVTX("gs_render_stage_a: project, SH, digest, depth order, tile count, scan");
  cudaStream_t side = use_side_stream() ? aux->side_stream : stream;
  const int rc = stage_a_impl(a, v_out, k_out, max_per_tile_out, stream, aux, side);
  if (rc != GS_OK) SideJoin(stream, side, aux->bwd_join);   // error after auxiliary work was enqueued: join before returning
  return rc;
}

// Waits for a count the device stores into mapped pinned memory (common.cuh: kWordPending).  The host spins on the
// word; every few thousand polls it asks the stream whether it failed or drained without publishing, so a faulting
// kernel ends in an error, not in a hang.
static int wait_mapped_word(volatile int32_t *word, cudaStream_t stream, const char *what, int64_t *out) {
  for (uint32_t spin = 1;; ++spin) {
    const int32_t value = *word;
    if (value != gs::kWordPending) { *out = value; return GS_OK; }
    if ((spin & 0xfffu) == 0) {
      const cudaError_t q = cudaStreamQuery(stream);
      if (q == cudaSuccess) {   // everything enqueued has finished: the store is visible now or never
        const int32_t last = *word;
        if (last != gs::kWordPending) { *out = last; return GS_OK; }
        gs::set_error("%s: the stream drained without publishing the count", what);
        return GS_ERR_CUDA;
      }
      if (q != cudaErrorNotReady) {
        gs::set_error("%s: %s", what, cudaGetErrorString(q));
        return GS_ERR_CUDA;
      }
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
}

// GS_COUNT_BESIDE_SORT=0: tile count after the depth sort on the caller's stream (A/B switch)
static bool count_beside_sort() {
  static const bool on = [] { const char *e = getenv("GS_COUNT_BESIDE_SORT"); return e == nullptr || e[0] != '0'; }();
  return on;
}

// Mapped count words are used while the process drives ONE device.  A single process rendering on two GPUs in turn
// (tests/test_gpu_multi.py::test_second_device_in_one_process) produced images that differed between the devices in
// 2 of 8 runs with the mapped words and in none of 6 runs with copy + synchronise (profiles/r02/r02ak_second_device.txt);
// not explained yet, so that configuration keeps the blocking read-back.  GS_MAPPED_COUNTS=0 turns them off altogether.
static bool use_mapped_words() {
  static const bool on = [] { const char *e = getenv("GS_MAPPED_COUNTS"); return e == nullptr || e[0] != '0'; }();
  return on && gs::g_devices_seen.load(std::memory_order_relaxed) <= 1;
}

// The tile count does not need the depth order, only the scan of its results does: the two-level ordering runs it on
// the auxiliary stream beside the four small radix passes of the depth sort (one wave of CTAs each, 2 us apart), with
// counts and hit records at the Gaussians' own indices; scan and key emission then read them through the order.
// A function of the arguments and the process-wide switches only, so stage A and stage B of a frame agree.
static bool hits_by_point(const gs_render_args *a) {
  return a->ordering != GS_ORDERING_BINNED && a->hits != nullptr && gs::use_side_stream() && count_beside_sort();
}

static int gs::stage_a_impl(const gs_render_args *a, int64_t *v_out, int64_t *k_out, int64_t *max_per_tile_out,
                            cudaStream_t stream, DeviceAux *aux, cudaStream_t side) {
  const gs_raster_config &c = a->config;
  const int64_t n = a->n;
  *v_out = 0; *k_out = 0; *max_per_tile_out = 0;

  // ---- projection: single-pass project + cull + ordered compaction (+ ndc depth) -> V ----
  const bool mapped = use_mapped_words();
  int64_t v = 0;
  if (mapped) {
    aux->host_words[0] = kWordPending;
    GS_TRY(project_compact_f32_mapped(a->position, a->log_scaling, a->rotation, a->alpha_logit, a->T_camera_world,
                                      a->projection, n, a->width, a->height, a->near_plane, a->far_plane, a->blur_cov,
                                      a->clamp_margin, c.alpha_threshold, a->ws_project, a->ws_project_bytes, a->points,
                                      a->depths, a->indexes, a->ndc, &aux->host_words[0], stream));
    if (a->use_sh) GS_TRY(gs_camera_position_f32(a->T_camera_world, a->camera_pos, stream));
    GS_TRY(wait_mapped_word(&aux->host_words[0], stream, "render_stage_a (V)", &v));
  } else {
    GS_TRY(gs_project_compact_f32(a->position, a->log_scaling, a->rotation, a->alpha_logit, a->T_camera_world,
                                  a->projection, n, a->width, a->height, a->near_plane, a->far_plane, a->blur_cov,
                                  a->clamp_margin, c.alpha_threshold, a->ws_project, a->ws_project_bytes, a->points,
                                  a->depths, a->indexes, a->ndc, &aux->host_words[0], stream));
    if (a->use_sh) GS_TRY(gs_camera_position_f32(a->T_camera_world, a->camera_pos, stream));
    GS_CUDA(cudaStreamSynchronize(stream));
    v = aux->host_words[0];
  }
  *v_out = v;

  // ---- auxiliary stream: [tile count,] features, zero fills, raster digest ----
  const int ts = c.tile_size;
  const int w_pad = pad_to(a->width, ts), h_pad = pad_to(a->height, ts);
  GS_CUDA(cudaEventRecord(aux->fence, stream));
  GS_CUDA(cudaStreamWaitEvent(side, aux->fence, 0));
  const bool by_point = hits_by_point(a);
  if (by_point) {
    GS_TRY(gs_tile_count_ordered_hits(a->points, nullptr, v, w_pad, h_pad, ts, c.alpha_threshold, a->tile_lo, a->tile_hi,
                                      a->counts, a->hits, side));
    GS_CUDA(cudaEventRecord(aux->count_done, side));
  }
  // the depth sort is enqueued before the bulk of the auxiliary work: its small kernels are then dispatched ahead of
  // the machine-filling feature / digest kernels instead of queueing behind them
  if (a->ordering != GS_ORDERING_BINNED)
    GS_TRY(gs_depth_order(a->ndc, v, a->use_depth16, a->order, a->ws_order, a->ws_order_bytes, stream));
  if (a->use_sh) {
    GS_TRY(gs_sh_fwd_f32(a->feature, a->position, a->indexes, a->camera_pos, v, a->channels, a->sh_degree,
                         a->features, side));
  } else if (v > 0) {
    const int64_t total = v * a->channels;
    gather_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, side>>>(a->feature, a->indexes, v, a->channels,
                                                                               a->features);
    GS_LAUNCH_CHECK();
  }
  if (a->visibility != nullptr && v > 0) GS_CUDA(cudaMemsetAsync(a->visibility, 0, sizeof(float) * v, side));
  if (a->heuristic != nullptr && v > 0) GS_CUDA(cudaMemsetAsync(a->heuristic, 0, sizeof(float) * 2 * v, side));
  GS_TRY(gs_raster_digest_f32(a->points, a->features, a->want_median ? a->depths : nullptr, v, a->channels, &a->config,
                              a->digest, side));
  GS_CUDA(cudaEventRecord(aux->side_done, side));

  if (a->ordering == GS_ORDERING_BINNED) {
    // ---- tile mapper, first half (binned ordering): per-tile counts -> tile ranges / slot cursors -> K ----
    const int64_t num_tiles = (int64_t)(w_pad / ts) * (h_pad / ts);
    GS_TRY(gs_tile_bin_count(a->points, v, w_pad, h_pad, ts, c.alpha_threshold, a->tile_counts, stream));
    GS_TRY(gs_tile_bin_offsets(a->tile_counts, num_tiles, a->tile_ranges, a->tile_cursor, a->tile_totals,
                               &aux->host_words[1], stream));
    GS_CUDA(cudaStreamSynchronize(stream));
    *k_out = (int64_t)aux->host_words[1];
    *max_per_tile_out = (int64_t)aux->host_words[2];
    return GS_OK;
  }
  GS_CHECK_ARG((a->tile_lo == 0 && a->tile_hi == 0) || a->hits != nullptr, "render_stage_a: a tile range needs the hit-record buffer");
  // ---- tile mapper, first half (two-level ordering): depth order (enqueued above) -> counts -> scan -> K ----
  if (by_point)
    GS_CUDA(cudaStreamWaitEvent(stream, aux->count_done, 0));
  else if (a->hits != nullptr)   // count and emit share one grid query through per-Gaussian hit records
    GS_TRY(gs_tile_count_ordered_hits(a->points, a->order, v, w_pad, h_pad, ts, c.alpha_threshold, a->tile_lo, a->tile_hi,
                                      a->counts, a->hits, stream));
  else
    GS_TRY(gs_tile_count_ordered(a->points, a->order, v, w_pad, h_pad, ts, c.alpha_threshold, a->counts, stream));
  if (mapped) {
    int64_t k = 0;
    aux->host_words[1] = kWordPending;
    GS_TRY(tile_scan_word(a->counts, v, a->cum, a->ws_scan, a->ws_scan_bytes, &aux->host_words[1], true,
                          by_point ? a->order : nullptr, stream));
    GS_TRY(wait_mapped_word(&aux->host_words[1], stream, "render_stage_a (K)", &k));
    *k_out = v > 0 ? k : 0;
    return GS_OK;
  }
  GS_TRY(tile_scan_word(a->counts, v, a->cum, a->ws_scan, a->ws_scan_bytes, &aux->host_words[1], false,
                        by_point ? a->order : nullptr, stream));
  GS_CUDA(cudaStreamSynchronize(stream));
  *k_out = v > 0 ? (int64_t)aux->host_words[1] : 0;
  return GS_OK;
}

extern "C" int gs_render_stage_b_f32(const gs_render_args *a, int64_t v, int64_t k, int64_t max_per_tile,
                                     int64_t k_stride, uint32_t *tiles, int32_t *o2p, void *ws_sort,
                                     size_t ws_sort_bytes, void *stream_) {
  using namespace gs;
  GS_CHECK_ARG(a != nullptr, "render_stage_b: NULL argument");
  GS_CHECK_ARG(render_supported(a->config, a->channels), "render_stage_b: unsupported raster configuration");
  GS_CHECK_ARG(k_stride >= k, "render_stage_b: k_stride %lld < k %lld", (long long)k_stride, (long long)k);
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceAux *aux = device_aux(stream);
  if (aux == nullptr) { set_error("render_stage_b: cannot create the auxiliary stream"); return GS_ERR_CUDA; }
  std::lock_guard<std::recursive_mutex> frame_lock(aux->frame_mu);
  GS_NVTX("gs_render_stage_b: key emit, tile sort, ranges, raster pack, raster forward");
  const int rc = stage_b_impl(a, v, k, max_per_tile, k_stride, tiles, o2p, ws_sort, ws_sort_bytes, stream, aux);
  if (rc != GS_OK) GS_CUDA(cudaStreamWaitEvent(stream, aux->side_done, 0));   // stage A's auxiliary work still holds caller buffers
  return rc;
}

static int gs::stage_b_impl(const gs_render_args *a, int64_t v, int64_t k, int64_t max_per_tile, int64_t k_stride,
                            uint32_t *tiles, int32_t *o2p, void *ws_sort, size_t ws_sort_bytes, cudaStream_t stream,
                            DeviceAux *aux) {
  const gs_raster_config &c = a->config;
  const int ts = c.tile_size;
  const int w_pad = pad_to(a->width, ts), h_pad = pad_to(a->height, ts);
  const int64_t num_tiles = (int64_t)(w_pad / ts) * (h_pad / ts);
  const int32_t *sorted_o2p = o2p + k_stride;
  const bool binned = a->ordering == GS_ORDERING_BINNED;
  bool two_level_sorted = false;
  if (binned && max_per_tile <= gs_tile_bin_max_per_tile()) {
    // binned ordering, second half: slot emission, then one shared-memory sort per tile.  `tiles` (2 k_stride u32)
    // holds the k 64-bit keys.
    if (k > 0) {
      GS_TRY(gs_tile_bin_emit(a->points, a->ndc, v, w_pad, h_pad, ts, c.alpha_threshold, a->use_depth16, a->tile_cursor,
                              reinterpret_cast<uint64_t *>(tiles), stream));
      GS_TRY(gs_tile_bin_sort(reinterpret_cast<const uint64_t *>(tiles), a->tile_ranges, num_tiles,
                              (int32_t)max_per_tile, o2p + k_stride, stream));
    }
  } else {
    // two-level ordering, second half (depth sort on V done in stage A, tile sort on K here); after a binned
    // stage A that met a tile too crowded for the shared-memory sort, its first half runs here too
    if (binned) {
      GS_TRY(gs_depth_order(a->ndc, v, a->use_depth16, a->order, a->ws_order, a->ws_order_bytes, stream));
      GS_TRY(gs_tile_count_ordered(a->points, a->order, v, w_pad, h_pad, ts, c.alpha_threshold, a->counts, stream));
      GS_TRY(gs_tile_scan(a->counts, v, a->cum, a->ws_scan, a->ws_scan_bytes, &aux->host_words[4], stream));
    }
    if (k > 0) {
      if (hits_by_point(a))
        GS_TRY(tile_emit_hits_by_point(a->points, a->order, a->cum, a->hits, v, w_pad, h_pad, ts, c.alpha_threshold,
                                       a->tile_lo, a->tile_hi, tiles, o2p, stream));
      else if (a->hits != nullptr && !binned)
        GS_TRY(gs_tile_emit_hits(a->points, a->order, a->cum, a->hits, v, w_pad, h_pad, ts, c.alpha_threshold, a->tile_lo,
                                 a->tile_hi, tiles, o2p, stream));
      else
        GS_TRY(gs_tile_emit_ordered(a->points, a->order, a->cum, v, w_pad, h_pad, ts, c.alpha_threshold, tiles, o2p,
                                    stream));
      GS_TRY(gs_sort_pairs(tiles, o2p, tiles + k_stride, o2p + k_stride, k, 4, 0, tile_bits(num_tiles), ws_sort,
                           ws_sort_bytes, stream));
    }
    // the tile ranges come out of the raster-pack pass below (same walk over the sorted tile ids); without packed
    // record buffers (or nothing to rasterise) they get their own kernel
    two_level_sorted = true;
    if (a->records == nullptr || k == 0)
      GS_TRY(gs_tile_ranges_from_tiles(tiles + k_stride, k, a->tile_ranges, num_tiles, stream));
  }
  GS_CUDA(cudaStreamWaitEvent(stream, aux->side_done, 0));
  if (a->records != nullptr || k == 0) {
    // per-overlap records in sorted order (kept for the backward), then the bulk-copy staged forward kernel
    if (!two_level_sorted)   // binned ordering: no sorted tile-id array, ranges already known
      GS_TRY(gs_raster_pack_f32(a->digest, a->tile_ranges, sorted_o2p, k, a->width, a->height, a->channels, a->records,
                                a->flush_records, stream));
    else if (k > 0)
      GS_TRY(gs_raster_pack_sorted_f32(a->digest, tiles + k_stride, sorted_o2p, k, a->width, a->height, a->channels,
                                       a->records, a->flush_records, a->tile_ranges, stream));
    if (a->ev_raster_start != nullptr) GS_CUDA(cudaEventRecord((cudaEvent_t)a->ev_raster_start, stream));
    GS_TRY(gs_raster_fwd_packed_f32(a->records, a->tile_ranges, sorted_o2p, v, k, a->width, a->height, a->channels,
                                    &a->config, a->median_threshold, a->image, a->image_alpha, a->visibility,
                                    a->want_median ? a->median_image : nullptr, stream));
  } else {
    if (a->ev_raster_start != nullptr) GS_CUDA(cudaEventRecord((cudaEvent_t)a->ev_raster_start, stream));
    GS_TRY(gs_raster_fwd_digest_f32(a->digest, a->tile_ranges, sorted_o2p, v, k, a->width, a->height, a->channels,
                                    &a->config, a->median_threshold, a->image, a->image_alpha, a->visibility,
                                    a->want_median ? a->median_image : nullptr, stream));
  }
  if (a->ev_raster_end != nullptr) GS_CUDA(cudaEventRecord((cudaEvent_t)a->ev_raster_end, stream));
  return GS_OK;
}

extern "C" int gs_render_forward_f32(const gs_render_args *a, int64_t k_capacity, uint32_t *tiles, int32_t *o2p,
                                     void *ws_sort, size_t ws_sort_bytes, int64_t *v_out, int64_t *k_out,
                                     int64_t *max_per_tile_out, int32_t *stage_b_done, void *stream) {
  GS_CHECK_ARG(stage_b_done != nullptr, "render_forward: NULL argument");
  *stage_b_done = 0;
  GS_TRY(gs_render_stage_a_f32(a, v_out, k_out, max_per_tile_out, stream));
  size_t need = 0;
  GS_TRY(gs_sort_pairs_workspace_bytes(*k_out, 4, &need));
  if (*k_out > k_capacity || (*k_out > 0 && (tiles == nullptr || o2p == nullptr || ws_sort_bytes < need)))
    return GS_OK;   // the caller allocates K-sized buffers and runs stage B itself
  GS_TRY(gs_render_stage_b_f32(a, *v_out, *k_out, *max_per_tile_out, k_capacity, tiles, o2p, ws_sort, ws_sort_bytes,
                               stream));
  *stage_b_done = 1;
  return GS_OK;
}

extern "C" int gs_render_backward_f32(const gs_render_bwd_args *a, void *stream_) {
  using namespace gs;
  GS_CHECK_ARG(a != nullptr, "render_backward: NULL argument");
  GS_CHECK_ARG(render_supported(a->config, a->channels), "render_backward: unsupported raster configuration");
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceAux *aux = device_aux(stream);
  if (aux == nullptr) { set_error("render_backward: cannot create the auxiliary stream"); return GS_ERR_CUDA; }
  std::lock_guard<std::recursive_mutex> frame_lock(aux->frame_mu);   // one frame at a time per (device, stream): events / side stream are shared
  GS_NVTX("gs_render_backward: raster backward, SH backward, projection backward");
  cudaStream_t side = use_side_stream() ? aux->side_stream : stream;
  const int64_t n = a->n, v = a->v;
  const int F = a->channels;
  const int D = a->use_sh ? (a->sh_degree + 1) * (a->sh_degree + 1) : 1;
  const int phases = a->phases == 0 ? (GS_BWD_RASTER | GS_BWD_FEATURE | GS_BWD_PROJECT) : a->phases;
  const bool need_geom = a->d_position || a->d_log_scaling || a->d_rotation || a->d_alpha_logit ||
                         a->d_T_camera_world || a->d_projection;
  SideJoin join(stream, side, aux->bwd_join);   // every exit path (errors included) joins the auxiliary stream

  if (phases & GS_BWD_RASTER) {
    // ---- auxiliary stream: zero fills of the dense parameter gradients, beside the raster backward ----
    GS_CUDA(cudaEventRecord(aux->fence, stream));
    GS_CUDA(cudaStreamWaitEvent(side, aux->fence, 0));
    if (a->d_position) GS_CUDA(cudaMemsetAsync(a->d_position, 0, sizeof(float) * 3 * n, side));
    if (a->d_log_scaling) GS_CUDA(cudaMemsetAsync(a->d_log_scaling, 0, sizeof(float) * 3 * n, side));
    if (a->d_rotation) GS_CUDA(cudaMemsetAsync(a->d_rotation, 0, sizeof(float) * 4 * n, side));
    if (a->d_alpha_logit) GS_CUDA(cudaMemsetAsync(a->d_alpha_logit, 0, sizeof(float) * n, side));
    if (a->d_T_camera_world) GS_CUDA(cudaMemsetAsync(a->d_T_camera_world, 0, sizeof(float) * 16, side));
    if (a->d_projection) GS_CUDA(cudaMemsetAsync(a->d_projection, 0, sizeof(float) * 4, side));
    if (a->d_camera_pos) GS_CUDA(cudaMemsetAsync(a->d_camera_pos, 0, sizeof(float) * 3, side));
    // the SH backward stores whole coefficient rows of every visible Gaussian: nothing to clear when all are visible
    if (a->d_feature && !(a->use_sh && v == n))
      GS_CUDA(cudaMemsetAsync(a->d_feature, 0, sizeof(float) * n * F * D, side));
    GS_CUDA(cudaEventRecord(aux->fills_done, side));

    // ---- raster backward ----
    if (v > 0) {
      if (a->grad_points && !a->grad_points_preset) GS_CUDA(cudaMemsetAsync(a->grad_points, 0, sizeof(float) * 7 * v, stream));
      if (a->grad_features && !a->grad_features_preset) GS_CUDA(cudaMemsetAsync(a->grad_features, 0, sizeof(float) * F * v, stream));
      if (a->d_image != nullptr) {
        if (a->ev_raster_start != nullptr) GS_CUDA(cudaEventRecord((cudaEvent_t)a->ev_raster_start, stream));
        float *gp = need_geom ? a->grad_points : nullptr, *gf = a->grad_features;
        if (a->records != nullptr && a->flush_records != nullptr)
          GS_TRY(gs_raster_bwd_packed_f32(a->records, a->flush_records, a->tile_ranges, a->overlap_to_point, a->image,
                                          a->d_image, a->d_image_strided ? a->d_image_strides : nullptr, v, a->k,
                                          a->width, a->height, F, &a->config, gp, gf, a->heuristic, stream));
        else if (a->d_image_strided)
          GS_TRY(gs_raster_bwd_digest_strided_f32(a->digest, a->tile_ranges, a->overlap_to_point, a->image, a->d_image,
                                                  a->d_image_strides, v, a->k, a->width, a->height, F, &a->config, gp, gf,
                                                  a->heuristic, stream));
        else
          GS_TRY(gs_raster_bwd_digest_f32(a->digest, a->tile_ranges, a->overlap_to_point, a->image, a->d_image, v, a->k,
                                          a->width, a->height, F, &a->config, gp, gf, a->heuristic, stream));
        if (a->ev_raster_end != nullptr) GS_CUDA(cudaEventRecord((cudaEvent_t)a->ev_raster_end, stream));
      }
    }
    // the fills must have landed before anything later on `stream` (this call or the next phase) accumulates
    GS_CUDA(cudaStreamWaitEvent(stream, aux->fills_done, 0));
  }

  if (v > 0) {
    // ---- auxiliary stream: feature gradient (SH backward / row scatter) ----
    if ((phases & GS_BWD_FEATURE) && a->d_feature) {
      GS_CUDA(cudaEventRecord(aux->raster_done, stream));
      GS_CUDA(cudaStreamWaitEvent(side, aux->raster_done, 0));
      if (a->use_sh) {
        GS_TRY(gs_sh_bwd_f32(a->feature, a->position, a->indexes, a->camera_pos, a->grad_features, a->features, v, F,
                             a->sh_degree, 1, a->d_feature, nullptr, a->d_camera_pos, side));
      } else {
        const int64_t total = v * F;
        scatter_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, side>>>(a->grad_features, a->indexes, v, F,
                                                                                    a->d_feature);
        GS_LAUNCH_CHECK();
      }
    }

    // ---- caller's stream: projection backward ----
    if ((phases & GS_BWD_PROJECT) && need_geom) {
      const float *dd = a->d_depths;
      if (dd == nullptr) {   // no incoming depth gradient: a zeroed column from library scratch
        float *z = (float *)stream_workspace(stream, sizeof(float) * v);
        if (z == nullptr) return GS_ERR_CUDA;
        GS_CUDA(cudaMemsetAsync(z, 0, sizeof(float) * v, stream));
        dd = z;
      }
      GS_TRY(gs_project_bwd_f32(a->position, a->log_scaling, a->rotation, a->alpha_logit, a->T_camera_world,
                                a->projection, a->indexes, v, a->width, a->height, a->blur_cov, a->clamp_margin,
                                a->grad_points, dd, a->d_position, a->d_log_scaling, a->d_rotation, a->d_alpha_logit,
                                a->d_T_camera_world, a->d_projection, stream));
    }
  }
  return GS_OK;   // ~SideJoin: the caller's stream waits for the auxiliary stream
}
