// common.cuh -- shared helpers for libgsplat_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gsplat_b200.h"

namespace gs {

void set_error(const char *fmt, ...);
void *stream_workspace(cudaStream_t stream, size_t bytes);   // api.cu: library-owned scratch per (device, stream)

// How a count (V, K) reaches the host.  Public entry points: cudaMemcpyAsync into the caller's host word, valid after a
// stream synchronisation.  Whole-frame drivers (render.cu): the words are MAPPED pinned memory, the producing kernel
// stores the count there itself and the host polls for it -- no copy-engine operation on the stream and no blocking
// synchronisation (together ~25 us of idle GPU per count in the step timeline, profiles/r02/r02ac_timeline.txt).
constexpr int32_t kWordPending = INT32_MIN;   // the host writes this before the producer is enqueued
int project_compact_f32_mapped(const float *position, const float *log_scaling, const float *rotation,
                               const float *alpha_logit, const float *T, const float *proj, int64_t n, int32_t width,
                               int32_t height, double near_plane, double far_plane, double blur_cov, double clamp_margin,
                               double alpha_threshold, void *workspace, size_t workspace_bytes, float *points,
                               float *depth, int64_t *indexes, float *ndc, int32_t *mapped_word, cudaStream_t stream);
// gs_tile_scan for the drivers.  word_is_mapped: the total is stored into the mapped word by the device (else copied
// into it, valid after a stream synchronisation).  order != NULL: the counts sit at the Gaussians' own indices and
// are scanned in depth order (counts[order[i]]).
int tile_scan_word(const int32_t *counts, int64_t v, int32_t *cum, void *workspace, size_t workspace_bytes,
                   int32_t *word, bool word_is_mapped, const int32_t *order, cudaStream_t stream);
// gs_tile_emit_hits with the hit records indexed by Gaussian instead of by depth rank
int tile_emit_hits_by_point(const float *gaussians, const int32_t *order, const int32_t *cum, const void *hits,
                            int64_t v, int32_t w_pad, int32_t h_pad, int32_t ts, double alpha_threshold,
                            int32_t tile_lo, int32_t tile_hi, uint32_t *tile_keys, int32_t *overlap_to_point,
                            cudaStream_t stream);

#define GS_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      gs::set_error(__VA_ARGS__);               \
      return GS_ERR_INVALID_ARGUMENT;           \
    }                                           \
  } while (0)

#define GS_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      gs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                    __LINE__);                                                         \
      return GS_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define GS_LAUNCH_CHECK()                                                              \
  do {                                                                                 \
    cudaError_t _e = cudaPeekAtLastError();                                            \
    if (_e != cudaSuccess) {                                                           \
      gs::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                    __FILE__, __LINE__);                                               \
      return GS_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

// NVTX ranges around the phases of the whole-frame drivers and the per-stage entry points (header-only NVTX v3: no
// library to link; a no-op unless a profiler -- nsys, ncu --nvtx -- is attached).  The reference wraps its Taichi
// launches in torch.profiler record_function ranges.
struct NvtxRange {
  explicit NvtxRange(const char *name);
  ~NvtxRange();
};
#define GS_NVTX(name) gs::NvtxRange _gs_nvtx_range(name)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename real> struct math;
template <> struct math<float> {
  static __device__ __forceinline__ float exp(float x) { return ::expf(x); }
  static __device__ __forceinline__ float fast_exp(float x) { return ::__expf(x); }
  static __device__ __forceinline__ float log(float x) { return ::logf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return ::sqrtf(x); }
  static __device__ __forceinline__ float abs(float x) { return ::fabsf(x); }
  static __device__ __forceinline__ float min(float a, float b) { return ::fminf(a, b); }
  static __device__ __forceinline__ float max(float a, float b) { return ::fmaxf(a, b); }
};
template <> struct math<double> {
  static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double fast_exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  static __device__ __forceinline__ double min(double a, double b) { return ::fmin(a, b); }
  static __device__ __forceinline__ double max(double a, double b) { return ::fmax(a, b); }
};

// Warp covers 8x4 pixels (same decomposition as rasterizer/tiling.py:34-51 with stride 1).
__device__ __forceinline__ void tile_pixel(int tid, int tile_size, int &u, int &v) {
  int warp_id = tid >> 5, lane = tid & 31;
  int warps_wide = tile_size >> 3;
  u = (warp_id % warps_wide) * 8 + (lane & 7);
  v = (warp_id / warps_wide) * 4 + (lane >> 3);
}

}  // namespace gs
