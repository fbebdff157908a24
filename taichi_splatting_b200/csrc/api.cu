// api.cu -- version + thread-local error string of the C ABI (include/gsplat_b200.h).
#include <stdarg.h>

#include <map>
#include <mutex>
#include <utility>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace gs {
NvtxRange::NvtxRange(const char *name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

static thread_local char g_error[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
}

// Library-owned scratch for the entry points that keep the reference's argument lists (no workspace parameter):
// one grow-only buffer per (device, stream), so that work enqueued on different streams never shares scratch.
// cudaMalloc only happens when a buffer has to grow.
void *stream_workspace(cudaStream_t stream, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, std::pair<void *, size_t>> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  auto &slot = cache[{dev, stream}];
  if (slot.second < bytes) {
    if (slot.first != nullptr) {
      cudaStreamSynchronize(stream);   // work already enqueued may still read the old buffer
      cudaFree(slot.first);
    }
    size_t want = bytes + bytes / 4;
    if (cudaMalloc(&slot.first, want) != cudaSuccess) {
      slot = {nullptr, 0};
      set_error("stream_workspace: cudaMalloc of %zu bytes failed", want);
      return nullptr;
    }
    slot.second = want;
  }
  return slot.first;
}
}  // namespace gs

extern "C" int gs_version(void) { return 120; }
extern "C" const char *gs_last_error_string(void) { return gs::g_error; }
