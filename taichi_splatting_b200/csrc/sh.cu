// sh.cu -- R2: real spherical harmonics (degree 0..3) evaluated at gathered indexes.
//
// Semantics: evaluate_sh_at_kernel (+ Taichi autodiff .grad), indexed_spherical_harmonics.py:118-160;
// basis constants :38-106.  out[i,c] = clamp(sum_d Y_d(normalize(p[idx]-cam)) * params[idx,c,d] + 0.5, 0, 1).
// Layout here: one thread per (visible point, channel) so the (M,C,D) coefficient rows are read and the
// (V,C) colours written fully coalesced; the backward is hand-derived.
#include "common.cuh"

namespace gs {

template <typename real, int DEG>
__device__ __forceinline__ void rsh(real x, real y, real z, real *Y) {
  Y[0] = real(0.282094791773878);
  if constexpr (DEG >= 1) {
    Y[1] = real(-0.48860251190292) * y;
    Y[2] = real(0.48860251190292) * z;
    Y[3] = real(-0.48860251190292) * x;
  }
  if constexpr (DEG >= 2) {
    real x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
    Y[4] = real(1.09254843059208) * xy;
    Y[5] = real(-1.09254843059208) * yz;
    Y[6] = real(0.94617469575756) * z2 - real(0.31539156525252);
    Y[7] = real(-1.09254843059208) * xz;
    Y[8] = real(0.54627421529604) * x2 - real(0.54627421529604) * y2;
    if constexpr (DEG >= 3) {
      Y[9] = real(-0.590043589926644) * y * (real(3.0) * x2 - y2);
      Y[10] = real(2.89061144264055) * xy * z;
      Y[11] = real(0.304697199642977) * y * (real(1.5) - real(7.5) * z2);
      Y[12] = real(1.24392110863372) * z * (real(1.5) * z2 - real(0.5)) - real(0.497568443453487) * z;
      Y[13] = real(0.304697199642977) * x * (real(1.5) - real(7.5) * z2);
      Y[14] = real(1.44530572132028) * z * (x2 - y2);
      Y[15] = real(-0.590043589926644) * x * (x2 - real(3.0) * y2);
    }
  }
}

// d(dir) = sum_d g[d] * dY_d/d(dir)
template <typename real, int DEG>
__device__ __forceinline__ void rsh_vjp(real x, real y, real z, const real *g, real *d) {
  d[0] = d[1] = d[2] = 0;
  if constexpr (DEG >= 1) {
    const real c1 = real(0.48860251190292);
    d[1] += -c1 * g[1];
    d[2] += c1 * g[2];
    d[0] += -c1 * g[3];
  }
  if constexpr (DEG >= 2) {
    const real c2 = real(1.09254843059208), c3 = real(0.94617469575756), c5 = real(0.54627421529604);
    d[0] += c2 * y * g[4]; d[1] += c2 * x * g[4];
    d[1] += -c2 * z * g[5]; d[2] += -c2 * y * g[5];
    d[2] += 2 * c3 * z * g[6];
    d[0] += -c2 * z * g[7]; d[2] += -c2 * x * g[7];
    d[0] += 2 * c5 * x * g[8]; d[1] += -2 * c5 * y * g[8];
  }
  if constexpr (DEG >= 3) {
    const real c6 = real(0.590043589926644), c7 = real(2.89061144264055), c8 = real(0.304697199642977),
               c9 = real(1.24392110863372), c10 = real(0.497568443453487), c11 = real(1.44530572132028);
    real x2 = x * x, y2 = y * y, z2 = z * z;
    d[0] += -6 * c6 * x * y * g[9]; d[1] += -c6 * (3 * x2 - 3 * y2) * g[9];
    d[0] += c7 * y * z * g[10]; d[1] += c7 * x * z * g[10]; d[2] += c7 * x * y * g[10];
    d[1] += c8 * (real(1.5) - real(7.5) * z2) * g[11]; d[2] += -15 * c8 * y * z * g[11];
    d[2] += (c9 * (real(4.5) * z2 - real(0.5)) - c10) * g[12];
    d[0] += c8 * (real(1.5) - real(7.5) * z2) * g[13]; d[2] += -15 * c8 * x * z * g[13];
    d[0] += 2 * c11 * x * z * g[14]; d[1] += -2 * c11 * y * z * g[14]; d[2] += c11 * (x2 - y2) * g[14];
    d[0] += -c6 * (3 * x2 - 3 * y2) * g[15]; d[1] += 6 * c6 * x * y * g[15];
  }
}

// Coefficient rows are D contiguous values; for fp32 with D % 4 == 0 (degree 1 and 3) they are 16-byte aligned
// and move as float4, which quarters the LSU instruction count of these HBM-bound kernels.
template <typename real, int D>
__device__ __forceinline__ void load_row(const real *__restrict__ p, real *r) {
  if constexpr (sizeof(real) == 4 && D % 4 == 0) {
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      float4 t = __ldg(p4 + q);
      r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = p[d];
  }
}

template <typename real, int D>
__device__ __forceinline__ void store_row(real *__restrict__ p, const real *r) {
  if constexpr (sizeof(real) == 4 && D % 4 == 0) {
    float4 *p4 = reinterpret_cast<float4 *>(p);
#pragma unroll
    for (int q = 0; q < D / 4; ++q) p4[q] = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) p[d] = r[d];
  }
}

template <typename real, int DEG>
__global__ void __launch_bounds__(256)
sh_fwd_kernel(const real *__restrict__ params, const real *__restrict__ positions,
              const int64_t *__restrict__ indexes, const real *__restrict__ camera_pos, int64_t total,
              int channels, real *__restrict__ out) {
  constexpr int D = (DEG + 1) * (DEG + 1);
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  int64_t i = t / channels;
  int c = (int)(t - i * channels);
  int64_t idx = indexes[i];
  real vx = positions[3 * idx] - camera_pos[0], vy = positions[3 * idx + 1] - camera_pos[1],
       vz = positions[3 * idx + 2] - camera_pos[2];
  real inv = real(1) / math<real>::sqrt(vx * vx + vy * vy + vz * vz);
  real Y[D];
  rsh<real, DEG>(vx * inv, vy * inv, vz * inv, Y);
  real p[D];
  load_row<real, D>(params + (idx * channels + c) * D, p);
  real acc = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) acc += Y[d] * p[d];
  acc += real(0.5);
  out[t] = math<real>::min(math<real>::max(acc, real(0)), real(1));
}

template <typename real, int DEG>
__global__ void __launch_bounds__(256)
sh_bwd_kernel(const real *__restrict__ params, const real *__restrict__ positions,
              const int64_t *__restrict__ indexes, const real *__restrict__ camera_pos,
              const real *__restrict__ d_out, const real *__restrict__ out, int64_t total, int channels,
              int unique, real *__restrict__ d_params, real *__restrict__ d_positions, real *__restrict__ d_camera_pos) {
  constexpr int D = (DEG + 1) * (DEG + 1);
  __shared__ real red[3];
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  real dcam[3] = {0, 0, 0};
  if (t < total) {
    int64_t i = t / channels;
    int c = (int)(t - i * channels);
    int64_t idx = indexes[i];
    real vx = positions[3 * idx] - camera_pos[0], vy = positions[3 * idx + 1] - camera_pos[1],
         vz = positions[3 * idx + 2] - camera_pos[2];
    real inv = real(1) / math<real>::sqrt(vx * vx + vy * vy + vz * vz);
    real x = vx * inv, y = vy * inv, z = vz * inv;
    real Y[D];
    rsh<real, DEG>(x, y, z, Y);
    // clamp mask: from the saved forward output when the caller has it (no need to re-read the D coefficients of
    // this row: the clamp was active iff the output sits exactly on 0 or 1), else recomputed
    const bool need_params = (out == nullptr) || ((d_positions || d_camera_pos) && DEG >= 1);
    real p[D];
    real g;
    if (need_params) load_row<real, D>(params + (idx * channels + c) * D, p);
    if (out != nullptr) {
      real o = out[t];
      g = (o > real(0) && o < real(1)) ? d_out[t] : real(0);
    } else {
      real pre = 0;
#pragma unroll
      for (int d = 0; d < D; ++d) pre += Y[d] * p[d];
      pre += real(0.5);
      g = (pre >= real(0) && pre <= real(1)) ? d_out[t] : real(0);
    }
    if (d_params) {
      real *dp = d_params + (idx * channels + c) * D;
      if (unique) {
        real row[D];
#pragma unroll
        for (int d = 0; d < D; ++d) row[d] = Y[d] * g;
        store_row<real, D>(dp, row);
      } else if (g != real(0)) {
#pragma unroll
        for (int d = 0; d < D; ++d) atomicAdd(dp + d, Y[d] * g);
      }
    }
    if ((d_positions || d_camera_pos) && DEG >= 1 && g != real(0)) {
      real gy[D], dd[3];
#pragma unroll
      for (int d = 0; d < D; ++d) gy[d] = p[d] * g;
      rsh_vjp<real, DEG>(x, y, z, gy, dd);
      real dot = x * dd[0] + y * dd[1] + z * dd[2];
      real dv[3] = {(dd[0] - x * dot) * inv, (dd[1] - y * dot) * inv, (dd[2] - z * dot) * inv};
      if (d_positions)
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(d_positions + 3 * idx + k, dv[k]);
#pragma unroll
      for (int k = 0; k < 3; ++k) dcam[k] = -dv[k];
    }
  }
  if (d_camera_pos) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) dcam[k] += __shfl_xor_sync(full, dcam[k], off);
    if (threadIdx.x < 3) red[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (dcam[k] != real(0)) atomicAdd(&red[k], dcam[k]);
    __syncthreads();
    if (threadIdx.x < 3 && red[threadIdx.x] != real(0)) atomicAdd(d_camera_pos + threadIdx.x, red[threadIdx.x]);
  }
}

// View-parallel gradient exchange.  The SH coefficient gradient of one view is rank-1 per Gaussian and channel:
// d_params[i, c, :] = Y(dir_view(i)) * g_view[i, c]  (g = dL/dcolour, zero where the clamp was active or the
// Gaussian was culled).  Summing it over W views therefore needs only the W (N, C) arrays g_w and the W camera
// centres -- 12 B per Gaussian and view instead of an all-reduce of the 192 B (N, C, 16) gradient.  This kernel
// rebuilds the sum: one thread per (Gaussian, channel), Y recomputed per view, one coalesced row store.
template <typename real, int DEG>
__global__ void __launch_bounds__(256)
sh_bwd_views_kernel(const real *__restrict__ positions, const real *__restrict__ cam_positions,
                    const real *__restrict__ g_all, int64_t n, int views, int channels, int64_t view_stride,
                    real *__restrict__ d_params) {
  constexpr int D = (DEG + 1) * (DEG + 1);
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n * channels) return;
  int64_t i = t / channels;
  real px = positions[3 * i], py = positions[3 * i + 1], pz = positions[3 * i + 2];
  real acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0;
  for (int w = 0; w < views; ++w) {
    real g = g_all[w * view_stride + t];
    if (g == real(0)) continue;
    real vx = px - cam_positions[3 * w], vy = py - cam_positions[3 * w + 1], vz = pz - cam_positions[3 * w + 2];
    real inv = real(1) / math<real>::sqrt(vx * vx + vy * vy + vz * vz);
    real Y[D];
    rsh<real, DEG>(vx * inv, vy * inv, vz * inv, Y);
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] += Y[d] * g;
  }
  store_row<real, D>(d_params + t * D, acc);
}

// Three channels (the render path): one thread per Gaussian, so the direction and the basis Y are evaluated once per
// view instead of once per (view, channel) -- the kernel is bound by that arithmetic, not by its 12 B read per view.
template <typename real, int DEG>
__global__ void __launch_bounds__(128)
sh_bwd_views_rgb_kernel(const real *__restrict__ positions, const real *__restrict__ cam_positions,
                        const real *__restrict__ g_all, int64_t n, int views, int64_t view_stride,
                        real *__restrict__ d_params) {
  constexpr int D = (DEG + 1) * (DEG + 1);
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const real px = positions[3 * i], py = positions[3 * i + 1], pz = positions[3 * i + 2];
  real acc[3][D];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < D; ++d) acc[c][d] = 0;
  for (int w = 0; w < views; ++w) {
    const real *g = g_all + w * view_stride + 3 * i;
    const real g0 = g[0], g1 = g[1], g2 = g[2];
    if (g0 == real(0) && g1 == real(0) && g2 == real(0)) continue;
    const real vx = px - cam_positions[3 * w], vy = py - cam_positions[3 * w + 1], vz = pz - cam_positions[3 * w + 2];
    const real inv = real(1) / math<real>::sqrt(vx * vx + vy * vy + vz * vz);
    real Y[D];
    rsh<real, DEG>(vx * inv, vy * inv, vz * inv, Y);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      acc[0][d] += Y[d] * g0;
      acc[1][d] += Y[d] * g1;
      acc[2][d] += Y[d] * g2;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) store_row<real, D>(d_params + (3 * i + c) * D, acc[c]);
}

template <typename real>
int sh_bwd_views(const real *positions, const real *cam_positions, const real *g_all, int64_t n, int views,
                 int channels, int64_t view_stride, int degree, real *d_params, cudaStream_t stream) {
  GS_CHECK_ARG(degree >= 0 && degree <= 3, "sh: degree %d not in 0..3", degree);
  int64_t total = n * channels;
  if (total == 0) return GS_OK;
  if (channels == 3) {
    const unsigned grid3 = (unsigned)ceil_div(n, 128);
    switch (degree) {
      case 0: sh_bwd_views_rgb_kernel<real, 0><<<grid3, 128, 0, stream>>>(positions, cam_positions, g_all, n, views, view_stride, d_params); break;
      case 1: sh_bwd_views_rgb_kernel<real, 1><<<grid3, 128, 0, stream>>>(positions, cam_positions, g_all, n, views, view_stride, d_params); break;
      case 2: sh_bwd_views_rgb_kernel<real, 2><<<grid3, 128, 0, stream>>>(positions, cam_positions, g_all, n, views, view_stride, d_params); break;
      default: sh_bwd_views_rgb_kernel<real, 3><<<grid3, 128, 0, stream>>>(positions, cam_positions, g_all, n, views, view_stride, d_params); break;
    }
    GS_LAUNCH_CHECK();
    return GS_OK;
  }
  unsigned grid = (unsigned)ceil_div(total, 256);
#define GS_SH_VIEWS(DEG) \
  sh_bwd_views_kernel<real, DEG><<<grid, 256, 0, stream>>>(positions, cam_positions, g_all, n, views, channels, view_stride, d_params)
  switch (degree) {
    case 0: GS_SH_VIEWS(0); break;
    case 1: GS_SH_VIEWS(1); break;
    case 2: GS_SH_VIEWS(2); break;
    default: GS_SH_VIEWS(3); break;
  }
#undef GS_SH_VIEWS
  GS_LAUNCH_CHECK();
  return GS_OK;
}

template <typename real>
int sh_fwd(const real *params, const real *positions, const int64_t *indexes, const real *camera_pos, int64_t v,
           int channels, int degree, real *out, cudaStream_t stream) {
  GS_CHECK_ARG(degree >= 0 && degree <= 3, "sh: degree %d not in 0..3", degree);
  GS_CHECK_ARG(channels >= 1, "sh: channels must be >= 1");
  int64_t total = v * channels;
  if (total == 0) return GS_OK;
  unsigned grid = (unsigned)ceil_div(total, 256);
  switch (degree) {
    case 0: sh_fwd_kernel<real, 0><<<grid, 256, 0, stream>>>(params, positions, indexes, camera_pos, total, channels, out); break;
    case 1: sh_fwd_kernel<real, 1><<<grid, 256, 0, stream>>>(params, positions, indexes, camera_pos, total, channels, out); break;
    case 2: sh_fwd_kernel<real, 2><<<grid, 256, 0, stream>>>(params, positions, indexes, camera_pos, total, channels, out); break;
    default: sh_fwd_kernel<real, 3><<<grid, 256, 0, stream>>>(params, positions, indexes, camera_pos, total, channels, out); break;
  }
  GS_LAUNCH_CHECK();
  return GS_OK;
}

template <typename real>
int sh_bwd(const real *params, const real *positions, const int64_t *indexes, const real *camera_pos,
           const real *d_out, const real *out, int64_t v, int channels, int degree, int unique, real *d_params,
           real *d_positions, real *d_camera_pos, cudaStream_t stream) {
  GS_CHECK_ARG(degree >= 0 && degree <= 3, "sh: degree %d not in 0..3", degree);
  int64_t total = v * channels;
  if (total == 0) return GS_OK;
  unsigned grid = (unsigned)ceil_div(total, 256);
#define GS_SH_BWD(DEG)                                                                                        \
  sh_bwd_kernel<real, DEG><<<grid, 256, 0, stream>>>(params, positions, indexes, camera_pos, d_out, out, total, \
                                                     channels, unique, d_params, d_positions, d_camera_pos)
  switch (degree) {
    case 0: GS_SH_BWD(0); break;
    case 1: GS_SH_BWD(1); break;
    case 2: GS_SH_BWD(2); break;
    default: GS_SH_BWD(3); break;
  }
#undef GS_SH_BWD
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // namespace gs

#define GS_SH_API(SUFFIX, real)                                                                                 \
  extern "C" int gs_sh_fwd_##SUFFIX(const real *params, const real *positions, const int64_t *indexes,          \
                                    const real *camera_pos, int64_t v, int32_t channels, int32_t degree,        \
                                    real *out, void *stream) {                                                  \
    return gs::sh_fwd<real>(params, positions, indexes, camera_pos, v, channels, degree, out,                   \
                            (cudaStream_t)stream);                                                              \
  }                                                                                                             \
  extern "C" int gs_sh_bwd_##SUFFIX(const real *params, const real *positions, const int64_t *indexes,          \
                                    const real *camera_pos, const real *d_out, const real *out, int64_t v,      \
                                    int32_t channels, int32_t degree, int32_t unique_indexes, real *d_params,   \
                                    real *d_positions, real *d_camera_pos, void *stream) {                      \
    return gs::sh_bwd<real>(params, positions, indexes, camera_pos, d_out, out, v, channels, degree,            \
                            unique_indexes, d_params, d_positions, d_camera_pos, (cudaStream_t)stream);         \
  }

namespace gs {
// This rank's factors for the view-parallel exchange, in one launch: out[idx * channels + c] = dL/dcolour[i, c] where
// the colour is inside the clamp (0 < colour < 1; 0 otherwise, like the SH backward), rows of culled Gaussians zero,
// camera centre in the last three words.  Replaces zeros + where + index_copy_ + slice assignment (four torch kernels).
__global__ void __launch_bounds__(256)
sh_pack_factors_kernel(const float *__restrict__ colours, const float *__restrict__ d_colours,
                       const int64_t *__restrict__ indexes, const float *__restrict__ camera_pos, int64_t v,
                       int channels, int64_t n, float *__restrict__ out) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < 3) out[n * channels + t] = camera_pos[t];
  if (t >= v * channels) return;
  const int64_t i = t / channels;
  const float col = colours[t];
  out[indexes[i] * channels + (t - i * channels)] = (col > 0.f && col < 1.f) ? d_colours[t] : 0.f;
}
}  // namespace gs

extern "C" int gs_sh_pack_factors_f32(const float *colours, const float *d_colours, const int64_t *indexes,
                                      const float *camera_pos, int64_t v, int32_t channels, int64_t n, float *out,
                                      void *stream_) {
  GS_CHECK_ARG(out != nullptr && camera_pos != nullptr && channels >= 1 && v >= 0 && v <= n, "sh_pack_factors: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (v < n) GS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n * channels, stream));   // rows of culled Gaussians
  const int64_t total = v * channels > 3 ? v * channels : 3;
  gs::sh_pack_factors_kernel<<<(unsigned)gs::ceil_div(total, 256), 256, 0, stream>>>(colours, d_colours, indexes, camera_pos,
                                                                                      v, channels, n, out);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

// ---- fused pack + all-gather over peer memory ------------------------------------------------------------------
// The view-parallel exchange needs every rank's factors on every rank.  Packing into a local buffer and calling an
// NCCL all-gather costs 0.29 ms at 8 ranks (96 MB, rank-0 timeline in profiles/r02/) with nothing to hide behind; here
// the pack kernel itself is the collective: each thread stores its value straight into slot `rank` of the gathered
// buffer of EVERY rank (peer buffers mapped by symmetric-memory rendezvous; the stores to the seven remote buffers
// travel over NVLink as plain st.global and are fire-and-forget), so the transfer runs while the kernel still computes
// and costs the NVLink egress time of 7 x 12 MB.  A symmetric-memory barrier (caller) orders it before the readers.
namespace gs {
constexpr int kMaxPeers = 16;
struct PeerTable {
  float *base[kMaxPeers];
  int world;
};

__global__ void __launch_bounds__(256)
sh_pack_factors_peers_kernel(const float *__restrict__ colours, const float *__restrict__ d_colours,
                             const int64_t *__restrict__ indexes, const float *__restrict__ camera_pos, int64_t v,
                             int channels, int64_t n, const __grid_constant__ PeerTable peers, int64_t slot_offset) {
  // grid-stride over a SMALL grid: the kernel is bound by NVLink egress, which a few SMs' store pipes saturate, and it
  // runs beside the projection backward, which needs the other SMs
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t0 < 3) {
    const float c = camera_pos[t0];
    for (int w = 0; w < peers.world; ++w) peers.base[w][slot_offset + n * channels + t0] = c;
  }
  for (int64_t t = t0; t < v * channels; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / channels;
    const float col = colours[t];
    const float g = (col > 0.f && col < 1.f) ? d_colours[t] : 0.f;
    const int64_t at = slot_offset + indexes[i] * channels + (t - i * channels);
    for (int w = 0; w < peers.world; ++w) peers.base[w][at] = g;
  }
}

__global__ void __launch_bounds__(256)
fill_peers_kernel(const __grid_constant__ PeerTable peers, int64_t slot_offset, int64_t count) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= count) return;
  for (int w = 0; w < peers.world; ++w) peers.base[w][slot_offset + t] = 0.f;
}
}  // namespace gs

extern "C" int gs_sh_pack_factors_peers_f32(const float *colours, const float *d_colours, const int64_t *indexes,
                                            const float *camera_pos, int64_t v, int32_t channels, int64_t n,
                                            const uint64_t *peer_bases_host, int32_t world, int64_t slot_offset,
                                            void *stream_) {
  GS_CHECK_ARG(peer_bases_host != nullptr && world >= 1 && world <= gs::kMaxPeers, "sh_pack_factors_peers: 1..%d peers, got %d", gs::kMaxPeers, world);
  GS_CHECK_ARG(camera_pos != nullptr && channels >= 1 && v >= 0 && v <= n, "sh_pack_factors_peers: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  gs::PeerTable peers;
  peers.world = world;
  for (int w = 0; w < gs::kMaxPeers; ++w) peers.base[w] = w < world ? reinterpret_cast<float *>(peer_bases_host[w]) : nullptr;
  if (v < n) {   // rows of culled Gaussians are zero in every peer's copy
    const int64_t count = n * channels;
    gs::fill_peers_kernel<<<(unsigned)gs::ceil_div(count, 256), 256, 0, stream>>>(peers, slot_offset, count);
    GS_LAUNCH_CHECK();
  }
  const int64_t total = v * channels > 3 ? v * channels : 3;
  const int64_t blocks = gs::ceil_div(total, 256);
  gs::sh_pack_factors_peers_kernel<<<(unsigned)(blocks < 96 ? blocks : 96), 256, 0, stream>>>(
      colours, d_colours, indexes, camera_pos, v, channels, n, peers, slot_offset);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

// ---- all-reduce (sum) of a flat buffer over peer memory, in place ---------------------------------------------------
// Rank r owns elements [r L, (r + 1) L) of the buffer: it loads that range from EVERY rank's copy (seven remote loads
// per element over NVLink), adds, and stores the sum back into every rank's copy.  No element is touched by two ranks, so
// the operation is in place; the caller brackets it with symmetric-memory barriers (all inputs written / all outputs
// landed).  Traffic per rank: 7/8 of the buffer in and out, against 2 x 7/8 through an NCCL ring with its per-step
// latency; used for the 44 MB geometry-gradient buffer of the view-parallel backward.
namespace gs {
__global__ void __launch_bounds__(256)
allreduce_peers_kernel(const __grid_constant__ PeerTable peers, int64_t begin4, int64_t end4) {
  const int64_t j = begin4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // float4 index
  if (j >= end4) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int w = 0; w < peers.world; ++w) {
    const float4 x = reinterpret_cast<const float4 *>(peers.base[w])[j];
    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
  }
#pragma unroll 8
  for (int w = 0; w < peers.world; ++w) reinterpret_cast<float4 *>(peers.base[w])[j] = acc;
}
}  // namespace gs

extern "C" int gs_allreduce_peers_f32(const uint64_t *peer_bases_host, int32_t world, int32_t rank, int64_t count,
                                      void *stream_) {
  GS_CHECK_ARG(peer_bases_host != nullptr && world >= 1 && world <= gs::kMaxPeers && rank >= 0 && rank < world,
               "allreduce_peers: bad peer table");
  GS_CHECK_ARG(count >= 0 && count % 4 == 0, "allreduce_peers: count must be a multiple of 4 floats");
  gs::PeerTable peers;
  peers.world = world;
  for (int w = 0; w < gs::kMaxPeers; ++w) {
    peers.base[w] = w < world ? reinterpret_cast<float *>(peer_bases_host[w]) : nullptr;
    GS_CHECK_ARG(w >= world || (peer_bases_host[w] & 15) == 0, "allreduce_peers: buffers must be 16-byte aligned");
  }
  const int64_t total4 = count / 4, per = (total4 + world - 1) / world;
  const int64_t begin4 = per * rank, end4 = begin4 + per < total4 ? begin4 + per : total4;
  if (end4 <= begin4) return GS_OK;
  gs::allreduce_peers_kernel<<<(unsigned)gs::ceil_div(end4 - begin4, 256), 256, 0, (cudaStream_t)stream_>>>(peers, begin4, end4);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_sh_bwd_views_f32(const float *positions, const float *cam_positions, const float *g_all, int64_t n,
                                   int32_t views, int32_t channels, int64_t view_stride, int32_t degree,
                                   float *d_params, void *stream) {
  return gs::sh_bwd_views<float>(positions, cam_positions, g_all, n, views, channels, view_stride, degree, d_params,
                                 (cudaStream_t)stream);
}

GS_SH_API(f32, float)
GS_SH_API(f64, double)
