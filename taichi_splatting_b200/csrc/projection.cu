// projection.cu -- R1 / R1b / R11: 3D Gaussian -> 2D (mean, axis, sigma, alpha) projection.
//
// What it computes follows the reference's project_kernel / indexed_project_kernel
// (perspective/projection.py:32-119) and the math in taichi_lib/generic.py:96-158,217-237,419-427.
// The backward is a hand-written reverse chain (the reference relies on Taichi autodiff).
// How it is computed is this library's own: one thread per Gaussian, camera held in registers,
// in-view flags -> single scan -> order-preserving compaction, camera gradients block-reduced.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <iterator>

#include "common.cuh"

namespace gs {

template <typename real>
struct Camera {
  real R[3][3], t[3];
  real fx, fy, cx, cy;
};

template <typename real>
__device__ __forceinline__ Camera<real> load_camera(const real *__restrict__ T, const real *__restrict__ proj) {
  Camera<real> cam;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) cam.R[i][j] = T[4 * i + j];
    cam.t[i] = T[4 * i + 3];
  }
  cam.fx = proj[0]; cam.fy = proj[1]; cam.cx = proj[2]; cam.cy = proj[3];
  return cam;
}

template <typename real>
struct Projected {
  real cam_xyz[3];     // point in camera frame
  real uv[2], tc[2];   // image position, clamped position used by the Jacobian
  bool inside[2];      // clamp inactive (gradient passes)
  real J00, J02, J11, J12;
  real qh[4], qnorm, s[3], Rq[3][3];
  real A[3][3], M[2][3];
  real a, b, c, tr, gap, sg, sigma[2], u[2], unorm, v1[2], alpha, sgn;
  bool swapped;           // eigenvector taken from (b, c - l2) instead of (a - l2, b)
};

template <typename real>
__device__ __forceinline__ void project_one(const Camera<real> &cam, const real *__restrict__ p,
                                            const real *__restrict__ ls, const real *__restrict__ q,
                                            real logit, real width, real height, real blur, real margin,
                                            Projected<real> &o) {
  using m = math<real>;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o.cam_xyz[i] = cam.R[i][0] * p[0] + cam.R[i][1] * p[1] + cam.R[i][2] * p[2] + cam.t[i];
  real z = o.cam_xyz[2];
  o.uv[0] = (cam.fx * o.cam_xyz[0]) / z + cam.cx;
  o.uv[1] = (cam.fy * o.cam_xyz[1]) / z + cam.cy;
  real lo[2] = {-width * margin, -height * margin};
  real hi[2] = {(width - 1) * (1 + margin), (height - 1) * (1 + margin)};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    o.tc[k] = m::min(m::max(o.uv[k], lo[k]), hi[k]);
    o.inside[k] = o.uv[k] >= lo[k] && o.uv[k] <= hi[k];
  }
  o.J00 = cam.fx / z; o.J02 = -(o.tc[0] - cam.cx) / z;
  o.J11 = cam.fy / z; o.J12 = -(o.tc[1] - cam.cy) / z;

  o.qnorm = m::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) o.qh[k] = q[k] / o.qnorm;
#pragma unroll
  for (int k = 0; k < 3; ++k) o.s[k] = m::exp(ls[k]);
  real x = o.qh[0], y = o.qh[1], zq = o.qh[2], w = o.qh[3];
  real x2 = x * x, y2 = y * y, z2 = zq * zq;
  o.Rq[0][0] = 1 - 2 * y2 - 2 * z2; o.Rq[0][1] = 2 * x * y - 2 * w * zq; o.Rq[0][2] = 2 * x * zq + 2 * w * y;
  o.Rq[1][0] = 2 * x * y + 2 * w * zq; o.Rq[1][1] = 1 - 2 * x2 - 2 * z2; o.Rq[1][2] = 2 * y * zq - 2 * w * x;
  o.Rq[2][0] = 2 * x * zq - 2 * w * y; o.Rq[2][1] = 2 * y * zq + 2 * w * x; o.Rq[2][2] = 1 - 2 * x2 - 2 * y2;
  // A = W * (Rq * diag(s))
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o.A[i][j] = (cam.R[i][0] * o.Rq[0][j] + cam.R[i][1] * o.Rq[1][j] + cam.R[i][2] * o.Rq[2][j]) * o.s[j];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    o.M[0][j] = o.J00 * o.A[0][j] + o.J02 * o.A[2][j];
    o.M[1][j] = o.J11 * o.A[1][j] + o.J12 * o.A[2][j];
  }
  real a0 = o.M[0][0] * o.M[0][0] + o.M[0][1] * o.M[0][1] + o.M[0][2] * o.M[0][2];
  real c0 = o.M[1][0] * o.M[1][0] + o.M[1][1] * o.M[1][1] + o.M[1][2] * o.M[1][2];
  o.a = a0 + blur;
  o.b = o.M[0][0] * o.M[1][0] + o.M[0][1] * o.M[1][1] + o.M[0][2] * o.M[1][2];
  o.c = c0 + blur;
  // eig (generic.py:217-230), same eigenvalues in a cancellation-free form: det(M M^T) is the sum of the squared
  // 2x2 minors of M (Cauchy-Binet), gap = tr^2 - 4 det = (a-c)^2 + 4 b^2, lambda2 = det / lambda1.  The reference's
  // (tr - sqrt(gap)) / 2 loses all fp32 digits for elongated splats.
  real m01 = o.M[0][0] * o.M[1][1] - o.M[0][1] * o.M[1][0];
  real m02 = o.M[0][0] * o.M[1][2] - o.M[0][2] * o.M[1][0];
  real m12 = o.M[0][1] * o.M[1][2] - o.M[0][2] * o.M[1][1];
  real det = m01 * m01 + m02 * m02 + m12 * m12 + blur * (a0 + c0) + blur * blur;
  o.tr = o.a + o.c;
  real dac = o.a - o.c;
  o.gap = dac * dac + 4 * o.b * o.b;
  o.sg = m::sqrt(o.gap);
  real l1 = (o.tr + o.sg) * real(0.5), l2 = det / l1;
  o.sigma[0] = m::sqrt(l1); o.sigma[1] = m::sqrt(l2);
  // major eigenvector.  The reference normalises (a - l2, b) (generic.py:227), which is 0/0 for an axis-aligned
  // covariance with a < c (D17) and ill-conditioned near it.  (a - l2, b) and (b, c - l2) are parallel; for a < c
  // the second is the well-conditioned one, scaled by sign(b) so that v1.x >= 0 exactly as in the reference.
  o.swapped = o.a < o.c;
  o.sgn = (o.swapped && o.b < 0) ? real(-1) : real(1);
  o.u[0] = o.swapped ? o.b : o.a - l2;
  o.u[1] = o.swapped ? o.c - l2 : o.b;
  o.unorm = m::sqrt(o.u[0] * o.u[0] + o.u[1] * o.u[1]);
  if (o.unorm > 0) {
    o.v1[0] = o.sgn * o.u[0] / o.unorm; o.v1[1] = o.sgn * o.u[1] / o.unorm;
  } else {  // isotropic covariance: any axis is an eigenvector
    o.v1[0] = real(1); o.v1[1] = real(0);
  }
  o.alpha = real(1) / (real(1) + m::exp(-logit));
}

template <typename real>
__device__ __forceinline__ bool in_view(const Projected<real> &o, real width, real height, real near_plane,
                                        real far_plane, real alpha_threshold) {
  using m = math<real>;
  real gscale = m::sqrt(2 * m::log(o.alpha / alpha_threshold));  // NaN when alpha < threshold -> culled
  real sx = o.sigma[0] * gscale, sy = o.sigma[1] * gscale;
  real e1x = o.v1[0] * sx, e1y = o.v1[1] * sx, e2x = -o.v1[1] * sy, e2y = o.v1[0] * sy;
  real ex = m::sqrt(e1x * e1x + e2x * e2x), ey = m::sqrt(e1y * e1y + e2y * e2y);
  real z = o.cam_xyz[2];
  return (z > near_plane) && (z < far_plane) && (o.uv[0] + ex > 0) && (o.uv[1] + ey > 0) &&
         (o.uv[0] - ex < width) && (o.uv[1] - ey < height);
}

template <typename real>
__global__ void __launch_bounds__(256)
project_cull_kernel(const real *__restrict__ position, const real *__restrict__ log_scaling,
                    const real *__restrict__ rotation, const real *__restrict__ alpha_logit,
                    const real *__restrict__ T, const real *__restrict__ proj, int64_t n, real width,
                    real height, real near_plane, real far_plane, real blur, real margin,
                    real alpha_threshold, int32_t *__restrict__ flags) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Camera<real> cam = load_camera(T, proj);
  Projected<real> o;
  project_one(cam, position + 3 * i, log_scaling + 3 * i, rotation + 4 * i, alpha_logit[i], width, height,
              blur, margin, o);
  flags[i] = in_view(o, width, height, near_plane, far_plane, alpha_threshold) ? 1 : 0;
}

template <typename real>
__global__ void __launch_bounds__(256)
project_write_kernel(const real *__restrict__ position, const real *__restrict__ log_scaling,
                     const real *__restrict__ rotation, const real *__restrict__ alpha_logit,
                     const real *__restrict__ T, const real *__restrict__ proj, int64_t n, real width,
                     real height, real blur, real margin, real inv_far, real ndc_denom,
                     const int32_t *__restrict__ flags, const int32_t *__restrict__ incl,
                     real *__restrict__ points, real *__restrict__ depth, int64_t *__restrict__ indexes,
                     real *__restrict__ ndc) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  Camera<real> cam = load_camera(T, proj);
  Projected<real> o;
  project_one(cam, position + 3 * i, log_scaling + 3 * i, rotation + 4 * i, alpha_logit[i], width, height,
              blur, margin, o);
  int64_t k = incl[i] - 1;
  real *out = points + 7 * k;
  out[0] = o.uv[0]; out[1] = o.uv[1]; out[2] = o.v1[0]; out[3] = o.v1[1];
  out[4] = o.sigma[0]; out[5] = o.sigma[1]; out[6] = o.alpha;
  real z = o.cam_xyz[2];
  depth[k] = z;
  indexes[k] = i;
  if (ndc) ndc[k] = real(1) - (real(1) / z - inv_far) / ndc_denom;  // torch_lib/projection.py:123
}

// ---- single-pass project + cull + ORDERED compaction ------------------------------------------------------------
// project_cull / project_write above project every Gaussian twice (flags -> scan -> recompute and write), because the
// output position of a visible Gaussian is only known after the scan.  Here the projection is the *load* of a stream
// compaction: cub::DeviceSelect::If pulls items through a transform iterator that projects Gaussian i on access, keeps
// the ones in view (stable: indexes stay ascending, as torch.nonzero gives them in the reference, projection.py:147) with
// its single-pass decoupled look-back scan, and scatters them through an output iterator that splits the item into
// the reference's separate arrays.  One projection per Gaussian, one kernel instead of three.
template <typename real>
struct ProjItem {
  real pts[7];
  real depth, ndc;
  int32_t index, visible;
};

template <typename real>
struct ProjectOp {
  const real *position, *log_scaling, *rotation, *alpha_logit, *T, *proj;   // camera stays in device memory (no host read)
  real width, height, near_plane, far_plane, blur, margin, alpha_threshold, inv_far, ndc_denom;
  __device__ __forceinline__ ProjItem<real> operator()(int i) const {
    const Camera<real> cam = load_camera(T, proj);
    Projected<real> o;
    project_one(cam, position + 3 * (int64_t)i, log_scaling + 3 * (int64_t)i, rotation + 4 * (int64_t)i, alpha_logit[i],
                width, height, blur, margin, o);
    ProjItem<real> it;
    it.pts[0] = o.uv[0]; it.pts[1] = o.uv[1]; it.pts[2] = o.v1[0]; it.pts[3] = o.v1[1];
    it.pts[4] = o.sigma[0]; it.pts[5] = o.sigma[1]; it.pts[6] = o.alpha;
    const real z = o.cam_xyz[2];
    it.depth = z;
    it.ndc = real(1) - (real(1) / z - inv_far) / ndc_denom;   // torch_lib/projection.py:123
    it.index = i;
    it.visible = in_view(o, width, height, near_plane, far_plane, alpha_threshold) ? 1 : 0;
    return it;
  }
};

template <typename real>
struct ProjVisible {
  __device__ __forceinline__ bool operator()(const ProjItem<real> &it) const { return it.visible != 0; }
};

// output iterator: `out[k] = item` writes row k of points / depth / indexes / ndc
template <typename real>
struct ProjSink {
  real *points, *depth, *ndc;
  int64_t *indexes;
  int64_t k;
  __device__ __forceinline__ const ProjSink &operator=(const ProjItem<real> &it) const {
    real *row = points + 7 * k;
#pragma unroll
    for (int c = 0; c < 7; ++c) row[c] = it.pts[c];
    depth[k] = it.depth;
    indexes[k] = it.index;
    if (ndc != nullptr) ndc[k] = it.ndc;
    return *this;
  }
};

template <typename real>
struct ProjOutIt {
  using iterator_category = std::random_access_iterator_tag;
  using value_type = ProjItem<real>;
  using difference_type = int64_t;
  using pointer = void;
  using reference = ProjSink<real>;
  real *points, *depth, *ndc;
  int64_t *indexes;
  int64_t base;
  __host__ __device__ __forceinline__ ProjSink<real> operator[](int64_t k) const {
    return ProjSink<real>{points, depth, ndc, indexes, base + k};
  }
  __host__ __device__ __forceinline__ ProjSink<real> operator*() const { return (*this)[0]; }
  __host__ __device__ __forceinline__ ProjOutIt operator+(int64_t d) const {
    ProjOutIt r = *this;
    r.base += d;
    return r;
  }
};

template <typename real>
static ProjectOp<real> make_project_op(const real *position, const real *log_scaling, const real *rotation,
                                       const real *alpha_logit, const real *T, const real *proj, int32_t width,
                                       int32_t height, double near_plane, double far_plane, double blur_cov,
                                       double clamp_margin, double alpha_threshold) {
  ProjectOp<real> op;
  op.position = position; op.log_scaling = log_scaling; op.rotation = rotation; op.alpha_logit = alpha_logit;
  op.T = T; op.proj = proj;
  op.width = (real)width; op.height = (real)height; op.near_plane = (real)near_plane; op.far_plane = (real)far_plane;
  op.blur = (real)blur_cov; op.margin = (real)clamp_margin; op.alpha_threshold = (real)alpha_threshold;
  // eager-torch semantics of ndc_depth: python-float scalars are rounded to the tensor dtype
  op.inv_far = (real)(1.0 / far_plane); op.ndc_denom = (real)(1.0 / near_plane - 1.0 / far_plane);
  return op;
}

template <typename real>
using ProjInIt = cub::TransformInputIterator<ProjItem<real>, ProjectOp<real>, cub::CountingInputIterator<int>>;

template <typename real>
size_t project_compact_temp_bytes(int64_t n) {
  size_t temp = 0;
  if (n > 0) {
    ProjInIt<real> in(cub::CountingInputIterator<int>(0), ProjectOp<real>{});
    ProjOutIt<real> out{nullptr, nullptr, nullptr, nullptr, 0};
    cub::DeviceSelect::If(nullptr, temp, in, out, (int32_t *)nullptr, (int)n, ProjVisible<real>());
  }
  return temp;
}

__global__ void publish_word_kernel(const int32_t *__restrict__ dev, volatile int32_t *mapped) {
  *mapped = *dev;
  __threadfence_system();
}

template <typename real>
int project_compact(const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,
                    const real *T, const real *proj, int64_t n, int32_t width, int32_t height, double near_plane,
                    double far_plane, double blur_cov, double clamp_margin, double alpha_threshold, void *workspace,
                    size_t workspace_bytes, real *points, real *depth, int64_t *indexes, real *ndc,
                    int32_t *num_visible_host, cudaStream_t stream, bool mapped = false) {
  GS_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "project: n=%lld out of range", (long long)n);
  GS_CHECK_ARG(num_visible_host != nullptr, "project: num_visible_host is NULL");
  if (n == 0) {
    *num_visible_host = 0;
    return GS_OK;
  }
  size_t temp_bytes = project_compact_temp_bytes<real>(n);
  if (workspace == nullptr || workspace_bytes < 256 + temp_bytes) {
    set_error("project_compact: workspace too small (%zu < %zu)", workspace_bytes, 256 + temp_bytes);
    return GS_ERR_WORKSPACE_TOO_SMALL;
  }
  int32_t *num_dev = (int32_t *)workspace;                 // first 256 bytes: the selected count
  void *temp = (unsigned char *)workspace + 256;
  ProjInIt<real> in(cub::CountingInputIterator<int>(0),
                    make_project_op<real>(position, log_scaling, rotation, alpha_logit, T, proj, width, height, near_plane,
                                          far_plane, blur_cov, clamp_margin, alpha_threshold));
  ProjOutIt<real> out{points, depth, ndc, indexes, 0};
  GS_CUDA(cub::DeviceSelect::If(temp, temp_bytes, in, out, num_dev, (int)n, ProjVisible<real>(), stream));
  if (mapped) {
    publish_word_kernel<<<1, 1, 0, stream>>>(num_dev, num_visible_host);
    GS_LAUNCH_CHECK();
  } else {
    GS_CUDA(cudaMemcpyAsync(num_visible_host, num_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  }
  return GS_OK;
}

int project_compact_f32_mapped(const float *position, const float *log_scaling, const float *rotation,
                               const float *alpha_logit, const float *T, const float *proj, int64_t n, int32_t width,
                               int32_t height, double near_plane, double far_plane, double blur_cov, double clamp_margin,
                               double alpha_threshold, void *workspace, size_t workspace_bytes, float *points,
                               float *depth, int64_t *indexes, float *ndc, int32_t *mapped_word, cudaStream_t stream) {
  return project_compact<float>(position, log_scaling, rotation, alpha_logit, T, proj, n, width, height, near_plane,
                                far_plane, blur_cov, clamp_margin, alpha_threshold, workspace, workspace_bytes, points,
                                depth, indexes, ndc, mapped_word, stream, true);
}

template <typename real, int N>
__device__ __forceinline__ void block_reduce_atomic(real (&v)[N], real *__restrict__ out, real *smem) {
  // warp tree reduce, then one shared slot per value, then one global atomic per block
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v[k] += __shfl_xor_sync(full, v[k], off);
  if (threadIdx.x < N) smem[threadIdx.x] = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (v[k] != real(0)) atomicAdd(&smem[k], v[k]);
  __syncthreads();
  if (threadIdx.x < N && smem[threadIdx.x] != real(0)) atomicAdd(&out[threadIdx.x], smem[threadIdx.x]);
}

template <typename real>
#ifndef GS_PBWD_MIN_BLOCKS
#define GS_PBWD_MIN_BLOCKS 8   // fp32: 64 registers, a few spilled words, 32 warps per SM (latency-bound kernel: 69 -> 61 us)
#endif
__global__ void __launch_bounds__(128, sizeof(real) == 4 ? GS_PBWD_MIN_BLOCKS : 2)
project_bwd_kernel(const real *__restrict__ position, const real *__restrict__ log_scaling,
                   const real *__restrict__ rotation, const real *__restrict__ alpha_logit,
                   const real *__restrict__ T, const real *__restrict__ proj,
                   const int64_t *__restrict__ indexes, int64_t v, real width, real height, real blur,
                   real margin, const real *__restrict__ d_points, const real *__restrict__ d_depth,
                   real *__restrict__ d_position, real *__restrict__ d_log_scaling,
                   real *__restrict__ d_rotation, real *__restrict__ d_alpha_logit,
                   real *__restrict__ d_T, real *__restrict__ d_proj) {
  __shared__ real red[16];
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  real dT[12], dP[4];
#pragma unroll
  for (int k = 0; k < 12; ++k) dT[k] = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) dP[k] = 0;

  if (i < v) {
    int64_t idx = indexes[i];
    Camera<real> cam = load_camera(T, proj);
    Projected<real> o;
    const real *p = position + 3 * idx;
    project_one(cam, p, log_scaling + 3 * idx, rotation + 4 * idx, alpha_logit[idx], width, height, blur,
                margin, o);
    const real *dp = d_points + 7 * i;
    real d_mean[2] = {dp[0], dp[1]}, d_v1[2] = {dp[2], dp[3]}, d_sigma[2] = {dp[4], dp[5]}, d_alpha = dp[6];
    real z = o.cam_xyz[2];

    if (d_alpha_logit) d_alpha_logit[idx] = d_alpha * o.alpha * (1 - o.alpha);

    // eigen-decomposition reverse
    real d_l1 = d_sigma[0] / (2 * o.sigma[0]);
    real d_l2 = d_sigma[1] / (2 * o.sigma[1]);
    real d_u[2] = {0, 0};
    if (o.unorm > 0) {
      real dot = o.v1[0] * d_v1[0] + o.v1[1] * d_v1[1];
      d_u[0] = o.sgn * (d_v1[0] - o.v1[0] * dot) / o.unorm;
      d_u[1] = o.sgn * (d_v1[1] - o.v1[1] * dot) / o.unorm;
    }
    real d_a = 0, d_b = 0, d_c_direct = 0;
    if (o.swapped) {   // u = (b, c - l2)
      d_b = d_u[0]; d_c_direct = d_u[1]; d_l2 -= d_u[1];
    } else {           // u = (a - l2, b)
      d_a = d_u[0]; d_b = d_u[1]; d_l2 -= d_u[0];
    }
    real d_tr = (d_l1 + d_l2) * real(0.5), d_sg = (d_l1 - d_l2) * real(0.5);
    real d_gap = o.gap > 0 ? d_sg / (2 * o.sg) : real(0);
    d_tr += 2 * o.tr * d_gap;
    real d_det = -4 * d_gap;
    d_a += d_tr + o.c * d_det;
    real d_c = d_c_direct + d_tr + o.a * d_det;
    d_b += -2 * o.b * d_det;

    real dM[2][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dM[0][j] = 2 * d_a * o.M[0][j] + d_b * o.M[1][j];
      dM[1][j] = d_b * o.M[0][j] + 2 * d_c * o.M[1][j];
    }
    real dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dJ00 += dM[0][j] * o.A[0][j]; dJ02 += dM[0][j] * o.A[2][j];
      dJ11 += dM[1][j] * o.A[1][j]; dJ12 += dM[1][j] * o.A[2][j];
    }
    real dA[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dA[0][j] = o.J00 * dM[0][j];
      dA[1][j] = o.J11 * dM[1][j];
      dA[2][j] = o.J02 * dM[0][j] + o.J12 * dM[1][j];
    }
    // B = Rq diag(s); dW = dA B^T ; dB = R^T dA
    real dW[3][3], dB[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dW[r][k] = dA[r][0] * o.Rq[k][0] * o.s[0] + dA[r][1] * o.Rq[k][1] * o.s[1] + dA[r][2] * o.Rq[k][2] * o.s[2];
        dB[r][k] = cam.R[0][r] * dA[0][k] + cam.R[1][r] * dA[1][k] + cam.R[2][r] * dA[2][k];
      }
    real G[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      real ds = o.Rq[0][j] * dB[0][j] + o.Rq[1][j] * dB[1][j] + o.Rq[2][j] * dB[2][j];
      if (d_log_scaling) d_log_scaling[3 * idx + j] = ds * o.s[j];
#pragma unroll
      for (int r = 0; r < 3; ++r) G[r][j] = dB[r][j] * o.s[j];
    }
    if (d_rotation) {
      real x = o.qh[0], y = o.qh[1], zq = o.qh[2], w = o.qh[3];
      real dq[4];
      dq[0] = 2 * y * (G[0][1] + G[1][0]) + 2 * zq * (G[0][2] + G[2][0]) - 4 * x * (G[1][1] + G[2][2]) + 2 * w * (G[2][1] - G[1][2]);
      dq[1] = 2 * x * (G[0][1] + G[1][0]) - 4 * y * (G[0][0] + G[2][2]) + 2 * zq * (G[1][2] + G[2][1]) + 2 * w * (G[0][2] - G[2][0]);
      dq[2] = 2 * x * (G[0][2] + G[2][0]) + 2 * y * (G[1][2] + G[2][1]) - 4 * zq * (G[0][0] + G[1][1]) + 2 * w * (G[1][0] - G[0][1]);
      dq[3] = 2 * zq * (G[1][0] - G[0][1]) + 2 * y * (G[0][2] - G[2][0]) + 2 * x * (G[2][1] - G[1][2]);
      real dot = o.qh[0] * dq[0] + o.qh[1] * dq[1] + o.qh[2] * dq[2] + o.qh[3] * dq[3];
#pragma unroll
      for (int k = 0; k < 4; ++k) d_rotation[4 * idx + k] = (dq[k] - o.qh[k] * dot) / o.qnorm;
    }

    // Jacobian / perspective reverse
    real z2 = z * z;
    real d_z = d_depth[i];
    real d_f[2] = {dJ00 / z, dJ11 / z};
    d_z -= dJ00 * cam.fx / z2 + dJ11 * cam.fy / z2;
    real d_tc[2] = {-dJ02 / z, -dJ12 / z};
    real d_pp[2] = {dJ02 / z, dJ12 / z};
    d_z += dJ02 * (o.tc[0] - cam.cx) / z2 + dJ12 * (o.tc[1] - cam.cy) / z2;
    real d_uv[2] = {d_mean[0] + (o.inside[0] ? d_tc[0] : real(0)), d_mean[1] + (o.inside[1] ? d_tc[1] : real(0))};
    d_f[0] += d_uv[0] * o.cam_xyz[0] / z; d_f[1] += d_uv[1] * o.cam_xyz[1] / z;
    d_pp[0] += d_uv[0]; d_pp[1] += d_uv[1];
    real d_cam[3];
    d_cam[0] = d_uv[0] * cam.fx / z;
    d_cam[1] = d_uv[1] * cam.fy / z;
    d_z -= (d_uv[0] * cam.fx * o.cam_xyz[0] + d_uv[1] * cam.fy * o.cam_xyz[1]) / z2;
    d_cam[2] = d_z;
    if (d_position)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        d_position[3 * idx + k] = cam.R[0][k] * d_cam[0] + cam.R[1][k] * d_cam[1] + cam.R[2][k] * d_cam[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) dT[4 * r + k] = d_cam[r] * p[k] + dW[r][k];
      dT[4 * r + 3] = d_cam[r];
    }
    dP[0] = d_f[0]; dP[1] = d_f[1]; dP[2] = d_pp[0]; dP[3] = d_pp[1];
  }
  if (d_T) block_reduce_atomic<real, 12>(dT, d_T, red);
  if (d_proj) {
    __syncthreads();
    block_reduce_atomic<real, 4>(dP, d_proj, red);
  }
}

// camera position = -A^-1 t for T_camera_world = [A t; 0 1]  (perspective/params.py:78-80 uses torch.inverse)
template <typename real>
__global__ void camera_position_kernel(const real *__restrict__ T, real *__restrict__ out) {
  real a = T[0], b = T[1], c = T[2], d = T[4], e = T[5], f = T[6], g = T[8], h = T[9], i = T[10];
  real tx = T[3], ty = T[7], tz = T[11];
  real c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
  real c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
  real c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
  real inv_det = real(1) / (a * c00 + b * c10 + c * c20);
  out[0] = -(c00 * tx + c01 * ty + c02 * tz) * inv_det;
  out[1] = -(c10 * tx + c11 * ty + c12 * tz) * inv_det;
  out[2] = -(c20 * tx + c21 * ty + c22 * tz) * inv_det;
}

template <typename real>
int project_cull(const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,
                 const real *T, const real *proj, int64_t n, int32_t width, int32_t height, double near_plane,
                 double far_plane, double blur_cov, double clamp_margin, double alpha_threshold, void *workspace,
                 size_t workspace_bytes, int32_t *num_visible_host, cudaStream_t stream) {
  GS_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "project: n=%lld out of range", (long long)n);
  GS_CHECK_ARG(num_visible_host != nullptr, "project: num_visible_host is NULL");
  size_t need = 0;
  gs_project_workspace_bytes(n, &need);
  if (workspace_bytes < need) {
    set_error("project: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GS_ERR_WORKSPACE_TOO_SMALL;
  }
  if (n == 0) {
    *num_visible_host = 0;
    return GS_OK;
  }
  int32_t *flags = (int32_t *)workspace;
  int32_t *incl = flags + align_up(n, 64);
  void *temp = (void *)(incl + align_up(n, 64));
  size_t temp_bytes = workspace_bytes - 2 * align_up(n, 64) * sizeof(int32_t);
  project_cull_kernel<real><<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(
      position, log_scaling, rotation, alpha_logit, T, proj, n, (real)width, (real)height, (real)near_plane,
      (real)far_plane, (real)blur_cov, (real)clamp_margin, (real)alpha_threshold, flags);
  GS_LAUNCH_CHECK();
  GS_CUDA(cub::DeviceScan::InclusiveSum(temp, temp_bytes, flags, incl, (int)n, stream));
  GS_CUDA(cudaMemcpyAsync(num_visible_host, incl + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  return GS_OK;
}

template <typename real>
int project_write(const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,
                  const real *T, const real *proj, int64_t n, int32_t width, int32_t height, double near_plane,
                  double far_plane, double blur_cov, double clamp_margin, const void *workspace, real *points,
                  real *depth, int64_t *indexes, real *ndc, cudaStream_t stream) {
  if (n == 0) return GS_OK;
  const int32_t *flags = (const int32_t *)workspace;
  const int32_t *incl = flags + align_up(n, 64);
  // eager-torch semantics of ndc_depth: python-float scalars are rounded to the tensor dtype
  real inv_far = (real)(1.0 / far_plane), denom = (real)(1.0 / near_plane - 1.0 / far_plane);
  project_write_kernel<real><<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(
      position, log_scaling, rotation, alpha_logit, T, proj, n, (real)width, (real)height, (real)blur_cov,
      (real)clamp_margin, inv_far, denom, flags, incl, points, depth, indexes, ndc);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

template <typename real>
int project_bwd(const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,
                const real *T, const real *proj, const int64_t *indexes, int64_t v, int32_t width,
                int32_t height, double blur_cov, double clamp_margin, const real *d_points, const real *d_depth,
                real *d_position, real *d_log_scaling, real *d_rotation, real *d_alpha_logit, real *d_T,
                real *d_proj, cudaStream_t stream) {
  if (v == 0) return GS_OK;
  project_bwd_kernel<real><<<(unsigned)ceil_div(v, 128), 128, 0, stream>>>(
      position, log_scaling, rotation, alpha_logit, T, proj, indexes, v, (real)width, (real)height,
      (real)blur_cov, (real)clamp_margin, d_points, d_depth, d_position, d_log_scaling, d_rotation,
      d_alpha_logit, d_T, d_proj);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // namespace gs

extern "C" int gs_project_workspace_bytes(int64_t n, size_t *bytes) {
  size_t temp = 0;
  if (n > 0) cub::DeviceScan::InclusiveSum(nullptr, temp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
  *bytes = 2 * gs::align_up((size_t)(n > 0 ? n : 0), 64) * sizeof(int32_t) + gs::align_up(temp, 256) + 256;
  // the same workspace serves gs_project_compact_*
  size_t compact = 0;
  gs_project_compact_workspace_bytes(n, 1, &compact);
  if (compact > *bytes) *bytes = compact;
  return GS_OK;
}

extern "C" int gs_project_compact_workspace_bytes(int64_t n, int32_t fp64, size_t *bytes) {
  GS_CHECK_ARG(bytes != nullptr && n >= 0, "project_compact_workspace_bytes: bad arguments");
  const size_t temp = fp64 ? gs::project_compact_temp_bytes<double>(n) : gs::project_compact_temp_bytes<float>(n);
  *bytes = 256 + gs::align_up(temp, 256) + 256;
  return GS_OK;
}

#define GS_PROJECT_API(SUFFIX, real)                                                                              \
  extern "C" int gs_project_cull_##SUFFIX(                                                                        \
      const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,               \
      const real *T, const real *projection, int64_t n, int32_t width, int32_t height, double near_plane,         \
      double far_plane, double blur_cov, double clamp_margin, double alpha_threshold, void *workspace,            \
      size_t workspace_bytes, int32_t *num_visible_host, void *stream) {                                          \
    return gs::project_cull<real>(position, log_scaling, rotation, alpha_logit, T, projection, n, width, height,  \
                                  near_plane, far_plane, blur_cov, clamp_margin, alpha_threshold, workspace,      \
                                  workspace_bytes, num_visible_host, (cudaStream_t)stream);                       \
  }                                                                                                               \
  extern "C" int gs_project_compact_##SUFFIX(                                                                     \
      const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,               \
      const real *T, const real *projection, int64_t n, int32_t width, int32_t height, double near_plane,         \
      double far_plane, double blur_cov, double clamp_margin, double alpha_threshold, void *workspace,            \
      size_t workspace_bytes, real *points, real *depth, int64_t *indexes, real *ndc_depth,                       \
      int32_t *num_visible_host, void *stream) {                                                                  \
    return gs::project_compact<real>(position, log_scaling, rotation, alpha_logit, T, projection, n, width,       \
                                     height, near_plane, far_plane, blur_cov, clamp_margin, alpha_threshold,      \
                                     workspace, workspace_bytes, points, depth, indexes, ndc_depth,               \
                                     num_visible_host, (cudaStream_t)stream);                                     \
  }                                                                                                               \
  extern "C" int gs_project_write_##SUFFIX(                                                                       \
      const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,               \
      const real *T, const real *projection, int64_t n, int32_t width, int32_t height, double near_plane,         \
      double far_plane, double blur_cov, double clamp_margin, const void *workspace, real *points, real *depth,   \
      int64_t *indexes, real *ndc_depth, void *stream) {                                                          \
    return gs::project_write<real>(position, log_scaling, rotation, alpha_logit, T, projection, n, width, height, \
                                   near_plane, far_plane, blur_cov, clamp_margin, workspace, points, depth,       \
                                   indexes, ndc_depth, (cudaStream_t)stream);                                     \
  }                                                                                                               \
  extern "C" int gs_project_bwd_##SUFFIX(                                                                         \
      const real *position, const real *log_scaling, const real *rotation, const real *alpha_logit,               \
      const real *T, const real *projection, const int64_t *indexes, int64_t v, int32_t width, int32_t height,    \
      double blur_cov, double clamp_margin, const real *d_points, const real *d_depth, real *d_position,          \
      real *d_log_scaling, real *d_rotation, real *d_alpha_logit, real *d_T, real *d_projection, void *stream) {  \
    return gs::project_bwd<real>(position, log_scaling, rotation, alpha_logit, T, projection, indexes, v, width,  \
                                 height, blur_cov, clamp_margin, d_points, d_depth, d_position, d_log_scaling,    \
                                 d_rotation, d_alpha_logit, d_T, d_projection, (cudaStream_t)stream);             \
  }

extern "C" int gs_camera_position_f32(const float *T_camera_world, float *camera_pos, void *stream) {
  gs::camera_position_kernel<float><<<1, 1, 0, (cudaStream_t)stream>>>(T_camera_world, camera_pos);
  GS_LAUNCH_CHECK();
  return GS_OK;
}
extern "C" int gs_camera_position_f64(const double *T_camera_world, double *camera_pos, void *stream) {
  gs::camera_position_kernel<double><<<1, 1, 0, (cudaStream_t)stream>>>(T_camera_world, camera_pos);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

GS_PROJECT_API(f32, float)
GS_PROJECT_API(f64, double)
