// optim.cu -- N3: sparse, visibility-weighted optimiser step on the visible set (the step right after backward).
//
// Semantics: the four Taichi kernels of the reference's optimisers,
//   optim/fractional_adam.py:7-44 (scalar), :46-86 (vector), optim/fractional_laprop.py:6-39 (scalar), :41-76 (vector),
// and, when `param` is given, the torch code that follows them in optim/fractional.py:131-147,184-199
// (clip, mask_lr, point_lr, non-finite -> 0, param[indexes] -= lr_step * saturate(weight)); also the running
// visibility update of optim/visibility_aware.py:37-48,90-91.
//
// "Fractional" = every visible point advances its own moment estimates by its weight w (beta^w) and carries its own
// step count (total_weight) for the bias correction; invisible points are not touched at all.  The reference does
// this in 6-10 separate launches per parameter group (kernel, clamp_, two multiplies, isfinite mask, exp, indexed
// subtract, each a full pass over (M, D)); here one launch reads grad / m / v / param rows once and writes m / v /
// param once: 2 x 4 D bytes read-modify-write of state, 4 D read of grad, 8 D of param per visible row -- HBM-bound.
#include "common.cuh"

namespace gs {

struct OptimParams {
  float lr, beta1, beta2, eps, clip, grad_smooth;
  int d;
  int64_t m_rows;
};

__device__ __forceinline__ float lerp_ref(float t, float a, float b) { return a * t + b * (1.0f - t); }  // generic.py:489-490

__device__ __forceinline__ float finish_step(float step, float w, int j, int64_t idx, const OptimParams &p,
                                             const float *__restrict__ mask_lr, const float *__restrict__ point_lr) {
  if (p.clip > 0.f) { const float mx = p.lr * p.clip; step = fminf(fmaxf(step, -mx), mx); }
  if (mask_lr != nullptr) step *= mask_lr[j];
  if (point_lr != nullptr) step *= point_lr[idx];
  if (!isfinite(step)) step = 0.f;
  return step;
}

// scalar kinds: one thread per (visible row, group of 4 components).  The per-row factors (beta^w, bias correction,
// saturate(w): five transcendental calls) are formed once per thread instead of once per component, and rows whose
// width is a multiple of 4 move as float4.
template <bool LAPROP, bool BIAS, bool VEC4>
__global__ void __launch_bounds__(256)
optim_scalar_kernel(const int64_t *__restrict__ indexes, const float *__restrict__ weight,
                    const float *__restrict__ grad_scale, float *__restrict__ m_arr, float *__restrict__ v_arr,
                    const float *__restrict__ total_weight, const float *__restrict__ grad, OptimParams p,
                    float *__restrict__ lr_step, float *__restrict__ param, const float *__restrict__ mask_lr,
                    const float *__restrict__ point_lr) {
  const int groups = (p.d + 3) >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.m_rows * groups) return;
  const int64_t i = t / groups;
  const int j0 = (int)(t - i * groups) * 4;
  const int cnt = min(4, p.d - j0);
  const int64_t idx = indexes[i];
  const float w = weight[i], tw = total_weight[idx];
  const float b1w = powf(p.beta1, w), b2w = powf(p.beta2, w);
  float bias1 = 1.0f, bias2 = 1.0f;   // ADAM: bias1 holds the combined factor
  if (BIAS) {
    if (LAPROP) { bias1 = 1.0f - powf(p.beta1, tw); bias2 = 1.0f - powf(p.beta2, tw); }
    else bias1 = sqrtf(1.0f - powf(p.beta2, tw)) / (1.0f - powf(p.beta1, tw));
  }
  const float sat = 1.0f - 1.0f / expf(2.0f * w);   // saturate, fractional.py:149-150
  const float gdiv = grad_scale != nullptr ? grad_scale[i] + p.grad_smooth : 1.0f;   // visibility_aware.py:97-99
  const int64_t e0 = idx * p.d + j0;
  float g[4] = {0.f, 0.f, 0.f, 0.f}, m[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f}, x[4] = {0.f, 0.f, 0.f, 0.f};
  if (VEC4) {
    const float4 g4 = *reinterpret_cast<const float4 *>(grad + e0), m4 = *reinterpret_cast<const float4 *>(m_arr + e0);
    const float4 v4 = *reinterpret_cast<const float4 *>(v_arr + e0);
    g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
    m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
    v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
    if (param != nullptr) {
      const float4 x4 = *reinterpret_cast<const float4 *>(param + e0);
      x[0] = x4.x; x[1] = x4.y; x[2] = x4.z; x[3] = x4.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < cnt) {
        g[c] = grad[e0 + c]; m[c] = m_arr[e0 + c]; v[c] = v_arr[e0 + c];
        if (param != nullptr) x[c] = param[e0 + c];
      }
  }
  float step[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float gc = g[c];
    if (grad_scale != nullptr) gc = gc / gdiv;
    if (LAPROP) {
      v[c] = lerp_ref(b2w, v[c], gc * gc);
      m[c] = lerp_ref(b1w, m[c], gc / fmaxf(sqrtf(v[c] / bias2), p.eps));
      step[c] = m[c] * p.lr / bias1;
    } else {
      m[c] = lerp_ref(b1w, m[c], gc);
      v[c] = lerp_ref(b2w, v[c], gc * gc);
      step[c] = m[c] / fmaxf(sqrtf(v[c]), p.eps) * bias1 * p.lr;
    }
    if (param != nullptr && c < cnt) x[c] -= finish_step(step[c], w, j0 + c, idx, p, mask_lr, point_lr) * sat;
  }
  if (VEC4) {
    *reinterpret_cast<float4 *>(m_arr + e0) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4 *>(v_arr + e0) = make_float4(v[0], v[1], v[2], v[3]);
    if (param != nullptr) *reinterpret_cast<float4 *>(param + e0) = make_float4(x[0], x[1], x[2], x[3]);
    if (lr_step != nullptr) *reinterpret_cast<float4 *>(lr_step + i * p.d + j0) = make_float4(step[0], step[1], step[2], step[3]);
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < cnt) {
        m_arr[e0 + c] = m[c]; v_arr[e0 + c] = v[c];
        if (param != nullptr) param[e0 + c] = x[c];
        if (lr_step != nullptr) lr_step[i * p.d + j0 + c] = step[c];
      }
  }
}

// vector kinds: one running second moment per row (squared gradient norm); one thread per visible row
template <bool LAPROP, bool BIAS>
__global__ void __launch_bounds__(256)
optim_vector_kernel(const int64_t *__restrict__ indexes, const float *__restrict__ weight,
                    const float *__restrict__ grad_scale, float *__restrict__ m_arr, float *__restrict__ v_arr,
                    const float *__restrict__ total_weight, const float *__restrict__ grad, OptimParams p,
                    float *__restrict__ lr_step, float *__restrict__ param, const float *__restrict__ mask_lr,
                    const float *__restrict__ point_lr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.m_rows) return;
  const int64_t idx = indexes[i];
  const float w = weight[i], tw = total_weight[idx];
  const float *g = grad + idx * p.d;
  float norm = 0.f;
  for (int j = 0; j < p.d; ++j) {
    const float gj = grad_scale != nullptr ? g[j] / (grad_scale[i] + p.grad_smooth) : g[j];
    norm = fmaf(gj, gj, norm);
  }
  const float b1w = powf(p.beta1, w), b2w = powf(p.beta2, w);
  const float v = lerp_ref(b2w, v_arr[idx], norm);
  float denom, scale;
  if (LAPROP) {
    const float bias1 = BIAS ? 1.0f - powf(p.beta1, tw) : 1.0f, bias2 = BIAS ? 1.0f - powf(p.beta2, tw) : 1.0f;
    denom = fmaxf(sqrtf(v / bias2), p.eps);
    scale = p.lr / bias1;
  } else {
    denom = fmaxf(sqrtf(v), p.eps);
    scale = (BIAS ? sqrtf(1.0f - powf(p.beta2, tw)) / (1.0f - powf(p.beta1, tw)) : 1.0f) * p.lr;
  }
  const float sat = 1.0f - 1.0f / expf(2.0f * w);
  for (int j = 0; j < p.d; ++j) {
    const int64_t e = idx * p.d + j;
    const float gj = grad_scale != nullptr ? g[j] / (grad_scale[i] + p.grad_smooth) : g[j];
    float m, step;
    if (LAPROP) {
      m = lerp_ref(b1w, m_arr[e], gj / denom);
      step = m * scale;
    } else {
      m = lerp_ref(b1w, m_arr[e], gj);
      step = m / denom * scale;
    }
    m_arr[e] = m;
    if (lr_step != nullptr) lr_step[i * p.d + j] = step;
    if (param != nullptr) param[e] -= finish_step(step, w, j, idx, p, mask_lr, point_lr) * sat;
  }
  v_arr[idx] = v;
}

// running visibility (visibility_aware.py:37-48) and the per-point step counter (:90-91)
__global__ void __launch_bounds__(256)
optim_visibility_kernel(float *__restrict__ running_vis, const float *__restrict__ visibility,
                        const int64_t *__restrict__ indexes, float *__restrict__ total_weight, float beta, float eps,
                        int64_t m_rows, float *__restrict__ weight_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m_rows) return;
  const int64_t idx = indexes[i];
  const float vis = visibility[i], rv = running_vis[idx];
  // power_lerp(beta, vis, rv, k = 4) = lerp(beta, vis^4, rv^4)^(1/4), lerp(t, a, b) = a + (b - a) t
  const float a = vis * vis * vis * vis, b = rv * rv * rv * rv;
  const float updated = powf(a + (b - a) * beta, 0.25f);
  running_vis[idx] = updated;
  const float w = vis / fmaxf(updated, eps);
  weight_out[i] = w;
  total_weight[idx] += w;
}

}  // namespace gs

extern "C" int gs_optim_step_f32(int32_t algorithm, int32_t vector, int32_t bias_correction, const int64_t *indexes,
                                 const float *weight, const float *grad_scale, double grad_smooth, int64_t m_rows,
                                 int32_t d, float *m_state, float *v_state, const float *total_weight,
                                 const float *grad, double lr, double beta1, double beta2, double eps, float *lr_step,
                                 float *param, double clip, const float *mask_lr, const float *point_lr,
                                 void *stream_) {
  using namespace gs;
  cudaStream_t stream = (cudaStream_t)stream_;
  GS_CHECK_ARG(algorithm == GS_OPTIM_ADAM || algorithm == GS_OPTIM_LAPROP, "optim_step: unknown algorithm %d", algorithm);
  GS_CHECK_ARG(m_rows >= 0 && d >= 1, "optim_step: bad shape (%lld, %d)", (long long)m_rows, d);
  GS_CHECK_ARG(lr_step != nullptr || param != nullptr, "optim_step: neither lr_step nor param given");
  if (m_rows == 0) return GS_OK;
  OptimParams p;
  p.lr = (float)lr; p.beta1 = (float)beta1; p.beta2 = (float)beta2; p.eps = (float)eps; p.clip = (float)clip;
  p.grad_smooth = (float)grad_smooth; p.d = d; p.m_rows = m_rows;
  const int64_t threads = vector ? m_rows : m_rows * ((d + 3) / 4);
  const unsigned grid = (unsigned)ceil_div(threads, 256);
  auto aligned16 = [](const void *q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec4 = !vector && d % 4 == 0 && aligned16(m_state) && aligned16(v_state) && aligned16(grad) &&
                    aligned16(param) && aligned16(lr_step);
#define GS_OPT(KERNEL, ...)                                                                                        \
  KERNEL<__VA_ARGS__><<<grid, 256, 0, stream>>>(indexes, weight, grad_scale, m_state, v_state, total_weight, grad, \
                                                p, lr_step, param, mask_lr, point_lr)
#define GS_OPT_SCALAR(LAPROP_, BIAS_) \
  do { if (vec4) GS_OPT(optim_scalar_kernel, LAPROP_, BIAS_, true); else GS_OPT(optim_scalar_kernel, LAPROP_, BIAS_, false); } while (0)
  const bool laprop = algorithm == GS_OPTIM_LAPROP, bias = bias_correction != 0;
  if (vector) {
    if (laprop) { if (bias) GS_OPT(optim_vector_kernel, true, true); else GS_OPT(optim_vector_kernel, true, false); }
    else        { if (bias) GS_OPT(optim_vector_kernel, false, true); else GS_OPT(optim_vector_kernel, false, false); }
  } else {
    if (laprop) { if (bias) GS_OPT_SCALAR(true, true); else GS_OPT_SCALAR(true, false); }
    else        { if (bias) GS_OPT_SCALAR(false, true); else GS_OPT_SCALAR(false, false); }
  }
#undef GS_OPT_SCALAR
#undef GS_OPT
  GS_LAUNCH_CHECK();
  return GS_OK;
}

extern "C" int gs_optim_update_visibility_f32(float *running_vis, const float *visibility, const int64_t *indexes,
                                              float *total_weight, double beta, double eps, int64_t m_rows,
                                              float *weight_out, void *stream_) {
  GS_CHECK_ARG(m_rows >= 0, "optim_update_visibility: bad row count");
  if (m_rows == 0) return GS_OK;
  gs::optim_visibility_kernel<<<(unsigned)gs::ceil_div(m_rows, 256), 256, 0, (cudaStream_t)stream_>>>(
      running_vis, visibility, indexes, total_weight, (float)beta, (float)eps, m_rows, weight_out);
  GS_LAUNCH_CHECK();
  return GS_OK;
}
