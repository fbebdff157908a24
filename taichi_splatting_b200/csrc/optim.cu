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

// scalar kinds: one thread per (visible row, component)
template <bool LAPROP, bool BIAS>
__global__ void __launch_bounds__(256)
optim_scalar_kernel(const int64_t *__restrict__ indexes, const float *__restrict__ weight,
                    const float *__restrict__ grad_scale, float *__restrict__ m_arr, float *__restrict__ v_arr,
                    const float *__restrict__ total_weight, const float *__restrict__ grad, OptimParams p,
                    float *__restrict__ lr_step, float *__restrict__ param, const float *__restrict__ mask_lr,
                    const float *__restrict__ point_lr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.m_rows * p.d) return;
  const int64_t i = t / p.d;
  const int j = (int)(t - i * p.d);
  const int64_t idx = indexes[i];
  const float w = weight[i], tw = total_weight[idx];
  const int64_t e = idx * p.d + j;
  float g = grad[e];
  if (grad_scale != nullptr) g = g / (grad_scale[i] + p.grad_smooth);   // visibility_aware.py:97-99
  const float b1w = powf(p.beta1, w), b2w = powf(p.beta2, w);
  float m, v, step;
  if (LAPROP) {
    const float bias1 = BIAS ? 1.0f - powf(p.beta1, tw) : 1.0f, bias2 = BIAS ? 1.0f - powf(p.beta2, tw) : 1.0f;
    v = lerp_ref(b2w, v_arr[e], g * g);
    m = lerp_ref(b1w, m_arr[e], g / fmaxf(sqrtf(v / bias2), p.eps));
    step = m * p.lr / bias1;
  } else {
    const float bias = BIAS ? sqrtf(1.0f - powf(p.beta2, tw)) / (1.0f - powf(p.beta1, tw)) : 1.0f;
    m = lerp_ref(b1w, m_arr[e], g);
    v = lerp_ref(b2w, v_arr[e], g * g);
    step = m / fmaxf(sqrtf(v), p.eps) * bias * p.lr;
  }
  m_arr[e] = m;
  v_arr[e] = v;
  if (lr_step != nullptr) lr_step[t] = step;
  if (param != nullptr) {
    step = finish_step(step, w, j, idx, p, mask_lr, point_lr);
    param[e] -= step * (1.0f - 1.0f / expf(2.0f * w));   // saturate, fractional.py:149-150
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
  const int64_t threads = vector ? m_rows : m_rows * d;
  const unsigned grid = (unsigned)ceil_div(threads, 256);
#define GS_OPT(KERNEL, LAPROP_, BIAS_)                                                                             \
  KERNEL<LAPROP_, BIAS_><<<grid, 256, 0, stream>>>(indexes, weight, grad_scale, m_state, v_state, total_weight,   \
                                                   grad, p, lr_step, param, mask_lr, point_lr)
  const bool laprop = algorithm == GS_OPTIM_LAPROP, bias = bias_correction != 0;
  if (vector) {
    if (laprop) { if (bias) GS_OPT(optim_vector_kernel, true, true); else GS_OPT(optim_vector_kernel, true, false); }
    else        { if (bias) GS_OPT(optim_vector_kernel, false, true); else GS_OPT(optim_vector_kernel, false, false); }
  } else {
    if (laprop) { if (bias) GS_OPT(optim_scalar_kernel, true, true); else GS_OPT(optim_scalar_kernel, true, false); }
    else        { if (bias) GS_OPT(optim_scalar_kernel, false, true); else GS_OPT(optim_scalar_kernel, false, false); }
  }
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
