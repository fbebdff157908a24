// raster_digest.cu -- per-Gaussian raster records ("digest"), written once per frame and gathered by the tuned
// rasteriser kernels instead of the raw (V,7) rows.
//
// Why: the rasteriser kernels stage every (tile, splat) overlap; from the reference layout that is 7 + F + 1
// scattered 4-byte loads on 28-byte rows (not 16-byte aligned, so no vector loads), two IEEE divisions, a log and
// three square roots per overlap -- ~20 % of the shared/L1 data-pipe time of the forward kernel, which is what
// bounds it.  The digest is 64 bytes per visible Gaussian, 64-byte aligned: four LDG.128 per overlap, and all the
// per-Gaussian arithmetic (reciprocal sigmas, exp-scaled basis, support radius) is done once per Gaussian
// instead of once per overlap (K / V = 3.8 at the bench workload).
//
//   R0 = { mean.x, mean.y, ux, wx }        u = axis / sigma.x * k,  w = perp(axis) / sigma.y * k,  k = sqrt(log2(e)/2)
//   R1 = { uy, wy, alpha, depth }          so that  exp(-0.5 (tx^2 + ty^2)) = 2^-( (u.d)^2 + (w.d)^2 )
//   R2 = { features[0..3] }                (zero padded)
//   R3 = { rcs, 1/sigma.x, 1/sigma.y, 0 }  rcs = conservative support radius of the alpha threshold in u/w units
//                                          (<= 0: the splat can never pass the threshold)
#include "raster_common.cuh"

namespace gs {

constexpr float kExpScaleD = 0.84932180028801904f;

__global__ void __launch_bounds__(256)
raster_digest_kernel(const float *__restrict__ points, const float *__restrict__ features,
                     const float *__restrict__ depths, int64_t v, int F, float thr, float4 *__restrict__ digest) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v) return;
  const float *g = points + 7 * i;
  const float mx = g[0], my = g[1], ax = g[2], ay = g[3], sx = g[4], sy = g[5], alpha = g[6];
  const float isx = 1.0f / sx, isy = 1.0f / sy;
  const float ux = ax * isx * kExpScaleD, uy = ay * isx * kExpScaleD;
  const float wx = -ay * isy * kExpScaleD, wy = ax * isy * kExpScaleD;
  float rcs = -1.0f;
  // margin covers fp32 / ex2.approx evaluation error of the kernels that use the record
  if (alpha > thr) rcs = (sqrtf(2.0f * __logf(alpha / thr)) * 1.001f + 0.01f) * kExpScaleD;
  float4 fv = make_float4(0.f, 0.f, 0.f, 0.f);
  const float *fp = features + (int64_t)F * i;
  fv.x = fp[0];
  if (F > 1) fv.y = fp[1];
  if (F > 2) fv.z = fp[2];
  if (F > 3) fv.w = fp[3];
  float4 *out = digest + 4 * i;
  out[0] = make_float4(mx, my, ux, wx);
  out[1] = make_float4(uy, wy, alpha, depths != nullptr ? depths[i] : 0.f);
  out[2] = fv;
  out[3] = make_float4(rcs, isx, isy, 0.f);
}

int raster_digest_f32(const float *points, const float *features, const float *depths, int64_t v, int F,
                      double alpha_threshold, void *digest, cudaStream_t stream) {
  GS_CHECK_ARG(F >= 1 && F <= 4, "raster_digest: 1..4 features, got %d", F);
  GS_CHECK_ARG(v == 0 || (points != nullptr && features != nullptr && digest != nullptr), "raster_digest: NULL buffer");
  GS_CHECK_ARG((reinterpret_cast<uintptr_t>(digest) & 63) == 0, "raster_digest: digest must be 64-byte aligned");
  if (v == 0) return GS_OK;
  raster_digest_kernel<<<(unsigned)ceil_div(v, 256), 256, 0, stream>>>(points, features, depths, v, F,
                                                                        (float)alpha_threshold,
                                                                        reinterpret_cast<float4 *>(digest));
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // namespace gs

extern "C" int gs_raster_digest_bytes(int64_t v, size_t *bytes) {
  GS_CHECK_ARG(bytes != nullptr && v >= 0, "raster_digest_bytes: bad arguments");
  *bytes = (size_t)v * 64;
  return GS_OK;
}

extern "C" int gs_raster_digest_f32(const float *points, const float *features, const float *depths, int64_t v,
                                    int32_t F, const gs_raster_config *cfg, void *digest, void *stream) {
  GS_CHECK_ARG(cfg != nullptr, "raster_digest: config is NULL");
  return gs::raster_digest_f32(points, features, depths, v, F, cfg->alpha_threshold, digest, (cudaStream_t)stream);
}
