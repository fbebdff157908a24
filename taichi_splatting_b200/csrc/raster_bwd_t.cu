// raster_bwd_t.cu -- R9, "transpose" variant of the tuned backward kernel (fp32, 16x16 tiles, plain pdf).
//
// Same semantics, tile / warp-rectangle / hit-list structure and moment formulation as raster_bwd.cu.  What
// changes is how the per-(pixel, splat) quantities become per-splat sums.  raster_bwd.cu reduces 11 values
// across the 32 pixel lanes with shuffles for EVERY splat (15 SHFL + ~30 issue slots per splat, and the LSU
// pipe -- LDS + SHFL at one warp instruction per cycle per SM -- co-limits the kernel).  Here a warp
//
//   phase 1  sweeps 8 splats of its hit list pixel-parallel (lane = pixel) and parks the four per-pixel scalars
//            {Gp, weight, G^2, |G dpdf/dmean|_1} of each in a padded shared-memory panel [splat][pixel] with one
//            conflict-free STS.128 per splat;
//   phase 2  re-reads the panel transposed (lane = (splat, pixel row)): each lane walks the 8 pixels of one row
//            of the 8x4 rectangle for ONE splat and accumulates in registers -- the row-local moments are just
//            sum Gp, sum Gp*i, sum Gp*i^2 with the column index i an immediate, features need the row's
//            dL/dimage (broadcast LDS) -- then a 2-stage transposed butterfly over the 4 rows (9 shuffles per
//            8 splats) leaves 3 finished sums per lane, added to the batch accumulators with 3 conflict-free
//            shared atomics per 8 splats.
//
// Per splat this is ~1 STS.128 + 2 LDS.128 + ~1 SHFL on the LSU pipe and ~20 issue slots for the reduction,
// against 15 SHFL and ~50 issue slots before.
#include "packed_f32.cuh"
#include "raster_common.cuh"

namespace gs {

#ifndef GS_BWDT_UNROLL
#define GS_BWDT_UNROLL 8
#endif
#ifndef GS_BWDT_MIN_BLOCKS
#define GS_BWDT_MIN_BLOCKS 3
#endif

namespace bwdt {

#ifdef GS_COUNT   // debug build only: work counters read by profiles/count_work.py
__device__ unsigned long long g_count[4];   // warp iterations (padded), hit-list entries, live (pixel, splat) lanes, batches
#endif

constexpr int kTile = 16;
constexpr int kBatch = 256;
constexpr int kChunk = 8;          // splats per phase-1 / phase-2 round
constexpr int kRow = 33;           // panel row stride in float4 (32 pixels + 1 pad: conflict-free transposed reads)
constexpr int kRowG = 9;           // gpix row stride in float4 (8 pixels + 1 pad)
constexpr int kAcc = 13;           // accumulator stride: 6 moments, 4 features, 2 heuristics, 1 pad (odd)
constexpr float kExpScale = 0.84932180028801904f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Smem {
  float4 a[kBatch + 1];            // mean.x, mean.y, (axis/sx)*k      (+1: null record that pads the hit lists)
  float4 b[kBatch + 1];            // (perp/sy)*k, alpha, unused
  float4 f[kBatch + 1];
  float acc[kBatch * kAcc];
  float4 gpix[8][4 * kRowG];       // dL/dimage of each warp's 32 pixels, rows padded (conflict-free phase-2 reads)
  float4 panel[8][kChunk * kRow];  // per-warp [splat][pixel] scratch
  unsigned short list[8][kBatch + kChunk];
  unsigned char mask[kBatch];
  int warp_done[8];
};

template <int F, bool GP, bool GF, bool HEUR>
__global__ void __launch_bounds__(kBatch, GS_BWDT_MIN_BLOCKS)
raster_bwd_t_kernel(const float *__restrict__ points, const float *__restrict__ features,
                    const int32_t *__restrict__ ranges, const int32_t *__restrict__ overlap_to_point,
                    const float *__restrict__ image, const float *__restrict__ grad_image, RasterParams<float> P,
                    float *__restrict__ grad_points, float *__restrict__ grad_features,
                    float *__restrict__ heuristic) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % P.tiles_wide) * kTile, tile_y0 = (tile / P.tiles_wide) * kTile;
  const int px = tile_x0 + (warp & 1) * 8 + (lane & 7), py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
  const bool in_bounds = px < P.width && py < P.height;
  // pixel centre relative to the tile centre: (tx, ty) = lx (ux, wx) + ly (uy, wy) + (tx0, ty0), two FFMA2
  const float lx = (float)((warp & 1) * 8 + (lane & 7)) - 7.5f, ly_pix = (float)((warp >> 1) * 4 + (lane >> 3)) - 7.5f;
  const f32x2 lx2 = pk(lx, lx), ly2 = pk(ly_pix, ly_pix);
  const float clamp_max = P.clamp_max, thr = P.thr;
  const float t_min = 1.0f - P.sat;   // a pixel is saturated (backward.py:131) once its transmittance is <= 1 - sat

  // The reference tracks remaining[c] = image[c] - sum_{j<=i} f_j[c] w_j per channel (backward.py:116-176), but it is
  // only ever used through its dot product with this pixel's dL/dimage, so one scalar carries the whole state:
  //   rem_dot = remaining . gpix ;  dL/dalpha = T (f . gpix) - rem_dot / (1 - alpha)
  float gpix[F];
#pragma unroll
  for (int c = 0; c < F; ++c) gpix[c] = 0.f;
  float rem_dot = 0.f;
  float trans = 0.f;                  // transmittance 1 - sum of weights; 0 outside the image: nothing contributes
  if (in_bounds) {
    const float *img = image + ((int64_t)py * P.width + px) * F;
    const float *gi = grad_image + ((int64_t)py * P.width + px) * F;
#pragma unroll
    for (int c = 0; c < F; ++c) { gpix[c] = gi[c]; rem_dot = fmaf(img[c], gpix[c], rem_dot); }
    trans = 1.0f;
  }
  {
    float4 gq = make_float4(gpix[0], F > 1 ? gpix[F > 1 ? 1 : 0] : 0.f, F > 2 ? gpix[F > 2 ? 2 : 0] : 0.f,
                            F > 3 ? gpix[F > 3 ? 3 : 0] : 0.f);
    sm.gpix[warp][(lane >> 3) * kRowG + (lane & 7)] = gq;
  }

  // phase-2 role of this lane: splat s of the chunk, pixel row q of the warp rectangle
  const int s = lane & 7, q = lane >> 3;
  const float bx = (float)((warp & 1) * 8) - 7.5f;            // tile-centred x of the row's first pixel
  const float ly = (float)((warp >> 1) * 4 + q) - 7.5f;       // tile-centred y of the row
  const int slot_base = ((lane & 16) ? 6 : 0) + ((lane & 8) ? 3 : 0);
  float4 *panel = sm.panel[warp];

  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  if (lane == 0) sm.warp_done[warp] = 0;
  if (tid == 0) {
    sm.a[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.b[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm.f[kBatch] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int base = start; base < end; base += kBatch) {
    const int nb = min(kBatch, end - base);
    __syncthreads();
    {
      int all_done = 1;
#pragma unroll
      for (int w = 0; w < 8; ++w) all_done &= sm.warp_done[w];
      if (all_done) break;
    }
    // ---- stage (thread j owns splat j of the batch, and flushes it at the end) ----
    int my_id = -1;
    float s_mx = 0.f, s_my = 0.f, s_ax = 0.f, s_ay = 0.f, s_isx = 0.f, s_isy = 0.f, s_alpha = 1.f;
    if (tid < nb) {
      my_id = overlap_to_point[base + tid];
      const float *g = points + 7 * (int64_t)my_id;
      float mx = g[0], my = g[1], ax = g[2], ay = g[3], sx = g[4], sy = g[5], alpha = g[6];
      float isx = 1.0f / sx, isy = 1.0f / sy;
      s_mx = mx; s_my = my; s_ax = ax; s_ay = ay; s_isx = isx; s_isy = isy; s_alpha = alpha;
      float ux = ax * isx * kExpScale, uy = ay * isx * kExpScale;
      float wx = -ay * isy * kExpScale, wy = ax * isy * kExpScale;
      {
        const float ddx = mx - ((float)tile_x0 + 8.0f), ddy = my - ((float)tile_y0 + 8.0f);
        sm.a[tid] = make_float4(-fmaf(ux, ddx, uy * ddy), -fmaf(wx, ddx, wy * ddy), ux, wx);
        sm.b[tid] = make_float4(uy, wy, alpha, 0.f);
      }
      unsigned mask = 0;
      if (alpha > thr) {
        float rc = sqrtf(2.0f * __logf(alpha / thr)) * 1.001f + 0.01f;
        float rcs = rc * kExpScale;
        float e1x = ax * sx, e1y = ay * sx, e2x = ay * sy, e2y = ax * sy;
        float ex = rc * sqrtf(e1x * e1x + e2x * e2x), ey = rc * sqrtf(e1y * e1y + e2y * e2y);
        float hu = fabsf(ux) * 3.5f + fabsf(uy) * 1.5f + rcs;
        float hw = fabsf(wx) * 3.5f + fabsf(wy) * 1.5f + rcs;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          float dcx = (float)(tile_x0 + (w & 1) * 8) + 4.0f - mx;
          float dcy = (float)(tile_y0 + (w >> 1) * 4) + 2.0f - my;
          bool hit = (fabsf(dcx) - 3.5f <= ex) && (fabsf(dcy) - 1.5f <= ey) &&
                     (fabsf(ux * dcx + uy * dcy) <= hu) && (fabsf(wx * dcx + wy * dcy) <= hw);
          mask |= hit ? (1u << w) : 0u;
        }
      }
      sm.mask[tid] = (unsigned char)mask;
      float4 fv = make_float4(0.f, 0.f, 0.f, 0.f);
      const float *fp = features + (int64_t)F * my_id;
      fv.x = fp[0];
      if (F > 1) fv.y = fp[1];
      if (F > 2) fv.z = fp[2];
      if (F > 3) fv.w = fp[3];
      sm.f[tid] = fv;
#pragma unroll
      for (int c = 0; c < 12; ++c) sm.acc[tid * kAcc + c] = 0.f;
    }
    __syncthreads();

    // ---- per-warp ordered hit list, padded to a multiple of the chunk with the null record ----
    int nhit = 0;
    if (!__all_sync(full, trans <= t_min)) {
      for (int c = 0; c < nb; c += 32) {
        int j = c + lane;
        bool hit = j < nb && ((sm.mask[j] >> warp) & 1);
        unsigned bal = __ballot_sync(full, hit);
        if (hit) sm.list[warp][nhit + __popc(bal & ((1u << lane) - 1))] = (unsigned short)j;
        nhit += __popc(bal);
      }
      if (lane < kChunk) sm.list[warp][nhit + lane] = (unsigned short)kBatch;
      __syncwarp();
    }
#ifdef GS_COUNT
    if (lane == 0) { atomicAdd(&g_count[1], (unsigned long long)nhit); if (warp == 0) atomicAdd(&g_count[3], 1ull); }
#endif

    for (int h0 = 0; h0 < nhit; h0 += kChunk) {
      // ---- phase 1: lane = pixel; 8 splats in depth order ----
      constexpr int kUnroll1 = GS_BWDT_UNROLL;
#pragma unroll kUnroll1
      for (int u = 0; u < kChunk; ++u) {
        const int j = sm.list[warp][h0 + u];
        const float4 A = sm.a[j], B = sm.b[j];
        const float4 fv = sm.f[j];
        const float feat[4] = {fv.x, fv.y, fv.z, fv.w};
        const f32x2 t2 = fma2(lx2, pk(A.z, A.w), fma2(ly2, pk(B.x, B.y), pk(A.x, A.y)));
        float tx, ty;
        upk(t2, tx, ty);
        float ga = ex2_approx(-(tx * tx + ty * ty));
        float alpha = B.z * ga;
        const bool has_grad = alpha > thr && trans > t_min;
        alpha = fminf(alpha, clamp_max);
        const float T_i = trans;
        float weight = has_grad ? alpha * T_i : 0.f;
        trans -= weight;
        float inv_1ma = rcp_approx(1.0f - alpha);
        float fg = feat[0] * gpix[0];
#pragma unroll
        for (int c = 1; c < F; ++c) fg = fmaf(feat[c], gpix[c], fg);
        rem_dot = fmaf(-weight, fg, rem_dot);
        float alpha_grad = fmaf(-rem_dot, inv_1ma, fg * T_i);
        float G = has_grad ? B.z * alpha_grad : 0.f;
        float Gp = G * ga;
        float h1 = 0.f;
        if (HEUR) {
          // |G dpdf/dmean|_1 = |Gp| (|tx ux + ty wx| + |tx uy + ty wy|) / k^2  (t and u, w carry one factor k each)
          const float inv_k2 = 1.0f / (kExpScale * kExpScale);
          float p0, p1, q0, q1;
          upk(mul2(t2, pk(A.z, A.w)), p0, p1);
          upk(mul2(t2, pk(B.x, B.y)), q0, q1);
          h1 = (fabsf(p0 + p1) + fabsf(q0 + q1)) * fabsf(Gp * inv_k2);
        }
        panel[u * kRow + lane] = make_float4(Gp, weight, G * G, h1);
#ifdef GS_COUNT
        {
          unsigned live = __ballot_sync(full, has_grad);
          if (lane == 0) { atomicAdd(&g_count[0], 1ull); atomicAdd(&g_count[2], (unsigned long long)__popc(live)); }
        }
#endif
      }
      __syncwarp();

      // ---- phase 2: lane = (splat s, pixel row q): walk the row's 8 pixels for one splat ----
      float m0 = 0.f, s1 = 0.f, s2 = 0.f, f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, hh0 = 0.f, hh1 = 0.f;
      f32x2 f01 = pk(0.f, 0.f), f23 = pk(0.f, 0.f), hh = pk(0.f, 0.f);
      const float4 *row = panel + s * kRow + q * 8;
      const float4 *grow = sm.gpix[warp] + q * kRowG;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = row[i];
        m0 += v.x;
        s1 = fmaf(v.x, (float)i, s1);
        s2 = fmaf(v.x, (float)(i * i), s2);
        if (GF) {
          const float4 g = grow[i];
          if (F == 1) f0 = fmaf(v.y, g.x, f0);
          if (F >= 2) f01 = fma2(pk(g.x, g.y), pk(v.y, v.y), f01);
          if (F == 3) f2 = fmaf(v.y, g.z, f2);
          if (F == 4) f23 = fma2(pk(g.z, g.w), pk(v.y, v.y), f23);
        }
        if (HEUR) hh = add2(hh, pk(v.z, v.w));
      }
      if (F >= 2) upk(f01, f0, f1);
      if (F == 4) upk(f23, f2, f3);
      if (HEUR) upk(hh, hh0, hh1);
      // row-local -> tile-centred moments (x = bx + i, y = ly)
      const float Lx = fmaf(bx, m0, s1);
      const float Lxx = fmaf(bx, fmaf(bx, m0, 2.0f * s1), s2);
      float v[12] = {m0, Lx, ly * m0, Lxx, ly * Lx, ly * ly * m0, f0, f1, f2, f3, hh0, hh1};
      // sum over the 4 rows with a 2-stage transposed butterfly: 6 + 3 shuffles, 3 finished sums per lane
      {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          float send = up ? v[i] : v[i + 6];
          float keep = up ? v[i + 6] : v[i];
          v[i] = keep + __shfl_xor_sync(full, send, 16);
        }
      }
      {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float send = up ? v[i] : v[i + 3];
          float keep = up ? v[i + 3] : v[i];
          v[i] = keep + __shfl_xor_sync(full, send, 8);
        }
      }
      if (h0 + s < nhit) {
        float *dst = sm.acc + sm.list[warp][h0 + s] * kAcc + slot_base;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (v[i] != 0.f) atomicAdd(dst + i, v[i]);
      }
      __syncwarp();
      if (__all_sync(full, trans <= t_min)) break;
    }
    if (__all_sync(full, trans <= t_min) && lane == 0) sm.warp_done[warp] = 1;

    // ---- flush: one thread per staged splat ----
    __syncthreads();
    if (tid < nb) {
      float S[12];
      bool any = false;
#pragma unroll
      for (int c = 0; c < 12; ++c) { S[c] = sm.acc[tid * kAcc + c]; any |= (S[c] != 0.f); }
      if (any) {
        if (GP) {
          // shift the tile-centred moments to the splat mean: d = l + c
          const float cx = (float)tile_x0 + 8.0f - s_mx, cy = (float)tile_y0 + 8.0f - s_my;
          const float M0 = S[0], Lx = S[1], Ly = S[2], Lxx = S[3], Lxy = S[4], Lyy = S[5];
          const float Mx = fmaf(cx, M0, Lx), My = fmaf(cy, M0, Ly);
          const float Mxx = Lxx + cx * (2.0f * Lx + cx * M0);
          const float Myy = Lyy + cy * (2.0f * Ly + cy * M0);
          const float Mxy = Lxy + cx * Ly + cy * Lx + cx * cy * M0;
          const float ux = s_ax * s_isx, uy = s_ay * s_isx, wx = -s_ay * s_isy, wy = s_ax * s_isy;
          const float S1 = ux * Mx + uy * My, S2 = wx * Mx + wy * My;
          const float S3 = ux * Mxx + uy * Mxy, S4 = ux * Mxy + uy * Myy;
          const float S5 = wx * Mxx + wy * Mxy, S6 = wx * Mxy + wy * Myy;
          float *gp = grad_points + 7 * (int64_t)my_id;
          atomicAdd(gp + 0, S1 * ux + S2 * wx);
          atomicAdd(gp + 1, S1 * uy + S2 * wy);
          atomicAdd(gp + 2, -s_isx * S3 - s_isy * S6);
          atomicAdd(gp + 3, -s_isx * S4 + s_isy * S5);
          atomicAdd(gp + 4, s_isx * (ux * S3 + uy * S4));
          atomicAdd(gp + 5, s_isy * (wx * S5 + wy * S6));
          atomicAdd(gp + 6, M0 / s_alpha);
        }
        if (GF) {
          float *gf = grad_features + (int64_t)F * my_id;
#pragma unroll
          for (int c = 0; c < F; ++c) atomicAdd(gf + c, S[6 + c]);
        }
        if (HEUR) {
          atomicAdd(heuristic + 2 * (int64_t)my_id, S[10]);
          atomicAdd(heuristic + 2 * (int64_t)my_id + 1, S[11]);
        }
      }
    }
  }
}

}  // namespace bwdt

template <int F>
int launch_bwd_transpose(const float *points, const float *features, const int32_t *ranges, const int32_t *o2p,
                         const float *image, const float *grad_image, const RasterParams<float> &P, int tiles,
                         float *grad_points, float *grad_features, float *heuristic, cudaStream_t stream) {
  const bool gp = grad_points != nullptr, gf = grad_features != nullptr, he = P.heur && heuristic != nullptr;
  const size_t smem = sizeof(bwdt::Smem);
#define GS_BWDT(GP_, GF_, HE_)                                                                                  \
  do {                                                                                                          \
    auto kern = bwdt::raster_bwd_t_kernel<F, GP_, GF_, HE_>;                                                    \
    static bool configured = false;                                                                             \
    if (!configured) {                                                                                          \
      GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
      configured = true;                                                                                        \
    }                                                                                                           \
    kern<<<tiles, bwdt::kBatch, smem, stream>>>(points, features, ranges, o2p, image, grad_image, P,            \
                                                grad_points, grad_features, heuristic);                         \
  } while (0)
  if (gp && gf && he) GS_BWDT(true, true, true);
  else if (gp && gf) GS_BWDT(true, true, false);
  else if (gp && he) GS_BWDT(true, false, true);
  else if (gp) GS_BWDT(true, false, false);
  else if (gf && he) GS_BWDT(false, true, true);
  else if (gf) GS_BWDT(false, true, false);
  else if (he) GS_BWDT(false, false, true);
  else return GS_OK;
#undef GS_BWDT
  GS_LAUNCH_CHECK();
  return GS_OK;
}

template int launch_bwd_transpose<1>(const float *, const float *, const int32_t *, const int32_t *, const float *,
                                     const float *, const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);
template int launch_bwd_transpose<2>(const float *, const float *, const int32_t *, const int32_t *, const float *,
                                     const float *, const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);
template int launch_bwd_transpose<3>(const float *, const float *, const int32_t *, const int32_t *, const float *,
                                     const float *, const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);
template int launch_bwd_transpose<4>(const float *, const float *, const int32_t *, const int32_t *, const float *,
                                     const float *, const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);

}  // namespace gs

#ifdef GS_COUNT
extern "C" int gs_debug_counters(unsigned long long *out4, int reset) {
  cudaMemcpyFromSymbol(out4, gs::bwdt::g_count, sizeof(unsigned long long) * 4);
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(gs::bwdt::g_count, z, sizeof z); }
  return 0;
}
#endif
