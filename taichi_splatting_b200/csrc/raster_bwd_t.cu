// raster_bwd_t.cu -- R9: the tuned backward kernel (fp32, 16x16 tiles, plain pdf, 1..4 features).
//
// Semantics: _backward_kernel, rasterizer/backward.py:50-225 (C ABI and dispatch: raster_bwd.cu).  Same tile / 8x8
// block per warp / two pixels per lane / digest / hit-list structure as raster_fwd.cu, and like it bound by the
// L1/shared-memory data pipe (DESIGN.md section 4 has the measurements).
//
// Per (pixel, splat) only six tile-local moments of Gp = alpha * dL/dalpha * pdf are needed, {1, lx, ly, lx^2, lx ly,
// ly^2} * Gp (l = pixel centre - tile centre), plus weight * dL/dimage and the two densification heuristics: every
// parameter gradient is linear in those sums (the pdf is exp of a quadratic form in the pixel position), so
// mean / axis / sigma enter once per (splat, tile) at flush time and dL/dalpha_point = M0 / alpha_point.  The sums
// over the 64 pixels of a block are formed by a shared-memory transpose:
//
//   phase 1  sweeps 8 splats of the hit list (lane = two pixels sharing a column), pre-sums the pair and parks
//            {S = Gp0 + Gp1, D = Gp1, sum G^2, sum |G dpdf/dmean|_1} and {sum_pixels weight * dL/dimage[c]} per lane
//            in two padded panel planes [splat][lane] (conflict-free STS.128); the NEXT splat's records are loaded
//            before the current one is computed, so their latency hides behind the arithmetic;
//   phase 2  re-reads the planes transposed -- lane = (splat s, row pair q) walks the 8 lanes of its row pair with
//            the column index an immediate (sum S, sum i S, sum i^2 S, sum D, sum i D as FADD2 / FFMA2) -- then a
//            2-stage transposed butterfly over the four row pairs (9 shuffles per 8 splats) leaves 3 finished sums
//            per lane, added to the batch accumulators with 3 conflict-free shared atomics per 8 splats.
//
// Per splat iteration the data pipe is charged 3 x 2.69 (records) + 0.7 (list) + 2 x 4.04 (panel stores) + 2 x 4.04
// (panel loads) + ~2 (shuffles, atomics) = ~27 cycles; the kernel runs at 81 % of that pipe.
//
// Staging: the tile's records come from raster_pack.cu (sorted order, tile-centred, block mask included), one
// `cp.async.bulk` per batch of 128 into one of two shared-memory buffers, signalled on an mbarrier (bulk_copy.cuh);
// batch b+1 is in flight while batch b is swept.  The reference stages with a synchronous cooperative gather
// (backward.py:100-118); the first version of this kernel gathered 64-byte digests and classified the blocks itself.
#include <atomic>

#include "bulk_copy.cuh"
#include "packed_f32.cuh"
#include "raster_common.cuh"

namespace gs {

#ifndef GS_EXACT_CULL
#define GS_EXACT_CULL 1   // exact block-vs-ellipse test after the box tests: 7.72 -> 7.33 hit entries per Gaussian (1.131 -> 1.118 ms)
#endif
#ifndef GS_BWDT_UNROLL
#define GS_BWDT_UNROLL 8
#endif
#ifndef GS_BWDT_MIN_BLOCKS
#define GS_BWDT_MIN_BLOCKS 4
#endif
#ifndef GS_BWDT_PAIRED
#define GS_BWDT_PAIRED 1   // lane pairs pre-add through shuffles, one panel plane (0: two planes, the round-1 form)
#endif

namespace bwdt {

#ifdef GS_COUNT   // debug build only: work counters read by profiles/count_work.py
__device__ unsigned long long g_count[4];   // warp iterations (padded), hit-list entries, live (pixel, splat) lanes, batches
#endif

constexpr int kTile = 16;
constexpr int kWarps = 4;          // one warp per 8x8 pixel block; every lane owns two pixels (rows r and r + 4)
constexpr int kThreads = kWarps * 32;
constexpr int kBatch = kThreads;   // splats staged per round: thread j stages and finally flushes splat j
#ifndef GS_BWDT_CHUNK
#define GS_BWDT_CHUNK 8
#endif
constexpr int kChunk = GS_BWDT_CHUNK;   // splats per phase-1 / phase-2 round: 8 | 4 (4: half the panel, more CTAs per SM)
static_assert(kChunk == 8 || kChunk == 4, "phase 2 is written for chunks of 8 or 4 splats");
constexpr int kRow = 33;           // panel row stride in float4 (32 lanes + 1 pad: conflict-free transposed reads)
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

// Feature part of a raster_pack record behind Q1 = {uy, wy, alpha, f0}: {f1, f2, depth, mask} (1..3 features) |
// {f1, f2, f3, depth} (4 features).  The backward never needs the depth: 0, 4, 8 or 16 bytes are read.
template <int F>
__device__ __forceinline__ float4 load_tail(const unsigned char *q2) {
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (F == 4) {
    t = *reinterpret_cast<const float4 *>(q2);
  } else if (F == 3) {
    const float2 h = *reinterpret_cast<const float2 *>(q2);
    t.x = h.x; t.y = h.y;
  } else if (F == 2) {
    t.x = *reinterpret_cast<const float *>(q2);
  }
  return t;
}

// lane pairs pre-add through shuffles and one panel plane carries everything (see the kernel)
constexpr bool paired_panel(int F, bool GF) { return GS_BWDT_PAIRED && GF && F <= 3 && kChunk == 8; }
constexpr int panel_planes(int F, bool GF) { return GF && !paired_panel(F, GF) ? 2 : 1; }
// accumulator stride (odd: conflict-free).  Two planes: 6 moments, 4 features, 2 heuristics, 1 pad; paired: 6 moments,
// sum G^2, sum |G dpdf/dmean|, 3 features
constexpr int acc_stride(int F, bool GF) { return paired_panel(F, GF) ? 11 : 13; }
// resident CTAs per SM the shared memory allows (37 KB | 38 KB | 55 KB per CTA); the register budget follows
constexpr int min_blocks(int F, bool GF) {
  return GS_BWDT_MIN_BLOCKS != 4 ? GS_BWDT_MIN_BLOCKS : paired_panel(F, GF) ? 6 : GF ? 4 : 5;
}

template <int RECW, int PLANES, int ACC>
struct Smem {
  // two landing buffers of raster_pack records {tx0, ty0, ux, wx | uy, wy, alpha, f0 | f1, f2, (f3,) depth, mask}
  // (+1: null record that pads the hit lists)
  float4 rec[2][(kBatch + 1) * RECW];
  float acc[kBatch * ACC];
  // per-warp [splat][lane] scratch.  Two planes: {S, D, sum G^2, sum |G dpdf/dmean|} and {sum weight dL/dimage[c]};
  // one plane: without feature gradients, or what the lane pairs left (paired_panel)
  float4 panel[PLANES][kWarps][kChunk * kRow];
  // byte offsets (16 RECW j) of the records a warp must visit; entry k lives at index k + 3, so that the eight
  // "next" entries of a chunk (k = h0 + 1 .. h0 + 8) are two aligned 16-byte loads
  alignas(16) unsigned list[kWarps][kBatch + kChunk + 4];
  alignas(8) uint64_t full[2];           // mbarriers: "buffer b holds its batch"
};

template <int F, bool GP, bool GF, bool HEUR, int RECW>
__global__ void __launch_bounds__(kThreads, min_blocks(F, GF))
raster_bwd_t_kernel(const float4 *__restrict__ records, const float4 *__restrict__ flush_records,
                    const int32_t *__restrict__ ranges, const int32_t *__restrict__ overlap_to_point,
                    const float *__restrict__ image, const float *__restrict__ grad_image, RasterParams<float> P,
                    float *__restrict__ grad_points, float *__restrict__ grad_features,
                    float *__restrict__ heuristic) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SmemT = Smem<RECW, panel_planes(F, GF), acc_stride(F, GF)>;
  SmemT &sm = *reinterpret_cast<SmemT *>(smem_raw);
  constexpr int kAcc = acc_stride(F, GF);
  // One panel plane instead of two (feature gradients of up to 3 channels, chunks of 8): lanes l and l ^ 16 -- same
  // column, rows {r, r + 4} and {r + 2, r + 6} -- first add up through 4 shuffles, the lower lane keeping the three
  // column moments {sum Gp, sum y Gp, sum y^2 Gp} and sum G^2, the upper one sum |G dpdf/dmean| and the 3 feature
  // sums, so a splat costs the shared-memory data pipe one STS.128 + one LDS.128 per lane instead of two each
  // (a shuffle is 0.58 cycles of that pipe, an STS.128 / LDS.128 4.04: profiles/r01p_micro.txt).  Without the
  // feature-gradient plane the kernel ran 1.01 -> 0.82 ms (profiles/r02/r02v_bwd_variants.txt).
  constexpr bool kPaired = paired_panel(F, GF);
  // accumulator slots per splat: {M0, Lx, Ly, Lxx, Lxy, Lyy} + features + heuristics
  constexpr int kSlots = kPaired ? 11 : 12;
  constexpr int kSlotF = kPaired ? 8 : 6, kSlotH0 = kPaired ? 6 : 10, kSlotH1 = kPaired ? 7 : 11;
  constexpr unsigned kRecBytes = 16u * RECW;
  constexpr int kMaskWord = RECW == 3 ? 11 : 12;
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int tile_x0 = (tile % P.tiles_wide) * kTile, tile_y0 = (tile / P.tiles_wide) * kTile;
  const int bx = (warp & 1) * 8 + (lane & 7), by = (warp >> 1) * 8 + (lane >> 3);
  const int px = tile_x0 + bx, py[2] = {tile_y0 + by, tile_y0 + by + 4};
  const bool in_bounds[2] = {px < P.width && py[0] < P.height, px < P.width && py[1] < P.height};
  // pixel centre relative to the tile centre: (tx, ty) = lx (ux, wx) + ly (uy, wy) + (tx0, ty0)
  const float lx = (float)bx - 7.5f, ly0 = (float)by - 7.5f;
  const f32x2 lx2 = pk(lx, lx), lyp = pk(ly0, ly0 + 4.0f);   // the lane's column; the rows of its two pixels
  // paired mode: the lane's rows enter the y moments here, y Gp0 + (y + 4) Gp1 = y S + 4 D and
  // y^2 Gp0 + (y + 4)^2 Gp1 = y^2 S + (8 y + 16) D
  const f32x2 ymom = pk(ly0, ly0 * ly0), dmom = pk(4.0f, fmaf(8.0f, ly0, 16.0f));
  const bool lower = (lane & 16) == 0;
  const float clamp_max = P.clamp_max, thr = P.thr;
  const float t_min = 1.0f - P.sat;   // a pixel is saturated (backward.py:131) once its transmittance is <= 1 - sat

  // Per-pixel state, packed over the lane's two pixels.  The reference tracks remaining[c] = image[c] -
  // sum_{j<=i} f_j[c] w_j per channel (backward.py:116-176), but it is only ever used through its dot product with
  // this pixel's dL/dimage, so one scalar carries it:  rneg = -(remaining . gpix);
  //   dL/dalpha = T (f . gpix) + rneg / (1 - alpha)
  f32x2 gpix2[F];                     // (dL/dimage[c] of pixel 0, of pixel 1)
  float rneg[2] = {0.f, 0.f};
  float trans[2] = {0.f, 0.f};        // transmittance 1 - sum of weights; 0 outside the image: nothing contributes
  {
    float g[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      if (in_bounds[p]) {
        const float *img = image + ((int64_t)py[p] * P.width + px) * F;
        const float *gi = grad_image + (int64_t)py[p] * P.gs_y + (int64_t)px * P.gs_x;   // any strides (expanded, CHW, ...)
#pragma unroll
        for (int c = 0; c < F; ++c) { g[p][c] = gi[c * P.gs_c]; rneg[p] = fmaf(-img[c], g[p][c], rneg[p]); }
        trans[p] = 1.0f;
      }
    }
#pragma unroll
    for (int c = 0; c < F; ++c) gpix2[c] = pk(g[0][c], g[1][c]);
  }

  // phase-2 role of this lane: splat s of the chunk, pixel rows q and q + 4 of the warp's block
  // phase-2 role of this lane.  Chunk 8: splat s = lane & 7, row pair q = lane >> 3 (8 lane entries).  Chunk 4: splat
  // s = lane & 3, half a row pair q = lane >> 2 (4 lane entries: columns 4 (q & 1) .. + 3 of row pair q >> 1).
  const int s = kChunk == 8 ? (lane & 7) : (lane & 3), q = kChunk == 8 ? (lane >> 3) : (lane >> 2);
  const int entries = kChunk == 8 ? 8 : 4;
  const float bx0 = (float)((warp & 1) * 8 + (kChunk == 8 ? 0 : 4 * (q & 1))) - 7.5f;   // tile-centred x of the first entry
  const float lya = (float)((warp >> 1) * 8 + (kChunk == 8 ? q : (q >> 1))) - 7.5f;   // y of the upper row (other: + 4)
  const int slot_base = kChunk == 8 ? ((lane & 16) ? 6 : 0) + ((lane & 8) ? 3 : 0)
                                    : ((lane & 16) ? 6 : 0) + ((lane & 8) ? 3 : 0);
  float4 *panel0 = sm.panel[0][warp], *panel1 = sm.panel[panel_planes(F, GF) - 1][warp];
  float g_scalar[2][4];               // dL/dimage of the two pixels as scalars (same registers as gpix2)
#pragma unroll
  for (int c = 0; c < 4; ++c) { g_scalar[0][c] = 0.f; g_scalar[1][c] = 0.f; }
#pragma unroll
  for (int c = 0; c < F; ++c) upk(gpix2[c], g_scalar[0][c], g_scalar[1][c]);

  const int start = ranges[2 * tile], end = ranges[2 * tile + 1];
  const int nbatches = (end - start + kBatch - 1) / kBatch;

  auto issue = [&](int b) {   // thread 0 only: arm the barrier with the byte count, hand the range to the copy engine
    const int base = start + b * kBatch;
    const uint32_t bytes = (uint32_t)min(kBatch, end - base) * kRecBytes;
    mbar_arrive_expect_tx(&sm.full[b & 1], bytes);
    bulk_copy_g2s(sm.rec[b & 1], records + (int64_t)RECW * base, bytes, &sm.full[b & 1]);
  };

  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_fence_init();
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int q = 0; q < RECW; ++q) sm.rec[b][kBatch * RECW + q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nbatches > 0) issue(0);
    if (nbatches > 1) issue(1);
  }
  int warp_saturated = 0;   // every pixel of this warp's block is saturated (uniform over the warp)
#pragma unroll
  for (int c = 0; c < kSlots; ++c) sm.acc[tid * kAcc + c] = 0.f;

  for (int b = 0; b < nbatches; ++b) {
    const int buf = b & 1;
    const int base = start + b * kBatch, nb = min(kBatch, end - base);
    // previous flush finished: its buffer is free, accumulators are zero again.  The barrier also decides, identically
    // for every thread, whether the whole tile is saturated (a shared flag per warp, re-read after the barrier while
    // fast warps already set theirs again, let warps disagree: compute-sanitizer racecheck).
    {
      const int all_done = __syncthreads_and(warp_saturated);
      // refill the buffer the previous batch just released (the flush still read its records, so not earlier); the
      // copy of batch b + 1 then runs beside the sweep of batch b.  Every issued copy is waited for before the CTA
      // exits: when every pixel is saturated nothing new is issued and batch b is the last one in flight.
      if (tid == 0 && b >= 1 && b + 1 < nbatches && !all_done) issue(b + 1);
      mbar_wait(&sm.full[buf], (uint32_t)(b >> 1) & 1u);   // the batch has landed
      if (all_done) break;
    }
    const unsigned char *rec = reinterpret_cast<const unsigned char *>(sm.rec[buf]);
    const unsigned *words = reinterpret_cast<const unsigned *>(sm.rec[buf]);

    // ---- per-warp ordered hit list, padded to a multiple of the chunk with the null record ----
    int nhit = 0;
    if (!__all_sync(full, trans[0] <= t_min && trans[1] <= t_min)) {
      for (int c = 0; c < nb; c += 32) {
        int j = c + lane;
        bool hit = j < nb && ((words[j * (4 * RECW) + kMaskWord] >> warp) & 1u);
        unsigned bal = __ballot_sync(full, hit);
        if (hit) sm.list[warp][3 + nhit + __popc(bal & ((1u << lane) - 1))] = kRecBytes * (unsigned)j;
        nhit += __popc(bal);
      }
    }
    // pad with the null record (also when the list is empty: the sweep prefetches entry 0 unconditionally)
    if (lane <= kChunk) sm.list[warp][3 + nhit + lane] = kRecBytes * (unsigned)kBatch;
    __syncwarp();
#ifdef GS_COUNT
    if (lane == 0) { atomicAdd(&g_count[1], (unsigned long long)nhit); if (warp == 0) atomicAdd(&g_count[3], 1ull); }
#endif

    // records of the first hit: every iteration loads the NEXT splat's records before it computes (the loads
    // then sit ahead of the panel stores in program order, so their latency hides behind the arithmetic)
    float4 A, B, fv;
    {
      const unsigned off = sm.list[warp][3];
      A = *reinterpret_cast<const float4 *>(rec + off);
      B = *reinterpret_cast<const float4 *>(rec + off + 16);
      fv = load_tail<F>(rec + off + 32);
    }
    for (int h0 = 0; h0 < nhit; h0 += kChunk) {
      // ---- phase 1: lane = two pixels; 8 splats in depth order ----
      const uint4 nx0 = *reinterpret_cast<const uint4 *>(&sm.list[warp][h0 + 4]);
      const uint4 nx1 = kChunk == 8 ? *reinterpret_cast<const uint4 *>(&sm.list[warp][h0 + 8]) : nx0;
      const unsigned next_off[8] = {nx0.x, nx0.y, nx0.z, nx0.w, nx1.x, nx1.y, nx1.z, nx1.w};
      constexpr int kUnroll1 = GS_BWDT_UNROLL;
#pragma unroll kUnroll1
      for (int u = 0; u < kChunk; ++u) {
        const unsigned off_next = next_off[u];
        const float4 An = *reinterpret_cast<const float4 *>(rec + off_next);
        const float4 Bn = *reinterpret_cast<const float4 *>(rec + off_next + 16);
        const float4 fn = load_tail<F>(rec + off_next + 32);
        const float feat[4] = {B.w, fv.x, fv.y, fv.z};
        // t = lx (ux, wx) + ly (uy, wy) + t0, kept component-major over the pixel pair -- tx2 = (tx of pixel 0, of
        // pixel 1) -- so that |t|^2 of both pixels is one FMUL2 + one FFMA2 (scalar operands broadcast for free)
        float cx, cy;
        upk(fma2(lx2, pk(A.z, A.w), pk(A.x, A.y)), cx, cy);
        const f32x2 tx2 = fma2(pk(B.x, B.x), lyp, pk(cx, cx)), ty2 = fma2(pk(B.y, B.y), lyp, pk(cy, cy));
        float q0, q1;
        upk(fma2(tx2, tx2, mul2(ty2, ty2)), q0, q1);
        const float g0 = ex2_approx(-q0), g1 = ex2_approx(-q1);
        const f32x2 ga2 = pk(g0, g1), alpha_pt2 = pk(B.z, B.z);
        float a0, a1;
        upk(mul2(ga2, alpha_pt2), a0, a1);
        const bool hg0 = a0 > thr && trans[0] > t_min, hg1 = a1 > thr && trans[1] > t_min;
        a0 = fminf(a0, clamp_max);
        a1 = fminf(a1, clamp_max);
        const f32x2 a2 = pk(a0, a1), T2 = pk(trans[0], trans[1]);
        float w0, w1;
        upk(mul2(a2, T2), w0, w1);
        w0 = hg0 ? w0 : 0.f;
        w1 = hg1 ? w1 : 0.f;
        const f32x2 w2 = pk(w0, w1);
        upk(sub2(T2, w2), trans[0], trans[1]);
        float om0, om1;
        upk(sub2(pk(1.0f, 1.0f), a2), om0, om1);
        const f32x2 inv2 = pk(rcp_approx(om0), rcp_approx(om1));
        f32x2 fg2 = mul2(pk(feat[0], feat[0]), gpix2[0]);
#pragma unroll
        for (int c = 1; c < F; ++c) fg2 = fma2(pk(feat[c], feat[c]), gpix2[c], fg2);
        const f32x2 rn2 = fma2(w2, fg2, pk(rneg[0], rneg[1]));
        upk(rn2, rneg[0], rneg[1]);
        float G0, G1;
        upk(mul2(alpha_pt2, fma2(rn2, inv2, mul2(fg2, T2))), G0, G1);
        G0 = hg0 ? G0 : 0.f;
        G1 = hg1 ? G1 : 0.f;
        float Gp0, Gp1;
        const f32x2 Gp2 = mul2(pk(G0, G1), ga2);
        upk(Gp2, Gp0, Gp1);
        // the pixel pair shares its column, so the pair enters the panel pre-summed: S = Gp0 + Gp1 carries the
        // 1, x, x^2 moments and D = Gp1 (the pixel 4 rows down) completes the y moments
        float4 e0 = make_float4(Gp0 + Gp1, Gp1, 0.f, 0.f);
        if (kPaired) {   // {S, y S + 4 D, y^2 S + (8 y + 16) D, sum G^2}
          upk(fma2(ymom, pk(e0.x, e0.x), mul2(dmom, pk(Gp1, Gp1))), e0.y, e0.z);
          e0.w = 0.f;
        }
        float heur_g2 = 0.f, heur_k = 0.f;
        if (HEUR) {
          // |G dpdf/dmean|_1 = |Gp| (|tx ux + ty wx| + |tx uy + ty wy|) / k^2  (t and u, w carry one factor k each;
          // the 1 / k^2 is applied once per splat in phase 2)
          float dx0, dx1, dy0, dy1, hk0, hk1;
          upk(fma2(pk(A.z, A.z), tx2, mul2(pk(A.w, A.w), ty2)), dx0, dx1);   // tx ux + ty wx of both pixels
          upk(fma2(pk(B.x, B.x), tx2, mul2(pk(B.y, B.y), ty2)), dy0, dy1);   // tx uy + ty wy
          upk(mul2(pk(fabsf(dx0) + fabsf(dy0), fabsf(dx1) + fabsf(dy1)), Gp2), hk0, hk1);
          heur_g2 = fmaf(G0, G0, G1 * G1);
          heur_k = fabsf(hk0) + fabsf(hk1);
        }
        if (kPaired) {
          const float4 mom = make_float4(e0.x, e0.y, e0.z, heur_g2);
          const float4 pln = make_float4(heur_k, fmaf(w0, g_scalar[0][0], w1 * g_scalar[1][0]),
                                         F > 1 ? fmaf(w0, g_scalar[0][1], w1 * g_scalar[1][1]) : 0.f,
                                         F > 2 ? fmaf(w0, g_scalar[0][2], w1 * g_scalar[1][2]) : 0.f);
          float4 out;
          out.x = (lower ? mom.x : pln.x) + __shfl_xor_sync(full, lower ? pln.x : mom.x, 16);
          out.y = (lower ? mom.y : pln.y) + __shfl_xor_sync(full, lower ? pln.y : mom.y, 16);
          out.z = (lower ? mom.z : pln.z) + __shfl_xor_sync(full, lower ? pln.z : mom.z, 16);
          out.w = (lower ? mom.w : pln.w) + __shfl_xor_sync(full, lower ? pln.w : mom.w, 16);
          panel0[u * kRow + lane] = out;
        } else {
          e0.z = heur_g2;
          e0.w = heur_k;
          panel0[u * kRow + lane] = e0;
        }
        if (GF && !kPaired) {
          float4 e1 = make_float4(0.f, 0.f, 0.f, 0.f);
          e1.x = fmaf(w0, g_scalar[0][0], w1 * g_scalar[1][0]);
          if (F > 1) e1.y = fmaf(w0, g_scalar[0][1], w1 * g_scalar[1][1]);
          if (F > 2) e1.z = fmaf(w0, g_scalar[0][2], w1 * g_scalar[1][2]);
          if (F > 3) e1.w = fmaf(w0, g_scalar[0][3], w1 * g_scalar[1][3]);
          panel1[u * kRow + lane] = e1;
        }
#ifdef GS_COUNT
        {
          unsigned live = __ballot_sync(full, hg0), live1 = __ballot_sync(full, hg1);
          if (lane == 0) { atomicAdd(&g_count[0], 1ull); atomicAdd(&g_count[2], (unsigned long long)(__popc(live) + __popc(live1))); }
        }
#endif
        A = An; B = Bn; fv = fn;
      }
      __syncwarp();

      if (kPaired) {
        // ---- phase 2, one plane: lane = (splat s, role q).  q = 0, 1: the column moments of the even / odd rows of the
        // block (entries of lanes 8 q .. 8 q + 7); q = 2, 3: the plain sums {|G dpdf/dmean|, features} of the same rows.
        // Every lane runs the same arithmetic over its 8 entries (entry i = column i of the block); the roles differ in
        // what the sums mean.
        f32x2 a01 = pk(0.f, 0.f), a23 = pk(0.f, 0.f), i01 = pk(0.f, 0.f);
        float ii0 = 0.f;
        {
          const float4 *row = panel0 + s * kRow + q * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 e = row[i];
            const f32x2 xy = pk(e.x, e.y);
            a01 = add2(a01, xy);                                       // (sum M0, sum My)      | (sum hk, sum f0)
            a23 = add2(a23, pk(e.z, e.w));                             // (sum Myy, sum G^2)    | (sum f1, sum f2)
            i01 = fma2(xy, pk((float)i, (float)i), i01);               // (sum i M0, sum i My)
            ii0 = fmaf(e.x, (float)(i * i), ii0);                      // sum i^2 M0
          }
        }
        float m0, my, s1, s1y, z0, z1;
        upk(a01, m0, my);
        upk(i01, s1, s1y);
        upk(a23, z0, z1);
        // column sums -> tile-centred moments (x = bx0 + i; the rows entered tile-centred in phase 1)
        const float Lx = fmaf(bx0, m0, s1);
        const float Lxx = fmaf(bx0, fmaf(bx0, m0, 2.0f * s1), ii0);
        const float Lxy = fmaf(bx0, my, s1y);
        const bool moments = q < 2;
        // moment role: {M0, Lx, Ly, Lxx | Lxy, Lyy, sum G^2, -};  plain role: {hk, f0, f1, f2 | -, -, -, -}
        float v[8] = {m0, moments ? Lx : my, moments ? my : z0, moments ? Lxx : z1, Lxy, z0, z1, 0.f};
        {   // sum over the two row parities: q = 0 and 2 end with v[0..3], q = 1 with v[4..7] (q = 3: nothing)
          const bool up = (lane & 8) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float send = up ? v[i] : v[i + 4];
            float keep = up ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(full, send, 8);
          }
        }
        if (h0 + s < nhit && q < 3) {
          // slots 0..3 | 4..6 (the lane's fourth sum is an exact zero) | 7..10
          float *dst = sm.acc + (sm.list[warp][3 + h0 + s] / kRecBytes) * kAcc + (q == 2 ? 7 : 4 * q);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (v[i] != 0.f) atomicAdd(dst + i, v[i]);
        }
      } else {
        // ---- phase 2: lane = (splat s, row pair q): walk the 8 lane entries of rows q and q + 4 for one splat ----
        f32x2 md = pk(0.f, 0.f), sd1 = pk(0.f, 0.f), hh = pk(0.f, 0.f), f01 = pk(0.f, 0.f), f23 = pk(0.f, 0.f);
        float s2 = 0.f;
        {
          const float4 *row0 = panel0 + s * kRow + q * entries, *row1 = panel1 + s * kRow + q * entries;
  #pragma unroll
          for (int i = 0; i < entries; ++i) {
            const float4 e0 = row0[i];
            const f32x2 sd = pk(e0.x, e0.y);
            md = add2(md, sd);                                         // (sum S, sum D)
            sd1 = fma2(sd, pk((float)i, (float)i), sd1);               // (sum i S, sum i D)
            s2 = fmaf(e0.x, (float)(i * i), s2);
            if (HEUR) hh = add2(hh, pk(e0.z, e0.w));
            if (GF) {
              const float4 e1 = row1[i];
              f01 = add2(f01, pk(e1.x, e1.y));
              if (F > 2) f23 = add2(f23, pk(e1.z, e1.w));
            }
          }
        }
        float m0, d0, s1, d1, f0, f1, f2, f3, hh0, hh1;
        upk(md, m0, d0);
        upk(sd1, s1, d1);
        upk(f01, f0, f1);
        upk(f23, f2, f3);
        upk(hh, hh0, hh1);
        hh1 *= 1.0f / (kExpScale * kExpScale);
        // lane-entry sums -> tile-centred moments (x = bx0 + i; y = lya for S - D, lya + 4 for D)
        const float Lx = fmaf(bx0, m0, s1);
        const float Lxx = fmaf(bx0, fmaf(bx0, m0, 2.0f * s1), s2);
        const float Ly = fmaf(lya, m0, 4.0f * d0);
        const float Lxy = fmaf(lya, Lx, 4.0f * fmaf(bx0, d0, d1));
        const float Lyy = fmaf(lya * lya, m0, fmaf(8.0f, lya, 16.0f) * d0);
        float v[12] = {m0, Lx, Ly, Lxx, Lxy, Lyy, f0, f1, f2, f3, hh0, hh1};
        // sum over the 4 row pairs with a 2-stage transposed butterfly: 6 + 3 shuffles, 3 finished sums per lane
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
        if (kChunk == 4) {   // third stage: the two halves of a row pair (both lanes end with the sums; one adds them)
  #pragma unroll
          for (int i = 0; i < 3; ++i) v[i] += __shfl_xor_sync(full, v[i], 4);
        }
        if (h0 + s < nhit && (kChunk == 8 || (lane & 4) == 0)) {
          float *dst = sm.acc + (sm.list[warp][3 + h0 + s] / kRecBytes) * kAcc + slot_base;
  #pragma unroll
          for (int i = 0; i < 3; ++i)
            if (v[i] != 0.f) atomicAdd(dst + i, v[i]);
        }
      }
      __syncwarp();
      if (__all_sync(full, trans[0] <= t_min && trans[1] <= t_min)) break;
    }
    warp_saturated = __all_sync(full, trans[0] <= t_min && trans[1] <= t_min);

    // ---- flush: one thread per splat of the batch ----
    __syncthreads();
    if (tid < nb) {
      float S[kSlots];
      bool any = false;
#pragma unroll
      for (int c = 0; c < kSlots; ++c) { S[c] = sm.acc[tid * kAcc + c]; any |= (S[c] != 0.f); }
      if (any) {
#pragma unroll
        for (int c = 0; c < kSlots; ++c) sm.acc[tid * kAcc + c] = 0.f;
        const int my_id = __ldg(overlap_to_point + base + tid);
        if (GP) {
          // shift the tile-centred moments to the splat mean: d = l + c
          const float4 RA = sm.rec[buf][tid * RECW], RB = sm.rec[buf][tid * RECW + 1];
          const float4 RC = __ldg(flush_records + base + tid);   // mean - tile centre, 1/sigma.x, 1/sigma.y
          const float inv_k = 1.0f / kExpScale;
          const float cx = -RC.x, cy = -RC.y, s_isx = RC.z, s_isy = RC.w, s_alpha = RB.z;
          const float M0 = S[0], Lx = S[1], Ly = S[2], Lxx = S[3], Lxy = S[4], Lyy = S[5];
          const float Mx = fmaf(cx, M0, Lx), My = fmaf(cy, M0, Ly);
          const float Mxx = Lxx + cx * (2.0f * Lx + cx * M0);
          const float Myy = Lyy + cy * (2.0f * Ly + cy * M0);
          const float Mxy = Lxy + cx * Ly + cy * Lx + cx * cy * M0;
          const float ux = RA.z * inv_k, uy = RB.x * inv_k, wx = RA.w * inv_k, wy = RB.y * inv_k;   // axis / sigma
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
          for (int c = 0; c < F; ++c) atomicAdd(gf + c, S[kSlotF + c]);
        }
        if (HEUR) {
          // paired mode leaves the 1 / k^2 of |G dpdf/dmean| (t and u, w carry one factor k each) to this point
          const float hk = kPaired ? S[kSlotH1] * (1.0f / (kExpScale * kExpScale)) : S[kSlotH1];
          atomicAdd(heuristic + 2 * (int64_t)my_id, S[kSlotH0]);
          atomicAdd(heuristic + 2 * (int64_t)my_id + 1, hk);
        }
      }
    }
  }
}

}  // namespace bwdt

template <int F>
int launch_bwd_transpose(const float4 *records, const float4 *flush_records, const int32_t *ranges, const int32_t *o2p,
                         const float *image, const float *grad_image, const RasterParams<float> &P, int tiles,
                         float *grad_points, float *grad_features, float *heuristic, cudaStream_t stream) {
  constexpr int RECW = F <= 3 ? 3 : 4;
  const bool gp = grad_points != nullptr, gf = grad_features != nullptr, he = P.heur && heuristic != nullptr;
#ifndef GS_BWDT_EXTRA_SMEM
#define GS_BWDT_EXTRA_SMEM 0   // profiling aid: extra dynamic shared memory per CTA lowers the residency
#endif
  int dev = 0;
  GS_CUDA(cudaGetDevice(&dev));
#define GS_BWDT(GP_, GF_, HE_)                                                                                  \
  do {                                                                                                          \
    auto kern = bwdt::raster_bwd_t_kernel<F, GP_, GF_, HE_, RECW>;                                              \
    const size_t smem = sizeof(bwdt::Smem<RECW, bwdt::panel_planes(F, GF_), bwdt::acc_stride(F, GF_)>) +        \
                        GS_BWDT_EXTRA_SMEM;                                                                     \
    /* the attribute is per device (and per kernel instantiation): one bit per device, set once each */         \
    static std::atomic<uint64_t> configured{0};                                                                 \
    const uint64_t dev_bit = 1ull << (dev & 63);                                                                \
    if (!(configured.load(std::memory_order_acquire) & dev_bit)) {                                              \
      GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
      configured.fetch_or(dev_bit, std::memory_order_release);                                                  \
    }                                                                                                           \
    kern<<<tiles, bwdt::kThreads, smem, stream>>>(records, flush_records, ranges, o2p, image, grad_image, P,    \
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

template int launch_bwd_transpose<1>(const float4 *, const float4 *, const int32_t *, const int32_t *, const float *, const float *,
                                     const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);
template int launch_bwd_transpose<2>(const float4 *, const float4 *, const int32_t *, const int32_t *, const float *, const float *,
                                     const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);
template int launch_bwd_transpose<3>(const float4 *, const float4 *, const int32_t *, const int32_t *, const float *, const float *,
                                     const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);
template int launch_bwd_transpose<4>(const float4 *, const float4 *, const int32_t *, const int32_t *, const float *, const float *,
                                     const RasterParams<float> &, int, float *, float *, float *, cudaStream_t);

}  // namespace gs

#ifdef GS_COUNT
extern "C" int gs_debug_counters(unsigned long long *out4, int reset) {
  cudaMemcpyFromSymbol(out4, gs::bwdt::g_count, sizeof(unsigned long long) * 4);
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(gs::bwdt::g_count, z, sizeof z); }
  return 0;
}
#endif
