// Microbenchmark: issue/throughput of packed fp32 (FFMA2/FADD2/FMUL2, sm_100a) against scalar FFMA, and the
// shared-memory wavefront cost of broadcast LDS.128.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

__global__ void k_ffma(float *out, float a, float b) {
  float v[CH];
  for (int i = 0; i < CH; ++i) v[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = fmaf(v[i], a, b);
  float s = 0; for (int i = 0; i < CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b) {
  u64 v[CH]; u64 A = pk(a, a * 1.0001f), B = pk(b, b + 1e-3f);
  for (int i = 0; i < CH; ++i) v[i] = pk(threadIdx.x * 1e-3f + i, i);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = fma2(v[i], A, B);
  float s = 0; for (int i = 0; i < CH; ++i) { float x, y; upk(v[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: FFMA2 interleaved with ALU-pipe work (FMNMX) to see whether packed ops free issue slots
__global__ void k_mix2(float *out, float a, float b) {
  u64 v[CH]; u64 A = pk(a, a * 1.0001f), B = pk(b, b + 1e-3f); float m[CH];
  for (int i = 0; i < CH; ++i) { v[i] = pk(threadIdx.x * 1e-3f + i, i); m[i] = i; }
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) { v[i] = fma2(v[i], A, B); m[i] = fminf(m[i] + 1.0f, a); }
  float s = 0; for (int i = 0; i < CH; ++i) { float x, y; upk(v[i], x, y); s += x + y + m[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix1(float *out, float a, float b) {
  float v[2 * CH], m[CH];
  for (int i = 0; i < 2 * CH; ++i) v[i] = threadIdx.x * 1e-3f + i;
  for (int i = 0; i < CH; ++i) m[i] = i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) { v[2 * i] = fmaf(v[2 * i], a, b); v[2 * i + 1] = fmaf(v[2 * i + 1], a, b); m[i] = fminf(m[i] + 1.0f, a); }
  float s = 0; for (int i = 0; i < 2 * CH; ++i) s += v[i];
  for (int i = 0; i < CH; ++i) s += m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// broadcast shared loads (every lane reads the same address): LSU cycles per warp instruction by width
template <int MODE>
__global__ void k_lds(float *out, int stride) {
  __shared__ float4 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float s = 0; int j = 0;
  const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      j = (j + stride) & 1023;
      const unsigned a = base + j * 16;
      float x, y, z, w;
      if (MODE == 0) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); s += x + y + z + w; }
      if (MODE == 1) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); s += x + y; }
      if (MODE == 2) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); s += x; }
      if (MODE == 3) {  // non-broadcast conflict-free 128-bit: lane l reads record (j + l)
        const unsigned b = base + ((j + (threadIdx.x & 31)) & 1023) * 16;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(b)); s += x + y + z + w; }
      if (MODE == 4) {  // two distinct addresses per warp (half-warps), 128-bit
        const unsigned b = base + ((j + ((threadIdx.x >> 4) & 1) * 37) & 1023) * 16;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(b)); s += x + y + z + w; }
      if (MODE == 5) { x = __shfl_sync(0xffffffffu, s + u, j & 31); s += x; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  const int blocks = 148 * 2, threads = 512;   // 32 warps / SM
  const double n = (double)blocks * threads * ITERS * CH;
  float t1 = time_it([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 1e-3f); });
  float t2 = time_it([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 1e-3f); });
  float t3 = time_it([&] { k_mix1<<<blocks, threads>>>(out, 1.0001f, 1e-3f); });
  float t4 = time_it([&] { k_mix2<<<blocks, threads>>>(out, 1.0001f, 1e-3f); });
  printf("FFMA   : %.3f ms  %.1f Gthread-inst/s  %.2f TFLOP/s\n", t1, n / t1 / 1e6, 2 * n / t1 / 1e9);
  printf("FFMA2  : %.3f ms  %.1f Gthread-inst/s  %.2f TFLOP/s\n", t2, n / t2 / 1e6, 4 * n / t2 / 1e9);
  printf("mix 2xFFMA+FADD+FMNMX : %.3f ms\n", t3);
  printf("mix FFMA2+FADD+FMNMX  : %.3f ms\n", t4);
  const double nl = (double)blocks * threads * ITERS * 8;
  const double clk = 1.965e9;
  const char *names[6] = {"broadcast LDS.128", "broadcast LDS.64", "broadcast LDS.32", "distinct LDS.128 (conflict-free)", "2-address LDS.128", "SHFL.IDX"};
  float l[6];
  l[0] = time_it([&] { k_lds<0><<<blocks, threads>>>(out, 7); });
  l[1] = time_it([&] { k_lds<1><<<blocks, threads>>>(out, 7); });
  l[2] = time_it([&] { k_lds<2><<<blocks, threads>>>(out, 7); });
  l[3] = time_it([&] { k_lds<3><<<blocks, threads>>>(out, 7); });
  l[4] = time_it([&] { k_lds<4><<<blocks, threads>>>(out, 7); });
  l[5] = time_it([&] { k_lds<5><<<blocks, threads>>>(out, 7); });
  for (int i = 0; i < 6; ++i)
    printf("%-34s: %.3f ms -> %.2f SM-cycles per warp instruction\n", names[i], l[i], l[i] * 1e-3 * clk / (nl / 32 / 148));
  return 0;
}
