// Pipe-rate microbenchmarks for the scan kernel's instruction mix on B200 (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
// Reports warp-instructions/clk/SM and lane-results/clk/SM for: MUFU.EX2, FFMA (3-reg), FFMA2, FMUL2,
// and the scan's per-step mix (16 EX2 + 16 FMUL2 + 16 FFMA2 + 8 LDS.128) with no global memory traffic.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b){ f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b){ asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x){ float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int ITERS = 4096;

__global__ void k_mufu(float* out, float seed) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ex2(v[i]);      // 8 independent chains
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, float seed) {
  float v[8], a = seed, b = seed * 0.5f;
  for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(a), "f"(b));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float seed) {
  f32x2 v[8], a = pk2(seed, seed * 0.9f), b = pk2(seed * 0.5f, seed * 0.25f);
  for (int i = 0; i < 8; ++i) v[i] = pk2(seed * (threadIdx.x + i), seed * i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fma2(v[i], a, b);
  }
  float s = 0; for (int i = 0; i < 8; ++i) { float x, y; upk2(v[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fmul2(float* out, float seed) {
  f32x2 v[8], a = pk2(seed, seed * 0.9f);
  for (int i = 0; i < 8; ++i) v[i] = pk2(seed * (threadIdx.x + i), seed * i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = mul2(v[i], a);
  }
  float s = 0; for (int i = 0; i < 8; ++i) { float x, y; upk2(v[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the scan's per-step state update, B/C from shared memory, per-step inputs synthesised in registers
template <int NEX>   // NEX of the 16 exps go through MUFU, the rest are replaced by a cheap FFMA2 stand-in
__global__ void k_scanmix(float* out, float seed, int steps) {
  __shared__ __align__(16) float bc[64][32];
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) (&bc[0][0])[i] = 0.001f * (i % 37);
  __syncthreads();
  f32x2 h[8], a2[8];
  for (int i = 0; i < 8; ++i) { h[i] = pk2(0.f, 0.f); a2[i] = pk2(-(2 * i + 1) * seed, -(2 * i + 2) * seed); }
  float dl = 0.01f * seed + 1e-4f * threadIdx.x, u = 0.5f, acc = 0.f;
  for (int s = 0; s < steps; ++s) {
    const float4* row = reinterpret_cast<const float4*>(&bc[s & 63][0]);
    const float du = dl * u;
    const f32x2 dl2 = pk2(dl, dl), du2 = pk2(du, du);
    f32x2 ya = pk2(u, 0.f), yb = pk2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 Bv = row[q], Cv = row[4 + q];
      f32x2 x0 = mul2(dl2, a2[2 * q]), x1 = mul2(dl2, a2[2 * q + 1]);
      float e0, e1, e2, e3; upk2(x0, e0, e1); upk2(x1, e2, e3);
      f32x2 dA0, dA1;
      if (4 * q < NEX) { dA0 = pk2(ex2(e0), ex2(e1)); dA1 = pk2(ex2(e2), ex2(e3)); }
      else { dA0 = fma2(x0, x0, pk2(1.f, 1.f)); dA1 = fma2(x1, x1, pk2(1.f, 1.f)); }
      h[2 * q] = fma2(dA0, h[2 * q], mul2(du2, pk2(Bv.x, Bv.y)));
      h[2 * q + 1] = fma2(dA1, h[2 * q + 1], mul2(du2, pk2(Bv.z, Bv.w)));
      ya = fma2(h[2 * q], pk2(Cv.x, Cv.y), ya);
      yb = fma2(h[2 * q + 1], pk2(Cv.z, Cv.w), yb);
    }
    float y0, y1, y2, y3; upk2(ya, y0, y1); upk2(yb, y2, y3);
    acc += (y0 + y1) + (y2 + y3);
    dl += 1e-6f; u = -u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sm = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %d MHz\n", p.name, sm, clk_khz / 1000);
  float* out; cudaMalloc(&out, sizeof(float) * sm * 8 * 1024);
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = warps * 32;   // one CTA per SM
    auto rep = [&](const char* name, float ms, double warp_instr_per_thread_iter, double lanes_per_instr) {
      double winst = (double)sm * warps * ITERS * warp_instr_per_thread_iter;
      double clk = ms * 1e-3 * clk_khz * 1e3;
      printf("%-8s warps/SM=%2d  %.3f ms  %.2f warp-instr/clk/SM  %.1f results/clk/SM (at max clock)\n", name, warps, ms,
             winst / clk / sm, winst * lanes_per_instr / clk / sm);
    };
    rep("MUFU.EX2", time_ms([&] { k_mufu<<<sm, threads>>>(out, 0.5f); }), 8, 32);
    rep("FFMA", time_ms([&] { k_ffma<<<sm, threads>>>(out, 0.5f); }), 8, 32);
    rep("FFMA2", time_ms([&] { k_ffma2<<<sm, threads>>>(out, 0.5f); }), 8, 64);
    rep("FMUL2", time_ms([&] { k_fmul2<<<sm, threads>>>(out, 0.5f); }), 8, 64);
  }
  const int steps = 16384;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = warps * 32;
    auto rep = [&](const char* name, float ms) {
      double clk = ms * 1e-3 * clk_khz * 1e3;
      double wsteps = (double)sm * warps * steps;
      printf("%-12s warps/SM=%2d  %.3f ms  %.1f clk per warp-step per SM  (%.2f Texp-equiv/s)\n", name, warps, ms,
             clk * 1.0 / (wsteps / sm), wsteps * 32 * 16 / (ms * 1e-3) / 1e12);
    };
    rep("scanmix16", time_ms([&] { k_scanmix<16><<<sm, threads>>>(out, 0.5f, steps); }));
    rep("scanmix12", time_ms([&] { k_scanmix<12><<<sm, threads>>>(out, 0.5f, steps); }));
    rep("scanmix8", time_ms([&] { k_scanmix<8><<<sm, threads>>>(out, 0.5f, steps); }));
    rep("scanmix0", time_ms([&] { k_scanmix<0><<<sm, threads>>>(out, 0.5f, steps); }));
  }
  return 0;
}
