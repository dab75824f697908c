// Microbenchmark for the round-2 scan restructuring: the per-step state update with CPT channels per thread sharing the
// broadcast B|C loads, with and without the per-step operand traffic of the real kernel (u, delta, z, parked partial
// loads and the y store, all shared memory).  No global memory traffic, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_scan2 microbench_scan2.cu && ./microbench_scan2
// Prints clk per (32-channel warp-step) per SM, directly comparable with microbench.cu's scanmix16 column
// (34.6 clk at 16 warps/SM = the 4.4 T exp/s MUFU ceiling).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b){ f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b){ asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x){ float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int CPT, bool EXTRAS>
__global__ void __launch_bounds__(512, 1) k_mix(float* out, float seed, int steps) {
  extern __shared__ __align__(16) unsigned char sm_[];
  float (*bc)[32] = reinterpret_cast<float (*)[32]>(sm_);                         // 64 rows of B|C
  __half* su = reinterpret_cast<__half*>(sm_ + 64 * 32 * 4);                      // 8 rows x (threads*CPT) halves: u, z, p
  const int nch = blockDim.x * CPT;
  __half* sz = su + 8 * nch; __half* sp = sz + 8 * nch;
  float* sd = reinterpret_cast<float*>(sp + 8 * nch);                             // 8 rows x nch floats: delta
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) (&bc[0][0])[i] = 0.001f * (i % 37);
  for (int i = threadIdx.x; i < 8 * nch; i += blockDim.x) { su[i] = __float2half(0.5f); sz[i] = __float2half(0.9f); sp[i] = __float2half(0.1f); sd[i] = 0.01f * seed + 1e-6f * i; }
  __syncthreads();
  f32x2 h[CPT][8], a2[CPT][8];
  for (int c = 0; c < CPT; ++c) for (int i = 0; i < 8; ++i) { h[c][i] = pk2(0.f, 0.f); a2[c][i] = pk2(-(2 * i + 1) * seed * (1 + c), -(2 * i + 2) * seed); }
  float acc = 0.f;
  float dl_r[CPT], u_r[CPT];
  for (int c = 0; c < CPT; ++c) { dl_r[c] = 0.01f * seed + 1e-4f * threadIdx.x; u_r[c] = 0.5f; }
#pragma unroll 1
  for (int s0 = 0; s0 < steps; s0 += 4) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int s = s0 + t;
      const float4* row = reinterpret_cast<const float4*>(&bc[s & 63][0]);
      float u[CPT], dl[CPT], zz[CPT], pp[CPT];
      if (EXTRAS) {
        const int r = s & 7;
        if (CPT == 2) {
          const __half2 uu = *reinterpret_cast<const __half2*>(su + r * nch + 2 * threadIdx.x);
          const __half2 z2 = *reinterpret_cast<const __half2*>(sz + r * nch + 2 * threadIdx.x);
          const __half2 p2 = *reinterpret_cast<const __half2*>(sp + r * nch + 2 * threadIdx.x);
          const float2 d2 = *reinterpret_cast<const float2*>(sd + r * nch + 2 * threadIdx.x);
          u[0] = __low2float(uu); u[CPT - 1] = __high2float(uu); zz[0] = __low2float(z2); zz[CPT - 1] = __high2float(z2);
          pp[0] = __low2float(p2); pp[CPT - 1] = __high2float(p2); dl[0] = d2.x; dl[CPT - 1] = d2.y;
        } else {
          u[0] = __half2float(su[r * nch + threadIdx.x]); zz[0] = __half2float(sz[r * nch + threadIdx.x]);
          pp[0] = __half2float(sp[r * nch + threadIdx.x]); dl[0] = sd[r * nch + threadIdx.x];
        }
      } else {
        for (int c = 0; c < CPT; ++c) { u[c] = u_r[c]; dl[c] = dl_r[c]; zz[c] = 1.f; pp[c] = 0.f; }
      }
      f32x2 dl2[CPT], du2[CPT], ya[CPT], yb[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) { const float du = dl[c] * u[c]; dl2[c] = pk2(dl[c], dl[c]); du2[c] = pk2(du, du); ya[c] = pk2(u[c], 0.f); yb[c] = pk2(0.f, 0.f); }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 Bv = row[q], Cv = row[4 + q];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          f32x2 x0 = mul2(dl2[c], a2[c][2 * q]), x1 = mul2(dl2[c], a2[c][2 * q + 1]);
          float e0, e1, e2, e3; upk2(x0, e0, e1); upk2(x1, e2, e3);
          const f32x2 dA0 = pk2(ex2(e0), ex2(e1)), dA1 = pk2(ex2(e2), ex2(e3));
          h[c][2 * q] = fma2(dA0, h[c][2 * q], mul2(du2[c], pk2(Bv.x, Bv.y)));
          h[c][2 * q + 1] = fma2(dA1, h[c][2 * q + 1], mul2(du2[c], pk2(Bv.z, Bv.w)));
          ya[c] = fma2(h[c][2 * q], pk2(Cv.x, Cv.y), ya[c]);
          yb[c] = fma2(h[c][2 * q + 1], pk2(Cv.z, Cv.w), yb[c]);
        }
      }
      float y[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) { float y0, y1, y2, y3; upk2(ya[c], y0, y1); upk2(yb[c], y2, y3); y[c] = ((y0 + y1) + (y2 + y3) + pp[c]) * zz[c]; }
      if (EXTRAS) {
        const int r = s & 7;
        if (CPT == 2) *reinterpret_cast<__half2*>(su + r * nch + 2 * threadIdx.x) = __floats2half2_rn(y[0], y[CPT - 1]);
        else su[r * nch + threadIdx.x] = __float2half(y[0]);
      } else {
        for (int c = 0; c < CPT; ++c) { acc += y[c]; dl_r[c] += 1e-6f; u_r[c] = -u_r[c]; }
      }
    }
  }
  if (EXTRAS) for (int c = 0; c < CPT; ++c) acc += __half2float(su[threadIdx.x * CPT + c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

template <int CPT, bool EXTRAS> void run(int sm, int clk_khz, float* out, int warps) {
  const int steps = 16384, threads = warps * 32;
  const int smem = 64 * 32 * 4 + 3 * 8 * threads * CPT * 2 + 8 * threads * CPT * 4;
  cudaFuncSetAttribute(k_mix<CPT, EXTRAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float ms = time_ms([&] { k_mix<CPT, EXTRAS><<<sm, threads, smem>>>(out, 0.5f, steps); });
  double clk = ms * 1e-3 * clk_khz * 1e3;
  double ch_wsteps_per_sm = (double)warps * CPT * steps;
  printf("CPT=%d extras=%d warps/SM=%2d  %.3f ms  %.1f clk per channel-warp-step per SM  (%.2f Texp/s)  %s\n", CPT, (int)EXTRAS, warps, ms,
         clk / ch_wsteps_per_sm, ch_wsteps_per_sm * sm * 32 * 16 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sm = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %d MHz\n", p.name, sm, clk_khz / 1000);
  float* out; cudaMalloc(&out, sizeof(float) * sm * 1024);
  for (int warps : {4, 8, 12, 16}) { run<1, false>(sm, clk_khz, out, warps); run<1, true>(sm, clk_khz, out, warps); }
  for (int warps : {4, 6, 8, 12, 16}) { run<2, false>(sm, clk_khz, out, warps); run<2, true>(sm, clk_khz, out, warps); }
  return 0;
}
