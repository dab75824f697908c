// Shared device/host helpers for libaum_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/aum_b200.h"

namespace aum {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);   // cudaGetLastError() -> error code + message

#define AUM_REQUIRE(cond, ...)                     \
  do {                                             \
    if (!(cond)) {                                 \
      ::aum::set_error(__VA_ARGS__);               \
      return 1;                                    \
    }                                              \
  } while (0)

// ---- dtype traits -----------------------------------------------------------------------------
template <typename T> struct DT;
template <> struct DT<float>         { static constexpr int id = AUM_F32;  };
template <> struct DT<__half>        { static constexpr int id = AUM_F16;  };
template <> struct DT<__nv_bfloat16> { static constexpr int id = AUM_BF16; };

__host__ __device__ inline int dtype_size(int dt) { return dt == AUM_F32 ? 4 : 2; }

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Runtime-typed scalar load/store (used on cold paths and small side inputs).
__device__ __forceinline__ float load_as_f(const void* p, int64_t idx, int dt) {
  if (dt == AUM_F32) return reinterpret_cast<const float*>(p)[idx];
  if (dt == AUM_F16) return __half2float(reinterpret_cast<const __half*>(p)[idx]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
}
__device__ __forceinline__ void store_from_f(void* p, int64_t idx, int dt, float v) {
  if (dt == AUM_F32) reinterpret_cast<float*>(p)[idx] = v;
  else if (dt == AUM_F16) reinterpret_cast<__half*>(p)[idx] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
}

// 8-element vector <-> 8 floats
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    a = make_float4(f[0], f[1], f[2], f[3]); b = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <> struct Vec8<__half> {
  uint4 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};

// ---- math -------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100: FMUL2 / FFMA2 / FADD2 process two fp32 lanes per issue slot)
typedef unsigned long long f32x2;   // two packed fp32 in one 64-bit register pair

__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// torch.nn.functional.softplus, beta=1, threshold=20 (selective_scan_interface.py:106-107)
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// x * sigmoid(x) with flush-to-zero MUFU ops: FMUL + EX2 + FADD + RCP + FMUL (the non-ftz forms add a range fix-up
// of 3 instructions per MUFU).  exp2 overflow -> rcp(inf) = 0 -> -0, underflow -> x.
__device__ __forceinline__ float silu_ftz(float x) {
  float r;
  const float e = ex2_approx(-1.4426950408889634f * x);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- device handling ----------------------------------------------------------------------------
// Every entry point runs on the device that owns its output pointer (what at::cuda::CUDAGuard on u.get_device() does
// in the reference's pybind kernels): if that is not the calling thread's current device, switch for the duration of
// the call.  Per-device state (cudaFuncSetAttribute done, SM count) is keyed by device ordinal.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const void* p) {
    cudaPointerAttributes a;
    if (p != nullptr && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
      int cur = 0;
      cudaGetDevice(&cur);
      if (cur != a.device) { prev = cur; cudaSetDevice(a.device); }
    } else {
      (void)cudaGetLastError();     // host / unregistered pointer: leave the error state clean, argument checks report it
    }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
constexpr int AUM_MAX_DEVICES = 64;
inline int current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < AUM_MAX_DEVICES) ? d : 0; }
template <typename V> struct PerDevice {
  V v[AUM_MAX_DEVICES] = {};
  V& cur() { return v[current_device()]; }
};

}  // namespace aum
