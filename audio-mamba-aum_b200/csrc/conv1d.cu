// Depthwise causal conv1d (+bias +SiLU) over token-major activations.
// Replaces causal_conv1d_cuda.causal_conv1d_fwd (call sites: /root/reference/vim-mamba_ssm/mamba_ssm/ops/
// selective_scan_interface.py:177,239,318,380,463,532); semantics = mamba_simple.py:272 with padding W-1.
//
// HBM-bound streaming kernel.  Each thread owns VEC (=2) adjacent channels and CONV_TL (=16) consecutive tokens:
// a warp therefore reads/writes one contiguous 128-byte (16-bit) or 256-byte (fp32) segment per token row, all
// TL+3 input rows of a thread are independent loads issued up front (one DRAM latency per thread), kept PACKED
// in registers (19 regs) so the kernel runs at full occupancy, and the 3-row halo is served by L1/L2.
// Algorithmic bytes per (token, channel): read s + write s (s = itemsize).
#include <stdlib.h>
#include <type_traits>

#include "scan_common.cuh"   // packed fp32x2 helpers

namespace aum {

constexpr int CONV_TL = 16;   // tokens per thread
constexpr int CONV_MAXW = 4;

template <typename T> struct Pair;        // two adjacent channels, packed as loaded
template <> struct Pair<float> {
  float2 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float2*>(p); }
  __device__ __forceinline__ void zero() { v = make_float2(0.f, 0.f); }
  __device__ __forceinline__ void set(float a, float b) { v = make_float2(a, b); }
  __device__ __forceinline__ float2 f() const { return v; }
  static __device__ __forceinline__ void store(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <> struct Pair<__half> {
  __half2 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const __half2*>(p); }
  __device__ __forceinline__ void zero() { v = __floats2half2_rn(0.f, 0.f); }
  __device__ __forceinline__ void set(float a, float b) { v = __floats2half2_rn(a, b); }
  __device__ __forceinline__ float2 f() const { return __half22float2(v); }
  static __device__ __forceinline__ void store(__half* p, float a, float b) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); }
};
template <> struct Pair<__nv_bfloat16> {
  __nv_bfloat162 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const __nv_bfloat162*>(p); }
  __device__ __forceinline__ void zero() { v = __floats2bfloat162_rn(0.f, 0.f); }
  __device__ __forceinline__ void set(float a, float b) { v = __floats2bfloat162_rn(a, b); }
  __device__ __forceinline__ float2 f() const { return __bfloat1622float2(v); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
conv1d_fwd_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                  const float* __restrict__ bias, T* __restrict__ out, int64_t ldo,
                  int batch, int L, int D, int W, int silu, int reverse, int n_cvec, int n_ltile) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)batch * n_ltile * n_cvec;
  if (gid >= total) return;
  const int cv = (int)(gid % n_cvec);
  const int lt = (int)((gid / n_cvec) % n_ltile);
  const int b = (int)(gid / ((int64_t)n_cvec * n_ltile));
  const int c0 = cv * VEC;

  // taps, zero-padded at the front to CONV_MAXW:  y[l] = bias + sum_j wk[j] * x[l - (MAXW-1) + j]
  float wk[CONV_MAXW][2];
  float bs[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int c = c0 + v;
    const bool ok = (v < VEC) && (c < D);
#pragma unroll
    for (int j = 0; j < CONV_MAXW; ++j) {
      const int k = j - (CONV_MAXW - W);
      wk[j][v] = (ok && k >= 0) ? __ldg(w + (int64_t)c * W + k) : 0.f;
    }
    bs[v] = (ok && bias != nullptr) ? __ldg(bias + c) : 0.f;
  }

  // Walk direction: causal -> ascending tokens with history of lower indices;
  // reverse (anti-causal) -> descending tokens with history of higher indices.
  const int l_begin = lt * CONV_TL;
  const int l_end = min(L, l_begin + CONV_TL);
  const int n = l_end - l_begin;
  const int step = reverse ? -1 : 1;
  const int l_first = reverse ? (l_end - 1) : l_begin;
  const T* xb = x + (int64_t)b * L * ldx + c0;
  T* ob = out + (int64_t)b * L * ldo + c0;

  constexpr int NR = CONV_TL + CONV_MAXW - 1;
  Pair<T> rows[NR];        // rows[j] = x at walk position (j - (MAXW-1)) relative to the first token
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const int pos = j - (CONV_MAXW - 1);                 // < 0: history, >= n: beyond this tile
    const int l = l_first + step * pos;
    const bool in = (pos < n) && (l >= 0) && (l < L);
    rows[j].zero();
    if (in) {
      if constexpr (VEC == 2) rows[j].load(xb + (int64_t)l * ldx);
      else rows[j].set(to_f(xb[(int64_t)l * ldx]), 0.f);      // scalar path: single channel in the low lane
    }
  }
#pragma unroll
  for (int i = 0; i < CONV_TL; ++i) {
    if (i < n) {
      const int l = l_first + step * i;
      float a0 = bs[0], a1 = bs[1];
#pragma unroll
      for (int j = 0; j < CONV_MAXW; ++j) {
        const float2 f = rows[i + j].f();
        a0 = fmaf(wk[j][0], f.x, a0);
        a1 = fmaf(wk[j][1], f.y, a1);
      }
      if (silu) { a0 = silu_f(a0); a1 = silu_f(a1); }
      if constexpr (VEC == 2) Pair<T>::store(ob + (int64_t)l * ldo, a0, a1);
      else ob[(int64_t)l * ldo] = from_f<T>(a0);
    }
  }
}

// ---- fast path: 4 adjacent channels per thread --------------------------------------------------------------
// The generic kernel above spends ~44 instructions per element (it re-converts every packed row once per tap and
// works on scalar lanes), which makes it issue-bound at 2.5 TB/s.  This one converts each input row once into two
// packed fp32x2 pairs, runs the taps as FFMA2 on the pairs (2 per element) and stores 8 or 16 bytes per thread:
// ~10 instructions per element, so the kernel is bound by HBM again.  Needs D, pitches and bases 4-element aligned.
template <typename T> struct Quad;      // 4 adjacent channels as loaded
template <> struct Quad<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void f(f32x2& a, f32x2& b) const { a = pk2(v.x, v.y); b = pk2(v.z, v.w); }
  static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  }
};
template <> struct Quad<__half> {
  uint2 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void zero() { v = make_uint2(0u, 0u); }
  __device__ __forceinline__ void f(f32x2& a, f32x2& b) const {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    a = pk2(lo.x, lo.y); b = pk2(hi.x, hi.y);
  }
  static __device__ __forceinline__ void store(__half* p, float a, float b, float c, float d) {
    __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
};
template <> struct Quad<__nv_bfloat16> {
  uint2 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void zero() { v = make_uint2(0u, 0u); }
  __device__ __forceinline__ void f(f32x2& a, f32x2& b) const {   // bf16 -> fp32 is a 16-bit shift
    a = pk2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
    b = pk2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
};

template <typename T>
__global__ void __launch_bounds__(128)
conv1d_fwd_vec4_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                       const float* __restrict__ bias, T* __restrict__ out, int64_t ldo,
                       int batch, int L, int D, int W, int silu, int reverse, int n_cvec, int n_ltile) {
  // grid: x = (sequence, 16-token tile), y = 128-thread slice of the channel quads (32-bit index math only)
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= n_cvec) return;
  const int lt = (int)(blockIdx.x % (unsigned)n_ltile);
  const int b = (int)(blockIdx.x / (unsigned)n_ltile);
  const int c0 = cv * 4;

  // taps as channel pairs, zero-padded at the front to CONV_MAXW
  f32x2 wa[CONV_MAXW], wb[CONV_MAXW];
#pragma unroll
  for (int j = 0; j < CONV_MAXW; ++j) {
    const int k = j - (CONV_MAXW - W);
    float t[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) t[v] = k >= 0 ? __ldg(w + (int64_t)(c0 + v) * W + k) : 0.f;
    wa[j] = pk2(t[0], t[1]); wb[j] = pk2(t[2], t[3]);
  }
  f32x2 ba = pk2(0.f, 0.f), bb = ba;
  if (bias != nullptr) { const float4 t = __ldg(reinterpret_cast<const float4*>(bias + c0)); ba = pk2(t.x, t.y); bb = pk2(t.z, t.w); }

  const int l_begin = lt * CONV_TL;
  const int l_end = min(L, l_begin + CONV_TL);
  const int n = l_end - l_begin;
  const int step = reverse ? -1 : 1;
  const int l_first = reverse ? (l_end - 1) : l_begin;
  const int64_t xstep = (int64_t)step * ldx, ostep = (int64_t)step * ldo;
  const T* xp = x + ((int64_t)b * L + l_first) * ldx + c0 - (CONV_MAXW - 1) * xstep;   // row of walk position -(MAXW-1)
  T* op = out + ((int64_t)b * L + l_first) * ldo + c0;

  constexpr int NR = CONV_TL + CONV_MAXW - 1;
  Quad<T> rows[NR];        // rows[j] = x at walk position (j - (MAXW-1)) relative to the first token; all loads up front
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const int pos = j - (CONV_MAXW - 1);
    const int l = l_first + step * pos;
    rows[j].zero();
    if ((pos < n) && (l >= 0) && (l < L)) rows[j].load(xp);
    xp += xstep;
  }
  f32x2 fa[NR], fb[NR];    // converted once; only a 4-row window is live at a time
#pragma unroll
  for (int j = 0; j < CONV_MAXW - 1; ++j) rows[j].f(fa[j], fb[j]);
#pragma unroll
  for (int i = 0; i < CONV_TL; ++i) {
    rows[i + CONV_MAXW - 1].f(fa[i + CONV_MAXW - 1], fb[i + CONV_MAXW - 1]);
    if (i < n) {
      f32x2 a = ba, c = bb;
#pragma unroll
      for (int j = 0; j < CONV_MAXW; ++j) { a = fma2(wa[j], fa[i + j], a); c = fma2(wb[j], fb[i + j], c); }
      float y0, y1, y2, y3;
      upk2(a, y0, y1); upk2(c, y2, y3);
      if (silu) { y0 = silu_ftz(y0); y1 = silu_ftz(y1); y2 = silu_ftz(y2); y3 = silu_ftz(y3); }
      Quad<T>::store(op, y0, y1, y2, y3);
    }
    op += ostep;
  }
}

template <typename T>
static int launch_conv(const void* x, int64_t ldx, const float* w, const float* bias, void* out, int64_t ldo,
                       int batch, int L, int D, int W, int silu, int reverse, cudaStream_t st) {
  const int n_ltile = ceil_div(L, CONV_TL);
  const int esz = (int)sizeof(T);
  const bool vec_ok = (D % 2 == 0) && (ldx % 2 == 0) && (ldo % 2 == 0) &&
                      (reinterpret_cast<uintptr_t>(x) % (2 * esz) == 0) && (reinterpret_cast<uintptr_t>(out) % (2 * esz) == 0);
  const bool vec4_ok = (D % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) &&
                       (reinterpret_cast<uintptr_t>(x) % (4 * esz) == 0) && (reinterpret_cast<uintptr_t>(out) % (4 * esz) == 0) &&
                       (bias == nullptr || reinterpret_cast<uintptr_t>(bias) % 16 == 0);
  if (vec4_ok && (int64_t)batch * n_ltile < (1ll << 31)) {
    const int n_cvec = D / 4;
    const dim3 grid((unsigned)(batch * n_ltile), (unsigned)ceil_div(n_cvec, 128));
    conv1d_fwd_vec4_kernel<T><<<grid, 128, 0, st>>>(
        (const T*)x, ldx, w, bias, (T*)out, ldo, batch, L, D, W, silu, reverse, n_cvec, n_ltile);
  } else if (vec_ok) {
    const int n_cvec = D / 2;
    const int64_t total = (int64_t)batch * n_ltile * n_cvec;
    conv1d_fwd_kernel<T, 2><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(
        (const T*)x, ldx, w, bias, (T*)out, ldo, batch, L, D, W, silu, reverse, n_cvec, n_ltile);
  } else {
    const int n_cvec = D;
    const int64_t total = (int64_t)batch * n_ltile * n_cvec;
    conv1d_fwd_kernel<T, 1><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(
        (const T*)x, ldx, w, bias, (T*)out, ldo, batch, L, D, W, silu, reverse, n_cvec, n_ltile);
  }
  return check_launch("aum_causal_conv1d_fwd");
}

}  // namespace aum

extern "C" int aum_causal_conv1d_fwd(const void* x, int64_t ldx, const float* w, const float* bias,
                                     void* out, int64_t ldo, int batch, int L, int D, int W,
                                     int dtype, int silu, int reverse, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(out);
  if (batch == 0 || L == 0 || D == 0) return 0;      // empty input: nothing to do (pointers may be null)
  AUM_REQUIRE(x && w && out, "aum_causal_conv1d_fwd: null pointer");
  AUM_REQUIRE(W >= 2 && W <= CONV_MAXW, "aum_causal_conv1d_fwd: width %d unsupported (2..4)", W);
  AUM_REQUIRE(batch >= 0 && L >= 0 && D >= 0, "aum_causal_conv1d_fwd: negative size");
  AUM_REQUIRE(ldx >= D && ldo >= D, "aum_causal_conv1d_fwd: leading dimension smaller than D");
  if (batch == 0 || L == 0 || D == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case AUM_F32:  return launch_conv<float>(x, ldx, w, bias, out, ldo, batch, L, D, W, silu, reverse, st);
    case AUM_F16:  return launch_conv<__half>(x, ldx, w, bias, out, ldo, batch, L, D, W, silu, reverse, st);
    case AUM_BF16: return launch_conv<__nv_bfloat16>(x, ldx, w, bias, out, ldo, batch, L, D, W, silu, reverse, st);
  }
  set_error("aum_causal_conv1d_fwd: bad dtype %d", dtype);
  return 1;
}

// ======================================================================================================
// Backward of the depthwise causal conv (+bias, +SiLU).
// Replaces causal_conv1d_cuda.causal_conv1d_bwd(x, w, bias, dout, None, dx, True)
// (selective_scan_interface.py:281-283, 425-427, 594-596).  With c = bias + conv(x) (recomputed here),
// dc = dout * silu'(c):  dx[p] = sum_k w[k] dc[p+(W-1)-k],  dw[k] = sum dc[p] x[p-(W-1)+k],  dbias = sum dc
// (p = position in walk order: token index for the causal conv, L-1-token for the anti-causal one).
// Block = 8 warps x 32 lanes: lane = channel pair, warp = one of 8 consecutive 16-token tiles of one sequence;
// dw/dbias partials are reduced over the 8 warps in shared memory before one atomicAdd per value.
// ======================================================================================================
namespace aum {

constexpr int CB_TL = 8;

template <typename T>
__global__ void __launch_bounds__(256, 3)
conv1d_bwd_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ dout, const float* __restrict__ dout2, const float* __restrict__ dout3, int64_t ldd,
                  T* __restrict__ dx, int64_t ld_dx, float* __restrict__ dw, float* __restrict__ dbias,
                  int batch, int L, int D, int W, int silu, int reverse, int n_cgrp, int n_tgrp) {
  __shared__ float red[8][10][32];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int cg = blockIdx.x % n_cgrp;
  const int tg = (blockIdx.x / n_cgrp) % n_tgrp;
  const int b = blockIdx.x / (n_cgrp * n_tgrp);
  const int c0 = (cg * 32 + lane) * 2;
  const bool ok0 = c0 < D, ok1 = c0 + 1 < D;
  const bool pair_ok = ok1 && (ldx % 2 == 0) && (ldd % 2 == 0) && (ld_dx % 2 == 0) &&
                       (reinterpret_cast<uintptr_t>(x) % (2 * sizeof(T)) == 0) &&
                       (reinterpret_cast<uintptr_t>(dx) % (2 * sizeof(T)) == 0) &&
                       (reinterpret_cast<uintptr_t>(dout) % 8 == 0) && (reinterpret_cast<uintptr_t>(dout2) % 8 == 0) &&
                       (reinterpret_cast<uintptr_t>(dout3) % 8 == 0);
  const int p0 = (tg * 8 + wrp) * CB_TL;           // first walk position of this warp's tile

  float wk[CONV_MAXW][2], bs[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int c = c0 + v;
    const bool ok = c < D;
#pragma unroll
    for (int j = 0; j < CONV_MAXW; ++j) {
      const int k = j - (CONV_MAXW - W);
      wk[j][v] = (ok && k >= 0) ? __ldg(w + (int64_t)c * W + k) : 0.f;
    }
    bs[v] = (ok && bias != nullptr) ? __ldg(bias + c) : 0.f;
  }
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;

  if (p0 < L && ok0) {
    const int64_t base = (int64_t)b * L;
    auto tok = [&](int p) { return reverse ? (L - 1 - p) : p; };
    // everything this tile touches is fetched up front (independent loads), kept packed:
    //   x at walk positions p0-3 .. p0+TL+2, dout at p0 .. p0+TL+2
    constexpr int NX = CB_TL + 2 * (CONV_MAXW - 1);
    constexpr int ND = CB_TL + (CONV_MAXW - 1);
    Pair<T> xr[NX];
    float2 gr[ND];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      const int p = p0 - (CONV_MAXW - 1) + j;
      xr[j].zero();
      if (p >= 0 && p < L) {
        const T* r = x + (base + tok(p)) * ldx + c0;
        if (pair_ok) xr[j].load(r); else xr[j].set(to_f(r[0]), ok1 ? to_f(r[1]) : 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int p = p0 + j;
      gr[j] = make_float2(0.f, 0.f);
      if (p < L) {
        const float* g = dout + (base + tok(p)) * ldd + c0;
        if (pair_ok) gr[j] = *reinterpret_cast<const float2*>(g); else gr[j] = make_float2(g[0], ok1 ? g[1] : 0.f);
        if (dout2 != nullptr) {          // second gradient term (same layout), summed on the fly
          const float* g2 = dout2 + (base + tok(p)) * ldd + c0;
          if (pair_ok) { const float2 t = *reinterpret_cast<const float2*>(g2); gr[j].x += t.x; gr[j].y += t.y; }
          else { gr[j].x += g2[0]; if (ok1) gr[j].y += g2[1]; }
        }
        if (dout3 != nullptr) {          // third term (the other scan direction's du when the two are kept apart)
          const float* g3 = dout3 + (base + tok(p)) * ldd + c0;
          if (pair_ok) { const float2 t = *reinterpret_cast<const float2*>(g3); gr[j].x += t.x; gr[j].y += t.y; }
          else { gr[j].x += g3[0]; if (ok1) gr[j].y += g3[1]; }
        }
      }
    }
    // dc[j] = dout * act'(c) at position p0 + j
    float2 dc[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      float ca = bs[0], cb = bs[1];
#pragma unroll
      for (int k = 0; k < CONV_MAXW; ++k) {
        const float2 f = xr[j + k].f();                     // x[p-3+k]
        ca = fmaf(wk[k][0], f.x, ca); cb = fmaf(wk[k][1], f.y, cb);
      }
      float ga = gr[j].x, gb = gr[j].y;
      if (silu) {
        const float sa = __fdividef(1.f, 1.f + __expf(-ca)), sb = __fdividef(1.f, 1.f + __expf(-cb));
        ga *= sa * (1.f + ca * (1.f - sa)); gb *= sb * (1.f + cb * (1.f - sb));
      }
      dc[j] = (p0 + j < L) ? make_float2(ga, gb) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < CB_TL; ++i) {
      const int p = p0 + i;
      if (p < L) {
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int k = 0; k < CONV_MAXW; ++k) {       // dx[p] = sum_k w[k] dc[p + 3 - k]
          d0 = fmaf(wk[k][0], dc[i + (CONV_MAXW - 1) - k].x, d0);
          d1 = fmaf(wk[k][1], dc[i + (CONV_MAXW - 1) - k].y, d1);
        }
        T* o = dx + (base + tok(p)) * ld_dx + c0;
        if (pair_ok) Pair<T>::store(o, d0, d1);
        else { o[0] = from_f<T>(d0); if (ok1) o[1] = from_f<T>(d1); }
#pragma unroll
        for (int k = 0; k < CONV_MAXW; ++k) {       // dw[k] += dc[p] x[p-3+k]
          const float2 f = xr[i + k].f();
          acc[k] = fmaf(dc[i].x, f.x, acc[k]);
          acc[5 + k] = fmaf(dc[i].y, f.y, acc[5 + k]);
        }
        acc[4] += dc[i].x;
        acc[9] += dc[i].y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) red[wrp][i][lane] = acc[i];
  __syncthreads();
  if (wrp == 0 && ok0) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][i][lane];
      const int v = i / 5, j = i % 5;
      const int c = c0 + v;
      if (c < D) {
        if (j < 4) { const int k = j - (CONV_MAXW - W); if (k >= 0) atomicAdd(dw + (int64_t)c * W + k, s); }
        else if (dbias != nullptr) atomicAdd(dbias + c, s);
      }
    }
  }
}


// ---- streaming form (the production path) --------------------------------------------------------------------
// The tile kernel above fetches everything a warp's 8-token tile touches up front (14 x rows + 11 rows of each gradient
// term) and then computes; at AuM-Base size with the three gradient terms of a Fo-Bi block it ran at 26 % of the HBM
// roofline: 80 registers = 24 warps per SM, each of them alternating between one burst of loads and a long dependent
// compute phase (long-scoreboard stall 9.1 per issue, profiles/r2_ncu_conv1d_bwd3_summary.txt), with 11/8 of the
// gradient bytes requested.  Here a thread (one channel pair) WALKS a segment of ~32-64 positions: per position one
// row of x and of each gradient term is loaded (the loads of the next positions are in flight while this one is
// computed), the 4-tap windows of x and dc slide through registers, dx[p-3] leaves as soon as dc[p] exists, and the
// weight / bias gradients accumulate in registers over the whole segment (8 warps = 8 consecutive segments per block,
// reduced in shared memory, one atomic per value and block).  Packed fp32x2 math on the channel pair.
// Needs W <= 4, D even and 8-byte-aligned rows (checked by the launcher; anything else takes the tile kernel).
// a channel pair of one gradient term as fp32 (TG: float, or the 16-bit activation dtype)
template <typename TG> __device__ __forceinline__ float2 ldg_pair(const TG* p);
template <> __device__ __forceinline__ float2 ldg_pair<float>(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
template <> __device__ __forceinline__ float2 ldg_pair<__half>(const __half* p) {
  const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
template <> __device__ __forceinline__ float2 ldg_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

// UNR / MINB: positions per loop trip / resident blocks per SM the register budget is set for: (2, 4) = 64 registers,
// 32 warps per SM; (4, 3) = 80 registers, 24 warps with twice the loads in flight per warp (AUM_CONV_BWD_VARIANT=1).
template <typename T, typename TG, int NT, int UNR, int MINB>
__global__ void __launch_bounds__(256, MINB)
conv1d_bwd_stream_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ bias,
                         const TG* __restrict__ g1, const TG* __restrict__ g2, const TG* __restrict__ g3, int64_t ldd,
                         T* __restrict__ dx, int64_t ld_dx, float* __restrict__ dw, float* __restrict__ dbias,
                         int L, int D, int W, int silu, int reverse, int n_cgrp, int n_tgrp, int seg) {
  __shared__ float red[8][10][32];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int cg = blockIdx.x % n_cgrp;
  const int tg = (blockIdx.x / n_cgrp) % n_tgrp;
  const int b = blockIdx.x / (n_cgrp * n_tgrp);
  const int c0 = (cg * 32 + lane) * 2;
  const bool ok = c0 < D;
  const int p0 = (tg * 8 + wrp) * seg;
  const int p1 = min(p0 + seg, L);

  f32x2 wk[CONV_MAXW];
  f32x2 bs = pk2(0.f, 0.f);
  if (ok) {
#pragma unroll
    for (int j = 0; j < CONV_MAXW; ++j) {
      const int k = j - (CONV_MAXW - W);
      wk[j] = k >= 0 ? pk2(__ldg(w + (int64_t)c0 * W + k), __ldg(w + (int64_t)(c0 + 1) * W + k)) : pk2(0.f, 0.f);
    }
    if (bias != nullptr) bs = pk2(__ldg(bias + c0), __ldg(bias + c0 + 1));
  }
  f32x2 acc[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) acc[i] = pk2(0.f, 0.f);

  if (ok && p0 < L) {
    // walk position p <-> token (reverse ? L-1-p : p); one step along the walk = sgn rows.  Element offsets are 32-bit
    // (the launcher checks that every buffer spans < 2^31 elements): half the registers of 64-bit pointers + strides.
    const int sgn = reverse ? -1 : 1;
    const int r0 = b * L + (reverse ? (L - 1 - p0) : p0);
    const T* const xb = x + c0;
    const TG* const q1 = g1 + c0;
    const TG* const q2 = NT >= 2 ? g2 + c0 : nullptr;
    const TG* const q3 = NT >= 3 ? g3 + c0 : nullptr;
    T* const ob = dx + c0;
    int ox = r0 * (int)ldx, og = r0 * (int)ldd, oo = r0 * (int)ld_dx;     // rows of x[p], g[p], dx[next row to emit]
    const int sx = sgn * (int)ldx, sg = sgn * (int)ldd, so = sgn * (int)ld_dx;
    auto ldx2 = [&](const T* r) { Pair<T> t; t.load(r); const float2 f = t.f(); return pk2(f.x, f.y); };
    const T* const xp = xb + ox;
    // x[p0-3 .. p0-1] (zeros before the sequence start)
    f32x2 xm3 = p0 >= 3 ? ldx2(xp - 3 * sx) : pk2(0.f, 0.f);
    f32x2 xm2 = p0 >= 2 ? ldx2(xp - 2 * sx) : pk2(0.f, 0.f);
    f32x2 xm1 = p0 >= 1 ? ldx2(xp - sx) : pk2(0.f, 0.f);
    f32x2 d1 = pk2(0.f, 0.f), d2 = d1, d3 = d1;    // dc[p-1], dc[p-2], dc[p-3]
    const f32x2 one = pk2(1.f, 1.f), mone = pk2(-1.f, -1.f);

    // one position with data (p < L).  ACC: p belongs to this segment (its dc feeds dw / dbias); EMIT: dx[p-3] is ours
    auto step = [&](const bool accum, const bool emit) {
      const f32x2 xn = ldx2(xb + ox);
      float2 g = ldg_pair<TG>(q1 + og);
      if (NT >= 2) { const float2 t = ldg_pair<TG>(q2 + og); g.x += t.x; g.y += t.y; }
      if (NT >= 3) { const float2 t = ldg_pair<TG>(q3 + og); g.x += t.x; g.y += t.y; }
      ox += sx; og += sg;
      f32x2 dcn = pk2(g.x, g.y);
      if (silu) {
        const f32x2 c = fma2(wk[3], xn, fma2(wk[2], xm1, fma2(wk[1], xm2, fma2(wk[0], xm3, bs))));
        float ca, cb;
        upk2(c, ca, cb);
        float sa, sb;
        const float ea = ex2_approx(-1.4426950408889634f * ca), eb = ex2_approx(-1.4426950408889634f * cb);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sa) : "f"(1.f + ea));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sb) : "f"(1.f + eb));
        const f32x2 sg2 = pk2(sa, sb);
        // silu'(c) = s (1 + c (1 - s))
        dcn = mul2(dcn, mul2(sg2, fma2(c, fma2(sg2, mone, one), one)));
      }
      if (accum) {
        acc[0] = fma2(dcn, xm3, acc[0]); acc[1] = fma2(dcn, xm2, acc[1]);
        acc[2] = fma2(dcn, xm1, acc[2]); acc[3] = fma2(dcn, xn, acc[3]);
        acc[4] = add2(acc[4], dcn);
      }
      if (emit) {                                   // dx[p-3] = w0 dc[p] + w1 dc[p-1] + w2 dc[p-2] + w3 dc[p-3]
        const f32x2 o = fma2(wk[3], d3, fma2(wk[2], d2, fma2(wk[1], d1, mul2(wk[0], dcn))));
        float oa, ob_;
        upk2(o, oa, ob_);
        Pair<T>::store(ob + oo, oa, ob_);
        oo += so;
      }
      xm3 = xm2; xm2 = xm1; xm1 = xn;
      d3 = d2; d2 = d1; d1 = dcn;
    };

    // a position past the sequence end: dc = 0 (no gradient arrives there); EMIT as above
    auto zstep = [&](const bool emit) {
      if (emit) {
        const f32x2 o = fma2(wk[3], d3, fma2(wk[2], d2, mul2(wk[1], d1)));
        float oa, ob_;
        upk2(o, oa, ob_);
        Pair<T>::store(ob + oo, oa, ob_);
        oo += so;
      }
      d3 = d2; d2 = d1; d1 = pk2(0.f, 0.f);
    };
    // every segment takes (p1 - p0) + 3 steps: the first three emit nothing (dx[p0] needs dc up to p0 + 3), the last
    // three are the halo - dc of the NEXT segment's first positions (no dw / dbias from them) or zeros past the end
    int p = p0;
#pragma unroll 1
    for (int i = 0; i < 3; ++i, ++p) {
      if (p < L) step(p < p1, false); else zstep(false);
    }
#pragma unroll UNR
    for (; p < p1; ++p) step(true, true);
    const int pend = p1 + 3;
#pragma unroll 1
    for (; p < pend; ++p) {
      if (p < L) step(false, true); else zstep(true);
    }
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    float lo, hi;
    upk2(acc[i], lo, hi);
    red[wrp][i][lane] = lo; red[wrp][5 + i][lane] = hi;
  }
  __syncthreads();
  if (wrp == 0 && ok) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += red[k][i][lane];
      const int v = i / 5, j = i % 5;
      const int c = c0 + v;
      if (j < 4) { const int k = j - (CONV_MAXW - W); if (k >= 0) atomicAdd(dw + (int64_t)c * W + k, sum); }
      else if (dbias != nullptr) atomicAdd(dbias + c, sum);
    }
  }
}

template <typename T, typename TG, int UNR, int MINB>
static void launch_conv_bwd_stream_v(const void* x, int64_t ldx, const float* w, const float* bias, const void* g1_, const void* g2_,
                                     const void* g3_, int64_t ldd, void* dx, int64_t ld_dx, float* dw, float* dbias,
                                     int batch, int L, int D, int W, int silu, int reverse, cudaStream_t st) {
  // 8 warps = 8 consecutive segments per block; segments of 32-64 positions (L = 513 -> 2 token groups of 8 x 33)
  const int n_tgrp = ceil_div(L, 512);
  const int seg = ceil_div(L, 8 * n_tgrp);
  const int n_cgrp = ceil_div(D, 64);
  const unsigned blocks = (unsigned)((int64_t)batch * n_cgrp * n_tgrp);
  const T* xx = reinterpret_cast<const T*>(x);
  T* dd = reinterpret_cast<T*>(dx);
  const TG *g1 = reinterpret_cast<const TG*>(g1_), *g2 = reinterpret_cast<const TG*>(g2_), *g3 = reinterpret_cast<const TG*>(g3_);
  if (g3 != nullptr)      conv1d_bwd_stream_kernel<T, TG, 3, UNR, MINB><<<blocks, 256, 0, st>>>(xx, ldx, w, bias, g1, g2, g3, ldd, dd, ld_dx, dw, dbias, L, D, W, silu, reverse, n_cgrp, n_tgrp, seg);
  else if (g2 != nullptr) conv1d_bwd_stream_kernel<T, TG, 2, UNR, MINB><<<blocks, 256, 0, st>>>(xx, ldx, w, bias, g1, g2, g3, ldd, dd, ld_dx, dw, dbias, L, D, W, silu, reverse, n_cgrp, n_tgrp, seg);
  else                    conv1d_bwd_stream_kernel<T, TG, 1, UNR, MINB><<<blocks, 256, 0, st>>>(xx, ldx, w, bias, g1, g2, g3, ldd, dd, ld_dx, dw, dbias, L, D, W, silu, reverse, n_cgrp, n_tgrp, seg);
}

// G16: the gradient terms are of the activation dtype T (else fp32)
template <typename T>
static void launch_conv_bwd_stream(const void* x, int64_t ldx, const float* w, const float* bias, const void* g1, const void* g2,
                                   const void* g3, int64_t ldd, bool g16, void* dx, int64_t ld_dx, float* dw, float* dbias,
                                   int batch, int L, int D, int W, int silu, int reverse, cudaStream_t st) {
  static int variant = -1, variant16 = -1;
  if (variant < 0) {
    const char* e = getenv("AUM_CONV_BWD_VARIANT"); variant = (e && atoi(e) == 1) ? 1 : 0;
    const char* f = getenv("AUM_CONV_BWD_VARIANT16"); variant16 = f ? atoi(f) : 0;
  }
  if (g16) {      // 16-bit terms.  Deeper unrolling measured SLOWER at AuM-Base size: 0.086 ms (2 positions per trip, 32 warps per
                  // SM) vs 0.104 (4 per trip) vs 0.123 (8 per trip, 24 warps) - AUM_CONV_BWD_VARIANT16=1 / 2 select those
    if constexpr (!std::is_same<T, float>::value) {
      if (variant16 == 0)      launch_conv_bwd_stream_v<T, T, 2, 4>(x, ldx, w, bias, g1, g2, g3, ldd, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st);
      else if (variant16 == 2) launch_conv_bwd_stream_v<T, T, 8, 3>(x, ldx, w, bias, g1, g2, g3, ldd, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st);
      else                     launch_conv_bwd_stream_v<T, T, 4, 4>(x, ldx, w, bias, g1, g2, g3, ldd, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st);
    }
    return;
  }
  if (variant == 1) launch_conv_bwd_stream_v<T, float, 4, 3>(x, ldx, w, bias, g1, g2, g3, ldd, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st);
  else              launch_conv_bwd_stream_v<T, float, 2, 4>(x, ldx, w, bias, g1, g2, g3, ldd, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st);
}

}  // namespace aum

extern "C" int aum_causal_conv1d_bwd(const void* x, int64_t ldx, const float* w, const float* bias,
                                     const void* dout_, const void* dout2_, const void* dout3_, int64_t ldd, int dout_dtype,
                                     void* dx, int64_t ld_dx,
                                     float* dw, float* dbias, int batch, int L, int D, int W,
                                     int dtype, int silu, int reverse, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(dx);
  AUM_REQUIRE(x && w && dout_ && dx && dw, "aum_causal_conv1d_bwd: null pointer");
  AUM_REQUIRE(dout_dtype == AUM_F32 || (dout_dtype == dtype && dtype != AUM_F32),
              "aum_causal_conv1d_bwd: the gradient terms must be fp32 or of the call's 16-bit dtype");
  const bool g16 = dout_dtype != AUM_F32;
  const float *dout = (const float*)dout_, *dout2 = (const float*)dout2_, *dout3 = (const float*)dout3_;   // (tile kernel: fp32 only)
  AUM_REQUIRE(W >= 2 && W <= CONV_MAXW, "aum_causal_conv1d_bwd: width %d unsupported (2..4)", W);
  AUM_REQUIRE(batch >= 0 && L >= 0 && D >= 0, "aum_causal_conv1d_bwd: negative size");
  AUM_REQUIRE(ldx >= D && ldd >= D && ld_dx >= D, "aum_causal_conv1d_bwd: leading dimension smaller than D");
  if (batch == 0 || L == 0 || D == 0) return 0;
  const int n_cgrp = ceil_div(D, 64), n_tgrp = ceil_div(L, 8 * CB_TL);
  const int64_t blocks = (int64_t)batch * n_cgrp * n_tgrp;
  AUM_REQUIRE(blocks < (1ll << 31), "aum_causal_conv1d_bwd: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // streaming kernel: channel pairs as 8-byte (fp32) / 4-byte (16-bit) vectors; a second gradient term must come with
    // the third slot free or filled, never third-only
    const int esz = dtype_size(dtype);
    auto al = [](const void* p, int a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
    static int force_tile = -1;
    if (force_tile < 0) force_tile = getenv("AUM_CONV_BWD_TILE") != nullptr ? 1 : 0;
    const bool stream_ok = !force_tile && D % 2 == 0 && ldx % 2 == 0 && ldd % 2 == 0 && ld_dx % 2 == 0 && al(x, 2 * esz) && al(dx, 2 * esz) &&
                           al(dout, g16 ? 4 : 8) && al(dout2, g16 ? 4 : 8) && al(dout3, g16 ? 4 : 8) && !(dout2 == nullptr && dout3 != nullptr) &&
                           (int64_t)batch * ceil_div(D, 64) * ceil_div(L, 512) < (1ll << 31) &&
                           ((int64_t)batch * L + 4) * (ldx > ldd ? (ldx > ld_dx ? ldx : ld_dx) : (ldd > ld_dx ? ldd : ld_dx)) < (1ll << 31);
    if (stream_ok) {
      switch (dtype) {
        case AUM_F32:  launch_conv_bwd_stream<float>(x, ldx, w, bias, dout, dout2, dout3, ldd, g16, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st); break;
        case AUM_F16:  launch_conv_bwd_stream<__half>(x, ldx, w, bias, dout, dout2, dout3, ldd, g16, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st); break;
        case AUM_BF16: launch_conv_bwd_stream<__nv_bfloat16>(x, ldx, w, bias, dout, dout2, dout3, ldd, g16, dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, st); break;
        default: set_error("aum_causal_conv1d_bwd: bad dtype %d", dtype); return 1;
      }
      return check_launch("aum_causal_conv1d_bwd");
    }
  }
  AUM_REQUIRE(!g16, "aum_causal_conv1d_bwd: 16-bit gradient terms need the streaming kernel (even D and pitches, W <= 4, 4-byte aligned rows)");
  switch (dtype) {
    case AUM_F32:  conv1d_bwd_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, ldx, w, bias, dout, dout2, dout3, ldd, (float*)dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, n_cgrp, n_tgrp); break;
    case AUM_F16:  conv1d_bwd_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)x, ldx, w, bias, dout, dout2, dout3, ldd, (__half*)dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, n_cgrp, n_tgrp); break;
    case AUM_BF16: conv1d_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, w, bias, dout, dout2, dout3, ldd, (__nv_bfloat16*)dx, ld_dx, dw, dbias, batch, L, D, W, silu, reverse, n_cgrp, n_tgrp); break;
    default: set_error("aum_causal_conv1d_bwd: bad dtype %d", dtype); return 1;
  }
  return check_launch("aum_causal_conv1d_bwd");
}
