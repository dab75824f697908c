// Depthwise causal conv1d (+bias +SiLU) over token-major activations.
// Replaces causal_conv1d_cuda.causal_conv1d_fwd (call sites: /root/reference/vim-mamba_ssm/mamba_ssm/ops/
// selective_scan_interface.py:177,239,318,380,463,532); semantics = mamba_simple.py:272 with padding W-1.
//
// HBM-bound streaming kernel: each thread owns 8 consecutive channels (one 16-byte vector for 16-bit
// dtypes) and walks TL consecutive tokens with a rolling (W-1)-deep window in registers, so every input row
// is read once per token tile (+ a 3-row halo) with fully coalesced 16 B accesses across the warp.
// Algorithmic bytes per (token, channel): read s + write s (s = itemsize).
#include "common.cuh"

namespace aum {

constexpr int CONV_TL = 8;    // tokens per thread
constexpr int CONV_MAXW = 4;

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
  float wk[CONV_MAXW][VEC];
  float bs[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c = c0 + v;
    const bool ok = c < D;
#pragma unroll
    for (int j = 0; j < CONV_MAXW; ++j) {
      const int k = j - (CONV_MAXW - W);
      wk[j][v] = (ok && k >= 0) ? w[(int64_t)c * W + k] : 0.f;
    }
    bs[v] = (ok && bias != nullptr) ? bias[c] : 0.f;
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

  float win[CONV_MAXW][VEC];   // win[j] = x at offset (j - (MAXW-1)) steps "behind" the current token
#pragma unroll
  for (int j = 0; j < CONV_MAXW - 1; ++j) {
    const int l = l_first - step * (CONV_MAXW - 1 - j);
    const bool in = (l >= 0 && l < L);
    if constexpr (VEC == 8) {
      if (in) { Vec8<T> t; t.load(xb + (int64_t)l * ldx); t.unpack(win[j]); }
      else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) win[j][v] = 0.f;
      }
    } else {
      win[j][0] = (in && c0 < D) ? to_f(xb[(int64_t)l * ldx]) : 0.f;
    }
  }

  for (int i = 0; i < n; ++i) {
    const int l = l_first + step * i;
    if constexpr (VEC == 8) { Vec8<T> t; t.load(xb + (int64_t)l * ldx); t.unpack(win[CONV_MAXW - 1]); }
    else win[CONV_MAXW - 1][0] = (c0 < D) ? to_f(xb[(int64_t)l * ldx]) : 0.f;
    float y[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float acc = bs[v];
#pragma unroll
      for (int j = 0; j < CONV_MAXW; ++j) acc = fmaf(wk[j][v], win[j][v], acc);
      y[v] = silu ? silu_f(acc) : acc;
    }
    if constexpr (VEC == 8) { Vec8<T> t; t.pack(y); t.store(ob + (int64_t)l * ldo); }
    else if (c0 < D) ob[(int64_t)l * ldo] = from_f<T>(y[0]);
#pragma unroll
    for (int j = 0; j < CONV_MAXW - 1; ++j)
#pragma unroll
      for (int v = 0; v < VEC; ++v) win[j][v] = win[j + 1][v];
  }
}

template <typename T>
static int launch_conv(const void* x, int64_t ldx, const float* w, const float* bias, void* out, int64_t ldo,
                       int batch, int L, int D, int W, int silu, int reverse, cudaStream_t st) {
  const int n_ltile = ceil_div(L, CONV_TL);
  const bool vec_ok = (D % 8 == 0) && (ldx % 8 == 0) && (ldo % 8 == 0) && aligned16(x) && aligned16(out);
  if (vec_ok) {
    const int n_cvec = D / 8;
    const int64_t total = (int64_t)batch * n_ltile * n_cvec;
    conv1d_fwd_kernel<T, 8><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(
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
