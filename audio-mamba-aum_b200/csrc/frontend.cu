// Front end of the block stack: the two data-movement steps either side of the patch-embed GEMM.
//
//  aum_patchify        spectrogram (B, T, F) fp32 -> im2col rows (B * gf * gt, pf * pt) in the activation dtype: the
//                      stride == kernel conv2d of FlexiPatchEmbed (/root/reference/src/utilities/tokenization.py:278-310,
//                      called with img[b, 0, f, t] = x[b, t, f], src/models/mamba_models.py:510-515) becomes one GEMM
//                      with the (Dm, 1, pf, pt) conv weight flattened to (Dm, pf * pt).  Row m = (b, f_blk, t_blk)
//                      in the order of `x.flatten(2).transpose(1, 2)` (f_blk major), column k = kf * pt + kt.
//  aum_assemble_tokens patch tokens (B, N, Dm) fp32 (+ conv bias already added by the GEMM) -> hidden (B, N + 1, Dm)
//                      fp32 with the cls token inserted at index N / 2 and the absolute position embedding added
//                      (mamba_models.py:525-541; tokenization.py:414-451: slot 0 of pos_embed belongs to the cls token).
// Both are HBM-bound copies; one launch each instead of the permute / cast / two adds / cls write of the torch version.
#include "common.cuh"

namespace aum {

// one thread per 4 consecutive kt of one (row m, kf): reads 4 strided floats x[b, t0 + kt.., f0 + kf], writes 4 packed
template <typename T>
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ x, T* __restrict__ cols, int64_t total4, int T_, int F_, int pf, int pt, int gf, int gt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int pt4 = pt >> 2;
  const int k4 = (int)(i % pt4);
  int64_t r = i / pt4;
  const int kf = (int)(r % pf); r /= pf;
  const int tb = (int)(r % gt); r /= gt;
  const int fb = (int)(r % gf);
  const int b = (int)(r / gf);
  const int t0 = tb * pt + 4 * k4, f = fb * pf + kf;
  const float* src = x + ((int64_t)b * T_ + t0) * F_ + f;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = __ldg(src + (int64_t)j * F_);
  const int64_t m = ((int64_t)b * gf + fb) * gt + tb;
  T* dst = cols + m * (int64_t)(pf * pt) + kf * pt + 4 * k4;
#pragma unroll
  for (int j = 0; j < 4; ++j) dst[j] = from_f<T>(v[j]);
}

__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const float4* __restrict__ tok, const float4* __restrict__ pos, const float4* __restrict__ cls,
                       float4* __restrict__ hidden, int64_t total4, int N, int Dm4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c = (int)(i % Dm4);
  const int64_t row = i / Dm4;
  const int t = (int)(row % (N + 1));
  const int64_t b = row / (N + 1);
  const int tp = N / 2;
  float4 a, p;
  if (t == tp) { a = __ldg(cls + c); p = __ldg(pos + c); }                        // cls token + its slot 0
  else {
    const int n = t < tp ? t : t - 1;                                             // patch index
    a = __ldg(tok + (b * N + n) * Dm4 + c);
    p = __ldg(pos + (int64_t)(t < tp ? t + 1 : t) * Dm4 + c);
  }
  hidden[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}

}  // namespace aum

extern "C" int aum_patchify(const float* x, void* cols, int batch, int T_, int F_, int pf, int pt, int dtype, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(cols);
  if (batch == 0) return 0;
  AUM_REQUIRE(x && cols, "aum_patchify: null pointer");
  AUM_REQUIRE(batch > 0 && pf > 0 && pt > 0 && pt % 4 == 0 && F_ % pf == 0 && T_ % pt == 0,
              "aum_patchify: need F %% pf == 0, T %% pt == 0, pt %% 4 == 0 (got F=%d T=%d pf=%d pt=%d)", F_, T_, pf, pt);
  const int gf = F_ / pf, gt = T_ / pt;
  const int64_t total4 = (int64_t)batch * gf * gt * pf * (pt / 4);
  const unsigned blocks = (unsigned)ceil_div64(total4, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case AUM_F32:  patchify_kernel<float><<<blocks, 256, 0, st>>>(x, (float*)cols, total4, T_, F_, pf, pt, gf, gt); break;
    case AUM_F16:  patchify_kernel<__half><<<blocks, 256, 0, st>>>(x, (__half*)cols, total4, T_, F_, pf, pt, gf, gt); break;
    case AUM_BF16: patchify_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(x, (__nv_bfloat16*)cols, total4, T_, F_, pf, pt, gf, gt); break;
    default: set_error("aum_patchify: bad dtype %d", dtype); return 1;
  }
  return check_launch("aum_patchify");
}

extern "C" int aum_assemble_tokens(const float* tok, const float* pos, const float* cls, float* hidden,
                                   int batch, int N, int Dm, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(hidden);
  if (batch == 0) return 0;
  AUM_REQUIRE(tok && pos && cls && hidden, "aum_assemble_tokens: null pointer");
  AUM_REQUIRE(batch > 0 && N > 0 && Dm > 0 && Dm % 4 == 0, "aum_assemble_tokens: Dm must be a multiple of 4");
  AUM_REQUIRE(aligned16(tok) && aligned16(pos) && aligned16(cls) && aligned16(hidden), "aum_assemble_tokens: 16-byte aligned buffers required");
  const int64_t total4 = (int64_t)batch * (N + 1) * (Dm / 4);
  assemble_tokens_kernel<<<(unsigned)ceil_div64(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)tok, (const float4*)pos, (const float4*)cls, (float4*)hidden, total4, N, Dm / 4);
  return check_launch("aum_assemble_tokens");
}
