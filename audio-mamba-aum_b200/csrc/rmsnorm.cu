// Fused residual-add + RMSNorm forward (fp32 residual stream).
// Replaces the Triton _layer_norm_fwd_1pass_kernel (/root/reference/vim-mamba_ssm/mamba_ssm/ops/triton/
// layernorm.py:65-120) as used through rms_norm_fn (:477) by src/models/mamba_models.py:77-97,646-657;
// semantics = rms_norm_ref (:35-48) with upcast.
//
// HBM-bound: one warp per token row, the row held in registers between the sum-of-squares pass and the
// scale pass, 16-byte accesses.  Algorithmic bytes per row: dim*(s_x + 4 [res in] + 4 [res out] + s_y).
#include "common.cuh"

namespace aum {

constexpr int RN_MAXC = 8;   // 8 chunks * 32 lanes * 8 elems = dim <= 2048 on the vector path

template <typename T, typename TY, typename RT>
__global__ void __launch_bounds__(256)
add_rmsnorm_vec_kernel(const T* __restrict__ x, int64_t ldx, const RT* __restrict__ rin, int64_t ldr,
                       const float* __restrict__ weight, const float* __restrict__ bias,
                       TY* __restrict__ y, int64_t ldy, RT* __restrict__ rout, int64_t ldro,
                       float* __restrict__ rstd_out, int rows, int dim, float eps) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nchunk = dim >> 3;
  const T* xr = x + (int64_t)warp * ldx;
  float v[RN_MAXC][8];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < RN_MAXC; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      Vec8<T> t; t.load(xr + ch * 8); t.unpack(v[c]);
      if (rin != nullptr) {
        Vec8<RT> r; r.load(rin + (int64_t)warp * ldr + ch * 8);
        float rf[8]; r.unpack(rf);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c][i] += rf[i];
      }
      if (rout != nullptr) { Vec8<RT> r; r.pack(v[c]); r.store(rout + (int64_t)warp * ldro + ch * 8); }
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(v[c][i], v[c][i], ss);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)dim + eps);
  if (rstd_out != nullptr && lane == 0) rstd_out[warp] = rstd;
#pragma unroll
  for (int c = 0; c < RN_MAXC; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      Vec8<float> wv; wv.load(weight + ch * 8);
      float wf[8]; wv.unpack(wf);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = v[c][i] * rstd * wf[i];
      if (bias != nullptr) {
        Vec8<float> bv; bv.load(bias + ch * 8);
        float bf[8]; bv.unpack(bf);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += bf[i];
      }
      Vec8<TY> t; t.pack(o); t.store(y + (int64_t)warp * ldy + ch * 8);
    }
  }
}

// Generic fallback: any dim / alignment / dtype mix; one block per row, two passes over global memory.
__global__ void __launch_bounds__(256)
add_rmsnorm_generic_kernel(const void* x, int64_t ldx, int x_dt, const void* rin, int64_t ldr, int r_dt,
                           const float* weight, const float* bias, void* y, int64_t ldy, int y_dt,
                           void* rout, int64_t ldro, int ro_dt, float* rstd_out, int rows, int dim, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  float ss = 0.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    float v = load_as_f(x, (int64_t)row * ldx + i, x_dt);
    if (rin) v += load_as_f(rin, (int64_t)row * ldr + i, r_dt);
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float rstd = rsqrtf(red[0] / (float)dim + eps);
  if (rstd_out && threadIdx.x == 0) rstd_out[row] = rstd;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    float v = load_as_f(x, (int64_t)row * ldx + i, x_dt);
    if (rin) v += load_as_f(rin, (int64_t)row * ldr + i, r_dt);
    if (rout) store_from_f(rout, (int64_t)row * ldro + i, ro_dt, v);
    float o = v * rstd * weight[i];
    if (bias) o += bias[i];
    store_from_f(y, (int64_t)row * ldy + i, y_dt, o);
  }
}

template <typename T, typename TY>
static void launch_vec(const void* x, int64_t ldx, const void* rin, int64_t ldr, const float* w, const float* b,
                       void* y, int64_t ldy, void* rout, int64_t ldro, float* rstd, int rows, int dim, float eps,
                       cudaStream_t st) {
  const int warps_per_block = 8;
  add_rmsnorm_vec_kernel<T, TY, float><<<ceil_div(rows, warps_per_block), warps_per_block * 32, 0, st>>>(
      (const T*)x, ldx, (const float*)rin, ldr, w, b, (TY*)y, ldy, (float*)rout, ldro, rstd, rows, dim, eps);
}

}  // namespace aum

extern "C" int aum_add_rmsnorm_fwd(const void* x, int64_t ldx, int x_dtype,
                                   const void* residual_in, int64_t ldr, int r_dtype,
                                   const float* weight, const float* bias,
                                   void* y, int64_t ldy, int y_dtype,
                                   void* residual_out, int64_t ldro, int ro_dtype,
                                   float* rstd_out, int rows, int dim, float eps, void* stream) {
  using namespace aum;
  AUM_REQUIRE(x && weight && y, "aum_add_rmsnorm_fwd: null pointer");
  AUM_REQUIRE(rows >= 0 && dim > 0, "aum_add_rmsnorm_fwd: bad shape rows=%d dim=%d", rows, dim);
  AUM_REQUIRE(ldx >= dim && ldy >= dim, "aum_add_rmsnorm_fwd: leading dimension smaller than dim");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = (dim % 8 == 0) && dim <= 8 * 32 * RN_MAXC && (x_dtype == y_dtype || x_dtype == AUM_F32) &&
                      (!residual_in || r_dtype == AUM_F32) && (!residual_out || ro_dtype == AUM_F32) &&
                      ldx % 8 == 0 && ldy % 8 == 0 && (!residual_in || ldr % 8 == 0) &&
                      (!residual_out || ldro % 8 == 0) && aligned16(x) && aligned16(y) &&
                      aligned16(weight) && (!bias || aligned16(bias)) &&
                      (!residual_in || aligned16(residual_in)) && (!residual_out || aligned16(residual_out));
  if (vec_ok) {
#define AUM_RN_ARGS x, ldx, residual_in, ldr, weight, bias, y, ldy, residual_out, ldro, rstd_out, rows, dim, eps, st
    if (x_dtype == AUM_F32 && y_dtype == AUM_F32) launch_vec<float, float>(AUM_RN_ARGS);
    else if (x_dtype == AUM_F32 && y_dtype == AUM_F16) launch_vec<float, __half>(AUM_RN_ARGS);
    else if (x_dtype == AUM_F32 && y_dtype == AUM_BF16) launch_vec<float, __nv_bfloat16>(AUM_RN_ARGS);
    else if (x_dtype == AUM_F16) launch_vec<__half, __half>(AUM_RN_ARGS);
    else if (x_dtype == AUM_BF16) launch_vec<__nv_bfloat16, __nv_bfloat16>(AUM_RN_ARGS);
    else { set_error("aum_add_rmsnorm_fwd: bad dtype %d", x_dtype); return 1; }
#undef AUM_RN_ARGS
  } else {
    add_rmsnorm_generic_kernel<<<rows, 256, 0, st>>>(x, ldx, x_dtype, residual_in, ldr, r_dtype, weight, bias,
                                                     y, ldy, y_dtype, residual_out, ldro, ro_dtype, rstd_out,
                                                     rows, dim, eps);
  }
  return check_launch("aum_add_rmsnorm_fwd");
}
