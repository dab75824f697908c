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

// MAXC: row chunks a lane may hold (32 lanes x 8 elements each); instantiated per width class so that a 768-wide row
// costs 24 value registers, not 64 (full occupancy = more loads in flight for this latency-bound kernel).
template <typename T, typename TY, typename RT, int MAXC>
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
  float v[MAXC][8];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
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
  for (int c = 0; c < MAXC; ++c) {
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
#define AUM_RN_LAUNCH(C) add_rmsnorm_vec_kernel<T, TY, float, C><<<ceil_div(rows, warps_per_block), warps_per_block * 32, 0, st>>>( \
      (const T*)x, ldx, (const float*)rin, ldr, w, b, (TY*)y, ldy, (float*)rout, ldro, rstd, rows, dim, eps)
  const int per_lane = ceil_div(dim >> 3, 32);
  if (per_lane <= 2) AUM_RN_LAUNCH(2);
  else if (per_lane <= 3) AUM_RN_LAUNCH(3);
  else if (per_lane <= 4) AUM_RN_LAUNCH(4);
  else AUM_RN_LAUNCH(RN_MAXC);
#undef AUM_RN_LAUNCH
}

}  // namespace aum

extern "C" int aum_add_rmsnorm_fwd(const void* x, int64_t ldx, int x_dtype,
                                   const void* residual_in, int64_t ldr, int r_dtype,
                                   const float* weight, const float* bias,
                                   void* y, int64_t ldy, int y_dtype,
                                   void* residual_out, int64_t ldro, int ro_dtype,
                                   float* rstd_out, int rows, int dim, float eps, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(y);
  if (rows == 0) return 0;                            // empty input: nothing to do (pointers may be null)
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

// ======================================================================================================
// Backward of fused add + RMSNorm (prenorm, fp32 residual stream).
// Replaces the Triton _layer_norm_bwd_kernel (/root/reference/vim-mamba_ssm/mamba_ssm/ops/triton/layernorm.py:196-290)
// for is_rms_norm=True, no bias.  With r = x + residual_in (saved), xhat = r*rstd, wdy = dy*w:
//   dr = (wdy - xhat * mean(xhat*wdy)) * rstd + dresidual_out ;  dx = dr (x dtype) ;  dresidual_in = dr (fp32)
//   dw += sum_rows dy * xhat
// One warp per row, 16 consecutive rows per warp with the dw partials of the warp kept in registers, one
// shared-memory reduction over the block's 8 warps and one atomicAdd per weight element per block.
// ======================================================================================================
namespace aum {

// rows per warp.  16 at first: 129 blocks of 8 warps for the 16 416 rows of a config-3 launch - fewer blocks than SMs, ~7 warps
// per SM, each walking its rows one after the other: latency-bound at 54 % of HBM.  4 rows per warp = 513 blocks; the
// shared-memory reduction + 768 atomics per block stay small against 32 rows of traffic.  (AUM_RNB_ROWS at build time.)
#ifndef AUM_RNB_ROWS
#define AUM_RNB_ROWS 4
#endif
constexpr int RNB_ROWS = AUM_RNB_ROWS;

template <typename TD, int NC>
__global__ void __launch_bounds__(256)
add_rmsnorm_bwd_kernel(const TD* __restrict__ dy, int64_t ld_dy, const float* __restrict__ dro, int64_t ld_dro,
                       const float* __restrict__ r, int64_t ld_r, const float* __restrict__ rstd,
                       const float* __restrict__ weight, TD* __restrict__ dx, int64_t ld_dx,
                       float* __restrict__ dri, int64_t ld_dri, float* __restrict__ dweight, int rows, int dim) {
  __shared__ float red[8][8 * 32];                      // per-warp dw partials of one 256-column slab
  const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = dim >> 3;
  const int row0 = (blockIdx.x * 8 + wrp) * RNB_ROWS;
  float dw[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) dw[c][i] = 0.f;
  float wv[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) { Vec8<float> t; t.load(weight + ch * 8); t.unpack(wv[c]); }
  }
  for (int rr = 0; rr < RNB_ROWS; ++rr) {
    const int row = row0 + rr;
    if (row >= rows) break;                               // warp-uniform
    const float rs = rstd[row];
    float xh[NC][8], wdy[NC][8];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        Vec8<TD> g; g.load(dy + (int64_t)row * ld_dy + ch * 8);
        float gf[8]; g.unpack(gf);
        Vec8<float> rv; rv.load(r + (int64_t)row * ld_r + ch * 8); rv.unpack(xh[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[c][i] *= rs;
          wdy[c][i] = gf[i] * wv[c][i];
          dot = fmaf(xh[c][i], wdy[c][i], dot);
          dw[c][i] = fmaf(gf[i], xh[c][i], dw[c][i]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float c1 = dot / (float)dim;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float dr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dr[i] = (wdy[c][i] - xh[c][i] * c1) * rs;
        if (dro != nullptr) {
          Vec8<float> t; t.load(dro + (int64_t)row * ld_dro + ch * 8);
          float tf[8]; t.unpack(tf);
#pragma unroll
          for (int i = 0; i < 8; ++i) dr[i] += tf[i];
        }
        { Vec8<TD> t; t.pack(dr); t.store(dx + (int64_t)row * ld_dx + ch * 8); }
        if (dri != nullptr) { Vec8<float> t; t.pack(dr); t.store(dri + (int64_t)row * ld_dri + ch * 8); }
      }
    }
  }
  // dweight: reduce the 8 warps of the block slab by slab (a slab = 32 lanes x 8 columns)
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (32 * c < nchunk) {                               // block-uniform
#pragma unroll
      for (int i = 0; i < 8; ++i) red[wrp][lane * 8 + i] = dw[c][i];
      __syncthreads();
      const int col = threadIdx.x;                       // 256 columns of this slab
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][col];
      const int gcol = c * 256 + col;
      if (gcol < dim) atomicAdd(dweight + gcol, s);
      __syncthreads();
    }
  }
}

}  // namespace aum

extern "C" int aum_add_rmsnorm_bwd(const void* dy, int64_t ld_dy, int dy_dtype,
                                   const float* dresidual_out, int64_t ld_dro,
                                   const float* r, int64_t ld_r, const float* rstd, const float* weight,
                                   void* dx, int64_t ld_dx, float* dresidual_in, int64_t ld_dri,
                                   float* dweight, int rows, int dim, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(dx);
  AUM_REQUIRE(dy && r && rstd && weight && dx && dweight, "aum_add_rmsnorm_bwd: null pointer");
  AUM_REQUIRE(rows >= 0 && dim > 0, "aum_add_rmsnorm_bwd: bad shape");
  AUM_REQUIRE(dim % 8 == 0 && dim <= 8 * 32 * RN_MAXC, "aum_add_rmsnorm_bwd: dim must be a multiple of 8 and <= %d", 8 * 32 * RN_MAXC);
  AUM_REQUIRE(ld_dy % 8 == 0 && ld_r % 8 == 0 && ld_dx % 8 == 0 && (!dresidual_out || ld_dro % 8 == 0) &&
              (!dresidual_in || ld_dri % 8 == 0) && aligned16(dy) && aligned16(r) && aligned16(dx) && aligned16(weight) &&
              (!dresidual_out || aligned16(dresidual_out)) && (!dresidual_in || aligned16(dresidual_in)),
              "aum_add_rmsnorm_bwd: 16-byte aligned rows required");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = ceil_div(rows, 8 * RNB_ROWS);
  const int nc = ceil_div(dim, 256);
#define AUM_RNB(TD, NCV) add_rmsnorm_bwd_kernel<TD, NCV><<<blocks, 256, 0, st>>>((const TD*)dy, ld_dy, dresidual_out, ld_dro, r, ld_r, rstd, weight, (TD*)dx, ld_dx, dresidual_in, ld_dri, dweight, rows, dim)
#define AUM_RNB_NC(TD) do { if (nc <= 1) AUM_RNB(TD, 1); else if (nc <= 2) AUM_RNB(TD, 2); else if (nc <= 3) AUM_RNB(TD, 3); else if (nc <= 4) AUM_RNB(TD, 4); else AUM_RNB(TD, 8); } while (0)
  switch (dy_dtype) {
    case AUM_F32:  AUM_RNB_NC(float); break;
    case AUM_F16:  AUM_RNB_NC(__half); break;
    case AUM_BF16: AUM_RNB_NC(__nv_bfloat16); break;
    default: set_error("aum_add_rmsnorm_bwd: bad dtype %d", dy_dtype); return 1;
  }
#undef AUM_RNB_NC
#undef AUM_RNB
  return check_launch("aum_add_rmsnorm_bwd");
}
