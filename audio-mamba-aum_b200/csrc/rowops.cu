// Row-wise helper of the training path:  out[r, c] = cast(a[r, c] + b[r, c]),  colsum[c] += sum_r (a + b)[r, c].
// One pass instead of the three torch launches it replaces in BiMambaInnerFn.backward's dt_proj chain
// (/root/reference/vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py:566-586: ddelta of the two directions summed
// (:556), delta_proj bias gradient = its sum over tokens, and the 16-bit operand of the d(dt_proj.weight) / d(x_dbl)
// products).  HBM-bound: reads 4 (+4) bytes - or 2 (+2) when the backward scan left ddelta in the activation dtype - and
// writes s bytes per element.
#include "common.cuh"

namespace aum {

constexpr int RO_ROWS = 64;      // rows per block
constexpr int RO_THREADS = 256;  // 64 column quads x 4 row lanes -> 256 columns per block

// 4 adjacent elements of a row as fp32
template <typename TI> __device__ __forceinline__ float4 ld4(const TI* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 ld4<__half>(const __half* p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
}

// Block = 64 column quads (256 columns) x 4 row lanes over RO_ROWS rows: a thread walks every fourth row of the block's row
// range with all its loads of one trip in flight (8 rows x 1-2 inputs), the four row lanes' column sums meet in shared
// memory, one atomic per column and block.  (The first version gave every thread 64 consecutive rows: 514 blocks = 28 warps
// per SM, 43 % of HBM on 16-bit inputs.)
constexpr int RO_LANES = 4;
template <typename T, typename TI>
__global__ void __launch_bounds__(RO_THREADS)
sum_cast_colsum_kernel(const TI* __restrict__ a, const TI* __restrict__ b, int64_t ld, T* __restrict__ out, int64_t ldo,
                       float* __restrict__ colsum, int rows, int cols) {
  __shared__ float4 red[RO_LANES][RO_THREADS / RO_LANES];
  const int cq = threadIdx.x % (RO_THREADS / RO_LANES), rl = threadIdx.x / (RO_THREADS / RO_LANES);
  const int c0 = (blockIdx.x * (RO_THREADS / RO_LANES) + cq) * 4;
  const bool ok = c0 < cols;
  const int r0 = blockIdx.y * RO_ROWS, r1 = min(rows, r0 + RO_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
#pragma unroll 8
    for (int r = r0 + rl; r < r1; r += RO_LANES) {
      float4 v = ld4<TI>(a + (int64_t)r * ld + c0);
      if (b != nullptr) {
        const float4 w = ld4<TI>(b + (int64_t)r * ld + c0);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      T* o = out + (int64_t)r * ldo + c0;
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(o) = v;
      } else {
        const T t0 = from_f<T>(v.x), t1 = from_f<T>(v.y), t2 = from_f<T>(v.z), t3 = from_f<T>(v.w);
        uint2 pk;
        pk.x = (uint32_t)(*reinterpret_cast<const unsigned short*>(&t0)) | ((uint32_t)(*reinterpret_cast<const unsigned short*>(&t1)) << 16);
        pk.y = (uint32_t)(*reinterpret_cast<const unsigned short*>(&t2)) | ((uint32_t)(*reinterpret_cast<const unsigned short*>(&t3)) << 16);
        *reinterpret_cast<uint2*>(o) = pk;
      }
    }
  }
  if (colsum != nullptr) {            // (uniform across the block)
    red[rl][cq] = s;
    __syncthreads();
    if (rl == 0 && ok) {
#pragma unroll
      for (int k = 1; k < RO_LANES; ++k) { const float4 t = red[k][cq]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
      atomicAdd(colsum + c0, s.x); atomicAdd(colsum + c0 + 1, s.y); atomicAdd(colsum + c0 + 2, s.z); atomicAdd(colsum + c0 + 3, s.w);
    }
  }
}

}  // namespace aum

namespace aum {
template <typename TI>
static void launch_sum_cast(const void* a, const void* b, int64_t ld, void* out, int64_t ld_out, int out_dtype, float* colsum,
                            int rows, int cols, cudaStream_t st) {
  dim3 grid(ceil_div(cols, (RO_THREADS / RO_LANES) * 4), ceil_div(rows, RO_ROWS));
  const TI *aa = reinterpret_cast<const TI*>(a), *bb = reinterpret_cast<const TI*>(b);
  switch (out_dtype) {
    case AUM_F32:  sum_cast_colsum_kernel<float, TI><<<grid, RO_THREADS, 0, st>>>(aa, bb, ld, (float*)out, ld_out, colsum, rows, cols); break;
    case AUM_F16:  sum_cast_colsum_kernel<__half, TI><<<grid, RO_THREADS, 0, st>>>(aa, bb, ld, (__half*)out, ld_out, colsum, rows, cols); break;
    default:       sum_cast_colsum_kernel<__nv_bfloat16, TI><<<grid, RO_THREADS, 0, st>>>(aa, bb, ld, (__nv_bfloat16*)out, ld_out, colsum, rows, cols); break;
  }
}
}  // namespace aum

extern "C" int aum_sum_cast_colsum(const void* a, const void* b, int64_t ld, int in_dtype, void* out, int64_t ld_out, int out_dtype,
                                   float* colsum, int rows, int cols, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(out);
  if (rows == 0 || cols == 0) return 0;
  AUM_REQUIRE(a && out, "aum_sum_cast_colsum: null pointer");
  AUM_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, "aum_sum_cast_colsum: cols must be a positive multiple of 4");
  AUM_REQUIRE(out_dtype >= AUM_F32 && out_dtype <= AUM_BF16, "aum_sum_cast_colsum: bad dtype %d", out_dtype);
  AUM_REQUIRE(in_dtype >= AUM_F32 && in_dtype <= AUM_BF16, "aum_sum_cast_colsum: bad input dtype %d", in_dtype);
  AUM_REQUIRE(ld >= cols && ld_out >= cols, "aum_sum_cast_colsum: leading dimension smaller than the row length");
  const int osz = dtype_size(out_dtype), isz = dtype_size(in_dtype);
  auto al = [](const void* p, int n) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % n == 0; };
  AUM_REQUIRE(al(a, 4 * isz) && al(b, 4 * isz) && ld % 4 == 0 && al(out, 4 * osz) && ld_out % 4 == 0,
              "aum_sum_cast_colsum: rows must be addressable in 4-element vectors");
  cudaStream_t st = (cudaStream_t)stream;
  switch (in_dtype) {
    case AUM_F32:  launch_sum_cast<float>(a, b, ld, out, ld_out, out_dtype, colsum, rows, cols, st); break;
    case AUM_F16:  launch_sum_cast<__half>(a, b, ld, out, ld_out, out_dtype, colsum, rows, cols, st); break;
    default:       launch_sum_cast<__nv_bfloat16>(a, b, ld, out, ld_out, out_dtype, colsum, rows, cols, st); break;
  }
  return check_launch("aum_sum_cast_colsum");
}
