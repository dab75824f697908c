// Batched 2-D transpose with dtype conversion: the layout adapter between the reference's channel-major
// (batch, C, L) tensors (selective_scan_interface.py:458-461 xz, :224 out_z) and this engine's token-major
// (batch, L, C) activations.  Classic 32x32 shared-memory tile (+1 padding), coalesced on both sides.
#include "common.cuh"

namespace aum {

template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
transpose_kernel(const TS* __restrict__ src, int64_t src_bs, int64_t src_ld,
                 TD* __restrict__ dst, int64_t dst_bs, int64_t dst_ld, int R, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const TS* s = src + (int64_t)b * src_bs;
  TD* d = dst + (int64_t)b * dst_bs;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < R && c < C) tile[ty + 8 * k][tx] = to_f(s[(int64_t)r * src_ld + c]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < R && c < C) d[(int64_t)c * dst_ld + r] = from_f<TD>(tile[tx][ty + 8 * k]);
  }
}

template <typename TS, typename TD>
static void launch_t(const void* src, int64_t sbs, int64_t sld, void* dst, int64_t dbs, int64_t dld,
                     int batch, int R, int C, cudaStream_t st) {
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32), batch);
  transpose_kernel<TS, TD><<<grid, 256, 0, st>>>((const TS*)src, sbs, sld, (TD*)dst, dbs, dld, R, C);
}

template <typename TS>
static int dispatch_dst(const void* src, int64_t sbs, int64_t sld, void* dst, int64_t dbs, int64_t dld,
                        int batch, int R, int C, int dd, cudaStream_t st) {
  switch (dd) {
    case AUM_F32:  launch_t<TS, float>(src, sbs, sld, dst, dbs, dld, batch, R, C, st); return 0;
    case AUM_F16:  launch_t<TS, __half>(src, sbs, sld, dst, dbs, dld, batch, R, C, st); return 0;
    case AUM_BF16: launch_t<TS, __nv_bfloat16>(src, sbs, sld, dst, dbs, dld, batch, R, C, st); return 0;
  }
  set_error("aum_transpose: bad dst dtype %d", dd);
  return 1;
}

}  // namespace aum

extern "C" int aum_transpose(const void* src, int64_t src_bs, int64_t src_ld,
                             void* dst, int64_t dst_bs, int64_t dst_ld,
                             int batch, int R, int C, int src_dtype, int dst_dtype, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(dst);
  if (batch == 0 || R == 0 || C == 0) return 0;
  AUM_REQUIRE(src && dst, "aum_transpose: null pointer");
  AUM_REQUIRE(batch >= 0 && R >= 0 && C >= 0, "aum_transpose: negative size");
  AUM_REQUIRE(batch <= 65535 && ceil_div(R, 32) <= 65535, "aum_transpose: grid too large");
  if (batch == 0 || R == 0 || C == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 1;
  switch (src_dtype) {
    case AUM_F32:  rc = dispatch_dst<float>(src, src_bs, src_ld, dst, dst_bs, dst_ld, batch, R, C, dst_dtype, st); break;
    case AUM_F16:  rc = dispatch_dst<__half>(src, src_bs, src_ld, dst, dst_bs, dst_ld, batch, R, C, dst_dtype, st); break;
    case AUM_BF16: rc = dispatch_dst<__nv_bfloat16>(src, src_bs, src_ld, dst, dst_bs, dst_ld, batch, R, C, dst_dtype, st); break;
    default: set_error("aum_transpose: bad src dtype %d", src_dtype); return 1;
  }
  if (rc) return rc;
  return check_launch("aum_transpose");
}
