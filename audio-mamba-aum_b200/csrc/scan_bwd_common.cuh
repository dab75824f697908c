// Shared definitions of the selective-scan backward kernels (generic: scan_bwd.cu, TMA-streamed: scan_bwd_tma.cu).
#pragma once
#include "scan_common.cuh"

namespace aum {

constexpr int SB_CH = 64;     // channels per CTA
constexpr int SB_TT = 8;      // checkpoint interval / chunk length
constexpr int SB_WARPS = SB_CH / 32;   // warps per direction group

struct ScanBwdDirDev {
  const void* u; int64_t ld_u;
  const void* delta; int64_t ld_delta;
  const float* A;
  const float* BC; int64_t ld_bc;
  const float* D;
  float* du; int64_t ld_du;
  float* ddelta; int64_t ld_dd;
  float* dA; float* dD;
  float* dBC; int64_t ld_dbc;
  float* dbc_ws;      // [part][batch*L][32] per-warp partial sums of dB|dC (no atomics), reduced by a second kernel
  float* ckpt;
  int ckpt_valid;
  int reverse;
  int dA_log;         // accumulate dA * A (the gradient w.r.t. A_log) instead of dA
};

struct ScanBwdParams {
  ScanBwdDirDev dir[2];
  int ndirs, shared_du;
  int g16;            // du / ddelta are of the activation dtype (specialised TMA kernel only), else fp32
  int d16;            // delta is of the activation dtype (TMA kernels only), else fp32
  const void* z; int64_t ld_z;
  const void* ypre; int64_t ld_y;
  const void* dout; int64_t ld_dout;
  void* dz; int64_t ld_dz;
  void* outz; int64_t ld_oz;
  int batch, L, Dch, nchunks;
  float scale;
  int softplus_grad;
};

// 32 values per lane -> lane i ends up with sum over the warp's lanes of value i (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = upper ? v[i + half] : v[i];
      const float send = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// TMA-streamed fast path (scan_bwd_tma.cu): -1 when the launch is not eligible (caller runs the generic kernel),
// 0 on success, > 0 on error.  Launches the kernel only; the caller reduces the dB|dC partial workspace.
int launch_scan_bwd_tma(const ScanBwdParams& p, int dtype, cudaStream_t st);

}  // namespace aum
