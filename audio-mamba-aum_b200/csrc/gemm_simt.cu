// fp32-accumulate CUDA-core GEMM  C[M,N] = A[M,K] @ W[N,K]^T  (+ shared epilogue).
// Role: the strict-parity (fp32 I/O) tier, shapes/alignments the tensor-core kernel does not take, and an
// independent cross-check of the tcgen05 kernel in the tests.  Not the production path for 16-bit inputs.
// 64x64x16 tiles, 256 threads, 4x4 micro-tile per thread, operands converted to fp32 in shared memory.
#include "gemm_common.cuh"

namespace aum {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ W, int64_t ldw,
                 EpiParams ep, int M, int N, int K) {
  __shared__ float sA[SG_BK][SG_BM + 4];
  __shared__ float sW[SG_BK][SG_BN + 4];
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tid = threadIdx.x;
  const int tr = tid / 16, tc = tid % 16;   // 16 x 16 thread grid, each 4 rows x 4 cols
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread; k fastest for coalescing
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      const int r = e / SG_BK, kk = e % SG_BK;
      const int gk = k0 + kk;
      const int gm = m0 + r, gn = n0 + r;
      sA[kk][r] = (gm < M && gk < K) ? to_f(A[(int64_t)gm * lda + gk]) : 0.f;
      sW[kk][r] = (gn < N && gk < K) ? to_f(W[(int64_t)gn * ldw + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][tr * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sW[kk][tc * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) epi_store1(ep, m0 + tr * 4 + i, n0 + tc * 4 + j, acc[i][j]);
}

int launch_gemm_simt(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dt, const EpiParams& ep,
                     int M, int N, int K, cudaStream_t st) {
  dim3 grid(ceil_div(N, SG_BN), ceil_div(M, SG_BM));
  if (grid.y > 65535) { set_error("aum_gemm_tn(simt): M too large"); return 1; }
  switch (ab_dt) {
    case AUM_F32:  gemm_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)A, lda, (const float*)W, ldw, ep, M, N, K); break;
    case AUM_F16:  gemm_simt_kernel<__half><<<grid, 256, 0, st>>>((const __half*)A, lda, (const __half*)W, ldw, ep, M, N, K); break;
    case AUM_BF16: gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, ldw, ep, M, N, K); break;
    default: set_error("aum_gemm_tn: bad ab_dtype %d", ab_dt); return 1;
  }
  return check_launch("aum_gemm_tn(simt)");
}

}  // namespace aum
