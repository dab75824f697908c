// aum_gemm_tn: C-ABI front end of the projection GEMMs (see include/aum_b200.h).
#include "gemm_common.cuh"

extern "C" int aum_gemm_tn(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dtype,
                           void* C, int64_t ldc, int c_dtype,
                           void* C2, int64_t ldc2, int c2_dtype, int split,
                           int M, int N, int K,
                           const float* bias, const float* row_scale, int act,
                           int backend, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(C);
  if (M == 0 || N == 0) return 0;                     // empty output: nothing to do (pointers may be null)
  AUM_REQUIRE(A && W && C, "aum_gemm_tn: null pointer");
  AUM_REQUIRE(M >= 0 && N >= 0 && K >= 0, "aum_gemm_tn: negative size");
  AUM_REQUIRE(ab_dtype >= AUM_F32 && ab_dtype <= AUM_BF16, "aum_gemm_tn: bad ab_dtype %d", ab_dtype);
  AUM_REQUIRE(c_dtype >= AUM_F32 && c_dtype <= AUM_BF16, "aum_gemm_tn: bad c_dtype %d", c_dtype);
  AUM_REQUIRE(lda >= K && ldw >= K, "aum_gemm_tn: lda/ldw smaller than K");
  const int act_kind = act & 0xff, act_col0 = act >> 8;
  AUM_REQUIRE(act_kind == AUM_ACT_NONE || act_kind == AUM_ACT_SOFTPLUS || act_kind == AUM_ACT_SILU, "aum_gemm_tn: bad activation %d", act);
  if (C2 == nullptr) { split = N; ldc2 = 0; c2_dtype = c_dtype; }
  AUM_REQUIRE(split >= 0 && split <= N, "aum_gemm_tn: split %d must lie in [0,N]", split);
  AUM_REQUIRE(ldc >= split, "aum_gemm_tn: ldc smaller than its column count");
  AUM_REQUIRE(C2 == nullptr || (ldc2 >= N - split && c2_dtype >= AUM_F32 && c2_dtype <= AUM_BF16), "aum_gemm_tn: bad second output");
  if (M == 0 || N == 0) return 0;
  AUM_REQUIRE(K > 0, "aum_gemm_tn: K must be positive");

  EpiParams ep;
  ep.C = C; ep.ldc = ldc; ep.c_dt = c_dtype;
  ep.C2 = C2; ep.ldc2 = ldc2; ep.c2_dt = c2_dtype; ep.split = split;
  ep.bias = bias; ep.row_scale = row_scale; ep.act = act_kind; ep.act_col0 = act_col0;
  ep.M = M; ep.N = N;
  auto ok16 = [](const void* p, int64_t ld, int dt) {
    return aligned16(p) && ((ld * dtype_size(dt)) % 16 == 0);
  };
  ep.vec_ok = ok16(C, ldc, c_dtype) && (C2 == nullptr || ok16(C2, ldc2, c2_dtype)) ? 1 : 0;

  cudaStream_t st = (cudaStream_t)stream;
  const bool elig = tcgen05_eligible(A, lda, W, ldw, ab_dtype, M, N, K);
  if (backend == AUM_GEMM_TCGEN05) {
    AUM_REQUIRE(elig, "aum_gemm_tn: tcgen05 backend needs fp16/bf16 operands, 16-byte aligned bases and row pitches");
    return launch_gemm_tcgen05(A, lda, W, ldw, ab_dtype, ep, M, N, K, st);
  }
  if (backend == AUM_GEMM_SIMT || !elig) return launch_gemm_simt(A, lda, W, ldw, ab_dtype, ep, M, N, K, st);
  return launch_gemm_tcgen05(A, lda, W, ldw, ab_dtype, ep, M, N, K, st);
}
