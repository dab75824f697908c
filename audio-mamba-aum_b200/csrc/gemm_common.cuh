// Shared GEMM epilogue:  C = act(row_scale[m] * acc + bias[n]), optionally split by column into two outputs
// of different dtype (x_proj: the dt slice stays 16-bit for the dt_proj tensor-core GEMM, the B/C slice is
// kept fp32 for the scan).
#pragma once
#include "common.cuh"

namespace aum {

struct EpiParams {
  void* C; int64_t ldc; int c_dt;
  void* C2; int64_t ldc2; int c2_dt; int split;
  const float* bias; const float* row_scale; int act; int act_col0;
  int M, N;
  int vec_ok;   // all of: ldc/ldc2/base pointers allow 16-byte stores of 8-column groups
};

__device__ __forceinline__ float epi_apply(const EpiParams& p, float acc, float rs, int col) {
  float v = acc * rs;
  if (p.bias != nullptr) v += __ldg(p.bias + col);
  if (p.act == AUM_ACT_SOFTPLUS) { if (col >= p.act_col0) v = softplus_f(v); }
  else if (p.act == AUM_ACT_SILU) { if (col >= p.act_col0) v = silu_f(v); }
  return v;
}

__device__ __forceinline__ void epi_store1(const EpiParams& p, int row, int col, float acc) {
  if (row >= p.M || col >= p.N) return;
  const float rs = p.row_scale ? __ldg(p.row_scale + row) : 1.f;
  const float v = epi_apply(p, acc, rs, col);
  if (col < p.split) store_from_f(p.C, (int64_t)row * p.ldc + col, p.c_dt, v);
  else store_from_f(p.C2, (int64_t)row * p.ldc2 + (col - p.split), p.c2_dt, v);
}

// 8 consecutive columns starting at col (col % 8 == 0) of one row.
__device__ __forceinline__ void epi_store8(const EpiParams& p, int row, int col, const float (&acc)[8], float rs) {
  if (row >= p.M || col >= p.N) return;
  // vector path only when the 8-column group lies entirely in one output and lands 16-byte aligned there
  const bool in_c = col + 8 <= p.split;
  const bool in_c2 = col >= p.split && ((col - p.split) & 7) == 0;
  if (p.vec_ok && col + 8 <= p.N && (in_c || in_c2)) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = epi_apply(p, acc[i], rs, col + i);
    void* base; int64_t off; int dt;
    if (in_c) { base = p.C;  off = (int64_t)row * p.ldc + col;               dt = p.c_dt;  }
    else               { base = p.C2; off = (int64_t)row * p.ldc2 + (col - p.split);  dt = p.c2_dt; }
    if (dt == AUM_F32)       { Vec8<float> t;         t.pack(v); t.store(reinterpret_cast<float*>(base) + off); }
    else if (dt == AUM_F16)  { Vec8<__half> t;        t.pack(v); t.store(reinterpret_cast<__half*>(base) + off); }
    else                     { Vec8<__nv_bfloat16> t; t.pack(v); t.store(reinterpret_cast<__nv_bfloat16*>(base) + off); }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = col + i;
      if (c < p.N) {
        const float v = epi_apply(p, acc[i], rs, c);
        if (c < p.split) store_from_f(p.C, (int64_t)row * p.ldc + c, p.c_dt, v);
        else store_from_f(p.C2, (int64_t)row * p.ldc2 + (c - p.split), p.c2_dt, v);
      }
    }
  }
}

int launch_gemm_simt(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dt, const EpiParams& ep,
                     int M, int N, int K, cudaStream_t st);
int launch_gemm_tcgen05(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dt, const EpiParams& ep,
                        int M, int N, int K, cudaStream_t st);
bool tcgen05_eligible(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dt, int M, int N, int K);

}  // namespace aum
