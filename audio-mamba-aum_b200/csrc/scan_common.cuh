// Shared definitions of the selective-scan kernels (generic kernel: scan_fwd.cu, TMA-streamed kernel: scan_fwd_tma.cu).
#pragma once
#include "common.cuh"

namespace aum {

constexpr int SCAN_NS = 16;   // states held in registers (d_state <= 16; padded with inert states)
constexpr int SCAN_TC = 64;   // tokens per staged B/C chunk
constexpr int SCAN_U = 4;     // tokens per prefetch batch
constexpr int SCAN_ROW = 2 * SCAN_NS;   // floats per staged token row: B[0..15] | C[0..15]

struct ScanDirDev {
  const void* u; int64_t ld_u;
  const void* delta; int64_t ld_delta;
  const float* A;
  const void* Bm; int64_t ld_B;
  const void* Cm; int64_t ld_C;
  int bc_dt;
  int bc_packed;   // fp32 rows of exactly [B(16) | C(16)], 16-byte aligned: eligible for bulk-async staging
  const float* D;
  const float* delta_bias;
  int delta_softplus;
  float* last_state;
  float* ckpt;     // optional: state before each checkpoint chunk, [batch][chunk][16][D]
  int reverse;     // 0: walks l = 0..L-1, 1: walks l = L-1..0
};

// Checkpoint chunking shared by the forward kernels (writers) and the backward kernel (reader): chunk 0 covers steps
// [0, first), chunk c >= 1 covers [first + 8(c-1), first + 8c); `first` makes the forward's phase switch S1 a chunk
// boundary (S1 = 0 for a unidirectional walk).
constexpr int SCAN_CK = 8;
__host__ __device__ inline int scan_S1(int L, bool bidir, bool reverse) { return bidir ? (reverse ? (L - L / 2) : (L / 2)) : 0; }
__host__ __device__ inline int scan_ck_first(int L, bool bidir, bool reverse) {
  const int S1 = scan_S1(L, bidir, reverse);
  return (S1 % SCAN_CK) ? (S1 % SCAN_CK) : SCAN_CK;
}
__host__ __device__ inline int scan_ck_count_max(int L) { return (L + SCAN_CK - 1) / SCAN_CK + 1; }

struct ScanParams {
  ScanDirDev dir[2];
  int ndirs;
  const void* z; int64_t ld_z;
  void* out; int64_t ld_out;
  void* ypre; int64_t ld_ypre;   // optional: pre-gate y_fwd + y_bwd (saved for the backward pass), activation dtype
  int batch, L, Dch, N;
  float out_scale;
  int z_pregated;                // z already holds silu(z)
};

// ---- small PTX helpers -------------------------------------------------------------------------------
// packed fp32x2 helpers (f32x2, pk2, upk2, fma2, mul2, add2): common.cuh

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (clock64() - t0 > (1ll << 32)) { printf("aum scan: mbarrier wait timed out\n"); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}


// TMA-streamed fast path (scan_fwd_tma.cu).  Returns -1 when the launch is not eligible (caller falls back to the
// generic kernel), 0 on success, >0 on error.
int launch_scan_tma(const ScanParams& p, int dtype, int delta_dt, cudaStream_t st);

}  // namespace aum
