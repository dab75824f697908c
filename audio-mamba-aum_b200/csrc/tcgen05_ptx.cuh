// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (gemm_tcgen05.cu, gemm_wgrad.cu).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace aum {

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must never hang the GPU (it traps instead, after ~2 s).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > (1ll << 32)) { printf("aum gemm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// softplus for the dt_proj epilogue, branch-free: max(x, 0) + log1p(e), e = exp(-|x|) in (0, 1].  log1p(e) is
// ln(1 + e) above e = 0.01 and its series below it (where 1 + e would lose the low bits of e); flush-to-zero MUFU
// forms (the default ones carry a 3-instruction range fix-up each).  Relative error < 1e-5 over the whole range;
// x > 20 returns x to fp32 precision, which is torch's threshold-20 behaviour.
__device__ __forceinline__ float softplus_fast(float x) {
  const float e = ex2_approx(-1.4426950408889634f * fabsf(x));
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + e));
  const float series = e * (1.f - e * (0.5f - e * 0.33333334f));
  const float lp = e < 0.01f ? series : l * 0.6931471805599453f;
  return fmaxf(x, 0.f) + lp;
}
// x * sigmoid(x) on a packed pair: FMUL2, 2 x EX2, FADD2, 2 x RCP, FMUL2 (7 issue slots per two elements against 10)
__device__ __forceinline__ f32x2 silu_ftz2(f32x2 x) {
  float t0, t1, r0, r1;
  upk2(mul2(x, pk2(-1.4426950408889634f, -1.4426950408889634f)), t0, t1);
  upk2(add2(pk2(ex2_approx(t0), ex2_approx(t1)), pk2(1.f, 1.f)), t0, t1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
  return mul2(x, pk2(r0, r1));
}

// Two softplus evaluations with the arithmetic on packed fp32 pairs (FADD2 / FFMA2 / FMUL2: one issue slot per two
// elements; the dt_proj epilogue - 32 K evaluations per tile after ONE k-block of MMA - is issue-bound, not MUFU- or
// HBM-bound).  Same formula as softplus_fast: max(x, 0) + log1p(e), e = exp(-|x|); log1p through lg2(1 + e), or its
// series where 1 + e would lose e's low bits (e < 0.01).  11.5 issue slots per pair against ~26 for two scalar calls.
__device__ __forceinline__ f32x2 softplus_fast2(f32x2 x) {
  float x0, x1;
  upk2(x, x0, x1);
  const float e0 = ex2_approx(-1.4426950408889634f * fabsf(x0)), e1 = ex2_approx(-1.4426950408889634f * fabsf(x1));
  const f32x2 e = pk2(e0, e1);
  float w0, w1, l0, l1;
  upk2(add2(e, pk2(1.f, 1.f)), w0, w1);
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(w0));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(w1));
  // series e (1 - e (1/2 - e/3)) = e (1 + e (e/3 - 1/2))
  const f32x2 ser = mul2(e, fma2(e, fma2(e, pk2(0.33333334f, 0.33333334f), pk2(-0.5f, -0.5f)), pk2(1.f, 1.f)));
  float s0, s1;
  upk2(ser, s0, s1);
  const f32x2 relu = pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  const f32x2 lg = fma2(pk2(l0, l1), pk2(0.6931471805599453f, 0.6931471805599453f), relu);
  float g0, g1, r0, r1;
  upk2(lg, g0, g1);
  upk2(add2(ser, relu), r0, r1);
  (void)s0; (void)s1;
  return pk2(e0 < 0.01f ? r0 : g0, e1 < 0.01f ? r1 : g1);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi, int dt) {
  if (dt == AUM_F16) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// One lane of a CONVERGED warp.  The MMA issuer runs its loop with the whole warp converged and elects a lane only
// around the tcgen05 instructions: inside an `if (lane == 0)` region the compiler wraps every uniform-datapath
// instruction (UTCHMMA, UTCBAR) in its own ELECT / BRA.U.ANY loop - ~10 instructions per MMA on the one thread whose
// issue rate bounds the tensor pipe.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO = 1024 B
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

}  // namespace aum
