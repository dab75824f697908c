// Tensor-core GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] @ W[N,K]^T), A/W fp16 or bf16 (K-contiguous),
// fp32 accumulation in TMEM.  This is the in_proj / x_proj / dt_proj / out_proj engine of the hot path
// (reference call sites: mamba_simple.py:185-189, selective_scan_interface.py:467,468,517).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (128B-swizzled) of A [128 x 64] and W [BN x 64]
//               into a STAGES-deep shared-memory ring, completion on "full" mbarriers.
//   warp 1      allocates TMEM, issues tcgen05.mma (128 x BN x 16, cta_group::1, kind::f16) from one thread;
//               tcgen05.commit releases smem stages ("empty") and publishes accumulators ("tmem_full").
//   warps 2..9  epilogue (two sets of 4 warps): tcgen05.ld 32x32b (lane == output row) -> registers ->
//               scale/bias/activation -> 128B-swizzled smem slab -> cp.async.bulk.tensor store (TMA).
//               TMEM holds two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1.
// M/N/K tails: TMA zero-fills out-of-bounds box elements (K tail adds zeros, M/N tails are masked at store).
#include <cuda.h>   // CUtensorMap types only; the encode entry point is fetched at run time
#include <stdlib.h>

#include "gemm_common.cuh"
#include "tcgen05_ptx.cuh"

namespace aum {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;              // 64 x 2 B = 128 B = one swizzle row
constexpr int TC_UMMA_K = 16;
constexpr int TC_SMEM_BUDGET = 192 * 1024;     // operand ring
constexpr int TC_CSTAGE_BYTES = 128 * 128;      // one epilogue staging buffer: 128 rows x 128 B (swizzled)
constexpr int TC_EPI_BAR = 1;                   // named barriers 1,2: the two epilogue warp-sets

// NSETS: epilogue warp-sets of 4 warps (one warp per TMEM lane quarter); each set owns a staging buffer, a named
// barrier and a store thread and takes every NSETS-th column slab of a tile.  2 sets for the MMA-bound projections;
// 4 sets (16 epilogue warps, four per scheduler) for dt_proj, whose tile is ONE k-block of MMA followed by 32 K
// bias + softplus evaluations: with 8 epilogue warps that epilogue ran at 0.25 IPC per scheduler and took 3x the
// time of the 202 MB delta write it feeds.
template <int BN, int NSETS = 2> struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;          // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int THREADS = 64 + 128 * NSETS;           // 2 control warps + 4 NSETS epilogue warps
  static constexpr int STAGES_RAW = (TC_SMEM_BUDGET - (NSETS - 2) * TC_CSTAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_STRIDE = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;   // TMEM columns / buffer
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NSETS * TC_CSTAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(B_BYTES % 1024 == 0, "W tile must keep 1024-byte stage alignment");
};

__device__ __forceinline__ void epi_barrier(int set) { asm volatile("bar.sync %0, %1;" ::"r"(TC_EPI_BAR + set), "n"(128) : "memory"); }

// Epilogue of one 128 x BN accumulator tile (TMEM lanes = rows), run by the NSETS warp-sets of a CTA: TMEM -> registers
// -> (row scale, bias, activation) -> convert -> 128B-swizzled staging slab -> TMA store, or per-thread vector stores
// (split outputs, odd pitches).  t_acc: TMEM address of the tile's accumulator (lane 0, first column).
template <int BN, int NSETS>
__device__ __forceinline__ void tc_epilogue_tile(const EpiParams& ep, const CUtensorMap* tmC, int tma_store,
                                                 uint32_t t_acc, int m0, int n0, int M, int N,
                                                 int q, int set, int lane, bool store_thread, uint32_t cstage_base) {
  const bool has_bias = ep.bias != nullptr, has_rs = ep.row_scale != nullptr;
  const int act = ep.act;
  const bool plain = !has_bias && !has_rs && act == AUM_ACT_NONE;
  const int row_in_tile = q * 32 + lane;
  const int row = m0 + row_in_tile;
  const float rs = (has_rs && row < M) ? __ldg(ep.row_scale + row) : 1.f;
  const uint32_t t_row = t_acc + ((uint32_t)(q * 32) << 16);
  if (tma_store) {
    // TMEM -> registers -> (scale, bias, activation) -> convert -> 128B-swizzled smem slab -> TMA store
    const int c_sz = (ep.c_dt == AUM_F32) ? 4 : 2;
    const int slab_cols = 128 / c_sz;            // 64 16-bit or 32 fp32 columns
    const uint32_t buf = cstage_base + (uint32_t)set * TC_CSTAGE_BYTES;
    const uint32_t srow = buf + (uint32_t)row_in_tile * 128u;
    const uint32_t sw = (uint32_t)(row_in_tile & 7);
#pragma unroll 1
    for (int c0 = set * slab_cols; c0 < BN; c0 += NSETS * slab_cols) {
      if (n0 + c0 >= N) break;                   // warp-uniform
      if (store_thread) tma_store_wait_read<0>();   // this set's previous store has drained its buffer
      epi_barrier(set);
#pragma unroll 1
      for (int cc = 0; cc < slab_cols; cc += 32) {
        uint32_t r[32];
        tc_ld_32x32b_x32(t_row + (uint32_t)(c0 + cc), r);
        tc_wait_ld();
        if (!plain) {
          // Every decision below is warp-uniform and hoisted out of the 32-element loops, so that each loop is
          // straight-line code whose MUFU chains the compiler can interleave (a per-element predicate made this
          // path latency-bound: 3x the MMA time per tile).
          const int colb = n0 + c0 + cc;
          if (has_rs) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * rs);
          }
          if (has_bias) {
            if (colb + 32 <= N) {
              const float4* bp = reinterpret_cast<const float4*>(ep.bias + colb);
              const bool al = (reinterpret_cast<uintptr_t>(bp) & 15) == 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 b4;
                if (al) b4 = __ldg(bp + i);
                else b4 = make_float4(__ldg(ep.bias + colb + 4 * i), __ldg(ep.bias + colb + 4 * i + 1),
                                      __ldg(ep.bias + colb + 4 * i + 2), __ldg(ep.bias + colb + 4 * i + 3));
                r[4 * i]     = __float_as_uint(__uint_as_float(r[4 * i]) + b4.x);
                r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + b4.y);
                r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + b4.z);
                r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + b4.w);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (colb + i < N) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldg(ep.bias + colb + i));
            }
          }
          if (act != AUM_ACT_NONE && colb + 32 > ep.act_col0) {
            if (colb >= ep.act_col0) {              // whole chunk inside the activated column range
              if (act == AUM_ACT_SILU) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  float v0, v1;
                  upk2(silu_ftz2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]))), v0, v1);
                  r[i] = __float_as_uint(v0); r[i + 1] = __float_as_uint(v1);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  float v0, v1;
                  upk2(softplus_fast2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]))), v0, v1);
                  r[i] = __float_as_uint(v0); r[i + 1] = __float_as_uint(v1);
                }
              }
            } else {                                // chunk straddles act_col0 (not 32-aligned): per element
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float v = __uint_as_float(r[i]);
                if (colb + i >= ep.act_col0) v = (act == AUM_ACT_SILU) ? silu_ftz(v) : softplus_fast(v);
                r[i] = __float_as_uint(v);
              }
            }
          }
        }
        if (c_sz == 4) {
#pragma unroll
          for (int k = 0; k < 8; ++k)            // 8 chunks of 4 floats (slab_cols == 32: cc == 0)
            st_shared_v4(srow + ((((uint32_t)k) ^ sw) << 4), r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        } else {
          const bool f16 = ep.c_dt == AUM_F16;
#pragma unroll
          for (int k = 0; k < 4; ++k) {          // 4 chunks of 8 halves per 32 columns
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float lo = __uint_as_float(r[8 * k + 2 * j]), hi = __uint_as_float(r[8 * k + 2 * j + 1]);
              if (f16) { __half2 h = __floats2half2_rn(lo, hi); pk[j] = *reinterpret_cast<uint32_t*>(&h); }
              else { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); pk[j] = *reinterpret_cast<uint32_t*>(&h); }
            }
            st_shared_v4(srow + ((((uint32_t)((cc >> 3) + k)) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      epi_barrier(set);
      if (store_thread) { tma_store_2d(tmC, buf, n0 + c0, m0); tma_store_commit(); }
    }
  } else {
    // per-thread vector stores (split outputs, odd pitches): each set takes alternate 32-column chunks
#pragma unroll 1
    for (int c0 = set * 32; c0 < BN; c0 += 32 * NSETS) {
      if (n0 + c0 >= N) break;                     // warp-uniform
      uint32_t r[32];
      tc_ld_32x32b_x32(t_row + (uint32_t)c0, r);
      tc_wait_ld();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
        epi_store8(ep, row, n0 + c0 + g * 8, v, rs);
      }
    }
  }
}

template <int BN, int NSETS>
__global__ void __launch_bounds__(TcCfg<BN, NSETS>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmC, int tma_store,
                    EpiParams ep, int M, int N, int K, uint32_t idesc) {
  using Cfg = TcCfg<BN, NSETS>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128-byte swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t cstage_base = smem_base + STAGES * Cfg::STAGE_BYTES;     // NSETS x 16 KB, 1024-aligned
  const uint32_t bar_base = cstage_base + NSETS * TC_CSTAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);     // 4-byte slot for the TMEM base address

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + TC_BM - 1) / TC_BM, n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128 * NSETS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, kb * TC_BK, m0, full_bar(stage));
          tma_load_2d(sb, &tmW, kb * TC_BK, n0, full_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp converged, one elected lane issues) =================
    {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_STRIDE);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
              // advance 16 elements = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
              tc_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit(empty_bar(stage));                    // smem stage reusable once these MMAs retire
            if (kb == k_blocks - 1) tc_commit(tfull_bar(acc));   // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================= epilogue warps (2..): NSETS independent sets of 4 warps =================
    // Each set covers all 128 accumulator rows (one warp per TMEM lane quarter) and takes every other
    // 128-byte-wide column slab of the tile; it owns one staging buffer, one named barrier and one store thread.
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int set = (warp - 2) >> 2;                // 0 or 1
    const bool store_thread = (lane == 0) && (((warp - 2) & 3) == 0);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      tc_epilogue_tile<BN, NSETS>(ep, &tmC, tma_store, tmem_base + (uint32_t)(acc * Cfg::ACC_STRIDE), m0, n0, M, N,
                                  q, set, lane, store_thread, cstage_base);
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (tma_store && store_thread) tma_store_wait_read<0>();   // smem must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// =====================================================================================================
// CTA-pair variant (cta_group::2) for the two MMA-bound projections (in_proj, out_proj).
// The single-CTA kernel above is bound by the shared-memory operand reads of its 128 x 256 x 16 MMAs (narrower tiles
// lose proportionally more); a pair of CTAs on the two SMs of a TPC computes a 256 x 256 tile with
// each SM holding 128 rows of A and 128 of the 256 rows of W, so every SM reads half as many operand bytes per FLOP.
//   * cluster (2,1,1); rank r loads A rows [m0 + 128 r, +128) and W rows [n0 + 128 r, +128) of every k-block into its
//     own ring (cp.async.bulk.tensor ... cta_group::2: the transaction bytes of both CTAs complete on the LEADER's
//     "full" barrier, whose expect_tx the leader posts for the pair);
//   * only the leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16); its tcgen05.commit is
//     multicast to both CTAs' "empty" (ring stage free) and "tmem_full" (accumulator ready) barriers;
//   * each CTA runs the usual epilogue on its own 128 TMEM lanes = its 128 rows; one thread per epilogue warp of both
//     CTAs arrives on the leader's "tmem_empty" barrier (remote mbarrier.arrive for the peer).
// Barrier layout and shared-memory offsets are identical in both CTAs: a shared::cta address with bit 24 cleared is the
// same location in the leader (cute::Sm100MmaPeerBitMask).
// =====================================================================================================
constexpr uint32_t TC_PEER_MASK = 0xFEFFFFFFu;
constexpr int TC2_BN = 256;                       // output tile 256 (pair) x 256
constexpr int TC2_BHALF = TC2_BN / 2;             // W rows per CTA

template <int NSETS_> struct Tc2Cfg {
  static constexpr int NSETS = NSETS_;
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;             // 16 KB
  static constexpr int B_BYTES = TC2_BHALF * TC_BK * 2;         // 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (TC_SMEM_BUDGET - (NSETS - 2) * TC_CSTAGE_BYTES) / STAGE_BYTES;   // 6 / 5
  static constexpr int THREADS = 64 + 128 * NSETS;
  static constexpr int ACC_STRIDE = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NSETS * TC_CSTAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {     // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // from either CTA: the leader's copy of `bar`
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & TC_PEER_MASK) : "memory");
}

template <int NSETS_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Cfg<NSETS_>::THREADS, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                         const __grid_constant__ CUtensorMap tmC, int tma_store,
                         EpiParams ep, int M, int N, int K, uint32_t idesc) {
  using Cfg = Tc2Cfg<NSETS_>;
  constexpr int STAGES = Cfg::STAGES, NSETS = Cfg::NSETS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t cstage_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = cstage_base + NSETS * TC_CSTAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_tiles = (M + 2 * TC_BM - 1) / (2 * TC_BM), n_tiles = (N + TC2_BN - 1) / TC2_BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * 4 * NSETS); }   // one arrival per epilogue warp of the pair
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers initialised and TMEM allocated before anyone signals across the pair
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer (both CTAs: own halves of A and W) =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        const int m0 = (tile / n_tiles) * (2 * TC_BM) + (int)rank * TC_BM;
        const int n0 = (tile % n_tiles) * TC2_BN + (int)rank * TC2_BHALF;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint32_t lbar = full_bar(stage) & TC_PEER_MASK;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          tma_load_2d_pair(sa, &tmA, kb * TC_BK, m0, lbar);
          tma_load_2d_pair(sb, &tmW, kb * TC_BK, n0, lbar);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only; whole warp converged, one elected lane issues) =================
    if (leader) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_STRIDE);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + Cfg::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
              tc_mma_f16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit_pair(empty_bar(stage));
            if (kb == k_blocks - 1) tc_commit_pair(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (both CTAs: own 128 rows) =================
    const int q = warp & 3;
    const int set = (warp - 2) >> 2;
    const bool store_thread = (lane == 0) && (((warp - 2) & 3) == 0);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs) {
      const int m0 = (tile / n_tiles) * (2 * TC_BM) + (int)rank * TC_BM, n0 = (tile % n_tiles) * TC2_BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      tc_epilogue_tile<TC2_BN, NSETS>(ep, &tmC, tma_store, tmem_base + (uint32_t)(acc * Cfg::ACC_STRIDE), m0, n0, M, N,
                                      q, set, lane, store_thread, cstage_base);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (tma_store && store_thread) tma_store_wait_read<0>();
  }

  tc_fence_before();
  cluster_sync_all();          // the leader's MMAs read the peer's shared memory and both TMEMs: leave together
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 2-D tensor map of the output [M, N] (row pitch ldc elements) for the TMA-store epilogue: box = 128 rows x 128 B.
static int make_tmap_out(CUtensorMap* tm, const void* base, int64_t M, int64_t N, int64_t ldc, int dt) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("aum_gemm_tn: cuTensorMapEncodeTiled unavailable (driver too old?)"); return 3; }
  const int sz = dtype_size(dt);
  cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t gstr[1] = {(cuuint64_t)ldc * sz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / sz), (cuuint32_t)TC_BM};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType t = dt == AUM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                        : dt == AUM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(tm, t, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("aum_gemm_tn: cuTensorMapEncodeTiled(out) failed (%d) M=%lld N=%lld ldc=%lld", (int)r, (long long)M, (long long)N, (long long)ldc); return 3; }
  return 0;
}

// 2-D tensor map over a K-contiguous [rows, K] matrix with row pitch ld (elements); box = [box_rows x 64].
static int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows, int dt) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("aum_gemm_tn: cuTensorMapEncodeTiled unavailable (driver too old?)"); return 3; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt == AUM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("aum_gemm_tn: cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r, (long long)rows, (long long)K, (long long)ld); return 3; }
  return 0;
}

bool tcgen05_eligible(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dt, int M, int N, int K) {
  if (ab_dt != AUM_F16 && ab_dt != AUM_BF16) return false;
  if (!aligned16(A) || !aligned16(W)) return false;
  if ((lda * 2) % 16 != 0 || (ldw * 2) % 16 != 0) return false;   // TMA global strides: multiples of 16 bytes
  if (M < 1 || N < 1 || K < 1) return false;
  return true;
}

static PerDevice<int> g_sm_count_dev;

template <int BN, int NSETS = 2>
static int launch_bn(const CUtensorMap& tmA, const void* W, int64_t ldw, int ab_dt, const EpiParams& ep,
                     int M, int N, int K, cudaStream_t st) {
  using Cfg = TcCfg<BN, NSETS>;
  CUtensorMap tmW, tmC;
  if (int rc = make_tmap(&tmW, W, N, K, ldw, BN, ab_dt)) return rc;
  // TMA-store epilogue when there is a single, 16-byte-pitched output; otherwise per-thread vector stores
  static const bool direct_store = getenv("AUM_GEMM_DIRECT_STORE") != nullptr;    // environment read once
  int tma_store = (ep.C2 == nullptr && ep.vec_ok && BN >= 64 && !direct_store) ? 1 : 0;
  if (tma_store) { if (int rc = make_tmap_out(&tmC, ep.C, M, N, ep.ldc, ep.c_dt)) return rc; }
  else tmC = tmW;
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, NSETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("aum_gemm_tn: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return 2; }
    attr_set = true;
  }
  const int fmt = (ab_dt == AUM_F16) ? 0 : 1;   // cute::UMMA::F16F32Format
  const uint32_t idesc = (1u << 4)              // D format: F32
                       | ((uint32_t)fmt << 7)   // A format
                       | ((uint32_t)fmt << 10)  // B format
                       | (0u << 15) | (0u << 16)          // A, B K-major
                       | ((uint32_t)(BN >> 3) << 17)      // N >> 3
                       | ((uint32_t)(TC_BM >> 4) << 24);  // M >> 4
  const int tiles = ceil_div(M, TC_BM) * ceil_div(N, BN);
  const int g_sm_count = g_sm_count_dev.cur();
  const int grid = tiles < g_sm_count ? tiles : g_sm_count;
  gemm_tcgen05_kernel<BN, NSETS><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW, tmC, tma_store, ep, M, N, K, idesc);
  return check_launch("aum_gemm_tn(tcgen05)");
}

template <int NSETS>
static int launch_pair(const CUtensorMap& tmA, const void* W, int64_t ldw, int ab_dt, const EpiParams& ep,
                       int M, int N, int K, cudaStream_t st) {
  using Cfg = Tc2Cfg<NSETS>;
  CUtensorMap tmW, tmC;
  if (int rc = make_tmap(&tmW, W, N, K, ldw, TC2_BHALF, ab_dt)) return rc;
  if (int rc = make_tmap_out(&tmC, ep.C, M, N, ep.ldc, ep.c_dt)) return rc;
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<NSETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("aum_gemm_tn: cudaFuncSetAttribute(pair, smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return 2; }
    attr_set = true;
  }
  const int fmt = (ab_dt == AUM_F16) ? 0 : 1;
  const uint32_t idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10)
                       | ((uint32_t)(TC2_BN >> 3) << 17)          // N >> 3
                       | ((uint32_t)((2 * TC_BM) >> 4) << 24);    // M >> 4 (256: the pair)
  const int tiles = ceil_div(M, 2 * TC_BM) * ceil_div(N, TC2_BN);
  int pairs = g_sm_count_dev.cur() / 2;
  if (tiles < pairs) pairs = tiles;
  gemm_tcgen05_pair_kernel<NSETS><<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW, tmC, 1, ep, M, N, K, idesc);
  return check_launch("aum_gemm_tn(tcgen05 pair)");
}

int launch_gemm_tcgen05(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dt, const EpiParams& ep,
                        int M, int N, int K, cudaStream_t st) {
  int& g_sm_count = g_sm_count_dev.cur();
  if (g_sm_count == 0) {
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, current_device());
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  CUtensorMap tmA;
  if (int rc = make_tmap(&tmA, A, M, K, lda, TC_BM, ab_dt)) return rc;
  // MMA-bound shapes with a single TMA-storable output: the CTA-pair kernel (AUM_GEMM_PAIR=0 turns it off)
  static int use_pair = -1;
  if (use_pair < 0) { const char* e = getenv("AUM_GEMM_PAIR"); use_pair = (e && atoi(e) == 0) ? 0 : 1; }
  // (with an activation epilogue - in_proj's SiLU(z) - the pair kernel measured no faster than the single-CTA one:
  //  0.125 vs 0.126 ms, against 0.116 vs 0.124 ms for the bare GEMM; 16 epilogue warps did not change that)
  if (use_pair && N >= 512 && K >= 256 && M >= 1024 && ep.C2 == nullptr && ep.vec_ok && ep.act == AUM_ACT_NONE)
    return launch_pair<2>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
  // Tile-N choice: the widest tile that does not waste more than ~12 % of the MMA on column padding.
  if (N <= 32)  return launch_bn<32>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
  if (N <= 64)  return launch_bn<64>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
  if (N <= 96)  return launch_bn<96>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
  if (N <= 128) return launch_bn<128>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
  // one k-block of MMA per tile and a wide output: the epilogue is the kernel (dt_proj) -> 16 epilogue warps
  static const bool epi2 = getenv("AUM_GEMM_EPI2") != nullptr;
  if (K <= TC_BK && N >= 512 && !epi2) return launch_bn<256, 4>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
  if (N % 256 == 0 || N > 1024) return launch_bn<256>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
  return launch_bn<128>(tmA, W, ldw, ab_dt, ep, M, N, K, st);
}

}  // namespace aum
