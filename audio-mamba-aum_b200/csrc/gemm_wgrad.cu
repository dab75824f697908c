// Weight-gradient GEMM on tcgen05 for sm_100a:   dW[No, Ki] += dY[T, No]^T @ X[T, Ki]
// dY, X: fp16 / bf16 token-major (T token rows, channels contiguous); dW: fp32, accumulated into (the flat gradient
// buffer of the trainer, or a zeroed temporary).  These are the four parameter-gradient products of one mixer block,
// which the reference computes with einsum / matmul over transposed views
// (selective_scan_interface.py:563 d(out_proj.weight), :586 d(dt_proj.weight), :589 d(x_proj.weight), and autograd's
// d(in_proj.weight) of mamba_simple.py:185-189).
//
// The reduction runs over the TOKEN axis, so neither operand is K-contiguous: both are fed to the tensor core as
// MN-major operands (instruction-descriptor bits 15/16), straight from token-major HBM - no transposed copies.
//   * TMA boxes of [64 tokens x 64 channels] (128 B rows, 128-byte swizzle): in shared memory one box is the canonical
//     MN-major SWIZZLE_128B atom column (8-token groups 1024 B apart = SBO); the 64-channel boxes of a tile sit 8 KB
//     apart (= LBO).  A tile is 128 dY channels (2 boxes, the MMA's M) x BN X channels (BN / 64 boxes, its N).
//   * one k-block = 64 tokens = 4 tcgen05.mma (K = 16 tokens each; descriptor start address + 2048 B per step).
//   * split-K over tokens: the grid is (tile, token-range) work items, ordered so that co-resident CTAs read the same
//     token range (operands hit in L2); each item accumulates its partial with red.global.add.v4.f32.  Every fp32
//     addend is exact, only the order of the atomic adds varies from run to run.
//   * warp roles as in gemm_tcgen05.cu: warp 0 TMA producer, warp 1 MMA issuer (TMEM double-buffered), 8 epilogue warps.
// T / No / Ki tails: TMA zero-fills out-of-bounds box elements; the epilogue masks rows >= No and columns >= Ki.
#include <cuda.h>
#include <stdlib.h>

#include "gemm_common.cuh"
#include "tcgen05_ptx.cuh"
#include "tma.cuh"

namespace aum {

constexpr int WG_BM = 128;            // dY channels per tile (MMA M)
constexpr int WG_BT = 64;             // tokens per k-block
constexpr int WG_BOX = 64 * 64 * 2;   // one [64 tokens x 64 channels] 16-bit box: 8 KB
constexpr int WG_SMEM_BUDGET = 192 * 1024;

template <int BN> struct WgCfg {
  static constexpr int A_BYTES = (WG_BM / 64) * WG_BOX;       // 16 KB
  static constexpr int B_BYTES = (BN / 64) * WG_BOX;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = WG_SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int THREADS = 64 + 256;                    // producer, issuer, 8 epilogue warps
  static constexpr int ACC_STRIDE = BN;                       // 64 / 128 / 256 TMEM columns per buffer
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN: whole 64-channel boxes, power-of-two TMEM allocation");
};

// MN-major operand, 128-byte swizzle: 64-element (128 B) rows, 8-row (8-token) groups SBO = 1024 B apart,
// 64-channel atoms LBO = 8 KB apart (cute::UMMA canonical layout  Sw<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units).
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(WG_BOX >> 4) << 16;      // LBO = 8192 B
  d |= (uint64_t)(1024 >> 4) << 32;        // SBO = 1024 B
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WgParams {
  float* dW; int64_t ld_dw;
  int T, No, Ki;
  int m_tiles, n_tiles, nsplit, kb_per_split, k_blocks;
  int vec_ok;
};

template <int BN>
__global__ void __launch_bounds__(WgCfg<BN>::THREADS, 1)
gemm_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                          const WgParams p, uint32_t idesc) {
  using Cfg = WgCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.m_tiles * p.n_tiles;
  const int nitems = ntiles * p.nsplit;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
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

  // work item -> (token range, tile): items of one token range are consecutive, so CTAs resident together share operands
  auto item_of = [&](int it, int& m0, int& n0, int& kb0, int& kb1) {
    const int split = it / ntiles, tile = it - split * ntiles;
    m0 = (tile / p.n_tiles) * WG_BM;
    n0 = (tile % p.n_tiles) * BN;
    kb0 = split * p.kb_per_split;
    kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        int m0, n0, kb0, kb1; item_of(it, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < WG_BM / 64; ++j) tma_load_2d(sa + j * WG_BOX, &tmY, m0 + 64 * j, kb * WG_BT, full_bar(stage));
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * WG_BOX, &tmX, n0 + 64 * j, kb * WG_BT, full_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp converged, one elected lane issues) =================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
      int m0, n0, kb0, kb1; item_of(it, m0, n0, kb0, kb1);
      if (kb1 <= kb0) continue;                      // (empty token range: the epilogue skips it too)
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_STRIDE);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
        const uint64_t da = make_smem_desc_mn_sw128(sa);
        const uint64_t db = make_smem_desc_mn_sw128(sa + Cfg::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < WG_BT / 16; ++k) {
            // 16 tokens = two 8-token groups = 2048 B further down every box: +128 in 16-byte units
            tc_mma_f16(d_tmem, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));
          if (kb == kb1 - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> red.global.add into dW =================
    // 8 warps: warp & 3 = TMEM lane quarter (hardware rule), two sets take alternate 32-column chunks
    const int q = warp & 3;
    const int set = (warp - 2) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
      int m0, n0, kb0, kb1; item_of(it, m0, n0, kb0, kb1);
      if (kb1 <= kb0) continue;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t)(acc * Cfg::ACC_STRIDE) + ((uint32_t)(q * 32) << 16);
      const int row = m0 + q * 32 + lane;                 // dY channel = row of dW
      float* wrow = p.dW + (int64_t)row * p.ld_dw;
#pragma unroll 1
      for (int c0 = set * 32; c0 < BN; c0 += 64) {
        if (n0 + c0 >= p.Ki) break;                      // warp-uniform
        uint32_t r[32];
        tc_ld_32x32b_x32(t_row + (uint32_t)c0, r);
        tc_wait_ld();
        if (row < p.No) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int col = n0 + c0 + 4 * g;
            if (p.vec_ok && col + 4 <= p.Ki) {
              red_add_v4(wrow + col, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                         __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (col + i < p.Ki) atomicAdd(wrow + col + i, __uint_as_float(r[4 * g + i]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

template <int BN>
static int launch_wg(const CUtensorMap& tmY, const CUtensorMap& tmX, WgParams p, int ab_dt, cudaStream_t st) {
  using Cfg = WgCfg<BN>;
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("aum_gemm_wgrad: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return 2; }
    attr_set = true;
  }
  static PerDevice<int> sms_dev;
  int& sms = sms_dev.cur();
  if (sms == 0) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device()); if (sms <= 0) sms = 148; }
  p.m_tiles = ceil_div(p.No, WG_BM);
  p.n_tiles = ceil_div(p.Ki, BN);
  p.k_blocks = ceil_div(p.T, WG_BT);
  const int ntiles = p.m_tiles * p.n_tiles;
  // split-K: minimise  rounds x (k-blocks per item + epilogue), the epilogue of a tile costing about as much as
  // BN / 32 k-blocks of MMA (atomic adds of a 128 x BN fp32 tile)
  const int epi = BN / 32 + 2;
  int best_ns = 1; long best_cost = -1;
  for (int ns = 1; ns <= 64 && ns <= p.k_blocks; ++ns) {
    const int kpb = ceil_div(p.k_blocks, ns);
    const long rounds = ceil_div(ntiles * ns, sms);
    const long cost = rounds * (kpb + epi);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_ns = ns; }
  }
  p.kb_per_split = ceil_div(p.k_blocks, best_ns);
  p.nsplit = ceil_div(p.k_blocks, p.kb_per_split);          // no empty token ranges
  const int nitems = ntiles * p.nsplit;
  const int grid = nitems < sms ? nitems : sms;
  const int fmt = (ab_dt == AUM_F16) ? 0 : 1;
  const uint32_t idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10)
                       | (1u << 15) | (1u << 16)                    // A and B MN-major
                       | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(WG_BM >> 4) << 24);
  gemm_wgrad_tcgen05_kernel<BN><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmY, tmX, p, idesc);
  return check_launch("aum_gemm_wgrad(tcgen05)");
}

}  // namespace aum

extern "C" int aum_gemm_wgrad(const void* dY, int64_t ld_dy, const void* X, int64_t ld_x, int ab_dtype,
                              float* dW, int64_t ld_dw, int T, int No, int Ki, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(dW);
  if (T == 0 || No == 0 || Ki == 0) return 0;        // nothing to accumulate
  AUM_REQUIRE(dY && X && dW, "aum_gemm_wgrad: null pointer");
  AUM_REQUIRE(T > 0 && No > 0 && Ki > 0, "aum_gemm_wgrad: negative size");
  AUM_REQUIRE(ab_dtype == AUM_F16 || ab_dtype == AUM_BF16, "aum_gemm_wgrad: operands must be fp16 or bf16 (got dtype %d)", ab_dtype);
  AUM_REQUIRE(ld_dy >= No && ld_x >= Ki && ld_dw >= Ki, "aum_gemm_wgrad: leading dimension smaller than the row length");
  AUM_REQUIRE(aligned16(dY) && aligned16(X) && (ld_dy * 2) % 16 == 0 && (ld_x * 2) % 16 == 0,
              "aum_gemm_wgrad: operands need 16-byte aligned bases and row pitches (TMA)");
  AUM_REQUIRE((reinterpret_cast<uintptr_t>(dW) & 3) == 0, "aum_gemm_wgrad: dW must be 4-byte aligned");
  AUM_REQUIRE(tma_available(), "aum_gemm_wgrad: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmY, tmX;
  if (int rc = tma_encode_2d(&tmY, dY, ab_dtype, T, No, ld_dy, WG_BT, 64, true, "aum_gemm_wgrad(dY)")) return rc;
  if (int rc = tma_encode_2d(&tmX, X, ab_dtype, T, Ki, ld_x, WG_BT, 64, true, "aum_gemm_wgrad(X)")) return rc;
  WgParams p;
  p.dW = dW; p.ld_dw = ld_dw; p.T = T; p.No = No; p.Ki = Ki;
  p.vec_ok = (aligned16(dW) && (ld_dw % 4) == 0) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (Ki <= 64)  return launch_wg<64>(tmY, tmX, p, ab_dtype, st);
  if (Ki <= 128) return launch_wg<128>(tmY, tmX, p, ab_dtype, st);
  return launch_wg<256>(tmY, tmX, p, ab_dtype, st);
}
