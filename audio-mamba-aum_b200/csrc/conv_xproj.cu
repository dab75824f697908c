// Fused  x --causal conv1d + bias + SiLU--> u --x_proj--> (dt | B | C)   for sm_100a (16-bit activations).
//
// Reference ops replaced, one launch instead of two and one pass over x instead of a write + re-read of u:
//   conv1d_out = causal_conv1d_cuda.causal_conv1d_fwd(x, w, b, None, True)      selective_scan_interface.py:463 (:177,:318)
//   x_dbl      = F.linear(rearrange(conv1d_out, 'b d l -> (b l) d'), x_proj_weight)                          :467 (:181,:322)
// (BASELINE.json's north star asks for the depthwise conv + SiLU to be fused into the kernel that consumes it; the
// consumer that needs a full cross-channel reduction of u is x_proj, so the conv is the PRODUCER of x_proj's A operand.
// u is still written once, because the scan reads it.)
//
// One persistent CTA per SM walks 128-token tiles; per tile it loops over the Di channels in blocks of 64:
//   warp 0      TMA producer: the raw x rows of the block (128 tokens + 3 halo rows x 64 channels, un-swizzled) and the
//               x_proj weight block W_x[:, 64 channels] (K-major, 128-byte swizzle) into a 4-stage ring.
//   warps 2..17 conv: each thread takes 4 channels x 4 tokens: 7 LDS.64 of raw rows, 4 taps as packed fp32 FMAs, bias, SiLU
//               (ftz MUFU), and writes the 16-bit results into the stage's A tile in the 128-byte-swizzled K-major layout
//               the tensor core reads.  One elected thread then (a) publishes the tile to the MMA warp and (b) sends the
//               SAME tile to HBM as u with one bulk tensor store.
//   warp 1      tcgen05.mma 128 x NB x 16 (NB = 96 for AuM-Base's 80 outputs), accumulating over all channel blocks in
//               TMEM; tcgen05.commit frees the stage.
//   epilogue    (conv warps 2..5, lane = token row): TMEM -> registers -> dt (16-bit, first R columns) | [B|C] (fp32).
// 16 conv warps (4 per scheduler) at <= 113 registers: with 8 warps of 8-channel threads (166 registers) the conv ran at
// 1 IPC per SM, its shared-memory and MUFU latencies exposed (profiles/r2_ncu_conv_xproj_v2_summary.txt).
// Sequence boundaries: the halo rows of a token near the start (causal) or end (anti-causal, Bi-Bi's second branch) of
// its sequence belong to the neighbouring sequence or lie outside the tensor: those taps are masked.
#include <cuda.h>
#include <stdlib.h>

#include "gemm_common.cuh"
#include "scan_common.cuh"     // packed fp32x2 helpers, bulk_g2s
#include "tcgen05_ptx.cuh"
#include "tma.cuh"

namespace aum {

constexpr int CX_BM = 128;                 // tokens per tile
constexpr int CX_BK = 64;                  // channels per block (128 B of 16-bit)
constexpr int CX_HALO = 3;                 // d_conv - 1
constexpr int CX_RAW_ROWS = CX_BM + CX_HALO;
constexpr int CX_RAW_BYTES = 17 * 1024;    // 131 rows x 128 B = 16768, padded to keep 1024-byte alignment of what follows
constexpr int CX_CW_BYTES = 2 * 1024;      // conv taps of the block's 64 channels (64 x 4 fp32 = 1 KB) + their biases (256 B)
constexpr int CX_CW_BIAS_OFF = 1024;
constexpr int CX_A_BYTES = CX_BM * 128;    // 16 KB, 128-byte swizzled
constexpr int CX_STAGES = 6;                // load ring (raw x rows + W_x block + conv taps): 31 KB per stage at NB = 96.  The
                                           // kernel is latency-bound below ~5 blocks of prefetch (ncu v1/v2: one DRAM round trip per
                                           // channel block at 4 stages that also held the A tiles)
constexpr int CX_ASTAGES = 2;              // A tiles (conv output = MMA operand = source of the u store)
constexpr int CX_CONV_WARPS = 16;          // 4 per scheduler: the conv is latency-bound below that (ncu v1-v3 at 8 warps: 1 IPC/SM)
constexpr int CX_THREADS = 64 + 32 * CX_CONV_WARPS;
constexpr int CX_BAR_ID = 3;               // named barrier of the conv warps

template <int NB> struct CxCfg {
  static constexpr int W_BYTES = NB * 128;
  static constexpr int STAGE_BYTES = CX_RAW_BYTES + W_BYTES + CX_CW_BYTES;
  static constexpr int STAGES = (NB <= 96) ? CX_STAGES : 5;        // NB = 128: 35 KB per stage
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + CX_ASTAGES * CX_A_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static constexpr int TMEM_COLS = NB <= 32 ? 32 : NB <= 64 ? 64 : 128;
  static_assert(NB % 16 == 0 && NB >= 16 && NB <= 128, "UMMA N");
  static_assert(W_BYTES % 1024 == 0 || NB % 8 == 0, "");
};

struct CxParams {
  const float* cw; const float* cb;        // conv weight (Di, 4) fp32, bias (Di) fp32 or null
  void* dt; int64_t ld_dt; int dt_dt;      // (M, >= R) activation dtype
  float* bc; int64_t ld_bc;                // (M, 2N) fp32
  int M, L, Di, R, Nout;                   // Nout = R + 2N
  int reverse;
};

__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ float4 lds_f4x(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t v);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

template <typename T, int NB, bool REV>
__global__ void __launch_bounds__(CX_THREADS, 1)
conv_xproj_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmU, const CxParams p, uint32_t idesc) {
  using Cfg = CxCfg<NB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int STAGES = Cfg::STAGES;
  const uint32_t a_base = smem_base + STAGES * Cfg::STAGE_BYTES;                  // 1024-aligned
  const uint32_t bar_base = a_base + CX_ASTAGES * CX_A_BYTES;
  auto raw_full = [&](int s) { return bar_base + 8u * s; };                       // TMA: raw x rows + W block + taps landed
  auto st_free = [&](int s) { return bar_base + 8u * (STAGES + s); };             // MMA: load stage consumed
  auto a_full = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };          // conv: A tile written
  auto a_done = [&](int s) { return bar_base + 8u * (2 * STAGES + CX_ASTAGES + s); };   // MMA: A tile consumed
  auto a_free = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 * CX_ASTAGES + s); };   // u store has read the A tile
  const uint32_t tfull = bar_base + 8u * (2 * STAGES + 3 * CX_ASTAGES);
  const uint32_t tempty = tfull + 8u;
  const uint32_t tmem_slot = tfull + 16u;
  auto s_raw = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto s_w = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + CX_RAW_BYTES; };
  auto s_cw = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + CX_RAW_BYTES + Cfg::W_BYTES; };
  auto s_a = [&](int s) { return a_base + s * CX_A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + CX_BM - 1) / CX_BM;
  const int k_blocks = (p.Di + CX_BK - 1) / CX_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(raw_full(s), 1); mbar_init(st_free(s), 1); }
    for (int s = 0; s < CX_ASTAGES; ++s) { mbar_init(a_full(s), CX_CONV_WARPS); mbar_init(a_done(s), 1); mbar_init(a_free(s), 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 128);
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
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = tile * CX_BM;
        // causal: rows [m0 - 3, m0 + 128); anti-causal: rows [m0, m0 + 131).  Out-of-range rows are zero-filled.
        const int row0 = REV ? m0 : m0 - CX_HALO;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(st_free(stage), phase ^ 1u);
          const int nch = min(CX_BK, p.Di - kb * CX_BK);                       // channels of this block (Di % 8 == 0)
          const uint32_t cw_bytes = (uint32_t)nch * 16u, cb_bytes = p.cb ? (uint32_t)nch * 4u : 0u;
          mbar_arrive_expect_tx(raw_full(stage), CX_RAW_ROWS * 128 + Cfg::W_BYTES + cw_bytes + cb_bytes);
          tma_load_2d(s_raw(stage), &tmX, kb * CX_BK, row0, raw_full(stage));
          tma_load_2d(s_w(stage), &tmW, kb * CX_BK, 0, raw_full(stage));
          // the block's conv taps and biases ride on the same barrier (a per-thread __ldg of them cost 4 long-scoreboard
          // stalls per issue: every channel block brings new channels, i.e. L1 misses, right before they are needed)
          bulk_g2s(s_cw(stage), p.cw + (int64_t)kb * CX_BK * 4, cw_bytes, raw_full(stage));
          if (cb_bytes) bulk_g2s(s_cw(stage) + CX_CW_BIAS_OFF, p.cb + (int64_t)kb * CX_BK, cb_bytes, raw_full(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer + u store =================
    // (whole warp converged; the elected lane - the same one every time - issues the MMAs of a block and then sends the
    //  block's A tile to HBM as u: both read the tile through the async proxy once all 16 conv warps have published it)
    int stage = 0, as = 0; uint32_t aphase = 0, tphase = 0;
    int nblk = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = tile * CX_BM;
      mbar_wait(tempty, tphase ^ 1u);                 // epilogue of the previous tile has drained the accumulator
      tc_fence_after();
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(a_full(as), aphase);                // (the conv warps waited for this block's raw_full: W_x is there too)
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(s_a(as));
        const uint64_t db = make_smem_desc_sw128(s_w(stage));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < CX_BK / 16; ++k)
            tc_mma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(st_free(stage));
          tc_commit(a_done(as));
          if (kb == k_blocks - 1) tc_commit(tfull);
          tma_store_2d(&tmU, s_a(as), kb * CX_BK, m0);          // u[m0 .. m0+127, 64 channels] (clipped at M / Di)
          tma_store_commit();
          tma_store_wait_read<1>();                             // the PREVIOUS block's store has left its A buffer
          if (nblk >= 1) mbar_arrive(a_free(as ^ 1));
        }
        __syncwarp();
        ++nblk;
        if (++stage == STAGES) stage = 0;
        if (++as == CX_ASTAGES) { as = 0; aphase ^= 1u; }
      }
      tphase ^= 1u;
    }
    if (elect_one()) tma_store_wait_read<0>();        // shared memory must outlive the last bulk store
    __syncwarp();
  } else {
    // ================= conv warps (2..17) + epilogue (2..5) =================
    // thread = 4 channels (8 bytes of a row) x 4 tokens: 512 threads cover the 64-channel x 128-token block
    const int ct = threadIdx.x - 64;                  // 0..511
    const int cq = ct & 15;                           // 4-channel group inside the 64-channel block
    const int tr = ct >> 4;                           // 0..31: tokens 4 tr .. 4 tr + 3 of the tile
    int stage = 0, as = 0; uint32_t phase = 0, aphase = 0, tphase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = tile * CX_BM;
      // position of this thread's first token in its sequence, and the per-token edge masks (bit j of mask[i]: tap row
      // j of token i is valid).  Causal: tap row j of a token at position l holds x[l - 3 + j]; anti-causal: x[l + j].
      int l0 = (m0 + 4 * tr) % p.L;
      uint32_t mask4 = 0;                             // 4 bits per token
      bool edge = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int l = l0 + i; if (l >= p.L) l -= p.L;
        uint32_t mk = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = REV ? (l + j < p.L) : (l - CX_HALO + j >= 0);
          mk |= ok ? (1u << j) : 0u;
        }
        mask4 |= mk << (4 * i); edge |= (mk != 0xfu);
      }
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(raw_full(stage), phase);
        // taps and biases of this thread's 4 channels, from the stage's weight slab, as channel pairs:
        // wp[k][c2] = (w_k[2 c2], w_k[2 c2 + 1])
        f32x2 wp[4][2], bp[2];
        {
          const bool okc = kb * CX_BK + cq * 4 < p.Di;                 // (channel tail: Di % 8 == 0, so all 4 or none)
          const uint32_t wbase = s_cw(stage) + (uint32_t)cq * 64u;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
            if (okc) { wa = lds_f4x(wbase + (uint32_t)(2 * c2) * 16u); wb = lds_f4x(wbase + (uint32_t)(2 * c2 + 1) * 16u); }
            wp[0][c2] = pk2(wa.x, wb.x); wp[1][c2] = pk2(wa.y, wb.y); wp[2][c2] = pk2(wa.z, wb.z); wp[3][c2] = pk2(wa.w, wb.w);
          }
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (okc && p.cb != nullptr) b0 = lds_f4x(s_cw(stage) + CX_CW_BIAS_OFF + (uint32_t)cq * 16u);
          bp[0] = pk2(b0.x, b0.y); bp[1] = pk2(b0.z, b0.w);
        }
        // raw rows 4 tr .. 4 tr + 6 of the block, this thread's 8-byte channel group, each converted once to 2 fp32 pairs
        const uint32_t rbase = s_raw(stage) + (uint32_t)(4 * tr) * 128u + (uint32_t)cq * 8u;
        f32x2 xr[7][2];
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          uint32_t q0, q1;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(q0), "=r"(q1) : "r"(rbase + (uint32_t)r * 128u));
          const float2 f0 = unpack2<T>(q0), f1 = unpack2<T>(q1);
          xr[r][0] = pk2(f0.x, f0.y); xr[r][1] = pk2(f1.x, f1.y);
        }
        uint32_t outp[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f32x2 acc[2];
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) acc[c2] = bp[c2];
          // tap row j of token i is raw row i + j; it multiplies tap k = j (causal) or k = 3 - j (anti-causal)
          if (!edge) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int c2 = 0; c2 < 2; ++c2) acc[c2] = fma2(wp[REV ? 3 - j : j][c2], xr[i + j][c2], acc[c2]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float mj = ((mask4 >> (4 * i + j)) & 1u) ? 1.f : 0.f;
              const f32x2 m2 = pk2(mj, mj);
#pragma unroll
              for (int c2 = 0; c2 < 2; ++c2) acc[c2] = fma2(mul2(wp[REV ? 3 - j : j][c2], m2), xr[i + j][c2], acc[c2]);
            }
          }
          // SiLU on pairs: x * rcp(1 + ex2(-x log2 e)), flush-to-zero MUFU forms
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            float t0, t1, a0, a1;
            upk2(mul2(acc[c2], pk2(-1.4426950408889634f, -1.4426950408889634f)), t0, t1);
            upk2(acc[c2], a0, a1);
            float r0, r1;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.f + ex2_approx(t0)));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.f + ex2_approx(t1)));
            outp[i][c2] = pack2<T>(a0 * r0, a1 * r1);
          }
        }
        // the A buffer is free once the MMAs (a_done) and the u store (a_free) that read it two blocks ago are done.
        // No CTA-wide barrier anywhere in this loop: every warp publishes its own rows, so the 16 warps drift apart and
        // one warp's LDS / FMA phase overlaps another's MUFU / store phase.
        mbar_wait(a_done(as), aphase ^ 1u);
        mbar_wait(a_free(as), aphase ^ 1u);
        // A tile: row = token (128 B = 64 channels), 16-byte chunk index XOR (row & 7)  (128-byte swizzle)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t row = (uint32_t)(4 * tr + i);
          const uint32_t addr = s_a(as) + row * 128u + ((((uint32_t)(cq >> 1)) ^ (row & 7u)) << 4) + (((uint32_t)cq & 1u) << 3);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(outp[i][0]), "r"(outp[i][1]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(as));                         // -> MMA warp (16 arrivals complete the phase)
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (++as == CX_ASTAGES) { as = 0; aphase ^= 1u; }
      }
      // ---- epilogue of the tile: the first four conv warps, lane = token row (warp & 3 = TMEM lane quarter)
      if (warp < 6) {
        mbar_wait(tfull, tphase);
        tc_fence_after();
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        EpiParams ep;
        ep.C = p.dt; ep.ldc = p.ld_dt; ep.c_dt = p.dt_dt;
        ep.C2 = p.bc; ep.ldc2 = p.ld_bc; ep.c2_dt = AUM_F32; ep.split = p.R;
        ep.bias = nullptr; ep.row_scale = nullptr; ep.act = AUM_ACT_NONE; ep.act_col0 = 0;
        ep.M = p.M; ep.N = p.Nout; ep.vec_ok = 1;
#pragma unroll 1
        for (int c0 = 0; c0 < NB; c0 += 32) {
          if (c0 >= p.Nout) break;
          uint32_t r[32];
          tc_ld_32x32b_x32(t_row + (uint32_t)c0, r);
          tc_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
            epi_store8(ep, row, c0 + g * 8, v, 1.f);
          }
        }
        tc_fence_before();
        mbar_arrive(tempty);
      }
      tphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

template <typename T, int NB, bool REV>
static int launch_cx(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmU, const CxParams& p, int dt, cudaStream_t st) {
  using Cfg = CxCfg<NB>;
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_xproj_kernel<T, NB, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("aum_conv_xproj_fwd: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return 2; }
    attr_set = true;
  }
  static PerDevice<int> sms_dev;
  int& sms = sms_dev.cur();
  if (sms == 0) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device()); if (sms <= 0) sms = 148; }
  const int n_tiles = ceil_div(p.M, CX_BM);
  const int grid = n_tiles < sms ? n_tiles : sms;
  const int fmt = (dt == AUM_F16) ? 0 : 1;
  const uint32_t idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10)
                       | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(CX_BM >> 4) << 24);
  conv_xproj_kernel<T, NB, REV><<<grid, CX_THREADS, Cfg::SMEM_BYTES, st>>>(tmX, tmW, tmU, p, idesc);
  return check_launch("aum_conv_xproj_fwd");
}

}  // namespace aum

extern "C" int aum_conv_xproj_fwd(const void* x, int64_t ldx, const float* conv_w, const float* conv_b,
                                  const void* Wx, int64_t ldw, void* u, int64_t ldu,
                                  void* dt, int64_t ld_dt, float* bc, int64_t ld_bc,
                                  int batch, int L, int Di, int R, int N2, int dtype, int reverse, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(u);
  if (batch == 0 || L == 0 || Di == 0) return 0;
  AUM_REQUIRE(x && conv_w && Wx && u && dt && bc, "aum_conv_xproj_fwd: null pointer");
  AUM_REQUIRE(batch > 0 && L > 0 && Di > 0 && R >= 0 && N2 > 0, "aum_conv_xproj_fwd: bad sizes");
  AUM_REQUIRE(dtype == AUM_F16 || dtype == AUM_BF16, "aum_conv_xproj_fwd: activations must be fp16 or bf16 (got dtype %d)", dtype);
  const int Nout = R + N2;
  AUM_REQUIRE(Nout <= 128, "aum_conv_xproj_fwd: R + 2N = %d exceeds 128", Nout);
  AUM_REQUIRE(R % 8 == 0, "aum_conv_xproj_fwd: dt_rank must be a multiple of 8 (16-byte stores of the split output)");
  AUM_REQUIRE(ldx >= Di && ldu >= Di && ldw >= Di && ld_dt >= R && ld_bc >= N2, "aum_conv_xproj_fwd: leading dimension too small");
  auto ok16 = [](const void* p_, int64_t ld, int sz) { return aligned16(p_) && (ld * sz) % 16 == 0; };
  AUM_REQUIRE(ok16(x, ldx, 2) && ok16(u, ldu, 2) && ok16(Wx, ldw, 2) && ok16(dt, ld_dt, 2) && ok16(bc, ld_bc, 4) && aligned16(conv_w) &&
              (conv_b == nullptr || aligned16(conv_b)),
              "aum_conv_xproj_fwd: 16-byte aligned bases and row pitches required");
  AUM_REQUIRE(Di % 8 == 0, "aum_conv_xproj_fwd: d_inner must be a multiple of 8");
  AUM_REQUIRE(tma_available(), "aum_conv_xproj_fwd: cuTensorMapEncodeTiled unavailable");
  const int64_t M = (int64_t)batch * L;
  AUM_REQUIRE(M < (1ll << 31) - 256, "aum_conv_xproj_fwd: too many tokens");
  const int NB = Nout <= 32 ? 32 : Nout <= 64 ? 64 : Nout <= 96 ? 96 : 128;
  CUtensorMap tmX, tmW, tmU;
  if (int rc = tma_encode_2d(&tmX, x, dtype, M, Di, ldx, CX_RAW_ROWS, CX_BK, false, "aum_conv_xproj_fwd(x)")) return rc;
  if (int rc = tma_encode_2d(&tmW, Wx, dtype, Nout, Di, ldw, NB, CX_BK, true, "aum_conv_xproj_fwd(W_x)")) return rc;
  if (int rc = tma_encode_2d(&tmU, u, dtype, M, Di, ldu, CX_BM, CX_BK, true, "aum_conv_xproj_fwd(u)")) return rc;
  CxParams p;
  p.cw = conv_w; p.cb = conv_b; p.dt = dt; p.ld_dt = ld_dt; p.dt_dt = dtype; p.bc = bc; p.ld_bc = ld_bc;
  p.M = (int)M; p.L = L; p.Di = Di; p.R = R; p.Nout = Nout; p.reverse = reverse ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
#define AUM_CX2(T_, R_) \
  (NB == 32 ? launch_cx<T_, 32, R_>(tmX, tmW, tmU, p, dtype, st) : NB == 64 ? launch_cx<T_, 64, R_>(tmX, tmW, tmU, p, dtype, st) \
   : NB == 96 ? launch_cx<T_, 96, R_>(tmX, tmW, tmU, p, dtype, st) : launch_cx<T_, 128, R_>(tmX, tmW, tmU, p, dtype, st))
#define AUM_CX(T_) (reverse ? AUM_CX2(T_, true) : AUM_CX2(T_, false))
  return dtype == AUM_F16 ? AUM_CX(__half) : AUM_CX(__nv_bfloat16);
#undef AUM_CX
#undef AUM_CX2
}
