// Selective scan backward — TMA-streamed kernel (the production path; math and reference citations: scan_bwd.cu).
//
// What changes against the generic kernel is how operands move and how the two directions are combined:
//   * grid = (ceil(D/128), batch, directions): a CTA owns 128 channels of one sequence in ONE time direction (one
//     thread per channel, the 16 states as 8 packed fp32x2 pairs).  The directions no longer meet inside a CTA:
//     where du / ddelta are shared (Fo-Bi) both directions red.global.add into buffers the entry point zeroes
//     first (two commutative fp32 additions per element: deterministic); dz / out_z are written by direction 0.
//   * per 8-step checkpoint chunk one elected thread issues bulk tensor copies into a 2-stage ring: the
//     checkpoint tile (16 x 128 fp32, state before the chunk, left by the forward kernel), delta, u, dout, z and
//     (direction 0) y_pre tiles (8 x 128) and the packed [B|C] rows; all complete on the stage's mbarrier.  The
//     loads of chunk c-1 are in flight while chunk c is processed, so no thread ever waits on a global load.
//   * the chunk's state history (8 slots x 16 states per channel) lives in TENSOR MEMORY, which this kernel has no
//     other use for: thread = TMEM lane (the CTA's four warps are the four lane quarters), slot j = columns
//     [16 j, 16 j + 16) of the CTA's 128-column allocation; one tcgen05.st (32x32b.x16) per replayed step, one
//     tcgen05.ld per reverse-time step.  In shared memory the history cost 512 B per thread and bounded residency at
//     8 warps per SM (the kernel is issue-bound: 446 instructions per warp-step at 48 % issue utilisation with 7
//     resident warps); with only the 2 x 21 KB stages left the register file is the limit: 3 CTAs = 12 warps.
// Eligibility (launch_scan_bwd_tma): d_state == 16, forward checkpoints present, packed fp32 [B|C] rows,
// 16-byte aligned bases and row pitches.  Anything else runs scan_bwd.cu.
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

#include "scan_bwd_common.cuh"
#include "tma.cuh"

namespace aum {

constexpr int BT_CH = 128;    // channels per CTA = TMEM lanes

constexpr int BT_TT = 8;      // steps per checkpoint chunk (== SCAN_CK)
static_assert(BT_TT == SCAN_CK, "chunk length must match the forward kernels' checkpoint interval");
constexpr int BT_TMEM_COLS = BT_TT * SCAN_NS;   // 128 columns: 8 history slots x 16 states (power of two >= 32)

// TD: element type of delta (float or T)
template <typename T, typename TD = float> struct BwdLayout {
  static constexpr int CK_BYTES = SCAN_NS * BT_CH * 4;             // checkpoint tile [n][ch]
  static constexpr int D_BYTES = BT_TT * BT_CH * (int)sizeof(TD);  // delta tile
  static constexpr int BC_BYTES = BT_TT * SCAN_ROW * 4;            // [B|C] rows
  static constexpr int A_BYTES = BT_TT * BT_CH * (int)sizeof(T);   // u, dout, z, y_pre tiles
  static constexpr int OFF_CK = 0, OFF_D = OFF_CK + CK_BYTES, OFF_BC = OFF_D + D_BYTES, OFF_U = OFF_BC + BC_BYTES;
  static constexpr int OFF_G = OFF_U + A_BYTES, OFF_Z = OFF_G + A_BYTES, OFF_Y = OFF_Z + A_BYTES;
  static constexpr int STAGE_BYTES = OFF_Y + A_BYTES;
  static constexpr int RED_PITCH = 34;                             // floats per lane row of the dB|dC transpose tile (see below)
  static constexpr int RED_BYTES = (BT_CH / 32) * 32 * RED_PITCH * 4;   // one 32 x 32 (+pad) tile per warp: 17 KB
  static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + RED_BYTES + 128 /*align slack*/ + 64 /*mbarriers, TMEM base slot*/;
};

struct ScanBwdMaps { CUtensorMap u[2], d[2], ck[2], g, z, y; };

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(col), "r"(row), "r"(bar) : "memory");
}
__device__ __forceinline__ float ldsf(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 ldsf4(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
// state history in tensor memory: 16 consecutive columns of this thread's lane = the 16 states of one slot
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const f32x2 (&h)[SCAN_NS / 2]) {
  uint32_t r[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) { float lo, hi; upk2(h[k], lo, hi); r[2 * k] = __float_as_uint(lo); r[2 * k + 1] = __float_as_uint(hi); }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, f32x2 (&h)[SCAN_NS / 2]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 8; ++k) h[k] = pk2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
}
template <typename T> __device__ __forceinline__ float ldst(uint32_t a);
template <> __device__ __forceinline__ float ldst<float>(uint32_t a) { return ldsf(a); }
template <> __device__ __forceinline__ float ldst<__half>(uint32_t a) {
  unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return __half2float(__ushort_as_half(v));
}
template <> __device__ __forceinline__ float ldst<__nv_bfloat16>(uint32_t a) {
  unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return __uint_as_float(((uint32_t)v) << 16);
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Replay one forward step: h <- exp(dl A) h + dl u B; the new state goes to history slot `slot` (a TMEM address) unless
// it is the chunk's last step, whose result stays in registers (it is the first h_s the reverse-time loop needs).
template <bool STORE>
__device__ __forceinline__ void replay_step(float u, float dl, uint32_t a_bc, uint32_t slot,
                                            f32x2 (&h)[SCAN_NS / 2], const f32x2 (&a2)[SCAN_NS / 2]) {
  const float du_ = dl * u;
  const f32x2 dl2 = pk2(dl, dl), du2 = pk2(du_, du_);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 Bv = ldsf4(a_bc + 16u * q);
    float e0, e1, e2, e3;
    upk2(mul2(dl2, a2[2 * q]), e0, e1); upk2(mul2(dl2, a2[2 * q + 1]), e2, e3);
    h[2 * q] = fma2(pk2(ex2_approx(e0), ex2_approx(e1)), h[2 * q], mul2(du2, pk2(Bv.x, Bv.y)));
    h[2 * q + 1] = fma2(pk2(ex2_approx(e2), ex2_approx(e3)), h[2 * q + 1], mul2(du2, pk2(Bv.z, Bv.w)));
  }
  if (STORE) tmem_st16(slot, h);
}

// Cross-channel sums of one token: every lane contributes 32 values (dB[0..15] | dC[0..15] of its channel) and the warp
// needs the 32 column sums.  Through a padded 32 x 32 shared-memory tile owned by the warp, with everything kept as packed
// fp32 pairs: each packed product is stored straight from its 64-bit register (16 x st.shared.b64, no unpacking, no
// staging array), then lanes 0-15 sum rows 0-15 and lanes 16-31 rows 16-31 of column PAIR (2 hl, 2 hl + 1) with 16 x
// ld.shared.b64 + add.f32x2, one shuffle pair joins the halves, and lanes 0-15 write the 32 sums as 16 float2.  Row pitch
// 34 floats: the 16 lanes of a 64-bit access hit 16 distinct bank pairs in both the row-wise stores and the column-wise
// loads.  ~55 instructions per step against 74 + 32 unpacking moves for the scalar version of round 1 (in a kernel that
// is issue-bound).
// (a plain C++ store: through an asm statement ptxas copied every packed product into one fixed register pair first,
// 2 MOVs per store, 32 per step)
__device__ __forceinline__ void sts_pair(uint8_t* a, f32x2 v) { *reinterpret_cast<volatile f32x2*>(a) = v; }
__device__ __forceinline__ f32x2 lds_pair(uint32_t a) { f32x2 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }

// After every lane has stored its 16 pairs into its row: the column sums.  Returns, in lanes 0-15, the sums of columns
// (2 lane, 2 lane + 1) as (lo, hi).
__device__ __forceinline__ void smem_column_sums(uint32_t tile, int lane, float& lo, float& hi) {
  constexpr int P = 34;
  __syncwarp();
  const uint32_t base = tile + (uint32_t)(((lane >> 4) * 16 * P + 2 * (lane & 15)) * 4);
  f32x2 s0 = pk2(0.f, 0.f), s1 = s0;
#pragma unroll
  for (int r = 0; r < 16; r += 2) {
    s0 = add2(s0, lds_pair(base + (uint32_t)(r * P * 4)));
    s1 = add2(s1, lds_pair(base + (uint32_t)((r + 1) * P * 4)));
  }
  __syncwarp();                     // the tile is rewritten at the next step
  upk2(add2(s0, s1), lo, hi);
  lo += __shfl_xor_sync(0xffffffffu, lo, 16);
  hi += __shfl_xor_sync(0xffffffffu, hi, 16);
}

// sigmoid(x) with flush-to-zero MUFU ops (exp2 overflow -> rcp(inf) = 0)
__device__ __forceinline__ float sigmoid_ftz(float x) {
  float r;
  const float e = ex2_approx(-1.4426950408889634f * x);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

// The walk of one CTA.  SPEC = the training configuration of the AuM mixers with every per-launch option folded into the
// instruction stream: direction 0 walks forwards and owns the gate (z, y_pre, dz and out_z all present), direction 1 walks
// backwards and has none; du / ddelta are plain stores (one pair per direction), ddelta leaves multiplied by softplus',
// every channel of the CTA exists (D % 128 == 0).  The general instantiation (SPEC = false) reads the same options from
// the parameter block at run time; in the 2-step unrolled reverse-time loop that is ~90 of ~340 instructions per step
// (uniform branches, predicate set-up, 64-bit pointer selects) and, worse for a latency-bound kernel, it cuts the step
// into basic blocks the scheduler cannot interleave across.
// TG: element type of du / ddelta (float, or T in the specialised instantiation when the caller asked for 16-bit gradients)
template <typename T, bool SPEC, bool REV_, bool GATE_, typename TG, typename TD>
__device__ __forceinline__ void scan_bwd_cta(const ScanBwdMaps& maps, const ScanBwdParams& p, uint8_t* smem_raw) {
  using BL = BwdLayout<T, TD>;
  const uint32_t smem0 = (s_u32(smem_raw) + 127u) & ~127u;
  const uint32_t stages = smem0;
  const uint32_t redt = stages + 2 * BL::STAGE_BYTES;
  const uint32_t bars = redt + BL::RED_BYTES;
  const uint32_t tmem_slot = bars + 16;

  const int g = blockIdx.z;
  const ScanBwdDirDev& d = p.dir[g];
  const int tig = threadIdx.x, lane = tig & 31;
  const int ch_raw = blockIdx.x * BT_CH + tig;
  const bool active = SPEC ? true : (ch_raw < p.Dch);
  const int ch = active ? ch_raw : (p.Dch - 1);
  const int b = blockIdx.y;
  const int L = p.L;
  const bool rev = SPEC ? REV_ : (d.reverse != 0);
  const int row0 = b * L;
  const bool bidir = p.ndirs == 2;
  const bool accumulate = SPEC ? false : (bidir && p.shared_du);
  const bool has_z = SPEC ? true : (p.z != nullptr);
  const bool gate = SPEC ? GATE_ : ((g == 0) && has_z && (p.dz != nullptr || p.outz != nullptr));
  const bool want_y = SPEC ? GATE_ : (gate && p.ypre != nullptr);
  const float scale = p.scale;
  const int64_t rows_total = (int64_t)p.batch * L;

  if (tig == 0) {
    sbar_init(bars, 1); sbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tig < 32) {      // warp 0 owns the TMEM allocation (128 columns: up to four CTAs per SM could hold one)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // checkpoint chunking (identical to the forward kernels'): chunk 0 = [0, first), chunk c = [first + 8(c-1), +8)
  const int first = min(scan_ck_first(L, bidir, rev), L);
  const int nchunks = 1 + (L - first + BT_TT - 1) / BT_TT;
  const int nck_max = scan_ck_count_max(L);
  auto chunk_range = [&](int c, int& s0, int& ns) {
    if (c == 0) { s0 = 0; ns = first; } else { s0 = first + (c - 1) * BT_TT; ns = min(BT_TT, L - s0); }
  };

  // ---- producer (thread 0): chunk c -> stage (nchunks-1-c) & 1
  auto issue = [&](int c) {
    int s0, ns; chunk_range(c, s0, ns);
    const int stage = (nchunks - 1 - c) & 1;
    const uint32_t st = stages + (uint32_t)stage * BL::STAGE_BYTES;
    const uint32_t bar = bars + 8u * stage;
    // box rows: forward [row0+s0, +8); reverse [row0+L-s0-8, +8) so that step j sits at tile row 7-j
    const int brow = rev ? (row0 + L - s0 - BT_TT) : (row0 + s0);
    const int col = blockIdx.x * BT_CH;
    const uint32_t bc_bytes = (uint32_t)ns * SCAN_ROW * 4u;
    const uint32_t tx = BL::CK_BYTES + BL::D_BYTES + 2 * BL::A_BYTES + (has_z ? BL::A_BYTES : 0) +
                        (want_y ? BL::A_BYTES : 0) + bc_bytes;
    sbar_expect_tx(bar, tx);
    tma_load_2d(st + BL::OFF_CK, &maps.ck[g], col, (b * nck_max + c) * SCAN_NS, bar);
    tma_load_2d(st + BL::OFF_D, &maps.d[g], col, brow, bar);
    tma_load_2d(st + BL::OFF_U, &maps.u[g], col, brow, bar);
    tma_load_2d(st + BL::OFF_G, &maps.g, col, brow, bar);
    if (has_z) tma_load_2d(st + BL::OFF_Z, &maps.z, col, brow, bar);
    if (want_y) tma_load_2d(st + BL::OFF_Y, &maps.y, col, brow, bar);
    const int bc_row_lo = rev ? (L - s0 - ns) : s0;
    const float* src = d.BC + (int64_t)(row0 + bc_row_lo) * SCAN_ROW;
    const uint32_t dst = st + BL::OFF_BC + (rev ? (uint32_t)(BT_TT - ns) * SCAN_ROW * 4u : 0u);
    bulk_g2s(dst, src, bc_bytes, bar);
  };
  if (tig == 0) {
    issue(nchunks - 1);
    if (nchunks > 1) issue(nchunks - 2);
  }

  // ---- per-channel constants and accumulators.  a2 = A log2(e) serves both the decay exp2(dl a2) and - with one
  // multiplication by ln 2 at the end of the step - the A factor of d(delta) (no second copy of A in registers).
  f32x2 a2[SCAN_NS / 2];
  {
    const float4* ap = reinterpret_cast<const float4*>(d.A + (int64_t)ch * SCAN_NS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldg(ap + i);
      a2[2 * i] = pk2(v.x * 1.4426950408889634f, v.y * 1.4426950408889634f);
      a2[2 * i + 1] = pk2(v.z * 1.4426950408889634f, v.w * 1.4426950408889634f);
    }
  }
  const float Dv = d.D ? __ldg(d.D + ch) : 0.f;
  f32x2 gcar[SCAN_NS / 2], dA_acc[SCAN_NS / 2];
#pragma unroll
  for (int k = 0; k < SCAN_NS / 2; ++k) { gcar[k] = pk2(0.f, 0.f); dA_acc[k] = pk2(0.f, 0.f); }
  float dD_acc = 0.f;

  // signed strides of one step forwards in time (tile rows / global rows run backwards for the reverse direction)
  const int s16 = rev ? -(BT_CH * (int)sizeof(T)) : (BT_CH * (int)sizeof(T));
  const int s32 = rev ? -(BT_CH * (int)sizeof(TD)) : (BT_CH * (int)sizeof(TD));     // delta rows
  const int sbc = rev ? -(SCAN_ROW * 4) : (SCAN_ROW * 4);
  const int rstep = rev ? -1 : 1;
  const int part = blockIdx.x * (BT_CH / 32) + (tig >> 5);     // this warp's slice of the dB|dC partial workspace
  const uint32_t hcol = tmem_base + ((uint32_t)((tig >> 5) * 32) << 16);   // slot 0 of this warp's TMEM lane quarter
  const bool spg_on = SPEC ? true : (p.softplus_grad != 0);
  const uint32_t red_tile = redt + (uint32_t)((tig >> 5) * 32 * BL::RED_PITCH * 4);
  uint8_t* const red_row = smem_raw + (red_tile - s_u32(smem_raw)) + lane * BL::RED_PITCH * 4;
  // one step backwards in time moves one global row against the walk direction
  const int64_t gdu = -(int64_t)rstep * d.ld_du, gdd = -(int64_t)rstep * d.ld_dd;
  const int64_t gdz = -(int64_t)rstep * p.ld_dz, goz = -(int64_t)rstep * p.ld_oz;
  const int gws = -rstep * 32;

  for (int q = 0; q < nchunks; ++q) {
    const int c = nchunks - 1 - q;
    int s0, ns; chunk_range(c, s0, ns);
    const int stage = q & 1;
    const uint32_t st = stages + (uint32_t)stage * BL::STAGE_BYTES;
    sbar_wait(bars + 8u * stage, (uint32_t)((q >> 1) & 1));

    const int row_first = rev ? (BT_TT - 1) : 0;                 // tile row of step 0 of the chunk
    const uint32_t e16 = (uint32_t)(row_first * BT_CH + tig) * (uint32_t)sizeof(T);
    const uint32_t e32 = (uint32_t)(row_first * BT_CH + tig) * (uint32_t)sizeof(TD);
    const uint32_t t_u = st + BL::OFF_U + e16, t_d = st + BL::OFF_D + e32;
    const uint32_t t_bc = st + BL::OFF_BC + (uint32_t)row_first * SCAN_ROW * 4u;

    // ---- replay: slot j = state before step j (slot 0 = the forward kernel's checkpoint)
    f32x2 hcur[SCAN_NS / 2];          // h_s of the reverse-time step in progress
    {
      f32x2 h[SCAN_NS / 2];
      const uint32_t ck = st + BL::OFF_CK + (uint32_t)tig * 4u;
#pragma unroll
      for (int k = 0; k < SCAN_NS / 2; ++k) h[k] = pk2(ldsf(ck + (uint32_t)(2 * k) * (BT_CH * 4)), ldsf(ck + (uint32_t)(2 * k + 1) * (BT_CH * 4)));
      tmem_st16(hcol, h);
      if (SPEC && ns == BT_TT) {       // full chunk: every shared-memory / tensor-memory address is base + immediate
#pragma unroll
        for (int j = 0; j < BT_TT - 1; ++j)
          replay_step<true>(ldst<T>(t_u + (uint32_t)(j * s16)), ldst<TD>(t_d + (uint32_t)(j * s32)), t_bc + (uint32_t)(j * sbc),
                            hcol + (uint32_t)(j + 1) * SCAN_NS, h, a2);
        replay_step<false>(ldst<T>(t_u + (uint32_t)((BT_TT - 1) * s16)), ldst<TD>(t_d + (uint32_t)((BT_TT - 1) * s32)),
                           t_bc + (uint32_t)((BT_TT - 1) * sbc), 0u, h, a2);
      } else {
        uint32_t a_u = t_u, a_d = t_d, a_bc = t_bc, slot = hcol + SCAN_NS;
#pragma unroll 1
        for (int j = 0; j < ns - 1; ++j) {
          replay_step<true>(ldst<T>(a_u), ldst<TD>(a_d), a_bc, slot, h, a2);
          a_u += s16; a_d += s32; a_bc += sbc; slot += SCAN_NS;
        }
        // the chunk's last step: its result h_{ns-1} is not history (it is the next chunk's checkpoint) but the
        // reverse-time loop below starts with it, and from there on carries h_s over from the h_{s-1} it loads
        replay_step<false>(ldst<T>(a_u), ldst<TD>(a_d), a_bc, slot, h, a2);
      }
#pragma unroll
      for (int k = 0; k < SCAN_NS / 2; ++k) hcur[k] = h[k];
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");   // the history is read back by the same thread below
    }

    // ---- reverse-time recurrence over the chunk, j = ns-1 .. 0
    {
      const int jl = ns - 1;
      uint32_t a_u = t_u + (uint32_t)(jl * s16), a_d = t_d + (uint32_t)(jl * s32), a_bc = t_bc + (uint32_t)(jl * sbc);
      uint32_t a_g = st + BL::OFF_G + e16 + (uint32_t)(jl * s16);
      uint32_t a_z = st + BL::OFF_Z + e16 + (uint32_t)(jl * s16);
      uint32_t a_y = st + BL::OFF_Y + e16 + (uint32_t)(jl * s16);
      uint32_t slot = hcol + (uint32_t)jl * SCAN_NS;
      const int64_t r_last = (int64_t)row0 + (rev ? (L - 1 - (s0 + jl)) : (s0 + jl));      // global row of step s0+jl
      TG* dup = reinterpret_cast<TG*>(d.du) + r_last * d.ld_du + ch;
      TG* ddp = reinterpret_cast<TG*>(d.ddelta) + r_last * d.ld_dd + ch;
      float* wsp = d.dbc_ws + ((int64_t)part * rows_total + r_last) * 32 + 2 * (lane & 15);   // lanes 0-15 store float2
      T* dzp = (SPEC ? GATE_ : (p.dz != nullptr)) ? reinterpret_cast<T*>(p.dz) + r_last * p.ld_dz + ch : nullptr;
      T* ozp = (SPEC ? GATE_ : (p.outz != nullptr)) ? reinterpret_cast<T*>(p.outz) + r_last * p.ld_oz + ch : nullptr;

      // one reverse-time step; every o* is the (compile-time, when unrolled) offset of the step from the a_* bases
      auto rstep_fn = [&](const uint32_t o16, const uint32_t o32, const uint32_t obc, const uint32_t oslot) {
        const float u = ldst<T>(a_u + o16), dl = ldst<TD>(a_d + o32), go = ldst<T>(a_g + o16) * scale;
        const float zv = has_z ? ldst<T>(a_z + o16) : 0.f;
        const float sg = has_z ? sigmoid_ftz(zv) : 1.f;          // sigmoid(z); silu(z) = z sg
        const float sz = has_z ? zv * sg : 1.f;
        const float dy = go * sz;
        dD_acc = fmaf(dy, u, dD_acc);
        const float dlu = dl * u;
        const f32x2 dl2 = pk2(dl, dl), dy2 = pk2(dy, dy), dlu2 = pk2(dlu, dlu);
        f32x2 sB2 = pk2(0.f, 0.f), dd2 = pk2(0.f, 0.f);
        f32x2 hprev[SCAN_NS / 2];
        tmem_ld16(slot + oslot, hprev);                                         // h_{s-1}, all 16 states
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float4 Bv = ldsf4(a_bc + obc + 16u * qq);
          const float4 Cv = ldsf4(a_bc + obc + 16u * (4 + qq));
          const f32x2 hp[2] = {hprev[2 * qq], hprev[2 * qq + 1]};
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const int k = 2 * qq + hq;                              // state pair (2k, 2k+1)
            const f32x2 Bp = hq ? pk2(Bv.z, Bv.w) : pk2(Bv.x, Bv.y);
            const f32x2 Cp = hq ? pk2(Cv.z, Cv.w) : pk2(Cv.x, Cv.y);
            float e0, e1; upk2(mul2(dl2, a2[k]), e0, e1);
            const f32x2 a = pk2(ex2_approx(e0), ex2_approx(e1));
            const f32x2 dh = fma2(Cp, dy2, gcar[k]);
            gcar[k] = mul2(a, dh);
            const f32x2 t1 = mul2(gcar[k], hp[hq]);                 // dh * a * h_{s-1}
            dA_acc[k] = fma2(t1, dl2, dA_acc[k]);
            dd2 = fma2(t1, a2[k], dd2);                             // x log2(e); undone below
            sB2 = fma2(dh, Bp, sB2);
            sts_pair(red_row + 8 * k, mul2(dh, dlu2));                    // dB contributions
            sts_pair(red_row + 8 * (SCAN_NS / 2 + k), mul2(hcur[k], dy2));   // dC contributions (h_s carried over)
            hcur[k] = hp[hq];                                       // h_{s-1} is the next (earlier) step's h_s
          }
        }
        float s0_, s1_, d0_, d1_;
        upk2(sB2, s0_, s1_); upk2(dd2, d0_, d1_);
        const float sB = s0_ + s1_;
        float dd = fmaf(sB, u, (d0_ + d1_) * 0.6931471805599453f);
        const float duv = fmaf(dl, sB, Dv * dy);
        // cross-channel sums of this token: lane i of each warp ends up with value i; one plain 128-byte store per
        // warp into this warp's slice of the partial workspace (summed over warps by dbc_reduce_kernel).
        // Channels past D read zero-filled tiles, so their contributions are exact zeros.
        {
          float c_lo, c_hi;
          smem_column_sums(red_tile, lane, c_lo, c_hi);
          if (lane < 16) *reinterpret_cast<float2*>(wsp) = make_float2(c_lo, c_hi);
        }
        if (spg_on) dd *= 1.f - ex2_approx(-1.4426950408889634f * dl);   // softplus'(pre) = 1 - exp(-delta), delta >= 0
        if (active) {
          if constexpr (std::is_same<TG, float>::value) {
            if (accumulate) { red_add_f32(dup, duv); red_add_f32(ddp, dd); }
            else { *dup = duv; *ddp = dd; }
          } else {
            *dup = from_f<TG>(duv); *ddp = from_f<TG>(dd);
          }
          if (gate) {
            const float yp = want_y ? ldst<T>(a_y + o16) : 0.f;
            if (dzp) *dzp = from_f<T>(go * yp * (sg * (1.f + zv * (1.f - sg))));
            if (ozp) *ozp = from_f<T>(scale * yp * sz);
          }
        }
        dup += gdu; ddp += gdd; wsp += gws;
        if (dzp) dzp += gdz;
        if (ozp) ozp += goz;
      };

      if (SPEC && ns == BT_TT) {
        // full chunk: two steps per trip, the second at compile-time offsets from the same bases
        constexpr int S16 = REV_ ? -(BT_CH * (int)sizeof(T)) : (BT_CH * (int)sizeof(T));
        constexpr int S32 = REV_ ? -(BT_CH * (int)sizeof(TD)) : (BT_CH * (int)sizeof(TD));
        constexpr int SBC = REV_ ? -(SCAN_ROW * 4) : (SCAN_ROW * 4);
#pragma unroll 1
        for (int j = 0; j < BT_TT / 2; ++j) {
          rstep_fn(0u, 0u, 0u, 0u);
          rstep_fn((uint32_t)(-S16), (uint32_t)(-S32), (uint32_t)(-SBC), (uint32_t)(-SCAN_NS));
          a_u -= 2 * S16; a_g -= 2 * S16; a_z -= 2 * S16; a_y -= 2 * S16; a_d -= 2 * S32; a_bc -= 2 * SBC; slot -= 2 * SCAN_NS;
        }
      } else {
#pragma unroll 1
        for (int j = jl; j >= 0; --j) {
          rstep_fn(0u, 0u, 0u, 0u);
          a_u -= s16; a_g -= s16; a_z -= s16; a_y -= s16; a_d -= s32; a_bc -= sbc; slot -= SCAN_NS;
        }
      }
    }

    // stage free: refill it with the chunk after next
    __syncthreads();
    if (tig == 0 && c >= 2) issue(c - 2);
  }

  if (active) {
#pragma unroll
    for (int k = 0; k < SCAN_NS / 2; ++k) {
      // (a2 = A log2(e): dA * A = dA * a2 * ln 2)
      const f32x2 ln2 = pk2(0.6931471805599453f, 0.6931471805599453f);
      float lo, hi; upk2(d.dA_log ? mul2(mul2(dA_acc[k], a2[k]), ln2) : dA_acc[k], lo, hi);
      atomicAdd(d.dA + (int64_t)ch * SCAN_NS + 2 * k, lo);
      atomicAdd(d.dA + (int64_t)ch * SCAN_NS + 2 * k + 1, hi);
    }
    if (d.dD) atomicAdd(d.dD + ch, dD_acc);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tig < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BT_TMEM_COLS) : "memory");
}

// G16: du / ddelta of the activation dtype; D16: delta of the activation dtype.  MINB: resident CTAs per SM the register
// budget is set for: 3 (168 registers) in general; 4 (128 registers, 16 warps per SM, all 512 TMEM columns) where the stages
// are small enough for four CTAs to fit in shared memory, i.e. with a 16-bit delta (56.5 KB per CTA) - the kernel is
// latency-bound (issue slots 45 % used at 12 warps).  Measured: no gain (see launch_bt), so 3 is what ships.
template <typename T, bool SPEC, bool G16, bool D16, int MINB>
__global__ void __launch_bounds__(BT_CH, MINB)
scan_bwd_tma_kernel(const __grid_constant__ ScanBwdMaps maps, const ScanBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  using TG = typename std::conditional<G16, T, float>::type;
  using TD = typename std::conditional<D16, T, float>::type;
  if (SPEC) {
    if (blockIdx.z == 0) scan_bwd_cta<T, true, false, true, TG, TD>(maps, p, smem_raw);
    else                 scan_bwd_cta<T, true, true, false, TG, TD>(maps, p, smem_raw);
  } else {
    scan_bwd_cta<T, false, false, false, float, TD>(maps, p, smem_raw);
  }
}

// SPEC eligibility: see scan_bwd_cta
static bool bwd_spec_ok(const ScanBwdParams& p) {
  if (p.z == nullptr || p.ypre == nullptr || p.dz == nullptr || p.outz == nullptr || !p.softplus_grad) return false;
  if (p.Dch % BT_CH != 0 || p.dir[0].reverse != 0) return false;
  if (p.ndirs == 2 && (p.dir[1].reverse == 0 || p.shared_du)) return false;
  static int off = -1;
  if (off < 0) off = getenv("AUM_SCAN_BWD_NOSPEC") != nullptr ? 1 : 0;
  return off == 0;
}

template <typename T, bool SPEC, bool G16, bool D16, int MINB>
static int launch_bt_v(const ScanBwdMaps& maps, const ScanBwdParams& p, cudaStream_t st) {
  using BL = BwdLayout<T, typename std::conditional<D16, T, float>::type>;
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(scan_bwd_tma_kernel<T, SPEC, G16, D16, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, BL::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("aum_selective_scan_bwd: cudaFuncSetAttribute(smem=%d): %s", BL::SMEM_BYTES, cudaGetErrorString(e)); return 2; }
    attr_set = true;
  }
  dim3 grid(ceil_div(p.Dch, BT_CH), p.batch, p.ndirs);
  scan_bwd_tma_kernel<T, SPEC, G16, D16, MINB><<<grid, BT_CH, BL::SMEM_BYTES, st>>>(maps, p);
  return check_launch("aum_selective_scan_bwd(tma)");
}

template <typename T>
static int launch_bt(const ScanBwdMaps& maps, const ScanBwdParams& p, cudaStream_t st) {
  constexpr bool T16 = !std::is_same<T, float>::value;
  const bool spec = bwd_spec_ok(p);
  if ((p.g16 && !spec) || ((p.g16 || p.d16) && !T16)) return -1;       // (the entry point turns this into an error)
  // 4 CTAs per SM (AUM_SCAN_BWD_4CTA=1) measured no faster than 3: 0.9575 vs 0.9544 ms per config-3 launch, 47.5 vs 47.7 ms
  // per training step (noise) - the 56 bytes the 128-register cap spills cost what the fourth CTA's latency hiding returns
  static int minb3 = -1;
  if (minb3 < 0) minb3 = getenv("AUM_SCAN_BWD_4CTA") != nullptr ? 0 : 1;
  if constexpr (T16) {
    if (p.d16) {
      if (spec && p.g16) return minb3 ? launch_bt_v<T, true, true, true, 3>(maps, p, st) : launch_bt_v<T, true, true, true, 4>(maps, p, st);
      if (spec) return launch_bt_v<T, true, false, true, 3>(maps, p, st);
      return launch_bt_v<T, false, false, true, 3>(maps, p, st);
    }
    if (spec && p.g16) return launch_bt_v<T, true, true, false, 3>(maps, p, st);
  }
  if (spec) return launch_bt_v<T, true, false, false, 3>(maps, p, st);
  return launch_bt_v<T, false, false, false, 3>(maps, p, st);
}

int launch_scan_bwd_tma(const ScanBwdParams& p, int dtype, cudaStream_t st) {
  if (!tma_available() || getenv("AUM_SCAN_BWD_GENERIC") != nullptr) return -1;
  const int esz = dtype_size(dtype);
  auto ok_mat = [](const void* base, int64_t ld, int sz) { return aligned16(base) && (ld * sz) % 16 == 0; };
  if (!ok_mat(p.dout, p.ld_dout, esz)) return -1;
  if (p.z && !ok_mat(p.z, p.ld_z, esz)) return -1;
  if (p.ypre && !ok_mat(p.ypre, p.ld_y, esz)) return -1;
  if (p.Dch % 4 != 0) return -1;                                   // checkpoint rows [.., D] fp32 must be 16-byte pitched
  for (int g = 0; g < p.ndirs; ++g) {
    const ScanBwdDirDev& d = p.dir[g];
    if (!d.ckpt_valid || d.ld_bc != SCAN_ROW || !aligned16(d.BC) || !aligned16(d.ckpt)) return -1;
    if (!ok_mat(d.u, d.ld_u, esz) || !ok_mat(d.delta, d.ld_delta, p.d16 ? esz : 4)) return -1;
  }
  ScanBwdMaps maps;
  const int64_t rows = (int64_t)p.batch * p.L;
  const int64_t ck_rows = (int64_t)p.batch * scan_ck_count_max(p.L) * SCAN_NS;
  for (int g = 0; g < 2; ++g) {
    const ScanBwdDirDev& d = p.dir[g < p.ndirs ? g : 0];
    if (int rc = tma_encode_2d(&maps.u[g], d.u, dtype, rows, p.Dch, d.ld_u, BT_TT, BT_CH, false, "aum_selective_scan_bwd(u)")) return rc;
    if (int rc = tma_encode_2d(&maps.d[g], d.delta, p.d16 ? dtype : AUM_F32, rows, p.Dch, d.ld_delta, BT_TT, BT_CH, false, "aum_selective_scan_bwd(delta)")) return rc;
    if (int rc = tma_encode_2d(&maps.ck[g], d.ckpt, AUM_F32, ck_rows, p.Dch, p.Dch, SCAN_NS, BT_CH, false, "aum_selective_scan_bwd(ckpt)")) return rc;
  }
  if (int rc = tma_encode_2d(&maps.g, p.dout, dtype, rows, p.Dch, p.ld_dout, BT_TT, BT_CH, false, "aum_selective_scan_bwd(dout)")) return rc;
  if (p.z) { if (int rc = tma_encode_2d(&maps.z, p.z, dtype, rows, p.Dch, p.ld_z, BT_TT, BT_CH, false, "aum_selective_scan_bwd(z)")) return rc; }
  else maps.z = maps.g;
  if (p.ypre) { if (int rc = tma_encode_2d(&maps.y, p.ypre, dtype, rows, p.Dch, p.ld_y, BT_TT, BT_CH, false, "aum_selective_scan_bwd(y_pre)")) return rc; }
  else maps.y = maps.g;
  // shared du / ddelta: both directions add into zeroed buffers
  if (p.ndirs == 2 && p.shared_du) {
    const ScanBwdDirDev& d = p.dir[0];
    cudaError_t e = cudaMemset2DAsync(d.du, (size_t)d.ld_du * 4, 0, (size_t)p.Dch * 4, (size_t)rows, st);
    if (e == cudaSuccess) e = cudaMemset2DAsync(d.ddelta, (size_t)d.ld_dd * 4, 0, (size_t)p.Dch * 4, (size_t)rows, st);
    if (e != cudaSuccess) { set_error("aum_selective_scan_bwd: cudaMemset2DAsync: %s", cudaGetErrorString(e)); return 2; }
  }
  switch (dtype) {
    case AUM_F32:  return launch_bt<float>(maps, p, st);
    case AUM_F16:  return launch_bt<__half>(maps, p, st);
    case AUM_BF16: return launch_bt<__nv_bfloat16>(maps, p, st);
  }
  return -1;
}

}  // namespace aum
