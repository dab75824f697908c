// Selective scan forward, both time directions in one launch — TMA-streamed kernel (the production path).
//
// Same math and the same two-directions-one-pass scheme as scan_fwd.cu (see there for the reference citations:
// selective_scan_interface.py:37,213,354,499,503-507 and selective_scan_ref :86-152); what changes is how the
// operands reach the math:
//   * grid = (ceil(D/128), batch); a CTA owns 128 channels of one sequence; 4 warps walk it forwards and, in
//     the bidirectional case, 4 more warps walk it backwards (one thread per channel and direction, the 16
//     recurrences of the channel in registers as 8 packed fp32x2 pairs, FMUL2/FFMA2 + ex2.approx).
//   * every per-token operand is fetched by the TMA unit into a 4-stage shared-memory ring per direction:
//     per 8-token tile one cp.async.bulk.tensor 2-D box each for u, delta (and z once the walk is in its
//     finalising half) plus one 1-D bulk copy of the packed fp32 [B|C] rows, all completing on the stage's
//     "full" mbarrier; warps hand stages back through an "empty" mbarrier and one elected thread refills them
//     one tile late, so 2-3 tiles (16-24 tokens) of loads are always in flight and the math warps never touch
//     a global-load scoreboard for them.
//   * results leave the same way: each thread overwrites its own u element of the stage with y, and the elected
//     thread sends the finished 8 x 128 tile to `out` with one bulk tensor store (no per-step global store, no
//     per-step address arithmetic: every shared-memory access in the unrolled tile body has an immediate offset,
//     the walk direction being a template parameter).
//   * the partial y of the first half-walk is parked in `out` exactly as in the generic kernel; the second
//     half receives it back as a fourth TMA tile per stage (issued after the CTA-wide phase barrier).
// Eligibility (checked in launch_scan_tma): d_state == 16, packed fp32 [B|C] rows, delta in fp32 or in the activation dtype, without
// bias/softplus left to apply, 16-byte aligned bases and row pitches.  Anything else runs scan_fwd.cu.
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

#include "scan_common.cuh"
#include "tma.cuh"

namespace aum {

constexpr int ST_TT = 8;      // tokens per tile
// steps of a full tile unrolled per loop trip (the rest of the tile is a rolled loop over trips): 4 = two trips per tile
#ifndef AUM_SCAN_UNROLL
#define AUM_SCAN_UNROLL 4
#endif

// Which lane of a (converged) warp does the TMA bookkeeping: the one elect.sync picks (default), so that ptxas emits the
// UTMALDG / UTMASTG / UBLKCP of the owner's duties straight instead of wrapping each one in an ELECT / BRA.U.ANY loop, as it
// does inside an `if (lane == 0)` region (the same change gave the tcgen05 GEMMs' MMA issuer 6 %).  Measured on B200
// (profiles/r2_scan_elect_ab.txt): 0.524 -> 0.510 ms per 64-sequence launch, 0.280 -> 0.272 ms per 32-sequence launch,
// every scan parity test green.  Every site runs with the warp converged and the full mask, so the elected lane is the
// same one each time (it owns the bulk async-groups).  -DAUM_SCAN_ELECT=0 restores the lane-0 build.
#ifndef AUM_SCAN_ELECT
#define AUM_SCAN_ELECT 1
#endif
#if AUM_SCAN_ELECT
__device__ __forceinline__ bool scan_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
#define AUM_LEAD(lane0_expr) scan_elect()
#define AUM_LEAD0(tig0_expr) (warp_in_group == 0 && scan_elect())
#else
#define AUM_LEAD(lane0_expr) (lane0_expr)
#define AUM_LEAD0(tig0_expr) (tig0_expr)
#endif
// channels per CTA (= threads per direction) is a template parameter CH: 128 (default: two CTAs = 16 warps per SM, the
// kernel owns the register file) or 192 (AUM_SCAN_TMA_CH=192: one 12-warp CTA per SM that leaves 16 K registers and
// ~95 KB of shared memory free).  Measured at config 2 (fp16, pre-gated z): 0.541 ms vs 0.582 ms - the XU pipe needs
// the 16 warps; 10 warps (CH = 160) do not divide over the 4 schedulers and take 0.73 ms.  Running one 8-warp scan CTA
// next to a half-size tcgen05 GEMM CTA on every SM (so that the XU and the tensor pipe overlap) was tried as well:
// the scan loses more (8 warps: -35 %) than the overlap wins, whole-model throughput -5 %; see DESIGN.md.

// NSTG: ring depth per direction.  TD: element type of delta (float, or T when dt_proj's epilogue already rounded
// delta to the activation dtype - the reference rounds it there too, selective_scan_interface.py:468 under autocast).
template <typename T, typename TD, int NSTG, int CH> struct StageLayout {
  static constexpr int U_BYTES = ST_TT * CH * (int)sizeof(T);
  static constexpr int D_BYTES = ST_TT * CH * (int)sizeof(TD);
  static constexpr int Z_BYTES = U_BYTES;
  static constexpr int BC_BYTES = ST_TT * SCAN_ROW * 4;
  static constexpr int P_BYTES = U_BYTES;                  // parked partials of the other direction (rows of `out`)
  static constexpr int OFF_U = 0, OFF_D = OFF_U + U_BYTES, OFF_Z = OFF_D + D_BYTES, OFF_BC = OFF_Z + Z_BYTES;
  static constexpr int OFF_P = OFF_BC + BC_BYTES;
  static constexpr int STAGE_BYTES = OFF_P + P_BYTES;
  static constexpr int GROUP_BYTES = NSTG * STAGE_BYTES;
  static constexpr int SMEM_BYTES = 2 * GROUP_BYTES + 128 /*align slack*/ + 2 * 3 * NSTG * 8 /*mbarriers*/;
};

struct ScanTmaMaps { CUtensorMap u[2], d[2], z, o; };

__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(col), "r"(row), "r"(bar) : "memory");
}
__device__ __forceinline__ void sbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float lds_f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
template <typename T> __device__ __forceinline__ float lds_t(uint32_t a);
template <> __device__ __forceinline__ float lds_t<float>(uint32_t a) { return lds_f(a); }
template <> __device__ __forceinline__ float lds_t<__half>(uint32_t a) {
  unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return __half2float(__ushort_as_half(v));
}
template <> __device__ __forceinline__ float lds_t<__nv_bfloat16>(uint32_t a) {
  unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return __uint_as_float(((uint32_t)v) << 16);
}

template <typename T> __device__ __forceinline__ void sts_t(uint32_t a, float v);
template <> __device__ __forceinline__ void sts_t<float>(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <> __device__ __forceinline__ void sts_t<__half>(uint32_t a, float v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(__half_as_ushort(__float2half_rn(v))) : "memory");
}
template <> __device__ __forceinline__ void sts_t<__nv_bfloat16>(uint32_t a, float v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(v))) : "memory");
}
__device__ __forceinline__ void tma_store_tile_2d(const CUtensorMap* tm, uint32_t src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(col), "r"(row) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// One recurrence step of one channel: consumes (u, delta') and the staged B|C row at `a_bc`, returns y (+D u).
__device__ __forceinline__ float scan_step(float u, float dl, uint32_t a_bc, float Dv,
                                           f32x2 (&h)[SCAN_NS / 2], const f32x2 (&a2)[SCAN_NS / 2]) {
  const float du = dl * u;
  const f32x2 dl2 = pk2(dl, dl), du2 = pk2(du, du);
  f32x2 ya = pk2(Dv * u, 0.f), yb = pk2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < SCAN_NS / 4; ++q) {
    const float4 Bv = lds_f4(a_bc + 16u * q);
    const float4 Cv = lds_f4(a_bc + 16u * (SCAN_NS / 4 + q));
    const f32x2 x0 = mul2(dl2, a2[2 * q]), x1 = mul2(dl2, a2[2 * q + 1]);
    float e0, e1, e2, e3;
    upk2(x0, e0, e1); upk2(x1, e2, e3);
    const f32x2 dA0 = pk2(ex2_approx(e0), ex2_approx(e1));
    const f32x2 dA1 = pk2(ex2_approx(e2), ex2_approx(e3));
    h[2 * q] = fma2(dA0, h[2 * q], mul2(du2, pk2(Bv.x, Bv.y)));
    h[2 * q + 1] = fma2(dA1, h[2 * q + 1], mul2(du2, pk2(Bv.z, Bv.w)));
    ya = fma2(h[2 * q], pk2(Cv.x, Cv.y), ya);
    yb = fma2(h[2 * q + 1], pk2(Cv.z, Cv.w), yb);
  }
  float y0, y1, y2, y3;
  upk2(ya, y0, y1); upk2(yb, y2, y3);
  return (y0 + y1) + (y2 + y3);
}

// A full 8-step tile, unguarded.  a_*: this thread's element in tile row 0 of the stage.
// REV: the walk runs down the token axis, so step t sits in tile row 7-t (compile-time: every access below has an
// immediate offset).  ZM: 0 no gate, 1 y *= silu(z), 2 y *= z (z pre-gated by the in_proj epilogue).
// PARTIAL: the other direction's parked partial of each step sits in the stage's P tile.  y overwrites u in place;
// the caller bulk-stores the tile.  YPRE: also save the pre-gate y (training) to pyp, the global row of step 0,
// ostep its signed stride.
// The body is unrolled by 4 steps and run twice: ~5.5 KB of SASS per instantiation, so that the four bodies a
// resident CTA pair can be in at once (2 directions x 2 phases) stay inside the 32 KB L1.5 instruction cache; the
// fully unrolled 8-step bodies did not and lost 7% to instruction-fetch stalls.
template <typename T, typename TD, int CH, bool FIN, bool PARTIAL, int ZM, bool REV, bool YPRE>
__device__ __forceinline__ void scan_tile_full(uint32_t a_u, uint32_t a_d, uint32_t a_z, uint32_t a_bc, uint32_t a_p,
                                               float Dv, float oscale, bool active,
                                               f32x2 (&h)[SCAN_NS / 2], const f32x2 (&a2)[SCAN_NS / 2],
                                               T* pyp, ptrdiff_t ostep) {
  constexpr int HALF = AUM_SCAN_UNROLL;             // steps per trip
  static_assert(ST_TT % HALF == 0, "AUM_SCAN_UNROLL must divide the tile");
  constexpr int P16 = CH * (int)sizeof(T), P32 = CH * (int)sizeof(TD), PBC = SCAN_ROW * 4;
  if (REV) { a_u += (ST_TT - HALF) * P16; a_z += (ST_TT - HALF) * P16; a_p += (ST_TT - HALF) * P16; a_d += (ST_TT - HALF) * P32; a_bc += (ST_TT - HALF) * PBC; }
#pragma unroll 1
  for (int hf = 0; hf < ST_TT / HALF; ++hf) {
#pragma unroll
    for (int t = 0; t < HALF; ++t) {
      const int r = REV ? (HALF - 1 - t) : t;
      const uint32_t o16 = (uint32_t)(r * P16);
      float y = scan_step(lds_t<T>(a_u + o16), lds_t<TD>(a_d + (uint32_t)(r * P32)), a_bc + (uint32_t)(r * PBC), Dv, h, a2);
      if (FIN) {
        if (PARTIAL) y += lds_t<T>(a_p + o16);
        if (YPRE) { if (active) *pyp = from_f<T>(y); pyp += ostep; }   // pre-gate y saved for the backward pass
        if (ZM == 1) y *= silu_f(lds_t<T>(a_z + o16));
        if (ZM == 2) y *= lds_t<T>(a_z + o16);
        y *= oscale;
      }
      sts_t<T>(a_u + o16, y);
    }
    if (REV) { a_u -= HALF * P16; a_z -= HALF * P16; a_p -= HALF * P16; a_d -= HALF * P32; a_bc -= HALF * PBC; }
    else     { a_u += HALF * P16; a_z += HALF * P16; a_p += HALF * P16; a_d += HALF * P32; a_bc += HALF * PBC; }
  }
}

// A short tile (first tile of the walk or its last): rolled loop, parked partials read directly.
template <typename T, typename TD>
__device__ __forceinline__ void scan_tile_tail(int nt, bool fin, bool partial, bool has_z,
                                               uint32_t a_u, uint32_t a_d, uint32_t a_z, uint32_t a_bc, uint32_t a_p,
                                               int su, int sd, int sbc, float Dv, float oscale, bool active,
                                               f32x2 (&h)[SCAN_NS / 2], const f32x2 (&a2)[SCAN_NS / 2],
                                               T* po, ptrdiff_t ostep, ptrdiff_t ypre_off, bool zpre) {
#pragma unroll 1
  for (int t = 0; t < nt; ++t) {
    float y = scan_step(lds_t<T>(a_u), lds_t<TD>(a_d), a_bc, Dv, h, a2);
    if (fin) {
      if (partial) y += lds_t<T>(a_p);
      if (ypre_off != 0 && active) po[ypre_off] = from_f<T>(y);
      if (has_z) { const float zv = lds_t<T>(a_z); y *= zpre ? zv : silu_f(zv); }
      y *= oscale;
    }
    if (active) *po = from_f<T>(y);
    po += ostep;
    a_u += su; a_d += sd; a_z += su; a_bc += sbc; a_p += su;
  }
}

// CKPT: the training instantiation (state checkpoints before every tile, for aum_selective_scan_bwd); inference
// launches run the CKPT = false instantiation, whose tile loops carry no checkpoint stores (smaller loop bodies: the
// steady-state loops of the resident CTAs compete for the 32 KB instruction cache).
template <typename T, typename TD, int MINB, int NSTG, int CH, bool CKPT>
__global__ void __launch_bounds__(CH <= 128 ? 2 * CH : 512, MINB)      // either way: at most 128 registers per thread
scan_fwd_tma_kernel(const __grid_constant__ ScanTmaMaps maps, const ScanParams p) {
  using SL = StageLayout<T, TD, NSTG, CH>;
  constexpr int NW = CH / 32;                 // warps per direction
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (s_u32(smem_raw) + 127u) & ~127u;

  // Direction slot = parity of the warp index when both directions run: the two directions execute different
  // (direction-specialised) tile bodies, and with warp w resident on scheduler w mod 4 this keeps each scheduler's
  // instruction stream to a single body (a [dir0 x4 | dir1 x4] split put both on every scheduler: 7% slower).
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = (p.ndirs == 2) ? (wid & 1) : 0;
  const int warp_in_group = (p.ndirs == 2) ? (wid >> 1) : wid;
  const int tig = warp_in_group * 32 + lane;
  const ScanDirDev& d = p.dir[g];
  const int ch_raw = blockIdx.x * CH + tig;
  const bool active = ch_raw < p.Dch;
  const int ch = active ? ch_raw : (p.Dch - 1);
  const int b = blockIdx.y;
  const int L = p.L;
  const bool bidir = p.ndirs == 2;
  const bool rev = d.reverse != 0;
  const int row0 = b * L;
  const bool has_z = p.z != nullptr;
  const bool zpre = p.z_pregated != 0;
  const int zmode = has_z ? (zpre ? 2 : 1) : 0;

  const uint32_t ring = smem0 + (uint32_t)g * SL::GROUP_BYTES;
  const uint32_t bars = smem0 + (uint32_t)p.ndirs * SL::GROUP_BYTES + (uint32_t)g * (3 * NSTG * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTG + s); };
  auto pfull_bar = [&](int s) { return bars + 8u * (2 * NSTG + s); };   // parked-partial tiles (phase 2 only)

  if (tig == 0) {
    for (int s = 0; s < NSTG; ++s) { sbar_init(full_bar(s), 1); sbar_init(empty_bar(s), NW); sbar_init(pfull_bar(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // ---- tiling of this direction's walk.  Steps s = 0..L-1 (token l = rev ? L-1-s : s).
  // Phase 1 = steps [0,S1) park partials, phase 2 = steps [S1,L) finalise; the first tile is shortened so that
  // S1 falls on a tile boundary.
  const int mid = L / 2;
  const int S1 = bidir ? (rev ? (L - mid) : mid) : 0;
  const int first = (S1 % ST_TT) ? (S1 % ST_TT) : ST_TT;          // length of tile 0 when S1 > 0
  const int n1t = S1 > 0 ? 1 + (S1 - first + ST_TT - 1) / ST_TT : 0;
  const int n2t = (L - S1 + ST_TT - 1) / ST_TT;
  const int ntiles = n1t + n2t;
  // S1 = first + 8 (n1t - 1), so one formula serves both phases: tile k >= 1 starts at first8 + 8 (k - 1)
  const int first8 = S1 > 0 ? first : ST_TT;
  auto tile_range = [&](int k, int& s0, int& nt) {
    s0 = k == 0 ? 0 : first8 + (k - 1) * ST_TT;
    nt = min(k == 0 ? first8 : ST_TT, (k < n1t ? S1 : L) - s0);
  };

  // ---- producers.  Tile j's bulk store and the refill of its stage are the job of lane 0 of warp (j mod 4) of the
  // direction, so the TMA bookkeeping is spread evenly over the four warps (a single producer thread made its warp
  // the slowest of the group and the other three waited for it at every stage).
  auto issue_partial = [&](int k) {
    int s0, nt; tile_range(k, s0, nt);
    const int stage = k % NSTG;
    const int brow = rev ? (row0 + L - s0 - ST_TT) : (row0 + s0);
    sbar_expect_tx(pfull_bar(stage), SL::P_BYTES);
    tma_tile_2d(ring + (uint32_t)stage * SL::STAGE_BYTES + SL::OFF_P, &maps.o, blockIdx.x * CH, brow, pfull_bar(stage));
  };
  // with_partial: the CTA is past its phase barrier, so a finalising tile may fetch its parked partials right away
  auto issue_tile = [&](int k, bool with_partial) {
    int s0, nt; tile_range(k, s0, nt);
    const int stage = k % NSTG;
    const uint32_t st = ring + (uint32_t)stage * SL::STAGE_BYTES;
    const uint32_t bar = full_bar(stage);
    const bool fin = k >= n1t;
    // box rows: forward [row0+s0, +TT); reverse [row0+L-s0-TT, +TT) so that step t sits at box row TT-1-t
    const int brow = rev ? (row0 + L - s0 - ST_TT) : (row0 + s0);
    const uint32_t bc_bytes = (uint32_t)nt * SCAN_ROW * 4u;
    const uint32_t tx = SL::U_BYTES + SL::D_BYTES + ((fin && has_z) ? SL::Z_BYTES : 0) + bc_bytes;
    sbar_expect_tx(bar, tx);
    const int col = blockIdx.x * CH;
    tma_tile_2d(st + SL::OFF_U, &maps.u[g], col, brow, bar);
    tma_tile_2d(st + SL::OFF_D, &maps.d[g], col, brow, bar);
    if (fin && has_z) tma_tile_2d(st + SL::OFF_Z, &maps.z, col, brow, bar);
    // packed [B|C] rows: only the nt valid rows, placed so that step t reads row (rev ? TT-1-t : t)
    const int bc_row_lo = rev ? (L - s0 - nt) : s0;
    const float* src = reinterpret_cast<const float*>(d.Bm) + (int64_t)(row0 + bc_row_lo) * SCAN_ROW;
    const uint32_t dst = st + SL::OFF_BC + (rev ? (uint32_t)(ST_TT - nt) * SCAN_ROW * 4u : 0u);
    bulk_g2s(dst, src, bc_bytes, bar);
    if (bidir && fin && with_partial) issue_partial(k);
  };
  // full tiles leave through one bulk tensor store of the stage's (overwritten) u tile, once all four warps have
  // released the stage; short tiles were stored directly
  auto store_tile = [&](int j) {
    sbar_wait(empty_bar(j % NSTG), (uint32_t)((j / NSTG) & 1));
    int s0, nt; tile_range(j, s0, nt);
    if (nt == ST_TT) {
      const int brow = rev ? (row0 + L - s0 - ST_TT) : (row0 + s0);
      tma_store_tile_2d(&maps.o, ring + (uint32_t)(j % NSTG) * SL::STAGE_BYTES + SL::OFF_U, blockIdx.x * CH, brow);
      bulk_commit();
    }
  };
  const bool my_lane0 = lane == 0;
#if AUM_SCAN_ELECT
  auto owns = [&](int j) { return (j % NW) == warp_in_group && scan_elect(); };     // (warp-uniform test first)
#else
  auto owns = [&](int j) { return my_lane0 && (j % NW) == warp_in_group; };
#endif
  if (AUM_LEAD0(tig == 0)) {
    for (int k = 0; k < NSTG && k < ntiles; ++k) issue_tile(k, !bidir);
  }
  // highest tile issued before iteration k ends its producer step (see the loop tail): NSTG - 1 up front, then k - 2 + NSTG
  auto issued_before = [&](int k) { return min(ntiles - 1, max(NSTG - 1, k - 3 + NSTG)); };

  // ---- per-channel constants
  f32x2 a2[SCAN_NS / 2], h[SCAN_NS / 2];
  {
    const float4* ap = reinterpret_cast<const float4*>(d.A + (int64_t)ch * SCAN_NS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldg(ap + i);
      a2[2 * i] = pk2(v.x * 1.4426950408889634f, v.y * 1.4426950408889634f);
      a2[2 * i + 1] = pk2(v.z * 1.4426950408889634f, v.w * 1.4426950408889634f);
    }
#pragma unroll
    for (int k = 0; k < SCAN_NS / 2; ++k) h[k] = pk2(0.f, 0.f);
  }
  const float Dv = d.D ? __ldg(d.D + ch) : 0.f;
  const float oscale = p.out_scale;
  const int ldo = (int)p.ld_out;
  T* ob = reinterpret_cast<T*>(p.out) + ch;
  const ptrdiff_t ostep = rev ? -(ptrdiff_t)ldo : (ptrdiff_t)ldo;
  // the pre-gate output (if requested) shares out's row pitch: address it as an element offset from the out row
  const ptrdiff_t ypre_off = p.ypre ? (reinterpret_cast<T*>(p.ypre) - reinterpret_cast<T*>(p.out)) : 0;
  const int su = rev ? -(int)(CH * sizeof(T)) : (int)(CH * sizeof(T));
  const int sd = rev ? -(CH * (int)sizeof(TD)) : (CH * (int)sizeof(TD));
  const int sbc = rev ? -(SCAN_ROW * 4) : (SCAN_ROW * 4);

  float* ckp = (CKPT && d.ckpt) ? d.ckpt + (int64_t)b * scan_ck_count_max(L) * SCAN_NS * p.Dch + ch : nullptr;

  // Loop state kept incrementally (no division / modulo / re-derivation of the tiling per trip: the bookkeeping
  // between two tile bodies is pure latency for the warp, and with four warps per scheduler it showed up as 19 % of
  // all stall samples): walk position, ring stage, the parity bits of the stage barriers, k mod NW.
  const int m_store = (warp_in_group + 1) % NW, m_refill = (warp_in_group + 2) % NW;   // k mod NW at which this warp stores / refills
  const uint32_t o_u = SL::OFF_U + (uint32_t)tig * (uint32_t)sizeof(T), o_d = SL::OFF_D + (uint32_t)tig * (uint32_t)sizeof(TD);
  const uint32_t o_z = SL::OFF_Z + (uint32_t)tig * (uint32_t)sizeof(T), o_p = SL::OFF_P + (uint32_t)tig * (uint32_t)sizeof(T);
  const bool save_ypre = ypre_off != 0;
  int s0 = 0, stage = 0, m = 0;
  uint32_t fpar = 0, ppar = 0;                     // bit s: parity of the next wait on full / partial barrier of stage s
  // Tiles at the edges of the two phases (tile 0, the first finalising tile with the phase barrier in front of it, the
  // last tile) go through this general step; everything in between - all of them full 8-step tiles - through the
  // specialised loops below.
  auto generic_step = [&](const int k) {
    const bool fin = k >= n1t;
    const int nt = min(k == 0 ? first8 : ST_TT, (fin ? L : S1) - s0);
    const uint32_t st = ring + (uint32_t)stage * SL::STAGE_BYTES;
    const bool partial = fin && bidir;
    if (bidir && k == n1t) {
      // every partial of both directions must be parked before anyone fetches one: the producer drains its bulk
      // stores (full tiles), the generic-proxy stores of short tiles are made visible to the async proxy, then the
      // CTA meets and the producer catches up on the partial tiles of the stages already in flight
      if (n1t >= 1 && owns(n1t - 1)) store_tile(n1t - 1);
      if (AUM_LEAD(my_lane0)) bulk_wait_all<0>();
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncthreads();
      if (AUM_LEAD0(tig == 0)) {
        const int upto = issued_before(n1t);
        for (int kk = n1t; kk <= upto; ++kk) issue_partial(kk);
      }
    }
    if (CKPT && ckp != nullptr && active) {        // training: state before this tile (tiles == checkpoint chunks)
      float* c = ckp + (int64_t)k * SCAN_NS * p.Dch;
#pragma unroll
      for (int i = 0; i < SCAN_NS / 2; ++i) {
        float lo, hi; upk2(h[i], lo, hi);
        c[(int64_t)(2 * i) * p.Dch] = lo; c[(int64_t)(2 * i + 1) * p.Dch] = hi;
      }
    }
    sbar_wait(full_bar(stage), (fpar >> stage) & 1u);
    fpar ^= 1u << stage;
    if (partial) { sbar_wait(pfull_bar(stage), (ppar >> stage) & 1u); ppar ^= 1u << stage; }

    // this thread's element in tile row 0 of each operand tile of the stage
    const uint32_t b_u = st + o_u, b_d = st + o_d, b_z = st + o_z, b_bc = st + SL::OFF_BC, b_p = st + o_p;
    if (nt == ST_TT) {
      T* pyp = nullptr;
      if (fin && save_ypre) pyp = ob + (int64_t)(row0 + (rev ? (L - 1 - s0) : s0)) * ldo + ypre_off;
#define AUM_TILE(F, P, Z, R, Y) scan_tile_full<T, TD, CH, F, P, Z, R, Y>(b_u, b_d, b_z, b_bc, b_p, Dv, oscale, active, h, a2, pyp, ostep)
#define AUM_TILE_Z(F, P, R) do { if (zmode == 0) AUM_TILE(F, P, 0, R, false); else if (zmode == 2) AUM_TILE(F, P, 2, R, false); \
                                 else if (save_ypre) AUM_TILE(F, P, 1, R, true); else AUM_TILE(F, P, 1, R, false); } while (0)
      if (rev) {
        if (!fin) AUM_TILE(false, false, 0, true, false);
        else if (partial) AUM_TILE_Z(true, true, true);
        else AUM_TILE_Z(true, false, true);
      } else {
        if (!fin) AUM_TILE(false, false, 0, false, false);
        else if (partial) AUM_TILE_Z(true, true, false);
        else AUM_TILE_Z(true, false, false);
      }
#undef AUM_TILE_Z
#undef AUM_TILE
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // y tile (generic proxy) -> bulk store (async proxy)
    } else {
      const int row_first = rev ? (ST_TT - 1) : 0;
      T* po = ob + (int64_t)(row0 + (rev ? (L - 1 - s0) : s0)) * ldo;    // global row of step s0
      scan_tile_tail<T, TD>(nt, fin, partial, has_z, b_u + (uint32_t)(row_first * CH * (int)sizeof(T)),
                        b_d + (uint32_t)(row_first * CH * (int)sizeof(TD)), b_z + (uint32_t)(row_first * CH * (int)sizeof(T)),
                        b_bc + (uint32_t)(row_first * SCAN_ROW * 4), b_p + (uint32_t)(row_first * CH * (int)sizeof(T)),
                        su, sd, sbc, Dv, oscale, active, h, a2, po, ostep, ypre_off, zpre);
    }

    // hand the stage back.  The elected thread then (a) bulk-stores the tile finished one iteration ago, once all
    // four warps have released it, and (b) refills the stage of the tile before that, once its store has drained.
    __syncwarp();
    if (AUM_LEAD(my_lane0)) {
      sbar_arrive(empty_bar(stage));
      if (m == m_store && k >= 1 && !(bidir && k == n1t)) store_tile(k - 1);     // (tile n1t-1 left at the phase barrier)
      if (m == m_refill && k >= 2) {
        const int kk = k - 2 + NSTG;                  // tile that reuses the stage of tile k-2
        if (kk < ntiles) {
          bulk_wait_read<0>();                        // this thread's store of tile k-2 has left shared memory
          issue_tile(kk, k >= n1t);
        }
      }
    }
    s0 += nt;
    stage = (stage + 1 == NSTG) ? 0 : stage + 1;
    m = (m + 1 == NW) ? 0 : m + 1;
  };

  // Steady-state loop over full tiles [k, kend) of ONE phase: walk direction, phase and gate are compile-time, so the
  // per-tile code between two tile bodies is the barrier handshake and the owner's TMA duties and nothing else (the
  // general step spends ~100 instructions there on re-deriving which case it is in: 16 % of the kernel's stall samples).
  int k = 0;
  auto fast_range = [&](auto fin_c, auto part_c, auto zm_c, auto rev_c, auto ypre_c, const int kend) {
    constexpr bool FIN = decltype(fin_c)::value, PART = decltype(part_c)::value, REV = decltype(rev_c)::value;
    constexpr bool YPRE = decltype(ypre_c)::value;
    constexpr int ZM = decltype(zm_c)::value;
    for (; k < kend; ++k) {
      const uint32_t st = ring + (uint32_t)stage * SL::STAGE_BYTES;
      if (CKPT && ckp != nullptr && active) {        // training: state before this tile (tiles == checkpoint chunks)
        float* c = ckp + (int64_t)k * SCAN_NS * p.Dch;
#pragma unroll
        for (int i = 0; i < SCAN_NS / 2; ++i) {
          float lo, hi; upk2(h[i], lo, hi);
          c[(int64_t)(2 * i) * p.Dch] = lo; c[(int64_t)(2 * i + 1) * p.Dch] = hi;
        }
      }
      sbar_wait(full_bar(stage), (fpar >> stage) & 1u);
      fpar ^= 1u << stage;
      if (PART) { sbar_wait(pfull_bar(stage), (ppar >> stage) & 1u); ppar ^= 1u << stage; }
      T* pyp = nullptr;
      if (YPRE) pyp = ob + (int64_t)(row0 + (REV ? (L - 1 - s0) : s0)) * ldo + ypre_off;
      scan_tile_full<T, TD, CH, FIN, PART, ZM, REV, YPRE>(st + o_u, st + o_d, st + o_z, st + SL::OFF_BC, st + o_p, Dv, oscale,
                                                      active, h, a2, pyp, ostep);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // y tile (generic proxy) -> bulk store (async proxy)
      __syncwarp();
      if (AUM_LEAD(my_lane0)) {
        sbar_arrive(empty_bar(stage));
        if (m == m_store) store_tile(k - 1);
        if (m == m_refill && k >= 2) {
          const int kk = k - 2 + NSTG;
          if (kk < ntiles) {
            bulk_wait_read<0>();
            issue_tile(kk, FIN);
          }
        }
      }
      s0 += ST_TT;
      stage = (stage + 1 == NSTG) ? 0 : stage + 1;
      m = (m + 1 == NW) ? 0 : m + 1;
    }
  };
  using std::integral_constant;
  typedef integral_constant<bool, true> T_; typedef integral_constant<bool, false> F_;
  auto fast_fin = [&](auto part_c, auto rev_c, const int kend) {       // finalising phase: pick the gate variant once
    if (zmode == 0) fast_range(T_{}, part_c, integral_constant<int, 0>{}, rev_c, F_{}, kend);
    else if (zmode == 2) fast_range(T_{}, part_c, integral_constant<int, 2>{}, rev_c, F_{}, kend);
    else if (save_ypre) fast_range(T_{}, part_c, integral_constant<int, 1>{}, rev_c, T_{}, kend);
    else fast_range(T_{}, part_c, integral_constant<int, 1>{}, rev_c, F_{}, kend);
  };
  // general step, then the run of full tiles of the same phase that follows it: tile 0 | tiles 1 .. n1t-1 | tile n1t
  // (phase barrier in front) | tiles n1t+1 .. ntiles-2 | last tile.  (One call site per lambda: everything inlines and
  // the recurrence state stays in registers.)
  while (k < ntiles) {
    generic_step(k);
    ++k;
    const int kend = (k <= n1t) ? n1t : ntiles - 1;
    if (k < kend) {
      if (k < n1t) {
        if (rev) fast_range(F_{}, F_{}, integral_constant<int, 0>{}, T_{}, F_{}, kend);
        else     fast_range(F_{}, F_{}, integral_constant<int, 0>{}, F_{}, F_{}, kend);
      } else if (bidir) { if (rev) fast_fin(T_{}, T_{}, kend); else fast_fin(T_{}, F_{}, kend); }
      else              { if (rev) fast_fin(F_{}, T_{}, kend); else fast_fin(F_{}, F_{}, kend); }
    }
  }
  if (ntiles > 0 && owns(ntiles - 1)) store_tile(ntiles - 1);
  if (AUM_LEAD(my_lane0)) bulk_wait_all<0>();  // shared memory must outlive the bulk stores
  if (bidir && n2t == 0) {                  // (degenerate) keep the CTA barrier count equal across directions
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncthreads();
  }

  if (active && d.last_state != nullptr) {
#pragma unroll
    for (int k = 0; k < SCAN_NS / 2; ++k) {
      float lo, hi; upk2(h[k], lo, hi);
      d.last_state[((int64_t)b * p.Dch + ch) * SCAN_NS + 2 * k] = lo;
      d.last_state[((int64_t)b * p.Dch + ch) * SCAN_NS + 2 * k + 1] = hi;
    }
  }
}

template <typename T, typename TD, int NSTG, int CH>
static int launch_n(const ScanTmaMaps& maps, const ScanParams& p, cudaStream_t st) {
  using SL = StageLayout<T, TD, NSTG, CH>;
  constexpr int MINB = CH <= 128 ? 2 : 1;
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.cur();
  // CH = 128: two resident CTAs per SM (128 registers, 90-110 KB of stages each); a 3-CTA / 80-register build measured slower
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(scan_fwd_tma_kernel<T, TD, MINB, NSTG, CH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(scan_fwd_tma_kernel<T, TD, MINB, NSTG, CH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("aum_selective_scan_fwd: cudaFuncSetAttribute(smem=%d): %s", SL::SMEM_BYTES, cudaGetErrorString(e)); return 2; }
    attr_set = true;
  }
  dim3 grid(ceil_div(p.Dch, CH), p.batch);
  const int smem = p.ndirs * SL::GROUP_BYTES + 128 + 2 * 3 * NSTG * 8;     // one ring per direction actually launched
  const bool ckpt = p.dir[0].ckpt != nullptr || (p.ndirs == 2 && p.dir[1].ckpt != nullptr);
  if (ckpt) scan_fwd_tma_kernel<T, TD, MINB, NSTG, CH, true><<<grid, CH * p.ndirs, smem, st>>>(maps, p);
  else      scan_fwd_tma_kernel<T, TD, MINB, NSTG, CH, false><<<grid, CH * p.ndirs, smem, st>>>(maps, p);
  return check_launch("aum_selective_scan_fwd(tma)");
}

// AUM_SCAN_TMA_CH=192 selects the one-CTA-per-SM build (see the note on CH above; ring depth 3 there: 98 KB).  The
// production build is CH = 128 with a 4-deep ring (3 and 5 were measured: 5 is 1-3 % slower, 3 starves the walk).
int scan_tma_ch() {
  static int ch = 0;
  if (ch == 0) { const char* e = getenv("AUM_SCAN_TMA_CH"); ch = (e && atoi(e) == 192) ? 192 : 128; }
  return ch;
}

template <typename T, typename TD>
static int launch_t(const ScanTmaMaps& maps, const ScanParams& p, cudaStream_t st) {
  if (scan_tma_ch() == 192) {
    if constexpr (std::is_same<TD, float>::value) return launch_n<T, float, 3, 192>(maps, p, st);
    else return -1;                                      // the 192-channel experiment build takes fp32 delta only
  }
  return launch_n<T, TD, 4, 128>(maps, p, st);
}

int launch_scan_tma(const ScanParams& p, int dtype, int delta_dt, cudaStream_t st) {
  if (p.N != SCAN_NS || (delta_dt != AUM_F32 && delta_dt != dtype) || !tma_available()) return -1;
  const int dsz = dtype_size(delta_dt);
  const int esz = dtype_size(dtype);
  auto ok_mat = [](const void* base, int64_t ld, int sz) { return aligned16(base) && (ld * sz) % 16 == 0; };
  if (!ok_mat(p.out, p.ld_out, esz) || p.ld_out * esz % 2 != 0) return -1;
  if (p.ypre && (p.ld_ypre != p.ld_out || p.ypre == p.out)) return -1;   // pre-gate rows must mirror the out rows
  if (p.ypre && (p.z == nullptr || p.z_pregated)) return -1;              // pre-gate output is specialised for y*silu(z)
  if (p.z && !ok_mat(p.z, p.ld_z, esz)) return -1;
  for (int g = 0; g < p.ndirs; ++g) {
    const ScanDirDev& d = p.dir[g];
    if (!d.bc_packed || d.delta_softplus || d.delta_bias != nullptr) return -1;
    if (!ok_mat(d.u, d.ld_u, esz) || !ok_mat(d.delta, d.ld_delta, dsz) || !aligned16(d.A)) return -1;
  }
  ScanTmaMaps maps;
  const int CH = scan_tma_ch();
  const int64_t rows = (int64_t)p.batch * p.L;
  for (int g = 0; g < 2; ++g) {
    const ScanDirDev& d = p.dir[g < p.ndirs ? g : 0];
    if (int rc = tma_encode_2d(&maps.u[g], d.u, dtype, rows, p.Dch, d.ld_u, ST_TT, CH, false, "aum_selective_scan_fwd(u)")) return rc;
    if (int rc = tma_encode_2d(&maps.d[g], d.delta, delta_dt, rows, p.Dch, d.ld_delta, ST_TT, CH, false, "aum_selective_scan_fwd(delta)")) return rc;
  }
  if (p.z) { if (int rc = tma_encode_2d(&maps.z, p.z, dtype, rows, p.Dch, p.ld_z, ST_TT, CH, false, "aum_selective_scan_fwd(z)")) return rc; }
  else maps.z = maps.u[0];
  if (int rc = tma_encode_2d(&maps.o, p.out, dtype, rows, p.Dch, p.ld_out, ST_TT, CH, false, "aum_selective_scan_fwd(out)")) return rc;
  switch (dtype) {
    case AUM_F32:  return launch_t<float, float>(maps, p, st);
    case AUM_F16:  return delta_dt == AUM_F32 ? launch_t<__half, float>(maps, p, st) : launch_t<__half, __half>(maps, p, st);
    case AUM_BF16: return delta_dt == AUM_F32 ? launch_t<__nv_bfloat16, float>(maps, p, st) : launch_t<__nv_bfloat16, __nv_bfloat16>(maps, p, st);
  }
  return -1;
}

}  // namespace aum
