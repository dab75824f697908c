// Selective scan forward, one or BOTH time directions in a single launch, token-major activations.
//
// Replaces selective_scan_cuda.fwd (call sites /root/reference/vim-mamba_ssm/mamba_ssm/ops/
// selective_scan_interface.py:37,213,354,499) and, for Fo-Bi / Bi-Bi, the second launch on five flip(-1)
// copies plus the un-flip and add (:503-507): here the reverse direction simply walks the tokens
// downwards, so no flipped copy of u / delta / B / C / z is ever materialised.
// Math per direction (selective_scan_ref, :86-152):
//     delta' = softplus(delta + delta_bias);  h_l = exp(delta'_l * A) h_prev + delta'_l B_l u_l;
//     y_l = <C_l, h_l> + D u_l;   out = out_scale * (y_fwd + y_bwd) * silu(z).
//
// Mapping.  grid = (ceil(D/CH), batch).  A CTA owns CH channels of one sequence for all L tokens; one thread
// per (channel, direction) keeps the 16 recurrences of its channel in registers as 8 packed fp32x2 pairs and
// advances them with FMUL2/FFMA2 (two states per instruction) and ex2.approx on A pre-scaled by log2(e).
// Token-major layout makes every per-token access of a warp one contiguous 64/128-byte segment; the next
// 4 tokens' u/delta/z are prefetched into registers while the current 4 are computed (ping-pong, no copies).
// B_l/C_l (shared by all channels of a token) stream through shared memory in 64-token chunks:
// cp.async.bulk (TMA 1-D) + mbarrier, double-buffered per direction, read back as broadcast 128-bit loads.
// Two directions, one pass: the forward threads walk l = 0..L-1 while the backward threads walk l = L-1..0.
// In the first half of its walk a direction parks its partial y (+D u) in `out`; after one CTA barrier at the
// midpoint each direction meets rows the other one has already visited, adds the parked partial, applies
// the SiLU(z) gate and writes the final value.  Extra traffic: one write + one (mostly L2) read of `out`.
//
// Roofline note: 16 ex2 per (token, channel, direction) make this kernel MUFU-bound
// (16 results/clk/SM) well before it is HBM-bound.
#include <stdlib.h>

#include "scan_common.cuh"

namespace aum {

// ---- chunk walk: a 4-deep rotating register window of per-channel stream values -----------------------
// Slot i holds the values of step t (t % 4 == i); right after step t is consumed the slot is refilled with
// step t+4, so every global load has three full steps of math in front of it and only 4 slots are live.
// Chunks are walked in unguarded groups of 4 steps: steps past the end of a chunk ("phantom" steps) get the
// neutral inputs delta' = 0, u = 0 (=> exp2(0) = 1, h unchanged) and read zero/stale-but-finite B/C rows from
// the padded staging buffer; only their store is predicated off.
struct Slots { float u[SCAN_U], d[SCAN_U], z[SCAN_U], p[SCAN_U]; };

template <typename T, typename TD, bool FINAL, bool PARTIAL, bool HASZ, bool SP>
__device__ __forceinline__ void scan_chunk(int ns, const float* bc_row, int row_step,
                                           f32x2 (&h)[SCAN_NS / 2], const f32x2 (&a2)[SCAN_NS / 2],
                                           float Dv, float dbias, float oscale,
                                           const T* ub, int ldu, const TD* db, int ldd,
                                           const T* zb, int ldz, T* ob, int ldo, T* ypb, int ldy, int r, int dr,
                                           float* ck, int ck_stride, int s_glob, int ck_first, bool zpre) {
  Slots w;
  // r: row of the step being computed; rl: row of the step being loaded (4 steps ahead)
  int rl = r;
  auto fill = [&](int i, int t) {
    if (t < ns) {
      w.u[i] = to_f(ub[(int64_t)rl * ldu]);
      w.d[i] = to_f(db[(int64_t)rl * ldd]);
      if (FINAL && HASZ) w.z[i] = to_f(zb[(int64_t)rl * ldz]);
      if (PARTIAL) w.p[i] = to_f(ob[(int64_t)rl * ldo]);
    } else {
      w.u[i] = 0.f;
      w.d[i] = -dbias;          // (-dbias) + dbias == 0 exactly
    }
    rl += dr;
  };
#pragma unroll
  for (int i = 0; i < SCAN_U; ++i) fill(i, i);

  for (int t0 = 0; t0 < ns; t0 += SCAN_U) {
#pragma unroll
    for (int i = 0; i < SCAN_U; ++i) {
      const int t = t0 + i;
      if (ck != nullptr && t < ns) {        // training: state BEFORE this step at checkpoint-chunk starts
        const int s = s_glob + t;
        if (s == 0 || (s >= ck_first && ((s - ck_first) % SCAN_CK) == 0)) {
          float* c = ck + (int64_t)(s == 0 ? 0 : 1 + (s - ck_first) / SCAN_CK) * SCAN_NS * ck_stride;
#pragma unroll
          for (int k = 0; k < SCAN_NS / 2; ++k) {
            float lo, hi; upk2(h[k], lo, hi);
            c[(int64_t)(2 * k) * ck_stride] = lo; c[(int64_t)(2 * k + 1) * ck_stride] = hi;
          }
        }
      }
      float dl = w.d[i] + dbias;
      if (SP) dl = (t < ns) ? softplus_f(dl) : 0.f;
      const float u = w.u[i];
      const float du = dl * u;
      const f32x2 dl2 = pk2(dl, dl), du2 = pk2(du, du);
      f32x2 ya = pk2(Dv * u, 0.f), yb = pk2(0.f, 0.f);
      const float4* row = reinterpret_cast<const float4*>(bc_row);
#pragma unroll
      for (int q = 0; q < SCAN_NS / 4; ++q) {
        const float4 Bv = row[q];
        const float4 Cv = row[SCAN_NS / 4 + q];
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
      float y = (y0 + y1) + (y2 + y3);
      if (FINAL) {
        if (PARTIAL) y += w.p[i];
        if (ypb != nullptr && t < ns) ypb[(int64_t)r * ldy] = from_f<T>(y);
        if (HASZ) y *= zpre ? w.z[i] : silu_f(w.z[i]);
        y *= oscale;
      }
      if (t < ns) ob[(int64_t)r * ldo] = from_f<T>(y);
      r += dr;
      bc_row += row_step;
      fill(i, t + SCAN_U);
    }
  }
}

constexpr int SCAN_PAD = 4;                       // zero rows before/after each staged chunk (phantom steps)
constexpr int SCAN_ROWS = SCAN_TC + 2 * SCAN_PAD;

template <typename T, typename TD, int CH, bool SP>
__global__ void __launch_bounds__(2 * CH, (CH == 64) ? 4 : 2)
scan_fwd_kernel(const ScanParams p) {
  __shared__ __align__(128) float bc_smem[2][2][SCAN_ROWS][SCAN_ROW];   // [direction group][stage][row][B|C]
  __shared__ __align__(8) unsigned long long mbar[2][2];

  const int g = threadIdx.x / CH;              // direction slot of this thread
  const int tig = threadIdx.x - g * CH;        // thread index within the direction group
  const ScanDirDev& d = p.dir[g];
  // lanes past the last channel shadow channel Dch-1: same inputs, same results, same (benign) stores
  const int ch = min(blockIdx.x * CH + tig, p.Dch - 1);
  const int b = blockIdx.y;
  const int L = p.L, N = p.N;
  const bool bidir = p.ndirs == 2;
  const bool rev = d.reverse != 0;
  const int row0 = b * L;
  const bool packed = d.bc_packed != 0;

  {  // zero the staging buffers once (pad rows and not-yet-filled rows must hold finite values)
    float4* z4 = reinterpret_cast<float4*>(&bc_smem[0][0][0][0]);
    const int n4 = 2 * 2 * SCAN_ROWS * SCAN_ROW / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tig == 0) {
    sbar_init(s_u32(&mbar[g][0]), 1);
    sbar_init(s_u32(&mbar[g][1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- per-channel constants
  f32x2 a2[SCAN_NS / 2], h[SCAN_NS / 2];
  {
    float a[SCAN_NS];
#pragma unroll
    for (int n = 0; n < SCAN_NS; ++n) a[n] = 0.f;
    if (N == SCAN_NS) {
      const float4* ap = reinterpret_cast<const float4*>(d.A + (int64_t)ch * SCAN_NS);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(ap + i);
        a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int n = 0; n < SCAN_NS; ++n)
        if (n < N) a[n] = __ldg(d.A + (int64_t)ch * N + n);
    }
#pragma unroll
    for (int k = 0; k < SCAN_NS / 2; ++k) {
      a2[k] = pk2(a[2 * k] * 1.4426950408889634f, a[2 * k + 1] * 1.4426950408889634f);
      h[k] = pk2(0.f, 0.f);
    }
  }
  const float Dv = d.D ? __ldg(d.D + ch) : 0.f;
  const float dbias = d.delta_bias ? __ldg(d.delta_bias + ch) : 0.f;
  const float oscale = p.out_scale;
  const bool has_z = p.z != nullptr;

  // phase 1: steps [0, S1) park partials; phase 2: steps [S1, L) finalise.  Unidirectional: S1 = 0.
  const int mid = L / 2;
  const int S1 = bidir ? (rev ? (L - mid) : mid) : 0;
  const int n1 = (S1 + SCAN_TC - 1) / SCAN_TC;
  const int n2 = (L - S1 + SCAN_TC - 1) / SCAN_TC;
  const int nchunks = n1 + n2;

  auto chunk_range = [&](int k, int& s0, int& ns) {
    if (k < n1) { s0 = k * SCAN_TC; ns = min(SCAN_TC, S1 - s0); }
    else { s0 = S1 + (k - n1) * SCAN_TC; ns = min(SCAN_TC, L - s0); }
  };
  // lowest token row of a chunk (chunks are staged in MEMORY order; the reverse walk reads them backwards)
  auto chunk_lo = [&](int s0, int ns) { return rev ? (L - s0 - ns) : s0; };

  auto issue_bulk = [&](int k) {      // one thread: TMA the chunk's [ns x 32] fp32 rows into stage k&1
    int s0, ns; chunk_range(k, s0, ns);
    const uint32_t bar = s_u32(&mbar[g][k & 1]);
    const uint32_t bytes = (uint32_t)ns * SCAN_ROW * 4u;
    const float* src = reinterpret_cast<const float*>(d.Bm) + (int64_t)(row0 + chunk_lo(s0, ns)) * SCAN_ROW;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    sbar_expect_tx(bar, bytes);
    bulk_g2s(s_u32(&bc_smem[g][k & 1][SCAN_PAD][0]), src, bytes, bar);
  };
  if (packed && tig == 0) {
    if (nchunks > 0) issue_bulk(0);
    if (nchunks > 1) issue_bulk(1);
  }

  const int ldu = (int)d.ld_u, ldd = (int)d.ld_delta, ldz = (int)p.ld_z, ldo = (int)p.ld_out;
  const int dr = rev ? -1 : 1;
  const T* ub = reinterpret_cast<const T*>(d.u) + ch;
  const TD* db = reinterpret_cast<const TD*>(d.delta) + ch;
  const T* zb = reinterpret_cast<const T*>(p.z) + ch;
  T* ob = reinterpret_cast<T*>(p.out) + ch;
  T* ypb = p.ypre ? reinterpret_cast<T*>(p.ypre) + ch : nullptr;
  const int ldy = (int)p.ld_ypre;
  const int ck_first = scan_ck_first(L, bidir, rev);
  // (lanes shadowing channel Dch-1 write the same values to the same checkpoint slots: benign)
  float* ckp = d.ckpt ? d.ckpt + (int64_t)b * scan_ck_count_max(L) * SCAN_NS * p.Dch + ch : nullptr;

  for (int k = 0; k < nchunks; ++k) {
    int s0, ns; chunk_range(k, s0, ns);
    const int stage = packed ? (k & 1) : 0;
    float (*bc)[SCAN_ROW] = bc_smem[g][stage] + SCAN_PAD;
    if (packed) {
      sbar_wait(s_u32(&mbar[g][stage]), (uint32_t)((k >> 1) & 1));
    } else {
      // generic staging (any dtype / strides): cooperative, synchronous, memory order
      group_barrier(1 + g, CH);
      const int lo = row0 + chunk_lo(s0, ns);
      const int total = ns * SCAN_ROW;
      for (int i0 = tig; i0 < total; i0 += 4 * CH) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = i0 + i * CH;
          v[i] = 0.f;
          if (idx < total) {
            const int m = idx / SCAN_ROW, j = idx % SCAN_ROW;
            if (j < SCAN_NS) { if (j < N) v[i] = load_as_f(d.Bm, (int64_t)(lo + m) * d.ld_B + j, d.bc_dt); }
            else { const int jj = j - SCAN_NS; if (jj < N) v[i] = load_as_f(d.Cm, (int64_t)(lo + m) * d.ld_C + jj, d.bc_dt); }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = i0 + i * CH;
          if (idx < total) bc[idx / SCAN_ROW][idx % SCAN_ROW] = v[i];
        }
      }
      group_barrier(1 + g, CH);
    }
    if (bidir && k == n1) __syncthreads();      // every partial of both directions is parked

    const int r = row0 + (rev ? (L - 1 - s0) : s0);
    const float* bc_row0 = &bc[rev ? (ns - 1) : 0][0];
    const int row_step = rev ? -SCAN_ROW : SCAN_ROW;
    const bool finalize = k >= n1;
#define AUM_SCAN_CHUNK(F, P, Z) scan_chunk<T, TD, F, P, Z, SP>(ns, bc_row0, row_step, h, a2, Dv, dbias, oscale, ub, ldu, db, ldd, zb, ldz, ob, ldo, ypb, ldy, r, dr, ckp, p.Dch, s0, ck_first, p.z_pregated != 0)
    if (!finalize) AUM_SCAN_CHUNK(false, false, false);
    else if (bidir) { if (has_z) AUM_SCAN_CHUNK(true, true, true); else AUM_SCAN_CHUNK(true, true, false); }
    else            { if (has_z) AUM_SCAN_CHUNK(true, false, true); else AUM_SCAN_CHUNK(true, false, false); }
#undef AUM_SCAN_CHUNK
    if (packed) {
      group_barrier(1 + g, CH);                 // every thread of the group is done with this stage
      if (tig == 0 && k + 2 < nchunks) issue_bulk(k + 2);
    }
  }
  if (bidir && n2 == 0) __syncthreads();        // (degenerate: no phase-2 chunk in this group) keep barrier counts equal

  if (d.last_state != nullptr) {
#pragma unroll
    for (int k = 0; k < SCAN_NS / 2; ++k) {
      float lo, hi; upk2(h[k], lo, hi);
      if (2 * k < N) d.last_state[((int64_t)b * p.Dch + ch) * N + 2 * k] = lo;
      if (2 * k + 1 < N) d.last_state[((int64_t)b * p.Dch + ch) * N + 2 * k + 1] = hi;
    }
  }
}

template <typename T, typename TD>
static int launch_scan_td(const ScanParams& p, int ch, bool sp, cudaStream_t st) {
  dim3 grid(ceil_div(p.Dch, ch), p.batch);
  if (ch == 64) {
    if (sp) scan_fwd_kernel<T, TD, 64, true><<<grid, 64 * p.ndirs, 0, st>>>(p);
    else    scan_fwd_kernel<T, TD, 64, false><<<grid, 64 * p.ndirs, 0, st>>>(p);
  } else {
    if (sp) scan_fwd_kernel<T, TD, 128, true><<<grid, 128 * p.ndirs, 0, st>>>(p);
    else    scan_fwd_kernel<T, TD, 128, false><<<grid, 128 * p.ndirs, 0, st>>>(p);
  }
  return check_launch("aum_selective_scan_fwd");
}

template <typename T>
static int launch_scan_t(const ScanParams& p, int delta_dt, int dtype, int ch, bool sp, cudaStream_t st) {
  if (delta_dt == dtype) return launch_scan_td<T, T>(p, ch, sp, st);
  if (delta_dt == AUM_F32) return launch_scan_td<T, float>(p, ch, sp, st);
  set_error("aum_selective_scan_fwd: delta dtype must equal the activation dtype or be fp32");
  return 1;
}

}  // namespace aum

extern "C" int aum_selective_scan_fwd(const aum_scan_dir_t* fwd, const aum_scan_dir_t* bwd,
                                      const void* z, int64_t ld_z, void* out, int64_t ld_out,
                                      int batch, int L, int D, int N, int dtype,
                                      float out_scale, void* y_pre, int64_t ld_ypre, int flags, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(out);
  AUM_REQUIRE(fwd || bwd, "aum_selective_scan_fwd: at least one direction is required");
  AUM_REQUIRE(out, "aum_selective_scan_fwd: null output");
  AUM_REQUIRE(N >= 1 && N <= SCAN_NS, "aum_selective_scan_fwd: d_state %d unsupported (1..%d)", N, SCAN_NS);
  AUM_REQUIRE(batch >= 0 && L >= 0 && D >= 0, "aum_selective_scan_fwd: negative size");
  AUM_REQUIRE(batch <= 65535, "aum_selective_scan_fwd: batch too large");
  AUM_REQUIRE(dtype >= AUM_F32 && dtype <= AUM_BF16, "aum_selective_scan_fwd: bad dtype %d", dtype);
  if (batch == 0 || L == 0 || D == 0) return 0;
  ScanParams p;
  memset(&p, 0, sizeof(p));
  int delta_dt = -1, sp_flag = -1;
  const aum_scan_dir_t* src[2] = {fwd, bwd};
  for (int i = 0; i < 2; ++i) {
    const aum_scan_dir_t* s = src[i];
    if (!s) continue;
    AUM_REQUIRE(s->u && s->delta && s->A && s->Bm && s->Cm, "aum_selective_scan_fwd: null pointer in direction %d", i);
    AUM_REQUIRE(s->ld_u >= D && s->ld_delta >= D && s->ld_B >= N && s->ld_C >= N, "aum_selective_scan_fwd: leading dimension too small");
    AUM_REQUIRE(s->bc_dtype >= AUM_F32 && s->bc_dtype <= AUM_BF16, "aum_selective_scan_fwd: bad bc_dtype");
    AUM_REQUIRE(delta_dt < 0 || delta_dt == s->delta_dtype, "aum_selective_scan_fwd: both directions must share delta_dtype");
    AUM_REQUIRE(N != SCAN_NS || aligned16(s->A), "aum_selective_scan_fwd: A must be 16-byte aligned");
    AUM_REQUIRE(sp_flag < 0 || sp_flag == (s->delta_softplus != 0), "aum_selective_scan_fwd: both directions must share delta_softplus");
    delta_dt = s->delta_dtype;
    sp_flag = s->delta_softplus != 0;
    ScanDirDev& d = p.dir[p.ndirs++];
    d.u = s->u; d.ld_u = s->ld_u; d.delta = s->delta; d.ld_delta = s->ld_delta; d.A = s->A;
    d.Bm = s->Bm; d.ld_B = s->ld_B; d.Cm = s->Cm; d.ld_C = s->ld_C; d.bc_dt = s->bc_dtype;
    d.bc_packed = (N == SCAN_NS && s->bc_dtype == AUM_F32 && s->ld_B == SCAN_ROW && s->ld_C == SCAN_ROW &&
                   reinterpret_cast<const float*>(s->Cm) == reinterpret_cast<const float*>(s->Bm) + SCAN_NS &&
                   aligned16(s->Bm)) ? 1 : 0;
    d.D = s->D; d.delta_bias = s->delta_bias; d.delta_softplus = s->delta_softplus;
    d.last_state = s->last_state; d.reverse = i;
    d.ckpt = s->ckpt;
    AUM_REQUIRE(!s->ckpt || N == SCAN_NS, "aum_selective_scan_fwd: checkpoints need d_state == 16");
  }
  p.z = z; p.ld_z = ld_z; p.out = out; p.ld_out = ld_out;
  p.ypre = y_pre; p.ld_ypre = ld_ypre;
  p.z_pregated = (flags & AUM_SCAN_Z_PREGATED) ? 1 : 0;
  AUM_REQUIRE(!y_pre || ld_ypre >= D, "aum_selective_scan_fwd: ld_ypre too small");
  p.batch = batch; p.L = L; p.Dch = D; p.N = N; p.out_scale = out_scale;
  AUM_REQUIRE(ld_out >= D && (!z || ld_z >= D), "aum_selective_scan_fwd: leading dimension too small");

  AUM_REQUIRE((int64_t)batch * L < (1ll << 31), "aum_selective_scan_fwd: batch*L must fit in int32");
  const bool sp = sp_flag > 0;
  int ch = 64;
  if (const char* e = getenv("AUM_SCAN_CH")) { int v = atoi(e); if (v == 64 || v == 128) ch = v; }
  cudaStream_t st = (cudaStream_t)stream;
  if (getenv("AUM_SCAN_GENERIC") == nullptr) {
    const int rc = launch_scan_tma(p, dtype, delta_dt, st);
    if (rc >= 0) return rc;            // -1: not eligible for the TMA-streamed kernel
  }
  switch (dtype) {
    case AUM_F32:  return launch_scan_t<float>(p, delta_dt, dtype, ch, sp, st);
    case AUM_F16:  return launch_scan_t<__half>(p, delta_dt, dtype, ch, sp, st);
    case AUM_BF16: return launch_scan_t<__nv_bfloat16>(p, delta_dt, dtype, ch, sp, st);
  }
  return 1;
}
