// Selective scan backward, one or both time directions in one launch, token-major.
//
// Replaces selective_scan_cuda.bwd (call sites /root/reference/vim-mamba_ssm/mamba_ssm/ops/
// selective_scan_interface.py:62,247,389,541,548) and, for Fo-Bi, the second call on seven flip(-1) copies plus
// the un-flips and adds of :554-561.  Gradients are those of selective_scan_ref (:86-152) under autograd; in
// particular d/dz includes BOTH directions (the shipped BiMambaInnerFn.backward drops the reverse direction's
// contribution, :560 after :537-538 — SURVEY.md Q2; we implement the mathematically correct gradient).
//
// Per direction (time order s = 0..L-1, token l = reverse ? L-1-s : s), with a_s = exp(delta_s A), b_s = delta_s B_s u_s:
//   forward:   h_s = a_s h_{s-1} + b_s,  y_s = <C_s, h_s> + D u_s,  out = scale * (y_f + y_b) * silu(z)
//   backward:  dy_s = scale * dout_s * silu(z_s);   dh_s = C_s dy_s + a_{s+1} dh_{s+1}
//              dC_s += dy_s h_s;  dB_s += dh_s delta_s u_s;  du_s = delta_s <dh_s, B_s> + D dy_s;  dD += dy_s u_s
//              ddelta_s = sum_n dh_s[n] (h_{s-1}[n] a_s[n] A[n] + B_s[n] u_s);  dA[n] += dh_s[n] h_{s-1}[n] a_s[n] delta_s
//   gate:      dz = scale * dout * y_pre * silu'(z),   out_z = scale * y_pre * silu(z)   (y_pre = y_f + y_b saved by fwd)
//
// Mapping: grid = (ceil(D/64), batch); a CTA owns 64 channels of one sequence; 64 threads per direction, one thread
// per channel with the 16 states in registers.  Sweep 1 replays the forward recurrence and checkpoints h every
// 8 steps to a caller-provided workspace; sweep 2 walks the 8-step chunks backwards: it recomputes the chunk's
// states from the checkpoint into shared memory, then runs the reverse-time recurrence over the chunk.
// dB/dC are sums over channels: a 31-shuffle butterfly leaves lane i with value i of the warp's 32 channels,
// one red.global.add.f32 per lane and token.  du / ddelta of the two directions (shared u, delta in Fo-Bi) meet
// through the same park-in-the-output trick as the forward kernel (one CTA barrier at the midpoint).
// This generic kernel (plain global loads) is the fallback of the TMA-streamed kernel in scan_bwd_tma.cu; it also
// regenerates the checkpoints when the forward pass did not leave them (ckpt_valid == 0).
#include "scan_bwd_common.cuh"

namespace aum {

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- per-step pieces (packed fp32x2: two states per instruction) --------------------------------------
__device__ __forceinline__ void bwd_recompute_step(float u, float dl, const float4* __restrict__ bq,
                                                   f32x2 (&h)[SCAN_NS / 2], const f32x2 (&a2)[SCAN_NS / 2],
                                                   float* hist_next /* [n][SB_CH] column of this thread */) {
  const float du_ = dl * u;
  const f32x2 dl2 = pk2(dl, dl), du2 = pk2(du_, du_);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 Bv = __ldg(bq + q);
    const f32x2 x0 = mul2(dl2, a2[2 * q]), x1 = mul2(dl2, a2[2 * q + 1]);
    float e0, e1, e2, e3;
    upk2(x0, e0, e1); upk2(x1, e2, e3);
    h[2 * q] = fma2(pk2(ex2_approx(e0), ex2_approx(e1)), h[2 * q], mul2(du2, pk2(Bv.x, Bv.y)));
    h[2 * q + 1] = fma2(pk2(ex2_approx(e2), ex2_approx(e3)), h[2 * q + 1], mul2(du2, pk2(Bv.z, Bv.w)));
    float h0, h1, h2, h3;
    upk2(h[2 * q], h0, h1); upk2(h[2 * q + 1], h2, h3);
    hist_next[(4 * q + 0) * SB_CH] = h0; hist_next[(4 * q + 1) * SB_CH] = h1;
    hist_next[(4 * q + 2) * SB_CH] = h2; hist_next[(4 * q + 3) * SB_CH] = h3;
  }
}

template <typename T>
__global__ void __launch_bounds__(2 * SB_CH, 3)
scan_bwd_kernel(const ScanBwdParams p) {
  // state history of the current chunk: hist[j] = state BEFORE step j of the chunk, hist[j+1] = state after it
  extern __shared__ float hist_raw[];
  float (*hist)[SB_TT + 1][SCAN_NS][SB_CH] = reinterpret_cast<float (*)[SB_TT + 1][SCAN_NS][SB_CH]>(hist_raw);
  const int64_t rows_total = (int64_t)p.batch * p.L;

  const int g = threadIdx.x / SB_CH;
  const int tig = threadIdx.x - g * SB_CH;
  const int lane = tig & 31;
  const ScanBwdDirDev& d = p.dir[g];
  const int ch_raw = blockIdx.x * SB_CH + tig;
  const bool active = ch_raw < p.Dch;
  const int ch = active ? ch_raw : (p.Dch - 1);
  const int b = blockIdx.y;
  const int L = p.L;
  const bool rev = d.reverse != 0;
  const int64_t row0 = (int64_t)b * L;
  const bool bidir = p.ndirs == 2;
  const bool shared = bidir && p.shared_du;
  const float scale = p.scale;

  f32x2 a2[SCAN_NS / 2], Av2[SCAN_NS / 2];
  {
    const float4* ap = reinterpret_cast<const float4*>(d.A + (int64_t)ch * SCAN_NS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldg(ap + i);
      Av2[2 * i] = pk2(v.x, v.y); Av2[2 * i + 1] = pk2(v.z, v.w);
      a2[2 * i] = pk2(v.x * 1.4426950408889634f, v.y * 1.4426950408889634f);
      a2[2 * i + 1] = pk2(v.z * 1.4426950408889634f, v.w * 1.4426950408889634f);
    }
  }
  const float Dv = d.D ? __ldg(d.D + ch) : 0.f;
  const T* ub = reinterpret_cast<const T*>(d.u) + ch;
  const float* db = reinterpret_cast<const float*>(d.delta) + ch;      // (fp32 only here: the entry point refuses a 16-bit delta)
  float* ck = d.ckpt + ((int64_t)b * scan_ck_count_max(L)) * SCAN_NS * p.Dch + ch;    // [b][chunk][n][D]
  auto token = [&](int s) { return rev ? (L - 1 - s) : s; };

  // checkpoint chunking (identical to the forward kernels'): chunk 0 = [0, first), chunk c = [first + 8(c-1), +8)
  const int first = min(scan_ck_first(L, bidir, rev), L);
  const int nchunks = 1 + (L - first + SB_TT - 1) / SB_TT;
  auto chunk_range = [&](int c, int& s0, int& ns) {
    if (c == 0) { s0 = 0; ns = first; } else { s0 = first + (c - 1) * SB_TT; ns = min(SB_TT, L - s0); }
  };
  float* hcol = &hist[g][0][0][tig];                 // this thread's column; [j][n] at hcol[(j*16 + n) * SB_CH]
  const int part = blockIdx.x * SB_WARPS + (tig >> 5);   // this warp's slice of the dB|dC partial workspace

  // ---------------- sweep 1 (only when the forward did not leave checkpoints) ----------------
  if (!d.ckpt_valid) {
    f32x2 h[SCAN_NS / 2];
#pragma unroll
    for (int k = 0; k < SCAN_NS / 2; ++k) h[k] = pk2(0.f, 0.f);
    for (int c = 0; c < nchunks; ++c) {
      int s0, ns; chunk_range(c, s0, ns);
      if (active) {
        float* cp = ck + (int64_t)c * SCAN_NS * p.Dch;
#pragma unroll
        for (int k = 0; k < SCAN_NS / 2; ++k) {
          float lo, hi; upk2(h[k], lo, hi);
          cp[(int64_t)(2 * k) * p.Dch] = lo; cp[(int64_t)(2 * k + 1) * p.Dch] = hi;
        }
      }
      for (int j = 0; j < ns; ++j) {
        const int64_t r = row0 + token(s0 + j);
        bwd_recompute_step(to_f(ub[r * d.ld_u]), db[r * d.ld_delta], reinterpret_cast<const float4*>(d.BC + r * d.ld_bc),
                           h, a2, hcol + SCAN_NS * SB_CH);      // (history slot 1 used as scratch here)
      }
    }
  }
  // the checkpoints are re-read by the same thread that wrote them (or were written by an earlier kernel)

  // ---------------- sweep 2: chunks in reverse time order ----------------
  const T* zb = p.z ? reinterpret_cast<const T*>(p.z) + ch : nullptr;
  const T* yb = p.ypre ? reinterpret_cast<const T*>(p.ypre) + ch : nullptr;
  const T* gb = reinterpret_cast<const T*>(p.dout) + ch;
  T* dzb = p.dz ? reinterpret_cast<T*>(p.dz) + ch : nullptr;
  T* ozb = p.outz ? reinterpret_cast<T*>(p.outz) + ch : nullptr;
  float* dub = d.du + ch;           // (fp32 only in this kernel: the entry point refuses 16-bit du / ddelta here)
  float* ddb = d.ddelta + ch;

  // visiting order q = 0..L-1 of sweep 2 is s = L-1-q.  With shared du/ddelta the first Q1 visits park partials.
  const int mid = L / 2;
  const int Q1 = shared ? (rev ? (L - mid) : mid) : 0;
  const bool writes_gate_all = !shared && g == 0;
  const bool gate_here_always = writes_gate_all || !bidir;

  f32x2 gcar[SCAN_NS / 2], dA_acc[SCAN_NS / 2];
#pragma unroll
  for (int k = 0; k < SCAN_NS / 2; ++k) { gcar[k] = pk2(0.f, 0.f); dA_acc[k] = pk2(0.f, 0.f); }
  float dD_acc = 0.f;
  bool synced = false;

  for (int c = nchunks - 1; c >= 0; --c) {
    int s0, ns; chunk_range(c, s0, ns);
    // ---- recompute the chunk's states into shared memory
    {
      f32x2 h[SCAN_NS / 2];
      const float* cp = ck + (int64_t)c * SCAN_NS * p.Dch;
#pragma unroll
      for (int k = 0; k < SCAN_NS / 2; ++k) {
        const float lo = cp[(int64_t)(2 * k) * p.Dch], hi = cp[(int64_t)(2 * k + 1) * p.Dch];
        h[k] = pk2(lo, hi);
        hcol[(2 * k) * SB_CH] = lo; hcol[(2 * k + 1) * SB_CH] = hi;
      }
      if (ns == SB_TT) {
        float uu[SB_TT], dd_[SB_TT];
#pragma unroll
        for (int j = 0; j < SB_TT; ++j) {
          const int64_t r = row0 + token(s0 + j);
          uu[j] = to_f(ub[r * d.ld_u]); dd_[j] = db[r * d.ld_delta];
        }
#pragma unroll
        for (int j = 0; j < SB_TT; ++j) {
          const int64_t r = row0 + token(s0 + j);
          bwd_recompute_step(uu[j], dd_[j], reinterpret_cast<const float4*>(d.BC + r * d.ld_bc), h, a2,
                             hcol + (j + 1) * SCAN_NS * SB_CH);
        }
      } else {
        for (int j = 0; j < ns; ++j) {
          const int64_t r = row0 + token(s0 + j);
          bwd_recompute_step(to_f(ub[r * d.ld_u]), db[r * d.ld_delta], reinterpret_cast<const float4*>(d.BC + r * d.ld_bc),
                             h, a2, hcol + (j + 1) * SCAN_NS * SB_CH);
        }
      }
    }
    // (each thread only reads back its own column of hist: no barrier needed)

    // ---- reverse-time recurrence over the chunk; the next step's operands are fetched one step ahead
    float nu, ndl, ngo, nz = 0.f, nyp = 0.f;
    {
      const int64_t r = row0 + token(s0 + ns - 1);
      nu = to_f(ub[r * d.ld_u]); ndl = db[r * d.ld_delta]; ngo = to_f(gb[r * p.ld_dout]);
      if (zb) nz = to_f(zb[r * p.ld_z]);
      if (yb) nyp = to_f(yb[r * p.ld_y]);
    }
    for (int j = ns - 1; j >= 0; --j) {
      const int s = s0 + j;
      const int qv = L - 1 - s;                       // visit index of sweep 2
      if (shared && !synced && qv == Q1) { __syncthreads(); synced = true; }   // all partials are parked
      const bool finalize = !shared || qv >= Q1;
      const int64_t r = row0 + token(s);
      const float u = nu, dl = ndl, go = ngo * scale, zv = nz, yp = nyp;
      if (j > 0) {
        const int64_t rn = row0 + token(s - 1);
        nu = to_f(ub[rn * d.ld_u]); ndl = db[rn * d.ld_delta]; ngo = to_f(gb[rn * p.ld_dout]);
        if (zb) nz = to_f(zb[rn * p.ld_z]);
        if (yb) nyp = to_f(yb[rn * p.ld_y]);
      }
      const float sz = zb ? silu_f(zv) : 1.f;
      const float dy = go * sz;
      dD_acc = fmaf(dy, u, dD_acc);
      const float4* bq = reinterpret_cast<const float4*>(d.BC + r * d.ld_bc);
      float red[32];
      const float dlu = dl * u;
      const f32x2 dl2 = pk2(dl, dl), dy2 = pk2(dy, dy), dlu2 = pk2(dlu, dlu);
      f32x2 sB2 = pk2(0.f, 0.f), dd2 = pk2(0.f, 0.f);
      const float* hb = hcol + j * SCAN_NS * SB_CH;              // h_{s-1}
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 Bv = __ldg(bq + q);
        const float4 Cv = __ldg(bq + 4 + q);
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          const int k = 2 * q + hq;                               // state pair (2k, 2k+1)
          const f32x2 Bp = hq ? pk2(Bv.z, Bv.w) : pk2(Bv.x, Bv.y);
          const f32x2 Cp = hq ? pk2(Cv.z, Cv.w) : pk2(Cv.x, Cv.y);
          float e0, e1; upk2(mul2(dl2, a2[k]), e0, e1);
          const f32x2 a = pk2(ex2_approx(e0), ex2_approx(e1));
          const f32x2 dh = fma2(Cp, dy2, gcar[k]);
          gcar[k] = mul2(a, dh);
          const f32x2 hprev = pk2(hb[(2 * k) * SB_CH], hb[(2 * k + 1) * SB_CH]);
          const f32x2 hcur = fma2(a, hprev, mul2(dlu2, Bp));      // h_s recomputed (cheaper than two more LDS)
          const f32x2 t1 = mul2(gcar[k], hprev);                  // dh * a * h_{s-1}
          dA_acc[k] = fma2(t1, dl2, dA_acc[k]);
          dd2 = fma2(t1, Av2[k], dd2);
          sB2 = fma2(dh, Bp, sB2);
          float r0, r1; upk2(mul2(dh, dlu2), r0, r1);             // dB contributions
          red[2 * k] = r0; red[2 * k + 1] = r1;
          upk2(mul2(hcur, dy2), r0, r1);                          // dC contributions
          red[SCAN_NS + 2 * k] = r0; red[SCAN_NS + 2 * k + 1] = r1;
        }
      }
      float s0_, s1_, d0_, d1_;
      upk2(sB2, s0_, s1_); upk2(dd2, d0_, d1_);
      const float sB = s0_ + s1_;
      float dd = fmaf(sB, u, d0_ + d1_);
      float duv = fmaf(dl, sB, Dv * dy);
      if (!active) {
#pragma unroll
        for (int i = 0; i < 32; ++i) red[i] = 0.f;
      }
      // cross-channel sums of this token: lane i of each warp ends up with value i; one plain 128-byte store per
      // warp into this warp's slice of the partial workspace (summed over warps by dbc_reduce_kernel)
      const float rsum = warp_transpose_reduce(red, lane);
      d.dbc_ws[((int64_t)part * rows_total + r) * 32 + lane] = rsum;

      const float spg = p.softplus_grad ? (1.f - __expf(-dl)) : 1.f;     // softplus'(pre) = 1 - exp(-delta)
      if (shared) {
        if (active) {
          if (finalize) { duv += dub[r * d.ld_du]; dd = (dd + ddb[r * d.ld_dd]) * spg; }
          dub[r * d.ld_du] = duv; ddb[r * d.ld_dd] = dd;
        }
      } else if (active) {
        dub[r * d.ld_du] = duv; ddb[r * d.ld_dd] = dd * spg;
      }
      if (active && ((shared && finalize) || gate_here_always)) {
        if (zb && (dzb || ozb)) {
          if (dzb) {
            const float sg = __fdividef(1.f, 1.f + __expf(-zv));             // sigmoid(z)
            dzb[r * p.ld_dz] = from_f<T>(go * yp * (sg * (1.f + zv * (1.f - sg))));
          }
          if (ozb) ozb[r * p.ld_oz] = from_f<T>(scale * yp * sz);
        }
      }
    }
  }
  if (shared && !synced) __syncthreads();   // (degenerate L) keep the CTA barrier count equal across directions

  if (active) {
#pragma unroll
    for (int k = 0; k < SCAN_NS / 2; ++k) {
      float lo, hi; upk2(d.dA_log ? mul2(dA_acc[k], Av2[k]) : dA_acc[k], lo, hi);
      atomicAdd(d.dA + (int64_t)ch * SCAN_NS + 2 * k, lo);
      atomicAdd(d.dA + (int64_t)ch * SCAN_NS + 2 * k + 1, hi);
    }
    if (d.dD) atomicAdd(d.dD + ch, dD_acc);
  }
}

// dBC[row][v] += sum over parts of ws[part][row][v]
__global__ void __launch_bounds__(256)
dbc_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dBC, int64_t ld_dbc, int64_t rows, int nparts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 32) return;
  float acc = 0.f;
  for (int pp = 0; pp < nparts; ++pp) acc += ws[(int64_t)pp * rows * 32 + i];
  dBC[(i >> 5) * ld_dbc + (i & 31)] += acc;
}

}  // namespace aum

extern "C" int64_t aum_selective_scan_bwd_dbc_ws_floats(int batch, int L, int D) {
  // sized for the wider of the two kernels' warp counts (generic: 2 per 64 channels, TMA-streamed: 4 per 128)
  const int parts_generic = aum::ceil_div(D, aum::SB_CH) * aum::SB_WARPS, parts_tma = aum::ceil_div(D, 128) * 4;
  return (int64_t)(parts_generic > parts_tma ? parts_generic : parts_tma) * (int64_t)batch * L * 32;
}

extern "C" int64_t aum_selective_scan_bwd_workspace_floats(int batch, int L, int D) {
  return (int64_t)batch * aum::scan_ck_count_max(L) * aum::SCAN_NS * D;
}

extern "C" int aum_selective_scan_bwd(const aum_scan_bwd_dir_t* fwd, const aum_scan_bwd_dir_t* bwd,
                                      const void* z, int64_t ld_z, const void* y_pre, int64_t ld_y,
                                      const void* dout, int64_t ld_dout,
                                      void* dz, int64_t ld_dz, void* out_z, int64_t ld_oz,
                                      int batch, int L, int D, int N, int dtype, float out_scale,
                                      int softplus_grad, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(dout);
  AUM_REQUIRE(fwd || bwd, "aum_selective_scan_bwd: at least one direction is required");
  AUM_REQUIRE(dout, "aum_selective_scan_bwd: null dout");
  AUM_REQUIRE(N == SCAN_NS, "aum_selective_scan_bwd: d_state must be %d", SCAN_NS);
  AUM_REQUIRE(batch >= 0 && L >= 0 && D >= 0 && batch <= 65535, "aum_selective_scan_bwd: bad sizes");
  AUM_REQUIRE(dtype >= AUM_F32 && dtype <= AUM_BF16, "aum_selective_scan_bwd: bad dtype %d", dtype);
  AUM_REQUIRE(!z || y_pre || (!dz && !out_z), "aum_selective_scan_bwd: dz/out_z need the saved pre-gate output y_pre");
  if (batch == 0 || L == 0 || D == 0) return 0;
  ScanBwdParams p;
  memset(&p, 0, sizeof(p));
  const aum_scan_bwd_dir_t* src[2] = {fwd, bwd};
  for (int i = 0; i < 2; ++i) {
    const aum_scan_bwd_dir_t* s = src[i];
    if (!s) continue;
    AUM_REQUIRE(s->u && s->delta && s->A && s->BC && s->du && s->ddelta && s->dA && s->dBC && s->ckpt && s->dbc_ws,
                "aum_selective_scan_bwd: null pointer in direction %d", i);
    AUM_REQUIRE(s->ld_u >= D && s->ld_delta >= D && s->ld_du >= D && s->ld_dd >= D && s->ld_bc >= 2 * N && s->ld_dbc >= 2 * N,
                "aum_selective_scan_bwd: leading dimension too small");
    AUM_REQUIRE(aligned16(s->A) && aligned16(s->BC) && (s->ld_bc % 4) == 0, "aum_selective_scan_bwd: A / BC must be 16-byte aligned");
    ScanBwdDirDev& d = p.dir[p.ndirs++];
    d.u = s->u; d.ld_u = s->ld_u; d.delta = s->delta; d.ld_delta = s->ld_delta; d.A = s->A;
    d.BC = s->BC; d.ld_bc = s->ld_bc; d.D = s->D; d.du = (float*)s->du; d.ld_du = s->ld_du;
    d.ddelta = (float*)s->ddelta; d.ld_dd = s->ld_dd; d.dA = s->dA; d.dD = s->dD; d.dBC = s->dBC; d.ld_dbc = s->ld_dbc;
    d.ckpt = s->ckpt; d.ckpt_valid = s->ckpt_valid; d.reverse = i; d.dbc_ws = s->dbc_ws; d.dA_log = s->dA_is_dAlog ? 1 : 0;
    AUM_REQUIRE(s->dgrad_dtype == AUM_F32 || (s->dgrad_dtype == dtype && dtype != AUM_F32),
                "aum_selective_scan_bwd: du / ddelta must be fp32 or of the call's 16-bit dtype");
    AUM_REQUIRE(p.ndirs == 1 || (s->dgrad_dtype != AUM_F32) == (p.g16 != 0), "aum_selective_scan_bwd: both directions must share dgrad_dtype");
    p.g16 = s->dgrad_dtype != AUM_F32 ? 1 : 0;
    AUM_REQUIRE(s->delta_dtype == AUM_F32 || (s->delta_dtype == dtype && dtype != AUM_F32),
                "aum_selective_scan_bwd: delta must be fp32 or of the call's 16-bit dtype");
    AUM_REQUIRE(p.ndirs == 1 || (s->delta_dtype != AUM_F32) == (p.d16 != 0), "aum_selective_scan_bwd: both directions must share delta_dtype");
    p.d16 = s->delta_dtype != AUM_F32 ? 1 : 0;
  }
  if (p.ndirs == 2) {
    const bool same_du = p.dir[0].du == p.dir[1].du, same_dd = p.dir[0].ddelta == p.dir[1].ddelta;
    AUM_REQUIRE(same_du == same_dd, "aum_selective_scan_bwd: du and ddelta must be shared together or not at all");
    AUM_REQUIRE(p.dir[0].ckpt != p.dir[1].ckpt && p.dir[0].dbc_ws != p.dir[1].dbc_ws,
                "aum_selective_scan_bwd: each direction needs its own checkpoint and dB|dC workspaces");
    p.shared_du = same_du ? 1 : 0;
  }
  p.z = z; p.ld_z = ld_z; p.ypre = y_pre; p.ld_y = ld_y; p.dout = dout; p.ld_dout = ld_dout;
  p.dz = dz; p.ld_dz = ld_dz; p.outz = out_z; p.ld_oz = ld_oz;
  p.batch = batch; p.L = L; p.Dch = D; p.nchunks = scan_ck_count_max(L); p.scale = out_scale;
  p.softplus_grad = softplus_grad;
  dim3 grid(ceil_div(D, SB_CH), batch);
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 2 * (SB_TT + 1) * SCAN_NS * SB_CH * (int)sizeof(float);     // 73 728 B
  static PerDevice<bool> attr_set_dev;
  bool& attr_set = attr_set_dev.cur();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(scan_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(scan_bwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(scan_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("aum_selective_scan_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 2; }
    attr_set = true;
  }
  const int fast = launch_scan_bwd_tma(p, dtype, st);     // TMA-streamed kernel when eligible
  if (fast > 0) return fast;
  AUM_REQUIRE(fast == 0 || !p.d16, "aum_selective_scan_bwd: a 16-bit delta needs the TMA-streamed kernel (forward checkpoints, aligned rows)");
  AUM_REQUIRE(fast == 0 || !p.g16, "aum_selective_scan_bwd: 16-bit du / ddelta need the specialised TMA-streamed kernel "
                                   "(forward checkpoints, z + y_pre + dz + out_z, softplus_grad, separate du / ddelta per direction, D %% 128 == 0)");
  if (fast < 0) {
    switch (dtype) {
      case AUM_F32:  scan_bwd_kernel<float><<<grid, SB_CH * p.ndirs, smem, st>>>(p); break;
      case AUM_F16:  scan_bwd_kernel<__half><<<grid, SB_CH * p.ndirs, smem, st>>>(p); break;
      case AUM_BF16: scan_bwd_kernel<__nv_bfloat16><<<grid, SB_CH * p.ndirs, smem, st>>>(p); break;
    }
    if (int rc = check_launch("aum_selective_scan_bwd")) return rc;
  }
  const int64_t rows = (int64_t)batch * L;
  // one workspace slice per warp of the kernel that ran: 2 per 64-channel CTA (generic), 4 per 128-channel CTA (TMA)
  const int nparts = fast < 0 ? ceil_div(D, SB_CH) * SB_WARPS : ceil_div(D, 128) * 4;
  for (int g = 0; g < p.ndirs; ++g)
    dbc_reduce_kernel<<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(p.dir[g].dbc_ws, p.dir[g].dBC, p.dir[g].ld_dbc, rows, nparts);
  return check_launch("aum_selective_scan_bwd(dbc reduce)");
}
