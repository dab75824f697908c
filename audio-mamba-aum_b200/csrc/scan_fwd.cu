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
// per (channel, direction), the d_state (<=16) recurrences of that channel held in registers, exp() through
// ex2.approx on A pre-scaled by log2(e).  Token-major layout makes every per-token access of a warp one
// contiguous 64/128-byte segment.  B_l/C_l (shared by all channels of a token) are staged per 64-token chunk
// into shared memory as fp32 and read back as broadcast 128-bit loads.
// Two directions, one pass: the forward threads walk l = 0..L-1 while the backward threads walk l = L-1..0.
// In the first half of its walk a direction parks its partial y (+D u) in `out`; after one CTA barrier at the
// midpoint each direction meets rows the other one has already visited, adds the parked partial, applies
// the SiLU(z) gate and writes the final value.  Extra traffic: one write + one (mostly L2) read of `out`.
//
// Roofline note: 16 ex2 per (token, channel, direction) make this kernel MUFU-bound
// (16 results/clk/SM) well before it is HBM-bound; algorithmic bytes per (token, channel) =
// s*(u + delta + z + out) + 2*N*4/ D-share of B,C.
#include <stdlib.h>

#include "common.cuh"

namespace aum {

constexpr int SCAN_NS = 16;   // states held in registers (d_state <= 16; padded with inert states)
constexpr int SCAN_TC = 64;   // tokens per staged B/C chunk
constexpr int SCAN_U = 4;     // tokens per software-pipelined batch

struct ScanDirDev {
  const void* u; int64_t ld_u;
  const void* delta; int64_t ld_delta;
  const float* A;
  const void* Bm; int64_t ld_B;
  const void* Cm; int64_t ld_C;
  int bc_dt;
  const float* D;
  const float* delta_bias;
  int delta_softplus;
  float* last_state;
  int reverse;   // 0: walks l = 0..L-1, 1: walks l = L-1..0
};

struct ScanParams {
  ScanDirDev dir[2];
  int ndirs;
  const void* z; int64_t ld_z;
  void* out; int64_t ld_out;
  int batch, L, Dch, N;
  float out_scale;
};

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <typename T, typename TD, int CH>
__global__ void __launch_bounds__(2 * CH, (CH == 64) ? 5 : 2)
scan_fwd_kernel(const ScanParams p) {
  __shared__ __align__(16) float bc_smem[2][SCAN_TC][2 * SCAN_NS];

  const int g = threadIdx.x / CH;              // direction slot of this thread
  if (g >= p.ndirs) return;                    // (single-direction launches use CH threads)
  const int tig = threadIdx.x - g * CH;        // thread index within the direction group
  const ScanDirDev& d = p.dir[g];
  const int ch = blockIdx.x * CH + tig;
  const bool active = ch < p.Dch;
  const int b = blockIdx.y;
  const int L = p.L, N = p.N;
  const bool bidir = p.ndirs == 2;
  const int64_t row0 = (int64_t)b * L;

  const T* __restrict__ up = reinterpret_cast<const T*>(d.u) + ch;
  const TD* __restrict__ dp = reinterpret_cast<const TD*>(d.delta) + ch;
  const T* __restrict__ zp = p.z ? reinterpret_cast<const T*>(p.z) + ch : nullptr;
  T* op = reinterpret_cast<T*>(p.out) + ch;

  float a2[SCAN_NS], h[SCAN_NS];
#pragma unroll
  for (int n = 0; n < SCAN_NS; ++n) { a2[n] = 0.f; h[n] = 0.f; }
  if (active) {
    if (N == SCAN_NS) {
      const float4* ap = reinterpret_cast<const float4*>(d.A + (int64_t)ch * SCAN_NS);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(ap + i);
        a2[4 * i + 0] = v.x * 1.4426950408889634f; a2[4 * i + 1] = v.y * 1.4426950408889634f;
        a2[4 * i + 2] = v.z * 1.4426950408889634f; a2[4 * i + 3] = v.w * 1.4426950408889634f;
      }
    } else {
#pragma unroll
      for (int n = 0; n < SCAN_NS; ++n)
        if (n < N) a2[n] = __ldg(d.A + (int64_t)ch * N + n) * 1.4426950408889634f;
    }
  }
  const float Dv = (active && d.D) ? __ldg(d.D + ch) : 0.f;
  const float dbias = (active && d.delta_bias) ? __ldg(d.delta_bias + ch) : 0.f;
  const bool sp = d.delta_softplus != 0;
  const float oscale = p.out_scale;

  // phase 1: steps [0, S1) park partials; phase 2: steps [S1, L) finalise.  Unidirectional: S1 = 0.
  const int mid = L / 2;
  const int S1 = bidir ? (d.reverse ? (L - mid) : mid) : 0;

  float (*bc)[2 * SCAN_NS] = bc_smem[g];

  for (int phase = 0; phase < 2; ++phase) {
    const int s_begin = phase == 0 ? 0 : S1;
    const int s_end = phase == 0 ? S1 : L;
    const bool finalize = phase == 1;
    const bool read_partial = finalize && bidir;

    for (int s0 = s_begin; s0 < s_end; s0 += SCAN_TC) {
      const int ns = min(SCAN_TC, s_end - s0);
      // ---- stage B_l, C_l of this chunk (step order) into shared memory as fp32
      group_barrier(1 + g, CH);   // previous chunk fully consumed
      for (int idx = tig; idx < ns * 2 * SCAN_NS; idx += CH) {
        const int t = idx / (2 * SCAN_NS), j = idx % (2 * SCAN_NS);
        const int s = s0 + t;
        const int l = d.reverse ? (L - 1 - s) : s;
        float v = 0.f;
        if (j < SCAN_NS) { if (j < N) v = load_as_f(d.Bm, (row0 + l) * d.ld_B + j, d.bc_dt); }
        else { const int jj = j - SCAN_NS; if (jj < N) v = load_as_f(d.Cm, (row0 + l) * d.ld_C + jj, d.bc_dt); }
        bc[t][j] = v;
      }
      group_barrier(1 + g, CH);

      // ---- walk the chunk, SCAN_U tokens per software-pipelined batch
      float cu[SCAN_U], cd[SCAN_U], cz[SCAN_U], cp[SCAN_U];
      float nu[SCAN_U], nd[SCAN_U], nz[SCAN_U], np_[SCAN_U];
      auto load_batch = [&](int t0, float (&lu)[SCAN_U], float (&ld)[SCAN_U], float (&lz)[SCAN_U], float (&lp)[SCAN_U]) {
#pragma unroll
        for (int i = 0; i < SCAN_U; ++i) {
          const int t = t0 + i;
          lu[i] = 0.f; ld[i] = 0.f; lz[i] = 0.f; lp[i] = 0.f;
          if (active && t < ns) {
            const int s = s0 + t;
            const int l = d.reverse ? (L - 1 - s) : s;
            const int64_t r = row0 + l;
            lu[i] = to_f(up[r * d.ld_u]);
            ld[i] = to_f(dp[r * d.ld_delta]);
            if (finalize && zp) lz[i] = to_f(zp[r * p.ld_z]);
            if (read_partial) lp[i] = to_f(op[r * p.ld_out]);
          }
        }
      };
      load_batch(0, cu, cd, cz, cp);
      for (int t0 = 0; t0 < ns; t0 += SCAN_U) {
        load_batch(t0 + SCAN_U, nu, nd, nz, np_);
#pragma unroll
        for (int i = 0; i < SCAN_U; ++i) {
          const int t = t0 + i;
          if (t < ns) {
            float dl = cd[i] + dbias;
            if (sp) dl = softplus_f(dl);
            const float u = cu[i];
            const float du = dl * u;
            float y = Dv * u;
            const float4* bq = reinterpret_cast<const float4*>(&bc[t][0]);
#pragma unroll
            for (int q = 0; q < SCAN_NS / 4; ++q) {
              const float4 Bv = bq[q];
              const float4 Cv = bq[SCAN_NS / 4 + q];
              const float bb[4] = {Bv.x, Bv.y, Bv.z, Bv.w};
              const float cc[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int n = 4 * q + k;
                const float dA = ex2_approx(dl * a2[n]);
                h[n] = fmaf(dA, h[n], du * bb[k]);
                y = fmaf(h[n], cc[k], y);
              }
            }
            if (active) {
              const int s = s0 + t;
              const int l = d.reverse ? (L - 1 - s) : s;
              const int64_t r = row0 + l;
              if (!finalize) {
                op[r * p.ld_out] = from_f<T>(y);
              } else {
                float tot = y + cp[i];
                if (zp) tot *= silu_f(cz[i]);
                op[r * p.ld_out] = from_f<T>(tot * oscale);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < SCAN_U; ++i) { cu[i] = nu[i]; cd[i] = nd[i]; cz[i] = nz[i]; cp[i] = np_[i]; }
      }
    }
    if (phase == 0 && bidir) __syncthreads();   // every partial of both directions is parked
  }

  if (active && d.last_state != nullptr) {
#pragma unroll
    for (int n = 0; n < SCAN_NS; ++n)
      if (n < N) d.last_state[((int64_t)b * p.Dch + ch) * N + n] = h[n];
  }
}

template <typename T, typename TD>
static int launch_scan_td(const ScanParams& p, int ch, cudaStream_t st) {
  dim3 grid(ceil_div(p.Dch, ch), p.batch);
  if (ch == 64) scan_fwd_kernel<T, TD, 64><<<grid, 64 * p.ndirs, 0, st>>>(p);
  else scan_fwd_kernel<T, TD, 128><<<grid, 128 * p.ndirs, 0, st>>>(p);
  return check_launch("aum_selective_scan_fwd");
}

template <typename T>
static int launch_scan_t(const ScanParams& p, int delta_dt, int dtype, int ch, cudaStream_t st) {
  if (delta_dt == dtype) return launch_scan_td<T, T>(p, ch, st);
  if (delta_dt == AUM_F32) return launch_scan_td<T, float>(p, ch, st);
  set_error("aum_selective_scan_fwd: delta dtype must equal the activation dtype or be fp32");
  return 1;
}

}  // namespace aum

extern "C" int aum_selective_scan_fwd(const aum_scan_dir_t* fwd, const aum_scan_dir_t* bwd,
                                      const void* z, int64_t ld_z, void* out, int64_t ld_out,
                                      int batch, int L, int D, int N, int dtype,
                                      float out_scale, void* stream) {
  using namespace aum;
  AUM_REQUIRE(fwd || bwd, "aum_selective_scan_fwd: at least one direction is required");
  AUM_REQUIRE(out, "aum_selective_scan_fwd: null output");
  AUM_REQUIRE(N >= 1 && N <= SCAN_NS, "aum_selective_scan_fwd: d_state %d unsupported (1..%d)", N, SCAN_NS);
  AUM_REQUIRE(batch >= 0 && L >= 0 && D >= 0, "aum_selective_scan_fwd: negative size");
  AUM_REQUIRE(batch <= 65535, "aum_selective_scan_fwd: batch too large");
  AUM_REQUIRE(dtype >= AUM_F32 && dtype <= AUM_BF16, "aum_selective_scan_fwd: bad dtype %d", dtype);
  if (batch == 0 || L == 0 || D == 0) return 0;
  ScanParams p;
  memset(&p, 0, sizeof(p));
  int delta_dt = -1;
  const aum_scan_dir_t* src[2] = {fwd, bwd};
  for (int i = 0; i < 2; ++i) {
    const aum_scan_dir_t* s = src[i];
    if (!s) continue;
    AUM_REQUIRE(s->u && s->delta && s->A && s->Bm && s->Cm, "aum_selective_scan_fwd: null pointer in direction %d", i);
    AUM_REQUIRE(s->ld_u >= D && s->ld_delta >= D && s->ld_B >= N && s->ld_C >= N, "aum_selective_scan_fwd: leading dimension too small");
    AUM_REQUIRE(s->bc_dtype >= AUM_F32 && s->bc_dtype <= AUM_BF16, "aum_selective_scan_fwd: bad bc_dtype");
    AUM_REQUIRE(delta_dt < 0 || delta_dt == s->delta_dtype, "aum_selective_scan_fwd: both directions must share delta_dtype");
    AUM_REQUIRE(N != SCAN_NS || aligned16(s->A), "aum_selective_scan_fwd: A must be 16-byte aligned");
    delta_dt = s->delta_dtype;
    ScanDirDev& d = p.dir[p.ndirs++];
    d.u = s->u; d.ld_u = s->ld_u; d.delta = s->delta; d.ld_delta = s->ld_delta; d.A = s->A;
    d.Bm = s->Bm; d.ld_B = s->ld_B; d.Cm = s->Cm; d.ld_C = s->ld_C; d.bc_dt = s->bc_dtype;
    d.D = s->D; d.delta_bias = s->delta_bias; d.delta_softplus = s->delta_softplus;
    d.last_state = s->last_state; d.reverse = i;
  }
  p.z = z; p.ld_z = ld_z; p.out = out; p.ld_out = ld_out;
  p.batch = batch; p.L = L; p.Dch = D; p.N = N; p.out_scale = out_scale;
  AUM_REQUIRE(ld_out >= D && (!z || ld_z >= D), "aum_selective_scan_fwd: leading dimension too small");

  int ch = 64;
  if (const char* e = getenv("AUM_SCAN_CH")) { int v = atoi(e); if (v == 64 || v == 128) ch = v; }
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case AUM_F32:  return launch_scan_t<float>(p, delta_dt, dtype, ch, st);
    case AUM_F16:  return launch_scan_t<__half>(p, delta_dt, dtype, ch, st);
    case AUM_BF16: return launch_scan_t<__nv_bfloat16>(p, delta_dt, dtype, ch, st);
  }
  return 1;
}
