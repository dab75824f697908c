/*
 * aum_b200.h — C ABI of libaum_b200.so: the B200 (sm_100a) engine behind Audio-Mamba's
 * bidirectional selective-scan hot path.
 *
 * This is the drop-in boundary.  The reference reaches its native code through
 * two pybind modules of un-vendored pip wheels (mamba_ssm==1.1.3.post1, causal_conv1d==1.1.3.post1,
 * /root/reference/README.md:58) that take at::Tensor; each entry point below names the reference
 * call site(s) it replaces (paths relative to /root/reference).  Here everything is plain C:
 * raw device pointers, explicit sizes / leading dimensions, a dtype enum, and the CUDA stream as void*.
 *
 * Conventions
 *   - Every function returns 0 on success, non-zero on error; aum_last_error() gives the message
 *     (thread-local).  No C++ exception crosses this boundary.  No function synchronises the device
 *     or allocates device memory; the caller owns all buffers (PyTorch's caching allocator in the
 *     Python host).  All launches go to the stream passed in; everything is CUDA-graph capturable.
 *   - Activations are TOKEN-MAJOR: a logical (batch, L, C) tensor is addressed as rows = batch*L tokens
 *     with a leading dimension ("ld", in ELEMENTS) between consecutive tokens and channels contiguous.
 *     (The reference's (batch, C, L) tensors are handled on the host side by aum_transpose_*.)
 *   - dtype: AUM_F32 / AUM_F16 / AUM_BF16 for activations; parameters named "float*" are always fp32
 *     as in the reference (A, D, delta_bias: vim-mamba_ssm/mamba_ssm/modules/mamba_simple.py:193,210-211).
 */
#ifndef AUM_B200_H_
#define AUM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUM_B200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define AUM_API __attribute__((visibility("default")))
#else
#define AUM_API
#endif

enum aum_dtype { AUM_F32 = 0, AUM_F16 = 1, AUM_BF16 = 2 };

/* GEMM epilogue activation */
enum aum_act { AUM_ACT_NONE = 0, AUM_ACT_SOFTPLUS = 1 /* torch softplus, threshold 20 */, AUM_ACT_SILU = 2 };
/* act argument of aum_gemm_tn: low 8 bits = aum_act, upper bits = first column the activation applies to
 * (AUM_ACT_FROM(AUM_ACT_SILU, Di) gates only the z half of in_proj's output). */
#define AUM_ACT_FROM(kind, col0) ((int)(kind) | ((int)(col0) << 8))

/* flags of aum_selective_scan_fwd */
enum aum_scan_flags { AUM_SCAN_Z_PREGATED = 1 /* z already holds silu(z) (in_proj epilogue applied it) */ };

/* GEMM backend selection */
enum aum_gemm_backend { AUM_GEMM_AUTO = 0, AUM_GEMM_TCGEN05 = 1, AUM_GEMM_SIMT = 2 };

AUM_API int aum_version(void);
AUM_API const char* aum_last_error(void);
/* sm_count / compute capability of the current device; 0 on success. */
AUM_API int aum_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Dense projections:  C[M,N] = act( row_scale[m] * (A[M,K] @ W[N,K]^T) + bias[n] )
 *   replaces the cuBLAS calls at mamba_simple.py:185-189 (in_proj), selective_scan_interface.py:467
 *   (x_proj), :468 (dt_proj), :517 (out_proj).  A and W are K-contiguous ("TN"), which is exactly
 *   nn.Linear's weight layout and the token-major activation layout.
 *   ab_dtype F16/BF16 -> tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue);
 *   ab_dtype F32     -> fp32 CUDA-core kernel (strict-parity tier).
 *   Optional split output: columns [0,split) go to C (c_dtype), columns [split,N) to C2 (c2_dtype) at
 *   column index n-split; pass C2=NULL, split=N for a single output (split%8==0 keeps 16-byte stores).
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_gemm_tn(const void* A, int64_t lda, const void* W, int64_t ldw, int ab_dtype,
                void* C, int64_t ldc, int c_dtype,
                void* C2, int64_t ldc2, int c2_dtype, int split,
                int M, int N, int K,
                const float* bias, const float* row_scale, int act,
                int backend, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight-gradient products of the projections:  dW[No, Ki] += dY[T, No]^T @ X[T, Ki]   (ACCUMULATES: zero dW first,
 * or point it into the trainer's flat gradient buffer).
 *   replaces the reductions over the token axis in BiMambaInnerFn / MambaInnerFn / MambaInnerFnNoOutProj .backward:
 *   d(out_proj.weight) = einsum("eB,dB->ed", dout, out_z)  (selective_scan_interface.py:563, :395, :254-256 caller),
 *   d(dt_proj.weight)  = einsum("dB,Br->dr", ddelta, x_dbl[:, :R])  (:586, :417, :273),
 *   d(x_proj.weight)   = einsum("Br,Bd->rd", dx_dbl, conv1d_out)    (:589, :420, :276),
 *   and autograd's d(in_proj.weight) of the matmul at mamba_simple.py:185-189.
 *   dY, X: fp16 / bf16 token-major rows (16-byte aligned bases and row pitches); dW fp32 with ld_dw >= Ki.
 *   tcgen05 with MN-major operand descriptors (no transposed copies), split-K over tokens with fp32 atomic adds:
 *   the summation ORDER of the partial sums is not deterministic (every addend is).
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_gemm_wgrad(const void* dY, int64_t ld_dy, const void* X, int64_t ld_x, int ab_dtype,
                   float* dW, int64_t ld_dw, int T, int No, int Ki, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Depthwise causal conv1d (+bias, +SiLU) along the token axis.
 *   replaces causal_conv1d_cuda.causal_conv1d_fwd(x, w, bias, None, silu)
 *   (selective_scan_interface.py:177,239,318,380,463,532) and causal_conv1d_fn (:646,:683).
 *   x, out: (batch*L, D) token-major with ldx / ldo; w: (D, W) fp32; bias: (D) fp32 or NULL; 2<=W<=4.
 *   reverse=1 computes the anti-causal conv  out[l] = b + sum_k w[k] x[l+(W-1)-k]  — what the Bi-Bi
 *   pipeline gets by running on xz.flip(-1) and flipping back (mamba_simple.py:229-241), without flips.
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_causal_conv1d_fwd(const void* x, int64_t ldx, const float* w, const float* bias,
                          void* out, int64_t ldo, int batch, int L, int D, int W,
                          int dtype, int silu, int reverse, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused  u = SiLU(causal_conv1d(x) + bias)  and  x_dbl = u @ W_x^T  split into dt | [B|C]   (16-bit activations).
 *   replaces causal_conv1d_cuda.causal_conv1d_fwd(x, w, b, None, True) followed by
 *   F.linear(rearrange(conv1d_out, "b d l -> (b l) d"), x_proj_weight)
 *   (selective_scan_interface.py:463+467, :177+181, :318+322) and the B / C slicing of :473-496.
 *   The conv is the producer of the x_proj tensor-core operand (conv results are written into the swizzled
 *   shared-memory tile tcgen05.mma reads, and the same tile leaves for HBM as u): one launch, x read once.
 *   x: (batch*L, Di) token-major, ldx; conv_w (Di, 4) fp32 [d_conv == 4]; conv_b (Di) fp32 or NULL;
 *   Wx: (R + N2, Di) dtype, K-contiguous; outputs u (batch*L, Di) dtype, dt (batch*L, >= R) dtype (columns [0, R)
 *   written), bc (batch*L, N2) fp32 (N2 = 2 * d_state).  R % 8 == 0, R + N2 <= 128, Di % 8 == 0.
 *   reverse = 1: anti-causal conv (Bi-Bi's second branch, mamba_simple.py:229-241, without flips).
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_conv_xproj_fwd(const void* x, int64_t ldx, const float* conv_w, const float* conv_b,
                       const void* Wx, int64_t ldw, void* u, int64_t ldu,
                       void* dt, int64_t ld_dt, float* bc, int64_t ld_bc,
                       int batch, int L, int Di, int R, int N2, int dtype, int reverse, void* stream);

/* Backward of the above.  replaces causal_conv1d_cuda.causal_conv1d_bwd(x, w, bias, dout, None, dx, silu)
 *   (selective_scan_interface.py:281-283, 425-427, 594-596).  dout (+ optional dout2, dout3, same pitch; summed on the
 *   fly — the scan's du of either direction (:556) and the x_proj term of :590): gradient w.r.t. the conv output
 *   (after the activation), all terms of ONE element type dout_dtype: AUM_F32, or the call's 16-bit `dtype` (what the
 *   reference accumulates dconv1d_out in under autocast; the sum is formed in fp32 here); dx: written (dtype); dw (D, W)
 *   and dbias (D) are ACCUMULATED (+=): zero them first. */
AUM_API int aum_causal_conv1d_bwd(const void* x, int64_t ldx, const float* w, const float* bias,
                          const void* dout, const void* dout2, const void* dout3, int64_t ldd, int dout_dtype,
                          void* dx, int64_t ld_dx,
                          float* dw, float* dbias, int batch, int L, int D, int W,
                          int dtype, int silu, int reverse, void* stream);

/* out[r, c] = (out_dtype)(a[r, c] + b[r, c]),  colsum[c] += sum_r (a + b)[r, c]   (b and colsum optional).
 *   One pass for the dt_proj chain of the three backward functions (selective_scan_interface.py:556 ddelta of the two
 *   directions summed, :583-586 delta_proj bias gradient = sum over tokens, and the 16-bit operand of the
 *   d(dt_proj.weight) / d(x_dbl) products).  a, b: (rows, cols) of in_dtype (fp32, or 16-bit as the backward scan leaves
 *   ddelta with dgrad_dtype) with pitch ld; the sum is formed in fp32; cols % 4 == 0. */
AUM_API int aum_sum_cast_colsum(const void* a, const void* b, int64_t ld, int in_dtype, void* out, int64_t ld_out, int out_dtype,
                        float* colsum, int rows, int cols, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Selective scan, one or both time directions in ONE launch.
 *   replaces selective_scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus)
 *   (selective_scan_interface.py:37,213,354,499) and — for the reverse direction — the second call on
 *   five flip(-1) copies plus the un-flip and add (:503-507).
 *   Per direction:  delta' = softplus?(delta + delta_bias);  h_l = exp(delta'_l A) h_{l∓1} + delta'_l B_l u_l;
 *                   y_l = <C_l, h_l> + D u_l.
 *   out = out_scale * (y_fwd + y_bwd) * silu(z)   (z optional; a NULL direction contributes 0).
 *   Fo-Bi (bimamba v1): both directions share u/delta/B/C/D/bias and differ in A (A vs A_b).
 *   Bi-Bi (v2): every field differs per direction; out_scale = 0.5 when if_devide_out.
 * ------------------------------------------------------------------------------------------- */
typedef struct aum_scan_dir {
  const void* u;     int64_t ld_u;                      /* (batch*L, D)  dtype            */
  const void* delta; int64_t ld_delta; int delta_dtype; /* (batch*L, D)  any aum_dtype    */
  const float* A;                                       /* (D, N) fp32, real, negative    */
  const void* Bm;    int64_t ld_B;                      /* (batch*L, N)  bc_dtype         */
  const void* Cm;    int64_t ld_C;                      /* (batch*L, N)  bc_dtype         */
  int bc_dtype;
  const float* D;                                       /* (D) fp32 or NULL               */
  const float* delta_bias;                              /* (D) fp32 or NULL               */
  int delta_softplus;                                   /* apply softplus(delta+bias)     */
  float* last_state;                                    /* (batch, D, N) fp32 or NULL     */
  float* ckpt;                                          /* optional (training, N == 16): state checkpoints for
                                                           aum_selective_scan_bwd, aum_selective_scan_bwd_workspace_floats()
                                                           floats, layout [batch][chunk][N][D]                */
} aum_scan_dir_t;

AUM_API int aum_selective_scan_fwd(const aum_scan_dir_t* fwd, const aum_scan_dir_t* bwd,
                           const void* z, int64_t ld_z,
                           void* out, int64_t ld_out,
                           int batch, int L, int D, int N, int dtype,
                           float out_scale,
                           void* y_pre, int64_t ld_ypre,   /* optional (training): pre-gate y_fwd+y_bwd, dtype */
                           int flags,                      /* aum_scan_flags */
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * Selective scan backward, one or both time directions in ONE launch.
 *   replaces selective_scan_cuda.bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, dz, softplus, recompute_out_z)
 *   (selective_scan_interface.py:62,247,389,541,548) and the flip/un-flip/add combination of :554-561.
 *   delta is the POST-softplus value (the softplus/bias chain rule is applied by the caller's dt_proj backward:
 *   d(pre) = ddelta * (1 - exp(-delta))).  [B|C] are packed fp32 rows (batch*L, 2N), N = 16.
 *   Outputs per direction: du, ddelta (fp32, (batch*L, D)) — pass the SAME pointers in both directions when u and
 *   delta are shared (Fo-Bi) and the kernel sums the two directions; dA (D,N), dD (D), dBC (batch*L, 2N) are
 *   ACCUMULATED (+=): zero them first.  ckpt: workspace of aum_selective_scan_bwd_workspace_floats() floats per
 *   direction.  Gate: dz and the recomputed out_z = out_scale*y_pre*silu(z) are written when z and y_pre are given.
 *   d/dz is the mathematically correct gradient (both directions), see SURVEY.md Q2.
 * ------------------------------------------------------------------------------------------- */
typedef struct aum_scan_bwd_dir {
  const void* u;      int64_t ld_u;      /* (batch*L, D) dtype */
  const void* delta;  int64_t ld_delta;  /* (batch*L, D) post-softplus; fp32, or dtype when delta_dtype says so */
  const float* A;                        /* (D, N) */
  const float* BC;    int64_t ld_bc;     /* (batch*L, 2N) fp32 [B|C] */
  const float* D;                        /* (D) or NULL */
  void* du;      int64_t ld_du;          /* out (batch*L, D) fp32, or dtype when dgrad_dtype says so */
  void* ddelta;  int64_t ld_dd;          /* out (batch*L, D) same dtype as du */
  float* dA;                             /* += (D, N) */
  float* dD;                             /* += (D) or NULL */
  float* dBC;    int64_t ld_dbc;         /* += (batch*L, 2N) */
  float* dbc_ws;                         /* workspace: aum_selective_scan_bwd_dbc_ws_floats() floats (per direction) */
  float* ckpt;                           /* workspace (same size as the forward's ckpt) */
  int ckpt_valid;                        /* 1: ckpt was filled by aum_selective_scan_fwd for the SAME direction slot and
                                            directionality (uni/bi) — the backward then skips its own forward sweep */
  int dgrad_dtype;                       /* dtype of du and ddelta: AUM_F32 (= 0, the default of a zeroed struct) or the
                                            call's 16-bit `dtype` - what the reference's kernels return under autocast
                                            (selective_scan_interface.py:541-561: du, ddelta in the input dtypes).  The
                                            16-bit form needs the training configuration the TMA-streamed kernel is
                                            specialised for (checkpoints, z, y_pre, dz, out_z, softplus_grad, one du /
                                            ddelta pair per direction, D % 128 == 0); anything else is refused. */
  int delta_dtype;                       /* dtype of delta: AUM_F32 (= 0) or the call's 16-bit `dtype` (delta as the forward
                                            call read it, see aum_scan_dir_t.delta_dtype); TMA-streamed kernel only. */
  int dA_is_dAlog;                       /* 1: dA += dA * A, the gradient w.r.t. A_log of A = -exp(A_log)
                                            (mamba_simple.py:193,197): lets the caller point dA at the A_log gradient itself */
} aum_scan_bwd_dir_t;

AUM_API int64_t aum_selective_scan_bwd_workspace_floats(int batch, int L, int D);
AUM_API int64_t aum_selective_scan_bwd_dbc_ws_floats(int batch, int L, int D);
AUM_API int aum_selective_scan_bwd(const aum_scan_bwd_dir_t* fwd, const aum_scan_bwd_dir_t* bwd,
                                   const void* z, int64_t ld_z, const void* y_pre, int64_t ld_y,
                                   const void* dout, int64_t ld_dout,
                                   void* dz, int64_t ld_dz, void* out_z, int64_t ld_oz,
                                   int batch, int L, int D, int N, int dtype, float out_scale,
                                   int softplus_grad, /* 1: ddelta := ddelta * (1 - exp(-delta)) = grad w.r.t. the
                                                         pre-softplus dt_proj output (softplus'(x) = 1 - e^{-softplus(x)}) */
                                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused residual-add + RMSNorm (fp32 residual stream) — the op on either side of the mixer.
 *   replaces Triton _layer_norm_fwd_1pass_kernel (vim-mamba_ssm/mamba_ssm/ops/triton/layernorm.py:65-120)
 *   as called by rms_norm_fn (:477) from src/models/mamba_models.py:77-97,646-657.
 *   r = x + residual_in (fp32);  y = r * rsqrt(mean(r^2)+eps) * weight (+bias);
 *   residual_out (fp32, optional) = r;  rstd_out (rows, optional) saved for backward.
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_add_rmsnorm_fwd(const void* x, int64_t ldx, int x_dtype,
                        const void* residual_in, int64_t ldr, int r_dtype,
                        const float* weight, const float* bias,
                        void* y, int64_t ldy, int y_dtype,
                        void* residual_out, int64_t ldro, int ro_dtype,
                        float* rstd_out, int rows, int dim, float eps, void* stream);

/* Backward of the above (RMSNorm, no bias).  replaces Triton _layer_norm_bwd_kernel (layernorm.py:196-290).
 *   dy: grad of y (dtype); dresidual_out: grad flowing into residual_out (fp32) or NULL; r: the saved residual_out
 *   (x + residual_in, fp32); rstd: saved by the forward.  Writes dx (same dtype as dy) and, optionally, the fp32
 *   dresidual_in (same values); dweight (dim) is ACCUMULATED (+=).  dim % 8 == 0, dim <= 2048. */
AUM_API int aum_add_rmsnorm_bwd(const void* dy, int64_t ld_dy, int dy_dtype,
                        const float* dresidual_out, int64_t ld_dro,
                        const float* r, int64_t ld_r, const float* rstd, const float* weight,
                        void* dx, int64_t ld_dx, float* dresidual_in, int64_t ld_dri,
                        float* dweight, int rows, int dim, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Front end of the block stack: data movement either side of the patch-embed GEMM.
 *   aum_patchify: x (batch, T, F) fp32 spectrogram -> im2col rows (batch * (F/pf) * (T/pt), pf * pt) in `dtype`, so that
 *     FlexiPatchEmbed's stride == kernel conv2d (src/utilities/tokenization.py:278-310; img[b,0,f,t] = x[b,t,f],
 *     src/models/mamba_models.py:510-515) is ONE aum_gemm_tn with the conv weight flattened to (Dm, pf * pt).
 *     Row order = x.flatten(2).transpose(1, 2) (frequency block major); column = kf * pt + kt.  pt % 4 == 0.
 *   aum_assemble_tokens: patch tokens (batch, N, Dm) fp32 -> hidden (batch, N + 1, Dm) fp32: cls token inserted at
 *     index N / 2, absolute position embedding added (mamba_models.py:525-541; tokenization.py:414-451: slot 0 of
 *     pos (N + 1, Dm) belongs to the cls token).  cls: (Dm).  Dm % 4 == 0, 16-byte aligned buffers.
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_patchify(const float* x, void* cols, int batch, int T, int F, int pf, int pt, int dtype, void* stream);
AUM_API int aum_assemble_tokens(const float* tok, const float* pos, const float* cls, float* hidden,
                        int batch, int N, int Dm, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Adam step over one flat fp32 buffer (the caller of the training path).
 *   replaces torch.optim.Adam(trainables, lr, weight_decay=5e-7, betas=(0.95, 0.999)).step()
 *   (src/traintest.py:32-34, :169) — same update rule (L2 weight decay folded into the gradient, bias-corrected
 *   moments, eps added to sqrt(v_hat)), one pass over p, g, m, v (n floats each, 16-byte aligned) instead of the
 *   multi-tensor implementation's ~10.  step counts from 1.  grad_scale multiplies g first (1/world_size after the
 *   SUM all-reduce of the flat gradient buffer).  p, m, v are updated in place; g is not modified.
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_adam_step(float* p, const float* g, float* m, float* v, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay,
                  int step, float grad_scale, void* stream);
/* The same update with the step number held in DEVICE memory: *step_dev is incremented first (a one-thread kernel), then
 * used for the bias corrections, so that a CUDA graph of the whole training step (src/traintest.py:144-169: forward,
 * loss, backward, all-reduce, optimizer.step()) can be replayed without host-side state.  Initialise *step_dev to the
 * number of steps already taken (0 for a fresh optimiser).  p16 (optional, may be NULL): a 16-bit shadow of the updated
 * parameters (p16_dtype AUM_F16 / AUM_BF16, n elements, 8-byte aligned) written in the same pass - the autocast copies
 * of the weights the next forward needs (the reference's autocast casts them once per use, mamba_simple.py:185-189).  */
AUM_API int aum_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n,
                      float lr, float beta1, float beta2, float eps, float weight_decay,
                      int* step_dev, float grad_scale, void* p16, int p16_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layout adapters for the reference's channel-major (batch, C, L) tensors:
 *   dst[b, j, i] = src[b, i, j]  for a (batch, R, C) -> (batch, C, R) transpose with explicit strides
 *   (elements): src element (b,i,j) at b*src_bs + i*src_ld + j; dst element (b,j,i) at b*dst_bs + j*dst_ld + i.
 *   Used where the functional API receives xz as (B, 2*Di, L) with stride(-1)==1
 *   (selective_scan_interface.py:458-461) or must return (B, Di, L) (:224).
 * ------------------------------------------------------------------------------------------- */
AUM_API int aum_transpose(const void* src, int64_t src_bs, int64_t src_ld,
                  void* dst, int64_t dst_bs, int64_t dst_ld,
                  int batch, int R, int C, int src_dtype, int dst_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUM_B200_H_ */
