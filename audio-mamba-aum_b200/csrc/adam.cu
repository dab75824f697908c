// Fused Adam step over ONE flat fp32 parameter buffer (and its flat gradient / moment buffers).
// Replaces the optimiser step of the reference's training loop (/root/reference/src/traintest.py:32-34:
// torch.optim.Adam(trainables, lr, weight_decay=5e-7, betas=(0.95, 0.999)); :169 optimizer.step()), whose
// multi-tensor implementation makes ~10 passes over parameters, gradients and moments; here one pass:
// 16 B read + 12 B written per parameter (p, g, m, v in; p, m, v out), 128-bit accesses.
// Semantics = torch.optim.Adam (L2 weight decay folded into the gradient, bias-corrected moments, eps added to
// sqrt(v_hat)); grad_scale multiplies the gradient first (1/world_size after a SUM all-reduce, or a loss-scale inverse).
#include "common.cuh"

namespace aum {

// step_dev (optional): the step number lives in device memory (incremented by adam_count_kernel just before), so that
// a captured CUDA graph of the whole training step can be replayed - the bias corrections are then formed here instead
// of on the host.  p16 (optional): a 16-bit shadow copy of the updated parameters, written in the same pass - the
// projections' tensor-core operands for the next forward, which would otherwise cost one cast kernel per weight per step.
template <typename T16>
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n4, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
                 float inv_bc1, float inv_sqrt_bc2, float grad_scale, const int* __restrict__ step_dev,
                 T16* __restrict__ p16) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (step_dev != nullptr) {
    const double t = (double)__ldg(step_dev);
    inv_bc1 = (float)(1.0 / (1.0 - pow((double)beta1, t)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, t)));
  }
  auto upd = [&](float& pv, float gv, float& mv, float& vv) {
    gv = fmaf(wd, pv, gv * grad_scale);
    mv = fmaf(beta1, mv, (1.f - beta1) * gv);
    vv = fmaf(beta2, vv, (1.f - beta2) * gv * gv);
    const float denom = fmaf(sqrtf(vv), inv_sqrt_bc2, eps);
    pv -= lr * inv_bc1 * mv / denom;
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    upd(pv.x, gv.x, mv.x, vv.x); upd(pv.y, gv.y, mv.y, vv.y); upd(pv.z, gv.z, mv.z, vv.z); upd(pv.w, gv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    if (p16 != nullptr) {
      T16 h[4] = {from_f<T16>(pv.x), from_f<T16>(pv.y), from_f<T16>(pv.z), from_f<T16>(pv.w)};
      reinterpret_cast<uint2*>(p16)[i] = *reinterpret_cast<const uint2*>(h);
    }
  }
  // tail (n not a multiple of 4)
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    upd(p[i], g[i], m[i], v[i]);
    if (p16 != nullptr) p16[i] = from_f<T16>(p[i]);
  }
}

__global__ void adam_count_kernel(int* step_dev) { *step_dev += 1; }

static int adam_launch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float inv_bc1, float inv_sqrt_bc2, float grad_scale, const int* step_dev,
                       void* p16, int p16_dtype, cudaStream_t st, const char* who) {
  const int64_t n4 = n / 4;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = ceil_div64(n4 > 0 ? n4 : n, 256);
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;      // grid-stride: 16 blocks of 256 threads per SM
  if (p16 != nullptr && p16_dtype == AUM_F16)
    adam_step_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n4, n, lr, beta1, beta2, eps, weight_decay, inv_bc1,
                                                               inv_sqrt_bc2, grad_scale, step_dev, reinterpret_cast<__half*>(p16));
  else
    adam_step_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n4, n, lr, beta1, beta2, eps, weight_decay, inv_bc1,
                                                                      inv_sqrt_bc2, grad_scale, step_dev,
                                                                      reinterpret_cast<__nv_bfloat16*>(p16));
  return check_launch(who);
}

}  // namespace aum

extern "C" int aum_adam_step(float* p, const float* g, float* m, float* v, int64_t n,
                             float lr, float beta1, float beta2, float eps, float weight_decay,
                             int step, float grad_scale, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(p);
  if (n == 0) return 0;
  AUM_REQUIRE(p && g && m && v, "aum_adam_step: null pointer");
  AUM_REQUIRE(n > 0 && step >= 1, "aum_adam_step: bad size / step (steps count from 1)");
  AUM_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "aum_adam_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  return adam_launch(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), grad_scale,
                     nullptr, nullptr, AUM_BF16, (cudaStream_t)stream, "aum_adam_step");
}

extern "C" int aum_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n,
                                 float lr, float beta1, float beta2, float eps, float weight_decay,
                                 int* step_dev, float grad_scale, void* p16, int p16_dtype, void* stream) {
  using namespace aum;
  DeviceGuard device_guard(p);
  if (n == 0) return 0;
  AUM_REQUIRE(p && g && m && v && step_dev, "aum_adam_step_dev: null pointer");
  AUM_REQUIRE(n > 0, "aum_adam_step_dev: bad size");
  AUM_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "aum_adam_step_dev: buffers must be 16-byte aligned");
  AUM_REQUIRE(p16 == nullptr || ((p16_dtype == AUM_F16 || p16_dtype == AUM_BF16) && (reinterpret_cast<uintptr_t>(p16) & 7) == 0),
              "aum_adam_step_dev: the shadow copy must be fp16 / bf16 and 8-byte aligned");
  adam_count_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  if (int rc = check_launch("aum_adam_step_dev(count)")) return rc;
  return adam_launch(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, grad_scale, step_dev, p16, p16_dtype,
                     (cudaStream_t)stream, "aum_adam_step_dev");
}
