// Fused global-norm clip + Adam(amsgrad) over flat fp32 parameter / gradient / state buffers.
// Replaces clip_grad_norm_ + torch.optim.Adam(amsgrad=True).step() (/root/reference/simvg/apis/train.py:81-83,
// /root/reference/simvg/core/optimizer.py:52-68; hyper-parameters configs/single/ViT-base/refcoco/refcoco_onestage.py:107-123),
// i.e. hundreds of small eager kernels, by two streaming passes: sum of squares, then the update (36 B / parameter).
// The exponential moving average of the weights (/root/reference/simvg/models/utils.py:148-173, called after every optimiser
// step at apis/train.py:85-86: a Python loop over the whole state dict) rides along in the update pass when enabled.
#include "common.cuh"
#include "simvg_b200.h"

namespace simvgb {

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float acc = 0.f;
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[n4 * 4 + threadIdx.x];
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

__global__ void adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, float* __restrict__ vmax, long long n, float lr, float beta1,
                                    float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                    const float* __restrict__ sumsq, float max_norm, const float* __restrict__ hyper,
                                    float* __restrict__ ema, float ema_decay) {
  if (hyper != nullptr) {   // step-dependent scalars read from the device: the launch can live inside a CUDA graph
    lr = __ldg(hyper);
    bc1 = __ldg(hyper + 1);
    bc2_sqrt = __ldg(hyper + 2);
    ema_decay = __ldg(hyper + 3);
  }
  float coef = 1.0f;
  if (sumsq != nullptr && max_norm > 0.f) coef = fminf(1.0f, max_norm / (sqrtf(__ldg(sumsq)) + 1e-6f));
  const float step = lr / bc1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 xx = reinterpret_cast<float4*>(vmax)[i];
#define SIMVGB_ADAM1(c)                                          \
    {                                                            \
      float gr = gg.c * coef + wd * pp.c;                        \
      mm.c = beta1 * mm.c + (1.f - beta1) * gr;                  \
      vv.c = beta2 * vv.c + (1.f - beta2) * gr * gr;             \
      xx.c = fmaxf(xx.c, vv.c);                                  \
      pp.c -= step * mm.c / (sqrtf(xx.c) / bc2_sqrt + eps);      \
    }
    SIMVGB_ADAM1(x) SIMVGB_ADAM1(y) SIMVGB_ADAM1(z) SIMVGB_ADAM1(w)
    reinterpret_cast<float4*>(p)[i] = pp;
    if (ema != nullptr) {   // ExponentialMovingAverage.update_params fused in: shadow = d * shadow + (1 - d) * p_new  (+8 B / parameter)
      float4 ee = reinterpret_cast<float4*>(ema)[i];
      ee.x = ema_decay * ee.x + (1.f - ema_decay) * pp.x;
      ee.y = ema_decay * ee.y + (1.f - ema_decay) * pp.y;
      ee.z = ema_decay * ee.z + (1.f - ema_decay) * pp.z;
      ee.w = ema_decay * ee.w + (1.f - ema_decay) * pp.w;
      reinterpret_cast<float4*>(ema)[i] = ee;
    }
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(vmax)[i] = xx;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    float gr = g[i] * coef + wd * p[i];
    m[i] = beta1 * m[i] + (1.f - beta1) * gr;
    v[i] = beta2 * v[i] + (1.f - beta2) * gr * gr;
    vmax[i] = fmaxf(vmax[i], v[i]);
    p[i] -= step * m[i] / (sqrtf(vmax[i]) / bc2_sqrt + eps);
    if (ema != nullptr) ema[i] = ema_decay * ema[i] + (1.f - ema_decay) * p[i];
  }
}

}  // namespace simvgb

using namespace simvgb;

extern "C" int simvgb_sumsq(const float* g, int64_t n, float* out, void* stream) {
  SIMVGB_CHECK(g && out, "simvgb_sumsq: null pointer");
  SIMVGB_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, "simvgb_sumsq: buffer must be 16-byte aligned");
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, n, out);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_adam_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, int64_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int step,
                                   const float* grad_sumsq, float max_norm, float* ema, float ema_decay, void* stream) {
  SIMVGB_CHECK(p && g && m && v && vmax, "simvgb_adam_amsgrad: null pointer");
  SIMVGB_CHECK(step >= 1, "simvgb_adam_amsgrad: step must be >= 1");
  SIMVGB_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vmax)) & 15) == 0,
               "simvgb_adam_amsgrad: buffers must be 16-byte aligned");
  if (n <= 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_amsgrad_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, vmax, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_sumsq, max_norm, nullptr, ema, ema_decay);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_adam_amsgrad_dev(float* p, const float* g, float* m, float* v, float* vmax, int64_t n,
                                       const float* hyper, float beta1, float beta2, float eps, float weight_decay,
                                       const float* grad_sumsq, float max_norm, float* ema, void* stream) {
  SIMVGB_CHECK(p && g && m && v && vmax && hyper, "simvgb_adam_amsgrad_dev: null pointer");
  SIMVGB_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vmax)) & 15) == 0,
               "simvgb_adam_amsgrad_dev: buffers must be 16-byte aligned");
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_amsgrad_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, vmax, n, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, grad_sumsq, max_norm, hyper, ema, 0.f);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}
