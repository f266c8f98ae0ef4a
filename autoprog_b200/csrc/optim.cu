// Fused AdamW step + up to 8 EMA shadow updates + bf16 compute copy, one pass over a flat parameter buffer.
//
// Replaces per step (main_prog.py:1019-1033): optimizer.step() [torch.optim.AdamW via timm create_optimizer] followed
// by `for ema in model_ema_list: ema.update(model)` [timm ModelEmaV2: e = d*e + (1-d)*p], i.e. 1 + 4 separate passes
// over 26.6 M parameters, with a single HBM pass: read p,g,m,v,e_1..e_k ; write p,m,v,e_1..e_k (+ bf16 p).
// The step-dependent scalars (lr, bias corrections) are read from DEVICE memory so a captured CUDA graph replays
// with fresh values.
#include "common.cuh"

namespace {

constexpr int MAX_EMA = 8;
struct EmaArgs {
  float* ptr[MAX_EMA];
  float decay[MAX_EMA];
  int n;
};

// hyper = {lr, bias_correction1, sqrt(bias_correction2)}
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        const float* __restrict__ hyper, float b1, float b2, float eps,
                                                        float wd, EmaArgs ema, bf16* __restrict__ shadow) {
  const float lr = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2];
  const float step_size = lr / bc1, decay = 1.f - lr * wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * decay;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
#pragma unroll
    for (int k = 0; k < MAX_EMA; ++k)
      if (k < ema.n) ema.ptr[k][i] = ema.decay[k] * ema.ptr[k][i] + (1.f - ema.decay[k]) * pi;
    if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pi);
  }
}

}  // namespace

int apb_adamw_ema(float* p, const float* g, float* m, float* v, long long n, const float* hyper_dev, float beta1,
                  float beta2, float eps, float weight_decay, float* const* ema_ptrs_host, const float* decay_host,
                  int n_ema, void* shadow_bf16, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(n_ema >= 0 && n_ema <= MAX_EMA, APB_ERR_ARG, "adamw_ema: n_ema=%d (max %d)", n_ema, MAX_EMA);
  APB_CHECK_ARG(hyper_dev != nullptr, APB_ERR_ARG, "adamw_ema: hyper_dev is NULL");
  if (n <= 0) return 0;
  EmaArgs e;
  e.n = n_ema;
  for (int k = 0; k < MAX_EMA; ++k) {
    e.ptr[k] = k < n_ema ? ema_ptrs_host[k] : nullptr;
    e.decay[k] = k < n_ema ? decay_host[k] : 0.f;
  }
  long long grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  adamw_ema_kernel<<<(int)grid, 256, 0, st>>>(p, g, m, v, n, hyper_dev, beta1, beta2, eps, weight_decay, e,
                                              (bf16*)shadow_bf16);
  APB_LAUNCH_CHECK("adamw_ema");
  return 0;
}
