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
__device__ __forceinline__ float adamw_one(float p, float g, float& m, float& v, float b1, float b2, float eps, float decay,
                                           float step_size, float bc2_sqrt) {
  float pi = p * decay;
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  return pi - step_size * (m / denom);
}

// VEC = 4: 16-byte accesses on all 4 + 2*n_ema fp32 streams (flat buffers are padded to multiples of 8 elements)
template <int VEC>
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        const float* __restrict__ hyper, float b1, float b2, float eps,
                                                        float wd, EmaArgs ema, bf16* __restrict__ shadow) {
  const float lr = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2];
  const float step_size = lr / bc1, decay = 1.f - lr * wd;
  const long long nv = n / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    if constexpr (VEC == 4) {
      const float4 g4 = reinterpret_cast<const float4*>(g)[i];
      float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
      float4 e4[MAX_EMA];
#pragma unroll
      for (int k = 0; k < MAX_EMA; ++k)
        if (k < ema.n) e4[k] = reinterpret_cast<float4*>(ema.ptr[k])[i];
      p4.x = adamw_one(p4.x, g4.x, m4.x, v4.x, b1, b2, eps, decay, step_size, bc2_sqrt);
      p4.y = adamw_one(p4.y, g4.y, m4.y, v4.y, b1, b2, eps, decay, step_size, bc2_sqrt);
      p4.z = adamw_one(p4.z, g4.z, m4.z, v4.z, b1, b2, eps, decay, step_size, bc2_sqrt);
      p4.w = adamw_one(p4.w, g4.w, m4.w, v4.w, b1, b2, eps, decay, step_size, bc2_sqrt);
      reinterpret_cast<float4*>(p)[i] = p4;
      reinterpret_cast<float4*>(m)[i] = m4;
      reinterpret_cast<float4*>(v)[i] = v4;
#pragma unroll
      for (int k = 0; k < MAX_EMA; ++k)
        if (k < ema.n) {
          const float d = ema.decay[k], c = 1.f - d;
          float4 e = e4[k];
          e.x = d * e.x + c * p4.x; e.y = d * e.y + c * p4.y; e.z = d * e.z + c * p4.z; e.w = d * e.w + c * p4.w;
          reinterpret_cast<float4*>(ema.ptr[k])[i] = e;
        }
      if (shadow != nullptr) {
        uint2 pk;
        *reinterpret_cast<__nv_bfloat162*>(&pk.x) = __floats2bfloat162_rn(p4.x, p4.y);
        *reinterpret_cast<__nv_bfloat162*>(&pk.y) = __floats2bfloat162_rn(p4.z, p4.w);
        reinterpret_cast<uint2*>(shadow)[i] = pk;
      }
    } else {
      float mi = m[i], vi = v[i];
      const float pi = adamw_one(p[i], g[i], mi, vi, b1, b2, eps, decay, step_size, bc2_sqrt);
      p[i] = pi;
      m[i] = mi;
      v[i] = vi;
#pragma unroll
      for (int k = 0; k < MAX_EMA; ++k)
        if (k < ema.n) ema.ptr[k][i] = ema.decay[k] * ema.ptr[k][i] + (1.f - ema.decay[k]) * pi;
      if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pi);
    }
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
  uintptr_t al = (uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v;
  for (int k = 0; k < n_ema; ++k) al |= (uintptr_t)ema_ptrs_host[k];
  const bool vec = (n % 4 == 0) && (al & 15) == 0 && ((uintptr_t)shadow_bf16 & 7) == 0;
  const long long work = vec ? n / 4 : n;
  long long grid = (work + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  if (vec)
    adamw_ema_kernel<4><<<(int)grid, 256, 0, st>>>(p, g, m, v, n, hyper_dev, beta1, beta2, eps, weight_decay, e, (bf16*)shadow_bf16);
  else
    adamw_ema_kernel<1><<<(int)grid, 256, 0, st>>>(p, g, m, v, n, hyper_dev, beta1, beta2, eps, weight_decay, e, (bf16*)shadow_bf16);
  APB_LAUNCH_CHECK("adamw_ema");
  return 0;
}
