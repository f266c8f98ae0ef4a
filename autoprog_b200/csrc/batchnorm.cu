// Train-mode BatchNorm2d + ReLU on channels-last activations, forward and backward (the PatchEmbed stem,
// models/volo.py:355-368: conv -> BatchNorm2d -> ReLU x3).  The convolutions stay in cuDNN; ATen's BN path costs six
// passes per layer (statistics, transform, clamp, relu-backward, backward-reduce, backward-elementwise) over a
// [B*112*112, 64] bf16 tensor -- here it is two passes forward (column statistics; normalise+ReLU) and two backward.
//
//   x, y, dy, dx : [rows, C] (NHWC flattened; bf16 or fp32), C % 8 == 0
//   stats        : fp32 {mean[C], invstd[C]} saved for backward; running stats updated like nn.BatchNorm2d
//                  (momentum m: running = (1-m)*running + m*batch, unbiased variance for running_var)
#include "common.cuh"

namespace {

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) { const float2 f = __bfloat1622float2(h[t]); v[2 * t] = f.x; v[2 * t + 1] = f.y; }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Column partial sums of two quantities.  MODE 0: (x, x^2).  MODE 1 (backward): g = dy * (y > 0); (g, g * xhat).
// The ReLU mask is recomputed from x with the forward's own expression (fma(x, a, b) > 0, a = invstd * gamma,
// b = beta - mean * a) when beta is given, which saves the read of y (a third of the backward's traffic); y is read
// only when beta == nullptr.
// block = 256 threads: lane -> 8 consecutive channels, (C/8) lanes cover a row, the rest of the block strides rows.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) bn_colstats_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                          const T* __restrict__ dy, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, long long rows, int C,
                                                          float* __restrict__ part, int rows_per_cta) {
  extern __shared__ float sm[];          // [groups][2][C]
  const int lanes_per_row = C >> 3;
  const int groups = 256 / lanes_per_row;                 // row slots per sweep
  const int slot = threadIdx.x / lanes_per_row, c0 = (threadIdx.x % lanes_per_row) * 8;
  const bool active = slot < groups;
  float a0[8], a1[8], mu[8], is[8], ma[8], mb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; mu[j] = 0.f; is[j] = 0.f; ma[j] = 0.f; mb[j] = 0.f; }
  const bool remask = MODE == 1 && beta != nullptr;
  if (MODE == 1 && active) {
    ld8(mean + c0, mu); ld8(invstd + c0, is);
    if (remask) {
      float ga[8], be[8];
      ld8(gamma + c0, ga); ld8(beta + c0, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) { ma[j] = is[j] * ga[j]; mb[j] = be[j] - mu[j] * ma[j]; }
    }
  }
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + (long long)rows_per_cta);
  if (active) {
    // UNR rows in flight per thread (all loads issued before the first use): the kernel is a pure HBM stream
    constexpr int UNR = (MODE == 0) ? 4 : 2;
    long long r = r0 + slot;
    for (; r + (long long)(UNR - 1) * groups < r1; r += (long long)UNR * groups) {
      float xv[UNR][8], yv[UNR][8], gv[UNR][8];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const size_t o = (size_t)(r + (long long)u * groups) * C + c0;
        ld8(x + o, xv[u]);
        if (MODE == 1) {
          ld8(dy + o, gv[u]);
          if (!remask) ld8(y + o, yv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) {
            a0[j] += xv[u][j];
            a1[j] = fmaf(xv[u][j], xv[u][j], a1[j]);
          } else {
            const float act = remask ? fmaf(xv[u][j], ma[j], mb[j]) : yv[u][j];
            const float g = act > 0.f ? gv[u][j] : 0.f;
            a0[j] += g;
            a1[j] = fmaf(g, (xv[u][j] - mu[j]) * is[j], a1[j]);
          }
        }
    }
    for (; r < r1; r += groups) {
      float xv[8];
      ld8(x + (size_t)r * C + c0, xv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += xv[j]; a1[j] = fmaf(xv[j], xv[j], a1[j]); }
      } else {
        float yv[8], gv[8];
        if (!remask) ld8(y + (size_t)r * C + c0, yv);
        ld8(dy + (size_t)r * C + c0, gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float act = remask ? fmaf(xv[j], ma[j], mb[j]) : yv[j];
          const float g = act > 0.f ? gv[j] : 0.f;
          a0[j] += g;
          a1[j] = fmaf(g, (xv[j] - mu[j]) * is[j], a1[j]);
        }
      }
    }
    float* s = sm + (size_t)slot * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[c0 + j] = a0[j]; s[C + c0 + j] = a1[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += sm[(size_t)g * 2 * C + c];
    part[(size_t)blockIdx.x * 2 * C + c] = t;
  }
}

// finalize forward statistics: mean, invstd; update running stats.  One warp per channel (fixed-order tree).
__global__ void __launch_bounds__(256) bn_finalize_fwd_kernel(const float* __restrict__ part, int nparts, int C, long long rows,
                                                              float eps, float momentum, float* __restrict__ mean,
                                                              float* __restrict__ invstd, float* __restrict__ running_mean,
                                                              float* __restrict__ running_var) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, ss = 0.0;
  for (int p = lane; p < nparts; p += 32) { s += (double)part[(size_t)p * 2 * C + c]; ss += (double)part[(size_t)p * 2 * C + C + c]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  if (lane != 0) return;
  const double m = s / (double)rows;
  double var = ss / (double)rows - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// finalize backward sums: dbeta = sum g, dgamma = sum g*xhat (one warp per channel, fixed order)
__global__ void __launch_bounds__(256) bn_finalize_bwd_kernel(const float* __restrict__ part, int nparts, int C,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float s = 0.f, ss = 0.f;
  for (int p = lane; p < nparts; p += 32) { s += part[(size_t)p * 2 * C + c]; ss += part[(size_t)p * 2 * C + C + c]; }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane == 0) { dbeta[c] = s; dgamma[c] = ss; }
}

// y = relu((x - mean) * invstd * gamma + beta)
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, T* __restrict__ y,
                                                            long long rows, int C) {
  // the grid stride (gridDim * 256 vectors) is a multiple of C/8, so a thread always works on the same 8 channels:
  // their scale / shift are computed once and the loop is a pure 16-byte load -> fma -> relu -> store stream
  const long long nvec = rows * (C >> 3);
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(i0 % (C >> 3)) * 8;
  float a[8], b[8];
  {
    float mu[8], is[8], ga[8], be[8];
    ld8(mean + c0, mu); ld8(invstd + c0, is); ld8(gamma + c0, ga); ld8(beta + c0, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = is[j] * ga[j]; b[j] = be[j] - mu[j] * a[j]; }
  }
  const long long step = (long long)gridDim.x * blockDim.x;
  long long i = i0;
  for (; i + step < nvec; i += 2 * step) {       // two vectors in flight
    float x0[8], x1[8];
    ld8(x + i * 8, x0);
    ld8(x + (i + step) * 8, x1);
#pragma unroll
    for (int j = 0; j < 8; ++j) { x0[j] = fmaxf(fmaf(x0[j], a[j], b[j]), 0.f); x1[j] = fmaxf(fmaf(x1[j], a[j], b[j]), 0.f); }
    st8(y + i * 8, x0);
    st8(y + (i + step) * 8, x1);
  }
  if (i < nvec) {
    float x0[8];
    ld8(x + i * 8, x0);
#pragma unroll
    for (int j = 0; j < 8; ++j) x0[j] = fmaxf(fmaf(x0[j], a[j], b[j]), 0.f);
    st8(y + i * 8, x0);
  }
}

// dx = gamma * invstd * (g - dbeta/rows - xhat * dgamma/rows),  g = dy * (y > 0)
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                           const T* __restrict__ dy, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ dgamma,
                                                           const float* __restrict__ dbeta, T* __restrict__ dx,
                                                           long long rows, int C) {
  // fixed 8 channels per thread (see bn_apply_relu_kernel): dx = k1 * g - k2 * x + k3 with per-channel constants
  //   k1 = gamma*invstd, k2 = k1 * invstd * dgamma / n, k3 = k2 * mean - k1 * dbeta / n
  const long long nvec = rows * (C >> 3);
  const float inv_n = 1.f / (float)rows;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(i0 % (C >> 3)) * 8;
  float k1[8], k2[8], k3[8], mb[8];
  const bool remask = beta != nullptr;           // ReLU mask from x (the forward's fma) instead of a read of y
  {
    float mu[8], is[8], ga[8], dg[8], db[8];
    ld8(mean + c0, mu); ld8(invstd + c0, is); ld8(gamma + c0, ga); ld8(dgamma + c0, dg); ld8(dbeta + c0, db);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k1[j] = ga[j] * is[j];                     // == the forward's scale a = invstd * gamma
      k2[j] = k1[j] * is[j] * dg[j] * inv_n;
      k3[j] = k2[j] * mu[j] - k1[j] * db[j] * inv_n;
      mb[j] = 0.f;
    }
    if (remask) {
      float be[8];
      ld8(beta + c0, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) mb[j] = be[j] - mu[j] * (is[j] * ga[j]);
    }
  }
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < nvec; i += step) {
    float xv[8], yv[8], gv[8];
    ld8(x + i * 8, xv); ld8(dy + i * 8, gv);
    if (!remask) ld8(y + i * 8, yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float act = remask ? fmaf(xv[j], k1[j], mb[j]) : yv[j];
      const float g = act > 0.f ? gv[j] : 0.f;
      xv[j] = fmaf(k1[j], g, fmaf(-k2[j], xv[j], k3[j]));
    }
    st8(dx + i * 8, xv);
  }
}

constexpr int BN_ROWS_PER_CTA = 2048;
inline int bn_parts(long long rows) { return (int)((rows + BN_ROWS_PER_CTA - 1) / BN_ROWS_PER_CTA); }
inline int bn_grid(long long nvec) {
  long long g = (nvec + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

long long apb_bn_workspace_floats(long long rows, int C) { return (long long)bn_parts(rows) * 2 * C; }

// training forward: batch statistics -> mean/invstd (saved), running stats updated (if non-NULL), y = relu(bn(x)).
// use_batch_stats = 0 (eval): mean/invstd must already hold running_mean and 1/sqrt(running_var + eps); only the apply runs.
int apb_bn_relu_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* invstd,
                    float* running_mean, float* running_var, float momentum, float eps, int use_batch_stats,
                    float* workspace, long long rows, int C, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(rows > 0 && C > 0 && (C % 8) == 0 && C <= 2048 && 256 % (C / 8) == 0, APB_ERR_SHAPE,
                "bn_relu_fwd: rows=%lld C=%d (C must be a multiple of 8 with 256 %% (C/8) == 0)", rows, C);
  APB_CHECK_ARG(dtype == APB_F32 || dtype == APB_BF16, APB_ERR_DTYPE, "bn_relu_fwd: dtype %d", dtype);
  const int parts = bn_parts(rows);
  const int groups = 256 / (C / 8);
  const size_t smem = (size_t)groups * 2 * C * sizeof(float);
  if (use_batch_stats) {
    if (dtype == APB_F32) {
      cudaFuncSetAttribute(bn_colstats_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      bn_colstats_kernel<float, 0><<<parts, 256, smem, st>>>((const float*)x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rows, C, workspace, BN_ROWS_PER_CTA);
    } else {
      cudaFuncSetAttribute(bn_colstats_kernel<bf16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      bn_colstats_kernel<bf16, 0><<<parts, 256, smem, st>>>((const bf16*)x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rows, C, workspace, BN_ROWS_PER_CTA);
    }
    APB_LAUNCH_CHECK("bn_colstats");
    bn_finalize_fwd_kernel<<<ceil_div(C, 8), 256, 0, st>>>(workspace, parts, C, rows, eps, momentum, mean, invstd, running_mean, running_var);
    APB_LAUNCH_CHECK("bn_finalize_fwd");
  }
  const long long nvec = rows * (C / 8);
  if (dtype == APB_F32) bn_apply_relu_kernel<float><<<bn_grid(nvec), 256, 0, st>>>((const float*)x, mean, invstd, gamma, beta, (float*)y, rows, C);
  else bn_apply_relu_kernel<bf16><<<bn_grid(nvec), 256, 0, st>>>((const bf16*)x, mean, invstd, gamma, beta, (bf16*)y, rows, C);
  APB_LAUNCH_CHECK("bn_apply_relu");
  return 0;
}

int apb_bn_relu_bwd(const void* x, const void* y, const void* dy, const float* gamma, const float* beta, const float* mean,
                    const float* invstd, void* dx, float* dgamma, float* dbeta, float* workspace, long long rows, int C, int dtype,
                    apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(rows > 0 && C > 0 && (C % 8) == 0 && C <= 2048 && 256 % (C / 8) == 0, APB_ERR_SHAPE, "bn_relu_bwd: rows=%lld C=%d", rows, C);
  APB_CHECK_ARG(dtype == APB_F32 || dtype == APB_BF16, APB_ERR_DTYPE, "bn_relu_bwd: dtype %d", dtype);
  APB_CHECK_ARG(y != nullptr || beta != nullptr, APB_ERR_ARG, "bn_relu_bwd: needs y or beta for the ReLU mask");
  const int parts = bn_parts(rows);
  const int groups = 256 / (C / 8);
  const size_t smem = (size_t)groups * 2 * C * sizeof(float);
  const long long nvec = rows * (C / 8);
  if (dtype == APB_F32) {
    cudaFuncSetAttribute(bn_colstats_kernel<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bn_colstats_kernel<float, 1><<<parts, 256, smem, st>>>((const float*)x, (const float*)y, (const float*)dy, mean, invstd, gamma, beta, rows, C, workspace, BN_ROWS_PER_CTA);
    APB_LAUNCH_CHECK("bn_colstats_bwd");
    bn_finalize_bwd_kernel<<<ceil_div(C, 8), 256, 0, st>>>(workspace, parts, C, dgamma, dbeta);
    bn_bwd_apply_kernel<float><<<bn_grid(nvec), 256, 0, st>>>((const float*)x, (const float*)y, (const float*)dy, mean, invstd, gamma, beta, dgamma, dbeta, (float*)dx, rows, C);
  } else {
    cudaFuncSetAttribute(bn_colstats_kernel<bf16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bn_colstats_kernel<bf16, 1><<<parts, 256, smem, st>>>((const bf16*)x, (const bf16*)y, (const bf16*)dy, mean, invstd, gamma, beta, rows, C, workspace, BN_ROWS_PER_CTA);
    APB_LAUNCH_CHECK("bn_colstats_bwd");
    bn_finalize_bwd_kernel<<<ceil_div(C, 8), 256, 0, st>>>(workspace, parts, C, dgamma, dbeta);
    bn_bwd_apply_kernel<bf16><<<bn_grid(nvec), 256, 0, st>>>((const bf16*)x, (const bf16*)y, (const bf16*)dy, mean, invstd, gamma, beta, dgamma, dbeta, (bf16*)dx, rows, C);
  }
  APB_LAUNCH_CHECK("bn_bwd_apply");
  return 0;
}
