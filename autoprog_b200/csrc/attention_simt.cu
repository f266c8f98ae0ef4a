// Multi-head self-attention core and class-attention core, CUDA-core fp32 path (parity mode, any head dim <= 64).
//
// mhsa  : models/volo.py:188-197   attn = softmax(q k^T * scale); out = attn @ v      (scores never hit HBM)
// class : models/volo.py:264-275   cls query (scaled) against all tokens
// Layouts: qkv [B,N,3,heads,D]; out [B,N,heads,D]; lse [B,heads,N]; kv [B,N,2,heads,D]; q/out(class) [B,heads,D].
#include "common.cuh"

namespace {

constexpr int AW = 8;         // warps per CTA
constexpr int MAXT = 32;      // key tiles of 32 per lane -> N <= 1024

// K and V of one (b, head) staged as fp32 with row stride D+1; one warp per query row.
template <typename T>
__global__ void __launch_bounds__(AW * 32) mhsa_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                           float* __restrict__ lse, int B, int N, int heads, int D,
                                                           float scale, int q_per_cta) {
  extern __shared__ float sm[];
  const int DS = D + 1;
  float* sK = sm;                       // [N][DS]
  float* sV = sK + (size_t)N * DS;      // [N][DS]
  float* sQ = sV + (size_t)N * DS;      // [AW][D]
  float* sP = sQ + AW * D;              // [AW][N]
  const int bh = blockIdx.y, b = bh / heads, hd = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t tok = (size_t)3 * heads * D;
  const T* base = qkv + (size_t)b * N * tok + (size_t)hd * D;
  for (int e = threadIdx.x; e < N * D; e += blockDim.x) {
    const int n = e / D, c = e % D;
    sK[n * DS + c] = to_f(base[(size_t)n * tok + (size_t)heads * D + c]);
    sV[n * DS + c] = to_f(base[(size_t)n * tok + (size_t)2 * heads * D + c]);
  }
  __syncthreads();
  const int q0 = blockIdx.x * q_per_cta;
  const int q1 = min(N, q0 + q_per_cta);
  float* myQ = sQ + warp * D;
  float* myP = sP + (size_t)warp * N;
  for (int i = q0 + warp; i < q1; i += AW) {
    for (int c = lane; c < D; c += 32) myQ[c] = to_f(base[(size_t)i * tok + c]) * scale;
    __syncwarp();
    float s[MAXT];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int k = t * 32 + lane;
      if (t * 32 < N) {
        float a = -INFINITY;
        if (k < N) {
          a = 0.f;
          for (int c = 0; c < D; ++c) a = fmaf(myQ[c], sK[k * DS + c], a);
        }
        s[t] = a;
        m = fmaxf(m, a);
      }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t * 32 < N) {
        const int k = t * 32 + lane;
        const float e = (k < N) ? expf(s[t] - m) : 0.f;
        s[t] = e;
        sum += e;
      }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t * 32 < N) {
        const int k = t * 32 + lane;
        if (k < N) myP[k] = s[t] * inv;
      }
    __syncwarp();
    for (int c = lane; c < D; c += 32) {
      float o = 0.f;
      for (int k = 0; k < N; ++k) o = fmaf(myP[k], sV[k * DS + c], o);
      out[((size_t)b * N + i) * heads * D + (size_t)hd * D + c] = from_f<T>(o);
    }
    if (lane == 0) lse[((size_t)b * heads + hd) * N + i] = m + logf(sum);
    __syncwarp();
  }
}

// dQ + row dots Drow[i] = <dO_i, O_i>.  One warp per query row, K/V staged in smem.
template <typename T>
__global__ void __launch_bounds__(AW * 32) mhsa_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ out,
                                                              const T* __restrict__ dout, const float* __restrict__ lse,
                                                              T* __restrict__ dqkv, float* __restrict__ drow, int B,
                                                              int N, int heads, int D, float scale, int q_per_cta) {
  extern __shared__ float sm[];
  const int DS = D + 1;
  float* sK = sm;
  float* sV = sK + (size_t)N * DS;
  float* sQ = sV + (size_t)N * DS;      // [AW][2*D] : q*scale, dO
  float* sP = sQ + AW * 2 * D;          // [AW][N]   : dS
  const int bh = blockIdx.y, b = bh / heads, hd = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t tok = (size_t)3 * heads * D;
  const T* base = qkv + (size_t)b * N * tok + (size_t)hd * D;
  for (int e = threadIdx.x; e < N * D; e += blockDim.x) {
    const int n = e / D, c = e % D;
    sK[n * DS + c] = to_f(base[(size_t)n * tok + (size_t)heads * D + c]);
    sV[n * DS + c] = to_f(base[(size_t)n * tok + (size_t)2 * heads * D + c]);
  }
  __syncthreads();
  const int q0 = blockIdx.x * q_per_cta;
  const int q1 = min(N, q0 + q_per_cta);
  float* myQ = sQ + warp * 2 * D;
  float* myG = myQ + D;
  float* myP = sP + (size_t)warp * N;
  for (int i = q0 + warp; i < q1; i += AW) {
    const size_t orow = ((size_t)b * N + i) * heads * D + (size_t)hd * D;
    float dd = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float g = to_f(dout[orow + c]);
      myQ[c] = to_f(base[(size_t)i * tok + c]) * scale;
      myG[c] = g;
      dd = fmaf(g, to_f(out[orow + c]), dd);
    }
    dd = warp_sum(dd);
    __syncwarp();
    const float L = lse[((size_t)b * heads + hd) * N + i];
    for (int k = lane; k < N; k += 32) {
      float a = 0.f, dp = 0.f;
      for (int c = 0; c < D; ++c) {
        a = fmaf(myQ[c], sK[k * DS + c], a);
        dp = fmaf(myG[c], sV[k * DS + c], dp);
      }
      const float p = expf(a - L);
      myP[k] = p * (dp - dd) * scale;
    }
    __syncwarp();
    for (int c = lane; c < D; c += 32) {
      float o = 0.f;
      for (int k = 0; k < N; ++k) o = fmaf(myP[k], sK[k * DS + c], o);
      dqkv[((size_t)b * N + i) * tok + (size_t)hd * D + c] = from_f<T>(o);
    }
    if (lane == 0) drow[((size_t)b * heads + hd) * N + i] = dd;
    __syncwarp();
  }
}

// dK, dV: one warp per key row; Q*scale and dO of the (b, head) staged in smem.
template <typename T>
__global__ void __launch_bounds__(AW * 32) mhsa_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                               const float* __restrict__ lse,
                                                               const float* __restrict__ drow, T* __restrict__ dqkv,
                                                               int B, int N, int heads, int D, float scale,
                                                               int k_per_cta) {
  extern __shared__ float sm[];
  const int DS = D + 1;
  float* sQ = sm;                        // [N][DS]  q * scale
  float* sG = sQ + (size_t)N * DS;       // [N][DS]  dO
  float* sL = sG + (size_t)N * DS;       // [N] lse
  float* sD = sL + N;                    // [N] row dots
  float* sKV = sD + N;                   // [AW][2*D]
  float* sP = sKV + AW * 2 * D;          // [AW][2*N] : p, dS
  const int bh = blockIdx.y, b = bh / heads, hd = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t tok = (size_t)3 * heads * D;
  const T* base = qkv + (size_t)b * N * tok + (size_t)hd * D;
  for (int e = threadIdx.x; e < N * D; e += blockDim.x) {
    const int n = e / D, c = e % D;
    sQ[n * DS + c] = to_f(base[(size_t)n * tok + c]) * scale;
    sG[n * DS + c] = to_f(dout[((size_t)b * N + n) * heads * D + (size_t)hd * D + c]);
  }
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    sL[n] = lse[((size_t)b * heads + hd) * N + n];
    sD[n] = drow[((size_t)b * heads + hd) * N + n];
  }
  __syncthreads();
  const int k0 = blockIdx.x * k_per_cta;
  const int k1 = min(N, k0 + k_per_cta);
  float* myK = sKV + warp * 2 * D;
  float* myV = myK + D;
  float* myP = sP + (size_t)warp * 2 * N;
  float* myS = myP + N;
  for (int k = k0 + warp; k < k1; k += AW) {
    for (int c = lane; c < D; c += 32) {
      myK[c] = to_f(base[(size_t)k * tok + (size_t)heads * D + c]);
      myV[c] = to_f(base[(size_t)k * tok + (size_t)2 * heads * D + c]);
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
      float a = 0.f, dp = 0.f;
      for (int c = 0; c < D; ++c) {
        a = fmaf(sQ[i * DS + c], myK[c], a);
        dp = fmaf(sG[i * DS + c], myV[c], dp);
      }
      const float p = expf(a - sL[i]);
      myP[i] = p;
      myS[i] = p * (dp - sD[i]);   // dS / (already includes q*scale in sQ below)
    }
    __syncwarp();
    for (int c = lane; c < D; c += 32) {
      float dv = 0.f, dk = 0.f;
      for (int i = 0; i < N; ++i) {
        dv = fmaf(myP[i], sG[i * DS + c], dv);
        dk = fmaf(myS[i], sQ[i * DS + c], dk);   // sQ holds q*scale -> dK = sum_i dS_i * scale * q_i
      }
      dqkv[((size_t)b * N + k) * tok + (size_t)heads * D + (size_t)hd * D + c] = from_f<T>(dk);
      dqkv[((size_t)b * N + k) * tok + (size_t)2 * heads * D + (size_t)hd * D + c] = from_f<T>(dv);
    }
    __syncwarp();
  }
}

// ---- class attention: one warp per (b, head)
template <typename T, bool BWD>
__global__ void __launch_bounds__(128) class_attn_kernel(const T* __restrict__ q, const T* __restrict__ kv,
                                                         const T* __restrict__ dout, T* __restrict__ out,
                                                         T* __restrict__ dq, T* __restrict__ dkv, int B, int N, int heads,
                                                         int D, float scale) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sP = sm + (size_t)warp * (2 * N + 2 * D);   // [N] p, [N] dS, [D] q*scale, [D] dO
  float* sS = sP + N;
  float* sQ = sS + N;
  float* sG = sQ + D;
  const int bh = blockIdx.x * 4 + warp;
  if (bh >= B * heads) return;
  const int b = bh / heads, hd = bh % heads;
  const size_t tok = (size_t)2 * heads * D;
  const T* kb = kv + (size_t)b * N * tok + (size_t)hd * D;
  const T* vb = kb + (size_t)heads * D;
  const size_t qoff = (size_t)b * heads * D + (size_t)hd * D;
  for (int c = lane; c < D; c += 32) {
    sQ[c] = to_f(q[qoff + c]) * scale;
    if (BWD) sG[c] = to_f(dout[qoff + c]);
  }
  __syncwarp();
  float m = -INFINITY;
  for (int k = lane; k < N; k += 32) {
    float a = 0.f;
    for (int c = 0; c < D; ++c) a = fmaf(sQ[c], to_f(kb[(size_t)k * tok + c]), a);
    sP[k] = a;
    m = fmaxf(m, a);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int k = lane; k < N; k += 32) { const float e = expf(sP[k] - m); sP[k] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int k = lane; k < N; k += 32) sP[k] *= inv;
  __syncwarp();
  if (!BWD) {
    for (int c = lane; c < D; c += 32) {
      float o = 0.f;
      for (int k = 0; k < N; ++k) o = fmaf(sP[k], to_f(vb[(size_t)k * tok + c]), o);
      out[qoff + c] = from_f<T>(o);
    }
  } else {
    float dsum = 0.f;
    for (int k = lane; k < N; k += 32) {
      float dp = 0.f;
      for (int c = 0; c < D; ++c) dp = fmaf(sG[c], to_f(vb[(size_t)k * tok + c]), dp);
      sS[k] = dp;
      dsum = fmaf(sP[k], dp, dsum);
    }
    dsum = warp_sum(dsum);
    for (int k = lane; k < N; k += 32) sS[k] = sP[k] * (sS[k] - dsum);
    __syncwarp();
    T* dkb = dkv + (size_t)b * N * tok + (size_t)hd * D;
    T* dvb = dkb + (size_t)heads * D;
    for (int c = lane; c < D; c += 32) {
      float a = 0.f;
      for (int k = 0; k < N; ++k) {
        a = fmaf(sS[k], to_f(kb[(size_t)k * tok + c]), a);
        dkb[(size_t)k * tok + c] = from_f<T>(sS[k] * sQ[c]);
        dvb[(size_t)k * tok + c] = from_f<T>(sP[k] * sG[c]);
      }
      dq[qoff + c] = from_f<T>(a * scale);
    }
  }
}

}  // namespace

static int mhsa_q_per_cta(int N) { return N <= 64 ? N : 64; }

int apb_mhsa_fwd_simt(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, int dtype,
                 apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && N > 0 && heads > 0 && D > 0, APB_ERR_SHAPE, "mhsa_fwd: bad shape");
  APB_CHECK_ARG(N <= 32 * MAXT, APB_ERR_UNSUPPORTED, "mhsa_fwd: N=%d > %d", N, 32 * MAXT);
  const int qpc = mhsa_q_per_cta(N);
  const size_t smem = ((size_t)2 * N * (D + 1) + AW * D + (size_t)AW * N) * sizeof(float);
  APB_CHECK_ARG(smem <= 227 * 1024, APB_ERR_UNSUPPORTED, "mhsa_fwd: N=%d D=%d needs %zu B smem", N, D, smem);
  dim3 grid(ceil_div(N, qpc), B * heads);
  if (dtype == APB_F32) {
    cudaFuncSetAttribute(mhsa_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mhsa_fwd_kernel<float><<<grid, AW * 32, smem, st>>>((const float*)qkv, (float*)out, lse, B, N, heads, D, scale, qpc);
  } else if (dtype == APB_BF16) {
    cudaFuncSetAttribute(mhsa_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mhsa_fwd_kernel<bf16><<<grid, AW * 32, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, B, N, heads, D, scale, qpc);
  } else { apb_set_error("mhsa_fwd: dtype %d", dtype); return APB_ERR_DTYPE; }
  APB_LAUNCH_CHECK("mhsa_fwd");
  return 0;
}

int apb_mhsa_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace,
                 int B, int N, int heads, int D, float scale, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && N > 0 && heads > 0 && D > 0, APB_ERR_SHAPE, "mhsa_bwd: bad shape");
  const int qpc = mhsa_q_per_cta(N);
  const size_t smem1 = ((size_t)2 * N * (D + 1) + AW * 2 * D + (size_t)AW * N) * sizeof(float);
  const size_t smem2 = ((size_t)2 * N * (D + 1) + 2 * (size_t)N + AW * 2 * D + (size_t)AW * 2 * N) * sizeof(float);
  APB_CHECK_ARG(smem2 <= 227 * 1024 && smem1 <= 227 * 1024, APB_ERR_UNSUPPORTED, "mhsa_bwd: N=%d D=%d needs %zu B smem", N, D, smem2);
  dim3 grid(ceil_div(N, qpc), B * heads);
#define MB(T_)                                                                                                        \
  do {                                                                                                                \
    cudaFuncSetAttribute(mhsa_bwd_dq_kernel<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);            \
    cudaFuncSetAttribute(mhsa_bwd_dkv_kernel<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);           \
    mhsa_bwd_dq_kernel<T_><<<grid, AW * 32, smem1, st>>>((const T_*)qkv, (const T_*)out, (const T_*)dout, lse,       \
                                                         (T_*)dqkv, workspace, B, N, heads, D, scale, qpc);           \
    mhsa_bwd_dkv_kernel<T_><<<grid, AW * 32, smem2, st>>>((const T_*)qkv, (const T_*)dout, lse, workspace, (T_*)dqkv, \
                                                          B, N, heads, D, scale, qpc);                                \
  } while (0)
  if (dtype == APB_F32) MB(float);
  else if (dtype == APB_BF16) MB(bf16);
  else { apb_set_error("mhsa_bwd: dtype %d", dtype); return APB_ERR_DTYPE; }
#undef MB
  APB_LAUNCH_CHECK("mhsa_bwd");
  return 0;
}

int apb_class_attn_fwd(const void* q, const void* kv, void* out, int B, int N, int heads, int D, float scale, int dtype,
                       apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && N > 0 && heads > 0 && D > 0, APB_ERR_SHAPE, "class_attn_fwd: bad shape");
  const size_t smem = (size_t)4 * (2 * N + 2 * D) * sizeof(float);
  APB_CHECK_ARG(smem <= 227 * 1024, APB_ERR_UNSUPPORTED, "class_attn: N=%d too large", N);
  const int grid = ceil_div((long long)B * heads, 4);
  if (dtype == APB_F32) {
    cudaFuncSetAttribute(class_attn_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    class_attn_kernel<float, false><<<grid, 128, smem, st>>>((const float*)q, (const float*)kv, nullptr, (float*)out, nullptr, nullptr, B, N, heads, D, scale);
  } else if (dtype == APB_BF16) {
    cudaFuncSetAttribute(class_attn_kernel<bf16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    class_attn_kernel<bf16, false><<<grid, 128, smem, st>>>((const bf16*)q, (const bf16*)kv, nullptr, (bf16*)out, nullptr, nullptr, B, N, heads, D, scale);
  } else { apb_set_error("class_attn_fwd: dtype %d", dtype); return APB_ERR_DTYPE; }
  APB_LAUNCH_CHECK("class_attn_fwd");
  return 0;
}

int apb_class_attn_bwd(const void* q, const void* kv, const void* dout, void* dq, void* dkv, int B, int N, int heads,
                       int D, float scale, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && N > 0 && heads > 0 && D > 0, APB_ERR_SHAPE, "class_attn_bwd: bad shape");
  const size_t smem = (size_t)4 * (2 * N + 2 * D) * sizeof(float);
  APB_CHECK_ARG(smem <= 227 * 1024, APB_ERR_UNSUPPORTED, "class_attn: N=%d too large", N);
  const int grid = ceil_div((long long)B * heads, 4);
  if (dtype == APB_F32) {
    cudaFuncSetAttribute(class_attn_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    class_attn_kernel<float, true><<<grid, 128, smem, st>>>((const float*)q, (const float*)kv, (const float*)dout, nullptr, (float*)dq, (float*)dkv, B, N, heads, D, scale);
  } else if (dtype == APB_BF16) {
    cudaFuncSetAttribute(class_attn_kernel<bf16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    class_attn_kernel<bf16, true><<<grid, 128, smem, st>>>((const bf16*)q, (const bf16*)kv, (const bf16*)dout, nullptr, (bf16*)dq, (bf16*)dkv, B, N, heads, D, scale);
  } else { apb_set_error("class_attn_bwd: dtype %d", dtype); return APB_ERR_DTYPE; }
  APB_LAUNCH_CHECK("class_attn_bwd");
  return 0;
}
