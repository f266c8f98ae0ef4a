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

// ---- class attention (models/volo.py:264-275): one 128-thread CTA per (b, head).  A thread owns keys t, t+128, ...
// for everything that is "per key" (scores, dP, the dK / dV rows: whole 2*D-byte rows moved with 16-byte accesses);
// the per-channel sums over keys (out, dq) are split over the four warps and combined through shared memory in a
// fixed order.  The previous one-warp-per-(b, head) version walked 197 keys serially with element-wise loads.
template <typename T>
__device__ __forceinline__ void ca_load_row(const T* p, int D, float* dst) {
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(p + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); dst[c + 2 * j] = f.x; dst[c + 2 * j + 1] = f.y; }
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; c += 4) {
      const float4 f = *reinterpret_cast<const float4*>(p + c);
      dst[c] = f.x; dst[c + 1] = f.y; dst[c + 2] = f.z; dst[c + 3] = f.w;
    }
  }
}
template <typename T>
__device__ __forceinline__ void ca_store_row(T* p, int D, const float* src, float mul) {
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(src[c + 2 * j] * mul, src[c + 2 * j + 1] * mul);
      *reinterpret_cast<uint4*>(p + c) = u;
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; c += 4)
      *reinterpret_cast<float4*>(p + c) = make_float4(src[c] * mul, src[c + 1] * mul, src[c + 2] * mul, src[c + 3] * mul);
  }
}
__device__ __forceinline__ float ca_block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();                                   // red[] may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return is_max ? fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) : ((red[0] + red[1]) + red[2]) + red[3];
}

constexpr int CA_MAXD = 64;
// CD = compile-time head dim (32: every VOLO variant; 64) so the per-key row lives in registers; 0 = runtime D
// kv0 / dkv0 != nullptr: SPLIT layout -- key 0 (the class token itself) lives in kv0 [B, 2C] and keys 1 .. N-1 (the patch
// tokens) in kv [B, N-1, 2C], so the caller never has to concatenate [cls ; tokens] into one buffer
template <typename T, bool BWD, int CD>
__global__ void __launch_bounds__(128) class_attn_kernel(const T* __restrict__ q, const T* __restrict__ kv,
                                                         const T* __restrict__ dout, T* __restrict__ out,
                                                         T* __restrict__ dq, T* __restrict__ dkv, int B, int N, int heads,
                                                         int Drt, float scale, const T* __restrict__ kv0,
                                                         T* __restrict__ dkv0) {
  const int D = CD ? CD : Drt;
  extern __shared__ float sm[];
  float* sP = sm;                 // [N] probabilities
  float* sS = sP + N;             // [N] dS (backward)
  float* sQ = sS + N;             // [D] q * scale
  float* sG = sQ + D;             // [D] dO (backward)
  float* sAcc = sG + D;           // [4][D] per-warp partial channel sums
  float* red = sAcc + 4 * D;      // [4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
  const size_t tok = (size_t)2 * heads * D;
  const int sp = kv0 != nullptr ? 1 : 0;                      // split layout: key 0 comes from kv0
  const T* kb = kv + (size_t)b * (N - sp) * tok + (size_t)hd * D - (size_t)sp * tok;      // kb + k * tok is key k >= sp
  const T* kb0 = sp ? kv0 + (size_t)b * tok + (size_t)hd * D : kb;                         // key 0
  const size_t voff = (size_t)heads * D;
  auto krow = [&](int k) -> const T* { return k == 0 ? kb0 : kb + (size_t)k * tok; };
  const size_t qoff = (size_t)b * heads * D + (size_t)hd * D;
  for (int c = tid; c < D; c += 128) {
    sQ[c] = to_f(q[qoff + c]) * scale;
    if (BWD) sG[c] = to_f(dout[qoff + c]);
  }
  __syncthreads();
  float row[CA_MAXD];
  // ---- scores and softmax
  float m = -INFINITY;
  for (int k = tid; k < N; k += 128) {
    ca_load_row(krow(k), D, row);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) a = fmaf(sQ[c], row[c], a);
    sP[k] = a;
    m = fmaxf(m, a);
  }
  m = ca_block_reduce(m, red, true);
  float sum = 0.f;
  for (int k = tid; k < N; k += 128) { const float e = expf(sP[k] - m); sP[k] = e; sum += e; }
  sum = ca_block_reduce(sum, red, false);
  const float inv = 1.f / sum;
  for (int k = tid; k < N; k += 128) sP[k] *= inv;
  float dsum = 0.f;
  if (BWD) {
    for (int k = tid; k < N; k += 128) {
      ca_load_row(krow(k) + voff, D, row);
      float dp = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) dp = fmaf(sG[c], row[c], dp);
      sS[k] = dp;
      dsum = fmaf(sP[k], dp, dsum);
    }
    dsum = ca_block_reduce(dsum, red, false);
    T* dkb = dkv + (size_t)b * (N - sp) * tok + (size_t)hd * D - (size_t)sp * tok;
    T* dkb0 = sp ? dkv0 + (size_t)b * tok + (size_t)hd * D : dkb;
    for (int k = tid; k < N; k += 128) {
      const float ds = sP[k] * (sS[k] - dsum);
      sS[k] = ds;
      T* drow = k == 0 ? dkb0 : dkb + (size_t)k * tok;
      ca_store_row(drow, D, sQ, ds);                           // dK[k] = dS[k] * (q * scale)
      ca_store_row(drow + voff, D, sG, sP[k]);                 // dV[k] = P[k] * dO
    }
  }
  __syncthreads();
  // ---- per-channel sums over keys: out[c] = sum_k P[k] V[k][c]  /  dq[c] = scale * sum_k dS[k] K[k][c]
  const float* wgt = BWD ? sS : sP;
  const size_t soff = BWD ? 0 : voff;
  for (int c0 = 0; c0 < D; c0 += 32) {
    const int c = c0 + lane;
    float a0 = 0.f, a1 = 0.f;
    if (c < D) {
      int k = warp;
      for (; k + 4 < N; k += 8) {
        a0 = fmaf(wgt[k], to_f(krow(k)[soff + c]), a0);
        a1 = fmaf(wgt[k + 4], to_f(krow(k + 4)[soff + c]), a1);
      }
      if (k < N) a0 = fmaf(wgt[k], to_f(krow(k)[soff + c]), a0);
      sAcc[warp * D + c] = a0 + a1;
    }
  }
  __syncthreads();
  for (int c = tid; c < D; c += 128) {
    const float t = ((sAcc[c] + sAcc[D + c]) + sAcc[2 * D + c]) + sAcc[3 * D + c];
    if (BWD) dq[qoff + c] = from_f<T>(t * scale);
    else out[qoff + c] = from_f<T>(t);
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

// one launcher for both directions and both key layouts (kv0 == nullptr: keys [B, N, 2C] in kv; else split, see the kernel)
template <bool BWD>
static int class_attn_launch(const void* q, const void* kv0, const void* kv, const void* dout, void* out, void* dq, void* dkv0,
                             void* dkv, int B, int N, int heads, int D, float scale, int dtype, cudaStream_t st) {
  const char* name = BWD ? "class_attn_bwd" : "class_attn_fwd";
  APB_CHECK_ARG(B > 0 && N > 0 && heads > 0 && D > 0, APB_ERR_SHAPE, "%s: bad shape", name);
  APB_CHECK_ARG(kv0 == nullptr || N >= 2, APB_ERR_SHAPE, "%s: the split layout needs at least one token besides the class token", name);
  const size_t smem = (size_t)(2 * N + 6 * D + 4) * sizeof(float);
  APB_CHECK_ARG(smem <= 227 * 1024 && D <= CA_MAXD && D % 8 == 0, APB_ERR_UNSUPPORTED, "class_attn: N=%d D=%d unsupported", N, D);
  const int grid = B * heads;
#define CA_LAUNCH(T_, CD_)                                                                                            \
  do {                                                                                                                \
    cudaFuncSetAttribute(class_attn_kernel<T_, BWD, CD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    class_attn_kernel<T_, BWD, CD_><<<grid, 128, smem, st>>>((const T_*)q, (const T_*)kv, (const T_*)dout, (T_*)out, (T_*)dq, \
                                                             (T_*)dkv, B, N, heads, D, scale, (const T_*)kv0, (T_*)dkv0); \
  } while (0)
#define CA_DISPATCH(T_)            \
  do {                             \
    if (D == 32) CA_LAUNCH(T_, 32); \
    else if (D == 64) CA_LAUNCH(T_, 64); \
    else CA_LAUNCH(T_, 0);         \
  } while (0)
  if (dtype == APB_F32) CA_DISPATCH(float);
  else if (dtype == APB_BF16) CA_DISPATCH(bf16);
  else { apb_set_error("class_attn: dtype %d", dtype); return APB_ERR_DTYPE; }
#undef CA_DISPATCH
#undef CA_LAUNCH
  APB_LAUNCH_CHECK(name);
  return 0;
}

int apb_class_attn_fwd(const void* q, const void* kv, void* out, int B, int N, int heads, int D, float scale, int dtype,
                       apb_stream_t stream) {
  return class_attn_launch<false>(q, nullptr, kv, nullptr, out, nullptr, nullptr, nullptr, B, N, heads, D, scale, dtype, APB_STREAM(stream));
}

int apb_class_attn_bwd(const void* q, const void* kv, const void* dout, void* dq, void* dkv, int B, int N, int heads,
                       int D, float scale, int dtype, apb_stream_t stream) {
  return class_attn_launch<true>(q, nullptr, kv, dout, nullptr, dq, nullptr, dkv, B, N, heads, D, scale, dtype, APB_STREAM(stream));
}

// split key layout: the class token's own k / v row in kv_cls [B, 2C], the N - 1 patch tokens in kv_tok [B, N-1, 2C]
int apb_class_attn_fwd_split(const void* q, const void* kv_cls, const void* kv_tok, void* out, int B, int N, int heads, int D,
                             float scale, int dtype, apb_stream_t stream) {
  APB_CHECK_ARG(kv_cls != nullptr, APB_ERR_ARG, "class_attn_fwd_split: kv_cls is required");
  return class_attn_launch<false>(q, kv_cls, kv_tok, nullptr, out, nullptr, nullptr, nullptr, B, N, heads, D, scale, dtype, APB_STREAM(stream));
}

int apb_class_attn_bwd_split(const void* q, const void* kv_cls, const void* kv_tok, const void* dout, void* dq, void* dkv_cls,
                             void* dkv_tok, int B, int N, int heads, int D, float scale, int dtype, apb_stream_t stream) {
  APB_CHECK_ARG(kv_cls != nullptr && dkv_cls != nullptr, APB_ERR_ARG, "class_attn_bwd_split: kv_cls / dkv_cls are required");
  return class_attn_launch<true>(q, kv_cls, kv_tok, dout, nullptr, dq, dkv_cls, dkv_tok, B, N, heads, D, scale, dtype, APB_STREAM(stream));
}
