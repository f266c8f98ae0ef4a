// Multi-head self-attention core on tensor cores (bf16 in, fp32 accumulate), head_dim 32 (every VOLO variant).
//
// Replaces models/volo.py:193-197 (q@k^T*scale -> softmax -> @v with the [B,heads,N,N] scores materialised) by
// flash-style kernels: scores live in registers, K/V (or Q/dO) of one (b, head) live in shared memory.
// N is 64..784 on this path and d=32, so one (b, head) problem is two 196x32x196 products: far below a tcgen05
// tile (128 x N x 16 per instruction, TMEM round trip per softmax) -- the warp-level mma.sync.m16n8k16 pipe keeps
// the whole softmax in the accumulator registers and is the right tool at this size.
//
//   fwd : CTA = one (b, head); each warp owns 16 query rows, loops over 64-key tiles with an online softmax.
//   bwd : (1) D[i] = <dO_i, O_i>            (2) dQ : CTA per (b, head), warp per 16 queries, K/V in smem
//         (3) dK,dV : CTA per (b, head), warp per 16 keys, Q/dO in smem.  No atomics -> deterministic.
// Layouts: qkv [B,N,3,heads,32], out/dout [B,N,heads,32], lse/D [B,heads,N] fp32.
#include "common.cuh"
#include <type_traits>

namespace {

// head dim D is a template parameter (32: every VOLO variant; 64: DeiT).  Shared-memory row pitch = D + 8 bf16
// (80 / 144 bytes): ldmatrix rows of 8 consecutive keys fall into distinct bank groups.

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {   // bare MUFU.EX2 (inputs are <= 0 here; denormal results flush to 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// rows [0,N) of a [N, 32] bf16 matrix with global row stride `gstride` (elements) -> smem [rows_pad][ROWP], tail zeroed
template <int D>
__device__ __forceinline__ void stage_rows(bf16* s, const bf16* g, size_t gstride, int N, int rows_pad) {
  constexpr int ROWP = D + 8, CPR = D / 8;
  for (int e = threadIdx.x; e < rows_pad * CPR; e += blockDim.x) {
    const int n = e / CPR, ch = e % CPR;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < N) v = *reinterpret_cast<const uint4*>(g + (size_t)n * gstride + ch * 8);
    *reinterpret_cast<uint4*>(s + n * ROWP + ch * 8) = v;
  }
}

// A-operand fragments (16 rows x 32 channels = 2 k-steps) straight from global; rows >= N read as zero
template <int D>
__device__ __forceinline__ void load_a_rows(uint32_t (&a)[D / 16][4], const bf16* g, size_t gstride, int row0, int N, int lane) {
  const int gi = lane >> 2, q = lane & 3;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks)
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int r = row0 + gi + (h & 1) * 8;
      const int c = ks * 16 + 2 * q + (h >> 1) * 8;
      a[ks][h] = (r < N) ? *reinterpret_cast<const uint32_t*>(g + (size_t)r * gstride + c) : 0u;
    }
}

// Per-lane ldmatrix byte offsets inside a [rows][ROWP] bf16 tile (loop invariant; all smem addressing is 32-bit):
//   offT : "rows^T" pattern -- 8 rows x 4 channel chunks (B operand "col" of S = A . M^T)
//   offR : "rows" pattern   -- 16 rows x 2 channel chunks, used with ldmatrix.trans (B operand of O = P . M)
template <int D>
__device__ __forceinline__ uint32_t lane_offT(int lane) { return (uint32_t)(((lane & 7) * (D + 8) + (lane >> 3) * 8) * 2); }
template <int D>
__device__ __forceinline__ uint32_t lane_offR(int lane) {
  const int mi = lane >> 3, r = lane & 7;
  return (uint32_t)((((mi & 1) * 8 + r) * (D + 8) + (mi >> 1) * 8) * 2);
}

// S[16 x 8] (+)= A[16 x D] . M[key0..key0+7][0..D)^T ; addr = tile + key0*ROWB + lane_offT
template <int D>
__device__ __forceinline__ void mma_rowsT(float (&c)[4], const uint32_t (&a)[D / 16][4], uint32_t addr) {
#pragma unroll
  for (int k2 = 0; k2 < D / 32; ++k2) {
    uint32_t b[4];
    ldsm_x4(b, addr + k2 * 64);
    mma16816(c, a[2 * k2], b[0], b[1]);
    mma16816(c, a[2 * k2 + 1], b[2], b[3]);
  }
}

// O[16 x D] += P[16 x 16 (keys key0..+15)] . M[key0..key0+15][0..D) ; addr = tile + key0*ROWB + lane_offR
template <int D>
__device__ __forceinline__ void mma_rows(float (&o)[D / 8][4], const uint32_t (&pa)[4], uint32_t addr) {
#pragma unroll
  for (int half = 0; half < D / 16; ++half) {
    uint32_t b[4];
    ldsm_x4_t(b, addr + half * 32);
    mma16816(o[2 * half], pa, b[0], b[1]);
    mma16816(o[2 * half + 1], pa, b[2], b[3]);
  }
}

template <int D, int MINB>
__global__ void __launch_bounds__(256, MINB) mhsa_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                           float* __restrict__ lse, int N, int heads, float scale,
                                                           int rows_pad) {
  constexpr int ROWP = D + 8;
  constexpr uint32_t ROWB = ROWP * 2;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* sK = reinterpret_cast<bf16*>(smraw);
  bf16* sV = sK + (size_t)rows_pad * ROWP;
  const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
  const size_t tok = (size_t)3 * heads * D;
  const bf16* qb = qkv + (size_t)b * N * tok + (size_t)hd * D;
  stage_rows<D>(sK, qb + heads * D, tok, N, rows_pad);
  stage_rows<D>(sV, qb + 2 * heads * D, tok, N, rows_pad);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int gi = lane >> 2, q = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
  const uint32_t sK_T = smem_u32(sK) + lane_offT<D>(lane), sV_R = smem_u32(sV) + lane_offR<D>(lane);
  const int nfull = N / 64;                 // key blocks that need no masking
  for (int row0 = warp * 16; row0 < N; row0 += nwarp * 16) {
    uint32_t qa[D / 16][4];
    load_a_rows<D>(qa, qb, tok, row0, N, lane);
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    // one 64-key block: S = Q K^T (registers), online softmax, O += P V.  TAIL = false: all 64 keys valid, no
    // predicates anywhere (the steady state).  TAIL = true: only the first `nbv` 8-key sub-blocks exist; everything
    // beyond them is skipped (warp-uniform), keys >= N inside the last sub-block are masked.
    auto block = [&](auto tail_tag, const int k0) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      const int nbv = TAIL ? (N - k0 + 7) >> 3 : 8;
      float s[8][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = TAIL ? -INFINITY : 0.f;
        if (!TAIL || nb < nbv) {
          s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
          mma_rowsT<D>(s[nb], qa, sK_T + (uint32_t)(k0 + nb * 8) * ROWB);
          if (TAIL) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k0 + nb * 8 + 2 * q + (j & 1) >= N) s[nb][j] = -INFINITY;
          }
        }
      }
      float mt[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        if (!TAIL || nb < nbv) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mt[j >> 1] = fmaxf(mt[j >> 1], s[nb][j]);
        }
      float corr[2], msc[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float mn = fmaxf(m_run[h], quad_max(mt[h]));
        corr[h] = ex2((m_run[h] - mn) * sl2);
        m_run[h] = mn;
        msc[h] = mn * sl2;
        l_run[h] *= corr[h];
      }
#pragma unroll
      for (int i = 0; i < D / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        if (!TAIL || nb < nbv) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float pv = ex2(fmaf(s[nb][j], sl2, -msc[j >> 1]));
            s[nb][j] = pv;
            l_run[j >> 1] += pv;
          }
        } else {
          s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
        }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (!TAIL || 2 * kk < nbv) {
          const uint32_t pa[4] = {pack_bf16(s[2 * kk][0], s[2 * kk][1]), pack_bf16(s[2 * kk][2], s[2 * kk][3]),
                                  pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3])};
          mma_rows<D>(o, pa, sV_R + (uint32_t)(k0 + kk * 16) * ROWB);
        }
      }
    };
    for (int kb = 0; kb < nfull; ++kb) block(std::false_type{}, kb * 64);
    if (nfull * 64 < N) block(std::true_type{}, nfull * 64);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float l = quad_sum(l_run[h]);
      const int r = row0 + gi + h * 8;
      if (r < N) {
        const float inv = 1.f / l;
        bf16* orow = out + ((size_t)b * N + r) * heads * D + (size_t)hd * D;
#pragma unroll
        for (int i = 0; i < D / 8; ++i)
          *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * q) = pack_bf16(o[i][2 * h] * inv, o[i][2 * h + 1] * inv);
        if (q == 0) lse[((size_t)b * heads + hd) * N + r] = m_run[h] * scale + logf(l);
      }
    }
  }
}

// D[i] = sum_c dO[i][c] * O[i][c]
template <int D>
__global__ void __launch_bounds__(256) mhsa_rowdot_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                                          float* __restrict__ drow, int B, int N, int heads) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, n, head)
  const long long total = (long long)B * N * heads;
  if (idx >= total) return;
  const int hd = (int)(idx % heads);
  const long long bn = idx / heads;
  const int n = (int)(bn % N);
  const int b = (int)(bn / N);
  const uint4* po = reinterpret_cast<const uint4*>(out + (size_t)idx * D);
  const uint4* pg = reinterpret_cast<const uint4*>(dout + (size_t)idx * D);
  float acc = 0.f;
#pragma unroll
  for (int v4 = 0; v4 < D / 8; ++v4) {
    const uint4 a = po[v4], g = pg[v4];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 fa = __bfloat1622float2(ha[t]), fg = __bfloat1622float2(hg[t]);
      acc = fmaf(fa.x, fg.x, acc);
      acc = fmaf(fa.y, fg.y, acc);
    }
  }
  drow[((size_t)b * heads + hd) * N + n] = acc;
}

// dQ: warp per 16 queries; S = Q K^T, P = exp(S*scale - lse), dP = dO V^T, dS = P*(dP - D), dQ = scale * dS K
template <int D, int MINB>
__global__ void __launch_bounds__(256, MINB) mhsa_bwd_dq_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                              const float* __restrict__ lse, const float* __restrict__ drow,
                                                              bf16* __restrict__ dqkv, int N, int heads, float scale,
                                                              int rows_pad) {
  constexpr int ROWP = D + 8;
  constexpr uint32_t ROWB = ROWP * 2;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* sK = reinterpret_cast<bf16*>(smraw);
  bf16* sV = sK + (size_t)rows_pad * ROWP;
  const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
  const size_t tok = (size_t)3 * heads * D, otok = (size_t)heads * D;
  const bf16* qb = qkv + (size_t)b * N * tok + (size_t)hd * D;
  const bf16* gb = dout + (size_t)b * N * otok + (size_t)hd * D;
  stage_rows<D>(sK, qb + heads * D, tok, N, rows_pad);
  stage_rows<D>(sV, qb + 2 * heads * D, tok, N, rows_pad);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int gi = lane >> 2, q = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
  const float* L = lse + ((size_t)b * heads + hd) * N;
  const float* Dr = drow + ((size_t)b * heads + hd) * N;
  const uint32_t sK_T = smem_u32(sK) + lane_offT<D>(lane), sV_T = smem_u32(sV) + lane_offT<D>(lane), sK_R = smem_u32(sK) + lane_offR<D>(lane);
  for (int row0 = warp * 16; row0 < N; row0 += nwarp * 16) {
    uint32_t qa[D / 16][4], ga[D / 16][4];
    load_a_rows<D>(qa, qb, tok, row0, N, lane);
    load_a_rows<D>(ga, gb, otok, row0, N, lane);
    float l2[2], dr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = row0 + gi + h * 8;
      l2[h] = (r < N) ? L[r] * 1.4426950408889634f : 0.f;
      dr[h] = (r < N) ? Dr[r] : 0.f;
    }
    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
    auto step = [&](auto tail_tag, const int k0) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      float s[2][4], dp[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
        dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
        mma_rowsT<D>(s[nb], qa, sK_T + (uint32_t)(k0 + nb * 8) * ROWB);
        mma_rowsT<D>(dp[nb], ga, sV_T + (uint32_t)(k0 + nb * 8) * ROWB);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float pv = ex2(fmaf(s[nb][j], sl2, -l2[j >> 1]));
          if (TAIL && k0 + nb * 8 + 2 * q + (j & 1) >= N) pv = 0.f;
          s[nb][j] = pv * (dp[nb][j] - dr[j >> 1]);
        }
      }
      const uint32_t pa[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]), pack_bf16(s[1][0], s[1][1]),
                              pack_bf16(s[1][2], s[1][3])};
      mma_rows<D>(dq, pa, sK_R + (uint32_t)k0 * ROWB);
    };
    const int kfull = N & ~15;
    for (int k0 = 0; k0 < kfull; k0 += 16) step(std::false_type{}, k0);
    if (kfull < N) step(std::true_type{}, kfull);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = row0 + gi + h * 8;
      if (r < N) {
        bf16* drowp = dqkv + ((size_t)b * N + r) * tok + (size_t)hd * D;
#pragma unroll
        for (int i = 0; i < D / 8; ++i)
          *reinterpret_cast<uint32_t*>(drowp + i * 8 + 2 * q) = pack_bf16(dq[i][2 * h] * scale, dq[i][2 * h + 1] * scale);
      }
    }
  }
}

// dK, dV: warp per 16 keys; S^T = K Q^T, P^T = exp(S^T*scale - lse[query]), dV = P^T dO, dP^T = V dO^T,
// dS^T = P^T*(dP^T - D[query]), dK = scale * dS^T Q.   Q, dO, lse, D of the (b, head) in smem.
template <int D>
__global__ void __launch_bounds__(256, (D == 32 ? 2 : 1)) mhsa_bwd_dkv_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                               const float* __restrict__ lse,
                                                               const float* __restrict__ drow, bf16* __restrict__ dqkv,
                                                               int N, int heads, float scale, int rows_pad) {
  constexpr int ROWP = D + 8;
  constexpr uint32_t ROWB = ROWP * 2;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* sQ = reinterpret_cast<bf16*>(smraw);
  bf16* sG = sQ + (size_t)rows_pad * ROWP;
  float* sL = reinterpret_cast<float*>(sG + (size_t)rows_pad * ROWP);
  float* sD = sL + rows_pad;
  const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
  const size_t tok = (size_t)3 * heads * D, otok = (size_t)heads * D;
  const bf16* qb = qkv + (size_t)b * N * tok + (size_t)hd * D;
  const bf16* gb = dout + (size_t)b * N * otok + (size_t)hd * D;
  stage_rows<D>(sQ, qb, tok, N, rows_pad);
  stage_rows<D>(sG, gb, otok, N, rows_pad);
  for (int n = threadIdx.x; n < rows_pad; n += blockDim.x) {
    sL[n] = (n < N) ? lse[((size_t)b * heads + hd) * N + n] * 1.4426950408889634f : 0.f;
    sD[n] = (n < N) ? drow[((size_t)b * heads + hd) * N + n] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int gi = lane >> 2, q = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
  const uint32_t sQ_T = smem_u32(sQ) + lane_offT<D>(lane), sG_T = smem_u32(sG) + lane_offT<D>(lane);
  const uint32_t sQ_R = smem_u32(sQ) + lane_offR<D>(lane), sG_R = smem_u32(sG) + lane_offR<D>(lane);
  for (int key0 = warp * 16; key0 < N; key0 += nwarp * 16) {
    uint32_t ka[D / 16][4], va[D / 16][4];
    load_a_rows<D>(ka, qb + heads * D, tok, key0, N, lane);
    load_a_rows<D>(va, qb + 2 * heads * D, tok, key0, N, lane);
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }
    auto step = [&](auto tail_tag, const int i0) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      float s[2][4], dp[2][4];
      uint32_t pp[4], ds[4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
        dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
        mma_rowsT<D>(s[nb], ka, sQ_T + (uint32_t)(i0 + nb * 8) * ROWB);     // S^T[key][query]
        mma_rowsT<D>(dp[nb], va, sG_T + (uint32_t)(i0 + nb * 8) * ROWB);    // dP^T[key][query]
        const int qi0 = i0 + nb * 8 + 2 * q;                                 // this lane's two queries: qi0, qi0 + 1
        const float2 lq = *reinterpret_cast<const float2*>(sL + qi0), dq2 = *reinterpret_cast<const float2*>(sD + qi0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float pv = ex2(fmaf(s[nb][j], sl2, -((j & 1) ? lq.y : lq.x)));
          if (TAIL && qi0 + (j & 1) >= N) pv = 0.f;
          s[nb][j] = pv;
          dp[nb][j] = pv * (dp[nb][j] - ((j & 1) ? dq2.y : dq2.x));
        }
        pp[nb * 2 + 0] = pack_bf16(s[nb][0], s[nb][1]);
        pp[nb * 2 + 1] = pack_bf16(s[nb][2], s[nb][3]);
        ds[nb * 2 + 0] = pack_bf16(dp[nb][0], dp[nb][1]);
        ds[nb * 2 + 1] = pack_bf16(dp[nb][2], dp[nb][3]);
      }
      mma_rows<D>(dv, pp, sG_R + (uint32_t)i0 * ROWB);
      mma_rows<D>(dk, ds, sQ_R + (uint32_t)i0 * ROWB);
    };
    const int ifull = N & ~15;
    for (int i0 = 0; i0 < ifull; i0 += 16) step(std::false_type{}, i0);
    if (ifull < N) step(std::true_type{}, ifull);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = key0 + gi + h * 8;
      if (r < N) {
        bf16* kp = dqkv + ((size_t)b * N + r) * tok + (size_t)heads * D + (size_t)hd * D;
        bf16* vp = kp + (size_t)heads * D;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
          *reinterpret_cast<uint32_t*>(kp + i * 8 + 2 * q) = pack_bf16(dk[i][2 * h] * scale, dk[i][2 * h + 1] * scale);
          *reinterpret_cast<uint32_t*>(vp + i * 8 + 2 * q) = pack_bf16(dv[i][2 * h], dv[i][2 * h + 1]);
        }
      }
    }
  }
}

int pick_warps(int N) {
  const int tiles = (N + 15) / 16;
  const int rounds = (tiles + 7) / 8;
  return (tiles + rounds - 1) / rounds;
}

}  // namespace

template <int D>
static int mhsa_fwd_mma_t(const void* qkv, void* out, float* lse, int B, int N, int heads, float scale, cudaStream_t st) {
  constexpr int ROWP = D + 8;
  const int rows_pad = (N + 63) / 64 * 64;
  const size_t smem = (size_t)2 * rows_pad * ROWP * sizeof(bf16);
  APB_CHECK_ARG(smem <= 227 * 1024, APB_ERR_UNSUPPORTED, "mhsa_fwd_mma: N=%d needs %zu B smem", N, smem);
  // D = 32: 3 CTAs / SM (80-register cap, a few spilled bytes) measured faster than 2 CTAs without spills (75 vs 86 us)
  constexpr int MINB = (D == 32) ? 3 : 1;
  cudaFuncSetAttribute(mhsa_fwd_mma_kernel<D, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mhsa_fwd_mma_kernel<D, MINB><<<B * heads, pick_warps(N) * 32, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, N, heads, scale, rows_pad);
  APB_LAUNCH_CHECK("mhsa_fwd_mma");
  return 0;
}

template <int D>
static int mhsa_bwd_mma_t(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace,
                          int B, int N, int heads, float scale, cudaStream_t st) {
  constexpr int ROWP = D + 8;
  const int rows_pad = (N + 15) / 16 * 16;
  const size_t smem1 = (size_t)2 * rows_pad * ROWP * sizeof(bf16);
  const size_t smem2 = smem1 + (size_t)2 * rows_pad * sizeof(float);
  APB_CHECK_ARG(smem2 <= 227 * 1024, APB_ERR_UNSUPPORTED, "mhsa_bwd_mma: N=%d needs %zu B smem", N, smem2);
  const long long total = (long long)B * N * heads;
  mhsa_rowdot_kernel<D><<<ceil_div(total, 256), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, workspace, B, N, heads);
  APB_LAUNCH_CHECK("mhsa_rowdot");
  constexpr int MINB = (D == 32) ? 3 : 1;
  cudaFuncSetAttribute(mhsa_bwd_dq_mma_kernel<D, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  cudaFuncSetAttribute(mhsa_bwd_dkv_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  const int threads = pick_warps(N) * 32;
  mhsa_bwd_dq_mma_kernel<D, MINB><<<B * heads, threads, smem1, st>>>((const bf16*)qkv, (const bf16*)dout, lse, workspace, (bf16*)dqkv,
                                                                    N, heads, scale, rows_pad);
  APB_LAUNCH_CHECK("mhsa_bwd_dq_mma");
  mhsa_bwd_dkv_mma_kernel<D><<<B * heads, threads, smem2, st>>>((const bf16*)qkv, (const bf16*)dout, lse, workspace, (bf16*)dqkv,
                                                               N, heads, scale, rows_pad);
  APB_LAUNCH_CHECK("mhsa_bwd_dkv_mma");
  return 0;
}

// D = 32 or 64; anything else -> APB_ERR_UNSUPPORTED (the dispatcher then uses the SIMT kernels)
int apb_mhsa_fwd_mma(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (D == 32) return mhsa_fwd_mma_t<32>(qkv, out, lse, B, N, heads, scale, st);
  if (D == 64) return mhsa_fwd_mma_t<64>(qkv, out, lse, B, N, heads, scale, st);
  return APB_ERR_UNSUPPORTED;
}

int apb_mhsa_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace,
                     int B, int N, int heads, int D, float scale, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (D == 32) return mhsa_bwd_mma_t<32>(qkv, out, dout, lse, dqkv, workspace, B, N, heads, scale, st);
  if (D == 64) return mhsa_bwd_mma_t<64>(qkv, out, dout, lse, dqkv, workspace, B, N, heads, scale, st);
  return APB_ERR_UNSUPPORTED;
}
