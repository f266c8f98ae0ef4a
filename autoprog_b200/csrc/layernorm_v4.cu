// Vectorised, branch-free variants of the LayerNorm backward and column-sum kernels (C % 4 == 0 / C % 8 == 0).
//
// The first versions (layernorm.cu) guard every element with `if (c < C)`; ptxas keeps those guards as branches, so
// the 36 loads of a row were issued one basic block at a time and the kernel sat at 10 % of HBM bandwidth waiting on
// the long scoreboard (profiles/r1_kernels.md).  Here every lane issues all of a row's 16-byte loads up front
// (clamped addresses + 0/1 masks instead of branches).
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  uint2 u;
  *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// NG groups of 4 consecutive channels per lane (group id = k*32 + lane)
template <int NG, typename TS, typename TC>
__global__ void __launch_bounds__(256) ln_bwd_v4_kernel(const TC* __restrict__ dy, const TS* __restrict__ xs,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, const TS* __restrict__ dres,
                                                        TS* __restrict__ dxs, TC* __restrict__ dr,
                                                        const float* __restrict__ rs, int rows_per_sample,
                                                        float* __restrict__ part_g, float* __restrict__ part_b,
                                                        long long rows, int C) {
  extern __shared__ float sm[];  // [nwarp][2][C]
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int ngroups = C >> 2;
  int off[NG];
  float okf[NG];
  float4 gam[NG], ag[NG], ab[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int g = k * 32 + lane;
    const bool ok = g < ngroups;
    off[k] = (ok ? g : 0) * 4;
    okf[k] = ok ? 1.f : 0.f;
    gam[k] = ld4(gamma + off[k]);
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invC = 1.f / (float)C;
  for (long long row = (long long)blockIdx.x * nwarp + warp; row < rows; row += (long long)gridDim.x * nwarp) {
    const size_t base = (size_t)row * C;
    float4 xv[NG], dv[NG], rv[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) {      // all loads of the row in flight before any use
      xv[k] = ld4(xs + base + off[k]);
      dv[k] = ld4(dy + base + off[k]);
      rv[k] = dres != nullptr ? ld4(dres + base + off[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mu = mean[row], rs_ = rstd[row];
    const float bs = (rs != nullptr) ? rs[row / rows_per_sample] : 1.f;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const float m = okf[k];
      xv[k].x = (xv[k].x - mu) * rs_; xv[k].y = (xv[k].y - mu) * rs_; xv[k].z = (xv[k].z - mu) * rs_; xv[k].w = (xv[k].w - mu) * rs_;
      dv[k].x *= m; dv[k].y *= m; dv[k].z *= m; dv[k].w *= m;
      ag[k].x = fmaf(dv[k].x, xv[k].x, ag[k].x); ag[k].y = fmaf(dv[k].y, xv[k].y, ag[k].y);
      ag[k].z = fmaf(dv[k].z, xv[k].z, ag[k].z); ag[k].w = fmaf(dv[k].w, xv[k].w, ag[k].w);
      ab[k].x += dv[k].x; ab[k].y += dv[k].y; ab[k].z += dv[k].z; ab[k].w += dv[k].w;
      dv[k].x *= gam[k].x; dv[k].y *= gam[k].y; dv[k].z *= gam[k].z; dv[k].w *= gam[k].w;   // g = dy * gamma
      s1 += (dv[k].x + dv[k].y) + (dv[k].z + dv[k].w);
      s2 += fmaf(dv[k].x, xv[k].x, dv[k].y * xv[k].y) + fmaf(dv[k].z, xv[k].z, dv[k].w * xv[k].w);
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      float4 d;
      d.x = fmaf(rs_, dv[k].x - s1 - xv[k].x * s2, rv[k].x);
      d.y = fmaf(rs_, dv[k].y - s1 - xv[k].y * s2, rv[k].y);
      d.z = fmaf(rs_, dv[k].z - s1 - xv[k].z * s2, rv[k].z);
      d.w = fmaf(rs_, dv[k].w - s1 - xv[k].w * s2, rv[k].w);
      if (okf[k] != 0.f) {
        st4(dxs + base + off[k], d);
        if (dr != nullptr) st4(dr + base + off[k], make_float4(bs * d.x, bs * d.y, bs * d.z, bs * d.w));
      }
    }
  }
  float* sg = sm + (size_t)warp * 2 * C;
#pragma unroll
  for (int k = 0; k < NG; ++k)
    if (okf[k] != 0.f) {
      st4(sg + off[k], ag[k]);
      st4(sg + C + off[k], ab[k]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float tg = 0.f, tb = 0.f;
    for (int w2 = 0; w2 < nwarp; ++w2) { tg += sm[(size_t)w2 * 2 * C + c]; tb += sm[(size_t)w2 * 2 * C + C + c]; }
    part_g[(size_t)blockIdx.x * C + c] = tg;
    part_b[(size_t)blockIdx.x * C + c] = tb;
  }
}

// forward: xs = x + rs[b]*r ; y = LN(xs)*gamma + beta.  NG groups of 4 channels per lane, all loads up front.
template <int NG, typename TS, typename TC>
__global__ void __launch_bounds__(256) ln_fwd_v4_kernel(const TS* __restrict__ x, const TC* __restrict__ r,
                                                        const float* __restrict__ rs, int rows_per_sample,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        TS* __restrict__ xs_out, TC* __restrict__ y,
                                                        float* __restrict__ mean, float* __restrict__ rstd, long long rows,
                                                        int C, float eps) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int ngroups = C >> 2;
  const size_t base = (size_t)row * C;
  int off[NG];
  bool ok[NG];
  float4 v[NG], rv[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int g = k * 32 + lane;
    ok[k] = g < ngroups;
    off[k] = (ok[k] ? g : 0) * 4;
    v[k] = ld4(x + base + off[k]);
    rv[k] = r != nullptr ? ld4(r + base + off[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float s = (rs != nullptr) ? rs[row / rows_per_sample] : 1.f;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    if (r != nullptr) {
      v[k].x = fmaf(s, rv[k].x, v[k].x); v[k].y = fmaf(s, rv[k].y, v[k].y);
      v[k].z = fmaf(s, rv[k].z, v[k].z); v[k].w = fmaf(s, rv[k].w, v[k].w);
    }
    if (xs_out != nullptr && ok[k]) st4(xs_out + base + off[k], v[k]);
    if (!ok[k]) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mu = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NG; ++k)
    if (ok[k]) {
      const float a = v[k].x - mu, b = v[k].y - mu, c = v[k].z - mu, d = v[k].w - mu;
      sq += fmaf(a, a, b * b) + fmaf(c, c, d * d);
    }
  const float rs_ = rsqrtf(warp_sum(sq) / (float)C + eps);
  if (y != nullptr) {
#pragma unroll
    for (int k = 0; k < NG; ++k)
      if (ok[k]) {
        const float4 ga = ld4(gamma + off[k]), be = ld4(beta + off[k]);
        st4(y + base + off[k], make_float4(fmaf((v[k].x - mu) * rs_, ga.x, be.x), fmaf((v[k].y - mu) * rs_, ga.y, be.y),
                                           fmaf((v[k].z - mu) * rs_, ga.z, be.z), fmaf((v[k].w - mu) * rs_, ga.w, be.w)));
      }
  }
  if (lane == 0 && mean != nullptr) { mean[row] = mu; rstd[row] = rs_; }
}

// column sums, stage 1.  A thread owns 8 consecutive columns (16-byte bf16 loads); TX threads span the CTA's column
// slice and the remaining 256 / TX thread rows stride the CTA's row range with 4 rows in flight each.  The host picks
// TX from C so narrow matrices (C = 192: TX = 24, 10 thread rows) keep every lane busy, and rows_per_cta so the grid
// is several CTAs per SM -- the kernel is a pure HBM / L2 stream.
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 lo = *reinterpret_cast<const float4*>(p), hi = *reinterpret_cast<const float4*>(p + 4);
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_v8_kernel(const T* __restrict__ a, long long rows, int C,
                                                        float* __restrict__ part, int rows_per_cta, int TX) {
  __shared__ __align__(16) float s[256 * 8];              // [TY][TX * 8]
  pdl_wait();
  pdl_trigger();
  const int TY = 256 / TX;
  const int ty = threadIdx.x / TX, tx = threadIdx.x - ty * TX;
  const int c0 = (blockIdx.x * TX + tx) * 8;
  const bool ok = ty < TY && c0 < C;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + (long long)rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ok) {
    const T* col = a + c0;
    long long r = r0 + ty;
    for (; r + 3 * TY < r1; r += 4 * TY) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) ld8(col + (size_t)(r + u * TY) * C, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; r < r1; r += TY) {
      float v[8];
      ld8(col + (size_t)r * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
  if (ty < TY) {
    float4* d = reinterpret_cast<float4*>(s + (ty * TX + tx) * 8);
    d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  const int W = TX * 8;
  for (int e = threadIdx.x; e < W; e += 256) {
    const int c = blockIdx.x * W + e;
    if (c < C) {
      float t = 0.f;
      for (int k = 0; k < TY; ++k) t += s[k * W + e];     // fixed order
      part[(size_t)blockIdx.y * C + c] = t;
    }
  }
}

// out[i] = sum_s parts[s][i] (fixed order): the deterministic tail of the split-K wgrad GEMM, one launch.  A second,
// small job (the bias-gradient partials of the same GEMM) rides along in the same grid.
// block = 64 float4 columns x 4 split phases: phase g sums the splits s = g, g+4, ... (two loads in flight), the four
// phase sums are then added in fixed order through shared memory -- the partials are a few MB spread over 8-63
// splits, so parallelism over the split dimension is what keeps the loads in flight.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ parts, float* __restrict__ out, int splits,
                                                            long long n4, const float* __restrict__ parts2,
                                                            float* __restrict__ out2, long long n4b, int splits2) {
  __shared__ float4 sm[4][64];
  pdl_wait();
  pdl_trigger();
  const int tx = threadIdx.x & 63, g = threadIdx.x >> 6;
  const long long i = (long long)blockIdx.x * 64 + tx;
  const bool live = i < n4 + n4b;
  const bool second = i >= n4;
  const float4* src = reinterpret_cast<const float4*>(second ? parts2 : parts);
  const long long j = second ? i - n4 : i, stride = second ? n4b : n4;
  const int ns = second ? splits2 : splits;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    int s = g;
    for (; s + 4 < ns; s += 8) {
      const float4 a = src[(long long)s * stride + j], b = src[(long long)(s + 4) * stride + j];
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (s < ns) {
      const float4 a = src[(long long)s * stride + j];
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  sm[g][tx] = acc;
  __syncthreads();
  if (g == 0 && live) {
    const float4 b = sm[1][tx], c = sm[2][tx], d = sm[3][tx];
    acc.x = ((acc.x + b.x) + c.x) + d.x;
    acc.y = ((acc.y + b.y) + c.y) + d.y;
    acc.z = ((acc.z + b.z) + c.z) + d.z;
    acc.w = ((acc.w + b.w) + c.w) + d.w;
    reinterpret_cast<float4*>(second ? out2 : out)[j] = acc;
  }
}

}  // namespace

int apb_splitk_reduce2(const float* parts, float* out, long long n, int splits, const float* parts2, float* out2, long long n2,
                       int splits2, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(splits >= 1 && n > 0 && (n % 4) == 0 && (((uintptr_t)parts | (uintptr_t)out) & 15) == 0, APB_ERR_ARG,
                "splitk_reduce: splits=%d n=%lld (n must be a multiple of 4, pointers 16-byte aligned)", splits, n);
  APB_CHECK_ARG(n2 == 0 || (parts2 != nullptr && out2 != nullptr && (n2 % 4) == 0 && (((uintptr_t)parts2 | (uintptr_t)out2) & 15) == 0),
                APB_ERR_ARG, "splitk_reduce: second job n2=%lld must be a multiple of 4 with 16-byte aligned pointers", n2);
  const long long n4 = n / 4, n4b = n2 / 4;
  const long long grid = (n4 + n4b + 63) / 64;
  APB_CHECK_ARG(grid <= 0x7fffffffLL, APB_ERR_SHAPE, "splitk_reduce: n too large");
  apb_launch_pdl(splitk_reduce_kernel, dim3((unsigned)grid), dim3(256), 0, st, parts, out, splits, n4, parts2, out2, n4b,
                 splits2 < 1 ? 1 : splits2);
  APB_LAUNCH_CHECK("splitk_reduce");
  return 0;
}

int apb_splitk_reduce(const float* parts, float* out, int splits, long long n, apb_stream_t stream) {
  return apb_splitk_reduce2(parts, out, n, splits, nullptr, nullptr, 0, 1, stream);
}

// returns 1 if handled, 0 if the caller must use the scalar kernel, <0 / >1 on error
int ln_bwd_v4_launch(const void* dy, const void* xs, const float* mean, const float* rstd, const float* gamma,
                     const void* dres, void* dxs, void* dr, const float* rs, int rows_per_sample, float* pg, float* pb,
                     int grid, long long rows, int C, int sdtype, int cdtype, cudaStream_t st) {
  if ((C & 3) != 0 || C > 1024) return 0;
  const uintptr_t al = (uintptr_t)dy | (uintptr_t)xs | (uintptr_t)gamma | (uintptr_t)dres | (uintptr_t)dxs | (uintptr_t)dr;
  if (al & 15) return 0;
  const int ng = (C / 4 + 31) / 32;
  const size_t smem = (size_t)8 * 2 * C * sizeof(float);
#define V4(NG_, TS_, TC_)                                                                                             \
  do {                                                                                                                \
    cudaFuncSetAttribute(ln_bwd_v4_kernel<NG_, TS_, TC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    apb_launch_pdl(ln_bwd_v4_kernel<NG_, TS_, TC_>, dim3(grid), dim3(256), smem, st, (const TC_*)dy, (const TS_*)xs, mean, rstd, \
                   gamma, (const TS_*)dres, (TS_*)dxs, (TC_*)dr, rs, rows_per_sample, pg, pb, rows, C);               \
  } while (0)
#define V4_T(TS_, TC_)                  \
  do {                                  \
    if (ng <= 1) V4(1, TS_, TC_);       \
    else if (ng <= 2) V4(2, TS_, TC_);  \
    else if (ng <= 3) V4(3, TS_, TC_);  \
    else if (ng <= 4) V4(4, TS_, TC_);  \
    else if (ng <= 6) V4(6, TS_, TC_);  \
    else V4(8, TS_, TC_);               \
  } while (0)
  if (sdtype == APB_F32 && cdtype == APB_F32) V4_T(float, float);
  else if (sdtype == APB_F32 && cdtype == APB_BF16) V4_T(float, bf16);
  else if (sdtype == APB_BF16 && cdtype == APB_BF16) V4_T(bf16, bf16);
  else return 0;
#undef V4_T
#undef V4
  return 1;
}

int ln_fwd_v4_launch(const void* x, const void* r, const float* rs, int rows_per_sample, const float* gamma, const float* beta,
                     void* xs_out, void* y, float* mean, float* rstd, long long rows, int C, float eps, int sdtype, int cdtype,
                     cudaStream_t st) {
  if ((C & 3) != 0 || C > 1024) return 0;
  const uintptr_t al = (uintptr_t)x | (uintptr_t)r | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)xs_out | (uintptr_t)y;
  if (al & 15) return 0;
  const int ng = (C / 4 + 31) / 32;
  const int grid = ceil_div(rows, 8);
#define F4(NG_, TS_, TC_)                                                                                            \
  apb_launch_pdl(ln_fwd_v4_kernel<NG_, TS_, TC_>, dim3(grid), dim3(256), 0, st, (const TS_*)x, (const TC_*)r, rs,        \
                 rows_per_sample, gamma, beta, (TS_*)xs_out, (TC_*)y, mean, rstd, rows, C, eps)
#define F4_T(TS_, TC_)                  \
  do {                                  \
    if (ng <= 1) F4(1, TS_, TC_);       \
    else if (ng <= 2) F4(2, TS_, TC_);  \
    else if (ng <= 3) F4(3, TS_, TC_);  \
    else if (ng <= 4) F4(4, TS_, TC_);  \
    else if (ng <= 6) F4(6, TS_, TC_);  \
    else F4(8, TS_, TC_);               \
  } while (0)
  if (sdtype == APB_F32 && cdtype == APB_F32) F4_T(float, float);
  else if (sdtype == APB_F32 && cdtype == APB_BF16) F4_T(float, bf16);
  else if (sdtype == APB_BF16 && cdtype == APB_BF16) F4_T(bf16, bf16);
  else return 0;
#undef F4_T
#undef F4
  return 1;
}

// stage-1 geometry shared by the launcher and apb_colsum_workspace_floats
void colsum_plan(long long rows, int C, int* gx, int* TX, int* rows_per_cta, int* parts) {
  const int groups = (C + 7) / 8;
  *gx = (groups + 255) / 256;
  *TX = (groups + *gx - 1) / *gx;
  const int TY = 256 / *TX;
  const long long want = (148 * 6) / *gx;                  // ~6 CTAs per SM
  long long rpc = (rows + want - 1) / want;
  if (rpc < 4LL * TY) rpc = 4LL * TY;
  *rows_per_cta = (int)rpc;
  *parts = (int)((rows + rpc - 1) / rpc);
}

int colsum_v8_launch(const void* a, long long rows, int C, float* part, int dtype, cudaStream_t st) {
  if ((C & 7) != 0 || ((uintptr_t)a & 15)) return 0;
  int gx, TX, rpc, parts;
  colsum_plan(rows, C, &gx, &TX, &rpc, &parts);
  dim3 grid(gx, parts);
  if (dtype == APB_F32) apb_launch_pdl(colsum_v8_kernel<float>, grid, dim3(256), 0, st, (const float*)a, rows, C, part, rpc, TX);
  else if (dtype == APB_BF16) apb_launch_pdl(colsum_v8_kernel<bf16>, grid, dim3(256), 0, st, (const bf16*)a, rows, C, part, rpc, TX);
  else return 0;
  return 1;
}
