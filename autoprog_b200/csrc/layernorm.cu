// Fused residual-add (+ per-sample DropPath scale) + LayerNorm, forward and backward.
//
// Replaces the `x = x + drop_path(branch); y = norm(x)` pairs of models/volo.py:142-143, 232-233, 306-307
// (nn.LayerNorm over the channel dim, eps 1e-5; timm DropPath = branch * mask[b] / keep):
//   fwd : xs = x + rs[b] * r      (r, rs optional)          -> xs_out (stream dtype TS)
//         y  = (xs - mean) * rstd * gamma + beta              -> y (compute dtype TC), mean/rstd fp32 per row
//   bwd : dxs = dres + LN'(dy)    (dres optional: gradient arriving on the residual stream)
//         dr  = rs[b] * dxs       (optional second output in the compute dtype: gradient of the branch r)
//         dgamma/dbeta: per-CTA partial sums, reduced by a second deterministic kernel
// One warp per row; the row lives in registers (C <= 32*MAXV).
#include "common.cuh"

namespace {

constexpr int MAXC = 1024;  // per-lane elements NV = ceil(C/32) is a template parameter (<= 32)

template <int MAXV, typename TS, typename TR, typename TC>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TS* __restrict__ x, const TR* __restrict__ r,
                                                     const float* __restrict__ rs, int rows_per_sample,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     TS* __restrict__ xs_out, TC* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, long long rows,
                                                     int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = (C + 31) / 32;
  float v[MAXV];
  const float s = (rs != nullptr) ? rs[row / rows_per_sample] : 1.f;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    if (k < nv) {
      const int c = k * 32 + lane;
      float t = 0.f;
      if (c < C) {
        t = to_f(x[row * C + c]);
        if (r != nullptr) t = fmaf(s, to_f(r[row * C + c]), t);
        if (xs_out != nullptr) xs_out[row * C + c] = from_f<TS>(t);
        if (xs_out != nullptr) t = to_f(from_f<TS>(t));  // normalise exactly what the stream stores
      }
      v[k] = t;
      sum += t;
    }
  }
  const float mu = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (k < nv) {
      const int c = k * 32 + lane;
      const float d = (c < C) ? v[k] - mu : 0.f;
      sq = fmaf(d, d, sq);
    }
  const float rs_ = rsqrtf(warp_sum(sq) / (float)C + eps);
  if (y != nullptr) {
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
      if (k < nv) {
        const int c = k * 32 + lane;
        if (c < C) y[row * C + c] = from_f<TC>((v[k] - mu) * rs_ * gamma[c] + beta[c]);
      }
  }
  if (lane == 0 && mean != nullptr) { mean[row] = mu; rstd[row] = rs_; }
}

// dxs = dres + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * gamma ; partial dgamma/dbeta per CTA
template <int MAXV, typename TS, typename TC, typename TG>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const TC* __restrict__ dy, const TS* __restrict__ xs,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const TG* __restrict__ dres,
                                                     TG* __restrict__ dxs, TC* __restrict__ dr,
                                                     const float* __restrict__ rs, int rows_per_sample,
                                                     float* __restrict__ part_g, float* __restrict__ part_b,
                                                     long long rows, int C) {
  extern __shared__ float sm[];  // [8 warps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nv = (C + 31) / 32;
  float ag[MAXV], ab[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) { ag[k] = 0.f; ab[k] = 0.f; }
  for (long long row = (long long)blockIdx.x * nwarp + warp; row < rows; row += (long long)gridDim.x * nwarp) {
    const float mu = mean[row], rs_ = rstd[row];
    const float bs = (rs != nullptr) ? rs[row / rows_per_sample] : 1.f;
    float xh[MAXV], g[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
      if (k < nv) {
        const int c = k * 32 + lane;
        float xv = 0.f, dv = 0.f, gm = 0.f;
        if (c < C) { xv = (to_f(xs[row * C + c]) - mu) * rs_; dv = to_f(dy[row * C + c]); gm = gamma[c]; }
        xh[k] = xv;
        g[k] = dv * gm;
        ag[k] = fmaf(dv, xv, ag[k]);
        ab[k] += dv;
        s1 += g[k];
        s2 = fmaf(g[k], xv, s2);
      }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
      if (k < nv) {
        const int c = k * 32 + lane;
        if (c < C) {
          float d = rs_ * (g[k] - s1 - xh[k] * s2);
          if (dres != nullptr) d += to_f(dres[row * C + c]);
          dxs[row * C + c] = from_f<TG>(d);
          if (dr != nullptr) dr[row * C + c] = from_f<TC>(bs * d);   // gradient of the pending residual branch
        }
      }
  }
  float* sg = sm + (size_t)warp * 2 * C;
#pragma unroll
  for (int k = 0; k < MAXV; ++k)
    if (k < nv) {
      const int c = k * 32 + lane;
      if (c < C) { sg[c] = ag[k]; sg[C + c] = ab[k]; }
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float tg = 0.f, tb = 0.f;
    for (int w2 = 0; w2 < nwarp; ++w2) { tg += sm[(size_t)w2 * 2 * C + c]; tb += sm[(size_t)w2 * 2 * C + C + c]; }
    part_g[(size_t)blockIdx.x * C + c] = tg;
    part_b[(size_t)blockIdx.x * C + c] = tb;
  }
}

// out[c] (+)= sum_r part[r][c]   (fixed order -> deterministic).  block = 32 columns x 32 row phases (a few hundred
// partial rows: the loads of a phase are independent, 4 in flight per thread);
// blockIdx.y selects one of up to two (partials, output) pairs so dgamma and dbeta share a launch.
__global__ void __launch_bounds__(1024) colsum_partials_kernel(const float* __restrict__ part0, float* __restrict__ out0,
                                                               const float* __restrict__ part1, float* __restrict__ out1,
                                                               int nparts, int C, int accumulate) {
  __shared__ float s[32][33];
  pdl_wait();
  pdl_trigger();
  const float* part = blockIdx.y == 0 ? part0 : part1;
  float* out = blockIdx.y == 0 ? out0 : out1;
  const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < C) {
    int r = ph;
    for (; r + 96 < nparts; r += 128) {
      const float v0 = part[(size_t)r * C + c], v1 = part[(size_t)(r + 32) * C + c];
      const float v2 = part[(size_t)(r + 64) * C + c], v3 = part[(size_t)(r + 96) * C + c];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; r < nparts; r += 32) acc += part[(size_t)r * C + c];
  }
  s[ph][lane] = acc;
  __syncthreads();
  if (ph == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += s[k][lane];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// column sums of a [rows, C] matrix: out[c] = sum_r a[r][c]; two-stage, deterministic.
template <typename T>
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const T* __restrict__ a, long long rows, int C,
                                                            float* __restrict__ part, int rows_per_cta) {
  // block = 32 x 8 : threadIdx.x -> column within a 32-wide strip, threadIdx.y -> row phase
  __shared__ float s[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + ry; r < r1; r += 8) acc += to_f(a[r * C + c]);
  s[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][threadIdx.x & 31];
    part[(size_t)blockIdx.y * C + c] = t;
  }
}

}  // namespace

int ln_bwd_v4_launch(const void* dy, const void* xs, const float* mean, const float* rstd, const float* gamma,
                     const void* dres, void* dxs, void* dr, const float* rs, int rows_per_sample, float* pg, float* pb,
                     int grid, long long rows, int C, int sdtype, int cdtype, cudaStream_t st);
int colsum_v8_launch(const void* a, long long rows, int C, float* part, int dtype, cudaStream_t st);
void colsum_plan(long long rows, int C, int* gx, int* TX, int* rows_per_cta, int* parts);
int ln_fwd_v4_launch(const void* x, const void* r, const float* rs, int rows_per_sample, const float* gamma, const float* beta,
                     void* xs_out, void* y, float* mean, float* rstd, long long rows, int C, float eps, int sdtype, int cdtype,
                     cudaStream_t st);

#define LN_GRID_BWD (148 * 4)   // persistent CTAs of the backward kernel (<= this many partial rows)

long long apb_ln_bwd_workspace_floats(int C) { return 2LL * LN_GRID_BWD * C; }

// sdtype: dtype of the residual stream (x, xs_out); cdtype: dtype of r and y.
int apb_ln_fwd(const void* x, const void* r, const float* rs, int rows_per_sample, const float* gamma, const float* beta,
               void* xs_out, void* y, float* mean, float* rstd, long long rows, int C, float eps, int sdtype, int cdtype,
               apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(C > 0 && C <= MAXC, APB_ERR_SHAPE, "ln_fwd: C=%d unsupported (max %d)", C, MAXC);
  if (rows <= 0) return 0;
  const int grid = ceil_div(rows, 8);
  const int nv = (C + 31) / 32;
  if (ln_fwd_v4_launch(x, r, rs, rows_per_sample, gamma, beta, xs_out, y, mean, rstd, rows, C, eps, sdtype, cdtype, st) == 1) {
    APB_LAUNCH_CHECK("ln_fwd_v4");
    return 0;
  }
#define LN_FWD_NV(NV_, TS_, TC_)                                                                                  \
  ln_fwd_kernel<NV_, TS_, TC_, TC_><<<grid, 256, 0, st>>>((const TS_*)x, (const TC_*)r, rs, rows_per_sample, gamma, \
                                                           beta, (TS_*)xs_out, (TC_*)y, mean, rstd, rows, C, eps)
#define LN_FWD(TS_, TC_)                                    \
  do {                                                      \
    if (nv <= 2) LN_FWD_NV(2, TS_, TC_);                    \
    else if (nv <= 4) LN_FWD_NV(4, TS_, TC_);               \
    else if (nv <= 6) LN_FWD_NV(6, TS_, TC_);               \
    else if (nv <= 8) LN_FWD_NV(8, TS_, TC_);               \
    else if (nv <= 12) LN_FWD_NV(12, TS_, TC_);             \
    else if (nv <= 16) LN_FWD_NV(16, TS_, TC_);             \
    else if (nv <= 24) LN_FWD_NV(24, TS_, TC_);             \
    else LN_FWD_NV(32, TS_, TC_);                           \
  } while (0)
  if (sdtype == APB_F32 && cdtype == APB_F32) LN_FWD(float, float);
  else if (sdtype == APB_F32 && cdtype == APB_BF16) LN_FWD(float, bf16);
  else if (sdtype == APB_BF16 && cdtype == APB_BF16) LN_FWD(bf16, bf16);
  else { apb_set_error("ln_fwd: unsupported dtype pair (%d,%d)", sdtype, cdtype); return APB_ERR_DTYPE; }
#undef LN_FWD
#undef LN_FWD_NV
  APB_LAUNCH_CHECK("ln_fwd");
  return 0;
}

// dgamma/dbeta are fp32 [C]; accumulate!=0 adds into them.  workspace: apb_ln_bwd_workspace_floats(C) floats.
int apb_ln_bwd(const void* dy, const void* xs, const float* mean, const float* rstd, const float* gamma, const void* dres,
               void* dxs, void* dr, const float* rs, int rows_per_sample, float* dgamma, float* dbeta, int accumulate,
               float* workspace, long long rows, int C, int sdtype, int cdtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(C > 0 && C <= MAXC, APB_ERR_SHAPE, "ln_bwd: C=%d unsupported (max %d)", C, MAXC);
  if (rows <= 0) return 0;
  int grid = ceil_div(rows, 8);
  if (grid > LN_GRID_BWD) grid = LN_GRID_BWD;
  float* pg = workspace;
  float* pb = workspace + (size_t)LN_GRID_BWD * C;
  const size_t smem = (size_t)8 * 2 * C * sizeof(float);
  const int nv = (C + 31) / 32;
  {
    const int handled = ln_bwd_v4_launch(dy, xs, mean, rstd, gamma, dres, dxs, dr, rs, rows_per_sample, pg, pb, grid, rows, C,
                                         sdtype, cdtype, st);
    if (handled == 1) {
      APB_LAUNCH_CHECK("ln_bwd_v4");
      apb_launch_pdl(colsum_partials_kernel, dim3(ceil_div(C, 32), 2), dim3(1024), 0, st, pg, dgamma, pb, dbeta, grid, C, accumulate);
      APB_LAUNCH_CHECK("ln_bwd_reduce");
      return 0;
    }
  }
#define LN_BWD_NV(NV_, TS_, TC_)                                                                                       \
  do {                                                                                                                 \
    cudaFuncSetAttribute(ln_bwd_kernel<NV_, TS_, TC_, TS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    ln_bwd_kernel<NV_, TS_, TC_, TS_><<<grid, 256, smem, st>>>((const TC_*)dy, (const TS_*)xs, mean, rstd, gamma,      \
                                                                (const TS_*)dres, (TS_*)dxs, (TC_*)dr, rs,             \
                                                                rows_per_sample, pg, pb, rows, C);                     \
  } while (0)
#define LN_BWD(TS_, TC_)                                    \
  do {                                                      \
    if (nv <= 2) LN_BWD_NV(2, TS_, TC_);                    \
    else if (nv <= 4) LN_BWD_NV(4, TS_, TC_);               \
    else if (nv <= 6) LN_BWD_NV(6, TS_, TC_);               \
    else if (nv <= 8) LN_BWD_NV(8, TS_, TC_);               \
    else if (nv <= 12) LN_BWD_NV(12, TS_, TC_);             \
    else if (nv <= 16) LN_BWD_NV(16, TS_, TC_);             \
    else if (nv <= 24) LN_BWD_NV(24, TS_, TC_);             \
    else LN_BWD_NV(32, TS_, TC_);                           \
  } while (0)
  if (sdtype == APB_F32 && cdtype == APB_F32) LN_BWD(float, float);
  else if (sdtype == APB_F32 && cdtype == APB_BF16) LN_BWD(float, bf16);
  else if (sdtype == APB_BF16 && cdtype == APB_BF16) LN_BWD(bf16, bf16);
  else { apb_set_error("ln_bwd: unsupported dtype pair (%d,%d)", sdtype, cdtype); return APB_ERR_DTYPE; }
#undef LN_BWD
#undef LN_BWD_NV
  APB_LAUNCH_CHECK("ln_bwd");
  apb_launch_pdl(colsum_partials_kernel, dim3(ceil_div(C, 32), 2), dim3(1024), 0, st, pg, dgamma, pb, dbeta, grid, C, accumulate);
  APB_LAUNCH_CHECK("ln_bwd_reduce");
  return 0;
}

long long apb_colsum_workspace_floats(long long rows, int C) {
  if (rows <= 0 || C <= 0) return 0;
  int gx, TX, rpc, parts;
  colsum_plan(rows, C, &gx, &TX, &rpc, &parts);
  const long long scalar_parts = (rows + 511) / 512;       // fallback kernel (C % 8 != 0 or unaligned)
  return (long long)(parts > scalar_parts ? parts : scalar_parts) * C;
}

// out[c] = sum_r a[r][c]  (bias gradients; batch reduction of the pos-embed gradient)
int apb_colsum(const void* a, long long rows, int C, float* out, int accumulate, float* workspace, int dtype,
               apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (rows <= 0 || C <= 0) return 0;
  int parts;
  if (colsum_v8_launch(a, rows, C, workspace, dtype, st) == 1) {
    int gx, TX, rpc;
    colsum_plan(rows, C, &gx, &TX, &rpc, &parts);
  } else {
    const int rows_per_cta = 512;
    parts = (int)((rows + rows_per_cta - 1) / rows_per_cta);
    dim3 grid(ceil_div(C, 32), parts);
    if (dtype == APB_F32) colsum_stage1_kernel<float><<<grid, 256, 0, st>>>((const float*)a, rows, C, workspace, rows_per_cta);
    else colsum_stage1_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)a, rows, C, workspace, rows_per_cta);
  }
  APB_LAUNCH_CHECK("colsum_stage1");
  apb_launch_pdl(colsum_partials_kernel, dim3(ceil_div(C, 32), 1), dim3(1024), 0, st, workspace, out, nullptr, nullptr, parts, C, accumulate);
  APB_LAUNCH_CHECK("colsum_stage2");
  return 0;
}
