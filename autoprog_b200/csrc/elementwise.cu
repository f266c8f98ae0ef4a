// Elementwise / layout kernels of the VOLO hot path (all HBM-bound, coalesced along the channel dim).
#include "common.cuh"

namespace {

constexpr int EW_THREADS = 256;
inline int ew_grid(long long n, int per_thread = 1) {
  long long g = (n + (long long)EW_THREADS * per_thread - 1) / ((long long)EW_THREADS * per_thread);
  const long long cap = 148LL * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- AvgPool2d(2,2,ceil_mode=True) on NHWC (models/volo.py:75,87): edge windows divide by in-bounds count
template <typename T>
__global__ void avgpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int h, int w) {
  const long long n = (long long)B * h * w * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int j = (int)(r % w); r /= w;
    const int i = (int)(r % h);
    const int b = (int)(r / h);
    float s = 0.f;
    int cnt = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * i + dy, xx = 2 * j + dx;
        if (yy < H && xx < W) { s += to_f(x[(((size_t)b * H + yy) * W + xx) * C + c]); ++cnt; }
      }
    y[idx] = from_f<T>(s / (float)cnt);
  }
}

// vector form: one thread = 8 contiguous bf16 channels (16 bytes) of one output pixel
__global__ void avgpool2_fwd_v8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C, int h, int w) {
  const int cv = C >> 3;
  const long long n = (long long)B * h * w * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * 8;
    long long r = idx / cv;
    const int j = (int)(r % w); r /= w;
    const int i = (int)(r % h);
    const int b = (int)(r / h);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * i + dy, xx = 2 * j + dx;
        if (yy < H && xx < W) {
          const uint4 u = *reinterpret_cast<const uint4*>(x + (((size_t)b * H + yy) * W + xx) * C + c);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int t = 0; t < 4; ++t) { const float2 f = __bfloat1622float2(h2[t]); s[2 * t] += f.x; s[2 * t + 1] += f.y; }
          ++cnt;
        }
      }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(s[2 * t] / (float)cnt, s[2 * t + 1] / (float)cnt);
    *reinterpret_cast<uint4*>(y + idx * 8) = o;
  }
}

// one thread = VEC contiguous channels (16 bytes)
template <typename T, int VEC>
__global__ void avgpool2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int H, int W, int C, int h, int w,
                                    int accumulate) {
  const int cv = C / VEC;
  const long long n = (long long)B * H * W * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * VEC;
    long long r = idx / cv;
    const int xx = (int)(r % W); r /= W;
    const int yy = (int)(r % H);
    const int b = (int)(r / H);
    const int i = yy >> 1, j = xx >> 1;
    const float inv = 1.f / (float)(((2 * i + 1 < H) ? 2 : 1) * ((2 * j + 1 < W) ? 2 : 1));
    const T* src = dy + (((size_t)b * h + i) * w + j) * C + c;
    T* dst = dx + (size_t)(idx / cv) * C + c;
    if (VEC * sizeof(T) == 16) {
      Vec16<T> g, o;
      g.load(src);
      if (accumulate) o.load(dst);
#pragma unroll
      for (int k = 0; k < Vec16<T>::N; ++k) o.set(k, g.get(k) * inv + (accumulate ? o.get(k) : 0.f));
      o.store(dst);
    } else {
      float g = to_f(*src) * inv;
      if (accumulate) g += to_f(*dst);
      *dst = from_f<T>(g);
    }
  }
}

// ---- mix-token / un-mix (models/volo.py:655-658, 687-689)
template <typename T, int VEC>
__global__ void flip_in_box_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int r0, int c0,
                                   int r1, int c1, const int* __restrict__ box_dev, int box_scale) {
  if (box_dev != nullptr) {   // CUDA-graph path: the box lives in device memory and changes between replays
    r0 = box_dev[0] * box_scale; c0 = box_dev[1] * box_scale; r1 = box_dev[2] * box_scale; c1 = box_dev[3] * box_scale;
  }
  const int cv = C / VEC;
  const long long n = (long long)B * H * W * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * VEC;
    long long r = idx / cv;
    const int xx = (int)(r % W); r /= W;
    const int yy = (int)(r % H);
    const int b = (int)(r / H);
    const bool inside = (yy >= r0 && yy < r1 && xx >= c0 && xx < c1);
    const size_t pix = (size_t)(idx / cv);
    const size_t spix = inside ? pix + (size_t)((long long)(B - 1 - 2 * b) * H * W) : pix;
    if (VEC * sizeof(T) == 16) *reinterpret_cast<uint4*>(y + pix * C + c) = *reinterpret_cast<const uint4*>(x + spix * C + c);
    else y[pix * C + c] = x[spix * C + c];
  }
}

// ---- patchify: [B,H,W,C] -> [B*(H/p)*(W/p), p*p*C], K order (kh,kw,c); floor semantics (conv stride p, no padding)
// One thread moves VEC contiguous channels (16 bytes when C allows); the (kw, c) run of a patch row is contiguous
// on both sides, so accesses are fully coalesced.
template <typename T, int VEC, bool INVERSE>
__global__ void patchify_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int H, int W, int C, int p, int hp,
                                int wp) {
  const int cv = C / VEC;
  const long long n = (long long)B * hp * wp * p * p * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * VEC;
    long long r = idx / cv;
    const int kw = (int)(r % p); r /= p;
    const int kh = (int)(r % p); r /= p;
    const int pj = (int)(r % wp); r /= wp;
    const int pi = (int)(r % hp);
    const int b = (int)(r / hp);
    const size_t img = (((size_t)b * H + (pi * p + kh)) * W + (pj * p + kw)) * C + c;
    const size_t row = (size_t)(idx / cv) * C + c;
    const T* s = INVERSE ? src + row : src + img;
    T* d = INVERSE ? dst + img : dst + row;
    if (VEC * sizeof(T) == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
    else *d = *s;
  }
}

template <typename T>
__global__ void zero_kernel(T* __restrict__ p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = from_f<T>(0.f);
}

// ---- bicubic (A = -0.75, align_corners=False, scale_factor semantics) -- SURVEY.md A.3
__device__ __forceinline__ void cubic_taps(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__device__ __forceinline__ void cubic_src(int o, float inv_sf, int n_in, int (&idx)[4], float (&w)[4]) {
  const float src = ((float)o + 0.5f) * inv_sf - 0.5f;
  const float fl = floorf(src);
  cubic_taps(src - fl, w);
  const int i0 = (int)fl;
#pragma unroll
  for (int k = 0; k < 4; ++k) idx[k] = min(max(i0 - 1 + k, 0), n_in - 1);
}

__global__ void bicubic_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst, int h, int w, int h0, int w0,
                                   int C, float inv_sy, float inv_sx) {
  const long long n = (long long)h0 * w0 * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int ox = (int)((idx / C) % w0);
    const int oy = (int)(idx / ((long long)C * w0));
    int iy[4], ix[4];
    float wy[4], wx[4];
    cubic_src(oy, inv_sy, h, iy, wy);
    cubic_src(ox, inv_sx, w, ix, wx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) row = fmaf(wx[b], src[((size_t)iy[a] * w + ix[b]) * C + c], row);
      acc = fmaf(wy[a], row, acc);
    }
    dst[idx] = acc;
  }
}

// transpose of the above as a gather over output pixels (deterministic): one thread per (src pixel, channel)
__global__ void bicubic_bwd_kernel(const float* __restrict__ ddst, float* __restrict__ dsrc, int h, int w, int h0, int w0,
                                   int C, float inv_sy, float inv_sx) {
  const long long n = (long long)h * w * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int sx = (int)((idx / C) % w);
    const int sy = (int)(idx / ((long long)C * w));
    float acc = 0.f;
    for (int oy = 0; oy < h0; ++oy) {
      int iy[4];
      float wy[4];
      cubic_src(oy, inv_sy, h, iy, wy);
      float cy = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) cy += (iy[a] == sy) ? wy[a] : 0.f;
      if (cy == 0.f) continue;
      for (int ox = 0; ox < w0; ++ox) {
        int ix[4];
        float wx[4];
        cubic_src(ox, inv_sx, w, ix, wx);
        float cx = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) cx += (ix[b] == sx) ? wx[b] : 0.f;
        if (cx != 0.f) acc = fmaf(cy * cx, ddst[((size_t)oy * w0 + ox) * C + c], acc);
      }
    }
    dsrc[idx] = acc;
  }
}

template <typename TI, typename TO>
__global__ void add_bcast_kernel(const TI* __restrict__ x, const float* __restrict__ p, TO* __restrict__ out, long long batch,
                                 long long inner) {
  const long long n = batch * inner;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = from_f<TO>(to_f(x[idx]) + p[idx % inner]);
}

template <typename TI, typename TO>
__global__ void scale_cast_kernel(const TI* __restrict__ in, const float* __restrict__ rs, TO* __restrict__ out,
                                  long long batch, long long inner) {
  const long long n = batch * inner;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = from_f<TO>(to_f(in[idx]) * (rs != nullptr ? rs[idx / inner] : 1.f));
}

template <typename TX, typename TR, typename TO>
__global__ void residual_add_kernel(const TX* __restrict__ x, const TR* __restrict__ r, const float* __restrict__ rs,
                                    TO* __restrict__ out, long long batch, long long inner) {
  const long long n = batch * inner;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = from_f<TO>(fmaf(rs != nullptr ? rs[idx / inner] : 1.f, to_f(r[idx]), to_f(x[idx])));
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = from_f<TO>(to_f(in[idx]));
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    out[idx] = from_f<T>(to_f(a[idx]) + to_f(b[idx]));
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <typename T>
__global__ void gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    y[idx] = from_f<T>(gelu_f(to_f(x[idx])));
}
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long n) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x)
    dx[idx] = from_f<T>(to_f(dy[idx]) * dgelu_f(to_f(x[idx])));
}


// Input resolution switch of the progressive schedule (main_prog.py:973-974, 1910): F.interpolate(x, size=(r, r),
// mode='bilinear', align_corners=False) on an NCHW batch.  ATen semantics: src = max((dst + 0.5) * in/out - 0.5, 0),
// i0 = floor(src), i1 = min(i0 + 1, in - 1), weights (1 - l, l).  One thread per output pixel, all planes of the image
// share the index / weight computation via the grid's y dimension (plane = b * C + c).  Output fp32 or bf16.
template <typename TO>
__global__ void __launch_bounds__(EW_THREADS) bilinear_resize_kernel(const float* __restrict__ src, TO* __restrict__ dst, int H,
                                                                     int W, int OH, int OW, float sy, float sx) {
  const float* sp = src + (size_t)blockIdx.y * H * W;
  TO* dp = dst + (size_t)blockIdx.y * OH * OW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < OH * OW; i += gridDim.x * blockDim.x) {
    const int oy = i / OW, ox = i - oy * OW;
    const float fy = fmaxf(((float)oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf(((float)ox + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float v = hy * (hx * sp[y0 * W + x0] + lx * sp[y0 * W + x1]) + ly * (hx * sp[y1 * W + x0] + lx * sp[y1 * W + x1]);
    dp[i] = from_f<TO>(v);
  }
}

// im2col for a small-channel convolution (the 7x7 / stride-2 stem conv, models/volo.py:352): row = output pixel
// (b, oy, ox), column k = c * KH * KW + ky * KW + kx (the memory order of an nn.Conv2d weight row), zero beyond
// C * KH * KW (columns are padded to a multiple of 8 for the TMA row pitch) and for taps outside the image.
// One CTA per ROWS output rows of one image: the C * ((ROWS-1)*stride + KH) input rows it needs are staged ONCE in
// shared memory (warp per row, coalesced, zero-filled borders); a per-CTA table maps column k to its offset in the
// staged tile, so a thread produces 8 consecutive columns with one 16-byte table read + 8 gathers and writes them as
// one 16-byte store.  The input is read through arbitrary element strides (NCHW or NHWC, fp32 or bf16).
constexpr int IM2COL_ROWS = 4;
template <typename T>
__global__ void __launch_bounds__(256) im2col_rows_kernel(const T* __restrict__ x, bf16* __restrict__ col, int C, int H, int W,
                                                          int KH, int KW, int stride, int pad, int OH, int OW, int Kpad,
                                                          long long sb, long long sc, long long sh, long long sw,
                                                          unsigned inv_chunks) {
  extern __shared__ __align__(16) unsigned char im2col_smem[];
  const int SW = W + 2 * pad, IR = (IM2COL_ROWS - 1) * stride + KH;
  short* lut = reinterpret_cast<short*>(im2col_smem);                      // [Kpad]
  float* srow = reinterpret_cast<float*>(im2col_smem + ((Kpad * 2 + 15) & ~15));   // [C * IR][SW]
  const int b = blockIdx.y, oy0 = blockIdx.x * IM2COL_ROWS;
  const int Kreal = C * KH * KW;
  for (int k = threadIdx.x; k < Kpad; k += 256) {
    int off = -1;
    if (k < Kreal) {
      const int c = k / (KH * KW), r = k - c * KH * KW, ky = r / KW, kx = r - ky * KW;
      off = (c * IR + ky) * SW + kx;
    }
    lut[k] = (short)off;
  }
  const T* xb = x + (long long)b * sb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iy0 = oy0 * stride - pad;
  for (int rr = warp; rr < C * IR; rr += 8) {
    const int c = rr / IR, iy = iy0 + (rr - c * IR);
    const bool yok = iy >= 0 && iy < H;
    const T* g = xb + c * sc + (long long)(yok ? iy : 0) * sh;
    float* d = srow + rr * SW;
    for (int xx = lane; xx < SW; xx += 32) {
      const int ix = xx - pad;
      d[xx] = (yok && ix >= 0 && ix < W) ? to_f(g[ix * sw]) : 0.f;
    }
  }
  __syncthreads();
  const int chunks = Kpad >> 3;
  for (int orow = 0; orow < IM2COL_ROWS; ++orow) {
    const int oy = oy0 + orow;
    if (oy >= OH) break;
    bf16* out = col + ((size_t)((size_t)b * OH + oy) * OW) * Kpad;
    const float* rbase = srow + orow * stride * SW;
    for (int e = threadIdx.x; e < OW * chunks; e += 256) {
      const int ox = (int)__umulhi((unsigned)e, inv_chunks);               // e / chunks
      const int ch = e - ox * chunks;
      const uint4 lt = *reinterpret_cast<const uint4*>(lut + ch * 8);
      const short* o8 = reinterpret_cast<const short*>(&lt);
      const float* base = rbase + ox * stride;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = o8[j] >= 0 ? base[o8[j]] : 0.f;
      uint4 pk;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4*>(out + (size_t)e * 8) = pk;
    }
  }
}

}  // namespace

#define DISPATCH_T(dtype, name, CALL_F32, CALL_BF16)                                 \
  do {                                                                               \
    if ((dtype) == APB_F32) { CALL_F32; }                                            \
    else if ((dtype) == APB_BF16) { CALL_BF16; }                                     \
    else { apb_set_error("%s: dtype %d", name, (int)(dtype)); return APB_ERR_DTYPE; } \
    APB_LAUNCH_CHECK(name);                                                          \
  } while (0)

int apb_avgpool2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  const int h = (H + 1) / 2, w = (W + 1) / 2;
  const long long n = (long long)B * h * w * C;
  if (n <= 0) return 0;
  if (dtype == APB_BF16 && C % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
    avgpool2_fwd_v8_kernel<<<ew_grid(n / 8), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)y, B, H, W, C, h, w);
    APB_LAUNCH_CHECK("avgpool2_fwd");
    return 0;
  }
  DISPATCH_T(dtype, "avgpool2_fwd",
             (avgpool2_fwd_kernel<float><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)x, (float*)y, B, H, W, C, h, w)),
             (avgpool2_fwd_kernel<bf16><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)y, B, H, W, C, h, w)));
  return 0;
}

int apb_avgpool2_bwd(const void* dy, void* dx, int B, int H, int W, int C, int accumulate, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  const int h = (H + 1) / 2, w = (W + 1) / 2;
  const long long n = (long long)B * H * W * C;
  if (n <= 0) return 0;
  const bool al = (((uintptr_t)dy | (uintptr_t)dx) & 15) == 0;
  if (dtype == APB_F32 && C % 4 == 0 && al) {
    avgpool2_bwd_kernel<float, 4><<<ew_grid(n / 4), EW_THREADS, 0, st>>>((const float*)dy, (float*)dx, B, H, W, C, h, w, accumulate);
    APB_LAUNCH_CHECK("avgpool2_bwd");
    return 0;
  }
  if (dtype == APB_BF16 && C % 8 == 0 && al) {
    avgpool2_bwd_kernel<bf16, 8><<<ew_grid(n / 8), EW_THREADS, 0, st>>>((const bf16*)dy, (bf16*)dx, B, H, W, C, h, w, accumulate);
    APB_LAUNCH_CHECK("avgpool2_bwd");
    return 0;
  }
  DISPATCH_T(dtype, "avgpool2_bwd",
             (avgpool2_bwd_kernel<float, 1><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)dy, (float*)dx, B, H, W, C, h, w, accumulate)),
             (avgpool2_bwd_kernel<bf16, 1><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)dy, (bf16*)dx, B, H, W, C, h, w, accumulate)));
  return 0;
}

static int flip_impl(const void* x, void* y, int B, int H, int W, int C, int r0, int c0, int r1, int c1, const int* box_dev,
                     int box_scale, int dtype, cudaStream_t st);

int apb_flip_in_box(const void* x, void* y, int B, int H, int W, int C, int r0, int c0, int r1, int c1, int dtype,
                    apb_stream_t stream) {
  return flip_impl(x, y, B, H, W, C, r0, c0, r1, c1, nullptr, 1, dtype, APB_STREAM(stream));
}

int apb_flip_in_box_dev(const void* x, void* y, int B, int H, int W, int C, const int* box_dev, int box_scale, int dtype,
                        apb_stream_t stream) {
  APB_CHECK_ARG(box_dev != nullptr && box_scale >= 1, APB_ERR_ARG, "flip_in_box_dev: box_dev/scale");
  return flip_impl(x, y, B, H, W, C, 0, 0, 0, 0, box_dev, box_scale, dtype, APB_STREAM(stream));
}

static int flip_impl(const void* x, void* y, int B, int H, int W, int C, int r0, int c0, int r1, int c1, const int* box_dev,
                     int box_scale, int dtype, cudaStream_t st) {
  APB_CHECK_ARG(x != y, APB_ERR_ARG, "flip_in_box: must be out of place");
  const long long n = (long long)B * H * W * C;
  if (n <= 0) return 0;
  const bool al = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  if (dtype == APB_F32 && C % 4 == 0 && al) {
    flip_in_box_kernel<float, 4><<<ew_grid(n / 4), EW_THREADS, 0, st>>>((const float*)x, (float*)y, B, H, W, C, r0, c0, r1, c1, box_dev, box_scale);
    APB_LAUNCH_CHECK("flip_in_box");
    return 0;
  }
  if (dtype == APB_BF16 && C % 8 == 0 && al) {
    flip_in_box_kernel<bf16, 8><<<ew_grid(n / 8), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)y, B, H, W, C, r0, c0, r1, c1, box_dev, box_scale);
    APB_LAUNCH_CHECK("flip_in_box");
    return 0;
  }
  DISPATCH_T(dtype, "flip_in_box",
             (flip_in_box_kernel<float, 1><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)x, (float*)y, B, H, W, C, r0, c0, r1, c1, box_dev, box_scale)),
             (flip_in_box_kernel<bf16, 1><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)y, B, H, W, C, r0, c0, r1, c1, box_dev, box_scale)));
  return 0;
}

int apb_patchify(const void* x, void* rows, int B, int H, int W, int C, int p, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(p > 0 && H >= p && W >= p, APB_ERR_SHAPE, "patchify: H=%d W=%d p=%d", H, W, p);
  const int hp = H / p, wp = W / p;
  const long long n = (long long)B * hp * wp * p * p * C;
  const bool al = (((uintptr_t)x | (uintptr_t)rows) & 15) == 0;
  if (dtype == APB_F32 && C % 4 == 0 && al) {
    patchify_kernel<float, 4, false><<<ew_grid(n / 4), EW_THREADS, 0, st>>>((const float*)x, (float*)rows, B, H, W, C, p, hp, wp);
    APB_LAUNCH_CHECK("patchify");
    return 0;
  }
  if (dtype == APB_BF16 && C % 8 == 0 && al) {
    patchify_kernel<bf16, 8, false><<<ew_grid(n / 8), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)rows, B, H, W, C, p, hp, wp);
    APB_LAUNCH_CHECK("patchify");
    return 0;
  }
  DISPATCH_T(dtype, "patchify",
             (patchify_kernel<float, 1, false><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)x, (float*)rows, B, H, W, C, p, hp, wp)),
             (patchify_kernel<bf16, 1, false><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)rows, B, H, W, C, p, hp, wp)));
  return 0;
}

int apb_unpatchify(const void* rows, void* x, int B, int H, int W, int C, int p, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(p > 0 && H >= p && W >= p, APB_ERR_SHAPE, "unpatchify: H=%d W=%d p=%d", H, W, p);
  const int hp = H / p, wp = W / p;
  const long long n = (long long)B * hp * wp * p * p * C;
  const long long nimg = (long long)B * H * W * C;
  if (hp * p != H || wp * p != W) {  // pixels not covered by any patch get zero gradient
    DISPATCH_T(dtype, "unpatchify_zero", (zero_kernel<float><<<ew_grid(nimg), EW_THREADS, 0, st>>>((float*)x, nimg)),
               (zero_kernel<bf16><<<ew_grid(nimg), EW_THREADS, 0, st>>>((bf16*)x, nimg)));
  }
  const bool al = (((uintptr_t)x | (uintptr_t)rows) & 15) == 0;
  if (dtype == APB_F32 && C % 4 == 0 && al) {
    patchify_kernel<float, 4, true><<<ew_grid(n / 4), EW_THREADS, 0, st>>>((const float*)rows, (float*)x, B, H, W, C, p, hp, wp);
    APB_LAUNCH_CHECK("unpatchify");
    return 0;
  }
  if (dtype == APB_BF16 && C % 8 == 0 && al) {
    patchify_kernel<bf16, 8, true><<<ew_grid(n / 8), EW_THREADS, 0, st>>>((const bf16*)rows, (bf16*)x, B, H, W, C, p, hp, wp);
    APB_LAUNCH_CHECK("unpatchify");
    return 0;
  }
  DISPATCH_T(dtype, "unpatchify",
             (patchify_kernel<float, 1, true><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)rows, (float*)x, B, H, W, C, p, hp, wp)),
             (patchify_kernel<bf16, 1, true><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)rows, (bf16*)x, B, H, W, C, p, hp, wp)));
  return 0;
}

int apb_bicubic_resize(const float* src, float* dst, int h, int w, int h0, int w0, int C, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(h > 0 && w > 0 && h0 > 0 && w0 > 0 && C > 0, APB_ERR_SHAPE, "bicubic_resize: bad shape");
  // F.interpolate(scale_factor=(h0+0.1)/h) maps with 1/scale_factor (models/volo.py:588-593)
  const float inv_sy = (float)((double)h / ((double)h0 + 0.1)), inv_sx = (float)((double)w / ((double)w0 + 0.1));
  const long long n = (long long)h0 * w0 * C;
  bicubic_fwd_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(src, dst, h, w, h0, w0, C, inv_sy, inv_sx);
  APB_LAUNCH_CHECK("bicubic_resize");
  return 0;
}

int apb_bicubic_resize_bwd(const float* ddst, float* dsrc, int h, int w, int h0, int w0, int C, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(h > 0 && w > 0 && h0 > 0 && w0 > 0 && C > 0, APB_ERR_SHAPE, "bicubic_resize_bwd: bad shape");
  const float inv_sy = (float)((double)h / ((double)h0 + 0.1)), inv_sx = (float)((double)w / ((double)w0 + 0.1));
  const long long n = (long long)h * w * C;
  bicubic_bwd_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(ddst, dsrc, h, w, h0, w0, C, inv_sy, inv_sx);
  APB_LAUNCH_CHECK("bicubic_resize_bwd");
  return 0;
}

int apb_add_bcast(const void* x, const float* p, void* out, long long batch, long long inner, int in_dtype,
                  int out_dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  const long long n = batch * inner;
  if (n <= 0) return 0;
  const int g = ew_grid(n);
#define ARGS_(TI_, TO_) (const TI_*)x, p, (TO_*)out, batch, inner
  if (in_dtype == APB_F32 && out_dtype == APB_F32) add_bcast_kernel<float, float><<<g, EW_THREADS, 0, st>>>(ARGS_(float, float));
  else if (in_dtype == APB_BF16 && out_dtype == APB_F32) add_bcast_kernel<bf16, float><<<g, EW_THREADS, 0, st>>>(ARGS_(bf16, float));
  else if (in_dtype == APB_BF16 && out_dtype == APB_BF16) add_bcast_kernel<bf16, bf16><<<g, EW_THREADS, 0, st>>>(ARGS_(bf16, bf16));
  else if (in_dtype == APB_F32 && out_dtype == APB_BF16) add_bcast_kernel<float, bf16><<<g, EW_THREADS, 0, st>>>(ARGS_(float, bf16));
  else { apb_set_error("add_bcast: dtypes %d->%d", in_dtype, out_dtype); return APB_ERR_DTYPE; }
#undef ARGS_
  APB_LAUNCH_CHECK("add_bcast");
  return 0;
}

int apb_scale_cast(const void* in, const float* rs, void* out, long long batch, long long inner, int in_dtype,
                   int out_dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  const long long n = batch * inner;
  if (n <= 0) return 0;
  const int g = ew_grid(n);
#define ARGS_(TI_, TO_) (const TI_*)in, rs, (TO_*)out, batch, inner
  if (in_dtype == APB_F32 && out_dtype == APB_F32) scale_cast_kernel<float, float><<<g, EW_THREADS, 0, st>>>(ARGS_(float, float));
  else if (in_dtype == APB_BF16 && out_dtype == APB_F32) scale_cast_kernel<bf16, float><<<g, EW_THREADS, 0, st>>>(ARGS_(bf16, float));
  else if (in_dtype == APB_BF16 && out_dtype == APB_BF16) scale_cast_kernel<bf16, bf16><<<g, EW_THREADS, 0, st>>>(ARGS_(bf16, bf16));
  else if (in_dtype == APB_F32 && out_dtype == APB_BF16) scale_cast_kernel<float, bf16><<<g, EW_THREADS, 0, st>>>(ARGS_(float, bf16));
  else { apb_set_error("scale_cast: dtypes %d->%d", in_dtype, out_dtype); return APB_ERR_DTYPE; }
#undef ARGS_
  APB_LAUNCH_CHECK("scale_cast");
  return 0;
}

int apb_residual_add(const void* x, const void* r, const float* rs, void* out, long long batch, long long inner,
                     int x_dtype, int r_dtype, int out_dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  const long long n = batch * inner;
  if (n <= 0) return 0;
  const int g = ew_grid(n);
#define RA_(TX_, TR_, TO_) \
  residual_add_kernel<TX_, TR_, TO_><<<g, EW_THREADS, 0, st>>>((const TX_*)x, (const TR_*)r, rs, (TO_*)out, batch, inner)
  if (x_dtype == APB_F32 && r_dtype == APB_F32 && out_dtype == APB_F32) RA_(float, float, float);
  else if (x_dtype == APB_F32 && r_dtype == APB_BF16 && out_dtype == APB_F32) RA_(float, bf16, float);
  else if (x_dtype == APB_F32 && r_dtype == APB_BF16 && out_dtype == APB_BF16) RA_(float, bf16, bf16);
  else if (x_dtype == APB_BF16 && r_dtype == APB_BF16 && out_dtype == APB_BF16) RA_(bf16, bf16, bf16);
  else { apb_set_error("residual_add: dtypes %d,%d->%d", x_dtype, r_dtype, out_dtype); return APB_ERR_DTYPE; }
#undef RA_
  APB_LAUNCH_CHECK("residual_add");
  return 0;
}

int apb_cast(const void* in, void* out, long long n, int in_dtype, int out_dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (n <= 0) return 0;
  const int g = ew_grid(n);
  if (in_dtype == APB_F32 && out_dtype == APB_BF16) cast_kernel<float, bf16><<<g, EW_THREADS, 0, st>>>((const float*)in, (bf16*)out, n);
  else if (in_dtype == APB_BF16 && out_dtype == APB_F32) cast_kernel<bf16, float><<<g, EW_THREADS, 0, st>>>((const bf16*)in, (float*)out, n);
  else if (in_dtype == APB_F32 && out_dtype == APB_F32) cast_kernel<float, float><<<g, EW_THREADS, 0, st>>>((const float*)in, (float*)out, n);
  else if (in_dtype == APB_BF16 && out_dtype == APB_BF16) cast_kernel<bf16, bf16><<<g, EW_THREADS, 0, st>>>((const bf16*)in, (bf16*)out, n);
  else { apb_set_error("cast: dtypes %d->%d", in_dtype, out_dtype); return APB_ERR_DTYPE; }
  APB_LAUNCH_CHECK("cast");
  return 0;
}

int apb_add(const void* a, const void* b, void* out, long long n, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (n <= 0) return 0;
  DISPATCH_T(dtype, "add", (add_kernel<float><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)a, (const float*)b, (float*)out, n)),
             (add_kernel<bf16><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n)));
  return 0;
}

int apb_gelu_fwd(const void* x, void* y, long long n, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (n <= 0) return 0;
  DISPATCH_T(dtype, "gelu_fwd", (gelu_fwd_kernel<float><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)x, (float*)y, n)),
             (gelu_fwd_kernel<bf16><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)x, (bf16*)y, n)));
  return 0;
}

int apb_gelu_bwd(const void* x, const void* dy, void* dx, long long n, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (n <= 0) return 0;
  DISPATCH_T(dtype, "gelu_bwd",
             (gelu_bwd_kernel<float><<<ew_grid(n), EW_THREADS, 0, st>>>((const float*)x, (const float*)dy, (float*)dx, n)),
             (gelu_bwd_kernel<bf16><<<ew_grid(n), EW_THREADS, 0, st>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n)));
  return 0;
}

int apb_im2col(const void* x, void* col, int B, int C, int H, int W, int KH, int KW, int stride, int pad, int Kpad,
               long long sb, long long sc, long long sh, long long sw, int in_dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, APB_ERR_SHAPE, "im2col: bad geometry");
  APB_CHECK_ARG(Kpad % 8 == 0 && Kpad >= C * KH * KW, APB_ERR_SHAPE, "im2col: Kpad=%d must be a multiple of 8 >= %d", Kpad,
                C * KH * KW);
  APB_CHECK_ARG(((uintptr_t)col & 15) == 0, APB_ERR_ARG, "im2col: col must be 16-byte aligned");
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  APB_CHECK_ARG(OH > 0 && OW > 0, APB_ERR_SHAPE, "im2col: empty output");
  const int SW = W + 2 * pad, IR = (IM2COL_ROWS - 1) * stride + KH;
  const size_t smem = (size_t)((Kpad * 2 + 15) & ~15) + (size_t)C * IR * SW * sizeof(float);
  APB_CHECK_ARG(smem <= 200 * 1024 && (long long)C * IR * SW < 32768, APB_ERR_UNSUPPORTED,
                "im2col: staged tile of %d x %d floats does not fit (shared memory / 16-bit offsets)", C * IR, SW);
  const int chunks = Kpad / 8;
  APB_CHECK_ARG((long long)OW * chunks < 65536 && B <= 65535, APB_ERR_UNSUPPORTED, "im2col: row of %d x %d chunks too long", OW, chunks);
  const unsigned inv_chunks = (unsigned)(0x100000000ULL / (unsigned)chunks) + 1u;   // exact for e < 65536
  const dim3 grid(ceil_div(OH, IM2COL_ROWS), B);
  if (in_dtype == APB_F32) {
    cudaFuncSetAttribute(im2col_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    im2col_rows_kernel<float><<<grid, 256, smem, st>>>((const float*)x, (bf16*)col, C, H, W, KH, KW, stride, pad, OH, OW, Kpad, sb,
                                                      sc, sh, sw, inv_chunks);
  } else if (in_dtype == APB_BF16) {
    cudaFuncSetAttribute(im2col_rows_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    im2col_rows_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)x, (bf16*)col, C, H, W, KH, KW, stride, pad, OH, OW, Kpad, sb,
                                                     sc, sh, sw, inv_chunks);
  } else { apb_set_error("im2col: unsupported dtype %d", in_dtype); return APB_ERR_DTYPE; }
  APB_LAUNCH_CHECK("im2col");
  return 0;
}

int apb_bilinear_resize(const float* src, void* dst, long long planes, int H, int W, int OH, int OW, int out_dtype,
                        apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(planes > 0 && planes <= 65535LL * 1024 && H > 0 && W > 0 && OH > 0 && OW > 0, APB_ERR_SHAPE, "bilinear_resize: bad shape");
  const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
  int gx = ceil_div(OH * OW, EW_THREADS);
  if (gx > 64) gx = 64;
  for (long long p0 = 0; p0 < planes; p0 += 65535) {     // grid.y limit
    const int np = (int)((planes - p0 < 65535) ? planes - p0 : 65535);
    const float* s = src + (size_t)p0 * H * W;
    if (out_dtype == APB_F32)
      bilinear_resize_kernel<float><<<dim3(gx, np), EW_THREADS, 0, st>>>(s, (float*)dst + (size_t)p0 * OH * OW, H, W, OH, OW, sy, sx);
    else if (out_dtype == APB_BF16)
      bilinear_resize_kernel<bf16><<<dim3(gx, np), EW_THREADS, 0, st>>>(s, (bf16*)dst + (size_t)p0 * OH * OW, H, W, OH, OW, sy, sx);
    else { apb_set_error("bilinear_resize: unsupported out dtype %d", out_dtype); return APB_ERR_DTYPE; }
    APB_LAUNCH_CHECK("bilinear_resize");
  }
  return 0;
}
