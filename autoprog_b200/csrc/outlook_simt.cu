// OutlookAttention core, reference-precision SIMT path (fp32 parity mode; also valid for bf16 I/O).
//
// Replaces, for one Outlooker layer, the reference chain models/volo.py:83-98
//   nn.Unfold(3, pad 1, stride 2) -> softmax(scale * logits) over the last 9 -> attn @ v -> F.fold
// with one gather kernel per direction: every output pixel sums the 1/2/4 window rows that cover it
// (SURVEY.md A.2), so there is no atomic scatter-add and the result is deterministic.
//
// Layouts (all contiguous):
//   v, y, dy, dv : [B, H, W, heads*32]            (NHWC, channel = head*32 + c)
//   logits, dlog : [B, h, w, lpitch]              (h=ceil(H/2), w=ceil(W/2); channel = head*81 + P*9 + Q;
//                                                  lpitch >= heads*81, the tail columns are padding: read never,
//                                                  written as zeros by the backward)
// Work unit: a "quad" = the 2x2 output pixels (2i..2i+1, 2j..2j+1) of one head, owned by one warp,
// lane = channel.  The quad needs the 5x5 pixel patch around it and the <=4 windows (i+a, j+b).
#include "common.cuh"

namespace {

constexpr int HD = 32;  // head dim, fixed by every VOLO variant (models/volo.py:697-821: dim/heads == 32)

// softmax of the 9x9 logits of up to 4 windows into per-warp shared memory sA[4][81]
template <typename T>
__device__ __forceinline__ void quad_softmax(const T* __restrict__ logits, float* sA, int b, int i, int j, int head,
                                             int h, int w, int lpitch, float scale, int lane) {
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const int wi = i + a, wj = j + bb;
      float* dst = sA + (a * 2 + bb) * 81;
      if (wi < h && wj < w) {
        if (lane < 9) {
          const T* src = logits + (((size_t)b * h + wi) * w + wj) * lpitch + head * 81 + lane * 9;
          float e[9], m = -INFINITY;
#pragma unroll
          for (int q = 0; q < 9; ++q) { e[q] = to_f(src[q]) * scale; m = fmaxf(m, e[q]); }
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < 9; ++q) { e[q] = expf(e[q] - m); s += e[q]; }
          const float inv = 1.f / s;
#pragma unroll
          for (int q = 0; q < 9; ++q) dst[lane * 9 + q] = e[q] * inv;
        }
      } else {
        for (int t = lane; t < 81; t += 32) dst[t] = 0.f;
      }
    }
  __syncwarp();
}

template <typename T>
__device__ __forceinline__ void load_patch(const T* __restrict__ src, float (&p)[5][5], int b, int i, int j, int head,
                                           int H, int W, int C, int lane) {
#pragma unroll
  for (int u = 0; u < 5; ++u)
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int y = 2 * i - 1 + u, x = 2 * j - 1 + t;
      p[u][t] = (y >= 0 && y < H && x >= 0 && x < W)
                    ? to_f(src[(((size_t)b * H + y) * W + x) * C + head * HD + lane])
                    : 0.f;
    }
}

// TRANS=false: out[pixel] = sum_{win,P->pixel} sum_Q A[P][Q] * src[pixel of Q]      (forward, src = v)
// TRANS=true : out[pixel] = sum_{win,Q->pixel} sum_P A[P][Q] * src[pixel of P]      (dV, src = dy)
template <typename T, bool TRANS>
__global__ void __launch_bounds__(128) outlook_gather_kernel(const T* __restrict__ src, const T* __restrict__ logits,
                                                             T* __restrict__ out, int B, int H, int W, int heads,
                                                             int h, int w, int lpitch, float scale) {
  __shared__ float sA_all[4][4 * 81];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sA = sA_all[warp];
  const int C = heads * HD;
  const long long nquads = (long long)B * h * w * heads;
  for (long long qd = (long long)blockIdx.x * 4 + warp; qd < nquads; qd += (long long)gridDim.x * 4) {
    const int head = (int)(qd % heads);
    long long r = qd / heads;
    const int j = (int)(r % w); r /= w;
    const int i = (int)(r % h);
    const int b = (int)(r / h);
    __syncwarp();
    quad_softmax<T>(logits, sA, b, i, j, head, h, w, lpitch, scale, lane);
    float p[5][5];
    load_patch<T>(src, p, b, i, j, head, H, W, C, lane);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a <= dy; ++a)
#pragma unroll
          for (int bb = 0; bb <= dx; ++bb) {
            const int fixed = (1 + dy - 2 * a) * 3 + (1 + dx - 2 * bb);  // P (fwd) or Q (dV) of this pixel in the window
            const float* Aw = sA + (a * 2 + bb) * 81;
#pragma unroll
            for (int qi = 0; qi < 3; ++qi)
#pragma unroll
              for (int qj = 0; qj < 3; ++qj) {
                const int run = qi * 3 + qj;
                const float coef = TRANS ? Aw[run * 9 + fixed] : Aw[fixed * 9 + run];
                acc = fmaf(coef, p[2 * a + qi][2 * bb + qj], acc);
              }
          }
        const int y = 2 * i + dy, x = 2 * j + dx;
        if (y < H && x < W) out[(((size_t)b * H + y) * W + x) * C + head * HD + lane] = from_f<T>(acc);
      }
  }
}

// dlogits[P][Q] = scale * A[P][Q] * (dA[P][Q] - sum_Q' A[P][Q'] dA[P][Q']),  dA[P][Q] = <dy[pix P], v[pix Q]>
template <typename T>
__global__ void __launch_bounds__(128) outlook_dlogits_kernel(const T* __restrict__ v, const T* __restrict__ logits,
                                                              const T* __restrict__ dy, T* __restrict__ dlogits,
                                                              int B, int H, int W, int heads, int h, int w, int lpitch,
                                                              float scale) {
  __shared__ float sD_all[4][81];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sD = sD_all[warp];
  const int C = heads * HD;
  const long long nwin = (long long)B * h * w * heads;
  for (long long wd = (long long)blockIdx.x * 4 + warp; wd < nwin; wd += (long long)gridDim.x * 4) {
    const int head = (int)(wd % heads);
    long long r = wd / heads;
    const int j = (int)(r % w); r /= w;
    const int i = (int)(r % h);
    const int b = (int)(r / h);
    float vv[9], gg[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const int y = 2 * i - 1 + q / 3, x = 2 * j - 1 + q % 3;
      const bool ok = (y >= 0 && y < H && x >= 0 && x < W);
      const size_t off = (((size_t)b * H + y) * W + x) * C + head * HD + lane;
      vv[q] = ok ? to_f(v[off]) : 0.f;
      gg[q] = ok ? to_f(dy[off]) : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int P = 0; P < 9; ++P)
#pragma unroll
      for (int Q = 0; Q < 9; ++Q) {
        const float s = warp_sum(gg[P] * vv[Q]);
        if (lane == 0) sD[P * 9 + Q] = s;
      }
    __syncwarp();
    const size_t win = (((size_t)b * h + i) * w + j) * lpitch;
    if (lane < 9) {
      const size_t base = win + head * 81 + lane * 9;
      float e[9], m = -INFINITY;
#pragma unroll
      for (int q = 0; q < 9; ++q) { e[q] = to_f(logits[base + q]) * scale; m = fmaxf(m, e[q]); }
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 9; ++q) { e[q] = expf(e[q] - m); s += e[q]; }
      const float inv = 1.f / s;
      float dot = 0.f;
#pragma unroll
      for (int q = 0; q < 9; ++q) { e[q] *= inv; dot = fmaf(e[q], sD[lane * 9 + q], dot); }
#pragma unroll
      for (int q = 0; q < 9; ++q) dlogits[base + q] = from_f<T>(scale * e[q] * (sD[lane * 9 + q] - dot));
    } else if (head == 0 && lane - 9 < lpitch - heads * 81) {
      dlogits[win + heads * 81 + (lane - 9)] = from_f<T>(0.f);   // padding columns (lpitch - heads*81 <= 7)
    }
  }
}

int grid_for(long long units) {
  long long g = (units + 3) / 4;
  const long long cap = 148LL * 16 * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

template <typename T>
static int outlook_fwd_simt_t(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale,
                              int lpitch, cudaStream_t st) {
  const int h = (H + 1) / 2, w = (W + 1) / 2;
  outlook_gather_kernel<T, false><<<grid_for((long long)B * h * w * heads), 128, 0, st>>>(
      (const T*)v, (const T*)logits, (T*)y, B, H, W, heads, h, w, lpitch, scale);
  APB_LAUNCH_CHECK("outlook_fwd_simt");
  return 0;
}

template <typename T>
static int outlook_bwd_simt_t(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H,
                              int W, int heads, float scale, int lpitch, cudaStream_t st) {
  const int h = (H + 1) / 2, w = (W + 1) / 2;
  const int g = grid_for((long long)B * h * w * heads);
  outlook_gather_kernel<T, true><<<g, 128, 0, st>>>((const T*)dy, (const T*)logits, (T*)dv, B, H, W, heads, h, w, lpitch, scale);
  APB_LAUNCH_CHECK("outlook_dv_simt");
  outlook_dlogits_kernel<T><<<g, 128, 0, st>>>((const T*)v, (const T*)logits, (const T*)dy, (T*)dlogits, B, H, W,
                                                heads, h, w, lpitch, scale);
  APB_LAUNCH_CHECK("outlook_dlogits_simt");
  return 0;
}

#define OUTLOOK_ARG_CHECK(name)                                                                                   \
  APB_CHECK_ARG(B > 0 && H > 0 && W > 0 && heads > 0, APB_ERR_SHAPE, name ": bad shape");                         \
  APB_CHECK_ARG(dtype == APB_F32 || dtype == APB_BF16, APB_ERR_DTYPE, name ": dtype %d", dtype);                  \
  APB_CHECK_ARG(lpitch >= heads * 81 && lpitch < heads * 81 + 8, APB_ERR_ARG, name ": logits pitch %d for %d heads", lpitch, heads)

int apb_outlook_fwd_simt(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale,
                         int lpitch, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  OUTLOOK_ARG_CHECK("outlook_fwd");
  if (dtype == APB_F32) return outlook_fwd_simt_t<float>(v, logits, y, B, H, W, heads, scale, lpitch, st);
  return outlook_fwd_simt_t<bf16>(v, logits, y, B, H, W, heads, scale, lpitch, st);
}

int apb_outlook_bwd_simt(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                         int heads, float scale, int lpitch, int dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  OUTLOOK_ARG_CHECK("outlook_bwd");
  if (dtype == APB_F32) return outlook_bwd_simt_t<float>(v, logits, dy, dv, dlogits, B, H, W, heads, scale, lpitch, st);
  return outlook_bwd_simt_t<bf16>(v, logits, dy, dv, dlogits, B, H, W, heads, scale, lpitch, st);
}
