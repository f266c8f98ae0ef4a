// Token-label TARGET builder: top-5 label maps + crop record -> soft token-label target [B, C, 2 + L*L] (class-major,
// the layout TokenLabelCrossEntropy consumes, loss/cross_entropy.py:146-148).
//
// Replaces `tlt.data.create_token_label_target` at the reference's call sites main_prog.py:983-1004 / 1919-1932 (tlt 0.1.0
// is a third-party package that is not vendored under /root/reference: the recipe below is restated from the published
// TokenLabeling code, see SURVEY.md Appendix B and oracle/token_label_cpu.py -- parity UNPINNED upstream):
//   dense[C, Hm, Wm] = scatter(top-5 scores by class id)  ->  RoIAlign(crop box * (Wm, Hm) - 0.5) to L x L (adaptive
//   sampling, aligned = false) [+ horizontal flip]  ->  softmax over classes  ->  * on + off;   slot 1 = the same with a
//   1 x 1 output, slot 0 = smoothed one-hot of the ground-truth class.
// The dense [B, C, Hm, Wm] map (C = 1000) is never materialised: a cell's RoIAlign touches at most grid^2 * 4 * 5
// (class, weight) pairs.  CTA = (image, 32 consecutive output slots): a warp owns a slot, accumulates its cell into a
// shared-memory row of C floats (untouched classes stay 0 = the value the dense map holds there), takes the softmax over
// the row in place (one exp per element) and applies the smoothing; the [32 slots][C] tile is then written out
// class-major, 32 contiguous floats per class row (odd row pitch in shared memory: conflict-free transposed reads).
#include "common.cuh"

namespace {

constexpr int TL_THREADS = 512;
constexpr int TL_WARPS = TL_THREADS / 32;
constexpr int TOPK = 5;
constexpr int TL_SLOTS = 32;   // output slots per CTA = floats per contiguous store run

struct TlParams {
  const float* maps;   // [B, 3, 5, Hm, Wm]
  float* out;          // [B, C, 2 + L*L]
  int B, C, Hm, Wm, L;
  float on, off;
  int softmax;
};

// accumulate the RoIAlign of output cell (ph, pw) of an (Lh x Lw)-cell pooling into vals[C]
__device__ __forceinline__ void accumulate_cell(const float* __restrict__ sc, const float* __restrict__ id, int Hm, int Wm, float x1,
                                                float y1, float x2, float y2, int Lh, int Lw, int ph, int pw, float* vals, int C,
                                                int lane) {
  const float roi_w = fmaxf(x2 - x1, 1.f), roi_h = fmaxf(y2 - y1, 1.f);
  const float bin_h = roi_h / (float)Lh, bin_w = roi_w / (float)Lw;
  const int gh = (int)ceilf(roi_h / (float)Lh), gw = (int)ceilf(roi_w / (float)Lw);
  const float inv_count = 1.f / (float)max(gh * gw, 1);
  const int total = gh * gw * 4 * TOPK;
  const int plane = Hm * Wm;
  for (int e = lane; e < total; e += 32) {
    const int k = e % TOPK, nb = (e / TOPK) & 3, pt = e / (TOPK * 4);
    const int iy = pt / gw, ix = pt % gw;
    float y = y1 + ph * bin_h + (iy + 0.5f) * bin_h / (float)gh;
    float x = x1 + pw * bin_w + (ix + 0.5f) * bin_w / (float)gw;
    if (y < -1.f || y > (float)Hm || x < -1.f || x > (float)Wm) continue;
    y = fmaxf(y, 0.f);
    x = fmaxf(x, 0.f);
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= Hm - 1) { yh = yl = Hm - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= Wm - 1) { xh = xl = Wm - 1; x = (float)xl; } else xh = xl + 1;
    const float ly = y - (float)yl, lx = x - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
    const int yy = (nb & 2) ? yh : yl, xx = (nb & 1) ? xh : xl;
    const float w = ((nb & 2) ? ly : hy) * ((nb & 1) ? lx : hx) * inv_count;
    const int pix = yy * Wm + xx;
    const int c = (int)id[k * plane + pix];
    if (c < 0 || c >= C) continue;
    atomicAdd(&vals[c], w * sc[k * plane + pix]);
  }
}

__global__ void __launch_bounds__(TL_THREADS) token_label_target_kernel(TlParams p) {
  extern __shared__ float tile[];                                         // [TL_SLOTS][pitch] | scores + ids [2][5][plane]
  const int C = p.C, N = p.L * p.L, row = 2 + N, pitch = C | 1;
  const int b = blockIdx.y, j0 = blockIdx.x * TL_SLOTS, nslots = min(TL_SLOTS, row - j0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane = p.Hm * p.Wm;
  const float* gsc = p.maps + (size_t)b * 3 * TOPK * plane;
  const float* rec = gsc + (size_t)2 * TOPK * plane;
  // the image's top-5 score and id planes -> shared memory once (the RoIAlign taps are data-dependent gathers)
  float* sc = tile + TL_SLOTS * pitch;
  const float* id = sc + TOPK * plane;
  for (int i = tid; i < 2 * TOPK * plane; i += TL_THREADS) sc[i] = gsc[i];
  const float x1 = rec[0] * p.Wm - 0.5f, y1 = rec[1] * p.Hm - 0.5f, x2 = rec[2] * p.Wm - 0.5f, y2 = rec[3] * p.Hm - 0.5f;
  const bool flip = rec[4] > 0.5f;
  const int gt = (int)rec[5];

  __syncthreads();
  // slot 0: smoothed one-hot of the ground truth; slot 1: the 1 x 1 class-level pooling; slot 2 + n: token n = (ph, pw) of
  // the L x L pooling (mirrored when flipped)
  for (int jj = warp; jj < nslots; jj += TL_WARPS) {
    const int j = j0 + jj;
    float* vals = tile + jj * pitch;
    if (j == 0) {
      for (int c = lane; c < C; c += 32) vals[c] = (c == gt) ? p.on : p.off;
      continue;
    }
    for (int c = lane; c < C; c += 32) vals[c] = 0.f;
    __syncwarp();
    if (j == 1) {
      accumulate_cell(sc, id, p.Hm, p.Wm, x1, y1, x2, y2, 1, 1, 0, 0, vals, C, lane);
    } else {
      const int n = j - 2, ph = n / p.L, pw_out = n % p.L, pw = flip ? p.L - 1 - pw_out : pw_out;
      accumulate_cell(sc, id, p.Hm, p.Wm, x1, y1, x2, y2, p.L, p.L, ph, pw, vals, C, lane);
    }
    __syncwarp();
    if (p.softmax) {
      float m = 0.f;                                   // classes the RoI never touched sit at 0
      for (int c = lane; c < C; c += 32) m = fmaxf(m, vals[c]);
      m = warp_max(m);
      float z = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float e = __expf(vals[c] - m);
        vals[c] = e;
        z += e;
      }
      const float scale = p.on / warp_sum(z);
      for (int c = lane; c < C; c += 32) vals[c] = fmaf(vals[c], scale, p.off);
    } else {
      for (int c = lane; c < C; c += 32) vals[c] = fmaf(vals[c], p.on, p.off);
    }
  }
  __syncthreads();
  float* outb = p.out + (size_t)b * C * row + j0;
  if (lane < nslots)
    for (int c = warp; c < C; c += TL_WARPS) outb[(size_t)c * row + lane] = tile[lane * pitch + c];
}

__global__ void onehot_smooth_kernel(const long long* __restrict__ labels, float* __restrict__ out, int B, int C, float on, float off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C) return;
  const int b = (int)(i / C), c = (int)(i % C);
  out[i] = (labels[b] == c) ? on : off;
}

}  // namespace

// maps: fp32 [B, 3, 5, Hm, Wm] (device); out: fp32 [B, C, 2 + L*L]
int apb_token_label_target(const float* maps, float* out, int B, int C, int Hm, int Wm, int label_size, float smoothing,
                           int apply_softmax, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && C > 0 && Hm > 0 && Wm > 0 && label_size > 0, APB_ERR_SHAPE, "token_label_target: B=%d C=%d Hm=%d Wm=%d L=%d", B,
                C, Hm, Wm, label_size);
  TlParams p;
  p.maps = maps; p.out = out; p.B = B; p.C = C; p.Hm = Hm; p.Wm = Wm; p.L = label_size;
  p.off = smoothing / (float)C;
  p.on = 1.f - smoothing + p.off;
  p.softmax = apply_softmax;
  const int row = 2 + label_size * label_size;
  const size_t smem = (size_t)TL_SLOTS * (C | 1) * 4 + (size_t)2 * 5 * Hm * Wm * 4;
  APB_CHECK_ARG(smem <= 227 * 1024, APB_ERR_UNSUPPORTED, "token_label_target: C=%d needs %zu B of shared memory", C, smem);
  cudaError_t e = cudaFuncSetAttribute(token_label_target_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { apb_set_error("token_label_target: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  token_label_target_kernel<<<dim3(ceil_div(row, TL_SLOTS), B), TL_THREADS, smem, st>>>(p);
  APB_LAUNCH_CHECK("token_label_target");
  return 0;
}

// labels: int64 [B] (device) -> smoothed one-hot fp32 [B, C]
int apb_onehot_smooth(const long long* labels, float* out, int B, int C, float smoothing, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(B > 0 && C > 0, APB_ERR_SHAPE, "onehot_smooth: B=%d C=%d", B, C);
  const float off = smoothing / (float)C, on = 1.f - smoothing + off;
  const long long n = (long long)B * C;
  onehot_smooth_kernel<<<ceil_div(n, 256), 256, 0, st>>>(labels, out, B, C, on, off);
  APB_LAUNCH_CHECK("onehot_smooth");
  return 0;
}
