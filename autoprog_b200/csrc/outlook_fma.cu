// OutlookAttention core FORWARD, gather formulation on the CUDA cores (bf16 I/O, fp32 softmax + accumulation).
//
//   reference: nn.Unfold(3,1,2) -> softmax(scale*logits) -> attn @ v -> F.fold          (models/volo.py:83-98)
//
// A (window, head) unit is a 9 x 9 x 32 product -- far below a tensor-core tile; the mma.sync formulation (outlook_mma.cu)
// pads it to 16 x 16 x 16 fragments and spends ~4x more instructions on fragment layout than on math (profiles/r1_kernels.md).
// This kernel instead computes every OUTPUT pixel directly (fold as a gather, no staging, no atomics):
//   Y[y, x, head, :] = sum over the 1 / 2 / 4 windows (i, j) covering (y, x), with P = position of (y, x) in the window,
//                      sum_Q softmax_Q(scale * logits[i, j, head, P, :])[Q] * V[pixel Q of window (i, j), head, :]
// CTA = two output rows (2r, 2r+1) x a range of columns x all heads.
//   stage 1: the 5-row pixel band of v (zero border) -> shared memory with 16-byte loads; the logits rows the tile needs
//            (window row r: P rows 3..8; window row r+1: P rows 0..2) -> shared memory as fp32, coalesced reads
//   stage 2: one thread per (window, head, P) row: softmax over its 9 logits in place (fp32, row pitch 12 floats)
//   stage 3: one warp per (2 x 2 output block, head pair): lane = (head of the pair, channel pair); the block's 5 x 5 pixel
//            patch of v is loaded ONCE into registers (25 x 32-bit shared loads per lane), the 9 weight rows that feed the
//            four output pixels arrive as broadcast 16-byte loads, 162 FFMA per lane, four 4-byte stores per lane
//            (64 contiguous bytes per head and pixel).
// Per (pixel, head): ~35 warp instructions instead of ~105 (418 per window unit / 4 pixels per unit).
#include "common.cuh"

namespace {

constexpr int HD = 32;
constexpr int WP = 12;                 // weight row pitch in floats (9 used)
constexpr int FMA_THREADS = 192;       // 6 warps: 14 blocks x 3 head pairs = 42 items = 7 per warp at 28 x 28 x 192

__device__ __forceinline__ uint32_t smem_u32f(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

struct OfParams {
  const bf16* v;
  const bf16* logits;
  bf16* y;
  int B, H, W, h, w, heads, lpitch;
  float scale;
  int tcw;          // output 2 x 2 blocks (= window columns) per CTA along x
  int xtiles;       // ceil(w / tcw)
};

__global__ void __launch_bounds__(FMA_THREADS) outlook_fwd_fma_kernel(OfParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.heads * HD;
  const int xt = blockIdx.x % p.xtiles;
  const int r = (blockIdx.x / p.xtiles) % p.h;            // output rows 2r, 2r+1 ; window rows r, r+1
  const int b = blockIdx.x / (p.xtiles * p.h);
  const int jb = xt * p.tcw;                             // first window column / output block column of the tile
  const int nblk = min(p.tcw, p.w - jb);                 // output blocks in this tile
  const int nwin = nblk + 1;                             // window columns jb .. jb + nblk (the last one feeds only odd x)
  const int BW = 2 * nblk + 3;                           // band pixel columns: 2 jb - 1 .. 2 (jb + nblk) + 1
  // shared memory: band [5][BW][C] bf16 | weights [nwin][heads][9][WP] fp32
  bf16* band = reinterpret_cast<bf16*>(smem_raw);
  const size_t band_bytes = ((size_t)5 * BW * C * sizeof(bf16) + 15) & ~(size_t)15;
  float* wts = reinterpret_cast<float*>(smem_raw + band_bytes);

  // ---- stage 1a: pixel band (rows 2r-1 .. 2r+3, columns 2jb-1 .. ), zeros outside the image.  No divisions in the loops:
  //      a thread walks (band column, 16-byte vector) with a precomputed stride; all five rows of a position are loaded
  //      before they are stored (five independent 16-byte loads in flight per thread)
  {
    const int vpp = C / 8;                                // 16-byte vectors per pixel
    const int step_bc = FMA_THREADS / vpp, step_cv = FMA_THREADS % vpp;
    int bc = tid / vpp, cv = tid % vpp;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4* band4 = reinterpret_cast<uint4*>(band);
    while (bc < BW) {
      const int xx = 2 * jb - 1 + bc;
      const bool xin = xx >= 0 && xx < p.W;
      uint4 val[5];
#pragma unroll
      for (int br = 0; br < 5; ++br) {
        const int yy = 2 * r - 1 + br;
        val[br] = z;
        if (xin && yy >= 0 && yy < p.H) val[br] = __ldg(reinterpret_cast<const uint4*>(p.v + (((size_t)b * p.H + yy) * p.W + xx) * C) + cv);
      }
#pragma unroll
      for (int br = 0; br < 5; ++br) band4[(br * BW + bc) * vpp + cv] = val[br];
      bc += step_bc;
      cv += step_cv;
      if (cv >= vpp) { cv -= vpp; ++bc; }
    }
  }
  // ---- stage 1b: raw logits -> fp32.  Per (window column, head): slot 0 = window row r, elements 27..80 (P = 3..8) ->
  //      weight rows 0..5; slot 1 = window row r+1, elements 0..26 (P = 0..2) -> weight rows 6..8.  One warp per (window
  //      column, head): lane l takes elements l, l+32, l+64 of the 81 (constant divisors, three independent loads)
  {
    const int nwh = nwin * p.heads;
    int hd = warp % p.heads, jl = warp / p.heads;
    const int dh = (FMA_THREADS / 32) % p.heads, dj = (FMA_THREADS / 32) / p.heads;
    for (int wh = warp; wh < nwh; wh += FMA_THREADS / 32) {
      const int jw = jb + jl;
      float val[3];
#pragma unroll
      for (int t3 = 0; t3 < 3; ++t3) {
        const int k = lane + 32 * t3;
        val[t3] = 0.f;
        if (k < 81) {
          const int slot = k >= 54 ? 1 : 0;
          const int el = slot ? k - 54 : k + 27;          // element inside the head's 81 logits
          const int iw = r + slot;
          if (iw < p.h && jw < p.w)
            val[t3] = __bfloat162float(p.logits[(((size_t)b * p.h + iw) * p.w + jw) * p.lpitch + hd * 81 + el]) * p.scale;
        }
      }
#pragma unroll
      for (int t3 = 0; t3 < 3; ++t3) {
        const int k = lane + 32 * t3;
        if (k < 81) wts[((size_t)wh * 9 + k / 9) * WP + k % 9] = val[t3];     // k / 9 = weight row: 0..5 (slot 0), 6..8 (slot 1)
      }
      hd += dh; jl += dj;
      if (hd >= p.heads) { hd -= p.heads; ++jl; }
    }
  }
  __syncthreads();
  // ---- stage 2: row softmax in place (rows of windows outside the grid become zeros); thread = (window-head, row)
  {
    const int nwh = nwin * p.heads;
    for (int e = tid; e < nwh * 9; e += FMA_THREADS) {
      const int ridx = e % 9, wh = e / 9;
      const int jl = wh / p.heads;
      const int iw = r + (ridx >= 6 ? 1 : 0), jw = jb + jl;
      float* row = wts + (size_t)e * WP;
      float4 a = *reinterpret_cast<float4*>(row), c4 = *reinterpret_cast<float4*>(row + 4);
      float l8 = row[8];
      if (iw < p.h && jw < p.w) {
        float m = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(c4.x, c4.y), fmaxf(fmaxf(c4.z, c4.w), l8)));
        a.x = __expf(a.x - m); a.y = __expf(a.y - m); a.z = __expf(a.z - m); a.w = __expf(a.w - m);
        c4.x = __expf(c4.x - m); c4.y = __expf(c4.y - m); c4.z = __expf(c4.z - m); c4.w = __expf(c4.w - m);
        l8 = __expf(l8 - m);
        const float inv = 1.f / (((a.x + a.y) + (a.z + a.w)) + ((c4.x + c4.y) + (c4.z + c4.w)) + l8);
        a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv; c4.x *= inv; c4.y *= inv; c4.z *= inv; c4.w *= inv; l8 *= inv;
      } else {
        a = make_float4(0.f, 0.f, 0.f, 0.f); c4 = a; l8 = 0.f;
      }
      *reinterpret_cast<float4*>(row) = a;
      *reinterpret_cast<float4*>(row + 4) = c4;
      row[8] = l8;
    }
  }
  __syncthreads();
  // ---- stage 3: one warp per (output block, head pair)
  const int npairs = (p.heads + 1) >> 1;
  const int items = nblk * npairs;
  const int half = lane >> 4, cp = lane & 15;
  const uint32_t band_a = smem_u32f(band), wts_a = smem_u32f(wts);
  const uint32_t row_pitch_b = (uint32_t)(BW * C * 2), col_pitch_b = (uint32_t)(C * 2);
  for (int item = warp; item < items; item += FMA_THREADS / 32) {
    const int hp = item % npairs, jl = item / npairs;
    const int hd = 2 * hp + half;
    const bool head_ok = hd < p.heads;
    const int hdc = head_ok ? hd : p.heads - 1;           // clamp: lanes of a missing head compute on valid memory, never store
    // 5 x 5 pixel patch at band rows 0..4, band columns 2 jl .. 2 jl + 4 ; this lane's channel pair
    float v0[25], v1[25];
    {
      const uint32_t base = band_a + (uint32_t)(((2 * jl) * C + hdc * HD + 2 * cp) * 2);
#pragma unroll
      for (int pr = 0; pr < 5; ++pr) {
        const uint32_t rb = base + pr * row_pitch_b;
#pragma unroll
        for (int pc = 0; pc < 5; ++pc) {
          const uint32_t u = lds32(rb + pc * col_pitch_b);
          v0[pr * 5 + pc] = __uint_as_float(u << 16);
          v1[pr * 5 + pc] = __uint_as_float(u & 0xFFFF0000u);
        }
      }
    }
    float acc[4][2];
#pragma unroll
    for (int o = 0; o < 4; ++o) { acc[o][0] = 0.f; acc[o][1] = 0.f; }
    // weight rows: (window column offset dj, row index, patch origin (oy, ox), output pixel o = dy * 2 + dx)
    //   window (r, j)    : P=(1,1)->o0 (ridx 1), (1,2)->o1 (ridx 2), (2,1)->o2 (ridx 4), (2,2)->o3 (ridx 5); origin (0,0)
    //   window (r, j+1)  : P=(1,0)->o1 (ridx 0), (2,0)->o3 (ridx 3);                                        origin (0,2)
    //   window (r+1, j)  : P=(0,1)->o2 (ridx 7), (0,2)->o3 (ridx 8);                                        origin (2,0)
    //   window (r+1, j+1): P=(0,0)->o3 (ridx 6);                                                            origin (2,2)
    const uint32_t wbase = wts_a + (uint32_t)((((jl * p.heads + hdc) * 9) * WP) * 4);
    const uint32_t wnext = (uint32_t)(p.heads * 9 * WP * 4);          // next window column
    auto apply = [&](int dj, int ridx, int oy, int ox, int o) {
      const uint32_t ra = wbase + dj * wnext + (uint32_t)(ridx * WP * 4);
      const float4 wa = lds128f(ra), wb = lds128f(ra + 16);
      const float w8 = __uint_as_float(lds32(ra + 32));
      const float wq[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, w8};
#pragma unroll
      for (int qi = 0; qi < 3; ++qi)
#pragma unroll
        for (int qj = 0; qj < 3; ++qj) {
          const int pidx = (oy + qi) * 5 + ox + qj;
          acc[o][0] = fmaf(wq[qi * 3 + qj], v0[pidx], acc[o][0]);
          acc[o][1] = fmaf(wq[qi * 3 + qj], v1[pidx], acc[o][1]);
        }
    };
    apply(0, 1, 0, 0, 0);
    apply(0, 2, 0, 0, 1);
    apply(0, 4, 0, 0, 2);
    apply(0, 5, 0, 0, 3);
    apply(1, 0, 0, 2, 1);
    apply(1, 3, 0, 2, 3);
    apply(0, 7, 2, 0, 2);
    apply(0, 8, 2, 0, 3);
    apply(1, 6, 2, 2, 3);
    if (head_ok) {
      const int y0 = 2 * r, x0 = 2 * (jb + jl);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int yy = y0 + (o >> 1), xx = x0 + (o & 1);
        if (yy < p.H && xx < p.W) {
          __nv_bfloat162 hv = __floats2bfloat162_rn(acc[o][0], acc[o][1]);
          *reinterpret_cast<__nv_bfloat162*>(p.y + (((size_t)b * p.H + yy) * p.W + xx) * C + hd * HD + 2 * cp) = hv;
        }
      }
    }
  }
}

}  // namespace

// returns APB_ERR_UNSUPPORTED when even a one-block-wide tile does not fit shared memory
int apb_outlook_fwd_fma(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int lpitch,
                        apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  OfParams p;
  p.v = (const bf16*)v; p.logits = (const bf16*)logits; p.y = (bf16*)y;
  p.B = B; p.H = H; p.W = W; p.h = (H + 1) / 2; p.w = (W + 1) / 2; p.heads = heads; p.lpitch = lpitch; p.scale = scale;
  const int C = heads * HD;
  auto smem_for = [&](int tcw) {
    const size_t band = ((size_t)5 * (2 * tcw + 3) * C * 2 + 15) & ~(size_t)15;
    return band + (size_t)(tcw + 1) * heads * 9 * WP * 4;
  };
  // widest tile that leaves room for two CTAs per SM; else the widest that fits at all
  int tcw = p.w;
  while (tcw > 1 && smem_for(tcw) > 113 * 1024) --tcw;
  if (smem_for(tcw) > 113 * 1024) {
    tcw = p.w;
    while (tcw > 1 && smem_for(tcw) > 227 * 1024) --tcw;
    if (smem_for(tcw) > 227 * 1024) return APB_ERR_UNSUPPORTED;
  }
  // balance the column tiles
  p.xtiles = ceil_div(p.w, tcw);
  p.tcw = ceil_div(p.w, p.xtiles);
  const size_t smem = smem_for(p.tcw);
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(outlook_fwd_fma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { apb_set_error("outlook_fwd_fma: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr = smem;
  }
  const long long grid = (long long)B * p.h * p.xtiles;
  outlook_fwd_fma_kernel<<<(unsigned)grid, FMA_THREADS, smem, st>>>(p);
  APB_LAUNCH_CHECK("outlook_fwd_fma");
  return 0;
}
