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
//   stage 1: the 5-row pixel band of v (zero border) and the raw logits rows of the tile's windows (window rows r and r+1,
//            whole 16-byte vectors of the padded row) -> shared memory with cp.async (zero-fill outside the image): every
//            load of the CTA is in flight at once and no register is staged (the register-staged version spent 40 % of its
//            time on long-scoreboard stalls, profiles/r2_kernels.md)
//   stage 2: one thread per (window, head, P) row: its 9 logits (window row r: P rows 3..8; window row r+1: P rows 0..2)
//            from the raw rows -> softmax in fp32 -> weight row (pitch 12 floats)
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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {     // src_bytes 0 -> 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
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

// CH > 0: heads (and the padded logits pitch that goes with it) are compile-time constants -- the kernel is issue-bound and
// every index split by a runtime head count costs a ~20-instruction integer division (CH = 0: fully dynamic)
template <int CH>
__global__ void __launch_bounds__(FMA_THREADS) outlook_fwd_fma_kernel(OfParams p) {
  const int heads = CH > 0 ? CH : p.heads;
  const int lpitch = CH > 0 ? (CH * 81 + 7) / 8 * 8 : p.lpitch;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = heads * HD;
  const int xt = blockIdx.x % p.xtiles;
  const int r = (blockIdx.x / p.xtiles) % p.h;            // output rows 2r, 2r+1 ; window rows r, r+1
  const int b = blockIdx.x / (p.xtiles * p.h);
  const int jb = xt * p.tcw;                             // first window column / output block column of the tile
  const int nblk = min(p.tcw, p.w - jb);                 // output blocks in this tile
  const int nwin = nblk + 1;                             // window columns jb .. jb + nblk (the last one feeds only odd x)
  const int BW = 2 * nblk + 3;                           // band pixel columns: 2 jb - 1 .. 2 (jb + nblk) + 1
  // shared memory: band [5][BW][C] bf16 | weights [nwin][heads][9][WP] fp32 | raw logits [2][nwin][lpitch] bf16
  bf16* band = reinterpret_cast<bf16*>(smem_raw);
  const size_t band_bytes = ((size_t)5 * BW * C * sizeof(bf16) + 15) & ~(size_t)15;
  float* wts = reinterpret_cast<float*>(smem_raw + band_bytes);
  const size_t wts_bytes = (size_t)nwin * heads * 9 * WP * sizeof(float);
  bf16* raw = reinterpret_cast<bf16*>(smem_raw + band_bytes + wts_bytes);

  // ---- stage 1a: pixel band (rows 2r-1 .. 2r+3, columns 2jb-1 .. ), zeros outside the image.  No divisions in the loop:
  //      a thread walks (band column, 16-byte vector) with a precomputed stride
  {
    const int vpp = C / 8;                                // 16-byte vectors per pixel
    const int step_bc = FMA_THREADS / vpp, step_cv = FMA_THREADS % vpp;
    int bc = tid / vpp, cv = tid % vpp;
    const uint32_t band_s = smem_u32f(band);
    while (bc < BW) {
      const int xx = 2 * jb - 1 + bc;
      const bool xin = xx >= 0 && xx < p.W;
#pragma unroll
      for (int br = 0; br < 5; ++br) {
        const int yy = 2 * r - 1 + br;
        const bool ok = xin && yy >= 0 && yy < p.H;
        const bf16* src = ok ? p.v + (((size_t)b * p.H + yy) * p.W + xx) * C + cv * 8 : p.v;
        cp_async16(band_s + (uint32_t)(((br * BW + bc) * vpp + cv) * 16), src, ok ? 16 : 0);
      }
      bc += step_bc;
      cv += step_cv;
      if (cv >= vpp) { cv -= vpp; ++bc; }
    }
  }
  // ---- stage 1b: raw logits rows of windows (r, jb ..) and (r+1, jb ..): lpitch * 2 / 16 vectors per window
  {
    const int vpw = lpitch / 8;
    const int step_w = FMA_THREADS / vpw, step_v = FMA_THREADS % vpw;
    int wi = tid / vpw, vv = tid % vpw;                   // wi = slot * nwin + jl
    const uint32_t raw_s = smem_u32f(raw);
    while (wi < 2 * nwin) {
      const int slot = wi >= nwin ? 1 : 0, jl = wi - slot * nwin;
      const int iw = r + slot, jw = jb + jl;
      const bool ok = iw < p.h && jw < p.w;
      const bf16* src = ok ? p.logits + (((size_t)b * p.h + iw) * p.w + jw) * lpitch + vv * 8 : p.logits;
      cp_async16(raw_s + (uint32_t)((wi * vpw + vv) * 16), src, ok ? 16 : 0);
      wi += step_w;
      vv += step_v;
      if (vv >= vpw) { vv -= vpw; ++wi; }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- stage 2: thread = (window-head, weight row): 9 raw logits -> softmax -> fp32 weight row (rows of windows outside the
  //      grid become zeros).  Weight rows 0..5 = window row r, P = 3..8; rows 6..8 = window row r+1, P = 0..2
  {
    const int nwh = nwin * heads;
    for (int e = tid; e < nwh * 9; e += FMA_THREADS) {
      const int ridx = e % 9, wh = e / 9;
      const int jl = wh / heads, hd = wh - jl * heads;
      const int slot = ridx >= 6 ? 1 : 0, P = slot ? ridx - 6 : ridx + 3;
      const int iw = r + slot, jw = jb + jl;
      float* row = wts + (size_t)e * WP;
      float4 a, c4;
      float l8;
      if (iw < p.h && jw < p.w) {
        const bf16* src = raw + (size_t)(slot * nwin + jl) * lpitch + hd * 81 + P * 9;
        float q[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] = __bfloat162float(src[i]) * p.scale;
        const float m = fmaxf(fmaxf(fmaxf(q[0], q[1]), fmaxf(q[2], q[3])), fmaxf(fmaxf(q[4], q[5]), fmaxf(fmaxf(q[6], q[7]), q[8])));
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] = __expf(q[i] - m);
        const float inv = 1.f / (((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7])) + q[8]);
        a = make_float4(q[0] * inv, q[1] * inv, q[2] * inv, q[3] * inv);
        c4 = make_float4(q[4] * inv, q[5] * inv, q[6] * inv, q[7] * inv);
        l8 = q[8] * inv;
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
  const int npairs = (heads + 1) >> 1;
  const int items = nblk * npairs;
  const int half = lane >> 4, cp = lane & 15;
  const uint32_t band_a = smem_u32f(band), wts_a = smem_u32f(wts);
  const uint32_t row_pitch_b = (uint32_t)(BW * C * 2), col_pitch_b = (uint32_t)(C * 2);
  for (int item = warp; item < items; item += FMA_THREADS / 32) {
    const int hp = item % npairs, jl = item / npairs;
    const int hd = 2 * hp + half;
    const bool head_ok = hd < heads;
    const int hdc = head_ok ? hd : heads - 1;           // clamp: lanes of a missing head compute on valid memory, never store
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
    const uint32_t wbase = wts_a + (uint32_t)((((jl * heads + hdc) * 9) * WP) * 4);
    const uint32_t wnext = (uint32_t)(heads * 9 * WP * 4);          // next window column
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
  if (lpitch % 8 != 0 || lpitch < heads * 81 || ((uintptr_t)logits & 15) != 0) return APB_ERR_UNSUPPORTED;   // raw rows are staged as 16-byte vectors
  auto smem_for = [&](int tcw) {
    const size_t band = ((size_t)5 * (2 * tcw + 3) * C * 2 + 15) & ~(size_t)15;
    return band + (size_t)(tcw + 1) * heads * 9 * WP * 4 + (size_t)2 * (tcw + 1) * lpitch * 2;
  };
  // widest tile that leaves room for three CTAs per SM (18 warps); else two; else the widest that fits at all
  int tcw = p.w;
  const size_t budgets[3] = {75 * 1024, 113 * 1024, 227 * 1024};
  int bi = 0;
  for (; bi < 3; ++bi) {
    tcw = p.w;
    while (tcw > 1 && smem_for(tcw) > budgets[bi]) --tcw;
    if (smem_for(tcw) <= budgets[bi] && (tcw >= 4 || tcw == p.w || bi == 2)) break;
  }
  if (bi == 3 || smem_for(tcw) > 227 * 1024) return APB_ERR_UNSUPPORTED;
  // balance the column tiles
  p.xtiles = ceil_div(p.w, tcw);
  p.tcw = ceil_div(p.w, p.xtiles);
  const size_t smem = smem_for(p.tcw);
  const long long grid = (long long)B * p.h * p.xtiles;
  const int ch = (heads == 6 || heads == 8 || heads == 12) && lpitch == (heads * 81 + 7) / 8 * 8 ? heads : 0;
#define OL_LAUNCH(CH_)                                                                                                      \
  do {                                                                                                                      \
    static size_t attr = 0;                                                                                                 \
    if (smem > attr) {                                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(outlook_fwd_fma_kernel<CH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) { apb_set_error("outlook_fwd_fma_kernel: smem attr: %s", cudaGetErrorString(e)); return (int)e; }  \
      attr = smem;                                                                                                          \
    }                                                                                                                       \
    outlook_fwd_fma_kernel<CH_><<<(unsigned)grid, FMA_THREADS, smem, st>>>(p);                                                 \
  } while (0)
  if (ch == 6) OL_LAUNCH(6);
  else if (ch == 8) OL_LAUNCH(8);
  else if (ch == 12) OL_LAUNCH(12);
  else OL_LAUNCH(0);
#undef OL_LAUNCH
  APB_LAUNCH_CHECK("outlook_fwd_fma");
  return 0;
}
