// OutlookAttention core on tensor cores (bf16 I/O, fp32 softmax + accumulation), one fused kernel per direction.
//
//   reference: nn.Unfold(3,1,2) -> softmax(scale*logits) -> attn @ v -> F.fold          (models/volo.py:83-98)
//
// A CTA owns a band of TR window rows of one image (all window columns when they fit, all heads).  It stages the
// (2TR+3)-row pixel band of v (and dy in the backward) ONCE in shared memory with coalesced 16-byte loads; every
// window/head unit then is a 16x16x16 mma.sync problem whose operands come straight from that tile:
//   forward : out[P][c]  = sum_Q A[P][Q] v[pix(Q)][c]        A = softmax fragment built in registers (warp shuffles)
//   backward: dA[P][Q]   = <dy[pix(P)], v[pix(Q)]>           ; dlogits = scale * A o (dA - rowsum(A o dA))
//             dvw[Q][c]  = sum_P A[P][Q] dy[pix(P)][c]       (A^T through movmatrix)
// ldmatrix row addresses do the 3x3 unfold for free (out-of-image pixels are zero rows of the tile).  The fold is a
// deterministic gather: per head, unit results are staged in shared memory and every output pixel sums the 1/2/4
// window rows that cover it -- no atomics.  HBM traffic is exactly one read of v / logits (/ dy) and one write of
// y (/ dv, dlogits); the halo rows re-read by neighbouring bands hit L2.
#include "common.cuh"

namespace {

constexpr int HD = 32;
constexpr int OS = 40;          // fp32 staging row pitch (32 channels + 8 pad: conflict-free 64-bit stores)
constexpr int NTHREADS = 512;

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
__device__ __forceinline__ uint32_t movmatrix_t(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct Geo {
  int B, H, W, heads, h, w, lpitch;
  int TR;              // window rows per CTA
  int nWR;             // window rows computed per CTA (TR + 1 halo)
  int PR, PC, Cp;      // staged pixel rows / cols, padded channel pitch (bf16 elements)
  float scale;
};

// stage pixel rows [y0, y0+PR) x cols [-1, -1+PC) of src[b] (NHWC bf16) into s[PR][PC][Cp]; outside the image -> zeros
__device__ __forceinline__ void stage_band(bf16* s, const bf16* __restrict__ src, const Geo& g, int b, int y0) {
  const int C = g.heads * HD;
  const int cpr = C / 8;                                   // 16-byte chunks per pixel
  const int total = g.PR * g.PC * cpr;
  for (int e = threadIdx.x; e < total; e += NTHREADS) {
    const int ch = e % cpr;
    const int px = (e / cpr) % g.PC;
    const int py = e / (cpr * g.PC);
    const int y = y0 + py, x = px - 1;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < g.H && x >= 0 && x < g.W)
      val = *reinterpret_cast<const uint4*>(src + (((size_t)b * g.H + y) * g.W + x) * C + ch * 8);
    *reinterpret_cast<uint4*>(s + ((size_t)py * g.PC + px) * g.Cp + ch * 8) = val;
  }
}

// raw logits of one unit in fragment order: (P=g: Q=2q, 2q+1, 8) and, for g == 0, (P=8: Q=2q, 2q+1, 8)
struct RawLogits { float e0, e1, e2, f0, f1, f2; };
__device__ __forceinline__ RawLogits load_logits(const bf16* __restrict__ L, int lane) {
  const int gi = lane >> 2, q = lane & 3;
  RawLogits r;
  r.e0 = to_f(L[gi * 9 + 2 * q]);
  r.e1 = to_f(L[gi * 9 + 2 * q + 1]);
  r.e2 = (q == 0) ? to_f(L[gi * 9 + 8]) : 0.f;
  r.f0 = r.f1 = r.f2 = 0.f;
  if (gi == 0) {
    r.f0 = to_f(L[72 + 2 * q]);
    r.f1 = to_f(L[72 + 2 * q + 1]);
    if (q == 0) r.f2 = to_f(L[80]);
  }
  return r;
}

// softmax of one unit's 9x9 logits, produced directly in mma fragment layout (fp32, warp shuffles inside each quad).
//   pf : probabilities in accumulator layout: pf[0][0..1] = (P=g, Q=2q,2q+1), pf[0][2..3] = (P=g+8, same Q),
//        pf[1][0..1] = (P=g, Q=8+2q, 9+2q), pf[1][2..3] = (P=g+8, ...); rows / cols >= 9 are zero
__device__ __forceinline__ void softmax_frag(const RawLogits& r, float scale, int lane, float (&pf)[2][4]) {
  const int gi = lane >> 2, q = lane & 3;
  const float NEG = -INFINITY;
  float e0 = r.e0 * scale, e1 = r.e1 * scale;
  float e2 = (q == 0) ? r.e2 * scale : NEG;
  float f0 = NEG, f1 = NEG, f2 = NEG;
  if (gi == 0) {
    f0 = r.f0 * scale;
    f1 = r.f1 * scale;
    if (q == 0) f2 = r.f2 * scale;
  }
  const float m0 = quad_max(fmaxf(fmaxf(e0, e1), e2));
  const float m1 = quad_max(fmaxf(fmaxf(f0, f1), f2));     // -inf for gi != 0
  e0 = __expf(e0 - m0); e1 = __expf(e1 - m0); e2 = (q == 0) ? __expf(e2 - m0) : 0.f;
  const float inv0 = 1.f / quad_sum(e0 + e1 + e2);
  float inv1 = 0.f;
  if (gi == 0) { f0 = __expf(f0 - m1); f1 = __expf(f1 - m1); f2 = (q == 0) ? __expf(f2 - m1) : 0.f; } else { f0 = f1 = f2 = 0.f; }
  const float s1 = quad_sum(f0 + f1 + f2);
  if (gi == 0) inv1 = 1.f / s1;
  pf[0][0] = e0 * inv0; pf[0][1] = e1 * inv0; pf[0][2] = f0 * inv1; pf[0][3] = f1 * inv1;
  pf[1][0] = e2 * inv0; pf[1][1] = 0.f;       pf[1][2] = f2 * inv1; pf[1][3] = 0.f;
}

// smem address (bf16 elements) of pixel-row `idx` (0..15; >= 9 -> zero row) of window (lr, lc) for head `hd`
__device__ __forceinline__ const bf16* win_row(const bf16* tile, const bf16* zero, const Geo& g, int lr, int lc, int hd, int idx) {
  if (idx >= 9) return zero;
  const int py = 2 * lr + idx / 3, px = 2 * lc + idx % 3;
  return tile + ((size_t)py * g.PC + px) * g.Cp + hd * HD;
}

// res[16 x 32] += Afrag[16 x 16] . rows(tile)[16 x 32]   (B operand through ldmatrix.trans; rows = window pixels)
__device__ __forceinline__ void mma_rows(float (&acc)[4][4], const uint32_t (&a)[4], const bf16* tile, const bf16* zero,
                                         const Geo& g, int lr, int lc, int hd, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  const bf16* rowp = win_row(tile, zero, g, lr, lc, hd, (mi & 1) * 8 + r);
  const int coff = (rowp == zero) ? 0 : (mi >> 1) * 8;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t b[4];
    ldsm_x4_t(b, smem_u32(rowp + (rowp == zero ? 0 : half * 16) + coff));
    mma16816(acc[half * 2 + 0], a, b[0], b[1]);
    mma16816(acc[half * 2 + 1], a, b[2], b[3]);
  }
}

// accumulator rows 0..8 -> staging sOut[unit][9][OS]
__device__ __forceinline__ void stage_unit(float* sOut, int unit, const float (&acc)[4][4], int lane) {
  const int gi = lane >> 2, q = lane & 3;
  float* base = sOut + (size_t)unit * 9 * OS;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    *reinterpret_cast<float2*>(base + gi * OS + nb * 8 + 2 * q) = make_float2(acc[nb][0], acc[nb][1]);
    if (gi == 0) *reinterpret_cast<float2*>(base + 8 * OS + nb * 8 + 2 * q) = make_float2(acc[nb][2], acc[nb][3]);
  }
}

// fold as a gather: output pixel (ly, lx) of the band (local coords) sums its covering (window, row) entries
__device__ __forceinline__ void gather_store(const float* sOut, bf16* __restrict__ dst, const Geo& g, int b, int i0, int hd,
                                             int nWR) {
  const int C = g.heads * HD;
  const int rows = min(2 * g.TR, g.H - 2 * i0);
  const int tasks = rows * g.W * 8;                       // 8 threads (4 channels each) per pixel
  for (int t = threadIdx.x; t < tasks; t += NTHREADS) {
    const int c4 = (t & 7) * 4;
    const int lx = (t >> 3) % g.W, ly = (t >> 3) / g.W;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      // row: even -> (lr = ly/2, k=1); odd -> (lr=(ly-1)/2, k=2) and (lr=(ly+1)/2, k=0)
      int lr, ki;
      if ((ly & 1) == 0) { if (a) continue; lr = ly >> 1; ki = 1; }
      else { lr = (ly >> 1) + a; ki = a ? 0 : 2; }
      if (lr >= nWR || i0 + lr >= g.h) continue;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        int lc, kj;
        if ((lx & 1) == 0) { if (bb) continue; lc = lx >> 1; kj = 1; }
        else { lc = (lx >> 1) + bb; kj = bb ? 0 : 2; }
        if (lc >= g.w) continue;
        const float4 v4 = *reinterpret_cast<const float4*>(sOut + ((size_t)(lr * g.w + lc) * 9 + ki * 3 + kj) * OS + c4);
        s.x += v4.x; s.y += v4.y; s.z += v4.z; s.w += v4.w;
      }
    }
    uint2 pk;
    pk.x = pack_bf16(s.x, s.y);
    pk.y = pack_bf16(s.z, s.w);
    *reinterpret_cast<uint2*>(dst + (((size_t)b * g.H + 2 * i0 + ly) * g.W + lx) * C + hd * HD + c4) = pk;
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) outlook_fwd_mma_kernel(const bf16* __restrict__ v, const bf16* __restrict__ logits,
                                                                     bf16* __restrict__ y, Geo g) {
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* zero = reinterpret_cast<bf16*>(smraw);                                     // 64 bytes of zeros
  bf16* sV = zero + 32;
  float* sOut = reinterpret_cast<float*>(sV + (size_t)g.PR * g.PC * g.Cp);
  const int b = blockIdx.y, i0 = blockIdx.x * g.TR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = NTHREADS / 32;
  if (threadIdx.x < 32) zero[threadIdx.x] = __float2bfloat16_rn(0.f);
  stage_band(sV, v, g, b, 2 * i0 - 1);
  const int nWR = min(g.nWR, g.h - i0);
  const int units = nWR * g.w;
  // the logits of a warp's NEXT unit are fetched while the current one is computed (they are the only global loads
  // inside the unit loop; the prefetch also runs across the per-head barriers)
  const size_t lrow0 = ((size_t)b * g.h + i0) * g.w;
  RawLogits nxt = load_logits(logits + (lrow0 + (warp < units ? warp : 0)) * g.lpitch, lane);
  __syncthreads();
  for (int hd = 0; hd < g.heads; ++hd) {
    for (int u = warp; u < units; u += nwarp) {
      const int lr = u / g.w, lc = u % g.w;
      const RawLogits cur = nxt;
      {
        int nu = u + nwarp, nh = hd;
        if (nu >= units) { nu = warp; ++nh; }
        if (nh < g.heads && nu < units) nxt = load_logits(logits + (lrow0 + nu) * g.lpitch + nh * 81, lane);
      }
      float pf[2][4];
      softmax_frag(cur, g.scale, lane, pf);
      const uint32_t a[4] = {pack_bf16(pf[0][0], pf[0][1]), pack_bf16(pf[0][2], pf[0][3]), pack_bf16(pf[1][0], pf[1][1]),
                             pack_bf16(pf[1][2], pf[1][3])};
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      mma_rows(acc, a, sV, zero, g, lr, lc, hd, lane);
      stage_unit(sOut, u, acc, lane);
    }
    __syncthreads();
    gather_store(sOut, y, g, b, i0, hd, nWR);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) outlook_bwd_mma_kernel(const bf16* __restrict__ v, const bf16* __restrict__ logits,
                                                                     const bf16* __restrict__ dy, bf16* __restrict__ dv,
                                                                     bf16* __restrict__ dlogits, Geo g) {
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* zero = reinterpret_cast<bf16*>(smraw);
  bf16* sV = zero + 32;
  bf16* sG = sV + (size_t)g.PR * g.PC * g.Cp;
  float* sOut = reinterpret_cast<float*>(sG + (size_t)g.PR * g.PC * g.Cp);
  const int b = blockIdx.y, i0 = blockIdx.x * g.TR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = NTHREADS / 32;
  const int gi = lane >> 2, q = lane & 3;
  if (threadIdx.x < 32) zero[threadIdx.x] = __float2bfloat16_rn(0.f);
  stage_band(sV, v, g, b, 2 * i0 - 1);
  stage_band(sG, dy, g, b, 2 * i0 - 1);
  const int nWR = min(g.nWR, g.h - i0);
  const int units = nWR * g.w;
  const size_t lrow0 = ((size_t)b * g.h + i0) * g.w;
  RawLogits nxt = load_logits(logits + (lrow0 + (warp < units ? warp : 0)) * g.lpitch, lane);
  __syncthreads();
  for (int hd = 0; hd < g.heads; ++hd) {
    for (int u = warp; u < units; u += nwarp) {
      const int lr = u / g.w, lc = u % g.w;
      const size_t lbase = (lrow0 + u) * g.lpitch + hd * 81;
      const RawLogits cur = nxt;
      {
        int nu = u + nwarp, nh = hd;
        if (nu >= units) { nu = warp; ++nh; }
        if (nh < g.heads && nu < units) nxt = load_logits(logits + (lrow0 + nu) * g.lpitch + nh * 81, lane);
      }
      float pf[2][4];
      softmax_frag(cur, g.scale, lane, pf);
      // ---- dA[P][Q] = sum_c dy[pix P][c] v[pix Q][c] : A operand = dy rows, B operand ("col") = v rows
      float da[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) da[nb][0] = da[nb][1] = da[nb][2] = da[nb][3] = 0.f;
      {
        const int mi = lane >> 3, r = lane & 7;
        // dy rows as A operand: matrix mi -> rows (mi&1)*8 + r, channels (mi>>1)*8 (+16 for the second k-step)
        const bf16* ga = win_row(sG, zero, g, lr, lc, hd, (mi & 1) * 8 + r);
        const int gco = (ga == zero) ? 0 : (mi >> 1) * 8;
        uint32_t a0[4], a1[4];
        ldsm_x4(a0, smem_u32(ga + gco));
        ldsm_x4(a1, smem_u32(ga + (ga == zero ? 0 : 16) + gco));
        // v rows as B operand: for n-block nb, rows Q = nb*8 + r, matrix mi -> channel chunk mi*8
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          const bf16* vb = win_row(sV, zero, g, lr, lc, hd, nb * 8 + r);
          uint32_t bq[4];
          ldsm_x4(bq, smem_u32(vb + (vb == zero ? 0 : mi * 8)));
          mma16816(da[nb], a0, bq[0], bq[1]);
          mma16816(da[nb], a1, bq[2], bq[3]);
        }
      }
      // ---- dlogits = scale * A o (dA - sum_Q A o dA)   (only the own unit's rows: the halo band belongs to the neighbour)
      float r0 = pf[0][0] * da[0][0] + pf[0][1] * da[0][1] + pf[1][0] * da[1][0];
      float r1 = pf[0][2] * da[0][2] + pf[0][3] * da[0][3] + pf[1][2] * da[1][2];
      r0 = quad_sum(r0);
      r1 = quad_sum(r1);
      if (lr < g.TR) {
        bf16* dl = dlogits + lbase;
        dl[gi * 9 + 2 * q] = __float2bfloat16_rn(g.scale * pf[0][0] * (da[0][0] - r0));
        dl[gi * 9 + 2 * q + 1] = __float2bfloat16_rn(g.scale * pf[0][1] * (da[0][1] - r0));
        if (q == 0) dl[gi * 9 + 8] = __float2bfloat16_rn(g.scale * pf[1][0] * (da[1][0] - r0));
        if (gi == 0) {
          dl[72 + 2 * q] = __float2bfloat16_rn(g.scale * pf[0][2] * (da[0][2] - r1));
          dl[72 + 2 * q + 1] = __float2bfloat16_rn(g.scale * pf[0][3] * (da[0][3] - r1));
          if (q == 0) dl[80] = __float2bfloat16_rn(g.scale * pf[1][2] * (da[1][2] - r1));
        }
        if (hd == 0 && lane >= 9 && lane - 9 < g.lpitch - g.heads * 81)
          dlogits[lbase + g.heads * 81 + (lane - 9)] = __float2bfloat16_rn(0.f);     // zero the TMA padding columns
      }
      // ---- dvw[Q][c] = sum_P A[P][Q] dy[pix P][c] : A operand = A^T (movmatrix), B operand = dy rows (ldmatrix.trans)
      const uint32_t at[4] = {movmatrix_t(pack_bf16(pf[0][0], pf[0][1])), movmatrix_t(pack_bf16(pf[1][0], pf[1][1])),
                              movmatrix_t(pack_bf16(pf[0][2], pf[0][3])), movmatrix_t(pack_bf16(pf[1][2], pf[1][3]))};
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      mma_rows(acc, at, sG, zero, g, lr, lc, hd, lane);
      stage_unit(sOut, u, acc, lane);
    }
    __syncthreads();
    gather_store(sOut, dv, g, b, i0, hd, nWR);
    __syncthreads();
  }
}

int plan(Geo& g, bool bwd, size_t& smem) {
  const int C = g.heads * HD;
  g.Cp = C + 8;
  g.PC = 2 * g.w + 3;
  // prefer a band small enough for two CTAs per SM (one CTA's staging / stores overlap the other's math)
  const size_t limits[2] = {(size_t)110 * 1024, (size_t)200 * 1024};
  for (int pass = 0; pass < 2; ++pass)
    for (int tr = 4; tr >= 1; --tr) {
      g.TR = tr;
      g.nWR = tr + 1;
      g.PR = 2 * tr + 3;
      const size_t tile = (size_t)g.PR * g.PC * g.Cp * sizeof(bf16);
      smem = 64 + tile * (bwd ? 2 : 1) + (size_t)g.nWR * g.w * 9 * OS * sizeof(float);
      if (smem <= limits[pass]) return 0;
    }
  return APB_ERR_UNSUPPORTED;
}

}  // namespace

int apb_outlook_fwd_mma(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int lpitch,
                        cudaStream_t st) {
  Geo g;
  g.B = B; g.H = H; g.W = W; g.heads = heads; g.h = (H + 1) / 2; g.w = (W + 1) / 2; g.lpitch = lpitch; g.scale = scale;
  size_t smem;
  if (plan(g, false, smem) != 0) return APB_ERR_UNSUPPORTED;
  cudaFuncSetAttribute(outlook_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ceil_div(g.h, g.TR), B);
  outlook_fwd_mma_kernel<<<grid, NTHREADS, smem, st>>>((const bf16*)v, (const bf16*)logits, (bf16*)y, g);
  APB_LAUNCH_CHECK("outlook_fwd_mma");
  return 0;
}

int apb_outlook_bwd_mma(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                        int heads, float scale, int lpitch, cudaStream_t st) {
  Geo g;
  g.B = B; g.H = H; g.W = W; g.heads = heads; g.h = (H + 1) / 2; g.w = (W + 1) / 2; g.lpitch = lpitch; g.scale = scale;
  size_t smem;
  if (plan(g, true, smem) != 0) return APB_ERR_UNSUPPORTED;
  cudaFuncSetAttribute(outlook_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ceil_div(g.h, g.TR), B);
  outlook_bwd_mma_kernel<<<grid, NTHREADS, smem, st>>>((const bf16*)v, (const bf16*)logits, (const bf16*)dy, (bf16*)dv,
                                                       (bf16*)dlogits, g);
  APB_LAUNCH_CHECK("outlook_bwd_mma");
  return 0;
}
