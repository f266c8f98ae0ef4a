// OutlookAttention core on tensor cores (bf16 I/O, fp32 softmax + accumulation), one fused kernel per direction.
//
//   reference: nn.Unfold(3,1,2) -> softmax(scale*logits) -> attn @ v -> F.fold          (models/volo.py:83-98)
//
// A CTA owns a band of TR window rows x TC window columns of one image (all heads; TC = all columns whenever that fits
// shared memory -- every 224-px VOLO stage -- else the grid is also tiled along x, e.g. the 48x48 grids of volo_d2@384).
// It stages the (2TR+3) x (2TC+3) pixel band of v (and dy in the backward) ONCE in shared memory with coalesced 16-byte
// loads; every (window, head) unit
// then is a 16x16x16 mma.sync problem whose operands come straight from that tile:
//   forward : out[P][c]  = sum_Q A[P][Q] v[pix(Q)][c]        A = softmax fragment built in registers (warp shuffles)
//   backward: dA[P][Q]   = <dy[pix(P)], v[pix(Q)]>           ; dlogits = scale * A o (dA - rowsum(A o dA))
//             dvw[Q][c]  = sum_P A[P][Q] dy[pix(P)][c]       (A^T through movmatrix)
// ldmatrix row addresses do the 3x3 unfold for free (out-of-image pixels are zero rows of the tile).  The fold is a
// deterministic gather: per head, unit results are staged in shared memory and every output pixel sums the 1/2/4
// window rows that cover it -- no atomics.  HBM traffic is exactly one read of v / logits (/ dy) and one write of
// y (/ dv, dlogits); the halo rows re-read by neighbouring bands hit L2.
// The kernel is issue-bound, so everything loop-invariant (per-lane ldmatrix / staging offsets) is hoisted out of the
// unit loop and all shared-memory addressing is 32-bit.
#include "common.cuh"

namespace {

constexpr int HD = 32;
constexpr int OS = 40;          // fp32 staging row pitch in floats (32 channels + 8 pad: conflict-free 64-bit stores)

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
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
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
  int TR, TC;          // window rows / columns owned per CTA
  int nWR, nWC;        // window rows / columns computed per CTA (+1 halo on the far side when the axis is tiled)
  int nCB;             // CTAs along x
  int PR, PC, Cp;      // staged pixel rows / cols, padded channel pitch (bf16 elements)
  float scale;
};

// Compile-time geometry: the kernels are bound by integer / address arithmetic on the runtime geometry, so the shapes
// of the AutoProg stages (square stage-1 grids of 16 / 20 / 24 / 28 pixels, 6 heads, untiled x axis) get instantiations in
// which every Geo field except B and H is a constant the compiler folds.  CW = 0: fully dynamic.
template <int CW, int CHEADS, int CTR>
__device__ __forceinline__ Geo fix_geo(Geo g) {
  if constexpr (CW > 0) {
    constexpr int cw = (CW + 1) / 2;
    g.W = CW; g.w = cw; g.heads = CHEADS; g.lpitch = (CHEADS * 81 + 7) / 8 * 8;
    g.TR = CTR; g.nWR = CTR + 1; g.PR = 2 * CTR + 3;
    g.TC = cw; g.nWC = cw; g.nCB = 1; g.PC = 2 * cw + 1; g.Cp = CHEADS * HD + 8;
  }
  return g;
}

// stage pixel rows [y0, y0+PR) x cols [x0, x0+PC) of src[b] (NHWC bf16) into s[PR][PC][Cp]; outside the image -> zeros
template <int NT>
__device__ __forceinline__ void stage_band(bf16* s, const bf16* __restrict__ src, const Geo& g, int b, int y0, int x0) {
  const int C = g.heads * HD;
  const int cpr = C >> 3;                                  // 16-byte chunks per pixel
  const int per_row = g.PC * cpr;
  const bf16* img = src + (size_t)b * g.H * g.W * C;
  for (int py = 0; py < g.PR; ++py) {
    const int y = y0 + py;
    const bool yok = (y >= 0 && y < g.H);
    const bf16* grow = img + ((ptrdiff_t)(yok ? y : 0) * g.W + x0) * C;   // column x = x0 + px
    bf16* srow = s + (size_t)py * g.PC * g.Cp;
    for (int e = threadIdx.x; e < per_row; e += NT) {
      const int px = e / cpr, ch = e - px * cpr;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (yok && x0 + px >= 0 && x0 + px < g.W) val = *reinterpret_cast<const uint4*>(grow + (ptrdiff_t)px * C + ch * 8);
      *reinterpret_cast<uint4*>(srow + px * g.Cp + ch * 8) = val;
    }
  }
}

// raw logits of one unit in fragment order: (P=g: Q=2q, 2q+1, 8) and, for g == 0, (P=8: Q=2q, 2q+1, 8)
struct RawLogits { float e0, e1, e2, f0, f1, f2; };
__device__ __forceinline__ RawLogits load_logits(const bf16* __restrict__ L, int gi, int q) {
  RawLogits r;
  const bf16* row = L + gi * 9 + 2 * q;
  r.e0 = to_f(row[0]);
  r.e1 = to_f(row[1]);
  r.e2 = (q == 0) ? to_f(row[8]) : 0.f;
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
__device__ __forceinline__ void softmax_frag(const RawLogits& r, float sl2, int gi, int q, float (&pf)[2][4]) {
  const float NEG = -INFINITY;
  float e0 = r.e0 * sl2, e1 = r.e1 * sl2;                 // logits * scale * log2(e)
  float e2 = (q == 0) ? r.e2 * sl2 : NEG;
  float f0 = (gi == 0) ? r.f0 * sl2 : NEG, f1 = (gi == 0) ? r.f1 * sl2 : NEG;
  float f2 = (gi == 0 && q == 0) ? r.f2 * sl2 : NEG;
  const float m0 = quad_max(fmaxf(fmaxf(e0, e1), e2));
  float m1 = quad_max(fmaxf(fmaxf(f0, f1), f2));
  m1 = (gi == 0) ? m1 : 0.f;                              // keeps exp2(-inf - m1) = 0 without NaNs
  e0 = exp2f(e0 - m0); e1 = exp2f(e1 - m0); e2 = exp2f(e2 - m0);
  f0 = exp2f(f0 - m1); f1 = exp2f(f1 - m1); f2 = exp2f(f2 - m1);
  const float inv0 = __frcp_rn(quad_sum(e0 + e1 + e2));
  const float s1 = quad_sum(f0 + f1 + f2);
  const float inv1 = (gi == 0) ? __frcp_rn(s1) : 0.f;
  pf[0][0] = e0 * inv0; pf[0][1] = e1 * inv0; pf[0][2] = f0 * inv1; pf[0][3] = f1 * inv1;
  pf[1][0] = e2 * inv0; pf[1][1] = 0.f;       pf[1][2] = f2 * inv1; pf[1][3] = 0.f;
}

// loop-invariant per-lane geometry
struct LaneGeo {
  uint32_t offT;      // byte offset (from the unit's base pixel) of this lane's ldmatrix row for "rows x channels" operands
  uint32_t offB[2];   // byte offsets of this lane's row for the B operand of dA (n-block 0 / 1)
  bool zT, zB[2];     // row index >= 9 -> read the zero row instead
  uint32_t offS;      // byte offset of this lane inside a unit's staging block
};
__device__ __forceinline__ uint32_t pix_off(const Geo& g, int idx) { return (uint32_t)(((idx / 3) * g.PC + idx % 3) * g.Cp * 2); }
__device__ __forceinline__ LaneGeo lane_geo(const Geo& g, int lane) {
  LaneGeo L;
  const int mi = lane >> 3, r = lane & 7;
  const int idxT = (mi & 1) * 8 + r;
  L.zT = idxT >= 9;
  L.offT = L.zT ? 0u : pix_off(g, idxT) + (uint32_t)((mi >> 1) * 16);
#pragma unroll
  for (int nb = 0; nb < 2; ++nb) {
    const int idxB = nb * 8 + r;
    L.zB[nb] = idxB >= 9;
    L.offB[nb] = L.zB[nb] ? 0u : pix_off(g, idxB) + (uint32_t)(mi * 16);
  }
  L.offS = (uint32_t)(((lane >> 2) * OS + 2 * (lane & 3)) * 4);
  return L;
}

// res[16 x 32] += Afrag[16 x 16] . rows(tile)[16 x 32]   (B operand through ldmatrix.trans; rows = window pixels)
__device__ __forceinline__ void mma_rows(float (&acc)[4][4], const uint32_t (&a)[4], uint32_t unit_base, uint32_t zero_s,
                                         const LaneGeo& L) {
  const uint32_t a0 = L.zT ? zero_s : unit_base + L.offT;
  const uint32_t a1 = L.zT ? zero_s : a0 + 32;           // channels 16..31
  uint32_t b[4];
  ldsm_x4_t(b, a0);
  mma16816(acc[0], a, b[0], b[1]);
  mma16816(acc[1], a, b[2], b[3]);
  ldsm_x4_t(b, a1);
  mma16816(acc[2], a, b[0], b[1]);
  mma16816(acc[3], a, b[2], b[3]);
}

// accumulator rows 0..8 -> staging block of the unit ([9][OS] floats)
__device__ __forceinline__ void stage_unit(uint32_t blk, const float (&acc)[4][4], const LaneGeo& L, int gi) {
  const uint32_t p = blk + L.offS;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) sts64(p + nb * 32, acc[nb][0], acc[nb][1]);
  if (gi == 0) {
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) sts64(p + 8 * OS * 4 + nb * 32, acc[nb][2], acc[nb][3]);
  }
}

// fold as a gather: output pixel (ly, lx) of the band (local coords) sums its covering (window, row) staging entries
template <int NT>
__device__ __forceinline__ void gather_store(uint32_t sOut_s, bf16* __restrict__ dst_band, const Geo& g, int rows, int cols,
                                             int nWR_eff, int wb, int hd, int ly0, int lx0) {
  const int C = g.heads * HD;
  const int c4 = (threadIdx.x & 7) * 4;
  constexpr int S = NT / 8;                                // pixels handled per sweep
  int ly = ly0, lx = lx0;
  while (ly < rows) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    // rows: even -> (lr = ly/2, ki = 1); odd -> (lr = (ly-1)/2, ki = 2) and (lr = (ly+1)/2, ki = 0); same for columns
    const int lr0 = ly >> 1, lc0 = lx >> 1;
    const int ki0 = (ly & 1) ? 2 : 1, kj0 = (lx & 1) ? 2 : 1;
    const bool r2 = (ly & 1) && (lr0 + 1 < nWR_eff), c2 = (lx & 1) && (lc0 + 1 < wb);
    const uint32_t e00 = sOut_s + (uint32_t)((((lr0 * wb + lc0) * 9 + ki0 * 3 + kj0) * OS + c4) * 4);
    {
      const float4 t = lds128(e00);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    if (c2) {   // window (lr0, lc0+1), kj = 0
      const float4 t = lds128(e00 + (uint32_t)((9 - kj0) * OS * 4));
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    if (r2) {   // window (lr0+1, lc0), ki = 0
      const uint32_t e10 = e00 + (uint32_t)((wb * 9 - ki0 * 3) * OS * 4);
      const float4 t = lds128(e10);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      if (c2) {
        const float4 t2 = lds128(e10 + (uint32_t)((9 - kj0) * OS * 4));
        s.x += t2.x; s.y += t2.y; s.z += t2.z; s.w += t2.w;
      }
    }
    uint2 pk;
    pk.x = pack_bf16(s.x, s.y);
    pk.y = pack_bf16(s.z, s.w);
    *reinterpret_cast<uint2*>(dst_band + ((size_t)ly * g.W + lx) * C + hd * HD + c4) = pk;
    lx += S;
    while (lx >= cols) { lx -= cols; ++ly; }
  }
}

template <int NT, int CW, int CHEADS, int CTR>
__global__ void __launch_bounds__(NT, 1) outlook_fwd_mma_kernel(const bf16* __restrict__ v, const bf16* __restrict__ logits,
                                                               bf16* __restrict__ y, Geo gin) {
  const Geo g = fix_geo<CW, CHEADS, CTR>(gin);
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* zero = reinterpret_cast<bf16*>(smraw);                                     // 64 bytes of zeros
  bf16* sV = zero + 32;
  const uint32_t tile_bytes = (uint32_t)(g.PR * g.PC * g.Cp * 2);
  const uint32_t zero_s = smem_u32(zero), sV_s = zero_s + 64, sOut_s = sV_s + tile_bytes;
  const int b = blockIdx.y, bi = blockIdx.x / g.nCB, bj = blockIdx.x - bi * g.nCB;
  const int i0 = bi * g.TR, j0 = bj * g.TC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int nwarp = NT / 32;
  const int gi = lane >> 2, q = lane & 3;
  if (threadIdx.x < 32) zero[threadIdx.x] = __float2bfloat16_rn(0.f);
  stage_band<NT>(sV, v, g, b, 2 * i0 - 1, 2 * j0 - 1);
  const int nWR = min(g.nWR, g.h - i0);
  const int wb = min(g.nWC, g.w - j0);
  const int units = nWR * wb;
  const LaneGeo L = lane_geo(g, lane);
  const float sl2 = g.scale * 1.4426950408889634f;
  const int rows = min(2 * g.TR, g.H - 2 * i0), cols = min(2 * g.TC, g.W - 2 * j0);
  const int ps = threadIdx.x >> 3;
  const int ly0 = ps / cols, lx0 = ps - ly0 * cols;
  bf16* ydst = y + (((size_t)b * g.H + 2 * i0) * g.W + 2 * j0) * (g.heads * HD);
  const uint32_t row_pitch = (uint32_t)(g.PC * g.Cp * 2);
  // logits of a warp's NEXT unit are fetched while the current one is computed (the only global loads in the loop)
  const bf16* lbase = logits + (((size_t)b * g.h + i0) * g.w + j0) * g.lpitch;
  const int lr_f = warp / wb, lc_f = warp - lr_f * wb;           // this warp's first unit of every head
  RawLogits nxt = load_logits(lbase + (size_t)(warp < units ? lr_f * g.w + lc_f : 0) * g.lpitch, gi, q);
  __syncthreads();
  for (int hd = 0; hd < g.heads; ++hd) {
    int lr = lr_f, lc = lc_f;
    for (int u = warp; u < units; u += nwarp) {
      const RawLogits cur = nxt;
      int nlr = lr, nlc = lc + nwarp;
      while (nlc >= wb) { nlc -= wb; ++nlr; }
      {
        int plr = nlr, plc = nlc, nh = hd;
        if (u + nwarp >= units) { plr = lr_f; plc = lc_f; ++nh; }
        if (nh < g.heads) nxt = load_logits(lbase + (size_t)(plr * g.w + plc) * g.lpitch + nh * 81, gi, q);
      }
      float pf[2][4];
      softmax_frag(cur, sl2, gi, q, pf);
      const uint32_t a[4] = {pack_bf16(pf[0][0], pf[0][1]), pack_bf16(pf[0][2], pf[0][3]), pack_bf16(pf[1][0], pf[1][1]),
                             pack_bf16(pf[1][2], pf[1][3])};
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      const uint32_t ub = sV_s + (uint32_t)(2 * lr) * row_pitch + (uint32_t)((2 * lc * g.Cp + hd * HD) * 2);
      mma_rows(acc, a, ub, zero_s, L);
      stage_unit(sOut_s + (uint32_t)(u * 9 * OS * 4), acc, L, gi);
      lr = nlr;
      lc = nlc;
    }
    __syncthreads();
    gather_store<NT>(sOut_s, ydst, g, rows, cols, nWR, wb, hd, ly0, lx0);
    __syncthreads();
  }
}

template <int NT, int CW, int CHEADS, int CTR>
__global__ void __launch_bounds__(NT, 1) outlook_bwd_mma_kernel(const bf16* __restrict__ v, const bf16* __restrict__ logits,
                                                               const bf16* __restrict__ dy, bf16* __restrict__ dv,
                                                               bf16* __restrict__ dlogits, Geo gin) {
  const Geo g = fix_geo<CW, CHEADS, CTR>(gin);
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* zero = reinterpret_cast<bf16*>(smraw);
  bf16* sV = zero + 32;
  const uint32_t tile_bytes = (uint32_t)(g.PR * g.PC * g.Cp * 2);
  bf16* sG = sV + (size_t)g.PR * g.PC * g.Cp;
  const uint32_t zero_s = smem_u32(zero), sV_s = zero_s + 64, sG_s = sV_s + tile_bytes, sOut_s = sG_s + tile_bytes;
  const int b = blockIdx.y, bi = blockIdx.x / g.nCB, bj = blockIdx.x - bi * g.nCB;
  const int i0 = bi * g.TR, j0 = bj * g.TC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int nwarp = NT / 32;
  const int gi = lane >> 2, q = lane & 3;
  if (threadIdx.x < 32) zero[threadIdx.x] = __float2bfloat16_rn(0.f);
  stage_band<NT>(sV, v, g, b, 2 * i0 - 1, 2 * j0 - 1);
  stage_band<NT>(sG, dy, g, b, 2 * i0 - 1, 2 * j0 - 1);
  const int nWR = min(g.nWR, g.h - i0);
  const int wb = min(g.nWC, g.w - j0);
  const int units = nWR * wb;
  const LaneGeo L = lane_geo(g, lane);
  const float sl2 = g.scale * 1.4426950408889634f;
  const int rows = min(2 * g.TR, g.H - 2 * i0), cols = min(2 * g.TC, g.W - 2 * j0);
  const int ps = threadIdx.x >> 3;
  const int ly0 = ps / cols, lx0 = ps - ly0 * cols;
  bf16* ddst = dv + (((size_t)b * g.H + 2 * i0) * g.W + 2 * j0) * (g.heads * HD);
  const uint32_t row_pitch = (uint32_t)(g.PC * g.Cp * 2);
  const int npad = g.lpitch - g.heads * 81;
  const size_t lband = (((size_t)b * g.h + i0) * g.w + j0) * g.lpitch;
  const bf16* lbase = logits + lband;
  bf16* dlbase = dlogits + lband;
  const int lr_f = warp / wb, lc_f = warp - lr_f * wb;
  RawLogits nxt = load_logits(lbase + (size_t)(warp < units ? lr_f * g.w + lc_f : 0) * g.lpitch, gi, q);
  __syncthreads();
  for (int hd = 0; hd < g.heads; ++hd) {
    int lr = lr_f, lc = lc_f;
    for (int u = warp; u < units; u += nwarp) {
      const RawLogits cur = nxt;
      int nlr = lr, nlc = lc + nwarp;
      while (nlc >= wb) { nlc -= wb; ++nlr; }
      {
        int plr = nlr, plc = nlc, nh = hd;
        if (u + nwarp >= units) { plr = lr_f; plc = lc_f; ++nh; }
        if (nh < g.heads) nxt = load_logits(lbase + (size_t)(plr * g.w + plc) * g.lpitch + nh * 81, gi, q);
      }
      float pf[2][4];
      softmax_frag(cur, sl2, gi, q, pf);
      const uint32_t uoff = (uint32_t)(2 * lr) * row_pitch + (uint32_t)((2 * lc * g.Cp + hd * HD) * 2);
      // ---- dA[P][Q] = sum_c dy[pix P][c] v[pix Q][c] : A operand = dy rows, B operand ("col") = v rows
      float da[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) da[nb][0] = da[nb][1] = da[nb][2] = da[nb][3] = 0.f;
      {
        uint32_t a0[4], a1[4], bq[4];
        const uint32_t ga = L.zT ? zero_s : sG_s + uoff + L.offT;
        ldsm_x4(a0, ga);
        ldsm_x4(a1, L.zT ? zero_s : ga + 32);
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          ldsm_x4(bq, L.zB[nb] ? zero_s : sV_s + uoff + L.offB[nb]);
          mma16816(da[nb], a0, bq[0], bq[1]);
          mma16816(da[nb], a1, bq[2], bq[3]);
        }
      }
      // ---- dlogits = scale * A o (dA - sum_Q A o dA)   (own windows only: halo windows belong to the neighbouring CTA)
      float r0 = pf[0][0] * da[0][0] + pf[0][1] * da[0][1] + pf[1][0] * da[1][0];
      float r1 = pf[0][2] * da[0][2] + pf[0][3] * da[0][3] + pf[1][2] * da[1][2];
      r0 = quad_sum(r0);
      r1 = quad_sum(r1);
      if (lr < g.TR && lc < g.TC) {
        const size_t lu = (size_t)(lr * g.w + lc) * g.lpitch;
        bf16* dl = dlbase + lu + hd * 81;
        bf16* drow = dl + gi * 9 + 2 * q;
        drow[0] = __float2bfloat16_rn(g.scale * pf[0][0] * (da[0][0] - r0));
        drow[1] = __float2bfloat16_rn(g.scale * pf[0][1] * (da[0][1] - r0));
        if (q == 0) drow[8] = __float2bfloat16_rn(g.scale * pf[1][0] * (da[1][0] - r0));
        if (gi == 0) {
          dl[72 + 2 * q] = __float2bfloat16_rn(g.scale * pf[0][2] * (da[0][2] - r1));
          dl[72 + 2 * q + 1] = __float2bfloat16_rn(g.scale * pf[0][3] * (da[0][3] - r1));
          if (q == 0) dl[80] = __float2bfloat16_rn(g.scale * pf[1][2] * (da[1][2] - r1));
        }
        if (hd == 0 && lane >= 9 && lane - 9 < npad)
          dlbase[lu + g.heads * 81 + (lane - 9)] = __float2bfloat16_rn(0.f);   // TMA padding columns
      }
      // ---- dvw[Q][c] = sum_P A[P][Q] dy[pix P][c] : A operand = A^T (movmatrix), B operand = dy rows (ldmatrix.trans)
      const uint32_t at[4] = {movmatrix_t(pack_bf16(pf[0][0], pf[0][1])), movmatrix_t(pack_bf16(pf[1][0], pf[1][1])),
                              movmatrix_t(pack_bf16(pf[0][2], pf[0][3])), movmatrix_t(pack_bf16(pf[1][2], pf[1][3]))};
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      mma_rows(acc, at, sG_s + uoff, zero_s, L);
      stage_unit(sOut_s + (uint32_t)(u * 9 * OS * 4), acc, L, gi);
      lr = nlr;
      lc = nlc;
    }
    __syncthreads();
    gather_store<NT>(sOut_s, ddst, g, rows, cols, nWR, wb, hd, ly0, lx0);
    __syncthreads();
  }
}

// Pick the (TR, TC) band with the least recomputation that fits shared memory.  A tiled axis recomputes one halo
// window row / column per CTA: work factor (TR+1)/TR * (TC+1)/TC; an untiled x axis (TC = w) has no column halo.
int plan(Geo& g, bool bwd, size_t& smem) {
  const int C = g.heads * HD;
  g.Cp = C + 8;
  float best = 1e30f;
  int btr = 0, btc = 0;
  for (int tr = 3; tr >= 1; --tr) {
    for (int nsplit = 1; nsplit <= 8; ++nsplit) {
      const int tc = ceil_div(g.w, nsplit);
      const int nwc = (nsplit == 1) ? g.w : tc + 1;
      const size_t tile = (size_t)(2 * tr + 3) * (2 * nwc + 1) * g.Cp * sizeof(bf16);
      const size_t need = 64 + tile * (bwd ? 2 : 1) + (size_t)(tr + 1) * nwc * 9 * OS * sizeof(float);
      if (need > 220 * 1024) continue;
      const float cost = (float)(tr + 1) / tr * (float)nwc / tc;
      if (cost < best - 1e-6f) { best = cost; btr = tr; btc = tc; }
      break;                                  // more column splits at this TR only add halo work
    }
  }
  if (btr == 0) return APB_ERR_UNSUPPORTED;
  g.TR = btr;
  g.nWR = btr + 1;
  g.PR = 2 * btr + 3;
  g.TC = btc;
  g.nCB = ceil_div(g.w, btc);
  g.nWC = (g.nCB == 1) ? g.w : btc + 1;
  g.PC = 2 * g.nWC + 1;
  smem = 64 + (size_t)g.PR * g.PC * g.Cp * sizeof(bf16) * (bwd ? 2 : 1) + (size_t)g.nWR * g.nWC * 9 * OS * sizeof(float);
  return 0;
}

Geo make_geo(int B, int H, int W, int heads, float scale, int lpitch) {
  Geo g;
  g.B = B; g.H = H; g.W = W; g.heads = heads; g.h = (H + 1) / 2; g.w = (W + 1) / 2; g.lpitch = lpitch; g.scale = scale;
  return g;
}

}  // namespace

int apb_outlook_fwd_mma(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int lpitch,
                        apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  Geo g = make_geo(B, H, W, heads, scale, lpitch);
  size_t smem;
  if (plan(g, false, smem) != 0) return APB_ERR_UNSUPPORTED;
  constexpr int NT = 1024;
  dim3 grid(ceil_div(g.h, g.TR) * g.nCB, B);
#define OL_FWD(CW_, CH_, CTR_)                                                                                             \
  do {                                                                                                                     \
    cudaFuncSetAttribute(outlook_fwd_mma_kernel<NT, CW_, CH_, CTR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    outlook_fwd_mma_kernel<NT, CW_, CH_, CTR_><<<grid, NT, smem, st>>>((const bf16*)v, (const bf16*)logits, (bf16*)y, g);     \
  } while (0)
  const bool std6 = heads == 6 && lpitch == 488 && g.nCB == 1 && g.TR == 3;
  if (std6 && W == 28) OL_FWD(28, 6, 3);
  else if (std6 && W == 24) OL_FWD(24, 6, 3);
  else if (std6 && W == 20) OL_FWD(20, 6, 3);
  else if (std6 && W == 16) OL_FWD(16, 6, 3);
  else OL_FWD(0, 0, 0);
#undef OL_FWD
  APB_LAUNCH_CHECK("outlook_fwd_mma");
  return 0;
}

int apb_outlook_bwd_mma(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                        int heads, float scale, int lpitch, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  Geo g = make_geo(B, H, W, heads, scale, lpitch);
  size_t smem;
  if (plan(g, true, smem) != 0) return APB_ERR_UNSUPPORTED;
  // thread count: measured at 28 x 28 x 192 (one CTA / SM, so warps per CTA = latency hiding): 448 -> 275 us, 640 -> 263,
  // 672 -> 242, 896 / 960 / 1024 -> 189 us.  The specialised geometries compile to <= 72 registers and run with 896;
  // the dynamic-geometry kernel (90 registers) stays at 640.
  dim3 grid(ceil_div(g.h, g.TR) * g.nCB, B);
#define OL_BWD_NT(NT_, CW_, CH_, CTR_)                                                                                        \
  do {                                                                                                                        \
    cudaFuncSetAttribute(outlook_bwd_mma_kernel<NT_, CW_, CH_, CTR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    outlook_bwd_mma_kernel<NT_, CW_, CH_, CTR_><<<grid, NT_, smem, st>>>((const bf16*)v, (const bf16*)logits, (const bf16*)dy, \
                                                                        (bf16*)dv, (bf16*)dlogits, g);                       \
  } while (0)
#define OL_BWD(CW_, CH_, CTR_) OL_BWD_NT(896, CW_, CH_, CTR_)
  const bool std6 = heads == 6 && lpitch == 488 && g.nCB == 1;
  if (std6 && W == 28 && g.TR == 2) OL_BWD(28, 6, 2);
  else if (std6 && W == 24 && g.TR == 2) OL_BWD(24, 6, 2);
  else if (std6 && W == 20 && g.TR == 3) OL_BWD(20, 6, 3);
  else if (std6 && W == 16 && g.TR == 3) OL_BWD(16, 6, 3);
  else OL_BWD_NT(640, 0, 0, 0);
#undef OL_BWD
#undef OL_BWD_NT
  APB_LAUNCH_CHECK("outlook_bwd_mma");
  return 0;
}
