// OutlookAttention core BACKWARD, gather formulation (bf16 I/O, fp32 softmax + accumulation), one kernel.
//
//   reference: autograd of nn.Unfold(3,1,2) -> softmax(scale*logits) -> attn @ v -> F.fold          (models/volo.py:83-98)
//
//   dV[y, x, head, :]  = sum over the 1 / 2 / 4 windows (i, j) covering (y, x), Q = position of (y, x) in the window,
//                        sum_P A[i, j, head, P, Q] * dY[pixel P of window (i, j), head, :]
//                        -- the forward gather (outlook_fma.cu) with A transposed and dY in place of V
//   dA[P, Q]           = < dY[pixel P], V[pixel Q] >  over the head's 32 channels
//   dlogits[P, Q]      = scale * A[P, Q] * (dA[P, Q] - sum_Q' A[P, Q'] dA[P, Q'])
// CTA = window row r x a range of window columns x all heads.  It owns the dlogits of those windows and the dV of output rows
// 2r, 2r+1 under them.
//   stage 1: the 5-row pixel bands of v and dy (zero border, channel pitch C + 8) and the raw logits rows of window rows r and
//            r+1 -> shared memory with cp.async (everything in flight at once, nothing staged in registers)
//   stage 2: one thread per (window row slot, window, head, P): softmax of its 9 logits -> scattered into the TRANSPOSED weight
//            rows the dV gather reads (window row r: Q = 3..8, window row r+1: Q = 0..2; fp32, pitch 12)
//   stage 3: warp items, two kinds:
//            (a) dV of one (2 x 2 output block, head pair): lane = (head of the pair, channel pair); 5 x 5 patch of dy in
//                registers, 9 broadcast weight rows, 162 FFMA, four 4-byte stores (as the forward)
//            (b) dlogits of one (window, head): dA as a 16 x 16 x 32 mma.sync product whose operands come straight from the
//                two bands through ldmatrix (row addresses do the unfold, rows 9..15 read a zero row); softmax rebuilt in
//                accumulator layout from the raw logits (quad shuffles), row sums by quad shuffles, 2-byte stores
// The mma.sync kernel it replaces (outlook_mma.cu) recomputed a halo window row per CTA for the fold (work factor 1.5 at
// 28 x 28), ran one 896-thread CTA per SM and staged through registers: 189 us at 128 x 28 x 28 x 192 (profiles/r2_kernels.md).
#include "common.cuh"

namespace {

constexpr int HD = 32;
constexpr int WP = 12;                 // weight row pitch in floats (9 used)
constexpr int BT = 256;                // threads per CTA

__device__ __forceinline__ uint32_t smem_u32b(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32b(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128b(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cp_async16b(uint32_t dst, const void* src, int src_bytes) {    // src_bytes 0 -> 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpa(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float qmax(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float qsum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct ObParams {
  const bf16* v;
  const bf16* logits;
  const bf16* dy;
  bf16* dv;
  bf16* dlogits;
  int B, H, W, h, w, heads, lpitch;
  float scale;
  int tcw;          // window columns (= 2 x 2 output blocks) owned per CTA
  int xtiles;       // ceil(w / tcw)
};

// CH > 0: heads (and the padded logits pitch that goes with it) are compile-time constants -- the kernel is issue-bound and
// every index split by a runtime head count costs a ~20-instruction integer division (CH = 0: fully dynamic)
template <int CH>
__global__ void __launch_bounds__(BT) outlook_bwd_fma_kernel(ObParams p) {
  const int heads = CH > 0 ? CH : p.heads;
  const int lpitch = CH > 0 ? (CH * 81 + 7) / 8 * 8 : p.lpitch;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = heads * HD, CP = C + 8;
  const int xt = blockIdx.x % p.xtiles;
  const int r = (blockIdx.x / p.xtiles) % p.h;            // window row r ; output rows 2r, 2r+1
  const int b = blockIdx.x / (p.xtiles * p.h);
  const int jb = xt * p.tcw;                             // first window column of the tile
  const int nblk = min(p.tcw, p.w - jb);                 // owned windows / output blocks
  const int nwin = nblk + 1;                             // window columns jb .. jb + nblk (the last one feeds only odd x)
  const int BW = 2 * nblk + 3;                           // band pixel columns: 2 jb - 1 .. 2 (jb + nblk) + 1
  // shared memory: zero row (64 B) | band v [5][BW][CP] | band dy [5][BW][CP] | weights^T [nwin][heads][9][WP] | raw [2][nwin][lpitch]
  const uint32_t zero_s = smem_u32b(smem_raw);
  const uint32_t band_bytes = (uint32_t)(5 * BW * CP * 2);            // CP * 2 is a multiple of 16
  const uint32_t bv_s = zero_s + 64, bg_s = bv_s + band_bytes;
  float* wts = reinterpret_cast<float*>(smem_raw + 64 + 2 * (size_t)band_bytes);
  const size_t wts_bytes = (size_t)nwin * heads * 9 * WP * sizeof(float);
  bf16* raw = reinterpret_cast<bf16*>(smem_raw + 64 + 2 * (size_t)band_bytes + wts_bytes);

  if (tid < 16) reinterpret_cast<uint32_t*>(smem_raw)[tid] = 0u;
  // ---- stage 1a: both pixel bands (rows 2r-1 .. 2r+3, columns 2jb-1 ..), zeros outside the image; the 16 pad bytes of a
  //      pixel are never read (ldmatrix rows and patch loads stay inside the head's 64 bytes)
  {
    const int vpp = C / 8;                                // 16-byte vectors per pixel
    const int step_bc = BT / vpp, step_cv = BT % vpp;
    int bc = tid / vpp, cv = tid % vpp;
    while (bc < BW) {
      const int xx = 2 * jb - 1 + bc;
      const bool xin = xx >= 0 && xx < p.W;
#pragma unroll
      for (int br = 0; br < 5; ++br) {
        const int yy = 2 * r - 1 + br;
        const bool ok = xin && yy >= 0 && yy < p.H;
        const size_t goff = ok ? (((size_t)b * p.H + yy) * p.W + xx) * C + cv * 8 : 0;
        const uint32_t soff = (uint32_t)(((br * BW + bc) * CP + cv * 8) * 2);
        cp_async16b(bv_s + soff, p.v + goff, ok ? 16 : 0);
        cp_async16b(bg_s + soff, p.dy + goff, ok ? 16 : 0);
      }
      bc += step_bc;
      cv += step_cv;
      if (cv >= vpp) { cv -= vpp; ++bc; }
    }
  }
  // ---- stage 1b: raw logits rows of windows (r, jb ..) and (r+1, jb ..): lpitch * 2 / 16 vectors per window
  {
    const int vpw = lpitch / 8;
    const int step_w = BT / vpw, step_v = BT % vpw;
    int wi = tid / vpw, vv = tid % vpw;                   // wi = slot * nwin + jl
    const uint32_t raw_s = smem_u32b(raw);
    while (wi < 2 * nwin) {
      const int slot = wi >= nwin ? 1 : 0, jl = wi - slot * nwin;
      const int iw = r + slot, jw = jb + jl;
      const bool ok = iw < p.h && jw < p.w;
      const bf16* src = ok ? p.logits + (((size_t)b * p.h + iw) * p.w + jw) * lpitch + vv * 8 : p.logits;
      cp_async16b(raw_s + (uint32_t)((wi * vpw + vv) * 16), src, ok ? 16 : 0);
      wi += step_w;
      vv += step_v;
      if (vv >= vpw) { vv -= vpw; ++wi; }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- stage 2: thread = (slot, window, head, P): softmax row P -> column entries of the transposed weight rows.
  //      Weight row ridx of a (window, head): ridx 0..5 = window row r, Q = 3..8 ; ridx 6..8 = window row r+1, Q = 0..2 ;
  //      row content = A[P = 0..8][Q]
  {
    const int nwh = nwin * heads;
    for (int e = tid; e < 2 * nwh * 9; e += BT) {
      const int P = e % 9, t = e / 9;
      const int slot = t >= nwh ? 1 : 0, wh = t - slot * nwh;
      const int jl = wh / heads, hd = wh - jl * heads;
      const int iw = r + slot, jw = jb + jl;
      float q[9];
      if (iw < p.h && jw < p.w) {
        const bf16* src = raw + (size_t)(slot * nwin + jl) * lpitch + hd * 81 + P * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] = __bfloat162float(src[i]) * p.scale;
        const float m = fmaxf(fmaxf(fmaxf(q[0], q[1]), fmaxf(q[2], q[3])), fmaxf(fmaxf(q[4], q[5]), fmaxf(fmaxf(q[6], q[7]), q[8])));
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] = __expf(q[i] - m);
        const float inv = 1.f / (((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7])) + q[8]);
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] *= inv;
      } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] = 0.f;
      }
      float* base = wts + (size_t)wh * 9 * WP + P;
      if (slot == 0) {
#pragma unroll
        for (int Q = 3; Q < 9; ++Q) base[(Q - 3) * WP] = q[Q];
      } else {
#pragma unroll
        for (int Q = 0; Q < 3; ++Q) base[(6 + Q) * WP] = q[Q];
      }
    }
  }
  __syncthreads();
  // ---- stage 3: warp items.  [0, nb_items): dlogits of (window jl < nblk, head); then dV of (output block, head pair)
  const int npairs = (heads + 1) >> 1;
  const int nb_items = nblk * heads, na_items = nblk * npairs;
  const uint32_t wts_a = smem_u32b(wts);
  const uint32_t row_pitch_b = (uint32_t)(BW * CP * 2), col_pitch_b = (uint32_t)(CP * 2);
  // per-lane ldmatrix geometry of the dA product (loop invariant)
  const int gi = lane >> 2, qd = lane & 3;
  uint32_t offT, offB[2];
  bool zT, zB[2];
  {
    const int mi = lane >> 3, r8 = lane & 7;
    const int idxT = (mi & 1) * 8 + r8;
    zT = idxT >= 9;
    offT = zT ? 0u : (uint32_t)(idxT / 3) * row_pitch_b + (uint32_t)(idxT % 3) * col_pitch_b + (uint32_t)((mi >> 1) * 16);
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      const int idxB = nb * 8 + r8;
      zB[nb] = idxB >= 9;
      offB[nb] = zB[nb] ? 0u : (uint32_t)(idxB / 3) * row_pitch_b + (uint32_t)(idxB % 3) * col_pitch_b + (uint32_t)(mi * 16);
    }
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  const int npad = lpitch - heads * 81;
  for (int item = warp; item < nb_items + na_items; item += BT / 32) {
    if (item < nb_items) {
      // ================= (b) dlogits of window (r, jb + jl), head hd =================
      const int jl = item / heads, hd = item - jl * heads;
      // softmax in accumulator layout: (P = gi: Q = 2q, 2q+1, 8) and, for gi == 0, (P = 8: Q = 2q, 2q+1, 8)
      const bf16* L = raw + (size_t)jl * lpitch + hd * 81;
      const float NEG = -INFINITY;
      float e0, e1, e2, f0, f1, f2;
      {
        const bf16* row = L + gi * 9 + 2 * qd;
        e0 = __bfloat162float(row[0]) * sl2;
        e1 = __bfloat162float(row[1]) * sl2;
        e2 = (qd == 0) ? __bfloat162float(row[8 - 2 * qd]) * sl2 : NEG;
        f0 = (gi == 0) ? __bfloat162float(L[72 + 2 * qd]) * sl2 : NEG;
        f1 = (gi == 0) ? __bfloat162float(L[72 + 2 * qd + 1]) * sl2 : NEG;
        f2 = (gi == 0 && qd == 0) ? __bfloat162float(L[80]) * sl2 : NEG;
      }
      const float m0 = qmax(fmaxf(fmaxf(e0, e1), e2));
      float m1 = qmax(fmaxf(fmaxf(f0, f1), f2));
      m1 = (gi == 0) ? m1 : 0.f;                              // keeps exp2(-inf - m1) = 0 without NaNs
      e0 = ex2a(e0 - m0); e1 = ex2a(e1 - m0); e2 = ex2a(e2 - m0);
      f0 = ex2a(f0 - m1); f1 = ex2a(f1 - m1); f2 = ex2a(f2 - m1);
      const float inv0 = rcpa(qsum(e0 + e1 + e2));
      const float s1 = qsum(f0 + f1 + f2);
      const float inv1 = (gi == 0) ? rcpa(s1) : 0.f;
      // probabilities: a00,a01 = (P=gi, Q=2q,2q+1) ; a02,a03 = (P=gi+8, same Q) ; a10 = (P=gi, Q=8) ; a12 = (P=gi+8, Q=8)
      const float a00 = e0 * inv0, a01 = e1 * inv0, a02 = f0 * inv1, a03 = f1 * inv1, a10 = e2 * inv0, a12 = f2 * inv1;
      // dA[P][Q] = sum_c dy[pix P][c] v[pix Q][c] : A operand = dy rows, B operand ("col") = v rows
      float da[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) da[nb][0] = da[nb][1] = da[nb][2] = da[nb][3] = 0.f;
      {
        const uint32_t uoff = (uint32_t)(2 * jl) * col_pitch_b + (uint32_t)(hd * HD * 2);
        uint32_t a0[4], a1[4], bq[4];
        const uint32_t ga = zT ? zero_s : bg_s + uoff + offT;
        ldsm4(a0, ga);
        ldsm4(a1, zT ? zero_s : ga + 32);
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          ldsm4(bq, zB[nb] ? zero_s : bv_s + uoff + offB[nb]);
          mma_bf16(da[nb], a0, bq[0], bq[1]);
          mma_bf16(da[nb], a1, bq[2], bq[3]);
        }
      }
      const float r0 = qsum(a00 * da[0][0] + a01 * da[0][1] + a10 * da[1][0]);
      const float r1 = qsum(a02 * da[0][2] + a03 * da[0][3] + a12 * da[1][2]);
      const int jw = jb + jl;
      bf16* dl = p.dlogits + (((size_t)b * p.h + r) * p.w + jw) * lpitch + hd * 81;
      bf16* drow = dl + gi * 9 + 2 * qd;
      drow[0] = __float2bfloat16_rn(p.scale * a00 * (da[0][0] - r0));
      drow[1] = __float2bfloat16_rn(p.scale * a01 * (da[0][1] - r0));
      if (qd == 0) drow[8] = __float2bfloat16_rn(p.scale * a10 * (da[1][0] - r0));
      if (gi == 0) {
        dl[72 + 2 * qd] = __float2bfloat16_rn(p.scale * a02 * (da[0][2] - r1));
        dl[72 + 2 * qd + 1] = __float2bfloat16_rn(p.scale * a03 * (da[0][3] - r1));
        if (qd == 0) dl[80] = __float2bfloat16_rn(p.scale * a12 * (da[1][2] - r1));
      }
      if (hd == 0 && lane >= 9 && lane - 9 < npad) dl[heads * 81 + (lane - 9)] = __float2bfloat16_rn(0.f);   // row padding
      continue;
    }
    // ================= (a) dV of output block jl, head pair hp =================
    const int it = item - nb_items;
    const int hp = it % npairs, jl = it / npairs;
    const int half = lane >> 4, cp = lane & 15;
    const int hd = 2 * hp + half;
    const bool head_ok = hd < heads;
    const int hdc = head_ok ? hd : heads - 1;           // clamp: lanes of a missing head compute on valid memory, never store
    float v0[25], v1[25];
    {
      const uint32_t base = bg_s + (uint32_t)(2 * jl) * col_pitch_b + (uint32_t)((hdc * HD + 2 * cp) * 2);
#pragma unroll
      for (int pr = 0; pr < 5; ++pr) {
        const uint32_t rb = base + pr * row_pitch_b;
#pragma unroll
        for (int pc = 0; pc < 5; ++pc) {
          const uint32_t u = lds32b(rb + pc * col_pitch_b);
          v0[pr * 5 + pc] = __uint_as_float(u << 16);
          v1[pr * 5 + pc] = __uint_as_float(u & 0xFFFF0000u);
        }
      }
    }
    float acc[4][2];
#pragma unroll
    for (int o = 0; o < 4; ++o) { acc[o][0] = 0.f; acc[o][1] = 0.f; }
    // transposed weight rows: (window column offset dj, row index, patch origin (oy, ox), output pixel o = dy * 2 + dx)
    //   window (r, j)    : Q=(1,1)->o0 (ridx 1), (1,2)->o1 (ridx 2), (2,1)->o2 (ridx 4), (2,2)->o3 (ridx 5); origin (0,0)
    //   window (r, j+1)  : Q=(1,0)->o1 (ridx 0), (2,0)->o3 (ridx 3);                                        origin (0,2)
    //   window (r+1, j)  : Q=(0,1)->o2 (ridx 7), (0,2)->o3 (ridx 8);                                        origin (2,0)
    //   window (r+1, j+1): Q=(0,0)->o3 (ridx 6);                                                            origin (2,2)
    const uint32_t wbase = wts_a + (uint32_t)((((jl * heads + hdc) * 9) * WP) * 4);
    const uint32_t wnext = (uint32_t)(heads * 9 * WP * 4);          // next window column
    auto apply = [&](int dj, int ridx, int oy, int ox, int o) {
      const uint32_t ra = wbase + dj * wnext + (uint32_t)(ridx * WP * 4);
      const float4 wa = lds128b(ra), wb = lds128b(ra + 16);
      const float w8 = __uint_as_float(lds32b(ra + 32));
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
          *reinterpret_cast<__nv_bfloat162*>(p.dv + (((size_t)b * p.H + yy) * p.W + xx) * C + hd * HD + 2 * cp) = hv;
        }
      }
    }
  }
}

}  // namespace

// returns APB_ERR_UNSUPPORTED when the tile does not fit shared memory or the logits rows are not 16-byte vectors
int apb_outlook_bwd_fma(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W, int heads,
                        float scale, int lpitch, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  ObParams p;
  p.v = (const bf16*)v; p.logits = (const bf16*)logits; p.dy = (const bf16*)dy; p.dv = (bf16*)dv; p.dlogits = (bf16*)dlogits;
  p.B = B; p.H = H; p.W = W; p.h = (H + 1) / 2; p.w = (W + 1) / 2; p.heads = heads; p.lpitch = lpitch; p.scale = scale;
  const int C = heads * HD;
  if (lpitch % 8 != 0 || lpitch < heads * 81 || ((uintptr_t)logits & 15) != 0) return APB_ERR_UNSUPPORTED;
  auto smem_for = [&](int tcw) {
    return (size_t)64 + (size_t)2 * 5 * (2 * tcw + 3) * (C + 8) * 2 + (size_t)(tcw + 1) * heads * 9 * WP * 4 +
           (size_t)2 * (tcw + 1) * lpitch * 2;
  };
  // widest tile that leaves room for three CTAs per SM; else two; else the widest that fits at all
  int tcw = p.w;
  const size_t budgets[3] = {75 * 1024, 113 * 1024, 227 * 1024};
  int bi = 0;
  for (; bi < 3; ++bi) {
    tcw = p.w;
    while (tcw > 1 && smem_for(tcw) > budgets[bi]) --tcw;
    if (smem_for(tcw) <= budgets[bi] && (tcw >= 4 || tcw == p.w || bi == 2)) break;
  }
  if (bi == 3 || smem_for(tcw) > 227 * 1024) return APB_ERR_UNSUPPORTED;
  p.xtiles = ceil_div(p.w, tcw);
  p.tcw = ceil_div(p.w, p.xtiles);
  const size_t smem = smem_for(p.tcw);
  const long long grid = (long long)B * p.h * p.xtiles;
  const int ch = (heads == 6 || heads == 8 || heads == 12) && lpitch == (heads * 81 + 7) / 8 * 8 ? heads : 0;
#define OL_LAUNCH(CH_)                                                                                                      \
  do {                                                                                                                      \
    static size_t attr = 0;                                                                                                 \
    if (smem > attr) {                                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(outlook_bwd_fma_kernel<CH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) { apb_set_error("outlook_bwd_fma_kernel: smem attr: %s", cudaGetErrorString(e)); return (int)e; }  \
      attr = smem;                                                                                                          \
    }                                                                                                                       \
    outlook_bwd_fma_kernel<CH_><<<(unsigned)grid, BT, smem, st>>>(p);                                                 \
  } while (0)
  if (ch == 6) OL_LAUNCH(6);
  else if (ch == 8) OL_LAUNCH(8);
  else if (ch == 12) OL_LAUNCH(12);
  else OL_LAUNCH(0);
#undef OL_LAUNCH
  APB_LAUNCH_CHECK("outlook_bwd_fma");
  return 0;
}
