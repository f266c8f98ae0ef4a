// Multi-head self-attention FORWARD and BACKWARD on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head_dim 32,
// N <= 224 tokens (every VOLO stage-2 grid of the progressive schedule: N = 64 ... 196).
//
//   reference: attn = softmax(q k^T * scale); out = attn @ v            (models/volo.py:193-197) and its autograd backward
//
// Common structure (both kernels): PERSISTENT CTAs (one per SM) of 18 warps
//   warps 0-15      : four softmax warpgroups; warpgroups 2s and 2s+1 share TMEM accumulator slot s: thread t of either
//                     owns TMEM lane t = one query row, and the two split the slot's COLUMNS (a warp may only touch the
//                     lane quarter warp % 4, so 16 warps = 2 slots x 2 column halves x 4 lane quarters)
//   warp 16         : TMA producer: per (batch, head) the q / k / v (/ dO) row blocks of that head are fetched straight
//                     from the packed qkv tensor as [rows][64 B] tiles (2-D tensor maps over [B*N, 3*heads*32], box
//                     {32 channels, rows}, 64-byte swizzle) into a 2-stage ring -> the next head streams in while this
//                     one computes; no thread-staged or transposed copies (V / dO / Q / K are consumed as MN-major
//                     operands where the product needs them transposed)
//   warp 17         : TMEM allocator + the single thread that issues EVERY tcgen05.mma (so all accumulations into shared
//                     accumulators are ordered) and signals completion through tcgen05.commit -> mbarrier.  It POLLS the
//                     two slots (mbarrier.test_wait) and serves whichever is ready, so the slots drift apart freely
// Keys fit one accumulator tile, so the softmax needs no online rescaling: each thread reads ITS row with tcgen05.ld.
//
// FORWARD  (item = (head, 128-row query tile); slot s takes tile (head + s) & 1 so that full and partial tiles alternate;
//   TMEM slot = S[Npad] + O[32]):
//   S = Q K^T (M=128, N=Npad, K=32)  ->  pass 1 row max (halves exchange through shared memory)  ->  pass 2
//   P = exp2(.) written back over the S columns the same thread has consumed, as packed bf16 (tcgen05.st)  ->  O = P V
//   with P as the TMEM A operand (one k-step per 16 keys, each with its own TMEM column address) and V MN-major from
//   shared memory  ->  out = O / l, lse.  The S product of a slot's NEXT item is issued right behind its P V product.
// BACKWARD (item = (query tile, key block of <= 80 keys), items alternate between the slots; TMEM: 2 x (S[80] + dP[80]) +
//   dQ[2][32] + 2 x (dK[32] + dV[32]) = 512 columns):
//   S = Q K_b^T, dP = dO V_b^T  ->  P = exp2(S*c - lse), dS = P (dP - D)  (registers; D = rowsum(dO o O) computed
//   in-kernel) -> P, dS as bf16 into shared memory (64B-swizzled 32-key blocks)  ->  dV_b += P^T dO, dK_b += dS^T Q
//   (A = the SAME shared tiles read MN-major; both query tiles accumulate into one TMEM accumulator), dQ += dS K_b
//   (A K-major).  dK_b / dV_b are double-buffered in TMEM and stored one key block LATE, so no warpgroup ever waits for
//   the products of the block it has just finished.  One kernel instead of the row-dot + dQ + dK/dV mma.sync kernels;
//   no atomics, fixed accumulation order.
#include "gemm_tc_common.cuh"

namespace {

constexpr int HD = 32;            // head dim
constexpr int ROWB = HD * 2;      // bytes per tile row (one head of one token)
constexpr int QT = 128;           // query rows per tile (UMMA M)
constexpr int NTHREADS = 576;     // 4 warpgroups + producer warp + MMA warp
constexpr int W_TMA = 16, W_MMA = 17;

__device__ __forceinline__ uint64_t desc64(uint32_t saddr, uint32_t lbo_bytes) {
  // 64-byte rows, 64B swizzle: 8-row groups are 512 B apart (SBO); LBO = distance between 32-element MN blocks (MN-major)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}
// byte offset of 16-byte chunk c (0..3) of row r inside a [rows][64 B] tile with the 64B swizzle
__device__ __forceinline__ uint32_t sw64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// lane 0's view of the barrier, broadcast: keeps polling loops warp-uniform
__device__ __forceinline__ bool mbar_test_u(uint64_t* bar, uint32_t parity) {
  return __shfl_sync(0xffffffffu, mbar_test(bar, parity) ? 1u : 0u, 0) != 0;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// keep a tcgen05.ld result from being consumed before wait::ld (see DESIGN.md "hardware lessons")
__device__ __forceinline__ void pin16(uint32_t (&r)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) asm volatile("" : "+r"(r[j])::"memory");
}

constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10);   // D = f32, A = B = bf16
__device__ __forceinline__ uint32_t idesc_mn(int M, int N, int a_mn, int b_mn) {
  return IDESC_BASE | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct AttParams {
  const bf16* qkv;     // [B, N, 3, heads, 32]
  const bf16* o;       // bwd: forward output [B, N, heads*32]
  const bf16* dout;    // bwd
  bf16* out;           // fwd: [B, N, heads*32]; bwd: dqkv
  float* lse;          // [B, heads, N]
  const float* rowdot; // bwd: [B, heads, N] D = sum_d dO * O (written by mhsa_rowdot_tc_kernel)
  int B, N, heads;
  int Npad;            // keys padded to a multiple of 16
  int ntq;             // query tiles of 128 rows
  int units;           // B * heads
  float scale;
  long long* trace;    // optional (diagnostic): [18 warps][TRACE_MAX] {event id << 48 | clock} written by CTA 0
};

constexpr int TRACE_MAX = 1024;
struct Tracer {
  long long* buf;
  int n;
  __device__ __forceinline__ void init(long long* base, int warp) {
    buf = (base != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) ? base + (size_t)warp * TRACE_MAX : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void ev(int id) {
    if (buf != nullptr && n < TRACE_MAX) buf[n++] = ((long long)id << 48) | (clock64() & 0xFFFFFFFFFFFFLL);
  }
};

// =====================================================================================================================
// forward
// =====================================================================================================================
constexpr int F_NST = 3;          // TMA ring stages (heads in flight)
// shared memory: stage = [Q: ntq*128 rows][K: Npad rows][V: Npad rows] x 64 B; then the exchange arrays and barriers
struct FwdBars {
  float xm[2][2][QT];             // [slot][column half][row]: partial row maxima
  float xl[2][2][QT];             // partial row sums
  uint64_t full[F_NST], empty[F_NST], s_ready[2], p_ready[2], o_ready[2];
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(NTHREADS, 1) mhsa_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                  const __grid_constant__ CUtensorMap map_kv, AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int qrows = p.ntq * QT;
  const uint32_t q_bytes = (uint32_t)qrows * ROWB, kv_bytes = (uint32_t)p.Npad * ROWB;
  const uint32_t stage_bytes = q_bytes + 2 * kv_bytes;
  FwdBars* bars = reinterpret_cast<FwdBars*>(smem + F_NST * stage_bytes);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int C = p.heads * HD;
  const int my_units = ((int)blockIdx.x < p.units) ? (p.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // item i of slot s:  two query tiles per head -> head i, tile (i + s) & 1;  one tile -> head 2 i + s, tile 0
  const int two = (p.ntq == 2);
  const int nch = p.Npad >> 4;                       // 16-key chunks
  const int h0 = (nch + 1) >> 1;                     // chunks [0, h0) belong to column half 0, [h0, nch) to half 1
  Tracer tr;
  tr.init(p.trace, warp);

  if (tid == 0) {
    for (int s = 0; s < F_NST; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->s_ready[s], 1); mbar_init(&bars->p_ready[s], 256); mbar_init(&bars->o_ready[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  constexpr uint32_t SLOT = 256, O_COL = 224;

  if (warp == W_TMA) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    {
      const bool leader = elect_one();
      for (int ul = 0; ul < my_units; ++ul) {
        const int st = ul % F_NST;
        mbar_wait(&bars->empty[st], ((ul / F_NST) & 1) ^ 1);
        const int unit = blockIdx.x + ul * gridDim.x;
        const int b = unit / p.heads, hd = unit % p.heads;
        uint8_t* sq = smem + st * stage_bytes;
        if (leader) {
          mbar_expect_tx(&bars->full[st], stage_bytes);
          tma_load_2d(sq, &map_q, &bars->full[st], hd * HD, b * p.N);
          tma_load_2d(sq + q_bytes, &map_kv, &bars->full[st], C + hd * HD, b * p.N);
          tma_load_2d(sq + q_bytes + kv_bytes, &map_kv, &bars->full[st], 2 * C + hd * HD, b * p.N);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer: polls the two slots =====================
    // One thread feeds the tensor pipe for the whole SM, so its instruction stream is kept short: item cursors advance
    // incrementally (no divisions) and the shared-memory descriptors are built once per item and stepped by constants.
    // The WHOLE warp runs this loop with warp-uniform state; only the tcgen05 instructions sit under the elected lane.
    {
      const bool leader = elect_one();
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
      const uint32_t idS = idesc_mn(QT, p.Npad, 0, 0), idO = idesc_mn(QT, HD, 0, 1);
      const uint32_t smem_a = smem_u32(smem);
      int n_items[2], s_issued[2] = {0, 0}, pv_issued[2] = {0, 0};
      n_items[0] = two ? my_units : (my_units + 1) / 2;
      n_items[1] = two ? my_units : my_units / 2;
      int pv_of_stage[F_NST];
#pragma unroll
      for (int i = 0; i < F_NST; ++i) pv_of_stage[i] = 0;
      // per slot: head (CTA-local), ring stage and ring phase of the next S item / next P V item
      int s_ul[2] = {0, two ? 0 : 1}, s_st[2], s_ph[2] = {0, 0}, v_st[2];
      const int ul_step = two ? 1 : 2;
#pragma unroll
      for (int s = 0; s < 2; ++s) { s_st[s] = s_ul[s] % F_NST; s_ph[s] = (s_ul[s] / F_NST) & 1; v_st[s] = s_st[s]; }
      while (pv_issued[0] < n_items[0] || pv_issued[1] < n_items[1]) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          // ---- S of the slot's next item: the slot is free once the P V product of its previous item has been issued
          if (s_issued[s] < n_items[s] && s_issued[s] == pv_issued[s]) {
            if (mbar_test_u(&bars->full[s_st[s]], (uint32_t)s_ph[s])) {
              fence_after();
              const int qt = two ? ((s_issued[s] + s) & 1) : 0;
              const uint32_t base = smem_a + (uint32_t)s_st[s] * stage_bytes;
              const uint64_t qd = desc64(base + (uint32_t)qt * QT * ROWB, 16), kd = desc64(base + q_bytes, 16);
              tr.ev(20 + s);
              if (leader) {
                umma_bf16(tmem_base + s * SLOT, qd, kd, idS, 0u);
                umma_bf16(tmem_base + s * SLOT, qd + 2, kd + 2, idS, 1u);       // + 32 B = next 16 channels
                umma_commit(&bars->s_ready[s]);
              }
              __syncwarp();
              ++s_issued[s];
              for (int k = 0; k < ul_step; ++k) {                             // advance the S cursor
                ++s_ul[s];
                if (++s_st[s] == F_NST) { s_st[s] = 0; s_ph[s] ^= 1; }
              }
            }
          }
          // ---- O = P V once both column halves have stored their P
          if (pv_issued[s] < s_issued[s] && mbar_test_u(&bars->p_ready[s], pv_issued[s] & 1)) {
            fence_after();
            const int st = v_st[s];
            const uint64_t vd0 = desc64(smem_a + (uint32_t)st * stage_bytes + q_bytes + kv_bytes, 16);
            const uint32_t to = tmem_base + s * SLOT + O_COL;
            const uint32_t pa0 = tmem_base + s * SLOT, pa1 = pa0 + 16 * h0 - 8 * h0;
            const bool last_of_head = ++pv_of_stage[st] == p.ntq;   // every product reading this head's tiles has been issued
            if (last_of_head) pv_of_stage[st] = 0;
            tr.ev(30 + s);
            if (leader) {
              // P of chunk kk: half 0 packs chunk c at columns 8c, half 1 at 16 h0 + 8 (c - h0) (inside its own S columns)
              umma_ts(to, pa0, vd0, idO, 0u);
              for (int kk = 1; kk < h0; ++kk) umma_ts(to, pa0 + 8 * kk, vd0 + (uint64_t)(kk * ((16 * ROWB) >> 4)), idO, 1u);
              for (int kk = h0; kk < nch; ++kk) umma_ts(to, pa1 + 8 * kk, vd0 + (uint64_t)(kk * ((16 * ROWB) >> 4)), idO, 1u);
              umma_commit(&bars->o_ready[s]);
              if (last_of_head) umma_commit(&bars->empty[st]);
            }
            __syncwarp();
            tr.ev(40 + s);
            ++pv_issued[s];
            for (int k = 0; k < ul_step; ++k)
              if (++v_st[s] == F_NST) v_st[s] = 0;
          }
        }
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int g = warp >> 2, s = g >> 1, hf = g & 1;  // warpgroup, slot, column half
    const int t = tid & 127;                          // row inside the query tile = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + s * SLOT;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int N = p.N;
    const int c_lo = hf ? h0 : 0, c_hi = hf ? nch : h0;          // this half's chunks
    const int n_items = two ? my_units : (s == 0 ? (my_units + 1) / 2 : my_units / 2);
    for (int i = 0; i < n_items; ++i) {
      const int ul = two ? i : 2 * i + s, qt = two ? ((i + s) & 1) : 0;
      const int unit = blockIdx.x + ul * gridDim.x;
      const int b = unit / p.heads, hd = unit % p.heads;
      const int row = qt * QT + t;
      const bool warp_live = (qt * QT + (warp & 3) * 32) < N;      // warp-uniform: any valid row in this warp
      tr.ev(1);
      mbar_wait(&bars->s_ready[s], i & 1);
      fence_after();
      tr.ev(2);
      float m = -INFINITY, l = 0.f;
      uint32_t ra[16], rb[16];
      if (warp_live && c_lo < c_hi) {
        // ---- pass 1: row maximum over this half's valid keys (chunk c + 1 in flight while chunk c is reduced)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto max_chunk = [&](const uint32_t (&rc)[16], int c) {
          if ((c + 1) * 16 <= N) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(rc[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c * 16 + j < N) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(rc[j]));
          }
        };
        tmem_ld16(lane_addr + (uint32_t)(c_lo * 16), ra);
        for (int c = c_lo; c < c_hi; c += 2) {
          wait_ld();
          pin16(ra);
          if (c + 1 < c_hi) tmem_ld16(lane_addr + (uint32_t)((c + 1) * 16), rb);
          max_chunk(ra, c);
          if (c + 1 < c_hi) {
            wait_ld();
            pin16(rb);
            if (c + 2 < c_hi) tmem_ld16(lane_addr + (uint32_t)((c + 2) * 16), ra);
            max_chunk(rb, c + 1);
          }
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      bars->xm[s][hf][t] = m;
      tr.ev(3);
      bar_sync(1 + s, 256);                            // the two column halves of this slot
      tr.ev(4);
      m = fmaxf(m, bars->xm[s][hf ^ 1][t]);
      if (warp_live && c_lo < c_hi) {
        // ---- pass 2: P = exp2((S - m) * scale * log2e) -> packed bf16 over S columns this thread has consumed
        const float msc = m * sl2;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t pbase = lane_addr + (uint32_t)(hf ? 16 * h0 : 0);
        auto exp_chunk = [&](const uint32_t (&rc)[16], int c) {
          uint32_t pk[8];
          if ((c + 1) * 16 <= N) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float p0 = ex2f(fmaf(__uint_as_float(rc[j]), sl2, -msc)), p1 = ex2f(fmaf(__uint_as_float(rc[j + 1]), sl2, -msc));
              l4[(j >> 1) & 3] += p0 + p1;
              pk[j >> 1] = pack2(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float p0 = ex2f(fmaf(__uint_as_float(rc[j]), sl2, -msc)), p1 = ex2f(fmaf(__uint_as_float(rc[j + 1]), sl2, -msc));
              if (c * 16 + j >= N) p0 = 0.f;
              if (c * 16 + j + 1 >= N) p1 = 0.f;
              l4[(j >> 1) & 3] += p0 + p1;
              pk[j >> 1] = pack2(p0, p1);
            }
          }
          tmem_st8(pbase + (uint32_t)((c - c_lo) * 8), pk);   // 8 columns at or below the 16 just consumed
        };
        tmem_ld16(lane_addr + (uint32_t)(c_lo * 16), ra);
        for (int c = c_lo; c < c_hi; c += 2) {
          wait_ld();
          pin16(ra);
          if (c + 1 < c_hi) tmem_ld16(lane_addr + (uint32_t)((c + 1) * 16), rb);
          exp_chunk(ra, c);
          if (c + 1 < c_hi) {
            wait_ld();
            pin16(rb);
            if (c + 2 < c_hi) tmem_ld16(lane_addr + (uint32_t)((c + 2) * 16), ra);
            exp_chunk(rb, c + 1);
          }
        }
        l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        wait_st();
      }
      bars->xl[s][hf][t] = l;
      fence_before();
      tr.ev(5);
      mbar_arrive(&bars->p_ready[s]);
      mbar_wait(&bars->o_ready[s], i & 1);
      fence_after();
      tr.ev(6);
      if (warp_live) {
        // ---- epilogue: this half normalises and stores 16 of the 32 output channels
        tmem_ld16(lane_addr + O_COL + (uint32_t)(hf * 16), ra);
        wait_ld();
        pin16(ra);
        l += bars->xl[s][hf ^ 1][t];
        if (row < N) {
          const float inv = 1.f / l;
          bf16* orow = p.out + ((size_t)b * N + row) * C + (size_t)hd * HD + hf * 16;
#pragma unroll
          for (int c4 = 0; c4 < 2; ++c4) {
            uint4 pkv;
            pkv.x = pack2(__uint_as_float(ra[c4 * 8 + 0]) * inv, __uint_as_float(ra[c4 * 8 + 1]) * inv);
            pkv.y = pack2(__uint_as_float(ra[c4 * 8 + 2]) * inv, __uint_as_float(ra[c4 * 8 + 3]) * inv);
            pkv.z = pack2(__uint_as_float(ra[c4 * 8 + 4]) * inv, __uint_as_float(ra[c4 * 8 + 5]) * inv);
            pkv.w = pack2(__uint_as_float(ra[c4 * 8 + 6]) * inv, __uint_as_float(ra[c4 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c4 * 8) = pkv;
          }
          if (hf == 0) p.lse[((size_t)b * p.heads + hd) * N + row] = m * p.scale + logf(l);
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
constexpr int B_NST = 2;                  // TMA ring stages (heads in flight)
constexpr int KB = 80;                    // keys per block (S and dP accumulators: 2 x 80 columns per slot)
constexpr uint32_t PBUF = 3 * QT * ROWB;  // one P or dS buffer: 128 query rows x (up to) 96 keys as three 32-key blocks of [128][64 B]
//   (the M = 128 MN-major products read a 4th block = whatever follows the buffer in shared memory; it only feeds
//    accumulator rows 96..127, which are never stored)

struct BwdBars {
  uint64_t full[B_NST], empty[B_NST];
  uint64_t sdp_ready[2], pds_ready[2], mma_done[2];     // per slot
  uint64_t dq_ready[2], dq_free[2];                     // per query tile
  uint64_t dkv_ready[2], dkv_free[2];                   // per dK / dV accumulator buffer
  uint32_t tmem_slot;
};

// TMEM columns: slot s: S [160 s, +80), dP [160 s + 80, +80); dQ of query tile q at 320 + 32 q; buffer j: dK 384 + 64 j, dV + 32
constexpr uint32_t B_SLOT = 2 * KB;
constexpr uint32_t B_DQ = 2 * B_SLOT;
constexpr uint32_t B_DKV = B_DQ + 64;

__global__ void __launch_bounds__(NTHREADS, 1) mhsa_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                  const __grid_constant__ CUtensorMap map_kv,
                                                                  const __grid_constant__ CUtensorMap map_do, AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // layout: [P0][dS0][P1][dS1] (4 x PBUF) | stage 0 | stage 1 | barriers ; stage = [Q][dO] (ntq*128 rows) [K][V] (Npad rows)
  const int qrows = p.ntq * QT;
  const uint32_t q_bytes = (uint32_t)qrows * ROWB, kv_bytes = (uint32_t)p.Npad * ROWB;
  const uint32_t stage_bytes = 2 * q_bytes + 2 * kv_bytes;
  uint8_t* stages = smem + 4 * PBUF;
  BwdBars* bars = reinterpret_cast<BwdBars*>(stages + B_NST * stage_bytes);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int C = p.heads * HD, N = p.N;
  const int my_units = ((int)blockIdx.x < p.units) ? (p.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nkb = (p.Npad + KB - 1) / KB;
  const int ntq = p.ntq;
  Tracer tr;
  tr.init(p.trace, warp);
  // items of a head in issue order: e = kb * ntq + qt.  Slot of item e of CTA-local head ul: two query tiles ->
  // qt ^ (ul & 1) (a slot alternates between the full and the partial tile); one tile -> kb & 1 (slots alternate key blocks)

  if (tid == 0) {
    for (int s = 0; s < B_NST; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->sdp_ready[s], 1); mbar_init(&bars->pds_ready[s], 256); mbar_init(&bars->mma_done[s], 1);
      mbar_init(&bars->dq_ready[s], 1); mbar_init(&bars->dq_free[s], 256);
      mbar_init(&bars->dkv_ready[s], 1); mbar_init(&bars->dkv_free[s], 512);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == W_TMA) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    {
      const bool leader = elect_one();
      for (int ul = 0; ul < my_units; ++ul) {
        const int st = ul % B_NST;
        mbar_wait(&bars->empty[st], ((ul / B_NST) & 1) ^ 1);
        const int unit = blockIdx.x + ul * gridDim.x;
        const int b = unit / p.heads, hd = unit % p.heads;
        uint8_t* sq = stages + st * stage_bytes;
        if (leader) {
          mbar_expect_tx(&bars->full[st], stage_bytes);
          tma_load_2d(sq, &map_q, &bars->full[st], hd * HD, b * N);
          tma_load_2d(sq + q_bytes, &map_do, &bars->full[st], hd * HD, b * N);
          tma_load_2d(sq + 2 * q_bytes, &map_kv, &bars->full[st], C + hd * HD, b * N);
          tma_load_2d(sq + 2 * q_bytes + kv_bytes, &map_kv, &bars->full[st], 2 * C + hd * HD, b * N);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer: polls the two slots =====================
    // One thread feeds the tensor pipe for the whole SM: item cursors advance incrementally (no divisions), descriptors
    // are built once per item and stepped by constants, and at every hand-over the NEXT item's S / dP products go out
    // before the current item's dV / dK / dQ products, so the slot's warpgroups resume while those still run.
    // The WHOLE warp runs this loop with warp-uniform state; only the tcgen05 instructions sit under the elected lane.
    {
      const bool leader = elect_one();
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
      const int items_per_unit = ntq * nkb;
      const uint32_t smem_a = smem_u32(smem), stages_a = smem_u32(stages);
      const uint32_t idT = idesc_mn(QT, HD, 1, 1);        // dV / dK: A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
      const uint32_t idQ = idesc_mn(QT, HD, 0, 1);        // dQ: A = dS (K-major), B = K (MN-major)
      struct Cur { int ul, kb, qt, st, ph; bool valid; };
      auto cur_init = [&](Cur& c, int s) {
        c.ul = 0; c.st = 0; c.ph = 0;
        c.kb = (ntq == 2) ? 0 : s;
        c.qt = (ntq == 2) ? s : 0;
        c.valid = my_units > 0 && c.kb < nkb;
      };
      auto cur_next = [&](Cur& c, int s) {
        c.kb += (ntq == 2) ? 1 : 2;
        if (c.kb >= nkb) {
          c.kb = (ntq == 2) ? 0 : s;
          ++c.ul;
          if (++c.st == B_NST) { c.st = 0; c.ph ^= 1; }
          if (ntq == 2) c.qt ^= 1;
        }
        c.valid = c.ul < my_units && c.kb < nkb;
      };
      Cur sd[2], mm[2];                                   // next S / dP item, next products item, per slot
      uint32_t n_mma[2] = {0, 0}, n_sdp[2] = {0, 0};      // products / S-dP pairs issued per slot (n_mma: parity source of pds_ready)
      bool sdp_pending[2], prod_ready[2] = {false, false};
      for (int s = 0; s < 2; ++s) { cur_init(sd[s], s); mm[s] = sd[s]; sdp_pending[s] = sd[s].valid; }
      int contrib_of[4] = {0, 0, 0, 0};                   // per (global key block & 3): contributions issued so far
      int dq_unit[2] = {-1, -1}, dq_cnt[2] = {0, 0};      // per query tile: head being accumulated, contributions issued
      int done_in_stage[B_NST];
#pragma unroll
      for (int i = 0; i < B_NST; ++i) done_in_stage[i] = 0;
      const uint64_t pd_[2] = {desc64(smem_a, QT * ROWB), desc64(smem_a + 2 * PBUF, QT * ROWB)};                  // P^T (MN-major A)
      const uint64_t dsd_[2] = {desc64(smem_a + PBUF, QT * ROWB), desc64(smem_a + 3 * PBUF, QT * ROWB)};          // dS^T (MN-major A)
      auto issue_sdp = [&](int s) -> bool {
        const Cur& c = sd[s];
        if (!mbar_test_u(&bars->full[c.st], (uint32_t)c.ph)) return false;
        fence_after();
        tr.ev(20 + s);
        const uint32_t sq = stages_a + (uint32_t)c.st * stage_bytes, sdo = sq + q_bytes, sk = sdo + q_bytes, sv = sk + kv_bytes;
        const int kw = min(KB, p.Npad - c.kb * KB);
        const uint32_t id = idesc_mn(QT, kw, 0, 0);
        const uint32_t ts = tmem_base + s * B_SLOT;
        const uint64_t qd = desc64(sq + (uint32_t)c.qt * QT * ROWB, 16), kd = desc64(sk + (uint32_t)c.kb * KB * ROWB, 16);
        const uint64_t od = desc64(sdo + (uint32_t)c.qt * QT * ROWB, 16), vd = desc64(sv + (uint32_t)c.kb * KB * ROWB, 16);
        if (leader) {
          umma_bf16(ts, qd, kd, id, 0u);
          umma_bf16(ts, qd + 2, kd + 2, id, 1u);
          umma_bf16(ts + KB, od, vd, id, 0u);
          umma_bf16(ts + KB, od + 2, vd + 2, id, 1u);
          umma_commit(&bars->sdp_ready[s]);
        }
        __syncwarp();
        cur_next(sd[s], s);
        sdp_pending[s] = sd[s].valid;
        return true;
      };
      int remaining = 0;
      for (int s = 0; s < 2; ++s) {
        Cur c; cur_init(c, s);
        while (c.valid) { ++remaining; cur_next(c, s); }
      }
      while (remaining > 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          // has the slot stored P / dS of its current item?  (never before that item's S / dP were issued)
          if (mm[s].valid && !prod_ready[s] && n_sdp[s] > n_mma[s] && mbar_test_u(&bars->pds_ready[s], n_mma[s] & 1)) prod_ready[s] = true;
          // S / dP of the slot's next item: allowed when the slot is idle, or at the hand-over (the warpgroups drained the
          // S / dP accumulators before they arrived on pds_ready) -- then it goes out BEFORE the current item's products
          if (sdp_pending[s] && (n_sdp[s] == n_mma[s] || (n_sdp[s] == n_mma[s] + 1 && prod_ready[s]))) {
            if (issue_sdp(s)) ++n_sdp[s];
          }
          if (!prod_ready[s]) continue;
          const Cur c = mm[s];
          const int gkb = c.ul * nkb + c.kb, j = gkb & 1;            // global key-block counter -> accumulator buffer
          // FIXED accumulation order (bitwise run-to-run determinism): dK_b / dV_b take the query tiles in order 0, 1 and
          // dQ takes the key blocks in order 0, 1, ... -- a slot that is early waits for the other one here
          if (ntq == 2 ? (contrib_of[gkb & 3] != c.qt) : (dq_unit[0] == c.ul ? dq_cnt[0] != c.kb : c.kb != 0)) continue;
          // first contribution to this key block: its accumulator buffer must have been stored (two blocks ago)
          if (contrib_of[gkb & 3] == 0 && gkb >= 2 && !mbar_test_u(&bars->dkv_free[j], ((gkb >> 1) - 1) & 1)) continue;
          // first contribution to dQ of this query tile in this head: the previous head's dQ must have been stored
          const bool dq_first = dq_unit[c.qt] != c.ul;
          if (dq_first && c.ul > 0 && !mbar_test_u(&bars->dq_free[c.qt], (c.ul - 1) & 1)) continue;
          fence_after();
          {
            const uint32_t sq = stages_a + (uint32_t)c.st * stage_bytes, sdo = sq + q_bytes, sk = sdo + q_bytes;
            const int kw = min(KB, p.Npad - c.kb * KB);
            const uint32_t tdk = tmem_base + B_DKV + 64 * j, tdv = tdk + 32;
            const uint32_t acc0 = contrib_of[gkb & 3] > 0 ? 1u : 0u;
            const uint64_t dod = desc64(sdo + (uint32_t)c.qt * QT * ROWB, 16), qd = desc64(sq + (uint32_t)c.qt * QT * ROWB, 16);
            const uint64_t dsk = desc64(smem_a + (2 * s + 1) * PBUF, 16);
            const uint64_t kd0 = desc64(sk + (uint32_t)(c.kb * KB) * ROWB, 16);
            const uint32_t tdq = tmem_base + B_DQ + c.qt * 32;
            const int nk = kw >> 4;
            const bool dkv_complete = ++contrib_of[gkb & 3] == ntq;      // dK_b / dV_b complete once these products retire
            if (dkv_complete) contrib_of[gkb & 3] = 0;
            if (dq_first) { dq_unit[c.qt] = c.ul; dq_cnt[c.qt] = 0; }
            const bool dq_complete = ++dq_cnt[c.qt] == nkb;
            const bool head_done = ++done_in_stage[c.st] == items_per_unit;   // every product reading this head's tiles issued
            if (head_done) done_in_stage[c.st] = 0;
            tr.ev(30 + s);
            if (leader) {
              // dV_b += P^T dO ; dK_b += dS^T Q : K = the 128 query rows of this tile (8 k-steps of 16 rows = 1024 B)
              umma_bf16(tdv, pd_[s], dod, idT, acc0);
#pragma unroll
              for (int kk = 1; kk < QT / 16; ++kk) umma_bf16(tdv, pd_[s] + kk * 64, dod + kk * 64, idT, 1u);
              umma_bf16(tdk, dsd_[s], qd, idT, acc0);
#pragma unroll
              for (int kk = 1; kk < QT / 16; ++kk) umma_bf16(tdk, dsd_[s] + kk * 64, qd + kk * 64, idT, 1u);
              // dQ_qt += dS K_b : K = the keys of this block (kw / 16 k-steps; 16 keys = half a 32-key block = 32 B)
              umma_bf16(tdq, dsk, kd0, idQ, dq_first ? 0u : 1u);
              for (int kk = 1; kk < nk; ++kk)
                umma_bf16(tdq, dsk + (uint64_t)((kk >> 1) * ((QT * ROWB) >> 4) + (kk & 1) * 2), kd0 + (uint64_t)(kk * ((16 * ROWB) >> 4)), idQ, 1u);
              umma_commit(&bars->mma_done[s]);
              if (dkv_complete) umma_commit(&bars->dkv_ready[j]);
              if (dq_complete) umma_commit(&bars->dq_ready[c.qt]);
              if (head_done) umma_commit(&bars->empty[c.st]);
            }
            __syncwarp();
            tr.ev(40 + s);
          }
          ++n_mma[s];
          prod_ready[s] = false;
          cur_next(mm[s], s);
          --remaining;
        }
      }
    }
  } else {
    // ===================== softmax-backward warpgroups =====================
    const int g = warp >> 2, s = g >> 1, hf = g & 1;
    const int t = tid & 127;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t ts = lane_base + s * B_SLOT;
    const float sl2 = p.scale * 1.4426950408889634f;
    const uint32_t sp = smem_u32(smem + (2 * s) * PBUF), sds = sp + PBUF;
    uint32_t n_item = 0;          // items this slot has processed (parity source of sdp_ready / mma_done)
    // per-row constants of a head: lse (log2 units) and D = sum_d dO * O (backward of models/volo.py:193-197:
    // dS = P (dP - D)); D comes from the row-dot pre-pass ([B, heads, N] like lse), both are fetched ONE HEAD AHEAD
    auto load_consts = [&](int ul, float& lse2, float& Dr) {
      lse2 = 0.f; Dr = 0.f;
      if (ul >= my_units) return;
      const int unit = blockIdx.x + ul * gridDim.x;
      const int qt = (ntq == 2) ? (s ^ (ul & 1)) : 0;
      const int row = qt * QT + t;
      if (row < N) {
        const size_t off = (size_t)unit * N + row;            // unit = b * heads + hd
        lse2 = __ldg(p.lse + off) * 1.4426950408889634f;
        Dr = __ldg(p.rowdot + off);
      }
    };
    // dK (warpgroups with hf == 0) / dV (hf == 1) of global key block gkb: lanes = its keys; slot s stores channels [16 s, +16)
    auto store_dkv = [&](int unit, int kb, int gkb) {
      const int b = unit / p.heads, hd = unit % p.heads;
      const int j = gkb & 1;
      const int k0 = kb * KB, kw = min(KB, p.Npad - k0);
      tr.ev(6);
      mbar_wait(&bars->dkv_ready[j], (gkb >> 1) & 1);
      fence_after();
      tr.ev(7);
      const bool key_live = (k0 + (warp & 3) * 32) < min(N, k0 + kw);      // warp-uniform
      if (key_live) {
        uint32_t o[16];
        tmem_ld16(lane_base + B_DKV + 64 * j + 32 * hf + 16 * s, o);
        wait_ld();
        pin16(o);
        const int key = k0 + t;
        if (t < kw && key < N) {
          const float f = (hf == 0) ? p.scale : 1.f;
          bf16* dst = p.out + ((size_t)b * N + key) * 3 * C + (size_t)(1 + hf) * C + (size_t)hd * HD + 16 * s;
#pragma unroll
          for (int c4 = 0; c4 < 2; ++c4) {
            uint4 pkv;
            pkv.x = pack2(__uint_as_float(o[c4 * 8 + 0]) * f, __uint_as_float(o[c4 * 8 + 1]) * f);
            pkv.y = pack2(__uint_as_float(o[c4 * 8 + 2]) * f, __uint_as_float(o[c4 * 8 + 3]) * f);
            pkv.z = pack2(__uint_as_float(o[c4 * 8 + 4]) * f, __uint_as_float(o[c4 * 8 + 5]) * f);
            pkv.w = pack2(__uint_as_float(o[c4 * 8 + 6]) * f, __uint_as_float(o[c4 * 8 + 7]) * f);
            *reinterpret_cast<uint4*>(dst + c4 * 8) = pkv;
          }
        }
      }
      fence_before();
      tr.ev(8);
      mbar_arrive(&bars->dkv_free[j]);
    };
    // dQ of the query tiles of CTA-local head ul: stored by the slot that owns item (qq, 0); halves store 16 channels each
    auto store_dq = [&](int ul) {
      const int unit = blockIdx.x + ul * gridDim.x;
      const int b = unit / p.heads, hd = unit % p.heads;
      for (int qq = 0; qq < ntq; ++qq) {
        const int owner = (ntq == 2) ? (qq ^ (ul & 1)) : 0;
        if (owner != s) continue;
        tr.ev(9);
        mbar_wait(&bars->dq_ready[qq], ul & 1);
        fence_after();
        tr.ev(10);
        const int rowq = qq * QT + t;
        if ((qq * QT + (warp & 3) * 32) < N) {
          uint32_t o[16];
          tmem_ld16(lane_base + B_DQ + qq * 32 + 16 * hf, o);
          wait_ld();
          pin16(o);
          if (rowq < N) {
            bf16* dst = p.out + ((size_t)b * N + rowq) * 3 * C + (size_t)hd * HD + 16 * hf;
            const float f = p.scale;
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
              uint4 pkv;
              pkv.x = pack2(__uint_as_float(o[c4 * 8 + 0]) * f, __uint_as_float(o[c4 * 8 + 1]) * f);
              pkv.y = pack2(__uint_as_float(o[c4 * 8 + 2]) * f, __uint_as_float(o[c4 * 8 + 3]) * f);
              pkv.z = pack2(__uint_as_float(o[c4 * 8 + 4]) * f, __uint_as_float(o[c4 * 8 + 5]) * f);
              pkv.w = pack2(__uint_as_float(o[c4 * 8 + 6]) * f, __uint_as_float(o[c4 * 8 + 7]) * f);
              *reinterpret_cast<uint4*>(dst + c4 * 8) = pkv;
            }
          }
        }
        fence_before();
        mbar_arrive(&bars->dq_free[qq]);
      }
    };
    float lse2, Dr, lse2_n, Dr_n;
    load_consts(0, lse2, Dr);
    // stores run one step LATE (also across heads): the products of a key block / a head retire while the next item is
    // being computed, so nobody waits for the tensor pipe
    int pend_unit = -1, pend_kb = 0, pend_gkb = 0, pend_dq_ul = -1;
    for (int ul = 0; ul < my_units; ++ul) {
      const int unit = blockIdx.x + ul * gridDim.x;
      const int qt = (ntq == 2) ? (s ^ (ul & 1)) : 0;          // the query tile this slot works on in this head
      const int row = qt * QT + t;
      const bool row_ok = row < N;
      const bool warp_live = (qt * QT + (warp & 3) * 32) < N;
      load_consts(ul + 1, lse2_n, Dr_n);
      for (int kb = 0; kb < nkb; ++kb) {
        const bool mine = (ntq == 2) || ((kb & 1) == s);       // does this slot own item (qt, kb)?
        if (mine) {
          const int k0 = kb * KB, kw = min(KB, p.Npad - k0);
          const int nch = kw >> 4, h0 = (nch + 1) >> 1;
          const int c_lo = hf ? h0 : 0, c_hi = hf ? nch : h0;
          tr.ev(1);
          mbar_wait(&bars->sdp_ready[s], n_item & 1);
          fence_after();
          tr.ev(2);
          // (the previous item's products must have read P / dS before they are overwritten: waited for right before the
          //  first shared-memory store, so the TMEM loads and the exponentials of the first chunk overlap those products)
          const bool wait_prev = n_item > 0;
          const uint32_t prev_par = (n_item & 1) ^ 1;
          if (warp_live) {
            for (int c = c_lo; c < c_hi; ++c) {
              uint32_t s_[16], d_[16];
              tmem_ld16(ts + (uint32_t)(c * 16), s_);
              tmem_ld16(ts + KB + (uint32_t)(c * 16), d_);
              wait_ld();
              pin16(s_);
              pin16(d_);
              uint32_t pk[8], dk[8];
              const bool full = (k0 + (c + 1) * 16 <= N) && row_ok;
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                float p0 = ex2f(fmaf(__uint_as_float(s_[i]), sl2, -lse2)), p1 = ex2f(fmaf(__uint_as_float(s_[i + 1]), sl2, -lse2));
                if (!full) {
                  if (!row_ok || k0 + c * 16 + i >= N) p0 = 0.f;
                  if (!row_ok || k0 + c * 16 + i + 1 >= N) p1 = 0.f;
                }
                const float g0 = p0 * (__uint_as_float(d_[i]) - Dr), g1 = p1 * (__uint_as_float(d_[i + 1]) - Dr);
                pk[i >> 1] = pack2(p0, p1);
                dk[i >> 1] = pack2(g0, g1);
              }
              // keys [16c, 16c+16) of this block -> 32-key block (c >> 1), 16-byte chunks 2 (c & 1) and 2 (c & 1) + 1 of row t
              const uint32_t blk = (uint32_t)(c >> 1) * (QT * ROWB);
              const int c0 = (c & 1) * 2;
              if (c == c_lo && wait_prev) { tr.ev(3); mbar_wait(&bars->mma_done[s], prev_par); tr.ev(4); }
              sts128(sp + blk + sw64(t, c0), pk[0], pk[1], pk[2], pk[3]);
              sts128(sp + blk + sw64(t, c0 + 1), pk[4], pk[5], pk[6], pk[7]);
              sts128(sds + blk + sw64(t, c0), dk[0], dk[1], dk[2], dk[3]);
              sts128(sds + blk + sw64(t, c0 + 1), dk[4], dk[5], dk[6], dk[7]);
            }
          } else {
            // rows of this warp are all beyond N: their P / dS must still be ZERO (they are K rows of the dV / dK products)
            if (wait_prev) mbar_wait(&bars->mma_done[s], prev_par);
            for (int c = c_lo; c < c_hi; ++c) {
              const uint32_t blk = (uint32_t)(c >> 1) * (QT * ROWB);
              const int c0 = (c & 1) * 2;
              sts128(sp + blk + sw64(t, c0), 0u, 0u, 0u, 0u);
              sts128(sp + blk + sw64(t, c0 + 1), 0u, 0u, 0u, 0u);
              sts128(sds + blk + sw64(t, c0), 0u, 0u, 0u, 0u);
              sts128(sds + blk + sw64(t, c0 + 1), 0u, 0u, 0u, 0u);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMAs
          fence_before();
          tr.ev(5);
          mbar_arrive(&bars->pds_ready[s]);
          ++n_item;
        }
        // late stores: the previous head's dQ (once, after this head's first item) and the previous key block's dK / dV
        if (pend_dq_ul >= 0) { store_dq(pend_dq_ul); pend_dq_ul = -1; }
        if (pend_unit >= 0) store_dkv(pend_unit, pend_kb, pend_gkb);
        pend_unit = unit; pend_kb = kb; pend_gkb = ul * nkb + kb;
      }
      pend_dq_ul = ul;
      lse2 = lse2_n; Dr = Dr_n;
    }
    if (pend_unit >= 0) store_dkv(pend_unit, pend_kb, pend_gkb);
    if (pend_dq_ul >= 0) store_dq(pend_dq_ul);
  }
  fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// D[b, h, n] = sum_d dO[b, n, h, d] * O[b, n, h, d]  (fp32): one thread per (b, n, h), 64 B of each operand; consecutive
// threads take consecutive heads, so a warp reads 2 KB contiguous per tensor.  Output in the [B, heads, N] layout of lse.
__global__ void __launch_bounds__(256) mhsa_rowdot_tc_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                             float* __restrict__ D, long long total, int N, int heads) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;      // (b * N + n) * heads + h
  if (i >= total) return;
  const uint4* po = reinterpret_cast<const uint4*>(o + i * HD);
  const uint4* pd = reinterpret_cast<const uint4*>(dout + i * HD);
  float acc = 0.f;
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const uint4 a = __ldg(po + c4), d = __ldg(pd + c4);
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __bfloat1622float2(ha[k]), fb = __bfloat1622float2(hb[k]);
      acc = fmaf(fa.x, fb.x, acc);
      acc = fmaf(fa.y, fb.y, acc);
    }
  }
  const int h = (int)(i % heads);
  const long long bn = i / heads;
  const int n = (int)(bn % N);
  const long long b = bn / N;
  D[(b * heads + h) * N + n] = acc;
}

// 2-D bf16 tensor [rows, cols] row-major, box {32 channels (64 B), box_rows}, 64B swizzle, zero fill outside the tensor
int make_map64(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { apb_set_error("mhsa_tc: cuTensorMapEncodeTiled entry point unavailable"); return APB_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)HD, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    apb_set_error("mhsa_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d", (int)r, rows, cols, box_rows);
    return APB_ERR_ARG;
  }
  return 0;
}

bool tc_envelope(const void* a, const void* b, int B, int N, int heads, int D) {
  return D == HD && B > 0 && heads > 0 && N >= 1 && N <= 224 && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

std::mutex g_attr_mu;
long long* g_trace = nullptr;

}  // namespace

// returns APB_ERR_UNSUPPORTED for shapes outside the single-tile envelope (head_dim 32, N <= 224)
int apb_mhsa_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (!tc_envelope(qkv, out, B, N, heads, D)) return APB_ERR_UNSUPPORTED;
  AttParams p{};
  p.qkv = (const bf16*)qkv; p.out = (bf16*)out; p.lse = lse; p.B = B; p.N = N; p.heads = heads; p.scale = scale;
  p.Npad = (N + 15) / 16 * 16;
  p.ntq = (N + QT - 1) / QT;
  p.units = B * heads;
  p.trace = g_trace;
  const int C = heads * HD;
  CUtensorMap mq, mkv;
  int rc = make_map64(&mq, qkv, (long long)B * N, 3LL * C, p.ntq * QT);
  if (rc) return rc;
  rc = make_map64(&mkv, qkv, (long long)B * N, 3LL * C, p.Npad);
  if (rc) return rc;
  const size_t smem = (size_t)F_NST * ((size_t)p.ntq * QT + 2 * p.Npad) * ROWB + sizeof(FwdBars) + 1024;
  {
    std::lock_guard<std::mutex> g(g_attr_mu);
    static size_t attr = 0;
    if (smem > attr) {
      cudaError_t e = cudaFuncSetAttribute(mhsa_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { apb_set_error("mhsa_fwd_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
      attr = smem;
    }
  }
  const int grid = p.units < num_sms() ? p.units : num_sms();
  mhsa_fwd_tc_kernel<<<grid, NTHREADS, smem, st>>>(mq, mkv, p);
  APB_LAUNCH_CHECK("mhsa_fwd_tc");
  return 0;
}

int apb_mhsa_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace, int B,
                    int N, int heads, int D, float scale, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (!tc_envelope(qkv, dqkv, B, N, heads, D) || (((uintptr_t)out | (uintptr_t)dout) & 15) != 0) return APB_ERR_UNSUPPORTED;
  APB_CHECK_ARG(workspace != nullptr, APB_ERR_ARG, "mhsa_bwd_tc: workspace of B*heads*N floats required");
  AttParams p{};
  p.qkv = (const bf16*)qkv; p.o = (const bf16*)out; p.dout = (const bf16*)dout; p.out = (bf16*)dqkv; p.lse = const_cast<float*>(lse);
  p.rowdot = workspace;
  p.B = B; p.N = N; p.heads = heads; p.scale = scale;
  p.Npad = (N + 15) / 16 * 16;
  p.ntq = (N + QT - 1) / QT;
  p.units = B * heads;
  p.trace = g_trace;
  const int C = heads * HD;
  CUtensorMap mq, mkv, mdo;
  int rc = make_map64(&mq, qkv, (long long)B * N, 3LL * C, p.ntq * QT);
  if (rc) return rc;
  rc = make_map64(&mkv, qkv, (long long)B * N, 3LL * C, p.Npad);
  if (rc) return rc;
  rc = make_map64(&mdo, dout, (long long)B * N, (long long)C, p.ntq * QT);
  if (rc) return rc;
  const size_t smem = 4 * (size_t)PBUF + (size_t)B_NST * (2 * (size_t)p.ntq * QT + 2 * p.Npad) * ROWB + sizeof(BwdBars) + 1024;
  if (smem > 227 * 1024) return APB_ERR_UNSUPPORTED;
  {
    std::lock_guard<std::mutex> g(g_attr_mu);
    static size_t attr = 0;
    if (smem > attr) {
      cudaError_t e = cudaFuncSetAttribute(mhsa_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { apb_set_error("mhsa_bwd_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
      attr = smem;
    }
  }
  {
    const long long total = (long long)B * N * heads;
    mhsa_rowdot_tc_kernel<<<ceil_div(total, 256), 256, 0, st>>>((const bf16*)out, (const bf16*)dout, workspace, total, N, heads);
    APB_LAUNCH_CHECK("mhsa_rowdot_tc");
  }
  const int grid = p.units < num_sms() ? p.units : num_sms();
  mhsa_bwd_tc_kernel<<<grid, NTHREADS, smem, st>>>(mq, mkv, mdo, p);
  APB_LAUNCH_CHECK("mhsa_bwd_tc");
  return 0;
}

// diagnostic: device buffer of 18 * 1024 int64 that CTA 0 of the next attention launches fills with (event, clock) pairs
extern "C" int apb_debug_mhsa_trace(long long* buf) {
  g_trace = buf;
  return 0;
}
