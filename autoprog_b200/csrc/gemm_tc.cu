// bf16 GEMM on the 5th-generation tensor cores: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma (kind::f16, fp32 accumulate in TMEM) -> tcgen05.ld epilogue (bias / GELU / dGELU / split-K partials).
//
// Serves every Linear and patchify-conv of the VOLO hot path in bf16 mode (models/volo.py:80,88,100,161-167,188,199,
// 253-256,370-373,389,547-554): forward (NT), dgrad (NN) and wgrad (TN) through the operand-major bits of the
// instruction descriptor -- no transposed copies of activations or weights are ever materialised.
//
//   C[M,N] = epi( sum_k A(m,k) B(n,k) + bias[n] )
//   trans_a = 0: A is [M,K] row-major  (K-major operand)    1: A is [K,M] row-major (MN-major operand)
//   trans_b = 0: B is [N,K] row-major  (K-major operand)    1: B is [K,N] row-major (MN-major operand)
//
// PERSISTENT kernel, one CTA per SM: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// 8 / 12 / 16 epilogue warps (one per TMEM lane quarter x column group).  Instantiations <BN, STAGES, EPI_WARPS, AUX>:
//   128 x 192 tile: 4 stages, 12 epilogue warps  (3 stages when AUX: the gelu' staging boxes take the 4th stage's room)
//   128 x 128 tile: 4 stages, 16 epilogue warps;  128 x 64 tile (stem conv, N = 64): 6 stages, 8 epilogue warps
// TWO accumulator stages in TMEM: the epilogue of tile i overlaps the main loop of tile i+1.  Epilogue warps stage their
// 32 x 32 sub-tiles in swizzled shared memory and write them with TMA stores (fully coalesced, asynchronous, M/N tails
// clipped by the TMA unit).  K is short on this path (3..18 k-blocks): the main loop is bound by L2->SM operand delivery
// (bytes in flight), the GELU epilogue by CUDA-core issue (A&S 7.1.26 erf: one MUFU.RCP + one MUFU.EX2 per element).
//   * epilogue 1 (GELU) also stores gelu'(u) for the backward; epilogue 2 (dGELU) multiplies by it, the gelu' tiles
//     being TMA-prefetched one tile ahead into two box sets per warp and the product written in place;
//   * split_k > 1 writes fp32 partial tiles [split][M][N] that the caller reduces in fixed order (deterministic wgrad);
//   * rowsum (apb_gemm_tc_rowsum): sum_k A(m,k) -- the bias gradient of a wgrad GEMM -- from one extra 128 x 16 x 16 MMA
//     per k-step against a tile of ones, shared between the n-tiles of an m-tile.
// The forward-shaped K >= 384 products run on CTA pairs instead (gemm_tc2.cu, cta_group::2).
#include "gemm_tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int ACC_STAGES = 2;

struct TcParams {
  void* C;
  const float* bias;
  void* aux;
  float* rowsum;        // optional [splits * tiles_n][M]: partial sum_k A(m,k) (bias gradient of a wgrad GEMM)
  int M, N, K;
  int a_mn, b_mn;       // operand majors (1 = MN-major)
  int epilogue;         // 0 none, 1 gelu, 2 dgelu, 4 split-k partial
  int out_f32;
  int kb_per_split;     // k-blocks per split
  int splits;
  int tiles_m, tiles_n;
  int dbg;              // diagnostics (apb_debug_gemm_switches, tools/gemm_bound.py): 1 = no TMA loads (stale operands), 2 = no MMAs, 4 = no stores
};

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7): cdf = Phi(x), e = exp(-x^2/2).  The epilogue is issue-bound on the
// CUDA cores, so both transcendentals are single MUFU ops (rcp.approx / ex2.approx: 1-2 ulp, far inside the bf16 output).
__device__ __forceinline__ void phi_fast(float x, float& cdf, float& e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.f));
  e = ex2_approx(x * x * -0.72134752044448170368f);     // exp(-x^2 / 2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_tail = 0.5f * poly * t * e;          // 0.5 * (1 - erf(z))
  cdf = x >= 0.f ? 1.f - half_tail : half_tail;
}

// AUX = the epilogue needs the per-warp gelu' boxes (GELU / dGELU).  Without them the shared memory they would take
// becomes one more pipeline stage: the main loop is bound by bytes in flight from L2, not by the tensor pipe.
// NARROW != 0: one 2 KB staging box per epilogue warp so that the shared memory saved holds a FIFTH pipeline stage (the k-block
// period is (load latency + MMA time of a stage) / STAGES, tools/gemm_bound.py).  1 = bf16 output, no row sums (no tile of
// ones either); 2 = fp32 output (wgrad / split-K partials) stored as 16-column half boxes, row sums kept, 8 epilogue warps.
template <int BN, int STAGES, int EPI_WARPS, bool AUX, int NARROW>
__global__ void __launch_bounds__(64 + EPI_WARPS * 32, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                             const __grid_constant__ CUtensorMap tma_b,
                                                             const __grid_constant__ CUtensorMap tma_c,
                                                             const __grid_constant__ CUtensorMap tma_x, TcParams p) {
  constexpr uint32_t A_BYTES = BM * BK * 2;   // 16 KB
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t RS_COLS = 16;                                        // row-sum accumulator: one N = 16 MMA per k-step
  constexpr uint32_t TMEM_COLS = (ACC_STAGES * (BN + RS_COLS) <= 256) ? 256 : 512;   // power of two >= 2 accumulator stages
  constexpr uint32_t ONES_BYTES = (AUX || NARROW == 1) ? 0 : 2048;                         // 16 rows x 128 B of bf16 1.0 (the "B operand" of the row sum)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // per epilogue warp: C boxes (AUX kernels write bf16 only: one 2 KB box per chunk; otherwise 4 KB so that the single
  // 32 x 128 B fp32 box fits) + one 2 KB gelu' box per chunk
  constexpr uint32_t NCH_ = (BN / 32) / (EPI_WARPS / 4);
  constexpr uint32_t STGC_BYTES = AUX ? NCH_ * 2048 : (NARROW ? 2048 : 4096);
  constexpr uint32_t STG_BYTES = STGC_BYTES + (AUX ? NCH_ * 2048 : 0);
  uint8_t* ones = smem + STAGES * STAGE_BYTES;
  uint8_t* stg_base = ones + ONES_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + EPI_WARPS * STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint64_t* aux_bar = tempty_bar + ACC_STAGES;          // [EPI_WARPS][2]: gelu' tiles of the dGELU epilogue (AUX only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + (AUX ? 2 * EPI_WARPS : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (p.K + BK - 1) / BK;
  const int n_items = p.tiles_m * p.tiles_n * p.splits;

  if (!AUX && NARROW != 1 && p.rowsum != nullptr) {
    for (int i = threadIdx.x; i < (int)(ONES_BYTES / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA's async proxy
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_x) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EPI_WARPS); }
    if (AUX) for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&aux_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    const bool leader = elect_one();
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int z = item % p.splits, t = item / p.splits;
      const int m0 = (t / p.tiles_n) * BM, n0 = (t % p.tiles_n) * BN;
      const int kb0 = z * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        const int k0 = kb * BK;
        if (leader && (p.dbg & 1)) {
          mbar_expect_tx(&full_bar[s], 0);
        } else if (leader) {
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          if (!p.a_mn) {
            tma_load_2d(sa, &tma_a, &full_bar[s], k0, m0);                 // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int h = 0; h < BM / 64; ++h) tma_load_2d(sa + h * (BK * 128), &tma_a, &full_bar[s], m0 + h * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tma_b, &full_bar[s], k0, n0);                 // box {64 k, BN n}
          } else {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h) tma_load_2d(sb + h * (BK * 128), &tma_b, &full_bar[s], n0 + h * 64, k0);
          }
        }
        __syncwarp();
      }
      // last tile's loads are in flight: let the next kernel in the stream get scheduled and run its prologue while this
      // CTA finishes (not earlier: a dependent CTA parked on this SM for the whole kernel competes for issue slots)
      if (item + (int)gridDim.x >= n_items) pdl_trigger();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    {
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
      // k-step of 16 inside a stage, in descriptor units (16 B): K-major +32 B, MN-major +16 rows of 128 B
      const uint64_t a_kstep = p.a_mn ? (16 * 128) >> 4 : 32 >> 4, b_kstep = p.b_mn ? (16 * 128) >> 4 : 32 >> 4;
      // row sums ride on the tensor pipe: one extra 128 x 16 x 16 MMA per k-step against a tile of ones.  The n-tiles
      // of an m-tile read the same A tiles, so they share the work: n-tile j covers the k-blocks with kb % tiles_n == j
      const uint32_t idesc_rs = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)(RS_COLS >> 3) << 17) |
                                ((uint32_t)(BM >> 4) << 24);
      const uint64_t ones_desc = make_smem_desc(smem_u32(ones), 16, 1024);
      const uint32_t smem_a0 = smem_u32(smem);
      uint32_t it = 0, ai = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ai) {
        const int z = item % p.splits;
        const int tn = (item / p.splits) % p.tiles_n;
        uint32_t rs_acc = 0;
        const int kb0 = z * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const uint32_t as = ai % ACC_STAGES;
        mbar_wait(&tempty_bar[as], ((ai / ACC_STAGES) & 1) ^ 1);      // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_u + as * BN, trs = tmem_u + ACC_STAGES * BN + as * RS_COLS;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_a0 + s * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t ad0 = make_smem_desc(sa, a_lbo, 1024), bd0 = make_smem_desc(sb, b_lbo, 1024);
          const uint32_t acc_first = kb > kb0 ? 1u : 0u;
          // two straight-line versions of the k-steps (no conditionally executed tensor-core instruction)
          const bool with_rs = !AUX && NARROW != 1 && p.rowsum != nullptr && (kb % p.tiles_n) == tn;
          if (leader) {
            if (p.dbg & 2) {
            } else if (with_rs) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_bf16(tacc, ad0 + k * a_kstep, bd0 + k * b_kstep, idesc, k > 0 ? 1u : acc_first);
                umma_bf16(trs, ad0 + k * a_kstep, ones_desc, idesc_rs, k > 0 ? 1u : rs_acc);
              }
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_bf16(tacc, ad0 + k * a_kstep, bd0 + k * b_kstep, idesc, k > 0 ? 1u : acc_first);
            }
            umma_commit(&empty_bar[s]);   // frees the smem stage once the MMAs above have read it
            if (kb + 1 == kb1) umma_commit(&tfull_bar[as]);    // accumulator of this tile complete
          }
          if (with_rs) rs_acc = 1u;
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> swizzled smem -> TMA store =====================
    constexpr int NCH = (BN / 32) / (EPI_WARPS / 4);   // 32-column chunks per warp
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const int cg = (warp - 2) >> 2;               // which column slice of the tile
    uint8_t* stgC = stg_base + (warp - 2) * STG_BYTES;   // bf16: NCH boxes of 32 x 64 B; fp32: one 32 x 128 B box
    uint8_t* stgX = stgC + STGC_BYTES;                   // aux boxes (bf16 32 x 64 B each)
    // dGELU: the gelu'(pre-activation) sub-tiles of tile i+1 are fetched by TMA (same 32 x 32 / 64B-swizzle boxes the
    // GELU forward stored them with) while tile i is in its epilogue.  Two box sets per warp, used alternately; the
    // product is written IN PLACE over the gelu' box and TMA-stored from there, so no extra shared memory is needed.
    uint64_t* my_aux_bar = aux_bar + (warp - 2) * 2;
    auto prefetch_aux = [&](int item_, int set) {
      if (lane != 0) return;
      const int t_ = item_ / p.splits;
      const int rb_ = (t_ / p.tiles_n) * BM + quarter * 32, cb_ = (t_ % p.tiles_n) * BN + cg * NCH * 32;
      uint32_t bytes = 0;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (rb_ < p.M && cb_ + c * 32 < p.N) bytes += 2048;
      if (bytes == 0) { mbar_arrive(&my_aux_bar[set]); return; }       // keep the phases of the two sets in lockstep
      mbar_expect_tx(&my_aux_bar[set], bytes);
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (cb_ + c * 32 < p.N) tma_load_3d((set ? stgX : stgC) + c * 2048, &tma_x, &my_aux_bar[set], cb_ + c * 32, rb_, 0);
    };
    const bool dgelu = AUX && p.epilogue == 2;
    if (dgelu && (int)blockIdx.x < n_items) prefetch_aux(blockIdx.x, 0);
    uint32_t ai = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ai) {
      const int z = item % p.splits, t = item / p.splits;
      const int m0 = (t / p.tiles_n) * BM, n0 = (t % p.tiles_n) * BN;
      const uint32_t as = ai % ACC_STAGES;
      const int rb = m0 + quarter * 32;           // first row of this warp's sub-tile
      const int cb0 = n0 + cg * NCH * 32;         // first column
      mbar_wait(&tfull_bar[as], (ai / ACC_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)((cg * NCH + c) * 32), r[c]);
      const bool rs_here = !AUX && p.rowsum != nullptr && cg == 0;   // warp-uniform
      // tcgen05.ld is asynchronous: its destination registers are valid only after wait::ld.  The row-sum load is issued
      // unconditionally (no select on its result that the compiler could evaluate before the wait) and pinned below.
      uint32_t rsv = 0;
      if (!AUX) rsv = tmem_ld1(tmem_base + ((uint32_t)(quarter * 32) << 16) + ACC_STAGES * BN + as * RS_COLS);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("" : "+r"(rsv)::"memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tempty_bar[as]);                                  // accumulator stage free for tile i+2
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous TMA stores have drained the staging boxes
      }
      __syncwarp();
      if (dgelu) {
        // the other box set is free (its TMA store was drained above): start fetching the next tile's gelu' now
        if (item + (int)gridDim.x < n_items) prefetch_aux(item + gridDim.x, (ai & 1) ^ 1);
        if (rb < p.M) mbar_wait(&my_aux_bar[ai & 1], (ai >> 1) & 1);
      }
      if (rs_here && rb + lane < p.M) {
        const int tn = t % p.tiles_n;
        const int kb0 = z * p.kb_per_split, kb1 = min(total_kb, kb0 + p.kb_per_split);
        const int first = kb0 + ((tn - kb0 % p.tiles_n) + p.tiles_n) % p.tiles_n;   // first k-block this n-tile summed
        p.rowsum[((size_t)z * p.tiles_n + tn) * p.M + rb + lane] = first < kb1 ? __uint_as_float(rsv) : 0.f;
      }
      if (rb >= p.M) continue;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int cb = cb0 + c * 32;
        if (cb >= p.N) continue;
        uint8_t* sx = dgelu ? (((ai & 1) ? stgX : stgC) + c * 2048) : stgX + c * 2048;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c][j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (cb + j < p.N) {                     // N % 8 == 0 on this path
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
        }
        if (AUX && p.epilogue == 2) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 pk = *reinterpret_cast<const uint4*>(stg64(sx, lane, ch));
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              v[ch * 8 + 2 * tt] *= __bfloat162float(h2[tt].x);
              v[ch * 8 + 2 * tt + 1] *= __bfloat162float(h2[tt].y);
            }
          }
        } else if (AUX && p.epilogue == 1) {
          // GELU: C = gelu(v); aux = gelu'(v) (what the dgrad epilogue multiplies by); Phi and the Gaussian are shared
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              float c0, e0, c1, e1;
              const float x0 = v[ch * 8 + 2 * tt], x1 = v[ch * 8 + 2 * tt + 1];
              phi_fast(x0, c0, e0);
              phi_fast(x1, c1, e1);
              h2[tt] = __floats2bfloat162_rn(fmaf(x0 * 0.39894228040143267794f, e0, c0), fmaf(x1 * 0.39894228040143267794f, e1, c1));
              v[ch * 8 + 2 * tt] = x0 * c0;
              v[ch * 8 + 2 * tt + 1] = x1 * c1;
            }
            *reinterpret_cast<uint4*>(stg64(sx, lane, ch)) = pk;
          }
        }
        const bool single_box = p.out_f32 || NARROW != 0;
        uint8_t* sc = dgelu ? sx : (single_box ? stgC : stgC + c * 2048);
        if (single_box && c > 0) {                  // the single box is reused: wait until the previous store has read it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        if (NARROW == 2) {
          // fp32 through the 2 KB box: two 16-column halves (32 rows x 64 B, 64B swizzle), one TMA store each
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            if (hb > 0) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              *reinterpret_cast<float4*>(stg64(sc, lane, ch)) =
                  make_float4(v[hb * 16 + ch * 4], v[hb * 16 + ch * 4 + 1], v[hb * 16 + ch * 4 + 2], v[hb * 16 + ch * 4 + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && !(p.dbg & 4) && cb + hb * 16 < p.N) {
              tma_store_3d(&tma_c, sc, cb + hb * 16, rb, z);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          continue;
        }
        if (p.out_f32) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<float4*>(stg128(sc, lane, ch)) = make_float4(v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
        } else {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) h2[tt] = __floats2bfloat162_rn(v[ch * 8 + 2 * tt], v[ch * 8 + 2 * tt + 1]);
            *reinterpret_cast<uint4*>(stg64(sc, lane, ch)) = pk;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && !(p.dbg & 4)) {
          tma_store_3d(&tma_c, sc, cb, rb, z);
          if (AUX && p.epilogue == 1) tma_store_3d(&tma_x, sx, cb, rb, 0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// tile width: 64 for narrow outputs (the stem conv, N = 64); 192 when it wastes no more columns than 128 (N = 192, 384,
// 576, 768, 1152, ...): 20 % less L2 operand traffic per FLOP and fewer tiles; otherwise 128
int pick_bn(int N) {
  if (N <= 64) return 64;
  return (ceil_div(N, 192) * 192 <= ceil_div(N, 128) * 128) ? 192 : 128;
}

template <int BN, int STAGES, int EPI_WARPS, bool AUX, int NARROW = 0>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mx, TcParams p, int splits,
           cudaStream_t st) {
  constexpr size_t smem = STAGES * (BM * BK * 2 + BN * BK * 2) + ((AUX || NARROW == 1) ? 0 : 2048) +
                          EPI_WARPS * (AUX ? 2 * ((BN / 32) / (EPI_WARPS / 4)) * 2048 : (NARROW ? 2048 : 4096)) + 1024 + 512;
  static_assert(smem <= 227 * 1024, "gemm_tc: shared memory budget");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPI_WARPS, AUX, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { apb_set_error("gemm_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  p.tiles_m = ceil_div(p.M, BM);
  p.tiles_n = ceil_div(p.N, BN);
  p.splits = splits;
  p.dbg = g_apb_gemm_dbg;
  const long long items = (long long)p.tiles_m * p.tiles_n * splits;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  apb_launch_pdl(gemm_tc_kernel<BN, STAGES, EPI_WARPS, AUX, NARROW>, dim3(grid), dim3(64 + EPI_WARPS * 32), smem, st, ma, mb, mc, mx, p);
  APB_LAUNCH_CHECK("gemm_tc");
  return 0;
}

}  // namespace

// split_k > 1: C must hold split_k fp32 partials [split_k][M][N]; bias/epilogue are ignored.
// rowsum_parts (optional): fp32 [apb_gemm_tc_rowsum_slots(N, split_k)][M] partial sums whose total over the slot dim is
// sum_k A(m,k) -- the bias gradient of a wgrad GEMM (A = dY^T), computed on the tensor pipe from the operand tiles
// already in shared memory.
int apb_gemm_tc_rowsum(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                       int trans_b, int epilogue, int in_dtype, int out_dtype, int split_k, float* rowsum_parts,
                       apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(in_dtype == APB_BF16, APB_ERR_DTYPE, "gemm_tc: inputs must be bf16 (got %d)", in_dtype);
  APB_CHECK_ARG(out_dtype == APB_BF16 || out_dtype == APB_F32, APB_ERR_DTYPE, "gemm_tc: out dtype %d", out_dtype);
  APB_CHECK_ARG(M > 0 && N > 0 && K > 0, APB_ERR_SHAPE, "gemm_tc: M=%d N=%d K=%d", M, N, K);
  APB_CHECK_ARG(epilogue >= 0 && epilogue <= 2, APB_ERR_ARG, "gemm_tc: epilogue %d", epilogue);
  APB_CHECK_ARG((epilogue == 0) || aux != nullptr, APB_ERR_ARG, "gemm_tc: GELU epilogues need aux");
  // TMA: row pitches must be multiples of 16 bytes
  const long long pitch_a = trans_a ? M : K, pitch_b = trans_b ? N : K;
  APB_CHECK_ARG(pitch_a % 8 == 0 && pitch_b % 8 == 0, APB_ERR_UNSUPPORTED,
                "gemm_tc: operand row pitches must be multiples of 8 elements (A %lld, B %lld)", pitch_a, pitch_b);
  APB_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0, APB_ERR_ARG,
                "gemm_tc: pointers must be 16-byte aligned");
  const int total_kb = (K + BK - 1) / BK;
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > total_kb) splits = total_kb;
  int kb_per = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per - 1) / kb_per;   // no empty split
  APB_CHECK_ARG(splits == 1 || (split_k == splits), APB_ERR_ARG,
                "gemm_tc: split_k=%d is not realisable for K=%d (use %d)", split_k, K, splits);
  CUtensorMap ma, mb;
  int rc;
  if (!trans_a) rc = make_map(&ma, A, M, K, BM, BK); else rc = make_map(&ma, A, K, M, BK, 64);
  if (rc) return rc;
  const int BN = pick_bn(N);
  if (!trans_b) rc = make_map(&mb, B, N, K, BN, BK); else rc = make_map(&mb, B, K, N, BK, 64);
  if (rc) return rc;
  TcParams p;
  p.C = C; p.bias = bias; p.aux = aux; p.rowsum = rowsum_parts; p.M = M; p.N = N; p.K = K; p.a_mn = trans_a ? 1 : 0; p.b_mn = trans_b ? 1 : 0;
  p.epilogue = splits > 1 ? 4 : epilogue;
  p.out_f32 = (out_dtype == APB_F32) ? 1 : 0;
  p.kb_per_split = kb_per;
  if (splits > 1) APB_CHECK_ARG(out_dtype == APB_F32, APB_ERR_DTYPE, "gemm_tc: split-K partials are fp32");
  APB_CHECK_ARG(!(p.out_f32 && (epilogue == 1 || epilogue == 2)), APB_ERR_UNSUPPORTED, "gemm_tc: GELU epilogues write bf16");
  APB_CHECK_ARG(N % 8 == 0, APB_ERR_UNSUPPORTED, "gemm_tc: N=%d must be a multiple of 8 (TMA store pitch)", N);
  APB_CHECK_ARG(aux == nullptr || ((uintptr_t)aux & 15) == 0, APB_ERR_ARG, "gemm_tc: aux must be 16-byte aligned");
  CUtensorMap mc, mx;
  const bool aux_epi0 = (p.epilogue == 1 || p.epilogue == 2);
  const bool narrow_ok = g_apb_gemm_narrow != 0;     // A/B switch for tools (apb_debug_gemm_switches)
  const bool narrow32 = BN == 192 && !aux_epi0 && p.out_f32 && narrow_ok;
  rc = make_map_out(&mc, C, p.out_f32 != 0, M, N, splits, narrow32);
  if (rc) return rc;
  const bool aux_map = (epilogue == 1 || epilogue == 2);      // GELU stores gelu' through it, dGELU loads gelu' through it
  rc = make_map_out(&mx, aux_map ? aux : C, aux_map ? false : (p.out_f32 != 0), M, N, aux_map ? 1 : splits);
  if (rc) return rc;
  const bool aux_epi = (p.epilogue == 1 || p.epilogue == 2);
  APB_CHECK_ARG(!(aux_epi && rowsum_parts != nullptr), APB_ERR_UNSUPPORTED, "gemm_tc: row sums are not available with GELU epilogues");
  if (BN == 64) return aux_epi ? launch<64, 6, 8, true>(ma, mb, mc, mx, p, splits, st) : launch<64, 6, 8, false>(ma, mb, mc, mx, p, splits, st);
  // (24 epilogue warps for the GELU kernels were measured on the same box: 19.83 vs 19.79 ms / step with 12 -> kept 12)
  // bf16 output without row sums (forward and dgrad products): 5 stages
  if (BN == 192 && !aux_epi && !p.out_f32 && rowsum_parts == nullptr && splits == 1 && narrow_ok)
    return launch<192, 5, 12, false, 1>(ma, mb, mc, mx, p, splits, st);
  // fp32 output (wgrad, split-K partials): 5 stages, 8 epilogue warps, half-box stores
  if (narrow32) return launch<192, 5, 8, false, 2>(ma, mb, mc, mx, p, splits, st);
  if (BN == 192) return aux_epi ? launch<192, 3, 12, true>(ma, mb, mc, mx, p, splits, st) : launch<192, 4, 12, false>(ma, mb, mc, mx, p, splits, st);
  return aux_epi ? launch<128, 4, 16, true>(ma, mb, mc, mx, p, splits, st) : launch<128, 4, 16, false>(ma, mb, mc, mx, p, splits, st);
}

int apb_gemm_tc_rowsum_slots(int N, int split_k) {
  return (split_k < 1 ? 1 : split_k) * ceil_div(N, pick_bn(N));
}

int apb_gemm_tc(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                int trans_b, int epilogue, int in_dtype, int out_dtype, int split_k, apb_stream_t stream) {
  return apb_gemm_tc_rowsum(A, B, C, bias, aux, M, N, K, trans_a, trans_b, epilogue, in_dtype, out_dtype, split_k, nullptr, stream);
}

// number of K splits the wgrad-shaped GEMM should use to fill the GPU (host helper for the binding)
int apb_gemm_tc_suggest_split(int M, int N, int K) {
  const int tiles = ceil_div(M, BM) * ceil_div(N, pick_bn(N));
  const int total_kb = (K + BK - 1) / BK;
  const int sms = 148;
  if (tiles >= sms || total_kb < 16) return 1;
  // choose the split count whose work items fill whole waves of the persistent grid best (>= 8 k-blocks per item)
  int best = 1;
  double best_score = 0.0;
  for (int s = 1; s <= 64 && s * 8 <= total_kb; ++s) {
    const int kb_per = (total_kb + s - 1) / s;
    if ((total_kb + kb_per - 1) / kb_per != s) continue;          // not realisable without an empty split
    const int items = tiles * s;
    const int waves = (items + sms - 1) / sms;
    const double eff = (double)items / (waves * sms);             // SM utilisation
    const double score = eff / (1.0 + 0.02 * s);                  // mild preference for fewer partials to reduce
    if (score > best_score) { best_score = score; best = s; }
  }
  return best;
}
