// bf16 NT GEMM on CTA PAIRS: tcgen05.mma.cta_group::2 (one 256 x BN x 16 MMA spans two SMs of a TPC).
//
//   C[M,N] = A[M,K] . B[N,K]^T (+ bias[n])     A, B K-major (nn.Linear forward: activations x weight)
//
// Why: with one 128 x 192 tile per SM (gemm_tc.cu) every k-block needs 40 KB of operands for 0.21 us of tensor work,
// i.e. 76.8 FLOP per byte read from L2 -- the main loop is bound by L2->SM bandwidth, not by the tensor pipe.  A CTA
// pair shares the B tile: CTA r stages its own 128 rows of A and only HALF of the B rows (n0 + r*BN/2 ...); the MMA
// reads both halves.  256 x 192 per pair = 28 KB per SM per k-block for the same work: 110 FLOP/B.
//
// Structure = gemm_tc.cu (persistent, warp 0 TMA producer, warp 1 MMA issuer, 12 epilogue warps, 2 TMEM accumulator
// stages, swizzled staging + TMA store), plus the pair plumbing:
//   * both CTAs issue their own TMA loads (.cta_group::2) but complete_tx on the LEADER's full barrier (mapa address);
//     the leader's producer arms it with the bytes of both CTAs;
//   * only the leader (cluster rank 0) issues MMAs; tcgen05.commit ... multicast::cluster arrives on the empty / tmem-full
//     barriers of BOTH CTAs;
//   * each CTA's epilogue drains its own 128 TMEM lanes; "accumulator drained" arrivals of both CTAs go to the leader's
//     tmem-empty barrier (remote mbarrier.arrive through a mapa address);
//   * TMEM is allocated / freed with .cta_group::2 by warp 1 of both CTAs; cluster barriers bracket the kernel.
#include "gemm_tc_common.cuh"

namespace {

constexpr int BMC = 128;   // rows per CTA (256 per pair)
constexpr int BK2 = 64;
constexpr int ACC2 = 2;

struct Tc2Params {
  const float* bias;
  int M, N, K;
  int out_f32;
  int tiles_m, tiles_n;   // tiles_m counts 256-row pair tiles
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int BN, int STAGES, int EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + EPI_WARPS * 32, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, Tc2Params p) {
  constexpr uint32_t A_BYTES = BMC * BK2 * 2;          // 16 KB: this CTA's 128 rows of A
  constexpr uint32_t B_BYTES = (BN / 2) * BK2 * 2;     // this CTA's half of the B rows
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (ACC2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t STG_BYTES = 4096;
  constexpr int NCH = (BN / 32) / (EPI_WARPS / 4);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + EPI_WARPS * STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_kb = (p.K + BK2 - 1) / BK2;
  const int n_items = p.tiles_m * p.tiles_n;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_c) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote signal
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // the prologue above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer (BOTH CTAs; whole warp runs the loop, one elected lane issues) =====================
    const bool leader = elect_one();
    uint32_t it = 0;
    for (int item = pair; item < n_items; item += npairs) {
      const int m0 = (item / p.tiles_n) * (2 * BMC) + (int)rank * BMC;
      const int n0 = (item % p.tiles_n) * BN + (int)rank * (BN / 2);
      for (int kb = 0; kb < total_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        const uint32_t lead_full = mapa_rank(smem_u32(&full_bar[s]), 0);
        if (leader) {
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);      // bytes of both CTAs land on the leader's barrier
          tma_load_2d_pair(sa, &tma_a, lead_full, kb * BK2, m0);            // box {64 k, 128 m}
          tma_load_2d_pair(sb, &tma_b, lead_full, kb * BK2, n0);            // box {64 k, BN/2 n}
        }
        __syncwarp();
      }
      if (item + npairs >= n_items) pdl_trigger();     // last tile's loads in flight: the next kernel may start its prologue
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (LEADER CTA; whole warp runs the loop, one elected lane issues) =====================
    if (rank == 0) {
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BMC) >> 4) << 24);
      const uint32_t smem_a0 = smem_u32(smem);
      uint32_t it = 0, ai = 0;
      for (int item = pair; item < n_items; item += npairs, ++ai) {
        const uint32_t as = ai % ACC2;
        mbar_wait(&tempty_bar[as], ((ai / ACC2) & 1) ^ 1);           // both CTAs' epilogues have drained this stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_u + as * BN;
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_a0 + s * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t ad0 = make_smem_desc(sa, 16, 1024), bd0 = make_smem_desc(sb, 16, 1024);
          const uint32_t acc_first = kb > 0 ? 1u : 0u;
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK2 / 16; ++k) umma_bf16_pair(tacc, ad0 + k * 2, bd0 + k * 2, idesc, k > 0 ? 1u : acc_first);
            umma_commit_pair(&empty_bar[s]);    // frees this smem stage in both CTAs
            if (kb + 1 == total_kb) umma_commit_pair(&tfull_bar[as]);     // accumulator complete: wakes the epilogues of both CTAs
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 TMEM lanes -> registers -> swizzled smem -> TMA store =====================
    const int quarter = warp & 3;
    const int cg = (warp - 2) >> 2;
    uint8_t* stgC = stg_base + (warp - 2) * STG_BYTES;
    uint32_t ai = 0;
    for (int item = pair; item < n_items; item += npairs, ++ai) {
      const int m0 = (item / p.tiles_n) * (2 * BMC) + (int)rank * BMC, n0 = (item % p.tiles_n) * BN;
      const uint32_t as = ai % ACC2;
      const int rb = m0 + quarter * 32;
      const int cb0 = n0 + cg * NCH * 32;
      mbar_wait(&tfull_bar[as], (ai / ACC2) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)((cg * NCH + c) * 32), r[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_remote(mapa_rank(smem_u32(&tempty_bar[as]), 0));     // leader's barrier collects both CTAs
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
      if (rb >= p.M) continue;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int cb = cb0 + c * 32;
        if (cb >= p.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c][j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (cb + j < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cb + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
        }
        uint8_t* sc = p.out_f32 ? stgC : stgC + c * 2048;
        if (p.out_f32) {
          if (c > 0) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
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
        if (lane == 0) {
          tma_store_3d(&tma_c, sc, cb, rb, 0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  // neither CTA may leave (or free TMEM) while its partner can still signal its barriers or read its shared memory
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int STAGES, int EPI_WARPS>
int launch2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, Tc2Params p, cudaStream_t st) {
  constexpr size_t smem = STAGES * (BMC * BK2 * 2 + (BN / 2) * BK2 * 2) + EPI_WARPS * 4096 + 1024 + 256;
  static_assert(smem <= 227 * 1024, "gemm_tc2: shared memory budget");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BN, STAGES, EPI_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { apb_set_error("gemm_tc2: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  p.tiles_m = ceil_div(p.M, 2 * BMC);
  p.tiles_n = ceil_div(p.N, BN);
  const long long items = (long long)p.tiles_m * p.tiles_n;
  const int pairs_max = num_sms() / 2;
  const int pairs = (int)(items < pairs_max ? items : pairs_max);
  apb_launch_pdl(gemm_tc2_kernel<BN, STAGES, EPI_WARPS>, dim3(2 * pairs), dim3(64 + EPI_WARPS * 32), smem, st, ma, mb, mc, p);
  APB_LAUNCH_CHECK("gemm_tc2");
  return 0;
}

}  // namespace

// C[M,N] = A[M,K] B[N,K]^T (+ bias): both operands K-major bf16, C bf16 or fp32.  N % 8 == 0, K % 8 == 0.
int apb_gemm_tc_pair(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int out_dtype,
                     apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(out_dtype == APB_BF16 || out_dtype == APB_F32, APB_ERR_DTYPE, "gemm_tc_pair: out dtype %d", out_dtype);
  APB_CHECK_ARG(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, APB_ERR_SHAPE, "gemm_tc_pair: M=%d N=%d K=%d", M, N, K);
  APB_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0, APB_ERR_ARG,
                "gemm_tc_pair: pointers must be 16-byte aligned");
  const bool wide = ceil_div(N, 192) * 192 <= ceil_div(N, 128) * 128;
  const int BN = wide ? 192 : 128;
  CUtensorMap ma, mb, mc;
  int rc = make_map(&ma, A, M, K, BMC, BK2);
  if (rc) return rc;
  rc = make_map(&mb, B, N, K, BN / 2, BK2);
  if (rc) return rc;
  rc = make_map_out(&mc, C, out_dtype == APB_F32, M, N, 1);
  if (rc) return rc;
  Tc2Params p;
  p.bias = bias; p.M = M; p.N = N; p.K = K; p.out_f32 = (out_dtype == APB_F32) ? 1 : 0;
  if (wide) return launch2<192, 6, 12>(ma, mb, mc, p, st);
  return launch2<128, 6, 16>(ma, mb, mc, p, st);
}
