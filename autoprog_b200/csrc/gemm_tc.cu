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
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue
// (one TMEM lane quarter each).  Tile 128 x BN x 64, 3-stage mbarrier ring, 2 CTAs per SM so one CTA's epilogue
// overlaps the other's main loop (K is short on this path: 3..18 k-blocks).  split_k > 1 writes fp32 partial tiles
// [split][M][N] that the caller reduces in fixed order (deterministic wgrad).
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace {

constexpr int BM = 128, BK = 64, STAGES = 3;
constexpr int NTHREADS = 192;

struct TcParams {
  void* C;
  const float* bias;
  void* aux;
  int M, N, K;
  int a_mn, b_mn;       // operand majors (1 = MN-major)
  int epilogue;         // 0 none, 1 gelu, 2 dgelu, 4 split-k partial
  int out_f32;
  int kb_per_split;     // k-blocks per blockIdx.z
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                             const __grid_constant__ CUtensorMap tma_b, TcParams p) {
  constexpr uint32_t A_BYTES = BM * BK * 2;   // 16 KB
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(total_kb, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;   // >= 1 by construction on the host

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(&empty_bar[s], ((i / STAGES) - 1) & 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int k0 = (kb0 + i) * BK;
        if (!p.a_mn) {
          tma_load_2d(sa, &tma_a, &full_bar[s], k0, m0);                 // box {64 k, 128 m}
        } else {
#pragma unroll
          for (int h = 0; h < BM / 64; ++h) tma_load_2d(sa + h * (BK * 128), &tma_a, &full_bar[s], m0 + h * 64, k0);  // box {64 m, 64 k}
        }
        if (!p.b_mn) {
          tma_load_2d(sb, &tma_b, &full_bar[s], k0, n0);                 // box {64 k, BN n}
        } else {
#pragma unroll
          for (int h = 0; h < BN / 64; ++h) tma_load_2d(sb + h * (BK * 128), &tma_b, &full_bar[s], n0 + h * 64, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? 16 * 128 : 32, b_kstep = p.b_mn ? 16 * 128 : 32;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbar_wait(&full_bar[s], (i / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
          const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);   // frees the smem stage once the MMAs above have read it
      }
      umma_commit(tmem_full_bar);     // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const int row = m0 + quarter * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool row_ok = row < p.M;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), r);
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const int ncol = min(32, p.N - col0);
      if (p.epilogue == 4) {
        float* dst = reinterpret_cast<float*>(p.C) + ((size_t)blockIdx.z * p.M + row) * p.N + col0;
        if (ncol == 32 && (p.N & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < ncol; ++j) dst[j] = v[j];
        }
        continue;
      }
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncol) v[j] += __ldg(p.bias + col0 + j);
      }
      const size_t off = (size_t)row * p.N + col0;
      if (p.out_f32) {
        float* dst = reinterpret_cast<float*>(p.C) + off;
        if (p.epilogue == 1) {
          float* ax = reinterpret_cast<float*>(p.aux) + off;
          for (int j = 0; j < ncol; ++j) { ax[j] = v[j]; v[j] = gelu_f(v[j]); }
        } else if (p.epilogue == 2) {
          const float* ax = reinterpret_cast<const float*>(p.aux) + off;
          for (int j = 0; j < ncol; ++j) v[j] *= dgelu_f(ax[j]);
        }
        if (ncol == 32 && (p.N & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < ncol; ++j) dst[j] = v[j];
        }
      } else {
        bf16* dst = reinterpret_cast<bf16*>(p.C) + off;
        const bool vec = (ncol == 32) && ((p.N & 7) == 0);
        if (p.epilogue == 1) {
          bf16* ax = reinterpret_cast<bf16*>(p.aux) + off;
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
              *reinterpret_cast<uint4*>(ax + j) = pk;
#pragma unroll
              for (int t = 0; t < 4; ++t) {   // GELU of the ROUNDED pre-activation (what backward will re-read)
                v[j + 2 * t] = gelu_f(__bfloat162float(h2[t].x));
                v[j + 2 * t + 1] = gelu_f(__bfloat162float(h2[t].y));
              }
            }
          } else {
            for (int j = 0; j < ncol; ++j) {
              const bf16 pre = __float2bfloat16_rn(v[j]);
              ax[j] = pre;
              v[j] = gelu_f(__bfloat162float(pre));
            }
          }
        } else if (p.epilogue == 2) {
          const bf16* ax = reinterpret_cast<const bf16*>(p.aux) + off;
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 pk = *reinterpret_cast<const uint4*>(ax + j);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                v[j + 2 * t] *= dgelu_f(__bfloat162float(h2[t].x));
                v[j + 2 * t + 1] *= dgelu_f(__bfloat162float(h2[t].y));
              }
            }
          } else {
            for (int j = 0; j < ncol; ++j) v[j] *= dgelu_f(__bfloat162float(ax[j]));
          }
        }
        if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
            *reinterpret_cast<uint4*>(dst + j) = pk;
          }
        } else {
          for (int j = 0; j < ncol; ++j) dst[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: TMA descriptors through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 2-D bf16 tensor [rows, cols] row-major, box {box_cols (<= 64 elements = 128 B), box_rows}, 128B swizzle, zero OOB fill
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { apb_set_error("gemm_tc: cuTensorMapEncodeTiled entry point unavailable"); return APB_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    apb_set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box=%dx%d base=%p", (int)r, rows, cols,
                  box_rows, box_cols, base);
    return APB_ERR_ARG;
  }
  return 0;
}

template <int BN>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int splits, cudaStream_t st) {
  constexpr size_t smem = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { apb_set_error("gemm_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), splits);
  gemm_tc_kernel<BN><<<grid, NTHREADS, smem, st>>>(ma, mb, p);
  APB_LAUNCH_CHECK("gemm_tc");
  return 0;
}

}  // namespace

// split_k > 1: C must hold split_k fp32 partials [split_k][M][N]; bias/epilogue are ignored.
int apb_gemm_tc(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                int trans_b, int epilogue, int in_dtype, int out_dtype, int split_k, apb_stream_t stream) {
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
  const int BN = 128;
  if (!trans_b) rc = make_map(&mb, B, N, K, BN, BK); else rc = make_map(&mb, B, K, N, BK, 64);
  if (rc) return rc;
  TcParams p;
  p.C = C; p.bias = bias; p.aux = aux; p.M = M; p.N = N; p.K = K; p.a_mn = trans_a ? 1 : 0; p.b_mn = trans_b ? 1 : 0;
  p.epilogue = splits > 1 ? 4 : epilogue;
  p.out_f32 = (out_dtype == APB_F32) ? 1 : 0;
  p.kb_per_split = kb_per;
  if (splits > 1) APB_CHECK_ARG(out_dtype == APB_F32, APB_ERR_DTYPE, "gemm_tc: split-K partials are fp32");
  return launch<128>(ma, mb, p, splits, st);
}

// number of K splits the wgrad-shaped GEMM should use to fill the GPU (host helper for the binding)
int apb_gemm_tc_suggest_split(int M, int N, int K) {
  const int tiles = ceil_div(M, BM) * ceil_div(N, 128);
  const int total_kb = (K + BK - 1) / BK;
  if (tiles >= 148 || total_kb < 8) return 1;
  int s = (2 * 148 + tiles - 1) / tiles;
  if (s > total_kb / 4) s = total_kb / 4;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  int kb_per = (total_kb + s - 1) / s;
  return (total_kb + kb_per - 1) / kb_per;
}
