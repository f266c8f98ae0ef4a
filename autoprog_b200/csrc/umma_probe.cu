// Diagnostic entry point: one tcgen05.mma against operands laid out by the caller's choice of shared-memory layout,
// used (tools/umma_probe.py) to pin down descriptor conventions the attention kernels rely on before they are built on:
//   mode 0: B MN-major, 128B swizzle, N = 32 columns taken at a byte offset INSIDE the 128-byte row (head pairs)
//   mode 1: A and B K-major with 64-byte rows / 64B swizzle (one head of 32 channels per row)
//   mode 2: B MN-major with 64-byte rows / 64B swizzle
//   mode 3: A MN-major with 128-byte rows / 128B swizzle, M = 128 spanning two 64-element blocks, A rows = K
// D[128 x 32] (fp32) = A[128 x K] * B[K x 32]; the host compares with a plain matmul.  Not on any product path.
#include "gemm_tc_common.cuh"

namespace {

__device__ __forceinline__ uint64_t make_desc_l(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t sw128o(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t sw64o(int r, int c) { return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// A: [128][K] bf16 row-major (global), Bm: [K][64] bf16 row-major (global), D: [128][32] fp32; K = 64
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const bf16* __restrict__ A, const bf16* __restrict__ Bm, float* __restrict__ D,
                                                            int mode, int off_elems) {
  constexpr int K = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;              // 32 KB region
  uint8_t* sB = smem + 32768;      // 16 KB region
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- stage A
  if (mode == 0 || mode == 2) {
    // K-major, 128-byte rows (64 k per row), 128B swizzle: 128 rows x 8 chunks
    for (int e = tid; e < 128 * 8; e += 128) {
      const int r = e >> 3, c = e & 7;
      *reinterpret_cast<uint4*>(sA + sw128o(r, c)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
    }
  } else if (mode == 1) {
    // K-major, 64-byte rows: two K blocks of 32 (block stride 128 rows * 64 B = 8 KB), 64B swizzle
    for (int e = tid; e < 128 * 8; e += 128) {
      const int r = e >> 3, c = e & 7, blk = c >> 2, cc = c & 3;
      *reinterpret_cast<uint4*>(sA + blk * 8192 + sw64o(r, cc)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
    }
  } else {
    // mode 3: MN-major A: smem holds A^T as [K rows][128 m] = two blocks of 64 m (block stride K * 128 B), 128B swizzle
    for (int e = tid; e < K * 16; e += 128) {
      const int k = e >> 4, c = e & 15, blk = c >> 3, cc = c & 7;     // 8 m-values m = c*8 .. c*8+7 of k-row k
      uint4 v;
      bf16* h = reinterpret_cast<bf16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = A[(size_t)(c * 8 + j) * K + k];
      *reinterpret_cast<uint4*>(sA + blk * (K * 128) + sw128o(k, cc)) = v;
    }
  }
  // ---- stage B: Bm [K][64]
  if (mode == 0 || mode == 3) {
    for (int e = tid; e < K * 8; e += 128) {      // [k rows][128 B], 128B swizzle (MN-major B: n contiguous)
      const int r = e >> 3, c = e & 7;
      *reinterpret_cast<uint4*>(sB + sw128o(r, c)) = *reinterpret_cast<const uint4*>(Bm + (size_t)r * 64 + c * 8);
    }
  } else if (mode == 1) {
    // K-major B with 64-byte rows: B^T rows n = 0..31 (columns off..off+31 of Bm), k contiguous: two K blocks of 32
    for (int e = tid; e < 32 * 8; e += 128) {
      const int n = e >> 3, c = e & 7, blk = c >> 2, cc = c & 3;
      uint4 v;
      bf16* h = reinterpret_cast<bf16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = Bm[(size_t)(c * 8 + j) * 64 + off_elems + n];
      *reinterpret_cast<uint4*>(sB + blk * 2048 + sw64o(n, cc)) = v;
    }
  } else {
    // mode 2: MN-major B with 64-byte rows: [k rows][32 n] (columns off..off+31), 64B swizzle
    for (int e = tid; e < K * 4; e += 128) {
      const int r = e >> 2, c = e & 3;
      *reinterpret_cast<uint4*>(sB + sw64o(r, c)) = *reinterpret_cast<const uint4*>(Bm + (size_t)r * 64 + off_elems + c * 8);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t a = smem_u32(sA), b = smem_u32(sB);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad, bd;
      uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      if (mode == 0) {
        ad = make_desc_l(a + ks * 32, 16, 1024, 2);
        bd = make_desc_l(b + ks * 2048 + off_elems * 2, 64 * 128, 1024, 2);
        idesc |= (1u << 16);
      } else if (mode == 1) {
        ad = make_desc_l(a + (ks >> 1) * 8192 + (ks & 1) * 32, 16, 512, 4);
        bd = make_desc_l(b + (ks >> 1) * 2048 + (ks & 1) * 32, 16, 512, 4);
      } else if (mode == 2) {
        ad = make_desc_l(a + ks * 32, 16, 1024, 2);
        bd = make_desc_l(b + ks * 1024, 16, 512, 4);
        idesc |= (1u << 16);
      } else {
        ad = make_desc_l(a + ks * 2048, K * 128, 1024, 2);
        bd = make_desc_l(b + ks * 2048 + off_elems * 2, 64 * 128, 1024, 2);
        idesc |= (1u << 15) | (1u << 16);
      }
      umma_bf16(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) D[(size_t)tid * 32 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32) : "memory");
}

}  // namespace

extern "C" int apb_debug_umma_probe(const void* A, const void* Bm, float* D, int mode, int off_elems, apb_stream_t stream) {
  APB_CHECK_ARG(mode >= 0 && mode <= 3 && (off_elems == 0 || off_elems == 32), APB_ERR_ARG, "umma_probe: mode %d off %d", mode, off_elems);
  const size_t smem = 49152 + 64 + 1024;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_probe_kernel<<<1, 128, smem, APB_STREAM(stream)>>>((const bf16*)A, (const bf16*)Bm, D, mode, off_elems);
  APB_LAUNCH_CHECK("umma_probe");
  return 0;
}
