// Diagnostic: cycles per tcgen05.mma for the operand layouts of the attention kernels (tools/umma_timing.py).
// One CTA, one issuing thread: `reps` back-to-back MMAs of one kind, then commit + wait; clock64 around it.
#include "gemm_tc_common.cuh"

namespace {
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma_ts_(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d), "r"(tmem_a),
               "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}

// kind: 0 SS K-major SW64 (A 128 x 16, B N x 16)          -- S = Q K^T style, N given
//       1 TS, B MN-major SW64 N = 32                      -- O = P V style
//       2 SS A MN-major SW64 (M = 128), B MN-major SW64   -- dV = P^T dO style
//       3 SS A K-major SW64, B MN-major SW64 N = 32       -- dQ = dS K style
//       4 SS K-major SW128 (gemm style), N given
//       5 SS A MN-major SW128, B MN-major SW128, N given  -- wgrad style
__global__ void __launch_bounds__(128, 1) umma_timing_kernel(long long* out, int kind, int N, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;   // finite bf16 pairs
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = __shfl_sync(0xffffffffu, *slot, 0);
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t ad = 0, bd = 0;
    if (kind == 0) { ad = mkdesc(a, 16, 512, 4); bd = mkdesc(b, 16, 512, 4); }
    if (kind == 1) { bd = mkdesc(b, 16, 512, 4); idesc |= 1u << 16; }
    if (kind == 2) { ad = mkdesc(a, 8192, 512, 4); bd = mkdesc(b, 16, 512, 4); idesc |= (1u << 15) | (1u << 16); }
    if (kind == 3) { ad = mkdesc(a, 16, 512, 4); bd = mkdesc(b, 16, 512, 4); idesc |= 1u << 16; }
    if (kind == 4) { ad = mkdesc(a, 16, 1024, 2); bd = mkdesc(b, 16, 1024, 2); }
    if (kind == 5) { ad = mkdesc(a, 8192, 1024, 2); bd = mkdesc(b, 8192, 1024, 2); idesc |= (1u << 15) | (1u << 16); }
    for (int rep = 0; rep < 2; ++rep) {     // rep 0 warms up
      const long long t0 = clock64();
      if (leader) {
        for (int i = 0; i < reps; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (kind == 1) umma_ts_(tm + 256, tm + (uint32_t)(u * 8), bd + (uint64_t)(u * 64), idesc, 1u);
            else umma_bf16(tm + 256, ad + (uint64_t)(u * 2), bd + (uint64_t)(u * 2), idesc, 1u);
          }
        }
      }
      __syncwarp();
      const long long t1 = clock64();
      if (leader) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      if (leader) {
        out[rep * 2] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
}  // namespace

extern "C" int apb_debug_umma_timing(long long* out4, int kind, int N, int reps, apb_stream_t stream) {
  const size_t smem = 160 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(umma_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_timing_kernel<<<1, 128, smem, APB_STREAM(stream)>>>(out4, kind, N, reps);
  APB_LAUNCH_CHECK("umma_timing");
  return 0;
}
