// Multi-head self-attention FORWARD on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 32, N <= 224 keys
// (every VOLO stage-2 grid of the progressive schedule: N = 64 ... 196, DeiT-style 197).
//
//   reference: attn = softmax(q k^T * scale); out = attn @ v            (models/volo.py:193-197)
//
// One CTA = one (batch, head), 128 threads; thread t owns query row t of the current 128-row tile, which is TMEM lane t:
//   S = Q K^T     : tcgen05.mma (M = 128, N = Npad, two K = 16 steps), accumulator S[128 x Npad] fp32 in TMEM
//   softmax       : all keys fit ONE accumulator tile, so it is a plain two-pass row softmax -- each thread reads ITS row
//                   with tcgen05.ld (no shuffles, no online rescaling), writes P = exp2(..) back to TMEM as packed bf16
//                   (tcgen05.st) over the columns of S it has already consumed
//   O = P V       : tcgen05.mma with the A operand read from TMEM (P) and V^T (K-major, staged transposed) from shared
//                   memory; O[128 x 32] fp32 in TMEM columns 224..255
//   out = O / l   : tcgen05.ld, one 64-byte row store per thread; lse = m * scale + ln(l) saved for the backward.
// Against the mma.sync formulation (attention_mma.cu: ~117 warp instructions per 16 x 16 score block, fragment
// shuffles, ldmatrix) the per-score work is ~5 thread instructions; the kernel is bound by the exp2 throughput.
// Operands are staged by the threads themselves into the 128-byte-swizzled K-major layout the UMMA descriptors of
// gemm_tc.cu use (rows are 64 bytes of data in a 128-byte pitch).
//
// STATUS: correct (parity tests) but NOT the default: 124 us against 72 us for attention_mma.cu at B=128, N=196, 12 heads.
// TMEM (512 columns / SM, 208 + 32 needed per 128 query rows) caps an SM at 256 resident query rows = 8 warps, and this
// first version serialises staging -> S MMA -> two softmax passes -> PV MMA -> store inside each CTA.  The exp2 work
// alone is ~21 us at the MUFU rate.  Next: one CTA per head PAIR (q/k/v of neighbouring heads are 128 contiguous bytes:
// no transposed V staging, V as an MN-major operand), two independent 128-thread halves per CTA so one half's MMAs
// overlap the other's softmax, tcgen05.ld of chunk c+1 issued before chunk c is processed, next CTA staging while it
// waits for TMEM.  Select with APB_MHSA_TC=1.
#include "gemm_tc_common.cuh"

namespace {

constexpr int HD = 32;            // head dim
constexpr int QT = 128;           // query rows per tile (UMMA M)
constexpr int O_COL = 224;        // TMEM column of the O accumulator (S / P use columns [0, Npad))
constexpr int TMEM_COLS_ATT = 256;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk `c` (0..7) of row `r` in a K-major tile with 128-byte rows and the 128B swizzle
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(128, 2) mhsa_fwd_tc_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                             float* __restrict__ lse, int N, int heads, float scale, int Npad) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                          // [2 * 128 rows][128 B]
  uint8_t* sK = sQ + 2 * QT * 128;             // [Npad (<= 224) rows][128 B]
  uint8_t* sVt = sK + 224 * 128;               // 4 key blocks of [32 rows (channels)][64 keys * 2 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sVt + 4 * 4096);   // [0]: S ready, [1]: O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int bh = blockIdx.x, b = bh / heads, hd = bh % heads;
  const size_t tok = (size_t)3 * heads * HD;
  const bf16* qb = qkv + (size_t)b * N * tok + (size_t)hd * HD;
  const bf16* kb = qb + (size_t)heads * HD;
  const bf16* vb = kb + (size_t)heads * HD;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS_ATT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- stage Q (2 tiles), K, V^T; rows / keys beyond N are zero
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int e = tid; e < 2 * QT * 4; e += 128) {
    const int r = e >> 2, c = e & 3;
    const uint4 v = (r < N) ? *reinterpret_cast<const uint4*>(qb + (size_t)r * tok + c * 8) : zero4;
    *reinterpret_cast<uint4*>(sQ + sw128(r, c)) = v;
  }
  for (int e = tid; e < Npad * 4; e += 128) {
    const int r = e >> 2, c = e & 3;
    const uint4 v = (r < N) ? *reinterpret_cast<const uint4*>(kb + (size_t)r * tok + c * 8) : zero4;
    *reinterpret_cast<uint4*>(sK + sw128(r, c)) = v;
  }
  for (int e = tid; e < Npad * 4; e += 128) {
    const int key = e >> 2, c = e & 3;                 // 8 channels c*8 .. c*8+7 of one key
    const uint4 v = (key < N) ? *reinterpret_cast<const uint4*>(vb + (size_t)key * tok + c * 8) : zero4;
    const uint16_t* h = reinterpret_cast<const uint16_t*>(&v);
    uint8_t* blk = sVt + (key >> 6) * 4096;            // key block of 64
    const int col = key & 63;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint16_t*>(blk + sw128(c * 8 + j, col >> 3) + (col & 7) * 2) = h[j];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's TMEM lane quarter

  const uint32_t idescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Npad >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
  const uint32_t idescO = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
  const float sl2 = scale * 1.4426950408889634f;
  const int nchunk = (Npad + 31) >> 5;
  const int ntiles = (N + QT - 1) / QT;

  for (int qt = 0; qt < ntiles; ++qt) {
    const uint32_t ph = (uint32_t)(qt & 1);
    if (tid == 0) {
      // ---- S = Q_tile K^T
      const uint32_t qa = smem_u32(sQ + qt * QT * 128), ka = smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        umma_bf16(tmem_base, make_smem_desc(qa + ks * 32, 16, 1024), make_smem_desc(ka + ks * 32, 16, 1024), idescS, ks > 0 ? 1u : 0u);
      umma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], ph);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- pass 1: row maximum
    float m = -INFINITY;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t r[32];
      tmem_ld32(lane_addr + (uint32_t)(c * 32), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < N) m = fmaxf(m, __uint_as_float(r[j]));
    }
    // ---- pass 2: P = exp2((S - m) * scale * log2e) -> bf16 pairs back into TMEM (columns [0, Npad / 2))
    const float msc = m * sl2;
    float l = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t r[32];
      tmem_ld32(lane_addr + (uint32_t)(c * 32), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float p0 = ex2f(fmaf(__uint_as_float(r[j]), sl2, -msc)), p1 = ex2f(fmaf(__uint_as_float(r[j + 1]), sl2, -msc));
        if (c * 32 + j >= N) p0 = 0.f;
        if (c * 32 + j + 1 >= N) p1 = 0.f;
        l += p0 + p1;
        __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
        pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st16(lane_addr + (uint32_t)(c * 16), pk);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      // ---- O = P V : A from TMEM (8 columns = 16 bf16 keys per step), B = V^T tile (K-major over keys)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t va = smem_u32(sVt);
      const int ksteps = Npad >> 4;
      for (int kk = 0; kk < ksteps; ++kk)
        umma_bf16_ts(tmem_base + O_COL, tmem_base + (uint32_t)(kk * 8),
                     make_smem_desc(va + (kk >> 2) * 4096 + (kk & 3) * 32, 16, 1024), idescO, kk > 0 ? 1u : 0u);
      umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], ph);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      uint32_t r[32];
      tmem_ld32(lane_addr + O_COL, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = qt * QT + tid;
      if (row < N) {
        const float inv = 1.f / l;
        bf16* orow = out + ((size_t)b * N + row) * heads * HD + (size_t)hd * HD;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4 pkv;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pkv);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            h2[t] = __floats2bfloat162_rn(__uint_as_float(r[c4 * 8 + 2 * t]) * inv, __uint_as_float(r[c4 * 8 + 2 * t + 1]) * inv);
          *reinterpret_cast<uint4*>(orow + c4 * 8) = pkv;
        }
        lse[((size_t)b * heads + hd) * N + row] = m * scale + logf(l);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                   // S / P / O columns are reused by the next tile
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS_ATT) : "memory");
}

}  // namespace

// returns APB_ERR_UNSUPPORTED for shapes outside the single-tile envelope (the caller falls back to attention_mma.cu)
int apb_mhsa_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  if (D != HD || N < 1 || N > 224 || (((uintptr_t)qkv | (uintptr_t)out) & 15) != 0) return APB_ERR_UNSUPPORTED;
  const int Npad = (N + 15) / 16 * 16;
  const size_t smem = (size_t)2 * QT * 128 + 224 * 128 + 4 * 4096 + 64 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mhsa_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { apb_set_error("mhsa_fwd_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  mhsa_fwd_tc_kernel<<<B * heads, 128, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, N, heads, scale, Npad);
  APB_LAUNCH_CHECK("mhsa_fwd_tc");
  return 0;
}
