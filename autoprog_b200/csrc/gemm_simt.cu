// CUDA-core GEMM with exact fp32 products/accumulation: the fp32 PARITY path for every Linear / patchify conv
// (tcgen05 has no true-fp32 MMA; kind::tf32 is ~1e-3, the north-star fp32 tolerance is 1e-5 -- SURVEY.md §7).
// Also used to validate the tcgen05 bf16 kernel (same epilogues, bf16 inputs are exact in fp32).
//   C[M,N] = epi( sum_k A(m,k) * B(n,k) + bias[n] )      see include/autoprog_b200.h for the flags.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TI* __restrict__ A, const TI* __restrict__ B, TO* __restrict__ C,
                                                        const float* __restrict__ bias, TO* __restrict__ aux, int M, int N,
                                                        int K, long long sa_m, long long sa_k, long long sb_n,
                                                        long long sb_k, int epilogue) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (sa_k == 1) { k = t & 15; m = (t >> 4) + 16 * i; }
      else { m = t & 63; k = (t >> 6) + 4 * i; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? to_f(A[(long long)gm * sa_m + (long long)gk * sa_k]) : 0.f;
      int n, kb;
      if (sb_k == 1) { kb = t & 15; n = (t >> 4) + 16 * i; }
      else { n = t & 63; kb = (t >> 6) + 4 * i; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < N && gkb < K) ? to_f(B[(long long)gn * sb_n + (long long)gkb * sb_k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const size_t o = (size_t)m * N + n;
      float v = acc[i][j] + (bias != nullptr ? bias[n] : 0.f);
      if (epilogue == 1) {
        aux[o] = from_f<TO>(dgelu_f(v));          // gelu'(pre-activation): all the backward epilogue needs
        C[o] = from_f<TO>(gelu_f(v));
      } else if (epilogue == 2) {
        C[o] = from_f<TO>(v * to_f(aux[o]));
      } else if (epilogue == 3) {
        C[o] = from_f<TO>(to_f(C[o]) + v);
      } else {
        C[o] = from_f<TO>(v);
      }
    }
  }
}

}  // namespace

int apb_gemm_simt(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                  int trans_b, int epilogue, int in_dtype, int out_dtype, apb_stream_t stream) {
  cudaStream_t st = APB_STREAM(stream);
  APB_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, APB_ERR_SHAPE, "gemm_simt: M=%d N=%d K=%d", M, N, K);
  APB_CHECK_ARG(epilogue >= 0 && epilogue <= 3, APB_ERR_ARG, "gemm_simt: epilogue %d", epilogue);
  APB_CHECK_ARG((epilogue != 1 && epilogue != 2) || aux != nullptr, APB_ERR_ARG, "gemm_simt: GELU epilogues need aux");
  if (M == 0 || N == 0) return 0;
  const long long sa_m = trans_a ? 1 : K, sa_k = trans_a ? M : 1;
  const long long sb_n = trans_b ? 1 : K, sb_k = trans_b ? N : 1;
  dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
#define GS(TI_, TO_)                                                                                              \
  gemm_simt_kernel<TI_, TO_><<<grid, 256, 0, st>>>((const TI_*)A, (const TI_*)B, (TO_*)C, bias, (TO_*)aux, M, N, K, sa_m, \
                                                   sa_k, sb_n, sb_k, epilogue)
  if (in_dtype == APB_F32 && out_dtype == APB_F32) GS(float, float);
  else if (in_dtype == APB_BF16 && out_dtype == APB_BF16) GS(bf16, bf16);
  else if (in_dtype == APB_BF16 && out_dtype == APB_F32) GS(bf16, float);
  else if (in_dtype == APB_F32 && out_dtype == APB_BF16) GS(float, bf16);
  else { apb_set_error("gemm_simt: dtypes %d/%d", in_dtype, out_dtype); return APB_ERR_DTYPE; }
#undef GS
  APB_LAUNCH_CHECK("gemm_simt");
  return 0;
}
