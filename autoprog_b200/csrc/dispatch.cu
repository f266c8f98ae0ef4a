// dtype dispatch for ops that have both a tensor-core bf16 kernel and a CUDA-core fp32 kernel.
#include "common.cuh"



// bf16 -> tensor-core kernels (outlook_mma.cu); fp32, or shapes whose band does not fit shared memory -> SIMT kernels
int apb_outlook_fwd(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int lpitch,
                    int dtype, apb_stream_t stream) {
  if (dtype == APB_BF16 && B > 0 && H > 0 && W > 0 && heads > 0 && lpitch >= heads * 81 && lpitch < heads * 81 + 8 &&
      (((uintptr_t)v | (uintptr_t)y) & 15) == 0) {
    // two bf16 kernels: the gather formulation on the CUDA cores (outlook_fma.cu) is the faster one on every AutoProg grid
    // (measured at B = 128, 6 heads: 16x16 29 vs 44 us, 20x20 42 vs 73, 24x24 59 vs 96, 28x28 78 vs 128); the mma.sync
    // fragments + staged fold (outlook_mma.cu) cover the tiles it declines (unaligned logits rows, oversized bands)
    int rc = apb_outlook_fwd_fma(v, logits, y, B, H, W, heads, scale, lpitch, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
    rc = apb_outlook_fwd_mma(v, logits, y, B, H, W, heads, scale, lpitch, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype == APB_BF16) apb_note_fallback("outlook_fwd", "shape / alignment outside the tensor-core kernel's envelope");
  return apb_outlook_fwd_simt(v, logits, y, B, H, W, heads, scale, lpitch, dtype, stream);
}

int apb_outlook_bwd(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                    int heads, float scale, int lpitch, int dtype, apb_stream_t stream) {
  if (dtype == APB_BF16 && B > 0 && H > 0 && W > 0 && heads > 0 && lpitch >= heads * 81 && lpitch < heads * 81 + 8 &&
      (((uintptr_t)v | (uintptr_t)dy | (uintptr_t)dv) & 15) == 0) {
    // gather kernel first (outlook_bwd_fma.cu), the band kernel (outlook_mma.cu) covers the tiles it declines
    int rc = apb_outlook_bwd_fma(v, logits, dy, dv, dlogits, B, H, W, heads, scale, lpitch, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
    rc = apb_outlook_bwd_mma(v, logits, dy, dv, dlogits, B, H, W, heads, scale, lpitch, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype == APB_BF16) apb_note_fallback("outlook_bwd", "shape / alignment outside the tensor-core kernel's envelope");
  return apb_outlook_bwd_simt(v, logits, dy, dv, dlogits, B, H, W, heads, scale, lpitch, dtype, stream);
}

// bf16, head_dim 32, N <= 224 (every VOLO stage-2 grid up to 224 px) -> tcgen05 / TMEM / TMA kernels (attention_tc.cu),
// forward and backward; bf16 with head_dim 64 (DeiT) or longer sequences (volo_d2 @ 384: N = 576) -> flash-style mma.sync
// kernels (attention_mma.cu); fp32 -> CUDA-core parity kernels.  A bf16 call that ends on the CUDA cores is counted and
// logged (apb_fallback_count).
int apb_mhsa_fwd(const void* qkv, void* out, float* lse, int B, int N, int heads, int D, float scale, int dtype,
                 apb_stream_t stream) {
  // (forward only: below ~80 tokens a 128-row query tile is mostly padding and the mma.sync kernel wins -- 14.7 vs 18.9 us at
  //  N = 64, the 8 x 8 grid of the first AutoProg stage; from N = 100 on the tcgen05 kernel is ahead, 25.0 vs 34.3 us)
  if (dtype == APB_BF16 && D == 32 && B > 0 && N >= 80 && N <= 224 && heads > 0) {
    const int rc = apb_mhsa_fwd_tc(qkv, out, lse, B, N, heads, D, scale, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype == APB_BF16 && (D == 32 || D == 64) && B > 0 && N > 0 && heads > 0) {
    const int rc = apb_mhsa_fwd_mma(qkv, out, lse, B, N, heads, D, scale, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype == APB_BF16) apb_note_fallback("mhsa_fwd", "head_dim not 32 / 64 or N too large for shared memory");
  return apb_mhsa_fwd_simt(qkv, out, lse, B, N, heads, D, scale, dtype, stream);
}

int apb_mhsa_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* workspace,
                 int B, int N, int heads, int D, float scale, int dtype, apb_stream_t stream) {
  if (dtype == APB_BF16 && D == 32 && B > 0 && N > 0 && N <= 224 && heads > 0) {
    const int rc = apb_mhsa_bwd_tc(qkv, out, dout, lse, dqkv, workspace, B, N, heads, D, scale, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype == APB_BF16 && (D == 32 || D == 64) && B > 0 && N > 0 && heads > 0) {
    const int rc = apb_mhsa_bwd_mma(qkv, out, dout, lse, dqkv, workspace, B, N, heads, D, scale, stream);
    if (rc != APB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype == APB_BF16) apb_note_fallback("mhsa_bwd", "head_dim not 32 / 64 or N too large for shared memory");
  return apb_mhsa_bwd_simt(qkv, out, dout, lse, dqkv, workspace, B, N, heads, D, scale, dtype, stream);
}
