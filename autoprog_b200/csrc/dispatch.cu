// dtype dispatch for ops that have both a tensor-core bf16 kernel and a CUDA-core fp32 kernel.
#include "common.cuh"

int apb_outlook_fwd_mma(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale,
                        cudaStream_t st);
int apb_outlook_bwd_mma(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                        int heads, float scale, cudaStream_t st);

int apb_outlook_fwd(const void* v, const void* logits, void* y, int B, int H, int W, int heads, float scale, int dtype,
                    apb_stream_t stream) {
  return apb_outlook_fwd_simt(v, logits, y, B, H, W, heads, scale, dtype, stream);
}

int apb_outlook_bwd(const void* v, const void* logits, const void* dy, void* dv, void* dlogits, int B, int H, int W,
                    int heads, float scale, int dtype, apb_stream_t stream) {
  return apb_outlook_bwd_simt(v, logits, dy, dv, dlogits, B, H, W, heads, scale, dtype, stream);
}
