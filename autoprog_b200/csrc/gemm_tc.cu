// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
int apb_gemm_tc(const void* A, const void* B, void* C, const float* bias, void* aux, int M, int N, int K, int trans_a,
                int trans_b, int epilogue, int in_dtype, int out_dtype, apb_stream_t stream) {
  apb_set_error("gemm_tc: not built yet");
  return APB_ERR_UNSUPPORTED;
}
