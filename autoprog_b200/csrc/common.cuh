// Shared device/host helpers for the autoprog_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/autoprog_b200.h"

// error codes returned through the C ABI (0 = ok, >0 = cudaError_t, <0 = argument check)
#define APB_ERR_ARG (-1)
#define APB_ERR_DTYPE (-2)
#define APB_ERR_SHAPE (-3)
#define APB_ERR_UNSUPPORTED (-4)

void apb_set_error(const char* fmt, ...);
#define APB_STREAM(s) ((cudaStream_t)(s))

#define APB_CHECK_ARG(cond, code, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      apb_set_error(__VA_ARGS__);                 \
      return (code);                              \
    }                                             \
  } while (0)

void apb_note_fallback(const char* what, const char* why);   // counted + logged: a bf16 call served by a CUDA-core kernel
extern long long g_apb_launches;   // kernel launches issued through the library (not thread-exact; evidence counter)
#define APB_LAUNCH_CHECK(name)                                              \
  do {                                                                      \
    ++g_apb_launches;                                                       \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      apb_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                      \
    }                                                                       \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// A kernel launched through apb_launch_pdl may begin (block scheduling, barrier / TMEM setup, tensor-map prefetch) while
// its predecessor in the stream is still draining; it must execute pdl_wait() BEFORE its first global-memory access (the
// wait returns once every prerequisite grid has completed and its writes are visible).  pdl_trigger() lets the NEXT
// kernel in the stream start its own prologue; it is placed after the wait, so at most two grids are in flight.
// Both instructions are no-ops for a kernel launched without the attribute.  g_apb_pdl (apb_set_pdl): off by default.
extern int g_apb_pdl, g_apb_gemm_dbg, g_apb_gemm_narrow;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t apb_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_apb_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return reinterpret_cast<const float*>(&raw)[i]; }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<float*>(&raw)[i] = v; }
  __device__ __forceinline__ void zero() { raw = make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const {
    return __bfloat162float(reinterpret_cast<const bf16*>(&raw)[i]);
  }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<bf16*>(&raw)[i] = __float2bfloat16_rn(v); }
  __device__ __forceinline__ void zero() { raw = make_uint4(0u, 0u, 0u, 0u); }
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
