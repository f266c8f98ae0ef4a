// Error reporting for the C ABI (thread-local message buffer).
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void apb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* apb_last_error(void) { return g_err; }
extern "C" int apb_abi_version(void) { return 1; }

long long g_apb_launches = 0;
extern "C" long long apb_launch_count(void) { return g_apb_launches; }

// bf16 calls that fell through to a CUDA-core kernel because the tensor-core kernel declined the shape
// Diagnostic switches: plain ints set through the ABI by tools (never read from the environment, never touched by the
// training path).  g_apb_pdl: programmatic dependent launch, off by default -- measured neutral on the whole step
// (profiles/r2_kernels.md).  g_apb_gemm_dbg: bit 0 no TMA loads, bit 1 no MMAs, bit 2 no stores (tools/gemm_bound.py).
// g_apb_gemm_narrow: 1 = five-stage 128 x 192 GEMM kernels (default), 0 = the four-stage ones (A/B runs).
int g_apb_pdl = 0;
int g_apb_gemm_dbg = 0;
int g_apb_gemm_narrow = 1;
extern "C" void apb_set_pdl(int on) { g_apb_pdl = on ? 1 : 0; }
extern "C" int apb_get_pdl(void) { return g_apb_pdl; }
extern "C" void apb_debug_gemm_switches(int dbg, int five_stage) { g_apb_gemm_dbg = dbg; g_apb_gemm_narrow = five_stage ? 1 : 0; }
long long g_apb_fallbacks = 0;
extern "C" long long apb_fallback_count(void) { return g_apb_fallbacks; }
void apb_note_fallback(const char* what, const char* why) {
  ++g_apb_fallbacks;
  if (g_apb_fallbacks <= 8) fprintf(stderr, "[autoprog_b200] %s: tensor-core kernel declined (%s); running the CUDA-core kernel\n", what, why);
}
