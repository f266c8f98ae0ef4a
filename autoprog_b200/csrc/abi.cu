// Error reporting for the C ABI (thread-local message buffer).
#include <stdlib.h>
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
// off by default: measured neutral on the whole step (profiles/r2_kernels.md "programmatic dependent launch")
static int pdl_from_env() { const char* e = getenv("APB_PDL"); return e && e[0] == '1'; }
int g_apb_pdl = pdl_from_env();
extern "C" void apb_set_pdl(int on) { g_apb_pdl = on ? 1 : 0; }
extern "C" int apb_get_pdl(void) { return g_apb_pdl; }
long long g_apb_fallbacks = 0;
extern "C" long long apb_fallback_count(void) { return g_apb_fallbacks; }
void apb_note_fallback(const char* what, const char* why) {
  ++g_apb_fallbacks;
  if (g_apb_fallbacks <= 8) fprintf(stderr, "[autoprog_b200] %s: tensor-core kernel declined (%s); running the CUDA-core kernel\n", what, why);
}
