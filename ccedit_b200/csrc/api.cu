// C-ABI plumbing shared by all kernels: error string, ABI version, launch counter.
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace ccedit {

std::atomic<long long> g_launch_count{0};
long long* g_trace_buf = nullptr;   // diagnostics buffer set by ccedit_gemm_trace (tap-GEMM and tcgen05 attention)

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace ccedit

extern "C" const char* ccedit_last_error(void) { return ccedit::g_err; }
extern "C" int ccedit_abi_version(void) { return 3; }
extern "C" int64_t ccedit_launch_count(void) { return ccedit::g_launch_count.load(); }
