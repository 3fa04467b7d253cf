// abi.cu — error reporting, launch accounting and version for the C-ABI (include/mamba_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace mb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();  // clear the sticky launch error so the next call starts clean
    return set_error(MAMBA_ELAUNCH, "%s: %s", what, cudaGetErrorString(e));
  }
  return MAMBA_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MAMBA_B200_PDL");   // opt-in: measured 7 % SLOWER on the decode chain (profiles/r02_decode.md)
    return e && e[0] == '1';
  }();
  return on;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace mb

extern "C" int mamba_abi_version(void) { return MAMBA_ABI_VERSION; }
extern "C" const char* mamba_last_error(void) { return mb::g_err; }
extern "C" uint64_t mamba_launch_count(void) { return mb::g_launches.load(std::memory_order_relaxed); }
