// Error plumbing, launch accounting and ABI version of libuof_b200.so.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace uof {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = getenv("UOF_NO_PDL") == nullptr;
  return on;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UOF_ERR_CUDA;
  }
  return UOF_OK;
}

}  // namespace uof

extern "C" {

int uof_abi_version(void) { return 1; }

const char* uof_last_error(void) { return uof::g_error; }

long long uof_launch_count(void) { return uof::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
