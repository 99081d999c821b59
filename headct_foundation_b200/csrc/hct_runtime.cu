// Error reporting, launch accounting and device queries shared by all libhct_b200 entry points.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/hct_b200.h"
#include "hct_common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void hct_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int hct_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    hct_set_error("%s: %s", what, cudaGetErrorString(e));
    return HCT_ERR_CUDA;
  }
  return HCT_OK;
}

int hct_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

extern "C" const char* hct_last_error(void) { return g_err; }
extern "C" int hct_abi_version(void) { return 1; }
extern "C" long long hct_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
