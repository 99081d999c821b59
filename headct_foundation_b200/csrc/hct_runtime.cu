// Error reporting, launch accounting and device queries shared by all libhct_b200 entry points.
#include <atomic>
#include <vector>
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

// ---- optional per-launch timing by kernel class (bench.py's roofline / roofline_secondary numbers) ----
namespace {
struct ProfRec { cudaEvent_t a, b; double work; };
unsigned g_prof_mask = 0;
std::vector<ProfRec> g_prof[HCT_PROF_CLASSES];
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t get_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

bool hct_prof_enabled(int cls) { return (g_prof_mask >> cls) & 1u; }
void* hct_prof_begin(cudaStream_t st) {
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  return e;
}
void hct_prof_end(void* begin_event, cudaStream_t st, double work, int cls) {
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  g_prof[cls].push_back(ProfRec{static_cast<cudaEvent_t>(begin_event), e, work});
}

// `on`: bit mask of 1 << class (1 = the GEMM kernel only, as in round 1; 0 = off)
extern "C" int hct_profile_enable(int on) { g_prof_mask = static_cast<unsigned>(on); return HCT_OK; }
// Sums elapsed time / work (flops or bytes) over all launches of a class recorded since its last collect (synchronises on
// their events).
extern "C" int hct_profile_collect_class(int cls, double* total_ms, double* total_work, long long* launches) {
  HCT_REQUIRE(cls >= 0 && cls < HCT_PROF_CLASSES, "profile_collect_class: bad class %d", cls);
  double ms = 0, wk = 0;
  for (auto& r : g_prof[cls]) {
    float t = 0.f;
    cudaEventSynchronize(r.b);
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; wk += r.work; }
    g_event_pool.push_back(r.a); g_event_pool.push_back(r.b);
  }
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = wk;
  if (launches) *launches = static_cast<long long>(g_prof[cls].size());
  g_prof[cls].clear();
  return HCT_OK;
}
extern "C" int hct_profile_collect(double* total_ms, double* total_flops, long long* launches) {
  return hct_profile_collect_class(HCT_PROF_GEMM, total_ms, total_flops, launches);
}

extern "C" const char* hct_last_error(void) { return g_err; }
extern "C" int hct_abi_version(void) { return 1; }
extern "C" long long hct_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Programmatic dependent launch of the kernels that support it (they call pdl_wait() themselves); 0 = plain stream order.
// Off by default: back-to-back GEMM launches gain 3-6 % (5-7 us per launch, tools/gemm_bench.py with HCT_PDL=1), but the
// batch-256 training step runs at the board power limit and does not get faster (95.6 vs 96.3 ms, tools/graph_vs_eager.py):
// the cycles it recovers between kernels come back as a lower clock.
static int g_pdl = 0;
bool hct_pdl_enabled() { return g_pdl != 0; }
extern "C" int hct_set_pdl(int enable) { g_pdl = enable != 0; return HCT_OK; }
