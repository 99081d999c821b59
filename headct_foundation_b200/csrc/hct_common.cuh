// Shared device/host helpers for the headct_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define HCT_OK 0
#define HCT_ERR_INVALID 1
#define HCT_ERR_CUDA 2
#define HCT_ERR_UNSUPPORTED 3

void hct_set_error(const char* fmt, ...);
int hct_check_launch(const char* what);   // cudaGetLastError -> HCT_ERR_CUDA + message
int hct_num_sms();
// per-launch timing classes of the measurement aid (hct_profile_enable takes a bit mask of 1 << class)
enum { HCT_PROF_GEMM = 0, HCT_PROF_ATTN_FWD = 1, HCT_PROF_ATTN_BWD = 2, HCT_PROF_LN_FWD = 3, HCT_PROF_LN_BWD = 4,
       HCT_PROF_LOSS = 5, HCT_PROF_ADAMW = 6, HCT_PROF_WINDOW = 7, HCT_PROF_PATCHIFY = 8, HCT_PROF_CLASSES = 9 };
bool hct_prof_enabled(int cls = HCT_PROF_GEMM);
void* hct_prof_begin(cudaStream_t st);
void hct_prof_end(void* begin_event, cudaStream_t st, double work, int cls = HCT_PROF_GEMM);   // work: flops or bytes
// RAII bracket for an entry point: records the event pair only when the class is enabled
struct HctProfScope {
  void* begin; cudaStream_t st; double work; int cls;
  HctProfScope(cudaStream_t s, int c, double w) : begin(hct_prof_enabled(c) ? hct_prof_begin(s) : nullptr), st(s), work(w), cls(c) {}
  ~HctProfScope() { if (begin != nullptr) hct_prof_end(begin, st, work, cls); }
};

#define HCT_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) {                            \
      hct_set_error(__VA_ARGS__);             \
      return HCT_ERR_INVALID;                 \
    }                                         \
  } while (0)

typedef __nv_bfloat16 bf16;

// Programmatic dependent launch (griddepcontrol): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may start before its predecessor in the stream has finished and MUST call pdl_wait() before its first access to global
// memory (it returns once the predecessor grid has completed and its writes are visible; immediately for a normal launch).
// pdl_launch_dependents() lets the successor grid be scheduled early; call it only where every CTA of this grid is already
// resident (persistent grids) or at the end of a CTA's work, so that waiting successor CTAs never take slots this grid needs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool hct_pdl_enabled();    // hct_set_pdl(0) turns the launch attribute off (A/B)
// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute; the kernel must call pdl_wait()
template <typename... KArgs, typename... Args>
inline cudaError_t hct_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = hct_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// exact-erf GELU (torch nn.GELU(approximate='none')) and its derivative, fp32.
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// block-wide sum for blockDim.x <= 1024 (result broadcast to all threads)
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
