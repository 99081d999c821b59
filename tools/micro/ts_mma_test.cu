// Semantics check for tcgen05.mma with the A operand in tensor memory (sm_100a):
//   A[128 x 64] bf16 is written to TMEM with tcgen05.st.32x32b (row = lane, 32-bit column j = elements 2j, 2j+1),
//   B[64 (N) x 64 (K)] bf16 sits in shared memory, K-major, 128-byte swizzle; D = A B^T is read back and compared.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/micro/ts_mma_test tools/micro/ts_mma_test.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../headct_foundation_b200/csrc/hct_tcgen05.cuh"
using namespace hct_tc;

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(192) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int a_col0, int n) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&slot, 256);
  // B tile -> smem, K-major SW128: row r (n index) at r*128 B, 16-byte chunk c at (c ^ (r & 7))
  for (int i = threadIdx.x; i < n * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(smem + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t tmA = tm + a_col0, tmD = tm + 128;
  if (warp < 4) {
    const int row = warp * 32 + lane;
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = *reinterpret_cast<const uint32_t*>(A + row * 64 + 2 * j);
    tmem_st32(tmA + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, n, false, false);
      const uint64_t db = make_sdesc_sw128(smem_u32(smem), false, 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tc_mma_ts(tmD, tmA + ks * 8, db + ks * 2, idesc, ks > 0);
      tc_commit(&bar);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < n; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmD + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 16; ++j) D[row * n + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

int main() {
  for (int n : {64, 48, 16}) for (int a_col0 : {0, 32, 64}) {
    std::vector<__nv_bfloat16> A(128 * 64), B(64 * 64);
    std::vector<float> Af(128 * 64), Bf(64 * 64), ref(128 * n), out(128 * n);
    srand(1);
    for (int i = 0; i < 128 * 64; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; A[i] = __float2bfloat16(v); Af[i] = __bfloat162float(A[i]); }
    for (int i = 0; i < 64 * 64; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; B[i] = __float2bfloat16(v); Bf[i] = __bfloat162float(B[i]); }
    for (int m = 0; m < 128; ++m) for (int j = 0; j < n; ++j) { float s = 0; for (int kk = 0; kk < 64; ++kk) s += Af[m * 64 + kk] * Bf[j * 64 + kk]; ref[m * n + j] = s; }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, out.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    k<<<1, 192, 32 * 1024>>>(dA, dB, dD, a_col0, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0; for (size_t i = 0; i < out.size(); ++i) mx = fmax(mx, fabs(out[i] - ref[i]));
    printf("TS MMA N=%d, A at TMEM column %d: max |D - ref| = %.3e  (%s)\n", n, a_col0, mx, mx < 1e-3 ? "OK" : "MISMATCH");
  }
  return 0;
}
