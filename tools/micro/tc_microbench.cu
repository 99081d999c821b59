// Micro-benchmarks that size the attention kernels (sm_100a): tcgen05.ld throughput and small-tile tcgen05.mma cost,
// operands from shared memory (SS) or A from tensor memory (TS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/micro/tc_microbench tools/micro/tc_microbench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../headct_foundation_b200/csrc/hct_tcgen05.cuh"
using namespace hct_tc;

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: every warp loops `iters` x (tcgen05.ld 32x32b.x32 + wait); reports cycles (max over warps, via block 0)
// mode 1: one thread issues `iters` x [nmma SS MMAs (M=128,N=n,K=16) + commit + wait]
// mode 2: same with A from TMEM
__global__ void __launch_bounds__(512) bench(int mode, int nwarps, int iters, int nmma, int n, long long* out, int nd) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&slot, 512);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (mode == 0) {
    uint32_t acc = 0;
    __syncthreads();
    t0 = clock64();
    if (warp < nwarps) {
      const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
      for (int it = 0; it < iters; ++it) {
        uint32_t v[32];
        tmem_ld32_issue(tm + lane_off + ((it * 32) & 511), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      }
    }
    t1 = clock64();
    if (acc == 0x12345678u) out[100] = acc;
  } else {
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_bf16(128, n, false, false);
      const uint64_t da = make_sdesc_sw128(smem_u32(smem), false, 0);
      const uint64_t db = make_sdesc_sw128(smem_u32(smem + 16384), false, 0);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < nmma; ++k) {
          // nd independent accumulators, round robin (nd = 1: one dependent accumulate chain)
          const uint32_t d = tm + 256 + (k % nd) * 64;
          if (mode == 1) tc_mma(d, da + (k & 3) * 2, db + (k & 3) * 2, idesc, k >= nd);
          else tc_mma_ts(d, tm + (k & 3) * 8, db + (k & 3) * 2, idesc, k >= nd);
        }
        tc_commit(&bar);
        mbar_wait(&bar, it & 1);
      }
      t1 = clock64();
    }
  }
  __syncthreads();
  if (lane == 0 && (mode == 0 ? warp < nwarps : threadIdx.x == 0) && blockIdx.x == 0) out[warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// The same measurement with the issue path the kernels use now: a CONVERGED warp, MMAs issued under elect_one(), the
// batch fully unrolled (descriptors in uniform registers, consecutive UTCHMMA in the SASS).
template <int NMMA, int N, bool TS>
__global__ void __launch_bounds__(128) bench_elect(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&slot, 512);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint64_t da = make_sdesc_sw128(smem_u32(smem), false, 0);
    const uint64_t db = make_sdesc_sw128(smem_u32(smem + 16384), false, 0);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < NMMA; ++k) {
          const uint32_t d = tm + 256 + (k & 1) * 64;
          if (TS) tc_mma_ts(d, tm + (k & 3) * 8, db + (k & 3) * 2, idesc, k >= 2 ? 1u : 0u);
          else tc_mma(d, da + (k & 3) * 2, db + (k & 3) * 2, idesc, k >= 2 ? 1u : 0u);
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, it & 1);
    }
    t1 = clock64();
    if (leader && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int NMMA, int N, bool TS>
static void run_elect(long long* d, int smem) {
  long long h = 0;
  const int iters = 500;
  cudaFuncSetAttribute(bench_elect<NMMA, N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench_elect<NMMA, N, TS><<<148, 128, smem>>>(iters, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("elect bench failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("elect_one %s M=128 N=%3d K=16 x %2d MMAs + commit + wait: %.0f cycles/batch (%.1f per MMA)\n", TS ? "TS" : "SS", N, NMMA,
         double(h) / iters, double(h) / iters / NMMA);
}

int main() {
  long long* d; cudaMalloc(&d, 1024 * 8); cudaMemset(d, 0, 1024 * 8);
  long long h[128];
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int nw : {1, 4, 8, 16}) {
    const int iters = 2000;
    bench<<<148, 512, smem>>>(0, nw, iters, 0, 0, d, 1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("ld bench failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("tcgen05.ld 32x32b.x32: %2d warps x %d loads (4 KiB each): %lld cycles -> %.1f cycles/load/warp, %.1f B/cycle/SM\n", nw, iters, mx,
           double(mx) / iters, double(nw) * iters * 4096 / mx);
  }
  run_elect<4, 64, false>(d, smem); run_elect<8, 64, false>(d, smem); run_elect<16, 64, false>(d, smem); run_elect<32, 64, false>(d, smem);
  run_elect<4, 48, true>(d, smem); run_elect<8, 48, true>(d, smem); run_elect<16, 48, true>(d, smem); run_elect<32, 48, true>(d, smem);
  run_elect<16, 16, false>(d, smem); run_elect<32, 16, false>(d, smem); run_elect<16, 128, false>(d, smem); run_elect<32, 128, false>(d, smem);
  for (int mode : {1, 2}) for (int n : {64}) for (int nd : {2}) for (int nmma : {4, 16}) {
    const int iters = 500;
    bench<<<148, 128, smem>>>(mode, 0, iters, nmma, n, d, nd);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma bench failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s M=128 N=%3d K=16 x %2d MMAs over %d accumulators + commit + wait: %.0f cycles/batch (%.1f per MMA)\n", mode == 1 ? "SS" : "TS", n, nmma, nd,
           double(h[0]) / iters, double(h[0]) / iters / nmma);
  }
  return 0;
}
