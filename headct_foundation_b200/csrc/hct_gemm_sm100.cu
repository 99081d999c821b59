// tcgen05 / TMEM / TMA GEMM core for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, 4 stages)
//   warp 1      MMA issuer    (one elected lane issues tcgen05.mma 128x256x16, cta_group::1)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulator stages)
//   warps 4-11  epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// The two accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// Operands may be K-major (row-major [rows,K]) or MN-major (row-major [K,rows]); the latter is
// what lets dgrad (B = W as stored) and wgrad (A = dY, B = X as stored) run without transposes.
//
// Replaces nn.Linear / Conv3d-as-GEMM call sites listed in include/hct_b200.h.
#include <cuda.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/hct_b200.h"
#include "hct_common.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int MN_CHUNK_BYTES = 64 * BK * 2;  // one 64(mn) x 64(k) MN-major box = 8 KiB
constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;
constexpr int NUM_THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;   // 384
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int num_m_tiles, num_n_tiles, splits, kb_per_split, total_kb;
  uint32_t idesc;
  void* out; long long ldo;
  void* out2; long long ldo2;
  const float* bias;
  const float* res; long long ldres;
  const bf16* aux; long long ldaux;
  const float* pos; long long ldpos;
  const int* pos_idx; int pos_period;
  int rows_in, rows_out, row_off;
  float alpha;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 0x3ff) == 0x3ff) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {   // ~2 s at 2 GHz
        printf("hct_gemm: mbarrier wait timeout (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x,
               parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bit set.
//   K-major : rows of 128 B (64 bf16 of K); 8-row groups every 1024 B (SBO); LBO unused (=1).
//   MN-major: 64-element MN chunks of [64 k][128 B]; LBO = chunk stride (8 KiB), SBO = 8 k-rows = 1024 B.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, bool mn_major) {
  uint64_t d = static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(mn_major ? (MN_CHUNK_BYTES >> 4) : 1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// erf with |err| < 1.5e-7 (Abramowitz-Stegun 7.1.26) -- far below bf16 output rounding; keeps the
// epilogue off the critical path (erff() costs ~3x more issue slots).
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float r = 1.0f - poly * t * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float cdf = 0.5f * (1.0f + erf_as(x * 0.70710678118654752f));
  return fmaf(x * 0.39894228040143268f, __expf(-0.5f * x * x), cdf);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ------------------------------------------------------------------ epilogue for one 32-column chunk of one row
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, long long orow, int row, int col0,
                                               const uint32_t (&acc)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  const int ncols = min(32, p.N - col0);   // multiple of 8 (N % 8 == 0)

  if (EPI == HCT_EPI_BF16 || EPI == HCT_EPI_F32 || EPI == HCT_EPI_ATOMIC_F32) {
    if (p.alpha != 1.0f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
    }
  }
  if (EPI != HCT_EPI_DGELU_BF16 && EPI != HCT_EPI_ATOMIC_F32) {
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
    }
  }

  if (EPI == HCT_EPI_BF16 || EPI == HCT_EPI_GELU_BF16 || EPI == HCT_EPI_DGELU_BF16) {
    if (EPI == HCT_EPI_GELU_BF16) {
      if (p.out2 != nullptr) {
        bf16* o2 = reinterpret_cast<bf16*>(p.out2) + orow * p.ldo2 + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j < ncols) {
            uint4 u;
            u.x = pack_bf16x2(v[j], v[j + 1]); u.y = pack_bf16x2(v[j + 2], v[j + 3]);
            u.z = pack_bf16x2(v[j + 4], v[j + 5]); u.w = pack_bf16x2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(o2 + j) = u;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
    }
    if (EPI == HCT_EPI_DGELU_BF16) {
      const bf16* a = p.aux + static_cast<long long>(row) * p.ldaux + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j < ncols) {
          const uint4 u = *reinterpret_cast<const uint4*>(a + j);
          const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z),
                       a3 = unpack_bf16x2(u.w);
          v[j] *= gelu_grad_fast(a0.x); v[j + 1] *= gelu_grad_fast(a0.y);
          v[j + 2] *= gelu_grad_fast(a1.x); v[j + 3] *= gelu_grad_fast(a1.y);
          v[j + 4] *= gelu_grad_fast(a2.x); v[j + 5] *= gelu_grad_fast(a2.y);
          v[j + 6] *= gelu_grad_fast(a3.x); v[j + 7] *= gelu_grad_fast(a3.y);
        }
      }
    }
    bf16* o = reinterpret_cast<bf16*>(p.out) + orow * p.ldo + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < ncols) {
        uint4 u;
        u.x = pack_bf16x2(v[j], v[j + 1]); u.y = pack_bf16x2(v[j + 2], v[j + 3]);
        u.z = pack_bf16x2(v[j + 4], v[j + 5]); u.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = u;
      }
    }
  } else {
    float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + col0;
    if (EPI == HCT_EPI_RES_F32) {
      const float* r = p.res + static_cast<long long>(row) * p.ldres + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) {
          const float4 t = *reinterpret_cast<const float4*>(r + j);
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      }
    }
    if (EPI == HCT_EPI_POS_F32) {
      const int pr = p.pos_idx != nullptr ? p.pos_idx[row] : (row % p.pos_period);
      const float* r = p.pos + static_cast<long long>(pr) * p.ldpos + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(r + j));
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      }
    }
    if (EPI == HCT_EPI_ATOMIC_F32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (j < ncols) red_add_v4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (j < ncols) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
}

// ------------------------------------------------------------------ the kernel
template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hct_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int total_work = tiles * p.splits;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w % tiles, split = w / tiles;
      const int n0 = (tile % p.num_n_tiles) * BN, m0 = (tile / p.num_n_tiles) * BM;
      const int kb0 = split * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.total_kb);
      const int a_chunks = p.a_mn ? min(BM / 64, (p.M - m0 + 63) / 64) : 0;
      const int b_chunks = p.b_mn ? min(BN / 64, (p.N - n0 + 63) / 64) : 0;
      const uint32_t bytes = (p.a_mn ? a_chunks * MN_CHUNK_BYTES : A_STAGE_BYTES) +
                             (p.b_mn ? b_chunks * MN_CHUNK_BYTES : B_STAGE_BYTES);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], bytes);
        const uint32_t a_dst = smem_u32(sA + stage * A_STAGE_BYTES);
        const uint32_t b_dst = smem_u32(sB + stage * B_STAGE_BYTES);
        if (!p.a_mn) {
          tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BK, m0);
        } else {
          for (int i = 0; i < a_chunks; ++i)
            tma_load_2d(a_dst + i * MN_CHUNK_BYTES, &tmA, &full_bar[stage], m0 + i * 64, kb * BK);
        }
        if (!p.b_mn) {
          tma_load_2d(b_dst, &tmB, &full_bar[stage], kb * BK, n0);
        } else {
          for (int i = 0; i < b_chunks; ++i)
            tma_load_2d(b_dst + i * MN_CHUNK_BYTES, &tmB, &full_bar[stage], n0 + i * 64, kb * BK);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t a_kstep = p.a_mn ? (16 * 128) >> 4 : 32 >> 4;   // descriptor units of 16 B per UMMA_K=16
    const uint32_t b_kstep = p.b_mn ? (16 * 128) >> 4 : 32 >> 4;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w / tiles;
      const int kb0 = split * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.total_kb);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = make_sdesc(smem_u32(sA + stage * A_STAGE_BYTES), p.a_mn);
        const uint64_t bdesc = make_sdesc(smem_u32(sB + stage * B_STAGE_BYTES), p.b_mn);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          tc_mma(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        tc_commit(&empty_bar[stage]);   // smem slot reusable once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      tc_commit(&tfull_bar[acc]);       // accumulator complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= FIRST_EPI_WARP) {
    // ===================== epilogue =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may touch
    const int half = (warp - FIRST_EPI_WARP) >> 2;        // which 128-column half
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w % tiles;
      const int n0 = (tile % p.num_n_tiles) * BN, m0 = (tile / p.num_n_tiles) * BM;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      bool row_ok = row < p.M;
      long long orow = row;
      if (p.rows_in > 0) {
        const int g = row / p.rows_in, r = row % p.rows_in + p.row_off;
        row_ok = row_ok && r >= 0 && r < p.rows_out;
        orow = static_cast<long long>(g) * p.rows_out + r;
      }
#pragma unroll 1
      for (int c = 0; c < (BN / 2) / 32; ++c) {
        const int col0 = n0 + half * (BN / 2) + c * 32;
        if (col0 < p.N) {   // warp-uniform
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * (BN / 2) + c * 32, v);
          if (row_ok) epilogue_chunk<EPI>(p, orow, row, col0, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with leading dimension ld (elements).
int make_tmap(CUtensorMap* tm, const void* base, long long inner, long long outer, long long ld, int box_inner,
              int box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) { hct_set_error("cuTensorMapEncodeTiled entry point unavailable"); return HCT_ERR_CUDA; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    hct_set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%lld outer=%lld ld=%lld box=%dx%d", (int)r, base,
                  inner, outer, ld, box_inner, box_outer);
    return HCT_ERR_CUDA;
  }
  return HCT_OK;
}

template <int EPI>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hct_gemm_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BYTES);
    if (e != cudaSuccess) { hct_set_error("cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
    configured = true;
  }
  hct_gemm_tcgen05_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
  return hct_check_launch("hct_gemm_tcgen05_kernel");
}

}  // namespace

extern "C" int hct_gemm_bf16(const hct_gemm_desc* d, hct_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HCT_REQUIRE(d != nullptr, "hct_gemm_bf16: null descriptor");
  HCT_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "hct_gemm_bf16: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
  HCT_REQUIRE(d->N % 8 == 0, "hct_gemm_bf16: N=%d must be a multiple of 8", d->N);
  HCT_REQUIRE(d->lda % 8 == 0 && d->ldb % 8 == 0, "hct_gemm_bf16: lda/ldb must be multiples of 8 elements");
  HCT_REQUIRE((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->B) & 15) == 0,
              "hct_gemm_bf16: A/B must be 16-byte aligned");
  HCT_REQUIRE(d->out != nullptr && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "hct_gemm_bf16: out misaligned");
  HCT_REQUIRE(d->epilogue >= 0 && d->epilogue <= HCT_EPI_ATOMIC_F32, "hct_gemm_bf16: bad epilogue %d", d->epilogue);
  const bool f32_out = d->epilogue == HCT_EPI_RES_F32 || d->epilogue == HCT_EPI_POS_F32 ||
                       d->epilogue == HCT_EPI_F32 || d->epilogue == HCT_EPI_ATOMIC_F32;
  HCT_REQUIRE(d->ldo % (f32_out ? 4 : 8) == 0, "hct_gemm_bf16: ldo=%lld breaks 16-byte row alignment", (long long)d->ldo);
  if (d->epilogue == HCT_EPI_RES_F32)
    HCT_REQUIRE(d->res != nullptr && d->ldres % 4 == 0, "hct_gemm_bf16: RES epilogue needs res");
  if (d->epilogue == HCT_EPI_POS_F32)
    HCT_REQUIRE(d->pos != nullptr && d->ldpos % 4 == 0 && (d->pos_idx != nullptr || d->pos_period > 0),
                "hct_gemm_bf16: POS epilogue needs pos and pos_idx/pos_period");
  if (d->epilogue == HCT_EPI_DGELU_BF16)
    HCT_REQUIRE(d->aux != nullptr && d->ldaux % 8 == 0, "hct_gemm_bf16: DGELU epilogue needs aux");
  if (d->epilogue == HCT_EPI_GELU_BF16 && d->out2 != nullptr)
    HCT_REQUIRE(d->ldo2 % 8 == 0, "hct_gemm_bf16: ldo2 misaligned");
  if (d->a_mn_major) HCT_REQUIRE(d->M % 8 == 0, "hct_gemm_bf16: MN-major A needs M %% 8 == 0");
  else HCT_REQUIRE(d->K % 8 == 0, "hct_gemm_bf16: K-major A needs K %% 8 == 0");
  if (!d->b_mn_major) HCT_REQUIRE(d->K % 8 == 0, "hct_gemm_bf16: K-major B needs K %% 8 == 0");

  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_mn = d->a_mn_major ? 1 : 0; p.b_mn = d->b_mn_major ? 1 : 0;
  p.num_m_tiles = (d->M + BM - 1) / BM;
  p.num_n_tiles = (d->N + BN - 1) / BN;
  p.total_kb = (d->K + BK - 1) / BK;
  const int sms = hct_num_sms();
  int splits = 1;
  if (d->epilogue == HCT_EPI_ATOMIC_F32) {
    splits = d->splits;
    if (splits <= 0) {
      const int tiles = p.num_m_tiles * p.num_n_tiles;
      splits = (2 * sms + tiles - 1) / tiles;            // ~2 work items per SM
      const int max_splits = (p.total_kb + 7) / 8;       // keep >= 8 k-blocks per split
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
    if (splits > p.total_kb) splits = p.total_kb;
  }
  p.kb_per_split = (p.total_kb + splits - 1) / splits;
  p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;   // no empty split
  // instruction descriptor: D=f32, A=B=bf16, majorness, N>>3, M>>4
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.a_mn) << 15) |
            (static_cast<uint32_t>(p.b_mn) << 16) | (static_cast<uint32_t>(BN >> 3) << 17) |
            (static_cast<uint32_t>(BM >> 4) << 24);
  p.out = d->out; p.ldo = d->ldo; p.out2 = d->out2; p.ldo2 = d->ldo2;
  p.bias = d->bias; p.res = d->res; p.ldres = d->ldres;
  p.aux = static_cast<const bf16*>(d->aux); p.ldaux = d->ldaux;
  p.pos = d->pos; p.ldpos = d->ldpos; p.pos_idx = d->pos_idx; p.pos_period = d->pos_period;
  p.rows_in = d->rows_in; p.rows_out = d->rows_out; p.row_off = d->row_off;
  p.alpha = d->alpha == 0.0f ? 1.0f : d->alpha;

  CUtensorMap tmA, tmB;
  int rc;
  if (!p.a_mn) rc = make_tmap(&tmA, d->A, d->K, d->M, d->lda, BK, BM);
  else rc = make_tmap(&tmA, d->A, d->M, d->K, d->lda, 64, BK);
  if (rc != HCT_OK) return rc;
  if (!p.b_mn) rc = make_tmap(&tmB, d->B, d->K, d->N, d->ldb, BK, BN);
  else rc = make_tmap(&tmB, d->B, d->N, d->K, d->ldb, 64, BK);
  if (rc != HCT_OK) return rc;

  const int total_work = p.num_m_tiles * p.num_n_tiles * p.splits;
  const int grid = total_work < sms ? total_work : sms;
  void* prof = hct_prof_enabled() ? hct_prof_begin(stream) : nullptr;
  switch (d->epilogue) {
    case HCT_EPI_BF16: rc = launch<HCT_EPI_BF16>(tmA, tmB, p, grid, stream); break;
    case HCT_EPI_GELU_BF16: rc = launch<HCT_EPI_GELU_BF16>(tmA, tmB, p, grid, stream); break;
    case HCT_EPI_RES_F32: rc = launch<HCT_EPI_RES_F32>(tmA, tmB, p, grid, stream); break;
    case HCT_EPI_POS_F32: rc = launch<HCT_EPI_POS_F32>(tmA, tmB, p, grid, stream); break;
    case HCT_EPI_DGELU_BF16: rc = launch<HCT_EPI_DGELU_BF16>(tmA, tmB, p, grid, stream); break;
    case HCT_EPI_F32: rc = launch<HCT_EPI_F32>(tmA, tmB, p, grid, stream); break;
    default: rc = launch<HCT_EPI_ATOMIC_F32>(tmA, tmB, p, grid, stream); break;
  }
  if (prof != nullptr) hct_prof_end(prof, stream, 2.0 * d->M * d->N * static_cast<double>(d->K));
  return rc;
}
