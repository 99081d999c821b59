// tcgen05 / TMEM / TMA GEMM core for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, 4 or 6 stages)
//   warp 1      MMA issuer    (one elected lane issues tcgen05.mma 128x256x16 / cta_group::2 256x256x16)
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
#include "hct_tcgen05.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int MN_CHUNK_BYTES = 64 * BK * 2;  // one 64(mn) x 64(k) MN-major box = 8 KiB
constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;
constexpr int NUM_THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;   // 384
constexpr int TMEM_COLS = 512;
constexpr int STG_BYTES = 4096;              // per epilogue warp: one 32 x 32 fp32 transpose buffer, or (TMA epilogue) two 2 KiB bf16 units

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
  float* colsum;
};

using namespace hct_tc;

__device__ long long* g_gemm_trace = nullptr;   // clock64 timeline of CTA 0 (tools/gemm_dbg.py); nullptr in production

__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, bool mn_major) {
  return make_sdesc_sw128(saddr, mn_major, MN_CHUNK_BYTES);
}

// GELU(x) = x * Phi(x) with the normal CDF evaluated as a logistic of an odd degree-7 polynomial, written with tanh
// so that it costs ONE special-function op:
//   Phi(x) ~= 1 / (1 + exp(-x p(x^2))) = 0.5 + 0.5 tanh(0.5 x p(x^2)),   p(t) = c0 + c1 t + c2 t^2 + c3 t^3,  |x| clamped to 6,
// p fitted (least squares on [-6, 6]) to erf-exact GELU: max |gelu error| 1.3e-5 with an exact tanh; tanh.approx
// (relative error <= 2^-11) keeps it below 8.5e-4 absolute -- under the bf16 rounding of the stored activations, so
// the outputs are those of exact-erf GELU (torch nn.GELU(approximate='none')) up to bf16 ties.
// The derivative uses the closed form Phi(x) + x phi(x) with the same Phi and an ex2 for the density.
// Cost: 8 FP32 ops + 1 MUFU (forward), 10 + 2 MUFU (derivative); the erff-based code needed ~20 / ~30 and made the
// GELU epilogues ALU-bound (a 128 x 256 tile has 32768 elements against a 6144-cycle K = 768 main loop).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kH0 = 0.5f * 1.59534958f, kH1 = 0.5f * 7.34984774e-02f, kH2 = 0.5f * -5.20865699e-04f,
                kH3 = 0.5f * -1.64242936e-05f;
__device__ __forceinline__ float tanh_half_arg(float x, float& x2_out) {   // tanh(0.5 x p(x^2))
  const float x2 = fminf(x * x, 36.0f);      // beyond |x| = 6 the argument keeps growing linearly: tanh saturates
  float q = fmaf(kH3, x2, kH2);
  q = fmaf(q, x2, kH1);
  q = fmaf(q, x2, kH0);
  x2_out = x2;
  return tanh_approx(x * q);
}
__device__ __forceinline__ float gelu_fast(float x) {
  float x2;
  const float t = tanh_half_arg(x, x2);
  const float h = 0.5f * x;
  return fmaf(h, t, h);                      // x (0.5 + 0.5 t)
}
__device__ __forceinline__ float gelu_and_grad_fast(float x, float& grad) {   // both from one tanh
  float x2;
  const float t = tanh_half_arg(x, x2);
  const float h = 0.5f * x;
  const float pdf = ex2_approx(x2 * (-0.5f * 1.4426950408889634f));
  grad = fmaf(x * 0.3989422804014327f, pdf, fmaf(0.5f, t, 0.5f));
  return fmaf(h, t, h);
}
// The same functions over NE elements, written stage by stage (NE independent dependency chains in flight) and with
// the packed two-lane fp32 instructions of sm_100 (fma.rn.f32x2 & co.): the epilogue of a K = 768 GEMM has ~20 fp32
// operations per output element against 0.19 tensor-pipe cycles, so the instruction count of this function sets the
// tile period of the GELU GEMMs.
template <int NE, bool WITH_GRAD>
__device__ __forceinline__ void gelu_multi(float (&xs)[NE], float (&grads)[NE]) {
  static_assert(NE % 2 == 0, "pairs");
  constexpr int NP = NE / 2;
  float2 x[NP], x2[NP], q[NP], t[NP];
#pragma unroll
  for (int e = 0; e < NP; ++e) x[e] = make_float2(xs[2 * e], xs[2 * e + 1]);
#pragma unroll
  for (int e = 0; e < NP; ++e) { x2[e] = __fmul2_rn(x[e], x[e]); x2[e].x = fminf(x2[e].x, 36.0f); x2[e].y = fminf(x2[e].y, 36.0f); }
#pragma unroll
  for (int e = 0; e < NP; ++e) q[e] = __ffma2_rn(make_float2(kH3, kH3), x2[e], make_float2(kH2, kH2));
#pragma unroll
  for (int e = 0; e < NP; ++e) q[e] = __ffma2_rn(q[e], x2[e], make_float2(kH1, kH1));
#pragma unroll
  for (int e = 0; e < NP; ++e) q[e] = __ffma2_rn(q[e], x2[e], make_float2(kH0, kH0));
#pragma unroll
  for (int e = 0; e < NP; ++e) { q[e] = __fmul2_rn(x[e], q[e]); t[e].x = tanh_approx(q[e].x); t[e].y = tanh_approx(q[e].y); }
  if (WITH_GRAD) {
    constexpr float kE = -0.5f * 1.4426950408889634f, kP = 0.3989422804014327f;
#pragma unroll
    for (int e = 0; e < NP; ++e) {
      x2[e] = __fmul2_rn(x2[e], make_float2(kE, kE));
      x2[e].x = ex2_approx(x2[e].x); x2[e].y = ex2_approx(x2[e].y);                    // exp(-x^2 / 2)
    }
#pragma unroll
    for (int e = 0; e < NP; ++e) {
      const float2 cdf = __ffma2_rn(make_float2(0.5f, 0.5f), t[e], make_float2(0.5f, 0.5f));
      const float2 g = __ffma2_rn(__fmul2_rn(x[e], make_float2(kP, kP)), x2[e], cdf);     // Phi(x) + x phi(x)
      grads[2 * e] = g.x; grads[2 * e + 1] = g.y;
    }
  }
#pragma unroll
  for (int e = 0; e < NP; ++e) {
    const float2 h = __fmul2_rn(x[e], make_float2(0.5f, 0.5f));
    const float2 y = __ffma2_rn(h, t[e], h);
    xs[2 * e] = y.x; xs[2 * e + 1] = y.y;
  }
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float x2;
  const float t = tanh_half_arg(x, x2);
  const float cdf = fmaf(0.5f, t, 0.5f);
  const float pdf = ex2_approx(x2 * (-0.5f * 1.4426950408889634f));       // exp(-x^2 / 2); 1.5e-8 at the clamp
  return fmaf(x * 0.3989422804014327f, pdf, cdf);                          // Phi(x) + x phi(x)
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {   // 16-byte aligned, bytes % 16 == 0
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ------------------------------------------------------------------ epilogue (coalesced domain)
// After the TMEM load each thread owns one accumulator ROW (32 fp32 columns).  Storing that way would touch 32
// different cache lines per instruction, so every epilogue warp transposes its 32x32 fp32 unit through a private
// 4 KiB XOR-swizzled smem buffer; afterwards lane l owns columns (l%8)*4..+3 of rows (l/8)+4i: one 16-byte
// vector per row, eight lanes = one full 128-byte line.  Bias lives in four registers; residual / aux / pos
// operands are fetched with the same coalesced mapping, all eight rows in flight before the math starts.
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---- phase 1: thread = row, write 8 x 16 B with chunk index XOR (row & 7)  (conflict-free)
__device__ __forceinline__ void epilogue_stage(uint32_t stg, int lane, const uint32_t (&acc)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  __syncwarp();
}

template <int EPI>
struct EpiTraits {
  static constexpr bool bf16_out = EPI == HCT_EPI_BF16 || EPI == HCT_EPI_GELU_BF16 || EPI == HCT_EPI_DGELU_BF16 ||
                                   EPI == HCT_EPI_GELU_DERIV_BF16 || EPI == HCT_EPI_MUL_BF16;
  static constexpr bool uses_aux = EPI == HCT_EPI_DGELU_BF16 || EPI == HCT_EPI_MUL_BF16;   // bf16 multiplicand, no bias
  static constexpr bool gelu_fwd = EPI == HCT_EPI_GELU_BF16 || EPI == HCT_EPI_GELU_DERIV_BF16;
  // bf16-output epilogues whose tiles leave (and whose bf16 multiplicand arrives) through TMA: see epilogue_tma_unit
  static constexpr bool tma = EPI == HCT_EPI_BF16 || EPI == HCT_EPI_GELU_BF16 || EPI == HCT_EPI_GELU_DERIV_BF16 ||
                              EPI == HCT_EPI_MUL_BF16 || EPI == HCT_EPI_RES_F32;
  static constexpr bool tma_in = EPI == HCT_EPI_MUL_BF16 || EPI == HCT_EPI_RES_F32;      // second operand tile loaded by TMA
  // unit = 32 rows x UC columns with 64-byte rows either way: 32 bf16 or 16 fp32 (2 KiB, the same swizzle and buffers)
  static constexpr int UC = EPI == HCT_EPI_RES_F32 ? 16 : 32;
};

// ------------------------------------------------------------------ TMA epilogue (bf16 outputs)
// The transposing epilogue above moves every accumulator element through the LSU four times (st.shared, ld.shared,
// st.global, plus the multiplicand's ld.global) and spends a third of its instructions on addresses and predicates: ncu
// shows the LSU data pipe at 51-54 % over the whole launch and the aux loads' latency exposed (stall_long_sb) for the K = 768
// GEMMs with two large bf16 streams (profiles/r02_ncu_gemm_epilogue.txt).  Here a thread keeps its accumulator ROW: the
// math runs in the TMEM register layout, results are written once to shared memory as a 32 x 32 bf16 box in the TMA
// 64-byte swizzle (conflict-free 16-byte stores), and one elected lane hands the box to cp.async.bulk.tensor -- no
// ld.shared, no st.global, no address arithmetic, M / N tails clipped by the tensor map.  The MUL epilogue's multiplicand
// comes in the same way (TMA load into the same swizzled layout, one unit ahead; the first unit of a tile while the main
// loop still runs), so its DRAM latency never meets a scoreboard.
__device__ __forceinline__ uint32_t sw64(int row, int chunk) {      // byte offset of 16-byte chunk (8 bf16) of a 64-byte row
  return static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_shared_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}

// One 32 x 32 unit: `acc` = this thread's row (32 fp32 columns).  buf: the unit's 4 KiB buffer (out1 at +0, out2 / aux at
// +2048).  Returns after the results sit in shared memory (the caller fences and issues the TMA store).
// the unit's 32 bias values (the same for every lane: broadcast loads), requested ahead of the accumulator wait
template <int EPI>
__device__ __forceinline__ void load_bias_row(const GemmParams& p, int col0, float4 (&b)[8]) {
  constexpr int NV = EpiTraits<EPI>::UC / 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!EpiTraits<EPI>::uses_aux && p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (col0 + 4 * j < p.N) b[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_tma_unit(const GemmParams& p, uint32_t buf, int lane, int col0, const uint32_t (&acc)[32],
                                                  const float4 (&bias)[8]) {
  using T = EpiTraits<EPI>;
#pragma unroll
  for (int j = 0; j < 4; ++j) {                         // 8 columns per pass
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(acc[8 * j + e]);
    if (EPI == HCT_EPI_BF16) {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] *= p.alpha;
    }
    if (!T::uses_aux) {
      const float4 b0 = bias[2 * j], b1 = bias[2 * j + 1];
      x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
    }
    const uint32_t off = sw64(lane, j);
    if (EPI == HCT_EPI_GELU_BF16 && p.out2 != nullptr) {      // pre-activation side output
      st_shared_v4(buf + 2048 + off, pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
    }
    if (T::gelu_fwd) {
      float d[8];
      gelu_multi<8, EPI == HCT_EPI_GELU_DERIV_BF16>(x, d);
      if (EPI == HCT_EPI_GELU_DERIV_BF16)
        st_shared_v4(buf + 2048 + off, pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
    }
    if (EPI == HCT_EPI_MUL_BF16) {
      const uint4 a = ld_shared_u4(buf + 2048 + off);
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      x[0] *= a0.x; x[1] *= a0.y; x[2] *= a1.x; x[3] *= a1.y; x[4] *= a2.x; x[5] *= a2.y; x[6] *= a3.x; x[7] *= a3.y;
    }
    st_shared_v4(buf + off, pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
  }
}
// fp32 variant (RES_F32: out = acc + bias + res, residual stream): 16 columns per unit, four 16-byte chunks of 4 floats
__device__ __forceinline__ void epilogue_tma_unit_res(uint32_t buf, int lane, const uint32_t (&acc)[16], const float4 (&bias)[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t off = sw64(lane, j);
    const float4 r = ld_shared_f4(buf + 2048 + off);
    const float4 b = bias[j];
    st_shared_v4(buf + off, __float_as_uint(__uint_as_float(acc[4 * j]) + b.x + r.x), __float_as_uint(__uint_as_float(acc[4 * j + 1]) + b.y + r.y),
                 __float_as_uint(__uint_as_float(acc[4 * j + 2]) + b.z + r.z), __float_as_uint(__uint_as_float(acc[4 * j + 3]) + b.w + r.w));
  }
}
// column sums of a staged bf16 unit (as the consumer will read it): lane = column, 32 two-byte reads down the rows
__device__ __forceinline__ void colsum_tma_unit(const GemmParams& p, uint32_t buf, int lane, int row_base, int col0) {
  const int rows = min(32, p.M - row_base);
  float s = 0.f;
  const int chunk = lane >> 3, within = (lane & 7) * 2;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const uint32_t u = ld_shared_u16(buf + sw64(r, chunk) + within);
    if (r < rows) s += __uint_as_float(u << 16);
  }
  if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, s);
}

// bf16 multiplicand rows of one interior 32 x 32 unit (lane = (row sub-index, 4-column group)), eight 8-byte loads
template <int EPI>
__device__ __forceinline__ bool unit_interior(const GemmParams& p, int row_base, int col0) {
  return EPI != HCT_EPI_POS_F32 && p.rows_in <= 0 && row_base + 32 <= p.M && col0 + 32 <= p.N;
}
__device__ __forceinline__ void load_aux_unit(const GemmParams& p, int lane, int row_base, int col0, uint2 (&aux)[8]) {
  const bf16* ap = p.aux + static_cast<long long>(row_base + (lane >> 3)) * p.ldaux + col0 + (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) aux[i] = *reinterpret_cast<const uint2*>(ap + static_cast<long long>(i) * 4 * p.ldaux);
}

// `pre`: the unit's aux operand already in registers (loaded one unit ahead by the caller), interior units only
// the four bias values of this lane's column group in a 32-column unit (zero where the epilogue has no bias)
template <int EPI>
__device__ __forceinline__ float4 load_bias_unit(const GemmParams& p, int lane, int col0) {
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!EpiTraits<EPI>::uses_aux && EPI != HCT_EPI_ATOMIC_F32) {
    const int col = col0 + (lane & 7) * 4;
    if (p.bias != nullptr && col < p.N) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  }
  return bias4;
}

template <int EPI>
__device__ __forceinline__ void epilogue_drain(const GemmParams& p, uint32_t stg, int lane, int row_base, int col0,
                                               const float4 bias4, const uint2 (&pre)[8], bool have_pre) {
  using T = EpiTraits<EPI>;
  // ---- phase 2: lane = (row sub-index, column group); rows handled in two groups of four to bound registers
  const int jj = lane & 7, rsub = lane >> 3;
  const int col = col0 + jj * 4;
  const bool col_ok = col < p.N;                       // N % 8 == 0 -> a group of 4 is all-in or all-out
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  // Interior units (all 32 rows and 32 columns inside the matrix, no row remapping) take a path without per-row
  // predicates and with incremental row pointers: the general path below spends ~40 % of its instructions on
  // 64-bit address arithmetic and reconvergence, which made the GELU epilogues (not the MMAs) set the tile period.
  const bool interior = unit_interior<EPI>(p, row_base, col0);
  if (interior) {
    const long long r0 = row_base + rsub;
    const bool has_colsum = p.colsum != nullptr;
    float4 extra[8];
    uint2 aux[8];
    if (EPI == HCT_EPI_RES_F32) {
      const float* rp = p.res + r0 * p.ldres + col;
#pragma unroll
      for (int i = 0; i < 8; ++i) extra[i] = *reinterpret_cast<const float4*>(rp + static_cast<long long>(i) * 4 * p.ldres);
    }
    if (T::uses_aux) {
      if (have_pre) {
#pragma unroll
        for (int i = 0; i < 8; ++i) aux[i] = pre[i];
      } else {
        load_aux_unit(p, lane, row_base, col0, aux);
      }
    }
    if (T::bf16_out) {
      bf16* op = reinterpret_cast<bf16*>(p.out) + r0 * p.ldo + col;
      bf16* op2 = (T::gelu_fwd && p.out2 != nullptr) ? reinterpret_cast<bf16*>(p.out2) + r0 * p.ldo2 + col : nullptr;
      const long long ostep = 4 * p.ldo, ostep2 = 4 * p.ldo2;
      // all eight staged rows are pulled into registers first: the shared-memory loads are volatile asm, so a load
      // placed inside the row loop could not move above the previous row's global store and the rows ran strictly
      // one after the other (one dependent chain, ~5 cycles per instruction)
      float4 vals[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + rsub;
        vals[i] = ld_shared_f4(stg + r * 128 + ((jj ^ (r & 7)) << 4));
      }
      if (T::gelu_fwd) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {                  // two rows = eight independent elements per pass
          float x[8] = {vals[i].x + bias4.x, vals[i].y + bias4.y, vals[i].z + bias4.z, vals[i].w + bias4.w,
                        vals[i + 1].x + bias4.x, vals[i + 1].y + bias4.y, vals[i + 1].z + bias4.z, vals[i + 1].w + bias4.w};
          float d[8];
          if (EPI == HCT_EPI_GELU_BF16 && op2 != nullptr) {
            uint2 u0, u1;
            u0.x = pack_bf16x2(x[0], x[1]); u0.y = pack_bf16x2(x[2], x[3]);
            u1.x = pack_bf16x2(x[4], x[5]); u1.y = pack_bf16x2(x[6], x[7]);
            *reinterpret_cast<uint2*>(op2 + i * ostep2) = u0;
            *reinterpret_cast<uint2*>(op2 + (i + 1) * ostep2) = u1;
          }
          gelu_multi<8, EPI == HCT_EPI_GELU_DERIV_BF16>(x, d);
          if (EPI == HCT_EPI_GELU_DERIV_BF16) {
            uint2 u0, u1;
            u0.x = pack_bf16x2(d[0], d[1]); u0.y = pack_bf16x2(d[2], d[3]);
            u1.x = pack_bf16x2(d[4], d[5]); u1.y = pack_bf16x2(d[6], d[7]);
            *reinterpret_cast<uint2*>(op2 + i * ostep2) = u0;
            *reinterpret_cast<uint2*>(op2 + (i + 1) * ostep2) = u1;
          }
          uint2 u0, u1;
          u0.x = pack_bf16x2(x[0], x[1]); u0.y = pack_bf16x2(x[2], x[3]);
          u1.x = pack_bf16x2(x[4], x[5]); u1.y = pack_bf16x2(x[6], x[7]);
          *reinterpret_cast<uint2*>(op + i * ostep) = u0;
          *reinterpret_cast<uint2*>(op + (i + 1) * ostep) = u1;
          if (has_colsum) {
            const float2 q0 = unpack_bf16x2(u0.x), q1 = unpack_bf16x2(u0.y), q2 = unpack_bf16x2(u1.x), q3 = unpack_bf16x2(u1.y);
            csum.x += q0.x + q2.x; csum.y += q0.y + q2.y; csum.z += q1.x + q3.x; csum.w += q1.y + q3.y;
          }
        }
      } else
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v = vals[i];
        if (EPI == HCT_EPI_BF16) { v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha; }
        if (!T::uses_aux) { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
        if (EPI == HCT_EPI_GELU_BF16) {
          if (op2 != nullptr) {
            uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(op2 + i * ostep2) = u;
          }
          v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w);
        }
        if (EPI == HCT_EPI_GELU_DERIV_BF16) {
          float4 d;
          v.x = gelu_and_grad_fast(v.x, d.x); v.y = gelu_and_grad_fast(v.y, d.y);
          v.z = gelu_and_grad_fast(v.z, d.z); v.w = gelu_and_grad_fast(v.w, d.w);
          uint2 u; u.x = pack_bf16x2(d.x, d.y); u.y = pack_bf16x2(d.z, d.w);
          *reinterpret_cast<uint2*>(op2 + i * ostep2) = u;
        }
        if (EPI == HCT_EPI_DGELU_BF16) {
          const float2 a0 = unpack_bf16x2(aux[i].x), a1 = unpack_bf16x2(aux[i].y);
          v.x *= gelu_grad_fast(a0.x); v.y *= gelu_grad_fast(a0.y); v.z *= gelu_grad_fast(a1.x); v.w *= gelu_grad_fast(a1.y);
        }
        if (EPI == HCT_EPI_MUL_BF16) {
          const float2 a0 = unpack_bf16x2(aux[i].x), a1 = unpack_bf16x2(aux[i].y);
          v.x *= a0.x; v.y *= a0.y; v.z *= a1.x; v.w *= a1.y;
        }
        uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(op + i * ostep) = u;
        if (has_colsum) {            // column sums of the values as the consumer will read them (bf16-rounded)
          const float2 q0 = unpack_bf16x2(u.x), q1 = unpack_bf16x2(u.y);
          csum.x += q0.x; csum.y += q0.y; csum.z += q1.x; csum.w += q1.y;
        }
      }
    } else {
      float* op = reinterpret_cast<float*>(p.out) + r0 * p.ldo + col;
      const long long ostep = 4 * p.ldo;
      float4 vals[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + rsub;
        vals[i] = ld_shared_f4(stg + r * 128 + ((jj ^ (r & 7)) << 4));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v = vals[i];
        if (EPI == HCT_EPI_F32 || EPI == HCT_EPI_ATOMIC_F32) { v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha; }
        if (EPI != HCT_EPI_ATOMIC_F32) { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
        if (EPI == HCT_EPI_RES_F32) { v.x += extra[i].x; v.y += extra[i].y; v.z += extra[i].z; v.w += extra[i].w; }
        if (EPI == HCT_EPI_ATOMIC_F32) red_add_v4(op + i * ostep, v.x, v.y, v.z, v.w);
        else *reinterpret_cast<float4*>(op + i * ostep) = v;
      }
    }
  } else
#pragma unroll
  for (int grp = 0; grp < 1; ++grp) {
    long long orow[8];
    bool ok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + (grp * 8 + i) * 4 + rsub;
      ok[i] = col_ok && row < p.M;
      orow[i] = row;
      if (p.rows_in > 0) {
        const int g = row / p.rows_in, r = row % p.rows_in + p.row_off;
        ok[i] = ok[i] && r >= 0 && r < p.rows_out;
        orow[i] = static_cast<long long>(g) * p.rows_out + r;
      }
    }
    // operands that do not depend on the accumulator: issue the group's loads first
    float4 extra[8];
    uint2 aux[8];
    if (EPI == HCT_EPI_RES_F32) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (ok[i]) extra[i] = *reinterpret_cast<const float4*>(p.res + static_cast<long long>(row_base + (grp * 8 + i) * 4 + rsub) * p.ldres + col);
    }
    if (EPI == HCT_EPI_POS_F32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (ok[i]) {
          const int row = row_base + (grp * 8 + i) * 4 + rsub;
          const int pr = p.pos_idx != nullptr ? p.pos_idx[row] : (row % p.pos_period);
          extra[i] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<long long>(pr) * p.ldpos + col));
        }
      }
    }
    if (T::uses_aux) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (ok[i]) aux[i] = *reinterpret_cast<const uint2*>(p.aux + static_cast<long long>(row_base + (grp * 8 + i) * 4 + rsub) * p.ldaux + col);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (grp * 8 + i) * 4 + rsub;
      float4 v = ld_shared_f4(stg + r * 128 + ((jj ^ (r & 7)) << 4));
      if (!ok[i]) continue;
      if (EPI == HCT_EPI_BF16 || EPI == HCT_EPI_F32 || EPI == HCT_EPI_ATOMIC_F32) {
        v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha;
      }
      v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
      if (T::bf16_out) {
        if (EPI == HCT_EPI_GELU_BF16) {
          if (p.out2 != nullptr) {
            uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out2) + orow[i] * p.ldo2 + col) = u;
          }
          v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w);
        }
        if (EPI == HCT_EPI_GELU_DERIV_BF16) {
          float4 d;
          v.x = gelu_and_grad_fast(v.x, d.x); v.y = gelu_and_grad_fast(v.y, d.y);
          v.z = gelu_and_grad_fast(v.z, d.z); v.w = gelu_and_grad_fast(v.w, d.w);
          uint2 u; u.x = pack_bf16x2(d.x, d.y); u.y = pack_bf16x2(d.z, d.w);
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out2) + orow[i] * p.ldo2 + col) = u;
        }
        if (EPI == HCT_EPI_DGELU_BF16) {
          const float2 a0 = unpack_bf16x2(aux[i].x), a1 = unpack_bf16x2(aux[i].y);
          v.x *= gelu_grad_fast(a0.x); v.y *= gelu_grad_fast(a0.y); v.z *= gelu_grad_fast(a1.x); v.w *= gelu_grad_fast(a1.y);
        }
        if (EPI == HCT_EPI_MUL_BF16) {
          const float2 a0 = unpack_bf16x2(aux[i].x), a1 = unpack_bf16x2(aux[i].y);
          v.x *= a0.x; v.y *= a0.y; v.z *= a1.x; v.w *= a1.y;
        }
        uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + orow[i] * p.ldo + col) = u;
        if (p.colsum != nullptr) {   // column sums of the values as the consumer will read them (bf16-rounded)
          const float2 r0 = unpack_bf16x2(u.x), r1 = unpack_bf16x2(u.y);
          csum.x += r0.x; csum.y += r0.y; csum.z += r1.x; csum.w += r1.y;
        }
      } else {
        float* o = reinterpret_cast<float*>(p.out) + orow[i] * p.ldo + col;
        if (EPI == HCT_EPI_RES_F32 || EPI == HCT_EPI_POS_F32) {
          v.x += extra[i].x; v.y += extra[i].y; v.z += extra[i].z; v.w += extra[i].w;
        }
        if (EPI == HCT_EPI_ATOMIC_F32) red_add_v4(o, v.x, v.y, v.z, v.w);
        else *reinterpret_cast<float4*>(o) = v;
      }
    }
  }
  if (T::bf16_out && p.colsum != nullptr) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o); csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
      csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o); csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
    }
    if (rsub == 0 && col_ok) red_add_v4(p.colsum + col, csum.x, csum.y, csum.z, csum.w);
  }
  __syncwarp();   // staging buffer is reused by the next unit
}

// ------------------------------------------------------------------ the kernel
// CTAS = 1: one CTA per tile of 128 x 256, 4-stage ring (48 KiB / stage).
// CTAS = 2: a CTA pair (cluster of 2 on one TPC) per tile of 256 x 256 with tcgen05.mma.cta_group::2: each CTA
//           stages its own 128 rows of A and HALF of B (128 of the 256 N rows), so a stage is 32 KiB per SM
//           (6-stage ring) and L2->SM operand traffic per flop drops by a third.  Only the leader (rank 0)
//           issues MMAs; TMA loads of both CTAs complete on the leader's `full` barrier; tcgen05.commit
//           multicasts to both CTAs' `empty` / `tmem full` barriers; both epilogues report to the leader's
//           `tmem empty` barrier.
template <int CTAS, int EPI>
struct Cfg {
  // TMA epilogues with a second bf16 stream (gelu' / pre-activation out, multiplicand in) double-buffer 2 x 2 KiB per warp
  // and give up one pipeline stage for it (their GEMMs have K = 768: the ring depth matters least there)
  static constexpr bool TWO_STREAMS = EPI == HCT_EPI_GELU_BF16 || EPI == HCT_EPI_GELU_DERIV_BF16 || EPI == HCT_EPI_MUL_BF16 ||
                                      EPI == HCT_EPI_RES_F32;
  static constexpr int WARP_STG = TWO_STREAMS ? 2 * STG_BYTES : STG_BYTES;
  static constexpr int STAGES = (CTAS == 2 ? 6 : 4) - (TWO_STREAMS ? 1 : 0);
  static constexpr int B_ROWS = BN / CTAS;                     // B rows staged by each CTA
  static constexpr int B_STAGE = B_ROWS * BK * 2;
  static constexpr int STAGE = A_STAGE_BYTES + B_STAGE;
  static constexpr int SMEM = STAGES * STAGE + EPI_WARPS * WARP_STG + 1024 + 512;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> rank 0

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tc_mma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // remote arrive on rank 0's barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

template <int EPI, int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
hct_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                        const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
  using C = Cfg<CTAS, EPI>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sStage = smem + STAGES * C::STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + EPI_WARPS * C::WARP_STG);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* aux_bar = tempty_bar + 2;                   // [EPI_WARPS][2]: TMA loads of the MUL epilogue's multiplicand units
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 2 * EPI_WARPS);

  // warp id and CTA rank are made warp-uniform by construction (shfl): the control warps below run their loops as whole
  // converged warps and issue TMA / MMA under elect_one(), so descriptors stay in uniform registers and the four
  // UTCHMMA of a stage issue back to back (a divergent `lane == 0` branch costs an ELECT loop + R2UR moves per MMA).
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int rank = CTAS == 2 ? __shfl_sync(0xffffffffu, static_cast<int>(cluster_ctarank()), 0) : 0;
  const int unit = blockIdx.x / CTAS, num_units = gridDim.x / CTAS;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int total_work = tiles * p.splits;
  constexpr int TILE_M = BM * CTAS;

  if (threadIdx.x == 0) {
    pdl_launch_dependents();    // persistent grid, every CTA resident from the start: the successor may queue behind us now
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (EpiTraits<EPI>::tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmOut)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EPI_WARPS * CTAS); }
    for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&aux_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  // Programmatic dependent launch: this grid may have been started while its predecessor in the stream was still running
  // (its CTAs take an SM as soon as the predecessor's CTA there has exited, and get the prologue above out of the way);
  // nothing before this line touches global memory the predecessor may still write or read.
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  long long* trace = (g_gemm_trace != nullptr && blockIdx.x == 0) ? g_gemm_trace : nullptr;
  // trace layout: [0,512): MMA warp, 8 events per tile; [512, 512+16*64): epilogue warp 4, 16 events per tile

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its share of B) =====================
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int w = unit; w < total_work; w += num_units) {
      const int tile = w % tiles, split = w / tiles;
      const int n0 = (tile % p.num_n_tiles) * BN, m0 = (tile / p.num_n_tiles) * TILE_M;
      const int kb0 = split * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.total_kb);
      const int my_m0 = m0 + rank * BM, my_n0 = n0 + rank * C::B_ROWS;
      // MN-major operands are staged in 64-column chunks; chunks entirely outside the matrix are skipped
      uint32_t bytes = 0;       // bytes that will land on the (leader's) full barrier per stage, all CTAs
      int a_chunks = 0, b_chunks = 0;
#pragma unroll
      for (int r = 0; r < CTAS; ++r) {
        const int ac = p.a_mn ? max(0, min(BM / 64, (p.M - (m0 + r * BM) + 63) / 64)) : 0;
        const int bc = p.b_mn ? max(0, min(C::B_ROWS / 64, (p.N - (n0 + r * C::B_ROWS) + 63) / 64)) : 0;
        bytes += (p.a_mn ? ac * MN_CHUNK_BYTES : A_STAGE_BYTES) + (p.b_mn ? bc * MN_CHUNK_BYTES : C::B_STAGE);
        if (r == rank) { a_chunks = ac; b_chunks = bc; }
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const uint32_t a_dst = smem_u32(sA + stage * A_STAGE_BYTES);
        const uint32_t b_dst = smem_u32(sB + stage * C::B_STAGE);
        if (!leader) {
        } else if (CTAS == 2) {
          if (rank == 0) mbar_expect_tx(&full_bar[stage], bytes);
          if (!p.a_mn) tma_load_2d_2sm(a_dst, &tmA, &full_bar[stage], kb * BK, my_m0);
          else for (int i = 0; i < a_chunks; ++i)
            tma_load_2d_2sm(a_dst + i * MN_CHUNK_BYTES, &tmA, &full_bar[stage], my_m0 + i * 64, kb * BK);
          if (!p.b_mn) tma_load_2d_2sm(b_dst, &tmB, &full_bar[stage], kb * BK, my_n0);
          else for (int i = 0; i < b_chunks; ++i)
            tma_load_2d_2sm(b_dst + i * MN_CHUNK_BYTES, &tmB, &full_bar[stage], my_n0 + i * 64, kb * BK);
        } else {
          mbar_expect_tx(&full_bar[stage], bytes);
          if (!p.a_mn) tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BK, my_m0);
          else for (int i = 0; i < a_chunks; ++i)
            tma_load_2d(a_dst + i * MN_CHUNK_BYTES, &tmA, &full_bar[stage], my_m0 + i * 64, kb * BK);
          if (!p.b_mn) tma_load_2d(b_dst, &tmB, &full_bar[stage], kb * BK, my_n0);
          else for (int i = 0; i < b_chunks; ++i)
            tma_load_2d(b_dst + i * MN_CHUNK_BYTES, &tmB, &full_bar[stage], my_n0 + i * 64, kb * BK);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA only) =====================
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t a_kstep = p.a_mn ? (16 * 128) >> 4 : 32 >> 4;   // descriptor units of 16 B per UMMA_K=16
    const uint32_t b_kstep = p.b_mn ? (16 * 128) >> 4 : 32 >> 4;
    for (int w = unit; w < total_work; w += num_units) {
      const int split = w / tiles;
      const int kb0 = split * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.total_kb);
      const int tix = (w - unit) / num_units;
      if (trace && lane == 0 && tix < 64) trace[tix * 8 + 0] = clock64();
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      if (trace && lane == 0 && tix < 64) trace[tix * 8 + 1] = clock64();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = make_sdesc(smem_u32(sA + stage * A_STAGE_BYTES), p.a_mn);
        const uint64_t bdesc = make_sdesc(smem_u32(sB + stage * C::B_STAGE), p.b_mn);
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (CTAS == 2) tc_mma_2sm(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else tc_mma(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CTAS == 2) tc_commit_2sm(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (leader) { if (CTAS == 2) tc_commit_2sm(&tfull_bar[acc]); else tc_commit(&tfull_bar[acc]); }
      __syncwarp();
      if (trace && lane == 0 && tix < 64) trace[tix * 8 + 2] = clock64();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= FIRST_EPI_WARP && EpiTraits<EPI>::tma) {
    // ===================== epilogue, TMA path (bf16 outputs; see epilogue_tma_unit) =====================
    using T = EpiTraits<EPI>;
    const int q = warp & 3;                               // TMEM lane quarter this warp may touch
    const int half = (warp - FIRST_EPI_WARP) >> 2;        // which 128-column half
    constexpr int UB = C::WARP_STG / 2;                   // bytes per unit buffer: out1 at +0, second stream (if any) at +2048
    const uint32_t stg = smem_u32(sStage + (warp - FIRST_EPI_WARP) * C::WARP_STG);
    uint64_t* my_aux_bar = aux_bar + 2 * (warp - FIRST_EPI_WARP);
    const bool two_out = EPI == HCT_EPI_GELU_DERIV_BF16 || (EPI == HCT_EPI_GELU_BF16 && p.out2 != nullptr);
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t u = 0;                                       // units handled by this warp: buffer u & 1, aux parity (u >> 1) & 1
    for (int w = unit; w < total_work; w += num_units) {
      const int tile = w % tiles;
      const int n0 = (tile % p.num_n_tiles) * BN, m0 = (tile / p.num_n_tiles) * TILE_M + rank * BM;
      const int row_base = m0 + q * 32;
      const int colw = n0 + half * (BN / 2);
      constexpr int UC = T::UC;
      int nunits = 0;
      if (row_base < p.M && colw < p.N) nunits = min((BN / 2) / UC, (p.N - colw + UC - 1) / UC);
      // the multiplicand of the first unit is requested before the tile's MMAs have finished (off the chain)
      if (T::tma_in && nunits > 0 && lane == 0) {
        const uint32_t b = stg + (u & 1) * UB + 2048;
        mbar_expect_tx(&my_aux_bar[u & 1], 2048);
        tma_load_2d(b, &tmAux, &my_aux_bar[u & 1], colw, row_base);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * (BN / 2);
      uint32_t v[UC];
      if (nunits > 0) { if constexpr (UC == 32) tmem_ld32_issue(t0, v); else tmem_ld16_issue(t0, v); }
#pragma unroll 1
      for (int c = 0; c < nunits; ++c, ++u) {
        const uint32_t buf = stg + (u & 1) * UB;
        const int col0 = colw + c * UC;
        // the store that read this buffer two units ago must have finished reading before it is overwritten
        if (lane == 0) {
          bulk_wait_read<1>();
          if (T::tma_in && c + 1 < nunits) {              // multiplicand of the next unit: its buffer half was last read one unit ago
            const uint32_t bn = stg + ((u + 1) & 1) * UB + 2048;
            mbar_expect_tx(&my_aux_bar[(u + 1) & 1], 2048);
            tma_load_2d(bn, &tmAux, &my_aux_bar[(u + 1) & 1], col0 + UC, row_base);
          }
        }
        __syncwarp();
        float4 bias[8];
        load_bias_row<EPI>(p, col0, bias);                // in flight across the accumulator wait
        tmem_ld_wait();
        uint32_t a[UC];
#pragma unroll
        for (int i = 0; i < UC; ++i) a[i] = v[i];
        if (c + 1 < nunits) {                             // next unit's accumulators in flight under this unit's math
          if constexpr (UC == 32) tmem_ld32_issue(t0 + (c + 1) * UC, v); else tmem_ld16_issue(t0 + (c + 1) * UC, v);
        } else {
          tc_fence_before(); __syncwarp();
          if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
        }
        if (T::tma_in) mbar_wait(&my_aux_bar[u & 1], (u >> 1) & 1);
        if constexpr (EPI == HCT_EPI_RES_F32) epilogue_tma_unit_res(buf, lane, a, bias);
        else epilogue_tma_unit<EPI>(p, buf, lane, col0, a, bias);
        fence_proxy_async_smem();                         // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, buf, col0, row_base);
          if (two_out) tma_store_2d(&tmOut2, buf + 2048, col0, row_base);
          bulk_commit();
        }
        if (UC == 32 && p.colsum != nullptr) colsum_tma_unit(p, buf, lane, row_base, col0);
        if (T::tma_in) __syncwarp();                      // every lane has read the multiplicand before its half is reloaded
      }
      if (nunits == 0) {
        tc_fence_before(); __syncwarp();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) bulk_wait_all();                       // shared memory (and the stores) outlive the CTA otherwise
    __syncwarp();
  } else if (warp >= FIRST_EPI_WARP) {
    // ===================== epilogue =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may touch
    const int half = (warp - FIRST_EPI_WARP) >> 2;        // which 128-column half
    const uint32_t stg = smem_u32(sStage + (warp - FIRST_EPI_WARP) * C::WARP_STG);
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = unit; w < total_work; w += num_units) {
      const int tile = w % tiles;
      const int n0 = (tile % p.num_n_tiles) * BN, m0 = (tile / p.num_n_tiles) * TILE_M + rank * BM;
      if (EPI == HCT_EPI_RES_F32 || EpiTraits<EPI>::uses_aux) {
        // the residual / pre-activation tile this warp will stream in its epilogue: pull it into L2 while the
        // main loop of this tile is still running (one 128-byte line per lane and step)
        const int prow = m0 + q * 32 + lane;
        const int pcol = n0 + half * (BN / 2);
        if (prow < p.M && pcol < p.N) {
          // one bulk L2 prefetch per row segment (a plain prefetch.global.L2 only pulls the addressed sector)
          const int ncols = min(BN / 2, p.N - pcol);
          if (EPI == HCT_EPI_RES_F32)
            prefetch_l2_bulk(p.res + static_cast<long long>(prow) * p.ldres + pcol, static_cast<uint32_t>(ncols) * 4u);
          else
            prefetch_l2_bulk(p.aux + static_cast<long long>(prow) * p.ldaux + pcol, static_cast<uint32_t>(ncols) * 2u);
        }
      }
      // the bf16 multiplicand of a MUL / DGELU epilogue does not depend on the accumulator: the first unit's rows are
      // requested before waiting for the tile, every further unit's while the previous one is drained
      uint2 aux_cur[8], aux_nxt[8];
      bool have_cur = false;
      if (EpiTraits<EPI>::uses_aux) {
        const int colw0 = n0 + half * (BN / 2);
        have_cur = unit_interior<EPI>(p, m0 + q * 32, colw0);
        if (have_cur) load_aux_unit(p, lane, m0 + q * 32, colw0, aux_cur);
      }
      const int tix = (w - unit) / num_units;
      long long* tr = (trace && warp == FIRST_EPI_WARP && lane == 0 && tix < 64) ? trace + 512 + tix * 16 : nullptr;
      if (tr) tr[0] = clock64();
      mbar_wait(&tfull_bar[acc], acc_phase);
      if (tr) tr[1] = clock64();
      tc_fence_after();
      // software pipeline over the warp's four 32-column units: the TMEM load of unit c+1 is in flight while unit c
      // is drained from the staging buffer; the accumulator stage is handed back to the MMA warp as soon as the
      // last load has landed (before the last drain).
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + half * (BN / 2);
      const int colw = n0 + half * (BN / 2);
      const bool tile_ok = m0 < p.M;
      int nunits = 0;
      if (tile_ok && colw < p.N) nunits = min(4, (p.N - colw + 31) / 32);
      uint32_t v[32];
      if (nunits > 0) tmem_ld32_issue(t0, v);
#pragma unroll 1
      for (int c = 0; c < nunits; ++c) {
        const float4 bias4 = load_bias_unit<EPI>(p, lane, colw + c * 32);   // requested before the TMEM wait: off the chain
        tmem_ld_wait();
        if (tr) tr[2 + c * 3] = clock64();
        epilogue_stage(stg, lane, v);                      // v is dead after staging: reuse it for the next load
        if (tr) tr[3 + c * 3] = clock64();
        if (c + 1 < nunits) {
          tmem_ld32_issue(t0 + (c + 1) * 32, v);
        } else {
          tc_fence_before(); __syncwarp();
          if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
        }
        bool have_nxt = false;
        if (EpiTraits<EPI>::uses_aux && c + 1 < nunits) {
          have_nxt = unit_interior<EPI>(p, m0 + q * 32, colw + (c + 1) * 32);
          if (have_nxt) load_aux_unit(p, lane, m0 + q * 32, colw + (c + 1) * 32, aux_nxt);
        }
        epilogue_drain<EPI>(p, stg, lane, m0 + q * 32, colw + c * 32, bias4, aux_cur, have_cur);
        if (EpiTraits<EPI>::uses_aux) {
#pragma unroll
          for (int i = 0; i < 8; ++i) aux_cur[i] = aux_nxt[i];
          have_cur = have_nxt;
        }
        if (tr) tr[4 + c * 3] = clock64();
      }
      if (nunits == 0) {
        tc_fence_before(); __syncwarp();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
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
              int box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
              CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) { hct_set_error("cuTensorMapEncodeTiled entry point unavailable"); return HCT_ERR_CUDA; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * (dtype == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    hct_set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%lld outer=%lld ld=%lld box=%dx%d", (int)r, base,
                  inner, outer, ld, box_inner, box_outer);
    return HCT_ERR_CUDA;
  }
  return HCT_OK;
}

struct EpiMaps { CUtensorMap out, out2, aux; };

template <int EPI, int CTAS>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em, const GemmParams& p, int grid, cudaStream_t stream) {
  static bool configured = false;
  auto kernel = hct_gemm_tcgen05_kernel<EPI, CTAS>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<CTAS, EPI>::SMEM);
    if (e != cudaSuccess) { hct_set_error("cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg<CTAS, EPI>::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // the kernel waits itself (pdl_wait)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = hct_pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, em.out, em.out2, em.aux, p);
  if (e != cudaSuccess) { hct_set_error("cudaLaunchKernelEx(gemm): %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return HCT_ERR_CUDA; }
  return hct_check_launch("hct_gemm_tcgen05_kernel");
}

template <int CTAS>
int dispatch(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em, const GemmParams& p, int grid, cudaStream_t st) {
  switch (epi) {
    case HCT_EPI_BF16: return launch<HCT_EPI_BF16, CTAS>(tmA, tmB, em, p, grid, st);
    case HCT_EPI_GELU_BF16: return launch<HCT_EPI_GELU_BF16, CTAS>(tmA, tmB, em, p, grid, st);
    case HCT_EPI_RES_F32: return launch<HCT_EPI_RES_F32, CTAS>(tmA, tmB, em, p, grid, st);
    case HCT_EPI_POS_F32: return launch<HCT_EPI_POS_F32, CTAS>(tmA, tmB, em, p, grid, st);
    case HCT_EPI_DGELU_BF16: return launch<HCT_EPI_DGELU_BF16, CTAS>(tmA, tmB, em, p, grid, st);
    case HCT_EPI_F32: return launch<HCT_EPI_F32, CTAS>(tmA, tmB, em, p, grid, st);
    case HCT_EPI_GELU_DERIV_BF16: return launch<HCT_EPI_GELU_DERIV_BF16, CTAS>(tmA, tmB, em, p, grid, st);
    case HCT_EPI_MUL_BF16: return launch<HCT_EPI_MUL_BF16, CTAS>(tmA, tmB, em, p, grid, st);
    default: return launch<HCT_EPI_ATOMIC_F32, CTAS>(tmA, tmB, em, p, grid, st);
  }
}

int g_gemm_ctas = 2;   // CTA-pair mode by default; hct_gemm_set_cta_pair(0) selects the single-CTA kernel

}  // namespace

int hct_make_tmap_bf16_2d(CUtensorMap* tm, const void* base, long long inner, long long outer, long long ld_elems,
                          int box_inner, int box_outer) {
  return make_tmap(tm, base, inner, outer, ld_elems, box_inner, box_outer);
}

int hct_make_tmap_bf16_2d_sw(CUtensorMap* tm, const void* base, long long inner, long long outer, long long ld_elems,
                             int box_inner, int box_outer, int swizzle128) {
  return make_tmap(tm, base, inner, outer, ld_elems, box_inner, box_outer,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

extern "C" int hct_gemm_set_cta_pair(int enable) { g_gemm_ctas = enable ? 2 : 1; return HCT_OK; }
extern "C" int hct_gemm_trace(void* buf) {     // device buffer of >= 1536 int64 (or NULL): timeline of CTA 0
  long long* p = static_cast<long long*>(buf);
  return cudaMemcpyToSymbol(g_gemm_trace, &p, sizeof(p)) == cudaSuccess ? HCT_OK : HCT_ERR_CUDA;
}

extern "C" int hct_gemm_bf16(const hct_gemm_desc* d, hct_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HCT_REQUIRE(d != nullptr, "hct_gemm_bf16: null descriptor");
  HCT_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "hct_gemm_bf16: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
  HCT_REQUIRE(d->N % 8 == 0, "hct_gemm_bf16: N=%d must be a multiple of 8", d->N);
  HCT_REQUIRE(d->lda % 8 == 0 && d->ldb % 8 == 0, "hct_gemm_bf16: lda/ldb must be multiples of 8 elements");
  HCT_REQUIRE((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->B) & 15) == 0,
              "hct_gemm_bf16: A/B must be 16-byte aligned");
  HCT_REQUIRE(d->out != nullptr && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0, "hct_gemm_bf16: out misaligned");
  HCT_REQUIRE(d->epilogue >= 0 && d->epilogue <= HCT_EPI_MUL_BF16, "hct_gemm_bf16: bad epilogue %d", d->epilogue);
  const bool f32_out = d->epilogue == HCT_EPI_RES_F32 || d->epilogue == HCT_EPI_POS_F32 ||
                       d->epilogue == HCT_EPI_F32 || d->epilogue == HCT_EPI_ATOMIC_F32;
  HCT_REQUIRE(d->ldo % (f32_out ? 4 : 8) == 0, "hct_gemm_bf16: ldo=%lld breaks 16-byte row alignment", (long long)d->ldo);
  if (d->epilogue == HCT_EPI_RES_F32)
    HCT_REQUIRE(d->res != nullptr && d->ldres % 4 == 0, "hct_gemm_bf16: RES epilogue needs res");
  if (d->epilogue == HCT_EPI_POS_F32)
    HCT_REQUIRE(d->pos != nullptr && d->ldpos % 4 == 0 && (d->pos_idx != nullptr || d->pos_period > 0),
                "hct_gemm_bf16: POS epilogue needs pos and pos_idx/pos_period");
  if (d->epilogue == HCT_EPI_DGELU_BF16 || d->epilogue == HCT_EPI_MUL_BF16)
    HCT_REQUIRE(d->aux != nullptr && d->ldaux % 8 == 0, "hct_gemm_bf16: DGELU / MUL epilogue needs aux");
  if (d->epilogue == HCT_EPI_GELU_DERIV_BF16)
    HCT_REQUIRE(d->out2 != nullptr, "hct_gemm_bf16: GELU_DERIV epilogue needs out2");
  if ((d->epilogue == HCT_EPI_GELU_BF16 || d->epilogue == HCT_EPI_GELU_DERIV_BF16) && d->out2 != nullptr)
    HCT_REQUIRE(d->ldo2 % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out2) & 15) == 0, "hct_gemm_bf16: out2 misaligned");
  HCT_REQUIRE(d->colsum == nullptr || !f32_out, "hct_gemm_bf16: colsum is only available with bf16-output epilogues");
  if (d->a_mn_major) HCT_REQUIRE(d->M % 8 == 0, "hct_gemm_bf16: MN-major A needs M %% 8 == 0");
  else HCT_REQUIRE(d->K % 8 == 0, "hct_gemm_bf16: K-major A needs K %% 8 == 0");
  if (!d->b_mn_major) HCT_REQUIRE(d->K % 8 == 0, "hct_gemm_bf16: K-major B needs K %% 8 == 0");

  const int ctas = g_gemm_ctas;
  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_mn = d->a_mn_major ? 1 : 0; p.b_mn = d->b_mn_major ? 1 : 0;
  p.num_m_tiles = (d->M + BM * ctas - 1) / (BM * ctas);
  p.num_n_tiles = (d->N + BN - 1) / BN;
  p.total_kb = (d->K + BK - 1) / BK;
  const int sms = hct_num_sms();
  const int units = sms / ctas;
  int splits = 1;
  if (d->epilogue == HCT_EPI_ATOMIC_F32) {
    splits = d->splits;
    if (splits <= 0) {
      // pick the split factor whose work-item count fills whole waves of CTA (pairs): fewest splits with >= 95 %
      // wave efficiency, else the most efficient one; keep >= 8 k-blocks per split.
      const int tiles = p.num_m_tiles * p.num_n_tiles;
      int max_splits = (p.total_kb + 7) / 8;
      if (max_splits > 64) max_splits = 64;
      if (max_splits < 1) max_splits = 1;
      double best_eff = -1.0;
      splits = 1;
      for (int sp = 1; sp <= max_splits; ++sp) {
        const long long items = static_cast<long long>(tiles) * sp;
        const long long waves = (items + units - 1) / units;
        const double eff = static_cast<double>(items) / static_cast<double>(waves * units);
        if (eff > best_eff + 1e-9) { best_eff = eff; splits = sp; }
        if (eff >= 0.95) { splits = sp; break; }
      }
    }
    if (splits > p.total_kb) splits = p.total_kb;
  }
  p.kb_per_split = (p.total_kb + splits - 1) / splits;
  p.splits = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;   // no empty split
  // instruction descriptor: D=f32, A=B=bf16, majorness, N>>3, M>>4 (M = 256 for the CTA pair)
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.a_mn) << 15) |
            (static_cast<uint32_t>(p.b_mn) << 16) | (static_cast<uint32_t>(BN >> 3) << 17) |
            (static_cast<uint32_t>((BM * ctas) >> 4) << 24);
  p.out = d->out; p.ldo = d->ldo; p.out2 = d->out2; p.ldo2 = d->ldo2;
  p.bias = d->bias; p.res = d->res; p.ldres = d->ldres;
  p.aux = static_cast<const bf16*>(d->aux); p.ldaux = d->ldaux;
  p.pos = d->pos; p.ldpos = d->ldpos; p.pos_idx = d->pos_idx; p.pos_period = d->pos_period;
  p.rows_in = d->rows_in; p.rows_out = d->rows_out; p.row_off = d->row_off;
  p.alpha = d->alpha == 0.0f ? 1.0f : d->alpha;
  p.colsum = d->colsum;

  CUtensorMap tmA, tmB;
  int rc;
  if (!p.a_mn) rc = make_tmap(&tmA, d->A, d->K, d->M, d->lda, BK, BM);
  else rc = make_tmap(&tmA, d->A, d->M, d->K, d->lda, 64, BK);
  if (rc != HCT_OK) return rc;
  if (!p.b_mn) rc = make_tmap(&tmB, d->B, d->K, d->N, d->ldb, BK, BN / ctas);
  else rc = make_tmap(&tmB, d->B, d->N, d->K, d->ldb, 64, BK);
  if (rc != HCT_OK) return rc;

  // bf16-output epilogues store (and read their bf16 multiplicand) through TMA: 32 x 32 boxes, 64-byte swizzle
  EpiMaps em;
  em.out = tmA; em.out2 = tmA; em.aux = tmA;
  const bool tma_epi = d->epilogue == HCT_EPI_BF16 || d->epilogue == HCT_EPI_GELU_BF16 ||
                       d->epilogue == HCT_EPI_GELU_DERIV_BF16 || d->epilogue == HCT_EPI_MUL_BF16;
  if (d->epilogue == HCT_EPI_RES_F32) {        // fp32 boxes of 16 columns x 32 rows (64-byte rows as well)
    HCT_REQUIRE(d->rows_in <= 0, "hct_gemm_bf16: row remapping is only available with the POS_F32 / F32 epilogues");
    HCT_REQUIRE((reinterpret_cast<uintptr_t>(d->res) & 15) == 0, "hct_gemm_bf16: res misaligned");
    rc = make_tmap(&em.out, d->out, d->N, d->M, d->ldo, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (rc != HCT_OK) return rc;
    rc = make_tmap(&em.aux, d->res, d->N, d->M, d->ldres, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (rc != HCT_OK) return rc;
  }
  if (tma_epi) {
    HCT_REQUIRE(d->rows_in <= 0, "hct_gemm_bf16: row remapping is only available with the POS_F32 / F32 epilogues");
    rc = make_tmap(&em.out, d->out, d->N, d->M, d->ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != HCT_OK) return rc;
    if (d->out2 != nullptr && (d->epilogue == HCT_EPI_GELU_BF16 || d->epilogue == HCT_EPI_GELU_DERIV_BF16)) {
      rc = make_tmap(&em.out2, d->out2, d->N, d->M, d->ldo2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != HCT_OK) return rc;
    }
    if (d->epilogue == HCT_EPI_MUL_BF16) {
      rc = make_tmap(&em.aux, d->aux, d->N, d->M, d->ldaux, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != HCT_OK) return rc;
    }
  }

  const int total_work = p.num_m_tiles * p.num_n_tiles * p.splits;
  const int grid = (total_work < units ? total_work : units) * ctas;
  void* prof = hct_prof_enabled() ? hct_prof_begin(stream) : nullptr;
  rc = ctas == 2 ? dispatch<2>(d->epilogue, tmA, tmB, em, p, grid, stream) : dispatch<1>(d->epilogue, tmA, tmB, em, p, grid, stream);
  if (prof != nullptr) hct_prof_end(prof, stream, 2.0 * d->M * d->N * static_cast<double>(d->K));
  return rc;
}
