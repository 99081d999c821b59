// Flash attention on tcgen05 / TMEM for sm_100a (F.scaled_dot_product_attention, attentionblock.py:61).
// Forward and backward kernels below; every matmul runs on the tensor core with fp32 accumulators in tensor memory,
// the softmax threads own one row (TMEM lane) each, and the bf16 probabilities / score gradients they produce go back
// to the tensor core through tensor memory (A-from-TMEM MMA), never through shared memory.
// Roles inside a CTA are warp-specialised and mbarrier-synchronised: TMA producer warp, MMA issuer warp (both run as
// converged warps and issue under elect_one), softmax warps.
#include "../../include/hct_b200.h"
#include "hct_tcgen05.cuh"

namespace {
using namespace hct_tc;

constexpr int TILE = 128;                  // query rows per CTA, keys per inner block
constexpr int TILE_BYTES = TILE * 128;     // 128 rows x 64 bf16
constexpr float LOG2E = 1.4426950408889634f;
__device__ long long* g_trace = nullptr;   // event timeline of one CTA (tools/attn_dbg.py); nullptr in production
#define TRACE(slot) do { if (trace) trace[slot] = clock64(); } while (0)

__device__ __forceinline__ float ex2f(float x) {   // bare MUFU.EX2 (ftz); ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int FWD_THREADS = 192;
// Warp roles: warps 0-3 softmax (TMEM lane quarter = warp id), then the two control warps.
constexpr int FWD_PRODUCER_WARP = 4, FWD_MMA_WARP = 5;
constexpr int FWD_TK = 64;                 // keys per inner block
constexpr int FWD_TMEM_COLS = 128;         // S: cols [0,64) (P overwrites [0,32) in place), O: cols [64, 64+hd)  -> four CTAs per SM
constexpr int FWD_KV_BYTES = FWD_TK * 128; // [64 keys][64 bf16]
constexpr int FWD_SMEM = TILE_BYTES + 4 * FWD_KV_BYTES + 1024 + 128;   // Q, 2 x K, 2 x V + align slack + barriers

// A operand from tensor memory: D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Forward: one CTA per (batch, head, 128-query tile), four CTAs per SM.
//   producer warp : Q once, then K_j / V_j tiles (64 keys, double-buffered) straight out of the qkv GEMM output
//   MMA warp      : S_0 = Q K_0^T;  then per block [O += P_j V_j (A = P_j from TMEM), S_{j+1} = Q K_{j+1}^T] back to back
//   warps 0-3     : one query row per thread: tcgen05.ld S_j, online max / exp2 / sum, P_j written as bf16 pairs IN PLACE
//                   over the first 32 of the 64 fp32 columns the thread has just read (tcgen05.st), O rescaled in TMEM
//                   when the running max moves, final O / l and log-sum-exp written out.
// S, P and O never leave the SM; P never touches shared memory.
template <int HD, int POLY>   // POLY of every 8 exponential pairs run on the FMA pipe (hct_tcgen05.cuh)
__global__ void __launch_bounds__(FWD_THREADS, 4)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const bf16* __restrict__ qkv,
                   bf16* __restrict__ out, float* __restrict__ lse, int S, int H, float scale, int fold_flags) {
  __shared__ float s_kt[64], s_vt[64];                       // the key / value row behind the last 64-key block (tail1)
  __shared__ float s_qt[64], s_p[64];                        // the query row behind the last full tile, its probabilities
  const int fold_tail_key = fold_flags & 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;                   // 2 x [64 keys][64]
  uint8_t* sV = sK + 2 * FWD_KV_BYTES;               // 2 x [64 keys][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * FWD_KV_BYTES);
  uint64_t *q_full = bars, *k_full = bars + 1 /*[2]*/, *k_empty = bars + 3 /*[2]*/, *v_full = bars + 5 /*[2]*/,
           *v_empty = bars + 7 /*[2]*/, *s_full = bars + 9, *p_full = bars + 10, *done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
  const int D = H * HD;
  const int nkb = (S + FWD_TK - 1) / FWD_TK;
  const bool tail16 = S - (nkb - 1) * FWD_TK <= 16;          // last key block is computed 16 columns wide
  // S = 64 k + 1 (the cls token): the ONE key behind the last full block does not get a chain step of its own (S MMA ->
  // softmax -> P V MMA for a single column: 1/9 of a tile's run time at S = 513, 1/3 at S = 129).  Its score is a dot
  // product per row, its contribution a rank-1 update; both are folded into the epilogue on the CUDA cores.
  const bool tail1 = fold_tail_key != 0 && nkb >= 2 && S - (nkb - 1) * FWD_TK == 1;
  const int nkb_eff = tail1 ? nkb - 1 : nkb;
  const float sl2 = scale * LOG2E;
  // S = 128 k + 1: the grid then holds the FULL query tiles only (host side), and the ONE query row behind them rides with
  // the CTA of the last full tile of its head: that CTA's producer warp, idle between its TMA issues, runs the row's online
  // softmax on the CUDA cores against the K_j / V_j blocks that pass through shared memory anyway.  As a tile of its own the
  // row cost a whole CTA (nine MMA -> softmax -> MMA steps with one live row): 0.51 against 0.41 ms at S = 513 / 512.
  const bool has_tail_row = (fold_flags & 2) != 0 && qt == static_cast<int>(gridDim.x) - 1;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, FWD_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                  // launched with programmatic stream serialization: nothing above touches global memory
  pdl_launch_dependents();
  if ((tail1 || has_tail_row) && threadIdx.x < HD) {
    const bf16* row = qkv + (static_cast<long long>(b) * S + (S - 1)) * (3LL * D) + h * HD + threadIdx.x;
    s_qt[threadIdx.x] = __bfloat162float(row[0]);
    s_kt[threadIdx.x] = __bfloat162float(row[D]);
    s_vt[threadIdx.x] = __bfloat162float(row[2 * D]);
  }
  if (tail1 || has_tail_row) __syncthreads();                // CTA-uniform
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64;

  if (warp == FWD_PRODUCER_WARP) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(smem_u32(sQ), &tmQ, q_full, h * HD, b * S + qt * TILE);
    }
    // online softmax of the tail query row (has_tail_row): keys lane and lane + 32 of a block per lane, output dims 2 lane, 2 lane + 1
    float tr_m = -INFINITY, tr_l = 0.f, tr_o0 = 0.f, tr_o1 = 0.f;
    const int keys_in_blocks = tail1 ? S - 1 : S;
    auto tail_row_block = [&](int j) {
      const int st = j & 1;
      mbar_wait(&k_full[st], (j >> 1) & 1);
      mbar_wait(&v_full[st], (j >> 1) & 1);
      const int nvalid = min(FWD_TK, keys_in_blocks - j * FWD_TK);
      const uint8_t* kb = sK + st * FWD_KV_BYTES;
      const uint8_t* vb = sV + st * FWD_KV_BYTES;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const float4 qa = *reinterpret_cast<const float4*>(s_qt + 8 * c), qb = *reinterpret_cast<const float4*>(s_qt + 8 * c + 4);
        const uint4 u0 = *reinterpret_cast<const uint4*>(kb + lane * 128 + ((c ^ (lane & 7)) << 4));
        const uint4 u1 = *reinterpret_cast<const uint4*>(kb + (lane + 32) * 128 + ((c ^ (lane & 7)) << 4));
        float2 f;
        f = unpack_bf16x2(u0.x); s0 = fmaf(qa.x, f.x, s0); s0 = fmaf(qa.y, f.y, s0);
        f = unpack_bf16x2(u0.y); s0 = fmaf(qa.z, f.x, s0); s0 = fmaf(qa.w, f.y, s0);
        f = unpack_bf16x2(u0.z); s0 = fmaf(qb.x, f.x, s0); s0 = fmaf(qb.y, f.y, s0);
        f = unpack_bf16x2(u0.w); s0 = fmaf(qb.z, f.x, s0); s0 = fmaf(qb.w, f.y, s0);
        f = unpack_bf16x2(u1.x); s1 = fmaf(qa.x, f.x, s1); s1 = fmaf(qa.y, f.y, s1);
        f = unpack_bf16x2(u1.y); s1 = fmaf(qa.z, f.x, s1); s1 = fmaf(qa.w, f.y, s1);
        f = unpack_bf16x2(u1.z); s1 = fmaf(qb.x, f.x, s1); s1 = fmaf(qb.y, f.y, s1);
        f = unpack_bf16x2(u1.w); s1 = fmaf(qb.z, f.x, s1); s1 = fmaf(qb.w, f.y, s1);
      }
      if (lane >= nvalid) s0 = -INFINITY;
      if (lane + 32 >= nvalid) s1 = -INFINITY;
      const float mn = fmaxf(tr_m, warp_max(fmaxf(s0, s1)));
      const float alpha = ex2f((tr_m - mn) * sl2);            // first block: ex2(-inf) = 0
      const float p0 = ex2f((s0 - mn) * sl2), p1 = ex2f((s1 - mn) * sl2);
      tr_l = tr_l * alpha + warp_sum(p0 + p1);
      tr_m = mn;
      s_p[lane] = p0;
      s_p[lane + 32] = p1;
      __syncwarp();
      tr_o0 *= alpha;
      tr_o1 *= alpha;
      if (lane < HD / 2) {
#pragma unroll 8
        for (int k = 0; k < FWD_TK; ++k) {
          const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vb + k * 128 + (((lane >> 2) ^ (k & 7)) << 4) + (lane & 3) * 4));
          const float pk = s_p[k];
          tr_o0 = fmaf(pk, v2.x, tr_o0);
          tr_o1 = fmaf(pk, v2.y, tr_o1);
        }
      }
      __syncwarp();
    };
    for (int j = 0; j < nkb_eff; ++j) {
      const int st = j & 1;
      const uint32_t ph = ((j >> 1) & 1) ^ 1u;
      mbar_wait(&k_empty[st], ph);
      if (leader) {
        mbar_expect_tx(&k_full[st], FWD_KV_BYTES);
        tma_load_2d(smem_u32(sK + st * FWD_KV_BYTES), &tmKV, &k_full[st], D + h * HD, b * S + j * FWD_TK);
      }
      mbar_wait(&v_empty[st], ph);
      if (leader) {
        mbar_expect_tx(&v_full[st], FWD_KV_BYTES);
        tma_load_2d(smem_u32(sV + st * FWD_KV_BYTES), &tmKV, &v_full[st], 2 * D + h * HD, b * S + j * FWD_TK);
      }
      __syncwarp();
      // block j - 1 for the tail query row: done before this warp issues the load that overwrites that stage (block j + 1)
      if (has_tail_row && j >= 1) tail_row_block(j - 1);
    }
    if (has_tail_row) {
      tail_row_block(nkb_eff - 1);
      if (tail1) {                                            // the row's own score against the folded tail key
        float part = 0.f;
        for (int d = lane; d < HD; d += 32) part = fmaf(s_qt[d], s_kt[d], part);
        const float s_t = warp_sum(part);
        const float mn = fmaxf(tr_m, s_t);
        const float alpha = ex2f((tr_m - mn) * sl2), p_t = ex2f((s_t - mn) * sl2);
        tr_l = tr_l * alpha + p_t;
        tr_m = mn;
        if (lane < HD / 2) { tr_o0 = fmaf(p_t, s_vt[2 * lane], tr_o0 * alpha); tr_o1 = fmaf(p_t, s_vt[2 * lane + 1], tr_o1 * alpha); }
      }
      const float inv = 1.0f / tr_l;
      if (lane < HD / 2)
        *reinterpret_cast<uint32_t*>(out + (static_cast<long long>(b) * S + (S - 1)) * D + h * HD + 2 * lane) =
            pack_bf16x2(tr_o0 * inv, tr_o1 * inv);
      if (lane == 0) lse[(static_cast<long long>(b) * H + h) * S + (S - 1)] = tr_m * scale + logf(tr_l);
    }
  } else if (warp == FWD_MMA_WARP) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(TILE, FWD_TK, false, false);
    const uint32_t idesc_s16 = make_idesc_bf16(TILE, 16, false, false);
    const uint32_t idesc_o = make_idesc_bf16(TILE, HD, false, true);
    const uint64_t dQ = make_sdesc_sw128(smem_u32(sQ), false, 0);
    mbar_wait(q_full, 0);
    // A last block with <= 16 valid keys (the cls token makes S = 64 k + 1) is computed 16 keys wide.
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(&k_full[st], (j >> 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t id = (j == nkb - 1 && tail16) ? idesc_s16 : idesc_s;
        const uint64_t dK = make_sdesc_sw128(smem_u32(sK + st * FWD_KV_BYTES), false, 0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) tc_mma(tmem_S, dQ + ks * 2, dK + ks * 2, id, ks > 0 ? 1u : 0u);
        tc_commit(s_full);
        tc_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < nkb_eff; ++j) {
      const int st = j & 1;
      mbar_wait(p_full, j & 1);                       // P_j sits in TMEM (and every softmax warp has read S_j)
      mbar_wait(&v_full[st], (j >> 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t dV = make_sdesc_sw128(smem_u32(sV + st * FWD_KV_BYTES), true, FWD_KV_BYTES);
        const uint32_t acc = j > 0 ? 1u : 0u;
        if (j == nkb - 1 && tail16) {
          tc_mma_ts(tmem_O, tmem_S, dV, idesc_o, acc);
        } else {
#pragma unroll
          for (int ks = 0; ks < FWD_TK / 16; ++ks) tc_mma_ts(tmem_O, tmem_S + ks * 8, dV + ks * 128, idesc_o, ks > 0 ? 1u : acc);
        }
        tc_commit(&v_empty[st]);
        if (j == nkb_eff - 1) tc_commit(done);
      }
      __syncwarp();
      if (j + 1 < nkb_eff) issue_s(j + 1);            // executes after P_j V_j (in-order pipe): safe to overwrite S / P
    }
  } else if (warp < 4) {
    // ===================== softmax / correction / epilogue: one query row per thread =====================
    const int q = warp & 3;
    const int row = qt * TILE + q * 32 + lane;
    const bool warp_active = qt * TILE + q * 32 < S;   // a warp whose 32 rows all lie past S only keeps the barriers moving
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float m = -INFINITY, l = 0.f;
    // Lazy rescaling: the reference point m only moves when the block maximum exceeds it by more than 2^8 in the
    // exponent domain (or on the first block).  Probabilities are then at most 2^8 -- harmless in fp32 / bf16 --
    // and the O accumulator in TMEM (whose read-back is the scarce resource here) is rescaled only rarely.
    auto move_max = [&](int j, float mb) -> float {
      float alpha = 1.0f;
      if (j == 0) {
        m = mb;
      } else if ((mb - m) * sl2 > 8.0f) {
        alpha = ex2f((m - mb) * sl2);
        m = mb;
      }
      return alpha;
    };
    // S_j complete implies P_{j-1} V_{j-1} complete (one in-order pipe, one commit): O is stable until this warp arrives
    auto rescale_o = [&](float alpha) {
      uint32_t o[16];
#pragma unroll 1
      for (int c0 = 0; c0 < HD; c0 += 16) {
        tmem_ld16(tmem_O + lane_off + c0, o);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st16(tmem_O + lane_off + c0, o);
      }
    };
    for (int j = 0; j < nkb_eff; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int nvalid = min(FWD_TK, S - j * FWD_TK);         // valid key columns in this block (>= 1)
      if (warp_active) {
        if (j == nkb - 1 && tail16) {
          // ---- 16-column tail block
          uint32_t v[16];
          tmem_ld16(tmem_S + lane_off, v);
          float mb = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) mb = fmaxf(mb, i < nvalid ? __uint_as_float(v[i]) : -INFINITY);
          const float alpha = move_max(j, mb);
          const float msc = m * sl2;
          float rsum = 0.f;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float a0 = i < nvalid ? ex2f(fmaf(__uint_as_float(v[i]), sl2, -msc)) : 0.f;
            const float a1 = i + 1 < nvalid ? ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -msc)) : 0.f;
            rsum += a0 + a1;
            pk[i >> 1] = pack_bf16x2(a0, a1);
          }
#pragma unroll
          for (int i = 8; i < 16; ++i) pk[i] = 0u;
          l = l * alpha + rsum;
          if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) rescale_o(alpha);
          tmem_st16(tmem_S + lane_off, pk);
        } else {
          if constexpr (POLY < 0) {
          // ---- full 64-column block (a partial one is padded with -inf first): scalar fp32, one pass
          uint32_t v0[32], v1[32];
          tmem_ld32_issue(tmem_S + lane_off, v0);
          tmem_ld32_issue(tmem_S + lane_off + 32, v1);
          tmem_ld_wait();
          if (nvalid < FWD_TK) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i >= nvalid) v0[i] = 0xff800000u;
              if (i + 32 >= nvalid) v1[i] = 0xff800000u;
            }
          }
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
#pragma unroll
            for (int t = 0; t < 4; ++t) mx[t] = fmaxf(mx[t], fmaxf(__uint_as_float(v0[i + t]), __uint_as_float(v1[i + t])));
          }
          const float mb = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          const float alpha = move_max(j, mb);
          const float msc = m * sl2;
          float rs0 = 0.f, rs1 = 0.f;
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = ex2f(fmaf(__uint_as_float(v0[i]), sl2, -msc)), a1 = ex2f(fmaf(__uint_as_float(v0[i + 1]), sl2, -msc));
            const float b0 = ex2f(fmaf(__uint_as_float(v1[i]), sl2, -msc)), b1 = ex2f(fmaf(__uint_as_float(v1[i + 1]), sl2, -msc));
            rs0 += a0 + a1;
            rs1 += b0 + b1;
            pk[i >> 1] = pack_bf16x2(a0, a1);
            pk[16 + (i >> 1)] = pack_bf16x2(b0, b1);
          }
          l = l * alpha + (rs0 + rs1);
          if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) rescale_o(alpha);
          tmem_st32(tmem_S + lane_off, pk);                   // keys 2c, 2c+1 -> column c: the A operand of P_j V_j
          } else {
          // ---- full 64-column block (a partial one is padded with -inf first).  Two passes over the row's scores: the
          // maximum first, then the exponentials 32 columns at a time -- re-reading tensor memory is cheaper than holding 64
          // scores plus the packed-pair temporaries in 80 registers (the one-pass form spilled, 0.505 -> 0.58 ms).
          float mb;
          {
            uint32_t v0[32], v1[32];
            tmem_ld32_issue(tmem_S + lane_off, v0);
            tmem_ld32_issue(tmem_S + lane_off + 32, v1);
            tmem_ld_wait();
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float x0 = i + t < nvalid ? __uint_as_float(v0[i + t]) : -INFINITY;
                const float x1 = i + t + 32 < nvalid ? __uint_as_float(v1[i + t]) : -INFINITY;
                mx[t] = fmaxf(mx[t], fmaxf(x0, x1));
              }
            }
            mb = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          }
          const float alpha = move_max(j, mb);
          const float msc = m * sl2;
          // packed fp32 (one issue slot per two elements) for scale-and-shift and row sums
          const float2 sl2v = make_float2(sl2, sl2), nmsc = make_float2(-msc, -msc);
          float2 rs = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t v[32], pk[16];
            tmem_ld32(tmem_S + lane_off + hf * 32, v);
            const int lim = nvalid - hf * 32;
            if (lim < 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i >= lim) v[i] = 0xff800000u;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 a = ex2_pair<(POLY < 0 ? 0 : POLY)>(__ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sl2v, nmsc), i >> 1);
              rs = __fadd2_rn(rs, a);
              pk[i >> 1] = pack_bf16x2(a.x, a.y);
            }
            // keys 2c, 2c+1 -> column c: the A operand of P_j V_j.  Half 0 lands in columns [0,16) (scores it has read),
            // half 1 in [16,32); the second half's scores sit in columns [32,64), untouched by either store
            tmem_st16(tmem_S + lane_off + hf * 16, pk);
          }
          l = l * alpha + (rs.x + rs.y);
          if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) rescale_o(alpha);
        }
          }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16, log-sum-exp
    // the tail key (tail1): score = q . k_tail on the CUDA cores while the last P V MMAs drain
    float s_t = 0.f;
    if (tail1) {
      if (warp_active) {
        const int rr = q * 32 + lane;
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(sQ + rr * 128 + ((c ^ (rr & 7)) << 4));
          const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
          const float* kk = s_kt + 8 * c;
          s_t = fmaf(a0.x, kk[0], s_t); s_t = fmaf(a0.y, kk[1], s_t); s_t = fmaf(a1.x, kk[2], s_t); s_t = fmaf(a1.y, kk[3], s_t);
          s_t = fmaf(a2.x, kk[4], s_t); s_t = fmaf(a2.y, kk[5], s_t); s_t = fmaf(a3.x, kk[6], s_t); s_t = fmaf(a3.y, kk[7], s_t);
        }
      }
    }
    mbar_wait(done, 0);
    tc_fence_after();
    if (warp_active) {
      float alpha_t = 1.0f, p_t = 0.f;                        // O_final = (O alpha_t + p_t v_tail) / l
      if (tail1) {
        const float mn = fmaxf(m, s_t);
        alpha_t = ex2f((m - mn) * sl2);
        p_t = ex2f((s_t - mn) * sl2);
        l = l * alpha_t + p_t;
        m = mn;
      }
      const float inv = 1.0f / l;
      const float oa = alpha_t * inv, pv = p_t * inv;
      bf16* orow = out + (static_cast<long long>(b) * S + row) * D + h * HD;
#pragma unroll 1
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t o[16];
        tmem_ld16(tmem_O + lane_off + c0, o);
        if (row < S) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f[i] = __uint_as_float(o[8 * g + i]) * oa;
              if (tail1) f[i] = fmaf(pv, s_vt[c0 + 8 * g + i], f[i]);
            }
            uint4 u;
            u.x = pack_bf16x2(f[0], f[1]);
            u.y = pack_bf16x2(f[2], f[3]);
            u.z = pack_bf16x2(f[4], f[5]);
            u.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(orow + c0 + 8 * g) = u;
          }
        }
      }
      if (row < S) lse[(static_cast<long long>(b) * H + h) * S + row] = m * scale + logf(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FWD_TMEM_COLS);
  }
}

// =====================================================================================================
// Backward.  Two kernels, both two CTAs per SM with a 256-column TMEM budget each; no atomics, deterministic.
// The bf16 operands the softmax threads produce (P^T, dS^T, dS) never touch shared memory: they are written to
// tensor memory with tcgen05.st and consumed as the A operand of tcgen05.mma (A-from-TMEM form), which removes the
// st.shared + proxy fence + 4 KiB-per-MMA shared-memory operand reads of the earlier version and frees the room for
// a 4-deep TMA ring (the 2-deep ring had the TMA round trip on the critical path, see tools/attn_dbg.py).
//
//   dK/dV kernel: CTA = (batch, head, 128 keys); rows (TMEM lanes) are KEYS.  Per 64-query block i:
//        S^T = K Q_i^T, dP^T = V dO_i^T                      -> TMEM cols [0,64) / [64,128)
//        P^T = exp2(S^T c - lse[q]),  dS^T = P^T (dP^T - delta[q])   written IN PLACE over the fp32 columns each
//                                                             softmax warp has just read (bf16 pairs, 16 columns)
//        dV += P^T dO_i,  dK += dS^T Q_i                      A from TMEM, B = the TMA-loaded dO / Q tiles read MN-major
//      The MMA warp issues [dV/dK(i), S^T/dP^T(i+1)] back to back: the tensor core executes in order, so block i+1
//      overwrites the S^T / dP^T columns only after block i's operands have been consumed.
//   dQ kernel:   CTA = (batch, head, 128 queries); rows are QUERIES.  Per 64-key block j:
//        S = Q K_j^T, dP = dO V_j^T -> TMEM; dS = P (dP - delta[row]) -> one of two TMEM operand buffers;
//        dQ += dS K_j (K tile read MN-major).  S/dP of block j+1 are issued as soon as block j sits in registers.
// =====================================================================================================
constexpr int HALF_BYTES = 64 * 128;       // 64 rows x 64 bf16
constexpr int BWD_THREADS = 320;           // 2 x 4 softmax warps (each group owns 32 of a block's 64 columns), TMA warp, MMA warp
constexpr int BWD_PRODUCER_WARP = 8, BWD_MMA_WARP = 9;
constexpr int BWD_TMEM_COLS = 256;
constexpr int BWD_STAGES = 4;              // TMA ring depth for the streamed 64-row tiles
constexpr int BWD_SMEM = 2 * TILE_BYTES + BWD_STAGES * 2 * HALF_BYTES + 8 * 512 + 1024 + 256;   // resident tiles, ring, per-warp stats, align, barriers

template <int HD>
__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                        const __grid_constant__ CUtensorMap tmDO64, const float* __restrict__ lse,
                        const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H, float scale, int merge_tail) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sK = smem;                               // [128 keys][64]
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;              // BWD_STAGES x [64 queries][64]
  uint8_t* sdO = sQ + BWD_STAGES * HALF_BYTES;      // BWD_STAGES x [64 queries][64]
  float* sStat = reinterpret_cast<float*>(sdO + BWD_STAGES * HALF_BYTES);   // [8 warps][2 buffers][lse 32 | delta 32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + 8 * 512);
  uint64_t *kv_full = bars, *qdo_full = bars + 1 /*[4]*/, *qdo_empty = bars + 5 /*[4]*/, *s_full = bars + 9,
           *p_full = bars + 10, *done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
  const int D = H * HD;
  const int nqb = (S + 63) / 64;
  const bool tail16 = S - (nqb - 1) * 64 <= 16;             // last query block is computed 16 columns wide
  // merged tail: the 16-wide last block is computed TOGETHER with block 0 -- its S^T / dP^T park in the dV / dK
  // accumulator columns, which are not live before block 0's second-stage MMAs -- so the MMA -> softmax -> MMA chain is
  // one step shorter (9 -> 8 at S = 513, 3 -> 2 at S = 129).  Ring entries: 0 = block 0, 1 = tail block, e >= 2 = block e - 1.
  const bool merge = merge_tail != 0 && tail16 && nqb >= 2;
  const int nsteps = merge ? nqb - 1 : nqb;
  const float sl2 = scale * LOG2E;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < BWD_STAGES; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, BWD_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64, tmem_dV = tmem_base + 128, tmem_dK = tmem_base + 192;
  long long* trace = (g_trace != nullptr && kt == 1 && h == 3 && b == gridDim.z / 2) ? g_trace : nullptr;
  // trace layout: [role 0 producer | 1 mma | 2 softmax warp 0 lane 0][block i][8 events]

  if (warp == BWD_PRODUCER_WARP) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_2d(smem_u32(sK), &tmQKV128, kv_full, D + h * HD, b * S + kt * TILE);
      tma_load_2d(smem_u32(sV), &tmQKV128, kv_full, 2 * D + h * HD, b * S + kt * TILE);
    }
    for (int i = 0; i < nqb; ++i) {                   // i = ring entry
      const int st = i % BWD_STAGES;
      const int blk = merge ? (i == 0 ? 0 : (i == 1 ? nqb - 1 : i - 1)) : i;
      TRACE(0 * 256 + i * 8 + 0);
      mbar_wait(&qdo_empty[st], ((i / BWD_STAGES) & 1) ^ 1u);
      TRACE(0 * 256 + i * 8 + 1);
      if (leader) {
        mbar_expect_tx(&qdo_full[st], 2 * HALF_BYTES);
        tma_load_2d(smem_u32(sQ + st * HALF_BYTES), &tmQKV64, &qdo_full[st], h * HD, b * S + blk * 64);
        tma_load_2d(smem_u32(sdO + st * HALF_BYTES), &tmDO64, &qdo_full[st], h * HD, b * S + blk * 64);
      }
      __syncwarp();
      TRACE(0 * 256 + i * 8 + 2);
    }
  } else if (warp == BWD_MMA_WARP) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(TILE, 64, false, false);
    const uint32_t idesc_s16 = make_idesc_bf16(TILE, 16, false, false);
    const uint32_t idesc_g = make_idesc_bf16(TILE, HD, false, true);
    const uint64_t dK_ = make_sdesc_sw128(smem_u32(sK), false, 0);
    const uint64_t dV_ = make_sdesc_sw128(smem_u32(sV), false, 0);
    mbar_wait(kv_full, 0);
    auto issue_sdp = [&](int i) {          // step i: S^T = K Q_i^T and dP^T = V dO_i^T into TMEM
      const int e = merge ? (i == 0 ? 0 : i + 1) : i;      // ring entry of the step's 64-wide block
      const int st = e % BWD_STAGES;
      TRACE(1 * 256 + i * 8 + 0);
      mbar_wait(&qdo_full[st], (e / BWD_STAGES) & 1);
      if (merge && i == 0) mbar_wait(&qdo_full[1], 0);     // the tail block's tiles (entry 1)
      TRACE(1 * 256 + i * 8 + 1);
      tc_fence_after();
      if (leader) {
        const uint32_t id = (!merge && i == nqb - 1 && tail16) ? idesc_s16 : idesc_s;
        const uint64_t dQk = make_sdesc_sw128(smem_u32(sQ + st * HALF_BYTES), false, 0);
        const uint64_t dOk = make_sdesc_sw128(smem_u32(sdO + st * HALF_BYTES), false, 0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {      // the two accumulate chains interleaved: consecutive MMAs independent
          tc_mma(tmem_S, dK_ + ks * 2, dQk + ks * 2, id, ks > 0 ? 1u : 0u);
          tc_mma(tmem_dP, dV_ + ks * 2, dOk + ks * 2, id, ks > 0 ? 1u : 0u);
        }
        if (merge && i == 0) {                      // tail block, 16 wide, parked in the (not yet live) dV / dK columns
          const uint64_t dQt = make_sdesc_sw128(smem_u32(sQ + 1 * HALF_BYTES), false, 0);
          const uint64_t dOt = make_sdesc_sw128(smem_u32(sdO + 1 * HALF_BYTES), false, 0);
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            tc_mma(tmem_dV, dK_ + ks * 2, dQt + ks * 2, idesc_s16, ks > 0 ? 1u : 0u);
            tc_mma(tmem_dK, dV_ + ks * 2, dOt + ks * 2, idesc_s16, ks > 0 ? 1u : 0u);
          }
        }
        tc_commit(s_full);
      }
      __syncwarp();
      TRACE(1 * 256 + i * 8 + 2);
    };
    issue_sdp(0);
    for (int i = 0; i < nsteps; ++i) {
      const int e = merge ? (i == 0 ? 0 : i + 1) : i;
      const int st = e % BWD_STAGES;
      TRACE(1 * 256 + i * 8 + 3);
      mbar_wait(p_full, i & 1);                       // P^T / dS^T of block i sit in TMEM (and every warp has read S^T / dP^T)
      TRACE(1 * 256 + i * 8 + 4);
      tc_fence_after();
      if (leader) {
        const uint64_t dQm = make_sdesc_sw128(smem_u32(sQ + st * HALF_BYTES), true, HALF_BYTES);
        const uint64_t dOm = make_sdesc_sw128(smem_u32(sdO + st * HALF_BYTES), true, HALF_BYTES);
        const uint32_t acc = i > 0 ? 1u : 0u;
        // k-step ks covers queries [16 ks, 16 ks + 16): packed by warp group ks / 2 at column 32 (ks / 2) + 8 (ks % 2)
        if (!merge && i == nqb - 1 && tail16) {
          tc_mma_ts(tmem_dV, tmem_S, dOm, idesc_g, acc);
          tc_mma_ts(tmem_dK, tmem_dP, dQm, idesc_g, acc);
        } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            tc_mma_ts(tmem_dV, tmem_S + (ks >> 1) * 32 + (ks & 1) * 8, dOm + ks * 128, idesc_g, ks > 0 ? 1u : acc);
            tc_mma_ts(tmem_dK, tmem_dP + (ks >> 1) * 32 + (ks & 1) * 8, dQm + ks * 128, idesc_g, ks > 0 ? 1u : acc);
          }
        }
        if (merge && i == 0) {      // the tail block's P^T / dS^T were packed into the free columns [16, 24) of each region
          const uint64_t dQt = make_sdesc_sw128(smem_u32(sQ + 1 * HALF_BYTES), true, HALF_BYTES);
          const uint64_t dOt = make_sdesc_sw128(smem_u32(sdO + 1 * HALF_BYTES), true, HALF_BYTES);
          tc_mma_ts(tmem_dV, tmem_S + 16, dOt, idesc_g, 1u);
          tc_mma_ts(tmem_dK, tmem_dP + 16, dQt, idesc_g, 1u);
          tc_commit(&qdo_empty[1]);
        }
        tc_commit(&qdo_empty[st]);
        if (i == nsteps - 1) tc_commit(done);
      }
      __syncwarp();
      TRACE(1 * 256 + i * 8 + 5);
      if (i + 1 < nsteps) issue_sdp(i + 1);
    }
  } else if (warp < 8) {
    // ===================== softmax-backward threads: one KEY row per thread =====================
    const int q = warp & 3;
    const int kvrow = kt * TILE + q * 32 + lane;
    const bool row_ok = kvrow < S;
    const bool warp_active = kt * TILE + q * 32 < S;          // warps without a valid key row only keep the barriers moving
    const int wg = warp >> 2;                                 // which 32-column half of each block this warp owns
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float* lse_g = lse + (static_cast<long long>(b) * H + h) * S;
    const float* delta_g = delta + (static_cast<long long>(b) * H + h) * S;
    // per-query statistics of this warp's 32 columns: lse*log2e | delta, loaded one block ahead (the global-load
    // latency stays off the critical path) into a private, double-buffered shared-memory slot -- no CTA-wide barrier
    float* wstat = sStat + warp * 128;
    auto load_lse = [&](int i) { const int qr = i * 64 + wg * 32 + lane; return qr < S ? lse_g[qr] : 0.f; };
    auto load_delta = [&](int i) { const int qr = i * 64 + wg * 32 + lane; return qr < S ? delta_g[qr] : 0.f; };
    float ls_n = load_lse(0), dl_n = load_delta(0);
    for (int i = 0; i < nsteps; ++i) {
      long long* tr = (warp == 0 && lane == 0) ? trace : nullptr;
      float* stat = wstat + (i & 1) * 64;
      stat[lane] = ls_n * LOG2E;
      stat[32 + lane] = dl_n;
      if (merge && i == 0 && wg == 0) {               // statistics of the tail block's queries: the other (still unused) buffer
        const int qr = (nqb - 1) * 64 + lane;
        wstat[64 + lane] = (lane < 16 && qr < S) ? lse_g[qr] * LOG2E : 0.f;
        wstat[96 + lane] = (lane < 16 && qr < S) ? delta_g[qr] : 0.f;
      }
      __syncwarp();
      if (i + 1 < nsteps) { ls_n = load_lse(i + 1); dl_n = load_delta(i + 1); }
      if (tr) tr[2 * 256 + i * 8 + 0] = clock64();
      mbar_wait(s_full, i & 1);
      if (tr) tr[2 * 256 + i * 8 + 1] = clock64();
      tc_fence_after();
      const int ncol = min(64, S - i * 64);                    // valid query columns in this block
      const bool t16 = !merge && (i == nqb - 1) && tail16;
      if (warp_active && !(t16 && wg == 1)) {
        uint32_t pk[16], dk[16];
        if (t16) {
          uint32_t sv[16], dv[16];
          tmem_ld16(tmem_S + lane_off, sv);
          tmem_ld16(tmem_dP + lane_off, dv);
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float2 ls = *reinterpret_cast<const float2*>(stat + e);
            const float2 dl = *reinterpret_cast<const float2*>(stat + 32 + e);
            const float p0 = e < ncol ? ex2f(fmaf(__uint_as_float(sv[e]), sl2, -ls.x)) : 0.f;
            const float p1 = e + 1 < ncol ? ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, -ls.y)) : 0.f;
            pk[e >> 1] = pack_bf16x2(p0, p1);
            dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - dl.x), p1 * (__uint_as_float(dv[e + 1]) - dl.y));
          }
#pragma unroll
          for (int e = 8; e < 16; ++e) { pk[e] = 0u; dk[e] = 0u; }
        } else {
          uint32_t sv[32], dv[32];
          tmem_ld32_issue(tmem_S + lane_off + wg * 32, sv);
          tmem_ld32_issue(tmem_dP + lane_off + wg * 32, dv);
          tmem_ld_wait();
          const int lim = ncol - wg * 32;                          // valid query columns in this warp's chunk
          if (lim < 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e >= lim) sv[e] = 0xff800000u;                   // exp2(-inf) = 0: P and dS vanish outside the problem
          }
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 ls = *reinterpret_cast<const float4*>(stat + e);
            const float4 dl = *reinterpret_cast<const float4*>(stat + 32 + e);
            const float p0 = ex2f(fmaf(__uint_as_float(sv[e]), sl2, -ls.x));
            const float p1 = ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, -ls.y));
            const float p2 = ex2f(fmaf(__uint_as_float(sv[e + 2]), sl2, -ls.z));
            const float p3 = ex2f(fmaf(__uint_as_float(sv[e + 3]), sl2, -ls.w));
            pk[e >> 1] = pack_bf16x2(p0, p1);
            pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
            dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - dl.x), p1 * (__uint_as_float(dv[e + 1]) - dl.y));
            dk[(e >> 1) + 1] = pack_bf16x2(p2 * (__uint_as_float(dv[e + 2]) - dl.z), p3 * (__uint_as_float(dv[e + 3]) - dl.w));
          }
        }
        if (tr) tr[2 * 256 + i * 8 + 2] = clock64();
        // in place: this thread's own lanes, inside the fp32 columns it has just read
        tmem_st16(tmem_S + lane_off + wg * 32, pk);
        tmem_st16(tmem_dP + lane_off + wg * 32, dk);
        tmem_st_wait();
      }
      if (merge && i == 0 && wg == 0 && warp_active) {
        // tail block: 16 columns out of the dV / dK accumulator columns, packed into columns [16, 24) of S^T / dP^T
        // (after the main block above: those columns held block 0's fp32 scores until this warp had loaded them)
        const int ncol_t = S - (nqb - 1) * 64;
        const float* tstat = wstat + 64;
        uint32_t sv[16], dv[16], pk[16], dk[16];
        tmem_ld16(tmem_dV + lane_off, sv);
        tmem_ld16(tmem_dK + lane_off, dv);
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float2 ls = *reinterpret_cast<const float2*>(tstat + e);
          const float2 dl = *reinterpret_cast<const float2*>(tstat + 32 + e);
          const float p0 = e < ncol_t ? ex2f(fmaf(__uint_as_float(sv[e]), sl2, -ls.x)) : 0.f;
          const float p1 = e + 1 < ncol_t ? ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, -ls.y)) : 0.f;
          pk[e >> 1] = pack_bf16x2(p0, p1);
          dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - dl.x), p1 * (__uint_as_float(dv[e + 1]) - dl.y));
        }
#pragma unroll
        for (int e = 8; e < 16; ++e) { pk[e] = 0u; dk[e] = 0u; }
        tmem_st16(tmem_S + lane_off + 16, pk);
        tmem_st16(tmem_dP + lane_off + 16, dk);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tr) tr[2 * 256 + i * 8 + 3] = clock64();
    }
    // ---- epilogue: dV, dK rows
    mbar_wait(done, 0);
    tc_fence_after();
    bf16* dkrow = dqkv + (static_cast<long long>(b) * S + kvrow) * (3LL * D) + D + h * HD;
    bf16* dvrow = dkrow + D;
    if (warp_active) {                       // both warp groups cover the same 128 rows: one writes dK, the other dV
      {
        const int which = wg;
        const uint32_t src = which == 0 ? tmem_dK : tmem_dV;
        bf16* dst = which == 0 ? dkrow : dvrow;
        const float sc = which == 0 ? scale : 1.0f;
#pragma unroll 1
        for (int c0 = 0; c0 < HD; c0 += 16) {
          uint32_t o[16];
          tmem_ld16(src + lane_off + c0, o);
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[8 * g]) * sc, __uint_as_float(o[8 * g + 1]) * sc);
              u.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * sc, __uint_as_float(o[8 * g + 3]) * sc);
              u.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * sc, __uint_as_float(o[8 * g + 5]) * sc);
              u.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * sc, __uint_as_float(o[8 * g + 7]) * sc);
              *reinterpret_cast<uint4*>(dst + c0 + 8 * g) = u;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, BWD_TMEM_COLS); }
}

// =====================================================================================================
// dK/dV kernel, pipelined variant: 32-query blocks with TWO S^T / dP^T buffer pairs in tensor memory
// (2 x (32 + 32) + dV + dK = 224..256 columns, still two CTAs per SM).  The MMA warp keeps S^T / dP^T of block i+1 in
// flight while the softmax warps work on block i (the 64-wide single-buffered kernel above is a serial
// MMA -> softmax -> MMA chain of ~2500 cycles per block, profiles/r01_attn_dkdv_timeline.txt).
// MEASURED: no faster (dec bwd 1.98 vs 1.90 ms) -- a softmax warp needs ~800 cycles for a 16-column chunk against ~950
// for 32 columns (tcgen05.ld round trip, dependent ex2 chain, tcgen05.st + wait, fence, arrive are mostly fixed cost), so
// with the same eight warps serving both buffers the softmax warps become the serial resource.  Kept as the off-by-default
// A/B variant (hct_attention_set_dkdv32) and as the starting point for a version with one warp group per buffer.  Buffer b = i & 1 is recycled by in-order issue: S^T / dP^T of block i+2 are
// issued behind the dV / dK MMAs of block i that read P^T / dS^T from the same columns.
// =====================================================================================================
constexpr int QB = 32;                        // queries per block
constexpr int QB_BYTES = QB * 128;            // [32 rows][64 bf16]
constexpr int DK32_STAGES = 6;
constexpr int DK32_SMEM = 2 * TILE_BYTES + DK32_STAGES * 2 * QB_BYTES + 8 * 256 + 1024 + 256;

template <int HD>
__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_bwd_dkdv32_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV32,
                          const __grid_constant__ CUtensorMap tmDO32, const float* __restrict__ lse,
                          const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sK = smem;                               // [128 keys][64]
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;              // DK32_STAGES x [32 queries][64]
  uint8_t* sdO = sQ + DK32_STAGES * QB_BYTES;
  float* sStat = reinterpret_cast<float*>(sdO + DK32_STAGES * QB_BYTES);    // [8 warps][2 buffers][lse 16 | delta 16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + 8 * 256);
  uint64_t *kv_full = bars, *qdo_full = bars + 1 /*[6]*/, *qdo_empty = bars + 7 /*[6]*/, *s_full = bars + 13 /*[2]*/,
           *p_full = bars + 15 /*[2]*/, *done = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
  const int D = H * HD;
  const int nqb = (S + QB - 1) / QB;
  const bool tail16 = S - (nqb - 1) * QB <= 16;             // last query block is computed 16 columns wide
  const float sl2 = scale * LOG2E;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < DK32_STAGES; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, BWD_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // S^T buffers: [0,32), [32,64)   dP^T buffers: [64,96), [96,128)   dV: [128, 128+HD)   dK: [192, 192+HD)
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64, tmem_dV = tmem_base + 128, tmem_dK = tmem_base + 192;

  if (warp == BWD_PRODUCER_WARP) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_2d(smem_u32(sK), &tmQKV128, kv_full, D + h * HD, b * S + kt * TILE);
      tma_load_2d(smem_u32(sV), &tmQKV128, kv_full, 2 * D + h * HD, b * S + kt * TILE);
    }
    for (int i = 0; i < nqb; ++i) {
      const int st = i % DK32_STAGES;
      mbar_wait(&qdo_empty[st], ((i / DK32_STAGES) & 1) ^ 1u);
      if (leader) {
        mbar_expect_tx(&qdo_full[st], 2 * QB_BYTES);
        tma_load_2d(smem_u32(sQ + st * QB_BYTES), &tmQKV32, &qdo_full[st], h * HD, b * S + i * QB);
        tma_load_2d(smem_u32(sdO + st * QB_BYTES), &tmDO32, &qdo_full[st], h * HD, b * S + i * QB);
      }
      __syncwarp();
    }
  } else if (warp == BWD_MMA_WARP) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(TILE, QB, false, false);
    const uint32_t idesc_s16 = make_idesc_bf16(TILE, 16, false, false);
    const uint32_t idesc_g = make_idesc_bf16(TILE, HD, false, true);
    const uint64_t dK_ = make_sdesc_sw128(smem_u32(sK), false, 0);
    const uint64_t dV_ = make_sdesc_sw128(smem_u32(sV), false, 0);
    mbar_wait(kv_full, 0);
    auto issue_sdp = [&](int i) {          // S^T = K Q_i^T and dP^T = V dO_i^T into buffer i & 1
      const int st = i % DK32_STAGES, bf = i & 1;
      mbar_wait(&qdo_full[st], (i / DK32_STAGES) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t id = (i == nqb - 1 && tail16) ? idesc_s16 : idesc_s;
        const uint64_t dQk = make_sdesc_sw128(smem_u32(sQ + st * QB_BYTES), false, 0);
        const uint64_t dOk = make_sdesc_sw128(smem_u32(sdO + st * QB_BYTES), false, 0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {      // the two accumulate chains interleaved: consecutive MMAs independent
          tc_mma(tmem_S + bf * QB, dK_ + ks * 2, dQk + ks * 2, id, ks > 0 ? 1u : 0u);
          tc_mma(tmem_dP + bf * QB, dV_ + ks * 2, dOk + ks * 2, id, ks > 0 ? 1u : 0u);
        }
        tc_commit(&s_full[bf]);
      }
      __syncwarp();
    };
    issue_sdp(0);
    if (nqb > 1) issue_sdp(1);
    for (int i = 0; i < nqb; ++i) {
      const int st = i % DK32_STAGES, bf = i & 1;
      mbar_wait(&p_full[bf], (i >> 1) & 1);           // P^T / dS^T of block i sit in TMEM (and every warp has read S^T / dP^T)
      tc_fence_after();
      if (leader) {
        const uint64_t dQm = make_sdesc_sw128(smem_u32(sQ + st * QB_BYTES), true, QB_BYTES);
        const uint64_t dOm = make_sdesc_sw128(smem_u32(sdO + st * QB_BYTES), true, QB_BYTES);
        const uint32_t acc = i > 0 ? 1u : 0u;
        // k-step ks covers queries [16 ks, 16 ks + 16): packed by column group ks into the first 8 of its 16 columns
        if (i == nqb - 1 && tail16) {
          tc_mma_ts(tmem_dV, tmem_S + bf * QB, dOm, idesc_g, acc);
          tc_mma_ts(tmem_dK, tmem_dP + bf * QB, dQm, idesc_g, acc);
        } else {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            tc_mma_ts(tmem_dV, tmem_S + bf * QB + ks * 16, dOm + ks * 128, idesc_g, ks > 0 ? 1u : acc);
            tc_mma_ts(tmem_dK, tmem_dP + bf * QB + ks * 16, dQm + ks * 128, idesc_g, ks > 0 ? 1u : acc);
          }
        }
        tc_commit(&qdo_empty[st]);
        if (i == nqb - 1) tc_commit(done);
      }
      __syncwarp();
      if (i + 2 < nqb) issue_sdp(i + 2);              // into the buffer block i has just released (in-order tensor pipe)
    }
  } else if (warp < 8) {
    // ===================== softmax-backward threads: one KEY row per thread, 16 query columns per warp =====================
    const int q = warp & 3;
    const int kvrow = kt * TILE + q * 32 + lane;
    const bool row_ok = kvrow < S;
    const bool warp_active = kt * TILE + q * 32 < S;          // warps without a valid key row only keep the barriers moving
    const int wg = warp >> 2;                                 // which 16-column half of each block this warp owns
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float* lse_g = lse + (static_cast<long long>(b) * H + h) * S;
    const float* delta_g = delta + (static_cast<long long>(b) * H + h) * S;
    // per-query statistics of this warp's 16 columns: lanes 0-15 fetch lse (scaled by log2 e), lanes 16-31 delta, one
    // block ahead, into a private double-buffered shared-memory slot
    float* wstat = sStat + warp * 64;
    auto load_stat = [&](int blk) {
      const int qr = blk * QB + wg * 16 + (lane & 15);
      if (qr >= S) return 0.f;
      return lane < 16 ? lse_g[qr] * LOG2E : delta_g[qr];
    };
    float st_n = load_stat(0);
    for (int i = 0; i < nqb; ++i) {
      const int bf = i & 1;
      float* stat = wstat + bf * 32;                           // [lse 16 | delta 16]
      stat[lane] = st_n;
      __syncwarp();
      if (i + 1 < nqb) st_n = load_stat(i + 1);
      mbar_wait(&s_full[bf], (i >> 1) & 1);
      tc_fence_after();
      const int ncol = min(QB, S - i * QB);                    // valid query columns in this block
      const bool t16 = (i == nqb - 1) && tail16;               // 16-wide tail block: column group 0 only
      if (warp_active && !(t16 && wg != 0)) {
        uint32_t sv[16], dv[16], pk[8], dk[8];
        tmem_ld16_issue(tmem_S + bf * QB + lane_off + wg * 16, sv);
        tmem_ld16_issue(tmem_dP + bf * QB + lane_off + wg * 16, dv);
        tmem_ld_wait();
        const int lim = ncol - wg * 16;                        // valid query columns in this warp's chunk
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 ls = *reinterpret_cast<const float4*>(stat + e);
          const float4 dl = *reinterpret_cast<const float4*>(stat + 16 + e);
          // P and dS vanish outside the problem
          const float p0 = e < lim ? ex2f(fmaf(__uint_as_float(sv[e]), sl2, -ls.x)) : 0.f;
          const float p1 = e + 1 < lim ? ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, -ls.y)) : 0.f;
          const float p2 = e + 2 < lim ? ex2f(fmaf(__uint_as_float(sv[e + 2]), sl2, -ls.z)) : 0.f;
          const float p3 = e + 3 < lim ? ex2f(fmaf(__uint_as_float(sv[e + 3]), sl2, -ls.w)) : 0.f;
          pk[e >> 1] = pack_bf16x2(p0, p1);
          pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
          dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - dl.x), p1 * (__uint_as_float(dv[e + 1]) - dl.y));
          dk[(e >> 1) + 1] = pack_bf16x2(p2 * (__uint_as_float(dv[e + 2]) - dl.z), p3 * (__uint_as_float(dv[e + 3]) - dl.w));
        }
        // in place: this thread's own lanes, the first 8 of the 16 fp32 columns it has just read
        tmem_st8(tmem_S + bf * QB + lane_off + wg * 16, pk);
        tmem_st8(tmem_dP + bf * QB + lane_off + wg * 16, dk);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[bf]);
    }
    // ---- epilogue: dV, dK rows
    mbar_wait(done, 0);
    tc_fence_after();
    bf16* dkrow = dqkv + (static_cast<long long>(b) * S + kvrow) * (3LL * D) + D + h * HD;
    bf16* dvrow = dkrow + D;
    if (warp_active) {                       // both warp groups cover the same 128 rows: one writes dK, the other dV
      const uint32_t src = wg == 0 ? tmem_dK : tmem_dV;
      bf16* dst = wg == 0 ? dkrow : dvrow;
      const float sc = wg == 0 ? scale : 1.0f;
#pragma unroll 1
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t o[16];
        tmem_ld16(src + lane_off + c0, o);
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * g]) * sc, __uint_as_float(o[8 * g + 1]) * sc);
            u.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * sc, __uint_as_float(o[8 * g + 3]) * sc);
            u.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * sc, __uint_as_float(o[8 * g + 5]) * sc);
            u.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * sc, __uint_as_float(o[8 * g + 7]) * sc);
            *reinterpret_cast<uint4*>(dst + c0 + 8 * g) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, BWD_TMEM_COLS); }
}

template <int HD>
__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                      const __grid_constant__ CUtensorMap tmDO128, const float* __restrict__ lse,
                      const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H, float scale, int merge_tail) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                               // [128 queries][64]
  uint8_t* sdO = smem + TILE_BYTES;
  uint8_t* sK = smem + 2 * TILE_BYTES;              // BWD_STAGES x [64 keys][64]
  uint8_t* sV = sK + BWD_STAGES * HALF_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + BWD_STAGES * HALF_BYTES + 8 * 512);
  uint64_t *qdo_full = bars, *kv_full = bars + 1 /*[4]*/, *kv_empty = bars + 5 /*[4]*/, *s_full = bars + 9,
           *s_empty = bars + 10, *p_full = bars + 11, *ds_free = bars + 12 /*[2]*/, *done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
  const int D = H * HD;
  const int nkb = (S + 63) / 64;
  const bool tail16 = S - (nkb - 1) * 64 <= 16;             // last key block is computed 16 columns wide
  // merged tail (see the dK/dV kernel): the 16-wide last key block rides along with block 0 -- its S / dP park in the dQ
  // accumulator columns [128,144) / [144,160), its dS goes to operand buffer 1 -- one chain step less.
  // Ring entries: 0 = block 0, 1 = tail block, e >= 2 = block e - 1.
  const bool merge = merge_tail != 0 && tail16 && nkb >= 2;
  const int nsteps = merge ? nkb - 1 : nkb;
  const float sl2 = scale * LOG2E;

  if (threadIdx.x == 0) {
    mbar_init(qdo_full, 1);
    for (int i = 0; i < BWD_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, 8); mbar_init(p_full, 8);
    mbar_init(&ds_free[0], 1); mbar_init(&ds_free[1], 1); mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, BWD_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // S: [0,64)  dP: [64,128)  dQ: [128,192)  dS operand buffers (bf16 pairs, 32 columns each): [192,224), [224,256)
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64, tmem_dQ = tmem_base + 128, tmem_dS = tmem_base + 192;

  if (warp == BWD_PRODUCER_WARP) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(qdo_full, 2 * TILE_BYTES);
      tma_load_2d(smem_u32(sQ), &tmQKV128, qdo_full, h * HD, b * S + qt * TILE);
      tma_load_2d(smem_u32(sdO), &tmDO128, qdo_full, h * HD, b * S + qt * TILE);
    }
    for (int j = 0; j < nkb; ++j) {                   // j = ring entry
      const int st = j % BWD_STAGES;
      const int blk = merge ? (j == 0 ? 0 : (j == 1 ? nkb - 1 : j - 1)) : j;
      mbar_wait(&kv_empty[st], ((j / BWD_STAGES) & 1) ^ 1u);
      if (leader) {
        mbar_expect_tx(&kv_full[st], 2 * HALF_BYTES);
        tma_load_2d(smem_u32(sK + st * HALF_BYTES), &tmQKV64, &kv_full[st], D + h * HD, b * S + blk * 64);
        tma_load_2d(smem_u32(sV + st * HALF_BYTES), &tmQKV64, &kv_full[st], 2 * D + h * HD, b * S + blk * 64);
      }
      __syncwarp();
    }
  } else if (warp == BWD_MMA_WARP) {
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(TILE, 64, false, false);
    const uint32_t idesc_s16 = make_idesc_bf16(TILE, 16, false, false);
    const uint32_t idesc_g = make_idesc_bf16(TILE, HD, false, true);
    const uint64_t dQ_ = make_sdesc_sw128(smem_u32(sQ), false, 0);
    const uint64_t dO_ = make_sdesc_sw128(smem_u32(sdO), false, 0);
    mbar_wait(qdo_full, 0);
    auto issue_sdp = [&](int j) {                   // step j
      const int e = merge ? (j == 0 ? 0 : j + 1) : j;
      const int st = e % BWD_STAGES;
      mbar_wait(&kv_full[st], (e / BWD_STAGES) & 1);
      if (merge && j == 0) mbar_wait(&kv_full[1], 0);
      if (j > 0) mbar_wait(s_empty, (j - 1) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t id = (!merge && j == nkb - 1 && tail16) ? idesc_s16 : idesc_s;
        const uint64_t dKk = make_sdesc_sw128(smem_u32(sK + st * HALF_BYTES), false, 0);
        const uint64_t dVk = make_sdesc_sw128(smem_u32(sV + st * HALF_BYTES), false, 0);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          tc_mma(tmem_S, dQ_ + ks * 2, dKk + ks * 2, id, ks > 0 ? 1u : 0u);
          tc_mma(tmem_dP, dO_ + ks * 2, dVk + ks * 2, id, ks > 0 ? 1u : 0u);
        }
        if (merge && j == 0) {                      // tail block: S_t -> dQ columns [0,16), dP_t -> dQ columns [16,32)
          const uint64_t dKt = make_sdesc_sw128(smem_u32(sK + 1 * HALF_BYTES), false, 0);
          const uint64_t dVt = make_sdesc_sw128(smem_u32(sV + 1 * HALF_BYTES), false, 0);
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            tc_mma(tmem_dQ, dQ_ + ks * 2, dKt + ks * 2, idesc_s16, ks > 0 ? 1u : 0u);
            tc_mma(tmem_dQ + 16, dO_ + ks * 2, dVt + ks * 2, idesc_s16, ks > 0 ? 1u : 0u);
          }
        }
        tc_commit(s_full);
      }
      __syncwarp();
    };
    issue_sdp(0);
    for (int j = 0; j < nsteps; ++j) {
      const int e = merge ? (j == 0 ? 0 : j + 1) : j;
      const int st = e % BWD_STAGES;
      if (j + 1 < nsteps) issue_sdp(j + 1);         // runs underneath the softmax threads' work on block j
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t dKm = make_sdesc_sw128(smem_u32(sK + st * HALF_BYTES), true, HALF_BYTES);
        const uint32_t a = tmem_dS + (j & 1) * 32;
        const uint32_t acc = j > 0 ? 1u : 0u;
        if (!merge && j == nkb - 1 && tail16) {
          tc_mma_ts(tmem_dQ, a, dKm, idesc_g, acc);
        } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc_mma_ts(tmem_dQ, a + ks * 8, dKm + ks * 128, idesc_g, ks > 0 ? 1u : acc);
        }
        if (merge && j == 0) {                      // the tail block's dS sits in operand buffer 1
          const uint64_t dKt = make_sdesc_sw128(smem_u32(sK + 1 * HALF_BYTES), true, HALF_BYTES);
          tc_mma_ts(tmem_dQ, tmem_dS + 32, dKt, idesc_g, 1u);
          tc_commit(&ds_free[1]);
          tc_commit(&kv_empty[1]);
        }
        tc_commit(&ds_free[j & 1]);
        tc_commit(&kv_empty[st]);
        if (j == nsteps - 1) tc_commit(done);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    const int q = warp & 3;
    const int row = qt * TILE + q * 32 + lane;
    const bool row_ok = row < S;
    const bool warp_active = qt * TILE + q * 32 < S;          // warps without a valid query row only keep the barriers moving
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int wg = warp >> 2;
    const long long sidx = (static_cast<long long>(b) * H + h) * S + row;
    const float lse_r = row_ok ? lse[sidx] * LOG2E : 0.f;
    const float delta_r = row_ok ? delta[sidx] : 0.f;
    for (int j = 0; j < nsteps; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int ncol = min(64, S - j * 64);                    // valid key columns in this block
      const bool t16 = !merge && (j == nkb - 1) && tail16;
      uint32_t dkt[16];
      if (merge && j == 0 && wg == 0 && warp_active) {         // tail block out of the dQ columns (not live before p_full)
        const int ncol_t = S - (nkb - 1) * 64;
        uint32_t svt[16], dvt[16];
        tmem_ld16(tmem_dQ + lane_off, svt);
        tmem_ld16(tmem_dQ + 16 + lane_off, dvt);
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float p0 = e < ncol_t ? ex2f(fmaf(__uint_as_float(svt[e]), sl2, -lse_r)) : 0.f;
          const float p1 = e + 1 < ncol_t ? ex2f(fmaf(__uint_as_float(svt[e + 1]), sl2, -lse_r)) : 0.f;
          dkt[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dvt[e]) - delta_r), p1 * (__uint_as_float(dvt[e + 1]) - delta_r));
        }
#pragma unroll
        for (int e = 8; e < 16; ++e) dkt[e] = 0u;
      }
      const bool works = warp_active && !(t16 && wg == 1);
      uint32_t dk[16];
      if (!works) {
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
      } else if (t16) {
        uint32_t sv[16], dv[16];
        tmem_ld16(tmem_S + lane_off, sv);
        tmem_ld16(tmem_dP + lane_off, dv);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float p0 = e < ncol ? ex2f(fmaf(__uint_as_float(sv[e]), sl2, -lse_r)) : 0.f;
          const float p1 = e + 1 < ncol ? ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, -lse_r)) : 0.f;
          dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - delta_r), p1 * (__uint_as_float(dv[e + 1]) - delta_r));
        }
#pragma unroll
        for (int e = 8; e < 16; ++e) dk[e] = 0u;
      } else {
        uint32_t sv[32], dv[32];
        tmem_ld32_issue(tmem_S + lane_off + wg * 32, sv);
        tmem_ld32_issue(tmem_dP + lane_off + wg * 32, dv);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);                   // S / dP are in registers: block j+1 may overwrite them
        const int lim = ncol - wg * 32;
        if (lim < 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e >= lim) sv[e] = 0xff800000u;
        }
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ex2f(fmaf(__uint_as_float(sv[e]), sl2, -lse_r));
          const float p1 = ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, -lse_r));
          dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - delta_r), p1 * (__uint_as_float(dv[e + 1]) - delta_r));
        }
      }
      // operand buffer j & 1 is free once the dQ MMAs of block j - 2 have retired (merged tail: buffer 1 is also used,
      // and released, at step 0, so the odd steps wait one completion earlier)
      if (merge && (j & 1)) { mbar_wait(&ds_free[1], (j >> 1) & 1); tc_fence_after(); }
      else if (j >= 2) { mbar_wait(&ds_free[j & 1], ((j >> 1) - 1) & 1); tc_fence_after(); }
      if (works) {
        tmem_st16(tmem_dS + (j & 1) * 32 + lane_off + wg * 16, dk);
        if (merge && j == 0 && wg == 0) tmem_st16(tmem_dS + 32 + lane_off, dkt);
        tmem_st_wait();
      }
      // a warp with nothing to compute could otherwise run two blocks ahead and arrive twice in one phase
      if (j > 0) mbar_wait(p_full, (j - 1) & 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    bf16* dqrow = dqkv + (static_cast<long long>(b) * S + row) * (3LL * D) + h * HD;
    if (warp_active) {                       // the two warp groups cover the same rows: alternate 16-column chunks
#pragma unroll 1
      for (int c0 = wg * 16; c0 < HD; c0 += 32) {
        uint32_t o[16];
        tmem_ld16(tmem_dQ + lane_off + c0, o);
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * g]) * scale, __uint_as_float(o[8 * g + 1]) * scale);
            u.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * scale, __uint_as_float(o[8 * g + 3]) * scale);
            u.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * scale, __uint_as_float(o[8 * g + 5]) * scale);
            u.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * scale, __uint_as_float(o[8 * g + 7]) * scale);
            *reinterpret_cast<uint4*>(dqrow + c0 + 8 * g) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, BWD_TMEM_COLS); }
}

template <typename K>
int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { hct_set_error("cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
  return HCT_OK;
}

}  // namespace

static int g_dkdv32 = 0;         // 1 = pipelined dK/dV kernel (32-query blocks, two buffer pairs); 0 (default) = 64-wide single-buffered
extern "C" int hct_attention_set_dkdv32(int enable) { g_dkdv32 = enable != 0; return HCT_OK; }
static int g_merge_tail = 1;     // 0 = the 16-wide tail block as its own chain step (A/B comparison)
extern "C" int hct_attention_set_merge_tail(int enable) { g_merge_tail = enable != 0; return HCT_OK; }

extern "C" int hct_attention_trace(void* buf) {      // device buffer of >= 768 int64 (or null): dK/dV kernel event timeline
  long long* p = static_cast<long long*>(buf);
  return cudaMemcpyToSymbol(g_trace, &p, sizeof(p)) == cudaSuccess ? HCT_OK : HCT_ERR_CUDA;
}

// Number of 128-row query tiles the tcgen05 kernels cover; the remaining (few) rows go to the mma.sync tail kernel.
int hct_attn_tc_tiles(int S, int tail_on_mma_sync) {
  const int rem = S % TILE;
  if (tail_on_mma_sync && S > TILE && rem != 0 && rem <= 32) return S / TILE;
  return (S + TILE - 1) / TILE;
}

static int g_fold_tail_key = 1;  // 1: a single key behind the last 64-key block is folded into the forward's epilogue
static int g_fold_tail_row = 1;  // 1: a single query row behind the last full tile rides with that tile's CTA (producer warp)
static int g_fwd_flags = 0;
extern "C" int hct_attention_set_tail_key(int fold) { g_fold_tail_key = (fold & 1) != 0; g_fold_tail_row = (fold & 2) != 0; return HCT_OK; }
static int g_fwd_poly = -1;      // forward softmax arithmetic: -1 scalar fp32, one pass (default: measured fastest, 0.516 vs 0.577 ms); 0 / 3: packed fp32, two passes, 0 or 3 of 8 exponential pairs on the FMA pipe
int hct_attention_bwd3_set_poly(int n);
extern "C" int hct_attention_set_poly(int fwd, int bwd) {     // -1: scalar fp32 arithmetic; -2: leave unchanged
  if (fwd >= -1) g_fwd_poly = fwd;
  if (bwd >= -1) hct_attention_bwd3_set_poly(bwd);
  return HCT_OK;
}
template <int POLY>
static int launch_fwd_tc(const CUtensorMap& tmq, const CUtensorMap& tmkv, const void* qkv, void* out, float* lse, int S, int H, int hd, float scale,
                         dim3 grid, cudaStream_t st) {
  static bool cfg = false;
  if (!cfg) {
    int rc = set_smem(attn_fwd_tc_kernel<64, POLY>, FWD_SMEM); if (rc) return rc;
    rc = set_smem(attn_fwd_tc_kernel<48, POLY>, FWD_SMEM); if (rc) return rc;
    cfg = true;
  }
  const bf16* q = static_cast<const bf16*>(qkv);
  if (hd == 64) hct_launch_pdl(attn_fwd_tc_kernel<64, POLY>, grid, dim3(FWD_THREADS), FWD_SMEM, st, tmq, tmkv, q, static_cast<bf16*>(out), lse, S, H, scale, g_fwd_flags);
  else hct_launch_pdl(attn_fwd_tc_kernel<48, POLY>, grid, dim3(FWD_THREADS), FWD_SMEM, st, tmq, tmkv, q, static_cast<bf16*>(out), lse, S, H, scale, g_fwd_flags);
  return HCT_OK;
}

int hct_attention_fwd_tc(const void* qkv, void* out, float* lse, int B, int S, int H, int hd, int n_tiles,
                         cudaStream_t st) {
  CUtensorMap tmq, tmkv;
  const long long D3 = 3LL * H * hd;
  int rc = hct_make_tmap_bf16_2d(&tmq, qkv, D3, static_cast<long long>(B) * S, D3, 64, TILE);
  if (rc != HCT_OK) return rc;
  rc = hct_make_tmap_bf16_2d(&tmkv, qkv, D3, static_cast<long long>(B) * S, D3, 64, FWD_TK);
  if (rc != HCT_OK) return rc;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  // asked to cover every row of an S = 128 k + 1 problem: full tiles only, the last row rides with the last full tile's CTA
  const bool fold_row = g_fold_tail_row && S > TILE && S % TILE == 1 && n_tiles * TILE >= S;
  dim3 grid(fold_row ? S / TILE : n_tiles, H, B);
  g_fwd_flags = (g_fold_tail_key ? 1 : 0) | (fold_row ? 2 : 0);
  rc = g_fwd_poly == 0   ? launch_fwd_tc<0>(tmq, tmkv, qkv, out, lse, S, H, hd, scale, grid, st)
       : g_fwd_poly == 3 ? launch_fwd_tc<3>(tmq, tmkv, qkv, out, lse, S, H, hd, scale, grid, st)
                         : launch_fwd_tc<-1>(tmq, tmkv, qkv, out, lse, S, H, hd, scale, grid, st);
  if (rc) return rc;
  return hct_check_launch("attn_fwd_tc_kernel");
}

template <int HD>
static int launch_bwd_tc(const CUtensorMap& q128, const CUtensorMap& q64, const CUtensorMap& do128, const CUtensorMap& do64,
                         const CUtensorMap& q32, const CUtensorMap& do32,
                         const float* lse, const float* delta, bf16* dqkv, int B, int S, int H, int n_tiles, cudaStream_t st) {
  static bool cfg = false;
  if (!cfg) {
    int rc = set_smem(attn_bwd_dkdv_tc_kernel<HD>, BWD_SMEM); if (rc) return rc;
    rc = set_smem(attn_bwd_dkdv32_tc_kernel<HD>, DK32_SMEM); if (rc) return rc;
    rc = set_smem(attn_bwd_dq_tc_kernel<HD>, BWD_SMEM); if (rc) return rc;
    cfg = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  dim3 grid(n_tiles, H, B);          // 128-row tiles of keys (dK/dV) resp. queries (dQ); rows behind them: hct_attention_tail.cu
  if (g_dkdv32)
    attn_bwd_dkdv32_tc_kernel<HD><<<grid, BWD_THREADS, DK32_SMEM, st>>>(q128, q32, do32, lse, delta, dqkv, S, H, scale);
  else
    attn_bwd_dkdv_tc_kernel<HD><<<grid, BWD_THREADS, BWD_SMEM, st>>>(q128, q64, do64, lse, delta, dqkv, S, H, scale, g_merge_tail);
  int rc = hct_check_launch("attn_bwd_dkdv_tc_kernel");
  if (rc) return rc;
  attn_bwd_dq_tc_kernel<HD><<<grid, BWD_THREADS, BWD_SMEM, st>>>(q128, q64, do128, lse, delta, dqkv, S, H, scale, g_merge_tail);
  return hct_check_launch("attn_bwd_dq_tc_kernel");
}

int hct_attention_bwd_tc(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int B, int S,
                         int H, int hd, int n_tiles, cudaStream_t st) {
  CUtensorMap q128, q64, do128, do64, q32, do32;
  const long long D = static_cast<long long>(H) * hd, D3 = 3 * D, rows = static_cast<long long>(B) * S;
  int rc = hct_make_tmap_bf16_2d(&q128, qkv, D3, rows, D3, 64, TILE); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&q64, qkv, D3, rows, D3, 64, 64); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&do128, dout, D, rows, D, 64, TILE); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&do64, dout, D, rows, D, 64, 64); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&q32, qkv, D3, rows, D3, 64, QB); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&do32, dout, D, rows, D, 64, QB); if (rc) return rc;
  if (hd == 64) return launch_bwd_tc<64>(q128, q64, do128, do64, q32, do32, lse, delta, static_cast<bf16*>(dqkv), B, S, H, n_tiles, st);
  return launch_bwd_tc<48>(q128, q64, do128, do64, q32, do32, lse, delta, static_cast<bf16*>(dqkv), B, S, H, n_tiles, st);
}
