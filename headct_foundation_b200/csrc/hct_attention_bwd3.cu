// Attention backward on tcgen05 / TMEM, pipelined variant ("bwd3"): one persistent CTA per SM that keeps THREE score
// buffer pairs in tensor memory and TWO softmax warp groups busy.
//
// Why: the two-CTA-per-SM kernels of hct_attention_sm100.cu run one serial chain per CTA
//     S/dP MMAs -> commit -> tcgen05.ld -> exp2 / dS -> tcgen05.st -> arrive -> dV/dK (dQ) MMAs -> next S/dP MMAs
// of ~2500 cycles per 64-wide block (profiles/r01_attn_dkdv_timeline.txt) with the tensor pipe busy ~480 of them; the SM
// therefore completes one block per ~1250 cycles.  Throughput tracks the score columns in flight and who waits for whom.
// Here a CTA owns all 512 TMEM columns: accumulators [0,128) + 3 x (S 64 | dP 64).  The MMA warp runs up to three blocks
// ahead with the S / dP MMAs (the tensor pipe executes in order, so a buffer is overwritten only after the accumulate MMAs
// that read its bf16 operands), two groups of eight softmax warps take alternate blocks, and neither waits for the MMA
// stages of its own block any more.  The CTA is persistent over (batch, head, 128-row tile) work items -- resident tiles
// double-buffered, streamed tiles in a 5-deep TMA ring that runs across item boundaries -- so that prologue / epilogue of an
// item (TMA round trip, accumulator drain) overlap the neighbouring items' work instead of being paid per CTA.
//
// One kernel template serves both passes (same chain, transposed roles):
//   KV = true   rows (TMEM lanes) = keys:   resident K, V;  streamed Q_i, dO_i (64 queries)
//               S^T = K Q_i^T, dP^T = V dO_i^T;  P^T = exp2(S^T c - lse[q]), dS^T = P^T (dP^T - delta[q]) written IN PLACE
//               as bf16 pairs;  dV += P^T dO_i,  dK += dS^T Q_i      (A from TMEM, B = streamed tile read MN-major)
//   KV = false  rows = queries:             resident Q, dO; streamed K_j, V_j (64 keys)
//               S = Q K_j^T, dP = dO V_j^T;  dS = P (dP - delta[row]) in place;  dQ += dS K_j
// Deterministic (no atomics).  Rows behind the last full 128-row tile are handled as in the two-kernel path
// (hct_attention_tail.cu), a last streamed block with <= 16 valid columns is computed 16 wide.
#include "../../include/hct_b200.h"
#include "hct_tcgen05.cuh"

namespace {
using namespace hct_tc;

constexpr int TILE = 128;
constexpr int TILE_BYTES = TILE * 128;     // 128 rows x 64 bf16
constexpr int HALF_BYTES = 64 * 128;       // 64 rows x 64 bf16
constexpr float LOG2E = 1.4426950408889634f;
constexpr int NSW = 16;                    // softmax warps: group = w >> 3, column half = (w >> 2) & 1, lane quarter = w & 3
constexpr int W_PROD = 16, W_MMA_S = 17, W_MMA_A = 18;     // TMA producer, S / dP issuer, accumulate issuer
constexpr int THREADS = 19 * 32;
constexpr int STAGES = 5;                  // TMA ring of streamed 64-row tile pairs
constexpr int NBUF = 3;                    // S / dP buffer pairs in tensor memory
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 128;              // [0,64): dV (dQ), [64,128): dK
constexpr int DRAIN_SLOTS = NSW, DRAIN_SLOT_BYTES = 32 * 128;   // per softmax warp: 32 rows x 64 bf16, 16-byte chunks XOR-swizzled
constexpr int SMEM_BYTES = 2 * 2 * TILE_BYTES + STAGES * 2 * HALF_BYTES + NSW * 512 + DRAIN_SLOTS * DRAIN_SLOT_BYTES + 1024 + 512;

static bool g_trace3_host = false;         // host mirror of g_trace3 != nullptr
__device__ long long* g_trace3 = nullptr;   // event timeline of CTA 0 (tools/attn_dbg3.py); nullptr in production
// layout: [role: 0 mma-issue, 1 mma-acc, 2 softmax group 0 (warp 0), 3 softmax group 1 (warp 8)][block gb < 64][16 events]
#define TR3(role, gb, ev) do { if (TRACE && trace && (gb) < 64) trace[((role) * 64 + (gb)) * 16 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (batch, head, row tile) of the CTA's items, advanced WITHOUT divisions: an integer division by a run-time divisor is a
// ~150-cycle dependent chain, and the first version of this kernel spent ~1400 cycles per block on them in every role
// (profiles/r02_attn_bwd3_timeline_v1.txt).
struct ItemIter {
  int b, h, t;              // current item
  int dt, dh, db;           // decomposition of the item stride gridDim.x
  int n_tiles, H;
  __device__ ItemIter(int first, int stride, int n_tiles_, int H_) : n_tiles(n_tiles_), H(H_) {
    t = first % n_tiles; h = (first / n_tiles) % H; b = first / (n_tiles * H);
    dt = stride % n_tiles; dh = (stride / n_tiles) % H; db = stride / (n_tiles * H);
  }
  __device__ __forceinline__ void next() {
    t += dt;
    int c = 0;
    if (t >= n_tiles) { t -= n_tiles; c = 1; }
    h += dh + c;
    c = 0;
    if (h >= H) { h -= H; c = 1; }
    b += db + c;
  }
};

// POLY of every 8 exponential pairs run on the FMA pipe (hct_tcgen05.cuh).  TRACE: the instrumented instantiation behind
// hct_attention_trace3 -- compiled apart because the live trace pointer and its predicates cost the production kernel
// registers (spilled: a local-memory load in front of every stamp).
template <int HD, bool KV, int POLY, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmDO128,
                 const __grid_constant__ CUtensorMap tmQKV64, const __grid_constant__ CUtensorMap tmDO64,
                 const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv,
                 float* __restrict__ colsum, int S, int H, int n_tiles, int n_items, float scale, int tma_drain_ok) {
  // accumulator sets in tensor memory: the dQ pass (64 columns per set) has room for two, so an item's drain overlaps the
  // next item's MMAs; the dK/dV pass (2 x HD columns) has one and pays a short bubble per item
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sR = smem;                                   // [2 buffers][R1 | R2][128 rows][64]
  uint8_t* sT1 = smem + 2 * 2 * TILE_BYTES;             // STAGES x [64 rows][64]
  uint8_t* sT2 = sT1 + STAGES * HALF_BYTES;
  float* sStat = reinterpret_cast<float*>(sT2 + STAGES * HALF_BYTES);       // [16 warps][2 slots][lse 32 | delta 32]
  uint8_t* sDrain = reinterpret_cast<uint8_t*>(sStat) + NSW * 512;          // [16 warps][32 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDrain + DRAIN_SLOTS * DRAIN_SLOT_BYTES);
  uint64_t *kv_full = bars /*[2]*/, *kv_empty = bars + 2 /*[2]*/, *ring_full = bars + 4 /*[STAGES]*/,
           *ring_empty = bars + 4 + STAGES /*[STAGES]*/, *s_full = bars + 4 + 2 * STAGES /*[NBUF]*/,
           *p_full = s_full + NBUF /*[NBUF]*/, *buf_free = p_full + NBUF /*[NBUF]*/, *done = buf_free + NBUF /*[2]*/,
           *acc_empty = done + 2 /*[2]*/;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
  const int D = H * HD;
  const int nblk = (S + 63) / 64;                               // streamed 64-row blocks per item
  const bool tail16 = S - (nblk - 1) * 64 <= 16;                // last block is computed 16 columns wide
  const float sl2 = scale * LOG2E;
  const int G = static_cast<int>(gridDim.x);
  const int nmy = (n_items - static_cast<int>(blockIdx.x) + G - 1) / G;      // items of this CTA: blockIdx.x + n * G
  const int total = nmy * nblk;                                 // its blocks, numbered gb = n * nblk + i
  const bool tma_drain = tma_drain_ok != 0 && n_tiles * TILE <= S;                 // every tile full: results leave through TMA stores

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&done[i], 1); mbar_init(&acc_empty[i], 8); }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
    for (int i = 0; i < NBUF; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); mbar_init(&buf_free[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                  // launched with programmatic stream serialization: nothing above touches global memory
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  // KV: dV at [0,HD), dK at [64,64+HD).  !KV: dQ of accumulator set a at [64 a, 64 a + HD)
  long long* trace = (TRACE && g_trace3 != nullptr && blockIdx.x == 0 && lane == 0 && KV) ? g_trace3 : nullptr;

  if (warp == W_PROD) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    ItemIter it(static_cast<int>(blockIdx.x), G, n_tiles, H);
    int st = 0; uint32_t ring_ph = 0;
    for (int n = 0; n < nmy; ++n, it.next()) {
      const int b = it.b, h = it.h, t = it.t;
      const int kb = n & 1;
      mbar_wait(&kv_empty[kb], ((n >> 1) & 1) ^ 1u);
      if (leader) {
        uint8_t* r1 = sR + kb * 2 * TILE_BYTES;
        mbar_expect_tx(&kv_full[kb], 2 * TILE_BYTES);
        if (KV) {
          tma_load_2d(smem_u32(r1), &tmQKV128, &kv_full[kb], D + h * HD, b * S + t * TILE);                    // K
          tma_load_2d(smem_u32(r1 + TILE_BYTES), &tmQKV128, &kv_full[kb], 2 * D + h * HD, b * S + t * TILE);   // V
        } else {
          tma_load_2d(smem_u32(r1), &tmQKV128, &kv_full[kb], h * HD, b * S + t * TILE);                        // Q
          tma_load_2d(smem_u32(r1 + TILE_BYTES), &tmDO128, &kv_full[kb], h * HD, b * S + t * TILE);            // dO
        }
      }
      __syncwarp();
      for (int i = 0; i < nblk; ++i) {
        mbar_wait(&ring_empty[st], ring_ph ^ 1u);
        if (leader) {
          mbar_expect_tx(&ring_full[st], 2 * HALF_BYTES);
          if (KV) {
            tma_load_2d(smem_u32(sT1 + st * HALF_BYTES), &tmQKV64, &ring_full[st], h * HD, b * S + i * 64);    // Q_i
            tma_load_2d(smem_u32(sT2 + st * HALF_BYTES), &tmDO64, &ring_full[st], h * HD, b * S + i * 64);     // dO_i
          } else {
            tma_load_2d(smem_u32(sT1 + st * HALF_BYTES), &tmQKV64, &ring_full[st], D + h * HD, b * S + i * 64);      // K_j
            tma_load_2d(smem_u32(sT2 + st * HALF_BYTES), &tmQKV64, &ring_full[st], 2 * D + h * HD, b * S + i * 64);  // V_j
          }
        }
        __syncwarp();
        if (++st == STAGES) { st = 0; ring_ph ^= 1u; }
      }
    }
  } else if (warp == W_MMA_S) {
    // ===================== MMA issuer 1: S / dP of block gb into buffer gb % NBUF, as far ahead as buffers and tiles allow ====
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(TILE, 64, false, false);
    const uint32_t idesc_s16 = make_idesc_bf16(TILE, 16, false, false);
    int st = 0; uint32_t ring_ph = 0;
    int buf = 0; uint32_t buf_ph = 0;                 // parity of the buffer's current use
    int gb = 0;
    for (int n = 0; n < nmy; ++n) {
      const int kb = n & 1;
      mbar_wait(&kv_full[kb], (n >> 1) & 1);
      const uint64_t dR1 = make_sdesc_sw128(smem_u32(sR + kb * 2 * TILE_BYTES), false, 0);
      const uint64_t dR2 = make_sdesc_sw128(smem_u32(sR + kb * 2 * TILE_BYTES + TILE_BYTES), false, 0);
      for (int i = 0; i < nblk; ++i, ++gb) {
        TR3(0, gb, 0);
        mbar_wait(&ring_full[st], ring_ph);
        mbar_wait(&buf_free[buf], buf_ph ^ 1u);       // the accumulate MMAs of the buffer's previous block have completed
        TR3(0, gb, 1);
        tc_fence_after();
        if (leader) {
          const uint32_t tS = tmem_base + ACC_COLS + buf * 128, tdP = tS + 64;
          const uint64_t dT1 = make_sdesc_sw128(smem_u32(sT1 + st * HALF_BYTES), false, 0);
          const uint64_t dT2 = make_sdesc_sw128(smem_u32(sT2 + st * HALF_BYTES), false, 0);
          const uint32_t id = (i == nblk - 1 && tail16) ? idesc_s16 : idesc_s;
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {       // the two accumulate chains interleaved: consecutive MMAs independent
            tc_mma(tS, dR1 + ks * 2, dT1 + ks * 2, id, ks > 0 ? 1u : 0u);
            tc_mma(tdP, dR2 + ks * 2, dT2 + ks * 2, id, ks > 0 ? 1u : 0u);
          }
          tc_commit(&s_full[buf]);
        }
        __syncwarp();
        TR3(0, gb, 2);
        if (++st == STAGES) { st = 0; ring_ph ^= 1u; }
        if (++buf == NBUF) { buf = 0; buf_ph ^= 1u; }
      }
    }
  } else if (warp == W_MMA_A) {
    // ===================== MMA issuer 2: accumulate MMAs (A = bf16 operands the softmax warps left in tensor memory) =========
    const bool leader = elect_one();
    const uint32_t idesc_g = make_idesc_bf16(TILE, HD, false, true);
    int st = 0;
    int buf = 0; uint32_t buf_ph = 0;
    int gb = 0;
    for (int n = 0; n < nmy; ++n) {
      const int ab = KV ? 0 : (n & 1);                 // accumulator set
      const int ause = KV ? n : (n >> 1);              // how often it has been used before
      const uint32_t tmem_acc1 = tmem_base, tmem_acc2 = tmem_base + 64 - (KV ? 0 : 64) + (KV ? 0 : 64 * ab);
      for (int i = 0; i < nblk; ++i, ++gb) {
        TR3(1, gb, 0);
        mbar_wait(&p_full[buf], buf_ph);              // bf16 operands of block gb sit in TMEM (and its fp32 scores were read)
        TR3(1, gb, 1);
        if (i == 0) mbar_wait(&acc_empty[ab], (ause & 1) ^ 1u);   // the set's previous item has been drained
        TR3(1, gb, 2);
        tc_fence_after();
        if (leader) {
          const uint32_t tS = tmem_base + ACC_COLS + buf * 128, tdP = tS + 64;
          const uint64_t dT1m = make_sdesc_sw128(smem_u32(sT1 + st * HALF_BYTES), true, HALF_BYTES);
          const uint64_t dT2m = make_sdesc_sw128(smem_u32(sT2 + st * HALF_BYTES), true, HALF_BYTES);
          const uint32_t acc = i > 0 ? 1u : 0u;
          // k-step ks covers streamed rows [16 ks, 16 ks + 16): packed by column half ks / 2 at column 32 (ks / 2) + 8 (ks % 2)
          if (i == nblk - 1 && tail16) {
            if (KV) tc_mma_ts(tmem_acc1, tS, dT2m, idesc_g, acc);
            tc_mma_ts(tmem_acc2, tdP, dT1m, idesc_g, acc);
          } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (KV) tc_mma_ts(tmem_acc1, tS + (ks >> 1) * 32 + (ks & 1) * 8, dT2m + ks * 128, idesc_g, ks > 0 ? 1u : acc);
              tc_mma_ts(tmem_acc2, tdP + (ks >> 1) * 32 + (ks & 1) * 8, dT1m + ks * 128, idesc_g, ks > 0 ? 1u : acc);
            }
          }
          tc_commit(&ring_empty[st]);
          tc_commit(&buf_free[buf]);
          // the item's accumulators are complete: tell the group that drains it (the owner of block gb + 1).  One barrier
          // per group, so that each waiter sees every phase of its barrier in order (parity waits must not skip phases)
          if (i == nblk - 1) { tc_commit(&done[(gb + 1) & 1]); tc_commit(&kv_empty[n & 1]); }
        }
        __syncwarp();
        TR3(1, gb, 3);
        if (++st == STAGES) st = 0;
        if (++buf == NBUF) { buf = 0; buf_ph ^= 1u; }
      }
    }
  } else {
    // ===================== softmax-backward warps: one row (TMEM lane) per thread =====================
    const int grp = warp >> 3, wg = (warp >> 2) & 1, q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float* wstat = sStat + warp * 128;
    if ((warp & 7) != 0) trace = nullptr;            // warps 0 and 8 (first warp of each group) report
    const int trole = 2 + grp;

    // accumulator drain of item m (coordinates b, h, t): done by ONE group (8 warps, two per lane quarter)
    uint32_t done_ph = 0;                              // parity of this group's next `done` phase
    auto drain = [&](int m, int b, int h, int t, int tgb) {
      const int ab = KV ? 0 : (m & 1);
      const bool warp_active = t * TILE + q * 32 < S;
      TR3(trole, tgb, 8);
      mbar_wait(&done[grp], done_ph);
      done_ph ^= 1u;
      tc_fence_after();
      TR3(trole, tgb, 9);
      const bool drains = warp_active && (KV || wg == 0);
      // KV: column half 0 -> dK, 1 -> dV (HD columns each).  !KV: half 0 -> dQ
      const bool second = KV && wg == 1;
      const float sc = second ? 1.0f : scale;
      // Each thread holds one ROW; stored from registers, a warp-wide 16-byte store would touch 32 different lines.
      // The rows go through a private staging slot.  When every tile is full (n_tiles * 128 <= S: the shipped S = 128 k + 1
      // shapes, whose last row has its own kernel) the slot is laid out as a TMA box (32 rows x HD: dense 96-byte rows for
      // HD = 48, the 128-byte swizzle for HD = 64) and ONE lane hands it to cp.async.bulk.tensor -- the first version's
      // per-lane LDS -> STG loop held the draining group for ~1600-2000 cycles per item (profiles/r02_attn_bwd3_timeline_v5.txt).
      // Otherwise (a partial last tile) eight lanes write one row's contiguous HD x 2 bytes from the swizzled slot.
      uint8_t* stg = sDrain + warp * DRAIN_SLOT_BYTES;     // private: the two groups' drains of consecutive items may overlap
      auto slot_off = [&](int row, int ch) -> int {
        return (tma_drain && HD == 48) ? row * 96 + ch * 16 : row * 128 + ((ch ^ (row & 7)) << 4);
      };
      if (drains) {
        if (tma_drain) {                                   // the slot's previous box has been read by the TMA unit
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        const uint32_t src = tmem_base + (KV ? (second ? 0 : 64) : 64 * ab);
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {            // 32 (+ 32 | + 16) columns: half the registers of one pass
          uint32_t o[32];
          if (c0 + 32 <= HD) tmem_ld32(src + lane_off + c0, o);
          else tmem_ld16(src + lane_off + c0, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c0 + 8 * c < HD) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[8 * c]) * sc, __uint_as_float(o[8 * c + 1]) * sc);
              u.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * sc, __uint_as_float(o[8 * c + 3]) * sc);
              u.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * sc, __uint_as_float(o[8 * c + 5]) * sc);
              u.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * sc, __uint_as_float(o[8 * c + 7]) * sc);
              *reinterpret_cast<uint4*>(stg + slot_off(lane, c0 / 8 + c)) = u;
            }
          }
        }
        if (tma_drain) fence_proxy_async_smem();           // the staged rows become visible to the TMA unit
      }
      // qkv-bias gradient (column sums of dqkv over all tokens, attentionblock.py:36 qkv_bias): summed here from the staged
      // bf16 rows -- two columns per lane, one red.global per column and warp tile -- instead of re-reading dqkv from HBM
      auto tile_colsum = [&](int col) {
        __syncwarp();
        if (lane < HD / 2) {
          const int row0c = t * TILE + q * 32;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(stg + slot_off(r, lane >> 2) + (lane & 3) * 4));
            if (row0c + r < S) { s0 += f.x; s1 += f.y; }
          }
          atomicAdd(colsum + col + 2 * lane, s0);
          atomicAdd(colsum + col + 2 * lane + 1, s1);
        }
      };
      // the accumulators have left tensor memory: hand the set back to the MMA issuer BEFORE the global-store path
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      TR3(trole, tgb, 10);
      if (drains) {
        const int row0 = t * TILE + q * 32;
        const int col = h * HD + (KV ? (second ? 2 * D : D) : 0);
        if (tma_drain) {
          if (lane == 0) {
            tma_store_2d(&tmOut, smem_u32(stg), col, b * S + row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
        } else {
          __syncwarp();
          const int chunk = lane & 7, rsub = lane >> 3;
          if (chunk < HD / 8) {
            bf16* dst0 = dqkv + (static_cast<long long>(b) * S + row0) * (3LL * D) + col + chunk * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = k * 4 + rsub;
              if (row0 + r < S) {
                const uint4 v = *reinterpret_cast<const uint4*>(stg + slot_off(r, chunk));
                *reinterpret_cast<uint4*>(dst0 + static_cast<long long>(r) * (3LL * D)) = v;
              }
            }
          }
          __syncwarp();
        }
        if (colsum != nullptr) tile_colsum(col);             // off the accumulator hand-over path; the slot is rewritten next item
      }
    };

    // Position of this group's current block and of its next one (statistics are fetched one own block ahead), advanced
    // incrementally.  `prev` = coordinates of item n - 1.
    struct Pos { int n, i; };
    Pos cur{0, grp};
    ItemIter cit(static_cast<int>(blockIdx.x), G, n_tiles, H);
    int pb = 0, ph_ = 0, pt = 0;                       // previous item's coordinates
    auto norm = [&](Pos& p, ItemIter& it, bool track_prev) {
      while (p.i >= nblk) {
        p.i -= nblk; ++p.n;
        if (track_prev) { pb = it.b; ph_ = it.h; pt = it.t; }
        it.next();
      }
    };
    norm(cur, cit, true);
    Pos nxt = cur;
    ItemIter nit = cit;
    // statistics of a block: KV -> per streamed column (queries of block i): lse * log2e | delta for this warp's 32 columns;
    //                        !KV -> per row (this thread's query row of the item's tile)
    auto load_stats = [&](const Pos& p, const ItemIter& it, float& a, float& d) {
      const long long base = (static_cast<long long>(it.b) * H + it.h) * S;
      // Nothing may depend on the loaded registers here -- neither a multiply nor the `else a = 0` of a bounds select (a
      // predicated move to the load's destination waits on its scoreboard): the in-order warp would sit out the whole
      // global-load latency on the spot (400-1100 cycles per block in profiles/r02_attn_bwd3_timeline_v1.txt).  The row
      // index is clamped instead; columns / rows beyond S are masked or never stored.
      const int r = min(KV ? (p.i * 64 + wg * 32 + lane) : (it.t * TILE + q * 32 + lane), S - 1);
      a = lse[base + r];
      d = delta[base + r];
    };
    // dK/dV pass: the per-column statistics of the NEXT own block travel global -> shared with cp.async (no registers: a
    // register destination was spilled by the compiler right behind the load, which parked the warp for the whole
    // global-load latency, ~900 cycles per block in profiles/r02_attn_bwd3_timeline_v5.txt); each lane then rewrites its
    // own element scaled and negated.  Slot (cnt & 1) of the warp's private 2 x [lse 32 | delta 32] floats.
    auto prefetch_stats = [&](const Pos& p, const ItemIter& it, float* slot) {
      const long long base = (static_cast<long long>(it.b) * H + it.h) * S;
      const int r = min(p.i * 64 + wg * 32 + lane, S - 1);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(slot + lane)), "l"(lse + base + r) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(slot + 32 + lane)), "l"(delta + base + r) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float st_a = 0.f, st_d = 0.f;
    if (cur.n < nmy) {
      if (KV) prefetch_stats(cur, cit, wstat);
      else load_stats(cur, cit, st_a, st_d);
    }
    int buf = grp % NBUF; uint32_t buf_ph = 0;         // buffer / use parity of block gb = grp, advanced by 2 per own block
    int cnt = 0;
    for (int gb = grp; gb < total; gb += 2, ++cnt) {
      const int n = cur.n, i = cur.i;
      TR3(trole, gb, 0);
      const bool warp_active = cit.t * TILE + q * 32 < S;    // warps without a valid row only keep the barriers moving
      const uint32_t tS = tmem_base + ACC_COLS + buf * 128, tdP = tS + 64;
      float* stat = wstat + (cnt & 1) * 64;
      // NEGATED statistics (!KV: of this thread's row): added, not subtracted
      const float row_a = KV ? 0.f : -st_a * LOG2E, row_d = KV ? 0.f : -st_d;
      if (KV) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        stat[lane] = -stat[lane] * LOG2E;
        stat[32 + lane] = -stat[32 + lane];
        __syncwarp();
      }
      TR3(trole, gb, 1);
      nxt.i += 2;
      norm(nxt, nit, false);
      TR3(trole, gb, 7);
      // one own block ahead: the global-load latency stays off the chain (dQ pass: row statistics change with the item only)
      if (nxt.n < nmy) {
        if (KV) prefetch_stats(nxt, nit, wstat + ((cnt + 1) & 1) * 64);
        else if (nxt.n != n) load_stats(nxt, nit, st_a, st_d);
      }
      TR3(trole, gb, 2);
      mbar_wait(&s_full[buf], buf_ph);
      TR3(trole, gb, 3);
      tc_fence_after();
      const int ncol = min(64, S - i * 64);                  // valid streamed columns in this block
      const bool t16 = (i == nblk - 1) && tail16;
      if (warp_active && !(t16 && wg == 1)) {
        uint32_t pk[16], dk[16];
        if (t16) {
          uint32_t sv[16], dv[16];
          tmem_ld16(tS + lane_off, sv);
          tmem_ld16(tdP + lane_off, dv);
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float a0 = KV ? stat[e] : row_a, a1 = KV ? stat[e + 1] : row_a;
            const float d0 = KV ? stat[32 + e] : row_d, d1 = KV ? stat[32 + e + 1] : row_d;
            const float p0 = e < ncol ? ex2f(fmaf(__uint_as_float(sv[e]), sl2, a0)) : 0.f;
            const float p1 = e + 1 < ncol ? ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, a1)) : 0.f;
            pk[e >> 1] = pack_bf16x2(p0, p1);
            dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) + d0), p1 * (__uint_as_float(dv[e + 1]) + d1));
          }
#pragma unroll
          for (int e = 8; e < 16; ++e) { pk[e] = 0u; dk[e] = 0u; }
        } else {
          uint32_t sv[32], dv[32];
          tmem_ld32_issue(tS + lane_off + wg * 32, sv);
          tmem_ld32_issue(tdP + lane_off + wg * 32, dv);
          tmem_ld_wait();
          const int lim = ncol - wg * 32;                    // valid columns in this warp's half
          if (lim < 32) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e >= lim) sv[e] = 0xff800000u;             // exp2(-inf) = 0: P and dS vanish outside the problem
          }
          const float2 sl2v = make_float2(sl2, sl2);
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            float4 ls, dl;
            if (KV) {
              ls = *reinterpret_cast<const float4*>(stat + e);
              dl = *reinterpret_cast<const float4*>(stat + 32 + e);
            } else {
              ls = make_float4(row_a, row_a, row_a, row_a);
              dl = make_float4(row_d, row_d, row_d, row_d);
            }
            if constexpr (POLY < 0) {     // scalar fp32 arithmetic (A/B)
              const float p0 = ex2f(fmaf(__uint_as_float(sv[e]), sl2, ls.x)), p1 = ex2f(fmaf(__uint_as_float(sv[e + 1]), sl2, ls.y));
              const float p2 = ex2f(fmaf(__uint_as_float(sv[e + 2]), sl2, ls.z)), p3 = ex2f(fmaf(__uint_as_float(sv[e + 3]), sl2, ls.w));
              if (KV) {
                pk[e >> 1] = pack_bf16x2(p0, p1);
                pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
              }
              dk[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) + dl.x), p1 * (__uint_as_float(dv[e + 1]) + dl.y));
              dk[(e >> 1) + 1] = pack_bf16x2(p2 * (__uint_as_float(dv[e + 2]) + dl.z), p3 * (__uint_as_float(dv[e + 3]) + dl.w));
            } else {
            // packed fp32: one issue slot per two elements for the shift, the (dP - delta) and the product
            const float2 p01 = ex2_pair<(POLY < 0 ? 0 : POLY)>(__ffma2_rn(make_float2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), sl2v, make_float2(ls.x, ls.y)), e >> 1);
            const float2 p23 = ex2_pair<(POLY < 0 ? 0 : POLY)>(__ffma2_rn(make_float2(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])), sl2v, make_float2(ls.z, ls.w)), (e >> 1) + 1);
            if (KV) {
              pk[e >> 1] = pack_bf16x2(p01.x, p01.y);
              pk[(e >> 1) + 1] = pack_bf16x2(p23.x, p23.y);
            }
            const float2 d01 = __fmul2_rn(p01, __fadd2_rn(make_float2(__uint_as_float(dv[e]), __uint_as_float(dv[e + 1])), make_float2(dl.x, dl.y)));
            const float2 d23 = __fmul2_rn(p23, __fadd2_rn(make_float2(__uint_as_float(dv[e + 2]), __uint_as_float(dv[e + 3])), make_float2(dl.z, dl.w)));
            dk[e >> 1] = pack_bf16x2(d01.x, d01.y);
            dk[(e >> 1) + 1] = pack_bf16x2(d23.x, d23.y);
            }
          }
        }
        TR3(trole, gb, 4);
        // in place: this thread's own lanes, inside the fp32 columns it has just read
        if (KV) tmem_st16(tS + lane_off + wg * 32, pk);
        tmem_st16(tdP + lane_off + wg * 32, dk);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);
      TR3(trole, gb, 5);
      // The group that owns the FIRST block of an item drains the previous item, after that block: its operands are in
      // place, so the accumulate issuer is held up by the drain alone (dK/dV pass) or not at all (dQ pass, two sets).
      if (i == 0 && n > 0) drain(n - 1, pb, ph_, pt, gb);
      TR3(trole, gb, 6);
      cur.i += 2;
      norm(cur, cit, true);
      buf += 2; if (buf >= NBUF) { buf -= NBUF; buf_ph ^= 1u; }
    }
    // the last item: drained by the group that would own the next block
    if ((total & 1) == grp) {
      // coordinates of item nmy - 1: cur has run past the end; cit is the item AFTER the last processed one only if it was
      // advanced, so recompute from scratch (once per kernel)
      ItemIter lit(static_cast<int>(blockIdx.x), G, n_tiles, H);
      for (int k = 0; k < nmy - 1; ++k) lit.next();
      drain(nmy - 1, lit.b, lit.h, lit.t, 63);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staged boxes outlive the CTA otherwise
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

int g_bwd3_tma_drain = 1;   // 0: results leave through the per-lane store loop (A/B, hct_attention_set_bwd3_drain)
template <int HD, bool KV, int POLY, bool TRACE = false>
int launch_one(const CUtensorMap& q128, const CUtensorMap& do128, const CUtensorMap& q64, const CUtensorMap& do64,
               const CUtensorMap& out, const float* lse, const float* delta, bf16* dqkv, float* colsum, int B, int S, int H, int n_tiles,
               cudaStream_t st) {
  static bool cfg = false;
  auto kernel = attn_bwd3_kernel<HD, KV, POLY, TRACE>;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { hct_set_error("cudaFuncSetAttribute(attn_bwd3): %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
    cfg = true;
  }
  const long long items = static_cast<long long>(B) * H * n_tiles;
  if (items <= 0) return HCT_OK;
  const int sms = hct_num_sms();
  const int grid = static_cast<int>(items < sms ? items : sms);
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  hct_launch_pdl(kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, st, q128, do128, q64, do64, out, lse, delta, dqkv, colsum, S, H, n_tiles,
                 static_cast<int>(items), scale, g_bwd3_tma_drain);
  return hct_check_launch(KV ? "attn_bwd3_kernel<dK/dV>" : "attn_bwd3_kernel<dQ>");
}

}  // namespace

extern "C" int hct_attention_set_bwd3_drain(int tma) { g_bwd3_tma_drain = tma != 0; return HCT_OK; }
extern "C" int hct_attention_trace3(void* buf) {      // device buffer of >= 4 * 64 * 16 int64 (or NULL): bwd3 dK/dV timeline
  long long* p = static_cast<long long*>(buf);
  g_trace3_host = p != nullptr;                      // hd = 48 dK/dV pass runs its instrumented instantiation while set
  return cudaMemcpyToSymbol(g_trace3, &p, sizeof(p)) == cudaSuccess ? HCT_OK : HCT_ERR_CUDA;
}

static int g_bwd3_poly = 0;      // -1: scalar fp32; 0 (default): packed fp32; 2: packed + 2 of 8 exponential pairs on the FMA pipe
int hct_attention_bwd3_set_poly(int n) { g_bwd3_poly = n; return HCT_OK; }
template <int POLY>
static int launch_both(const CUtensorMap& q128, const CUtensorMap& do128, const CUtensorMap& q64, const CUtensorMap& do64,
                       const CUtensorMap& out, const float* lse, const float* delta, bf16* dq, float* colsum, int B, int S, int H, int hd,
                       int n_tiles, cudaStream_t st) {
  int rc;
  if (hd == 64) {
    rc = launch_one<64, true, POLY>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, n_tiles, st); if (rc) return rc;
    return launch_one<64, false, POLY>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, n_tiles, st);
  }
  if (POLY == 0 && g_trace3_host) rc = launch_one<48, true, 0, true>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, n_tiles, st);
  else rc = launch_one<48, true, POLY>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, n_tiles, st);
  if (rc) return rc;
  return launch_one<48, false, POLY>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, n_tiles, st);
}

// n_tiles full 128-row tiles per (batch, head) on tcgen05 (rows behind them: hct_attention_tail.cu)
// colsum (may be NULL): fp32 [3 * H * hd], the column sums of the written dqkv rows are ADDED to it
int hct_attention_bwd3(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, float* colsum, int B, int S,
                       int H, int hd, int n_tiles, cudaStream_t st) {
  CUtensorMap q128, q64, do128, do64;
  const long long D = static_cast<long long>(H) * hd, D3 = 3 * D, rows = static_cast<long long>(B) * S;
  HCT_REQUIRE(static_cast<long long>(B) * H * n_tiles <= 2147483647LL, "attention_bwd3: too many work items");
  int rc = hct_make_tmap_bf16_2d(&q128, qkv, D3, rows, D3, 64, TILE); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&q64, qkv, D3, rows, D3, 64, 64); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&do128, dout, D, rows, D, 64, TILE); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&do64, dout, D, rows, D, 64, 64); if (rc) return rc;
  // results: 32-row x hd boxes of dqkv (dense rows for hd = 48, 128-byte swizzle for hd = 64)
  CUtensorMap out;
  rc = hct_make_tmap_bf16_2d_sw(&out, dqkv, D3, rows, D3, hd, 32, hd == 64 ? 1 : 0); if (rc) return rc;
  bf16* dq = static_cast<bf16*>(dqkv);
  return g_bwd3_poly == 0   ? launch_both<0>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, hd, n_tiles, st)
         : g_bwd3_poly == 2 ? launch_both<2>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, hd, n_tiles, st)
                            : launch_both<-1>(q128, do128, q64, do64, out, lse, delta, dq, colsum, B, S, H, hd, n_tiles, st);
}
