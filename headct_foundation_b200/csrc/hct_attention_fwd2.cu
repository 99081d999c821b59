// Flash-attention forward on tcgen05 / TMEM, pipelined persistent variant ("fwd2").
//
// Why: the forward of hct_attention_sm100.cu runs four CTAs per SM, each a serial chain
//     S = Q K_j^T  ->  commit  ->  tcgen05.ld, max, exp2, tcgen05.st  ->  arrive  ->  O += P_j V_j  ->  S_{j+1} ...
// whose four softmax warps idle while the MMAs of their own block run (and vice versa): 0.52 ms at the decoder shape against a
// 0.25 ms floor set by the exponentials (16 MUFU results per clock and SM).  Here ONE persistent CTA per SM works on TWO query
// tiles of a head at a time (slot g = 0 / 1, 128 rows each), every slot with TWO score buffers in tensor memory: the MMA warp
// issues S(j + 1) of a slot before it waits for P(j), so a slot's softmax warps find their next scores ready when they have
// handed a block over, and the only serial chain left is the softmax arithmetic itself.  Sixteen softmax warps: a slot's
// eight split the 64 columns of a block in halves (two warps per TMEM lane quarter) and exchange their half-row maxima
// through shared memory; O is rescaled lazily (only when the running maximum moves by more than 2^8) by both halves.
//   tensor memory: slot g: S buffers at 192 g + {0, 64} (P overwrites the first 32 columns in place), O at 192 g + 128
//   warps: 0-15 softmax (slot = w >> 3, column half = (w >> 2) & 1, lane quarter = w & 3), 16 TMA producer, 17 MMA issuer
// Work items: (batch, head, pair of 128-row query tiles); K_j / V_j are streamed ONCE per pair through a 4-deep TMA ring.
#include "../../include/hct_b200.h"
#include "hct_tcgen05.cuh"

namespace {
using namespace hct_tc;

constexpr int TILE = 128;
constexpr int TILE_BYTES = TILE * 128;     // 128 rows x 64 bf16
constexpr int KB = 64;                     // keys per block
constexpr int KV_BYTES = KB * 128;         // [64 keys][64 bf16]
constexpr float LOG2E = 1.4426950408889634f;
constexpr int NSW = 16;
constexpr int W_PROD = 16, W_MMA = 17;
constexpr int THREADS = 18 * 32;
constexpr int STAGES = 4;                  // K / V ring
constexpr int TMEM_COLS = 512;
constexpr int SLOT_COLS = 192;             // S0 | S1 | O
constexpr int OUT_SLOT_BYTES = 32 * 64;    // per softmax warp: 32 rows x (hd / 2) bf16
constexpr int SMEM_BYTES = 2 * 2 * TILE_BYTES + STAGES * 2 * KV_BYTES + NSW * 2 * 32 * 4 * 2 + NSW * OUT_SLOT_BYTES + 1024 + 512;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void pair_sync(int id) {      // the two warps that share 32 rows
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmOut, bf16* __restrict__ out, float* __restrict__ lse, int S, int H,
                 int n_tiles, int n_pairs, int n_items, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                    // [2 item buffers][2 slots][128 rows][64]
  uint8_t* sK = smem + 2 * 2 * TILE_BYTES;               // STAGES x [64 keys][64]
  uint8_t* sV = sK + STAGES * KV_BYTES;
  float* sX = reinterpret_cast<float*>(sV + STAGES * KV_BYTES);             // [16 warps][2 uses][max 32 | sum 32]
  uint8_t* sOut = reinterpret_cast<uint8_t*>(sX) + NSW * 2 * 32 * 4 * 2;    // [16 warps][32 rows][hd bytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + NSW * OUT_SLOT_BYTES);
  uint64_t *q_full = bars /*[2]*/, *q_empty = bars + 2 /*[2]*/, *kv_full = bars + 4 /*[STAGES]*/, *kv_empty = kv_full + STAGES,
           *s_full = kv_empty + STAGES /*[2 slots][2 bufs]*/, *p_full = s_full + 4 /*[2][2]*/, *pv_done = p_full + 4 /*[2]*/,
           *o_full = pv_done + 2 /*[2]*/, *o_empty = o_full + 2 /*[2]*/;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int D = H * HD;
  const int nkb = (S + KB - 1) / KB;
  const bool tail16 = S - (nkb - 1) * KB <= 16;            // last key block is computed 16 columns wide
  const float sl2 = scale * LOG2E;
  const int G = static_cast<int>(gridDim.x);
  const int nmy = (n_items - static_cast<int>(blockIdx.x) + G - 1) / G;      // items of this CTA: blockIdx.x + n * G

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&pv_done[i], 1); mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 8);
    }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;

  // item n of this CTA -> (b, h, pair): pairs of a head are neighbours in the item order (their K / V meet in L2)
  auto item_coords = [&](int n, int& b, int& h, int& p) {
    const int it = static_cast<int>(blockIdx.x) + n * G;
    p = it % n_pairs;
    const int bh = it / n_pairs;
    h = bh % H;
    b = bh / H;
  };

  if (warp == W_PROD) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int r = 0;                                             // K / V blocks loaded so far
    for (int n = 0; n < nmy; ++n) {
      int b, h, p;
      item_coords(n, b, h, p);
      const int kb = n & 1;
      const int nt = (2 * p + 1 < n_tiles) ? 2 : 1;        // query tiles of this pair
      mbar_wait(&q_empty[kb], ((n >> 1) & 1) ^ 1u);
      if (leader) {
        mbar_expect_tx(&q_full[kb], nt * TILE_BYTES);
        for (int g = 0; g < nt; ++g)
          tma_load_2d(smem_u32(sQ + (kb * 2 + g) * TILE_BYTES), &tmQ, &q_full[kb], h * HD, b * S + (2 * p + g) * TILE);
      }
      __syncwarp();
      for (int j = 0; j < nkb; ++j, ++r) {
        const int st = r % STAGES;
        mbar_wait(&kv_empty[st], ((r / STAGES) & 1) ^ 1u);
        if (leader) {
          mbar_expect_tx(&kv_full[st], 2 * KV_BYTES);
          tma_load_2d(smem_u32(sK + st * KV_BYTES), &tmKV, &kv_full[st], D + h * HD, b * S + j * KB);
          tma_load_2d(smem_u32(sV + st * KV_BYTES), &tmKV, &kv_full[st], 2 * D + h * HD, b * S + j * KB);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(TILE, KB, false, false);
    const uint32_t idesc_s16 = make_idesc_bf16(TILE, 16, false, false);
    const uint32_t idesc_o = make_idesc_bf16(TILE, HD, false, true);
    int use[2] = {0, 0};                                   // blocks whose S has been issued, per slot
    int usep[2] = {0, 0};                                  // blocks whose P V has been issued, per slot
    int itc[2] = {0, 0};                                   // items per slot
    int rs = 0, rp = 0;                                    // ring positions of the S stream and of the P V stream
    for (int n = 0; n < nmy; ++n) {
      int b, h, p;
      item_coords(n, b, h, p);
      const int kb = n & 1;
      const int nt = (2 * p + 1 < n_tiles) ? 2 : 1;
      mbar_wait(&q_full[kb], (n >> 1) & 1);
      auto issue_s = [&](int j) {                          // S_g(j) = Q_g K_j^T for the pair's tiles
        const int st = rs % STAGES;
        mbar_wait(&kv_full[st], (rs / STAGES) & 1);
        tc_fence_after();
        if (leader) {
          const uint64_t dK = make_sdesc_sw128(smem_u32(sK + st * KV_BYTES), false, 0);
          const uint32_t id = (j == nkb - 1 && tail16) ? idesc_s16 : idesc_s;
          for (int g = 0; g < nt; ++g) {
            const uint64_t dQ = make_sdesc_sw128(smem_u32(sQ + (kb * 2 + g) * TILE_BYTES), false, 0);
            const uint32_t tS = tmem_base + g * SLOT_COLS + (use[g] & 1) * 64;
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) tc_mma(tS, dQ + ks * 2, dK + ks * 2, id, ks > 0 ? 1u : 0u);
            tc_commit(&s_full[g * 2 + (use[g] & 1)]);
          }
        }
        __syncwarp();
        for (int g = 0; g < nt; ++g) ++use[g];
        ++rs;
      };
      issue_s(0);
      for (int j = 0; j < nkb; ++j) {
        // the next block's scores go out BEFORE this block's probabilities are waited for: the score buffer they land in
        // held P(j - 1), whose P V MMAs this thread has issued already (one in-order pipe)
        if (j + 1 < nkb) issue_s(j + 1);
        const int st = rp % STAGES;
        for (int g = 0; g < nt; ++g) {
          mbar_wait(&p_full[g * 2 + (usep[g] & 1)], (usep[g] >> 1) & 1);
          if (j == 0) mbar_wait(&o_empty[g], (itc[g] & 1) ^ 1u);       // the slot's previous output has left tensor memory
        }
        tc_fence_after();
        if (leader) {
          const uint64_t dV = make_sdesc_sw128(smem_u32(sV + st * KV_BYTES), true, KV_BYTES);
          const uint32_t acc = j > 0 ? 1u : 0u;
          for (int g = 0; g < nt; ++g) {
            const uint32_t tP = tmem_base + g * SLOT_COLS + (usep[g] & 1) * 64;
            const uint32_t tO = tmem_base + g * SLOT_COLS + 128;
            if (j == nkb - 1 && tail16) {
              tc_mma_ts(tO, tP, dV, idesc_o, acc);
            } else {
#pragma unroll
              for (int ks = 0; ks < KB / 16; ++ks) tc_mma_ts(tO, tP + ks * 8, dV + ks * 128, idesc_o, ks > 0 ? 1u : acc);
            }
            tc_commit(&pv_done[g]);
            if (j == nkb - 1) tc_commit(&o_full[g]);
          }
          tc_commit(&kv_empty[st]);
          if (j == nkb - 1) tc_commit(&q_empty[kb]);
        }
        __syncwarp();
        for (int g = 0; g < nt; ++g) ++usep[g];
        ++rp;
      }
      for (int g = 0; g < nt; ++g) ++itc[g];
    }
  } else if (warp < NSW) {
    // ===================== softmax warps =====================
    const int g = warp >> 3, wg = (warp >> 2) & 1, q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tslot = tmem_base + g * SLOT_COLS;
    const uint32_t tO = tslot + 128;
    float* xme = sX + warp * 128;                          // [2 uses][max 32 | sum 32]
    float* xpartner = sX + (warp ^ 4) * 128;
    const int pair_id = 1 + g * 4 + q;
    uint8_t* stg = sOut + warp * OUT_SLOT_BYTES;
    constexpr int HC = HD / 2;                             // O columns of this warp: [wg * HC, wg * HC + HC)
    int use = 0, itc = 0;
    for (int n = 0; n < nmy; ++n) {
      int b, h, p;
      item_coords(n, b, h, p);
      const int t = 2 * p + g;
      if (t >= n_tiles) continue;                          // this slot has no tile in a head's last, odd pair
      const int row = t * TILE + q * 32 + lane;
      const bool warp_active = t * TILE + q * 32 < S;      // a warp whose 32 rows all lie past S only keeps the barriers moving
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < nkb; ++j, ++use) {
        const int buf = use & 1;
        const uint32_t tS = tslot + buf * 64;
        mbar_wait(&s_full[g * 2 + buf], (use >> 1) & 1);
        tc_fence_after();
        const int nvalid = min(KB, S - j * KB);
        const bool t16 = j == nkb - 1 && tail16;
        const int lim = nvalid - wg * 32;                  // valid columns of this warp's half (<= 0: none)
        uint32_t v[32];
        float pmax = -INFINITY;
        const bool has_cols = warp_active && !(t16 && wg == 1);   // masked columns still get their zeros written
        if (has_cols) {
          if (t16) tmem_ld16(tS + lane_off, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          else tmem_ld32(tS + lane_off + wg * 32, v);
          const int ncol = t16 ? 16 : 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < ncol) {
              if (i >= lim) v[i] = 0xff800000u;
              pmax = fmaxf(pmax, __uint_as_float(v[i]));
            }
          }
        }
        // row maximum over both halves
        xme[buf * 64 + lane] = pmax;
        pair_sync(pair_id);
        const float mb = fmaxf(pmax, xpartner[buf * 64 + lane]);
        // lazy rescaling: the reference point only moves when the block maximum exceeds it by more than 2^8
        float alpha = 1.0f;
        if (j == 0) {
          m = mb;
        } else if ((mb - m) * sl2 > 8.0f) {
          alpha = ex2f((m - mb) * sl2);
          m = mb;
        }
        const float msc = m * sl2;
        float rsum = 0.f;
        if (has_cols) {
          if (t16) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float a0 = ex2f(fmaf(__uint_as_float(v[i]), sl2, -msc)), a1 = ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -msc));
              rsum += a0 + a1;
              pk[i >> 1] = pack_bf16x2(a0, a1);
            }
            tmem_st8(tS + lane_off, pk);
          } else {
            uint32_t pk[16];
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float a0 = ex2f(fmaf(__uint_as_float(v[i]), sl2, -msc)), a1 = ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -msc));
              const float a2 = ex2f(fmaf(__uint_as_float(v[i + 2]), sl2, -msc)), a3 = ex2f(fmaf(__uint_as_float(v[i + 3]), sl2, -msc));
              rs0 += a0 + a1;
              rs1 += a2 + a3;
              pk[i >> 1] = pack_bf16x2(a0, a1);
              pk[(i >> 1) + 1] = pack_bf16x2(a2, a3);
            }
            rsum = rs0 + rs1;
            // keys 2c, 2c+1 -> column c: this half's 32 keys land in columns [16 wg, 16 wg + 16); the partner has read its
            // scores (it passed the pair barrier with them in registers), so overwriting its first columns is safe
            tmem_st16(tS + lane_off + wg * 16, pk);
          }
        }
        l = l * alpha + rsum;
        // O is touched by the previous block's P V MMAs until pv_done; rescale (rarely) behind it, own half of the columns
        if (use > 0) mbar_wait(&pv_done[g], (use - 1) & 1);
        if (j > 0 && warp_active && !__all_sync(0xffffffffu, alpha == 1.0f)) {
          tc_fence_after();
          uint32_t o[8];
#pragma unroll 1
          for (int c0 = 0; c0 < HC; c0 += 8) {
            tmem_ld8(tO + lane_off + wg * HC + c0, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st8(tO + lane_off + wg * HC + c0, o);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g * 2 + buf]);
      }
      // ---- epilogue of the tile: O / l -> bf16, log-sum-exp
      mbar_wait(&o_full[g], itc & 1);
      tc_fence_after();
      xme[(itc & 1) * 64 + 32 + lane] = l;
      pair_sync(pair_id);
      const float lt = l + xpartner[(itc & 1) * 64 + 32 + lane];
      const bool full_tile = t * TILE + TILE <= S;
      if (warp_active) {
        const float inv = 1.0f / lt;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slot's previous box has been read
        __syncwarp();
        uint32_t o[HC];
        if constexpr (HC == 32) {
          tmem_ld32(tO + lane_off + wg * HC, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        } else {
          tmem_ld16(tO + lane_off + wg * HC, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
          tmem_ld8(tO + lane_off + wg * HC + 16, *reinterpret_cast<uint32_t(*)[8]>(&o[16]));
        }
        uint4 u[HC / 8];
#pragma unroll
        for (int c = 0; c < HC / 8; ++c) {
          u[c].x = pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
          u[c].y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
          u[c].z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
          u[c].w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
        }
        if (full_tile) {
          // dense [32 rows][HC bf16] box for one cp.async.bulk.tensor store per warp
#pragma unroll
          for (int c = 0; c < HC / 8; ++c) *reinterpret_cast<uint4*>(stg + lane * (HC * 2) + c * 16) = u[c];
          fence_proxy_async_smem();
        } else if (row < S) {
          bf16* orow = out + (static_cast<long long>(b) * S + row) * D + h * HD + wg * HC;
#pragma unroll
          for (int c = 0; c < HC / 8; ++c) *reinterpret_cast<uint4*>(orow + 8 * c) = u[c];
        }
        if (wg == 0 && row < S) lse[(static_cast<long long>(b) * H + h) * S + row] = m * scale + logf(lt);
      }
      // O has left tensor memory: the slot's next tile may accumulate
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[g]);
      if (warp_active && full_tile && lane == 0) {
        tma_store_2d(&tmOut, smem_u32(stg), h * HD + wg * HC, b * S + t * TILE + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
      ++itc;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staged boxes outlive the CTA otherwise
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <int HD>
int launch_fwd2(const CUtensorMap& tmq, const CUtensorMap& tmkv, const CUtensorMap& tmo, bf16* out, float* lse, int B, int S, int H,
                int n_tiles, cudaStream_t st) {
  static bool cfg = false;
  auto kernel = attn_fwd2_kernel<HD>;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { hct_set_error("cudaFuncSetAttribute(attn_fwd2): %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
    cfg = true;
  }
  const int n_pairs = (n_tiles + 1) / 2;
  const long long items = static_cast<long long>(B) * H * n_pairs;
  if (items <= 0) return HCT_OK;
  const int sms = hct_num_sms();
  const int grid = static_cast<int>(items < sms ? items : sms);
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  cudaError_t e = hct_launch_pdl(kernel, dim3(grid), dim3(THREADS), SMEM_BYTES, st, tmq, tmkv, tmo, out, lse, S, H, n_tiles, n_pairs,
                                 static_cast<int>(items), scale);
  if (e != cudaSuccess) { hct_set_error("launch(attn_fwd2): %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return HCT_ERR_CUDA; }
  return hct_check_launch("attn_fwd2_kernel");
}

}  // namespace

// n_tiles 128-row query tiles per (batch, head), the last one possibly partial
int hct_attention_fwd2(const void* qkv, void* out, float* lse, int B, int S, int H, int hd, int n_tiles, cudaStream_t st) {
  CUtensorMap tmq, tmkv, tmo;
  const long long D = static_cast<long long>(H) * hd, D3 = 3 * D, rows = static_cast<long long>(B) * S;
  HCT_REQUIRE(static_cast<long long>(B) * H * ((n_tiles + 1) / 2) <= 2147483647LL, "attention_fwd2: too many work items");
  int rc = hct_make_tmap_bf16_2d(&tmq, qkv, D3, rows, D3, 64, TILE); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d(&tmkv, qkv, D3, rows, D3, 64, KB); if (rc) return rc;
  rc = hct_make_tmap_bf16_2d_sw(&tmo, out, D, rows, D, hd / 2, 32, 0); if (rc) return rc;
  if (hd == 64) return launch_fwd2<64>(tmq, tmkv, tmo, static_cast<bf16*>(out), lse, B, S, H, n_tiles, st);
  return launch_fwd2<48>(tmq, tmkv, tmo, static_cast<bf16*>(out), lse, B, S, H, n_tiles, st);
}
