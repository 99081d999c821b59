// Flash-style self-attention, forward and backward, for the short ragged sequences of the 3D ViT
// (S = 129 / 513 / 517, head dim 64 or 48; F.scaled_dot_product_attention at attentionblock.py:61).
//
// v1 data path: bf16 mma.sync.m16n8k16 with fp32 accumulation, online softmax in registers,
// K/V (or Q/dO) tiles streamed through shared memory with cp.async double buffering.  It reads q/k/v
// straight out of the qkv GEMM output ([B,S,3,H,hd]) and writes [B,S,H*hd] -- no head transposes.
// Backward is split in two kernels (dK/dV per key block, dQ per query block) so that no atomics are
// needed and results are deterministic.
#include "../../include/hct_b200.h"
#include "hct_common.cuh"

namespace {

constexpr int BLK = 64;        // query rows / key rows per tile
constexpr int NTHREADS = 128;  // 4 warps x 16 rows
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD>
struct Tile {
  static constexpr int PITCH = HD + 8;            // elements; keeps ldmatrix rows on distinct banks
  static constexpr int BYTES = BLK * PITCH * 2;
  static constexpr int CHUNKS = HD / 8;           // 16-byte chunks per row
};

// copy a [64 x HD] tile (rows row0..row0+63 of a strided global matrix) into padded smem; rows >= S zero-filled
template <int HD>
__device__ __forceinline__ void load_tile(bf16* smem_tile, const bf16* gbase, long long row_stride, int row0, int S) {
  constexpr int CH = Tile<HD>::CHUNKS;
  for (int idx = threadIdx.x; idx < BLK * CH; idx += NTHREADS) {
    const int r = idx / CH, c = idx % CH;
    const bool ok = row0 + r < S;
    const bf16* src = gbase + static_cast<long long>(ok ? row0 + r : 0) * row_stride + c * 8;
    cp_async16(smem_u32(smem_tile + r * Tile<HD>::PITCH + c * 8), src, ok ? 16 : 0);
  }
}

// A fragments (16 rows x 16 k) at (row0, k0) of a padded tile
template <int HD>
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* tile, int row0, int k0, int lane) {
  ldsm_x4(a, smem_u32(tile + (row0 + (lane & 15)) * Tile<HD>::PITCH + k0 + (lane >> 4) * 8));
}
// B fragments for two adjacent n-tiles (n0..n0+15) x 16 k, storage [n][k]
template <int HD>
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], const bf16* tile, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  ldsm_x4(b, smem_u32(tile + (n0 + (mi >> 1) * 8 + (lane & 7)) * Tile<HD>::PITCH + k0 + (mi & 1) * 8));
}
// B fragments for two adjacent n-tiles (n0..n0+15) x 16 k, storage [k][n]  (transposing load)
template <int HD>
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], const bf16* tile, int k0, int n0, int lane) {
  const int mi = lane >> 3;
  ldsm_x4_trans(b, smem_u32(tile + (k0 + (mi & 1) * 8 + (lane & 7)) * Tile<HD>::PITCH + n0 + (mi >> 1) * 8));
}

// ------------------------------------------------------------------ forward
template <int HD>
__global__ void __launch_bounds__(NTHREADS)
attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int S, int H,
                float scale, int q_start) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + BLK * Tile<HD>::PITCH;                 // 2 stages
  bf16* sV = sK + 2 * BLK * Tile<HD>::PITCH;             // 2 stages
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * HD;
  const long long rs = 3LL * D;
  const bf16* qg = qkv + static_cast<long long>(b) * S * rs + h * HD;
  const bf16* kg = qg + D;
  const bf16* vg = qg + 2 * D;
  const int nkb = (S + BLK - 1) / BLK;
  const float sl2 = scale * LOG2E;

  const int q_base = q_start + qb * BLK;
  load_tile<HD>(sQ, qg, rs, q_base, S);
  load_tile<HD>(sK, kg, rs, 0, S);
  load_tile<HD>(sV, vg, rs, 0, S);
  cp_async_commit();

  uint32_t qf[HD / 16][4];
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};

  for (int kb = 0; kb < nkb; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < nkb) {
      load_tile<HD>(sK + (st ^ 1) * BLK * Tile<HD>::PITCH, kg, rs, (kb + 1) * BLK, S);
      load_tile<HD>(sV + (st ^ 1) * BLK * Tile<HD>::PITCH, vg, rs, (kb + 1) * BLK, S);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) load_a<HD>(qf[kk], sQ, warp * 16, kk * 16, lane);
    }
    const bf16* tK = sK + st * BLK * Tile<HD>::PITCH;
    const bf16* tV = sV + st * BLK * Tile<HD>::PITCH;

    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        load_b_nk<HD>(bb, tK, np * 16, kk * 16, lane);
        mma16816(s[2 * np], qf[kk], bb[0], bb[1]);
        mma16816(s[2 * np + 1], qf[kk], bb[2], bb[3]);
      }
    }
    // mask key columns beyond S
    const int col_base = kb * BLK + (lane & 3) * 2;
    if (kb * BLK + BLK > S) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = col_base + j * 8;
        if (c >= S) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (c + 1 >= S) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float alpha[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_i[r], mx[r]);
      alpha[r] = exp2f((m_i[r] - m_new) * sl2);
      m_i[r] = m_new;
      msc[r] = m_new * sl2;
    }
    float rsum[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(fmaf(s[j][0], sl2, -msc[0])); s[j][1] = exp2f(fmaf(s[j][1], sl2, -msc[0]));
      s[j][2] = exp2f(fmaf(s[j][2], sl2, -msc[1])); s[j][3] = exp2f(fmaf(s[j][3], sl2, -msc[1]));
      rsum[0] += s[j][0] + s[j][1];
      rsum[1] += s[j][2] + s[j][3];
    }
    l_i[0] = l_i[0] * alpha[0] + rsum[0];
    l_i[1] = l_i[1] * alpha[1] + rsum[1];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] *= alpha[0]; o[i][1] *= alpha[0]; o[i][2] *= alpha[1]; o[i][3] *= alpha[1]; }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]); a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]); a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        uint32_t bb[4];
        load_b_kn<HD>(bb, tV, kk * 16, dp * 16, lane);
        mma16816(o[2 * dp], a, bb[0], bb[1]);
        mma16816(o[2 * dp + 1], a, bb[2], bb[3]);
      }
    }
    __syncthreads();   // all warps done with stage st before it is refilled
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
  }
  const int row0 = q_base + warp * 16 + (lane >> 2);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + r * 8;
    if (row < S) {
      const float inv = 1.f / l_i[r];
      bf16* orow = out + (static_cast<long long>(b) * S + row) * D + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8) = pack_bf16x2(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
      if ((lane & 3) == 0) lse[(static_cast<long long>(b) * H + h) * S + row] = m_i[r] * scale + logf(l_i[r]);
    }
  }
}

// ------------------------------------------------------------------ backward: delta = rowsum(dO * O)
// Eight lanes share one (token, head) slice (one 16-byte load of O and of dO each: a warp reads 4 consecutive heads =
// 384 / 512 contiguous bytes per array), reduce with three shuffles, and the CTA transposes its 32 tokens x H heads of
// results through shared memory so that delta [B, H, S] is written in 128-byte rows.  (One thread per slice walked its
// 96 bytes alone and scattered 4-byte writes S floats apart: 0.46 of HBM bandwidth.)
constexpr int DELTA_TOK = 32;
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                  int S, int H, int HD, long long total) {
  pdl_wait();                  // launched with programmatic stream serialization (hct_common.cuh)
  pdl_launch_dependents();
  extern __shared__ float sdelta[];                       // [H][DELTA_TOK + 1]
  const int b = blockIdx.y;
  const int s0 = blockIdx.x * DELTA_TOK;
  const int ntok = min(DELTA_TOK, S - s0);
  const int c = threadIdx.x & 7, grp = threadIdx.x >> 3;  // 16-byte chunk of the head slice, (token, head) group
  const bool live = c * 8 < HD;
  const int pairs = ntok * H;
  const int D = H * HD;
  const long long base = (static_cast<long long>(b) * S + s0) * D;
  for (int p0 = 0; p0 < pairs; p0 += 32) {                // warp-uniform trip count (shuffles below)
    const int p = p0 + grp;
    float acc = 0.f;
    if (p < pairs && live) {
      const long long off = base + static_cast<long long>(p) * HD + c * 8;      // (token, head) slices are contiguous
      const uint4 a = *reinterpret_cast<const uint4*>(o + off), g = *reinterpret_cast<const uint4*>(dout + off);
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z), g3 = unpack_bf16x2(g.w);
      acc = (a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y) + (a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (c == 0 && p < pairs) sdelta[(p % H) * (DELTA_TOK + 1) + p / H] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * DELTA_TOK; i += blockDim.x) {
    const int h = i / DELTA_TOK, t = i - h * DELTA_TOK;
    if (t < ntok) delta[(static_cast<long long>(b) * H + h) * S + s0 + t] = sdelta[h * (DELTA_TOK + 1) + t];
  }
}

// ------------------------------------------------------------------ backward: dK, dV  (one CTA per key block)
template <int HD>
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dkdv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                     const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int TP = BLK * Tile<HD>::PITCH;
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);
  bf16* sV = sK + TP;
  bf16* sQ = sV + TP;        // 2 stages
  bf16* sdO = sQ + 2 * TP;   // 2 stages
  float* sLse = reinterpret_cast<float*>(sdO + 2 * TP);   // [2][64]
  float* sDelta = sLse + 2 * BLK;                          // [2][64]
  const int kvb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * HD;
  const long long rs = 3LL * D;
  const bf16* qg = qkv + static_cast<long long>(b) * S * rs + h * HD;
  const bf16* kg = qg + D;
  const bf16* vg = qg + 2 * D;
  const bf16* dog = dout + static_cast<long long>(b) * S * D + h * HD;
  const float* lse_g = lse + (static_cast<long long>(b) * H + h) * S;
  const float* delta_g = delta + (static_cast<long long>(b) * H + h) * S;
  const int nqb = (S + BLK - 1) / BLK;
  const float sl2 = scale * LOG2E;

  auto load_stats = [&](int stage, int qb) {
    if (threadIdx.x < BLK) {
      const int r = qb * BLK + threadIdx.x;
      sLse[stage * BLK + threadIdx.x] = r < S ? lse_g[r] * LOG2E : 0.f;
      sDelta[stage * BLK + threadIdx.x] = r < S ? delta_g[r] : 0.f;
    }
  };

  load_tile<HD>(sK, kg, rs, kvb * BLK, S);
  load_tile<HD>(sV, vg, rs, kvb * BLK, S);
  load_tile<HD>(sQ, qg, rs, 0, S);
  load_tile<HD>(sdO, dog, D, 0, S);
  cp_async_commit();
  load_stats(0, 0);

  uint32_t kf[HD / 16][4], vf[HD / 16][4];
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  const int kvrow0 = kvb * BLK + warp * 16 + (lane >> 2);

  for (int qb = 0; qb < nqb; ++qb) {
    const int st = qb & 1;
    if (qb + 1 < nqb) {
      load_tile<HD>(sQ + (st ^ 1) * TP, qg, rs, (qb + 1) * BLK, S);
      load_tile<HD>(sdO + (st ^ 1) * TP, dog, D, (qb + 1) * BLK, S);
      cp_async_commit();
      load_stats(st ^ 1, qb + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qb == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        load_a<HD>(kf[kk], sK, warp * 16, kk * 16, lane);
        load_a<HD>(vf[kk], sV, warp * 16, kk * 16, lane);
      }
    }
    const bf16* tQ = sQ + st * TP;
    const bf16* tdO = sdO + st * TP;
    const float* tL = sLse + st * BLK;
    const float* tD = sDelta + st * BLK;

    float pt[8][4], dpt[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      pt[j][0] = pt[j][1] = pt[j][2] = pt[j][3] = 0.f;
      dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bq[4], bo[4];
        load_b_nk<HD>(bq, tQ, np * 16, kk * 16, lane);
        mma16816(pt[2 * np], kf[kk], bq[0], bq[1]);
        mma16816(pt[2 * np + 1], kf[kk], bq[2], bq[3]);
        load_b_nk<HD>(bo, tdO, np * 16, kk * 16, lane);
        mma16816(dpt[2 * np], vf[kk], bo[0], bo[1]);
        mma16816(dpt[2 * np + 1], vf[kk], bo[2], bo[3]);
      }
    }
    // P^T = exp(S^T*scale - lse[q]); dS^T = P^T * (dP^T - delta[q]); zero outside the S x S problem
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ql = j * 8 + (lane & 3) * 2;          // local q column
      const int qgidx = qb * BLK + ql;
      const float l0 = tL[ql], l1 = tL[ql + 1], d0 = tD[ql], d1 = tD[ql + 1];
      const bool c0 = qgidx < S, c1 = qgidx + 1 < S;
      const bool r0 = kvrow0 < S, r1 = kvrow0 + 8 < S;
      const float p00 = (c0 && r0) ? exp2f(fmaf(pt[j][0], sl2, -l0)) : 0.f;
      const float p01 = (c1 && r0) ? exp2f(fmaf(pt[j][1], sl2, -l1)) : 0.f;
      const float p10 = (c0 && r1) ? exp2f(fmaf(pt[j][2], sl2, -l0)) : 0.f;
      const float p11 = (c1 && r1) ? exp2f(fmaf(pt[j][3], sl2, -l1)) : 0.f;
      pt[j][0] = p00; pt[j][1] = p01; pt[j][2] = p10; pt[j][3] = p11;
      dpt[j][0] = p00 * (dpt[j][0] - d0); dpt[j][1] = p01 * (dpt[j][1] - d1);
      dpt[j][2] = p10 * (dpt[j][2] - d0); dpt[j][3] = p11 * (dpt[j][3] - d1);
    }
    // dV += P^T dO ; dK += dS^T Q      (k index = q rows of the tile, B stored [k][n] -> transposing loads)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ap[4], ad[4];
      ap[0] = pack_bf16x2(pt[2 * kk][0], pt[2 * kk][1]); ap[1] = pack_bf16x2(pt[2 * kk][2], pt[2 * kk][3]);
      ap[2] = pack_bf16x2(pt[2 * kk + 1][0], pt[2 * kk + 1][1]); ap[3] = pack_bf16x2(pt[2 * kk + 1][2], pt[2 * kk + 1][3]);
      ad[0] = pack_bf16x2(dpt[2 * kk][0], dpt[2 * kk][1]); ad[1] = pack_bf16x2(dpt[2 * kk][2], dpt[2 * kk][3]);
      ad[2] = pack_bf16x2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]); ad[3] = pack_bf16x2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        uint32_t bo[4], bq[4];
        load_b_kn<HD>(bo, tdO, kk * 16, dp * 16, lane);
        mma16816(dv[2 * dp], ap, bo[0], bo[1]);
        mma16816(dv[2 * dp + 1], ap, bo[2], bo[3]);
        load_b_kn<HD>(bq, tQ, kk * 16, dp * 16, lane);
        mma16816(dk[2 * dp], ad, bq[0], bq[1]);
        mma16816(dk[2 * dp + 1], ad, bq[2], bq[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = kvrow0 + r * 8;
    if (row < S) {
      bf16* dkrow = dqkv + (static_cast<long long>(b) * S + row) * rs + D + h * HD + (lane & 3) * 2;
      bf16* dvrow = dkrow + D;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        *reinterpret_cast<uint32_t*>(dkrow + i * 8) = pack_bf16x2(dk[i][2 * r] * scale, dk[i][2 * r + 1] * scale);
        *reinterpret_cast<uint32_t*>(dvrow + i * 8) = pack_bf16x2(dv[i][2 * r], dv[i][2 * r + 1]);
      }
    }
  }
}

// ------------------------------------------------------------------ backward: dQ  (one CTA per query block)
template <int HD>
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                   const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int TP = BLK * Tile<HD>::PITCH;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + TP;
  bf16* sK = sdO + TP;      // 2 stages
  bf16* sV = sK + 2 * TP;   // 2 stages
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * HD;
  const long long rs = 3LL * D;
  const bf16* qg = qkv + static_cast<long long>(b) * S * rs + h * HD;
  const bf16* kg = qg + D;
  const bf16* vg = qg + 2 * D;
  const bf16* dog = dout + static_cast<long long>(b) * S * D + h * HD;
  const int nkb = (S + BLK - 1) / BLK;
  const float sl2 = scale * LOG2E;

  load_tile<HD>(sQ, qg, rs, qb * BLK, S);
  load_tile<HD>(sdO, dog, D, qb * BLK, S);
  load_tile<HD>(sK, kg, rs, 0, S);
  load_tile<HD>(sV, vg, rs, 0, S);
  cp_async_commit();

  const int qrow0 = qb * BLK + warp * 16 + (lane >> 2);
  float lse_r[2], delta_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = qrow0 + r * 8;
    const long long idx = (static_cast<long long>(b) * H + h) * S + row;
    lse_r[r] = row < S ? lse[idx] * LOG2E : 0.f;
    delta_r[r] = row < S ? delta[idx] : 0.f;
  }
  uint32_t qf[HD / 16][4], dof[HD / 16][4];
  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

  for (int kb = 0; kb < nkb; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < nkb) {
      load_tile<HD>(sK + (st ^ 1) * TP, kg, rs, (kb + 1) * BLK, S);
      load_tile<HD>(sV + (st ^ 1) * TP, vg, rs, (kb + 1) * BLK, S);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        load_a<HD>(qf[kk], sQ, warp * 16, kk * 16, lane);
        load_a<HD>(dof[kk], sdO, warp * 16, kk * 16, lane);
      }
    }
    const bf16* tK = sK + st * TP;
    const bf16* tV = sV + st * TP;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bk[4], bv[4];
        load_b_nk<HD>(bk, tK, np * 16, kk * 16, lane);
        mma16816(s[2 * np], qf[kk], bk[0], bk[1]);
        mma16816(s[2 * np + 1], qf[kk], bk[2], bk[3]);
        load_b_nk<HD>(bv, tV, np * 16, kk * 16, lane);
        mma16816(dp[2 * np], dof[kk], bv[0], bv[1]);
        mma16816(dp[2 * np + 1], dof[kk], bv[2], bv[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = kb * BLK + j * 8 + (lane & 3) * 2;
      const bool c0 = c < S, c1 = c + 1 < S;
      const float p00 = c0 ? exp2f(fmaf(s[j][0], sl2, -lse_r[0])) : 0.f;
      const float p01 = c1 ? exp2f(fmaf(s[j][1], sl2, -lse_r[0])) : 0.f;
      const float p10 = c0 ? exp2f(fmaf(s[j][2], sl2, -lse_r[1])) : 0.f;
      const float p11 = c1 ? exp2f(fmaf(s[j][3], sl2, -lse_r[1])) : 0.f;
      dp[j][0] = p00 * (dp[j][0] - delta_r[0]); dp[j][1] = p01 * (dp[j][1] - delta_r[0]);
      dp[j][2] = p10 * (dp[j][2] - delta_r[1]); dp[j][3] = p11 * (dp[j][3] - delta_r[1]);
    }
    // dQ += dS K    (k index = key rows of the tile; K stored [k][n] -> transposing loads)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]); a[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
      a[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]); a[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int dd = 0; dd < HD / 16; ++dd) {
        uint32_t bk[4];
        load_b_kn<HD>(bk, tK, kk * 16, dd * 16, lane);
        mma16816(dq[2 * dd], a, bk[0], bk[1]);
        mma16816(dq[2 * dd + 1], a, bk[2], bk[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = qrow0 + r * 8;
    if (row < S) {
      bf16* dqrow = dqkv + (static_cast<long long>(b) * S + row) * rs + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
        *reinterpret_cast<uint32_t*>(dqrow + i * 8) = pack_bf16x2(dq[i][2 * r] * scale, dq[i][2 * r + 1] * scale);
    }
  }
}

template <typename K>
int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { hct_set_error("cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
  return HCT_OK;
}

template <int HD>
int launch_fwd(const bf16* qkv, bf16* out, float* lse, int B, int S, int H, int q_start, cudaStream_t st) {
  const int smem = 5 * Tile<HD>::BYTES;
  static bool configured = false;
  if (!configured) { int rc = set_smem(attn_fwd_kernel<HD>, smem); if (rc) return rc; configured = true; }
  dim3 grid((S - q_start + BLK - 1) / BLK, H, B);
  attn_fwd_kernel<HD><<<grid, NTHREADS, smem, st>>>(qkv, out, lse, S, H, 1.0f / sqrtf(static_cast<float>(HD)), q_start);
  return hct_check_launch("attn_fwd_kernel");
}

template <int HD>
int launch_bwd(const bf16* qkv, const bf16* dout, const float* lse, const float* delta, bf16* dqkv, int B, int S, int H,
               cudaStream_t st) {
  const int smem_kv = 6 * Tile<HD>::BYTES + 4 * BLK * static_cast<int>(sizeof(float));
  const int smem_q = 6 * Tile<HD>::BYTES;
  static bool configured = false;
  if (!configured) {
    int rc = set_smem(attn_bwd_dkdv_kernel<HD>, smem_kv); if (rc) return rc;
    rc = set_smem(attn_bwd_dq_kernel<HD>, smem_q); if (rc) return rc;
    configured = true;
  }
  dim3 grid((S + BLK - 1) / BLK, H, B);
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  attn_bwd_dkdv_kernel<HD><<<grid, NTHREADS, smem_kv, st>>>(qkv, dout, lse, delta, dqkv, S, H, scale);
  int rc = hct_check_launch("attn_bwd_dkdv_kernel");
  if (rc) return rc;
  attn_bwd_dq_kernel<HD><<<grid, NTHREADS, smem_q, st>>>(qkv, dout, lse, delta, dqkv, S, H, scale);
  return hct_check_launch("attn_bwd_dq_kernel");
}

}  // namespace

// tcgen05 kernels (hct_attention_sm100.cu)
int hct_attn_tc_tiles(int S, int tail_on_mma_sync);
int hct_attention_fwd_tc(const void* qkv, void* out, float* lse, int B, int S, int H, int hd, int n_tiles, cudaStream_t st);
int hct_attention_bwd_tc(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, int B, int S,
                         int H, int hd, int n_tiles, cudaStream_t st);
// row kernel for the one row behind the last full tile in the backward (S = 128 k + 1; hct_attention_tail.cu)
int hct_attention_bwd3(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, float* colsum, int B, int S, int H,
                       int hd, int n_tiles, cudaStream_t st);
bool hct_attention_bwd_tail_supported(int S, int hd, int r0);
int hct_attention_fwd2(const void* qkv, void* out, float* lse, int B, int S, int H, int hd, int n_tiles, cudaStream_t st);
static int g_attn_fwd2 = 0;      // 1: forward on the pipelined persistent kernel of hct_attention_fwd2.cu
extern "C" int hct_attention_set_fwd2(int enable) { g_attn_fwd2 = enable != 0; return HCT_OK; }
int hct_attention_bwd_tail(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, float* colsum, int B, int S,
                           int H, int hd, int r0, cudaStream_t st);
// 0: mma.sync kernels only; 1: tcgen05, forward tail rows (S % 128 <= 32) on mma.sync; 2 (default): tcgen05 for every
// forward tile, a single backward tail row (S % 128 == 1) on the row kernel; 3: tcgen05 for every tile, backward included
static int g_attn_tc = 2;
// 1: pipelined persistent backward (hct_attention_bwd3.cu: three score buffers, two softmax groups per SM);
// 0: the two-CTA-per-SM kernels of hct_attention_sm100.cu
static int g_attn_bwd3 = 1;
extern "C" int hct_attention_set_bwd3(int enable) { g_attn_bwd3 = enable != 0; return HCT_OK; }
extern "C" int hct_attention_set_tcgen05(int mode) { g_attn_tc = mode < 0 ? 0 : (mode > 3 ? 3 : mode); return HCT_OK; }

extern "C" int hct_attention_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t S, int32_t H, int32_t hd,
                                 hct_stream_t s) {
  HCT_REQUIRE(B > 0 && S > 0 && H > 0 && H <= 65535 && B <= 65535, "attention_fwd: bad B=%d S=%d H=%d", B, S, H);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  HctProfScope prof(st, HCT_PROF_ATTN_FWD, 4.0 * B * static_cast<double>(S) * S * H * hd);
  const bf16* q = static_cast<const bf16*>(qkv);
  bf16* o = static_cast<bf16*>(out);
  int q_start = 0;
  if (g_attn_tc && (hd == 64 || hd == 48)) {
    // full 128-row query tiles on tcgen05; the few rows behind them (the cls token makes S = 128 k + 1) on mma.sync
    const int n_tiles = hct_attn_tc_tiles(S, g_attn_tc == 1);   // modes 2 and 3: every forward tile on tcgen05
    int rc = g_attn_fwd2 ? hct_attention_fwd2(qkv, out, lse, B, S, H, hd, n_tiles, st)
                         : hct_attention_fwd_tc(qkv, out, lse, B, S, H, hd, n_tiles, st);
    if (rc != HCT_OK) return rc;
    q_start = n_tiles * 128;
    if (q_start >= S) return HCT_OK;
  }
  switch (hd) {
    case 64: return launch_fwd<64>(q, o, lse, B, S, H, q_start, st);
    case 48: return launch_fwd<48>(q, o, lse, B, S, H, q_start, st);
    case 32: return launch_fwd<32>(q, o, lse, B, S, H, q_start, st);
    default: hct_set_error("attention_fwd: head dim %d unsupported (32/48/64)", hd); return HCT_ERR_UNSUPPORTED;
  }
}

// hct_attention_bwd plus the qkv-bias gradient: the column sums of dqkv over all B * S tokens are ADDED to dqkv_colsum
// (fp32 [3 * H * hd], may be NULL).  The pipelined tcgen05 path sums them from the tiles it has staged for its stores; every
// other path runs the column-sum kernel over the finished dqkv.
extern "C" int hct_colsum(const void* x, int32_t x_bf16, int64_t ld, float* out, int64_t rows, int32_t cols, hct_stream_t s);
extern "C" int hct_attention_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                      float* delta_ws, float* dqkv_colsum, int32_t B, int32_t S, int32_t H, int32_t hd,
                                      hct_stream_t s) {
  HCT_REQUIRE(B > 0 && S > 0 && H > 0 && H <= 65535 && B <= 65535, "attention_bwd: bad B=%d S=%d H=%d", B, S, H);
  HCT_REQUIRE(hd == 64 || hd == 48 || hd == 32, "attention_bwd: head dim %d unsupported (32/48/64)", hd);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const long long total = static_cast<long long>(B) * S * H;
  int rc;
  bool fused = false;
  {
    HctProfScope prof(st, HCT_PROF_ATTN_BWD, 8.0 * B * static_cast<double>(S) * S * H * hd);   // 2 x forward
    HCT_REQUIRE(hd % 8 == 0 && hd <= 64, "attention_bwd: head dim %d unsupported by the delta pre-pass", hd);
    hct_launch_pdl(attn_delta_kernel, dim3((S + DELTA_TOK - 1) / DELTA_TOK, B), dim3(256), H * (DELTA_TOK + 1) * sizeof(float), st,
                   static_cast<const bf16*>(out), static_cast<const bf16*>(dout), delta_ws, S, H, hd, total);
    rc = hct_check_launch("attn_delta_kernel");
    if (rc) return rc;
    if (g_attn_tc && (hd == 64 || hd == 48)) {
      const int full = S / 128, r0 = full * 128;
      fused = g_attn_bwd3 != 0;                          // the pipelined kernels (and the row kernel behind them) sum as they store
      float* cs = fused ? dqkv_colsum : nullptr;
      auto tc_bwd = [&](int n_tiles) {
        return g_attn_bwd3 ? hct_attention_bwd3(qkv, dout, lse, delta_ws, dqkv, cs, B, S, H, hd, n_tiles, st)
                           : hct_attention_bwd_tc(qkv, dout, lse, delta_ws, dqkv, B, S, H, hd, n_tiles, st);
      };
      if (g_attn_tc != 3 && r0 < S && hct_attention_bwd_tail_supported(S, hd, r0)) {
        if (full > 0) {
          rc = tc_bwd(full);
          if (rc) return rc;
        }
        rc = hct_attention_bwd_tail(qkv, dout, lse, delta_ws, dqkv, cs, B, S, H, hd, r0, st);
      } else {
        rc = tc_bwd((S + 127) / 128);
      }
    } else {
      const bf16* q = static_cast<const bf16*>(qkv);
      const bf16* d = static_cast<const bf16*>(dout);
      bf16* dq = static_cast<bf16*>(dqkv);
      switch (hd) {
        case 64: rc = launch_bwd<64>(q, d, lse, delta_ws, dq, B, S, H, st); break;
        case 48: rc = launch_bwd<48>(q, d, lse, delta_ws, dq, B, S, H, st); break;
        default: rc = launch_bwd<32>(q, d, lse, delta_ws, dq, B, S, H, st); break;
      }
    }
  }
  if (rc) return rc;
  if (dqkv_colsum != nullptr && !fused)
    return hct_colsum(dqkv, 1, 3LL * H * hd, dqkv_colsum, static_cast<long long>(B) * S, 3 * H * hd, s);
  return HCT_OK;
}

extern "C" int hct_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                 float* delta_ws, int32_t B, int32_t S, int32_t H, int32_t hd, hct_stream_t s) {
  return hct_attention_bwd_bias(qkv, out, dout, lse, dqkv, delta_ws, nullptr, B, S, H, hd, s);
}
