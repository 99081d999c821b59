// Attention backward for the row behind the last full 128-row tile (S = 128 k + 1: the cls token makes S = 129 / 513).  The tcgen05 backward kernels give every 128-row tile of keys (dK/dV kernel) and
// of queries (dQ kernel) its own CTA, and a CTA's run time is set by the mbarrier-chained MMA -> softmax -> MMA loop
// over ALL blocks of the other dimension, not by how many of its 128 rows are real: the one-row tail tile cost as much
// as a full one (+22 % at S = 513, +50-68 % at S = 129, tools/attn_tail_probe.py).  Those r rows need r x S scores only,
// so they are done here on the CUDA cores: one CTA per (head, sample), fp32 throughout, the K / V (part A) and Q / dO
// (part B) head slices each read exactly once.  The full tiles still sweep over all S rows of the other dimension, so nothing else changes.
//   part A (tail QUERIES x all keys):   dQ_r = scale * sum_k dS[r,k] K_k
//   part B (tail KEYS x all queries):   dV_r = sum_q P[q,r] dO_q,   dK_r = scale * sum_q dS[q,r] Q_q
//   with P = exp(scale * q.k - lse_q), dS = P * (dO_q.v_k - delta_q)          (F.scaled_dot_product_attention backward)
#include "../../include/hct_b200.h"
#include "hct_common.cuh"

namespace {

constexpr int TAIL_THREADS = 256;
constexpr int TAIL_GROUPS = TAIL_THREADS / 8;          // 8 lanes share a row (16 bytes each); 32 rows in flight per CTA
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void unpack8(const uint4 u, float* __restrict__ f) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ float dot8f(const float* __restrict__ a, const float* __restrict__ b) {
  return ((a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3])) + ((a[4] * b[4] + a[5] * b[5]) + (a[6] * b[6] + a[7] * b[7]));
}
__device__ __forceinline__ float group8_sum(float v) {      // sum over the 8 lanes that share a row
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// Eight lanes own one row of the streamed matrix pair (lane c holds elements [8c, 8c+8) of the head slice, one 16-byte
// load each: a row is one contiguous 96 / 128-byte segment), reduce their partial dot products with three shuffles, and
// accumulate the rank-1 update of the output rows straight from the registers that hold the row -- every matrix is
// read exactly once and no score ever touches shared memory.
template <int HD, int RMAX>
__global__ void __launch_bounds__(TAIL_THREADS)
attn_bwd_tail_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                     const float* __restrict__ delta, bf16* __restrict__ dqkv, float* __restrict__ colsum, int S, int H, int r0, int R,
                     float scale) {
  pdl_wait();                  // launched with programmatic stream serialization (hct_common.cuh)
  pdl_launch_dependents();
  __shared__ float red[TAIL_GROUPS][HD + 1];
  __shared__ float sl[RMAX], sdl[RMAX];                 // rows [r0, r0 + R), 1 <= R <= RMAX
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int c = tid & 7, grp = tid >> 3;                // 16-byte chunk of the head slice, row group
  const bool live = c * 8 < HD;                         // HD = 48: lanes 6, 7 of each group only take part in the shuffles
  const int D = H * HD;
  const long long rs = 3LL * D;
  const bf16* qg = qkv + static_cast<long long>(b) * S * rs + h * HD + c * 8;
  const bf16* dog = dout + static_cast<long long>(b) * S * D + h * HD + c * 8;
  const float* lse_g = lse + (static_cast<long long>(b) * H + h) * S;
  const float* delta_g = delta + (static_cast<long long>(b) * H + h) * S;
  const float sl2 = scale * LOG2E;
  if (tid < R) { sl[tid] = lse_g[r0 + tid] * LOG2E; sdl[tid] = delta_g[r0 + tid]; }

  // this lane's chunk of the tail rows: q, k, v, dO
  float tq[RMAX][8], tk[RMAX][8], tv[RMAX][8], tdo[RMAX][8];
#pragma unroll
  for (int r = 0; r < RMAX; ++r) {
    uint4 z = make_uint4(0, 0, 0, 0), uq = z, uk = z, uv = z, uo = z;
    if (r < R && live) {
      const bf16* row = qg + static_cast<long long>(r0 + r) * rs;
      uq = *reinterpret_cast<const uint4*>(row); uk = *reinterpret_cast<const uint4*>(row + D);
      uv = *reinterpret_cast<const uint4*>(row + 2 * D);
      uo = *reinterpret_cast<const uint4*>(dog + static_cast<long long>(r0 + r) * D);
    }
    unpack8(uq, tq[r]); unpack8(uk, tk[r]); unpack8(uv, tv[r]); unpack8(uo, tdo[r]);
  }
  __syncthreads();

  auto reduce_store = [&](const float* acc8, bf16* dst, float mul, int col) {     // sum over row groups, write one output row
    __syncthreads();
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[grp][c * 8 + j] = acc8[j];
    }
    __syncthreads();
    if (tid < HD) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < TAIL_GROUPS; ++g) t += red[g][tid];
      const bf16 v = __float2bfloat16_rn(t * mul);
      dst[tid] = v;
      if (colsum != nullptr) atomicAdd(colsum + col + tid, __bfloat162float(v));   // qkv-bias gradient: this row's share
    }
  };

  // ---------------- part A: the tail queries against every key  ->  dQ rows
  {
    float acc[RMAX][8];
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
#pragma unroll 2
    for (int k0 = 0; k0 < S; k0 += TAIL_GROUPS) {         // warp-uniform trip count: the shuffles below need all 32 lanes
      const int k = k0 + grp;
      const bool valid = k < S;
      uint4 uk = make_uint4(0, 0, 0, 0), uv = uk;
      if (live && valid) {
        const bf16* krow = qg + static_cast<long long>(k) * rs + D;
        uk = *reinterpret_cast<const uint4*>(krow); uv = *reinterpret_cast<const uint4*>(krow + D);
      }
      float kf[8], vf[8];
      unpack8(uk, kf); unpack8(uv, vf);
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R) {
          const float sc = group8_sum(dot8f(kf, tq[r])), dp = group8_sum(dot8f(vf, tdo[r]));
          const float ds = valid ? exp2f(fmaf(sc, sl2, -sl[r])) * (dp - sdl[r]) : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(ds, kf[j], acc[r][j]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
      if (r < R) reduce_store(acc[r], dqkv + (static_cast<long long>(b) * S + r0 + r) * rs + h * HD, scale, h * HD);
  }

  // ---------------- part B: every query against the tail keys  ->  dK, dV rows
  {
    float av[RMAX][8], ak[RMAX][8];
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) { av[r][j] = 0.f; ak[r][j] = 0.f; }
#pragma unroll 2
    for (int q0 = 0; q0 < S; q0 += TAIL_GROUPS) {
      const int q = q0 + grp;
      const bool valid = q < S;
      uint4 uq = make_uint4(0, 0, 0, 0), uo = uq;
      if (live && valid) {
        uq = *reinterpret_cast<const uint4*>(qg + static_cast<long long>(q) * rs);
        uo = *reinterpret_cast<const uint4*>(dog + static_cast<long long>(q) * D);
      }
      const float lq = valid ? lse_g[q] * LOG2E : 0.f, dq = valid ? delta_g[q] : 0.f;
      float qf[8], of[8];
      unpack8(uq, qf); unpack8(uo, of);
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R) {
          const float sc = group8_sum(dot8f(qf, tk[r])), dp = group8_sum(dot8f(of, tv[r]));
          const float p = valid ? exp2f(fmaf(sc, sl2, -lq)) : 0.f, ds = p * (dp - dq);
#pragma unroll
          for (int j = 0; j < 8; ++j) { av[r][j] = fmaf(p, of[j], av[r][j]); ak[r][j] = fmaf(ds, qf[j], ak[r][j]); }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R) {
        bf16* o = dqkv + (static_cast<long long>(b) * S + r0 + r) * rs + h * HD;
        reduce_store(ak[r], o + D, scale, D + h * HD);
        reduce_store(av[r], o + 2 * D, 1.0f, 2 * D + h * HD);
      }
    }
  }
}

template <int HD, int RMAX>
int launch_tail(const bf16* qkv, const bf16* dout, const float* lse, const float* delta, bf16* dqkv, float* colsum, int B, int S, int H,
                int r0, int R, cudaStream_t st) {
  hct_launch_pdl(attn_bwd_tail_kernel<HD, RMAX>, dim3(H, B), dim3(TAIL_THREADS), 0, st, qkv, dout, lse, delta, dqkv, colsum, S, H, r0, R,
                 1.0f / sqrtf(static_cast<float>(HD)));
  return hct_check_launch("attn_bwd_tail_kernel");
}

}  // namespace

// Row r0 = S - 1 of dQ, dK and dV (r0 = the number of rows covered by full 128-row tiles).  Measured on B200 at B = 256:
// one tail row costs 0.26 ms here against 0.28 ms as a fifth tcgen05 tile at S = 513 (16 x 48), 0.085 against 0.10 ms
// at S = 129 (12 x 64); with several tail rows (DINO's S = 517) the tcgen05 tile wins, so only R == 1 is routed here.
bool hct_attention_bwd_tail_supported(int S, int hd, int r0) {
  return (hd == 64 || hd == 48) && r0 > 0 && S - r0 == 1;
}

int hct_attention_bwd_tail(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, float* colsum, int B,
                           int S, int H, int hd, int r0, cudaStream_t st) {
  const bf16* q = static_cast<const bf16*>(qkv);
  const bf16* d = static_cast<const bf16*>(dout);
  bf16* o = static_cast<bf16*>(dqkv);
  if (hd == 64) return launch_tail<64, 1>(q, d, lse, delta, o, colsum, B, S, H, r0, S - r0, st);
  return launch_tail<48, 1>(q, d, lse, delta, o, colsum, B, S, H, r0, S - r0, st);
}
