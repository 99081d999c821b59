// HBM-bound row kernels: LayerNorm fwd/bwd, casts, column sums, token-row broadcast/reduce.
// Roofline for all of them is HBM bandwidth; they use 16-byte accesses and warp-per-row layouts.
#include "../../include/hct_b200.h"
#include "hct_common.cuh"
#include "hct_tcgen05.cuh"

namespace {

constexpr int LN_MAXV = 16;            // float4 per lane held in registers -> dim <= 2048
constexpr int LN_WARPS = 8;
constexpr int LN_SMEM_MAX = 226 * 1024;   // dynamic shared memory budget: the 227 KiB a CTA may opt in to on sm_100 minus the static part

// ------------------------------------------------------------------ LayerNorm forward
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              void* __restrict__ y, int y_bf16, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              long long rows, int dim, float eps) {
  pdl_wait();                  // launched with programmatic stream serialization (hct_common.cuh)
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = dim >> 2;   // float4 per row
  const bool rms = beta == nullptr;
  for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * LN_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) { v[i] = xr[c]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    }
    const float mean = rms ? 0.f : warp_sum(s) / dim;     // RMSNorm (layers.py:29-40): no centring
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / dim + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = rms ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(beta) + c);
        const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
        if (y_bf16) {
          uint2 u; u.x = pack_bf16x2(o0, o1); u.y = pack_bf16x2(o2, o3);
          reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + row * dim)[c] = u;
        } else {
          reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * dim)[c] = make_float4(o0, o1, o2, o3);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// One warp per row, rows software-pipelined through shared memory with cp.async (next row's x / dy / dres land
// while the current row is reduced), so the kernel keeps ~8 KB per warp in flight without register cost.
// dgamma, dbeta and (optionally) the column sums of the bf16-rounded dx output -- the bias gradient of the
// Linear that consumes dx -- are accumulated in registers and flushed once per CTA.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc) : "memory");
}

// 1-D bulk copy global -> shared through the TMA unit, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(hct_tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(hct_tc::smem_u32(bar)) : "memory");
}

// USE_BULK: each row's x / dres / dy land through three cp.async.bulk copies issued by one lane (TMA unit, mbarrier
// completion) instead of 18 per-lane cp.async instructions.
// LN_BWD_WARPS: 12 warps x 2 stages where that fits (160 registers per thread cap an SM at 12 warps): the kernel is bound
// by how much latency its few resident warps can hide, not by bytes in flight -- 8 warps x 3 stages ran at 0.73 of copy
// bandwidth, 12 x 2 (the same bytes in flight) at 0.89.
template <int NV, bool DY_BF16, int LN_BWD_STAGES, bool USE_BULK, int LN_BWD_WARPS>
__global__ void __launch_bounds__(LN_BWD_WARPS * 32)
ln_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres_in,
              float* __restrict__ dx_f32, bf16* __restrict__ dx_bf16, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dxsum, long long rows, int dim) {
  pdl_wait();                  // launched with programmatic stream serialization (hct_common.cuh)
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sbuf[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = dim >> 2;
  // per warp, per stage: x [dim] f32 | dres [dim] f32 | dy [dim] (f32, or bf16 in half the space)
  // LN_BWD_STAGES-deep cp.async ring per warp: two rows (15 KiB at dim 768) in flight per warp, ~120 KiB per SM --
  // a 2-deep ring (60 KiB per SM) sat right at the bandwidth-delay product of HBM under load.
  const int stage_floats = DY_BF16 ? 2 * dim + dim / 2 : 3 * dim;
  float* wbuf = sbuf + static_cast<size_t>(warp) * LN_BWD_STAGES * stage_floats;
  __shared__ uint64_t bars[LN_BWD_WARPS][4];
  if (USE_BULK) {
    if (lane == 0) {
      for (int k = 0; k < LN_BWD_STAGES; ++k) hct_tc::mbar_init(&bars[warp][k], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  float4 dg[NV], db[NV], ds[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { dg[i] = make_float4(0, 0, 0, 0); db[i] = dg[i]; ds[i] = dg[i]; }

  const long long row0 = static_cast<long long>(blockIdx.x) * LN_BWD_WARPS + warp;
  const long long rstride = static_cast<long long>(gridDim.x) * LN_BWD_WARPS;
  auto prefetch = [&](long long row, int stage) {          // always commits a group (possibly empty): uniform counting
    if (USE_BULK) {
      if (row < rows && lane == 0) {
        float* sx = wbuf + stage * stage_floats;
        const uint32_t bx = dim * 4u, bd = DY_BF16 ? dim * 2u : dim * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warp's generic reads of this stage precede the refill
        hct_tc::mbar_expect_tx(&bars[warp][stage], bx + (dres_in ? bx : 0u) + bd);
        bulk_load_1d(sx, x + row * dim, bx, &bars[warp][stage]);
        if (dres_in) bulk_load_1d(sx + dim, dres_in + row * dim, bx, &bars[warp][stage]);
        bulk_load_1d(sx + 2 * dim, DY_BF16 ? static_cast<const void*>(reinterpret_cast<const bf16*>(dy) + row * dim)
                                           : static_cast<const void*>(reinterpret_cast<const float*>(dy) + row * dim),
                     bd, &bars[warp][stage]);
      }
      return;
    }
    if (row < rows) {
      float* sx = wbuf + stage * stage_floats;
      float* sr = sx + dim;
      float* sd = sr + dim;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = i * 32 + lane;
        if (c < nv) {
          cp_async_16(sx + 4 * c, x + row * dim + 4 * c);
          if (dres_in) cp_async_16(sr + 4 * c, dres_in + row * dim + 4 * c);
          if (DY_BF16) cp_async_8(reinterpret_cast<bf16*>(sd) + 4 * c, reinterpret_cast<const bf16*>(dy) + row * dim + 4 * c);
          else cp_async_16(sd + 4 * c, reinterpret_cast<const float*>(dy) + row * dim + 4 * c);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

#pragma unroll
  for (int k = 0; k < LN_BWD_STAGES - 1; ++k) prefetch(row0 + k * rstride, k);
  int stage = 0;
  uint32_t it = 0;
  for (long long row = row0; row < rows; row += rstride, stage = (stage + 1 == LN_BWD_STAGES ? 0 : stage + 1), ++it) {
    prefetch(row + (LN_BWD_STAGES - 1) * rstride, (stage + LN_BWD_STAGES - 1) % LN_BWD_STAGES);
    if (USE_BULK) {
      hct_tc::mbar_wait(&bars[warp][stage], (it / LN_BWD_STAGES) & 1u);
    } else {
      asm volatile("cp.async.wait_group %0;" ::"n"(LN_BWD_STAGES - 1) : "memory");
    }
    __syncwarp();
    const float mu = mean ? mean[row] : 0.f, rs = rstd[row];      // mean == NULL: RMSNorm backward
    const float* sx = wbuf + stage * stage_floats;
    const float* sr = sx + dim;
    const float* sd = sr + dim;
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        const float4 xv = reinterpret_cast<const float4*>(sx)[c];
        float4 d;
        if (DY_BF16) {
          const uint2 u = reinterpret_cast<const uint2*>(sd)[c];
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
          d = make_float4(a.x, a.y, b.x, b.y);
        } else {
          d = reinterpret_cast<const float4*>(sd)[c];
        }
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      }
    }
    const float c1 = mean ? warp_sum(s1) / dim : 0.f, c2 = warp_sum(s2) / dim;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        float4 o = make_float4(rs * (g[i].x - c1 - xh[i].x * c2), rs * (g[i].y - c1 - xh[i].y * c2),
                               rs * (g[i].z - c1 - xh[i].z * c2), rs * (g[i].w - c1 - xh[i].w * c2));
        if (dres_in) {
          const float4 r = reinterpret_cast<const float4*>(sr)[c];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (dx_f32) reinterpret_cast<float4*>(dx_f32 + row * dim)[c] = o;
        if (dx_bf16) {
          uint2 u; u.x = pack_bf16x2(o.x, o.y); u.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(dx_bf16 + row * dim)[c] = u;
          if (dxsum) {
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
            ds[i].x += a.x; ds[i].y += a.y; ds[i].z += b.x; ds[i].w += b.y;
          }
        }
      }
    }
    __syncwarp();   // everyone done reading this stage before it is refilled by the next iteration's prefetch
  }
  // cross-warp reduction of the column accumulators (reuses the staging memory), one atomic per column per CTA
  __syncthreads();
  if (dgamma != nullptr || dxsum != nullptr) {
    float* my = sbuf + static_cast<size_t>(warp) * 3 * dim;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        reinterpret_cast<float4*>(my)[c] = dg[i];
        reinterpret_cast<float4*>(my + dim)[c] = db[i];
        reinterpret_cast<float4*>(my + 2 * dim)[c] = ds[i];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * dim; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < LN_BWD_WARPS; ++w) s += sbuf[static_cast<size_t>(w) * 3 * dim + c];
      if (c < dim) { if (dgamma) atomicAdd(dgamma + c, s); }
      else if (c < 2 * dim) { if (dbeta) atomicAdd(dbeta + (c - dim), s); }
      else if (dxsum) atomicAdd(dxsum + (c - 2 * dim), s);
    }
  }
}

// ------------------------------------------------------------------ casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n8 = n >> 3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
  for (long long i = (n8 << 3) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n8 = n >> 3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 u = reinterpret_cast<const uint4*>(src)[i];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
  }
  for (long long i = (n8 << 3) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __bfloat162float(src[i]);
}

// ------------------------------------------------------------------ column sum (bias gradients)
// block = 32 column-groups (8 columns each) x 8 row lanes; grid.y strides over rows.
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, int x_bf16, long long ld, float* __restrict__ out, long long rows,
              int cols) {
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < cols) {
    const long long rstep = static_cast<long long>(gridDim.y) * 8;
    long long r = static_cast<long long>(blockIdx.y) * 8 + rl;
    if (x_bf16) {
      // four independent 16-byte loads in flight per thread (one load per iteration left the kernel latency-bound)
      const bf16* xb = reinterpret_cast<const bf16*>(x) + col;
      for (; r + 3 * rstep < rows; r += 4 * rstep) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (r + k * rstep) * ld);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = unpack_bf16x2(u[k].x), b = unpack_bf16x2(u[k].y), c = unpack_bf16x2(u[k].z), d = unpack_bf16x2(u[k].w);
          acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
        }
      }
    }
    for (; r < rows; r += rstep) {
      if (x_bf16) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(x) + r * ld + col);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      } else {
        const float* p = reinterpret_cast<const float*>(x) + r * ld + col;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;   // 256 columns of this block
  const int gc = blockIdx.x * 256 + c;
  if (gc < cols) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(out + gc, s);
  }
}

__global__ void broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int batch, int nrows,
                                      long long dst_rows_per_batch, int row_off, int dim) {
  const int nv = dim >> 2;
  const long long total = static_cast<long long>(batch) * nrows * nv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const long long t = i / nv;
    const int r = static_cast<int>(t % nrows);
    const long long b = t / nrows;
    reinterpret_cast<float4*>(dst + (b * dst_rows_per_batch + row_off + r) * dim)[c] =
        __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(r) * dim) + c);
  }
}

// out[r, c] += sum_b src[b, row_off + r, c]; one thread per (r, c), loop over batch (coalesced over c)
__global__ void reduce_rows_kernel(const float* __restrict__ src, float* __restrict__ out, int batch, int nrows,
                                   long long src_rows_per_batch, int row_off, int dim) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(nrows) * dim) return;
  const int c = static_cast<int>(i % dim), r = static_cast<int>(i / dim);
  float s = 0.f;
  for (int b = 0; b < batch; ++b) s += src[(static_cast<long long>(b) * src_rows_per_batch + row_off + r) * dim + c];
  out[i] += s;
}

__global__ void copy_rows_f32_bf16_kernel(const float* __restrict__ src, long long src_ld, long long src_rows_per_group,
                                          int src_row_off, bf16* __restrict__ dst, long long dst_ld, long long groups,
                                          int rows_per_group, int dim) {
  const int nv = dim >> 2;
  const long long total = groups * rows_per_group * nv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const long long r = i / nv;
    const long long g = r / rows_per_group;
    const int rr = static_cast<int>(r % rows_per_group);
    const float4 v = reinterpret_cast<const float4*>(src + (g * src_rows_per_group + src_row_off + rr) * src_ld)[c];
    uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst + r * dst_ld)[c] = u;
  }
}

__global__ void scatter_add_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ idx, float* __restrict__ out,
                                        long long rows, int dim) {
  const int nv = dim >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const long long r = i / nv;
    const uint2 u = reinterpret_cast<const uint2*>(src + r * dim)[c];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    float* o = out + static_cast<long long>(idx[r]) * dim + 4 * c;
    atomicAdd(o, a.x); atomicAdd(o + 1, a.y); atomicAdd(o + 2, b.x); atomicAdd(o + 3, b.y);
  }
}

__global__ void gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pre, bf16* __restrict__ out, long long n) {
  const long long n8 = n >> 3;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 d = reinterpret_cast<const uint4*>(dy)[i], a = reinterpret_cast<const uint4*>(pre)[i];
    const uint32_t dd[4] = {d.x, d.y, d.z, d.w}, aa[4] = {a.x, a.y, a.z, a.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = unpack_bf16x2(dd[k]), x = unpack_bf16x2(aa[k]);
      o[k] = pack_bf16x2(g.x * gelu_erf_grad(x.x), g.y * gelu_erf_grad(x.y));
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

inline int grid_for(long long work_items, int threads, int max_blocks) {
  long long g = (work_items + threads - 1) / threads;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

extern "C" int hct_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_bf16,
                                 float* mean, float* rstd, int64_t rows, int32_t dim, float eps, hct_stream_t s) {
  HCT_REQUIRE(rows >= 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAXV * 128, "layernorm_fwd: dim=%d unsupported", dim);
  if (rows == 0) return HCT_OK;
  // algorithmic bytes: x fp32 in, y out, statistics out
  HctProfScope prof(static_cast<cudaStream_t>(s), HCT_PROF_LN_FWD,
                    static_cast<double>(rows) * (dim * (4.0 + (y_bf16 ? 2.0 : 4.0)) + (mean ? 4.0 : 0.0) + (rstd ? 4.0 : 0.0)));
  const int grid = grid_for(rows, LN_WARPS, hct_num_sms() * 8);
#define HCT_LN_FWD(NV) \
  hct_launch_pdl(ln_fwd_kernel<NV>, dim3(grid), dim3(LN_WARPS * 32), 0, static_cast<cudaStream_t>(s), x, gamma, beta, y, y_bf16, mean, rstd, rows, dim, eps)
  if (dim <= 256) HCT_LN_FWD(2); else if (dim <= 768) HCT_LN_FWD(6); else if (dim <= 1024) HCT_LN_FWD(8); else HCT_LN_FWD(16);
#undef HCT_LN_FWD
  return hct_check_launch("ln_fwd_kernel");
}

static int g_ln_bulk = 1;
extern "C" int hct_layernorm_set_bulk(int enable) { g_ln_bulk = enable != 0; return HCT_OK; }

extern "C" int hct_layernorm_bwd(const void* dy, int dy_bf16, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, const float* dres_in, float* dx_out_f32, void* dx_out_bf16,
                                 float* dgamma, float* dbeta, float* dxsum, int64_t rows, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(rows >= 0 && dim > 0 && dim % 4 == 0 && dim <= LN_MAXV * 128, "layernorm_bwd: dim=%d unsupported", dim);
  HCT_REQUIRE(dbeta == nullptr || dgamma != nullptr, "layernorm_bwd: dbeta without dgamma");
  HCT_REQUIRE(mean != nullptr || dbeta == nullptr, "layernorm_bwd: RMSNorm (mean == NULL) has no beta gradient");
  HCT_REQUIRE(dxsum == nullptr || dx_out_bf16 != nullptr, "layernorm_bwd: dxsum needs the bf16 output");
  if (rows == 0) return HCT_OK;
  const size_t stage_bytes = (dy_bf16 ? 2 * static_cast<size_t>(dim) + dim / 2 : 3 * static_cast<size_t>(dim)) * sizeof(float);
  // 12 warps with a 2-deep ring when that fits, else 8 warps with the deepest ring that fits
  const int warps = 12 * 2 * stage_bytes <= static_cast<size_t>(LN_SMEM_MAX) ? 12 : 8;
  int stages = 2;
  if (warps == 8 && 8 * 3 * stage_bytes <= static_cast<size_t>(LN_SMEM_MAX)) stages = 3;
  const int grid = grid_for(rows, warps, hct_num_sms());
  const size_t scratch = static_cast<size_t>(warps) * 3 * dim * sizeof(float);        // final cross-warp reduction
  size_t smem = static_cast<size_t>(warps) * stages * stage_bytes;
  if (smem < scratch) smem = scratch;
  HCT_REQUIRE(smem <= static_cast<size_t>(LN_SMEM_MAX), "layernorm_bwd: dim=%d needs %zu bytes of shared memory", dim, smem);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  bf16* dx16 = static_cast<bf16*>(dx_out_bf16);
  // algorithmic bytes: dy + x + (residual gradient) in, dx fp32 (+ bf16 copy) out
  HctProfScope prof(st, HCT_PROF_LN_BWD,
                    static_cast<double>(rows) * dim * ((dy_bf16 ? 2.0 : 4.0) + 4.0 + (dres_in ? 4.0 : 0.0) + 4.0 + (dx16 ? 2.0 : 0.0)));
  // bulk (TMA) staging needs 16-byte aligned rows of a multiple of 16 bytes in every stream
  const bool bulk = g_ln_bulk && dim % 8 == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
                    (reinterpret_cast<uintptr_t>(dy) % 16) == 0 && (dres_in == nullptr || reinterpret_cast<uintptr_t>(dres_in) % 16 == 0);
#define HCT_LN_BWD(NV, BF, ST, BULK, NW)                                                                                  \
  do {                                                                                                                    \
    static bool configured = false;                                                                                       \
    if (!configured) {                                                                                                    \
      cudaFuncSetAttribute(ln_bwd_kernel<NV, BF, ST, BULK, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM_MAX); \
      configured = true;                                                                                                  \
    }                                                                                                                     \
    hct_launch_pdl(ln_bwd_kernel<NV, BF, ST, BULK, NW>, dim3(grid), dim3(NW * 32), smem, st, dy, x, gamma, mean, rstd, dres_in, dx_out_f32,        \
                                                                     dx16, dgamma, dbeta, dxsum, rows, dim);              \
  } while (0)
#define HCT_LN_BWD_B(NV, BF, ST, NW) do { if (bulk) HCT_LN_BWD(NV, BF, ST, true, NW); else HCT_LN_BWD(NV, BF, ST, false, NW); } while (0)
#define HCT_LN_BWD_S(NV, BF)                                                                         \
  do {                                                                                               \
    if (warps == 12) HCT_LN_BWD_B(NV, BF, 2, 12);                                                    \
    else if (stages == 3) HCT_LN_BWD_B(NV, BF, 3, 8);                                                \
    else HCT_LN_BWD_B(NV, BF, 2, 8);                                                                 \
  } while (0)
#define HCT_LN_BWD_D(BF)                                                                                        \
  do {                                                                                                         \
    if (dim <= 256) HCT_LN_BWD_S(2, BF); else if (dim <= 768) HCT_LN_BWD_S(6, BF);                              \
    else if (dim <= 1024) HCT_LN_BWD_S(8, BF); else HCT_LN_BWD_S(16, BF);                                       \
  } while (0)
  if (dy_bf16) HCT_LN_BWD_D(true); else HCT_LN_BWD_D(false);
#undef HCT_LN_BWD_D
#undef HCT_LN_BWD_S
#undef HCT_LN_BWD_B
#undef HCT_LN_BWD
  return hct_check_launch("ln_bwd_kernel");
}

extern "C" int hct_cast_f32_to_bf16(const float* src, void* dst, int64_t n, hct_stream_t s) {
  if (n <= 0) return HCT_OK;
  HCT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "cast_f32_to_bf16: pointers must be 16-byte aligned");
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      src, static_cast<bf16*>(dst), n);
  return hct_check_launch("cast_f32_bf16_kernel");
}
extern "C" int hct_cast_bf16_to_f32(const void* src, float* dst, int64_t n, hct_stream_t s) {
  if (n <= 0) return HCT_OK;
  HCT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "cast_bf16_to_f32: pointers must be 16-byte aligned");
  cast_bf16_f32_kernel<<<grid_for(n / 8 + 1, 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      static_cast<const bf16*>(src), dst, n);
  return hct_check_launch("cast_bf16_f32_kernel");
}

extern "C" int hct_colsum(const void* x, int x_bf16, int64_t ld, float* out, int64_t rows, int32_t cols,
                          hct_stream_t s) {
  HCT_REQUIRE(cols > 0 && cols % 8 == 0 && ld % 8 == 0, "colsum: cols=%d ld=%lld must be multiples of 8", cols, (long long)ld);
  if (rows <= 0) return HCT_OK;
  dim3 grid((cols + 255) / 256, 1);
  long long gy = (rows + 63) / 64;
  const long long max_gy = (4LL * hct_num_sms() + grid.x - 1) / grid.x;
  if (gy > max_gy) gy = max_gy;
  grid.y = static_cast<unsigned>(gy < 1 ? 1 : gy);
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(s)>>>(x, x_bf16, ld, out, rows, cols);
  return hct_check_launch("colsum_kernel");
}

extern "C" int hct_broadcast_rows(const float* src, float* dst, int32_t batch, int32_t nrows, int64_t dst_rows_per_batch,
                                  int32_t row_off, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "broadcast_rows: dim %% 4");
  if (batch <= 0 || nrows <= 0) return HCT_OK;
  const long long total = static_cast<long long>(batch) * nrows * (dim / 4);
  broadcast_rows_kernel<<<grid_for(total, 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      src, dst, batch, nrows, dst_rows_per_batch, row_off, dim);
  return hct_check_launch("broadcast_rows_kernel");
}

extern "C" int hct_reduce_rows(const float* src, float* out, int32_t batch, int32_t nrows, int64_t src_rows_per_batch,
                               int32_t row_off, int32_t dim, hct_stream_t s) {
  if (batch <= 0 || nrows <= 0) return HCT_OK;
  const long long total = static_cast<long long>(nrows) * dim;
  reduce_rows_kernel<<<static_cast<int>((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(s)>>>(
      src, out, batch, nrows, src_rows_per_batch, row_off, dim);
  return hct_check_launch("reduce_rows_kernel");
}

extern "C" int hct_copy_rows_f32_to_bf16(const float* src, int64_t src_ld, int64_t src_rows_per_group,
                                         int32_t src_row_off, void* dst, int64_t dst_ld, int64_t groups,
                                         int32_t rows_per_group, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0, "copy_rows_f32_to_bf16: alignment");
  if (groups <= 0 || rows_per_group <= 0) return HCT_OK;
  copy_rows_f32_bf16_kernel<<<grid_for(groups * rows_per_group * (dim / 4), 256, hct_num_sms() * 8), 256, 0,
                              static_cast<cudaStream_t>(s)>>>(src, src_ld, src_rows_per_group, src_row_off,
                                                              static_cast<bf16*>(dst), dst_ld, groups, rows_per_group, dim);
  return hct_check_launch("copy_rows_f32_bf16_kernel");
}

extern "C" int hct_scatter_add_rows(const void* src, const int32_t* idx, float* out, int64_t rows, int32_t dim,
                                    hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "scatter_add_rows: dim %% 4");
  if (rows <= 0) return HCT_OK;
  scatter_add_rows_kernel<<<grid_for(rows * (dim / 4), 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      static_cast<const bf16*>(src), idx, out, rows, dim);
  return hct_check_launch("scatter_add_rows_kernel");
}

extern "C" int hct_gelu_bwd(const void* dy, const void* pre, void* out, int64_t n, hct_stream_t s) {
  HCT_REQUIRE(n % 8 == 0, "gelu_bwd: n must be a multiple of 8");
  if (n <= 0) return HCT_OK;
  gelu_bwd_kernel<<<grid_for(n / 8, 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      static_cast<const bf16*>(dy), static_cast<const bf16*>(pre), static_cast<bf16*>(out), n);
  return hct_check_launch("gelu_bwd_kernel");
}
