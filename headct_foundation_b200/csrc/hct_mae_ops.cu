// MAE-specific HBM-bound kernels: HU windowing, patchify (im2col for Conv3d k=s), random-masking
// indices (stable rank sort), token gather/scatter, decoder input assembly, masked-MSE loss.
#include "../../include/hct_b200.h"
#include <cuda_fp16.h>

#include "hct_common.cuh"

namespace {

inline int grid_for(long long work_items, int threads, int max_blocks) {
  long long g = (work_items + threads - 1) / threads;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------ a1: HU windowing
struct WindowParams { float a_min[8]; float inv_w[8]; int nwin; };

template <bool IN_I16, bool OUT_BF16>
__global__ void window_kernel(const void* __restrict__ hu, void* __restrict__ out, long long nvol, long long vox,
                              WindowParams wp) {
  const long long vox4 = vox >> 2;
  const long long total = nvol * vox4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = i / vox4, k = i % vox4;
    float4 x;
    if (IN_I16) {
      const short4 s = reinterpret_cast<const short4*>(reinterpret_cast<const short*>(hu) + v * vox)[k];
      x = make_float4(s.x, s.y, s.z, s.w);
    } else {
      x = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(hu) + v * vox)[k];
    }
#pragma unroll 3
    for (int w = 0; w < wp.nwin; ++w) {
      // (x - a_min) / (a_max - a_min) then clip: a true division keeps bit parity with the MONAI formula
      const float lo = wp.a_min[w], width = wp.inv_w[w];
      const float y0 = fminf(fmaxf((x.x - lo) / width, 0.f), 1.f), y1 = fminf(fmaxf((x.y - lo) / width, 0.f), 1.f);
      const float y2 = fminf(fmaxf((x.z - lo) / width, 0.f), 1.f), y3 = fminf(fmaxf((x.w - lo) / width, 0.f), 1.f);
      const long long o = (v * wp.nwin + w) * vox;
      if (OUT_BF16) {
        uint2 u; u.x = pack_bf16x2(y0, y1); u.y = pack_bf16x2(y2, y3);
        reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out) + o)[k] = u;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o)[k] = make_float4(y0, y1, y2, y3);
      }
    }
  }
}

// ------------------------------------------------------------------ a2: patchify (im2col, K order c,ph,pw,pd)
// one CTA per output row; thread t copies run t = (c, ph, pw): p contiguous floats along D.
template <bool OUT_F32>
__global__ void __launch_bounds__(128)
patchify_kernel(const float* __restrict__ x, void* __restrict__ cols_, const long long* __restrict__ patch_ids,
                int* __restrict__ pos_idx_out, int C, int H, int W, int D, int p, int rows_per_vol) {
  bf16* cols = static_cast<bf16*>(cols_);
  const long long row = blockIdx.x;
  const int b = static_cast<int>(row / rows_per_vol), j = static_cast<int>(row % rows_per_vol);
  const int gw = W / p, gd = D / p;
  const int pid = patch_ids ? static_cast<int>(patch_ids[row]) : j;
  if (threadIdx.x == 0 && pos_idx_out) pos_idx_out[row] = pid;
  const int pd0 = (pid % gd) * p, pw0 = ((pid / gd) % gw) * p, ph0 = (pid / (gd * gw)) * p;
  const int runs = C * p * p;
  const long long K = static_cast<long long>(runs) * p;
  bf16* dst = cols + row * K;
  for (int r = threadIdx.x; r < runs; r += blockDim.x) {
    const int c = r / (p * p), ph = (r / p) % p, pw = r % p;
    const float* src = x + (((static_cast<long long>(b) * C + c) * H + ph0 + ph) * W + pw0 + pw) * D + pd0;
    if (OUT_F32) {                                   // fp32 mode: the patch rows stay fp32 (split into bf16 terms later)
      float* d32 = static_cast<float*>(cols_) + row * K + static_cast<long long>(r) * p;
      for (int k = 0; k < p; ++k) d32[k] = src[k];
      continue;
    }
    bf16* d = dst + static_cast<long long>(r) * p;
    if ((p & 3) == 0 && (D & 3) == 0) {
      for (int k = 0; k < p; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + k);
        uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(d + k) = u;
      }
    } else {
      for (int k = 0; k < p; ++k) d[k] = __float2bfloat16_rn(src[k]);
    }
  }
}

// ------------------------------------------------------------------ a4: masking indices (stable rank sort)
// rank[i] = #{j : noise[j] < noise[i]  or (noise[j] == noise[i] and j < i)}  == ids_restore[i]
// (stable ascending argsort followed by its inverse, mae.py:208-209); ids_shuffle[rank[i]] = i.
__global__ void mask_indices_kernel(const float* __restrict__ noise, long long* __restrict__ ids_restore,
                                    long long* __restrict__ ids_keep, float* __restrict__ mask, int L, int len_keep) {
  extern __shared__ float snoise[];
  const long long row = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) snoise[i] = noise[row * L + i];
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float v = snoise[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const float u = snoise[j];
      rank += (u < v) || (u == v && j < i);
    }
    ids_restore[row * L + i] = rank;
    mask[row * L + i] = rank >= len_keep ? 1.f : 0.f;
    if (rank < len_keep) ids_keep[row * len_keep + rank] = i;
  }
}

// ------------------------------------------------------------------ token gather / scatter (fp32 rows)
__global__ void gather_tokens_kernel(const float* __restrict__ src, const long long* __restrict__ ids,
                                     float* __restrict__ dst, int L, int n_ids, long long dst_rows_per_batch,
                                     int row_off, int dim, long long total_rows) {
  const int nv = dim >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_rows * nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const long long r = i / nv;
    const long long n = r / n_ids;
    const int j = static_cast<int>(r % n_ids);
    const long long s = ids[r];
    reinterpret_cast<float4*>(dst + (n * dst_rows_per_batch + row_off + j) * dim)[c] =
        reinterpret_cast<const float4*>(src + (n * L + s) * dim)[c];
  }
}
__global__ void scatter_tokens_kernel(const float* __restrict__ ddst, const long long* __restrict__ ids,
                                      float* __restrict__ dsrc, int L, int n_ids, long long ddst_rows_per_batch,
                                      int row_off, int dim, long long total_rows) {
  const int nv = dim >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_rows * nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const long long r = i / nv;
    const long long n = r / n_ids;
    const int j = static_cast<int>(r % n_ids);
    const long long s = ids[r];
    reinterpret_cast<float4*>(dsrc + (n * L + s) * dim)[c] =
        reinterpret_cast<const float4*>(ddst + (n * ddst_rows_per_batch + row_off + j) * dim)[c];
  }
}

// ------------------------------------------------------------------ a7: decoder input assembly
template <bool Y_F32>
__device__ __forceinline__ float4 load_row4(const void* base, long long row, int dim, int c) {
  if (Y_F32) return reinterpret_cast<const float4*>(static_cast<const float*>(base) + row * dim)[c];
  const uint2 u = reinterpret_cast<const uint2*>(static_cast<const bf16*>(base) + row * dim)[c];
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <bool Y_F32>
__global__ void decoder_assemble_kernel(const void* __restrict__ y, const long long* __restrict__ ids_restore,
                                        const float* __restrict__ mask_token, const float* __restrict__ dec_cls,
                                        const float* __restrict__ dec_pos, float* __restrict__ out, int L, int keep,
                                        int dim, long long total_rows) {
  const int nv = dim >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_rows * nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const long long r = i / nv;            // output row in [0, N*(L+1))
    const long long n = r / (L + 1);
    const int t = static_cast<int>(r % (L + 1));
    float4 v, add;
    if (t == 0) {
      v = load_row4<Y_F32>(y, n * (keep + 1), dim, c);
      add = __ldg(reinterpret_cast<const float4*>(dec_cls) + c);
    } else {
      const long long src = ids_restore[n * L + (t - 1)];
      if (src < keep) {
        v = load_row4<Y_F32>(y, n * (keep + 1) + 1 + src, dim, c);
      } else {
        v = __ldg(reinterpret_cast<const float4*>(mask_token) + c);
      }
      add = __ldg(reinterpret_cast<const float4*>(dec_pos + static_cast<long long>(t - 1) * dim) + c);
    }
    reinterpret_cast<float4*>(out + r * dim)[c] = make_float4(v.x + add.x, v.y + add.y, v.z + add.z, v.w + add.w);
  }
}

// block (dim/4 threads, <= 512) walks rows; masked rows accumulate into dmask_token, row 0 into ddec_cls
template <bool Y_F32>
__global__ void decoder_assemble_bwd_kernel(const float* __restrict__ dout, const long long* __restrict__ ids_restore,
                                            void* __restrict__ dy, float* __restrict__ dmask_token,
                                            float* __restrict__ ddec_cls, int L, int keep, int dim,
                                            long long total_rows) {
  const int nv = dim >> 2;
  for (int c = threadIdx.x; c < nv; c += blockDim.x) {
    float4 am = make_float4(0, 0, 0, 0), ac = make_float4(0, 0, 0, 0);
    for (long long r = blockIdx.x; r < total_rows; r += gridDim.x) {
      const long long n = r / (L + 1);
      const int t = static_cast<int>(r % (L + 1));
      const float4 g = reinterpret_cast<const float4*>(dout + r * dim)[c];
      long long dst_row = -1;
      if (t == 0) {
        ac.x += g.x; ac.y += g.y; ac.z += g.z; ac.w += g.w;
        dst_row = n * (keep + 1);
      } else {
        const long long src = ids_restore[n * L + (t - 1)];
        if (src < keep) dst_row = n * (keep + 1) + 1 + src;
        else { am.x += g.x; am.y += g.y; am.z += g.z; am.w += g.w; }
      }
      if (dst_row >= 0) {
        if (Y_F32) {
          reinterpret_cast<float4*>(static_cast<float*>(dy) + dst_row * dim)[c] = g;
        } else {
          uint2 u; u.x = pack_bf16x2(g.x, g.y); u.y = pack_bf16x2(g.z, g.w);
          reinterpret_cast<uint2*>(static_cast<bf16*>(dy) + dst_row * dim)[c] = u;
        }
      }
    }
    if (dmask_token) {
      atomicAdd(dmask_token + 4 * c, am.x); atomicAdd(dmask_token + 4 * c + 1, am.y);
      atomicAdd(dmask_token + 4 * c + 2, am.z); atomicAdd(dmask_token + 4 * c + 3, am.w);
    }
    if (ddec_cls) {
      atomicAdd(ddec_cls + 4 * c, ac.x); atomicAdd(ddec_cls + 4 * c + 1, ac.y);
      atomicAdd(ddec_cls + 4 * c + 2, ac.z); atomicAdd(ddec_cls + 4 * c + 3, ac.w);
    }
  }
}

// ------------------------------------------------------------------ a8: masked MSE loss (fwd and bwd share the staging)
// One CTA per MASKED patch.  The target patch is staged into shared memory in the prediction's
// (ph,pw,pd,c) order from the volume's (c,ph,pw,pd) runs, optionally normalised, then compared with
// the bf16 prediction row using 16-byte loads.
// PS: patch side as a compile-time constant (12 = every shipped yaml; 0 = generic).  The target gather computes four
// quotients per 16-byte load; with a run-time divisor that arithmetic (ALU 50 %, issue slots 62 % busy under ncu) and not
// DRAM bounded the kernel.
// PF32: prediction (and its gradient) are fp32 rows instead of bf16 (fp32 mode)
template <bool BWD, int PS, bool PF32 = false>
__global__ void __launch_bounds__(256)
mae_loss_kernel(const void* __restrict__ pred_, const float* __restrict__ imgs, const float* __restrict__ mask,
                float* __restrict__ per_patch, const float* __restrict__ dloss, const float* __restrict__ mask_sum,
                void* __restrict__ dpred_, int L, int C, int H, int W, int D, int p_arg, int norm_pix, int prefix) {
  const bf16* pred = static_cast<const bf16*>(pred_);
  bf16* dpred = static_cast<bf16*>(dpred_);
  const float* pred32 = static_cast<const float*>(pred_);
  float* dpred32 = static_cast<float*>(dpred_);
  const int p = PS > 0 ? PS : p_arg;
  extern __shared__ __align__(16) float tgt[];     // [P]
  __shared__ float red[33];
  const long long patch = blockIdx.x;          // n * L + l
  const int P = C * p * p * p;
  const float m = mask[patch];
  const int n = static_cast<int>(patch / L), l = static_cast<int>(patch % L);
  // prediction rows carry `prefix` extra rows (the cls token) per sample: row = n*(L+prefix) + prefix + l
  const long long prow = static_cast<long long>(n) * (L + prefix) + prefix + l;
  if (BWD && l == 0 && prefix > 0) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (PF32) {
      for (int i = threadIdx.x; i < prefix * (P >> 2); i += blockDim.x)
        reinterpret_cast<uint4*>(dpred32 + static_cast<long long>(n) * (L + prefix) * P)[i] = z;
    } else {
      for (int i = threadIdx.x; i < prefix * (P >> 3); i += blockDim.x)
        reinterpret_cast<uint4*>(dpred + static_cast<long long>(n) * (L + prefix) * P)[i] = z;
    }
  }
  if (m == 0.f) {
    if (BWD) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      if (PF32) { for (int i = threadIdx.x; i < (P >> 2); i += blockDim.x) reinterpret_cast<uint4*>(dpred32 + prow * P)[i] = z; }
      else { for (int i = threadIdx.x; i < (P >> 3); i += blockDim.x) reinterpret_cast<uint4*>(dpred + prow * P)[i] = z; }
    } else if (threadIdx.x == 0) {
      per_patch[patch] = 0.f;
    }
    return;
  }
  const int gw = W / p, gd = D / p;
  const int pd0 = (l % gd) * p, pw0 = ((l / gd) % gw) * p, ph0 = (l / (gd * gw)) * p;
  const int runs = C * p * p;
  // the prediction row is fetched FIRST, into registers: its DRAM round trip overlaps the target gather below instead of
  // following it behind the barrier (one exposed memory latency per CTA instead of two)
  constexpr int PRE = 4;
  const bf16* pr = pred + prow * P;
  const int nvec = P >> 3;
  const bool preload = !PF32 && nvec <= PRE * static_cast<int>(blockDim.x);
  uint4 pu[PRE];
  if (preload) {
#pragma unroll
    for (int k = 0; k < PRE; ++k) {
      const int i = threadIdx.x + k * blockDim.x;
      pu[k] = i < nvec ? __ldg(reinterpret_cast<const uint4*>(pr) + i) : make_uint4(0, 0, 0, 0);
    }
  }
  float lsum = 0.f;
  if ((p & 3) == 0 && (D & 3) == 0) {
    // a patch row along D is p contiguous floats, 16-byte aligned: fetch it as p/4 vectors (a quarter of the
    // load instructions / L1 wavefronts of the scalar loop, which was what bounded this kernel)
    const int p4 = p >> 2;
    for (int idx = threadIdx.x; idx < runs * p4; idx += blockDim.x) {
      const int r = idx / p4, k4 = idx - r * p4;
      const int c = r / (p * p), ph = (r / p) % p, pw = r % p;
      const float4 v = __ldg(reinterpret_cast<const float4*>(
          imgs + (((static_cast<long long>(n) * C + c) * H + ph0 + ph) * W + pw0 + pw) * D + pd0) + k4);
      float* d = tgt + (static_cast<long long>(ph * p + pw) * p + 4 * k4) * C + c;
      d[0] = v.x; d[C] = v.y; d[2 * C] = v.z; d[3 * C] = v.w;
      lsum += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (int r = threadIdx.x; r < runs; r += blockDim.x) {
      const int c = r / (p * p), ph = (r / p) % p, pw = r % p;
      const float* src = imgs + (((static_cast<long long>(n) * C + c) * H + ph0 + ph) * W + pw0 + pw) * D + pd0;
      float* d = tgt + (static_cast<long long>(ph * p + pw) * p) * C + c;
      for (int k = 0; k < p; ++k) { const float v = src[k]; d[k * C] = v; lsum += v; }
    }
  }
  __syncthreads();
  float mean = 0.f, inv_std = 1.f;
  if (norm_pix) {
    mean = block_sum(lsum, red) / P;
    float q = 0.f;
    for (int i = threadIdx.x; i < P; i += blockDim.x) { const float dlt = tgt[i] - mean; q += dlt * dlt; }
    const float var = block_sum(q, red) / (P - 1);     // unbiased (mae.py:292)
    inv_std = rsqrtf(var + 1e-6f);
  }
  if (!BWD) {
    float s = 0.f;
    for (int i = threadIdx.x, k = 0; i < nvec; i += blockDim.x, ++k) {
      float pv[8];
      if (PF32) {
        const float4 a = reinterpret_cast<const float4*>(pred32 + prow * P)[2 * i], b = reinterpret_cast<const float4*>(pred32 + prow * P)[2 * i + 1];
        pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; pv[4] = b.x; pv[5] = b.y; pv[6] = b.z; pv[7] = b.w;
      } else {
        uint4 u;
        if (preload) { u = k == 0 ? pu[0] : (k == 1 ? pu[1] : (k == 2 ? pu[2] : pu[3])); }
        else u = reinterpret_cast<const uint4*>(pr)[i];
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z), d2 = unpack_bf16x2(u.w);
        pv[0] = a.x; pv[1] = a.y; pv[2] = b.x; pv[3] = b.y; pv[4] = c2.x; pv[5] = c2.y; pv[6] = d2.x; pv[7] = d2.y;
      }
      const float4 t0 = *reinterpret_cast<const float4*>(tgt + i * 8), t1 = *reinterpret_cast<const float4*>(tgt + i * 8 + 4);
      const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float e = pv[k] - (tv[k] - mean) * inv_std; s += e * e; }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) per_patch[patch] = s / P;
  } else {
    const float scale = dloss[0] * m * 2.f / (static_cast<float>(P) * mask_sum[0]);
    for (int i = threadIdx.x, k = 0; i < nvec; i += blockDim.x, ++k) {
      float pv[8];
      if (PF32) {
        const float4 a = reinterpret_cast<const float4*>(pred32 + prow * P)[2 * i], b = reinterpret_cast<const float4*>(pred32 + prow * P)[2 * i + 1];
        pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; pv[4] = b.x; pv[5] = b.y; pv[6] = b.z; pv[7] = b.w;
      } else {
        uint4 u;
        if (preload) { u = k == 0 ? pu[0] : (k == 1 ? pu[1] : (k == 2 ? pu[2] : pu[3])); }
        else u = reinterpret_cast<const uint4*>(pr)[i];
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z), d2 = unpack_bf16x2(u.w);
        pv[0] = a.x; pv[1] = a.y; pv[2] = b.x; pv[3] = b.y; pv[4] = c2.x; pv[5] = c2.y; pv[6] = d2.x; pv[7] = d2.y;
      }
      const float4 t0 = *reinterpret_cast<const float4*>(tgt + i * 8), t1 = *reinterpret_cast<const float4*>(tgt + i * 8 + 4);
      const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
      float g[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = scale * (pv[k] - (tv[k] - mean) * inv_std);
      if (PF32) {
        reinterpret_cast<float4*>(dpred32 + prow * P)[2 * i] = make_float4(g[0], g[1], g[2], g[3]);
        reinterpret_cast<float4*>(dpred32 + prow * P)[2 * i + 1] = make_float4(g[4], g[5], g[6], g[7]);
      } else {
        uint4 o;
        o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]); o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
        reinterpret_cast<uint4*>(dpred + prow * P)[i] = o;
      }
    }
  }
}

// deterministic final reduction: loss_out[0] = sum(per_patch * mask), loss_out[1] = sum(mask)
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(const float* __restrict__ per_patch, const float* __restrict__ mask, float* __restrict__ loss_out,
                   long long n) {
  __shared__ float red[33];
  float a = 0.f, b = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { const float m = mask[i]; a += per_patch[i] * m; b += m; }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) { loss_out[0] = a; loss_out[1] = b; loss_out[2] = a / b; }
}

// ------------------------------------------------------------------ train-time augmentation of cached volumes
// RandFlipd x 3 + RandShiftIntensityd of mae3d_transforms (src/data/transforms.py:200-223) on a batch that already sits
// in HBM: out[v, c, i, j, k] = in[v, c, i', j', k'] + offset[v], with an axis reversed where its flip bit is set.
// One pass, 16-byte vectors along the innermost axis (reversed in registers when that axis flips).
template <bool IN_F16>
__global__ void flip_shift_kernel(const void* __restrict__ in, float* __restrict__ out, const unsigned char* __restrict__ flips,
                                  const float* __restrict__ offsets, int C, int D0, int D1, int D2) {
  // one CTA per (sample, channel, i0) plane of D1 x D2 voxels: no per-element 64-bit index arithmetic
  const int plane = blockIdx.x;                               // (v * C + c) * D0 + i0
  const int i0 = plane % D0;
  const int vc = plane / D0;
  const int v = vc / C;
  const unsigned f = flips != nullptr ? flips[v] : 0u;
  const float off = offsets != nullptr ? offsets[v] : 0.f;
  const int si = (f & 1u) ? D0 - 1 - i0 : i0;
  const long long src_plane = (static_cast<long long>(vc) * D0 + si) * D1 * D2;
  const long long dst_plane = static_cast<long long>(plane) * D1 * D2;
  const int w8n = D2 >> 3;
  for (int t = threadIdx.x; t < D1 * w8n; t += blockDim.x) {
    const int j = t / w8n, w8 = t - j * w8n;
    const int sj = (f & 2u) ? D1 - 1 - j : j, sw = (f & 4u) ? w8n - 1 - w8 : w8;
    const long long src = src_plane + static_cast<long long>(sj) * D2 + 8 * sw;
    float x[8];
    if (IN_F16) {
      const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(in) + src);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        x[2 * k] = a.x; x[2 * k + 1] = a.y;
      }
    } else {
      const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + src);
      const float4 b2 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + src + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b2.x; x[5] = b2.y; x[6] = b2.z; x[7] = b2.w;
    }
    float4 o0, o1;
    if (f & 4u) {
      o0 = make_float4(x[7] + off, x[6] + off, x[5] + off, x[4] + off);
      o1 = make_float4(x[3] + off, x[2] + off, x[1] + off, x[0] + off);
    } else {
      o0 = make_float4(x[0] + off, x[1] + off, x[2] + off, x[3] + off);
      o1 = make_float4(x[4] + off, x[5] + off, x[6] + off, x[7] + off);
    }
    float* dst = out + dst_plane + static_cast<long long>(j) * D2 + 8 * w8;
    *reinterpret_cast<float4*>(dst) = o0;
    *reinterpret_cast<float4*>(dst + 4) = o1;
  }
}

// RandGaussianSmoothd (transforms.py:228-236): separable Gaussian with one sigma per sample and axis, zero padding.
// One launch per axis over the n SMOOTHED samples only (prob 0.2): sample s is read from volume in_idx[s] of `in` and
// written to volume out_idx[s] of `out` (NULL = s), so the three passes run batch -> compact scratch -> compact scratch
// -> batch without touching the other samples.  `taps` = each sample's kernel for this axis ([n][2 R + 1]).
// One CTA per (sample, channel, i0) plane; the kernel taps sit in shared memory.
__global__ void gauss_axis_kernel(const float* __restrict__ in, const int* __restrict__ in_idx, float* __restrict__ out,
                                  const int* __restrict__ out_idx, const float* __restrict__ taps, int R,
                                  int C, int D0, int D1, int D2, int axis) {
  __shared__ float st[129];
  const int plane = blockIdx.x;                               // (s * C + c) * D0 + i0
  const int i0 = plane % D0;
  const int sc = plane / D0;
  const int s = sc / C, c = sc - s * C;
  for (int k = threadIdx.x; k < 2 * R + 1; k += blockDim.x) st[k] = taps[static_cast<long long>(s) * (2 * R + 1) + k];
  __syncthreads();
  const long long vol = static_cast<long long>(C) * D0 * D1 * D2;
  const long long off = (static_cast<long long>(c) * D0 + i0) * D1 * D2;
  const float* src = in + (in_idx != nullptr ? in_idx[s] : s) * vol + off;
  float* dst = out + (out_idx != nullptr ? out_idx[s] : s) * vol + off;
  const int n_axis = axis == 0 ? D0 : (axis == 1 ? D1 : D2);
  const int stride = axis == 0 ? D1 * D2 : (axis == 1 ? D2 : 1);
  for (int t = threadIdx.x; t < D1 * D2; t += blockDim.x) {
    const int pos = axis == 0 ? i0 : (axis == 1 ? t / D2 : t % D2);
    const int lo = max(-R, -pos), hi = min(R, n_axis - 1 - pos);
    float acc = 0.f;
    for (int k = lo; k <= hi; ++k) acc = fmaf(st[k + R], src[t + k * stride], acc);
    dst[t] = acc;
  }
}

// ------------------------------------------------------------------ DINO multi-crop (DataAugmentationDINO3D)
// ResizeWithPadOrCrop + (Center)SpatialCrop + RandSpatialCrop + Resize(mode="area") of src/data/transforms.py:75-105
// as ONE gather: crop n takes the box [start, start + size) (in SOURCE voxel coordinates; whatever falls outside the
// source volume is the zero padding of ResizeWithPadOrCrop) of sample src_idx[n] and area-resizes it to T0 x T1 x T2.
// Area interpolation = adaptive average pooling: out[i] = mean over [floor(i a / T), ceil((i + 1) a / T)) per axis.
struct CropBox { int sample, s0, s1, s2, n0, n1, n2, flips; };   // flips: bit k = reverse OUTPUT axis k (RandFlip after the resize)
template <bool IN_F16>
__global__ void crop_resize_area_kernel(const void* __restrict__ src, const CropBox* __restrict__ boxes,
                                        const float* __restrict__ offsets, float* __restrict__ out,
                                        int C, int S0, int S1, int S2, int T0, int T1, int T2) {
  const int plane = blockIdx.x;                               // (crop * C + c) * T0 + i_out
  const int i_out = plane % T0;
  const int nc = plane / T0;
  const int n = nc / C, c = nc - n * C;
  const CropBox bx = boxes[n];
  const float off = offsets != nullptr ? offsets[n] : 0.f;    // RandShiftIntensity after the resize
  const int i = (bx.flips & 1) ? T0 - 1 - i_out : i_out;
  const long long vbase = (static_cast<long long>(bx.sample) * C + c) * S0;
  const int a0 = static_cast<int>((static_cast<long long>(i) * bx.n0) / T0);
  const int b0 = static_cast<int>((static_cast<long long>(i + 1) * bx.n0 + T0 - 1) / T0);
  // the part of each averaging window that lies inside the source volume (the rest is zero padding: it only counts
  // in the divisor), as global index ranges -> no bounds checks in the loops
  const int g0a = max(bx.s0 + a0, 0), g0b = min(bx.s0 + b0, S0);
  float* dst = out + static_cast<long long>(plane) * T1 * T2;
  for (int t = threadIdx.x; t < T1 * T2; t += blockDim.x) {
    const int j_out = t / T2, k_out = t - j_out * T2;
    const int j = (bx.flips & 2) ? T1 - 1 - j_out : j_out, k = (bx.flips & 4) ? T2 - 1 - k_out : k_out;
    const int a1 = (j * bx.n1) / T1, b1 = ((j + 1) * bx.n1 + T1 - 1) / T1;
    const int a2 = (k * bx.n2) / T2, b2 = ((k + 1) * bx.n2 + T2 - 1) / T2;
    const int g1a = max(bx.s1 + a1, 0), g1b = min(bx.s1 + b1, S1);
    const int g2a = max(bx.s2 + a2, 0), g2b = min(bx.s2 + b2, S2);
    float acc = 0.f;
    for (int g0 = g0a; g0 < g0b; ++g0) {
      for (int g1 = g1a; g1 < g1b; ++g1) {
        const long long row = ((vbase + g0) * S1 + g1) * S2;
        if (IN_F16) {
          const __half* r = reinterpret_cast<const __half*>(src) + row;
          for (int g2 = g2a; g2 < g2b; ++g2) acc += __half2float(__ldg(r + g2));
        } else {
          const float* r = reinterpret_cast<const float*>(src) + row;
          for (int g2 = g2a; g2 < g2b; ++g2) acc += __ldg(r + g2);
        }
      }
    }
    dst[t] = acc / static_cast<float>((b0 - a0) * (b1 - a1) * (b2 - a2)) + off;
  }
}

// Row-staged version (source rows of <= 256 elements, 16-byte aligned).  One warp per OUTPUT row (j, all k) of one
// output plane: the <= 3 x 3 source rows of its averaging window are fetched with one 16-byte (fp16) / 32-byte (fp32)
// load per lane and summed element-wise in registers -- each source element is touched by ONE instruction instead of
// once per overlapping output window -- then the summed row goes through a per-warp shared-memory line and every lane
// adds up the <= 3 entries of its three output columns.  No CTA-wide synchronisation.
constexpr int CROP_ROW_MAX = 256;
constexpr int CROP_MAX_K = 4;                                 // output columns per lane: T2 <= 128
template <bool IN_F16>
__global__ void __launch_bounds__(256)
crop_resize_area_rows_kernel(const void* __restrict__ src, const CropBox* __restrict__ boxes,
                             const float* __restrict__ offsets, float* __restrict__ out,
                             int C, int S0, int S1, int S2, int T0, int T1, int T2) {
  __shared__ __align__(16) float rowbuf[8][CROP_ROW_MAX + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- per plane (once per CTA): which crop / channel / output slice, its source planes
  const int plane = blockIdx.x;                               // (crop * C + c) * T0 + i_out
  const int i_out = plane % T0;
  const int nc = plane / T0;
  const int n = nc / C, c = nc - n * C;
  const CropBox bx = boxes[n];
  const float off = offsets != nullptr ? offsets[n] : 0.f;
  const int i = (bx.flips & 1) ? T0 - 1 - i_out : i_out;
  const long long vbase = (static_cast<long long>(bx.sample) * C + c) * S0;
  const int a0 = static_cast<int>((static_cast<long long>(i) * bx.n0) / T0);
  const int b0 = static_cast<int>((static_cast<long long>(i + 1) * bx.n0 + T0 - 1) / T0);
  const int g0a = max(bx.s0 + a0, 0), g0b = min(bx.s0 + b0, S0);
  // the columns this crop can touch, widened to whole 8-element groups
  const int c_lo = max(bx.s2, 0) & ~7, c_hi = min(bx.s2 + bx.n2, S2);
  const int e0 = c_lo + lane * 8;
  const bool active = e0 < c_hi;
  // ---- per lane (once): the column windows of its output columns k_out = lane + 32 m, as offsets into the row line
  int w_lo[CROP_MAX_K], w_hi[CROP_MAX_K];
  float w_inv[CROP_MAX_K];
#pragma unroll
  for (int m = 0; m < CROP_MAX_K; ++m) {
    const int k_out = lane + 32 * m;
    w_lo[m] = 0; w_hi[m] = 0; w_inv[m] = 0.f;
    if (k_out < T2) {
      const int k = (bx.flips & 4) ? T2 - 1 - k_out : k_out;
      const int a2 = (k * bx.n2) / T2, b2 = ((k + 1) * bx.n2 + T2 - 1) / T2;
      w_lo[m] = max(bx.s2 + a2, 0) - c_lo;
      w_hi[m] = min(bx.s2 + b2, S2) - c_lo;
      w_inv[m] = 1.f / static_cast<float>((b0 - a0) * (b2 - a2));
    }
  }
  float* rb = rowbuf[warp];
  float* dplane = out + static_cast<long long>(plane) * T1 * T2;
  // ---- one output row per warp and iteration
  for (int j_out = warp; j_out < T1; j_out += 8) {
    const int j = (bx.flips & 2) ? T1 - 1 - j_out : j_out;
    const int a1 = (j * bx.n1) / T1, b1 = ((j + 1) * bx.n1 + T1 - 1) / T1;
    const int g1a = max(bx.s1 + a1, 0), g1b = min(bx.s1 + b1, S1);
    float sum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
      // (issuing the window's loads in predicated batches of eight was tried: the extra predication / conversion work
      // cost more than the exposed latency it hid -- 6.5 ms against 5.05 ms for the 256-crop DINO batch)
      for (int g0 = g0a; g0 < g0b; ++g0) {
        const long long prow = ((vbase + g0) * S1) * S2 + e0;
        for (int g1 = g1a; g1 < g1b; ++g1) {
          const long long row = prow + static_cast<long long>(g1) * S2;
          if (IN_F16) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(src) + row));
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(h[q]); sum[2 * q] += f.x; sum[2 * q + 1] += f.y; }
          } else {
            const float4 x = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + row));
            const float4 y = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + row) + 1);
            sum[0] += x.x; sum[1] += x.y; sum[2] += x.z; sum[3] += x.w; sum[4] += y.x; sum[5] += y.y; sum[6] += y.z; sum[7] += y.w;
          }
        }
      }
    }
    __syncwarp();                                             // the previous row's readers are done with the line
    *reinterpret_cast<float4*>(rb + lane * 8) = make_float4(sum[0], sum[1], sum[2], sum[3]);
    *reinterpret_cast<float4*>(rb + lane * 8 + 4) = make_float4(sum[4], sum[5], sum[6], sum[7]);
    __syncwarp();
    const float inv1 = 1.f / static_cast<float>(b1 - a1);
    float* dst = dplane + j_out * T2;
#pragma unroll
    for (int m = 0; m < CROP_MAX_K; ++m) {
      const int k_out = lane + 32 * m;
      if (k_out < T2) {
        float acc = 0.f;
        for (int g2 = w_lo[m]; g2 < w_hi[m]; ++g2) acc += rb[g2];
        dst[k_out] = acc * w_inv[m] * inv1 + off;
      }
    }
  }
}

// RandAdjustContrast (transforms.py:92): ((x - min) / (range + 1e-7))^gamma * range + min over the whole sample.
// Pass 1: per-sample min / max (ordered-int atomics); pass 2: apply to the samples whose gamma is set (> 0).
__device__ __forceinline__ int float_ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__global__ void minmax_init_kernel(int* __restrict__ mm, int n) {      // {ordered(+inf), ordered(-inf)} per sample
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) { mm[2 * s] = float_ordered(INFINITY); mm[2 * s + 1] = float_ordered(-INFINITY); }
}
__global__ void sample_minmax_kernel(const float* __restrict__ x, const float* __restrict__ gamma, int* __restrict__ mm,
                                     long long per_sample) {
  const int s = blockIdx.y;
  if (gamma[s] <= 0.f) return;
  const float4* p = reinterpret_cast<const float4*>(x + static_cast<long long>(s) * per_sample);
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < (per_sample >> 2);
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = p[i];
    lo = fminf(lo, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
    hi = fmaxf(hi, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMin(mm + 2 * s, float_ordered(lo)); atomicMax(mm + 2 * s + 1, float_ordered(hi)); }
}
__global__ void gamma_kernel(float* __restrict__ x, const float* __restrict__ gamma, const int* __restrict__ mm, long long per_sample) {
  const int s = blockIdx.y;
  const float g = gamma[s];
  if (g <= 0.f) return;
  const float lo = ordered_float(mm[2 * s]), range = ordered_float(mm[2 * s + 1]) - lo;
  const float inv = 1.0f / (range + 1e-7f);
  float4* p = reinterpret_cast<float4*>(x + static_cast<long long>(s) * per_sample);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < (per_sample >> 2);
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = p[i];
    v.x = powf((v.x - lo) * inv, g) * range + lo; v.y = powf((v.y - lo) * inv, g) * range + lo;
    v.z = powf((v.z - lo) * inv, g) * range + lo; v.w = powf((v.w - lo) * inv, g) * range + lo;
    p[i] = v;
  }
}

}  // namespace

extern "C" int hct_window_scale_stack(const void* hu, int hu_i16, void* out, int out_bf16, int64_t nvol, int64_t vox,
                                      int32_t nwin, const float* a_min, const float* a_max, hct_stream_t s) {
  HCT_REQUIRE(nwin >= 1 && nwin <= 8, "window_scale_stack: nwin=%d", nwin);
  HCT_REQUIRE(vox % 4 == 0, "window_scale_stack: voxels per volume must be a multiple of 4");
  if (nvol <= 0) return HCT_OK;
  WindowParams wp;
  wp.nwin = nwin;
  for (int i = 0; i < nwin; ++i) { wp.a_min[i] = a_min[i]; wp.inv_w[i] = a_max[i] - a_min[i]; }
  const int grid = grid_for(nvol * (vox / 4), 256, hct_num_sms() * 16);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  HctProfScope prof(st, HCT_PROF_WINDOW, static_cast<double>(nvol) * vox * ((hu_i16 ? 2.0 : 4.0) + nwin * (out_bf16 ? 2.0 : 4.0)));
  if (hu_i16) {
    if (out_bf16) window_kernel<true, true><<<grid, 256, 0, st>>>(hu, out, nvol, vox, wp);
    else window_kernel<true, false><<<grid, 256, 0, st>>>(hu, out, nvol, vox, wp);
  } else {
    if (out_bf16) window_kernel<false, true><<<grid, 256, 0, st>>>(hu, out, nvol, vox, wp);
    else window_kernel<false, false><<<grid, 256, 0, st>>>(hu, out, nvol, vox, wp);
  }
  return hct_check_launch("window_kernel");
}

extern "C" int hct_patchify(const float* x, void* cols, const int64_t* patch_ids, int32_t* pos_idx_out, int32_t B,
                            int32_t C, int32_t H, int32_t W, int32_t D, int32_t p, int32_t rows_per_vol,
                            hct_stream_t s) {
  HCT_REQUIRE(p > 0 && H % p == 0 && W % p == 0 && D % p == 0, "patchify: volume %dx%dx%d not divisible by patch %d", H, W, D, p);
  HCT_REQUIRE(rows_per_vol > 0 && rows_per_vol <= (H / p) * (W / p) * (D / p), "patchify: rows_per_vol=%d", rows_per_vol);
  if (B <= 0) return HCT_OK;
  HctProfScope prof(static_cast<cudaStream_t>(s), HCT_PROF_PATCHIFY,
                    static_cast<double>(B) * rows_per_vol * C * p * p * p * (4.0 + 2.0));
  patchify_kernel<false><<<static_cast<unsigned>(static_cast<long long>(B) * rows_per_vol), 128, 0, static_cast<cudaStream_t>(s)>>>(
      x, cols, reinterpret_cast<const long long*>(patch_ids), pos_idx_out, C, H, W, D, p, rows_per_vol);
  return hct_check_launch("patchify_kernel");
}

extern "C" int hct_patchify_f32(const float* x, float* cols, const int64_t* patch_ids, int32_t* pos_idx_out, int32_t B,
                                int32_t C, int32_t H, int32_t W, int32_t D, int32_t p, int32_t rows_per_vol, hct_stream_t s) {
  HCT_REQUIRE(p > 0 && H % p == 0 && W % p == 0 && D % p == 0, "patchify_f32: volume %dx%dx%d not divisible by patch %d", H, W, D, p);
  HCT_REQUIRE(rows_per_vol > 0 && rows_per_vol <= (H / p) * (W / p) * (D / p), "patchify_f32: rows_per_vol=%d", rows_per_vol);
  if (B <= 0) return HCT_OK;
  patchify_kernel<true><<<static_cast<unsigned>(static_cast<long long>(B) * rows_per_vol), 128, 0, static_cast<cudaStream_t>(s)>>>(
      x, cols, reinterpret_cast<const long long*>(patch_ids), pos_idx_out, C, H, W, D, p, rows_per_vol);
  return hct_check_launch("patchify_kernel<f32>");
}

extern "C" int hct_mask_indices(const float* noise, int64_t* ids_restore, int64_t* ids_keep, float* mask, int32_t N,
                                int32_t L, int32_t len_keep, hct_stream_t s) {
  HCT_REQUIRE(L > 0 && L <= 12288 && len_keep >= 0 && len_keep <= L, "mask_indices: L=%d len_keep=%d", L, len_keep);
  if (N <= 0) return HCT_OK;
  const int threads = L >= 512 ? 512 : ((L + 31) / 32) * 32;
  mask_indices_kernel<<<N, threads, L * sizeof(float), static_cast<cudaStream_t>(s)>>>(
      noise, reinterpret_cast<long long*>(ids_restore), reinterpret_cast<long long*>(ids_keep), mask, L, len_keep);
  return hct_check_launch("mask_indices_kernel");
}

extern "C" int hct_gather_tokens(const float* src, const int64_t* ids, float* dst, int32_t N, int32_t L, int32_t n_ids,
                                 int64_t dst_rows_per_batch, int32_t row_off, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "gather_tokens: dim %% 4");
  if (N <= 0 || n_ids <= 0) return HCT_OK;
  const long long rows = static_cast<long long>(N) * n_ids;
  gather_tokens_kernel<<<grid_for(rows * (dim / 4), 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      src, reinterpret_cast<const long long*>(ids), dst, L, n_ids, dst_rows_per_batch, row_off, dim, rows);
  return hct_check_launch("gather_tokens_kernel");
}

extern "C" int hct_scatter_tokens(const float* ddst, const int64_t* ids, float* dsrc, int32_t N, int32_t L,
                                  int32_t n_ids, int64_t ddst_rows_per_batch, int32_t row_off, int32_t dim,
                                  hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "scatter_tokens: dim %% 4");
  if (N <= 0) return HCT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  cudaError_t e = cudaMemsetAsync(dsrc, 0, sizeof(float) * static_cast<size_t>(N) * L * dim, st);
  if (e != cudaSuccess) { hct_set_error("scatter_tokens memset: %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
  if (n_ids <= 0) return HCT_OK;
  const long long rows = static_cast<long long>(N) * n_ids;
  scatter_tokens_kernel<<<grid_for(rows * (dim / 4), 256, hct_num_sms() * 8), 256, 0, st>>>(
      ddst, reinterpret_cast<const long long*>(ids), dsrc, L, n_ids, ddst_rows_per_batch, row_off, dim, rows);
  return hct_check_launch("scatter_tokens_kernel");
}

extern "C" int hct_decoder_assemble(const void* y, const int64_t* ids_restore, const float* mask_token,
                                    const float* dec_cls, const float* dec_pos, float* out, int32_t N, int32_t L,
                                    int32_t keep, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "decoder_assemble: dim %% 4");
  if (N <= 0) return HCT_OK;
  const long long rows = static_cast<long long>(N) * (L + 1);
  decoder_assemble_kernel<false><<<grid_for(rows * (dim / 4), 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      y, reinterpret_cast<const long long*>(ids_restore), mask_token, dec_cls, dec_pos, out, L, keep, dim, rows);
  return hct_check_launch("decoder_assemble_kernel");
}
extern "C" int hct_decoder_assemble_f32(const float* y, const int64_t* ids_restore, const float* mask_token,
                                        const float* dec_cls, const float* dec_pos, float* out, int32_t N, int32_t L,
                                        int32_t keep, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "decoder_assemble_f32: dim %% 4");
  if (N <= 0) return HCT_OK;
  const long long rows = static_cast<long long>(N) * (L + 1);
  decoder_assemble_kernel<true><<<grid_for(rows * (dim / 4), 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      y, reinterpret_cast<const long long*>(ids_restore), mask_token, dec_cls, dec_pos, out, L, keep, dim, rows);
  return hct_check_launch("decoder_assemble_kernel<f32>");
}

extern "C" int hct_decoder_assemble_bwd(const float* dout, const int64_t* ids_restore, void* dy, float* dmask_token,
                                        float* ddec_cls, int32_t N, int32_t L, int32_t keep, int32_t dim,
                                        hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "decoder_assemble_bwd: dim %% 4");
  if (N <= 0) return HCT_OK;
  const long long rows = static_cast<long long>(N) * (L + 1);
  int threads = ((dim / 4 + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  long long grid = rows < 4LL * hct_num_sms() ? rows : 4LL * hct_num_sms();
  decoder_assemble_bwd_kernel<false><<<static_cast<int>(grid), threads, 0, static_cast<cudaStream_t>(s)>>>(
      dout, reinterpret_cast<const long long*>(ids_restore), dy, dmask_token, ddec_cls, L, keep, dim, rows);
  return hct_check_launch("decoder_assemble_bwd_kernel");
}
extern "C" int hct_decoder_assemble_bwd_f32(const float* dout, const int64_t* ids_restore, float* dy, float* dmask_token,
                                            float* ddec_cls, int32_t N, int32_t L, int32_t keep, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(dim % 4 == 0, "decoder_assemble_bwd_f32: dim %% 4");
  if (N <= 0) return HCT_OK;
  const long long rows = static_cast<long long>(N) * (L + 1);
  int threads = ((dim / 4 + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  long long grid = rows < 4LL * hct_num_sms() ? rows : 4LL * hct_num_sms();
  decoder_assemble_bwd_kernel<true><<<static_cast<int>(grid), threads, 0, static_cast<cudaStream_t>(s)>>>(
      dout, reinterpret_cast<const long long*>(ids_restore), dy, dmask_token, ddec_cls, L, keep, dim, rows);
  return hct_check_launch("decoder_assemble_bwd_kernel<f32>");
}

static int loss_common_checks(int32_t C, int32_t H, int32_t W, int32_t D, int32_t p) {
  HCT_REQUIRE(p > 0 && H % p == 0 && W % p == 0 && D % p == 0, "mae_loss: volume not divisible by patch");
  const long long P = static_cast<long long>(C) * p * p * p;
  HCT_REQUIRE(P % 8 == 0 && P * 4 <= 200 * 1024, "mae_loss: patch dim %lld unsupported", P);
  return HCT_OK;
}

// loss_out layout: [0] = sum(mask*mse), [1] = sum(mask), [2] = loss, [3..3+N*L) = per-patch workspace
template <bool BWD, int PS, bool PF32>
static void loss_set_smem() {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(mae_loss_kernel<BWD, PS, PF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
}
template <bool PF32>
static int mae_loss_fwd_impl(const void* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask, float* loss_out,
                             int32_t N, int32_t C, int32_t H, int32_t W, int32_t D, int32_t p, int32_t norm_pix, hct_stream_t s) {
  int rc = loss_common_checks(C, H, W, D, p);
  if (rc != HCT_OK) return rc;
  if (N <= 0) return HCT_OK;
  const int L = (H / p) * (W / p) * (D / p);
  const int P = C * p * p * p;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  loss_set_smem<false, 0, PF32>(); loss_set_smem<false, 12, PF32>();
  float* per_patch = loss_out + 4;
  const unsigned grid = static_cast<unsigned>(static_cast<long long>(N) * L);
  // algorithmic bytes as in SURVEY 8(d): pred + target (fp32) of every patch row (an upper bound: unmasked rows are skipped)
  HctProfScope prof(st, HCT_PROF_LOSS, static_cast<double>(N) * L * P * ((PF32 ? 4.0 : 2.0) + 4.0));
  if (p == 12)
    mae_loss_kernel<false, 12, PF32><<<grid, 256, P * sizeof(float), st>>>(pred, imgs, mask, per_patch, nullptr, nullptr, nullptr, L, C,
                                                                           H, W, D, p, norm_pix, pred_prefix_rows);
  else
    mae_loss_kernel<false, 0, PF32><<<grid, 256, P * sizeof(float), st>>>(pred, imgs, mask, per_patch, nullptr, nullptr, nullptr, L, C,
                                                                          H, W, D, p, norm_pix, pred_prefix_rows);
  rc = hct_check_launch("mae_loss_kernel<fwd>");
  if (rc != HCT_OK) return rc;
  loss_reduce_kernel<<<1, 1024, 0, st>>>(per_patch, mask, loss_out, static_cast<long long>(N) * L);
  return hct_check_launch("loss_reduce_kernel");
}
template <bool PF32>
static int mae_loss_bwd_impl(const void* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask, const float* dloss,
                             const float* mask_sum, void* dpred, int32_t N, int32_t C, int32_t H, int32_t W, int32_t D, int32_t p,
                             int32_t norm_pix, hct_stream_t s) {
  int rc = loss_common_checks(C, H, W, D, p);
  if (rc != HCT_OK) return rc;
  if (N <= 0) return HCT_OK;
  const int L = (H / p) * (W / p) * (D / p);
  const int P = C * p * p * p;
  loss_set_smem<true, 0, PF32>(); loss_set_smem<true, 12, PF32>();
  const unsigned grid = static_cast<unsigned>(static_cast<long long>(N) * L);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  HctProfScope prof(st, HCT_PROF_LOSS, static_cast<double>(N) * L * P * (2.0 * (PF32 ? 4.0 : 2.0) + 4.0));      // + dpred out
  if (p == 12)
    mae_loss_kernel<true, 12, PF32><<<grid, 256, P * sizeof(float), st>>>(pred, imgs, mask, nullptr, dloss, mask_sum, dpred, L, C, H, W,
                                                                          D, p, norm_pix, pred_prefix_rows);
  else
    mae_loss_kernel<true, 0, PF32><<<grid, 256, P * sizeof(float), st>>>(pred, imgs, mask, nullptr, dloss, mask_sum, dpred, L, C, H, W,
                                                                         D, p, norm_pix, pred_prefix_rows);
  return hct_check_launch("mae_loss_kernel<bwd>");
}

extern "C" int hct_mae_loss_fwd(const void* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask,
                                float* loss_out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t D, int32_t p,
                                int32_t norm_pix, hct_stream_t s) {
  return mae_loss_fwd_impl<false>(pred, pred_prefix_rows, imgs, mask, loss_out, N, C, H, W, D, p, norm_pix, s);
}
extern "C" int hct_mae_loss_bwd(const void* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask,
                                const float* dloss, const float* mask_sum, void* dpred, int32_t N, int32_t C, int32_t H,
                                int32_t W, int32_t D, int32_t p, int32_t norm_pix, hct_stream_t s) {
  return mae_loss_bwd_impl<false>(pred, pred_prefix_rows, imgs, mask, dloss, mask_sum, dpred, N, C, H, W, D, p, norm_pix, s);
}
/* fp32 mode: prediction rows (and their gradient) are fp32 */
extern "C" int hct_mae_loss_fwd_f32(const float* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask,
                                    float* loss_out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t D, int32_t p,
                                    int32_t norm_pix, hct_stream_t s) {
  return mae_loss_fwd_impl<true>(pred, pred_prefix_rows, imgs, mask, loss_out, N, C, H, W, D, p, norm_pix, s);
}
extern "C" int hct_mae_loss_bwd_f32(const float* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask,
                                    const float* dloss, const float* mask_sum, float* dpred, int32_t N, int32_t C, int32_t H,
                                    int32_t W, int32_t D, int32_t p, int32_t norm_pix, hct_stream_t s) {
  return mae_loss_bwd_impl<true>(pred, pred_prefix_rows, imgs, mask, dloss, mask_sum, dpred, N, C, H, W, D, p, norm_pix, s);
}

extern "C" int hct_flip_shift(const void* in, int32_t in_f16, float* out, const uint8_t* flip_bits, const float* offsets,
                              int64_t nvol, int32_t C, int32_t D0, int32_t D1, int32_t D2, hct_stream_t s) {
  HCT_REQUIRE(C > 0 && D0 > 0 && D1 > 0 && D2 > 0 && D2 % 8 == 0, "flip_shift: bad volume %dx%dx%dx%d (innermost %% 8)", C, D0, D1, D2);
  HCT_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "flip_shift: pointers must be 16-byte aligned");
  if (nvol <= 0) return HCT_OK;
  const long long planes = nvol * C * D0;
  HCT_REQUIRE(planes <= 2147483647LL, "flip_shift: too many planes (%lld)", planes);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  if (in_f16) flip_shift_kernel<true><<<static_cast<unsigned>(planes), 256, 0, st>>>(in, out, flip_bits, offsets, C, D0, D1, D2);
  else flip_shift_kernel<false><<<static_cast<unsigned>(planes), 256, 0, st>>>(in, out, flip_bits, offsets, C, D0, D1, D2);
  return hct_check_launch("flip_shift_kernel");
}

extern "C" int hct_gaussian_smooth_axis(const float* in, const int32_t* in_idx, float* out, const int32_t* out_idx,
                                        const float* taps, int32_t radius, int64_t n, int32_t C, int32_t D0, int32_t D1,
                                        int32_t D2, int32_t axis, hct_stream_t s) {
  HCT_REQUIRE(axis >= 0 && axis <= 2 && radius >= 0 && radius <= 64, "gaussian_smooth_axis: axis=%d radius=%d", axis, radius);
  HCT_REQUIRE(in != out, "gaussian_smooth_axis: in-place filtering is not supported");
  if (n <= 0) return HCT_OK;
  const long long planes = n * C * D0;
  HCT_REQUIRE(planes <= 2147483647LL, "gaussian_smooth_axis: too many planes (%lld)", planes);
  gauss_axis_kernel<<<static_cast<unsigned>(planes), 256, 0, static_cast<cudaStream_t>(s)>>>(
      in, in_idx, out, out_idx, taps, radius, C, D0, D1, D2, axis);
  return hct_check_launch("gauss_axis_kernel");
}

static int g_crop_rows = 1;      // 0: per-voxel gather only (A/B comparison, tools/augment_bench.py)
extern "C" int hct_crop_resize_set_rows(int enable) { g_crop_rows = enable != 0; return HCT_OK; }

extern "C" int hct_crop_resize_area(const void* src, int32_t src_f16, const int32_t* boxes, const float* offsets, float* out, int64_t ncrops,
                                    int32_t C, int32_t S0, int32_t S1, int32_t S2, int32_t T0, int32_t T1, int32_t T2,
                                    hct_stream_t s) {
  HCT_REQUIRE(C > 0 && S0 > 0 && S1 > 0 && S2 > 0 && T0 > 0 && T1 > 0 && T2 > 0, "crop_resize_area: bad shapes");
  if (ncrops <= 0) return HCT_OK;
  const long long planes = ncrops * C * T0;
  HCT_REQUIRE(planes <= 2147483647LL, "crop_resize_area: too many planes (%lld)", planes);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const CropBox* bx = reinterpret_cast<const CropBox*>(boxes);
  const size_t elt = src_f16 ? 2 : 4;
  if (g_crop_rows && S2 % 8 == 0 && S2 <= CROP_ROW_MAX && (reinterpret_cast<uintptr_t>(src) % 16) == 0 && T2 <= 32 * CROP_MAX_K &&
      (static_cast<size_t>(S2) * elt) % 16 == 0) {
    dim3 grid(static_cast<unsigned>(planes));
    if (src_f16) crop_resize_area_rows_kernel<true><<<grid, 256, 0, st>>>(src, bx, offsets, out, C, S0, S1, S2, T0, T1, T2);
    else crop_resize_area_rows_kernel<false><<<grid, 256, 0, st>>>(src, bx, offsets, out, C, S0, S1, S2, T0, T1, T2);
    return hct_check_launch("crop_resize_area_rows_kernel");
  }
  if (src_f16) crop_resize_area_kernel<true><<<static_cast<unsigned>(planes), 256, 0, st>>>(src, bx, offsets, out, C, S0, S1, S2, T0, T1, T2);
  else crop_resize_area_kernel<false><<<static_cast<unsigned>(planes), 256, 0, st>>>(src, bx, offsets, out, C, S0, S1, S2, T0, T1, T2);
  return hct_check_launch("crop_resize_area_kernel");
}

extern "C" int hct_adjust_contrast(float* x, const float* gamma, int32_t* minmax_ws, int64_t nsamples, int64_t per_sample,
                                   hct_stream_t s) {
  HCT_REQUIRE(per_sample > 0 && per_sample % 4 == 0 && nsamples <= 65535, "adjust_contrast: per_sample=%lld nsamples=%lld",
              (long long)per_sample, (long long)nsamples);
  if (nsamples <= 0) return HCT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  minmax_init_kernel<<<static_cast<unsigned>((nsamples + 127) / 128), 128, 0, st>>>(minmax_ws, static_cast<int>(nsamples));
  const dim3 grid(64, static_cast<unsigned>(nsamples));
  sample_minmax_kernel<<<grid, 256, 0, st>>>(x, gamma, minmax_ws, per_sample);
  int rc = hct_check_launch("sample_minmax_kernel");
  if (rc) return rc;
  gamma_kernel<<<grid, 256, 0, st>>>(x, gamma, minmax_ws, per_sample);
  return hct_check_launch("gamma_kernel");
}
