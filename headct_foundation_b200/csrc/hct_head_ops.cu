// Downstream-head kernels: LoRA q/v adapters (the reshape quirk of attentionblock.py:57-59),
// BatchNorm1d over token rows and the attentive-pooling read-out of AttentionClassifier (classifier.py:84-100).
// All of them are HBM-bound passes over [rows, dim] matrices; 16-byte accesses, grids sized from the SM count.
#include "../../include/hct_b200.h"
#include "hct_common.cuh"

namespace {

inline int grid_for(long long work_items, int threads, int max_blocks) {
  long long g = (work_items + threads - 1) / threads;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------ LoRA add / its adjoint
// The reference adds lora_q(x) [B,N,C] to q [B,H,N,hd] after a plain reshape (no head transpose), i.e. element
// f = (h*N + n)*hd + d of the flat per-sample LoRA output lands on q[b, h, n, d].  qkv is our [B,N,3,H,hd] buffer.
// forward : qkv[b,n,{0,2},h,d] += l{q,v}[b].flat[f]          backward: dl{q,v}[b].flat[f] = dqkv[b,n,{0,2},h,d]
__global__ void __launch_bounds__(256)
lora_shuffle_kernel(bf16* __restrict__ qkv, bf16* __restrict__ lq, bf16* __restrict__ lv, long long batch, int seq,
                    int heads, int hd, int backward) {
  const int C = heads * hd, hv = hd >> 3;                       // 8-element (16-byte) groups per head row
  const long long per_sample = static_cast<long long>(seq) * heads * hv;
  const long long total = batch * per_sample;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / per_sample;
    const long long g = i - b * per_sample;                     // group index inside the flat [H, N, hd/8] view
    const int dv = static_cast<int>(g % hv);
    const long long t = g / hv;
    const int n = static_cast<int>(t % seq), h = static_cast<int>(t / seq);
    const long long lo = b * seq * C + g * 8;
    const long long qo = ((b * seq + n) * 3) * C + h * hd + dv * 8;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* l = which ? lv : lq;
      bf16* q = qkv + qo + (which ? 2 * C : 0);
      if (backward) {
        *reinterpret_cast<uint4*>(l + lo) = *reinterpret_cast<const uint4*>(q);
      } else {
        const uint4 a = *reinterpret_cast<const uint4*>(q), d = *reinterpret_cast<const uint4*>(l + lo);
        const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, dd[4] = {d.x, d.y, d.z, d.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 x = unpack_bf16x2(aa[k]), y = unpack_bf16x2(dd[k]);
          o[k] = pack_bf16x2(x.x + y.x, x.y + y.y);
        }
        *reinterpret_cast<uint4*>(q) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ------------------------------------------------------------------ BatchNorm1d(affine=False) over rows
// classifier.py:64-65,89,96: channels = columns, statistics over every row (batch x tokens).
// sums[0:dim] += column sums of a, sums[dim:2dim] += column sums of a*b, where
//   forward  (x_hat == 0): a = x,  b = x            -> sum x, sum x^2
//   backward (x_hat == 1): a = dy, b = (x-mean)*invstd -> sum dy, sum dy*x_hat
template <bool DY_BF16>
__global__ void __launch_bounds__(256)
colnorm_sums_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ invstd, float* __restrict__ sums, long long rows, int dim, int backward) {
  __shared__ float4 red[2][4][64];
  const int cq = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int col = (blockIdx.x * 64 + cq) * 4;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = s1;
  if (col < dim) {
    float4 mu = make_float4(0, 0, 0, 0), is = make_float4(1, 1, 1, 1);
    if (backward) { mu = *reinterpret_cast<const float4*>(mean + col); is = *reinterpret_cast<const float4*>(invstd + col); }
    for (long long r = static_cast<long long>(blockIdx.y) * 4 + rl; r < rows; r += static_cast<long long>(gridDim.y) * 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + r * dim + col);
      float4 a, b;
      if (backward) {
        if (DY_BF16) {
          const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy) + r * dim + col);
          const float2 p = unpack_bf16x2(u.x), q = unpack_bf16x2(u.y);
          a = make_float4(p.x, p.y, q.x, q.y);
        } else {
          a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + r * dim + col);
        }
        b = make_float4((xv.x - mu.x) * is.x, (xv.y - mu.y) * is.y, (xv.z - mu.z) * is.z, (xv.w - mu.w) * is.w);
      } else {
        a = xv; b = xv;
      }
      s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
      s2.x += a.x * b.x; s2.y += a.y * b.y; s2.z += a.z * b.z; s2.w += a.w * b.w;
    }
  }
  red[0][rl][cq] = s1; red[1][rl][cq] = s2;
  __syncthreads();
  if (rl < 2 && col < dim) {
    float4 t = red[rl][0][cq];
#pragma unroll
    for (int w = 1; w < 4; ++w) { const float4 u = red[rl][w][cq]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    float* o = sums + rl * dim + col;
    atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
  }
}

// batch statistics -> mean / invstd (biased variance, as F.batch_norm normalises with) and the running-stat update
// (momentum form of nn.BatchNorm1d: unbiased variance into running_var).
__global__ void colnorm_finalize_kernel(const float* __restrict__ sums, long long rows, int dim, float eps, float momentum,
                                        float* __restrict__ mean, float* __restrict__ invstd,
                                        float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dim) return;
  const float n = static_cast<float>(rows);
  const float mu = sums[c] / n;
  const float var = fmaxf(sums[dim + c] / n - mu * mu, 0.f);
  mean[c] = mu;
  invstd[c] = rsqrtf(var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (rows > 1 ? var * n / (n - 1.f) : var);
}

// eval mode: invstd from the running variance
__global__ void colnorm_invstd_kernel(const float* __restrict__ var, float eps, float* __restrict__ invstd, int dim) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < dim) invstd[c] = rsqrtf(var[c] + eps);
}

// y = (x - mean) * invstd  (bf16 or fp32 out)
__global__ void __launch_bounds__(256)
colnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                     void* __restrict__ y, int y_bf16, long long rows, int dim) {
  const int nv = dim >> 2;
  const long long total = rows * nv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c), is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 o = make_float4((v.x - mu.x) * is.x, (v.y - mu.y) * is.y, (v.z - mu.z) * is.z, (v.w - mu.w) * is.w);
    if (y_bf16) {
      uint2 u; u.x = pack_bf16x2(o.x, o.y); u.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y)[i] = u;
    } else {
      reinterpret_cast<float4*>(y)[i] = o;
    }
  }
}

// dx = invstd * (dy - s1/R - x_hat * s2/R)   (training);   dx = invstd * dy  (eval: sums == NULL)
template <bool DY_BF16>
__global__ void __launch_bounds__(256)
colnorm_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ invstd, const float* __restrict__ sums, float* __restrict__ dx,
                   long long rows, int dim) {
  const int nv = dim >> 2;
  const long long total = rows * nv;
  const float rn = 1.f / static_cast<float>(rows);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    float4 d;
    if (DY_BF16) {
      const uint2 u = reinterpret_cast<const uint2*>(dy)[i];
      const float2 p = unpack_bf16x2(u.x), q = unpack_bf16x2(u.y);
      d = make_float4(p.x, p.y, q.x, q.y);
    } else {
      d = reinterpret_cast<const float4*>(dy)[i];
    }
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    float4 o;
    if (sums) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c);
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(sums) + c), s2 = __ldg(reinterpret_cast<const float4*>(sums + dim) + c);
      o = make_float4(is.x * (d.x - s1.x * rn - (v.x - mu.x) * is.x * s2.x * rn), is.y * (d.y - s1.y * rn - (v.y - mu.y) * is.y * s2.y * rn),
                      is.z * (d.z - s1.z * rn - (v.z - mu.z) * is.z * s2.z * rn), is.w * (d.w - s1.w * rn - (v.w - mu.w) * is.w * s2.w * rn));
    } else {
      o = make_float4(is.x * d.x, is.y * d.y, is.z * d.z, is.w * d.w);
    }
    reinterpret_cast<float4*>(dx)[i] = o;
  }
}

// ------------------------------------------------------------------ attentive pooling (classifier.py:84-94)
// A handful of learned queries (cls_token, shared by the whole batch) attend over the N tokens of each sample.
// kv = wkv(bn1(x)) in our [B, N, 2, H, hd] bf16 layout.  q_eff = cls_token * scale_total where
// scale_total = self.scale * (1/sqrt(hd)): the reference scales q itself AND lets SDPA apply its default scale.
// One CTA per (head, sample); N * nq <= POOL_MAX_SCORES scores live in shared memory.
constexpr int POOL_MAX_SCORES = 8192;
constexpr int POOL_MAX_Q = 8;
constexpr int POOL_THREADS = 128;

__device__ __forceinline__ float dot8(const uint4 u, const float* __restrict__ q) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  return (a.x * q[0] + a.y * q[1]) + (b.x * q[2] + b.y * q[3]) + (c.x * q[4] + c.y * q[5]) + (d.x * q[6] + d.y * q[7]);
}

__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__global__ void __launch_bounds__(POOL_THREADS)
pool_attn_fwd_kernel(const float* __restrict__ cls, const bf16* __restrict__ kv, float* __restrict__ out,
                     float* __restrict__ probs, int seq, int heads, int hd, int nq, float scale_total) {
  __shared__ float s[POOL_MAX_SCORES];
  __shared__ float qs[POOL_MAX_Q * 128];
  __shared__ float red[33];
  __shared__ float part[POOL_THREADS];
  const int h = blockIdx.x, b = blockIdx.y, C = heads * hd, tid = threadIdx.x;
  for (int i = tid; i < nq * hd; i += POOL_THREADS) qs[i] = cls[(i / hd) * C + h * hd + (i % hd)] * scale_total;
  __syncthreads();
  const bf16* kbase = kv + (static_cast<long long>(b) * seq * 2) * C + h * hd;
  for (int n = tid; n < seq; n += POOL_THREADS) {
    const bf16* kr = kbase + static_cast<long long>(n) * 2 * C;
    float acc[POOL_MAX_Q];
#pragma unroll
    for (int q = 0; q < POOL_MAX_Q; ++q) acc[q] = 0.f;
    for (int d = 0; d < hd; d += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(kr + d);
#pragma unroll
      for (int q = 0; q < POOL_MAX_Q; ++q) if (q < nq) acc[q] += dot8(u, qs + q * hd + d);
    }
#pragma unroll
    for (int q = 0; q < POOL_MAX_Q; ++q) if (q < nq) s[q * seq + n] = acc[q];
  }
  __syncthreads();
  for (int q = 0; q < nq; ++q) {
    float m = -INFINITY;
    for (int n = tid; n < seq; n += POOL_THREADS) m = fmaxf(m, s[q * seq + n]);
    m = block_max(m, red);
    float z = 0.f;
    for (int n = tid; n < seq; n += POOL_THREADS) { const float e = __expf(s[q * seq + n] - m); s[q * seq + n] = e; z += e; }
    z = block_sum(z, red);
    const float rz = 1.f / z;
    float* pr = probs + ((static_cast<long long>(b) * heads + h) * nq + q) * seq;
    for (int n = tid; n < seq; n += POOL_THREADS) { const float p = s[q * seq + n] * rz; s[q * seq + n] = p; pr[n] = p; }
    __syncthreads();
    // out[d] = sum_n p[n] * v[n, d]: thread = (row group, d); consecutive threads read consecutive d
    const int groups = POOL_THREADS / hd, d = tid % hd, gsel = tid / hd;
    float acc = 0.f;
    if (gsel < groups) {
      const bf16* vbase = kbase + C + d;
      for (int n = gsel; n < seq; n += groups) acc += s[q * seq + n] * __bfloat162float(vbase[static_cast<long long>(n) * 2 * C]);
    }
    part[tid] = acc;
    __syncthreads();
    if (tid < hd) {
      float t = 0.f;
      for (int g = 0; g < groups; ++g) t += part[g * hd + tid];
      out[(static_cast<long long>(b) * nq + q) * C + h * hd + tid] = t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(POOL_THREADS)
pool_attn_bwd_kernel(const float* __restrict__ cls, const bf16* __restrict__ kv, const float* __restrict__ out,
                     const float* __restrict__ probs, const float* __restrict__ dout, float* __restrict__ dcls,
                     bf16* __restrict__ dkv, int seq, int heads, int hd, int nq, float scale_total) {
  __shared__ float ds[POOL_MAX_SCORES];
  __shared__ float qs[POOL_MAX_Q * 128];
  __shared__ float dos[POOL_MAX_Q * 128];
  __shared__ float delta[POOL_MAX_Q];
  __shared__ float part[POOL_THREADS];
  const int h = blockIdx.x, b = blockIdx.y, C = heads * hd, tid = threadIdx.x;
  for (int i = tid; i < nq * hd; i += POOL_THREADS) {
    const int q = i / hd, d = i % hd;
    qs[i] = cls[q * C + h * hd + d] * scale_total;
    dos[i] = dout[(static_cast<long long>(b) * nq + q) * C + h * hd + d];
  }
  __syncthreads();
  if (tid < nq) {                                  // delta_q = sum_n p dp = <dout_q, out_q> over this head's slice
    float t = 0.f;
    for (int d = 0; d < hd; ++d) t += dos[tid * hd + d] * out[(static_cast<long long>(b) * nq + tid) * C + h * hd + d];
    delta[tid] = t;
  }
  __syncthreads();
  const bf16* kbase = kv + (static_cast<long long>(b) * seq * 2) * C + h * hd;
  bf16* dkbase = dkv + (static_cast<long long>(b) * seq * 2) * C + h * hd;
  const float* pr = probs + (static_cast<long long>(b) * heads + h) * nq * seq;
  for (int n = tid; n < seq; n += POOL_THREADS) {
    const bf16* vr = kbase + static_cast<long long>(n) * 2 * C + C;
    float dp[POOL_MAX_Q], p[POOL_MAX_Q];
#pragma unroll
    for (int q = 0; q < POOL_MAX_Q; ++q) { dp[q] = 0.f; p[q] = q < nq ? pr[q * seq + n] : 0.f; }
    for (int d = 0; d < hd; d += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(vr + d);
#pragma unroll
      for (int q = 0; q < POOL_MAX_Q; ++q) if (q < nq) dp[q] += dot8(u, dos + q * hd + d);
    }
#pragma unroll
    for (int q = 0; q < POOL_MAX_Q; ++q) if (q < nq) { dp[q] = p[q] * (dp[q] - delta[q]); ds[q * seq + n] = dp[q]; }   // dp now holds dS
    bf16* dkr = dkbase + static_cast<long long>(n) * 2 * C;
    for (int d = 0; d < hd; d += 8) {
      float k8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, v8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < POOL_MAX_Q; ++q) if (q < nq) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { k8[j] += dp[q] * qs[q * hd + d + j]; v8[j] += p[q] * dos[q * hd + d + j]; }
      }
      *reinterpret_cast<uint4*>(dkr + d) = make_uint4(pack_bf16x2(k8[0], k8[1]), pack_bf16x2(k8[2], k8[3]),
                                                      pack_bf16x2(k8[4], k8[5]), pack_bf16x2(k8[6], k8[7]));
      *reinterpret_cast<uint4*>(dkr + C + d) = make_uint4(pack_bf16x2(v8[0], v8[1]), pack_bf16x2(v8[2], v8[3]),
                                                          pack_bf16x2(v8[4], v8[5]), pack_bf16x2(v8[6], v8[7]));
    }
  }
  __syncthreads();
  // dcls[q, h*hd + d] += scale_total * sum_n dS[q, n] * k[n, d]   (summed over the batch with atomics)
  const int groups = POOL_THREADS / hd, d = tid % hd, gsel = tid / hd;
  for (int q = 0; q < nq; ++q) {
    float acc = 0.f;
    if (gsel < groups)
      for (int n = gsel; n < seq; n += groups) acc += ds[q * seq + n] * __bfloat162float(kbase[static_cast<long long>(n) * 2 * C + d]);
    part[tid] = acc;
    __syncthreads();
    if (tid < hd) {
      float t = 0.f;
      for (int g = 0; g < groups; ++g) t += part[g * hd + tid];
      atomicAdd(dcls + q * C + h * hd + tid, t * scale_total);
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int hct_lora_shuffle(void* qkv, void* lq, void* lv, int64_t batch, int32_t seq, int32_t heads, int32_t head_dim,
                                int32_t backward, hct_stream_t s) {
  HCT_REQUIRE(batch >= 0 && seq > 0 && heads > 0 && head_dim > 0 && head_dim % 8 == 0,
              "lora_shuffle: head_dim=%d must be a multiple of 8", head_dim);
  if (batch == 0) return HCT_OK;
  const long long total = batch * seq * heads * (head_dim / 8);
  lora_shuffle_kernel<<<grid_for(total, 256, hct_num_sms() * 8), 256, 0, static_cast<cudaStream_t>(s)>>>(
      static_cast<bf16*>(qkv), static_cast<bf16*>(lq), static_cast<bf16*>(lv), batch, seq, heads, head_dim, backward);
  return hct_check_launch("lora_shuffle_kernel");
}

extern "C" int hct_colnorm_stats(const float* x, float* sums, int64_t rows, int32_t dim, float eps, float momentum,
                                 float* mean, float* invstd, float* running_mean, float* running_var, hct_stream_t s) {
  HCT_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0, "colnorm_stats: rows=%lld dim=%d unsupported", (long long)rows, dim);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  dim3 grid((dim + 255) / 256, grid_for(rows, 4 * 16, 2 * hct_num_sms()));
  colnorm_sums_kernel<false><<<grid, 256, 0, st>>>(nullptr, x, nullptr, nullptr, sums, rows, dim, 0);
  int rc = hct_check_launch("colnorm_sums_kernel");
  if (rc) return rc;
  colnorm_finalize_kernel<<<(dim + 255) / 256, 256, 0, st>>>(sums, rows, dim, eps, momentum, mean, invstd, running_mean, running_var);
  return hct_check_launch("colnorm_finalize_kernel");
}

extern "C" int hct_colnorm_apply(const float* x, const float* mean, const float* invstd_or_var, int32_t is_var, float eps,
                                 float* invstd_scratch, void* y, int32_t y_bf16, int64_t rows, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(rows >= 0 && dim > 0 && dim % 4 == 0, "colnorm_apply: dim=%d unsupported", dim);
  HCT_REQUIRE(!is_var || invstd_scratch != nullptr, "colnorm_apply: running-variance mode needs invstd_scratch");
  if (rows == 0) return HCT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const float* invstd = invstd_or_var;
  if (is_var) {
    colnorm_invstd_kernel<<<(dim + 255) / 256, 256, 0, st>>>(invstd_or_var, eps, invstd_scratch, dim);
    int rc = hct_check_launch("colnorm_invstd_kernel");
    if (rc) return rc;
    invstd = invstd_scratch;
  }
  colnorm_apply_kernel<<<grid_for(rows * (dim / 4), 256, hct_num_sms() * 8), 256, 0, st>>>(x, mean, invstd, y, y_bf16, rows, dim);
  return hct_check_launch("colnorm_apply_kernel");
}

extern "C" int hct_colnorm_bwd(const void* dy, int32_t dy_bf16, const float* x, const float* mean, const float* invstd,
                               float* sums, float* dx, int64_t rows, int32_t dim, hct_stream_t s) {
  HCT_REQUIRE(rows >= 0 && dim > 0 && dim % 4 == 0, "colnorm_bwd: dim=%d unsupported", dim);
  if (rows == 0) return HCT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(s);
  if (sums) {
    dim3 grid((dim + 255) / 256, grid_for(rows, 4 * 16, 2 * hct_num_sms()));
    if (dy_bf16) colnorm_sums_kernel<true><<<grid, 256, 0, st>>>(dy, x, mean, invstd, sums, rows, dim, 1);
    else colnorm_sums_kernel<false><<<grid, 256, 0, st>>>(dy, x, mean, invstd, sums, rows, dim, 1);
    int rc = hct_check_launch("colnorm_sums_kernel");
    if (rc) return rc;
  }
  const int g = grid_for(rows * (dim / 4), 256, hct_num_sms() * 8);
  if (dy_bf16) colnorm_bwd_kernel<true><<<g, 256, 0, st>>>(dy, x, mean, invstd, sums, dx, rows, dim);
  else colnorm_bwd_kernel<false><<<g, 256, 0, st>>>(dy, x, mean, invstd, sums, dx, rows, dim);
  return hct_check_launch("colnorm_bwd_kernel");
}

extern "C" int hct_pool_attention_fwd(const float* cls, const void* kv, float* out, float* probs, int32_t batch, int32_t seq,
                                      int32_t heads, int32_t head_dim, int32_t num_queries, float scale_total, hct_stream_t s) {
  HCT_REQUIRE(head_dim % 8 == 0 && head_dim <= 128 && num_queries >= 1 && num_queries <= POOL_MAX_Q &&
                  static_cast<long long>(seq) * num_queries <= POOL_MAX_SCORES && seq > 0,
              "pool_attention: seq=%d queries=%d head_dim=%d unsupported", seq, num_queries, head_dim);
  if (batch == 0) return HCT_OK;
  pool_attn_fwd_kernel<<<dim3(heads, batch), POOL_THREADS, 0, static_cast<cudaStream_t>(s)>>>(
      cls, static_cast<const bf16*>(kv), out, probs, seq, heads, head_dim, num_queries, scale_total);
  return hct_check_launch("pool_attn_fwd_kernel");
}

extern "C" int hct_pool_attention_bwd(const float* cls, const void* kv, const float* out, const float* probs, const float* dout,
                                      float* dcls, void* dkv, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim,
                                      int32_t num_queries, float scale_total, hct_stream_t s) {
  HCT_REQUIRE(head_dim % 8 == 0 && head_dim <= 128 && num_queries >= 1 && num_queries <= POOL_MAX_Q &&
                  static_cast<long long>(seq) * num_queries <= POOL_MAX_SCORES && seq > 0,
              "pool_attention: seq=%d queries=%d head_dim=%d unsupported", seq, num_queries, head_dim);
  if (batch == 0) return HCT_OK;
  pool_attn_bwd_kernel<<<dim3(heads, batch), POOL_THREADS, 0, static_cast<cudaStream_t>(s)>>>(
      cls, static_cast<const bf16*>(kv), out, probs, dout, dcls, static_cast<bf16*>(dkv), seq, heads, head_dim, num_queries,
      scale_total);
  return hct_check_launch("pool_attn_bwd_kernel");
}
