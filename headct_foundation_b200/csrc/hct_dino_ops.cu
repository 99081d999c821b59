// DINO-specific HBM-bound kernels (L2-normalise, weight-norm, centered/sharpened softmax CE, center EMA,
// multi-tensor teacher EMA) and the multi-tensor train-step glue (per-parameter clip + AdamW).
#include <math.h>

#include "../../include/hct_b200.h"
#include "hct_common.cuh"

namespace {

inline int grid_for(long long work_items, int threads, int max_blocks) {
  long long g = (work_items + threads - 1) / threads;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------ F.normalize(p=2) fwd / bwd (warp per row)
__global__ void l2norm_fwd_kernel(const void* __restrict__ x, int x_bf16, bf16* __restrict__ y, float* __restrict__ inv_norm,
                                  long long rows, int dim) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float v = x_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(x)[row * dim + c])
                           : reinterpret_cast<const float*>(x)[row * dim + c];
    ss += v * v;
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  for (int c = lane; c < dim; c += 32) {
    const float v = x_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(x)[row * dim + c])
                           : reinterpret_cast<const float*>(x)[row * dim + c];
    y[row * dim + c] = __float2bfloat16_rn(v * inv);
  }
}
__global__ void l2norm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, const float* __restrict__ inv_norm,
                                  bf16* __restrict__ dx, long long rows, int dim) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float dot = 0.f;
  for (int c = lane; c < dim; c += 32) dot += __bfloat162float(dy[row * dim + c]) * __bfloat162float(y[row * dim + c]);
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  for (int c = lane; c < dim; c += 32) {
    const float g = __bfloat162float(dy[row * dim + c]), yy = __bfloat162float(y[row * dim + c]);
    dx[row * dim + c] = __float2bfloat16_rn(inv * (g - yy * dot));
  }
}

// ------------------------------------------------------------------ weight_norm fwd / bwd (warp per prototype row)
__global__ void weightnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, bf16* __restrict__ w,
                                      float* __restrict__ inv_norm, long long rows, int dim) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) { const float t = v[row * dim + c]; ss += t * t; }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  const float sc = g[row] * inv;
  for (int c = lane; c < dim; c += 32) w[row * dim + c] = __float2bfloat16_rn(v[row * dim + c] * sc);
}
__global__ void weightnorm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                                      const float* __restrict__ inv_norm, float* __restrict__ dv, float* __restrict__ dg,
                                      long long rows, int dim) {
  const int lane = threadIdx.x & 31;
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int c = lane; c < dim; c += 32) dot += dw[row * dim + c] * v[row * dim + c] * inv;
  dot = warp_sum(dot);
  if (dg != nullptr && lane == 0) dg[row] = dot;          // d/dg of g * v / ||v||  (norm_last_layer=False trains the gain)
  if (dv == nullptr) return;
  const float sc = g[row] * inv;
  for (int c = lane; c < dim; c += 32) dv[row * dim + c] = sc * (dw[row * dim + c] - dot * v[row * dim + c] * inv);
}

// ------------------------------------------------------------------ DINO loss
// stats[row] = {max, logsumexp} of the temperature-scaled (and centered, for teacher rows) logits.
// rows [0, 2B) are teacher rows, rows [2B, 2B + ncrops*B) student rows.
__global__ void __launch_bounds__(256)
dino_stats_kernel(const float* __restrict__ student, const float* __restrict__ teacher, const float* __restrict__ center,
                  float* __restrict__ stats, int B, int K, float inv_ts, float inv_tt) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const bool is_teacher = row < 2 * B;
  const float* src = is_teacher ? teacher + static_cast<long long>(row) * K : student + static_cast<long long>(row - 2 * B) * K;
  const float sc = is_teacher ? inv_tt : inv_ts;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 4; c < K; c += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(src + c);
    if (is_teacher) { const float4 ce = *reinterpret_cast<const float4*>(center + c); v.x -= ce.x; v.y -= ce.y; v.z -= ce.z; v.w -= ce.w; }
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) * sc);
  }
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  float se = 0.f;
  for (int c = threadIdx.x * 4; c < K; c += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(src + c);
    if (is_teacher) { const float4 ce = *reinterpret_cast<const float4*>(center + c); v.x -= ce.x; v.y -= ce.y; v.z -= ce.z; v.w -= ce.w; }
    se += __expf(v.x * sc - mx) + __expf(v.y * sc - mx) + __expf(v.z * sc - mx) + __expf(v.w * sc - mx);
  }
  se = block_sum(se, red);
  if (threadIdx.x == 0) { stats[2 * row] = mx; stats[2 * row + 1] = mx + logf(se); }
}

// grid (K / (256*4), B).  Each CTA covers 1024 columns of one sample b: builds q0,q1 (teacher probs),
// accumulates the cross terms and (optionally) writes the student gradient.
constexpr int DINO_MAX_CROPS = 12;
__global__ void __launch_bounds__(256)
dino_loss_kernel(const float* __restrict__ student, const float* __restrict__ teacher, const float* __restrict__ center,
                 const float* __restrict__ stats, float* __restrict__ loss_out, bf16* __restrict__ dstudent,
                 const float* __restrict__ dloss, int B, int ncrops, int K, float inv_ts, float inv_tt) {
  __shared__ float red[33];
  const int b = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n_terms = 2 * (ncrops - 1);
  const float w_loss = 1.f / (static_cast<float>(n_terms) * B);
  const float w_grad = w_loss * (dloss != nullptr ? dloss[0] : 1.f);
  float acc = 0.f;
  if (c < K) {
    float q[2][4];
    const float4 ce = *reinterpret_cast<const float4*>(center + c);
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      const int row = iq * B + b;
      const float lse = stats[2 * row + 1];
      const float4 t = *reinterpret_cast<const float4*>(teacher + static_cast<long long>(row) * K + c);
      q[iq][0] = __expf((t.x - ce.x) * inv_tt - lse); q[iq][1] = __expf((t.y - ce.y) * inv_tt - lse);
      q[iq][2] = __expf((t.z - ce.z) * inv_tt - lse); q[iq][3] = __expf((t.w - ce.w) * inv_tt - lse);
    }
    for (int v = 0; v < ncrops; ++v) {
      const int srow = v * B + b;
      const float lse = stats[2 * (2 * B + srow) + 1];
      const float4 s4 = *reinterpret_cast<const float4*>(student + static_cast<long long>(srow) * K + c);
      const float s[4] = {s4.x * inv_ts, s4.y * inv_ts, s4.z * inv_ts, s4.w * inv_ts};
      float qs[4];   // sum of teacher probs over views != v
      float cnt = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) qs[k] = 0.f;
#pragma unroll
      for (int iq = 0; iq < 2; ++iq) {
        if (iq != v) {
          cnt += 1.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) qs[k] += q[iq][k];
        }
      }
      // sum_k q (lse - s')  ; the lse part is added once per (iq,v) pair below via sum(q)=1 -> use exact form
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += qs[k] * (lse - s[k]);
      if (dstudent != nullptr) {
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] = w_grad * inv_ts * (cnt * __expf(s[k] - lse) - qs[k]);
        uint2 u; u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]);
        *reinterpret_cast<uint2*>(dstudent + static_cast<long long>(srow) * K + c) = u;
      }
    }
  }
  if (loss_out != nullptr) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss_out, acc * w_loss);
  }
}

__global__ void center_ema_kernel(float* __restrict__ center, const float* __restrict__ bc, float inv_denom, float m, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) center[i] = center[i] * m + bc[i] * inv_denom * (1.f - m);
}

// ------------------------------------------------------------------ multi-tensor kernels (grid.y = tensor index)
__global__ void ema_multi_kernel(const long long* __restrict__ table, float m) {
  const long long* e = table + 4LL * blockIdx.y;
  float* pk = reinterpret_cast<float*>(e[0]);
  const float* pq = reinterpret_cast<const float*>(e[1]);
  const long long n = e[2];
  bf16* sh = reinterpret_cast<bf16*>(e[3]);              // bf16 shadow of the teacher tensor (GEMM operand copy) or null
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const float om = 1.f - m;
  if (((reinterpret_cast<uintptr_t>(pk) | reinterpret_cast<uintptr_t>(pq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(sh) & 7) == 0) {
    const long long n4 = n >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 a = reinterpret_cast<float4*>(pk)[i];
      const float4 q = reinterpret_cast<const float4*>(pq)[i];
      // param_k.mul_(m).add_((1 - m) * param_q)   (misc.py:397)
      a.x = a.x * m + om * q.x; a.y = a.y * m + om * q.y; a.z = a.z * m + om * q.z; a.w = a.w * m + om * q.w;
      reinterpret_cast<float4*>(pk)[i] = a;
      if (sh != nullptr) {
        uint2 u; u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
        reinterpret_cast<uint2*>(sh)[i] = u;
      }
    }
    for (long long i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float a = pk[i] * m + om * pq[i];
      pk[i] = a;
      if (sh != nullptr) sh[i] = __float2bfloat16(a);
    }
  } else {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float a = pk[i] * m + om * pq[i];
      pk[i] = a;
      if (sh != nullptr) sh[i] = __float2bfloat16(a);
    }
  }
}

__global__ void grad_sqnorm_multi_kernel(const long long* __restrict__ table, float* __restrict__ norms) {
  __shared__ float red[33];
  const long long* e = table + 7LL * blockIdx.y;
  const float* g = reinterpret_cast<const float*>(e[1]);
  const long long n = e[4];
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) { const float v = g[i]; s += v * v; }
  s = block_sum(s, red);
  if (threadIdx.x == 0 && s != 0.f) atomicAdd(norms + blockIdx.y, s);
}

__global__ void adamw_multi_kernel(const long long* __restrict__ table, const float* __restrict__ sqnorms, float clip,
                                   float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                   float step, const float* __restrict__ hyper) {
  if (hyper != nullptr) {      // per-step scalars from device memory: the launch can live in a CUDA graph
    lr = hyper[0]; wd = hyper[1]; bc1 = hyper[2]; bc2_sqrt = hyper[3]; step = hyper[4];
  }
  const long long* e = table + 7LL * blockIdx.y;
  // torch.optim.AdamW bias-corrects every parameter with ITS OWN step count.  Column 6 holds how many steps this tensor
  // is behind the table's first tensor (a tensor whose gradient was None for a while -- cancel_gradients_last_layer,
  // misc.py:366-371 -- joins late with zero moments): bc1 / bc2_sqrt above are the first tensor's, recomputed here otherwise.
  const long long behind = e[6];
  if (behind != 0) {
    const float s = step - static_cast<float>(behind);
    bc1 = 1.f - powf(beta1, s);
    bc2_sqrt = sqrtf(1.f - powf(beta2, s));
  }
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  float* m = reinterpret_cast<float*>(e[2]);
  float* v = reinterpret_cast<float*>(e[3]);
  const long long n = e[4];
  bf16* sh = reinterpret_cast<bf16*>(e[5]);              // bf16 shadow of the parameter (GEMM operand copy) or null
  float coef = 1.f;
  if (clip > 0.f) {
    // clip_coef = clip / (||g|| + 1e-6), applied per tensor when < 1 (misc.py:379-382)
    const float c = clip / (sqrtf(sqnorms[blockIdx.y]) + 1e-6f);
    if (c < 1.f) coef = c;
  }
  const float step_size = lr / bc1;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  // 16-byte accesses on the four fp32 streams (8-byte on the bf16 copy) where the tensor allows it; the scalar loop below
  // takes the remainder (and everything, for a tensor that is not 16-byte aligned)
  long long done4 = 0;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(sh) & 7) == 0) {
    const long long n4 = n >> 2;
    const float decay = 1.f - lr * wd, ob1 = 1.f - beta1, ob2 = 1.f - beta2;     // same arithmetic as the scalar loop
    for (long long i = tid; i < n4; i += nthreads) {
      const float4 g4 = reinterpret_cast<const float4*>(g)[i];
      float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
      float pe[4] = {p4.x, p4.y, p4.z, p4.w}, me[4] = {m4.x, m4.y, m4.z, m4.w}, ve[4] = {v4.x, v4.y, v4.z, v4.w};
      const float ge[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float pi = pe[k] * decay;
        me[k] = beta1 * me[k] + ob1 * ge[k];
        ve[k] = beta2 * ve[k] + ob2 * ge[k] * ge[k];
        pi -= step_size * me[k] / (sqrtf(ve[k]) / bc2_sqrt + eps);
        pe[k] = pi;
      }
      reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
      reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
      reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
      if (sh != nullptr) {
        uint2 u; u.x = pack_bf16x2(pe[0], pe[1]); u.y = pack_bf16x2(pe[2], pe[3]);
        reinterpret_cast<uint2*>(sh)[i] = u;
      }
    }
    done4 = n4 << 2;
  }
  for (long long i = done4 + tid; i < n; i += nthreads) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    pi -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
    if (sh != nullptr) sh[i] = __float2bfloat16(pi);     // the next forward's GEMM operand, refreshed in the same pass
  }
}

}  // namespace

extern "C" int hct_l2norm_fwd(const void* x, int x_bf16, void* y, float* inv_norm, int64_t rows, int32_t dim, hct_stream_t s) {
  if (rows <= 0) return HCT_OK;
  l2norm_fwd_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(s)>>>(
      x, x_bf16, static_cast<bf16*>(y), inv_norm, rows, dim);
  return hct_check_launch("l2norm_fwd_kernel");
}
extern "C" int hct_l2norm_bwd(const void* dy, const void* y, const float* inv_norm, void* dx, int64_t rows, int32_t dim,
                              hct_stream_t s) {
  if (rows <= 0) return HCT_OK;
  l2norm_bwd_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(s)>>>(
      static_cast<const bf16*>(dy), static_cast<const bf16*>(y), inv_norm, static_cast<bf16*>(dx), rows, dim);
  return hct_check_launch("l2norm_bwd_kernel");
}
extern "C" int hct_weightnorm_fwd(const float* v, const float* g, void* w, float* inv_norm, int64_t rows, int32_t dim,
                                  hct_stream_t s) {
  if (rows <= 0) return HCT_OK;
  weightnorm_fwd_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(s)>>>(
      v, g, static_cast<bf16*>(w), inv_norm, rows, dim);
  return hct_check_launch("weightnorm_fwd_kernel");
}
extern "C" int hct_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv,
                                  float* dg, int64_t rows, int32_t dim, hct_stream_t s) {
  if (rows <= 0) return HCT_OK;
  weightnorm_bwd_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(s)>>>(
      dw, v, g, inv_norm, dv, dg, rows, dim);
  return hct_check_launch("weightnorm_bwd_kernel");
}

extern "C" int hct_dino_loss(const float* student, const float* teacher, const float* center, float* loss_out,
                             float* stats_ws, void* dstudent, const float* dloss, int32_t B, int32_t ncrops, int32_t K,
                             float student_temp, float teacher_temp, hct_stream_t s) {
  HCT_REQUIRE(B > 0 && ncrops >= 2 && ncrops <= DINO_MAX_CROPS && K > 0 && K % 4 == 0, "dino_loss: B=%d ncrops=%d K=%d", B,
              ncrops, K);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  const float inv_ts = 1.f / student_temp, inv_tt = 1.f / teacher_temp;
  dino_stats_kernel<<<(2 + ncrops) * B, 256, 0, st>>>(student, teacher, center, stats_ws, B, K, inv_ts, inv_tt);
  int rc = hct_check_launch("dino_stats_kernel");
  if (rc) return rc;
  dim3 grid((K + 1023) / 1024, B);
  dino_loss_kernel<<<grid, 256, 0, st>>>(student, teacher, center, stats_ws, loss_out, static_cast<bf16*>(dstudent), dloss,
                                         B, ncrops, K, inv_ts, inv_tt);
  return hct_check_launch("dino_loss_kernel");
}

extern "C" int hct_center_ema(float* center, const float* batch_center_sum, float denom, float momentum, int32_t K,
                              hct_stream_t s) {
  if (K <= 0) return HCT_OK;
  center_ema_kernel<<<(K + 255) / 256, 256, 0, static_cast<cudaStream_t>(s)>>>(center, batch_center_sum, 1.f / denom,
                                                                              momentum, K);
  return hct_check_launch("center_ema_kernel");
}

extern "C" int hct_ema_multi(const int64_t* table, int32_t n, float m, hct_stream_t s) {
  if (n <= 0) return HCT_OK;
  HCT_REQUIRE(n <= 65535, "ema_multi: too many tensors (%d)", n);
  ema_multi_kernel<<<dim3(64, n), 256, 0, static_cast<cudaStream_t>(s)>>>(reinterpret_cast<const long long*>(table), m);
  return hct_check_launch("ema_multi_kernel");
}

extern "C" int hct_grad_norms_multi(const int64_t* table, int32_t n, float* norms_ws, hct_stream_t s) {
  if (n <= 0) return HCT_OK;
  HCT_REQUIRE(n <= 65535, "grad_norms_multi: too many tensors (%d)", n);
  cudaStream_t st = static_cast<cudaStream_t>(s);
  HctProfScope prof(st, HCT_PROF_ADAMW, 0.0);            // bytes are unknown here (the table lives on the device): the host adds them
  cudaError_t e = cudaMemsetAsync(norms_ws, 0, sizeof(float) * n, st);
  if (e != cudaSuccess) { hct_set_error("grad_norms_multi memset: %s", cudaGetErrorString(e)); return HCT_ERR_CUDA; }
  grad_sqnorm_multi_kernel<<<dim3(32, n), 256, 0, st>>>(reinterpret_cast<const long long*>(table), norms_ws);
  return hct_check_launch("grad_sqnorm_multi_kernel");
}

extern "C" int hct_adamw_multi(const int64_t* table, int32_t n, const float* norms_ws, float clip, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int32_t step, hct_stream_t s) {
  if (n <= 0) return HCT_OK;
  HCT_REQUIRE(n <= 65535 && step >= 1, "adamw_multi: n=%d step=%d", n, step);
  HCT_REQUIRE(clip <= 0.f || norms_ws != nullptr, "adamw_multi: clip > 0 needs norms_ws");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  HctProfScope prof(static_cast<cudaStream_t>(s), HCT_PROF_ADAMW, 0.0);
  adamw_multi_kernel<<<dim3(64, n), 256, 0, static_cast<cudaStream_t>(s)>>>(
      reinterpret_cast<const long long*>(table), norms_ws, clip, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2),
      static_cast<float>(step), nullptr);
  return hct_check_launch("adamw_multi_kernel");
}

extern "C" int hct_adamw_multi_dev(const int64_t* table, int32_t n, const float* norms_ws, float clip, const float* hyper,
                                   float beta1, float beta2, float eps, hct_stream_t s) {
  if (n <= 0) return HCT_OK;
  HCT_REQUIRE(n <= 65535 && hyper != nullptr, "adamw_multi_dev: n=%d hyper=%p", n, (const void*)hyper);
  HCT_REQUIRE(clip <= 0.f || norms_ws != nullptr, "adamw_multi_dev: clip > 0 needs norms_ws");
  adamw_multi_kernel<<<dim3(64, n), 256, 0, static_cast<cudaStream_t>(s)>>>(
      reinterpret_cast<const long long*>(table), norms_ws, clip, 0.f, beta1, beta2, eps, 0.f, 1.f, 1.f, 1.f, hyper);
  return hct_check_launch("adamw_multi_kernel");
}
