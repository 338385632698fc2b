// spectral_norm.cuh — torch.nn.utils.spectral_norm as libsrk kernels (models/discriminator_swin.py:10,25,49-52,67-69:
// every convolution of UNetDiscriminatorSN is wrapped in it).  Per forward and layer the hook computes, on the weight seen as
// a matrix Wm [U][V] (nn.Conv2d: U = Cout, V = Cin*kh*kw; nn.ConvTranspose2d, dim = 1: U = Cout = second axis, V = Cin*kh*kw):
//     training:  v <- normalize(Wm^T u),  u <- normalize(Wm v)        (one power iteration, in place on the buffers)
//     always:    sigma = u . (Wm v),      W_sn = W / sigma
// and autograd differentiates W / sigma with u and v held constant:
//     dL/dW = dL/dW_sn / sigma  -  (<dL/dW_sn, W> / sigma^2) * (u v^T laid out like W).
// In stock PyTorch this is ~12 small launches and several full passes over the weight per layer (16.8 M parameters in the
// discriminator: the weight-side work is as large as the activation-side work at micro-batch 2).  Here: two passes over W
// forward (Wm^T u, Wm v), 1/sigma folded into the bf16 operand packing, two passes backward.
//
// The weight is addressed as W [A][B][KK] fp32 (KK = kh*kw):
//   dim 0 (Conv2d):          u index = a,  v index = b*KK + k
//   dim 1 (ConvTranspose2d): u index = b,  v index = a*KK + k
#pragma once
#include "srk_ptx.cuh"

namespace srk {

struct SnDims { int A, B, KK, dim; };

__device__ __forceinline__ long long sn_addr(const SnDims& d, int ui, int vi) {
  if (d.dim == 0) return (long long)ui * d.B * d.KK + vi;
  const int a = vi / d.KK, k = vi - a * d.KK;
  return ((long long)a * d.B + ui) * d.KK + k;
}

__device__ __forceinline__ float block_sum_256(float x, float* red /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

// part[split][vi] = sum over the split's range of ui of Wm[ui][vi] * u[ui]   (grid: (ceil(V/256), splits))
static __global__ void __launch_bounds__(256) sn_t_partial_kernel(const float* __restrict__ w, const float* __restrict__ u, SnDims d, int U,
                                                                  int V, float* __restrict__ part) {
  const int vi = blockIdx.x * 256 + threadIdx.x;
  if (vi >= V) return;
  const int chunk = (U + gridDim.y - 1) / gridDim.y;
  const int u0 = blockIdx.y * chunk;
  const int u1 = (u0 + chunk < U) ? u0 + chunk : U;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int ui = u0;
  for (; ui + 4 <= u1; ui += 4) {
    acc0 = fmaf(w[sn_addr(d, ui, vi)], u[ui], acc0);
    acc1 = fmaf(w[sn_addr(d, ui + 1, vi)], u[ui + 1], acc1);
    acc2 = fmaf(w[sn_addr(d, ui + 2, vi)], u[ui + 2], acc2);
    acc3 = fmaf(w[sn_addr(d, ui + 3, vi)], u[ui + 3], acc3);
  }
  for (; ui < u1; ++ui) acc0 = fmaf(w[sn_addr(d, ui, vi)], u[ui], acc0);
  part[(long long)blockIdx.y * V + vi] = (acc0 + acc1) + (acc2 + acc3);
}

// t_raw[vi] = sum_s part[s][vi];  ssq[block] = sum of t_raw^2 over the block's 256 entries      (grid: ceil(V/256))
static __global__ void __launch_bounds__(256) sn_t_reduce_kernel(const float* __restrict__ part, int splits, int V, float* __restrict__ t_raw,
                                                                 float* __restrict__ ssq) {
  __shared__ float red[8];
  const int vi = blockIdx.x * 256 + threadIdx.x;
  float t = 0.f;
  if (vi < V) {
    for (int s = 0; s < splits; ++s) t += part[(long long)s * V + vi];
    t_raw[vi] = t;
  }
  const float ss = block_sum_256(t * t, red);
  if (threadIdx.x == 0) ssq[blockIdx.x] = ss;
}

// s_raw[ui] = sum_vi Wm[ui][vi] * v[vi]     (grid: U blocks of 256)
// t_raw != nullptr (power iteration): v = t_raw / max(||t_raw||, eps) is formed on the fly from the nblk block sums of
// sn_t_reduce_kernel, and block 0 also stores it into the module's weight_v buffer (nobody reads that buffer here).
static __global__ void __launch_bounds__(256) sn_s_rows_kernel(const float* __restrict__ w, float* __restrict__ v, const float* __restrict__ t_raw,
                                                               const float* __restrict__ ssq, int nblk, float eps, SnDims d, int V,
                                                               float* __restrict__ s_raw) {
  __shared__ float red[8];
  const int ui = blockIdx.x;
  float inv = 1.f;
  const float* src = v;
  if (t_raw != nullptr) {
    float p = 0.f;
    for (int i = threadIdx.x; i < nblk; i += 256) p += ssq[i];
    inv = 1.f / fmaxf(sqrtf(block_sum_256(p, red)), eps);
    src = t_raw;
  }
  float acc0 = 0.f, acc1 = 0.f;
  int vi = threadIdx.x;
  for (; vi + 256 < V; vi += 512) {
    acc0 = fmaf(w[sn_addr(d, ui, vi)], src[vi], acc0);
    acc1 = fmaf(w[sn_addr(d, ui, vi + 256)], src[vi + 256], acc1);
  }
  if (vi < V) acc0 = fmaf(w[sn_addr(d, ui, vi)], src[vi], acc0);
  const float t = block_sum_256(acc0 + acc1, red) * inv;
  if (threadIdx.x == 0) s_raw[ui] = t;
  if (t_raw != nullptr && ui == 0)
    for (int i = threadIdx.x; i < V; i += 256) v[i] = t_raw[i] * inv;
}

// update: u = s / max(||s||, eps); sigma = u . s       (one block of 256; U <= a few thousand)
static __global__ void __launch_bounds__(256) sn_s_finish_kernel(const float* __restrict__ s_raw, int U, float eps, int update,
                                                                 float* __restrict__ u, float* __restrict__ sigma) {
  __shared__ float red[8];
  float inv = 1.f;
  if (update) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < U; i += 256) ss = fmaf(s_raw[i], s_raw[i], ss);
    inv = 1.f / fmaxf(sqrtf(block_sum_256(ss, red)), eps);
  }
  float dot = 0.f;
  for (int i = threadIdx.x; i < U; i += 256) {
    float ui = u[i];
    if (update) { ui = s_raw[i] * inv; u[i] = ui; }
    dot = fmaf(ui, s_raw[i], dot);
  }
  dot = block_sum_256(dot, red);
  if (threadIdx.x == 0) *sigma = dot;
}

// out = w / sigma (fp32), for the three 3x3 layers whose kernels take fp32 filters
static __global__ void __launch_bounds__(256) sn_scale_kernel(const float* __restrict__ w, long long n, const float* __restrict__ sigma,
                                                              float* __restrict__ out) {
  const float inv = 1.f / *sigma;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) out[i] = w[i] * inv;
}

// part[block] = sum_i a[i] * b[i]
static __global__ void __launch_bounds__(256) sn_dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                                    float* __restrict__ part) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) acc = fmaf(a[i], b[i], acc);
  acc = block_sum_256(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// dw = dw_sn / sigma - (<dw_sn, W> / sigma^2) * u[ui] v[vi]      (dw may alias dw_sn)
static __global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const float* dw_sn, const float* __restrict__ part, int nparts,
                                                                  const float* __restrict__ sigma, const float* __restrict__ u,
                                                                  const float* __restrict__ v, SnDims d, float* dw) {
  __shared__ float red[8];
  float p = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) p += part[i];
  const float dot = block_sum_256(p, red);
  const float inv = 1.f / *sigma;
  const float coef = dot * inv * inv;
  const long long n = (long long)d.A * d.B * d.KK;
  const int bk = d.B * d.KK;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int a = int(i / bk);
    const int r = int(i - (long long)a * bk);
    const int b = r / d.KK, k = r - b * d.KK;
    const int ui = d.dim == 0 ? a : b;
    const int vi = d.dim == 0 ? r : a * d.KK + k;
    dw[i] = dw_sn[i] * inv - coef * u[ui] * v[vi];
  }
}

}  // namespace srk
