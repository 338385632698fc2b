// conv_aux.cuh — memory-bound helpers around the implicit-GEMM convolution:
//   weight packing / gradient unpacking (incl. the PixelShuffle channel permutation), bias gradients,
//   and the two degenerate convolutions of the SR head/tail that have a single input or output channel
//   (conv_first 1->C and conv_last 64->1: no GEMM shape to speak of; they are pure HBM streams).
// Reference: models/architecture_swin.py:202 (conv_first), :230 (conv_last), :175-190 (Upsample).
#pragma once
#include "srk_ptx.cuh"

namespace srk {

// PixelShuffle(2) channel permutation: conv output channel n = c*4 + i*2 + j  <->  packed n' = (i*2+j)*64 + c
__device__ __forceinline__ int ps_pack(int n, int cgrp) { return (n & 3) * cgrp + (n >> 2); }

// W[Cout][Cin][3][3] fp32 -> Wf [Cout_p][9*Cin_p] (k = tap*Cin_p + ci) and Wt [Cin_p][9*Cout_p]
// (flipped taps, for the input gradient).  ps: permute output channels for a fused PixelShuffle(2).
static __global__ void conv_prep_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                         __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wt,
                                         float* __restrict__ bias_packed, int Cout, int Cin, int Cout_p, int Cin_p,
                                         int ps) {
  const int total = Cout_p * 9 * Cin_p;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int np = idx / (9 * Cin_p), rem = idx % (9 * Cin_p), tap = rem / Cin_p, ci = rem % Cin_p;
    // which real output channel lives in packed row np?
    int n = np;
    if (ps) { const int grp = np / (Cout_p / 4), c = np % (Cout_p / 4); n = c * 4 + grp; }
    float v = 0.f;
    if (n < Cout && ci < Cin) v = w[(size_t(n) * Cin + ci) * 9 + tap];
    wf[idx] = __float2bfloat16_rn(v);
    if (wt) wt[size_t(ci) * 9 * Cout_p + (8 - tap) * Cout_p + np] = __float2bfloat16_rn(v);
    if (bias_packed && tap == 0 && ci == 0) bias_packed[np] = (bias && n < Cout) ? bias[n] : 0.f;
  }
}

// partials [splits][9][co_pad][Cin_p] -> dW [Cout][Cin][3][3]
static __global__ void conv_unpack_wgrad_kernel(const float* __restrict__ part, int splits, int co_pad, int Cin_p,
                                         float* __restrict__ dw, int Cout, int Cin, int Cout_p, int ps) {
  const int total = Cout * Cin * 9;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i / (Cin * 9), ci = (i / 9) % Cin, tap = i % 9;
  const int np = ps ? ps_pack(n, Cout_p / 4) : n;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += part[((size_t(s) * 9 + tap) * co_pad + np) * Cin_p + ci];
  dw[i] = acc;
}

// db[n] = sum over pixels of dY; dY is NHWC [B,H,W,C] or, when ps, the shuffled tensor [B,2H,2W,C/4]
// (n = c*4 + i*2 + j).  One block per channel group; two-stage via atomics on a zeroed output is avoided:
// grid = (nblocks), partial[blockIdx][C] then reduced by colsum_finish_kernel.
static __global__ void colsum_nhwc_kernel(const __nv_bfloat16* __restrict__ dy, long long npix, int C, int ld,
                                   float* __restrict__ partial) {
  // thread t handles channel pair (t % (C/2)); rows strided by blockDim/(C/2) * gridDim
  const int pairs = C / 2;
  const int lanes_per_row = pairs;
  const int rows_per_block = blockDim.x / lanes_per_row;
  const int cp = threadIdx.x % lanes_per_row, rl = threadIdx.x / lanes_per_row;
  float a0 = 0.f, a1 = 0.f;
  if (rl < rows_per_block) {
    for (long long r = (long long)blockIdx.x * rows_per_block + rl; r < npix; r += (long long)gridDim.x * rows_per_block) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(dy + r * ld + cp * 2);
      a0 += bf16_lo(v);
      a1 += bf16_hi(v);
    }
  }
  extern __shared__ float s_cs[];  // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_cs[i] = 0.f;
  __syncthreads();
  if (rl < rows_per_block) {
    atomicAdd(&s_cs[cp * 2], a0);
    atomicAdd(&s_cs[cp * 2 + 1], a1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) partial[size_t(blockIdx.x) * C + i] = s_cs[i];
}
// ps layout: pixel index p of the shuffled image -> sub-pixel (i,j) = (Y&1, X&1); handled by running
// colsum_nhwc_kernel per sub-pixel view is wasteful, so the finish kernel receives 4 partial sets instead.
static __global__ void colsum_finish_kernel(const float* __restrict__ partial, int nparts, int C, float* __restrict__ out,
                                     int n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  float acc = 0.f;
  for (int k = 0; k < nparts; ++k) acc += partial[size_t(k) * C + i];
  out[i] = acc;
}
// shuffled-gradient bias sum: dYs [B,2H,2W,Cg]; out[c*4 + i*2 + j] = sum over pixels with (Y&1,X&1) == (i,j).
// 8 lanes x 16 B cover the Cg = 64 channels of a pixel; a thread always visits pixels of ONE sub-pixel lattice
// (sub = pixel lane & 3), so it needs 8 accumulators only, and 4 independent 16-byte loads are kept in flight.
static __global__ void colsum_ps_kernel(const __nv_bfloat16* __restrict__ dys, int B, int H2, int W2, int Cg,
                                 float* __restrict__ partial /*[grid][4*Cg]*/) {
  extern __shared__ float s_cs[];  // [4*Cg]
  for (int i = threadIdx.x; i < 4 * Cg; i += blockDim.x) s_cs[i] = 0.f;
  __syncthreads();
  const int groups = Cg / 8;                       // 16-byte channel groups per pixel
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int lanes = blockDim.x / groups;           // pixel lanes per block (multiple of 4)
  const int sub = pl & 3, si = sub >> 1, sj = sub & 1;
  const int Hq = H2 >> 1, Wq = W2 >> 1;
  const long long nq = (long long)B * Hq * Wq;     // quarter-resolution positions
  const long long qstride = (long long)gridDim.x * (lanes >> 2);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto load = [&](long long q) -> uint4 {
    const int qx = int(q % Wq), qy = int((q / Wq) % Hq);
    const long long b = q / ((long long)Wq * Hq);
    const long long pix = (b * H2 + (2 * qy + si)) * W2 + (2 * qx + sj);
    return *reinterpret_cast<const uint4*>(dys + pix * Cg + cg * 8);
  };
  auto add = [&](const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[2 * e] += bf16_lo(w[e]); acc[2 * e + 1] += bf16_hi(w[e]); }
  };
  long long q = (long long)blockIdx.x * (lanes >> 2) + (pl >> 2);
  for (; q + 3 * qstride < nq; q += 4 * qstride) {
    const uint4 v0 = load(q), v1 = load(q + qstride), v2 = load(q + 2 * qstride), v3 = load(q + 3 * qstride);
    add(v0); add(v1); add(v2); add(v3);
  }
  for (; q < nq; q += qstride) add(load(q));
#pragma unroll
  for (int e = 0; e < 8; ++e) atomicAdd(&s_cs[(cg * 8 + e) * 4 + sub], acc[e]);
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * Cg; i += blockDim.x) partial[size_t(blockIdx.x) * 4 * Cg + i] = s_cs[i];
}

// ------------------------------------------------------------------ conv_first: Cin = 1 -> C, output token-major
// y[p][co] = b[co] + sum_tap x[p + tap] * w[co][tap];  y: [B*H*W, Cp] bf16 (pads zero), x: [B,H,W] fp32
// Idx: unsigned when B*H*W*Cp/8 fits 32 bits (the 64-bit divisions of the index decomposition otherwise cost more than
// the arithmetic: 87 -> us at 512^2 x 2 x 64 channels, the discriminator's first layer)
template <typename Idx>
static __global__ void conv_in1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                                    int C, int Cp) {
  extern __shared__ __align__(16) float s_w[];  // [10][Cp]: 9 taps + bias, tap-major (a [Cp][10] layout puts the 24 channel
  for (int i = threadIdx.x; i < Cp * 10; i += blockDim.x) {   // groups of a warp on two bank sets: 12-way conflicts)
    const int t = i / Cp, co = i % Cp;
    s_w[i] = (co < C) ? (t < 9 ? w[co * 9 + t] : bias[co]) : 0.f;
  }
  __syncthreads();
  const Idx groups = Idx(Cp / 8);
  const Idx total = Idx(B) * H * W * groups;
  for (Idx idx = Idx(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += Idx(gridDim.x) * blockDim.x) {
    const int g = int(idx % groups);
    const Idx pi = idx / groups;
    const Idx ri = pi / Idx(W);
    const int xx = int(pi - ri * Idx(W)), yy = int(ri % Idx(H));
    const long long p = (long long)pi;
    const long long base = p - (long long)yy * W - xx;  // start of image b
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int sy = yy + t / 3 - 1, sx = xx + t % 3 - 1;
      v[t] = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? x[base + (long long)sy * W + sx] : 0.f;
    }
    float o[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(s_w + 9 * Cp + g * 8), b1 = *reinterpret_cast<const float4*>(s_w + 9 * Cp + g * 8 + 4);
      o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y; o[6] = b1.z; o[7] = b1.w;
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + t * Cp + g * 8), w1 = *reinterpret_cast<const float4*>(s_w + t * Cp + g * 8 + 4);
      o[0] = fmaf(v[t], w0.x, o[0]); o[1] = fmaf(v[t], w0.y, o[1]); o[2] = fmaf(v[t], w0.z, o[2]); o[3] = fmaf(v[t], w0.w, o[3]);
      o[4] = fmaf(v[t], w1.x, o[4]); o[5] = fmaf(v[t], w1.y, o[5]); o[6] = fmaf(v[t], w1.z, o[6]); o[7] = fmaf(v[t], w1.w, o[7]);
    }
    *reinterpret_cast<uint4*>(y + p * Cp + g * 8) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

// Same arithmetic (bias first, taps in the same order: bit-identical results), four horizontally adjacent pixels per thread:
// the 18 LDS.128 of the filter taps are paid once per four pixels instead of once per pixel, and the 3 x 6 input window is
// three float4 + six scalar loads.  At 512^2 (the discriminators' first layer) the one-pixel kernel is bound by exactly
// those shared-memory reads (87 us for a 67 MB output).  W % 4 == 0.
template <typename Idx>
static __global__ void __launch_bounds__(256) conv_in1_fwd_x4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                                     int B, int H, int W, int C, int Cp) {
  extern __shared__ __align__(16) float s_w[];  // [10][Cp]: 9 taps + bias, tap-major
  for (int i = threadIdx.x; i < Cp * 10; i += blockDim.x) {
    const int t = i / Cp, co = i % Cp;
    s_w[i] = (co < C) ? (t < 9 ? w[co * 9 + t] : bias[co]) : 0.f;
  }
  __syncthreads();
  const Idx groups = Idx(Cp / 8);
  const Idx Wq = Idx(W / 4);
  const Idx total = Idx(B) * H * Wq * groups;
  for (Idx idx = Idx(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += Idx(gridDim.x) * blockDim.x) {
    const int g = int(idx % groups);
    const Idx q = idx / groups;
    const Idx r = q / Wq;
    const int xx0 = int(q - r * Wq) * 4, yy = int(r % Idx(H));
    const long long img = (long long)(r / Idx(H)) * H * W;
    float v[3][6];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int sy = yy + dy - 1;
      if (sy >= 0 && sy < H) {
        const float* row = x + img + (long long)sy * W + xx0;
        const float4 m = *reinterpret_cast<const float4*>(row);
        v[dy][0] = (xx0 > 0) ? row[-1] : 0.f;
        v[dy][1] = m.x; v[dy][2] = m.y; v[dy][3] = m.z; v[dy][4] = m.w;
        v[dy][5] = (xx0 + 4 < W) ? row[4] : 0.f;
      } else {
#pragma unroll
        for (int dx = 0; dx < 6; ++dx) v[dy][dx] = 0.f;
      }
    }
    float o[4][8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(s_w + 9 * Cp + g * 8), b1 = *reinterpret_cast<const float4*>(s_w + 9 * Cp + g * 8 + 4);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        o[p][0] = b0.x; o[p][1] = b0.y; o[p][2] = b0.z; o[p][3] = b0.w; o[p][4] = b1.x; o[p][5] = b1.y; o[p][6] = b1.z; o[p][7] = b1.w;
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + t * Cp + g * 8), w1 = *reinterpret_cast<const float4*>(s_w + t * Cp + g * 8 + 4);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float xv = v[t / 3][t % 3 + p];
        o[p][0] = fmaf(xv, w0.x, o[p][0]); o[p][1] = fmaf(xv, w0.y, o[p][1]); o[p][2] = fmaf(xv, w0.z, o[p][2]); o[p][3] = fmaf(xv, w0.w, o[p][3]);
        o[p][4] = fmaf(xv, w1.x, o[p][4]); o[p][5] = fmaf(xv, w1.y, o[p][5]); o[p][6] = fmaf(xv, w1.z, o[p][6]); o[p][7] = fmaf(xv, w1.w, o[p][7]);
      }
    }
    __nv_bfloat16* dst = y + (img + (long long)yy * W + xx0) * Cp + g * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p)
      *reinterpret_cast<uint4*>(dst + (long long)p * Cp) =
          make_uint4(pack_bf16(o[p][0], o[p][1]), pack_bf16(o[p][2], o[p][3]), pack_bf16(o[p][4], o[p][5]), pack_bf16(o[p][6], o[p][7]));
  }
}

// dW[co][tap] = sum_p dY[p][co] * x[p+tap], db[co] = sum_p dY[p][co]; partial[blockIdx][Cp][10]
template <typename Idx>
static __global__ void conv_in1_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                      float* __restrict__ partial, int B, int H, int W, int Cp) {
  extern __shared__ float s_acc[];  // [Cp][10]
  for (int i = threadIdx.x; i < Cp * 10; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int groups = Cp / 8;
  const int pix_per_block = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  float acc[8][10];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[e][t] = 0.f;
  const Idx npix = Idx(B) * H * W;
  if (pl < pix_per_block) {
    // two pixels per iteration: both pixels' ten loads are issued before the 160 FMAs (the one-pixel loop was bound by the
    // latency of its dependent load -> FMA chain: 128 us for 67 MB at 512^2)
    const Idx stride = Idx(gridDim.x) * pix_per_block;
    auto load = [&](Idx pi, float (&v)[9], uint4& d) {
      const Idx ri = pi / Idx(W);
      const int xx = int(pi - ri * Idx(W)), yy = int(ri % Idx(H));
      const long long p = (long long)pi;
      const long long base = p - (long long)yy * W - xx;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int sy = yy + t / 3 - 1, sx = xx + t % 3 - 1;
        v[t] = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? x[base + (long long)sy * W + sx] : 0.f;
      }
      d = *reinterpret_cast<const uint4*>(dy + p * Cp + g * 8);
    };
    auto fma_all = [&](const float (&v)[9], const uint4& d) {
      const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dv = (e & 1) ? bf16_hi(dw[e >> 1]) : bf16_lo(dw[e >> 1]);
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[e][t] = fmaf(dv, v[t], acc[e][t]);
        acc[e][9] += dv;
      }
    };
    Idx pi = Idx(blockIdx.x) * pix_per_block + pl;
    for (; pi + stride < npix; pi += 2 * stride) {
      float va[9], vb[9];
      uint4 da, db_;
      load(pi, va, da);
      load(pi + stride, vb, db_);
      fma_all(va, da);
      fma_all(vb, db_);
    }
    if (pi < npix) {
      float va[9];
      uint4 da;
      load(pi, va, da);
      fma_all(va, da);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int t = 0; t < 10; ++t) atomicAdd(&s_acc[(g * 8 + e) * 10 + t], acc[e][t]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cp * 10; i += blockDim.x) partial[size_t(blockIdx.x) * Cp * 10 + i] = s_acc[i];
}
// one warp per output element: lanes stride over the per-block partials (592 of them for the discriminator's first layer: the
// one-thread-per-output loop took 36 us of dependent loads)
static __global__ void conv_in1_wgrad_finish_kernel(const float* __restrict__ partial, int nparts, int C, int Cp,
                                             float* __restrict__ dw, float* __restrict__ db) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= C * 10) return;
  const int co = i / 10, t = i % 10;
  float acc = 0.f;
  for (int k = lane; k < nparts; k += 32) acc += partial[size_t(k) * Cp * 10 + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (t < 9) dw[co * 9 + t] = acc; else db[co] = acc;
  }
}

// ------------------------------------------------------------------ conv_last: C(=64) -> 1, NHWC bf16 in, fp32 out
// 8 lanes per output pixel, each lane owns 8 input channels; 4 pixels per warp.
template <int C>
static __global__ void conv_out1_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w /*[1][C][3][3]*/,
                                     const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W) {
  static_assert(C == 64, "conv_out1 is specialised for 64 input channels");
  __shared__ float s_w[9][C];
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i / C][i % C] = w[(i % C) * 9 + i / C];
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane >> 3, cl = lane & 7;
  const long long npix = (long long)B * H * W;
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float b0 = bias[0];
  for (long long p0 = warp_id * 4; p0 < npix; p0 += nwarps * 4) {
    const long long p = p0 + sub;
    float acc = 0.f;
    if (p < npix) {
      const int xx = int(p % W), yy = int((p / W) % H);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int sy = yy + t / 3 - 1, sx = xx + t % 3 - 1;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
          const long long q = p + (long long)(t / 3 - 1) * W + (t % 3 - 1);
          const uint4 v = *reinterpret_cast<const uint4*>(x + q * C + cl * 8);
          const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc = fmaf(bf16_lo(vw[e]), s_w[t][cl * 8 + 2 * e], acc);
            acc = fmaf(bf16_hi(vw[e]), s_w[t][cl * 8 + 2 * e + 1], acc);
          }
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (cl == 0 && p < npix) y[p] = acc + b0;
  }
}

// dX[p][c] = sum_tap dY[p - off(tap)] * w[c][tap]   (dY fp32 [B,H,W], dX NHWC bf16); HBM-write bound.
template <int C, typename Idx>   // Idx = int whenever B*H*W*C/8 < 2^31: 64-bit div/mod costs ~10x the useful work here
static __global__ void conv_out1_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                       __nv_bfloat16* __restrict__ dx, int B, int H, int W) {
  __shared__ __align__(16) float s_w[9][C];
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i / C][i % C] = w[(i % C) * 9 + i / C];
  __syncthreads();
  constexpr int G = C / 8;
  const Idx total = (Idx)B * H * W * G;
  for (Idx idx = (Idx)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (Idx)gridDim.x * blockDim.x) {
    const int g = int(idx % G);
    const Idx p = idx / G;
    const int xx = int(p % W), yy = int((p / W) % H);
    float d[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      // output pixel q = p - off(tap) received x[p] through tap t
      const int sy = yy - (t / 3 - 1), sx = xx - (t % 3 - 1);
      d[t] = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? __ldg(dy + p - (Idx)(t / 3 - 1) * W - (t % 3 - 1)) : 0.f;
    }
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&s_w[t][g * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&s_w[t][g * 8 + 4]);
      o[0] = fmaf(d[t], w0.x, o[0]); o[1] = fmaf(d[t], w0.y, o[1]); o[2] = fmaf(d[t], w0.z, o[2]); o[3] = fmaf(d[t], w0.w, o[3]);
      o[4] = fmaf(d[t], w1.x, o[4]); o[5] = fmaf(d[t], w1.y, o[5]); o[6] = fmaf(d[t], w1.z, o[6]); o[7] = fmaf(d[t], w1.w, o[7]);
    }
    *reinterpret_cast<uint4*>(dx + (size_t)p * C + g * 8) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

// dW[c][tap] = sum_p dY[p] * x[p + off(tap)][c], db = sum_p dY[p]; partial[blockIdx][C*9 + 1]
template <int C, typename Idx>
static __global__ void conv_out1_wgrad_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                       float* __restrict__ partial, int B, int H, int W) {
  __shared__ float s_acc[C * 9 + 1];
  for (int i = threadIdx.x; i < C * 9 + 1; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  constexpr int G = C / 8;
  const int g = threadIdx.x % G, pl = threadIdx.x / G, ppb = blockDim.x / G;
  float acc[8][9];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[e][t] = 0.f;
  float accb = 0.f;
  const Idx npix = (Idx)B * H * W;
  const Idx pstride = (Idx)gridDim.x * ppb;
  Idx p = (Idx)blockIdx.x * ppb + pl;
  uint4 v_next = make_uint4(0u, 0u, 0u, 0u);
  if (p < npix) v_next = *reinterpret_cast<const uint4*>(x + (size_t)p * C + g * 8);
  for (; p < npix; p += pstride) {
    // gather form: input pixel p contributes to output pixel q = p - off(tap) through tap t
    const uint4 v = v_next;
    if (p + pstride < npix) v_next = *reinterpret_cast<const uint4*>(x + (size_t)(p + pstride) * C + g * 8);  // prefetch
    const int xx = int(p % W), yy = int((p / W) % H);
    const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
    float xv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) xv[e] = (e & 1) ? bf16_hi(vw[e >> 1]) : bf16_lo(vw[e >> 1]);
    if (g == 0) accb += __ldg(dy + p);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int sy = yy - (t / 3 - 1), sx = xx - (t % 3 - 1);
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
        const float d = __ldg(dy + p - (Idx)(t / 3 - 1) * W - (t % 3 - 1));
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e][t] = fmaf(d, xv[e], acc[e][t]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(&s_acc[(g * 8 + e) * 9 + t], acc[e][t]);
  if (g == 0) atomicAdd(&s_acc[C * 9], accb);
  __syncthreads();
  for (int i = threadIdx.x; i < C * 9 + 1; i += blockDim.x) partial[size_t(blockIdx.x) * (C * 9 + 1) + i] = s_acc[i];
}


// ------------------------------------------------------------------ conv_last backward, row-walking versions (round 2)
// The two kernels above spend most of their instructions on index arithmetic (two integer divisions and nine four-way
// bounds tests per 8 outputs) and keep 16 bytes per thread in flight.  Here a warp owns (image, four rows, x segment):
// lane = (row r = lane / 8, channel group g = lane % 8); a thread walks x with a sliding 3x3 window of dY in registers
// (three new values per pixel, row validity decided once per task), its 8 x 9 weights (dgrad) or accumulators (wgrad) in
// registers.  Loads and stores of a warp cover four full 128-byte pixel rows.  H % 4 == 0.
__device__ __forceinline__ float ldrow(const float* row, int x) { return row ? __ldg(row + x) : 0.f; }

template <int C>
static __global__ void __launch_bounds__(256) conv_out1_dgrad_rows_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                       __nv_bfloat16* __restrict__ dx, int B, int H, int W, int seg) {
  static_assert(C == 64, "lane mapping: 8 channel groups of 8");
  const int lane = threadIdx.x & 31, g = lane & 7, r = lane >> 3;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[t][e] = w[(g * 8 + e) * 9 + t];
  const int nseg = (W + seg - 1) / seg, hq = H >> 2;
  const long long total = (long long)B * hq * nseg;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < total; task += nwarps) {
    const int sgm = int(task % nseg);
    const int yq = int((task / nseg) % hq), b = int(task / ((long long)nseg * hq));
    const int y = yq * 4 + r, x0 = sgm * seg, x1 = min(W, x0 + seg);
    const float* rows[3];
    rows[1] = dy + ((size_t)b * H + y) * W;
    rows[0] = y > 0 ? rows[1] - W : nullptr;
    rows[2] = y < H - 1 ? rows[1] + W : nullptr;
    float d[3][3];   // d[i][j] = dY[y - 1 + i][x - 1 + j]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      d[i][1] = (x0 > 0) ? ldrow(rows[i], x0 - 1) : 0.f;
      d[i][2] = ldrow(rows[i], x0);
    }
    __nv_bfloat16* out = dx + (((size_t)b * H + y) * W + x0) * C + g * 8;
    for (int x = x0; x < x1; ++x, out += C) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        d[i][0] = d[i][1];
        d[i][1] = d[i][2];
        d[i][2] = (x + 1 < W) ? ldrow(rows[i], x + 1) : 0.f;
      }
      float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 9; ++t) {   // output pixel q = p - off(t) received x[p] through tap t: dY at (y + 1 - t/3, x + 1 - t%3)
        const float dv = d[2 - t / 3][2 - t % 3];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(dv, wr[t][e], o[e]);
      }
      *reinterpret_cast<uint4*>(out) =
          make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
}

// y[p] = bias + sum_tap sum_c x[p + off(tap)][c] * w[c][tap]   (x NHWC bf16, y fp32 [B,H,W]; weights rounded to bf16 like
// the tensor-core path this replaces: a 64 -> 1 convolution is N = 16 for tcgen05.mma, i.e. 36 M = 128 instructions per 128
// pixels for one useful column, 613 us at 512^2 x 16).  Same task mapping as above; the 3x3 window holds packed bf16.
template <int C>
static __global__ void __launch_bounds__(256, 2) conv_out1_fwd_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, float* __restrict__ y,
                                                                     int B, int H, int W, int seg) {
  static_assert(C == 64, "lane mapping: 8 channel groups of 8");
  const int lane = threadIdx.x & 31, g = lane & 7, r = lane >> 3;
  __shared__ __align__(16) float s_w[9][C];   // weights stay in shared memory: 72 more registers would halve the occupancy
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i / C][i % C] = round_bf16(w[(i % C) * 9 + i / C]);
  __syncthreads();
  const float b0 = bias[0];
  const int nseg = (W + seg - 1) / seg, hq = H >> 2;
  const long long total = (long long)B * hq * nseg;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < total; task += nwarps) {
    const int sgm = int(task % nseg);
    const int yq = int((task / nseg) % hq), b = int(task / ((long long)nseg * hq));
    const int yy = yq * 4 + r, x0 = sgm * seg, x1 = min(W, x0 + seg);
    const __nv_bfloat16* rows[3];
    rows[1] = x + ((size_t)b * H + yy) * W * C + g * 8;
    rows[0] = yy > 0 ? rows[1] - (size_t)W * C : nullptr;
    rows[2] = yy < H - 1 ? rows[1] + (size_t)W * C : nullptr;
    auto ld = [&](int i, int xx) { return rows[i] ? *reinterpret_cast<const uint4*>(rows[i] + (size_t)xx * C) : zero4; };
    uint4 d[3][4];   // d[i][j] = x[yy - 1 + i][xx - 1 + j][8 channels]; column 3 is the load in flight for the next pixel
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      d[i][1] = (x0 > 0) ? ld(i, x0 - 1) : zero4;
      d[i][2] = ld(i, x0);
      d[i][3] = (x0 + 1 < W) ? ld(i, x0 + 1) : zero4;
    }
    float* out = y + ((size_t)b * H + yy) * W;
    for (int xx = x0; xx < x1; ++xx) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        d[i][0] = d[i][1];
        d[i][1] = d[i][2];
        d[i][2] = d[i][3];
        d[i][3] = (xx + 2 < W) ? ld(i, xx + 2) : zero4;   // consumed one iteration later
      }
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const uint4 v = d[t / 3][t % 3];
        const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[t][g * 8]), w1 = *reinterpret_cast<const float4*>(&s_w[t][g * 8 + 4]);
        a0 = fmaf(bf16_lo(vw[0]), w0.x, a0); a1 = fmaf(bf16_hi(vw[0]), w0.y, a1);
        a0 = fmaf(bf16_lo(vw[1]), w0.z, a0); a1 = fmaf(bf16_hi(vw[1]), w0.w, a1);
        a0 = fmaf(bf16_lo(vw[2]), w1.x, a0); a1 = fmaf(bf16_hi(vw[2]), w1.y, a1);
        a0 = fmaf(bf16_lo(vw[3]), w1.z, a0); a1 = fmaf(bf16_hi(vw[3]), w1.w, a1);
      }
      float acc = a0 + a1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (g == 0) out[xx] = acc + b0;
    }
  }
}

// dW[c][tap] = sum_p dY[p - off(tap)] * x[p][c], db = sum_p dY[p]; partial[blockIdx][C*9 + 1]
template <int C>
static __global__ void __launch_bounds__(256) conv_out1_wgrad_rows_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                       float* __restrict__ partial, int B, int H, int W, int seg) {
  static_assert(C == 64, "lane mapping: 8 channel groups of 8");
  __shared__ float s_acc[C * 9 + 1];
  for (int i = threadIdx.x; i < C * 9 + 1; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane & 7, r = lane >> 3;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
  float accb = 0.f;
  const int nseg = (W + seg - 1) / seg, hq = H >> 2;
  const long long total = (long long)B * hq * nseg;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < total; task += nwarps) {
    const int sgm = int(task % nseg);
    const int yq = int((task / nseg) % hq), b = int(task / ((long long)nseg * hq));
    const int y = yq * 4 + r, x0 = sgm * seg, x1 = min(W, x0 + seg);
    const float* rows[3];
    rows[1] = dy + ((size_t)b * H + y) * W;
    rows[0] = y > 0 ? rows[1] - W : nullptr;
    rows[2] = y < H - 1 ? rows[1] + W : nullptr;
    float d[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      d[i][1] = (x0 > 0) ? ldrow(rows[i], x0 - 1) : 0.f;
      d[i][2] = ldrow(rows[i], x0);
    }
    const __nv_bfloat16* in = x + (((size_t)b * H + y) * W + x0) * C + g * 8;
    for (int xb = x0; xb < x1; xb += 4, in += 4 * C) {
      uint4 xq[4];   // four pixels of this thread's 8 channels in flight
#pragma unroll
      for (int k = 0; k < 4; ++k)
        xq[k] = (xb + k < x1) ? *reinterpret_cast<const uint4*>(in + k * C) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xx = xb + k;
        if (xx < x1) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            d[i][0] = d[i][1];
            d[i][1] = d[i][2];
            d[i][2] = (xx + 1 < W) ? ldrow(rows[i], xx + 1) : 0.f;
          }
          const uint32_t vw[4] = {xq[k].x, xq[k].y, xq[k].z, xq[k].w};
          float xv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[e] = (e & 1) ? bf16_hi(vw[e >> 1]) : bf16_lo(vw[e >> 1]);
          if (g == 0) accb += d[1][1];
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float dv = d[2 - t / 3][2 - t % 3];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[t][e] = fmaf(dv, xv[e], acc[t][e]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_acc[(g * 8 + e) * 9 + t], acc[t][e]);
  if (g == 0) atomicAdd(&s_acc[C * 9], accb);
  __syncthreads();
  for (int i = threadIdx.x; i < C * 9 + 1; i += blockDim.x) partial[size_t(blockIdx.x) * (C * 9 + 1) + i] = s_acc[i];
}

}  // namespace srk

namespace srk {
// ------------------------------------------------------------------ channel-slice ("view") elementwise helpers
// A view is (pointer to the first channel, channel count C, pixel pitch in elements): a channel slice of an NHWC
// tensor.  They serve the dense blocks of the hybrid generator (ResidualDenseBlock, models/hybridmodels_hat.py:21-44),
// whose torch.cat inputs are slices of ONE [pixels, nf + 4*gc] buffer here (virtual concat).  C % 8 == 0, 16-byte
// aligned slices: every thread moves 16 bytes.

// g[:, :C] *= (f[:, :C] > 0 ? 1 : slope)     (backward through LeakyReLU; f = forward output)
// Optionally also the column sums of the masked gradient (= the bias gradient of the layer that produced f): the total
// thread count is a multiple of C/8, so a thread always meets the same 8 channels and keeps their sums in registers;
// per-block partials [gridDim.x][C] are reduced by colsum_finish_kernel.
static __global__ void view_lrelu_mask_kernel(__nv_bfloat16* __restrict__ g, int ldg, const __nv_bfloat16* __restrict__ f,
                                              int ldf, int C, long long npix, float slope, float* __restrict__ partial) {
  extern __shared__ float s_cs[];  // [blockDim.x][8]
  const int groups = C >> 3;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = int(tid % groups) * 8;
  const long long pstride = nthreads / groups;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long p = tid / groups; p < npix; p += pstride) {
    uint4* gp = reinterpret_cast<uint4*>(g + p * ldg + c);
    const uint4 fv = *reinterpret_cast<const uint4*>(f + p * ldf + c);
    uint4 gv = *gp;
    uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
    const uint32_t fw[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = bf16_lo(fw[e]) > 0.f ? bf16_lo(gw[e]) : bf16_lo(gw[e]) * slope;
      const float hi = bf16_hi(fw[e]) > 0.f ? bf16_hi(gw[e]) : bf16_hi(gw[e]) * slope;
      gw[e] = pack_bf16(lo, hi);
      acc[2 * e] += bf16_lo(gw[e]);      // sums of the values as stored (what the weight-gradient kernel will read)
      acc[2 * e + 1] += bf16_hi(gw[e]);
    }
    *gp = make_uint4(gw[0], gw[1], gw[2], gw[3]);
  }
  if (partial != nullptr) {   // block-level sums without atomics: every thread parks its 8 sums, C threads fold them
#pragma unroll
    for (int e = 0; e < 8; ++e) s_cs[threadIdx.x * 8 + e] = acc[e];
    __syncthreads();
    if (threadIdx.x < C) {
      const int grp = threadIdx.x >> 3, e = threadIdx.x & 7;
      const int base = int(((long long)blockIdx.x * blockDim.x) % groups);
      int t0 = grp - base;             // first thread of this block whose channel group is grp
      if (t0 < 0) t0 += groups;
      float sum = 0.f;
      for (int t = t0; t < int(blockDim.x); t += groups) sum += s_cs[t * 8 + e];
      partial[size_t(blockIdx.x) * C + threadIdx.x] = sum;
    }
  }
}
// y[:, :C] = alpha * a[:, :C] + (x ? x[:, :C] : 0)      (y may alias a or x)
static __global__ void view_axpy_kernel(__nv_bfloat16* y, int ldy, const __nv_bfloat16* a, int lda, const __nv_bfloat16* x,
                                        int ldx, int C, long long npix, float alpha) {
  const int groups = C >> 3;
  const long long total = npix * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / groups;
    const int c = int(i - p * groups) * 8;
    const uint4 av = *reinterpret_cast<const uint4*>(a + p * lda + c);
    uint4 xv = make_uint4(0u, 0u, 0u, 0u);
    if (x != nullptr) xv = *reinterpret_cast<const uint4*>(x + p * ldx + c);
    const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, xw[4] = {xv.x, xv.y, xv.z, xv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = pack_bf16(round_bf16(bf16_lo(aw[e]) * alpha) + bf16_lo(xw[e]), round_bf16(bf16_hi(aw[e]) * alpha) + bf16_hi(xw[e]));
    *reinterpret_cast<uint4*>(y + p * ldy + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// nearest-neighbour x2 (F.interpolate(scale_factor=2, mode='nearest'), hybridmodels_hat.py:127):
// y[b, 2h+i, 2w+j, :] = x[b, h, w, :]
static __global__ void nearest2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y,
                                           int ldy, int C, int B, int H, int W) {
  const int groups = C >> 3;
  const long long total = (long long)B * H * W * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / groups;
    const int c = int(i - p * groups) * 8;
    const int w = int(p % W), h = int((p / W) % H);
    const long long b = p / ((long long)W * H);
    const uint4 v = *reinterpret_cast<const uint4*>(x + p * ldx + c);
    const long long q = (b * 2 * H + 2 * h) * (2LL * W) + 2 * w;
    *reinterpret_cast<uint4*>(y + q * ldy + c) = v;
    *reinterpret_cast<uint4*>(y + (q + 1) * ldy + c) = v;
    *reinterpret_cast<uint4*>(y + (q + 2 * W) * ldy + c) = v;
    *reinterpret_cast<uint4*>(y + (q + 2 * W + 1) * ldy + c) = v;
  }
}
// dx[b, h, w, :] = sum of the four dy pixels it was copied to
static __global__ void nearest2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, __nv_bfloat16* __restrict__ dx,
                                           int lddx, int C, int B, int H, int W) {
  const int groups = C >> 3;
  const long long total = (long long)B * H * W * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / groups;
    const int c = int(i - p * groups) * 8;
    const int w = int(p % W), h = int((p / W) % H);
    const long long b = p / ((long long)W * H);
    const long long q = (b * 2 * H + 2 * h) * (2LL * W) + 2 * w;
    const long long qs[4] = {q, q + 1, q + 2 * W, q + 2 * W + 1};
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 v = *reinterpret_cast<const uint4*>(dy + qs[k] * lddy + c);
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc[2 * e] += bf16_lo(wv[e]); acc[2 * e + 1] += bf16_hi(wv[e]); }
    }
    *reinterpret_cast<uint4*>(dx + p * lddx + c) =
        make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}
// single-channel image <-> 8-channel bf16 NHWC rows (channel 0 = the image, channels 1..7 zero): lets the 1 -> nf and
// nf -> 1 convolutions of the hybrid generator (conv_adapt / conv_last, hybridmodels_hat.py:94,105) and their
// gradients run through the tcgen05 implicit-GEMM kernel (TMA zero-fills the other 56 K columns).
static __global__ void img1_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long npix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x)
    *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(pack_bf16(x[i], 0.f), 0u, 0u, 0u);
}
static __global__ void img1_unpack_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long npix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i * 8]);
}
}  // namespace srk
