// cab_aux.cuh — the memory-bound pieces of HAT's channel-attention block (CAB) around its two implicit-GEMM convs:
//   global average pool per image, the squeeze/excite MLP (C -> C/squeeze -> C, ReLU, sigmoid), the channel scaling
//   fused with the conv_scale residual add, and their backward passes.
// Reference: ChannelAttention hat_arch.py:40-58, CAB :61-74, `shortcut + attn + conv_x * conv_scale` in HAB.forward :306.
// Layout: activations token-major bf16 [B*HW, Cp] (== NHWC), per-image vectors fp32 [B, C].
#pragma once
#include "srk_ptx.cuh"

namespace srk {

// partial[(b*nchunk + chunk)*Cp + c] = sum over the chunk's rows of image b of x[r][c] (* y[r][c] when y != nullptr)
// grid (nchunk, B); block = (Cp/8) * rows_per_block threads (a thread owns 8 channels: 16-byte loads, four rows in
// flight); dynamic smem Cp floats.
static __global__ void cab_colsum_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y, int HW,
                                         int Cp, float* __restrict__ partial) {
  extern __shared__ float s_cs[];
  const int groups = Cp / 8;
  const int rows_per_block = blockDim.x / groups;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  const int nchunk = gridDim.x, chunk = blockIdx.x, b = blockIdx.y;
  const int r_begin = int((long long)chunk * HW / nchunk), r_end = int((long long)(chunk + 1) * HW / nchunk);
  for (int i = threadIdx.x; i < Cp; i += blockDim.x) s_cs[i] = 0.f;
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto add = [&](const uint4& v, const uint4& u) {
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w}, wu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float f0 = bf16_lo(wv[e]), f1 = bf16_hi(wv[e]);
      if (y != nullptr) { f0 *= bf16_lo(wu[e]); f1 *= bf16_hi(wu[e]); }
      acc[2 * e] += f0; acc[2 * e + 1] += f1;
    }
  };
  if (rl < rows_per_block) {
    const size_t base = (size_t)b * HW * Cp + cg * 8;
    int r = r_begin + rl;
    for (; r + 3 * rows_per_block < r_end; r += 4 * rows_per_block) {
      uint4 v[4], u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t off = base + (size_t)(r + k * rows_per_block) * Cp;
        v[k] = *reinterpret_cast<const uint4*>(x + off);
        u[k] = (y != nullptr) ? *reinterpret_cast<const uint4*>(y + off) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) add(v[k], u[k]);
    }
    for (; r < r_end; r += rows_per_block) {
      const size_t off = base + (size_t)r * Cp;
      const uint4 v = *reinterpret_cast<const uint4*>(x + off);
      const uint4 u = (y != nullptr) ? *reinterpret_cast<const uint4*>(y + off) : make_uint4(0u, 0u, 0u, 0u);
      add(v, u);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_cs[cg * 8 + e], acc[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cp; i += blockDim.x) partial[((size_t)b * nchunk + chunk) * Cp + i] = s_cs[i];
}

// One block per image: pool = mean, hidden = relu(W1 pool + b1), scale = sigmoid(W2 hidden + b2).
// w1 [S][C], w2 [C][S] (the 1x1 conv weights of ChannelAttention.attention.1 / .3 viewed as matrices).
static __global__ void cab_se_fwd_kernel(const float* __restrict__ partial, int nchunk, int Cp, int HW, int C, int S,
                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                         float* __restrict__ pool, float* __restrict__ hidden, float* __restrict__ scale) {
  extern __shared__ float s_se[];  // [C] pool, [S] hidden
  float* s_pool = s_se;
  float* s_hid = s_se + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < nchunk; ++k) acc += partial[((size_t)b * nchunk + k) * Cp + c];
    acc /= float(HW);
    s_pool[c] = acc;
    pool[(size_t)b * C + c] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int s = warp; s < S; s += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += w1[s * C + c] * s_pool[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      acc = fmaxf(acc + b1[s], 0.f);
      s_hid[s] = acc;
      hidden[(size_t)b * S + s] = acc;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float z = b2[c];
    for (int s = 0; s < S; ++s) z += w2[c * S + s] * s_hid[s];
    scale[(size_t)b * C + c] = 1.0f / (1.0f + __expf(-z));
  }
}

// out[p][c] = x[p][c] + alpha * y[p][c] * scale[b][c]   (c < C; pad columns copy x)
static __global__ void cab_combine_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                              const float* __restrict__ scale, float alpha, __nv_bfloat16* __restrict__ out,
                                              long long T, int HW, int C, int Cp) {
  const int groups = Cp / 8;
  const long long total = T * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gidx = int(i % groups);
    const long long p = i / groups;
    const int b = int(p / HW);
    const uint4 xv = *reinterpret_cast<const uint4*>(x + p * Cp + gidx * 8);
    const uint4 yv = *reinterpret_cast<const uint4*>(y + p * Cp + gidx * 8);
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = gidx * 8 + e;
      const float xf = (e & 1) ? bf16_hi(xw[e >> 1]) : bf16_lo(xw[e >> 1]);
      const float yf = (e & 1) ? bf16_hi(yw[e >> 1]) : bf16_lo(yw[e >> 1]);
      o[e] = (c < C) ? fmaf(alpha * yf, scale[(size_t)b * C + c], xf) : xf;
    }
    *reinterpret_cast<uint4*>(out + p * Cp + gidx * 8) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

// Backward of the squeeze/excite MLP for all images (single block; parameter gradients summed over images).
// dsc_partial: per-image column sums of g * y (cab_colsum_kernel with y), alpha applied here.
// Outputs: dpool_hw[b][c] = dL/dpool / HW (added to every pixel of d_y), dw1 [S][C], db1 [S], dw2 [C][S], db2 [C].
static __global__ void cab_se_bwd_kernel(const float* __restrict__ dsc_partial, int nchunk, int Cp, int HW, int B, int C,
                                         int S, float alpha, const float* __restrict__ pool, const float* __restrict__ hidden,
                                         const float* __restrict__ scale, const float* __restrict__ w1,
                                         const float* __restrict__ w2, float* __restrict__ dpool_hw,
                                         float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                         float* __restrict__ db2) {
  extern __shared__ float s_se[];  // [C] dz2, [S] dz1, [S] dhid scratch
  float* s_dz2 = s_se;
  float* s_dz1 = s_se + C;
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) { dw1[i] = 0.f; dw2[i] = 0.f; }
  for (int i = threadIdx.x; i < C; i += blockDim.x) db2[i] = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) db1[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int b = 0; b < B; ++b) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float ds = 0.f;
      for (int k = 0; k < nchunk; ++k) ds += dsc_partial[((size_t)b * nchunk + k) * Cp + c];
      const float sc = scale[(size_t)b * C + c];
      const float dz = alpha * ds * sc * (1.f - sc);
      s_dz2[c] = dz;
      db2[c] += dz;                      // thread-private column: no race
      for (int s = 0; s < S; ++s) dw2[c * S + s] += dz * hidden[(size_t)b * S + s];
    }
    __syncthreads();
    for (int s = warp; s < S; s += nwarps) {
      float acc = 0.f;
      for (int c = lane; c < C; c += 32) acc += s_dz2[c] * w2[c * S + s];
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        const float dz1 = hidden[(size_t)b * S + s] > 0.f ? acc : 0.f;
        s_dz1[s] = dz1;
        db1[s] += dz1;
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float dp = 0.f;
      const float pl = pool[(size_t)b * C + c];
      for (int s = 0; s < S; ++s) {
        dp += s_dz1[s] * w1[s * C + c];
        dw1[s * C + c] += s_dz1[s] * pl;  // thread-private column
      }
      dpool_hw[(size_t)b * C + c] = dp / float(HW);
    }
    __syncthreads();
  }
}

// dy[p][c] = alpha * g[p][c] * scale[b][c] + dpool_hw[b][c]   (c < C; pad columns 0)
static __global__ void cab_combine_bwd_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ scale,
                                              const float* __restrict__ dpool_hw, float alpha,
                                              __nv_bfloat16* __restrict__ dy, long long T, int HW, int C, int Cp) {
  const int groups = Cp / 8;
  const long long total = T * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gidx = int(i % groups);
    const long long p = i / groups;
    const int b = int(p / HW);
    const uint4 gv = *reinterpret_cast<const uint4*>(g + p * Cp + gidx * 8);
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = gidx * 8 + e;
      const float gf = (e & 1) ? bf16_hi(gw[e >> 1]) : bf16_lo(gw[e >> 1]);
      o[e] = (c < C) ? fmaf(alpha * gf, scale[(size_t)b * C + c], dpool_hw[(size_t)b * C + c]) : 0.f;
    }
    *reinterpret_cast<uint4*>(dy + p * Cp + gidx * 8) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

}  // namespace srk
