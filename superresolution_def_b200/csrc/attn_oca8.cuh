// attn_oca8.cuh — overlapping cross-attention core of OCAB for 8x8 query windows with a 12x12 key/value halo window
// (window_size 8, overlap_ratio 0.5: the configuration train_hat.py / infer_hat.py build, hybridmodels_hat.py:80-91),
// forward + backward.
//
// Replaces (reference, models/hat_arch/hat_arch.py): window_partition :97-109 of q, nn.Unfold(k=12, stride=8, pad=2)
// + einops rearrange of k/v :377,408-409 (ZERO outside the image, applied after the qkv projection), the q@k^T bmm,
// the (8+12-1)^2-entry bias gather + add with the reference's wrap-around negative indices :896-919, softmax, attn@v
// and window_reverse :419-428.  No mask.  S and P never touch HBM.
//
// Layouts as attn_win16.cuh: qkv [T, 3*heads*32] bf16 token-major (head slots padded to 32, q pre-scaled), out
// [T, heads*32].  One CTA keeps one head and walks windows; 4 warps x 16 query rows; the whole 64 x 144 logit tile is
// register-resident (18 n-tiles), so the softmax is exact (no online rescaling) and the backward recomputes it.
// The bias is expanded once per CTA into fragment order (one LDS.128 per n-tile, no index arithmetic per window).
#pragma once
#include "attn_win16.cuh"

namespace srk {

constexpr int O8_NK = 144;        // 12 x 12 key slots, slot = ky * 12 + kx
constexpr int O8_NT = 18;         // key n-tiles of 8
constexpr int O8_KS = 9;          // key k-steps of 16
constexpr int O8_TBL = 361;       // (8 + 12 - 1)^2
constexpr int O8_PROW = 304;      // bytes per row of the [64 q][144 keys] bf16 P / dS tiles (+16 B pad: conflict-free ldmatrix)
constexpr int O8_THREADS = 128;
constexpr int O8_BIAS_BYTES = 4 * O8_NT * 32 * 16;   // [warp][nt][lane] float4
constexpr int O8_FWD_SMEM = O8_BIAS_BYTES + 4096 + 2 * O8_NK * 64;
constexpr int O8_BWD_SMEM = O8_BIAS_BYTES + 2 * 4096 + 2 * O8_NK * 64 + 2 * 64 * O8_PROW;

struct O8Win { int b, wy, wx; };
__device__ __forceinline__ O8Win o8_win(const Attn16Args& a, int w) {
  const int nwx = a.W >> 3, nwy = a.H >> 3;
  O8Win p;
  p.b = w / (nwx * nwy);
  const int r = w - p.b * nwx * nwy;
  p.wy = r / nwx;
  p.wx = r - p.wy * nwx;
  return p;
}
__device__ __forceinline__ long long o8_qtok(const Attn16Args& a, const O8Win& p, int i) {
  return ((long long)p.b * a.H + p.wy * 8 + (i >> 3)) * a.W + p.wx * 8 + (i & 7);
}
// reference index (ky - qy + ws - wse + 1) * 19 + (kx - qx + ws - wse + 1); negative values wrap around the table end
__device__ __forceinline__ int o8_bias_index(int q, int k) {
  const int qy = q >> 3, qx = q & 7, ky = k / 12, kx = k - ky * 12;
  int idx = (ky - qy - 3) * 19 + (kx - qx - 3);
  return idx < 0 ? idx + O8_TBL : idx;
}
// bias of this head in fragment order: entry (warp, nt, lane) = logits (rows g / g+8, cols 2t, 2t+1) of n-tile nt
__device__ __forceinline__ void o8_fill_bias(float4* s_bias, const float* table, int heads, int h) {
  for (int i = threadIdx.x; i < 4 * O8_NT * 32; i += O8_THREADS) {
    const int wl = i / (O8_NT * 32), nt = (i >> 5) % O8_NT, ln = i & 31;
    const int q0 = wl * 16 + (ln >> 2), k0 = nt * 8 + 2 * (ln & 3);
    s_bias[i] = make_float4(table[o8_bias_index(q0, k0) * heads + h], table[o8_bias_index(q0, k0 + 1) * heads + h],
                            table[o8_bias_index(q0 + 8, k0) * heads + h], table[o8_bias_index(q0 + 8, k0 + 1) * heads + h]);
  }
}
__device__ __forceinline__ void o8_load_q(const Attn16Args& a, const O8Win& p, int h, uint32_t dst, const __nv_bfloat16* src,
                                          int ld) {
  const int ch = threadIdx.x & 3;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = (threadIdx.x >> 2) + 32 * k;
    cp_async16(dst + t32_off(i, ch), src + o8_qtok(a, p, i) * ld + h * 32 + ch * 8);
  }
}
__device__ __forceinline__ void o8_load_kv(const Attn16Args& a, const O8Win& p, int h, uint32_t sK, uint32_t sV) {
  const int hw = a.heads * 32;
  for (int c = threadIdx.x; c < O8_NK * 4; c += O8_THREADS) {
    const int slot = c >> 2, ch = c & 3;
    const int ky = slot / 12, kx = slot - ky * 12;
    const int y = p.wy * 8 - 2 + ky, x = p.wx * 8 - 2 + kx;
    const bool ok = y >= 0 && y < a.H && x >= 0 && x < a.W;
    const long long tok = ok ? ((long long)p.b * a.H + y) * a.W + x : 0;
    const __nv_bfloat16* src = a.qkv + tok * a.ld_qkv + hw + h * 32 + ch * 8;
    cp_async16_zfill(sK + t32_off(slot, ch), src, ok);
    cp_async16_zfill(sV + t32_off(slot, ch), src + hw, ok);
  }
}
__device__ __forceinline__ void o8_logits(const uint32_t (&aq)[2][4], uint32_t k_tile, const float4* bias, int lane,
                                          float (&s)[O8_NT][4]) {
#pragma unroll
  for (int nt = 0; nt < O8_NT; ++nt) {
    uint32_t b0, b1, b2, b3;
    ldsm_x4(k_tile + t32_off(nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
    const float4 bv = bias[nt * 32];
    s[nt][0] = bv.x; s[nt][1] = bv.y; s[nt][2] = bv.z; s[nt][3] = bv.w;
    mma_bf16(s[nt], aq[0], b0, b1);
    mma_bf16(s[nt], aq[1], b2, b3);
  }
}
__device__ __forceinline__ void o8_softmax(float (&s)[O8_NT][4]) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < O8_NT; ++nt) {
    m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
    m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  constexpr float kLog2e = 1.4426950408889634f;
  const float n0 = -m0 * kLog2e, n1 = -m1 * kLog2e;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < O8_NT; ++nt) {
    s[nt][0] = fast_ex2(fmaf(s[nt][0], kLog2e, n0));
    s[nt][1] = fast_ex2(fmaf(s[nt][1], kLog2e, n0));
    s[nt][2] = fast_ex2(fmaf(s[nt][2], kLog2e, n1));
    s[nt][3] = fast_ex2(fmaf(s[nt][3], kLog2e, n1));
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = fast_rcp(l0), i1 = fast_rcp(l1);
#pragma unroll
  for (int nt = 0; nt < O8_NT; ++nt) { s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1; }
}
// o[16 x 32] = A[16 x 144] (bf16 fragments built from fp32 s) * Bt[144 x 32]
__device__ __forceinline__ void o8_frag_times_tile(const float (&s)[O8_NT][4], uint32_t bt_tile, int lane, float (&o)[4][4]) {
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < O8_KS; ++kt) {
    uint32_t af[4];
    af[0] = pack_bf16(s[2 * kt][0], s[2 * kt][1]);
    af[1] = pack_bf16(s[2 * kt][2], s[2 * kt][3]);
    af[2] = pack_bf16(s[2 * kt + 1][0], s[2 * kt + 1][1]);
    af[3] = pack_bf16(s[2 * kt + 1][2], s[2 * kt + 1][3]);
    const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(bt_tile + t32_off(row, np * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma_bf16(o[2 * np], af, b0, b1);
      mma_bf16(o[2 * np + 1], af, b2, b3);
    }
  }
}
__device__ __forceinline__ uint32_t o8_p_off(int row, int chunk) { return uint32_t(row) * O8_PROW + (uint32_t(chunk) << 4); }

// ============================================================================ forward
__global__ void __launch_bounds__(O8_THREADS) win_attn_oca8_fwd_kernel(const Attn16Args a) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  float4* s_bias = reinterpret_cast<float4*>(smem_dyn);
  const uint32_t sQ = smem_u32(smem_dyn) + O8_BIAS_BYTES, sK = sQ + 4096, sV = sK + O8_NK * 64;
  const int h = blockIdx.y;
  const int nwin = a.B * (a.H >> 3) * (a.W >> 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  o8_fill_bias(s_bias, a.bias_table, a.heads, h);
  const bool ones_here = a.ones_col >= h * 32 && a.ones_col < h * 32 + 32;
  const int ones_c = a.ones_col - h * 32;
  for (int w = blockIdx.x; w < nwin; w += gridDim.x) {
    const O8Win p = o8_win(a, w);
    __syncthreads();   // previous window's tiles fully consumed; bias table visible
    o8_load_q(a, p, h, sQ, a.qkv, a.ld_qkv);
    o8_load_kv(a, p, h, sK, sV);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int r0 = warp * 16;
    uint32_t aq[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldsm_x4(sQ + t32_off(row, ks * 2 + (lane >> 4)), aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3]);
    }
    float s[O8_NT][4];
    o8_logits(aq, sK, s_bias + warp * O8_NT * 32 + lane, lane, s);
    o8_softmax(s);
    float o[4][4];
    o8_frag_times_tile(s, sV, lane, o);
    __syncwarp();   // this warp's Q rows are dead (fragments in registers): reuse them as the output staging rows
    store_frag_t32(sQ, r0, lane, o);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = lane + 32 * k, i = c >> 2, ch = c & 3;
      const long long tok = o8_qtok(a, p, r0 + i);
      uint4 v = lds128(sQ + t32_off(r0 + i, ch));
      if (ones_here && ch == (ones_c >> 3)) {
        const int word = (ones_c & 7) >> 1;
        const uint32_t keep = (ones_c & 1) ? 0x0000FFFFu : 0xFFFF0000u;
        const uint32_t one = (ones_c & 1) ? 0x3F800000u : 0x00003F80u;
        v.x = (word == 0) ? ((v.x & keep) | one) : v.x;
        v.y = (word == 1) ? ((v.y & keep) | one) : v.y;
        v.z = (word == 2) ? ((v.z & keep) | one) : v.z;
        v.w = (word == 3) ? ((v.w & keep) | one) : v.w;
      }
      *reinterpret_cast<uint4*>(a.out + tok * a.ld_o + h * 32 + ch * 8) = v;
    }
  }
}

// ============================================================================ backward
// dst[16 key slots x 32] = A^T B: A = [64 q][144 keys] tile (row stride O8_PROW; this call: keys k0..k0+15),
// B = t32 tile [64 q][32]; contraction over the 64 queries.
__device__ __forceinline__ void o8_tileT_times_tile(uint32_t a_tile, uint32_t b_tile, int k0, int lane, float (&o)[4][4]) {
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t af[4];
    {
      const int i = lane >> 3;
      const int row = kt * 16 + (lane & 7) + ((i >> 1) & 1) * 8;  // query
      const int col = k0 + (i & 1) * 8;                           // key
      ldsm_x4_t(a_tile + o8_p_off(row, col >> 3), af[0], af[1], af[2], af[3]);
    }
    const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(b_tile + t32_off(row, np * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma_bf16(o[2 * np], af, b0, b1);
      mma_bf16(o[2 * np + 1], af, b2, b3);
    }
  }
}
// write a [16 x 32] fp32 fragment tile as bf16 rows of a row-major [*, 32] global matrix
__device__ __forceinline__ void o8_store_frag_global(__nv_bfloat16* row_g, __nv_bfloat16* row_g8, int lane,
                                                     const float (&o)[4][4]) {
  const int t = lane & 3;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    *reinterpret_cast<uint32_t*>(row_g + n * 8 + 2 * t) = pack_bf16(o[n][0], o[n][1]);
    *reinterpret_cast<uint32_t*>(row_g8 + n * 8 + 2 * t) = pack_bf16(o[n][2], o[n][3]);
  }
}

__global__ void __launch_bounds__(O8_THREADS, 2) win_attn_oca8_bwd_kernel(const Attn16Args a) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  __shared__ float s_delta[64];
  float4* s_bias = reinterpret_cast<float4*>(smem_dyn);
  const uint32_t sQ = smem_u32(smem_dyn) + O8_BIAS_BYTES, sDO = sQ + 4096, sK = sDO + 4096, sV = sK + O8_NK * 64,
                 sP = sV + O8_NK * 64, sDS = sP + 64 * O8_PROW;
  const int h = blockIdx.y;
  const int nwin = a.B * (a.H >> 3) * (a.W >> 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2;
  o8_fill_bias(s_bias, a.bias_table, a.heads, h);
  float dbacc[O8_NT][4];
#pragma unroll
  for (int nt = 0; nt < O8_NT; ++nt) dbacc[nt][0] = dbacc[nt][1] = dbacc[nt][2] = dbacc[nt][3] = 0.f;

  for (int w = blockIdx.x; w < nwin; w += gridDim.x) {
    const O8Win p = o8_win(a, w);
    __syncthreads();
    o8_load_q(a, p, h, sQ, a.qkv, a.ld_qkv);
    o8_load_q(a, p, h, sDO, a.dout, a.ld_o);
    o8_load_kv(a, p, h, sK, sV);
    cp_async_commit();
    if (threadIdx.x < 64) {  // delta_i = sum_d dO[i,d] * O[i,d]   (the forward's ones column has a zero gradient)
      const long long tok = o8_qtok(a, p, threadIdx.x);
      const uint4* po = reinterpret_cast<const uint4*>(a.osave + tok * a.ld_o + h * 32);
      const uint4* pd = reinterpret_cast<const uint4*>(a.dout + tok * a.ld_o + h * 32);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 vo = po[k], vd = pd[k];
        const uint32_t wo[4] = {vo.x, vo.y, vo.z, vo.w}, wd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) acc += bf16_lo(wo[e]) * bf16_lo(wd[e]) + bf16_hi(wo[e]) * bf16_hi(wd[e]);
      }
      s_delta[threadIdx.x] = acc;
    }
    cp_async_wait<0>();
    __syncthreads();
    const int r0 = warp * 16;
    uint32_t aq[2][4], ad[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldsm_x4(sQ + t32_off(row, ks * 2 + (lane >> 4)), aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3]);
      ldsm_x4(sDO + t32_off(row, ks * 2 + (lane >> 4)), ad[ks][0], ad[ks][1], ad[ks][2], ad[ks][3]);
    }
    const float d0 = s_delta[r0 + g], d1 = s_delta[r0 + g + 8];
    // ---- phase A: this warp's 16 query rows x 144 keys
    float s[O8_NT][4];
    o8_logits(aq, sK, s_bias + warp * O8_NT * 32 + lane, lane, s);
    o8_softmax(s);   // s = P
#pragma unroll
    for (int nt = 1; nt < O8_NT; nt += 2) store_frag_pair(sP, r0, nt - 1, lane, s[nt - 1], s[nt], o8_p_off);   // P (bf16)
#pragma unroll
    for (int nt = 0; nt < O8_NT; ++nt) {  // dP = dO V^T; dS = P * (dP - delta), in place
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sV + t32_off(nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
      float dp[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16(dp, ad[0], b0, b1);
      mma_bf16(dp, ad[1], b2, b3);
      s[nt][0] *= (dp[0] - d0);
      s[nt][1] *= (dp[1] - d0);
      s[nt][2] *= (dp[2] - d1);
      s[nt][3] *= (dp[3] - d1);
#pragma unroll
      for (int e = 0; e < 4; ++e) dbacc[nt][e] += s[nt][e];
      if (nt & 1) store_frag_pair(sDS, r0, nt - 1, lane, s[nt - 1], s[nt], o8_p_off);
    }
    // dQ rows = dS (bf16) * K, written straight from the fragments (a quad covers 16 contiguous bytes per row)
    {
      float dq[4][4];
      o8_frag_times_tile(s, sK, lane, dq);
      __nv_bfloat16* q0 = a.dqkv + o8_qtok(a, p, r0 + g) * a.ld_qkv + h * 32;
      __nv_bfloat16* q1 = a.dqkv + o8_qtok(a, p, r0 + g + 8) * a.ld_qkv + h * 32;
      o8_store_frag_global(q0, q1, lane, dq);
    }
    __syncthreads();
    // ---- phase B: dK = dS^T Q, dV = P^T dO for 9 groups of 16 key slots (per-window rows, gathered afterwards)
    for (int rg = warp; rg < O8_KS; rg += 4) {
      float o[4][4];
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        o8_tileT_times_tile(m ? sP : sDS, m ? sDO : sQ, rg * 16, lane, o);
        __nv_bfloat16* base = a.dkv_win + ((((size_t)w * a.heads + h) * 2 + m) * O8_NK + rg * 16 + g) * 32;
        o8_store_frag_global(base, base + 8 * 32, lane, o);
      }
    }
  }
  // per-CTA sums of dS in fragment order: [gridDim.x][heads][4 warps][18][32 lanes] float4
  float4* scratch = reinterpret_cast<float4*>(a.dbias_scratch) +
                    (((size_t)blockIdx.x * a.heads + h) * 4 + warp) * (O8_NT * 32) + lane;
#pragma unroll
  for (int nt = 0; nt < O8_NT; ++nt) scratch[nt * 32] = make_float4(dbacc[nt][0], dbacc[nt][1], dbacc[nt][2], dbacc[nt][3]);
}

// dense[h][q][k] = sum over CTAs of the fragment-ordered partial sums
__global__ void oca8_dbias_dense_kernel(const float* __restrict__ scratch, int gx, int heads, float* __restrict__ dense) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * 64 * O8_NK) return;
  const int h = idx / (64 * O8_NK), r = idx % (64 * O8_NK), q = r / O8_NK, k = r % O8_NK;
  const int warp = q >> 4, rq = q & 15, g = rq & 7, rowsel = rq >> 3;
  const int nt = k >> 3, t = (k & 7) >> 1, e = rowsel * 2 + (k & 1);
  const size_t off = (((size_t)h * 4 + warp) * O8_NT + nt) * 128 + (g * 4 + t) * 4 + e;
  const size_t stride = (size_t)heads * 4 * O8_NT * 128;
  float acc = 0.f;
  for (int c = 0; c < gx; ++c) acc += scratch[c * stride + off];
  dense[idx] = acc;
}
// d_table[tbl][h] = sum of dense[h][q][k] over the (q, k) pairs whose (wrapped) reference index is tbl
__global__ void oca8_dbias_table_kernel(const float* __restrict__ dense, int heads, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O8_TBL * heads) return;
  const int tbl = idx / heads, h = idx % heads;
  // un-wrap: raw = ry*19 + rx with ry, rx in [-10, 8]; raw in [-200, 160]; negative raws were stored at raw + 361
  const int raw = tbl > 160 ? tbl - O8_TBL : tbl;
  const int num = raw + 10;                       // = ry*19 + (rx + 10), rx + 10 in [0, 18]
  const int ry = (num >= 0) ? num / 19 : -((-num + 18) / 19);
  const int rx = raw - ry * 19;
  float acc = 0.f;
  if (ry >= -10 && ry <= 8 && rx >= -10 && rx <= 8) {
    for (int qy = 0; qy < 8; ++qy) {
      const int ky = qy + ry + 3;
      if (ky < 0 || ky >= 12) continue;
      for (int qx = 0; qx < 8; ++qx) {
        const int kx = qx + rx + 3;
        if (kx < 0 || kx >= 12) continue;
        acc += dense[((size_t)h * 64 + qy * 8 + qx) * O8_NK + ky * 12 + kx];
      }
    }
  }
  out[idx] = acc;
}
// d_qkv[tok][K|V] = sum over the (up to 4) overlapping 12x12 key windows of their per-window dK/dV rows
__global__ void oca8_kv_gather_kernel(const __nv_bfloat16* __restrict__ dkv_win, __nv_bfloat16* __restrict__ dqkv,
                                      int ld_qkv, int B, int H, int W, int heads) {
  const long long total = (long long)B * H * W * heads * 2 * 4;
  const int nwy = H >> 3, nwx = W >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i & 3);
    long long r = i >> 2;
    const int h = int(r % heads); r /= heads;
    const int m = int(r & 1); r >>= 1;
    const long long tok = r;
    const int x = int(tok % W), y = int((tok / W) % H), b = int(tok / ((long long)W * H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int wy_hi = min((y + 2) >> 3, nwy - 1), wx_hi = min((x + 2) >> 3, nwx - 1);
    for (int wy = wy_hi; wy >= 0 && wy * 8 + 10 > y; --wy)
      for (int wx = wx_hi; wx >= 0 && wx * 8 + 10 > x; --wx) {
        const int slot = (y - wy * 8 + 2) * 12 + (x - wx * 8 + 2);
        const size_t win = ((size_t)b * nwy + wy) * nwx + wx;
        const uint4 v = *reinterpret_cast<const uint4*>(dkv_win + (((win * heads + h) * 2 + m) * O8_NK + slot) * 32 + ch * 8);
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[2 * e] += bf16_lo(wv[e]); acc[2 * e + 1] += bf16_hi(wv[e]); }
      }
    *reinterpret_cast<uint4*>(dqkv + tok * ld_qkv + (1 + m) * heads * 32 + h * 32 + ch * 8) =
        make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}

}  // namespace srk
