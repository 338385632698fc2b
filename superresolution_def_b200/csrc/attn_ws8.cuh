// attn_ws8.cuh — fused (shifted-)window attention core for 8x8 windows (64 tokens), forward + backward.
//
// Replaces (reference, models/architecture_swin.py): torch.roll :130-133,143-146; window_partition :27-31;
// window_reverse :33-37; and inside WindowAttention.forward :75-93 the q@k^T bmm, relative-position-bias
// gather+add, softmax, attn@v bmm and the head-merge transpose.  Shift / partition / reverse are pure
// address arithmetic here; S and P never touch HBM.
//
// Layouts (token-major, bf16):  qkv [T, 3*heads*32]  (q|k|v, each head padded 30->32 with zeros, q already
// scaled by head_dim^-0.5 through the projection weights);  out [T, heads*32].
// This kernel is HBM-bound by construction (32 FLOP/B), so it uses register-resident
// mma.sync tiles: a 64x64 logit tile fits the register file and needs no TMEM round trip.
// Grid: (window groups, heads); each CTA keeps one head and walks windows, double-buffering its loads.
#pragma once
#include "srk_ptx.cuh"

namespace srk {

struct AttnArgs {
  const __nv_bfloat16* qkv;   // [T, ld_qkv]
  const __nv_bfloat16* dout;  // bwd: [T, ld_o] gradient of out
  __nv_bfloat16* out;         // fwd: [T, ld_o]
  __nv_bfloat16* dqkv;        // bwd: [T, ld_qkv]
  const float* bias_table;    // [(2*8-1)^2 = 225][heads] fp32 (reference layout)
  float* dbias_partials;      // bwd: [gridDim.x][heads][225]
  int B, H, W, heads, shift;
  int ld_qkv, ld_o;
  int ones_col;               // fwd: column of `out` forced to 1.0 (bias-folding column), -1: none
  int mask;                   // 1: add HAT's 0/-100 shifted-window mask (hat_arch.py:921-940); SwinIR never masks (:138)
};

constexpr int ATT_THREADS = 128;
constexpr int ATT_TILE = 64 * 32 * 2;  // one [64 tokens x 32] bf16 tile (4 KB)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// [64 x 32] bf16 tile, 64-byte rows, 16-byte chunks XOR-swizzled so ldmatrix is conflict-free.
__device__ __forceinline__ uint32_t t32_off(int row, int chunk) {
  return uint32_t(row) * 64u + (uint32_t(chunk ^ ((row >> 1) & 3)) << 4);
}
// [64 x 64] bf16 tile, 128-byte rows.
__device__ __forceinline__ uint32_t t64_off(int row, int chunk) {
  return uint32_t(row) * 128u + (uint32_t(chunk ^ (row & 7)) << 4);
}

// token row index (in the un-shifted [B,H,W] token grid) of local token i of window w
__device__ __forceinline__ int window_token(const AttnArgs& a, int w, int i) {
  const int nwx = a.W >> 3, nwy = a.H >> 3;
  const int b = w / (nwx * nwy);
  const int r = w - b * nwx * nwy;
  const int wy = r / nwx, wx = r - wy * nwx;
  int y = wy * 8 + (i >> 3) + a.shift;
  int x = wx * 8 + (i & 7) + a.shift;
  if (y >= a.H) y -= a.H;
  if (x >= a.W) x -= a.W;
  return (b * a.H + y) * a.W + x;
}

__device__ __forceinline__ int rel_index(int i, int j) {
  return ((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7);
}

// S[16 x 64] (this warp's 16 query rows) = Q K^T + bias ; returns fp32 logits in s[nt][4]
// Shift mask of HAT.calculate_mask for ws = 8, shift = 4 on fragments: in the shifted frame the last window row /
// column is split at local token 4, so query (qy, qx) and key (ky, kx) lie in different regions iff
// last_y && (qy>=4) != (ky>=4)  or  last_x && (qx>=4) != (kx>=4).  qy = 2*warp + rowsel, qx = g, ky = nt, kx = 2t + e.
__device__ __forceinline__ void add_shift_mask(float (&s)[8][4], int r0, int lane, bool last_y, bool last_x) {
  const int g = lane >> 2, t = lane & 3;
  const bool xd = last_x && ((g >= 4) != (t >= 2));
  const bool qlow = r0 < 32;   // query rows 0..3
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const bool yd = last_y && (qlow != (nt < 4));
    const float m = (yd || xd) ? -100.0f : 0.0f;
    s[nt][0] += m; s[nt][1] += m; s[nt][2] += m; s[nt][3] += m;
  }
}
// (last window row, last window column) flags of window w of a shifted, masked block
__device__ __forceinline__ void window_edge_flags(const AttnArgs& a, int w, bool& last_y, bool& last_x) {
  const int nwx = a.W >> 3, nwy = a.H >> 3;
  const int r = w % (nwx * nwy);
  last_y = (r / nwx) == nwy - 1;
  last_x = (r % nwx) == nwx - 1;
}

__device__ __forceinline__ void qk_logits(uint32_t q_tile, uint32_t k_tile, const float* s_bias, int r0, int lane,
                                          float (&s)[8][4]) {
  uint32_t aq[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int chunk = ks * 2 + (lane >> 4);
    ldsm_x4(q_tile + t32_off(row, chunk), aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3]);
  }
  const int g = lane >> 2, t = lane & 3;
  // rel_index(i, j) for i = r0 + g + 8*rowsel (window row 2*warp + rowsel, column g) and j = nt*8 + 2t + e (window row
  // nt, column 2t + e) is  base + (rowsel - nt) * 15 - e : one LDS with an immediate offset per logit
  const float* bp = s_bias + ((r0 >> 3) + 7) * 15 + (g - 2 * t + 7);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b0, b1, b2, b3;
    ldsm_x4(k_tile + t32_off(nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
    s[nt][0] = bp[(0 - nt) * 15];
    s[nt][1] = bp[(0 - nt) * 15 - 1];
    s[nt][2] = bp[(1 - nt) * 15];
    s[nt][3] = bp[(1 - nt) * 15 - 1];
    mma_bf16(s[nt], aq[0], b0, b1);
    mma_bf16(s[nt], aq[1], b2, b3);
  }
}

// in-place row softmax of the 16x64 fragment tile (rows g and g+8 of the quad)
__device__ __forceinline__ void softmax_rows(float (&s)[8][4]) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
    m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;
  const float n0 = -m0 * kLog2e, n1 = -m1 * kLog2e;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
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
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1;
  }
}

// O[16 x 32] = A[16 x 64] (bf16 fragments built from fp32 s) * Bt[64 x 32] where Bt is a [key][d] tile
__device__ __forceinline__ void frag_times_tile(const float (&s)[8][4], uint32_t bt_tile, int lane, float (&o)[4][4]) {
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t a[4];
    a[0] = pack_bf16(s[2 * kt][0], s[2 * kt][1]);
    a[1] = pack_bf16(s[2 * kt][2], s[2 * kt][3]);
    a[2] = pack_bf16(s[2 * kt + 1][0], s[2 * kt + 1][1]);
    a[3] = pack_bf16(s[2 * kt + 1][2], s[2 * kt + 1][3]);
    const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(bt_tile + t32_off(row, np * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma_bf16(o[2 * np], a, b0, b1);
      mma_bf16(o[2 * np + 1], a, b2, b3);
    }
  }
}

__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};"
               ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// Store two adjacent 16x8 fp32 accumulator n-tiles (c0 = n-tile n, c1 = n-tile n+1; rows r0..r0+15) as bf16 with ONE
// stmatrix.x4: matrices (rows 0-7, n), (rows 8-15, n), (rows 0-7, n+1), (rows 8-15, n+1).  `row_off(row, chunk)` is the
// tile's swizzled byte offset of 16-byte chunk `chunk` of row `row`.
template <typename OffFn>
__device__ __forceinline__ void store_frag_pair(uint32_t tile, int r0, int n, int lane, const float (&c0)[4],
                                                const float (&c1)[4], OffFn row_off) {
  const int m = lane >> 3;
  const uint32_t addr = tile + row_off(r0 + (m & 1) * 8 + (lane & 7), n + (m >> 1));
  stsm_x4(addr, pack_bf16(c0[0], c0[1]), pack_bf16(c0[2], c0[3]), pack_bf16(c1[0], c1[1]), pack_bf16(c1[2], c1[3]));
}
// store a [16 x 32] fp32 fragment tile as bf16 into a swizzled t32 tile (rows r0..r0+15)
__device__ __forceinline__ void store_frag_t32(uint32_t tile, int r0, int lane, const float (&o)[4][4]) {
  store_frag_pair(tile, r0, 0, lane, o[0], o[1], t32_off);
  store_frag_pair(tile, r0, 2, lane, o[2], o[3], t32_off);
}

// Each thread always handles the same two token rows of a window -- i0 = 16 * warp + lane / 4 and i0 + 8, i.e. rows of the 16-row
// block its own warp computes -- and 16-byte chunk lane % 4 of every [64 x 32] tile, so a window's token addresses are
// computed once per thread and window (two wrapped coordinates) and serve the loads AND the warp-private result stores.
struct WinToks { long long t0, t1; };
__device__ __forceinline__ int tok_row0() { return ((threadIdx.x >> 5) << 4) + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ WinToks window_toks(const AttnArgs& a, int w) {
  WinToks r;
  r.t0 = window_token(a, w, tok_row0());
  r.t1 = window_token(a, w, tok_row0() + 8);
  return r;
}
// async-load the q,k,v tiles (column offsets col0 + {0,1,2}*heads*32 of src0) and optionally a 4th tile from src1
__device__ __forceinline__ void load_window_tiles(const AttnArgs& a, const WinToks& tk, uint32_t dst,
                                                  const __nv_bfloat16* src0, int ld0, int col0,
                                                  const __nv_bfloat16* src1, int ld1, int col1) {
  const int hw = a.heads * 32;
  const int i0 = tok_row0(), ch = threadIdx.x & 3;
  const uint32_t d0 = dst + t32_off(i0, ch), d1 = dst + t32_off(i0 + 8, ch);
  const __nv_bfloat16* p0 = src0 + tk.t0 * ld0 + col0 + ch * 8;
  const __nv_bfloat16* p1 = src0 + tk.t1 * ld0 + col0 + ch * 8;
#pragma unroll
  for (int tile = 0; tile < 3; ++tile) {
    cp_async16(d0 + tile * ATT_TILE, p0 + tile * hw);
    cp_async16(d1 + tile * ATT_TILE, p1 + tile * hw);
  }
  if (src1 != nullptr) {
    cp_async16(d0 + 3 * ATT_TILE, src1 + tk.t0 * ld1 + col1 + ch * 8);
    cp_async16(d1 + 3 * ATT_TILE, src1 + tk.t1 * ld1 + col1 + ch * 8);
  }
}
// A warp's finished [16 x 32] fp32 fragment tile -> bf16 in its private 1 KB staging tile -> 16-byte global stores of its own
// 16 token rows (columns col .. col+31 of `dst`).  Only the warp itself touches the staging tile: __syncwarp, no CTA barrier.
__device__ __forceinline__ void store_rows16(uint32_t stage, const uint8_t* stage_ptr, int lane, const float (&o)[4][4],
                                             __nv_bfloat16* dst, int ld, int col, const WinToks& tk) {
  __syncwarp();
  store_frag_t32(stage, 0, lane, o);
  __syncwarp();
  const int i = lane >> 2, ch = lane & 3;
  const uint4 v0 = *reinterpret_cast<const uint4*>(stage_ptr + t32_off(i, ch));
  const uint4 v1 = *reinterpret_cast<const uint4*>(stage_ptr + t32_off(i + 8, ch));
  *reinterpret_cast<uint4*>(dst + tk.t0 * ld + col + ch * 8) = v0;
  *reinterpret_cast<uint4*>(dst + tk.t1 * ld + col + ch * 8) = v1;
}

// ============================================================================ forward
__global__ void __launch_bounds__(ATT_THREADS) win_attn_ws8_fwd_kernel(const AttnArgs a) {
  __shared__ __align__(128) uint8_t s_in[2][3 * ATT_TILE];
  __shared__ __align__(128) uint8_t s_out[4][1024];   // per-warp staging of its 16 output rows
  __shared__ float s_bias[225];
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.y;
  const int nwin = a.B * (a.H >> 3) * (a.W >> 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 225; i += ATT_THREADS) s_bias[i] = a.bias_table[i * a.heads + h];
  const bool ones_here = a.ones_col >= h * 32 && a.ones_col < h * 32 + 32;
  const int ones_c = a.ones_col - h * 32;
  const bool masked = a.mask != 0 && a.shift > 0;

  int w = blockIdx.x;
  int buf = 0;
  WinToks tk{0, 0}, tkn{0, 0};
  if (w < nwin) {
    tk = window_toks(a, w);
    load_window_tiles(a, tk, smem_u32(s_in[0]), a.qkv, a.ld_qkv, h * 32, nullptr, 0, 0);
  }
  cp_async_commit();
  for (; w < nwin; w += gridDim.x, buf ^= 1, tk = tkn) {
    cp_async_wait<0>();
    __syncthreads();   // this window's tiles have landed; every warp is done with the previous window (s_in[buf ^ 1] is free)
    const int wn = w + gridDim.x;
    if (wn < nwin) {
      tkn = window_toks(a, wn);
      load_window_tiles(a, tkn, smem_u32(s_in[buf ^ 1]), a.qkv, a.ld_qkv, h * 32, nullptr, 0, 0);
    }
    cp_async_commit();
    const uint32_t qt = smem_u32(s_in[buf]), kt = qt + ATT_TILE, vt = qt + 2 * ATT_TILE;
    const int r0 = warp * 16;
    float s[8][4];
    qk_logits(qt, kt, s_bias, r0, lane, s);
    if (masked) {
      bool ly, lx;
      window_edge_flags(a, w, ly, lx);
      if (ly || lx) add_shift_mask(s, r0, lane, ly, lx);
    }
    softmax_rows(s);
    float o[4][4];
    frag_times_tile(s, vt, lane, o);
    __syncwarp();
    store_frag_t32(smem_u32(s_out[warp]), 0, lane, o);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = (lane >> 2) + 8 * k, ch = lane & 3;
      const long long tok = k ? tk.t1 : tk.t0;
      uint4 v = *reinterpret_cast<const uint4*>(s_out[warp] + t32_off(i, ch));
      if (ones_here && ch == (ones_c >> 3)) {  // bias-folding column of the following projection := 1.0
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
  cp_async_wait<0>();
}

// ============================================================================ backward
struct AttnBwdSmem {  // 61.8 KB: three CTAs per SM (the register file allows three as well)
  uint8_t in[2][4 * ATT_TILE];  // q, k, v, dO (double-buffered)
  uint8_t p[64 * 128];          // P  (bf16) [q][key]
  uint8_t ds[64 * 128];         // dS (bf16) [q][key]
  uint8_t stage[4][3][1024];    // per warp: its 16 rows of dq, dk, dv on their way to global memory
  float bias[225];
  float dbias[225];
};

// C[16 keys x 32] += A^T * B  with A stored as a t64 tile [q][key] (this warp's keys k0..k0+15) and
// B a t32 tile [q][d]; contraction over the 64 queries.
__device__ __forceinline__ void tileT_times_tile(uint32_t a_tile, uint32_t b_tile, int k0, int lane, float (&o)[4][4]) {
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t af[4];
    {
      const int i = lane >> 3;
      const int row = kt * 16 + (lane & 7) + ((i >> 1) & 1) * 8;  // query
      const int col = k0 + (i & 1) * 8;                           // key
      ldsm_x4_t(a_tile + t64_off(row, col >> 3), af[0], af[1], af[2], af[3]);
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

__global__ void __launch_bounds__(ATT_THREADS, 3) win_attn_ws8_bwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  AttnBwdSmem& sm = *reinterpret_cast<AttnBwdSmem*>(smem_dyn);
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.y;
  const int nwin = a.B * (a.H >> 3) * (a.W >> 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < 225; i += ATT_THREADS) {
    sm.bias[i] = a.bias_table[i * a.heads + h];
    sm.dbias[i] = 0.f;
  }
  const bool masked = a.mask != 0 && a.shift > 0;
  float dbacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dbacc[nt][0] = dbacc[nt][1] = dbacc[nt][2] = dbacc[nt][3] = 0.f;

  int w = blockIdx.x;
  int buf = 0;
  WinToks tk{0, 0}, tkn{0, 0};
  if (w < nwin) {
    tk = window_toks(a, w);
    load_window_tiles(a, tk, smem_u32(sm.in[0]), a.qkv, a.ld_qkv, h * 32, a.dout, a.ld_o, h * 32);
  }
  cp_async_commit();
  const int hw = a.heads * 32;
  const uint32_t stg = smem_u32(sm.stage[warp][0]);
  for (; w < nwin; w += gridDim.x, buf ^= 1, tk = tkn) {
    cp_async_wait<0>();
    __syncthreads();   // (1) this window's tiles have landed; every warp is done with the previous window (in[buf ^ 1], P, dS free)
    const int wn = w + gridDim.x;
    if (wn < nwin) {
      tkn = window_toks(a, wn);
      load_window_tiles(a, tkn, smem_u32(sm.in[buf ^ 1]), a.qkv, a.ld_qkv, h * 32, a.dout, a.ld_o, h * 32);
    }
    cp_async_commit();
    const uint32_t qt = smem_u32(sm.in[buf]), kt = qt + ATT_TILE, vt = qt + 2 * ATT_TILE, dot = qt + 3 * ATT_TILE;
    const uint32_t pt = smem_u32(sm.p), dst = smem_u32(sm.ds);
    const int r0 = warp * 16;
    // ---- phase A: this warp's 16 query rows
    float s[8][4];
    qk_logits(qt, kt, sm.bias, r0, lane, s);
    if (masked) {
      bool ly, lx;
      window_edge_flags(a, w, ly, lx);
      if (ly || lx) add_shift_mask(s, r0, lane, ly, lx);
    }
    softmax_rows(s);  // s = P (fp32)
    // dP = dO V^T  (A = dO rows, B = V as [key][d], same access pattern as K in QK^T)
    float dp[8][4];
    {
      uint32_t ad[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(dot + t32_off(row, ks * 2 + (lane >> 4)), ad[ks][0], ad[ks][1], ad[ks][2], ad[ks][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(vt + t32_off(nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        mma_bf16(dp[nt], ad[0], b0, b1);
        mma_bf16(dp[nt], ad[1], b2, b3);
      }
    }
    // P (bf16) -> smem; delta = rowsum(dP * P); dS = P * (dP - delta)
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      // the forward multiplies V by bf16(P); its gradient dP is therefore taken w.r.t. the rounded P
      d0 += dp[nt][0] * s[nt][0] + dp[nt][1] * s[nt][1];
      d1 += dp[nt][2] * s[nt][2] + dp[nt][3] * s[nt][3];
      if (nt & 1) store_frag_pair(pt, r0, nt - 1, lane, s[nt - 1], s[nt], t64_off);
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] *= (dp[nt][0] - d0);
      s[nt][1] *= (dp[nt][1] - d0);
      s[nt][2] *= (dp[nt][2] - d1);
      s[nt][3] *= (dp[nt][3] - d1);
#pragma unroll
      for (int e = 0; e < 4; ++e) dbacc[nt][e] += s[nt][e];
      if (nt & 1) store_frag_pair(dst, r0, nt - 1, lane, s[nt - 1], s[nt], t64_off);
    }
    // dQ rows = dS (bf16) * K, straight out through this warp's staging tile
    {
      float dq[4][4];
      frag_times_tile(s, kt, lane, dq);
      store_rows16(stg, sm.stage[warp][0], lane, dq, a.dqkv, a.ld_qkv, h * 32, tk);
    }
    __syncthreads();   // (2) P and dS of all 64 queries are in shared memory
    // ---- phase B: this warp's 16 key rows
    {
      float dk[4][4];
      tileT_times_tile(dst, qt, r0, lane, dk);   // dK = dS^T Q
      store_rows16(stg + 1024, sm.stage[warp][1], lane, dk, a.dqkv, a.ld_qkv, hw + h * 32, tk);
    }
    {
      float dv[4][4];
      tileT_times_tile(pt, dot, r0, lane, dv);   // dV = P^T dO
      store_rows16(stg + 2048, sm.stage[warp][2], lane, dv, a.dqkv, a.ld_qkv, 2 * hw + h * 32, tk);
    }
  }
  cp_async_wait<0>();
  // fold the per-thread dS sums into the (2*8-1)^2 table entries of this head
  {
    const int r0 = warp * 16;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int i0 = r0 + g, j0 = nt * 8 + 2 * t;
      atomicAdd(&sm.dbias[rel_index(i0, j0)], dbacc[nt][0]);
      atomicAdd(&sm.dbias[rel_index(i0, j0 + 1)], dbacc[nt][1]);
      atomicAdd(&sm.dbias[rel_index(i0 + 8, j0)], dbacc[nt][2]);
      atomicAdd(&sm.dbias[rel_index(i0 + 8, j0 + 1)], dbacc[nt][3]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 225; i += ATT_THREADS)
    a.dbias_partials[(size_t(blockIdx.x) * a.heads + h) * 225 + i] = sm.dbias[i];
}

}  // namespace srk
