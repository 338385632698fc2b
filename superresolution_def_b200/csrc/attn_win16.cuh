// attn_win16.cuh — fused window attention core for HAT's 16x16 windows (256 query tokens), forward + backward:
//   MODE_SELF : (S)W-MSA of HAB   — keys = the same (cyclically shifted) window, shift mask 0/-100 computed from
//               coordinates in-kernel (reference: HAT.calculate_mask hat_arch.py:921-940, added at :183-187)
//   MODE_OCA  : overlapping cross-attention of OCAB — keys/values = the 24x24 halo window around the query window,
//               ZERO outside the image after projection (nn.Unfold padding, hat_arch.py:377,408-410), no mask,
//               (16+24-1)^2 bias table indexed with the reference's wrap-around offsets (:896-919).
//
// Replaces (reference, models/hat_arch/hat_arch.py): window_partition/window_reverse :97-126, torch.roll :280-302,
// nn.Unfold + einops rearrange :408-409, and inside WindowAttention.forward :175-193 / OCAB.forward :419-428 the
// q@k^T bmm, relative-position-bias gather+add, mask add, softmax, attn@v bmm and the head-merge transpose.
// Shift / partition / reverse / unfold are address arithmetic; S and P never touch HBM.
//
// Layouts: qkv [T, 3*heads*32] bf16 token-major (q|k|v, heads padded 30->32, q pre-scaled by the projection weights),
// out [T, heads*32] bf16, lse [heads][T] fp32 (row log-sum-exp, consumed by the backward).
//
// Design: flash-style tiles of 64 key slots on register-resident mma.sync fragments (exp-bound, not MMA-bound: 256x256
// logits need 64 Ki ex2 per (window, head) against 8.4 MFLOP of MMA, so TMEM round trips would buy nothing here).
// MODE_OCA lays the 24 key columns out as two 16-wide bands (the second half-empty, masked), i.e. 12 tiles of
// (2 key rows x 2 bands x 16): every 16x16 (query-row, key-row) block then has the same shape as in MODE_SELF, which
// lets the bias-table gradient be a tensor-core diagonal sum (see diag_mma below).
#pragma once
#include "attn_ws8.cuh"

namespace srk {

enum { MODE_SELF = 0, MODE_OCA = 1 };

template <int MODE>
struct A16 {
  static constexpr int WS = 16;
  static constexpr int WSE = (MODE == MODE_SELF) ? 16 : 24;
  static constexpr int NKT = (MODE == MODE_SELF) ? 4 : 12;   // key tiles of 64 slots
  static constexpr int NS = NKT * 64;                          // key slots (MODE_OCA: 768, 576 real)
  static constexpr int TSIDE = WS + WSE - 1;                   // 31 / 39
  static constexpr int TBL = TSIDE * TSIDE;                    // 961 / 1521
  static constexpr int PAD = (WSE - WS) / 2;                   // 0 / 4
};

struct Attn16Args {
  const __nv_bfloat16* qkv;   // [T, ld_qkv]
  __nv_bfloat16* out;         // fwd: [T, ld_o]
  float* lse;                 // [heads][T]
  const __nv_bfloat16* dout;  // bwd: [T, ld_o]
  const __nv_bfloat16* osave; // bwd: forward output [T, ld_o]
  __nv_bfloat16* dqkv;        // bwd: [T, ld_qkv] (MODE_OCA: only the q third is written here)
  __nv_bfloat16* dkv_win;     // bwd MODE_OCA: [nwin*heads][2][768][32] per-window dK/dV (gathered afterwards)
  const float* bias_table;    // [TBL][heads]
  float* dbias_scratch;       // bwd: [gridDim.x][heads][NKT][16 warps][32 lanes][4]
  int B, H, W, heads, shift;
  int ld_qkv, ld_o;
  int ones_col;
  long long T;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// window (b, wy, wx) of window index w
struct WinPos { int b, wy, wx; };
__device__ __forceinline__ WinPos win_pos(const Attn16Args& a, int w) {
  const int nwx = a.W >> 4, nwy = a.H >> 4;
  WinPos p;
  p.b = w / (nwx * nwy);
  const int r = w - p.b * nwx * nwy;
  p.wy = r / nwx;
  p.wx = r - p.wy * nwx;
  return p;
}
// token row of query (ty, tx) of the window (cyclic shift applied: shifted-frame coordinate -> image coordinate)
__device__ __forceinline__ long long q_token(const Attn16Args& a, const WinPos& p, int ty, int tx) {
  int y = p.wy * 16 + ty + a.shift, x = p.wx * 16 + tx + a.shift;
  if (y >= a.H) y -= a.H;
  if (x >= a.W) x -= a.W;
  return ((long long)p.b * a.H + y) * a.W + x;
}
// key slot -> (ky, kx) inside the key window, validity (MODE_OCA pads 24 columns to 2 x 16)
template <int MODE>
__device__ __forceinline__ void slot_key(int slot, int& ky, int& kx, bool& valid) {
  if (MODE == MODE_SELF) {
    ky = slot >> 4; kx = slot & 15; valid = true;
  } else {
    ky = (slot >> 6) * 2 + ((slot >> 5) & 1);
    kx = ((slot >> 4) & 1) * 16 + (slot & 15);
    valid = kx < 24;
  }
}
// token row of a key slot, or -1 when the key lies outside the image / is a padding slot (=> k = v = 0)
template <int MODE>
__device__ __forceinline__ long long k_token(const Attn16Args& a, const WinPos& p, int slot, bool& slot_valid) {
  int ky, kx;
  slot_key<MODE>(slot, ky, kx, slot_valid);
  if (MODE == MODE_SELF) return q_token(a, p, ky, kx);
  if (!slot_valid) return -1;
  const int y = p.wy * 16 - A16<MODE>::PAD + ky, x = p.wx * 16 - A16<MODE>::PAD + kx;
  if (y < 0 || y >= a.H || x < 0 || x >= a.W) return -1;
  return ((long long)p.b * a.H + y) * a.W + x;
}

// Relative-position bias table in shared memory, laid out so that the index of element (query row qy, column
// qx = g + 8*rowsel; key n-tile nt, column 2t + e of key tile kt) is  base(qy, kt, g, t) + CONST(nt, e, rowsel):
// one LDS with an immediate offset per logit, no per-element index arithmetic.
//   MODE_SELF: index = (qy-ky+15)*31 + (qx-kx+15)                                  (reference :882-894)
//   MODE_OCA : index = (ky-qy-7)*39 + (kx-qx-7) + 880; the reference's negative indices (which PyTorch wraps around
//              the table end, :896-919) are materialised by loading the table rotated by 880 entries.
template <int MODE>
__device__ __forceinline__ void load_bias_table(float* s_bias, const float* table, int heads, int h, int nthreads) {
  constexpr int TBL = A16<MODE>::TBL;
  for (int i = threadIdx.x; i < TBL; i += nthreads) {
    int src = i;
    if (MODE == MODE_OCA) { src = i - 880; if (src < 0) src += TBL; }
    s_bias[i] = table[src * heads + h];
  }
}
template <int MODE>
__device__ __forceinline__ int bias_base(int qy, int kt, int g, int t) {
  if (MODE == MODE_SELF) return (qy - kt * 4 + 15) * 31 + (g - 2 * t + 15);
  return (kt * 2 - qy - 7) * 39 + (2 * t - g - 7) + 880;
}
template <int MODE>
__device__ __forceinline__ constexpr int bias_const(int nt, int e, int rowsel) {
  return (MODE == MODE_SELF) ? (-(nt >> 1) * 31 - (nt & 1) * 8 - e + 8 * rowsel)
                             : ((nt >> 2) * 39 + ((nt >> 1) & 1) * 16 + (nt & 1) * 8 + e - 8 * rowsel);
}
// MODE_OCA: n-tiles 3 and 7 of every key tile are the padding half of the second 16-wide band (kx >= 24)
template <int MODE>
__device__ __forceinline__ constexpr bool nt_valid(int nt) { return MODE == MODE_SELF || (nt & 3) != 3; }

// Shift mask of HAT.calculate_mask (:921-940) for ws = 16, shift = 8, expressed on fragments.  In the shifted frame
// the regions split every window of the last window row / column at token 8, so query (qy, qx) and key (ky, kx) lie
// in different regions iff  last_y && (qy>=8) != (ky>=8)  or  last_x && (qx>=8) != (kx>=8).  On fragments (qy>=8) is
// warp-uniform, (ky>=8) is tile-uniform (kt>=2), (qx>=8) is the fragment row half and (kx>=8) the n-tile parity.
struct MaskCtx {
  bool any;       // window touches the last window row / column of a shifted block
  float same;     // additive term where row half == n-tile parity
  float diff;     // additive term where they differ
};
__device__ __forceinline__ MaskCtx mask_ctx(bool last_y, bool last_x, int qy, int kt) {
  MaskCtx m;
  const bool ydiff = last_y && ((qy >= 8) != (kt >= 2));
  m.any = last_y || last_x;
  m.same = ydiff ? -100.0f : 0.0f;
  m.diff = (ydiff || last_x) ? -100.0f : 0.0f;
  return m;
}

// S[16 x 64] logits of this warp's 16 query rows against key tile kt: Q K^T + bias (+ mask)
template <int MODE>
__device__ __forceinline__ void qk_tile(const uint32_t (&aq)[2][4], uint32_t k_tile, const float* bp, const MaskCtx& mk,
                                        int lane, float (&s)[8][4]) {
  // K fragments are double-buffered: the ldmatrix of n-tile nt+1 is issued before the MMAs of n-tile nt
  const uint32_t kb_base = k_tile + t32_off(lane & 7, lane >> 3);   // + nt * 8 rows = nt * 512 B (swizzle unchanged)
  uint32_t kb[2][4];
  ldsm_x4(kb_base, kb[0][0], kb[0][1], kb[0][2], kb[0][3]);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int cur = nt & 1, nxt = cur ^ 1;
    if (nt + 1 < 8 && nt_valid<MODE>(nt + 1))
      ldsm_x4(kb_base + (nt + 1) * 512, kb[nxt][0], kb[nxt][1], kb[nxt][2], kb[nxt][3]);
    if (!nt_valid<MODE>(nt)) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = -INFINITY;
      continue;
    }
    s[nt][0] = bp[bias_const<MODE>(nt, 0, 0)];
    s[nt][1] = bp[bias_const<MODE>(nt, 1, 0)];
    s[nt][2] = bp[bias_const<MODE>(nt, 0, 1)];
    s[nt][3] = bp[bias_const<MODE>(nt, 1, 1)];
    mma_bf16(s[nt], aq[0], kb[cur][0], kb[cur][1]);
    mma_bf16(s[nt], aq[1], kb[cur][2], kb[cur][3]);
  }
  if (MODE == MODE_SELF && mk.any) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float m01 = (nt & 1) ? mk.diff : mk.same, m23 = (nt & 1) ? mk.same : mk.diff;
      s[nt][0] += m01; s[nt][1] += m01; s[nt][2] += m23; s[nt][3] += m23;
    }
  }
}

// o[16 x 32] += A[16 x 64] (bf16 fragments built from fp32 s) * Bt[64 x 32]   (accumulating frag_times_tile)
__device__ __forceinline__ void frag_times_tile_acc(const float (&s)[8][4], uint32_t bt_tile, int lane, float (&o)[4][4]) {
  // 8 transposed B fragments (kk = 0..3, np = 0..1), double-buffered one step ahead of the MMAs
  const uint32_t b_base = bt_tile + t32_off((lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);
  auto frag_addr = [&](int step) {  // step = kk * 2 + np: +16 rows per kk (swizzle unchanged), chunk pair np toggles bit 1
    const int kk = step >> 1, np = step & 1;
    return (b_base + kk * 16 * 64) ^ (uint32_t(np) << 5);
  };
  uint32_t bf[2][4];
  ldsm_x4_t(frag_addr(0), bf[0][0], bf[0][1], bf[0][2], bf[0][3]);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const int step = kk * 2 + np, cur = step & 1, nxt = cur ^ 1;
      if (step + 1 < 8) ldsm_x4_t(frag_addr(step + 1), bf[nxt][0], bf[nxt][1], bf[nxt][2], bf[nxt][3]);
      mma_bf16(o[2 * np], a, bf[cur][0], bf[cur][1]);
      mma_bf16(o[2 * np + 1], a, bf[cur][2], bf[cur][3]);
    }
  }
}

// Token-row table of one window in shared memory: s_qtok[256] (query tokens == MODE_SELF key tokens), filled once per
// window so that loads and stores are table look-ups instead of per-transfer coordinate arithmetic.  MODE_OCA key
// slots are decoded on the fly (a table would cost 3 KB and the second resident CTA of the forward kernel).
__device__ __forceinline__ void fill_token_table(const Attn16Args& a, const WinPos& p, int* s_qtok, int nthreads) {
  for (int i = threadIdx.x; i < 256; i += nthreads) s_qtok[i] = int(q_token(a, p, i >> 4, i & 15));
}
// cooperative async load of the K and V slot tiles of one window (rows = key slots, zero-filled where absent)
template <int MODE>
__device__ __forceinline__ void load_kv(const Attn16Args& a, const WinPos& p, const int* s_qtok, int h, uint32_t sK,
                                        uint32_t sV, int nthreads) {
  constexpr int NS = A16<MODE>::NS;
  const int hw = a.heads * 32;
  const int ch = threadIdx.x & 3;
  const __nv_bfloat16* base = a.qkv + hw + h * 32 + ch * 8;
  for (int slot = threadIdx.x >> 2; slot < NS; slot += nthreads >> 2) {
    long long tok;
    if (MODE == MODE_SELF) {
      tok = s_qtok[slot];
    } else {
      bool sv;
      tok = k_token<MODE>(a, p, slot, sv);
    }
    const bool ok = tok >= 0;
    const __nv_bfloat16* src = base + (ok ? tok : 0) * a.ld_qkv;
    cp_async16_zfill(sK + t32_off(slot, ch), src, ok);
    cp_async16_zfill(sV + t32_off(slot, ch), src + hw, ok);
  }
}

// ============================================================================ forward
// grid (gx, heads); 256 threads = 8 warps x 16 query rows = half a window per work item; 2 CTAs / SM.
constexpr int A16_FWD_THREADS = 256;
template <int MODE>
constexpr int a16_fwd_smem() { return 128 * 64 + 2 * A16<MODE>::NS * 64; }

template <int MODE>
__global__ void __launch_bounds__(A16_FWD_THREADS, 2) win_attn16_fwd_kernel(const Attn16Args a) {
  using G = A16<MODE>;
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  __shared__ float s_bias[G::TBL];
  __shared__ int s_qtok[256];
  const uint32_t sQ = smem_u32(smem_dyn), sK = sQ + 128 * 64, sV = sK + G::NS * 64;
  const int h = blockIdx.y;
  const int nwin = a.B * (a.H >> 4) * (a.W >> 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  load_bias_table<MODE>(s_bias, a.bias_table, a.heads, h, A16_FWD_THREADS);
  const bool ones_here = a.ones_col >= h * 32 && a.ones_col < h * 32 + 32;
  const int ones_c = a.ones_col - h * 32;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int item = blockIdx.x; item < nwin * 2; item += gridDim.x) {
    const int w = item >> 1, half = item & 1;
    const WinPos p = win_pos(a, w);
    __syncthreads();  // previous item fully consumed (tiles and token tables)
    fill_token_table(a, p, s_qtok, A16_FWD_THREADS);
    __syncthreads();
    {
      const int ch = threadIdx.x & 3;
      const __nv_bfloat16* base = a.qkv + h * 32 + ch * 8;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = (threadIdx.x >> 2) + 64 * k;  // 128 query rows of this half window
        cp_async16(sQ + t32_off(i, ch), base + (long long)s_qtok[half * 128 + i] * a.ld_qkv);
      }
    }
    load_kv<MODE>(a, p, s_qtok, h, sK, sV, A16_FWD_THREADS);
    cp_async_commit();
    const bool shifted = (MODE == MODE_SELF) && a.shift > 0;
    const bool last_y = shifted && p.wy == (a.H >> 4) - 1, last_x = shifted && p.wx == (a.W >> 4) - 1;
    cp_async_wait<0>();
    __syncthreads();
    const int r0 = warp * 16, qy = half * 8 + warp;
    uint32_t aq[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldsm_x4(sQ + t32_off(row, ks * 2 + (lane >> 4)), aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3]);
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll 1
    for (int kt = 0; kt < G::NKT; ++kt) {
      float s[8][4];
      qk_tile<MODE>(aq, sK + kt * 64 * 64, s_bias + bias_base<MODE>(qy, kt, g, t), mask_ctx(last_y, last_x, qy, kt), lane, s);
      float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        t0 = fmaxf(t0, fmaxf(s[nt][0], s[nt][1]));
        t1 = fmaxf(t1, fmaxf(s[nt][2], s[nt][3]));
      }
      t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
      t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
      t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
      t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
      const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);  // finite: every tile has at least one valid key per row
      const float c0 = fast_ex2((m0 - n0) * kLog2e), c1 = fast_ex2((m1 - n1) * kLog2e);
      m0 = n0; m1 = n1;
      float r0s = 0.f, r1s = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = fast_ex2((s[nt][0] - n0) * kLog2e);
        s[nt][1] = fast_ex2((s[nt][1] - n0) * kLog2e);
        s[nt][2] = fast_ex2((s[nt][2] - n1) * kLog2e);
        s[nt][3] = fast_ex2((s[nt][3] - n1) * kLog2e);
        r0s += s[nt][0] + s[nt][1];
        r1s += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + r0s;
      l1 = l1 * c1 + r1s;
#pragma unroll
      for (int n = 0; n < 4; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
      frag_times_tile_acc(s, sV + kt * 64 * 64, lane, o);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int n = 0; n < 4; ++n) { o[n][0] *= i0; o[n][1] *= i0; o[n][2] *= i1; o[n][3] *= i1; }
    if (t == 0 && a.lse != nullptr) {
      a.lse[(long long)h * a.T + s_qtok[qy * 16 + g]] = m0 + logf(l0);
      a.lse[(long long)h * a.T + s_qtok[qy * 16 + g + 8]] = m1 + logf(l1);
    }
    // this warp's Q rows are dead (fragments live in registers): reuse them as the output staging rows
    __syncwarp();
    store_frag_t32(sQ, r0, lane, o);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = lane + 32 * k, i = c >> 2, ch = c & 3;  // 16 rows x 4 chunks
      const long long tok = s_qtok[qy * 16 + i];
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
// grid (gx, heads); 256 threads = 8 warps, warp w owns query row-groups 2w, 2w+1 of one whole window per work item
// (two m-tiles per warp: halves the shared-memory fragment traffic, which is this kernel's limiter).
constexpr int A16_BWD_THREADS = 256;
template <int MODE>
struct A16BwdSmem {
  static constexpr int kQ = 0;                                  // [256][32]
  static constexpr int kDO = kQ + 256 * 64;                     // [256][32]
  static constexpr int kK = kDO + 256 * 64;                     // [NS][32]
  static constexpr int kV = kK + A16<MODE>::NS * 64;            // [NS][32]
  static constexpr int kP = kV + A16<MODE>::NS * 64;            // [256][64]  (later: dQ staging)
  static constexpr int kDS = kP + 256 * 128;                    // [256][64]
  static constexpr int kOut = kDS + 256 * 128;                  // dK, dV staging 2 x [64][32]
  static constexpr int kBytes = kOut + 2 * 64 * 64;
};

// dst[16 slots x 32 cols] = A^T B: A = t64 tile [256 q][64 slots] (this warp: slots k0..k0+15), B = t32 tile [256 q][32],
// contraction over the 256 queries.  One transposed A fragment feeds four MMAs (all four 8-column n-tiles); the
// fragments of step kk+1 are fetched before the MMAs of step kk are issued.
__device__ __forceinline__ void tileT_times_tile256(uint32_t a_tile, uint32_t b_tile, int k0, int lane, float (&o)[4][4]) {
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  const int ia = lane >> 3;
  const uint32_t a_base = a_tile + t64_off((lane & 7) + ((ia >> 1) & 1) * 8, (k0 + (ia & 1) * 8) >> 3);
  const uint32_t b_base = b_tile + t32_off((lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);
  // rows advance by 16 per step: +16*128 B in the t64 tile, +16*64 B in the t32 tile; the swizzle terms depend on
  // (row & 7) and ((row >> 1) & 3) only, which a multiple of 16 rows leaves unchanged; chunk pair 2,3 = address ^ 32
  uint32_t af[2][4], b0[2][4], b1[2][4];
  auto fetch = [&](int kk, int buf) {
    ldsm_x4_t(a_base + kk * 16 * 128, af[buf][0], af[buf][1], af[buf][2], af[buf][3]);
    ldsm_x4_t(b_base + kk * 16 * 64, b0[buf][0], b0[buf][1], b0[buf][2], b0[buf][3]);
    ldsm_x4_t((b_base + kk * 16 * 64) ^ 32u, b1[buf][0], b1[buf][1], b1[buf][2], b1[buf][3]);
  };
  fetch(0, 0);
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    const int cur = kk & 1;
    if (kk + 1 < 16) fetch(kk + 1, cur ^ 1);
    mma_bf16(o[0], af[cur], b0[cur][0], b0[cur][1]);
    mma_bf16(o[1], af[cur], b0[cur][2], b0[cur][3]);
    mma_bf16(o[2], af[cur], b1[cur][0], b1[cur][1]);
    mma_bf16(o[3], af[cur], b1[cur][2], b1[cur][3]);
  }
}

// Bias-table gradient as a tensor-core diagonal sum.  For every (query row qy, key row / band) pair the 16x16 block
// X[qx][kx] of dS contributes  D[b] = sum_{qx,kx} R[b; qx,kx] X[qx][kx]  with R = [b == 15 + qx - kx] (MODE_SELF) or
// [b == 15 + kx - qx] (MODE_OCA).  As an MMA: M = b (32 diagonals, two m-tiles), K = (qx, kx) (16 k-steps of 16),
// N = 8 pairs.  Per key tile there are 64 pairs = 8 n-tiles; warp j takes n-tile j (its own two query rows) and both m-tiles.
// Pair n of n-tile j: query row qy = 2j + (n>>2), in-tile key row/band index r = n&3 (its 16 slots = columns r*16..).
// The A operand (the 0/1 Toeplitz selector R) is never materialised: each fragment word is zero or one-hot.
template <int MODE>
__device__ __forceinline__ void diag_mma(uint32_t ds_tile, int j, int lane, float (&acc)[2][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
  const int n = lane & 7;                       // ldmatrix row provider (B operand): pair index
  const int prow = (2 * j + (n >> 2)) * 16;     // first dS row of that pair's query row-group
  const int pchunk = (n & 3) * 2 + ((lane >> 3) & 1);
  // A operand built in registers (shared-memory bandwidth is this kernel's limiter, ALU is not): fragment word
  // (row b, k pair 2t, 2t+1) is one-hot iff the selected k = qx + c equals 2t or 2t+1.
  const int g = lane >> 2, t = lane & 3;
  // p = (selected k) - 2t for row b = g of m-tile 0 at step qx = 0; rows b + 8 / + 16 / + 24 shift it by -/+ 8, 16, 24
  const int p0 = ((MODE == MODE_SELF) ? (15 - g) : (g - 15)) - 2 * t;
  constexpr int RS = (MODE == MODE_SELF) ? -8 : 8;   // change of the selected k per +8 rows
  auto onehot2 = [](int p) -> uint32_t { return (uint32_t(p) < 2u) ? (0x3F80u << (16 * p)) : 0u; };
  // prow is a multiple of 16, so the swizzle term of row prow + qx is (qx & 7): a compile-time constant per step
  const uint32_t ds_row0 = ds_tile + uint32_t(prow) * 128u;
  auto fetch = [&](int qx, uint32_t (&b)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(b[0]), "=r"(b[1]) : "r"(ds_row0 + uint32_t(qx) * 128u + (uint32_t(pchunk ^ (qx & 7)) << 4)));
  };
  uint32_t bq[2][2];
  fetch(0, bq[0]);
#pragma unroll
  for (int qx = 0; qx < 16; ++qx) {
    const int cur = qx & 1, nxt = cur ^ 1;
    if (qx + 1 < 16) fetch(qx + 1, bq[nxt]);
    const int p = p0 + qx;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int pm = p + 2 * mt * RS;   // rows b = 16*mt + g
      uint32_t af[4];
      af[0] = onehot2(pm);              // (row b,     k 2t..2t+1)
      af[1] = onehot2(pm + RS);         // (row b + 8, k 2t..2t+1)
      af[2] = onehot2(pm - 8);          // (row b,     k 2t+8..2t+9)
      af[3] = onehot2(pm + RS - 8);     // (row b + 8, k 2t+8..2t+9)
      mma_bf16(acc[mt], af, bq[cur][0], bq[cur][1]);
    }
  }
}

// Two-m-tile variants (backward): a warp owns 32 query rows (two row-groups), so every K / V fragment fetched from
// shared memory feeds four MMAs instead of two.
template <int MODE>
__device__ __forceinline__ void qk_tile2(const uint32_t (&aq)[2][2][4], uint32_t k_tile, const float* bp0, const float* bp1,
                                         const MaskCtx& mk, int lane, float (&s)[2][8][4]) {
  const uint32_t kb_base = k_tile + t32_off(lane & 7, lane >> 3);   // + nt * 8 rows = nt * 512 B (swizzle unchanged)
  uint32_t kb[2][4];
  ldsm_x4(kb_base, kb[0][0], kb[0][1], kb[0][2], kb[0][3]);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int cur = nt & 1, nxt = cur ^ 1;
    if (nt + 1 < 8 && nt_valid<MODE>(nt + 1))
      ldsm_x4(kb_base + (nt + 1) * 512, kb[nxt][0], kb[nxt][1], kb[nxt][2], kb[nxt][3]);
    if (!nt_valid<MODE>(nt)) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = -INFINITY;
      continue;
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float* bp = mt ? bp1 : bp0;
      s[mt][nt][0] = bp[bias_const<MODE>(nt, 0, 0)];
      s[mt][nt][1] = bp[bias_const<MODE>(nt, 1, 0)];
      s[mt][nt][2] = bp[bias_const<MODE>(nt, 0, 1)];
      s[mt][nt][3] = bp[bias_const<MODE>(nt, 1, 1)];
      mma_bf16(s[mt][nt], aq[mt][0], kb[cur][0], kb[cur][1]);
      mma_bf16(s[mt][nt], aq[mt][1], kb[cur][2], kb[cur][3]);
    }
  }
  if (MODE == MODE_SELF && mk.any) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float m01 = (nt & 1) ? mk.diff : mk.same, m23 = (nt & 1) ? mk.same : mk.diff;
        s[mt][nt][0] += m01; s[mt][nt][1] += m01; s[mt][nt][2] += m23; s[mt][nt][3] += m23;
      }
  }
}

// o[mt][16 x 32] += A_mt[16 x 64] (bf16 fragments built from fp32 s[mt]) * Bt[64 x 32], both m-tiles per B fragment
__device__ __forceinline__ void frag_times_tile_acc2(const float (&s)[2][8][4], uint32_t bt_tile, int lane,
                                                     float (&o)[2][4][4]) {
  const uint32_t b_base = bt_tile + t32_off((lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);
  auto frag_addr = [&](int step) { return (b_base + (step >> 1) * 16 * 64) ^ (uint32_t(step & 1) << 5); };
  uint32_t bf[2][4];
  ldsm_x4_t(frag_addr(0), bf[0][0], bf[0][1], bf[0][2], bf[0][3]);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      a[mt][0] = pack_bf16(s[mt][2 * kk][0], s[mt][2 * kk][1]);
      a[mt][1] = pack_bf16(s[mt][2 * kk][2], s[mt][2 * kk][3]);
      a[mt][2] = pack_bf16(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
      a[mt][3] = pack_bf16(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
    }
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const int step = kk * 2 + np, cur = step & 1, nxt = cur ^ 1;
      if (step + 1 < 8) ldsm_x4_t(frag_addr(step + 1), bf[nxt][0], bf[nxt][1], bf[nxt][2], bf[nxt][3]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_bf16(o[mt][2 * np], a[mt], bf[cur][0], bf[cur][1]);
        mma_bf16(o[mt][2 * np + 1], a[mt], bf[cur][2], bf[cur][3]);
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(A16_BWD_THREADS, 1) win_attn16_bwd_kernel(const Attn16Args a) {
  using G = A16<MODE>;
  using L = A16BwdSmem<MODE>;
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  __shared__ float s_bias[G::TBL];
  __shared__ float s_lse[256];
  __shared__ float s_delta[256];
  __shared__ int s_qtok[256];
  const uint32_t sm0 = smem_u32(smem_dyn);
  const uint32_t sQ = sm0 + L::kQ, sDO = sm0 + L::kDO, sK = sm0 + L::kK, sV = sm0 + L::kV, sP = sm0 + L::kP,
                 sDS = sm0 + L::kDS, sOut = sm0 + L::kOut;
  const int h = blockIdx.y;
  const int nwin = a.B * (a.H >> 4) * (a.W >> 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 8 warps; warp w owns query row-groups 2w, 2w+1
  const int g = lane >> 2, t = lane & 3;
  const int hw = a.heads * 32;
  constexpr float kLog2e = 1.4426950408889634f;
  load_bias_table<MODE>(s_bias, a.bias_table, a.heads, h, A16_BWD_THREADS);
  // this CTA's slice of the diagonal-sum scratch: [NKT][16 = m-tile*8 + n-tile][32 lanes] float4, owned per thread
  float4* scratch = reinterpret_cast<float4*>(a.dbias_scratch) +
                    ((size_t)blockIdx.x * a.heads + h) * (G::NKT * 16 * 32);
  for (int kt = 0; kt < G::NKT; ++kt) {
    scratch[(kt * 16 + warp) * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    scratch[(kt * 16 + 8 + warp) * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int w = blockIdx.x; w < nwin; w += gridDim.x) {
    const WinPos p = win_pos(a, w);
    __syncthreads();
    fill_token_table(a, p, s_qtok, A16_BWD_THREADS);
    __syncthreads();
    {
      const int ch = threadIdx.x & 3;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = (threadIdx.x >> 2) + 64 * k;
        const long long tok = s_qtok[i];
        cp_async16(sQ + t32_off(i, ch), a.qkv + tok * a.ld_qkv + h * 32 + ch * 8);
        cp_async16(sDO + t32_off(i, ch), a.dout + tok * a.ld_o + h * 32 + ch * 8);
      }
    }
    load_kv<MODE>(a, p, s_qtok, h, sK, sV, A16_BWD_THREADS);
    cp_async_commit();
    const bool shifted = (MODE == MODE_SELF) && a.shift > 0;
    const bool last_y = shifted && p.wy == (a.H >> 4) - 1, last_x = shifted && p.wx == (a.W >> 4) - 1;
    {  // delta_i = sum_d dO[i,d] * O[i,d]  (one thread per query row), lse
      const int i = threadIdx.x;
      const long long tok = s_qtok[i];
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
      s_delta[i] = acc;
      s_lse[i] = a.lse[(long long)h * a.T + tok] * kLog2e;
    }
    cp_async_wait<0>();
    __syncthreads();
    const int r0 = warp * 32;
    uint32_t aq[2][2][4], ad[2][2][4];
    float lse_[2][2], dl_[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int row = r0 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(sQ + t32_off(row, ks * 2 + (lane >> 4)), aq[mt][ks][0], aq[mt][ks][1], aq[mt][ks][2], aq[mt][ks][3]);
        ldsm_x4(sDO + t32_off(row, ks * 2 + (lane >> 4)), ad[mt][ks][0], ad[mt][ks][1], ad[mt][ks][2], ad[mt][ks][3]);
      }
      lse_[mt][0] = s_lse[r0 + mt * 16 + g]; lse_[mt][1] = s_lse[r0 + mt * 16 + g + 8];
      dl_[mt][0] = s_delta[r0 + mt * 16 + g]; dl_[mt][1] = s_delta[r0 + mt * 16 + g + 8];
    }
    float dq[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < 4; ++n) dq[mt][n][0] = dq[mt][n][1] = dq[mt][n][2] = dq[mt][n][3] = 0.f;

#pragma unroll 1
    for (int kt = 0; kt < G::NKT; ++kt) {
      const uint32_t kT = sK + kt * 64 * 64, vT = sV + kt * 64 * 64;
      // ---- phase A: this warp's 32 query rows x 64 key slots
      float s[2][8][4];
      // (qy >= 8) is the same for both row-groups of a warp (2w, 2w+1), so one mask context serves both
      qk_tile2<MODE>(aq, kT, s_bias + bias_base<MODE>(2 * warp, kt, g, t), s_bias + bias_base<MODE>(2 * warp + 1, kt, g, t),
                     mask_ctx(last_y, last_x, 2 * warp, kt), lane, s);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {  // P = exp(S - lse), normalised
          s[mt][nt][0] = fast_ex2(fmaf(s[mt][nt][0], kLog2e, -lse_[mt][0]));
          s[mt][nt][1] = fast_ex2(fmaf(s[mt][nt][1], kLog2e, -lse_[mt][0]));
          s[mt][nt][2] = fast_ex2(fmaf(s[mt][nt][2], kLog2e, -lse_[mt][1]));
          s[mt][nt][3] = fast_ex2(fmaf(s[mt][nt][3], kLog2e, -lse_[mt][1]));
        }
      uint32_t pp[2][2][2];
      const uint32_t vb_base = vT + t32_off(lane & 7, lane >> 3);
      uint32_t vb[2][4];
      ldsm_x4(vb_base, vb[0][0], vb[0][1], vb[0][2], vb[0][3]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {  // dP = dO V^T, dS = P * (dP - delta); P, dS -> smem (bf16)
        if (nt + 1 < 8 && nt_valid<MODE>(nt + 1))
          ldsm_x4(vb_base + (nt + 1) * 512, vb[(nt + 1) & 1][0], vb[(nt + 1) & 1][1], vb[(nt + 1) & 1][2], vb[(nt + 1) & 1][3]);
        const int m = lane >> 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t off = t64_off(r0 + mt * 16 + (m & 1) * 8 + (lane & 7), nt - 1 + (m >> 1));  // used when nt is odd
          if (!nt_valid<MODE>(nt)) {    // padding keys (always an odd n-tile): P = dS = 0, stored with its even partner
            s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
            stsm_x4(sP + off, pp[mt][0][0], pp[mt][0][1], 0u, 0u);
            stsm_x4(sDS + off, pack_bf16(s[mt][nt - 1][0], s[mt][nt - 1][1]), pack_bf16(s[mt][nt - 1][2], s[mt][nt - 1][3]), 0u, 0u);
            continue;
          }
          float dp[4] = {0.f, 0.f, 0.f, 0.f};
          mma_bf16(dp, ad[mt][0], vb[nt & 1][0], vb[nt & 1][1]);
          mma_bf16(dp, ad[mt][1], vb[nt & 1][2], vb[nt & 1][3]);
          pp[mt][nt & 1][0] = pack_bf16(s[mt][nt][0], s[mt][nt][1]);
          pp[mt][nt & 1][1] = pack_bf16(s[mt][nt][2], s[mt][nt][3]);
          s[mt][nt][0] *= (dp[0] - dl_[mt][0]);
          s[mt][nt][1] *= (dp[1] - dl_[mt][0]);
          s[mt][nt][2] *= (dp[2] - dl_[mt][1]);
          s[mt][nt][3] *= (dp[3] - dl_[mt][1]);
          if (nt & 1) {  // one stmatrix.x4 per n-tile pair and tile
            stsm_x4(sP + off, pp[mt][0][0], pp[mt][0][1], pp[mt][1][0], pp[mt][1][1]);
            stsm_x4(sDS + off, pack_bf16(s[mt][nt - 1][0], s[mt][nt - 1][1]), pack_bf16(s[mt][nt - 1][2], s[mt][nt - 1][3]),
                    pack_bf16(s[mt][nt][0], s[mt][nt][1]), pack_bf16(s[mt][nt][2], s[mt][nt][3]));
          }
        }
      }
      frag_times_tile_acc2(s, kT, lane, dq);  // dQ += dS K
      __syncthreads();
      // ---- phase B: dK / dV of this key tile (contraction over all 256 queries) + bias-gradient diagonal sums
      {
        const int m = warp >> 2, rg = warp & 3;   // matrix (dK / dV), 16-slot row group
        float o[4][4];
        tileT_times_tile256(m ? sP : sDS, m ? sDO : sQ, rg * 16, lane, o);
        store_frag_t32(sOut + m * (64 * 64), rg * 16, lane, o);
        float acc[2][4];
        diag_mma<MODE>(sDS, warp, lane, acc);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          float4* sp = scratch + (kt * 16 + mt * 8 + warp) * 32 + lane;
          float4 v = *sp;
          v.x += acc[mt][0]; v.y += acc[mt][1]; v.z += acc[mt][2]; v.w += acc[mt][3];
          *sp = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 2; ++k) {  // store dK, dV rows of this tile: 2 matrices x 64 slots x 4 chunks = 512 x 16 B
        const int c = threadIdx.x + 256 * k, m = c >> 8, rem = c & 255, sl = rem >> 2, ch = rem & 3;
        const int slot = kt * 64 + sl;
        const uint4 v = lds128(sOut + m * (64 * 64) + t32_off(sl, ch));
        if (MODE == MODE_SELF) {
          const long long tok = s_qtok[slot];
          *reinterpret_cast<uint4*>(a.dqkv + tok * a.ld_qkv + (1 + m) * hw + h * 32 + ch * 8) = v;
        } else {
          __nv_bfloat16* dst = a.dkv_win + ((((size_t)w * a.heads + h) * 2 + m) * G::NS + slot) * 32 + ch * 8;
          *reinterpret_cast<uint4*>(dst) = v;
        }
      }
      // sOut is rewritten in the next tile's phase B, i.e. after the next tile's first barrier; sP / sDS are rewritten
      // by the next tile's phase A, after the barrier above (all phase-B reads precede it).
    }
    // dQ -> staging (this warp's rows of the P tile region, dead after the last barrier) -> global
    __syncwarp();
    store_frag_t32(sP, r0, lane, dq[0]);
    store_frag_t32(sP, r0 + 16, lane, dq[1]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane + 32 * k, i = c >> 2, ch = c & 3;   // 32 rows x 4 chunks
      const long long tok = s_qtok[r0 + i];
      *reinterpret_cast<uint4*>(a.dqkv + tok * a.ld_qkv + h * 32 + ch * 8) = lds128(sP + t32_off(r0 + i, ch));
    }
  }
}

// d_rpb_table[t*heads + h] = sum over CTAs and (qy, key row) pairs of the diagonal sums that map to table entry t.
// One thread per (table entry, head).  Accumulator addressing mirrors diag_mma's fragment layout.
template <int MODE>
__global__ void attn16_dbias_finish_kernel(const float* __restrict__ scratch, int gx, int heads, float* __restrict__ out) {
  using G = A16<MODE>;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G::TBL * heads) return;
  const int tbl = idx / heads, h = idx % heads;
  float total = 0.f;
  // enumerate (qy, ky, band, b) accumulators that feed table entry `tbl`
  for (int qy = 0; qy < 16; ++qy) {
    for (int band = 0; band < (MODE == MODE_SELF ? 1 : 2); ++band) {
      int ky, b;
      if (MODE == MODE_SELF) {
        const int ry = tbl / 31, rx = tbl % 31;   // ry = qy - ky + 15, rx = qx - kx + 15 = b
        ky = qy + 15 - ry; b = rx;
        if (ky < 0 || ky >= 16) continue;
      } else {
        // idx_ref = ry*39 + rx with ry = ky-qy-7, rx = kx-qx-7, both in [-22,16]; a negative idx_ref was wrapped (+1521)
        const int v = (tbl > 16 * 39 + 16) ? tbl - 1521 : tbl;
        const int num = v + 22;  // floor division by 39
        const int ry = (num >= 0) ? num / 39 : -((-num + 38) / 39);
        const int rx = v - ry * 39;  // in [-22, 16]
        ky = ry + qy + 7;
        if (ky < 0 || ky >= 24) continue;
        b = rx + 22 - band * 16;   // b = 15 + kx_in_band - qx, kx = band*16 + kx_in_band
        if (b < 0 || b >= 32) continue;
      }
      int kt, r;
      if (MODE == MODE_SELF) { kt = ky >> 2; r = ky & 3; }
      else { kt = ky >> 1; r = (ky & 1) * 2 + band; }
      const int j = qy >> 1, n = ((qy & 1) << 2) | r;
      const int mt = b >> 4, rowin = b & 15;
      const int lane = (rowin & 7) * 4 + (n >> 1), reg = (rowin >> 3) * 2 + (n & 1);
      const int warp = mt * 8 + j;
      for (int c = 0; c < gx; ++c)
        total += scratch[((((size_t)c * heads + h) * G::NKT + kt) * 16 + warp) * 128 + lane * 4 + reg];
    }
  }
  out[idx] = total;
}

// MODE_OCA: d_qkv[tok][K|V] = sum over the (up to 4) overlapping windows of their per-window dK/dV rows.
__global__ void oca_kv_gather_kernel(const __nv_bfloat16* __restrict__ dkv_win, __nv_bfloat16* __restrict__ dqkv,
                                     int ld_qkv, int B, int H, int W, int heads) {
  const long long total = (long long)B * H * W * heads * 2 * 4;
  const int nwy = H >> 4, nwx = W >> 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = int(i & 3);
    long long r = i >> 2;
    const int h = int(r % heads); r /= heads;
    const int m = int(r & 1); r >>= 1;
    const long long tok = r;
    const int x = int(tok % W), y = int((tok / W) % H), b = int(tok / ((long long)W * H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int wy_hi = min((y + 4) >> 4, nwy - 1), wx_hi = min((x + 4) >> 4, nwx - 1);
    for (int wy = wy_hi; wy >= 0 && wy * 16 + 20 > y; --wy)
      for (int wx = wx_hi; wx >= 0 && wx * 16 + 20 > x; --wx) {
        const int ky = y - wy * 16 + 4, kx = x - wx * 16 + 4;
        const int slot = (ky >> 1) * 64 + (ky & 1) * 32 + (kx >> 4) * 16 + (kx & 15);
        const size_t win = ((size_t)b * nwy + wy) * nwx + wx;
        const uint4 v = *reinterpret_cast<const uint4*>(dkv_win + (((win * heads + h) * 2 + m) * 768 + slot) * 32 + ch * 8);
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc[2 * e] += bf16_lo(wv[e]); acc[2 * e + 1] += bf16_hi(wv[e]); }
      }
    *reinterpret_cast<uint4*>(dqkv + tok * ld_qkv + (1 + m) * heads * 32 + h * 32 + ch * 8) =
        make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}

}  // namespace srk
