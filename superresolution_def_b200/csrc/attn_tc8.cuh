// attn_tc8.cuh — (shifted-)window attention core for 8x8 windows on the Blackwell tensor path:
// window tokens TMA-staged into shared memory, QK^T and PV on tcgen05 with the logits / output in TMEM, bias + mask +
// softmax on tcgen05.ld fragments (one thread owns one query row, so the row max / sum need no shuffles at all).
//
// Replaces (reference): torch.roll (models/architecture_swin.py:130-133,143-146), window_partition / window_reverse
// (:27-37) and, inside WindowAttention.forward (:75-93), q@k^T, the relative-position-bias gather + add, softmax, attn@v and
// the head-merge transpose; with `mask` also HAT's (S)W-MSA at window 8 (hat_arch.py:165-196, mask :921-940).
//
// Work unit = (pair of windows, pair of heads): the M = 128 rows of one tcgen05.mma are the 2 x 64 query tokens of two
// windows, the 64-channel TMA boxes carry two 32-wide head slots.
//   TMA    per window and operand four quadrant boxes (64 ch, 4 x, 4 y) of the 4-D view [c, x, y, b] of qkv: a cyclic shift
//          of 4 moves window borders onto quadrant borders, so torch.roll + window_partition are box coordinates
//          (wrapped per quadrant); tokens sit in shared memory in quadrant-major order r = quad*16 + yl*4 + xl.
//   S      [128 x 128] = Q_h K_h^T, K = 32: the two heads are K sub-ranges (+64 B) of the same 128B-swizzled rows; only the
//          two diagonal 64 x 64 blocks (query and key of the same window) are read back.
//   P      bf16, written into a [128 x 128] K-major tile whose off-diagonal blocks stay zero, so ONE M = 128 MMA with
//          K = 128 keys computes both windows' P V;  V is fed as an MN-major operand straight from its token-major box.
//   O      overwrites the S columns in TMEM (S is dead once P is in shared memory); 32 fp32 per thread -> bf16 -> global.
// S is double-buffered in TMEM (2 x 256 columns), the loads run 3 stages ahead.
// Algorithmic HBM bytes per token and head: read q, k, v (3 x 64 B), write out (64 B) — same as the mma.sync kernel.
#pragma once
#include "attn_ws8.cuh"

namespace srk {

constexpr int TC8_THREADS = 64 + 256;      // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (2 heads x 4 lane quarters)
constexpr int TC8_NST = 3;
constexpr int TC8_TILE = 128 * 128;        // [128 tokens x 64 ch] bf16
constexpr int TC8_STAGE = 3 * TC8_TILE;    // q, k, v
constexpr int TC8_OFF_P = TC8_NST * TC8_STAGE;
constexpr int TC8_OFF_BAR = TC8_OFF_P + 2 * 2 * TC8_TILE;
constexpr int TC8_SMEM = TC8_OFF_BAR + 512 + 1024;
constexpr int TC8_MAX_HEADS = 8;

// window-local coordinates of quadrant-major token r (0..63)
__host__ __device__ constexpr int tc8_y(int r) { return ((r >> 5) & 1) * 4 + ((r >> 2) & 3); }
__host__ __device__ constexpr int tc8_x(int r) { return ((r >> 4) & 1) * 4 + (r & 3); }
// bias-table offset of key j relative to the query's base (y_i * 15 + x_i): idx = (y_i - y_j + 7) * 15 + (x_i - x_j + 7)
__host__ __device__ constexpr int tc8_boff(int j) { return (7 - tc8_y(j)) * 15 + (7 - tc8_x(j)); }

// byte offset of 16-byte chunk `ch` of row `r` inside a 128B-swizzled [rows x 128 B] tile
__device__ __forceinline__ uint32_t tc8_swz(int r, int ch) { return uint32_t(r) * 128u + (uint32_t(ch ^ (r & 7)) << 4); }

struct Tc8Win { int b, y0, x0, last_y, last_x; };
__device__ __forceinline__ Tc8Win tc8_window(const AttnArgs& a, int w) {
  const int nwx = a.W >> 3, nwy = a.H >> 3;
  Tc8Win p;
  p.b = w / (nwx * nwy);
  const int r = w - p.b * nwx * nwy;
  const int wy = r / nwx, wx = r - wy * nwx;
  p.y0 = wy * 8 + a.shift;
  p.x0 = wx * 8 + a.shift;
  p.last_y = (wy == nwy - 1);
  p.last_x = (wx == nwx - 1);
  return p;
}

__global__ void __launch_bounds__(TC8_THREADS, 1)
win_attn_tc8_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC8_OFF_BAR;
  auto ld_full = [&](int s) { return bar_base + 8u * s; };
  auto ld_empty = [&](int s) { return bar_base + 8u * (TC8_NST + s); };
  auto s_full = [&](int buf, int hh) { return bar_base + 8u * (2 * TC8_NST + buf * 2 + hh); };
  auto s_empty = [&](int buf, int hh) { return bar_base + 8u * (2 * TC8_NST + 4 + buf * 2 + hh); };
  auto o_full = [&](int buf, int hh) { return bar_base + 8u * (2 * TC8_NST + 8 + buf * 2 + hh); };
  auto p_full = [&](int hh) { return bar_base + 8u * (2 * TC8_NST + 12 + hh); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC8_NST + 14);
  __shared__ float s_table[TC8_MAX_HEADS * 225];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwin = a.B * (a.H >> 3) * (a.W >> 3);
  const int npairs = nwin >> 1;
  const int nhp = a.heads >> 1;
  const int my_pairs = (npairs - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int U = my_pairs * nhp;            // units of this CTA
  const int AW = a.heads * 32;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < TC8_NST; ++s) { mbar_init(ld_full(s), 1); mbar_init(ld_empty(s), 1); }
    for (int b = 0; b < 2; ++b)
      for (int hh = 0; hh < 2; ++hh) { mbar_init(s_full(b, hh), 1); mbar_init(s_empty(b, hh), 4); mbar_init(o_full(b, hh), 1); }
    mbar_init(p_full(0), 4); mbar_init(p_full(1), 4);
    fence_mbar_init();
  }
  pdl_launch_dependents();
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  pdl_wait();
  for (int i = threadIdx.x; i < 225 * a.heads; i += TC8_THREADS) {
    const int h = i / 225, t = i - h * 225;
    s_table[i] = a.bias_table[t * a.heads + h];
  }
  // the off-diagonal blocks of the P tiles are never written again: zero everything once
  for (int i = threadIdx.x; i < 4 * TC8_TILE / 16; i += TC8_THREADS) sts128(smem_base + TC8_OFF_P + i * 16, make_uint4(0, 0, 0, 0));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < my_pairs; ++i) {
        const int wp = int(blockIdx.x) + i * int(gridDim.x);
        const Tc8Win w0 = tc8_window(a, 2 * wp), w1 = tc8_window(a, 2 * wp + 1);
        for (int hp = 0; hp < nhp; ++hp) {
          mbar_wait(ld_empty(stage), phase ^ 1u);
          mbar_arrive_expect_tx(ld_full(stage), TC8_STAGE);
          const uint32_t st = smem_base + stage * TC8_STAGE;
#pragma unroll
          for (int wi = 0; wi < 2; ++wi) {
            const Tc8Win& w = wi ? w1 : w0;
#pragma unroll
            for (int quad = 0; quad < 4; ++quad) {
              int y = w.y0 + (quad >> 1) * 4, x = w.x0 + (quad & 1) * 4;
              if (y >= a.H) y -= a.H;
              if (x >= a.W) x -= a.W;
              const uint32_t dst = st + uint32_t(wi * 64 + quad * 16) * 128u;
#pragma unroll
              for (int t = 0; t < 3; ++t) tma_load_4d(dst + t * TC8_TILE, &tmQKV, ld_full(stage), t * AW + hp * 64, x, y, w.b);
            }
          }
          if (++stage == TC8_NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // B = V, MN-major
      auto issue_s = [&](int v) {
        const int st = v % TC8_NST;
        mbar_wait(ld_full(st), uint32_t(v / TC8_NST) & 1u);
        const uint32_t qt = smem_base + st * TC8_STAGE, kt = qt + TC8_TILE;
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(s_empty(v & 1, hh), ((uint32_t(v) >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d = tmem_base + uint32_t((v & 1) * 256 + hh * 128);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(d, make_smem_desc(qt + hh * 64 + k * 32, 16, 1024), make_smem_desc(kt + hh * 64 + k * 32, 16, 1024),
                      idesc_s, k);
          umma_commit(s_full(v & 1, hh));
        }
      };
      if (U > 0) issue_s(0);
      for (int u = 0; u < U; ++u) {
        if (u + 1 < U) issue_s(u + 1);
        const int st = u % TC8_NST;
        const uint32_t vt = smem_base + st * TC8_STAGE + 2 * TC8_TILE;
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(p_full(hh), uint32_t(u) & 1u);
          tc_fence_after();
          const uint32_t pt = smem_base + TC8_OFF_P + hh * 2 * TC8_TILE;
          const uint32_t d = tmem_base + uint32_t((u & 1) * 256 + hh * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16(d, make_smem_desc(pt + (j >> 2) * TC8_TILE + (j & 3) * 32, 16, 1024),
                      make_smem_desc_mn(vt + j * 2048, TC8_TILE), idesc_o, j);
          umma_commit(o_full(u & 1, hh));
        }
        umma_commit(ld_empty(st));   // q, k (S of this unit completed long ago) and v are dead
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / output warps
    const int hh = (warp - 2) >> 2;          // head within the pair
    const int q = warp & 3;                  // TMEM lane quarter
    const int row = q * 32 + lane;           // 0..127: window (row >> 6), token r = row & 63
    const int wi = row >> 6, r = row & 63;
    const int yi = tc8_y(r), xi = tc8_x(r);
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    const uint32_t p_row = smem_base + TC8_OFF_P + hh * 2 * TC8_TILE + wi * TC8_TILE;
    const bool masked = a.mask != 0 && a.shift > 0;
    constexpr float kLog2e = 1.4426950408889634f;
    int u = 0;
#ifdef SRK_TC8_PROFILE
    long long t_swait = 0, t_ld = 0, t_soft = 0, t_pwrite = 0, t_owait = 0, t_epi = 0, t0 = clock64(), tk;
#define TC8_TICK(acc) do { tk = clock64(); acc += tk - t0; t0 = tk; } while (0)
#else
#define TC8_TICK(acc) do { } while (0)
#endif
    for (int i = 0; i < my_pairs; ++i) {
      const int wp = int(blockIdx.x) + i * int(gridDim.x);
      const Tc8Win w = tc8_window(a, 2 * wp + wi);
      int y = w.y0 + yi, x = w.x0 + xi;
      if (y >= a.H) y -= a.H;
      if (x >= a.W) x -= a.W;
      const long long tok = (long long)(w.b * a.H + y) * a.W + x;
      float mq[4] = {0.f, 0.f, 0.f, 0.f};   // additive mask per key quadrant (HAT, windows on the last row / column)
      if (masked && (w.last_y || w.last_x)) {
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
          const bool yd = w.last_y && ((yi >> 2) != (kq >> 1));
          const bool xd = w.last_x && ((xi >> 2) != (kq & 1));
          mq[kq] = (yd || xd) ? -100.0f : 0.0f;
        }
      }
      for (int hp = 0; hp < nhp; ++hp, ++u) {
        const int head = hp * 2 + hh;
        const int buf = u & 1;
        const float* tb = s_table + head * 225 + yi * 15 + xi;
        TC8_TICK(t_epi);
        mbar_wait(s_full(buf, hh), (uint32_t(u) >> 1) & 1u);
        tc_fence_after();
        TC8_TICK(t_swait);
        uint32_t sv[64];
        tmem_ld_x64(tmem_base + lane_sel + uint32_t(buf * 256 + hh * 128 + wi * 64), sv);
        tmem_ld_wait();
        TC8_TICK(t_ld);
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float s = __uint_as_float(sv[j]) + tb[tc8_boff(j)] + mq[j >> 4];
          sv[j] = __float_as_uint(s);
          mx = fmaxf(mx, s);
        }
        const float nm = -mx * kLog2e;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float e = fast_ex2(fmaf(__uint_as_float(sv[j]), kLog2e, nm));
          sv[j] = __float_as_uint(e);
          sum += e;
        }
        const float inv = fast_rcp(sum);
        TC8_TICK(t_soft);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = pack_bf16(__uint_as_float(sv[c * 8 + 2 * e]) * inv, __uint_as_float(sv[c * 8 + 2 * e + 1]) * inv);
          sts128(p_row + tc8_swz(row, c), make_uint4(o[0], o[1], o[2], o[3]));
        }
        tc_fence_before();      // S has been read: PV may overwrite its columns
        fence_proxy_async();    // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(hh));
        TC8_TICK(t_pwrite);
        // ---- O = P V (this head's 32 columns of the 64-channel result)
        mbar_wait(o_full(buf, hh), (uint32_t(u) >> 1) & 1u);
        tc_fence_after();
        TC8_TICK(t_owait);
        uint32_t ov[32];
        tmem_ld_x32(tmem_base + lane_sel + uint32_t(buf * 256 + hh * 128 + hh * 32), ov);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(buf, hh));
        uint32_t ob[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) ob[e] = pack_bf16(__uint_as_float(ov[2 * e]), __uint_as_float(ov[2 * e + 1]));
        const int oc = a.ones_col - head * 32;   // bias-folding column of the following projection := 1.0
        if (oc >= 0 && oc < 32) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (e == (oc >> 1)) ob[e] = (oc & 1) ? ((ob[e] & 0x0000FFFFu) | 0x3F800000u) : ((ob[e] & 0xFFFF0000u) | 0x00003F80u);
        }
        uint4* op = reinterpret_cast<uint4*>(a.out + tok * a.ld_o + head * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) op[c] = make_uint4(ob[4 * c], ob[4 * c + 1], ob[4 * c + 2], ob[4 * c + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ============================================================================ backward
// Same unit (pair of windows x pair of heads), five tensor-core products per head, all with M = 128:
//   S  = Q K^T, dP = dO V^T            K-major operands straight from the TMA boxes            -> TMEM [0,128), [128,256)
//   P, dS                              one thread per query row: softmax recomputed, delta = sum_j P dP, dS = P (dP - delta)
//   dV = P^T dO                        A = P tile as an MN-major operand (contraction over the query ROWS), B = dO MN-major
//   dK = dS^T Q, dQ = dS K             dS replaces P in the same shared-memory tile once dV has completed;
//                                      dK: both MN-major, dQ: A = dS K-major, B = K MN-major  -> TMEM [0,64), [64,128), [128,192)
// Block-diagonal zeros of the P / dS tile make the two windows of a pair independent inside one M = 128 instruction.
// The relative-position-bias gradient is accumulated in registers (a thread always owns the same (query, key) pairs of a
// head while the CTA walks its windows with the head pair as the OUTER loop) and folded into per-CTA partials at the end.
constexpr int TC8B_NST = 2;
constexpr int TC8B_STAGE = 4 * TC8_TILE;   // q, k, v, dO
constexpr int TC8B_OFF_P = TC8B_NST * TC8B_STAGE;
constexpr int TC8B_OFF_BAR = TC8B_OFF_P + 2 * 2 * TC8_TILE;
constexpr int TC8B_SMEM = TC8B_OFF_BAR + 512 + 1024;

__global__ void __launch_bounds__(TC8_THREADS, 1)
win_attn_tc8_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC8B_OFF_BAR;
  auto ld_full = [&](int s) { return bar_base + 8u * s; };
  auto ld_empty = [&](int s) { return bar_base + 8u * (TC8B_NST + s); };
  auto sdp_full = [&](int hh) { return bar_base + 8u * (2 * TC8B_NST + hh); };
  auto p_full = [&](int hh) { return bar_base + 8u * (2 * TC8B_NST + 2 + hh); };
  auto dv_done = [&](int hh) { return bar_base + 8u * (2 * TC8B_NST + 4 + hh); };
  auto ds_full = [&](int hh) { return bar_base + 8u * (2 * TC8B_NST + 6 + hh); };
  auto out_full = [&](int hh) { return bar_base + 8u * (2 * TC8B_NST + 8 + hh); };
  auto t_empty = [&](int hh) { return bar_base + 8u * (2 * TC8B_NST + 10 + hh); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC8B_NST + 12);
  __shared__ float s_table[TC8_MAX_HEADS * 225];
  __shared__ float s_dbias[TC8_MAX_HEADS * 225];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwin = a.B * (a.H >> 3) * (a.W >> 3);
  const int npairs = nwin >> 1;
  const int nhp = a.heads >> 1;
  const int my_pairs = (npairs - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int U = my_pairs * nhp;   // unit u = hp * my_pairs + i  (head pair outer, window pair inner)
  const int AW = a.heads * 32;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmDO);
    for (int s = 0; s < TC8B_NST; ++s) { mbar_init(ld_full(s), 1); mbar_init(ld_empty(s), 1); }
    for (int hh = 0; hh < 2; ++hh) {
      mbar_init(sdp_full(hh), 1); mbar_init(p_full(hh), 4); mbar_init(dv_done(hh), 1);
      mbar_init(ds_full(hh), 4); mbar_init(out_full(hh), 1); mbar_init(t_empty(hh), 4);
    }
    fence_mbar_init();
  }
  pdl_launch_dependents();
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  pdl_wait();
  for (int i = threadIdx.x; i < 225 * a.heads; i += TC8_THREADS) {
    const int h = i / 225, t = i - h * 225;
    s_table[i] = a.bias_table[t * a.heads + h];
    s_dbias[i] = 0.f;
  }
  for (int i = threadIdx.x; i < 4 * TC8_TILE / 16; i += TC8_THREADS) sts128(smem_base + TC8B_OFF_P + i * 16, make_uint4(0, 0, 0, 0));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int hp = 0; hp < nhp; ++hp) {
        for (int i = 0; i < my_pairs; ++i) {
          const int wp = int(blockIdx.x) + i * int(gridDim.x);
          mbar_wait(ld_empty(stage), phase ^ 1u);
          mbar_arrive_expect_tx(ld_full(stage), TC8B_STAGE);
          const uint32_t st = smem_base + stage * TC8B_STAGE;
#pragma unroll
          for (int wi = 0; wi < 2; ++wi) {
            const Tc8Win w = tc8_window(a, 2 * wp + wi);
#pragma unroll
            for (int quad = 0; quad < 4; ++quad) {
              int y = w.y0 + (quad >> 1) * 4, x = w.x0 + (quad & 1) * 4;
              if (y >= a.H) y -= a.H;
              if (x >= a.W) x -= a.W;
              const uint32_t dst = st + uint32_t(wi * 64 + quad * 16) * 128u;
#pragma unroll
              for (int t = 0; t < 3; ++t) tma_load_4d(dst + t * TC8_TILE, &tmQKV, ld_full(stage), t * AW + hp * 64, x, y, w.b);
              tma_load_4d(dst + 3 * TC8_TILE, &tmDO, ld_full(stage), hp * 64, x, y, w.b);
            }
          }
          if (++stage == TC8B_NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_tt = make_idesc_bf16(128, 64, 1, 1);   // dV, dK: both operands MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);    // dQ: A = dS K-major, B = K MN-major
      int stage = 0; uint32_t phase = 0;
      for (int u = 0; u < U; ++u) {
        const uint32_t up = uint32_t(u) & 1u;
        mbar_wait(ld_full(stage), phase);
        const uint32_t qt = smem_base + stage * TC8B_STAGE, kt = qt + TC8_TILE, vt = qt + 2 * TC8_TILE, dot = qt + 3 * TC8_TILE;
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(t_empty(hh), up ^ 1u);
          tc_fence_after();
          const uint32_t d = tmem_base + uint32_t(hh * 256);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(d, make_smem_desc(qt + hh * 64 + k * 32, 16, 1024), make_smem_desc(kt + hh * 64 + k * 32, 16, 1024),
                      idesc_s, k);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(d + 128, make_smem_desc(dot + hh * 64 + k * 32, 16, 1024), make_smem_desc(vt + hh * 64 + k * 32, 16, 1024),
                      idesc_s, k);
          umma_commit(sdp_full(hh));
        }
        for (int hh = 0; hh < 2; ++hh) {   // dV = P^T dO
          mbar_wait(p_full(hh), up);
          tc_fence_after();
          const uint32_t pt = smem_base + TC8B_OFF_P + hh * 2 * TC8_TILE;
          const uint32_t d = tmem_base + uint32_t(hh * 256);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16(d, make_smem_desc_mn(pt + j * 2048, TC8_TILE), make_smem_desc_mn(dot + j * 2048, TC8_TILE), idesc_tt, j);
          umma_commit(dv_done(hh));
        }
        for (int hh = 0; hh < 2; ++hh) {   // dK = dS^T Q, dQ = dS K
          mbar_wait(ds_full(hh), up);
          tc_fence_after();
          const uint32_t pt = smem_base + TC8B_OFF_P + hh * 2 * TC8_TILE;
          const uint32_t d = tmem_base + uint32_t(hh * 256);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16(d + 64, make_smem_desc_mn(pt + j * 2048, TC8_TILE), make_smem_desc_mn(qt + j * 2048, TC8_TILE), idesc_tt, j);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16(d + 128, make_smem_desc(pt + (j >> 2) * TC8_TILE + (j & 3) * 32, 16, 1024),
                      make_smem_desc_mn(kt + j * 2048, TC8_TILE), idesc_q, j);
          umma_commit(out_full(hh));
        }
        umma_commit(ld_empty(stage));
        if (++stage == TC8B_NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / gradient warps
    const int hh = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wi = row >> 6, r = row & 63;
    const int yi = tc8_y(r), xi = tc8_x(r);
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    const uint32_t p_row = smem_base + TC8B_OFF_P + hh * 2 * TC8_TILE + wi * TC8_TILE;
    const uint32_t treg = tmem_base + lane_sel + uint32_t(hh * 256);
    const bool masked = a.mask != 0 && a.shift > 0;
    constexpr float kLog2e = 1.4426950408889634f;
    int u = 0;
    for (int hp = 0; hp < nhp; ++hp) {
      const int head = hp * 2 + hh;
      const float* tb = s_table + head * 225 + yi * 15 + xi;
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      for (int i = 0; i < my_pairs; ++i, ++u) {
        const uint32_t up = uint32_t(u) & 1u;
        const int wp = int(blockIdx.x) + i * int(gridDim.x);
        const Tc8Win w = tc8_window(a, 2 * wp + wi);
        int y = w.y0 + yi, x = w.x0 + xi;
        if (y >= a.H) y -= a.H;
        if (x >= a.W) x -= a.W;
        const long long tok = (long long)(w.b * a.H + y) * a.W + x;
        float mq[4] = {0.f, 0.f, 0.f, 0.f};
        if (masked && (w.last_y || w.last_x)) {
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) {
            const bool yd = w.last_y && ((yi >> 2) != (kq >> 1));
            const bool xd = w.last_x && ((xi >> 2) != (kq & 1));
            mq[kq] = (yd || xd) ? -100.0f : 0.0f;
          }
        }
        mbar_wait(sdp_full(hh), up);
        tc_fence_after();
        uint32_t pv[64];
        tmem_ld_x64(treg + uint32_t(wi * 64), pv);
        tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float s = __uint_as_float(pv[j]) + tb[tc8_boff(j)] + mq[j >> 4];
          pv[j] = __float_as_uint(s);
          mx = fmaxf(mx, s);
        }
        const float nm = -mx * kLog2e;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float e = fast_ex2(fmaf(__uint_as_float(pv[j]), kLog2e, nm));
          pv[j] = __float_as_uint(e);
          sum += e;
        }
        const float inv = fast_rcp(sum);
#pragma unroll
        for (int j = 0; j < 64; ++j) pv[j] = __float_as_uint(__uint_as_float(pv[j]) * inv);   // P (fp32)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = pack_bf16(__uint_as_float(pv[c * 8 + 2 * e]), __uint_as_float(pv[c * 8 + 2 * e + 1]));
          sts128(p_row + tc8_swz(row, c), make_uint4(o[0], o[1], o[2], o[3]));
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(hh));
        // delta = sum_j P_ij dP_ij ; dS = P (dP - delta)      (dP is read from TMEM twice, 32 columns at a time)
        float delta = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t dp[32];
          tmem_ld_x32(treg + uint32_t(128 + wi * 64 + half * 32), dp);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) delta = fmaf(__uint_as_float(pv[half * 32 + j]), __uint_as_float(dp[j]), delta);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t dp[32];
          tmem_ld_x32(treg + uint32_t(128 + wi * 64 + half * 32), dp);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float ds = __uint_as_float(pv[half * 32 + j]) * (__uint_as_float(dp[j]) - delta);
            pv[half * 32 + j] = __float_as_uint(ds);
            acc[half * 32 + j] += ds;
          }
        }
        mbar_wait(dv_done(hh), up);     // the tensor core has finished reading P: the tile may take dS
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = pack_bf16(__uint_as_float(pv[c * 8 + 2 * e]), __uint_as_float(pv[c * 8 + 2 * e + 1]));
          sts128(p_row + tc8_swz(row, c), make_uint4(o[0], o[1], o[2], o[3]));
        }
        tc_fence_before();      // dP has been read: dQ may overwrite its columns
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full(hh));
        // ---- dV, dK (rows = keys of this token), dQ (row = query): this head's 32 columns each
        mbar_wait(out_full(hh), up);
        tc_fence_after();
        __nv_bfloat16* gp = a.dqkv + tok * a.ld_qkv + head * 32;
#pragma unroll
        for (int t = 0; t < 3; ++t) {     // t = 0: dV -> v slot, 1: dK -> k slot, 2: dQ -> q slot
          uint32_t ov[32];
          tmem_ld_x32(treg + uint32_t(t * 64 + hh * 32), ov);
          tmem_ld_wait();
          if (t == 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty(hh));
          }
          uint4* op = reinterpret_cast<uint4*>(gp + (2 - t) * AW);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            op[c] = make_uint4(pack_bf16(__uint_as_float(ov[8 * c]), __uint_as_float(ov[8 * c + 1])),
                               pack_bf16(__uint_as_float(ov[8 * c + 2]), __uint_as_float(ov[8 * c + 3])),
                               pack_bf16(__uint_as_float(ov[8 * c + 4]), __uint_as_float(ov[8 * c + 5])),
                               pack_bf16(__uint_as_float(ov[8 * c + 6]), __uint_as_float(ov[8 * c + 7])));
        }
      }
      // fold this thread's (query, key) sums of head `head` into the table entries
      float* db = s_dbias + head * 225 + yi * 15 + xi;
#pragma unroll
      for (int j = 0; j < 64; ++j) atomicAdd(db + tc8_boff(j), acc[j]);
    }
  }

  tc_fence_before();
  __syncthreads();
  for (int i = threadIdx.x; i < 225 * a.heads; i += TC8_THREADS)
    a.dbias_partials[size_t(blockIdx.x) * a.heads * 225 + i] = s_dbias[i];
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace srk
