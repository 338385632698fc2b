// attn_tc16.cuh — forward of HAT's 16x16 window attention (W-MSA / SW-MSA of HAB and the overlapping cross-attention of
// OCAB) on the Blackwell tensor path: window tokens arrive as TMA boxes, Q K^T and P V are tcgen05.mma with the logits and
// the output in TMEM, bias + mask + softmax run on tcgen05.ld fragments (one thread owns one query row: the row max and
// row sum need no shuffles), P goes back to the tensor core through 128B-swizzled shared-memory chunks.
//
// Replaces (reference, models/hat_arch/hat_arch.py): window_partition / window_reverse :97-126, torch.roll :280-302,
// nn.Unfold + rearrange :408-409 and, inside WindowAttention.forward :175-193 / OCAB.forward :419-428, q@k^T, the
// relative-position-bias gather + add, the 0/-100 shift mask (:921-940), softmax, attn@v and the head-merge transpose.
//
// Geometry.  A window is four 8x8 quadrants; the cyclic shift is 0 or 8 = one quadrant, so torch.roll + window_partition
// are the (wrapped) origins of four TMA boxes [64 channels = 2 heads, 8 x, 8 y] of the 4-D view [c, x, y, b] of qkv, and
// tokens sit in shared memory quadrant-major: row = quad * 64 + yl * 8 + xl.  OCAB's 24x24 key window is nine such boxes
// at origin (-4, -4) of the (unshifted) query window; the part of a box outside the image is zero-filled by TMA, which IS
// the reference's zero padding of the projected k / v (nn.Unfold(padding), :377,408).
//
// Work unit = (window, pair of heads); grid (windows strided, head pairs).  Inside a unit FOUR independent LANES,
// lane = (query half hf, head hh of the pair): 128 query rows = the M of one tcgen05.mma, 128 TMEM columns each.  Every lane
// walks the same stream of key PARTS (two 64-key chunks = 128 keys; OCA's ninth chunk is a part of its own):
//   S   [128 x 128] = Q_hf K_part^T, K = 32 (two K = 16 steps; the two heads are +64 B sub-ranges of the same swizzled rows).
//   P   per 64-key chunk: p = 2^(s*log2e + bias' + mask' - m'), m' = (row max of s)*log2e + (max of the bias table) >= the
//                    true row maximum of the part, so p <= 1 without a second pass over the bias; bf16 [128 x 64] -> smem.
//                    (the row max is taken over s*log2e + mask': the mask is constant per 64-key chunk.)
//   O   [128 x 64] += P_chunk V_chunk (V as an MN-major operand straight from its token-major boxes, both heads'
//                    channels; the 32 columns of the lane's head are read back).  O overwrites columns 0..63 of the lane's
//                    S buffer: chunk 0 of S is dead once its P chunk has been handed over.
//   The parts of a row are combined by the online-softmax recurrence in REGISTERS (32 output values per thread) — never by
//   rescaling TMEM; S of a whole window (256 / 576 columns x 4 lanes) would not fit the 512 TMEM columns anyway.
// K / V travel per part through a 3-stage ring (a stage = 2 K + 2 V quadrant boxes, 32 KB), Q per unit.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (an event loop over the four lanes: it issues whatever a lane is
// ready for, so the waits of one lane — S latency, the tail of P V — are covered by the softmax of the other three),
// warps 2-17 softmax + output (lane = (warp - 2) >> 2, TMEM lane quarter = warp & 3).
// Algorithmic HBM bytes per (token, head): read q, k, v (3 x 64 B; OCA k, v 2.25x), write out 64 B + lse 4 B.
#pragma once
#include "attn_win16.cuh"
#include "attn_tc8.cuh"

namespace srk {

constexpr int TC16_THREADS = 64 + 256;
constexpr int TC16_QUAD = 64 * 128;          // one [64 tokens x 64 ch] quadrant box (8 KB)
constexpr int TC16_TSTRIDE = 40;             // padded row stride of the bias table in shared memory: a warp's 4 x 8
                                             // (query row, query column) offsets fall into 32 distinct banks

template <int MODE>
struct TC16 {
  static constexpr int NCHUNK = (MODE == MODE_SELF) ? 4 : 9;   // 64-key chunks (quadrant boxes) of the key window
  static constexpr int NPART = (MODE == MODE_SELF) ? 1 : 3;    // key parts per (head, half window)
  static constexpr int PCH = NCHUNK / NPART;                   // chunks per part: S has N = PCH * 64 columns
  static constexpr int SPU = 2 * NPART;                        // sub-rounds per unit and lane
  static constexpr int VST = (MODE == MODE_SELF) ? 2 : 1;      // V stages
  static constexpr int PSLOTS = (MODE == MODE_SELF) ? 2 : 1;   // P chunk slots per lane
  static constexpr int TSIDE = A16<MODE>::TSIDE;               // 31 / 39 table rows
  static constexpr int TBLP = TSIDE * TC16_TSTRIDE;            // padded table entries per head
  static constexpr int KV_BYTES = NCHUNK * TC16_QUAD;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = 4 * TC16_QUAD;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_P = OFF_V + VST * KV_BYTES;
  static constexpr int OFF_TBL = OFF_P + 2 * PSLOTS * 128 * 128;
  static constexpr int OFF_BAR = OFF_TBL + 2 * TBLP * 4;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
};

template <int MODE>
__global__ void __launch_bounds__(TC16_THREADS, 1)
win_attn_tc16_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const Attn16Args a) {
  using G = TC16<MODE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_ptr = smem_raw + (smem_base - smem_u32(smem_raw));
  float* s_table = reinterpret_cast<float*>(smem_ptr + G::OFF_TBL);
  const uint32_t bar_base = smem_base + G::OFF_BAR;
  const uint32_t qk_full = bar_base, qk_empty = bar_base + 8;
  auto v_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 32u + 8u * s; };
  auto s_full = [&](int l) { return bar_base + 48u + 8u * l; };
  auto s_empty = [&](int l) { return bar_base + 64u + 8u * l; };
  auto o_full = [&](int l) { return bar_base + 80u + 8u * l; };
  auto p_full = [&](int l, int s) { return bar_base + 96u + 8u * (l * 2 + s); };
  auto p_empty = [&](int l, int s) { return bar_base + 128u + 8u * (l * 2 + s); };
  const uint32_t tmem_slot = bar_base + 160u;
  __shared__ float s_tmax[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwx = a.W >> 4, nwy = a.H >> 4;
  const int nwin = a.B * nwx * nwy;
  const int hp = blockIdx.y;
  const int my_wins = (nwin - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int AW = a.heads * 32;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(qk_full, 1); mbar_init(qk_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    for (int l = 0; l < 2; ++l) {
      mbar_init(s_full(l), 1); mbar_init(s_empty(l), 4); mbar_init(o_full(l), 1);
      for (int s = 0; s < 2; ++s) { mbar_init(p_full(l, s), 4); mbar_init(p_empty(l, s), 1); }
    }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // bias table of the two heads, rows padded to TC16_TSTRIDE, pre-multiplied by log2(e).  MODE_OCA: rotated by 880 entries
  // so that the reference's negative indices (wrapped around the table end by PyTorch, hat_arch.py:896-919) are plain offsets.
  constexpr float kLog2e = 1.4426950408889634f;
  {
    constexpr int TBL = A16<MODE>::TBL;
    float mx0 = -INFINITY, mx1 = -INFINITY;
    for (int i = threadIdx.x; i < TBL; i += TC16_THREADS) {
      int src = i;
      if (MODE == MODE_OCA) { src = i - 880; if (src < 0) src += TBL; }
      const int dst = (i / G::TSIDE) * TC16_TSTRIDE + (i % G::TSIDE);
      const float v0 = a.bias_table[src * a.heads + 2 * hp] * kLog2e, v1 = a.bias_table[src * a.heads + 2 * hp + 1] * kLog2e;
      s_table[dst] = v0;
      s_table[G::TBLP + dst] = v1;
      mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    }
    __shared__ float s_red[2][TC16_THREADS / 32];
    if (lane == 0) { s_red[0][warp] = mx0; s_red[1][warp] = mx1; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 2) {
      float m = -INFINITY;
      for (int w = 0; w < TC16_THREADS / 32; ++w) m = fmaxf(m, s_red[threadIdx.x][w]);
      s_tmax[threadIdx.x] = m;
    }
    __syncthreads();
  }
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // window origin (image coordinates of window-local (0, 0) in the frame the boxes are taken from)
  auto win_origin = [&](int i, int& b, int& y0, int& x0, bool& last_y, bool& last_x) {
    const int w = int(blockIdx.x) + i * int(gridDim.x);
    b = w / (nwx * nwy);
    const int r = w - b * nwx * nwy;
    const int wy = r / nwx, wx = r - wy * nwx;
    y0 = wy * 16; x0 = wx * 16;
    last_y = wy == nwy - 1; last_x = wx == nwx - 1;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < my_wins; ++i) {
        int b, y0, x0; bool ly, lx;
        win_origin(i, b, y0, x0, ly, lx);
        mbar_wait(qk_empty, (uint32_t(i) & 1u) ^ 1u);
        mbar_arrive_expect_tx(qk_full, 4 * TC16_QUAD + G::KV_BYTES);
#pragma unroll
        for (int quad = 0; quad < 4; ++quad) {
          int y = y0 + a.shift + (quad >> 1) * 8, x = x0 + a.shift + (quad & 1) * 8;
          if (y >= a.H) y -= a.H;
          if (x >= a.W) x -= a.W;
          tma_load_4d(smem_base + G::OFF_Q + quad * TC16_QUAD, &tmQKV, qk_full, hp * 64, x, y, b);
          if (MODE == MODE_SELF) tma_load_4d(smem_base + G::OFF_K + quad * TC16_QUAD, &tmQKV, qk_full, AW + hp * 64, x, y, b);
        }
        if (MODE == MODE_OCA) {
#pragma unroll
          for (int c = 0; c < 9; ++c)
            tma_load_4d(smem_base + G::OFF_K + c * TC16_QUAD, &tmQKV, qk_full, AW + hp * 64, x0 - 4 + (c % 3) * 8,
                        y0 - 4 + (c / 3) * 8, b);
        }
        const int vs = i % G::VST;
        mbar_wait(v_empty(vs), (uint32_t(i / G::VST) & 1u) ^ 1u);
        mbar_arrive_expect_tx(v_full(vs), G::KV_BYTES);
        const uint32_t vdst = smem_base + G::OFF_V + vs * G::KV_BYTES;
        if (MODE == MODE_SELF) {
#pragma unroll
          for (int quad = 0; quad < 4; ++quad) {
            int y = y0 + a.shift + (quad >> 1) * 8, x = x0 + a.shift + (quad & 1) * 8;
            if (y >= a.H) y -= a.H;
            if (x >= a.W) x -= a.W;
            tma_load_4d(vdst + quad * TC16_QUAD, &tmQKV, v_full(vs), 2 * AW + hp * 64, x, y, b);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 9; ++c)
            tma_load_4d(vdst + c * TC16_QUAD, &tmQKV, v_full(vs), 2 * AW + hp * 64, x0 - 4 + (c % 3) * 8, y0 - 4 + (c / 3) * 8, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: event loop over the two lanes
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, G::PCH * 64, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // B = V, MN-major
      const int total = my_wins * G::SPU;
      int n[2] = {0, 0};        // sub-round counter of the lane
      int j[2] = {-1, -1};      // -1: S of sub-round n not issued yet; 0..PCH-1: next P V chunk
      int s_issued = 0;         // S products issued for the unit whose Q / K are resident
      int pv_done[2] = {0, 0};  // finished sub-rounds per V stage
      while (n[0] < total || n[1] < total) {
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          if (n[l] >= total) continue;
          const int i = n[l] / G::SPU, sr = n[l] - i * G::SPU;
          const int hh = sr / G::NPART, part = sr - hh * G::NPART;
          const uint32_t d = tmem_base + uint32_t(l * 256);
          if (j[l] < 0) {
            if (!mbar_test_wait(s_empty(l), (uint32_t(n[l]) & 1u) ^ 1u)) continue;
            if (!mbar_test_wait(qk_full, uint32_t(i) & 1u)) continue;
            tc_fence_after();
            const uint32_t qa = smem_base + G::OFF_Q + l * 2 * TC16_QUAD + hh * 64;
            const uint32_t kb = smem_base + G::OFF_K + part * G::PCH * TC16_QUAD + hh * 64;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d, make_smem_desc(qa + k * 32, 16, 1024), make_smem_desc(kb + k * 32, 16, 1024), idesc_s, k);
            umma_commit(s_full(l));
            if (++s_issued == 2 * G::SPU) { s_issued = 0; umma_commit(qk_empty); }   // Q and K of this unit are dead
            j[l] = 0;
          } else {
            const int c = n[l] * G::PCH + j[l];           // running chunk counter of the lane
            const int slot = c % G::PSLOTS;
            if (!mbar_test_wait(p_full(l, slot), uint32_t(c / G::PSLOTS) & 1u)) continue;
            const int vs = i % G::VST;
            if (!mbar_test_wait(v_full(vs), uint32_t(i / G::VST) & 1u)) continue;
            tc_fence_after();
            const uint32_t pt = smem_base + G::OFF_P + (l * G::PSLOTS + slot) * 128 * 128;
            const uint32_t vt = smem_base + G::OFF_V + vs * G::KV_BYTES + (part * G::PCH + j[l]) * TC16_QUAD;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d, make_smem_desc(pt + k * 32, 16, 1024), make_smem_desc_mn(vt + k * 2048, TC16_QUAD), idesc_o,
                        (j[l] | k) != 0);
            umma_commit(p_empty(l, slot));
            if (++j[l] == G::PCH) {
              umma_commit(o_full(l));
              j[l] = -1;
              ++n[l];
              if (++pv_done[vs] == 2 * G::SPU) { pv_done[vs] = 0; umma_commit(v_empty(vs)); }   // V of this unit is dead
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / output warps
    const int l = (warp - 2) >> 2;           // lane = query half (quadrant row) of the window
    const int q = warp & 3;                  // TMEM lane quarter
    const int row = q * 32 + lane;           // 0..127 within the half: quadrant column row >> 6, yl, xl
    const int qxq = row >> 6, yl = (row >> 3) & 7, xl = row & 7;
    const int qy = l * 8 + yl, qx = qxq * 8 + xl;   // window-local query coordinates
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    const uint32_t tm_lane = tmem_base + lane_sel + uint32_t(l * 256);
    // table offset of key (ky, kx):  SELF (qy - ky + 15, qx - kx + 15);  OCA rotated index (ky - qy - 7)*39 + (kx - qx - 7) + 880
    // = (ky - qy + 15, kx - qx + 15) in (row, column) form since 880 = 22 * 39 + 22.
    const int base_q = (MODE == MODE_SELF) ? (qy + 15) * TC16_TSTRIDE + (qx + 15) : (15 - qy) * TC16_TSTRIDE + (15 - qx);
    const bool shifted = (MODE == MODE_SELF) && a.shift > 0;
    int c_run = 0;   // running chunk counter of this lane (same sequence as the MMA issuer's)
    int n_run = 0;   // running sub-round counter
    for (int i = 0; i < my_wins; ++i) {
      int b, y0, x0; bool last_y, last_x;
      win_origin(i, b, y0, x0, last_y, last_x);
      int y = y0 + a.shift + qy, x = x0 + a.shift + qx;
      if (y >= a.H) y -= a.H;
      if (x >= a.W) x -= a.W;
      const long long tok = (long long)(b * a.H + y) * a.W + x;
      float mq[4] = {0.f, 0.f, 0.f, 0.f};   // additive mask (log2 domain) per key quadrant, hat_arch.py:921-940
      if (shifted && (last_y || last_x)) {
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
          const bool yd = last_y && (l != (kq >> 1));
          const bool xd = last_x && (qxq != (kq & 1));
          mq[kq] = (yd || xd) ? -100.0f * kLog2e : 0.0f;
        }
      }
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int head = 2 * hp + hh;
        const float* tb = s_table + hh * G::TBLP + base_q;
        const float tmax = s_tmax[hh];
        float m_run = -INFINITY, l_run = 0.f;
        float oacc[32];
        if (G::NPART > 1) {
#pragma unroll
          for (int e = 0; e < 32; ++e) oacc[e] = 0.f;
        }
#pragma unroll 1
        for (int part = 0; part < G::NPART; ++part, ++n_run) {
          mbar_wait(s_full(l), uint32_t(n_run) & 1u);
          tc_fence_after();
          // ---- pass 1: row maximum of the raw logits -> an upper bound of the biased, masked row maximum
          // (the shift mask is constant over a 64-key chunk, so it enters the bound per chunk: without it a row whose largest
          //  raw logit sits on a masked key would be normalised against a maximum 100 too high, and with logits spread
          //  over hundreds the whole row underflows: sum = 0, 0 * inf = NaN in the output)
          float mrow = -INFINITY;
#pragma unroll 1
          for (int jj = 0; jj < G::PCH; ++jj) {
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;   // four independent chains
            uint32_t sv[64];
            tmem_ld_x64(tm_lane + uint32_t(jj * 64), sv);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 64; e += 8) {
              mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
              mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
              mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[e + 4]), __uint_as_float(sv[e + 5])));
              mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[e + 6]), __uint_as_float(sv[e + 7])));
            }
            float mqc = 0.f;
            if (MODE == MODE_SELF) {
              const int cj = part * G::PCH + jj;
              mqc = (cj == 0) ? mq[0] : (cj == 1) ? mq[1] : (cj == 2) ? mq[2] : mq[3];
            }
            mrow = fmaxf(mrow, fmaf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), kLog2e, mqc));
          }
          const float ml2 = mrow + tmax;
          // ---- pass 2: P chunk by chunk
          float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll 1
          for (int jj = 0; jj < G::PCH; ++jj, ++c_run) {
            const int cj = part * G::PCH + jj;   // chunk of the key window: SELF quadrant (cj >> 1, cj & 1), OCA (cj / 3, cj % 3)
            const float* tj;
            float cb;
            if (MODE == MODE_SELF) {
              tj = tb - ((cj >> 1) * 8 * TC16_TSTRIDE + (cj & 1) * 8);
              cb = ((cj == 0) ? mq[0] : (cj == 1) ? mq[1] : (cj == 2) ? mq[2] : mq[3]) - ml2;
            } else {
              tj = tb + (part * 8 * TC16_TSTRIDE + jj * 8);
              cb = -ml2;
            }
            uint32_t sv[64];
            tmem_ld_x64(tm_lane + uint32_t(jj * 64), sv);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 64; ++e) {
              const int koff = (e >> 3) * TC16_TSTRIDE + (e & 7);
              const float bias = (MODE == MODE_SELF) ? tj[-koff] : tj[koff];
              const float p = fast_ex2(fmaf(__uint_as_float(sv[e]), kLog2e, bias + cb));
              if ((e & 3) == 0) sum0 += p; else if ((e & 3) == 1) sum1 += p; else if ((e & 3) == 2) sum2 += p; else sum3 += p;
              sv[e] = __float_as_uint(p);
            }
            const int slot = c_run % G::PSLOTS;
            mbar_wait(p_empty(l, slot), (uint32_t(c_run / G::PSLOTS) & 1u) ^ 1u);
            const uint32_t p_row = smem_base + G::OFF_P + (l * G::PSLOTS + slot) * 128 * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o[e] = pack_bf16(__uint_as_float(sv[c * 8 + 2 * e]), __uint_as_float(sv[c * 8 + 2 * e + 1]));
              sts128(p_row + tc8_swz(row, c), make_uint4(o[0], o[1], o[2], o[3]));
            }
            tc_fence_before();      // this chunk of S has been read: P V may overwrite columns 0..63 once chunk 0 is handed over
            fence_proxy_async();    // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(l, slot));
          }
          // ---- O of this part
          const float sum = (sum0 + sum1) + (sum2 + sum3);
          mbar_wait(o_full(l), uint32_t(n_run) & 1u);
          tc_fence_after();
          uint32_t ov[32];
          tmem_ld_x32(tm_lane + uint32_t(hh * 32), ov);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty(l));
          if (G::NPART == 1) {
            m_run = ml2; l_run = sum;
#pragma unroll
            for (int e = 0; e < 32; ++e) oacc[e] = __uint_as_float(ov[e]);
          } else {
            const float mn = fmaxf(m_run, ml2);
            const float ca = fast_ex2(m_run - mn), cbb = fast_ex2(ml2 - mn);
            l_run = l_run * ca + sum * cbb;
#pragma unroll
            for (int e = 0; e < 32; ++e) oacc[e] = oacc[e] * ca + __uint_as_float(ov[e]) * cbb;
            m_run = mn;
          }
        }
        const float inv = 1.0f / l_run;
        uint32_t ob[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) ob[e] = pack_bf16(oacc[2 * e] * inv, oacc[2 * e + 1] * inv);
        const int oc = a.ones_col - head * 32;   // bias-folding column of the following projection := 1.0
        if (oc >= 0 && oc < 32) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (e == (oc >> 1)) ob[e] = (oc & 1) ? ((ob[e] & 0x0000FFFFu) | 0x3F800000u) : ((ob[e] & 0xFFFF0000u) | 0x00003F80u);
        }
        uint4* op = reinterpret_cast<uint4*>(a.out + tok * a.ld_o + head * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) op[c] = make_uint4(ob[4 * c], ob[4 * c + 1], ob[4 * c + 2], ob[4 * c + 3]);
        if (a.lse != nullptr) a.lse[(long long)head * a.T + tok] = (m_run + log2f(l_run)) * 0.6931471805599453f;
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

}  // namespace srk
