// conv3x3_halo.cuh — halo-resident variant of the implicit-GEMM 3x3 convolution (conv3x3.cuh) for plain NHWC views.
//
// conv3x3_kernel fetches one TMA box per (tap, 64-channel chunk): the input crosses the L2 -> SM crossbar nine times,
// and ncu shows the thin dense-block layers of the hybrid generator (hybridmodels_hat.py:21-44) bound by exactly that
// (l1tex__m_xbar2l1tex_read_bytes 1.8 GB for a 150 MB layer, 11 TB/s; profiles/r01_ncu_full_rdb_*.txt).
// Here the CTA loads the input tile ONCE per 64-channel chunk including its 1-pixel border (box 64 ch x 16 x 18
// pixels at (x0-1, y0-1); TMA zero-fill is still the padding) and issues the nine taps as nine row-shifted UMMA
// descriptors over the same shared-memory tile:
//   output tile  8 wide x 16 tall = 128 pixels, accumulator row r = y*8 + x  (one 8-row swizzle group per image row)
//   halo tile    rows hy*16 + hx (hx 0..15, of which 0..9 are used; hy 0..17), 128 B per row, 128B-swizzled by TMA
//   tap (ky,kx)  descriptor start = tile + ((ky*16 + kx) * 128) B, SBO = 2048 B (next image row).  The start is not
//                1024-byte aligned any more; measured on B200: the UMMA unit applies the 128B swizzle to the absolute
//                shared-memory address (as TMA does when it writes the tile), so the descriptor's matrix-base-offset
//                field must stay 0 — setting it to kx produced wrong results (round-1 probe, since removed from the library).
// Weights stream through their own ring, one [BN x 64] box per (chunk, tap).  Epilogues as conv3x3_kernel.
#pragma once
#include "conv3x3.cuh"

namespace srk {

constexpr int HALO_TW = 8, HALO_TH = 16;     // output tile
constexpr int HALO_BW = 16, HALO_BH = 18;    // input box (pixels), incl. border and the pad columns that keep SBO uniform
constexpr int HALO_A_BYTES = HALO_BW * HALO_BH * 128;

template <int BN, int EPI>
struct HaloCfg {
  static constexpr int kWBytes = BN * 128;
  static constexpr int kBoxes = (BN == 16) ? 0 : (BN + 63) / 64;   // BN = 32: half a box (the store is clipped by the view)
  static constexpr bool kAux = (EPI == CEPI_BIAS_RES || EPI == CEPI_MASK_LRELU || EPI == CEPI_MUL);
  // halo ring depth: three tiles in flight where shared memory allows (thin layers: the TMA latency of a 36 KB box is
  // what the two-stage ring exposed), two otherwise
  static constexpr int kAStages = (BN <= 64 && !kAux) ? 3 : 2;
  static constexpr int kOutPerBox = 1;
  static constexpr int kEpiBytes = (kAux ? 2 * BOX_BYTES : 0) + 2 * BOX_BYTES;
  static constexpr int kBudget = 232448 - 1024 - 1024 - 1280;
  static constexpr int kWRaw = (kBudget - kEpiBytes - kAStages * HALO_A_BYTES) / kWBytes;
  static constexpr int kWCap = (BN <= 32) ? 18 : 9;
  static constexpr int kWStages = kWRaw > kWCap ? kWCap : kWRaw;
  static constexpr int kSmemBytes = kAStages * HALO_A_BYTES + kWStages * kWBytes + kEpiBytes + 1024 + 1024;
  static_assert(kWStages >= 2, "weight ring needs two stages");
  static_assert(BN % 64 == 0 || BN == 16 || BN == 32, "BN: multiple of 64, 32 (thin layers) or 16 (single channel)");
  static_assert(EPI != CEPI_BIAS_GELU, "the GELU epilogue (two outputs) stays on conv3x3_kernel");
};

// K-major 128B-swizzled operand whose start is not 1024-byte aligned: matrix base offset = (start >> 7) & 7
__device__ __forceinline__ uint64_t make_smem_desc_bo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_off) {
  return make_smem_desc(saddr, lbo_bytes, sbo_bytes) | (static_cast<uint64_t>(base_off & 7u) << 49);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args) {
  using Cfg = HaloCfg<BN, EPI>;
  constexpr int SW = Cfg::kWStages, SA = Cfg::kAStages;
  constexpr int S = SA + SW;  // barrier slots: [0, SA) halo ring, [SA, S) weight ring
  constexpr int NBOX = Cfg::kBoxes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base + SA * HALO_A_BYTES;
  const uint32_t epi_base = w_base + SW * Cfg::kWBytes;
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  auto aux_bar = [&](int b) { return bar_base + 8u * (2 * S + 4 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 6);
  __shared__ float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = args.W / HALO_TW, tiles_y = args.H / HALO_TH;
  const int n_tiles = args.Cout_p / BN;
  const int m_tiles = args.B * tiles_y * tiles_x;
  const int num_tiles = m_tiles * n_tiles;
  const int kc_per_tap = args.Cin_p / 64;
  // All nine taps of every 64-channel chunk fit the weight ring and there is one N tile: load the weights once per CTA
  // and keep them (the ring slots are never released), instead of re-fetching them from L2 for every spatial tile.
  const bool w_resident = (9 * kc_per_tap <= SW) && (n_tiles == 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), GEMM_EPI_THREADS / 32);
      mbar_init(aux_bar(a), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 256; i += GEMM_THREADS)
    s_bias[i] = (args.bias != nullptr && i < args.n_real) ? args.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto tile_coords = [&](int tile, int& b, int& y0, int& x0, int& n0) {
    const int mt = tile / n_tiles;
    n0 = (tile % n_tiles) * BN;
    b = mt / (tiles_y * tiles_x);
    const int r = mt % (tiles_y * tiles_x);
    y0 = (r / tiles_x) * HALO_TH;
    x0 = (r % tiles_x) * HALO_TW;
  };

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, ws = 0; uint32_t aph = 0, wph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int b, y0, x0, n0;
        tile_coords(tile, b, y0, x0, n0);
        for (int kc = 0; kc < kc_per_tap; ++kc) {
          mbar_wait(empty_bar(as), aph ^ 1u);
          {
            mbar_arrive_expect_tx(full_bar(as), HALO_A_BYTES);
            // three boxes of 6 image rows each (same bytes as one 18-row box; no measurable difference on B200)
#pragma unroll
            for (int part = 0; part < 3; ++part)
              tma_load_4d(smem_base + as * HALO_A_BYTES + part * (6 * HALO_BW * 128), &maps.a[0], full_bar(as), kc * 64,
                          x0 - 1, y0 - 1 + part * 6, b);
          }
          if (++as == SA) { as = 0; aph ^= 1u; }
          if (w_resident && tile != int(blockIdx.x)) continue;   // weights already sit in shared memory
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(empty_bar(SA + ws), wph ^ 1u);
            mbar_arrive_expect_tx(full_bar(SA + ws), Cfg::kWBytes);
            tma_load_2d(w_base + ws * Cfg::kWBytes, &maps.w, full_bar(SA + ws), (tap * kc_per_tap + kc) * 64, n0);
            if (++ws == SW) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
      int as = 0, ws = 0; uint32_t aph = 0, wph = 0; int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kc = 0; kc < kc_per_tap; ++kc) {
          mbar_wait(full_bar(as), aph);
          tc_fence_after();
          const uint32_t sa = smem_base + as * HALO_A_BYTES;
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            if (w_resident) ws = kc * 9 + tap;            // slot of this (chunk, tap); its first phase stays complete
            mbar_wait(full_bar(SA + ws), w_resident ? 0u : wph);
            tc_fence_after();
            const uint32_t a0 = sa + uint32_t(ky * HALO_BW + kx) * 128u;
            const uint32_t sb = w_base + ws * Cfg::kWBytes;
            const uint32_t bo = 0u;   // matrix base offset stays 0: the 128B swizzle is applied to absolute smem addresses
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, make_smem_desc_bo(a0 + k * 32, 16, HALO_BW * 128, bo), make_smem_desc(sb + k * 32, 16, 1024),
                        idesc, (kc | tap | k) != 0 ? 1u : 0u);
            if (!w_resident) {
              umma_commit(empty_bar(SA + ws));
              if (++ws == SW) { ws = 0; wph ^= 1u; }
            }
          }
          umma_commit(empty_bar(as));
          if (++as == SA) { as = 0; aph ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool elected = (threadIdx.x == 64);
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    constexpr int kAuxOff = 0;
    constexpr int kOutOff = Cfg::kAux ? 2 * BOX_BYTES : 0;
    constexpr int kOutPerBox = Cfg::kOutPerBox;
    uint32_t box_counter = 0, aux_count = 0;
    int it = 0;
    if constexpr (Cfg::kAux) {
      if (elected && blockIdx.x < num_tiles) {
        int b, y0, x0, n0;
        tile_coords(blockIdx.x, b, y0, x0, n0);
        mbar_arrive_expect_tx(aux_bar(0), BOX_BYTES);
        tma_load_4d(epi_base, &maps.r, aux_bar(0), n0, x0, y0, b);
      }
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int b, y0, x0, n0;
      tile_coords(tile, b, y0, x0, n0);
      const int acc = it & 1;
      const uint32_t taddr = tmem_base + lane_sel + uint32_t(acc * BN);
      const int next_tile = tile + gridDim.x;
      mbar_wait(tfull_bar(acc), (it >> 1) & 1u);
      tc_fence_after();
      if constexpr (EPI == CEPI_OUT1) {
        // one real output channel: column 0 of the accumulator, row = pixel of the 16 x 8 spatial tile
        float v = 0.f;
        if (half == 0) {
          v = __uint_as_float(tmem_ld_x1(taddr));
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (half == 0) {
          const int py = y0 + (row >> 3), px = x0 + (row & 7);
          args.y32[((size_t)b * args.H + py) * args.W + px] = v + s_bias[0];
        }
      }
#pragma unroll 1
      for (int j = 0; j < NBOX; ++j) {
        const uint32_t ring = box_counter & 1u;
        const uint32_t out0 = epi_base + kOutOff + ring * (kOutPerBox * BOX_BYTES);
        if (elected) {
          tma_store_wait_read<1>();
          if constexpr (Cfg::kAux) {
            int nb_, ny0 = y0, nx0 = x0, nn0 = n0 + (j + 1) * 64, nbb = b;
            bool have = true;
            if (j + 1 == NBOX) {
              have = next_tile < num_tiles;
              if (have) tile_coords(next_tile, nbb, ny0, nx0, nn0);
            }
            (void)nb_;
            if (have) {
              const uint32_t nb = (aux_count + 1) & 1u;
              mbar_arrive_expect_tx(aux_bar(nb), BOX_BYTES);
              tma_load_4d(epi_base + kAuxOff + nb * BOX_BYTES, &maps.r, aux_bar(nb), nn0, nx0, ny0, nbb);
            }
          }
        }
        named_bar_sync(1, GEMM_EPI_THREADS);
        uint32_t aux_addr = 0;
        if constexpr (Cfg::kAux) {
          const uint32_t ab = aux_count & 1u;
          mbar_wait(aux_bar(ab), (aux_count >> 1) & 1u);
          aux_addr = epi_base + kAuxOff + ab * BOX_BYTES;
        }
        const bool active = (BN >= 64) || (half == 0);   // BN = 32: the accumulator has one 32-column half only
        uint32_t r[32];
        if (active) {
          tmem_ld_x32(taddr + uint32_t(j * 64 + half * 32), r);
          tmem_ld_wait();
        }
        if (j == NBOX - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (active)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = half * 4 + i;
          const uint32_t off = swz(row, ch);
          const int col0 = n0 + j * 64 + ch * 8;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[i * 8 + e]);
          if constexpr (EPI == CEPI_BIAS || EPI == CEPI_BIAS_LRELU || EPI == CEPI_BIAS_RES || EPI == CEPI_BIAS_GELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += s_bias[(col0 + e) & 255];
          }
          if constexpr (EPI == CEPI_BIAS_LRELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = round_bf16(v[e]) > 0.f ? v[e] : v[e] * args.slope;
          }
          if constexpr (Cfg::kAux) {
            const uint4 g = lds128(aux_addr + off);
            const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float a = (e & 1) ? bf16_hi(gw[e >> 1]) : bf16_lo(gw[e >> 1]);
              if constexpr (EPI == CEPI_BIAS_RES) v[e] = round_bf16(v[e] * args.alpha) + a;
              else if constexpr (EPI == CEPI_MASK_LRELU) v[e] = (a > 0.f) ? v[e] : v[e] * args.slope;
              else v[e] = round_bf16(v[e]) * a;
            }
          }
          if constexpr (EPI == CEPI_BIAS_GELU) {
            float a[8], g[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) gelu_pair(v[e], a[e], g[e]);
            sts128(out0 + off, make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]),
                                          pack_bf16(a[6], a[7])));
            sts128(out0 + BOX_BYTES + off, make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]),
                                                      pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7])));
          } else {
            sts128(out0 + off, make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                          pack_bf16(v[6], v[7])));
          }
        }
        fence_proxy_async();
        named_bar_sync(1, GEMM_EPI_THREADS);
        if (elected) {
          tma_store_4d(&maps.c[0], out0, n0 + j * 64, x0, y0, b);
          tma_store_commit();
        }
        ++box_counter;
        if constexpr (Cfg::kAux) ++aux_count;
      }
    }
    if (elected) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


}  // namespace srk

namespace srk {
// ------------------------------------------------------------------------------------------------ thin wgrad
// Weight gradient of a 3x3 convolution with FEW output channels (Cout <= 32: conv1..conv4 of a residual dense block,
// hybridmodels_hat.py:24-27), halo-resident:  dW[tap][ci][co] = sum over pixels of X[p + tap][ci] * dY[p][co].
// conv3x3_wgrad_kernel runs one CTA per (tap, split) and so pulls X and dY through the crossbar nine times (ncu: 2.6 GB
// for a 250 MB layer).  Here a CTA owns a pixel range and ALL nine taps: per 16 x 4 pixel patch it loads dY once
// (B operand, N = 32) and X once with its border (A operand, M = 128 input channels, box 64 ch x 18 x 6 pixels), and
// issues 9 taps x 4 image rows of K = 16 pixels as row-shifted MN-major descriptors into nine [128 x 32] fp32
// accumulators that sit side by side in TMEM (288 columns).  grid = ci_tiles x splits; partials reduced afterwards.
constexpr int WT_XBOX = 14 * 1024;                  // 18 * 6 * 128 B = 13824, padded so every box base is 1024-aligned
constexpr int WT_XBOX_TX = 18 * 6 * 128;
constexpr int WT_DYBOX = 16 * 4 * 128;
constexpr int WT_STAGE = 2 * WT_XBOX + WT_DYBOX;    // 36864
constexpr int WT_STAGES = 5;
constexpr int WT_SMEM = WT_STAGES * WT_STAGE + 256 + 1024;
constexpr int WT_NCO_MAX = 48;   // instances: N = 32 (conv1..conv4, 24 or 32 channels) and N = 48 (conv5 at num_feat 48)

struct ConvWgradThinArgs {
  int B, H, W, ci_tiles, splits;
  float* partials;  // [splits][ci_tiles][128][9][NCO]
};
struct ConvWgradThinMaps {
  CUtensorMap dy;  // box (64, 16, 4, 1)
  CUtensorMap x;   // box (64, 18, 6, 1)
};

template <int NCO>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv3x3_wgrad_thin_kernel(const __grid_constant__ ConvWgradThinMaps maps, const ConvWgradThinArgs args) {
  constexpr int S = WT_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * WT_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci_tile = blockIdx.x % args.ci_tiles;
  const int split = blockIdx.x / args.ci_tiles;
  const int px = args.W / 16, py = args.H / 4;
  const int total_iters = args.B * py * px;  // 64-pixel patches
  const int it_begin = int((long long)split * total_iters / args.splits);
  const int it_end = int((long long)(split + 1) * total_iters / args.splits);
  const int k_iters = it_end - it_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < k_iters; ++kb) {
        const int p = it_begin + kb;
        const int b = p / (py * px), r = p % (py * px);
        const int y0 = (r / px) * 4, x0 = (r % px) * 16;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sx = smem_base + stage * WT_STAGE;
        mbar_arrive_expect_tx(full_bar(stage), 2 * WT_XBOX_TX + WT_DYBOX);
        tma_load_4d(sx, &maps.x, full_bar(stage), (ci_tile * 2) * 64, x0 - 1, y0 - 1, b);
        tma_load_4d(sx + WT_XBOX, &maps.x, full_bar(stage), (ci_tile * 2 + 1) * 64, x0 - 1, y0 - 1, b);
        tma_load_4d(sx + 2 * WT_XBOX, &maps.dy, full_bar(stage), 0, x0, y0, b);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, NCO, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < k_iters; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sx = smem_base + stage * WT_STAGE;
        const uint32_t sdy = sx + 2 * WT_XBOX;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
          for (int y = 0; y < 4; ++y)
            umma_bf16(tmem_base + uint32_t(tap * NCO),
                      make_smem_desc(sx + uint32_t((y + ky) * 18 + kx) * 128u, WT_XBOX, 1024),
                      make_smem_desc(sdy + uint32_t(y) * 2048u, WT_DYBOX, 1024), idesc, (kb | y) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
    float* out = args.partials + ((size_t(split) * args.ci_tiles + ci_tile) * 128 + row) * (9 * NCO);
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      uint32_t r[32];
      tmem_ld_x32(taddr + uint32_t(tap * NCO), r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<uint4*>(out + tap * NCO)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
      if constexpr (NCO > 32) {
        uint32_t r2[16];
        tmem_ld_x16(taddr + uint32_t(tap * NCO + 32), r2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<uint4*>(out + tap * NCO + 32)[i] = make_uint4(r2[4 * i], r2[4 * i + 1], r2[4 * i + 2], r2[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// partials [splits][ci_tiles][128][9][nco] -> dW [Cout][Cin][3][3]; one thread per partial element (coalesced over splits)
static __global__ void conv_unpack_wgrad_thin_kernel(const float* __restrict__ part, int splits, int ci_tiles, int nco,
                                                     float* __restrict__ dw, int Cout, int Cin) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_split = ci_tiles * 128 * 9 * nco;
  if (e >= per_split) return;
  const int co = e % nco, tap = (e / nco) % 9, ci = e / (9 * nco);
  if (co >= Cout || ci >= Cin) return;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += part[size_t(s) * per_split + e];
  dw[(size_t(co) * Cin + ci) * 9 + tap] = acc;
}
}  // namespace srk
