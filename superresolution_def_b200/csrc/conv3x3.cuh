// conv3x3.cuh — 3x3 / stride 1 / pad 1 convolution on NHWC bf16 activations as an implicit GEMM on tcgen05.
//
//   Y[b,y,x,n] = epilogue( sum_{ky,kx,c} X[b, y+ky-1, x+kx-1, c] * Wk[n, (ky*3+kx)*Cin_p + c] )
//
// No im2col buffer exists: for every tap the A tile is one 4-D TMA box (64 ch x 16 x 8 pixels) fetched at the
// shifted coordinates; coordinates outside the image are zero-filled by the TMA unit, which *is* the padding.
// The same kernel computes the input gradient (X := dY, Wk := flipped/transposed weights).
// PixelShuffle(2) is free: the caller hands four tensor maps (one per sub-pixel (i,j)) that view the shuffled
// tensor [B,2H,2W,64] with doubled pixel strides, and each 64-channel box of the tile is stored (or, for the
// gradient, loaded) through the map of its sub-pixel — the shuffle is pure address arithmetic.
//
// Replaces (reference): nn.Conv2d 3x3 in SwinIR.conv_after_body / conv_before_upsample(+LeakyReLU) / Upsample
// (+nn.PixelShuffle) (models/architecture_swin.py:222-229,175-190), HAT's CAB / RHAG / head convs
// (models/hat_arch/hat_arch.py:66-74,608,859-868), and their autograd input gradients.
#pragma once
#include "gemm_tn.cuh"
#include "gemm_wgrad.cuh"

namespace srk {

constexpr int CONV_TW = 16, CONV_TH = 8;  // 128-pixel spatial tile

enum ConvEpilogue : int {
  CEPI_BIAS = 0,       // y = acc + bias
  CEPI_BIAS_LRELU = 1, // y = leaky_relu(acc + bias, slope)
  CEPI_BIAS_RES = 2,   // y = alpha * (acc + bias) + R  (R: same layout as Y; alpha = 1 except the RDB's 0.2 residual scale)
  CEPI_MASK_LRELU = 3, // y = acc * (R > 0 ? 1 : slope)  (backward through LeakyReLU; R = forward output)
  CEPI_BIAS_GELU = 4,  // y = gelu(acc + bias), y2 = gelu'(acc + bias)   (HAT CAB, hat_arch.py:69)
  CEPI_MUL = 5,        // y = acc * R                    (backward through GELU; R = gelu')
  CEPI_OUT1 = 6,       // BN = 16, one real output channel: y32[pixel] = acc[:,0] + bias[0]   (conv_last, fp32 out)
};

struct ConvArgs {
  int B, H, W;        // output spatial size == input spatial size
  int Cin_p, Cout_p;  // padded channel counts (multiples of 64)
  int n_real;         // real output channels (bias length)
  const float* bias;  // [n_real] or nullptr
  float slope;
  int a_split;        // 1: A k-chunk kc is loaded through tmA[kc] at channel 0 (pixel-shuffled input gradient)
  int c_split;        // 1: output box j is stored through tmC[j] at channel 0 (pixel-shuffled output)
  float* y32;         // CEPI_OUT1: fp32 output [B,H,W]
  float alpha;        // CEPI_BIAS_RES: scale of the convolution branch
};

struct ConvMaps {
  CUtensorMap a[4];   // input  [B,H,W,C] 4-D maps, box (64, 16, 8, 1)
  CUtensorMap c[4];   // output 4-D maps
  CUtensorMap c2;     // second output (GELU')
  CUtensorMap r;      // residual / mask input
  CUtensorMap w;      // weights 2-D [Cout_p, 9*Cin_p], box (64, BN)
};

template <int BN, int EPI>
struct ConvCfg {
  static constexpr int kStageBytes = GEMM_BM * 128 + BN * 128;
  static constexpr int kBoxes = BN / 64;  // 0 for the BN = 16 single-channel variant
  static constexpr bool kAux = (EPI == CEPI_BIAS_RES || EPI == CEPI_MASK_LRELU || EPI == CEPI_MUL);
  static constexpr int kOutPerBox = (EPI == CEPI_BIAS_GELU) ? 2 : 1;
  static constexpr int kEpiBytes = (kAux ? 2 * BOX_BYTES : 0) + 2 * kOutPerBox * BOX_BYTES;
  static constexpr int kBudget = 232448 - 1024 - 512 - 1280;
  static constexpr int kStagesRaw = (kBudget - kEpiBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 512 + 1024;
  static_assert(kStages >= 2, "conv pipeline needs two stages");
  static_assert(BN % 64 == 0 || BN == 16, "BN: multiple of 64, or 16 for the single-channel variant");
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv3x3_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args) {
  using Cfg = ConvCfg<BN, EPI>;
  constexpr int S = Cfg::kStages;
  constexpr int NBOX = Cfg::kBoxes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + S * Cfg::kStageBytes;
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  auto aux_bar = [&](int b) { return bar_base + 8u * (2 * S + 4 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 6);
  __shared__ float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = args.W / CONV_TW, tiles_y = args.H / CONV_TH;
  const int n_tiles = args.Cout_p / BN;
  const int m_tiles = args.B * tiles_y * tiles_x;
  const int num_tiles = m_tiles * n_tiles;
  const int kc_per_tap = args.Cin_p / 64;
  const int k_iters = 9 * kc_per_tap;

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
    y0 = (r / tiles_x) * CONV_TH;
    x0 = (r % tiles_x) * CONV_TW;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int b, y0, x0, n0;
        tile_coords(tile, b, y0, x0, n0);
        for (int kb = 0; kb < k_iters; ++kb) {
          const int tap = kb / kc_per_tap, kc = kb % kc_per_tap;
          const int ky = tap / 3, kx = tap % 3;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + GEMM_BM * 128;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          if (args.a_split) tma_load_4d(sa, &maps.a[kc], full_bar(stage), 0, x0 + kx - 1, y0 + ky - 1, b);
          else tma_load_4d(sa, &maps.a[0], full_bar(stage), kc * 64, x0 + kx - 1, y0 + ky - 1, b);
          tma_load_2d(sb, &maps.w, full_bar(stage), kb * 64, n0);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < k_iters; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + GEMM_BM * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
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
          const int py = y0 + (row >> 4), px = x0 + (row & 15);
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
        uint32_t r[32];
        tmem_ld_x32(taddr + uint32_t(j * 64 + half * 32), r);
        tmem_ld_wait();
        if (j == NBOX - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
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
          if (args.c_split) tma_store_4d(&maps.c[(n0 >> 6) + j], out0, 0, x0, y0, b);
          else tma_store_4d(&maps.c[0], out0, n0 + j * 64, x0, y0, b);
          if constexpr (EPI == CEPI_BIAS_GELU) tma_store_4d(&maps.c2, out0 + BOX_BYTES, n0 + j * 64, x0, y0, b);
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

// ------------------------------------------------------------------------------------------------ wgrad
// dW[co][tap][ci] (fp32 partials) = sum over pixels of dY[p][co] * X[p + tap][ci].
// grid = (co_tiles * 9 taps) x splits; operands are MN-major [64 pixels x 64 ch] boxes (pixel patch 16 x 4).
struct ConvWgradArgs {
  int B, H, W, Cin_p, Cout_p, co_tiles, splits;
  float* partials;  // [splits][9][co_tiles*128][Cin_p]
  int a_split;      // dY comes pixel-shuffled: co chunk c64 is loaded through tmA[c64] at channel 0
};
struct ConvWgradMaps {
  CUtensorMap a[4];  // dY maps, box (64, 16, 4, 1)
  CUtensorMap b;     // X map,   box (64, 16, 4, 1)
};

template <int BNW>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv3x3_wgrad_kernel(const __grid_constant__ ConvWgradMaps maps, const ConvWgradArgs args) {
  using Cfg = WgradCfg<BNW, 1>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  constexpr uint32_t kTmemCols = (BNW <= 64) ? 64 : (BNW <= 128) ? 128 : 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x % (args.co_tiles * 9);
  const int split = blockIdx.x / (args.co_tiles * 9);
  const int co_tile = unit / 9, tap = unit % 9;
  const int ky = tap / 3, kx = tap % 3;
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
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
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
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + 2 * WG_SUBBOX;
        mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c64 = co_tile * 2 + h;
          if (args.a_split) tma_load_4d(sa + h * WG_SUBBOX, &maps.a[c64 & 3], full_bar(stage), 0, x0, y0, b);
          else tma_load_4d(sa + h * WG_SUBBOX, &maps.a[0], full_bar(stage), c64 * 64, x0, y0, b);
        }
#pragma unroll
        for (int c = 0; c < BNW / 64; ++c)
          tma_load_4d(sb + c * WG_SUBBOX, &maps.b, full_bar(stage), c * 64, x0 + kx - 1, y0 + ky - 1, b);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BNW, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < k_iters; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + 2 * WG_SUBBOX;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, make_smem_desc(sa + k * 2048, WG_SUBBOX, 1024),
                    make_smem_desc(sb + k * 2048, WG_SUBBOX, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
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
    float* out = args.partials +
                 ((size_t(split) * 9 + tap) * args.co_tiles * 128 + size_t(co_tile) * 128 + row) * BNW;
#pragma unroll 1
    for (int c32 = 0; c32 < BNW / 32; ++c32) {
      uint32_t r[32];
      tmem_ld_x32(taddr + uint32_t(c32 * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<uint4*>(out + c32 * 32)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace srk
