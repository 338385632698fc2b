// conv3x3_swap.cuh — role-swapped halo-resident 3x3 convolution for layers with few output channels.
//
// Why: a tcgen05.mma with M = 128, K = 16 occupies the tensor unit ~128 cycles whatever N <= 256 is (measured on B200,
// DESIGN.md section 4): with pixels as M and 24..64 output channels as N (conv3x3_kernel / conv3x3_halo_kernel) a thin
// layer issues one instruction per 128 pixels x 16 input channels x tap and is bound by that instruction rate.  Here the
// roles are swapped:
//   A operand (M = 128) = weights  [128 output channels x 64 input channels] of one (tap, chunk), K-major; rows beyond
//                         the layer's channel count are zero-filled by TMA (tensor rows < box rows)
//   B operand (N = 256) = 256 pixels of the haloed input tile (8 wide x 32 tall), K-major, row-shifted per tap exactly
//                         as in conv3x3_halo_kernel (box 64 ch x 16 x 34 pixels, SBO = one image row = 2048 B)
//   D [128 x 256] fp32   = TMEM lane = output channel, column = pixel; double-buffered (512 columns)
// so the same instruction now covers 256 pixels: half the tensor-unit cycles per pixel.  The epilogue transposes while it
// stores: thread (lane = channel) writes its pixels as 2-byte elements into [64 pixels x 64 channels] 128B-swizzled boxes
// (a warp covers 64 contiguous bytes of a pixel row: conflict-free), which leave through TMA as NHWC.
//
// Serves (reference): the 24-channel conv1..conv4 and conv5 of ResidualDenseBlock and their input gradients
// (models/hybridmodels_hat.py:21-44), conv_adapt / conv_body / conv_up / conv_hr (:94-105).
#pragma once
#include "conv3x3_halo.cuh"

namespace srk {

constexpr int SWP_TW = 8, SWP_TH = 32;              // output tile: 256 pixels
constexpr int SWP_BW = 16, SWP_BH = 34;             // haloed input box (pixels)
constexpr int SWP_B_BYTES = SWP_BW * SWP_BH * 128;  // 69632
constexpr int SWP_BOX = 64 * 128;                   // output / aux box: 64 pixels x 64 channels
constexpr int SWP_B_STAGES = 2;

// MM = 128: one weight tile covers 128 output channels (two 64-channel boxes per pixel group).
// MM = 64 : layers with <= 64 output channels; the M = 64 instruction streams half the A tile.  Its accumulator rows
//           live in TMEM lanes 0-15 of each 32-lane quarter (row r -> lane 32*(r/16) + r%16; probed on B200 in
//           round 1 against the alternative r -> lane r, which fails parity), so 16 lanes per epilogue warp carry data.
template <int EPI, int MM>
struct SwapCfg {
  static constexpr bool kAux = (EPI == CEPI_BIAS_RES || EPI == CEPI_MASK_LRELU);
  static constexpr int kABytes = MM * 128;
  static constexpr int kAStages = (MM == 64) ? (kAux ? 6 : 8) : (kAux ? 3 : 4);
  static constexpr int kEpiBytes = (kAux ? 2 * SWP_BOX : 0) + 2 * SWP_BOX;
  static constexpr int kSmemBytes = SWP_B_STAGES * SWP_B_BYTES + kAStages * kABytes + kEpiBytes + 1024 + 1024;
  static_assert(kSmemBytes <= 232448 - 1280, "shared memory budget");
  static_assert(EPI == CEPI_BIAS || EPI == CEPI_BIAS_LRELU || EPI == CEPI_BIAS_RES || EPI == CEPI_MASK_LRELU,
                "epilogues of the role-swapped kernel");
};

template <int EPI, int MM>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv3x3_swap_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args) {
  using Cfg = SwapCfg<EPI, MM>;
  constexpr int SWP_A_BYTES = Cfg::kABytes;
  constexpr int SA = Cfg::kAStages, SB = SWP_B_STAGES, S = SA + SB;   // barrier slots: [0, SB) halo ring, [SB, S) weights
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base + SB * SWP_B_BYTES;
  const uint32_t epi_base = a_base + SA * SWP_A_BYTES;
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  auto aux_bar = [&](int b) { return bar_base + 8u * (2 * S + 4 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 6);
  __shared__ float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = args.W / SWP_TW, tiles_y = args.H / SWP_TH;
  const int co_tiles = (args.n_real + MM - 1) / MM;   // MM-channel tiles that hold real output channels
  const int m_tiles = args.B * tiles_y * tiles_x;
  const int num_tiles = m_tiles * co_tiles;
  const int kc_per_tap = args.Cin_p / 64;

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

  auto tile_coords = [&](int tile, int& b, int& y0, int& x0, int& c0) {
    const int mt = tile / co_tiles;
    c0 = (tile % co_tiles) * MM;
    b = mt / (tiles_y * tiles_x);
    const int r = mt % (tiles_y * tiles_x);
    y0 = (r / tiles_x) * SWP_TH;
    x0 = (r % tiles_x) * SWP_TW;
  };

  if (warp == 0) {
    if (lane == 0) {
      int bs = 0, as = 0; uint32_t bph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int b, y0, x0, c0;
        tile_coords(tile, b, y0, x0, c0);
        for (int kc = 0; kc < kc_per_tap; ++kc) {
          mbar_wait(empty_bar(bs), bph ^ 1u);
          mbar_arrive_expect_tx(full_bar(bs), SWP_B_BYTES);
          if (args.a_split)   // pixel-shuffled input gradient: chunk kc is sub-pixel plane kc (its own strided map)
            tma_load_4d(smem_base + bs * SWP_B_BYTES, &maps.a[kc & 3], full_bar(bs), 0, x0 - 1, y0 - 1, b);
          else
            tma_load_4d(smem_base + bs * SWP_B_BYTES, &maps.a[0], full_bar(bs), kc * 64, x0 - 1, y0 - 1, b);
          if (++bs == SB) { bs = 0; bph ^= 1u; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(empty_bar(SB + as), aph ^ 1u);
            mbar_arrive_expect_tx(full_bar(SB + as), SWP_A_BYTES);
            tma_load_2d(a_base + as * SWP_A_BYTES, &maps.w, full_bar(SB + as), (tap * kc_per_tap + kc) * 64, c0);
            if (++as == SA) { as = 0; aph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(MM, 256, 0, 0);
      int bs = 0, as = 0; uint32_t bph = 0, aph = 0; int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * 256);
        for (int kc = 0; kc < kc_per_tap; ++kc) {
          mbar_wait(full_bar(bs), bph);
          tc_fence_after();
          const uint32_t sb = smem_base + bs * SWP_B_BYTES;
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            mbar_wait(full_bar(SB + as), aph);
            tc_fence_after();
            const uint32_t sa = a_base + as * SWP_A_BYTES;
            const uint32_t b0 = sb + uint32_t(ky * SWP_BW + kx) * 128u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(b0 + k * 32, 16, SWP_BW * 128), idesc,
                        (kc | tap | k) != 0 ? 1u : 0u);
            umma_commit(empty_bar(SB + as));
            if (++as == SA) { as = 0; aph ^= 1u; }
          }
          umma_commit(empty_bar(bs));
          if (++bs == SB) { bs = 0; bph ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 8 warps, transposing store
    const int q = warp & 3, half = (warp - 2) >> 2;
    // accumulator row (= output channel within the tile) held by this thread's TMEM lane, or -1
    int row;
    if (MM == 128) row = q * 32 + lane;
    else row = (lane < 16) ? q * 16 + lane : -1;                        // row r -> lane 32*(r/16) + r%16
    const int cbox = (row < 0) ? -1 : (row >> 6);   // which 64-channel box this row belongs to
    const int cc = row & 63;                        // channel inside the box
    const bool elected = (threadIdx.x == 64);
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    constexpr int kOutOff = Cfg::kAux ? 2 * SWP_BOX : 0;
    uint32_t box_counter = 0, aux_count = 0;
    int it = 0;
    // number of 64-channel boxes that hold real channels in co-tile c0: ceil(min(n_real - c0, 128) / 64)
    auto boxes_of = [&](int c0) { const int rem = args.n_real - c0; return (MM == 128 && rem > 64) ? 2 : (rem > 0 ? 1 : 0); };
    auto step_coords = [&](int tile, int step, int& b, int& yy, int& x0, int& ch0) {
      int y0, c0;
      tile_coords(tile, b, y0, x0, c0);
      const int nb = boxes_of(c0);
      yy = y0 + (step / nb) * 8;
      ch0 = c0 + (step % nb) * 64;
    };
    if constexpr (Cfg::kAux) {
      if (elected && blockIdx.x < num_tiles) {
        int b, yy, x0, ch0;
        step_coords(blockIdx.x, 0, b, yy, x0, ch0);
        mbar_arrive_expect_tx(aux_bar(0), SWP_BOX);
        tma_load_4d(epi_base, &maps.r, aux_bar(0), ch0, x0, yy, b);
      }
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int b, y0, x0, c0;
      tile_coords(tile, b, y0, x0, c0);
      const int acc = it & 1;
      const uint32_t taddr = tmem_base + lane_sel + uint32_t(acc * 256);
      const int next_tile = tile + gridDim.x;
      const int nb = boxes_of(c0);
      const int nsteps = 4 * nb;   // 4 groups of 8 image rows x nb channel boxes
      const float bias_v = s_bias[(c0 + (row < 0 ? 0 : row)) & 255];
      mbar_wait(tfull_bar(acc), (it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int step = 0; step < nsteps; ++step) {
        const int g = step / nb, cb = step % nb;
        const uint32_t ring = box_counter & 1u;
        const uint32_t out0 = epi_base + kOutOff + ring * SWP_BOX;
        if (elected) {
          tma_store_wait_read<1>();
          if constexpr (Cfg::kAux) {
            bool have = true;
            int nbb, nyy, nx0, nch0;
            if (step + 1 < nsteps) step_coords(tile, step + 1, nbb, nyy, nx0, nch0);
            else {
              have = next_tile < num_tiles;
              if (have) step_coords(next_tile, 0, nbb, nyy, nx0, nch0);
            }
            if (have) {
              const uint32_t nbuf = (aux_count + 1) & 1u;
              mbar_arrive_expect_tx(aux_bar(nbuf), SWP_BOX);
              tma_load_4d(epi_base + nbuf * SWP_BOX, &maps.r, aux_bar(nbuf), nch0, nx0, nyy, nbb);
            }
          }
        }
        named_bar_sync(1, GEMM_EPI_THREADS);
        uint32_t aux_addr = 0;
        if constexpr (Cfg::kAux) {
          const uint32_t ab = aux_count & 1u;
          mbar_wait(aux_bar(ab), (aux_count >> 1) & 1u);
          aux_addr = epi_base + ab * SWP_BOX;
        }
        const bool active = (cbox == cb);                     // this lane holds a channel of box cb
        // tcgen05.ld is warp-collective: the whole warp loads when any of its lanes can hold data of this box
        const bool warp_active = (MM == 128) ? ((q >> 1) == cb) : (args.c_split ? (q < 2) : true);
        uint32_t r[32];
        if (warp_active) {
          tmem_ld_x32(taddr + uint32_t(g * 64 + half * 32), r);
          tmem_ld_wait();
        }
        if (step == nsteps - 1) {   // last read of this accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (active) {
          // element (pixel j of the box, channel cc): byte offset j*128 + ((cc/8 ^ j%8) * 16) + (cc%8)*2
          const uint32_t col_off = uint32_t(cc & 7) * 2u;
          const uint32_t chunk = uint32_t(cc >> 3);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = half * 32 + i;
            const uint32_t off = uint32_t(j) * 128u + ((chunk ^ uint32_t(j & 7)) << 4) + col_off;
            float v = __uint_as_float(r[i]);
            if constexpr (EPI == CEPI_BIAS || EPI == CEPI_BIAS_LRELU || EPI == CEPI_BIAS_RES) v += bias_v;
            if constexpr (EPI == CEPI_BIAS_LRELU) v = round_bf16(v) > 0.f ? v : v * args.slope;
            if constexpr (Cfg::kAux) {
              uint16_t a16;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(a16) : "r"(aux_addr + off));
              const float a = __uint_as_float(uint32_t(a16) << 16);
              if constexpr (EPI == CEPI_BIAS_RES) v = round_bf16(v * args.alpha) + a;
              else v = (a > 0.f) ? v : v * args.slope;
            }
            const uint16_t o16 = uint16_t(pack_bf16(v, 0.f) & 0xFFFFu);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(out0 + off), "h"(o16) : "memory");
          }
        }
        fence_proxy_async();
        named_bar_sync(1, GEMM_EPI_THREADS);
        if (elected) {
          tma_store_4d(&maps.c[0], out0, c0 + cb * 64, x0, y0 + g * 8, b);
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
