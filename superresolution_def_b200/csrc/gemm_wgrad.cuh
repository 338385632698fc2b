// gemm_wgrad.cuh — weight-gradient GEMM  P[split][Ca][Cb] = sum_{tokens in split} A[t,ca] * B[t,cb]
// for token-major activations A:[T,Ca], B:[T,Cb] (bf16).  The contraction runs over the *strided*
// dimension of both operands, so both are fed to tcgen05.mma as MN-major 128B-swizzled tiles.
//
// Replaces (reference): autograd's weight/bias gradients of nn.Linear (qkv/proj/fc1/fc2,
// models/architecture_swin.py:73,94,19-25).  Bias gradients come for free: the B operand carries
// a constant-one column (see DESIGN.md "bias folding"), so one column of the result is sum_t A[t,:].
//
// grid = ca_tiles * splits; each CTA owns one [128 x BNW] fp32 accumulator in TMEM and a token range.
#pragma once
#include "srk_ptx.cuh"

namespace srk {

constexpr int WG_THREADS = 192;
constexpr int WG_TOK = 64;                     // tokens per pipeline stage
constexpr int WG_SUBBOX = 64 * 128;            // one [64 tokens x 64 channels] swizzled box (8 KB)

struct WgradArgs {
  int T;        // tokens (multiple of 64); split s owns k-iterations [s*I/splits, (s+1)*I/splits), I = T/64
  int Ca, Cb;   // channel counts (Cb == BNW)
  int ca_tiles; // ceil(Ca / 128)
  int splits;
  float* partials;  // [splits][ca_tiles*128][BNW]
  // debug knobs (validated once on hardware, then fixed): descriptor LBO/SBO in bytes
  int lbo_bytes, sbo_bytes;
};

template <int BNW>
struct WgradCfg {
  static constexpr int kStageBytes = 2 * WG_SUBBOX + (BNW / 64) * WG_SUBBOX;
  static constexpr int kStages = 5;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
};

template <int BNW>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradArgs args) {
  using Cfg = WgradCfg<BNW>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  constexpr uint32_t kTmemCols = (BNW <= 32) ? 32 : (BNW <= 64) ? 64 : (BNW <= 128) ? 128 : 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x % args.ca_tiles;
  const int split = blockIdx.x / args.ca_tiles;
  const int total_iters = args.T / WG_TOK;
  const int it_begin = int((long long)split * total_iters / args.splits);
  const int it_end = int((long long)(split + 1) * total_iters / args.splits);
  const int t_begin = it_begin * WG_TOK;
  const int k_iters = it_end - it_begin;  // >= 1 because splits <= total_iters

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  pdl_launch_dependents();
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < k_iters; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + 2 * WG_SUBBOX;
        const int t0 = t_begin + kb * WG_TOK;
        mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
        tma_load_2d(sa, &tmA, full_bar(stage), a_tile * 128, t0);
        tma_load_2d(sa + WG_SUBBOX, &tmA, full_bar(stage), a_tile * 128 + 64, t0);
#pragma unroll
        for (int b = 0; b < BNW / 64; ++b) tma_load_2d(sb + b * WG_SUBBOX, &tmB, full_bar(stage), b * 64, t0);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BNW, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < k_iters; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + 2 * WG_SUBBOX;
#pragma unroll
        for (int k = 0; k < WG_TOK / 16; ++k) {
          // 16 tokens = 16 rows of 128 B inside each [64 tok x 128 B] box
          const uint64_t adesc = make_smem_desc(sa + k * 16 * 128, args.lbo_bytes, args.sbo_bytes);
          const uint64_t bdesc = make_smem_desc(sb + k * 16 * 128, args.lbo_bytes, args.sbo_bytes);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
    float* out = args.partials + (size_t(split) * args.ca_tiles * 128 + size_t(a_tile) * 128 + row) * BNW;
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Sum the per-split partials: out[r][c] = sum_s partials[s][r][c]   (rows = ca_tiles*128, cols = Cb)
static __global__ void wgrad_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out,
                                    int splits, int n_elems) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elems) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partials[size_t(k) * n_elems + i];
  out[i] = s;
}

}  // namespace srk
