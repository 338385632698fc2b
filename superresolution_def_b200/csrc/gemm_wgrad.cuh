// gemm_wgrad.cuh — weight-gradient GEMM  P[split][Ca][Cb] = sum_{tokens in split} A[t,ca] * B[t,cb]
// for token-major activations A:[T,Ca], B:[T,Cb] (bf16).  The contraction runs over the *strided*
// dimension of both operands, so both are fed to tcgen05.mma as MN-major 128B-swizzled tiles.
//
// Replaces (reference): autograd's weight/bias gradients of nn.Linear (qkv/proj/fc1/fc2,
// models/architecture_swin.py:73,94,19-25).  Bias gradients come for free: the B operand carries
// a constant-one column (see DESIGN.md "bias folding"), so one column of the result is sum_t A[t,:].
//
// grid = ca_groups * splits; each CTA owns AT (1 or 2) [128 x BNW] fp32 accumulators in TMEM -- AT * 128 channels of A -- and a
// token range.  With AT = 2 every B box that crosses the L2 -> SM crossbar feeds two accumulators: the Ca = 768 / 576 shapes
// were bound by that crossbar (A once + B once per 128-channel tile = 1.0 GB per launch at 11 TB/s), not by HBM.
#pragma once
#include "srk_ptx.cuh"

namespace srk {

constexpr int WG_THREADS = 192;
constexpr int WG_TOK = 64;                     // tokens per pipeline stage
constexpr int WG_SUBBOX = 64 * 128;            // one [64 tokens x 64 channels] swizzled box (8 KB)

struct WgradArgs {
  int T;        // tokens (multiple of 64); split s owns the 64-token blocks s, s + splits, s + 2 splits, ... of the I = T/64
                // blocks: all CTAs move through the tensors as one band, like the persistent GEMMs (adjacent CTAs read
                // adjacent 8 KB pieces at the same time: fc 87.4 -> 84.9 us, qkv 75.9 -> 72.6 us)
  int Ca, Cb;   // channel counts (Cb == BNW)
  int ca_groups; // ceil(Ca / (AT * 128))
  int splits;
  float* partials;  // [splits][ca_groups * AT * 128][BNW]
  // debug knobs (validated once on hardware, then fixed): descriptor LBO/SBO in bytes
  int lbo_bytes, sbo_bytes;
};

template <int BNW, int AT>
struct WgradCfg {
  static constexpr int kStageBytes = 2 * AT * WG_SUBBOX + (BNW / 64) * WG_SUBBOX;
  static constexpr int kStagesRaw = (232448 - 256 - 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 5 ? 5 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
  static_assert(kStages >= 3, "wgrad pipeline too shallow");
  static_assert(AT * BNW <= 512, "accumulators must fit the 512 TMEM columns");
};

template <int BNW, int AT>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradArgs args) {
  using Cfg = WgradCfg<BNW, AT>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  constexpr uint32_t kTmemCols = (AT * BNW <= 32) ? 32 : (AT * BNW <= 64) ? 64 : (AT * BNW <= 128) ? 128 : (AT * BNW <= 256) ? 256 : 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x % args.ca_groups;   // group of AT * 128 channels of A
  const int split = blockIdx.x / args.ca_groups;
  const int total_iters = args.T / WG_TOK;
  const int k_iters = (total_iters - split + args.splits - 1) / args.splits;  // >= 1 because splits <= total_iters

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
        const uint32_t sb = sa + 2 * AT * WG_SUBBOX;
        const int t0 = (kb * args.splits + split) * WG_TOK;
        mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
#pragma unroll
        for (int b = 0; b < 2 * AT; ++b)   // channels beyond Ca: TMA zero-fills (and still counts the bytes)
          tma_load_2d(sa + b * WG_SUBBOX, &tmA, full_bar(stage), a_tile * (AT * 128) + b * 64, t0);
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
        const uint32_t sb = sa + 2 * AT * WG_SUBBOX;
#pragma unroll
        for (int k = 0; k < WG_TOK / 16; ++k) {
          // 16 tokens = 16 rows of 128 B inside each [64 tok x 128 B] box
          const uint64_t bdesc = make_smem_desc(sb + k * 16 * 128, args.lbo_bytes, args.sbo_bytes);
#pragma unroll
          for (int t = 0; t < AT; ++t) {
            const uint64_t adesc = make_smem_desc(sa + t * 2 * WG_SUBBOX + k * 16 * 128, args.lbo_bytes, args.sbo_bytes);
            umma_bf16(tmem_base + uint32_t(t * BNW), adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
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
    // Accumulator row -> this thread's padded row of a staging tile in the (now idle) operand ring -> ONE bulk copy of the
    // BNW * 4 contiguous bytes of that row in the partials buffer.  (Row-per-thread 16-byte global stores touched 32
    // different 128-byte lines per instruction: ~14 us of exposed epilogue per [128 x 192] tile.)
    constexpr uint32_t kRowBytes = BNW * 4 + 16;   // +16: consecutive rows start 4 banks apart, 128-bit stores conflict-free
    static_assert(128 * kRowBytes <= Cfg::kStages * Cfg::kStageBytes, "staging tile must fit the operand ring");
#pragma unroll 1
    for (int t = 0; t < AT; ++t) {
      float* out = args.partials + (size_t(split) * args.ca_groups * (AT * 128) + size_t(a_tile) * (AT * 128) + t * 128 + row) * BNW;
      const uint32_t srow = smem_base + uint32_t(row) * kRowBytes;
      if (t > 0) tma_store_wait_read<0>();   // the previous tile's copy has left this row
#pragma unroll 1
      for (int c32 = 0; c32 < BNW / 32; ++c32) {
        uint32_t r[32];
        tmem_ld_x32(taddr + uint32_t(t * BNW + c32 * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sts128(srow + uint32_t(c32 * 128 + i * 16), make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
      }
      fence_proxy_async();
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out), "r"(srow), "n"(BNW * 4) : "memory");
      tma_store_commit();
    }
    tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Sum the per-split partials: out[i] = sum_s partials[s * split_stride + i], i < n_elems   (row-major [rows][Cb] per split)
static __global__ void wgrad_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out,
                                    int splits, int n_elems, size_t split_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_elems) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partials[size_t(k) * split_stride + i];
  out[i] = s;
}

}  // namespace srk
