// gemm_tn.cuh — persistent warp-specialised tcgen05 GEMM for the token-major linears of the
// Swin/HAT blocks:  C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  bf16 operands, fp32 accumulate in TMEM.
//
// Replaces (reference): nn.Linear calls in WindowAttention.qkv/proj (models/architecture_swin.py:73,94),
// Mlp.fc1/fc2 (:19-25), and their autograd input-gradients; the fused epilogues replace
// LayerNorm (:127,150), GELU (:20), the residual adds (:149-150) and their backward passes.
//
// Roles (1 CTA/SM, persistent over output tiles):
//   warp 0   : TMA producer   (A [128 x 64] + B [BN x 64] bf16 boxes, 128B swizzle, mbarrier ring)
//   warp 1   : MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16; accumulators double-buffered in TMEM)
//   row epilogues (RES_LN, LNBWD; 320 threads): warps 2-9: tcgen05.ld -> registers -> math -> swizzled smem -> TMA store.
//   elementwise ("box") epilogues (STORE, GELU*, MUL*; 640 threads): warp 2 = store warp (one lane issues the TMA store
//              of every finished 64-column box and hands the staging slot back), warp 3 = aux warp (MUL: one lane streams
//              the multiplier boxes through a ring of kAuxSlots buffers, several boxes ahead of the epilogue), warps
//              4-19 = 16 epilogue warps.  The epilogue warps never synchronise with each other: every hand-over is an
//              mbarrier (staging slot full / empty, aux box full / empty), so a warp only ever waits for data.
//   A warp reaches TMEM lanes 32*(warp%4)..+31 only, so 2 (or 4) warps share each lane quarter and split the columns of
//   every 64-column box.
#pragma once
#include <cuda_fp16.h>
#include "srk_ptx.cuh"

namespace srk {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int GEMM_EPI_THREADS = 256;
constexpr int BOX_BYTES = 128 * 128;  // one [128 rows x 64 bf16] swizzled box

enum GemmEpilogue : int {
  EPI_STORE = 0,   // C = bf16(acc)
  EPI_GELU2 = 1,   // C = gelu(u) (bf16), C2 = gelu'(u) stored as FP16 (11 significant bits, no conversion)   (fc1 forward)
  EPI_MUL = 2,     // C = bf16(bf16(acc) * X1), X1 = the FP16 gelu' tensor GELU2 stored   (fc2 dgrad -> dU)
  EPI_RES_LN = 3,  // C = v = bf16(bf16(acc) + X1); C2 = LayerNorm(v)           (proj / fc2 forward)
  EPI_LNBWD = 4,   // C = X2 + LayerNormBackward(acc | X1, stats)               (fc1 / qkv dgrad)
  EPI_GELU1 = 5,   // C = gelu(u), u = acc                 (fc1 forward when the backward recomputes gelu', see MULG)
  EPI_MULG = 6,    // two GEMMs of the same shape per tile: acc0 = A B^T, acc1 = A2 B2^T (A2 -> tmX1, B2 -> tmX2);
                   // C = bf16(bf16(acc0) * bf16(gelu'(acc1)))   (fc2 dgrad -> dU with u = xn2 W1^T recomputed on the
                   // tensor cores instead of reading a stored gelu'(u): the block is HBM-bound, the MMA pipe is idle)
  EPI_LRELU = 7,   // C = bf16(leaky_relu(acc, slope))    (4x4 stride-2 convolutions of UNetDiscriminatorSN as GEMMs over
                   // gathered patches: models/discriminator_swin.py:10-11; slope = GemmArgs::slope)
};

struct GemmArgs {
  int M, N, K;        // padded sizes: M % 128 == 0, N % BN == 0, K % 64 == 0
  int n_real;         // number of real channels normalised by the LN epilogues (e.g. 180)
  int ones_col;       // column forced to 1.0 in the LN output / GELU output (bias-folding column), -1: none
  const float* gamma; // LN weight [n_real]
  const float* beta;  // LN bias   [n_real]
  float* stats;       // [M][2] (mean, rstd): written by EPI_RES_LN, read by EPI_LNBWD
  float* partials;    // EPI_LNBWD: [gridDim.x][2][BN] per-CTA column sums (dgamma, dbeta)
  float eps;
  const float* row_scale;  // EPI_RES_LN, optional: per-sample stochastic-depth factor applied to bf16(acc)
  int rows_per_scale;
  const __nv_bfloat16* x2;  // EPI_LNBWD: the residual-gradient tensor X2 [M, ldx2] (read straight from global memory)
  int ldx2;
  float slope;        // EPI_LRELU: negative slope
  int b_resident;     // K <= 192: every CTA keeps ONE N tile of B ([BN x K], loaded once) in shared memory and walks M tiles
                      // only; the operand ring then holds A boxes alone (twice to six times as many bytes of A in flight)
};

template <int BN, int EPI>
struct GemmCfg {
  static constexpr bool kBoxEpi = (EPI == EPI_STORE || EPI == EPI_GELU2 || EPI == EPI_MUL || EPI == EPI_GELU1 || EPI == EPI_MULG ||
                                    EPI == EPI_LRELU);
  static constexpr int kAccs = (EPI == EPI_MULG) ? 2 : 1;  // accumulators per tile (each double-buffered in TMEM)
  static constexpr int kEpiWarps = (kBoxEpi || EPI == EPI_LNBWD) ? 16 : 8;
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kFirstEpiWarp = kBoxEpi ? 4 : 2;   // box epilogues: warp 2 = store warp, warp 3 = aux warp
  static constexpr int kThreads = 32 * kFirstEpiWarp + kEpiThreads;
  static constexpr int kOutPerBox = (EPI == EPI_GELU2) ? 2 : 1;   // boxes written per 64-column step
  static constexpr int kOutSlots = 2;                             // staging ring (slot = kOutPerBox boxes)
  static constexpr int kAuxSlots = (EPI == EPI_MUL) ? 4 : 0;      // multiplier boxes in flight (HBM latency x bandwidth)
  static constexpr int kParts = kEpiWarps / 4;        // warps sharing one TMEM lane quarter
  static constexpr int kColsPerPart = 64 / kParts;    // columns of a 64-column box handled by one warp
  static constexpr int kStageBytes = GEMM_BM * 128 + BN * 128;
  static constexpr int kBoxes = BN / 64;
  static constexpr int kEpiBytes = kBoxEpi ? (kAuxSlots + kOutSlots * kOutPerBox) * BOX_BYTES : 2 * kBoxes * BOX_BYTES;
  static constexpr int kRedBytes = 2 * 4 * 128 * 4;  // cross-part row reductions: [2][parts <= 4][128 rows] floats
  static constexpr int kBudget = 232448 - 1024 /*align slack*/ - 512 /*barriers*/ - 2304 /*static smem*/ - kRedBytes;
  static constexpr int kStagesRaw = (kBudget - kEpiBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 512 + kRedBytes + 1024;
  // B-resident mode (K <= 192): [B: 3 boxes of BN x 128 B][A ring][epilogue buffers] inside the same allocation
  static constexpr int kResBBytes = 3 * BN * 128;
  static constexpr int kResStagesRaw = (kStages * kStageBytes - kResBBytes) / (GEMM_BM * 128);
  static constexpr int kResStages = kResStagesRaw > 8 ? 8 : kResStagesRaw;
  static constexpr bool kResOk = (kAccs == 1) && kResStages >= 2;
  static_assert(kStages >= 2, "not enough shared memory for a 2-stage pipeline");
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64, <= 256");
  static_assert(2 * kAccs * BN <= 512, "double-buffered accumulators must fit the 512 TMEM columns");
};

// GELU (exact-erf semantics, nn.GELU default) and its derivative with ONE MUFU op per element:
//   erf(u/sqrt2) = tanh(u * P(u^2)),  P(t) = a0 + a1 t + a2 t^2   (least-squares fit of atanh(erf) on |u| <= 5;
//   |gelu - exact| <= 4.9e-5, |gelu' - exact| <= 1.3e-4 before the hardware tanh.approx error of 2^-11 relative,
//   i.e. <= 2.5e-4 in Phi — an order of magnitude below the bf16 resolution of the stored outputs).
//   gelu(u) = u * Phi,  Phi = 0.5 + 0.5 tanh(w),  gelu'(u) = Phi + u * 0.5 (1 - tanh^2 w) * d(w)/du.
// The fit polynomial turns over beyond |u| ~ 11, so the tanh argument is evaluated at clamp(u, +-8) (tanh(14.7) = 1).
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_pair(float u, float& a, float& g) {
  constexpr float a0 = 7.97703653e-01f, a1 = 3.68205808e-02f, a2 = -3.20923304e-04f;
  const float uc = fminf(fmaxf(u, -8.0f), 8.0f);
  const float t = uc * uc;
  const float p = fmaf(fmaf(a2, t, a1), t, a0);                            // P(t)
  const float dp = fmaf(fmaf(2.5f * a2, t, 1.5f * a1), t, 0.5f * a0);      // 0.5 * d(u P(u^2))/du
  const float th = fast_tanh(uc * p);
  const float cdf = fmaf(0.5f, th, 0.5f);
  a = u * cdf;
  g = fmaf(uc * fmaf(-th, th, 1.0f), dp, cdf);
}

// Two elements per instruction: the same fit evaluated in packed fp16 (HFMA2 / one tanh.approx.f16x2 per pair).  fp16 carries
// 11 significant bits, the results are stored as bf16 (8 bits): |gelu - fp32 path| <= 5e-4 absolute on |u| <= 8, below the
// bf16 spacing of the stored value wherever the value exceeds 0.25 and below 2.5e-4 in Phi elsewhere (the bound the fp32
// path already has from tanh.approx).  u^2 is clamped instead of u (t = min(u^2, 64)): beyond |u| = 8 the tanh argument is
// u * P(64) = 1.84 u >= 14.7, i.e. tanh = +-1, Phi in {0, 1}, gelu' = Phi exactly.
// Returns gelu(u) as a packed bf16x2 word (a GEMM operand) and gelu'(u) as the packed fp16x2 word it was computed in.
__device__ __forceinline__ void gelu_pair_h2(float x0, float x1, uint32_t& a_bf, uint32_t& g_h2) {
  const __half2 a0 = __float2half2_rn(7.97703653e-01f), a1 = __float2half2_rn(3.68205808e-02f), a2 = __float2half2_rn(-3.20923304e-04f);
  const __half2 d0 = __float2half2_rn(0.5f * 7.97703653e-01f), d1 = __float2half2_rn(1.5f * 3.68205808e-02f), d2 = __float2half2_rn(2.5f * -3.20923304e-04f);
  const __half2 half_ = __float2half2_rn(0.5f), one_ = __float2half2_rn(1.0f);
  const __half2 u = __floats2half2_rn(x0, x1);
  const __half2 t = __hmin2(__hmul2(u, u), __float2half2_rn(64.0f));
  const __half2 p = __hfma2(__hfma2(a2, t, a1), t, a0);
  const __half2 dp = __hfma2(__hfma2(d2, t, d1), t, d0);
  const __half2 w = __hmul2(u, p);
  uint32_t thw;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(thw) : "r"(*reinterpret_cast<const uint32_t*>(&w)));
  const __half2 th = *reinterpret_cast<const __half2*>(&thw);
  const __half2 cdf = __hfma2(half_, th, half_);
  const __half2 a = __hmul2(u, cdf);
  const __half2 s1 = __hfma2(__hneg2(th), th, one_);
  const __half2 g = __hfma2(__hmul2(u, s1), dp, cdf);
  const float2 af = __half22float2(a);
  a_bf = pack_bf16(af.x, af.y);
  g_h2 = *reinterpret_cast<const uint32_t*>(&g);   // gelu' is only ever read back by EPI_MUL: it stays fp16
}

// Transposing butterfly: on entry lane l holds v[0..31] (32 columns of its row); on exit v[0] of
// lane l is the sum over the 32 lanes of column l.  31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Same for 16 columns: on exit v[0] of lane l is the sum over the 32 lanes of column (l >> 1) (both lanes of a pair hold it).
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 16; off >= 2; off >>= 1) {
    const int half = off >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// byte offset of 16-byte chunk `ch` (0..7) of row `r` inside a 128B-swizzled [rows x 128 B] box
__device__ __forceinline__ uint32_t swz(int r, int ch) { return uint32_t(r) * 128u + (uint32_t(ch ^ (r & 7)) << 4); }

template <int BN, int EPI>
__global__ void __launch_bounds__((GemmCfg<BN, EPI>::kThreads), 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
               const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
               const GemmArgs args) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr int NBOX = Cfg::kBoxes;
  const bool bres = Cfg::kResOk && args.b_resident != 0;
  const int S = bres ? Cfg::kResStages : Cfg::kStages;                  // operand ring depth
  const uint32_t stage_bytes = bres ? uint32_t(GEMM_BM * 128) : uint32_t(Cfg::kStageBytes);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring_base = smem_base + (bres ? uint32_t(Cfg::kResBBytes) : 0u);   // resident B sits in front of the ring
  const uint32_t epi_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  // barrier slots (8 B each, 64 slots)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };              // 0..7
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };       // 8..15
  auto tfull_bar = [&](int a) { return bar_base + 8u * (16 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (18 + a); };
  auto aux_bar = [&](int b) { return bar_base + 8u * (20 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 22;
  const uint32_t bres_bar = bar_base + 8u * 23;
  // box epilogues: staging-slot and aux-box hand-over barriers
  auto ofull_bar = [&](int s_) { return bar_base + 8u * (24 + s_); };
  auto oempty_bar = [&](int s_) { return bar_base + 8u * (28 + s_); };
  auto afull_bar = [&](int s_) { return bar_base + 8u * (32 + s_); };
  auto aempty_bar = [&](int s_) { return bar_base + 8u * (40 + s_); };
  const uint32_t red_base = bar_base + 512;

  __shared__ __align__(16) float s_gamma[256];
  __shared__ __align__(16) float s_beta[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = args.M / GEMM_BM;
  const int n_tiles = args.N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_iters = args.K / GEMM_BK;
  // Tile walk: tile = blockIdx.x + i * tile_step, (m, n) = (tile / n_tiles, tile % n_tiles).  Resident mode keeps n fixed
  // per CTA: the CTAs with the same n (every n_tiles-th one) share the M tiles among themselves.
  const int tile_step = bres ? ((int(gridDim.x) - int(blockIdx.x) % n_tiles + n_tiles - 1) / n_tiles) * n_tiles : int(gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < 8; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(bres_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), Cfg::kEpiWarps);
      mbar_init(aux_bar(a), 1);
    }
    if constexpr (Cfg::kBoxEpi) {
      for (int a = 0; a < Cfg::kOutSlots; ++a) {
        mbar_init(ofull_bar(a), Cfg::kEpiWarps);
        mbar_init(oempty_bar(a), 1);
      }
      for (int a = 0; a < Cfg::kAuxSlots; ++a) {
        mbar_init(afull_bar(a), 1);
        mbar_init(aempty_bar(a), Cfg::kEpiWarps);
      }
    }
    fence_mbar_init();
  }
  pdl_launch_dependents();
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_wait();  // everything above overlaps the previous kernel's tail; nothing below may run before it has finished
  if constexpr (EPI == EPI_RES_LN || EPI == EPI_LNBWD) {
    for (int i = threadIdx.x; i < 256; i += Cfg::kThreads) {
      s_gamma[i] = (i < args.n_real) ? args.gamma[i] : 0.f;
      s_beta[i] = (i < args.n_real) ? (args.beta != nullptr ? args.beta[i] : 0.f) : (i == args.ones_col ? 1.f : 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (bres) {   // this CTA's N tile of B: loaded once
        const int n0 = (int(blockIdx.x) % n_tiles) * BN;
        mbar_arrive_expect_tx(bres_bar, uint32_t(k_iters) * BN * 128);
        for (int kb = 0; kb < k_iters; ++kb) tma_load_2d(smem_base + kb * (BN * 128), &tmB, bres_bar, kb * GEMM_BK, n0);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile / n_tiles) * GEMM_BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < Cfg::kAccs * k_iters; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = ring_base + stage * stage_bytes;
          const uint32_t sb = sa + GEMM_BM * 128;
          mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
          if (Cfg::kAccs == 2 && kb >= k_iters) {  // second GEMM of the tile: operands A2 / B2
            tma_load_2d(sa, &tmX1, full_bar(stage), (kb - k_iters) * GEMM_BK, m0);
            tma_load_2d(sb, &tmX2, full_bar(stage), (kb - k_iters) * GEMM_BK, n0);
          } else {
            tma_load_2d(sa, &tmA, full_bar(stage), kb * GEMM_BK, m0);
            if (!bres) tma_load_2d(sb, &tmB, full_bar(stage), kb * GEMM_BK, n0);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (bres) mbar_wait(bres_bar, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < Cfg::kAccs * k_iters; ++kb) {
          const int g2 = (Cfg::kAccs == 2 && kb >= k_iters) ? 1 : 0;
          const int kk = kb - g2 * k_iters;
          const uint32_t d_tmem = tmem_base + uint32_t((acc * Cfg::kAccs + g2) * BN);
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = ring_base + stage * stage_bytes;
          const uint32_t sb = bres ? smem_base + kb * (BN * 128) : sa + GEMM_BM * 128;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kk | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else if (Cfg::kBoxEpi && warp == 2) {
    // ------------------------------------------------------------------ store warp (box epilogues)
    if (lane == 0) {
      uint32_t g = 0;   // global box counter: the epilogue warps walk the same sequence
      for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile / n_tiles) * GEMM_BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int j = 0; j < NBOX; ++j, ++g) {
          const uint32_t slot = g % Cfg::kOutSlots;
          const uint32_t src = epi_base + (Cfg::kAuxSlots + slot * Cfg::kOutPerBox) * BOX_BYTES;
          mbar_wait(ofull_bar(slot), (g / Cfg::kOutSlots) & 1u);
          tma_store_2d(&tmC, src, n0 + j * 64, m0);
          if constexpr (EPI == EPI_GELU2) tma_store_2d(&tmC2, src + BOX_BYTES, n0 + j * 64, m0);
          tma_store_commit();
          tma_store_wait_read<0>();          // the slot has left shared memory: hand it back
          mbar_arrive(oempty_bar(slot));
        }
      }
      tma_store_wait_all<0>();
    }
  } else if (Cfg::kBoxEpi && warp == 3) {
    // ------------------------------------------------------------------ aux warp (MUL: multiplier boxes, kAuxSlots deep)
    if constexpr (EPI == EPI_MUL) {
      if (lane == 0) {
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
          const int m0 = (tile / n_tiles) * GEMM_BM;
          const int n0 = (tile % n_tiles) * BN;
          for (int j = 0; j < NBOX; ++j, ++g) {
            const uint32_t slot = g % Cfg::kAuxSlots;
            mbar_wait(aempty_bar(slot), ((g / Cfg::kAuxSlots) & 1u) ^ 1u);
            mbar_arrive_expect_tx(afull_bar(slot), BOX_BYTES);
            tma_load_2d(epi_base + slot * BOX_BYTES, &tmX1, afull_bar(slot), n0 + j * 64, m0);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Two (or four) warps share each TMEM lane quarter (hardware: a warp reaches lanes 32*(warp%4)..+31) and split the
    // columns between them ("half"), so every SM sub-partition hosts two (four) epilogue warps.
    const int q = warp & 3;
    const int half = (warp - Cfg::kFirstEpiWarp) >> 2;   // column part of this warp inside every 64-column box
    const int row = q * 32 + lane;      // accumulator row owned by this thread (shared with the other parts)
    const bool elected = (threadIdx.x == 32 * Cfg::kFirstEpiWarp);
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    int it = 0;
    uint32_t box_counter = 0;           // box-granular staging ring position
    uint32_t aux_count = 0;             // number of aux loads consumed (parity tracking)
    constexpr int NC = BN / 32;         // 32-column chunks per row
    constexpr int NCH = NC / 2;         // RES_LN: chunks per half (needs an even chunk count)
    constexpr int PC = BN / Cfg::kParts;   // LNBWD: columns per thread (its part of the row), walked in 16-column pieces
    constexpr int NPC = PC / 16;
    static_assert(EPI != EPI_LNBWD || (PC % 16 == 0), "LNBWD: BN / 4 must be a multiple of 16");
    float acc_g[(EPI == EPI_LNBWD) ? NPC : 1];
    float acc_b[(EPI == EPI_LNBWD) ? NPC : 1];
#pragma unroll
    for (int i = 0; i < ((EPI == EPI_LNBWD) ? NPC : 1); ++i) { acc_g[i] = 0.f; acc_b[i] = 0.f; }
    // cross-half row reductions (row epilogues): red[k][half][row]
    float* s_red = reinterpret_cast<float*>(smem_raw + (red_base - smem_u32(smem_raw)));

    // aux prefetch for the first tile (row epilogues)
    if constexpr (EPI == EPI_RES_LN || EPI == EPI_LNBWD) {
      if (elected && blockIdx.x < num_tiles) {
        const int m0 = (blockIdx.x / n_tiles) * GEMM_BM, n0 = (blockIdx.x % n_tiles) * BN;
        mbar_arrive_expect_tx(aux_bar(0), NBOX * BOX_BYTES);
        for (int b = 0; b < NBOX; ++b) tma_load_2d(epi_base + b * BOX_BYTES, &tmX1, aux_bar(0), n0 + b * 64, m0);
      }
    }

    for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step, ++it) {
      const int m0 = (tile / n_tiles) * GEMM_BM;
      const int n0 = (tile % n_tiles) * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const uint32_t taddr = tmem_base + lane_sel + uint32_t(acc * Cfg::kAccs * BN);
      const int next_tile = tile + tile_step;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();

      if constexpr (Cfg::kBoxEpi) {
        // ---------------------------------------------------------- box-granular elementwise epilogues
        // Per 64-column box: accumulator columns (prefetched from TMEM during the previous box) -> math -> this warp's
        // 16 columns of the staging slot -> arrive on the slot's "full" barrier (the store warp issues the TMA store).
        constexpr int CPP = Cfg::kColsPerPart;
        uint32_t rb[2][CPP];
        uint32_t rub[2][(EPI == EPI_MULG) ? CPP : 1];
        tmem_ld_cols(taddr + uint32_t(half * CPP), rb[0]);
        if constexpr (EPI == EPI_MULG) tmem_ld_cols(taddr + uint32_t(BN + half * CPP), rub[0]);
#pragma unroll
        for (int j = 0; j < NBOX; ++j) {
          uint32_t (&r)[CPP] = rb[j & 1];
          uint32_t (&ru)[(EPI == EPI_MULG) ? CPP : 1] = rub[j & 1];
          tmem_ld_wait();
          if (j + 1 < NBOX) {   // next box's columns travel while this box is computed
            tmem_ld_cols(taddr + uint32_t((j + 1) * 64 + half * CPP), rb[(j + 1) & 1]);
            if constexpr (EPI == EPI_MULG) tmem_ld_cols(taddr + uint32_t(BN + (j + 1) * 64 + half * CPP), rub[(j + 1) & 1]);
          } else {              // accumulator fully drained into registers: hand TMEM back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          const uint32_t slot = box_counter % Cfg::kOutSlots;
          const uint32_t out0 = epi_base + (Cfg::kAuxSlots + slot * Cfg::kOutPerBox) * BOX_BYTES;
          uint32_t aux_addr = 0;
          if constexpr (EPI == EPI_MUL) {
            const uint32_t as = box_counter % Cfg::kAuxSlots;
            mbar_wait(afull_bar(as), (box_counter / Cfg::kAuxSlots) & 1u);
            aux_addr = epi_base + as * BOX_BYTES;
          }
          mbar_wait(oempty_bar(slot), ((box_counter / Cfg::kOutSlots) & 1u) ^ 1u);
#pragma unroll
          for (int i = 0; i < CPP / 8; ++i) {
            const int ch = half * (CPP / 8) + i;
            const uint32_t off = swz(row, ch);
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[i * 8 + e]);
            if constexpr (EPI == EPI_STORE) {
              sts128(out0 + off, make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                            pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7])));
            } else if constexpr (EPI == EPI_LRELU) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * args.slope;
              sts128(out0 + off, make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                            pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7])));
            } else if constexpr (EPI == EPI_MUL) {
              const uint4 g = lds128(aux_addr + off);
              const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 gf = __half22float2(*reinterpret_cast<const __half2*>(&gw[e]));
                o[e] = pack_bf16(round_bf16(v[2 * e]) * gf.x, round_bf16(v[2 * e + 1]) * gf.y);
              }
              sts128(out0 + off, make_uint4(o[0], o[1], o[2], o[3]));
            } else if constexpr (EPI == EPI_MULG) {
              const int col0 = n0 + j * 64 + ch * 8;
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float a_, g_;
                gelu_pair(__uint_as_float(ru[i * 8 + e]), a_, g_);   // same fp32 accumulator value the forward saw
                o[e] = round_bf16(v[e]) * round_bf16(g_);
              }
              if (args.ones_col >= col0 && args.ones_col < col0 + 8) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (col0 + e == args.ones_col) o[e] = 0.0f;   // the forward's constant 1.0 column has no gradient
              }
              sts128(out0 + off, make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]),
                                            pack_bf16(o[6], o[7])));
            } else if constexpr (EPI == EPI_GELU1) {
              float a[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { float g_; gelu_pair(v[e], a[e], g_); }
              const int col0 = n0 + j * 64 + ch * 8;
              if (args.ones_col >= col0 && args.ones_col < col0 + 8) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (col0 + e == args.ones_col) a[e] = 1.0f;
              }
              sts128(out0 + off, make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]),
                                            pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7])));
            } else {  // EPI_GELU2
              uint32_t a[4], g[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) gelu_pair_h2(v[2 * e], v[2 * e + 1], a[e], g[e]);
              const unsigned rel = unsigned(args.ones_col - (n0 + j * 64 + ch * 8));   // warp-uniform: one test per 8 columns
              if (rel < 8u) {   // constant 1.0 column (bias folding): act = 1, gelu' = 0
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (rel == unsigned(2 * e)) { a[e] = (a[e] & 0xFFFF0000u) | 0x3F80u; g[e] &= 0xFFFF0000u; }
                  if (rel == unsigned(2 * e + 1)) { a[e] = (a[e] & 0x0000FFFFu) | 0x3F800000u; g[e] &= 0x0000FFFFu; }
                }
              }
              sts128(out0 + off, make_uint4(a[0], a[1], a[2], a[3]));
              sts128(out0 + BOX_BYTES + off, make_uint4(g[0], g[1], g[2], g[3]));
            }
          }
          fence_proxy_async();   // this thread's staging stores -> visible to the TMA store (async proxy)
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(ofull_bar(slot));
            if constexpr (EPI == EPI_MUL) mbar_arrive(aempty_bar(box_counter % Cfg::kAuxSlots));
          }
          ++box_counter;
        }
      } else {
        // ---------------------------------------------------------- full-row epilogues (BN covers the row)
        // A thread owns one accumulator row (tcgen05.ld 32x32b) and half of its columns; the other half lives in the
        // warp 4 positions up, so row reductions cross warps through s_red.  Pad columns (>= n_real) hold exact zeros
        // in the accumulator and in every aux tile (zero weight rows / zero pad activations), which lets the hot loops
        // run without per-element column tests.
        const uint32_t T0 = epi_base;                      // RES_LN: residual -> v ; LNBWD: x (LN input)
        const uint32_t T1 = epi_base + NBOX * BOX_BYTES;   // RES_LN: LN output   ; LNBWD: dres -> out
        const float inv_n = 1.0f / float(args.n_real);
        const int c_begin = half * NCH * 32, c_end = c_begin + NCH * 32;  // this thread's column range
        if constexpr (EPI == EPI_RES_LN) {
          mbar_wait(aux_bar(0), aux_count & 1u);
          ++aux_count;
        }
        if constexpr (EPI == EPI_RES_LN) {
          const float rs = (args.row_scale != nullptr) ? args.row_scale[(m0 + row) / args.rows_per_scale] : 1.0f;
          uint32_t vp[NCH * 16];  // this thread's half row of v = bf16(bf16(acc) * rs + residual), packed pairs
          float sum = 0.f;
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) {
            const int c32 = half * NCH + ci;
            uint32_t r[32];
            tmem_ld_x32(taddr + uint32_t(c32 * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = c32 * 32 + i * 8;
              const uint32_t addr = T0 + (c >> 6) * BOX_BYTES + swz(row, (c & 63) >> 3);
              const uint4 rv = lds128(addr);
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t a = pack_bf16(__uint_as_float(r[i * 8 + 2 * e]), __uint_as_float(r[i * 8 + 2 * e + 1]));
                const uint32_t pv = pack_bf16(fmaf(bf16_lo(a), rs, bf16_lo(rw[e])), fmaf(bf16_hi(a), rs, bf16_hi(rw[e])));
                vp[ci * 16 + i * 4 + e] = pv;
                sum += bf16_lo(pv) + bf16_hi(pv);
              }
              sts128(addr, make_uint4(vp[ci * 16 + i * 4], vp[ci * 16 + i * 4 + 1], vp[ci * 16 + i * 4 + 2],
                                      vp[ci * 16 + i * 4 + 3]));
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          s_red[(0 * 2 + half) * 128 + row] = sum;
          fence_proxy_async();
          named_bar_sync(1, Cfg::kEpiThreads);
          if (elected) {  // v (= new residual stream) is complete in T0; nobody reads T0 again (v stays in registers)
            for (int b = 0; b < NBOX; ++b) tma_store_2d(&tmC, T0 + b * BOX_BYTES, n0 + b * 64, m0);
            tma_store_commit();
            tma_store_wait_read<0>();  // v and the previous tile's LN output have left shared memory
            if (next_tile < num_tiles) {  // prefetch the next tile's residual under this tile's passes 2 and 3
              const int nm0 = (next_tile / n_tiles) * GEMM_BM, nn0 = (next_tile % n_tiles) * BN;
              mbar_arrive_expect_tx(aux_bar(0), NBOX * BOX_BYTES);
              for (int b = 0; b < NBOX; ++b) tma_load_2d(T0 + b * BOX_BYTES, &tmX1, aux_bar(0), nn0 + b * 64, nm0);
            }
          }
          const float mean = (s_red[(0 * 2 + 0) * 128 + row] + s_red[(0 * 2 + 1) * 128 + row]) * inv_n;
          float var = 0.f;
#pragma unroll
          for (int k = 0; k < NCH * 16; ++k) {
            const float d0 = bf16_lo(vp[k]) - mean, d1 = bf16_hi(vp[k]) - mean;
            var = fmaf(d0, d0, var);
            var = fmaf(d1, d1, var);
          }
          {  // pad columns contributed (0 - mean)^2 each
            const int lo = c_begin > args.n_real ? c_begin : args.n_real;
            const int npad = c_end > lo ? c_end - lo : 0;
            var -= float(npad) * mean * mean;
          }
          s_red[(1 * 2 + half) * 128 + row] = var;
          named_bar_sync(1, Cfg::kEpiThreads);
          var = s_red[(1 * 2 + 0) * 128 + row] + s_red[(1 * 2 + 1) * 128 + row];
          const float rstd = rsqrtf(fmaxf(var, 0.f) * inv_n + args.eps);
          if (half == 0 && args.stats != nullptr)
            reinterpret_cast<float2*>(args.stats)[m0 + row] = make_float2(mean, rstd);
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = (half * NCH + ci) * 32 + i * 8;
              const uint32_t boff = (c >> 6) * BOX_BYTES + swz(row, (c & 63) >> 3);
              // pad columns: s_gamma = 0, s_beta = (column == ones_col) -> the affine below yields exactly 0 / 1 there
              const float4 g0 = *reinterpret_cast<const float4*>(&s_gamma[c]), g1 = *reinterpret_cast<const float4*>(&s_gamma[c + 4]);
              const float4 b0 = *reinterpret_cast<const float4*>(&s_beta[c]), b1 = *reinterpret_cast<const float4*>(&s_beta[c + 4]);
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t pv = vp[ci * 16 + i * 4 + e];
                const float y0 = fmaf((bf16_lo(pv) - mean) * rstd, gg[2 * e], bb[2 * e]);
                const float y1 = fmaf((bf16_hi(pv) - mean) * rstd, gg[2 * e + 1], bb[2 * e + 1]);
                o[e] = pack_bf16(y0, y1);
              }
              sts128(T1 + boff, make_uint4(o[0], o[1], o[2], o[3]));
            }
          }
          fence_proxy_async();
          named_bar_sync(1, Cfg::kEpiThreads);
          if (elected) {
            for (int b = 0; b < NBOX; ++b) tma_store_2d(&tmC2, T1 + b * BOX_BYTES, n0 + b * 64, m0);
            tma_store_commit();
          }
        } else {  // EPI_LNBWD: 16 warps, a thread owns one accumulator row and PC = BN / 4 of its columns
          // Shared-memory plan: two [128 x BN] buffers.  x (the LayerNorm input) of tile i sits in buffer i & 1 and the
          // result is written over it IN PLACE (same thread, same bytes), so the other buffer is free for the whole
          // epilogue: the next tile's x is requested at the very start of this one (a full epilogue of lead time -- with a
          // single x buffer the warps spent 31 % of their time waiting for this load, ncu source view).  The residual
          // gradient X2 is only added element by element in pass 2: it comes straight from global memory into registers.
          const uint32_t Tx = epi_base + uint32_t(it & 1) * (NBOX * BOX_BYTES);
          if (elected) {
            tma_store_wait_read<0>();   // the previous tile's result has left the other buffer
            if (next_tile < num_tiles) {
              const int nm0 = (next_tile / n_tiles) * GEMM_BM, nn0 = (next_tile % n_tiles) * BN;
              const uint32_t Tn = epi_base + uint32_t((it + 1) & 1) * (NBOX * BOX_BYTES);
              mbar_arrive_expect_tx(aux_bar((it + 1) & 1), NBOX * BOX_BYTES);
              for (int b = 0; b < NBOX; ++b) tma_load_2d(Tn + b * BOX_BYTES, &tmX1, aux_bar((it + 1) & 1), nn0 + b * 64, nm0);
            }
          }
          if (half == 0 && next_tile < num_tiles) {   // next tile's X2 row and row statistics -> L2 (they are read from
            const int nm0 = (next_tile / n_tiles) * GEMM_BM, nn0 = (next_tile % n_tiles) * BN;   // global memory at its start)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(args.x2 + size_t(nm0 + row) * args.ldx2 + nn0), "n"(BN * 2) : "memory");
            if (lane == 0)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(args.stats + size_t(nm0 + q * 32) * 2), "n"(32 * 8) : "memory");
          }
          const int cbase = half * PC;                   // `half` is this warp's column part (0..3) here
          const bool warp_has_pad = __shfl_sync(0xffffffffu, int(cbase + PC > args.n_real), 0) != 0;   // warp-uniform: only the
          uint4 dres[2 * NPC];                                                                        // last part masks pads                           // this thread's PC columns of X2 (consumed in pass 2)
          {
            const uint4* gp = reinterpret_cast<const uint4*>(args.x2 + size_t(m0 + row) * args.ldx2 + n0 + cbase);
#pragma unroll
            for (int i = 0; i < 2 * NPC; ++i) dres[i] = __ldg(gp + i);
          }
          const float2 st = reinterpret_cast<const float2*>(args.stats)[m0 + row];
          const float rstd = st.y, nmr = -st.x * st.y;   // xhat = x * rstd + nmr
          mbar_wait(aux_bar(it & 1), (uint32_t(it) >> 1) & 1u);
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int pc = 0; pc < NPC; ++pc) {
            const int c16 = cbase + pc * 16;
            uint32_t r[16];
            tmem_ld_x16(taddr + uint32_t(c16), r);
            tmem_ld_wait();
            float pg[16], pb[16];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int c = c16 + i * 8;
              const uint4 xv = lds128(Tx + (c >> 6) * BOX_BYTES + swz(row, (c & 63) >> 3));
              const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
              const float4 g0 = *reinterpret_cast<const float4*>(&s_gamma[c]), g1 = *reinterpret_cast<const float4*>(&s_gamma[c + 4]);
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t a = pack_bf16(__uint_as_float(r[i * 8 + 2 * e]), __uint_as_float(r[i * 8 + 2 * e + 1]));
                const float d0 = bf16_lo(a), d1 = bf16_hi(a);   // dxn = bf16(acc); exact 0 in pad columns
                const float h0 = fmaf(bf16_lo(xw[e]), rstd, nmr), h1 = fmaf(bf16_hi(xw[e]), rstd, nmr);   // xhat
                const float q0 = d0 * gg[2 * e], q1 = d1 * gg[2 * e + 1];
                s1 += q0 + q1;
                s2 = fmaf(q0, h0, fmaf(q1, h1, s2));
                pg[i * 8 + 2 * e] = d0 * h0; pg[i * 8 + 2 * e + 1] = d1 * h1;
                pb[i * 8 + 2 * e] = d0;      pb[i * 8 + 2 * e + 1] = d1;
              }
            }
            acc_g[pc] += warp_colsum16(pg, lane);
            acc_b[pc] += warp_colsum16(pb, lane);
          }
          s_red[(0 * 4 + half) * 128 + row] = s1;
          s_red[(1 * 4 + half) * 128 + row] = s2;
          named_bar_sync(1, Cfg::kEpiThreads);
          // dx = rstd * (dxn * gamma - c1 - xhat * c2), with rstd folded into the row constants
          const float c1r = (s_red[(0 * 4 + 0) * 128 + row] + s_red[(0 * 4 + 1) * 128 + row] + s_red[(0 * 4 + 2) * 128 + row] +
                             s_red[(0 * 4 + 3) * 128 + row]) * inv_n * rstd;
          const float c2r = (s_red[(1 * 4 + 0) * 128 + row] + s_red[(1 * 4 + 1) * 128 + row] + s_red[(1 * 4 + 2) * 128 + row] +
                             s_red[(1 * 4 + 3) * 128 + row]) * inv_n * rstd;
#pragma unroll
          for (int pc = 0; pc < NPC; ++pc) {
            const int c16 = cbase + pc * 16;
            uint32_t r[16];
            tmem_ld_x16(taddr + uint32_t(c16), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int c = c16 + i * 8;
              const uint32_t boff = (c >> 6) * BOX_BYTES + swz(row, (c & 63) >> 3);
              const uint4 xv = lds128(Tx + boff);
              const uint4 dv = dres[pc * 2 + i];
              const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
              const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
              const float4 g0 = *reinterpret_cast<const float4*>(&s_gamma[c]), g1 = *reinterpret_cast<const float4*>(&s_gamma[c + 4]);
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t a = pack_bf16(__uint_as_float(r[i * 8 + 2 * e]), __uint_as_float(r[i * 8 + 2 * e + 1]));
                const float h0 = fmaf(bf16_lo(xw[e]), rstd, nmr), h1 = fmaf(bf16_hi(xw[e]), rstd, nmr);
                const float x0 = fmaf(-h0, c2r, fmaf(bf16_lo(a) * rstd, gg[2 * e], -c1r));
                const float x1 = fmaf(-h1, c2r, fmaf(bf16_hi(a) * rstd, gg[2 * e + 1], -c1r));
                const uint32_t dx = pack_bf16(x0, x1);   // LayerNorm input gradient, rounded like the reference's bf16 tensor
                o[e] = pack_bf16(bf16_lo(dw[e]) + bf16_lo(dx), bf16_hi(dw[e]) + bf16_hi(dx));
              }
              if (warp_has_pad && c + 8 > args.n_real) {  // group touches pad columns: they must stay exactly zero
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (c + 2 * e >= args.n_real) o[e] &= 0xFFFF0000u;
                  if (c + 2 * e + 1 >= args.n_real) o[e] &= 0x0000FFFFu;
                }
              }
              sts128(Tx + boff, make_uint4(o[0], o[1], o[2], o[3]));   // in place: this thread just read exactly these bytes
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          fence_proxy_async();
          named_bar_sync(1, Cfg::kEpiThreads);
          if (elected) {
            for (int b = 0; b < NBOX; ++b) tma_store_2d(&tmC, Tx + b * BOX_BYTES, n0 + b * 64, m0);
            tma_store_commit();
          }
        }
      }
    }
    if constexpr (EPI == EPI_LNBWD) {
      // per-CTA column sums: 8 warps -> smem (re-using the idle tile buffers) -> one row of partials per CTA
      float* s_part = reinterpret_cast<float*>(smem_raw + (epi_base - smem_u32(smem_raw)));  // [4][2][BN]
      if (elected) tma_store_wait_read<0>();   // the last tile's result may still be leaving the buffer s_part aliases
      named_bar_sync(1, Cfg::kEpiThreads);
      if ((lane & 1) == 0) {   // warp_colsum16: lanes 2k and 2k+1 both hold column k of the piece
#pragma unroll
        for (int k = 0; k < NPC; ++k) {
          s_part[(q * 2 + 0) * BN + half * PC + k * 16 + (lane >> 1)] = acc_g[k];
          s_part[(q * 2 + 1) * BN + half * PC + k * 16 + (lane >> 1)] = acc_b[k];
        }
      }
      named_bar_sync(1, Cfg::kEpiThreads);
      const int e = threadIdx.x - 32 * Cfg::kFirstEpiWarp;
      for (int i = e; i < 2 * BN; i += Cfg::kEpiThreads) {
        const int w = i / BN, c = i % BN;
        args.partials[(size_t(blockIdx.x) * 2 + w) * BN + c] =
            s_part[(0 * 2 + w) * BN + c] + s_part[(1 * 2 + w) * BN + c] + s_part[(2 * 2 + w) * BN + c] +
            s_part[(3 * 2 + w) * BN + c];
      }
    }
    if (!Cfg::kBoxEpi && elected) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace srk
