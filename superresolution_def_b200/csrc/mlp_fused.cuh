// mlp_fused.cuh — the MLP half of a Swin / HAT block as ONE persistent tcgen05 kernel:
//
//     x_out = x_mid + drop * fc2( gelu( fc1(xn2) ) ),   xn_out = LayerNorm_next(x_out)            (forward)
//
// Replaces (reference): Mlp.forward fc1 -> GELU -> fc2 (models/architecture_swin.py:19-25), the residual add and the
// next LayerNorm (:149-150, :127), HAT's identical Mlp (models/hat_arch/hat_arch.py:76-96,306-307).
//
// Why one kernel: the unfused pair (gemm_tn<GELU2> + gemm_tn<RES_LN>) writes the hidden activation [T, Hp] to HBM
// and reads it straight back (2 x 403 MB of the 1.61 GB the pair moves at batch 16).  Here the hidden tile never leaves
// the SM on its way to fc2: per 128-token tile and 128-wide hidden chunk
//     GEMM1  acc1[128 x 128] = xn2_tile[128 x 192] . W1[chunk]^T          (TMEM, double-buffered)
//     epi-1  gelu / gelu' on tcgen05.ld fragments -> bf16 [128 x 64] swizzled boxes in shared memory
//     GEMM2  acc2[128 x 192] += act_box[128 x 64] . W2[:, box]^T            (the box is the A operand, straight from smem)
// and the boxes are TMA-stored to `act` / `dact` only because the backward needs them (training); inference stores
// nothing of the hidden tensor.  acc2 ends in the residual + LayerNorm row epilogue.
// Algorithmic HBM bytes per token (training): read xn2 + x_mid (2 x 384 B), write act + dact (2 x 2*Hp B) and
// x_out + xn_out (2 x 384 B): 4.6 KB at Hp = 768 vs 6.1 KB unfused; weights (2 x Hp x 192 bf16) stream from L2.
//
// Roles (576 threads, 1 CTA/SM, persistent over 128-token tiles):
//   warp 0      TMA producer: xn2 tile (3 boxes, single buffer) + weight ring (4 x 24 KB: W1 boxes [128 x 64],
//               W2 boxes [192 x 64]) in exactly the order the MMA warp consumes them
//   warp 1      MMA issuer, software-pipelined over a flat chunk index g:  G1(g), then G2(g-1)  — the tensor core runs
//               chunk g's first GEMM while the 16 epilogue warps turn chunk g-1 into bf16 boxes
//   warps 2-17  epilogue-1 per 64-column box (ring of 2 act + 2 dact boxes), then the row epilogue per tile
#pragma once
#include "gemm_tn.cuh"

namespace srk {

constexpr int MF_THREADS = 64 + 512;
constexpr int MF_EPI_THREADS = 512;
constexpr int MF_EPI_WARPS = 16;
constexpr int MF_C = 192;                 // Cp: K of GEMM1, N of GEMM2
constexpr int MF_CH = 128;                // hidden chunk (N of GEMM1)
constexpr int MF_WSTAGES = 4;
constexpr int MF_WSTAGE_BYTES = 192 * 128;  // one W2 box [192 x 64]; a W1 box [128 x 64] uses the first 16 KB
constexpr int MF_OFF_X = 0;
constexpr int MF_OFF_W = MF_OFF_X + 3 * BOX_BYTES;
constexpr int MF_OFF_ACT = MF_OFF_W + MF_WSTAGES * MF_WSTAGE_BYTES;
constexpr int MF_OFF_DACT = MF_OFF_ACT + 2 * BOX_BYTES;
constexpr int MF_OFF_BAR = MF_OFF_DACT + 2 * BOX_BYTES;
constexpr int MF_OFF_RED = MF_OFF_BAR + 256;
constexpr int MF_SMEM_BYTES = MF_OFF_RED + 2 * 4 * 128 * 4 + 1024;

struct MlpFwdArgs {
  int M;                 // tokens (multiple of 128)
  int Hp;                // padded hidden (multiple of 128)
  int n_real;            // real channels normalised by the LayerNorm epilogue (180 / 90)
  int hid_ones_col;      // column of act forced to 1.0 (bias-folding column of fc2), gelu' there = 0
  int ln_ones_col;       // column of xn_out forced to 1.0
  const float* gamma;    // next LayerNorm weight [n_real]
  const float* beta;     // next LayerNorm bias [n_real]
  float* stats;          // [M][2] mean, rstd of that LayerNorm (may be null)
  float eps;
  const float* row_scale;  // optional stochastic-depth factor per sample
  int rows_per_scale;
  const __nv_bfloat16* resid;  // x_mid [M, ld_res]
  int ld_res;
  __nv_bfloat16* x_out;        // [M, ld_xo]
  int ld_xo;
  int store_act, store_dact;   // training: 1, 1; inference: 0, 0
};

__global__ void __launch_bounds__(MF_THREADS, 1)
mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmAct,
               const __grid_constant__ CUtensorMap tmDact, const __grid_constant__ CUtensorMap tmXn,
               const MlpFwdArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + MF_OFF_BAR;
  auto wfull = [&](int s) { return bar_base + 8u * s; };
  auto wempty = [&](int s) { return bar_base + 8u * (MF_WSTAGES + s); };
  const uint32_t xfull = bar_base + 8u * (2 * MF_WSTAGES), xempty = xfull + 8u;
  auto a1full = [&](int b) { return bar_base + 8u * (2 * MF_WSTAGES + 2 + b); };
  auto a1empty = [&](int b) { return bar_base + 8u * (2 * MF_WSTAGES + 4 + b); };
  auto actfull = [&](int b) { return bar_base + 8u * (2 * MF_WSTAGES + 6 + b); };
  auto actempty = [&](int b) { return bar_base + 8u * (2 * MF_WSTAGES + 8 + b); };
  const uint32_t a2full = bar_base + 8u * (2 * MF_WSTAGES + 10), a2empty = a2full + 8u;
  const uint32_t tmem_slot = bar_base + 8u * (2 * MF_WSTAGES + 12);
  float* s_red = reinterpret_cast<float*>(smem_raw + (smem_base + MF_OFF_RED - smem_u32(smem_raw)));  // [2][4][128]

  __shared__ __align__(16) float s_gamma[MF_C];
  __shared__ __align__(16) float s_beta[MF_C];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = args.M / GEMM_BM;
  const int nch = args.Hp / MF_CH;                                   // hidden chunks per tile
  const int my_tiles = (m_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int G = my_tiles * nch;                                      // flat chunk count of this CTA

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmAct); tma_prefetch_desc(&tmDact); tma_prefetch_desc(&tmXn);
    for (int s = 0; s < MF_WSTAGES; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
    mbar_init(xfull, 1); mbar_init(xempty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(a1full(b), 1); mbar_init(a1empty(b), MF_EPI_WARPS);
      mbar_init(actfull(b), 1); mbar_init(actempty(b), 1);
    }
    mbar_init(a2full, 1); mbar_init(a2empty, MF_EPI_WARPS);
    fence_mbar_init();
  }
  pdl_launch_dependents();
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  pdl_wait();
  for (int i = threadIdx.x; i < MF_C; i += MF_THREADS) {
    s_gamma[i] = (i < args.n_real) ? args.gamma[i] : 0.f;
    s_beta[i] = (i < args.n_real) ? (args.beta != nullptr ? args.beta[i] : 0.f) : (i == args.ln_ones_col ? 1.f : 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tm_acc1 = tmem_base, tm_acc2 = tmem_base + 256u;   // acc1: 2 x 128 columns, acc2: 192 columns

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int g = 0; g <= G; ++g) {
        if (g < G) {
          const int i = g / nch, cc = g - i * nch;
          const int m0 = (int(blockIdx.x) + i * int(gridDim.x)) * GEMM_BM;
          if (cc == 0) {
            mbar_wait(xempty, (uint32_t(i) & 1u) ^ 1u);
            mbar_arrive_expect_tx(xfull, 3 * BOX_BYTES);
            for (int kb = 0; kb < 3; ++kb) tma_load_2d(smem_base + MF_OFF_X + kb * BOX_BYTES, &tmX, xfull, kb * 64, m0);
          }
          for (int kb = 0; kb < 3; ++kb) {   // W1 rows [cc*128, +128), K columns [kb*64, +64)
            mbar_wait(wempty(stage), phase ^ 1u);
            mbar_arrive_expect_tx(wfull(stage), BOX_BYTES);
            tma_load_2d(smem_base + MF_OFF_W + stage * MF_WSTAGE_BYTES, &tmW1, wfull(stage), kb * 64, cc * MF_CH);
            if (++stage == MF_WSTAGES) { stage = 0; phase ^= 1u; }
          }
        }
        if (g >= 1) {
          const int cc = (g - 1) % nch;
          for (int h = 0; h < 2; ++h) {      // W2 all 192 rows, K columns [(cc*2+h)*64, +64)
            mbar_wait(wempty(stage), phase ^ 1u);
            mbar_arrive_expect_tx(wfull(stage), MF_WSTAGE_BYTES);
            tma_load_2d(smem_base + MF_OFF_W + stage * MF_WSTAGE_BYTES, &tmW2, wfull(stage), (cc * 2 + h) * 64, 0);
            if (++stage == MF_WSTAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_bf16(GEMM_BM, MF_CH, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_bf16(GEMM_BM, MF_C, 0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int g = 0; g <= G; ++g) {
        if (g < G) {   // ---- GEMM1 of chunk g
          const int i = g / nch, cc = g - i * nch;
          if (cc == 0) { mbar_wait(xfull, uint32_t(i) & 1u); }
          mbar_wait(a1empty(g & 1), ((uint32_t(g) >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d = tm_acc1 + uint32_t((g & 1) * MF_CH);
          for (int kb = 0; kb < 3; ++kb) {
            mbar_wait(wfull(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + MF_OFF_X + kb * BOX_BYTES;
            const uint32_t sb = smem_base + MF_OFF_W + stage * MF_WSTAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc1,
                        (kb | k) != 0 ? 1u : 0u);
            umma_commit(wempty(stage));
            if (++stage == MF_WSTAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(a1full(g & 1));
          if (cc == nch - 1) umma_commit(xempty);   // the xn2 tile is dead once this tile's last GEMM1 has completed
        }
        if (g >= 1) {  // ---- GEMM2 of chunk g-1
          const int gp = g - 1;
          const int i = gp / nch, cc = gp - i * nch;
          for (int h = 0; h < 2; ++h) {
            const int b = gp * 2 + h, slot = b & 1;
            if (cc == 0 && h == 0) { mbar_wait(a2empty, (uint32_t(i) & 1u) ^ 1u); }
            mbar_wait(actfull(slot), (uint32_t(b) >> 1) & 1u);
            mbar_wait(wfull(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + MF_OFF_ACT + slot * BOX_BYTES;
            const uint32_t sb = smem_base + MF_OFF_W + stage * MF_WSTAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tm_acc2, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc2,
                        (cc | h | k) != 0 ? 1u : 0u);
            umma_commit(wempty(stage));
            umma_commit(actempty(slot));
            if (++stage == MF_WSTAGES) { stage = 0; phase ^= 1u; }
          }
          if (cc == nch - 1) umma_commit(a2full);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..17
    const int q = warp & 3;                 // TMEM lane quarter this warp can reach
    const int part = (warp - 2) >> 2;       // 0..3: column part
    const int row = q * 32 + lane;
    const bool elected = (threadIdx.x == 64);
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    const float inv_n = 1.0f / float(args.n_real);
    uint32_t b = 0;                         // global box counter of this CTA
    for (int i = 0; i < my_tiles; ++i) {
      const int m0 = (int(blockIdx.x) + i * int(gridDim.x)) * GEMM_BM;
      // ================= epilogue 1: hidden chunks -> bf16 boxes
      for (int cc = 0; cc < nch; ++cc) {
        const uint32_t g = uint32_t(i) * uint32_t(nch) + uint32_t(cc);
        mbar_wait(a1full(g & 1u), (g >> 1) & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h, ++b) {
          const uint32_t slot = b & 1u;
          const uint32_t act_s = smem_base + MF_OFF_ACT + slot * BOX_BYTES;
          const uint32_t dact_s = smem_base + MF_OFF_DACT + slot * BOX_BYTES;
          if (elected) {
            // boxes 0/1 of a tile overwrite the row epilogue's staging area: all earlier bulk stores must have been read
            if (cc == 0) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
          }
          mbar_wait(actempty(slot), ((b >> 1) & 1u) ^ 1u);   // GEMM2 has finished reading this slot's previous box
          named_bar_sync(1, MF_EPI_THREADS);
          uint32_t r[16];
          tmem_ld_x16(tm_acc1 + lane_sel + uint32_t((g & 1u) * MF_CH + h * 64 + part * 16), r);
          tmem_ld_wait();
          if (h == 1) {   // chunk fully drained into registers: hand acc1[g&1] back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a1empty(g & 1u));
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int ch = part * 2 + k;
            const uint32_t off = swz(row, ch);
            uint32_t a[4], d[4];   // same packed-fp16 GELU pair as the fc1 epilogue of gemm_tn (bit-identical outputs)
#pragma unroll
            for (int e = 0; e < 4; ++e) gelu_pair_h2(__uint_as_float(r[k * 8 + 2 * e]), __uint_as_float(r[k * 8 + 2 * e + 1]), a[e], d[e]);
            const int col0 = cc * MF_CH + h * 64 + ch * 8;
            if (args.hid_ones_col >= col0 && args.hid_ones_col < col0 + 8) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (col0 + 2 * e == args.hid_ones_col) { a[e] = (a[e] & 0xFFFF0000u) | 0x3F80u; d[e] &= 0xFFFF0000u; }
                if (col0 + 2 * e + 1 == args.hid_ones_col) { a[e] = (a[e] & 0x0000FFFFu) | 0x3F800000u; d[e] &= 0x0000FFFFu; }
              }
            }
            sts128(act_s + off, make_uint4(a[0], a[1], a[2], a[3]));
            if (args.store_dact) sts128(dact_s + off, make_uint4(d[0], d[1], d[2], d[3]));
          }
          fence_proxy_async();
          named_bar_sync(1, MF_EPI_THREADS);
          if (elected) {
            mbar_arrive(actfull(slot));
            if (args.store_act) tma_store_2d(&tmAct, act_s, cc * MF_CH + h * 64, m0);
            if (args.store_dact) tma_store_2d(&tmDact, dact_s, cc * MF_CH + h * 64, m0);
            tma_store_commit();
          }
        }
      }
      // ================= row epilogue: v = bf16(bf16(acc2) * rs + residual); x_out = v; xn_out = LayerNorm(v)
      // thread = (row, part): 48 columns [48*part, +48).  The residual and x_out move straight between registers and
      // global memory (96 contiguous bytes per thread); xn_out is staged through the (now idle) act/dact boxes.
      const int c0 = part * 48;
      uint4 res[6];
      {
        const uint4* rp = reinterpret_cast<const uint4*>(args.resid + size_t(m0 + row) * args.ld_res + c0);
#pragma unroll
        for (int k = 0; k < 6; ++k) res[k] = __ldg(rp + k);
      }
      const float rs = (args.row_scale != nullptr) ? args.row_scale[(m0 + row) / args.rows_per_scale] : 1.0f;
      mbar_wait(a2full, uint32_t(i) & 1u);
      tc_fence_after();
      uint32_t vp[24];   // packed v
      {
        uint32_t r0[32], r1[16];
        tmem_ld_x32(tm_acc2 + lane_sel + uint32_t(c0), r0);
        tmem_ld_x16(tm_acc2 + lane_sel + uint32_t(c0 + 32), r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2empty);
        const uint32_t rw[24] = {res[0].x, res[0].y, res[0].z, res[0].w, res[1].x, res[1].y, res[1].z, res[1].w,
                                 res[2].x, res[2].y, res[2].z, res[2].w, res[3].x, res[3].y, res[3].z, res[3].w,
                                 res[4].x, res[4].y, res[4].z, res[4].w, res[5].x, res[5].y, res[5].z, res[5].w};
#pragma unroll
        for (int k = 0; k < 24; ++k) {
          const float a0 = __uint_as_float(k < 16 ? r0[2 * k] : r1[2 * k - 32]);
          const float a1 = __uint_as_float(k < 16 ? r0[2 * k + 1] : r1[2 * k + 1 - 32]);
          const uint32_t a = pack_bf16(a0, a1);
          vp[k] = pack_bf16(fmaf(bf16_lo(a), rs, bf16_lo(rw[k])), fmaf(bf16_hi(a), rs, bf16_hi(rw[k])));
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 24; ++k) sum += bf16_lo(vp[k]) + bf16_hi(vp[k]);
      {
        uint4* op = reinterpret_cast<uint4*>(args.x_out + size_t(m0 + row) * args.ld_xo + c0);
#pragma unroll
        for (int k = 0; k < 6; ++k) op[k] = make_uint4(vp[4 * k], vp[4 * k + 1], vp[4 * k + 2], vp[4 * k + 3]);
      }
      s_red[(0 * 4 + part) * 128 + row] = sum;
      if (elected) tma_store_wait_read<0>();   // the act / dact boxes of this tile have left shared memory
      named_bar_sync(1, MF_EPI_THREADS);
      const float mean = (s_red[(0 * 4 + 0) * 128 + row] + s_red[(0 * 4 + 1) * 128 + row] + s_red[(0 * 4 + 2) * 128 + row] +
                          s_red[(0 * 4 + 3) * 128 + row]) * inv_n;
      float var = 0.f;
#pragma unroll
      for (int k = 0; k < 24; ++k) {
        const float d0 = bf16_lo(vp[k]) - mean, d1 = bf16_hi(vp[k]) - mean;
        var = fmaf(d0, d0, var);
        var = fmaf(d1, d1, var);
      }
      {  // pad columns (exact zeros in v) contributed (0 - mean)^2 each
        const int lo = c0 > args.n_real ? c0 : args.n_real;
        const int npad = (c0 + 48) > lo ? (c0 + 48) - lo : 0;
        var -= float(npad) * mean * mean;
      }
      s_red[(1 * 4 + part) * 128 + row] = var;
      named_bar_sync(1, MF_EPI_THREADS);
      var = s_red[(1 * 4 + 0) * 128 + row] + s_red[(1 * 4 + 1) * 128 + row] + s_red[(1 * 4 + 2) * 128 + row] +
            s_red[(1 * 4 + 3) * 128 + row];
      const float rstd = rsqrtf(fmaxf(var, 0.f) * inv_n + args.eps);
      if (part == 0 && args.stats != nullptr) reinterpret_cast<float2*>(args.stats)[m0 + row] = make_float2(mean, rstd);
      const uint32_t T1 = smem_base + MF_OFF_ACT;   // 3 boxes [128 x 64] over the act + dact slots
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int c = c0 + k * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(&s_gamma[c]), g1 = *reinterpret_cast<const float4*>(&s_gamma[c + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&s_beta[c]), b1 = *reinterpret_cast<const float4*>(&s_beta[c + 4]);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t pv = vp[k * 4 + e];
          o[e] = pack_bf16(fmaf((bf16_lo(pv) - mean) * rstd, gg[2 * e], bb[2 * e]),
                           fmaf((bf16_hi(pv) - mean) * rstd, gg[2 * e + 1], bb[2 * e + 1]));
        }
        sts128(T1 + (c >> 6) * BOX_BYTES + swz(row, (c & 63) >> 3), make_uint4(o[0], o[1], o[2], o[3]));
      }
      fence_proxy_async();
      named_bar_sync(1, MF_EPI_THREADS);
      if (elected) {
        for (int bx = 0; bx < 3; ++bx) tma_store_2d(&tmXn, T1 + bx * BOX_BYTES, bx * 64, m0);
        tma_store_commit();
      }
    }
    if (elected) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace srk
