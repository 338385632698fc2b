// api_gemm.cu — C-ABI entry points for the tcgen05 GEMMs (forward/dgrad and wgrad).
#include <cstdlib>
#include "gemm_tn.cuh"
#include "gemm_wgrad.cuh"
#include "mlp_fused.cuh"
#include "srk_host.h"

namespace srk {

std::atomic<long long> g_launches{0};

template <int BN, int EPI>
static int launch_gemm_tn(const GemmArgs& a, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tC,
                          const CUtensorMap& tC2, const CUtensorMap& tX1, const CUtensorMap& tX2,
                          cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::kSmemBytes));
    configured.done();
  }
  const int tiles = (a.M / GEMM_BM) * (a.N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  GemmArgs args = a;
  // K <= 192: B ([BN x K] per N tile) stays resident in shared memory, the ring carries A only (SRK_GEMM_BRES=0: A/B switch)
  static const bool bres_enabled = [] { const char* e = getenv("SRK_GEMM_BRES"); return !(e && e[0] == '0'); }();
  args.b_resident = (bres_enabled && Cfg::kResOk && a.K <= 3 * GEMM_BK && grid >= a.N / BN) ? 1 : 0;
  SRK_CUDA_OK(launch_pdl(gemm_tn_kernel<BN, EPI>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, tA, tB, tC, tC2,
                         tX1, tX2, args));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

static int pick_bn(int epi, int N) {
  if (epi == EPI_RES_LN || epi == EPI_LNBWD) return (N == 192) ? 192 : (N == 128 ? 128 : -1);
  if (epi == EPI_MUL && N % 192 == 0) return 192;  // 3-stage operand ring + 4 multiplier boxes in flight fit shared memory
  if (epi == EPI_GELU2 || epi == EPI_MUL || epi == EPI_GELU1) return (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : -1);
  if (epi == EPI_MULG) return (N % 128 == 0) ? 128 : -1;  // two double-buffered accumulators: 4 * BN <= 512 TMEM columns
  if (N % 192 == 0) return 192;
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return -1;
}

}  // namespace srk

using namespace srk;

extern "C" const char* srk_version(void) { return "libsrk 0.1 (sm_100a: tcgen05/TMEM/TMA)"; }

extern "C" long long srk_launch_count(void) { return g_launches.load(); }

extern "C" int srk_gemm_grid(int M, int N) {
  (void)N;
  const int tiles = M / GEMM_BM;  // row epilogues have a single N tile
  return tiles < num_sms() ? tiles : num_sms();
}

static int gemm_tn_impl(int epi, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C,
                        int ldc, void* C2, int ldc2, const void* X1, int ldx1, const void* X2, int ldx2,
                        const SrkLnArgs* ln, float slope, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0 || M % GEMM_BM != 0 || K % GEMM_BK != 0)
    return fail(SRK_ERR_ARG, "srk_gemm_tn: M must be a multiple of 128 and K of 64");
  const int bn = pick_bn(epi, N);
  if (bn < 0) return fail(SRK_ERR_UNSUPPORTED, "srk_gemm_tn: unsupported N for this epilogue");
  if (!A || !B || !C) return fail(SRK_ERR_ARG, "srk_gemm_tn: null operand");
  GemmArgs a{};
  a.M = M; a.N = N; a.K = K;
  a.n_real = N; a.ones_col = -1; a.gamma = nullptr; a.beta = nullptr; a.stats = nullptr; a.partials = nullptr;
  a.eps = 1e-5f; a.row_scale = nullptr; a.rows_per_scale = 1; a.slope = slope;
  if (ln) {
    a.n_real = ln->n_real; a.ones_col = ln->ones_col; a.gamma = ln->gamma; a.beta = ln->beta;
    a.stats = ln->stats; a.partials = ln->partials; a.eps = ln->eps;
    a.row_scale = ln->row_scale; a.rows_per_scale = ln->rows_per_scale > 0 ? ln->rows_per_scale : 1;
  }
  a.x2 = static_cast<const __nv_bfloat16*>(X2); a.ldx2 = ldx2;
  if (epi == EPI_LNBWD && X2 && ((reinterpret_cast<uintptr_t>(X2) & 15) || ldx2 % 8)) return fail(SRK_ERR_ARG, "LNBWD: X2 rows must be 16-byte aligned");
  if ((epi == EPI_RES_LN || epi == EPI_LNBWD) && (!ln || !ln->gamma)) return fail(SRK_ERR_ARG, "LN epilogue needs SrkLnArgs");
  if (epi == EPI_LNBWD && (!ln->stats || !ln->partials || !X1 || !X2)) return fail(SRK_ERR_ARG, "LNBWD needs stats, partials, X1, X2");
  if (epi == EPI_RES_LN && (!X1 || !C2)) return fail(SRK_ERR_ARG, "RES_LN needs X1 and C2");
  if (epi == EPI_GELU2 && !C2) return fail(SRK_ERR_ARG, "GELU2 needs C2");
  if (epi == EPI_MUL && !X1) return fail(SRK_ERR_ARG, "MUL needs X1");
  if (epi == EPI_MULG && (!X1 || !X2)) return fail(SRK_ERR_ARG, "MULG needs the second GEMM's operands in X1 (A2 [M,K]) and X2 (B2 [N,K])");

  CUtensorMap tA, tB, tC, tC2, tX1, tX2;
  int rc;
  if ((rc = make_tmap_2d(&tA, A, M, K, lda, GEMM_BM))) return rc;
  if ((rc = make_tmap_2d(&tB, B, N, K, ldb, bn))) return rc;
  if ((rc = make_tmap_2d(&tC, C, M, N, ldc, GEMM_BM))) return rc;
  tC2 = tC; tX1 = tC; tX2 = tC;
  if (C2 && (rc = make_tmap_2d(&tC2, C2, M, N, ldc2, GEMM_BM))) return rc;
  if (epi == EPI_MULG) {
    if ((rc = make_tmap_2d(&tX1, X1, M, K, ldx1, GEMM_BM))) return rc;
    if ((rc = make_tmap_2d(&tX2, X2, N, K, ldx2, bn))) return rc;
  } else {
    if (X1 && (rc = make_tmap_2d(&tX1, X1, M, N, ldx1, GEMM_BM))) return rc;
    if (X2 && (rc = make_tmap_2d(&tX2, X2, M, N, ldx2, GEMM_BM))) return rc;
  }

#define SRK_CASE(BN_, EPI_) \
  if (bn == BN_ && epi == EPI_) return launch_gemm_tn<BN_, EPI_>(a, tA, tB, tC, tC2, tX1, tX2, stream);
  SRK_CASE(64, EPI_STORE)
  SRK_CASE(128, EPI_STORE)
  SRK_CASE(192, EPI_STORE)
  SRK_CASE(256, EPI_STORE)
  SRK_CASE(64, EPI_LRELU)
  SRK_CASE(128, EPI_LRELU)
  SRK_CASE(192, EPI_LRELU)
  SRK_CASE(256, EPI_LRELU)
  SRK_CASE(128, EPI_GELU2)
  SRK_CASE(256, EPI_GELU2)
  SRK_CASE(128, EPI_MUL)
  SRK_CASE(192, EPI_MUL)
  SRK_CASE(256, EPI_MUL)
  SRK_CASE(128, EPI_GELU1)
  SRK_CASE(256, EPI_GELU1)
  SRK_CASE(128, EPI_MULG)
  SRK_CASE(128, EPI_RES_LN)
  SRK_CASE(192, EPI_RES_LN)
  SRK_CASE(128, EPI_LNBWD)
  SRK_CASE(192, EPI_LNBWD)
#undef SRK_CASE
  return fail(SRK_ERR_UNSUPPORTED, "srk_gemm_tn: no kernel instance for (BN, epilogue)");
}

extern "C" int srk_gemm_tn(int epi, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C,
                           int ldc, void* C2, int ldc2, const void* X1, int ldx1, const void* X2, int ldx2,
                           const SrkLnArgs* ln, void* stream) {
  if (epi == EPI_LRELU) return fail(SRK_ERR_ARG, "srk_gemm_tn: SRK_EPI_LRELU is served by srk_gemm_tn_lrelu (it carries the slope)");
  return gemm_tn_impl(epi, M, N, K, A, lda, B, ldb, C, ldc, C2, ldc2, X1, ldx1, X2, ldx2, ln, 0.f, stream);
}

extern "C" int srk_gemm_tn_lrelu(int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                                 float slope, void* stream) {
  return gemm_tn_impl(EPI_LRELU, M, N, K, A, lda, B, ldb, C, ldc, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, slope, stream);
}

template <int BNW, int AT>
static int launch_wgrad(const WgradArgs& a, const CUtensorMap& tA, const CUtensorMap& tB, cudaStream_t stream) {
  using Cfg = WgradCfg<BNW, AT>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(gemm_wgrad_kernel<BNW, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::kSmemBytes));
    configured.done();
  }
  SRK_CUDA_OK(launch_pdl(gemm_wgrad_kernel<BNW, AT>, dim3(a.ca_groups * a.splits), dim3(WG_THREADS), Cfg::kSmemBytes, stream, tA,
                         tB, a));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

// channels of A per CTA: two 128-row accumulators whenever A has more than 128 channels (one B box feeds both)
// channels of A per CTA.  Two 128-row accumulators per CTA (every B box feeds both) were measured SLOWER at the block's
// shapes (fc 95 vs 87 us, qkv 80 vs 76 us, proj 53 vs 45 us: same tensor time, one pipeline stage less), so one tile per
// CTA is the default and SRK_WGRAD_AT=2 the switch that reproduces the measurement.
static int wgrad_at(int Ca) {
  static const int forced = getenv("SRK_WGRAD_AT") ? atoi(getenv("SRK_WGRAD_AT")) : 0;
  return (forced == 2 && Ca > 128) ? 2 : 1;
}

// MN-major, 128B swizzle: LBO = distance between 64-channel groups (one [64 tok x 128 B] box),
// SBO = distance between 8-token groups (8 rows x 128 B); validated on hardware in round 1.
static int gemm_wgrad_impl(int T, int Ca, int Cb, const void* A, int lda, const void* B, int ldb,
                           float* workspace, int splits, float* out, int lbo_bytes, int sbo_bytes,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (splits <= 0 || T <= 0 || T % WG_TOK != 0 || splits > T / WG_TOK)
    return fail(SRK_ERR_ARG, "srk_gemm_wgrad: T must be a multiple of 64 and 1 <= splits <= T/64");
  if (!A || !B || !workspace) return fail(SRK_ERR_ARG, "srk_gemm_wgrad: null pointer");
  const int at = wgrad_at(Ca);
  WgradArgs a{};
  a.T = T; a.Ca = Ca; a.Cb = Cb; a.ca_groups = (Ca + at * 128 - 1) / (at * 128); a.splits = splits; a.partials = workspace;
  a.lbo_bytes = lbo_bytes; a.sbo_bytes = sbo_bytes;
  CUtensorMap tA, tB;
  int rc;
  if ((rc = make_tmap_2d(&tA, A, T, Ca, lda, WG_TOK))) return rc;
  if ((rc = make_tmap_2d(&tB, B, T, Cb, ldb, WG_TOK))) return rc;
#define SRK_WG(CB_) \
  case CB_: rc = (at == 2) ? launch_wgrad<CB_, 2>(a, tA, tB, stream) : launch_wgrad<CB_, 1>(a, tA, tB, stream); break;
  switch (Cb) {
    SRK_WG(64)
    SRK_WG(128)
    SRK_WG(192)
    SRK_WG(256)
    default: return fail(SRK_ERR_UNSUPPORTED, "srk_gemm_wgrad: Cb must be 64/128/192/256");
  }
#undef SRK_WG
  if (rc) return rc;
  if (out == nullptr) return SRK_OK;   // internal callers that sum the per-split partials themselves (block backward)
  const int n = ((Ca + 127) / 128) * 128 * Cb;
  const size_t split_stride = size_t(a.ca_groups) * at * 128 * Cb;
  SRK_CUDA_OK(launch_pdl(wgrad_reduce_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, workspace, out, splits, n,
                         split_stride));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_gemm_wgrad_splits(int T, int Ca) {
  const int at = wgrad_at(Ca);
  const int groups = (Ca + at * 128 - 1) / (at * 128);
  int s = num_sms() / groups;
  if (s > T / WG_TOK) s = T / WG_TOK;
  return s < 1 ? 1 : s;
}

extern "C" long long srk_gemm_wgrad_workspace_elems(int Ca, int Cb, int splits) {
  const int at = wgrad_at(Ca);
  const int groups = (Ca + at * 128 - 1) / (at * 128);
  return (long long)splits * groups * at * 128 * Cb;
}

extern "C" int srk_gemm_wgrad(int T, int Ca, int Cb, const void* A, int lda, const void* B, int ldb,
                              float* workspace, int splits, float* out, void* stream) {
  if (!out) return fail(SRK_ERR_ARG, "srk_gemm_wgrad: null pointer");
  return gemm_wgrad_impl(T, Ca, Cb, A, lda, B, ldb, workspace, splits, out, WG_SUBBOX, 1024, stream);
}

namespace srk {
// Per-split partials only: [splits][rows_per_split][Cb] fp32 in `workspace`, rows_per_split = srk_gemm_wgrad_workspace_elems
// / (splits * Cb); the caller folds the splits (the block backward does it inside its gradient-unpack kernel).
int gemm_wgrad_partials(int T, int Ca, int Cb, const void* A, int lda, const void* B, int ldb, float* workspace, int splits,
                        void* stream) {
  return gemm_wgrad_impl(T, Ca, Cb, A, lda, B, ldb, workspace, splits, nullptr, WG_SUBBOX, 1024, stream);
}
}  // namespace srk

extern "C" int srk_mlp_fwd(int T, int Cp, int Hp, const void* xn2, const void* w1, const void* w2, const void* resid,
                           void* act, void* dact, void* x_out, void* xn_out, int hid_ones_col, const SrkLnArgs* ln,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (T <= 0 || T % GEMM_BM != 0 || Cp != MF_C || Hp <= 0 || Hp % MF_CH != 0)
    return fail(SRK_ERR_UNSUPPORTED, "srk_mlp_fwd: T % 128 == 0, Cp == 192, Hp % 128 == 0");
  if (!xn2 || !w1 || !w2 || !resid || !x_out || !xn_out || !ln || !ln->gamma) return fail(SRK_ERR_ARG, "srk_mlp_fwd: null argument");
  if (dact && !act) return fail(SRK_ERR_ARG, "srk_mlp_fwd: dact without act");
  MlpFwdArgs a{};
  a.M = T; a.Hp = Hp; a.n_real = ln->n_real; a.hid_ones_col = hid_ones_col; a.ln_ones_col = ln->ones_col;
  a.gamma = ln->gamma; a.beta = ln->beta; a.stats = ln->stats; a.eps = ln->eps;
  a.row_scale = ln->row_scale; a.rows_per_scale = ln->rows_per_scale > 0 ? ln->rows_per_scale : 1;
  a.resid = static_cast<const __nv_bfloat16*>(resid); a.ld_res = Cp;
  a.x_out = static_cast<__nv_bfloat16*>(x_out); a.ld_xo = Cp;
  a.store_act = act ? 1 : 0; a.store_dact = dact ? 1 : 0;
  CUtensorMap tX, tW1, tW2, tAct, tDact, tXn;
  int rc;
  if ((rc = make_tmap_2d(&tX, xn2, T, Cp, Cp, GEMM_BM))) return rc;
  if ((rc = make_tmap_2d(&tW1, w1, Hp, Cp, Cp, MF_CH))) return rc;
  if ((rc = make_tmap_2d(&tW2, w2, Cp, Hp, Hp, MF_C))) return rc;
  if ((rc = make_tmap_2d(&tXn, xn_out, T, Cp, Cp, GEMM_BM))) return rc;
  tAct = tXn; tDact = tXn;
  if (act && (rc = make_tmap_2d(&tAct, act, T, Hp, Hp, GEMM_BM))) return rc;
  if (dact && (rc = make_tmap_2d(&tDact, dact, T, Hp, Hp, GEMM_BM))) return rc;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MF_SMEM_BYTES));
    configured.done();
  }
  const int tiles = T / GEMM_BM;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  SRK_CUDA_OK(launch_pdl(mlp_fwd_kernel, dim3(grid), dim3(MF_THREADS), MF_SMEM_BYTES, stream, tX, tW1, tW2, tAct, tDact, tXn, a));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}
