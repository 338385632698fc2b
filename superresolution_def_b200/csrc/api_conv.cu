// api_conv.cu — C-ABI entry points for the convolutional paths (implicit-GEMM conv3x3 fwd/dgrad/wgrad with
// fused bias / LeakyReLU / residual / PixelShuffle / GELU epilogues, and the 1-channel head/tail convolutions).
#include "conv3x3_swap.cuh"
#include <cstdlib>
#include "conv_aux.cuh"
#include "srk_host.h"

using namespace srk;

namespace {

// maps for an NHWC tensor [B,H,W,Cp]; when `ps`, the memory is the pixel-shuffled tensor [B,2H,2W,Cp/4] and the
// four maps view its sub-pixel lattices (i,j) with 64 channels each.
int make_act_maps(CUtensorMap* m4, const void* ptr, int B, int H, int W, int Cp, int ps, int bw, int bh) {
  if (!ps) return make_tmap_nhwc(&m4[0], ptr, Cp, W, H, B, Cp, (uint64_t)W * Cp, (uint64_t)H * W * Cp, bw, bh);
  if (Cp != 256) return fail(SRK_ERR_UNSUPPORTED, "pixel-shuffle views need 256 conv channels (4 x 64)");
  const uint64_t Cg = 64, W2 = 2 * (uint64_t)W, H2 = 2 * (uint64_t)H;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      const char* base = static_cast<const char*>(ptr) + ((uint64_t)i * W2 + j) * Cg * 2;
      int rc = make_tmap_nhwc(&m4[i * 2 + j], base, Cg, W, H, B, 2 * Cg, 2 * W2 * Cg, H2 * W2 * Cg, bw, bh);
      if (rc) return rc;
    }
  return SRK_OK;
}

template <int BN, int EPI>
int launch_conv(const ConvMaps& maps, const ConvArgs& a, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, EPI>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(conv3x3_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::kSmemBytes));
    configured.done();
  }
  const int tiles = a.B * (a.H / CONV_TH) * (a.W / CONV_TW) * (a.Cout_p / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv3x3_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(maps, a);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

template <int BN, int EPI>
int launch_conv_halo(const ConvMaps& maps, const ConvArgs& a, cudaStream_t stream) {
  using Cfg = HaloCfg<BN, EPI>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(conv3x3_halo_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::kSmemBytes));
    configured.done();
  }
  const int tiles = a.B * (a.H / HALO_TH) * (a.W / HALO_TW) * (a.Cout_p / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv3x3_halo_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(maps, a);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

// SRK_SWAP_M64: "off" = always the M = 128 instance; default = M = 64 for layers with <= 64 output channels
int swap_m64_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("SRK_SWAP_M64");
    mode = !e ? 1 : (e[0] == 'o' ? 0 : 1);
  }
  return mode;
}

template <int EPI, int MM>
int launch_conv_swap(const ConvMaps& maps, const ConvArgs& a, cudaStream_t stream) {
  using Cfg = SwapCfg<EPI, MM>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(conv3x3_swap_kernel<EPI, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured.done();
  }
  const int tiles = a.B * (a.H / SWP_TH) * (a.W / SWP_TW) * ((a.n_real + MM - 1) / MM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv3x3_swap_kernel<EPI, MM><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(maps, a);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

// 0: per-tap boxes (conv3x3_kernel); 1 (default): halo-resident tiles where the shape allows (H % 16 == 0, W % 8 == 0)
int conv_halo_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("SRK_CONV_HALO");
    mode = e ? std::atoi(e) : 1;
  }
  return mode;
}

template <int BNW>
int launch_conv_wgrad(const ConvWgradMaps& maps, const ConvWgradArgs& a, cudaStream_t stream) {
  using Cfg = WgradCfg<BNW, 1>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(conv3x3_wgrad_kernel<BNW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::kSmemBytes));
    configured.done();
  }
  conv3x3_wgrad_kernel<BNW><<<a.co_tiles * 9 * a.splits, WG_THREADS, Cfg::kSmemBytes, stream>>>(maps, a);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

int conv_wgrad_splits(int B, int H, int W, int co_tiles) {
  int s = num_sms() / (co_tiles * 9);
  const int iters = B * (H / 4) * (W / 16);
  if (s > iters) s = iters;
  return s < 1 ? 1 : s;
}

}  // namespace

extern "C" int srk_conv3x3_prep_weights(const float* w, const float* bias, int Cout, int Cin, int Cout_p, int Cin_p,
                                        int ps, void* wf, void* wt, float* bias_packed, void* stream_) {
  if ((Cout_p % 64 && Cout_p != 16) || Cin_p % 64 || Cout > Cout_p || Cin > Cin_p) return fail(SRK_ERR_ARG, "conv prep: bad padding");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  conv_prep_weights_kernel<<<num_sms() * 2, 256, 0, stream>>>(w, bias, static_cast<__nv_bfloat16*>(wf),
                                                              static_cast<__nv_bfloat16*>(wt), bias_packed, Cout, Cin,
                                                              Cout_p, Cin_p, ps);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_conv3x3_igemm(int epi, int B, int H, int W, int Cin_p, int Cout_p, int n_real, const void* x,
                                 int x_ps, const void* wk, const float* bias, float slope, void* y, int y_ps,
                                 void* y2, const void* r, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (H % CONV_TH || W % CONV_TW) return fail(SRK_ERR_UNSUPPORTED, "conv3x3: H % 8 == 0 and W % 16 == 0 required");
  const bool out1 = (epi == CEPI_OUT1);
  if (Cin_p % 64 || (Cout_p % 64 && !(out1 && Cout_p == 16)) || Cout_p > 256 || Cin_p > 256)
    return fail(SRK_ERR_UNSUPPORTED, "conv3x3: channels must be multiples of 64, <= 256");
  if (!x || !wk || !y) return fail(SRK_ERR_ARG, "conv3x3: null pointer");
  if (out1 && (x_ps || y_ps || !bias)) return fail(SRK_ERR_ARG, "conv3x3: OUT1 needs plain layouts and a bias");
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  {
    // Few output channels (<= 64 real ones here): role-swapped kernel, as in the view entry point below -- the 256 -> 64
    // input gradients of the two upsampler stages (pixel-shuffled gradient, one strided map per 64-channel chunk) and
    // conv_before_upsample are UMMA-instruction-rate bound as [pixels x 64] tiles (627 us at 256^2 x 16 for 309 GFLOP).
    const int hmode = conv_halo_mode();
    const bool swap_aux = (epi == CEPI_BIAS_RES || epi == CEPI_MASK_LRELU);
    const bool swap_ok = !out1 && !y_ps && !y2 && n_real > 0 && n_real <= 64 && Cout_p == 64 && H % SWP_TH == 0 && W % SWP_TW == 0 &&
                         (epi == CEPI_BIAS || epi == CEPI_BIAS_LRELU || swap_aux) && (!swap_aux || Cin_p >= 128);
    if (hmode != 0 && hmode != 5 && swap_ok) {
      if (swap_aux && !r) return fail(SRK_ERR_ARG, "conv3x3: epilogue needs an aux tensor");
      if ((rc = make_act_maps(maps.a, x, B, H, W, Cin_p, x_ps, SWP_BW, SWP_BH))) return rc;
      if (!x_ps) for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
      if ((rc = make_tmap_nhwc(&maps.c[0], y, Cout_p, W, H, B, Cout_p, (uint64_t)W * Cout_p, (uint64_t)H * W * Cout_p, 8, 8))) return rc;
      for (int i = 1; i < 4; ++i) maps.c[i] = maps.c[0];
      maps.c2 = maps.c[0];
      maps.r = maps.c[0];
      if (swap_aux && (rc = make_tmap_nhwc(&maps.r, r, Cout_p, W, H, B, Cout_p, (uint64_t)W * Cout_p, (uint64_t)H * W * Cout_p, 8, 8))) return rc;
      const bool use64 = swap_m64_mode() != 0;
      if ((rc = make_tmap_2d(&maps.w, wk, Cout_p, 9 * (uint64_t)Cin_p, 9 * (uint64_t)Cin_p, use64 ? 64 : 128))) return rc;
      ConvArgs a{};
      a.B = B; a.H = H; a.W = W; a.Cin_p = Cin_p; a.Cout_p = Cout_p; a.n_real = n_real; a.bias = bias; a.slope = slope;
      a.alpha = 1.0f; a.a_split = x_ps; a.c_split = 0;
      if (use64) {
        switch (epi) {
          case CEPI_BIAS: return launch_conv_swap<CEPI_BIAS, 64>(maps, a, stream);
          case CEPI_BIAS_LRELU: return launch_conv_swap<CEPI_BIAS_LRELU, 64>(maps, a, stream);
          case CEPI_BIAS_RES: return launch_conv_swap<CEPI_BIAS_RES, 64>(maps, a, stream);
          default: return launch_conv_swap<CEPI_MASK_LRELU, 64>(maps, a, stream);
        }
      }
      switch (epi) {
        case CEPI_BIAS: return launch_conv_swap<CEPI_BIAS, 128>(maps, a, stream);
        case CEPI_BIAS_LRELU: return launch_conv_swap<CEPI_BIAS_LRELU, 128>(maps, a, stream);
        case CEPI_BIAS_RES: return launch_conv_swap<CEPI_BIAS_RES, 128>(maps, a, stream);
        default: return launch_conv_swap<CEPI_MASK_LRELU, 128>(maps, a, stream);
      }
    }
  }
  if ((rc = make_act_maps(maps.a, x, B, H, W, Cin_p, x_ps, CONV_TW, CONV_TH))) return rc;
  if (out1) {
    for (int i = 0; i < 4; ++i) maps.c[i] = maps.a[0];  // unused: the fp32 output is written with plain stores
  } else if ((rc = make_act_maps(maps.c, y, B, H, W, Cout_p, y_ps, CONV_TW, CONV_TH))) {
    return rc;
  }
  for (int i = 1; i < 4; ++i) {
    if (!x_ps) maps.a[i] = maps.a[0];
    if (!y_ps) maps.c[i] = maps.c[0];
  }
  maps.c2 = maps.c[0];
  maps.r = maps.c[0];
  if (y2 && (rc = make_tmap_nhwc(&maps.c2, y2, Cout_p, W, H, B, Cout_p, (uint64_t)W * Cout_p, (uint64_t)H * W * Cout_p, CONV_TW, CONV_TH))) return rc;
  if (r && (rc = make_tmap_nhwc(&maps.r, r, Cout_p, W, H, B, Cout_p, (uint64_t)W * Cout_p, (uint64_t)H * W * Cout_p, CONV_TW, CONV_TH))) return rc;
  const int bn = Cout_p;  // one N tile covers all output channels (64 / 128 / 192 / 256)
  if ((rc = make_tmap_2d(&maps.w, wk, Cout_p, 9 * (uint64_t)Cin_p, 9 * (uint64_t)Cin_p, bn))) return rc;
  ConvArgs a{};
  a.B = B; a.H = H; a.W = W; a.Cin_p = Cin_p; a.Cout_p = Cout_p; a.n_real = n_real; a.bias = bias; a.slope = slope;
  a.a_split = x_ps; a.c_split = y_ps; a.alpha = 1.0f;
  a.y32 = out1 ? static_cast<float*>(y) : nullptr;
  const bool need_r = (epi == CEPI_BIAS_RES || epi == CEPI_MASK_LRELU || epi == CEPI_MUL);
  if (need_r && (!r || y_ps)) return fail(SRK_ERR_ARG, "conv3x3: epilogue needs an aux tensor (and a plain output)");
  if (epi == CEPI_BIAS_GELU && (!y2 || y_ps)) return fail(SRK_ERR_ARG, "conv3x3: GELU epilogue needs y2");
#define SRK_CCASE(BN_, EPI_) if (bn == BN_ && epi == EPI_) return launch_conv<BN_, EPI_>(maps, a, stream);
  SRK_CCASE(64, CEPI_BIAS) SRK_CCASE(128, CEPI_BIAS) SRK_CCASE(192, CEPI_BIAS) SRK_CCASE(256, CEPI_BIAS)
  SRK_CCASE(64, CEPI_BIAS_LRELU) SRK_CCASE(192, CEPI_BIAS_LRELU)
  SRK_CCASE(192, CEPI_BIAS_RES) SRK_CCASE(64, CEPI_BIAS_RES)
  SRK_CCASE(64, CEPI_MASK_LRELU) SRK_CCASE(192, CEPI_MASK_LRELU)
  SRK_CCASE(64, CEPI_BIAS_GELU) SRK_CCASE(128, CEPI_BIAS_GELU)
  SRK_CCASE(64, CEPI_MUL) SRK_CCASE(128, CEPI_MUL)
  SRK_CCASE(16, CEPI_OUT1)
#undef SRK_CCASE
  return fail(SRK_ERR_UNSUPPORTED, "conv3x3: no kernel instance for (Cout_p, epilogue)");
}

extern "C" long long srk_conv3x3_wgrad_ws_floats(int Cin_p, int Cout_p) {
  const long long a = (long long)num_sms() * 128 * Cin_p + 16;   // co_tiles*9*splits <= num_sms tiles of [128 x Cin_p]
  const long long b = (long long)num_sms() * 128 * 9 * WT_NCO_MAX;   // thin variant: ci_tiles*splits <= num_sms tiles of [128 x 9 x NCO]
  return a > b ? a : b;
}

extern "C" int srk_conv3x3_wgrad(int B, int H, int W, int Cin, int Cout, int Cin_p, int Cout_p, int ps,
                                 const void* dy, const void* x, float* ws, float* dw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (H % 4 || W % 16 || Cin_p % 64 || Cout_p % 64 || Cin_p > 256) return fail(SRK_ERR_UNSUPPORTED, "conv wgrad: shape");
  ConvWgradMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = make_act_maps(maps.a, dy, B, H, W, Cout_p, ps, 16, 4))) return rc;
  if (!ps) for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
  if ((rc = make_tmap_nhwc(&maps.b, x, Cin_p, W, H, B, Cin_p, (uint64_t)W * Cin_p, (uint64_t)H * W * Cin_p, 16, 4))) return rc;
  ConvWgradArgs a{};
  a.B = B; a.H = H; a.W = W; a.Cin_p = Cin_p; a.Cout_p = Cout_p; a.co_tiles = (Cout_p + 127) / 128;
  a.splits = conv_wgrad_splits(B, H, W, a.co_tiles); a.partials = ws; a.a_split = ps;
  switch (Cin_p) {
    case 64: rc = launch_conv_wgrad<64>(maps, a, stream); break;
    case 128: rc = launch_conv_wgrad<128>(maps, a, stream); break;
    case 192: rc = launch_conv_wgrad<192>(maps, a, stream); break;
    case 256: rc = launch_conv_wgrad<256>(maps, a, stream); break;
    default: return fail(SRK_ERR_UNSUPPORTED, "conv wgrad: Cin_p");
  }
  if (rc) return rc;
  const int total = Cout * Cin * 9;
  conv_unpack_wgrad_kernel<<<(total + 255) / 256, 256, 0, stream>>>(ws, a.splits, a.co_tiles * 128, Cin_p, dw, Cout, Cin,
                                                                    Cout_p, ps);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_bias_grad_nhwc(const void* dy, int B, int H, int W, int C, int ps, float* ws, float* db, int n_out,
                                  void* stream_) {
  // dy: [B,H,W,C] bf16, or when ps the shuffled [B,2H,2W,C/4]; db[n_out]
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = num_sms() * 2;
  if (!ps) {
    const int threads = (C / 2) * (C >= 128 ? 4 : 8);
    colsum_nhwc_kernel<<<grid, threads, C * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(dy),
                                                                       (long long)B * H * W, C, C, ws);
    SRK_LAUNCHED(1);
    colsum_finish_kernel<<<(n_out + 127) / 128, 128, 0, stream>>>(ws, grid, C, db, n_out);
  } else {
    const int Cg = C / 4;
    colsum_ps_kernel<<<grid, (Cg / 2) * 8, 4 * Cg * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(dy), B,
                                                                              2 * H, 2 * W, Cg, ws);
    SRK_LAUNCHED(1);
    colsum_finish_kernel<<<(n_out + 127) / 128, 128, 0, stream>>>(ws, grid, C, db, n_out);
  }
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" long long srk_small_ws_floats(void) { return (long long)num_sms() * 2 * 4096; }

extern "C" int srk_conv_in1_fwd(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int C,
                                int Cp, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool small = (long long)B * H * W * (Cp / 8) + (long long)num_sms() * 8 * 256 < (1LL << 32);
  if (small && W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    conv_in1_fwd_x4_kernel<unsigned><<<num_sms() * 8, 256, Cp * 10 * sizeof(float), stream>>>(x, w, bias, static_cast<__nv_bfloat16*>(y),
                                                                                          B, H, W, C, Cp);
  else if (small)
    conv_in1_fwd_kernel<unsigned><<<num_sms() * 8, 256, Cp * 10 * sizeof(float), stream>>>(x, w, bias, static_cast<__nv_bfloat16*>(y),
                                                                                       B, H, W, C, Cp);
  else
    conv_in1_fwd_kernel<unsigned long long><<<num_sms() * 8, 256, Cp * 10 * sizeof(float), stream>>>(
        x, w, bias, static_cast<__nv_bfloat16*>(y), B, H, W, C, Cp);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_conv_in1_wgrad(const float* x, const void* dy, float* ws, float* dw, float* db, int B, int H, int W,
                                  int C, int Cp, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // 8 pixels in flight per block at the generators' widths (Cp = 192 / 64 at 128^2); the discriminator's first layer
  // (Cp = 64 at 512^2: 16x the pixels) gets 16 pixel lanes per block and three (resident) blocks per SM
  const bool wide = Cp <= 64 && (long long)B * H * W >= (1 << 18);
  // (130 registers per thread: 128-thread blocks keep three of them resident per SM)
  const int grid = wide ? num_sms() * 3 : num_sms();   // one resident wave: the 80 shared-memory atomics per thread of the epilogue are paid once
  const int threads = (Cp / 8) * (wide ? 16 : 8);
  if ((long long)B * H * W + (long long)grid * threads < (1LL << 32))
    conv_in1_wgrad_kernel<unsigned><<<grid, threads, Cp * 10 * sizeof(float), stream>>>(x, static_cast<const __nv_bfloat16*>(dy), ws,
                                                                                    B, H, W, Cp);
  else
    conv_in1_wgrad_kernel<unsigned long long><<<grid, threads, Cp * 10 * sizeof(float), stream>>>(
        x, static_cast<const __nv_bfloat16*>(dy), ws, B, H, W, Cp);
  SRK_LAUNCHED(1);
  conv_in1_wgrad_finish_kernel<<<(C * 10 * 32 + 255) / 256, 256, 0, stream>>>(ws, grid, C, Cp, dw, db);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_conv_out1_fwd(const void* x, const float* w, const float* bias, float* y, int B, int H, int W, int C,
                                 void* stream_) {
  if (C != 64) return fail(SRK_ERR_UNSUPPORTED, "conv_out1: 64 input channels");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (H % 4 == 0)
    conv_out1_fwd_rows_kernel<64><<<num_sms() * 3, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), w, bias, y, B, H, W, 64);
  else
    conv_out1_fwd_kernel<64><<<num_sms() * 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), w, bias, y, B, H, W);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_conv_out1_bwd(const float* dy, const void* x, const float* w, void* dx, float* ws, float* dw, float* db,
                                 int B, int H, int W, int C, void* stream_) {
  if (C != 64) return fail(SRK_ERR_UNSUPPORTED, "conv_out1: 64 input channels");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool small = (long long)B * H * W * 8 < (1LL << 31);
  const bool rows4 = (H % 4) == 0;   // row-walking kernels: a warp owns four image rows
  const int seg = 64;
  if (rows4) conv_out1_dgrad_rows_kernel<64><<<num_sms() * 2, 256, 0, stream>>>(dy, w, static_cast<__nv_bfloat16*>(dx), B, H, W, seg);
  else if (small) conv_out1_dgrad_kernel<64, int><<<num_sms() * 8, 256, 0, stream>>>(dy, w, static_cast<__nv_bfloat16*>(dx), B, H, W);
  else conv_out1_dgrad_kernel<64, long long><<<num_sms() * 8, 256, 0, stream>>>(dy, w, static_cast<__nv_bfloat16*>(dx), B, H, W);
  SRK_LAUNCHED(1);
  const int grid = num_sms() * 2;
  if (rows4) conv_out1_wgrad_rows_kernel<64><<<grid, 256, 0, stream>>>(dy, static_cast<const __nv_bfloat16*>(x), ws, B, H, W, seg);
  else if (small) conv_out1_wgrad_kernel<64, int><<<grid, 256, 0, stream>>>(dy, static_cast<const __nv_bfloat16*>(x), ws, B, H, W);
  else conv_out1_wgrad_kernel<64, long long><<<grid, 256, 0, stream>>>(dy, static_cast<const __nv_bfloat16*>(x), ws, B, H, W);
  SRK_LAUNCHED(1);
  colsum_finish_kernel<<<(64 * 9 + 1 + 127) / 128, 128, 0, stream>>>(ws, grid, 64 * 9 + 1, ws + (size_t)grid * (64 * 9 + 1), 64 * 9 + 1);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaMemcpyAsync(dw, ws + (size_t)grid * (64 * 9 + 1), 64 * 9 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  SRK_CUDA_OK(cudaMemcpyAsync(db, ws + (size_t)grid * (64 * 9 + 1) + 64 * 9, sizeof(float), cudaMemcpyDeviceToDevice, stream));
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Channel-slice ("view") variants: every activation operand is (pointer to first channel, visible channels, pixel
// pitch).  TMA clips the 64-channel boxes to the visible channels (zero-fill on load, no write on store), so a
// convolution can read the first Cin channels of a wider buffer and write its Cout channels at an offset of the same
// or another buffer: torch.cat of the dense blocks (hybridmodels_hat.py:38-43) costs no copy.
namespace {
int view_map(CUtensorMap* m, const SrkView* v, int B, int H, int W, int bw, int bh) {
  if (!v || !v->ptr || v->C <= 0 || v->C % 8 || v->pitch % 8 || v->C > v->pitch)
    return fail(SRK_ERR_ARG, "view: channels / pitch must be positive multiples of 8 (16-byte slices)");
  return make_tmap_nhwc(m, v->ptr, (uint64_t)v->C, W, H, B, (uint64_t)v->pitch, (uint64_t)W * v->pitch,
                        (uint64_t)H * W * v->pitch, bw, bh);
}
int ew_grid(long long items) {
  long long g = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace

extern "C" int srk_conv3x3_igemm_v(int epi, int B, int H, int W, int Cin_p, int Cout_p, int n_real, const SrkView* x,
                                   const void* wk, const float* bias, float slope, float alpha, const SrkView* y,
                                   const SrkView* r, float* y32, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (H % CONV_TH || W % CONV_TW) return fail(SRK_ERR_UNSUPPORTED, "conv3x3: H % 8 == 0 and W % 16 == 0 required");
  const bool out1 = (epi == CEPI_OUT1);
  if (Cin_p % 64 || (Cout_p % 64 && !(out1 && Cout_p == 16)) || Cout_p > 256 || Cin_p > 256)
    return fail(SRK_ERR_UNSUPPORTED, "conv3x3: channels must be multiples of 64, <= 256");
  if (!x || !wk || (!out1 && !y) || (out1 && (!y32 || !bias))) return fail(SRK_ERR_ARG, "conv3x3_v: null pointer");
  if (x->C > Cin_p || (!out1 && y->C > Cout_p)) return fail(SRK_ERR_ARG, "conv3x3_v: view wider than the padded channel count");
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  const int hmode = conv_halo_mode();
  // Few output channels (<= 128): role-swapped kernel — weights as the M = 128 operand, 256 pixels as N (conv3x3_swap.cuh)
  // Measured (hybrid step, B200): it wins where the tensor unit is the limit — plain epilogues (105 vs 140-160 us per
  // dense-block conv) and aux epilogues with <= 64 output channels over >= 2 input chunks (conv5); the 24 -> 72..144
  // input-gradient layers (one chunk, two channel boxes per pixel group, aux loads) are bound by its transposing
  // epilogue instead and stay on the halo / per-tap kernels.  SRK_CONV_HALO=6 forces it for every eligible layer (tests).
  const bool swap_aux = (epi == CEPI_BIAS_RES || epi == CEPI_MASK_LRELU);
  const bool swap_ok = !out1 && n_real <= 128 && n_real > 0 && H % SWP_TH == 0 && W % SWP_TW == 0 &&
                       (epi == CEPI_BIAS || epi == CEPI_BIAS_LRELU || swap_aux);
  const bool swap_pick = !swap_aux || (n_real <= 64 && Cin_p >= 128);
  if (hmode != 0 && hmode != 5 && swap_ok && (swap_pick || hmode == 6)) {
    const bool need_aux = swap_aux;
    if (need_aux && !r) return fail(SRK_ERR_ARG, "conv3x3_v: epilogue needs an aux view");
    if ((rc = view_map(&maps.a[0], x, B, H, W, SWP_BW, SWP_BH))) return rc;
    if ((rc = view_map(&maps.c[0], y, B, H, W, 8, 8))) return rc;
    for (int i = 1; i < 4; ++i) { maps.a[i] = maps.a[0]; maps.c[i] = maps.c[0]; }
    maps.c2 = maps.c[0];
    maps.r = maps.c[0];
    if (need_aux && (rc = view_map(&maps.r, r, B, H, W, 8, 8))) return rc;
    const int m64 = swap_m64_mode();
    const bool use64 = m64 != 0 && n_real <= 64;
    if ((rc = make_tmap_2d(&maps.w, wk, Cout_p, 9 * (uint64_t)Cin_p, 9 * (uint64_t)Cin_p, use64 ? 64 : 128))) return rc;
    ConvArgs a{};
    a.B = B; a.H = H; a.W = W; a.Cin_p = Cin_p; a.Cout_p = Cout_p; a.n_real = n_real; a.bias = bias; a.slope = slope;
    a.alpha = alpha; a.c_split = 0;
    if (use64) {
      switch (epi) {
        case CEPI_BIAS: return launch_conv_swap<CEPI_BIAS, 64>(maps, a, stream);
        case CEPI_BIAS_LRELU: return launch_conv_swap<CEPI_BIAS_LRELU, 64>(maps, a, stream);
        case CEPI_BIAS_RES: return launch_conv_swap<CEPI_BIAS_RES, 64>(maps, a, stream);
        default: return launch_conv_swap<CEPI_MASK_LRELU, 64>(maps, a, stream);
      }
    }
    switch (epi) {
      case CEPI_BIAS: return launch_conv_swap<CEPI_BIAS, 128>(maps, a, stream);
      case CEPI_BIAS_LRELU: return launch_conv_swap<CEPI_BIAS_LRELU, 128>(maps, a, stream);
      case CEPI_BIAS_RES: return launch_conv_swap<CEPI_BIAS_RES, 128>(maps, a, stream);
      default: return launch_conv_swap<CEPI_MASK_LRELU, 128>(maps, a, stream);
    }
  }
  // Measured on B200 (hybrid step, tools/gpu_probe_hybrid_prof.py): the halo-resident kernel wins where one 64-channel
  // chunk feeds all nine taps and the weight ring stays deep (Cin_p == 64, BN <= 128: 151 -> 100 us, 189 -> 166 us); with
  // two or three chunks per tile its two-stage halo ring exposes the TMA latency and the per-tap kernel is as fast or
  // faster, so those shapes stay on conv3x3_kernel.  SRK_CONV_HALO=3 forces it everywhere (tests), 0 disables it.
  const bool halo_shape = H % HALO_TH == 0 && W % HALO_TW == 0;
  // thin layers (<= 32 real output channels, e.g. the 24-channel conv1..conv4 of a dense block): N = 32 instance with a
  // three-stage halo ring and a whole tile of weights in flight
  const bool thin = halo_shape && hmode != 0 && !out1 && (epi == CEPI_BIAS || epi == CEPI_BIAS_LRELU) && y->C <= 32 &&
                    n_real <= 32 && Cout_p == 64;
  const bool halo = halo_shape && (thin || hmode >= 2 || (hmode == 1 && Cin_p == 64 && Cout_p <= 128));
  const int cbw = halo ? HALO_TW : CONV_TW, cbh = halo ? HALO_TH : CONV_TH;
  if ((rc = view_map(&maps.a[0], x, B, H, W, halo ? HALO_BW : CONV_TW, halo ? HALO_BH / 3 : CONV_TH))) return rc;
  if (out1) maps.c[0] = maps.a[0];
  else if ((rc = view_map(&maps.c[0], y, B, H, W, cbw, cbh))) return rc;
  for (int i = 1; i < 4; ++i) { maps.a[i] = maps.a[0]; maps.c[i] = maps.c[0]; }
  maps.c2 = maps.c[0];
  maps.r = maps.c[0];
  const bool need_r = (epi == CEPI_BIAS_RES || epi == CEPI_MASK_LRELU || epi == CEPI_MUL);
  if (need_r) {
    if (!r) return fail(SRK_ERR_ARG, "conv3x3_v: epilogue needs an aux view");
    if ((rc = view_map(&maps.r, r, B, H, W, cbw, cbh))) return rc;
  }
  if (epi == CEPI_BIAS_GELU) return fail(SRK_ERR_UNSUPPORTED, "conv3x3_v: GELU epilogue is served by srk_conv3x3_igemm");
  const int bn = thin ? 32 : Cout_p;
  if ((rc = make_tmap_2d(&maps.w, wk, Cout_p, 9 * (uint64_t)Cin_p, 9 * (uint64_t)Cin_p, bn))) return rc;
  ConvArgs a{};
  a.B = B; a.H = H; a.W = W; a.Cin_p = Cin_p; a.Cout_p = thin ? 32 : Cout_p; a.n_real = n_real; a.bias = bias; a.slope = slope;
  a.a_split = 0; a.c_split = 0; a.alpha = alpha;
  a.y32 = out1 ? y32 : nullptr;
  if (halo) {
#define SRK_HCASE(BN_, EPI_) if (bn == BN_ && epi == EPI_) return launch_conv_halo<BN_, EPI_>(maps, a, stream);
    SRK_HCASE(32, CEPI_BIAS) SRK_HCASE(32, CEPI_BIAS_LRELU)
    SRK_HCASE(64, CEPI_BIAS) SRK_HCASE(128, CEPI_BIAS) SRK_HCASE(192, CEPI_BIAS) SRK_HCASE(256, CEPI_BIAS)
    SRK_HCASE(64, CEPI_BIAS_LRELU) SRK_HCASE(192, CEPI_BIAS_LRELU)
    SRK_HCASE(64, CEPI_BIAS_RES) SRK_HCASE(128, CEPI_BIAS_RES) SRK_HCASE(192, CEPI_BIAS_RES) SRK_HCASE(256, CEPI_BIAS_RES)
    SRK_HCASE(64, CEPI_MASK_LRELU) SRK_HCASE(192, CEPI_MASK_LRELU)
    SRK_HCASE(16, CEPI_OUT1)
#undef SRK_HCASE
    return fail(SRK_ERR_UNSUPPORTED, "conv3x3_v: no halo kernel instance for (Cout_p, epilogue)");
  }
#define SRK_CCASE(BN_, EPI_) if (bn == BN_ && epi == EPI_) return launch_conv<BN_, EPI_>(maps, a, stream);
  SRK_CCASE(64, CEPI_BIAS) SRK_CCASE(128, CEPI_BIAS) SRK_CCASE(192, CEPI_BIAS) SRK_CCASE(256, CEPI_BIAS)
  SRK_CCASE(64, CEPI_BIAS_LRELU) SRK_CCASE(192, CEPI_BIAS_LRELU)
  SRK_CCASE(64, CEPI_BIAS_RES) SRK_CCASE(128, CEPI_BIAS_RES) SRK_CCASE(192, CEPI_BIAS_RES) SRK_CCASE(256, CEPI_BIAS_RES)
  SRK_CCASE(64, CEPI_MASK_LRELU) SRK_CCASE(192, CEPI_MASK_LRELU)
  SRK_CCASE(16, CEPI_OUT1)
#undef SRK_CCASE
  return fail(SRK_ERR_UNSUPPORTED, "conv3x3_v: no kernel instance for (Cout_p, epilogue)");
}

extern "C" int srk_conv3x3_wgrad_v(int B, int H, int W, int Cin, int Cout, int Cin_p, int Cout_p, const SrkView* dy,
                                   const SrkView* x, float* ws, float* dw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (H % 4 || W % 16 || Cin_p % 64 || Cout_p % 64 || Cin_p > 256) return fail(SRK_ERR_UNSUPPORTED, "conv wgrad: shape");
  if (!dy || !x || dy->C > Cout_p || x->C > Cin_p || Cout > Cout_p || Cin > Cin_p) return fail(SRK_ERR_ARG, "conv wgrad_v: views");
  int rc;
  if (conv_halo_mode() != 0 && Cout <= WT_NCO_MAX && dy->C <= 64) {
    // few output channels: one CTA per pixel range computes all nine taps from one halo load (conv3x3_wgrad_thin_kernel)
    ConvWgradThinMaps tm;
    memset(&tm, 0, sizeof(tm));
    if ((rc = view_map(&tm.dy, dy, B, H, W, 16, 4))) return rc;
    if ((rc = view_map(&tm.x, x, B, H, W, 18, 6))) return rc;
    const int nco = Cout <= 32 ? 32 : 48;
    static DeviceOnce configured;
    if (configured.need()) {
      SRK_CUDA_OK(cudaFuncSetAttribute(conv3x3_wgrad_thin_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
      SRK_CUDA_OK(cudaFuncSetAttribute(conv3x3_wgrad_thin_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
      configured.done();
    }
    ConvWgradThinArgs t{};
    t.B = B; t.H = H; t.W = W; t.ci_tiles = (Cin_p + 127) / 128;
    const int iters = B * (H / 4) * (W / 16);
    t.splits = num_sms() / t.ci_tiles;
    if (t.splits > iters) t.splits = iters;
    if (t.splits < 1) t.splits = 1;
    t.partials = ws;
    if (nco == 32) conv3x3_wgrad_thin_kernel<32><<<t.ci_tiles * t.splits, WG_THREADS, WT_SMEM, stream>>>(tm, t);
    else conv3x3_wgrad_thin_kernel<48><<<t.ci_tiles * t.splits, WG_THREADS, WT_SMEM, stream>>>(tm, t);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
    const int total = t.ci_tiles * 128 * 9 * nco;
    conv_unpack_wgrad_thin_kernel<<<(total + 255) / 256, 256, 0, stream>>>(ws, t.splits, t.ci_tiles, nco, dw, Cout, Cin);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
    return SRK_OK;
  }
  ConvWgradMaps maps;
  memset(&maps, 0, sizeof(maps));
  if ((rc = view_map(&maps.a[0], dy, B, H, W, 16, 4))) return rc;
  for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
  if ((rc = view_map(&maps.b, x, B, H, W, 16, 4))) return rc;
  ConvWgradArgs a{};
  a.B = B; a.H = H; a.W = W; a.Cin_p = Cin_p; a.Cout_p = Cout_p; a.co_tiles = (Cout_p + 127) / 128;
  a.splits = conv_wgrad_splits(B, H, W, a.co_tiles); a.partials = ws; a.a_split = 0;
  switch (Cin_p) {
    case 64: rc = launch_conv_wgrad<64>(maps, a, stream); break;
    case 128: rc = launch_conv_wgrad<128>(maps, a, stream); break;
    case 192: rc = launch_conv_wgrad<192>(maps, a, stream); break;
    case 256: rc = launch_conv_wgrad<256>(maps, a, stream); break;
    default: return fail(SRK_ERR_UNSUPPORTED, "conv wgrad: Cin_p");
  }
  if (rc) return rc;
  const int total = Cout * Cin * 9;
  conv_unpack_wgrad_kernel<<<(total + 255) / 256, 256, 0, stream>>>(ws, a.splits, a.co_tiles * 128, Cin_p, dw, Cout, Cin,
                                                                    Cout_p, 0);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_bias_grad_v(const SrkView* dy, long long npix, float* ws, float* db, int n_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dy || !dy->ptr || dy->C % 2 || dy->C > 256 || n_out > dy->C) return fail(SRK_ERR_ARG, "bias_grad_v: view");
  const int grid = num_sms() * 2, C = dy->C;
  const int threads = (C / 2) * (C >= 128 ? 4 : 8);
  colsum_nhwc_kernel<<<grid, threads, C * sizeof(float), stream>>>(static_cast<const __nv_bfloat16*>(dy->ptr), npix, C,
                                                                     dy->pitch, ws);
  SRK_LAUNCHED(1);
  colsum_finish_kernel<<<(n_out + 127) / 128, 128, 0, stream>>>(ws, grid, C, db, n_out);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_view_lrelu_mask(const SrkView* g, const SrkView* f, long long npix, float slope, float* ws, float* colsum,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!g || !f || g->C != f->C || g->C % 8 || g->pitch % 8 || f->pitch % 8 || g->C > 256)
    return fail(SRK_ERR_ARG, "view_lrelu_mask: views");
  if (colsum && !ws) return fail(SRK_ERR_ARG, "view_lrelu_mask: column sums need the small workspace");
  const int groups = g->C / 8;
  int grid = ew_grid(npix * groups);
  if (colsum && grid > num_sms() * 4) grid = num_sms() * 4;   // few partial rows for the finishing kernel
  grid = (grid + groups - 1) / groups * groups;   // total threads divisible by `groups`: a thread keeps its channel group
  view_lrelu_mask_kernel<<<grid, 256, 256 * 8 * sizeof(float), stream>>>(
      static_cast<__nv_bfloat16*>(const_cast<void*>(g->ptr)), g->pitch, static_cast<const __nv_bfloat16*>(f->ptr), f->pitch,
      g->C, npix, slope, colsum ? ws : nullptr);
  SRK_LAUNCHED(1);
  if (colsum) {
    colsum_finish_kernel<<<(g->C + 127) / 128, 128, 0, stream>>>(ws, grid, g->C, colsum, g->C);
    SRK_LAUNCHED(1);
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_view_axpy(const SrkView* y, const SrkView* a, const SrkView* x, long long npix, float alpha, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!y || !a || y->C != a->C || (x && x->C != y->C) || y->C % 8 || y->pitch % 8 || a->pitch % 8 || (x && x->pitch % 8))
    return fail(SRK_ERR_ARG, "view_axpy: views");
  view_axpy_kernel<<<ew_grid(npix * (y->C / 8)), 256, 0, stream>>>(
      static_cast<__nv_bfloat16*>(const_cast<void*>(y->ptr)), y->pitch, static_cast<const __nv_bfloat16*>(a->ptr), a->pitch,
      x ? static_cast<const __nv_bfloat16*>(x->ptr) : nullptr, x ? x->pitch : 0, y->C, npix, alpha);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_nearest2_fwd(const SrkView* x, const SrkView* y, int B, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || x->C != y->C || x->C % 8 || x->pitch % 8 || y->pitch % 8) return fail(SRK_ERR_ARG, "nearest2_fwd: views");
  nearest2_fwd_kernel<<<ew_grid((long long)B * H * W * (x->C / 8)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x->ptr), x->pitch, static_cast<__nv_bfloat16*>(const_cast<void*>(y->ptr)), y->pitch,
      x->C, B, H, W);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_nearest2_bwd(const SrkView* dy, const SrkView* dx, int B, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dx || !dy || dx->C != dy->C || dx->C % 8 || dx->pitch % 8 || dy->pitch % 8) return fail(SRK_ERR_ARG, "nearest2_bwd: views");
  nearest2_bwd_kernel<<<ew_grid((long long)B * H * W * (dx->C / 8)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dy->ptr), dy->pitch, static_cast<__nv_bfloat16*>(const_cast<void*>(dx->ptr)),
      dx->pitch, dx->C, B, H, W);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_img1_pack(const float* x, void* y8, long long npix, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  img1_pack_kernel<<<ew_grid(npix), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(y8), npix);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_img1_unpack(const void* x8, float* y, long long npix, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  img1_unpack_kernel<<<ew_grid(npix), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x8), y, npix);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}
