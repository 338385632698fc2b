// api_disc.cu — C-ABI entry points for the data movement around the discriminator's 4x4 stride-2 (transposed)
// convolutions (disc_aux.cuh).  The arithmetic itself is srk_gemm_tn / srk_gemm_tn_lrelu / srk_gemm_wgrad.
#include "disc_aux.cuh"
#include "spectral_norm.cuh"
#include "srk_host.h"

using namespace srk;

namespace {
int check_view(const SrkView* v, const char* what) {
  if (!v || !v->ptr || v->C <= 0 || v->C % 8 || v->pitch % 8 || v->C > v->pitch || (reinterpret_cast<uintptr_t>(v->ptr) & 15))
    return fail(SRK_ERR_ARG, what);
  return SRK_OK;
}
int stream_grid(long long vectors) {
  long long blocks = (vectors + 255) / 256;
  const long long cap = (long long)num_sms() * 16;   // grid-stride beyond 16 resident blocks of 256 threads per SM
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : int(blocks);
}
}  // namespace

extern "C" int srk_disc_patches_k4s2(const SrkView* x, const SrkView* f, float slope, int B, int H, int W, void* patches,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = check_view(x, "disc_patches: x view (channels / pitch multiples of 8, 16-byte aligned)"))) return rc;
  if (f && ((rc = check_view(f, "disc_patches: f view")) || f->C != x->C)) return rc ? rc : fail(SRK_ERR_ARG, "disc_patches: f and x differ in channels");
  if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || !patches || (reinterpret_cast<uintptr_t>(patches) & 15))
    return fail(SRK_ERR_ARG, "disc_patches: even H and W and a 16-byte aligned patch matrix required");
  const long long vectors = (long long)B * (H / 2) * (W / 2) * (x->C / 8);   // threads: one per (patch row, 8-channel group)
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x->ptr);
  const __nv_bfloat16* fp = f ? static_cast<const __nv_bfloat16*>(f->ptr) : nullptr;
  // + one grid stride of headroom so that the 32-bit loop counter cannot wrap
  if (vectors + (long long)num_sms() * 16 * 256 < (1LL << 32))
    disc_patches_k4s2_kernel<unsigned><<<stream_grid(vectors), 256, 0, stream>>>(xp, x->pitch, fp, f ? f->pitch : 0, slope, B, H, W,
                                                                             x->C, static_cast<__nv_bfloat16*>(patches));
  else
    disc_patches_k4s2_kernel<unsigned long long><<<stream_grid(vectors), 256, 0, stream>>>(xp, x->pitch, fp, f ? f->pitch : 0, slope, B,
                                                                                       H, W, x->C, static_cast<__nv_bfloat16*>(patches));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_disc_fold_k4s2(const void* taps, int B, int Hi, int Wi, const SrkView* add, const SrkView* f, int act,
                                  float slope, const SrkView* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = check_view(y, "disc_fold: y view (channels / pitch multiples of 8, 16-byte aligned)"))) return rc;
  if (add && ((rc = check_view(add, "disc_fold: add view")) || add->C != y->C)) return rc ? rc : fail(SRK_ERR_ARG, "disc_fold: add and y differ in channels");
  if (act < DISC_ACT_NONE || act > DISC_ACT_MASK) return fail(SRK_ERR_ARG, "disc_fold: act must be 0, 1 or 2");
  if (act == DISC_ACT_MASK) {
    if ((rc = check_view(f, "disc_fold: the mask needs the forward activation view f"))) return rc;
    if (f->C != y->C) return fail(SRK_ERR_ARG, "disc_fold: f and y differ in channels");
  }
  if (B <= 0 || Hi <= 0 || Wi <= 0 || !taps || (reinterpret_cast<uintptr_t>(taps) & 15)) return fail(SRK_ERR_ARG, "disc_fold: shape / pointer");
  const long long vectors = (long long)B * (2 * Hi) * (2 * Wi) * (y->C / 8);
  const __nv_bfloat16* ap = add ? static_cast<const __nv_bfloat16*>(add->ptr) : nullptr;
  const __nv_bfloat16* fp = (act == DISC_ACT_MASK) ? static_cast<const __nv_bfloat16*>(f->ptr) : nullptr;
  const int ldf = (act == DISC_ACT_MASK) ? f->pitch : 0;
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(const_cast<void*>(y->ptr));
  if (vectors + (long long)num_sms() * 16 * 256 < (1LL << 32))
    disc_fold_k4s2_kernel<unsigned><<<stream_grid(vectors), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(taps), B, Hi, Wi, y->C, ap,
                                                                          add ? add->pitch : 0, fp, ldf, act, slope, yp, y->pitch);
  else
    disc_fold_k4s2_kernel<unsigned long long><<<stream_grid(vectors), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(taps), B, Hi, Wi,
                                                                                    y->C, ap, add ? add->pitch : 0, fp, ldf, act, slope,
                                                                                    yp, y->pitch);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_view_lrelu(const SrkView* y, long long npix, float slope, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = check_view(y, "view_lrelu: view (channels / pitch multiples of 8, 16-byte aligned)"))) return rc;
  if (npix <= 0) return fail(SRK_ERR_ARG, "view_lrelu: npix");
  view_lrelu_kernel<<<stream_grid(npix * (y->C / 8)), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(const_cast<void*>(y->ptr)),
                                                                       y->pitch, y->C, npix, slope);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_disc_prep_w4(const float* w, int P, int Q, const float* sigma, void* a, void* at, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!w || !a || P <= 0 || Q <= 0 || P % 32 || Q % 32 || (reinterpret_cast<uintptr_t>(w) & 15))
    return fail(SRK_ERR_ARG, "disc_prep_w4: P and Q must be multiples of 32, w 16-byte aligned");
  disc_prep_w4_kernel<<<stream_grid((long long)P * Q), 256, 0, stream>>>(w, P, Q, sigma, static_cast<__nv_bfloat16*>(a));
  SRK_LAUNCHED(1);
  if (at) {
    const long long tiles = (long long)(P / 32) * (16 * Q / 32);
    const long long cap = (long long)num_sms() * 8;
    transpose_bf16_kernel<<<int(tiles < cap ? tiles : cap), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(a), P, 16 * Q,
                                                                             static_cast<__nv_bfloat16*>(at));
    SRK_LAUNCHED(1);
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" long long srk_disc_wgrad4_ws_floats(int T, int R, int Cb) {
  const int Ca = 16 * R;
  const int cb = Cb < 256 ? Cb : 256;
  return srk_gemm_wgrad_workspace_elems(Ca, cb, srk_gemm_wgrad_splits(T, Ca));
}

extern "C" int srk_disc_wgrad4(int T, int R, int Cb, const void* A, int lda, const void* B, int ldb, float* ws, float* dw,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!A || !B || !ws || !dw || R <= 0 || R % 8 || Cb <= 0 || (Cb > 256 && Cb % 256) || (reinterpret_cast<uintptr_t>(dw) & 15))
    return fail(SRK_ERR_ARG, "disc_wgrad4: Cb must be 64 / 128 / 192 / 256 or a multiple of 256");
  const int Ca = 16 * R;
  const int splits = srk_gemm_wgrad_splits(T, Ca);
  const int ca_pad = ((Ca + 127) / 128) * 128;
  for (int c0 = 0; c0 < Cb; c0 += 256) {
    const int cb = (Cb - c0) < 256 ? (Cb - c0) : 256;
    int rc = gemm_wgrad_partials(T, Ca, cb, A, lda, static_cast<const __nv_bfloat16*>(B) + c0, ldb, ws, splits, stream_);
    if (rc) return rc;
    disc_unpack_wgrad4_kernel<<<stream_grid(4LL * R * cb), 256, 0, stream>>>(ws, splits, (long long)ca_pad * cb, R, cb, c0, dw);
    SRK_LAUNCHED(1);
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// torch.nn.utils.spectral_norm (spectral_norm.cuh)
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr long long SN_WS_FLOATS = 1LL << 19;
int sn_check(const SrkSnLayer& l) {
  if (!l.w || !l.u || !l.v || !l.sigma || l.A <= 0 || l.B <= 0 || l.KK <= 0 || (l.dim != 0 && l.dim != 1))
    return fail(SRK_ERR_ARG, "spectral_norm: layer descriptor (w, u, v, sigma, A, B, KK, dim)");
  return SRK_OK;
}
}  // namespace

extern "C" long long srk_spectral_norm_ws_floats(void) { return SN_WS_FLOATS; }

extern "C" int srk_spectral_norm(const SrkSnLayer* layers, int n, int power_iteration, float eps, float* ws, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!layers || n <= 0 || !ws) return fail(SRK_ERR_ARG, "spectral_norm: null argument");
  for (int i = 0; i < n; ++i) {
    const SrkSnLayer& l = layers[i];
    int rc = sn_check(l);
    if (rc) return rc;
    const SnDims d{l.A, l.B, l.KK, l.dim};
    const int U = l.dim == 0 ? l.A : l.B;
    const int V = (l.dim == 0 ? l.B : l.A) * l.KK;
    const int nb = (V + 255) / 256;
    int splits = (2 * num_sms() + nb - 1) / nb;
    if (splits > U) splits = U;
    if (splits < 1) splits = 1;
    if ((long long)(splits + 1) * V + U + nb > SN_WS_FLOATS) return fail(SRK_ERR_UNSUPPORTED, "spectral_norm: layer too large for the workspace");
    float* t_raw = ws + (long long)splits * V;   // [V]
    float* s_raw = t_raw + V;                    // [U]
    float* ssq = s_raw + U;                      // [nb]
    if (power_iteration) {
      sn_t_partial_kernel<<<dim3(nb, splits), 256, 0, stream>>>(l.w, l.u, d, U, V, ws);
      sn_t_reduce_kernel<<<nb, 256, 0, stream>>>(ws, splits, V, t_raw, ssq);
      SRK_LAUNCHED(2);
    }
    sn_s_rows_kernel<<<U, 256, 0, stream>>>(l.w, l.v, power_iteration ? t_raw : nullptr, ssq, nb, eps, d, V, s_raw);
    sn_s_finish_kernel<<<1, 256, 0, stream>>>(s_raw, U, eps, power_iteration ? 1 : 0, l.u, l.sigma);
    SRK_LAUNCHED(2);
    if (l.w_sn) {
      const long long cnt = (long long)l.A * l.B * l.KK;
      sn_scale_kernel<<<stream_grid(cnt), 256, 0, stream>>>(l.w, cnt, l.sigma, l.w_sn);
      SRK_LAUNCHED(1);
    }
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_spectral_norm_bwd(const SrkSnLayer* layers, int n, const float* const* dw_sn, float* const* dw, float* ws,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!layers || n <= 0 || !dw_sn || !dw || !ws) return fail(SRK_ERR_ARG, "spectral_norm_bwd: null argument");
  for (int i = 0; i < n; ++i) {
    const SrkSnLayer& l = layers[i];
    if (!dw_sn[i] || !dw[i]) continue;   // this layer's gradient was not requested
    int rc = sn_check(l);
    if (rc) return rc;
    const SnDims d{l.A, l.B, l.KK, l.dim};
    const long long cnt = (long long)l.A * l.B * l.KK;
    int g = stream_grid(cnt);
    if (g > 1024) g = 1024;
    float* part = ws + (long long)(i & 1) * 1024;   // two slots: layer i+1's partials must not overwrite what layer i's apply reads
    sn_dot_partial_kernel<<<g, 256, 0, stream>>>(dw_sn[i], l.w, cnt, part);
    sn_bwd_apply_kernel<<<stream_grid(cnt), 256, 0, stream>>>(dw_sn[i], part, g, l.sigma, l.u, l.v, d, dw[i]);
    SRK_LAUNCHED(2);
  }
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// models/discriminator_hat.py: bilinear x2 resize between the decoder's 3x3 convolutions
// ---------------------------------------------------------------------------------------------------------------
extern "C" int srk_bilinear2x_fwd(const SrkView* x, const SrkView* s, const SrkView* y, int B, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = check_view(x, "bilinear2x_fwd: x view")) || (rc = check_view(y, "bilinear2x_fwd: y view"))) return rc;
  if (s && ((rc = check_view(s, "bilinear2x_fwd: s view")) || s->C != x->C)) return rc ? rc : fail(SRK_ERR_ARG, "bilinear2x_fwd: s and x differ in channels");
  if (x->C != y->C || B <= 0 || H <= 0 || W <= 0) return fail(SRK_ERR_ARG, "bilinear2x_fwd: shape");
  const long long vectors = (long long)B * 4 * H * W * (x->C / 8);
  bilinear2x_fwd_kernel<<<stream_grid(vectors), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x->ptr), x->pitch, s ? static_cast<const __nv_bfloat16*>(s->ptr) : nullptr, s ? s->pitch : 0,
      static_cast<__nv_bfloat16*>(const_cast<void*>(y->ptr)), y->pitch, B, H, W, x->C);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_bilinear2x_bwd(const SrkView* dy, const SrkView* dx, int B, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = check_view(dy, "bilinear2x_bwd: dy view")) || (rc = check_view(dx, "bilinear2x_bwd: dx view"))) return rc;
  if (dx->C != dy->C || B <= 0 || H <= 0 || W <= 0) return fail(SRK_ERR_ARG, "bilinear2x_bwd: shape");
  const long long vectors = (long long)B * H * W * (dx->C / 8);
  bilinear2x_bwd_kernel<<<stream_grid(vectors), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy->ptr), dy->pitch,
                                                                 static_cast<__nv_bfloat16*>(const_cast<void*>(dx->ptr)), dx->pitch, B, H, W,
                                                                 dx->C);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}
