// api_disc.cu — C-ABI entry points for the data movement around the discriminator's 4x4 stride-2 (transposed)
// convolutions (disc_aux.cuh).  The arithmetic itself is srk_gemm_tn / srk_gemm_tn_lrelu / srk_gemm_wgrad.
#include "disc_aux.cuh"
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
  const long long vectors = (long long)B * (H / 2) * (W / 2) * 16 * (x->C / 8);
  disc_patches_k4s2_kernel<<<stream_grid(vectors), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x->ptr), x->pitch, f ? static_cast<const __nv_bfloat16*>(f->ptr) : nullptr,
      f ? f->pitch : 0, slope, B, H, W, x->C, static_cast<__nv_bfloat16*>(patches));
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
  disc_fold_k4s2_kernel<<<stream_grid(vectors), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(taps), B, Hi, Wi, y->C, add ? static_cast<const __nv_bfloat16*>(add->ptr) : nullptr,
      add ? add->pitch : 0, (act == DISC_ACT_MASK) ? static_cast<const __nv_bfloat16*>(f->ptr) : nullptr,
      (act == DISC_ACT_MASK) ? f->pitch : 0, act, slope, static_cast<__nv_bfloat16*>(const_cast<void*>(y->ptr)), y->pitch);
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
