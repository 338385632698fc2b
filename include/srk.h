/* srk.h — C ABI of libsrk.so, the sm_100a kernel library behind superresolution_def_b200.
 *
 * The reference (GDev96/SuperResolution_Def) has no FFI: its hot path is nn.Module.forward code that
 * calls ATen.  Each entry point below names the reference lines whose arithmetic it replaces; the
 * Python binding (superresolution_def_b200/_capi.py) is the ctypes stub a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 * the caller owns every buffer (the library never allocates device memory and never synchronises);
 * `stream` is a cudaStream_t passed as void*; return value 0 = OK, negative = error (see SRK_ERR_*),
 * with a diagnostic on stderr.  Activations are bf16, row-major, "token-major": [tokens, channels].
 */
#ifndef SRK_H_
#define SRK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRK_OK 0
#define SRK_ERR_ARG (-1)
#define SRK_ERR_CUDA (-2)
#define SRK_ERR_UNSUPPORTED (-3)

/* Library / build identification ("sm_100a tcgen05"). */
const char* srk_version(void);

/* ---- epilogues of srk_gemm_tn (values match csrc/gemm_tn.cuh) ---- */
#define SRK_EPI_STORE 0  /* C = bf16(acc)                                                       */
#define SRK_EPI_GELU2 1  /* C = gelu(u), C2 = gelu'(u)  — Mlp.act, architecture_swin.py:20       */
#define SRK_EPI_MUL 2    /* C = acc * X1               — backward of Mlp.act                    */
#define SRK_EPI_RES_LN 3 /* C = acc + X1, C2 = LN(C)    — residual :149-150 + norm :127,150     */
#define SRK_EPI_LNBWD 4  /* C = X2 + LNbackward(acc)    — backward of the same                  */

typedef struct SrkLnArgs {
  int n_real;         /* real channel count normalised (180 / 90)                         */
  int ones_col;       /* column forced to 1.0 in the LN / GELU output, -1 for none        */
  const float* gamma; /* LN weight [n_real]                                               */
  const float* beta;  /* LN bias [n_real] (may be NULL for LNBWD)                         */
  float* stats;       /* [M][2] mean,rstd: written by RES_LN, read by LNBWD               */
  float* partials;    /* LNBWD: [srk_gemm_grid(...)][2][N] per-CTA sums for dgamma, dbeta */
  float eps;
} SrkLnArgs;

/* C[M,N] = epilogue(A[M,K] * B[N,K]^T): tcgen05 GEMM, bf16 in, fp32 accumulate.
 * Replaces nn.Linear forward / input-gradient: architecture_swin.py:73 (qkv), :94 (proj), :19-25 (fc1, fc2).
 * M % 128 == 0, K % 64 == 0, N % 64 == 0 (N <= 256 or a multiple of 192 / 256); ld* in elements. */
int srk_gemm_tn(int epi, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                void* C2, int ldc2, const void* X1, int ldx1, const void* X2, int ldx2, const SrkLnArgs* ln,
                void* stream);
/* Number of CTAs srk_gemm_tn launches for this shape (size of the LNBWD partials buffer). */
int srk_gemm_grid(int M, int N);

/* dW[Ca,Cb] (fp32) = sum_t A[t,ca] * B[t,cb] over T tokens: tcgen05 GEMM with MN-major operands.
 * Replaces autograd's weight/bias gradient of nn.Linear (same lines as above).
 * Cb in {64,128,192,256}; T % (64*splits) == 0; workspace >= splits*ceil(Ca/128)*128*Cb floats;
 * out is [ceil(Ca/128)*128, Cb] fp32 (rows >= Ca are zero). */
int srk_gemm_wgrad(int T, int Ca, int Cb, const void* A, int lda, const void* B, int ldb, float* workspace,
                   int splits, float* out, void* stream);
/* debug variant with explicit UMMA descriptor byte offsets (used once to validate the layout on hardware) */
int srk_gemm_wgrad_dbg(int T, int Ca, int Cb, const void* A, int lda, const void* B, int ldb, float* workspace,
                       int splits, float* out, int lbo_bytes, int sbo_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRK_H_ */
