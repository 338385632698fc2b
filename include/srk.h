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
/* Number of CUDA kernels this library has launched since it was loaded (bench.py's gpu_launches). */
long long srk_launch_count(void);

/* ---- epilogues of srk_gemm_tn (values match csrc/gemm_tn.cuh) ---- */
#define SRK_EPI_STORE 0  /* C = bf16(acc)                                                       */
#define SRK_EPI_GELU2 1  /* C = gelu(u) (bf16), C2 = gelu'(u) stored as FP16 — Mlp.act, architecture_swin.py:20; C2 is an
                            opaque 2-byte-per-element tensor whose only consumer is SRK_EPI_MUL              */
#define SRK_EPI_MUL 2    /* C = bf16(acc) * X1, X1 = the FP16 gelu' tensor of SRK_EPI_GELU2 — backward of Mlp.act */
#define SRK_EPI_RES_LN 3 /* C = acc + X1, C2 = LN(C)    — residual :149-150 + norm :127,150     */
#define SRK_EPI_LNBWD 4  /* C = X2 + LNbackward(acc)    — backward of the same                  */
#define SRK_EPI_GELU1 5  /* C = gelu(u)                 — Mlp.act when gelu'(u) is recomputed by MULG   */
#define SRK_EPI_MULG 6   /* C = (A B^T) * gelu'(X1 X2^T): two GEMMs per tile, X1 = A2 [M,K] (lda = ldx1), X2 = B2
                            [N,K]; backward of Mlp.act with u = fc1(xn2) recomputed on the tensor cores instead of
                            a stored gelu' tensor (saves 2 x [T, hidden] of HBM traffic and memory per block)   */

typedef struct SrkLnArgs {
  int n_real;         /* real channel count normalised (180 / 90)                         */
  int ones_col;       /* column forced to 1.0 in the LN / GELU output, -1 for none        */
  const float* gamma; /* LN weight [n_real]                                               */
  const float* beta;  /* LN bias [n_real] (may be NULL for LNBWD)                         */
  float* stats;       /* [M][2] mean,rstd: written by RES_LN, read by LNBWD               */
  float* partials;    /* LNBWD: [srk_gemm_grid(...)][2][N] per-CTA sums for dgamma, dbeta */
  float eps;
  const float* row_scale; /* RES_LN, optional: C = acc * row_scale[row / rows_per_scale] + X1 — stochastic depth
                             (drop_path, hat_arch.py:11-23,306-307): one factor 0 or 1/keep_prob per sample       */
  int rows_per_scale;     /* tokens per sample (H*W)                                                            */
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
 * Cb in {64,128,192,256}; T % 64 == 0, 1 <= splits <= T/64 (token ranges reduced by a second tiny kernel;
 * srk_gemm_wgrad_splits gives the count that fills the GPU); workspace >= srk_gemm_wgrad_workspace_elems floats;
 * out is [ceil(Ca/128)*128, Cb] fp32 (rows >= Ca are zero). */
int srk_gemm_wgrad(int T, int Ca, int Cb, const void* A, int lda, const void* B, int ldb, float* workspace,
                   int splits, float* out, void* stream);
int srk_gemm_wgrad_splits(int T, int Ca);
long long srk_gemm_wgrad_workspace_elems(int Ca, int Cb, int splits);

/* Fused MLP half of a block, forward: x_out = resid + drop * fc2(gelu(fc1(xn2))), xn_out = LayerNorm_next(x_out), in ONE
 * tcgen05 kernel — the hidden activation goes from the fc1 accumulator (TMEM) through GELU into shared-memory boxes that
 * are the A operand of fc2; it is written to HBM (act = gelu(u), dact = gelu'(u), both [T,Hp]) only when the pointers are
 * given (training).  Replaces Mlp.forward + residual + norm: architecture_swin.py:19-25,149-150,127; hat_arch.py:76-96,306.
 * xn2, resid, x_out, xn_out: [T,Cp] bf16; w1 = prepared fc1 operand [Hp,Cp], w2 = prepared fc2 operand [Cp,Hp];
 * T % 128 == 0, Cp == 192, Hp % 128 == 0; ln = the next LayerNorm (gamma, beta, stats, eps, row_scale as in SRK_EPI_RES_LN);
 * hid_ones_col = column of act forced to 1.0 (bias column of fc2).  act / dact may be NULL (inference). */
int srk_mlp_fwd(int T, int Cp, int Hp, const void* xn2, const void* w1, const void* w2, const void* resid, void* act,
                void* dact, void* x_out, void* xn_out, int hid_ones_col, const SrkLnArgs* ln, void* stream);

/* ======================================================================================================
 * Swin / HAT transformer-block level API (what the mirrored nn.Modules call).
 * ====================================================================================================== */

typedef struct SrkBlockDims {
  int C;      /* real channels (embed_dim: 180)                                  */
  int Cp;     /* padded channels (192); column C of normalised activations is 1.0 */
  int heads;  /* 6                                                               */
  int dh;     /* real head dim (30)                                              */
  int ds;     /* padded head slot (32)                                           */
  int hidden; /* MLP hidden (720)                                                */
  int Hp;     /* padded hidden (768); column `hidden` of the activation is 1.0    */
} SrkBlockDims;

typedef struct SrkGeom { int B, H, W, ws, shift; } SrkGeom; /* tokens T = B*H*W, row-major (b,y,x) */

/* fp32 master parameters of one block, reference shapes (SwinTransformerBlock, architecture_swin.py:113-121) */
typedef struct SrkBlockParams {
  const float *norm1_w, *norm1_b, *rpb_table, *qkv_w, *qkv_b, *proj_w, *proj_b, *norm2_w, *norm2_b, *fc1_w, *fc1_b,
      *fc2_w, *fc2_b;
} SrkBlockParams;
typedef struct SrkBlockGrads {
  float *norm1_w, *norm1_b, *rpb_table, *qkv_w, *qkv_b, *proj_w, *proj_b, *norm2_w, *norm2_b, *fc1_w, *fc1_b, *fc2_w,
      *fc2_b;
} SrkBlockGrads;
/* bf16 GEMM operands derived from the parameters (see srk_block_weight_elems for sizes) */
typedef struct SrkBlockWeights { void *qkv_f, *qkv_t, *proj_f, *proj_t, *fc1_f, *fc1_t, *fc2_f, *fc2_t; } SrkBlockWeights;

/* activations of one block (all bf16 token-major unless noted); the forward fills qkv..stats_out */
typedef struct SrkBlockActs {
  const void* x_in;  /* [T,Cp] residual stream entering the block                     */
  const void* xn1;   /* [T,Cp] LayerNorm1(x_in) (+ ones column)                       */
  const float* stats1; /* [T,2] mean,rstd of LayerNorm1 (needed by the backward only)  */
  void* qkv;         /* [T,3*heads*ds]                                                */
  void* ao;          /* [T,heads*ds] attention output before proj                     */
  void* x_mid;       /* [T,Cp] x_in + proj(ao)                                        */
  void* xn2;         /* [T,Cp] LayerNorm2(x_mid)                                      */
  float* stats2;     /* [T,2]                                                         */
  void* act;         /* [T,Hp] gelu(fc1)                                              */
  void* dact;        /* [T,Hp] gelu'(fc1) as FP16, or NULL: not stored, the backward recomputes it (SRK_EPI_MULG) */
  void* x_out;       /* [T,Cp] x_mid + fc2(act)                                       */
  void* xn_out;      /* [T,Cp] LayerNorm_next(x_out): next block's norm1 or the model's final norm */
  float* stats_out;  /* [T,2]                                                         */
} SrkBlockActs;

typedef struct SrkBlockScratch { /* backward scratch, caller-owned, reusable across blocks */
  void* d_act;     /* [T,Hp]  bf16  dU                                   */
  void* d_ao;      /* [T,heads*ds] bf16                                  */
  void* d_qkv;     /* [T,3*heads*ds] bf16                                */
  void* g_mid;     /* [T,Cp] bf16 gradient at x_mid                      */
  float* wg_ws;    /* srk_block_bwd_scratch_floats(...) floats           */
} SrkBlockScratch;

/* number of bf16 elements in the 8 operand buffers of SrkBlockWeights, in struct order */
void srk_block_weight_elems(const SrkBlockDims* d, long long out[8]);
/* floats needed for SrkBlockScratch.wg_ws */
long long srk_block_bwd_scratch_floats(const SrkBlockDims* d, const SrkGeom* g);

/* params -> bf16 operands (bias folding, head padding, q pre-scaling, transposes for dgrad) */
int srk_block_prep_weights(const SrkBlockDims* d, const SrkBlockParams* p, const SrkBlockWeights* w, void* stream);

/* Forward of one SwinTransformerBlock (architecture_swin.py:123-151) given x_in and xn1 = LN1(x_in).
 * next_norm_w/b: affine of the LayerNorm that consumes x_out (next block's norm1, or SwinIR.norm :247). */
int srk_swin_block_fwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w, const SrkBlockParams* p,
                       const float* next_norm_w, const float* next_norm_b, const SrkBlockActs* a, void* stream);
/* Backward: g_out = dL/dx_out [T,Cp] bf16 -> g_in = dL/dx_in [T,Cp] bf16 and all 13 parameter gradients
 * (written, or accumulated when accumulate != 0). */
int srk_swin_block_bwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w, const SrkBlockParams* p,
                       const SrkBlockActs* a, const void* g_out, const SrkBlockScratch* s, void* g_in,
                       const SrkBlockGrads* grads, int accumulate, void* stream);

/* Window attention core on packed qkv (shift/partition/reverse as address arithmetic), ws == 8.
 * Replaces architecture_swin.py:27-37 (partition/reverse), :130-146 (roll), :75-93 (attention math). */
int srk_win_attn_fwd(const SrkGeom* g, int heads, const void* qkv, int ld_qkv, const float* rpb_table, void* out,
                     int ld_out, int ones_col, void* stream);
/* dbias_ws: srk_win_attn_bwd_ws_floats(heads) floats; d_rpb_table [225,heads] written if non-NULL */
int srk_win_attn_bwd(const SrkGeom* g, int heads, const void* qkv, int ld_qkv, const float* rpb_table,
                     const void* d_out, int ld_out, void* d_qkv, float* dbias_ws, float* d_rpb_table, void* stream);
long long srk_win_attn_bwd_ws_floats(int heads);

/* Stand-alone LayerNorm over token-major bf16 rows (nn.LayerNorm, architecture_swin.py:113,119,221). */
int srk_layernorm_fwd(const void* x, int ldx, void* y, int ldy, float* stats, const float* gamma, const float* beta,
                      int rows, int C, int Cp, int ones_col, float eps, void* stream);
/* part_ws: srk_layernorm_bwd_ws_floats(Cp) floats */
int srk_layernorm_bwd(const void* dy, int lddy, const void* x, int ldx, const float* stats, const float* gamma,
                      const void* dres, int lddres, void* dx, int lddx, float* part_ws, float* dgamma, float* dbeta,
                      int rows, int C, int Cp, void* stream);
long long srk_layernorm_bwd_ws_floats(int Cp);

/* ======================================================================================================
 * Convolutional paths: 3x3 / stride 1 / pad 1 on NHWC bf16 activations, implicit GEMM on tcgen05.
 * Replace nn.Conv2d(3x3) + nn.LeakyReLU + nn.PixelShuffle + residual adds in SwinIR's head/tail
 * (architecture_swin.py:202,222-230,175-190) and HAT's CAB / RHAG / head (hat_arch.py:66-74,608,859-869).
 * ====================================================================================================== */
#define SRK_CEPI_BIAS 0       /* y = conv + bias                                         */
#define SRK_CEPI_BIAS_LRELU 1 /* y = leaky_relu(conv + bias, slope)                      */
#define SRK_CEPI_BIAS_RES 2   /* y = conv + bias + r                                     */
#define SRK_CEPI_MASK_LRELU 3 /* y = conv * (r > 0 ? 1 : slope)   (LeakyReLU backward)   */
#define SRK_CEPI_BIAS_GELU 4  /* y = gelu(conv + bias), y2 = gelu'(conv + bias)          */
#define SRK_CEPI_MUL 5        /* y = conv * r                     (GELU backward)        */
#define SRK_CEPI_OUT1 6       /* Cout_p == 16, one real output channel: y is fp32 [B,H,W] = conv[:,0] + bias[0]
                                 (conv_last, architecture_swin.py:230): tcgen05 N = 16 instead of a CUDA-core pass */

/* w [Cout,Cin,3,3] fp32 -> wf [Cout_p, 9*Cin_p] bf16 (forward operand), wt [Cin_p, 9*Cout_p] bf16 (flipped /
 * transposed operand of the input gradient, may be NULL), bias_packed [Cout_p] fp32 (may be NULL).
 * ps != 0: output channels are permuted so that a following PixelShuffle(2) becomes a strided store. */
int srk_conv3x3_prep_weights(const float* w, const float* bias, int Cout, int Cin, int Cout_p, int Cin_p, int ps,
                             void* wf, void* wt, float* bias_packed, void* stream);
/* y = epilogue(conv3x3(x, wk)).  x: [B,H,W,Cin_p]; if x_ps, x is a pixel-shuffled tensor [B,2H,2W,64] read as
 * 256 channels (input-gradient of a PixelShuffle layer).  y: [B,H,W,Cout_p]; if y_ps (Cout_p == 256), y is written
 * directly in pixel-shuffled form [B,2H,2W,64].  H % 8 == 0, W % 16 == 0, channels multiples of 64 (<= 256). */
int srk_conv3x3_igemm(int epi, int B, int H, int W, int Cin_p, int Cout_p, int n_real, const void* x, int x_ps,
                      const void* wk, const float* bias, float slope, void* y, int y_ps, void* y2, const void* r,
                      void* stream);
/* dw [Cout,Cin,3,3] fp32 = weight gradient; dy: [B,H,W,Cout_p] (or pixel-shuffled if ps), x: [B,H,W,Cin_p]. */
int srk_conv3x3_wgrad(int B, int H, int W, int Cin, int Cout, int Cin_p, int Cout_p, int ps, const void* dy,
                      const void* x, float* ws, float* dw, void* stream);
long long srk_conv3x3_wgrad_ws_floats(int Cin_p, int Cout_p);
/* db[n_out] = sum over pixels of dy ([B,H,W,C] bf16, or pixel-shuffled [B,2H,2W,C/4] if ps). ws: srk_small_ws_floats */
int srk_bias_grad_nhwc(const void* dy, int B, int H, int W, int C, int ps, float* ws, float* db, int n_out,
                       void* stream);
long long srk_small_ws_floats(void);
/* conv_first (1 -> C): x fp32 [B,H,W] -> y token-major bf16 [B*H*W, Cp]; and its weight/bias gradient */
int srk_conv_in1_fwd(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int Cp,
                     void* stream);
int srk_conv_in1_wgrad(const float* x, const void* dy, float* ws, float* dw, float* db, int B, int H, int W, int C,
                       int Cp, void* stream);
/* conv_last (64 -> 1): x NHWC bf16 [B,H,W,64] -> y fp32 [B,H,W]; backward gives dx (bf16), dw [1,64,3,3], db [1] */
int srk_conv_out1_fwd(const void* x, const float* w, const float* bias, float* y, int B, int H, int W, int C,
                      void* stream);
int srk_conv_out1_bwd(const float* dy, const void* x, const float* w, void* dx, float* ws, float* dw, float* db, int B,
                      int H, int W, int C, void* stream);


/* ------------------------------------------------------------------------------------------------------
 * Channel-slice ("view") variants for the dense blocks and tail of the hybrid generator
 * (models/hybridmodels_hat.py:21-131: ResidualDenseBlock / RRDBBlock / HybridHATRealESRGAN).
 * A view = (pointer to the slice's first channel, visible channels C, pixel pitch in elements) of an NHWC bf16
 * tensor; C % 8 == 0, pitch % 8 == 0, 16-byte aligned.  The implicit-GEMM kernels clip their 64-channel TMA boxes
 * to the view, so torch.cat((x, x1, ...), 1) (:40-43) is a slice of one [pixels, nf + 4*gc] buffer and never copied.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* ptr;
  int C;
  int pitch;
} SrkView;
/* y = epilogue(conv3x3(x, wk)); BIAS_RES: y = alpha * (conv + bias) + r (r may alias y); OUT1: y32 fp32 [B,H,W]. */
int srk_conv3x3_igemm_v(int epi, int B, int H, int W, int Cin_p, int Cout_p, int n_real, const SrkView* x,
                        const void* wk, const float* bias, float slope, float alpha, const SrkView* y, const SrkView* r,
                        float* y32, void* stream);
int srk_conv3x3_wgrad_v(int B, int H, int W, int Cin, int Cout, int Cin_p, int Cout_p, const SrkView* dy,
                        const SrkView* x, float* ws, float* dw, void* stream);
int srk_bias_grad_v(const SrkView* dy, long long npix, float* ws, float* db, int n_out, void* stream);
/* g *= (f > 0 ? 1 : slope): backward through nn.LeakyReLU (hybridmodels_hat.py:29), f = forward output */
/* colsum (may be NULL): [g->C] column sums of the masked gradient = bias gradient of the layer; ws: srk_small_ws_floats */
int srk_view_lrelu_mask(const SrkView* g, const SrkView* f, long long npix, float slope, float* ws, float* colsum,
                        void* stream);
/* y = alpha * a + x (x may be NULL; y may alias a or x): the 0.2-scaled residuals (:44,:58) and their gradients */
int srk_view_axpy(const SrkView* y, const SrkView* a, const SrkView* x, long long npix, float alpha, void* stream);
/* F.interpolate(scale_factor=2, mode='nearest') (:127) on NHWC views, x [B,H,W,C] -> y [B,2H,2W,C], and its adjoint */
int srk_nearest2_fwd(const SrkView* x, const SrkView* y, int B, int H, int W, void* stream);
int srk_nearest2_bwd(const SrkView* dy, const SrkView* dx, int B, int H, int W, void* stream);
/* fp32 single-channel image [npix] <-> bf16 rows of 8 channels (channel 0 = image): operand form of the 1 -> nf and
 * nf -> 1 convolutions (conv_adapt :94, conv_last :105) for the implicit-GEMM kernel */
int srk_img1_pack(const float* x, void* y8, long long npix, void* stream);
int srk_img1_unpack(const void* x8, float* y, long long npix, void* stream);

/* ======================================================================================================
 * HAT (models/hat_arch/hat_arch.py): 16x16-window attention cores, block orchestration, channel attention.
 * ====================================================================================================== */
#define SRK_ATTN_SELF 0 /* (S)W-MSA of HAB: keys = same window, shift mask 0/-100 from coordinates (:183-187,921-940) */
#define SRK_ATTN_OCA 1  /* OCAB: keys/values = 24x24 halo window, zero outside the image, no mask (:400-428)          */

/* Window attention core on packed qkv for ws == 16 (g->ws must be 16; g->shift in {0, 8} for SELF, 0 for OCA).
 * Replaces window_partition/reverse (hat_arch.py:97-126), torch.roll (:280-302), nn.Unfold + rearrange (:408-409),
 * and the attention math of WindowAttention.forward (:175-193) / OCAB.forward (:419-428).
 * rpb_table: [(16+wse-1)^2, heads] fp32 (961 / 1521 rows); lse: [heads][T] fp32 written by fwd, read by bwd. */
int srk_win_attn16_fwd(const SrkGeom* g, int mode, int heads, const void* qkv, int ld_qkv, const float* rpb_table,
                       void* out, int ld_out, float* lse, int ones_col, void* stream);
/* out: the forward output (for the softmax-gradient row term); ws: srk_win_attn16_bwd_ws_bytes(...) bytes;
 * d_rpb_table [(16+wse-1)^2, heads] written if non-NULL. */
int srk_win_attn16_bwd(const SrkGeom* g, int mode, int heads, const void* qkv, int ld_qkv, const float* rpb_table,
                       const void* out, const void* d_out, int ld_out, const float* lse, void* d_qkv, void* ws,
                       float* d_rpb_table, void* stream);
long long srk_win_attn16_bwd_ws_bytes(const SrkGeom* g, int mode, int heads);

/* What distinguishes a HAT block from a Swin block at the orchestration level. */
typedef struct SrkHatExtra {
  int mode;           /* SRK_ATTN_SELF (HAB) or SRK_ATTN_OCA (OCAB)                                                  */
  const void* res_in; /* [T,Cp] residual added by the proj epilogue: HAB x_in + conv_scale*CAB(xn1) (:306), OCAB x_in */
  float* lse;         /* [heads][T]                                                                                   */
  void* attn_ws;      /* backward only: srk_win_attn16_bwd_ws_bytes bytes                                             */
  void* d_xn1;        /* backward, optional: if non-NULL receives d_qkv @ Wqkv ([T,Cp] bf16) and the LayerNorm-1
                         backward (g_in, norm1 grads) is left to the caller, who first adds the CAB branch gradient    */
  const float* drop_attn; /* optional [B]: stochastic-depth factor (0 or 1/keep) of the attention branch per sample   */
  const float* drop_mlp;  /* optional [B]: same for the MLP branch                                                    */
  void* gs_buf;           /* backward, required when a drop_* is given: [T,Cp] bf16 scratch for the scaled gradient   */
} SrkHatExtra;

/* Forward / backward of one HAB or OCAB given xn1 = LN1(x_in): same GEMM chain as srk_swin_block_fwd/bwd
 * (HAB.forward hat_arch.py:266-309 minus the CAB branch, OCAB.forward :392-438). */
int srk_hat_block_fwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w, const SrkBlockParams* p,
                      const float* next_norm_w, const float* next_norm_b, const SrkBlockActs* a, const SrkHatExtra* x,
                      void* stream);
int srk_hat_block_bwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w, const SrkBlockParams* p,
                      const SrkBlockActs* a, const void* g_out, const SrkBlockScratch* s, void* g_in,
                      const SrkBlockGrads* grads, const SrkHatExtra* x, void* stream);

/* Channel attention of CAB (ChannelAttention hat_arch.py:40-58) on token-major bf16 y = conv2 output [B*HW, Cp]:
 *   pool[b,c] = mean_p y; hidden = relu(W1 pool + b1); scale = sigmoid(W2 hidden + b2);
 *   out = x + alpha * y * scale   (alpha = HAB.conv_scale, x = the block's shortcut; :306)
 * w1 [S,C], w2 [C,S] are the 1x1 conv weights viewed as matrices.  ws: srk_small_ws_floats() floats. */
int srk_cab_se_fwd(const void* y, const void* x, int B, int HW, int C, int Cp, int S, const float* w1, const float* b1,
                   const float* w2, const float* b2, float alpha, float* ws, float* pool, float* hidden, float* scale,
                   void* out, void* stream);
/* Backward: g = dL/dout [B*HW,Cp] -> dy (bf16 [B*HW,Cp]) and the four parameter gradients (dL/dx = g, by identity). */
int srk_cab_se_bwd(const void* g, const void* y, int B, int HW, int C, int Cp, int S, const float* w1, const float* w2,
                   float alpha, const float* pool, const float* hidden, const float* scale, float* ws, void* dy,
                   float* dw1, float* db1, float* dw2, float* db2, void* stream);

/* ======================================================================================================
 * UNetDiscriminatorSN (models/discriminator_swin.py:43-84, models/discriminator_hat.py:8-49):
 * the eight 4x4 / stride-2 / pad-1 convolutions and transposed convolutions run on srk_gemm_tn / srk_gemm_wgrad through a
 * patch matrix; the 3x3 layers reuse srk_conv_in1_*, srk_conv3x3_igemm / _wgrad and srk_conv_out1_*.
 *   Conv2d(4,2,1) + LeakyReLU          y  = srk_gemm_tn_lrelu(patches(x), Wf)          Wf [Cout, 16*Cin], k = (ky*4+kx)*Cin + ci
 *   its input gradient                 dx = fold(srk_gemm_tn(STORE, dy_pre, Wt))        Wt [16*Cin, Cout]
 *   ConvTranspose2d(4,2,1) + LeakyReLU y  = fold(srk_gemm_tn(STORE, x, Wu), LRELU)      Wu [16*Cout, Cin]
 *   its input gradient                 dx = srk_gemm_tn(STORE, patches(dy * mask), Wd)  Wd [Cin, 16*Cout]
 *   weight gradients                   srk_gemm_wgrad(patches, dy_pre) / (x, patches)
 * ====================================================================================================== */

/* C[M,N] = bf16(leaky_relu(A[M,K] * B[N,K]^T, slope)); same kernel and shape rules as srk_gemm_tn(SRK_EPI_STORE). */
#define SRK_EPI_LRELU 7
int srk_gemm_tn_lrelu(int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc, float slope,
                      void* stream);

/* patches [B*(H/2)*(W/2), 16*x->C] bf16 (contiguous) = the 4x4 / stride-2 / pad-1 patches of the NHWC view x [B,H,W,C];
 * if f != NULL the gathered value is x * (f > 0 ? 1 : slope) (LeakyReLU backward applied to a gradient image while it is
 * gathered; f = the forward activation at the same pixels).  Replaces the unfold inside nn.Conv2d(.., 4, 2, 1)
 * (discriminator_swin.py:10,52) and inside the input gradient of nn.ConvTranspose2d(.., 4, 2, 1) (:25). */
int srk_disc_patches_k4s2(const SrkView* x, const SrkView* f, float slope, int B, int H, int W, void* patches, void* stream);

#define SRK_FOLD_NONE 0  /* y = fold(taps) + add                                     */
#define SRK_FOLD_LRELU 1 /* y = leaky_relu(fold(taps) + add, slope)                  */
#define SRK_FOLD_MASK 2  /* y = (fold(taps) + add) * (f > 0 ? 1 : slope)             */
/* Fold of a tap matrix [B*Hi*Wi, 16*y->C] (bf16, contiguous) onto the NHWC view y [B,2Hi,2Wi,C]: every output pixel sums
 * the (at most four) taps that reach it — gather form, deterministic, no atomics; add (optional view, same channels) is
 * added before the activation.  This is nn.ConvTranspose2d(.., 4, 2, 1) after its GEMM (:25) and the input gradient of
 * nn.Conv2d(.., 4, 2, 1) (:10), with the skip-connection gradient of torch.cat (:40) as `add`. */
int srk_disc_fold_k4s2(const void* taps, int B, int Hi, int Wi, const SrkView* add, const SrkView* f, int act, float slope,
                       const SrkView* y, void* stream);
/* Operand forms of a 4x4 weight W [P][Q][4][4] fp32: a [P, 16*Q] bf16 with column (ky*4+kx)*Q + q, and at = a^T [16*Q, P]
 * (may be NULL).  nn.Conv2d: P = Cout, Q = Cin (a = Wf, at = Wt); nn.ConvTranspose2d: P = Cin, Q = Cout (a = Wd, at = Wu).
 * P, Q multiples of 32.  sigma (device pointer to one float, may be NULL): the operands hold W / sigma — spectral
 * normalisation folded into the packing. */
int srk_disc_prep_w4(const float* w, int P, int Q, const float* sigma, void* a, void* at, void* stream);
/* Weight gradient of a 4x4 layer in the parameter's own layout: dw [Cb][R][4][4] fp32, dw[c][r][ky][kx] =
 * sum_t A[t][(ky*4+kx)*R + r] * B[t][c]  (A [T, 16*R], B [T, Cb] bf16, T % 64 == 0; Cb 64 / 128 / 192 / 256 or a multiple of
 * 256).  nn.Conv2d: A = patches(x), B = dy_pre, R = Cin; nn.ConvTranspose2d: A = patches(dy_pre), B = x, R = Cout.
 * srk_gemm_wgrad per 256 columns of B, the token splits folded by the un-permuting kernel (no reduce launch, no copy).
 * ws: srk_disc_wgrad4_ws_floats floats. */
int srk_disc_wgrad4(int T, int R, int Cb, const void* A, int lda, const void* B, int ldb, float* ws, float* dw, void* stream);
long long srk_disc_wgrad4_ws_floats(int T, int R, int Cb);
/* y = leaky_relu(y, slope) in place on a view (nn.LeakyReLU(0.2, inplace=True) after the 1 -> nf convolution, :49-50) */
int srk_view_lrelu(const SrkView* y, long long npix, float slope, void* stream);

/* F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) on NHWC views (models/discriminator_hat.py:31,36,41):
 * y [B,2H,2W,C] = resize(x + s), s optional (the skip connection added just before the resize, :35,:40); and its adjoint
 * dx [B,H,W,C] from dy [B,2H,2W,C] (gather form, deterministic). */
int srk_bilinear2x_fwd(const SrkView* x, const SrkView* s, const SrkView* y, int B, int H, int W, void* stream);
int srk_bilinear2x_bwd(const SrkView* dy, const SrkView* dx, int B, int H, int W, void* stream);

/* torch.nn.utils.spectral_norm (discriminator_swin.py:10,25,49-52,67-69; torch/nn/utils/spectral_norm.py) on the weight
 * W [A][B][KK] fp32 (KK = kh*kw) seen as the matrix Wm [U][V]:  dim 0 (nn.Conv2d): U = A, V = B*KK;  dim 1
 * (nn.ConvTranspose2d): U = B, V = A*KK.   power_iteration != 0 (module.training):  v <- normalize(Wm^T u),
 * u <- normalize(Wm v) in place on the module's weight_u / weight_v buffers (normalize = x / max(||x||, eps));  always:
 * *sigma = u . (Wm v);  w_sn (optional) = W / sigma in fp32.  One call serves all layers of a network (n descriptors,
 * HOST array).  ws: srk_spectral_norm_ws_floats() floats. */
typedef struct SrkSnLayer {
  const float* w; /* weight_orig                                         */
  float* u;       /* weight_u [U]                                        */
  float* v;       /* weight_v [V]                                        */
  int A, B, KK, dim;
  float* sigma;   /* out: 1 float                                        */
  float* w_sn;    /* out, optional: W / sigma [A][B][KK] fp32            */
} SrkSnLayer;
int srk_spectral_norm(const SrkSnLayer* layers, int n, int power_iteration, float eps, float* ws, void* stream);
long long srk_spectral_norm_ws_floats(void);
/* Backward of W_sn = W / sigma(W) with u, v constant (what autograd derives for the hook):
 *   dw[i] = dw_sn[i] / sigma - (<dw_sn[i], W> / sigma^2) * (u v^T laid out like W);   u, v, sigma as the forward left them.
 * dw_sn / dw: HOST arrays of n device pointers (a NULL entry skips the layer; dw[i] may alias dw_sn[i]).
 * ws: srk_spectral_norm_ws_floats() floats. */
int srk_spectral_norm_bwd(const SrkSnLayer* layers, int n, const float* const* dw_sn, float* const* dw, float* ws, void* stream);

/* ======================================================================================================
 * Data formats on either side of the path: 16-bit image planes.
 * ====================================================================================================== */

/* dst[b] (float32, n x n) = augment(src[b] (uint16, n x n)) / 65535: the dataset's load + augmentation in one pass.
 * Replaces dataset/astronomical_dataset_swin.py:34-39 (/65535 -> float32) and :58-67 (flip(-1), flip(-2), rot90(k)).
 * codes: device int[B], per sample  fh | fv << 1 | k << 2  (NULL: no augmentation); n % 32 == 0. */
int srk_u16_to_f32_aug(const void* src_u16, float* dst, const int* codes, int B, int n, void* stream);

/* dst (uint16) = trunc(clip(src, 0, 1) * 65535): the quantisation of save_as_tiff16 (infer_hat.py:42-50), on the device,
 * so that a super-resolved frame leaves the GPU at 2 bytes per pixel.  src 16-byte, dst 8-byte aligned. */
int srk_f32_to_u16(const float* src, void* dst_u16, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRK_H_ */
