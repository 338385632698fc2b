// disc_aux.cuh — patch gather / scatter-free fold for the 4x4, stride-2, pad-1 convolutions and transposed convolutions of
// UNetDiscriminatorSN (models/discriminator_swin.py:10, :25, :52; models/discriminator_hat.py:8-49), so that all eight of
// them run on the persistent tcgen05 GEMM (gemm_tn.cuh) and their weight gradients on the MN-major tcgen05 GEMM
// (gemm_wgrad.cuh):
//
//   Conv2d(4,2,1)           y  = lrelu( patches(x) @ Wf^T )          patches: [B*H/2*W/2, 16*C], k = (ky*4 + kx)*C + c
//   its input gradient      dx = fold( dy_pre @ Wt^T )               fold: every input pixel sums its (at most) 4 taps
//   ConvTranspose2d(4,2,1)  y  = lrelu( fold( x @ Wu^T ) )           (the transposed convolution IS the fold of a GEMM)
//   its input gradient      dx = patches(dy_pre) @ Wd^T
//
// Both helpers are pure HBM streams (16-byte vectors, channel-contiguous): the patch matrix is 4x the activation it is
// gathered from, which at the discriminator's sizes (<= 268 MB at 512^2 x 2) costs less than the launch of a dedicated
// implicit-GEMM kernel family would save; LeakyReLU and its backward mask ride on the passes that exist anyway.
#pragma once
#include "srk_ptx.cuh"

namespace srk {

__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { f[2 * e] = bf16_lo(w[e]); f[2 * e + 1] = bf16_hi(w[e]); }
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// patches[m][(ky*4+kx)*C + c] = g[b][2*oy-1+ky][2*ox-1+kx][c]  (zero outside the image), m = (b*Ho + oy)*Wo + ox,
// g = x, or x * (f > 0 ? 1 : slope) when f is given (LeakyReLU backward applied while gathering a gradient image).
// One thread per (patch row m, 8-channel group): the index decomposition is paid once for sixteen 16-byte vectors, whose
// loads are independent (16 in flight per thread) and whose stores are C*2-byte contiguous runs across the channel lanes.
// (First version: one thread per vector with 64-bit divisions — 140 us of integer work for the 268 MB patch matrix of the
// first stride-2 layer against 55 us of HBM time; 32-bit indices halved the kernel, this form removes the rest.)
template <typename Idx>
static __global__ void __launch_bounds__(256) disc_patches_k4s2_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                                       const __nv_bfloat16* __restrict__ f, int ldf, float slope,
                                                                       int B, int H, int W, int C,
                                                                       __nv_bfloat16* __restrict__ patches) {
  const Idx groups = Idx(C >> 3);
  const Idx Ho = Idx(H >> 1), Wo = Idx(W >> 1);
  const Idx total = Idx(B) * Ho * Wo * groups;
  for (Idx idx = Idx(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += Idx(gridDim.x) * blockDim.x) {
    const Idx m = idx / groups;
    const int cg = int(idx - m * groups);
    const Idx r = m / Wo;
    const int ox = int(m - r * Wo);
    const int oy = int(r % Ho);
    const int b = int(r / Ho);
    const int iy0 = 2 * oy - 1, ix0 = 2 * ox - 1;
    const __nv_bfloat16* xb = x + ((long long)b * H * W) * ldx + cg * 8;
    const __nv_bfloat16* fb = f ? f + ((long long)b * H * W) * ldf + cg * 8 : nullptr;
    uint4 v[16];
#pragma unroll
    for (int tap = 0; tap < 16; ++tap) {
      const int iy = iy0 + (tap >> 2), ix = ix0 + (tap & 3);
      v[tap] = make_uint4(0u, 0u, 0u, 0u);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v[tap] = *reinterpret_cast<const uint4*>(xb + (long long)(iy * W + ix) * ldx);
    }
    if (fb != nullptr) {
#pragma unroll
      for (int tap = 0; tap < 16; ++tap) {
        const int iy = iy0 + (tap >> 2), ix = ix0 + (tap & 3);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const uint4 fv = *reinterpret_cast<const uint4*>(fb + (long long)(iy * W + ix) * ldf);
          float xv[8], fw[8];
          bf16x8_to_f32(v[tap], xv);
          bf16x8_to_f32(fv, fw);
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[e] = fw[e] > 0.f ? xv[e] : xv[e] * slope;
          v[tap] = f32_to_bf16x8(xv);
        }
      }
    }
    __nv_bfloat16* dst = patches + (long long)m * 16 * C + cg * 8;
#pragma unroll
    for (int tap = 0; tap < 16; ++tap) *reinterpret_cast<uint4*>(dst + tap * C) = v[tap];
  }
}

// Fold of a [B*Hi*Wi, 16*C] tap matrix onto the [B, 2Hi, 2Wi, C] image (gather form, no atomics):
//   y[b][Y][X][c] = sum over (ky, kx) with (Y+1-ky), (X+1-kx) even and in range of taps[(b, (Y+1-ky)/2, (X+1-kx)/2)][(ky*4+kx)*C + c]
//   then  + add (optional)  then  act: 0 none, 1 LeakyReLU(slope), 2 multiply by (f > 0 ? 1 : slope)  (LeakyReLU backward).
enum { DISC_ACT_NONE = 0, DISC_ACT_LRELU = 1, DISC_ACT_MASK = 2 };
template <typename Idx>
static __global__ void __launch_bounds__(256) disc_fold_k4s2_kernel(const __nv_bfloat16* __restrict__ taps, int B, int Hi, int Wi, int C,
                                                                    const __nv_bfloat16* __restrict__ add, int ldadd,
                                                                    const __nv_bfloat16* __restrict__ f, int ldf, int act, float slope,
                                                                    __nv_bfloat16* __restrict__ y, int ldy) {
  const Idx groups = Idx(C >> 3);
  const Idx Ho = Idx(2 * Hi), Wo = Idx(2 * Wi);
  const Idx total = Idx(B) * Ho * Wo * groups;
  const long long row = 16LL * C;   // elements per row of the tap matrix
  for (Idx idx = Idx(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += Idx(gridDim.x) * blockDim.x) {
    const int cg = int(idx % groups);
    const Idx pixi = idx / groups;
    const long long pix = (long long)pixi;
    const int X = int(pixi % Wo);
    const Idx rowi = pixi / Wo;
    const int Y = int(rowi % Ho);
    const int b = int(rowi / Ho);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int ky0 = (Y + 1) & 1, kx0 = (X + 1) & 1;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ky0 + 2 * a;
      const int ny = Y + 1 - ky;          // even by construction; may be -2 at the top border
      const int iy = ny >> 1;
      if (ny < 0 || iy >= Hi) continue;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int kx = kx0 + 2 * c2;
        const int nx = X + 1 - kx;
        const int ix = nx >> 1;
        if (nx < 0 || ix >= Wi) continue;
        const long long m = ((long long)b * Hi + iy) * Wi + ix;
        const uint4 v = *reinterpret_cast<const uint4*>(taps + m * row + (long long)(ky * 4 + kx) * C + cg * 8);
        float fv[8];
        bf16x8_to_f32(v, fv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += fv[e];
      }
    }
    if (add != nullptr) {
      float av[8];
      bf16x8_to_f32(*reinterpret_cast<const uint4*>(add + pix * ldadd + cg * 8), av);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += av[e];
    }
    if (act == DISC_ACT_LRELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = acc[e] > 0.f ? acc[e] : acc[e] * slope;
    } else if (act == DISC_ACT_MASK) {
      float fw[8];
      bf16x8_to_f32(*reinterpret_cast<const uint4*>(f + pix * ldf + cg * 8), fw);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fw[e] > 0.f ? acc[e] : acc[e] * slope;
    }
    *reinterpret_cast<uint4*>(y + pix * ldy + cg * 8) = f32_to_bf16x8(acc);
  }
}

// y[:, :C] = leaky_relu(y[:, :C], slope) in place (after the 1 -> nf convolution, discriminator_swin.py:49-50)
static __global__ void __launch_bounds__(256) view_lrelu_kernel(__nv_bfloat16* __restrict__ y, int ldy, int C, long long npix, float slope) {
  const int groups = C >> 3;
  const long long total = npix * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / groups;
    const int c = int(i - p * groups) * 8;
    uint4* yp = reinterpret_cast<uint4*>(y + p * ldy + c);
    float v[8];
    bf16x8_to_f32(*yp, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * slope;
    *yp = f32_to_bf16x8(v);
  }
}

}  // namespace srk

namespace srk {

// Operand form of a 4x4 weight: W [P][Q][4][4] fp32 -> A [P][16*Q] bf16 with column k = (ky*4+kx)*Q + q.
// (Conv2d: P = Cout, Q = Cin, A = Wf; ConvTranspose2d: P = Cin, Q = Cout, A = Wd.)  One thread per (p, q): 64 contiguous
// bytes in, sixteen 2-byte stores that are contiguous across the warp for each tap.
static __global__ void __launch_bounds__(256) disc_prep_w4_kernel(const float* __restrict__ w, int P, int Q, const float* __restrict__ sigma,
                                                                  __nv_bfloat16* __restrict__ a) {
  const float sc = sigma ? 1.f / *sigma : 1.f;   // spectral normalisation folded into the packing: W / sigma
  const long long total = (long long)P * Q;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = int(i % Q);
    const long long p = i / Q;
    const float4* src = reinterpret_cast<const float4*>(w + i * 16);
    __nv_bfloat16* dst = a + p * 16 * Q + q;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = src[j];
      dst[(long long)(4 * j + 0) * Q] = __float2bfloat16_rn(v.x * sc);
      dst[(long long)(4 * j + 1) * Q] = __float2bfloat16_rn(v.y * sc);
      dst[(long long)(4 * j + 2) * Q] = __float2bfloat16_rn(v.z * sc);
      dst[(long long)(4 * j + 3) * Q] = __float2bfloat16_rn(v.w * sc);
    }
  }
}

// out [C][R] = in [R][C]^T (bf16), 32 x 32 tiles through padded shared memory; R, C multiples of 32.
static __global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int R, int C,
                                                                    __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int tiles_c = C / 32;
  const long long ntiles = (long long)(R / 32) * tiles_c;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int r0 = int(t / tiles_c) * 32, c0 = int(t % tiles_c) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[ty + 8 * j][tx] = in[(long long)(r0 + ty + 8 * j) * C + c0 + tx];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(long long)(c0 + ty + 8 * j) * R + r0 + tx] = tile[tx][ty + 8 * j];
    __syncthreads();
  }
}

// Weight gradient of a 4x4 layer out of the split partials of the MN-major GEMM:
//   dw[(c * R + rr) * 16 + tap] = sum_s part[s][tap * R + rr][c - c0]      c in [c0, c0 + cb)
// (Conv2d: rows = patch columns (tap, ci), c = co -> dW[co][ci][ky][kx]; ConvTranspose2d: rows = (tap, co), c = ci ->
//  dW[ci][co][ky][kx]: the same index map with R = Cin resp. Cout.)  A thread owns one (c, rr) pair: 16 taps = 64
// contiguous output bytes; across the warp c is the fast index, so the partial reads are contiguous.
static __global__ void __launch_bounds__(256) disc_unpack_wgrad4_kernel(const float* __restrict__ part, int splits, long long split_stride,
                                                                        int R, int cb, int c0, float* __restrict__ dw) {
  // one thread per (c, rr, group of four taps): c is the fast index (contiguous partial reads), four times the threads of a
  // (c, rr) mapping — the first version ran 64-512 blocks with 16 x splits dependent-latency loads each (6-17 % of DRAM
  // throughput under ncu)
  const long long total = 4LL * R * cb;
  const long long tap_stride = (long long)R * cb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % cb);
    const long long t = i / cb;
    const int rr = int(t % R);
    const int q = int(t / R);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* p0 = part + (long long)(4 * q) * tap_stride + (long long)rr * cb + c;
    for (int s = 0; s < splits; ++s) {
      const float* ps = p0 + s * split_stride;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += ps[k * tap_stride];
    }
    *reinterpret_cast<float4*>(dw + ((long long)(c0 + c) * R + rr) * 16 + 4 * q) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

}  // namespace srk

namespace srk {

// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) on NHWC bf16 (models/discriminator_hat.py:31,36,41):
// output row Y reads input rows (i0, i1) with weights (w0, 1 - w0):  src = Y/2 - 0.25 clamped at 0, i1 clamped at H-1, i.e.
//   Y = 2i:   0.25 x[max(i-1, 0)] + 0.75 x[i]          Y = 2i+1:   0.75 x[i] + 0.25 x[min(i+1, H-1)]
__device__ __forceinline__ void bil2_taps(int Y, int H, int& i0, int& i1, float& w0) {
  const int i = Y >> 1;
  if (Y & 1) { i0 = i; i1 = (i + 1 < H) ? i + 1 : H - 1; w0 = 0.75f; }
  else { i0 = (i > 0) ? i - 1 : 0; i1 = i; w0 = 0.25f; }
}
// weight of input index i in output index Y (0 when Y is outside [0, 2H) or does not read i)
__device__ __forceinline__ float bil2_weight(int Y, int i, int H) {
  if (Y < 0 || Y >= 2 * H) return 0.f;
  int i0, i1;
  float w0;
  bil2_taps(Y, H, i0, i1, w0);
  return (i0 == i ? w0 : 0.f) + (i1 == i ? 1.f - w0 : 0.f);
}

// y [B,2H,2W,C] = bilinear2x(x + s)   (s optional: the skip connection added before the resize, discriminator_hat.py:35,40)
static __global__ void __launch_bounds__(256) bilinear2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ s,
                                                                    int lds, __nv_bfloat16* __restrict__ y, int ldy, int B, int H, int W, int C) {
  const unsigned groups = unsigned(C >> 3);
  const unsigned Wo = 2u * W, Ho = 2u * H;
  const unsigned long long total = (unsigned long long)B * Ho * Wo * groups;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned cg = unsigned(idx % groups);
    const unsigned long long pix = idx / groups;
    const int X = int(pix % Wo);
    const unsigned long long rowi = pix / Wo;
    const int Y = int(rowi % Ho);
    const long long b = (long long)(rowi / Ho);
    int y0, y1, x0, x1;
    float wy, wx;
    bil2_taps(Y, H, y0, y1, wy);
    bil2_taps(X, W, x0, x1, wx);
    const long long base = b * H * W;
    const long long p[4] = {base + (long long)y0 * W + x0, base + (long long)y0 * W + x1, base + (long long)y1 * W + x0,
                            base + (long long)y1 * W + x1};
    const float wt[4] = {wy * wx, wy * (1.f - wx), (1.f - wy) * wx, (1.f - wy) * (1.f - wx)};
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      bf16x8_to_f32(*reinterpret_cast<const uint4*>(x + p[k] * ldx + cg * 8), v);
      if (s != nullptr) {
        float sv[8];
        bf16x8_to_f32(*reinterpret_cast<const uint4*>(s + p[k] * lds + cg * 8), sv);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += sv[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt[k], v[e], acc[e]);
    }
    *reinterpret_cast<uint4*>(y + (long long)pix * ldy + cg * 8) = f32_to_bf16x8(acc);
  }
}

// dx [B,H,W,C] = adjoint of bilinear2x applied to dy [B,2H,2W,C]  (gather form: <= 4 x 4 output pixels read an input pixel)
static __global__ void __launch_bounds__(256) bilinear2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, __nv_bfloat16* __restrict__ dx,
                                                                    int lddx, int B, int H, int W, int C) {
  const unsigned groups = unsigned(C >> 3);
  const unsigned long long total = (unsigned long long)B * H * W * groups;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned cg = unsigned(idx % groups);
    const unsigned long long pix = idx / groups;
    const int j = int(pix % unsigned(W));
    const unsigned long long rowi = pix / unsigned(W);
    const int i = int(rowi % unsigned(H));
    const long long b = (long long)(rowi / unsigned(H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = -1; a <= 2; ++a) {
      const int Y = 2 * i + a;
      const float wy = bil2_weight(Y, i, H);
      if (wy == 0.f) continue;
#pragma unroll
      for (int c = -1; c <= 2; ++c) {
        const int X = 2 * j + c;
        const float wx = bil2_weight(X, j, W);
        if (wx == 0.f) continue;
        float v[8];
        bf16x8_to_f32(*reinterpret_cast<const uint4*>(dy + ((b * 2 * H + Y) * 2 * W + X) * lddy + cg * 8), v);
        const float wgt = wy * wx;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, v[e], acc[e]);
      }
    }
    *reinterpret_cast<uint4*>(dx + (long long)pix * lddx + cg * 8) = f32_to_bf16x8(acc);
  }
}

}  // namespace srk
