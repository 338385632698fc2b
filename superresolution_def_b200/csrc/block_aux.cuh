// block_aux.cuh — small memory-bound kernels around the Swin block GEMMs:
//   * weight preparation (fp32 master params -> padded bf16 GEMM operands, bias folded into a "ones" column)
//   * gradient unpacking (padded fp32 wgrad results + per-CTA partials -> reference-shaped fp32 grads)
//   * standalone LayerNorm forward / backward over token-major bf16 rows
// Layout conventions are documented in DESIGN.md ("HBM layouts").
#pragma once
#include "srk_ptx.cuh"

namespace srk {

struct BlockDims {
  int C;       // real channels (180)
  int Cp;      // padded channels (192), ones column at index C
  int heads;   // 6
  int dh;      // real head dim (30)
  int ds;      // head slot (32); attention-output ones column at index dh (head 0)
  int hidden;  // 720
  int Hp;      // padded hidden (768), ones column at index hidden
  int QW() const { return 3 * heads * ds; }
  int AW() const { return heads * ds; }
};

struct BlockParamPtrs {  // fp32 master parameters, reference shapes (models/architecture_swin.py:113-121)
  const float* norm1_w; const float* norm1_b;
  const float* rpb_table;                       // [(2ws-1)^2, heads]
  const float* qkv_w; const float* qkv_b;       // [3C, C], [3C]
  const float* proj_w; const float* proj_b;     // [C, C], [C]
  const float* norm2_w; const float* norm2_b;
  const float* fc1_w; const float* fc1_b;       // [hidden, C], [hidden]
  const float* fc2_w; const float* fc2_b;       // [C, hidden], [C]
};

struct BlockGradPtrs {  // fp32 gradients, same shapes
  float* norm1_w; float* norm1_b; float* rpb_table; float* qkv_w; float* qkv_b; float* proj_w; float* proj_b;
  float* norm2_w; float* norm2_b; float* fc1_w; float* fc1_b; float* fc2_w; float* fc2_b;
};

struct BlockWeightPtrs {  // prepared bf16 operands
  __nv_bfloat16* qkv_f;   // [QW, Cp]   forward  (rows: s,h,d slots; col C = bias; q rows pre-scaled)
  __nv_bfloat16* qkv_t;   // [Cp, QW]   dgrad    (transposed, bias row zero)
  __nv_bfloat16* proj_f;  // [Cp, AW]   forward  (col dh = bias)
  __nv_bfloat16* proj_t;  // [AW, Cp]   dgrad
  __nv_bfloat16* fc1_f;   // [Hp, Cp]   forward  (col C = bias)
  __nv_bfloat16* fc1_t;   // [Cp, Hp]   dgrad
  __nv_bfloat16* fc2_f;   // [Cp, Hp]   forward  (col hidden = bias)
  __nv_bfloat16* fc2_t;   // [Hp, Cp]   dgrad
};

// ------------------------------------------------------------------ weight preparation
static __global__ void prep_block_weights_kernel(BlockDims d, BlockParamPtrs p, BlockWeightPtrs w) {
  pdl_launch_dependents();
  pdl_wait();
  const int QW = 3 * d.heads * d.ds, AW = d.heads * d.ds;
  const int n0 = QW * d.Cp, n1 = d.Cp * AW, n2 = d.Hp * d.Cp, n3 = d.Cp * d.Hp;
  const int total = n0 + n1 + n2 + n3;
  const float qscale = rsqrtf(float(d.dh));
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    if (idx < n0) {  // qkv: ext row r = (s, h, dd), col c
      const int r = idx / d.Cp, c = idx % d.Cp;
      const int s = r / AW, h = (r % AW) / d.ds, dd = r % d.ds;
      float v = 0.f, vt = 0.f;
      if (dd < d.dh) {
        const int n = s * d.C + h * d.dh + dd;
        const float sc = (s == 0) ? qscale : 1.f;
        if (c < d.C) { v = p.qkv_w[n * d.C + c] * sc; vt = v; }
        else if (c == d.C) { v = p.qkv_b[n] * sc; }
      }
      w.qkv_f[idx] = __float2bfloat16_rn(v);
      w.qkv_t[c * QW + r] = __float2bfloat16_rn(vt);
    } else if (idx < n0 + n1) {  // proj: row n, ext col c = (h, dd)
      const int i = idx - n0, n = i / AW, c = i % AW;
      const int h = c / d.ds, dd = c % d.ds;
      float v = 0.f, vt = 0.f;
      if (n < d.C) {
        if (dd < d.dh) { v = p.proj_w[n * d.C + h * d.dh + dd]; vt = v; }
        else if (h == 0 && dd == d.dh) { v = p.proj_b[n]; }
      }
      w.proj_f[i] = __float2bfloat16_rn(v);
      w.proj_t[c * d.Cp + n] = __float2bfloat16_rn(vt);
    } else if (idx < n0 + n1 + n2) {  // fc1: row n (hidden), col c
      const int i = idx - n0 - n1, n = i / d.Cp, c = i % d.Cp;
      float v = 0.f, vt = 0.f;
      if (n < d.hidden) {
        if (c < d.C) { v = p.fc1_w[n * d.C + c]; vt = v; }
        else if (c == d.C) { v = p.fc1_b[n]; }
      }
      w.fc1_f[i] = __float2bfloat16_rn(v);
      w.fc1_t[c * d.Hp + n] = __float2bfloat16_rn(vt);
    } else {  // fc2: row n (C), col c (hidden)
      const int i = idx - n0 - n1 - n2, n = i / d.Hp, c = i % d.Hp;
      float v = 0.f, vt = 0.f;
      if (n < d.C) {
        if (c < d.hidden) { v = p.fc2_w[n * d.hidden + c]; vt = v; }
        else if (c == d.hidden) { v = p.fc2_b[n]; }
      }
      w.fc2_f[i] = __float2bfloat16_rn(v);
      w.fc2_t[c * d.Cp + n] = __float2bfloat16_rn(vt);
    }
  }
}

// ------------------------------------------------------------------ gradient unpacking
struct SplitSum {          // a weight-gradient GEMM result still in per-split form: value(i) = sum_k p[k * stride + i]
  const float* p;
  int splits;
  long long stride;
  __device__ __forceinline__ float at(long long i) const {   // eight loads in flight per thread (the partials sit in L2)
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* q = p + i;
    int k = 0;
    for (; k + 8 <= splits; k += 8, q += 8 * stride) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += __ldg(q + (long long)j * stride);
    }
    for (; k < splits; ++k, q += stride) a[0] += __ldg(q);
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
};
struct UnpackSrc {
  SplitSum dqkv_ext;   // [ceil128(QW), Cp]  rows = qkv ext rows, cols = xn1 channels (col C = bias grad)
  SplitSum dproj_ext;  // [ceil128(Cp), AW]  rows = out channels n, cols = ao ext channels (col dh = bias grad)
  SplitSum dfc1_ext;   // [Hp, Cp]           rows = hidden, cols = xn2 channels (col C = bias grad)
  SplitSum dfc2T_ext;  // [Hp, Cp]           rows = hidden (row `hidden` = bias grad), cols = out channels
  const float* ln1_part;   // [n_ln_part][2][Cp] (dgamma, dbeta) partials of norm1
  const float* ln2_part;   // same for norm2
  const float* rpb_part;   // [n_rpb_part][heads][T2] partials of the bias-table gradient
  int n_ln_part, n_rpb_part, table_rows;  // table_rows = (2ws-1)^2
};

static __global__ void unpack_block_grads_kernel(BlockDims d, UnpackSrc s, BlockGradPtrs g, float accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int C = d.C, AW = d.heads * d.ds;
  const int n_qkv_w = 3 * C * C, n_qkv_b = 3 * C, n_proj_w = C * C, n_proj_b = C;
  const int n_fc1_w = d.hidden * C, n_fc1_b = d.hidden, n_fc2_w = C * d.hidden, n_fc2_b = C;
  const int n_ln = 4 * C, n_rpb = s.table_rows * d.heads;
  const int total = n_qkv_w + n_qkv_b + n_proj_w + n_proj_b + n_fc1_w + n_fc1_b + n_fc2_w + n_fc2_b + n_ln + n_rpb;
  const float qscale = rsqrtf(float(d.dh));
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int i = idx;
    float v;
    float* dst;
    if (i < n_qkv_w + n_qkv_b) {
      const bool is_b = i >= n_qkv_w;
      const int n = is_b ? (i - n_qkv_w) : (i / C), c = is_b ? C : (i % C);
      const int sidx = n / C, h = (n % C) / d.dh, dd = n % d.dh;
      const int r = sidx * AW + h * d.ds + dd;
      v = s.dqkv_ext.at(r * d.Cp + c) * (sidx == 0 ? qscale : 1.f);
      dst = is_b ? (g.qkv_b + n) : (g.qkv_w + i);
    } else if ((i -= n_qkv_w + n_qkv_b) < n_proj_w + n_proj_b) {
      const bool is_b = i >= n_proj_w;
      const int n = is_b ? (i - n_proj_w) : (i / C), k = is_b ? 0 : (i % C);
      const int c = is_b ? d.dh : ((k / d.dh) * d.ds + (k % d.dh));
      v = s.dproj_ext.at(n * AW + c);
      dst = is_b ? (g.proj_b + n) : (g.proj_w + i);
    } else if ((i -= n_proj_w + n_proj_b) < n_fc1_w + n_fc1_b) {
      const bool is_b = i >= n_fc1_w;
      const int n = is_b ? (i - n_fc1_w) : (i / C), c = is_b ? C : (i % C);
      v = s.dfc1_ext.at(n * d.Cp + c);
      dst = is_b ? (g.fc1_b + n) : (g.fc1_w + i);
    } else if ((i -= n_fc1_w + n_fc1_b) < n_fc2_w + n_fc2_b) {
      const bool is_b = i >= n_fc2_w;
      const int n = is_b ? (i - n_fc2_w) : (i % C), k = is_b ? d.hidden : (i / C);   // source order: reads coalesced
      v = s.dfc2T_ext.at(k * d.Cp + n);
      dst = is_b ? (g.fc2_b + n) : (g.fc2_w + n * d.hidden + k);
    } else if ((i -= n_fc2_w + n_fc2_b) < n_ln) {
      const int which = i / C, c = i % C;  // 0: norm1_w, 1: norm1_b, 2: norm2_w, 3: norm2_b
      const float* part = (which < 2) ? s.ln1_part : s.ln2_part;
      if (part == nullptr) continue;  // LayerNorm-1 backward done by the caller (HAB)
      float acc = 0.f;
      for (int k = 0; k < s.n_ln_part; ++k) acc += part[(size_t(k) * 2 + (which & 1)) * d.Cp + c];
      v = acc;
      dst = (which == 0 ? g.norm1_w : which == 1 ? g.norm1_b : which == 2 ? g.norm2_w : g.norm2_b) + c;
    } else {
      i -= n_ln;
      if (s.rpb_part == nullptr) continue;  // table gradient written by the attention core's own finish kernel
      const int t = i / d.heads, h = i % d.heads;  // reference layout [table_rows, heads]
      float acc = 0.f;
      for (int k = 0; k < s.n_rpb_part; ++k) acc += s.rpb_part[(size_t(k) * d.heads + h) * s.table_rows + t];
      v = acc;
      dst = g.rpb_table + i;
    }
    *dst = (accumulate != 0.f) ? (*dst + v) : v;
  }
}

// ------------------------------------------------------------------ per-sample row scaling (stochastic depth backward)
// out[r][:] = in[r][:] * scale[r / rows_per_scale]   (bf16 rows of `cols` elements, cols % 8 == 0)
static __global__ void row_scale_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                        const float* __restrict__ scale, long long rows, int cols, int rows_per_scale) {
  const int groups = cols / 8;
  const long long total = rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const float sc = scale[r / rows_per_scale];
    const uint4 v = *reinterpret_cast<const uint4*>(in + i * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = pack_bf16(bf16_lo(w[e]) * sc, bf16_hi(w[e]) * sc);
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------ standalone LayerNorm (warp per row)
// y = LN(x[:, :C]) * gamma + beta, y[:, C] = 1 (ones column, if ones_col >= 0), other pads 0; stats = (mean, rstd)
static __global__ void ln_fwd_rows_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y,
                                   int ldy, float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int rows, int C, int Cp, int ones_col, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const __nv_bfloat16* xr = x + size_t(warp) * ldx;
  float v[8];  // Cp <= 256
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = lane + 32 * k;
    v[k] = (c < C) ? __bfloat162float(xr[c]) : 0.f;
    sum += v[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / float(C);
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = lane + 32 * k;
    const float dlt = v[k] - mean;
    if (c < C) var += dlt * dlt;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / float(C) + eps);
  if (lane == 0 && stats) reinterpret_cast<float2*>(stats)[warp] = make_float2(mean, rstd);
  __nv_bfloat16* yr = y + size_t(warp) * ldy;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = lane + 32 * k;
    if (c < Cp) {
      float o = 0.f;
      if (c < C) o = (v[k] - mean) * rstd * gamma[c] + beta[c];
      else if (c == ones_col) o = 1.f;
      yr[c] = __float2bfloat16_rn(o);
    }
  }
}

// dx = (dres ? dres : 0) + LNbackward(dy | x, stats, gamma); partial dgamma/dbeta per CTA: part[blockIdx][2][Cp]
static __global__ void ln_bwd_rows_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ x,
                                   int ldx, const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const __nv_bfloat16* __restrict__ dres, int lddres, __nv_bfloat16* __restrict__ dx,
                                   int lddx, float* __restrict__ part, int rows, int C, int Cp) {
  // warp per row, lane l owns channels 8l .. 8l+7 (one 16-byte access per tensor and row; Cp <= 256)
  __shared__ float s_acc[2][256];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int c0 = lane * 8;
  const bool has = c0 < Cp;
  float g8[8], ag[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ag[k] = ab[k] = 0.f;
    g8[k] = (c0 + k < C) ? gamma[c0 + k] : 0.f;   // gamma = 0 on pad columns: they drop out of every sum below
  }
  const float inv_c = 1.0f / float(C);
  const int row_step = gridDim.x * warps_per_block;
  int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  // software pipeline: the next row's three 16-byte loads are in flight while this row is reduced and stored
  uint4 nx = make_uint4(0u, 0u, 0u, 0u), nd = nx, nr = nx;
  float2 nst = make_float2(0.f, 0.f);
  auto fetch = [&](int r) {
    if (r < rows) {
      nst = reinterpret_cast<const float2*>(stats)[r];
      if (has) {
        nx = *reinterpret_cast<const uint4*>(x + size_t(r) * ldx + c0);
        nd = *reinterpret_cast<const uint4*>(dy + size_t(r) * lddy + c0);
        if (dres) nr = *reinterpret_cast<const uint4*>(dres + size_t(r) * lddres + c0);
      }
    }
  };
  fetch(row);
  for (; row < rows; row += row_step) {
    const float2 st = nst;
    const uint4 vx = nx, vd = nd, vr = nr;
    fetch(row + row_step);
    const uint32_t wx[4] = {vx.x, vx.y, vx.z, vx.w}, wd[4] = {vd.x, vd.y, vd.z, vd.w}, wr[4] = {vr.x, vr.y, vr.z, vr.w};
    float xh[8], dn[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xv = (k & 1) ? bf16_hi(wx[k >> 1]) : bf16_lo(wx[k >> 1]);
      const bool real = c0 + k < C;
      xh[k] = real ? (xv - st.x) * st.y : 0.f;
      dn[k] = real ? ((k & 1) ? bf16_hi(wd[k >> 1]) : bf16_lo(wd[k >> 1])) : 0.f;
      const float dh = dn[k] * g8[k];
      s1 += dh;
      s2 += dh * xh[k];
      ag[k] += dn[k] * xh[k];
      ab[k] += dn[k];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float c1 = s1 * inv_c, c2 = s2 * inv_c;
    if (has) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o[k] = 0.f;
        if (c0 + k < C) {
          o[k] = round_bf16(st.y * (dn[k] * g8[k] - c1 - xh[k] * c2));
          if (dres) o[k] += (k & 1) ? bf16_hi(wr[k >> 1]) : bf16_lo(wr[k >> 1]);
        }
      }
      *reinterpret_cast<uint4*>(dx + size_t(row) * lddx + c0) =
          make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
  if (has) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&s_acc[0][c0 + k], ag[k]);
      atomicAdd(&s_acc[1][c0 + k], ab[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cp; i += blockDim.x) {
    const int w = i / Cp, c = i % Cp;
    part[(size_t(blockIdx.x) * 2 + w) * Cp + c] = s_acc[w][c];
  }
}

// out[which][c] = sum_k part[k][which][c]  (tiny finishing reduction for the standalone LN backward)
static __global__ void ln_param_grad_reduce_kernel(const float* __restrict__ part, int nparts, int Cp, int C,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  // one warp per output element: lanes stride over the per-CTA partials, then a shuffle reduction
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= 2 * C) return;
  const int w = i / C, c = i % C;
  float acc = 0.f;
  for (int k = lane; k < nparts; k += 32) acc += part[(size_t(k) * 2 + w) * Cp + c];
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) (w == 0 ? dgamma : dbeta)[c] = acc;
}

}  // namespace srk
