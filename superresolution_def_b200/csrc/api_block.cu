// api_block.cu — C-ABI entry points at Swin-block granularity: weight prep, block forward / backward
// (a fixed sequence of the tcgen05 GEMMs, the window-attention core and the small aux kernels), plus the
// stand-alone window-attention and LayerNorm ops.  Host code only orchestrates launches on the caller's
// stream; it never allocates or synchronises.
#include "attn_oca8.cuh"
#include "attn_tc8.cuh"
#include "attn_tc16.cuh"
#include "cab_aux.cuh"
#include "block_aux.cuh"
#include "srk_host.h"
#include <cstdlib>

using namespace srk;

namespace {

inline BlockDims to_dims(const SrkBlockDims* d) {
  BlockDims o;
  o.C = d->C; o.Cp = d->Cp; o.heads = d->heads; o.dh = d->dh; o.ds = d->ds; o.hidden = d->hidden; o.Hp = d->Hp;
  return o;
}

int check_dims(const SrkBlockDims* d, const SrkGeom* g, bool hat = false) {
  if (!d) return fail(SRK_ERR_ARG, "null dims");
  if (d->Cp != 192 || d->heads * d->ds != 192 || d->ds != 32)
    return fail(SRK_ERR_UNSUPPORTED, "block kernels are specialised for Cp == heads*ds == 192, ds == 32");
  if (d->C >= d->Cp || d->dh >= d->ds || d->hidden >= d->Hp || d->Hp % 256 != 0 || d->C != d->heads * d->dh)
    return fail(SRK_ERR_UNSUPPORTED, "need C < Cp, dh < ds, hidden < Hp, Hp % 256 == 0, C == heads*dh");
  if (g) {
    const int ws = g->ws;
    if (ws != 8 && !(hat && ws == 16))
      return fail(SRK_ERR_UNSUPPORTED, "window attention cores are specialised for ws == 8 (Swin, HAT) / 16 (HAT)");
    if (g->H % ws || g->W % ws || g->shift < 0 || g->shift >= ws) return fail(SRK_ERR_ARG, "bad geometry");
    if ((long long)g->B * g->H * g->W % 128 != 0) return fail(SRK_ERR_ARG, "B*H*W must be a multiple of 128");
  }
  return SRK_OK;
}

// SRK_ATTN_TC=1 selects the tcgen05 / TMEM / TMA window-attention kernels (attn_tc8.cuh); they take every shape the models
// produce (cyclic shift 0 or 4, an even number of windows and heads).  Default: the register-resident mma.sync kernels,
// which are faster at head_dim 30 — measured on B200 at the bench shape (tools/gpu_probe_attn_tc.py,
// profiles/r02_attn_tc8_vs_mma.txt): forward 85.8 vs 142.8 us, backward 199.0 vs 332.7 us.
bool attn_tc_eligible(const SrkGeom* g, int heads, int ld_qkv, int ld_o) {
  static const bool on = getenv("SRK_ATTN_TC") && getenv("SRK_ATTN_TC")[0] == '1';
  const long long nwin = (long long)g->B * (g->H / 8) * (g->W / 8);
  return on && (g->shift % 4) == 0 && (heads % 2) == 0 && heads <= TC8_MAX_HEADS && (nwin % 2) == 0 &&
         ld_qkv == 3 * heads * 32 && ld_o == heads * 32;
}

int qkv_window_map(CUtensorMap* tm, const void* p, const SrkGeom* g, int ld) {
  return make_tmap_nhwc(tm, p, ld, g->W, g->H, g->B, ld, (uint64_t)g->W * ld, (uint64_t)g->H * g->W * ld, 4, 4);
}

int wgrad_splits(int T, int Ca) { return srk_gemm_wgrad_splits(T, Ca); }

struct WsLayout {  // offsets (floats) into SrkBlockScratch.wg_ws
  long long part_qkv, part_proj, part_fc1, part_fc2, ln1, ln2, rpb, total;   // per-split weight-gradient partials
  int rpb_gx, ln_grid;
};

int attn_bwd_gx(int nwin, int heads) {
  static const int per_sm = getenv("SRK_ATTN_BWD_CTAS") ? atoi(getenv("SRK_ATTN_BWD_CTAS")) : 3;
  int gx = num_sms() * per_sm / heads;
  if (gx > nwin) gx = nwin;
  return gx < 1 ? 1 : gx;
}
int attn_tc_grid(const SrkGeom* g) {
  const int npairs = g->B * (g->H / 8) * (g->W / 8) / 2;
  return npairs < num_sms() ? npairs : num_sms();
}
// rows of per-CTA bias-table partials the ws-8 backward writes for this geometry (block layouts: ld = 3*heads*32 / heads*32)
int attn_bwd_parts_ld(const SrkGeom* g, int heads, int ld_qkv, int ld_o) {
  if (attn_tc_eligible(g, heads, ld_qkv, ld_o)) return attn_tc_grid(g);
  return attn_bwd_gx(g->B * (g->H / 8) * (g->W / 8), heads);
}
int attn_bwd_parts(const SrkGeom* g, int heads) { return attn_bwd_parts_ld(g, heads, 3 * heads * 32, heads * 32); }
int attn_fwd_gx(int nwin, int heads) {
  int gx = num_sms() * 6 / heads;
  if (gx > nwin) gx = nwin;
  return gx < 1 ? 1 : gx;
}

WsLayout ws_layout(const SrkBlockDims* d, const SrkGeom* g) {
  WsLayout L{};
  const int QW = 3 * d->heads * d->ds, AW = d->heads * d->ds;
  const long long T = (long long)g->B * g->H * g->W;
  // every weight-gradient GEMM of the block keeps its per-split partials until the unpack kernel folds them
  long long o = 0;
  L.part_qkv = o; o += srk_gemm_wgrad_workspace_elems(QW, d->Cp, wgrad_splits(int(T), QW));
  L.part_proj = o; o += srk_gemm_wgrad_workspace_elems(d->Cp, AW, wgrad_splits(int(T), d->Cp));
  L.part_fc1 = o; o += srk_gemm_wgrad_workspace_elems(d->Hp, d->Cp, wgrad_splits(int(T), d->Hp));
  L.part_fc2 = o; o += srk_gemm_wgrad_workspace_elems(d->Hp, d->Cp, wgrad_splits(int(T), d->Hp));
  L.ln_grid = srk_gemm_grid(int(T), d->Cp);
  L.ln1 = o; o += (long long)L.ln_grid * 2 * d->Cp;
  L.ln2 = o; o += (long long)L.ln_grid * 2 * d->Cp;
  L.rpb_gx = attn_bwd_parts(g, d->heads);
  {  // sized for either kernel (the scratch is allocated once per geometry, whatever the block's shift)
    const int a0 = attn_tc_grid(g), a1 = attn_bwd_gx(g->B * (g->H / 8) * (g->W / 8), d->heads);
    L.rpb = o; o += (long long)(a0 > a1 ? a0 : a1) * d->heads * 225;
  }
  L.total = o;
  return L;
}

int launch_attn_fwd(const SrkGeom* g, int heads, const void* qkv, int ld_qkv, const float* table, void* out, int ld_o,
                    int ones_col, cudaStream_t stream, int mask = 0) {
  if (attn_tc_eligible(g, heads, ld_qkv, ld_o)) {
    AttnArgs a{};
    a.mask = mask;
    a.qkv = static_cast<const __nv_bfloat16*>(qkv);
    a.out = static_cast<__nv_bfloat16*>(out);
    a.bias_table = table;
    a.B = g->B; a.H = g->H; a.W = g->W; a.heads = heads; a.shift = g->shift;
    a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.ones_col = ones_col;
    CUtensorMap tm;
    int rc = qkv_window_map(&tm, qkv, g, ld_qkv);
    if (rc) return rc;
    static DeviceOnce configured;
    if (configured.need()) {
      SRK_CUDA_OK(cudaFuncSetAttribute(win_attn_tc8_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC8_SMEM));
      configured.done();
    }
    const int npairs = g->B * (g->H / 8) * (g->W / 8) / 2;
    const int grid = npairs < num_sms() ? npairs : num_sms();
    SRK_CUDA_OK(launch_pdl(win_attn_tc8_fwd_kernel, dim3(grid), dim3(TC8_THREADS), TC8_SMEM, stream, tm, a));
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
    return SRK_OK;
  }
  AttnArgs a{};
  a.mask = mask;
  a.qkv = static_cast<const __nv_bfloat16*>(qkv);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.bias_table = table;
  a.B = g->B; a.H = g->H; a.W = g->W; a.heads = heads; a.shift = g->shift;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.ones_col = ones_col;
  const int nwin = g->B * (g->H / 8) * (g->W / 8);
  dim3 grid(attn_fwd_gx(nwin, heads), heads);
  SRK_CUDA_OK(launch_pdl(win_attn_ws8_fwd_kernel, grid, dim3(ATT_THREADS), 0, stream, a));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

int launch_attn_bwd(const SrkGeom* g, int heads, const void* qkv, int ld_qkv, const float* table, const void* dout,
                    int ld_o, void* dqkv, float* partials, int gx, cudaStream_t stream, int mask = 0) {
  if (attn_tc_eligible(g, heads, ld_qkv, ld_o)) {
    if (gx != attn_tc_grid(g)) return fail(SRK_ERR_ARG, "win_attn bwd: partial-row count does not match the tcgen05 kernel's grid");
    AttnArgs a{};
    a.qkv = static_cast<const __nv_bfloat16*>(qkv);
    a.dout = static_cast<const __nv_bfloat16*>(dout);
    a.dqkv = static_cast<__nv_bfloat16*>(dqkv);
    a.bias_table = table;
    a.dbias_partials = partials;
    a.mask = mask;
    a.B = g->B; a.H = g->H; a.W = g->W; a.heads = heads; a.shift = g->shift;
    a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.ones_col = -1;
    CUtensorMap tq, td;
    int rc = qkv_window_map(&tq, qkv, g, ld_qkv);
    if (rc) return rc;
    if ((rc = qkv_window_map(&td, dout, g, ld_o))) return rc;
    static DeviceOnce configured_tc;
    if (configured_tc.need()) {
      SRK_CUDA_OK(cudaFuncSetAttribute(win_attn_tc8_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC8B_SMEM));
      configured_tc.done();
    }
    SRK_CUDA_OK(launch_pdl(win_attn_tc8_bwd_kernel, dim3(gx), dim3(TC8_THREADS), TC8B_SMEM, stream, tq, td, a));
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
    return SRK_OK;
  }
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(win_attn_ws8_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(sizeof(AttnBwdSmem))));
    configured.done();
  }
  AttnArgs a{};
  a.qkv = static_cast<const __nv_bfloat16*>(qkv);
  a.dout = static_cast<const __nv_bfloat16*>(dout);
  a.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  a.bias_table = table;
  a.dbias_partials = partials;
  a.mask = mask;
  a.B = g->B; a.H = g->H; a.W = g->W; a.heads = heads; a.shift = g->shift;
  a.ld_qkv = ld_qkv; a.ld_o = ld_o; a.ones_col = -1;
  dim3 grid(gx, heads);
  SRK_CUDA_OK(launch_pdl(win_attn_ws8_bwd_kernel, grid, dim3(ATT_THREADS), sizeof(AttnBwdSmem), stream, a));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}


template <int MODE>
int launch_attn16_fwd_t(const Attn16Args& a, cudaStream_t stream) {
  static DeviceOnce configured;
  constexpr int smem = a16_fwd_smem<MODE>();
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(win_attn16_fwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.done();
  }
  const int nitems = a.B * (a.H / 16) * (a.W / 16) * 2;
  int gx = num_sms() * 2 / a.heads;
  if (gx > nitems) gx = nitems;
  if (gx < 1) gx = 1;
  win_attn16_fwd_kernel<MODE><<<dim3(gx, a.heads), A16_FWD_THREADS, smem, stream>>>(a);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

// Window-16 forward on tcgen05 / TMEM / TMA (attn_tc16.cuh): the default for every shape the HAT models produce (even head
// count, packed q|k|v rows, window-aligned image).  SRK_ATTN16_TC=0 selects the mma.sync kernel (A/B measurements, tests).
bool attn16_tc_eligible(const SrkGeom* g, int heads, int ld_qkv, int ld_o) {
  const char* e = getenv("SRK_ATTN16_TC");
  if (e && e[0] == '0') return false;
  return g->ws == 16 && (g->shift == 0 || g->shift == 8) && (heads % 2) == 0 && ld_qkv >= 3 * heads * 32 && ld_o >= heads * 32 &&
         g->H % 16 == 0 && g->W % 16 == 0;
}
template <int MODE>
int launch_attn16_tc_fwd_t(const CUtensorMap& tm, const Attn16Args& a, cudaStream_t stream) {
  using G = TC16<MODE>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(win_attn_tc16_fwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
    configured.done();
  }
  const int nwin = a.B * (a.H / 16) * (a.W / 16), nhp = a.heads / 2;
  int gx = num_sms() / nhp;
  if (gx > nwin) gx = nwin;
  if (gx < 1) gx = 1;
  win_attn_tc16_fwd_kernel<MODE><<<dim3(gx, nhp), TC16_THREADS, G::SMEM, stream>>>(tm, a);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

int attn16_bwd_gx(int nwin, int heads) {
  int gx = num_sms() / heads;
  if (gx > nwin) gx = nwin;
  return gx < 1 ? 1 : gx;
}

struct Attn16Ws { size_t scratch_off, dkv_off, dense_off, total; int gx; };
int oca8_bwd_gx(int nwin, int heads) {
  int gx = num_sms() * 2 / heads;
  if (gx > nwin) gx = nwin;
  return gx < 1 ? 1 : gx;
}
Attn16Ws attn16_ws_layout(const SrkGeom* g, int mode, int heads) {
  Attn16Ws L{};
  if (g->ws == 8) {  // HAT at window 8: SELF runs on the ws-8 core (partials live in the block scratch); OCA = 12x12 keys
    const int nwin = g->B * (g->H / 8) * (g->W / 8);
    L.gx = oca8_bwd_gx(nwin, heads);
    size_t o = 0;
    if (mode == MODE_SELF) {  // either ws-8 backward kernel: the larger of the two partial-row counts
      const int a0 = attn_tc_grid(g), a1 = attn_bwd_gx(nwin, heads);
      o = (size_t)(a0 > a1 ? a0 : a1) * heads * 225 * sizeof(float);
    }
    if (mode == MODE_OCA) {
      L.scratch_off = o; o += (size_t)L.gx * heads * 4 * O8_NT * 32 * 4 * sizeof(float);
      L.dense_off = o; o += (size_t)heads * 64 * O8_NK * sizeof(float);
      L.dkv_off = o; o += (size_t)nwin * heads * 2 * O8_NK * 32 * sizeof(__nv_bfloat16);
    }
    L.total = o < 256 ? 256 : o;
    return L;
  }
  const int nwin = g->B * (g->H / 16) * (g->W / 16);
  L.gx = attn16_bwd_gx(nwin, heads);
  const int nkt = mode == MODE_SELF ? 4 : 12;
  L.scratch_off = 0;
  size_t o = (size_t)L.gx * heads * nkt * 16 * 32 * 4 * sizeof(float);
  L.dkv_off = o;
  if (mode == MODE_OCA) o += (size_t)nwin * heads * 2 * 768 * 32 * sizeof(__nv_bfloat16);
  L.total = o;
  return L;
}

template <int MODE>
int launch_attn16_bwd_t(Attn16Args a, void* ws, const Attn16Ws& L, float* d_table, cudaStream_t stream) {
  using SM = A16BwdSmem<MODE>;
  static DeviceOnce configured;
  if (configured.need()) {
    SRK_CUDA_OK(cudaFuncSetAttribute(win_attn16_bwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
    configured.done();
  }
  a.dbias_scratch = reinterpret_cast<float*>(static_cast<char*>(ws) + L.scratch_off);
  a.dkv_win = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + L.dkv_off);
  win_attn16_bwd_kernel<MODE><<<dim3(L.gx, a.heads), A16_BWD_THREADS, SM::kBytes, stream>>>(a);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  if (MODE == MODE_OCA) {
    oca_kv_gather_kernel<<<num_sms() * 8, 256, 0, stream>>>(a.dkv_win, a.dqkv, a.ld_qkv, a.B, a.H, a.W, a.heads);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
  }
  if (d_table) {
    const int n = A16<MODE>::TBL * a.heads;
    attn16_dbias_finish_kernel<MODE><<<(n + 127) / 128, 128, 0, stream>>>(a.dbias_scratch, L.gx, a.heads, d_table);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
  }
  return SRK_OK;
}

int check_attn16(const SrkGeom* g, int mode, int ld_qkv, int ld_out) {
  if (!g || (g->ws != 16 && g->ws != 8) || g->H % g->ws || g->W % g->ws)
    return fail(SRK_ERR_UNSUPPORTED, "srk_win_attn16: HAT window attention supports ws 16 and ws 8");
  if (mode != MODE_SELF && mode != MODE_OCA) return fail(SRK_ERR_ARG, "srk_win_attn16: mode");
  if (mode == MODE_OCA && g->shift != 0) return fail(SRK_ERR_ARG, "srk_win_attn16: OCA has no shift");
  if (mode == MODE_SELF && g->shift != 0 && g->shift != g->ws / 2) return fail(SRK_ERR_UNSUPPORTED, "srk_win_attn16: shift must be 0 or ws/2");
  if (ld_qkv % 8 || ld_out % 8) return fail(SRK_ERR_ARG, "srk_win_attn16: rows must be 16-byte aligned");
  return SRK_OK;
}

int attn16_fwd(const SrkGeom* g, int mode, int heads, const void* qkv, int ld_qkv, const float* table, void* out,
               int ld_out, float* lse, int ones_col, cudaStream_t stream) {
  Attn16Args a{};
  a.qkv = static_cast<const __nv_bfloat16*>(qkv);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.lse = lse;
  a.bias_table = table;
  a.B = g->B; a.H = g->H; a.W = g->W; a.heads = heads; a.shift = g->shift;
  a.ld_qkv = ld_qkv; a.ld_o = ld_out; a.ones_col = ones_col;
  a.T = (long long)g->B * g->H * g->W;
  if (g->ws == 8) {
    if (mode == MODE_SELF) return launch_attn_fwd(g, heads, qkv, ld_qkv, table, out, ld_out, ones_col, stream, 1);
    static DeviceOnce configured;
    if (configured.need()) {
      SRK_CUDA_OK(cudaFuncSetAttribute(win_attn_oca8_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, O8_FWD_SMEM));
      configured.done();
    }
    const int nwin = g->B * (g->H / 8) * (g->W / 8);
    int gx = num_sms() * 3 / heads;
    if (gx > nwin) gx = nwin;
    if (gx < 1) gx = 1;
    win_attn_oca8_fwd_kernel<<<dim3(gx, heads), O8_THREADS, O8_FWD_SMEM, stream>>>(a);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
    return SRK_OK;
  }
  if (attn16_tc_eligible(g, heads, ld_qkv, ld_out)) {
    CUtensorMap tm;
    const int rc = make_tmap_nhwc(&tm, qkv, ld_qkv, g->W, g->H, g->B, ld_qkv, (uint64_t)g->W * ld_qkv,
                                  (uint64_t)g->H * g->W * ld_qkv, 8, 8);
    if (rc) return rc;
    return mode == MODE_SELF ? launch_attn16_tc_fwd_t<MODE_SELF>(tm, a, stream) : launch_attn16_tc_fwd_t<MODE_OCA>(tm, a, stream);
  }
  return mode == MODE_SELF ? launch_attn16_fwd_t<MODE_SELF>(a, stream) : launch_attn16_fwd_t<MODE_OCA>(a, stream);
}

int attn16_bwd(const SrkGeom* g, int mode, int heads, const void* qkv, int ld_qkv, const float* table, const void* out,
               const void* d_out, int ld_out, const float* lse, void* d_qkv, void* ws, float* d_table,
               cudaStream_t stream) {
  Attn16Args a{};
  a.qkv = static_cast<const __nv_bfloat16*>(qkv);
  a.osave = static_cast<const __nv_bfloat16*>(out);
  a.dout = static_cast<const __nv_bfloat16*>(d_out);
  a.dqkv = static_cast<__nv_bfloat16*>(d_qkv);
  a.lse = const_cast<float*>(lse);
  a.bias_table = table;
  a.B = g->B; a.H = g->H; a.W = g->W; a.heads = heads; a.shift = g->shift;
  a.ld_qkv = ld_qkv; a.ld_o = ld_out; a.ones_col = -1;
  a.T = (long long)g->B * g->H * g->W;
  const Attn16Ws L = attn16_ws_layout(g, mode, heads);
  if (g->ws == 8) {
    if (mode != MODE_OCA) return fail(SRK_ERR_ARG, "attn16_bwd: the ws-8 self-attention backward runs through srk_win_attn_bwd");
    static DeviceOnce configured;
    if (configured.need()) {
      SRK_CUDA_OK(cudaFuncSetAttribute(win_attn_oca8_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, O8_BWD_SMEM));
      configured.done();
    }
    char* wsb = static_cast<char*>(ws);
    a.dbias_scratch = reinterpret_cast<float*>(wsb + L.scratch_off);
    a.dkv_win = reinterpret_cast<__nv_bfloat16*>(wsb + L.dkv_off);
    float* dense = reinterpret_cast<float*>(wsb + L.dense_off);
    win_attn_oca8_bwd_kernel<<<dim3(L.gx, heads), O8_THREADS, O8_BWD_SMEM, stream>>>(a);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
    oca8_kv_gather_kernel<<<num_sms() * 8, 256, 0, stream>>>(a.dkv_win, a.dqkv, a.ld_qkv, a.B, a.H, a.W, heads);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
    if (d_table) {
      const int nd = heads * 64 * O8_NK;
      oca8_dbias_dense_kernel<<<(nd + 255) / 256, 256, 0, stream>>>(a.dbias_scratch, L.gx, heads, dense);
      SRK_LAUNCHED(1);
      oca8_dbias_table_kernel<<<(O8_TBL * heads + 127) / 128, 128, 0, stream>>>(dense, heads, d_table);
      SRK_LAUNCHED(1);
      SRK_CUDA_OK(cudaGetLastError());
    }
    return SRK_OK;
  }
  return mode == MODE_SELF ? launch_attn16_bwd_t<MODE_SELF>(a, ws, L, d_table, stream)
                           : launch_attn16_bwd_t<MODE_OCA>(a, ws, L, d_table, stream);
}

__global__ void rpb_partials_reduce_kernel(const float* __restrict__ part, int nparts, int heads, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = t*heads + h (reference layout)
  if (i >= 225 * heads) return;
  const int t = i / heads, h = i % heads;
  float acc = 0.f;
  for (int k = 0; k < nparts; ++k) acc += part[(size_t(k) * heads + h) * 225 + t];
  out[i] = acc;
}

}  // namespace

extern "C" void srk_block_weight_elems(const SrkBlockDims* d, long long out[8]) {
  const long long QW = 3LL * d->heads * d->ds, AW = 1LL * d->heads * d->ds;
  out[0] = QW * d->Cp; out[1] = d->Cp * QW;
  out[2] = d->Cp * AW; out[3] = AW * d->Cp;
  out[4] = 1LL * d->Hp * d->Cp; out[5] = 1LL * d->Cp * d->Hp;
  out[6] = 1LL * d->Cp * d->Hp; out[7] = 1LL * d->Hp * d->Cp;
}

extern "C" long long srk_block_bwd_scratch_floats(const SrkBlockDims* d, const SrkGeom* g) {
  return ws_layout(d, g).total;
}

extern "C" int srk_block_prep_weights(const SrkBlockDims* d, const SrkBlockParams* p, const SrkBlockWeights* w,
                                      void* stream_) {
  int rc = check_dims(d, nullptr);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BlockParamPtrs pp{p->norm1_w, p->norm1_b, p->rpb_table, p->qkv_w, p->qkv_b, p->proj_w, p->proj_b,
                    p->norm2_w, p->norm2_b, p->fc1_w,     p->fc1_b, p->fc2_w,  p->fc2_b};
  BlockWeightPtrs ww{static_cast<__nv_bfloat16*>(w->qkv_f),  static_cast<__nv_bfloat16*>(w->qkv_t),
                     static_cast<__nv_bfloat16*>(w->proj_f), static_cast<__nv_bfloat16*>(w->proj_t),
                     static_cast<__nv_bfloat16*>(w->fc1_f),  static_cast<__nv_bfloat16*>(w->fc1_t),
                     static_cast<__nv_bfloat16*>(w->fc2_f),  static_cast<__nv_bfloat16*>(w->fc2_t)};
  SRK_CUDA_OK(launch_pdl(prep_block_weights_kernel, dim3(296), dim3(256), 0, stream, to_dims(d), pp, ww));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

static int block_fwd_impl(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w,
                          const SrkBlockParams* p, const float* next_norm_w, const float* next_norm_b,
                          const SrkBlockActs* a, const SrkHatExtra* x, void* stream) {
  int rc = check_dims(d, g, x != nullptr);
  if (rc) return rc;
  const int T = g->B * g->H * g->W;
  const int QW = 3 * d->heads * d->ds, AW = d->heads * d->ds, Cp = d->Cp, Hp = d->Hp;
  // qkv = xn1 @ Wqkv^T (+bias via ones column)
  if ((rc = srk_gemm_tn(SRK_EPI_STORE, T, QW, Cp, a->xn1, Cp, w->qkv_f, Cp, a->qkv, QW, nullptr, 0, nullptr, 0,
                        nullptr, 0, nullptr, stream)))
    return rc;
  // attention core (shift / partition / reverse by address arithmetic); ao[:, dh] = 1 (proj bias column)
  if (x) {
    if ((rc = check_attn16(g, x->mode, QW, AW))) return rc;
    if (!x->res_in || (!x->lse && g->ws == 16)) return fail(SRK_ERR_ARG, "srk_hat_block_fwd: res_in and lse are required");
    rc = attn16_fwd(g, x->mode, d->heads, a->qkv, QW, p->rpb_table, a->ao, AW, x->lse, d->dh, static_cast<cudaStream_t>(stream));
  } else {
    rc = launch_attn_fwd(g, d->heads, a->qkv, QW, p->rpb_table, a->ao, AW, d->dh, static_cast<cudaStream_t>(stream));
  }
  if (rc) return rc;
  // x_mid = residual + proj(ao); xn2 = LN2(x_mid)   (residual: x_in, or for HAB x_in + conv_scale * CAB(xn1))
  const void* resid = x ? x->res_in : a->x_in;
  SrkLnArgs ln2{d->C, d->C, p->norm2_w, p->norm2_b, a->stats2, nullptr, 1e-5f, x ? x->drop_attn : nullptr, g->H * g->W};
  if ((rc = srk_gemm_tn(SRK_EPI_RES_LN, T, Cp, AW, a->ao, AW, w->proj_f, AW, a->x_mid, Cp, a->xn2, Cp, resid, Cp,
                        nullptr, 0, &ln2, stream)))
    return rc;
  // MLP half.  SRK_FUSED_MLP=1: one fused kernel (fc1 -> GELU -> fc2 -> residual -> next LayerNorm; the hidden tile goes from
  // TMEM through GELU into shared-memory boxes that are fc2's A operand) — bit-identical act / dact / x_out, but measured
  // SLOWER than the two-kernel path on B200 (410 vs 329 us per block at batch 16, tools/gpu_probe_mlp_fused.py): the GELU
  // epilogue's instruction stream, not the re-read of the hidden tensor, is what bounds this half.
  SrkLnArgs lnn{d->C, d->C, next_norm_w, next_norm_b, a->stats_out, nullptr, 1e-5f, x ? x->drop_mlp : nullptr, g->H * g->W};
  static const bool fused_mlp = getenv("SRK_FUSED_MLP") && getenv("SRK_FUSED_MLP")[0] == '1';
  if (fused_mlp && (a->dact || !a->act) && Hp % 128 == 0)
    return srk_mlp_fwd(T, Cp, Hp, a->xn2, w->fc1_f, w->fc2_f, a->x_mid, a->act, a->dact, a->x_out, a->xn_out, d->hidden, &lnn,
                       stream);
  // act = gelu(fc1(xn2)); gelu'(.) is stored only if the caller provides `dact` (otherwise the backward recomputes it)
  SrkLnArgs ge{Hp, d->hidden, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, 1};
  if (a->dact) {
    rc = srk_gemm_tn(SRK_EPI_GELU2, T, Hp, Cp, a->xn2, Cp, w->fc1_f, Cp, a->act, Hp, a->dact, Hp, nullptr, 0, nullptr, 0,
                     &ge, stream);
  } else {
    rc = srk_gemm_tn(SRK_EPI_GELU1, T, Hp, Cp, a->xn2, Cp, w->fc1_f, Cp, a->act, Hp, nullptr, 0, nullptr, 0, nullptr, 0,
                     &ge, stream);
  }
  if (rc) return rc;
  // x_out = x_mid + fc2(act); xn_out = LN_next(x_out)
  if ((rc = srk_gemm_tn(SRK_EPI_RES_LN, T, Cp, Hp, a->act, Hp, w->fc2_f, Hp, a->x_out, Cp, a->xn_out, Cp, a->x_mid, Cp,
                        nullptr, 0, &lnn, stream)))
    return rc;
  return SRK_OK;
}

static int block_bwd_impl(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w,
                          const SrkBlockParams* p, const SrkBlockActs* a, const void* g_out,
                          const SrkBlockScratch* s, void* g_in, const SrkBlockGrads* grads, int accumulate,
                          const SrkHatExtra* x, void* stream_) {
  int rc = check_dims(d, g, x != nullptr);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int T = g->B * g->H * g->W;
  const int QW = 3 * d->heads * d->ds, AW = d->heads * d->ds, Cp = d->Cp, Hp = d->Hp;
  SrkGeom g8 = *g;
  g8.ws = 8;  // the fp32 scratch layout only depends on T; its rpb partial area is used by the ws == 8 core alone
  const WsLayout L = ws_layout(d, &g8);
  float* ws = s->wg_ws;

  // stochastic depth: the branch sees s_b * g (per-sample factor); the residual path keeps g
  auto scaled = [&](const void* gsrc, const float* factors) -> const void* {
    if (factors == nullptr) return gsrc;
    row_scale_kernel<<<num_sms() * 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(gsrc),
                                                      static_cast<__nv_bfloat16*>(x->gs_buf), factors, T, Cp, g->H * g->W);
    SRK_LAUNCHED(1);
    return x->gs_buf;
  };
  if (x && (x->drop_attn || x->drop_mlp) && !x->gs_buf) return fail(SRK_ERR_ARG, "srk_hat_block_bwd: gs_buf is required with drop_*");
  const void* g_mlp = scaled(g_out, x ? x->drop_mlp : nullptr);
  // dU = (g_out @ W2) * gelu'(u): gelu' read back if the forward stored it, else u = xn2 @ W1^T recomputed (MULG)
  if (a->dact) {
    rc = srk_gemm_tn(SRK_EPI_MUL, T, Hp, Cp, g_mlp, Cp, w->fc2_t, Cp, s->d_act, Hp, nullptr, 0, a->dact, Hp, nullptr, 0,
                     nullptr, stream_);
  } else {
    SrkLnArgs ge{Hp, d->hidden, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, 1};
    rc = srk_gemm_tn(SRK_EPI_MULG, T, Hp, Cp, g_mlp, Cp, w->fc2_t, Cp, s->d_act, Hp, nullptr, 0, a->xn2, Cp, w->fc1_f, Cp,
                     &ge, stream_);
  }
  if (rc) return rc;
  // dW2^T (+db2 in row `hidden`) = act^T @ g_out
  const int s_fc = wgrad_splits(T, Hp);
  if ((rc = gemm_wgrad_partials(T, Hp, Cp, a->act, Hp, g_mlp, Cp, ws + L.part_fc2, s_fc, stream_))) return rc;
  // g_mid = g_out + LN2bwd(dU @ W1)
  SrkLnArgs ln2{d->C, -1, p->norm2_w, nullptr, a->stats2, ws + L.ln2, 1e-5f, nullptr, 1};
  if ((rc = srk_gemm_tn(SRK_EPI_LNBWD, T, Cp, Hp, s->d_act, Hp, w->fc1_t, Hp, s->g_mid, Cp, nullptr, 0, a->x_mid, Cp,
                        g_out, Cp, &ln2, stream_)))
    return rc;
  // dW1 (+db1 in column C) = dU^T @ xn2
  if ((rc = gemm_wgrad_partials(T, Hp, Cp, s->d_act, Hp, a->xn2, Cp, ws + L.part_fc1, s_fc, stream_))) return rc;
  // d_ao = g_mid @ Wproj
  const void* g_att = scaled(s->g_mid, x ? x->drop_attn : nullptr);
  if ((rc = srk_gemm_tn(SRK_EPI_STORE, T, AW, Cp, g_att, Cp, w->proj_t, Cp, s->d_ao, AW, nullptr, 0, nullptr, 0,
                        nullptr, 0, nullptr, stream_)))
    return rc;
  // dWproj (+dbproj in column dh) = g_mid^T @ ao
  const int s_proj = wgrad_splits(T, Cp);
  if ((rc = gemm_wgrad_partials(T, Cp, AW, g_att, Cp, a->ao, AW, ws + L.part_proj, s_proj, stream_))) return rc;
  // attention backward -> d_qkv, rpb-table gradient (ws 8: per-CTA partials folded by the unpack kernel below)
  const bool ws8_self = !x || (g->ws == 8 && x->mode == MODE_SELF);   // rpb-table gradient arrives as per-CTA partials
  if (x && ws8_self) {
    if ((rc = check_attn16(g, x->mode, QW, AW))) return rc;
    rc = launch_attn_bwd(g, d->heads, a->qkv, QW, p->rpb_table, s->d_ao, AW, s->d_qkv, ws + L.rpb, L.rpb_gx, stream, 1);
  } else if (x) {
    if ((rc = check_attn16(g, x->mode, QW, AW))) return rc;
    if ((!x->lse && g->ws == 16) || !x->attn_ws) return fail(SRK_ERR_ARG, "srk_hat_block_bwd: lse and attn_ws are required");
    rc = attn16_bwd(g, x->mode, d->heads, a->qkv, QW, p->rpb_table, a->ao, s->d_ao, AW, x->lse, s->d_qkv, x->attn_ws,
                    grads->rpb_table, stream);
  } else {
    rc = launch_attn_bwd(g, d->heads, a->qkv, QW, p->rpb_table, s->d_ao, AW, s->d_qkv, ws + L.rpb, L.rpb_gx, stream);
  }
  if (rc) return rc;
  const bool defer_ln1 = x && x->d_xn1;
  if (defer_ln1) {
    // d_xn1 = d_qkv @ Wqkv; the caller adds the CAB branch and runs the LayerNorm-1 backward itself
    if ((rc = srk_gemm_tn(SRK_EPI_STORE, T, Cp, QW, s->d_qkv, QW, w->qkv_t, QW, x->d_xn1, Cp, nullptr, 0, nullptr, 0,
                          nullptr, 0, nullptr, stream_)))
      return rc;
  } else {
    // g_in = g_mid + LN1bwd(d_qkv @ Wqkv)
    SrkLnArgs ln1{d->C, -1, p->norm1_w, nullptr, const_cast<float*>(a->stats1), ws + L.ln1, 1e-5f, nullptr, 1};
    if ((rc = srk_gemm_tn(SRK_EPI_LNBWD, T, Cp, QW, s->d_qkv, QW, w->qkv_t, QW, g_in, Cp, nullptr, 0, a->x_in, Cp,
                          s->g_mid, Cp, &ln1, stream_)))
      return rc;
  }
  // dWqkv (+dbqkv in column C) = d_qkv^T @ xn1
  const int s_qkv = wgrad_splits(T, QW);
  if ((rc = gemm_wgrad_partials(T, QW, Cp, s->d_qkv, QW, a->xn1, Cp, ws + L.part_qkv, s_qkv, stream_))) return rc;
  // scatter everything into reference-shaped fp32 gradients
  auto split_sum = [&](long long off, int Ca, int Cb, int splits) {
    return SplitSum{ws + off, splits, srk_gemm_wgrad_workspace_elems(Ca, Cb, splits) / splits};
  };
  UnpackSrc us{split_sum(L.part_qkv, QW, Cp, s_qkv), split_sum(L.part_proj, Cp, AW, s_proj),
               split_sum(L.part_fc1, Hp, Cp, s_fc), split_sum(L.part_fc2, Hp, Cp, s_fc), defer_ln1 ? nullptr : ws + L.ln1,
               ws + L.ln2, ws8_self ? ws + L.rpb : nullptr, L.ln_grid, L.rpb_gx, 225};
  BlockGradPtrs gp{grads->norm1_w, grads->norm1_b, grads->rpb_table, grads->qkv_w, grads->qkv_b, grads->proj_w,
                   grads->proj_b,  grads->norm2_w, grads->norm2_b,   grads->fc1_w, grads->fc1_b, grads->fc2_w,
                   grads->fc2_b};
  SRK_CUDA_OK(launch_pdl(unpack_block_grads_kernel, dim3(num_sms() * 8), dim3(256), 0, stream, to_dims(d), us, gp,
                         accumulate ? 1.f : 0.f));
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_swin_block_fwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w,
                                  const SrkBlockParams* p, const float* next_norm_w, const float* next_norm_b,
                                  const SrkBlockActs* a, void* stream) {
  return block_fwd_impl(d, g, w, p, next_norm_w, next_norm_b, a, nullptr, stream);
}

extern "C" int srk_swin_block_bwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w,
                                  const SrkBlockParams* p, const SrkBlockActs* a, const void* g_out,
                                  const SrkBlockScratch* s, void* g_in, const SrkBlockGrads* grads, int accumulate,
                                  void* stream) {
  return block_bwd_impl(d, g, w, p, a, g_out, s, g_in, grads, accumulate, nullptr, stream);
}

extern "C" int srk_hat_block_fwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w,
                                 const SrkBlockParams* p, const float* next_norm_w, const float* next_norm_b,
                                 const SrkBlockActs* a, const SrkHatExtra* x, void* stream) {
  if (!x) return fail(SRK_ERR_ARG, "srk_hat_block_fwd: SrkHatExtra is required");
  return block_fwd_impl(d, g, w, p, next_norm_w, next_norm_b, a, x, stream);
}

extern "C" int srk_hat_block_bwd(const SrkBlockDims* d, const SrkGeom* g, const SrkBlockWeights* w,
                                 const SrkBlockParams* p, const SrkBlockActs* a, const void* g_out,
                                 const SrkBlockScratch* s, void* g_in, const SrkBlockGrads* grads, const SrkHatExtra* x,
                                 void* stream) {
  if (!x) return fail(SRK_ERR_ARG, "srk_hat_block_bwd: SrkHatExtra is required");
  return block_bwd_impl(d, g, w, p, a, g_out, s, g_in, grads, 0, x, stream);
}

extern "C" int srk_win_attn16_fwd(const SrkGeom* g, int mode, int heads, const void* qkv, int ld_qkv,
                                  const float* rpb_table, void* out, int ld_out, float* lse, int ones_col, void* stream) {
  int rc = check_attn16(g, mode, ld_qkv, ld_out);
  if (rc) return rc;
  return attn16_fwd(g, mode, heads, qkv, ld_qkv, rpb_table, out, ld_out, lse, ones_col, static_cast<cudaStream_t>(stream));
}

extern "C" long long srk_win_attn16_bwd_ws_bytes(const SrkGeom* g, int mode, int heads) {
  return (long long)attn16_ws_layout(g, mode, heads).total;
}

extern "C" int srk_win_attn16_bwd(const SrkGeom* g, int mode, int heads, const void* qkv, int ld_qkv,
                                  const float* rpb_table, const void* out, const void* d_out, int ld_out,
                                  const float* lse, void* d_qkv, void* ws, float* d_rpb_table, void* stream) {
  int rc = check_attn16(g, mode, ld_qkv, ld_out);
  if (rc) return rc;
  if ((!lse && g->ws == 16) || !ws || !out) return fail(SRK_ERR_ARG, "srk_win_attn16_bwd: lse, ws and the forward output are required");
  if (g->ws == 8 && mode == MODE_SELF) {  // ws-8 core with HAT's shift mask; per-CTA table partials in ws
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int gx = attn_bwd_parts_ld(g, heads, ld_qkv, ld_out);
    rc = launch_attn_bwd(g, heads, qkv, ld_qkv, rpb_table, d_out, ld_out, d_qkv, static_cast<float*>(ws), gx, st, 1);
    if (rc) return rc;
    if (d_rpb_table) {
      rpb_partials_reduce_kernel<<<(225 * heads + 127) / 128, 128, 0, st>>>(static_cast<float*>(ws), gx, heads, d_rpb_table);
      SRK_LAUNCHED(1);
      SRK_CUDA_OK(cudaGetLastError());
    }
    return SRK_OK;
  }
  return attn16_bwd(g, mode, heads, qkv, ld_qkv, rpb_table, out, d_out, ld_out, lse, d_qkv, ws, d_rpb_table,
                    static_cast<cudaStream_t>(stream));
}

extern "C" int srk_cab_se_fwd(const void* y, const void* x, int B, int HW, int C, int Cp, int S, const float* w1,
                              const float* b1, const float* w2, const float* b2, float alpha, float* ws, float* pool,
                              float* hidden, float* scale, void* out, void* stream_) {
  if (Cp % 8 || C > Cp || S < 1 || S > 32 || Cp > 256) return fail(SRK_ERR_UNSUPPORTED, "srk_cab_se_fwd: shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int nchunk = 32, groups = Cp / 8, rpb = 16;
  cab_colsum_kernel<<<dim3(nchunk, B), groups * rpb, Cp * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(y), nullptr, HW, Cp, ws);
  SRK_LAUNCHED(1);
  cab_se_fwd_kernel<<<B, 256, (C + S) * sizeof(float), stream>>>(ws, nchunk, Cp, HW, C, S, w1, b1, w2, b2, pool, hidden, scale);
  SRK_LAUNCHED(1);
  cab_combine_fwd_kernel<<<num_sms() * 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                            static_cast<const __nv_bfloat16*>(y), scale, alpha,
                                                            static_cast<__nv_bfloat16*>(out), (long long)B * HW, HW, C, Cp);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_cab_se_bwd(const void* g, const void* y, int B, int HW, int C, int Cp, int S, const float* w1,
                              const float* w2, float alpha, const float* pool, const float* hidden, const float* scale,
                              float* ws, void* dy, float* dw1, float* db1, float* dw2, float* db2, void* stream_) {
  if (Cp % 8 || C > Cp || S < 1 || S > 32 || Cp > 256) return fail(SRK_ERR_UNSUPPORTED, "srk_cab_se_bwd: shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int nchunk = 32, groups = Cp / 8, rpb = 16;
  float* dpool_hw = ws + (size_t)B * nchunk * Cp;
  cab_colsum_kernel<<<dim3(nchunk, B), groups * rpb, Cp * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(y), HW, Cp, ws);
  SRK_LAUNCHED(1);
  cab_se_bwd_kernel<<<1, 256, (C + 2 * S) * sizeof(float), stream>>>(ws, nchunk, Cp, HW, B, C, S, alpha, pool, hidden, scale,
                                                                     w1, w2, dpool_hw, dw1, db1, dw2, db2);
  SRK_LAUNCHED(1);
  cab_combine_bwd_kernel<<<num_sms() * 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(g), scale, dpool_hw, alpha,
                                                            static_cast<__nv_bfloat16*>(dy), (long long)B * HW, HW, C, Cp);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_win_attn_fwd(const SrkGeom* g, int heads, const void* qkv, int ld_qkv, const float* rpb_table,
                                void* out, int ld_out, int ones_col, void* stream) {
  if (!g || g->ws != 8 || g->H % 8 || g->W % 8) return fail(SRK_ERR_UNSUPPORTED, "srk_win_attn_fwd: ws must be 8");
  if (ld_qkv % 8 || ld_out % 8) return fail(SRK_ERR_ARG, "srk_win_attn_fwd: rows must be 16-byte aligned");
  return launch_attn_fwd(g, heads, qkv, ld_qkv, rpb_table, out, ld_out, ones_col, static_cast<cudaStream_t>(stream));
}

extern "C" long long srk_win_attn_bwd_ws_floats(int heads) {
  return (long long)num_sms() * (heads > 4 ? heads : 4) * 225 + 225LL * heads;   // per-CTA partials of either backward kernel
}

extern "C" int srk_win_attn_bwd(const SrkGeom* g, int heads, const void* qkv, int ld_qkv, const float* rpb_table,
                                const void* d_out, int ld_out, void* d_qkv, float* dbias_ws, float* d_rpb_table,
                                void* stream_) {
  if (!g || g->ws != 8 || g->H % 8 || g->W % 8) return fail(SRK_ERR_UNSUPPORTED, "srk_win_attn_bwd: ws must be 8");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int gx = attn_bwd_parts_ld(g, heads, ld_qkv, ld_out);
  int rc = launch_attn_bwd(g, heads, qkv, ld_qkv, rpb_table, d_out, ld_out, d_qkv, dbias_ws, gx, stream);
  if (rc) return rc;
  if (d_rpb_table) {
    rpb_partials_reduce_kernel<<<(225 * heads + 127) / 128, 128, 0, stream>>>(dbias_ws, gx, heads, d_rpb_table);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
  }
  return SRK_OK;
}

extern "C" int srk_layernorm_fwd(const void* x, int ldx, void* y, int ldy, float* stats, const float* gamma,
                                 const float* beta, int rows, int C, int Cp, int ones_col, float eps, void* stream_) {
  if (Cp > 256 || C > Cp) return fail(SRK_ERR_UNSUPPORTED, "srk_layernorm_fwd: Cp <= 256");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int wpb = 8;
  ln_fwd_rows_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy, stats, gamma, beta, rows, C, Cp,
      ones_col, eps);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" long long srk_layernorm_bwd_ws_floats(int Cp) { return (long long)num_sms() * 4 * 2 * Cp; }

extern "C" int srk_layernorm_bwd(const void* dy, int lddy, const void* x, int ldx, const float* stats,
                                 const float* gamma, const void* dres, int lddres, void* dx, int lddx, float* part_ws,
                                 float* dgamma, float* dbeta, int rows, int C, int Cp, void* stream_) {
  if (Cp > 256 || C > Cp) return fail(SRK_ERR_UNSUPPORTED, "srk_layernorm_bwd: Cp <= 256");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int grid = num_sms() * 4;
  if (grid > (rows + 7) / 8) grid = (rows + 7) / 8;
  ln_bwd_rows_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy), lddy,
                                               static_cast<const __nv_bfloat16*>(x), ldx, stats, gamma,
                                               static_cast<const __nv_bfloat16*>(dres), lddres,
                                               static_cast<__nv_bfloat16*>(dx), lddx, part_ws, rows, C, Cp);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  if (dgamma && dbeta) {
    ln_param_grad_reduce_kernel<<<(2 * C * 32 + 255) / 256, 256, 0, stream>>>(part_ws, grid, Cp, C, dgamma, dbeta);
    SRK_LAUNCHED(1);
    SRK_CUDA_OK(cudaGetLastError());
  }
  return SRK_OK;
}
