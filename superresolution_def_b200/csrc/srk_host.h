// srk_host.h — host-side helpers: error codes, TMA tensor-map encoding, device properties.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <utility>

#include "../../include/srk.h"

namespace srk {

// number of kernels launched by this library since load (reported by srk_launch_count)
extern std::atomic<long long> g_launches;
#define SRK_LAUNCHED(n) (::srk::g_launches.fetch_add((n), std::memory_order_relaxed))

inline int fail(int code, const char* what) {
  fprintf(stderr, "[srk] error %d: %s\n", code, what);
  return code;
}

#define SRK_CUDA_OK(expr)                                                          \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      fprintf(stderr, "[srk] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),  \
              __FILE__, __LINE__, cudaGetErrorString(_e));                         \
      return SRK_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows x 64 cols],
// 128B swizzle, OOB reads return zero, OOB writes are clipped.
inline int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols = 64) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SRK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(SRK_ERR_ARG, "TMA tensor must be 16-byte aligned with a 16-byte-multiple row pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[srk] cuTensorMapEncodeTiled failed: %d (rows=%llu cols=%llu ld=%llu box=%ux%u)\n", int(r),
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return SRK_ERR_CUDA;
  }
  return SRK_OK;
}

// 4-D bf16 NHWC view [B,H,W,C] with arbitrary pixel / row / image pitches (elements); box = (64 ch, bw, bh, 1).
inline int make_tmap_nhwc(CUtensorMap* out, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t B,
                          uint64_t pix_pitch, uint64_t row_pitch, uint64_t img_pitch, uint32_t bw, uint32_t bh) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SRK_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (pix_pitch * 2) % 16 != 0)
    return fail(SRK_ERR_ARG, "NHWC tensor must be 16-byte aligned with 16-byte-multiple pixel pitch");
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {pix_pitch * 2, row_pitch * 2, img_pitch * 2};
  cuuint32_t box[4] = {64, bw, bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[srk] cuTensorMapEncodeTiled(4d) failed: %d\n", int(r));
    return SRK_ERR_CUDA;
  }
  return SRK_OK;
}

// SRK_PDL=1: per-block kernels are launched with programmatic stream serialization (they all call pdl_wait()).
// Default off — measured on B200 (SwinIR step as a CUDA graph): 71.4 ms with PDL vs 70.1 ms without; the kernels are
// persistent, one CTA per SM and HBM-bound, so an early-resident successor only spins in griddepcontrol.wait.
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SRK_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}
// Launch `kernel` so that it may overlap the tail of its stream predecessor.  ONLY for kernels that execute
// pdl_wait() before their first read of global memory written by earlier kernels.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// Function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) and the SM count are per device: a process that
// touches a second GPU (tiled inference across devices without torchrun) must configure each kernel there too.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  static int cur() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool need() const { return ((mask.load(std::memory_order_acquire) >> cur()) & 1ull) == 0; }
  void done() { mask.fetch_or(1ull << cur(), std::memory_order_release); }
};

// api_gemm.cu: weight-gradient GEMM that leaves its per-split partials in `workspace` (no reduce kernel)
int gemm_wgrad_partials(int T, int Ca, int Cb, const void* A, int lda, const void* B, int ldb, float* workspace, int splits,
                        void* stream);

inline int num_sms() {
  static std::atomic<int> n[64];
  const int dev = DeviceOnce::cur();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

}  // namespace srk
