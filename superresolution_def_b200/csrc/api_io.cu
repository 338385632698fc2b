// api_io.cu — the data formats on either side of the hot path: 16-bit image planes in and out.
//
//   srk_u16_to_f32_aug  uint16 planes -> float32 /65535 with the dataset's augmentation (horizontal flip, vertical flip,
//                       rot90^k) applied as ADDRESS ARITHMETIC in the same pass.  Replaces, per sample,
//                       dataset/astronomical_dataset_swin.py:34-39 (np.array(img, float32) / 65535.0) and :58-67
//                       (torch.flip / torch.rot90 on the host, one tensor copy each) — the host now only reads the TIFF into a
//                       pinned uint16 buffer (half the H2D bytes of the float tensor the reference ships).
//   srk_f32_to_u16      float32 [0,1] -> uint16, the quantisation of infer_hat.py:42-50 / infer_swin.py (clip, * 65535,
//                       truncate), so that the D2H copy of a super-resolved frame is 2 bytes per pixel.
// Both are HBM-bound byte kernels: 32 x 32 tiles through shared memory so that reads and writes are coalesced for every
// orientation (a rot90 turns rows into columns), 2 + 4 bytes per pixel, no re-reads.
#include "srk_host.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace srk {

// Source pixel of output pixel (i, j) of an n x n plane for code = fh | fv << 1 | k << 2, where the reference applies
// flip(-1) if fh, then flip(-2) if fv, then rot90(k) counter-clockwise in the (-2, -1) plane:
//   rot90^1: out[i][j] = in[j][n-1-i];  rot90^2: out[i][j] = in[n-1-i][n-1-j];  rot90^3: out[i][j] = in[n-1-j][i].
__device__ __forceinline__ void aug_source(int code, int n, int i, int j, int& sy, int& sx) {
  const int k = (code >> 2) & 3;
  int a, b;
  if (k == 0) { a = i; b = j; }
  else if (k == 1) { a = j; b = n - 1 - i; }
  else if (k == 2) { a = n - 1 - i; b = n - 1 - j; }
  else { a = n - 1 - j; b = i; }
  sy = (code & 2) ? n - 1 - a : a;
  sx = (code & 1) ? n - 1 - b : b;
}

__global__ void __launch_bounds__(256) u16_to_f32_aug_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst,
                                                              const int* __restrict__ codes, int n) {
  __shared__ uint16_t tile[32][33];
  const int b = blockIdx.z;
  const int code = codes ? codes[b] : 0;
  const int oy0 = blockIdx.y * 32, ox0 = blockIdx.x * 32;
  // the isometry maps this 32 x 32 output tile onto one 32 x 32 source tile: find its origin from two opposite corners
  int y0, x0, y1, x1;
  aug_source(code, n, oy0, ox0, y0, x0);
  aug_source(code, n, oy0 + 31, ox0 + 31, y1, x1);
  const int sy0 = y0 < y1 ? y0 : y1, sx0 = x0 < x1 ? x0 : x1;
  const uint16_t* s = src + (size_t)b * n * n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
#pragma unroll
  for (int r = 0; r < 4; ++r) tile[ty + 8 * r][tx] = s[(size_t)(sy0 + ty + 8 * r) * n + sx0 + tx];   // coalesced rows
  __syncthreads();
  float* d = dst + (size_t)b * n * n;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = oy0 + ty + 8 * r, j = ox0 + tx;
    int sy, sx;
    aug_source(code, n, i, j, sy, sx);
    // the reference divides in float32: arr / 65535.0 (numpy float32 array / python float) — IEEE division, not a reciprocal
    d[(size_t)i * n + j] = __fdiv_rn(float(tile[sy - sy0][sx - sx0]), 65535.0f);
  }
}

__global__ void __launch_bounds__(256) f32_to_u16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  const long long i = (long long)(blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    ushort4 o;
    uint16_t* po = reinterpret_cast<uint16_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) po[e] = (uint16_t)(fminf(fmaxf(f[e], 0.f), 1.f) * 65535.0f);   // np.clip, * 65535, astype: truncation
    *reinterpret_cast<ushort4*>(dst + i) = o;
  } else {
    for (long long k = i; k < n; ++k) dst[k] = (uint16_t)(fminf(fmaxf(src[k], 0.f), 1.f) * 65535.0f);
  }
}

}  // namespace srk

using namespace srk;

extern "C" int srk_u16_to_f32_aug(const void* src_u16, float* dst, const int* codes, int B, int n, void* stream_) {
  if (!src_u16 || !dst || B <= 0 || n <= 0 || n % 32 != 0) return fail(SRK_ERR_ARG, "srk_u16_to_f32_aug: square planes, n % 32 == 0");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  u16_to_f32_aug_kernel<<<dim3(n / 32, n / 32, B), 256, 0, stream>>>(static_cast<const uint16_t*>(src_u16), dst, codes, n);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}

extern "C" int srk_f32_to_u16(const float* src, void* dst_u16, long long n, void* stream_) {
  if (!src || !dst_u16 || n <= 0 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst_u16) & 7))
    return fail(SRK_ERR_ARG, "srk_f32_to_u16: aligned pointers, n > 0");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long threads = (n + 3) / 4;
  f32_to_u16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(src, static_cast<uint16_t*>(dst_u16), n);
  SRK_LAUNCHED(1);
  SRK_CUDA_OK(cudaGetLastError());
  return SRK_OK;
}
