// swb200_pack_weights: a reference-schema SwinV2 checkpoint (fp32 parameters, the names and layouts of
// `SwinV2.state_dict()`, models/swinv2.py:278-292) -> the packed device layouts of `struct swb200_model`
// (include/swift_b200.h), so a host in any language can build the model from raw parameter pointers.
// One-time per checkpoint (and per optimiser step when training): a few dozen gather / convert launches.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace swb {

namespace {
__device__ __forceinline__ uint16_t cvt16(float x, int f16) {
  if (f16) {
    __half h = __float2half_rn(x);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&b);
}

enum RowMode { ROWS_PLAIN = 0, ROWS_QKV = 1, ROWS_W1 = 2 };

// dst[r, c] = cvt(src[srcrow(r), c]) for c < K; dst row pitch ldd; optional duplicate of the row at column offset dup
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int rows, int K,
                                                        int ldd, int dup, int mode, int a, int b, int f16) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * K) return;
  const int r = static_cast<int>(idx / K), c = static_cast<int>(idx - static_cast<long long>(r) * K);
  int sr = r;
  if (mode == ROWS_QKV) {
    // packed row part*D + h*hd + d  <-  reference row h*3*hd + part*hd + d   (a = heads, b = head dim)
    const int D = a * b, part = r / D, rem = r - part * D, h = rem / b, d = rem - h * b;
    sr = (h * 3 + part) * b + d;
  } else if (mode == ROWS_W1) {
    // per GEMM tile of 2*half rows: [half gate rows | half up rows]   (a = half, b = dff); reference rows [gate(dff) | up(dff)]
    const int t = r / (2 * a), j = r - t * 2 * a;
    sr = j < a ? t * a + j : b + t * a + (j - a);
  }
  const uint16_t v = cvt16(__ldg(src + static_cast<size_t>(sr) * K + c), f16);
  dst[static_cast<size_t>(r) * ldd + c] = v;
  if (dup > 0) dst[static_cast<size_t>(r) * ldd + dup + c] = v;
}

// patch-embed weight: reference input-feature order "(p1 p2 c)" -> "(c p1 p2)", zero padded to k_embed, duplicated for the
// [hi | lo] split operand when split
__global__ void __launch_bounds__(256) pack_embed_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int D, int C, int pp,
                                                         int k_embed, int split, int f16) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(D) * k_embed) return;
  const int n = static_cast<int>(idx / k_embed), k = static_cast<int>(idx - static_cast<long long>(n) * k_embed);
  float v = 0.f;
  if (k < C * pp) {
    const int c = k / pp, q = k - c * pp;
    v = __ldg(src + static_cast<size_t>(n) * C * pp + q * C + c);
  }
  const int ld = k_embed * (split ? 2 : 1);
  const uint16_t h = cvt16(v, f16);
  dst[static_cast<size_t>(n) * ld + k] = h;
  if (split) dst[static_cast<size_t>(n) * ld + k_embed + k] = h;
}

__global__ void __launch_bounds__(256) pack_pos_kernel(const float* __restrict__ pos, const float* __restrict__ bias,
                                                       float* __restrict__ out, long long n, int D) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) out[i] = pos[i] + bias[i % D];
}

__global__ void pack_qscale_kernel(const float* __restrict__ scale, float* __restrict__ out, int heads) {
  const int i = threadIdx.x;
  if (i < heads) out[i] = expf(fminf(scale[i], 4.605170185988092f));       // exp(clamp(scale, max=ln 100)), swinv2.py:125-126
}

// dst[k, n] = cvt(src[n, k]) (src fp32 [N, K] row-major, dst 16-bit [K, ldd]): the transposed copies the dgrad GEMMs read
__global__ void __launch_bounds__(256) pack_transposed_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int N, int K,
                                                              int ldd, int f16) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + 8 * i, k = k0 + tx;
    tile[ty + 8 * i][tx] = (n < N && k < K) ? __ldg(src + static_cast<size_t>(n) * K + k) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty + 8 * i, n = n0 + tx;
    if (k < K && n < N) dst[static_cast<size_t>(k) * ldd + n] = cvt16(tile[tx][ty + 8 * i], f16);
  }
}

inline unsigned nblk(long long n) { return static_cast<unsigned>((n + 255) / 256); }
}  // namespace

int launch_pack_transposed(const float* src, void* dst, int N, int K, int ldd, int f16, cudaStream_t st) {
  dim3 grid((K + 31) / 32, (N + 31) / 32);
  pack_transposed_kernel<<<grid, 256, 0, st>>>(src, static_cast<uint16_t*>(dst), N, K, ldd, f16);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

int launch_pack_rows(const float* src, void* dst, int rows, int K, int ldd, int dup, int mode, int a, int b, int f16,
                     cudaStream_t st) {
  pack_rows_kernel<<<nblk(static_cast<long long>(rows) * K), 256, 0, st>>>(src, static_cast<uint16_t*>(dst), rows, K, ldd, dup, mode,
                                                                          a, b, f16);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

int launch_pack_embed(const float* src, void* dst, int D, int C, int pp, int k_embed, int split, int f16, cudaStream_t st) {
  pack_embed_kernel<<<nblk(static_cast<long long>(D) * k_embed), 256, 0, st>>>(src, static_cast<uint16_t*>(dst), D, C, pp, k_embed,
                                                                              split, f16);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

int launch_pack_pos(const float* pos, const float* bias, float* out, long long n, int D, cudaStream_t st) {
  pack_pos_kernel<<<nblk(n), 256, 0, st>>>(pos, bias, out, n, D);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

int launch_pack_qscale(const float* scale, float* out, int heads, cudaStream_t st) {
  SWB_REQUIRE(heads <= 1024, "pack_qscale: too many heads");
  pack_qscale_kernel<<<1, 1024, 0, st>>>(scale, out, heads);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
