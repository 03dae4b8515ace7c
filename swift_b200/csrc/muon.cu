// Optimiser step of the sCM training configuration on the device (reference: training/optimizers/muon.py, selected by
// configs/experiment/era5-swinv2-1.4-scm.yaml "override /optimizer: muon").
//
// Muon (muon.py:5-45, :218-241): SGD momentum, Nesterov blend, then the update matrix is orthogonalised by five steps of the
// quintic Newton-Schulz iteration in bf16
//     X <- G / (|G|_F + 1e-7)  (transposed if tall);   A = X X^T;   B = b A + c A A;   X <- a X + B X
// -- a chain of bf16 GEMMs, here the tcgen05 GEMM of the forecast path (fp32 accumulation, split-K over the long
// contraction of X X^T, whose output has only a handful of tiles) with the roundings of the reference's bf16 tensors applied
// by the small combination kernels below.  AuxAdam (muon.py:147-152, :243-266): one fused elementwise kernel per tensor.
#include "common.h"
#include "gemm_sm100.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>
#include <math.h>

namespace swb {

namespace {
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint16_t bf16_bits(float x) {
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ float bits_f(uint16_t u) { return __uint_as_float(static_cast<uint32_t>(u) << 16); }

// momentum <- lerp(momentum, grad, 1 - beta);  u = nesterov ? lerp(grad, momentum, beta) : momentum;  U16 = bf16(u);
// part[block] = sum of bf16(u)^2 over the block (muon.py:37-39, :18, :23)
__global__ void __launch_bounds__(256) muon_prepare_kernel(const float* __restrict__ grad, float* __restrict__ mom,
                                                           uint16_t* __restrict__ U16, float* __restrict__ part, long long n,
                                                           float beta, int nesterov) {
  float ss = 0.f;
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;      // group of 4 elements (n % 4 == 0)
  if (4 * i < n) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(grad) + i);
    float4 m4 = reinterpret_cast<float4*>(mom)[i];
    const float g[4] = {g4.x, g4.y, g4.z, g4.w};
    float m[4] = {m4.x, m4.y, m4.z, m4.w};
    uint16_t b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m[j] = fmaf(1.0f - beta, g[j] - m[j], m[j]);
      const float u = nesterov ? fmaf(beta, m[j] - g[j], g[j]) : m[j];
      b[j] = bf16_bits(u);
      const float r = bits_f(b[j]);
      ss = fmaf(r, r, ss);
    }
    reinterpret_cast<float4*>(mom)[i] = make_float4(m[0], m[1], m[2], m[3]);
    reinterpret_cast<uint2*>(U16)[i] = make_uint2(static_cast<uint32_t>(b[0]) | (static_cast<uint32_t>(b[1]) << 16),
                                                  static_cast<uint32_t>(b[2]) | (static_cast<uint32_t>(b[3]) << 16));
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w];
    part[blockIdx.x] = a;
  }
}

// one block: norm2[0] = sum of the partials (fixed order, fp64 accumulation)
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  __shared__ double red[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += static_cast<double>(part[i]);
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = static_cast<float>(red[0]);
}

// X16 <- bf16(X16 / (sqrt(norm2) + 1e-7))
__global__ void __launch_bounds__(256) muon_normalize_kernel(uint16_t* __restrict__ X16, const float* __restrict__ norm2, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;      // group of 8 elements
  if (8 * i >= n) return;
  const float inv = 1.0f / (bf16_round(sqrtf(norm2[0])) + 1e-7f);
  uint4 r = reinterpret_cast<uint4*>(X16)[i];
  uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    w[j] = pack_bf16x2(__uint_as_float(w[j] << 16) * inv, __uint_as_float(w[j] & 0xffff0000u) * inv);
  reinterpret_cast<uint4*>(X16)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// out16[i] = bf16(alpha * x16[i] + beta * bf16(sum_s y[s * stride + i]))     (x16 may be NULL: alpha term dropped)
// 8 elements per thread: 16-byte loads / stores (n and stride are multiples of 8: every extent is a multiple of 8)
__global__ void __launch_bounds__(256) muon_combine_kernel(uint16_t* out16, float alpha, const uint16_t* x16,
                                                           float beta, const float* __restrict__ y, int splits, long long stride,
                                                           long long n8, long long per8) {
  // batched: problem b = i / per8 holds its `splits` partial products at y[(b * splits + s) * stride ..]; stride == 8 * per8
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const long long b = i / per8, j = i - b * per8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < splits; ++s) {
    const float4* p = reinterpret_cast<const float4*>(y + (b * splits + s) * stride) + 2 * j;
    const float4 a = __ldg(p), b = __ldg(p + 1);
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
  float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (x16) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(x16) + i);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xv[2 * j] = __uint_as_float(w[j] << 16);
      xv[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v0 = beta * bf16_round(acc[2 * j]), v1 = beta * bf16_round(acc[2 * j + 1]);
    if (x16) {
      v0 = bf16_round(v0) + bf16_round(alpha * xv[2 * j]);
      v1 = bf16_round(v1) + bf16_round(alpha * xv[2 * j + 1]);
    }
    o[j] = pack_bf16x2(v0, v1);
  }
  reinterpret_cast<uint4*>(out16)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// p <- p (1 - lr wd) - lr scale X   (muon.py:42, :232-233); X is the orthogonalised update in the parameter's orientation
__global__ void __launch_bounds__(256) muon_apply_kernel(float* __restrict__ p, const uint16_t* __restrict__ X16, long long n, float decay,
                                                         float step) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;      // group of 4 elements
  if (4 * i >= n) return;
  float4 v = reinterpret_cast<float4*>(p)[i];
  const uint2 x = __ldg(reinterpret_cast<const uint2*>(X16) + i);
  v.x = fmaf(v.x, decay, -step * __uint_as_float(x.x << 16));
  v.y = fmaf(v.y, decay, -step * __uint_as_float(x.x & 0xffff0000u));
  v.z = fmaf(v.z, decay, -step * __uint_as_float(x.y << 16));
  v.w = fmaf(v.w, decay, -step * __uint_as_float(x.y & 0xffff0000u));
  reinterpret_cast<float4*>(p)[i] = v;
}

// AuxAdam (muon.py:147-152, :261-266): buf1 <- lerp(buf1, g, 1-b1); buf2 <- lerp(buf2, g^2, 1-b2);
// p <- p (1 - lr wd) - lr (buf1 / c1) / (sqrt(buf2 / c2) + eps)
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                        float wd, float c1, float c2) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = fmaf(1.0f - b1, gi - m[i], m[i]);
  const float vi = fmaf(1.0f - b2, gi * gi - v[i], v[i]);
  m[i] = mi;
  v[i] = vi;
  const float upd = (mi / c1) / (sqrtf(vi / c2) + eps);
  p[i] = fmaf(p[i], 1.0f - lr * wd, -lr * upd);
}

// Muon for a VECTOR-shaped parameter (1 x n or n x 1; Swift-B: the [1, heads, 1, 1] logit scales): with one row the
// Newton-Schulz products collapse to scalars -- A = X X^T = |X|^2, B = b A + c A A, X <- a X + B X -- evaluated with the same
// bf16 roundings as the matrix path.  One block per vector.
__global__ void __launch_bounds__(256) muon_vector_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ mom,
                                                          int n, float beta, int nesterov, int ns_steps, float decay, float step) {
  __shared__ float red[256];
  auto block_sum = [&](float v) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    const float r = red[0];
    __syncthreads();
    return r;
  };
  constexpr int kMax = 16;                             // n <= 4096
  float x[kMax];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMax; ++j) {
    const int i = threadIdx.x + 256 * j;
    x[j] = 0.f;
    if (i < n) {
      const float g = grad[i];
      const float m = fmaf(1.0f - beta, g - mom[i], mom[i]);
      mom[i] = m;
      x[j] = bf16_round(nesterov ? fmaf(beta, m - g, g) : m);
      ss = fmaf(x[j], x[j], ss);
    }
  }
  const float inv = 1.0f / (bf16_round(sqrtf(block_sum(ss))) + 1e-7f);
#pragma unroll
  for (int j = 0; j < kMax; ++j) x[j] = bf16_round(x[j] * inv);
  const float a = 3.4445f, b = -4.7750f, c = 2.0315f;
  for (int it = 0; it < ns_steps; ++it) {
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kMax; ++j) s2 = fmaf(x[j], x[j], s2);
    const float A = bf16_round(block_sum(s2));
    const float B = bf16_round(bf16_round(b * A) + bf16_round(c * bf16_round(A * A)));
#pragma unroll
    for (int j = 0; j < kMax; ++j) x[j] = bf16_round(bf16_round(a * x[j]) + bf16_round(B * x[j]));
  }
#pragma unroll
  for (int j = 0; j < kMax; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < n) p[i] = fmaf(p[i], decay, -step * x[j]);
  }
}

inline unsigned blocks_for(long long n) { return static_cast<unsigned>((n + 255) / 256); }
inline size_t up(size_t v) { return (v + 1023) / 1024 * 1024; }

int ns_splits(int m, int k, int batch) {          // X X^T: [m, m] output tiles of 256 x 176 per problem, contraction k
  const int tiles = ((m + 255) / 256) * ((m + 175) / 176) * batch;
  int s = 1;
  while (s < 16 && tiles * s < 96 && k % (2 * s * kBlockK) == 0 && k / (2 * s) >= 512) s *= 2;
  return s;
}

struct MuonWs {
  size_t U, UT, A32, A16, B16, BX32, part, total;
};
MuonWs carve_muon(int rows, int cols, int batch) {
  const size_t m = std::min(rows, cols), n = std::max(rows, cols), nb = batch;
  MuonWs w;
  size_t off = 0;
  auto take = [&](size_t b) {
    size_t o = off;
    off = up(off + b);
    return o;
  };
  w.U = take(nb * m * n * 2);
  w.UT = take(nb * m * n * 2);
  w.A32 = take(nb * static_cast<size_t>(ns_splits(static_cast<int>(m), static_cast<int>(n), batch)) * m * m * 4);
  w.A16 = take(nb * m * m * 2);
  w.B16 = take(nb * m * m * 2);
  w.BX32 = take(nb * m * n * 4);
  w.part = take(nb * (static_cast<size_t>(blocks_for(static_cast<long long>(m) * n / 4)) + 4) * 4);
  w.total = off;
  return w;
}

int gemm_f32(int tile, const void* A, int lda, const void* W, int ldw, float* out, int M, int N, int K, int splits, int batch,
             cudaStream_t st) {
  GemmParams p = {};
  p.M = M;
  p.N = N;
  p.K = K / splits;
  p.out0 = out;
  p.ldo = N;
  p.splits = splits;
  p.batch = batch;
  return launch_gemm(EPI_STORE_F32, tile, 0, A, lda, W, ldw, p, st);
}

}  // namespace

size_t muon_workspace_bytes(int rows, int cols, int batch) { return carve_muon(rows, cols, batch).total; }

// `batch` parameters of one shape [rows, cols] (HOST arrays of device pointers): the elementwise kernels run per matrix,
// the Newton-Schulz GEMMs / combinations / transposes once for the whole stack (batched GEMM: enough tiles to fill the GPU).
int launch_muon_step(float* const* param, const float* const* grad, float* const* momentum, int batch, int rows, int cols, float lr,
                     float weight_decay, float beta, int nesterov, int ns_steps, void* workspace, size_t ws_bytes, cudaStream_t st) {
  SWB_REQUIRE(rows >= 8 && cols >= 8 && rows % 8 == 0 && cols % 8 == 0,
              "muon_step: %d x %d unsupported (both extents multiples of 8; vectors are handled by the host wrapper)", rows, cols);
  SWB_REQUIRE(batch >= 1 && batch <= 4096, "muon_step: batch %d out of range", batch);
  const MuonWs w = carve_muon(rows, cols, batch);
  SWB_REQUIRE(ws_bytes >= w.total && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0,
              "muon_step: workspace too small (%zu < %zu) or not 1024-byte aligned", ws_bytes, w.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const long long n_el = static_cast<long long>(rows) * cols;
  uint16_t* U = reinterpret_cast<uint16_t*>(ws + w.U);       // [batch] parameter orientation [rows, cols]
  uint16_t* UT = reinterpret_cast<uint16_t*>(ws + w.UT);     // [batch] [cols, rows]
  float* part = reinterpret_cast<float*>(ws + w.part);
  const unsigned nb = blocks_for(n_el / 4);
  for (int i = 0; i < batch; ++i) {
    SWB_REQUIRE(param[i] && grad[i] && momentum[i], "muon_step: NULL pointer for matrix %d", i);
    float* pi = part + static_cast<size_t>(i) * (nb + 4);
    muon_prepare_kernel<<<nb, 256, 0, st>>>(grad[i], momentum[i], U + i * n_el, pi, n_el, beta, nesterov);
    sum_partials_kernel<<<1, 256, 0, st>>>(pi, static_cast<int>(nb), pi + nb);
    muon_normalize_kernel<<<blocks_for(n_el / 8), 256, 0, st>>>(U + i * n_el, pi + nb, n_el);
  }
  SWB_CHECK_CUDA(cudaGetLastError());
  int rc = launch_transpose16_batched(U, rows, cols, UT, batch, st);
  if (rc) return rc;
  // X: the wide orientation [m, n] (m <= n);  XT: [n, m]
  const bool tall = rows > cols;
  uint16_t* X = tall ? UT : U;
  uint16_t* XT = tall ? U : UT;
  const int m = tall ? cols : rows, n = tall ? rows : cols;
  float* A32 = reinterpret_cast<float*>(ws + w.A32);
  uint16_t* A16 = reinterpret_cast<uint16_t*>(ws + w.A16);
  uint16_t* B16 = reinterpret_cast<uint16_t*>(ws + w.B16);
  float* BX32 = reinterpret_cast<float*>(ws + w.BX32);
  const float a = 3.4445f, b = -4.7750f, c = 2.0315f;
  const long long mm = static_cast<long long>(m) * m, mn = static_cast<long long>(m) * n;
  const int S = ns_splits(m, n, batch);
  for (int it = 0; it < ns_steps; ++it) {
    if ((rc = gemm_f32(2, X, n, X, n, A32, m, m, n, S, batch, st))) return rc;                        // A = X X^T
    muon_combine_kernel<<<blocks_for(batch * mm / 8), 256, 0, st>>>(A16, 0.f, nullptr, 1.0f, A32, S, mm, batch * mm / 8, mm / 8);
    if ((rc = gemm_f32(2, A16, m, A16, m, A32, m, m, m, 1, batch, st))) return rc;                    // A A (A symmetric)
    muon_combine_kernel<<<blocks_for(batch * mm / 8), 256, 0, st>>>(B16, b, A16, c, A32, 1, mm, batch * mm / 8, mm / 8);   // B = b A + c A A
    if ((rc = gemm_f32(3, B16, m, XT, m, BX32, m, n, m, 1, batch, st))) return rc;                    // B X
    muon_combine_kernel<<<blocks_for(batch * mn / 8), 256, 0, st>>>(X, a, X, 1.0f, BX32, 1, mn, batch * mn / 8, mn / 8);   // X = a X + B X
    SWB_CHECK_CUDA(cudaGetLastError());
    if ((rc = launch_transpose16_batched(X, m, n, XT, batch, st))) return rc;
  }
  const float scale = sqrtf(fmaxf(1.0f, static_cast<float>(rows) / static_cast<float>(cols)));
  for (int i = 0; i < batch; ++i)
    muon_apply_kernel<<<blocks_for(n_el / 4), 256, 0, st>>>(param[i], U + i * n_el, n_el, 1.0f - lr * weight_decay, lr * scale);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// rows x cols with min(rows, cols) == 1
int launch_muon_vector(float* param, const float* grad, float* momentum, int rows, int cols, float lr, float weight_decay, float beta,
                       int nesterov, int ns_steps, cudaStream_t st) {
  SWB_REQUIRE((rows == 1 || cols == 1) && rows * cols >= 1 && rows * cols <= 4096,
              "muon_vector: %d x %d is not a vector of at most 4096 elements", rows, cols);
  const float scale = sqrtf(fmaxf(1.0f, static_cast<float>(rows) / static_cast<float>(cols)));
  muon_vector_kernel<<<1, 256, 0, st>>>(param, grad, momentum, rows * cols, beta, nesterov, ns_steps, 1.0f - lr * weight_decay,
                                        lr * scale);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

int launch_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
                     int step, cudaStream_t st) {
  SWB_REQUIRE(step >= 1 && n > 0, "adam_step: step must be >= 1");
  const float c1 = 1.0f - powf(b1, static_cast<float>(step)), c2 = 1.0f - powf(b2, static_cast<float>(step));
  adam_step_kernel<<<blocks_for(n), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, c1, c2);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
