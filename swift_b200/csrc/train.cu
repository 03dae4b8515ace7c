// Reverse-mode companions of the denoiser kernels: what `F_x.backward(cot)` of the sCM training step needs between the
// tcgen05 GEMMs (reference: training/loss.py:226-260 calls net(...) grad-enabled, trainer.py:199-219 runs .backward()).
//
// The training path computes in bf16 operands with fp32 accumulation (the reference trains under bf16 autocast,
// configs/experiment/era5-swinv2-1.4-scm.yaml), the residual stream and every gradient between kernels in fp32 or as
// bf16 GEMM operands:
//     dgrad  dX = dY W          the forecast GEMM kernel over transposed bf16 weight copies
//     wgrad  dW = dY^T X        the same kernel with split-K over the tokens, operands transposed by transpose16_kernel
// and the kernels below apply the derivative rules of LayerNorm-modulation-residual, SwiGLU, the q/k cosine normalisation
// (inside the attention backward, attention_bwd.cu), the head's pixel shuffle and the conditioning MLP.
// Every cross-row reduction (modulation gain / bias gradients, bias gradients, logit-scale gradients) is two-stage with
// a fixed summation order: results are bit-reproducible.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace swb {

namespace {
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf2f(uint16_t u) { return __uint_as_float(static_cast<uint32_t>(u) << 16); }
__device__ __forceinline__ uint16_t f2bf(float x) {
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ void unpack8_bf16(uint4 r, float* o) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[2 * j] = __uint_as_float(w[j] << 16);
    o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
}  // namespace

// =========================================================================================================
// SwiGLU (models/swinv2.py:99-100) on the un-fused training path: gu [M, 2*Dff] bf16 = [gate | up] in the reference's
// column order (w1 is used un-permuted here), h = silu(gate) * up.
__global__ void __launch_bounds__(256) swiglu_fwd_train_kernel(const uint16_t* __restrict__ gu, uint16_t* __restrict__ h,
                                                               long long M, int Dff) {
  const int c8 = Dff >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * c8) return;
  const long long r = idx / c8;
  const int c = static_cast<int>(idx - r * c8);
  const uint16_t* row = gu + r * 2 * Dff;
  float g[8], u[8], o[8];
  unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(row) + c), g);
  unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(row + Dff) + c), u);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = g[j] * sigmoid_f(g[j]) * u[j];
  reinterpret_cast<uint4*>(h + r * Dff)[c] = pack8_bf16(o);
}

int launch_swiglu_fwd_train(const void* gu, void* h, int M, int Dff, cudaStream_t stream) {
  SWB_REQUIRE(Dff % 8 == 0, "swiglu_fwd_train: Dff %% 8 != 0");
  const long long total = static_cast<long long>(M) * (Dff / 8);
  swiglu_fwd_train_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint16_t*>(gu), static_cast<uint16_t*>(h), M, Dff);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// d gate = dh * up * silu'(gate), d up = dh * silu(gate);  dh fp32 [M, Dff] -> dgu bf16 [M, 2*Dff] ([d gate | d up])
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const float* __restrict__ dh, const uint16_t* __restrict__ gu,
                                                         uint16_t* __restrict__ dgu, long long M, int Dff) {
  const int c8 = Dff >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= M * c8) return;
  const long long r = idx / c8;
  const int c = static_cast<int>(idx - r * c8);
  const uint16_t* row = gu + r * 2 * Dff;
  float g[8], u[8], dg[8], du[8];
  unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(row) + c), g);
  unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(row + Dff) + c), u);
  const float4 a = __ldg(reinterpret_cast<const float4*>(dh + r * Dff) + 2 * c);
  const float4 b = __ldg(reinterpret_cast<const float4*>(dh + r * Dff) + 2 * c + 1);
  const float d[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = sigmoid_f(g[j]);
    du[j] = d[j] * g[j] * s;
    dg[j] = d[j] * u[j] * s * (1.0f + g[j] * (1.0f - s));
  }
  uint16_t* orow = dgu + r * 2 * Dff;
  reinterpret_cast<uint4*>(orow)[c] = pack8_bf16(dg);
  reinterpret_cast<uint4*>(orow + Dff)[c] = pack8_bf16(du);
}

int launch_swiglu_bwd(const float* dh, const void* gu, void* dgu, int M, int Dff, cudaStream_t stream) {
  SWB_REQUIRE(Dff % 8 == 0, "swiglu_bwd: Dff %% 8 != 0");
  const long long total = static_cast<long long>(M) * (Dff / 8);
  swiglu_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      dh, static_cast<const uint16_t*>(gu), static_cast<uint16_t*>(dgu), M, Dff);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// =========================================================================================================
// q / k cosine normalisation + logit scale + per-head packing on the training path (models/swinv2.py:119-127).
// raw fp32 [M, 3*D] straight from the to_qkv GEMM with the reference's column order (head, part, d)
//   -> packed bf16 [3][heads][M][pad]: q_hat * qscale[head], k_hat, v (columns hd..pad-1 zero), what both attention
//      kernels read, and invn fp32 [2][heads][M] = 1 / max(|q|, 1e-12), 1 / max(|k|, 1e-12) for the backward.
// One block per token row; a warp per (head, part).
__global__ void __launch_bounds__(128) qkv_pack_train_kernel(const float* __restrict__ raw, const float* __restrict__ qscale,
                                                             uint16_t* __restrict__ packed, float* __restrict__ invn, int M,
                                                             int heads, int hd, int pad) {
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* r = raw + static_cast<size_t>(row) * 3 * heads * hd;
  for (int hp = warp; hp < heads * 3; hp += 4) {
    const int head = hp / 3, part = hp - head * 3;
    const float* src = r + (head * 3 + part) * hd;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = lane + 32 * i;
      v[i] = d < hd ? __ldg(src + d) : 0.f;
    }
    float mul = 1.0f;
    if (part < 2) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) ss = fmaf(v[i], v[i], ss);
      ss = wsum(ss);
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      if (lane == 0) invn[(static_cast<size_t>(part) * heads + head) * M + row] = inv;
      mul = part == 0 ? inv * __ldg(qscale + head) : inv;
    }
    uint16_t* dst = packed + ((static_cast<size_t>(part) * heads + head) * M + row) * pad;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = lane + 32 * i;
      if (d < pad) dst[d] = d < hd ? f2bf(v[i] * mul) : static_cast<uint16_t>(0);
    }
  }
}

int launch_qkv_pack_train(const float* raw, const float* qscale, void* packed, float* invn, int M, int heads, int hd,
                          int pad, cudaStream_t stream) {
  SWB_REQUIRE(hd <= 128 && pad <= 128 && pad >= hd, "qkv_pack_train: head dim %d / pad %d unsupported", hd, pad);
  qkv_pack_train_kernel<<<M, 128, 0, stream>>>(raw, qscale, static_cast<uint16_t*>(packed), invn, M, heads, hd, pad);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// =========================================================================================================
// 16-bit transpose: in [R, C] (row pitch ldi) -> out [C, R] (row pitch ldo).  The weight-gradient GEMMs contract over the
// token axis, and the tcgen05 GEMM takes both operands K-major: dY^T [features, tokens] and X^T [features, tokens].
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ in, int R, int C, long long ldi,
                                                          uint16_t* __restrict__ out, long long ldo, long long batch_in,
                                                          long long batch_out) {
  __shared__ uint16_t tile[64][66];
  in += blockIdx.z * batch_in;               // batched: problem z at element offsets z * batch_in / z * batch_out
  out += blockIdx.z * batch_out;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = warp * 8 + i;
    uint32_t w = 0u;
    if (r0 + r < R && c0 + 2 * lane < C)       // C is even: a pair never straddles the edge
      w = __ldg(reinterpret_cast<const uint32_t*>(in + static_cast<long long>(r0 + r) * ldi + c0) + lane);
    *reinterpret_cast<uint32_t*>(&tile[r][2 * lane]) = w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = warp * 8 + i;
    if (c0 + c < C && r0 + 2 * lane < R) {
      const uint32_t w = static_cast<uint32_t>(tile[2 * lane][c]) | (static_cast<uint32_t>(tile[2 * lane + 1][c]) << 16);
      reinterpret_cast<uint32_t*>(out + static_cast<long long>(c0 + c) * ldo + r0)[lane] = w;
    }
  }
}

int launch_transpose16(const void* in, int R, int C, long long ldi, void* out, long long ldo, cudaStream_t stream) {
  SWB_REQUIRE(R % 2 == 0 && C % 2 == 0 && ldi % 2 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 3) == 0,
              "transpose16: even extents / pitches and 4-byte aligned buffers required (R=%d C=%d)", R, C);
  dim3 grid((C + 63) / 64, (R + 63) / 64);
  transpose16_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(in), R, C, ldi, static_cast<uint16_t*>(out), ldo, 0, 0);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// `batch` contiguous [R, C] matrices -> `batch` contiguous [C, R] matrices
int launch_transpose16_batched(const void* in, int R, int C, void* out, int batch, cudaStream_t stream) {
  SWB_REQUIRE(R % 2 == 0 && C % 2 == 0 && batch >= 1 && batch <= 65535, "transpose16_batched: even extents required (R=%d C=%d)", R, C);
  dim3 grid((C + 63) / 64, (R + 63) / 64, batch);
  transpose16_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(in), R, C, C, static_cast<uint16_t*>(out), R,
                                               static_cast<long long>(R) * C, static_cast<long long>(R) * C);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// =========================================================================================================
// Backward of  y = LayerNorm(b) * gain[s] + bias[s]  (ModulatedNorm with the affine and the modulation folded into
// per-sample gain / bias, models/swinv2.py:83-86) for rows of sample s = row / tokens:
//     n = (b - mean) rstd,  dn = dy gain,  db = rstd (dn - mean(dn) - n mean(dn n)),
//     d gain[s] += sum_rows dy n,  d bias[s] += sum_rows dy.
// dy = dx (+ add): the gradient arriving at the residual stream; with `add` (the dX a dgrad GEMM has just produced for
// the same rows) the sum is also written back to dx, so the residual accumulation costs no extra pass.
// Block = 8 warps x 4 rows = 32 rows of one sample; the block's column sums go to part[blockIdx][2][D].
constexpr int kLnBwdRowsPerWarp = 4;
constexpr int kLnBwdRowsPerBlock = 8 * kLnBwdRowsPerWarp;
constexpr int kLnBwdMaxIter = 9;      // D <= 9 * 128

__global__ void __launch_bounds__(256) ln_bwd_kernel(float* __restrict__ dx, const float* __restrict__ add,
                                                     const float* __restrict__ branch, const float* __restrict__ gain,
                                                     uint16_t* __restrict__ db16, float* __restrict__ part, int M, int D,
                                                     int tokens, float eps) {
  extern __shared__ float red[];                       // [8 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_base = blockIdx.x * kLnBwdRowsPerBlock + warp * kLnBwdRowsPerWarp;
  const int s = (blockIdx.x * kLnBwdRowsPerBlock) / tokens;
  const float4* g4 = reinterpret_cast<const float4*>(gain + static_cast<size_t>(s) * D);
  const int nv = D >> 2;                               // float4 groups per row
  float ag[kLnBwdMaxIter][4], ab[kLnBwdMaxIter][4];
#pragma unroll
  for (int i = 0; i < kLnBwdMaxIter; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ag[i][j] = ab[i][j] = 0.f;
  const float invD = 1.0f / static_cast<float>(D);
  for (int rr = 0; rr < kLnBwdRowsPerWarp; ++rr) {
    const int row = row_base + rr;
    if (row >= M) break;
    const float4* b4 = reinterpret_cast<const float4*>(branch + static_cast<size_t>(row) * D);
    float4* dx4 = reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * D);
    const float4* a4 = add ? reinterpret_cast<const float4*>(add + static_cast<size_t>(row) * D) : nullptr;
    float bv[kLnBwdMaxIter][4], dv[kLnBwdMaxIter][4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kLnBwdMaxIter; ++i) {
      const int c = lane + 32 * i;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f), d = b;
      if (c < nv) {
        b = __ldg(b4 + c);
        d = dx4[c];
        if (a4) {
          const float4 a = __ldg(a4 + c);
          d.x += a.x; d.y += a.y; d.z += a.z; d.w += a.w;
          dx4[c] = d;
        }
      }
      bv[i][0] = b.x; bv[i][1] = b.y; bv[i][2] = b.z; bv[i][3] = b.w;
      dv[i][0] = d.x; dv[i][1] = d.y; dv[i][2] = d.z; dv[i][3] = d.w;
      sum += (b.x + b.y) + (b.z + b.w);
    }
    const float mean = wsum(sum) * invD;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kLnBwdMaxIter; ++i)
      if (lane + 32 * i < nv)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = bv[i][j] - mean;
          var = fmaf(t, t, var);
        }
    const float rstd = rsqrtf(wsum(var) * invD + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnBwdMaxIter; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 g = __ldg(g4 + c);
        const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float n = (bv[i][j] - mean) * rstd;
          const float dy = dv[i][j];
          ag[i][j] = fmaf(dy, n, ag[i][j]);
          ab[i][j] += dy;
          const float dn = dy * gg[j];
          bv[i][j] = n;                   // keep n
          dv[i][j] = dn;                  // keep dn
          s1 += dn;
          s2 = fmaf(dn, n, s2);
        }
      }
    }
    s1 = wsum(s1) * invD;
    s2 = wsum(s2) * invD;
    uint2* o2 = reinterpret_cast<uint2*>(db16 + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < kLnBwdMaxIter; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = rstd * (dv[i][j] - s1 - bv[i][j] * s2);
        o2[c] = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
      }
    }
  }
  // block reduction of the column sums, fixed order over the 8 warps
  float* mine = red + static_cast<size_t>(warp) * 2 * D;
#pragma unroll
  for (int i = 0; i < kLnBwdMaxIter; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      reinterpret_cast<float4*>(mine)[c] = make_float4(ag[i][0], ag[i][1], ag[i][2], ag[i][3]);
      reinterpret_cast<float4*>(mine + D)[c] = make_float4(ab[i][0], ab[i][1], ab[i][2], ab[i][3]);
    }
  }
  __syncthreads();
  float* out = part + static_cast<size_t>(blockIdx.x) * 2 * D;
  for (int c = threadIdx.x; c < 2 * D; c += 256) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[static_cast<size_t>(w) * 2 * D + c];
    out[c] = a;
  }
}

// out[g, c] (+)= sum_{p < per} part[(g * per + p) * width + c]   (fixed order)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int per, int width,
                                                              float* __restrict__ out, int groups, int accumulate) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(groups) * width) return;
  const int g = static_cast<int>(idx / width), c = static_cast<int>(idx - static_cast<long long>(g) * width);
  const float* p = part + static_cast<size_t>(g) * per * width + c;
  float a = 0.f;
  for (int i = 0; i < per; ++i) a += p[static_cast<size_t>(i) * width];
  out[idx] = accumulate ? out[idx] + a : a;
}

int launch_reduce_partials(const float* part, int per, int width, float* out, int groups, int accumulate,
                           cudaStream_t stream) {
  const long long total = static_cast<long long>(groups) * width;
  if (total <= 0 || per <= 0) return SWB_OK;
  reduce_partials_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(part, per, width, out, groups,
                                                                                       accumulate);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

size_t ln_bwd_partial_floats(int M, int D) {
  return static_cast<size_t>((M + kLnBwdRowsPerBlock - 1) / kLnBwdRowsPerBlock) * 2 * D;
}

// dgain_dbias: [2][B][D] is NOT the layout; outputs are dgain [B, D] and dbias [B, D] (accumulated when `accumulate`).
int launch_ln_bwd(float* dx, const float* add, const float* branch, const float* gain, void* db16, float* part,
                  float* dgain, float* dbias, int M, int D, int tokens, float eps, int accumulate, cudaStream_t stream) {
  SWB_REQUIRE(D % 4 == 0 && D <= kLnBwdMaxIter * 128, "ln_bwd: D=%d unsupported (multiple of 4, <= %d)", D,
              kLnBwdMaxIter * 128);
  SWB_REQUIRE(tokens % kLnBwdRowsPerBlock == 0 && M % tokens == 0, "ln_bwd: tokens per sample (%d) must be a multiple of %d",
              tokens, kLnBwdRowsPerBlock);
  const int blocks = M / kLnBwdRowsPerBlock;
  const size_t smem = static_cast<size_t>(8) * 2 * D * sizeof(float);
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * kLnBwdMaxIter * 128 * 4));
    attr_done.set(true);
  }
  ln_bwd_kernel<<<blocks, 256, smem, stream>>>(dx, add, branch, gain, static_cast<uint16_t*>(db16), part, M, D, tokens, eps);
  SWB_CHECK_CUDA(cudaGetLastError());
  // part: [blocks][2][D] -> per sample [2][D]; written as dgain[s] / dbias[s] by two strided reductions
  const int B = M / tokens, per = tokens / kLnBwdRowsPerBlock;
  // view part as [B][per][2*D]: reduce over `per` -> tmp [B][2*D]; dgain/dbias are separate [B, D] arrays, so reduce twice
  // with width D over the two halves is not contiguous; instead reduce into the first 2*D*B floats of a scratch region
  // that follows the partials (the caller sizes `part` with ln_bwd_partial_floats + 2*B*D).
  float* tmp = part + ln_bwd_partial_floats(M, D);
  int rc = launch_reduce_partials(part, per, 2 * D, tmp, B, 0, stream);
  if (rc) return rc;
  for (int s = 0; s < B; ++s) {
    rc = launch_reduce_partials(tmp + static_cast<size_t>(s) * 2 * D, 1, D, dgain + static_cast<size_t>(s) * D, 1, accumulate, stream);
    if (rc) return rc;
    rc = launch_reduce_partials(tmp + static_cast<size_t>(s) * 2 * D + D, 1, D, dbias + static_cast<size_t>(s) * D, 1, accumulate, stream);
    if (rc) return rc;
  }
  return SWB_OK;
}

// =========================================================================================================
// Column sums of an fp32 [R, W] matrix (bias gradients, the position-table gradient over samples): partial sums per
// block of 64 rows, then reduce_partials.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ in, int R, int W, float* __restrict__ part) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= W) return;
  const int r0 = blockIdx.y * 64;
  float a = 0.f;
  for (int r = r0; r < r0 + 64 && r < R; ++r) a += __ldg(in + static_cast<size_t>(r) * W + c);
  part[static_cast<size_t>(blockIdx.y) * W + c] = a;
}

size_t colsum_partial_floats(int R, int W) { return static_cast<size_t>((R + 63) / 64) * W; }

int launch_colsum(const float* in, int R, int W, float* part, float* out, int accumulate, cudaStream_t stream) {
  dim3 grid((W + 255) / 256, (R + 63) / 64);
  colsum_partial_kernel<<<grid, 256, 0, stream>>>(in, R, W, part);
  SWB_CHECK_CUDA(cudaGetLastError());
  return launch_reduce_partials(part, (R + 63) / 64, W, out, 1, accumulate, stream);
}

// out[i] = (accumulate ? out[i] : 0) + sum_s part[s * stride + i]: the split-K partial products of a weight gradient
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits, long long stride,
                                                            float* __restrict__ out, long long n4, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = accumulate ? reinterpret_cast<float4*>(out)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + s * stride) + i);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = a;
}

int launch_splitk_reduce(const float* part, int splits, long long stride, float* out, long long n, int accumulate,
                         cudaStream_t stream) {
  SWB_REQUIRE(n % 4 == 0 && stride % 4 == 0, "splitk_reduce: element counts must be multiples of 4");
  const long long n4 = n / 4;
  splitk_reduce_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(part, splits, stride, out, n4, accumulate);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// dst[r, c] = a[r, c] + b[r, c] (fp32), optionally also as bf16 (the A operand of a following wgrad / the embed backward)
__global__ void __launch_bounds__(256) add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      float* __restrict__ dst, uint16_t* __restrict__ dst16, long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 x = __ldg(reinterpret_cast<const float4*>(a) + i);
  if (b) {
    const float4 y = __ldg(reinterpret_cast<const float4*>(b) + i);
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
  }
  if (dst) reinterpret_cast<float4*>(dst)[i] = x;
  if (dst16) reinterpret_cast<uint2*>(dst16)[i] = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
}

int launch_add_f32(const float* a, const float* b, float* dst, void* dst16, long long n, cudaStream_t stream) {
  SWB_REQUIRE(n % 4 == 0, "add_f32: n %% 4 != 0");
  add_f32_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, stream>>>(a, b, dst, static_cast<uint16_t*>(dst16), n / 4);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// =========================================================================================================
// Head backward input: the cotangent of the network output, NCHW fp32 [B, C, H, W], as token rows of the head GEMM's
// output space: dF[row = (b, gy, gx), col = c*p1*p2 + i*p2 + j] (the reference's "(c p1 p2)" order, models/swinv2.py:242),
// bf16 [M, Kp] with the columns C*p1*p2 .. Kp-1 zero.
__global__ void __launch_bounds__(256) cot_patchify_kernel(const float* __restrict__ cot, uint16_t* __restrict__ dF, int B, int C,
                                                           int H, int W, int p1, int p2, int Kp) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * C * H * W;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W);
  const int y = static_cast<int>((idx / W) % H);
  const int c = static_cast<int>((idx / (static_cast<long long>(W) * H)) % C);
  const int b = static_cast<int>(idx / (static_cast<long long>(W) * H * C));
  const int gw = W / p2, gh = H / p1;
  const int gy = y / p1, i = y - gy * p1, gx = x / p2, j = x - gx * p2;
  const long long row = (static_cast<long long>(b) * gh + gy) * gw + gx;
  dF[row * Kp + (c * p1 + i) * p2 + j] = f2bf(__ldg(cot + idx));
}

int launch_cot_patchify(const float* cot, void* dF, int B, int C, int H, int W, int p1, int p2, int Kp, cudaStream_t stream) {
  SWB_REQUIRE(Kp >= C * p1 * p2 && H % p1 == 0 && W % p2 == 0, "cot_patchify: bad geometry");
  const long long rows = static_cast<long long>(B) * (H / p1) * (W / p2);
  if (Kp > C * p1 * p2) SWB_CHECK_CUDA(cudaMemsetAsync(dF, 0, static_cast<size_t>(rows) * Kp * 2, stream));
  const long long total = static_cast<long long>(B) * C * H * W;
  cot_patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(cot, static_cast<uint16_t*>(dF), B, C, H, W,
                                                                                      p1, p2, Kp);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// dpos[t, c] (+)= sum_b dx[b * tokens + t, c]
__global__ void __launch_bounds__(256) sum_over_samples_kernel(const float* __restrict__ dx, int B, long long per,
                                                               float* __restrict__ out, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per) return;
  float a = accumulate ? out[i] : 0.f;
  for (int b = 0; b < B; ++b) a += __ldg(dx + b * per + i);
  out[i] = a;
}

int launch_sum_over_samples(const float* dx, int B, long long per, float* out, int accumulate, cudaStream_t stream) {
  sum_over_samples_kernel<<<static_cast<unsigned>((per + 255) / 256), 256, 0, stream>>>(dx, B, per, out, accumulate);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// =========================================================================================================
// Conditioning backward (models/swinv2.py:44-74, :83-86, :316-321).  Forward, per sample b:
//   emb = temb(t) + aux_embed(aux sqrt(a));  z1 = l1 emb;  h1 = silu(z1);  z2 = l2 h1;  c = silu(z2);
//   [scale | shift]_l = mod_l c;  gain_l = gamma_l (1 + scale_l);  bias_l = beta_l (1 + scale_l) + shift_l.
// All of it is O(B * D^2): tiny next to the token work, so the kernels are simple (one warp per output row, fixed order).

// y[b, n] = bias[n] + W[n, :] . x[b, :]      (pre-activations are recomputed in the backward instead of being saved)
__global__ void __launch_bounds__(256) gemv_plain_kernel(const float* __restrict__ Wm, const float* __restrict__ bias,
                                                         const float* __restrict__ x, float* __restrict__ y, int N, int K, int B) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  for (int b = 0; b < B; ++b) {
    float a = 0.f;
    for (int k = lane; k < K; k += 32) a = fmaf(__ldg(Wm + static_cast<size_t>(n) * K + k), __ldg(x + static_cast<size_t>(b) * K + k), a);
    a = wsum(a);
    if (lane == 0) y[static_cast<size_t>(b) * N + n] = a + (bias ? bias[n] : 0.f);
  }
}

// dx[b, k] = sum_n W[n, k] dy[b, n]: partial over a chunk of 256 rows n per blockIdx.y -> part[chunk][b][k]
__global__ void __launch_bounds__(256) gemv_t_partial_kernel(const float* __restrict__ Wm, const float* __restrict__ dy,
                                                             float* __restrict__ part, int N, int K, int B) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int n0 = blockIdx.y * 256;
  const int n1 = min(N, n0 + 256);
  for (int b = 0; b < B; ++b) {
    float a = 0.f;
    if (k < K)
      for (int n = n0; n < n1; ++n) a = fmaf(__ldg(Wm + static_cast<size_t>(n) * K + k), __ldg(dy + static_cast<size_t>(b) * N + n), a);
    if (k < K) part[(static_cast<size_t>(blockIdx.y) * B + b) * K + k] = a;
  }
}

// dW[n, k] (+)= sum_b dy[b, n] x[b, k];  db[n] (+)= sum_b dy[b, n]
__global__ void __launch_bounds__(256) outer_accum_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                          float* __restrict__ dW, float* __restrict__ db, int N, int K, int B,
                                                          int accumulate) {
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += 256) {
    float a = accumulate ? dW[static_cast<size_t>(n) * K + k] : 0.f;
    for (int b = 0; b < B; ++b) a = fmaf(__ldg(dy + static_cast<size_t>(b) * N + n), __ldg(x + static_cast<size_t>(b) * K + k), a);
    dW[static_cast<size_t>(n) * K + k] = a;
  }
  if (db && threadIdx.x == 0) {
    float a = accumulate ? db[n] : 0.f;
    for (int b = 0; b < B; ++b) a += dy[static_cast<size_t>(b) * N + n];
    db[n] = a;
  }
}

// dz[b, i] = dy[b, i] * silu'(z[b, i])
__global__ void __launch_bounds__(256) silu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                       float* __restrict__ dz, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float zz = z[i], s = 1.0f / (1.0f + expf(-zz));
  dz[i] = dy[i] * s * (1.0f + zz * (1.0f - s));
}

// dmod[b, l, 0:D] = dgain[l,b] gamma[l] + dbias[l,b] beta[l];  dmod[b, l, D:2D] = dbias[l,b];
// dgamma[l, i] (+)= sum_b dgain (1 + scale);  dbeta[l, i] (+)= sum_b dbias (1 + scale)        mod: [B, L*2D] saved by the forward
__global__ void __launch_bounds__(256) mod_finalize_bwd_kernel(const float* __restrict__ dgain, const float* __restrict__ dbias,
                                                               const float* __restrict__ mod, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ dmod,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta, int L, int B,
                                                               int D, int accumulate) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= L * D) return;
  const int l = idx / D, i = idx - l * D;
  float ag = accumulate ? dgamma[idx] : 0.f, ab = accumulate ? dbeta[idx] : 0.f;
  const float g = gamma[idx], be = beta[idx];
  for (int b = 0; b < B; ++b) {
    const float dg = dgain[(static_cast<size_t>(l) * B + b) * D + i], dbv = dbias[(static_cast<size_t>(l) * B + b) * D + i];
    const float* m = mod + (static_cast<size_t>(b) * L + l) * 2 * D;
    const float sc = 1.0f + m[i];
    ag = fmaf(dg, sc, ag);
    ab = fmaf(dbv, sc, ab);
    float* dm = dmod + (static_cast<size_t>(b) * L + l) * 2 * D;
    dm[i] = fmaf(dg, g, dbv * be);
    dm[D + i] = dbv;
  }
  dgamma[idx] = ag;
  dbeta[idx] = ab;
}

// daux_w[i, j] (+)= sum_b demb[b, i] aux[b, j] sqrt(a);  daux_b[i] (+)= sum_b demb[b, i]
__global__ void __launch_bounds__(256) aux_embed_bwd_kernel(const float* __restrict__ demb, const float* __restrict__ aux,
                                                            float* __restrict__ daux_w, float* __restrict__ daux_b, int D,
                                                            int aux_dim, int B, int accumulate) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= D) return;
  const float s = sqrtf(static_cast<float>(aux_dim));
  float ab = accumulate ? daux_b[i] : 0.f;
  for (int b = 0; b < B; ++b) ab += demb[static_cast<size_t>(b) * D + i];
  daux_b[i] = ab;
  for (int j = 0; j < aux_dim; ++j) {
    float a = accumulate ? daux_w[i * aux_dim + j] : 0.f;
    for (int b = 0; b < B; ++b) a = fmaf(demb[static_cast<size_t>(b) * D + i], aux[b * aux_dim + j] * s, a);
    daux_w[i * aux_dim + j] = a;
  }
}

size_t conditioning_bwd_scratch_floats(int B, int D, int L) {
  const size_t chunks = (static_cast<size_t>(L) * 2 * D + 255) / 256;
  // dmod [B, L*2D] | z [B,D] | dvec x3 [B,D] | gemv_t partials [chunks][B][D]
  return static_cast<size_t>(B) * L * 2 * D + 4 * static_cast<size_t>(B) * D + chunks * B * D;
}

// fwd_scratch: what launch_conditioning left behind: emb [B,D] | h1 [B,D] | c [B,D] | mod [B, L*2D]
// dc[b, k] += dlv[b] * lv_w[k]: the logvar head's contribution to the gradient of the conditioning vector
__global__ void __launch_bounds__(256) logvar_dcond_kernel(float* __restrict__ dc, const float* __restrict__ dlv,
                                                           const float* __restrict__ lv_w, int B, int D) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= B * D) return;
  dc[i] = fmaf(__ldg(dlv + i / D), __ldg(lv_w + i % D), dc[i]);
}

int launch_logvar_head(const float* fwd_scratch, const float* lv_w, const float* lv_b, int B, int D, float* logvar,
                       cudaStream_t stream) {
  const float* c = fwd_scratch + static_cast<size_t>(2) * B * D;          // [emb | h1 | c | mod ...]
  gemv_plain_kernel<<<1, 256, 0, stream>>>(lv_w, lv_b, c, logvar, 1, D, B);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

int launch_conditioning_bwd(const CondWeights& w, const CondGrads& g, const float* aux, const float* fwd_scratch,
                            const float* dgain, const float* dbias, int B, int D, int L, float* scratch, int accumulate,
                            cudaStream_t stream, const float* lv_w, const float* dlogvar, float* g_lv_w, float* g_lv_b) {
  const float* emb = fwd_scratch;
  const float* h1 = emb + static_cast<size_t>(B) * D;
  const float* c = h1 + static_cast<size_t>(B) * D;
  const float* mod = c + static_cast<size_t>(B) * D;
  const int NM = L * 2 * D;
  float* dmod = scratch;
  float* z = dmod + static_cast<size_t>(B) * NM;
  float* v0 = z + static_cast<size_t>(B) * D;
  float* v1 = v0 + static_cast<size_t>(B) * D;
  float* v2 = v1 + static_cast<size_t>(B) * D;
  float* part = v2 + static_cast<size_t>(B) * D;
  mod_finalize_bwd_kernel<<<(L * D + 255) / 256, 256, 0, stream>>>(dgain, dbias, mod, w.ln_gamma, w.ln_beta, dmod, g.ln_gamma,
                                                                   g.ln_beta, L, B, D, accumulate);
  outer_accum_kernel<<<NM, 256, 0, stream>>>(dmod, c, g.mod_w, g.mod_b, NM, D, B, accumulate);
  auto gemv_t = [&](const float* Wm, const float* dy, float* out, int N, int K) -> int {
    const int chunks = (N + 255) / 256;
    gemv_t_partial_kernel<<<dim3((K + 255) / 256, chunks), 256, 0, stream>>>(Wm, dy, part, N, K, B);
    return launch_reduce_partials(part, chunks, B * K, out, 1, 0, stream);
  };
  int rc = gemv_t(w.mod_w, dmod, v0, NM, D);                                       // v0 = dc
  if (rc) return rc;
  if (lv_w != nullptr && dlogvar != nullptr) {
    // logvar = logvar_embed(c) (models/swinv2.py:326-327): its weight / bias gradients and its share of dc
    if (g_lv_w != nullptr) outer_accum_kernel<<<1, 256, 0, stream>>>(dlogvar, c, g_lv_w, g_lv_b, 1, D, B, accumulate);
    logvar_dcond_kernel<<<(B * D + 255) / 256, 256, 0, stream>>>(v0, dlogvar, lv_w, B, D);
  }
  gemv_plain_kernel<<<(D + 7) / 8, 256, 0, stream>>>(w.l2_w, w.l2_b, h1, z, D, D, B);      // z2
  silu_bwd_kernel<<<(B * D + 255) / 256, 256, 0, stream>>>(v0, z, v1, B * D);               // v1 = dz2
  outer_accum_kernel<<<D, 256, 0, stream>>>(v1, h1, g.l2_w, g.l2_b, D, D, B, accumulate);
  rc = gemv_t(w.l2_w, v1, v0, D, D);                                               // v0 = dh1
  if (rc) return rc;
  gemv_plain_kernel<<<(D + 7) / 8, 256, 0, stream>>>(w.l1_w, w.l1_b, emb, z, D, D, B);     // z1
  silu_bwd_kernel<<<(B * D + 255) / 256, 256, 0, stream>>>(v0, z, v1, B * D);               // v1 = dz1
  outer_accum_kernel<<<D, 256, 0, stream>>>(v1, emb, g.l1_w, g.l1_b, D, D, B, accumulate);
  if (w.aux_w != nullptr && aux != nullptr && w.aux_dim > 0 && g.aux_w != nullptr) {
    rc = gemv_t(w.l1_w, v1, v2, D, D);                                             // v2 = demb
    if (rc) return rc;
    aux_embed_bwd_kernel<<<(D + 255) / 256, 256, 0, stream>>>(v2, aux, g.aux_w, g.aux_b, D, w.aux_dim, B, accumulate);
  }
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
