// HBM-bound glue kernels of the Swift denoiser: patch gather (fused concat + patchify + bf16 cast),
// LayerNorm + TrigFlow-time modulation + residual add, and the conditioning micro-kernels
// (sinusoidal embedding, fp32 GEMVs, modulation folding).
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace swb {

// =========================================================================================================
// Patch gather: A[(b, gy, gx), k] with k = c*(p1*p2) + py*p2 + px  <-  cat([src0*scale0, src1], 1)[b, c, gy*p1+py, gx*p2+px]
//
// Replaces precond.py:139-141 (torch.cat) + swinv2.py:224-229 (einops rearrange) + the fp32->bf16 cast of the
// GEMM operand.  The reference feature order is "(p1 p2 c)"; we use "(c p1 p2)" and permute the columns of the
// patch-embed weight once at pack time, which makes the gather write 2*p1*p2-byte runs per channel.
// SPLIT writes a second operand half lo = bf16(x - float(bf16(x))) at column offset Kp so that
// [hi | lo] * [W | W]^T reproduces the fp32 input to ~2^-17 relative instead of 2^-9.
constexpr int kGatherCh = 16;   // channels staged per pass: 16 * p1*p2 output elements (128 B for 2x2 patches) per token

template <bool SPLIT, bool F16>
__global__ void __launch_bounds__(256) patch_gather_kernel(const float* __restrict__ src0, int C0, float scale0,
                                                           const float* __restrict__ src1, int C1,
                                                           uint16_t* __restrict__ A, int lda, int Kp, int H,
                                                           int W, int p1, int p2) {
  extern __shared__ float tile[];   // [kGatherCh * p1][W + 2]
  const int gy = blockIdx.x, b = blockIdx.y;
  const int gh = H / p1, gw = W / p2;
  const int pp = p1 * p2;
  const int C = C0 + C1;
  const int pitch = W + 2;
  const int cvirt = (Kp + pp - 1) / pp;            // channels >= C are virtual zero padding up to Kp
  const int nchunks = (cvirt + kGatherCh - 1) / kGatherCh;
  const size_t row_base = (static_cast<size_t>(b) * gh + gy) * gw;
  const int kchunk = kGatherCh * pp;               // output elements per token per pass (multiple of 8)
  const int groups = kchunk / 8;                   // 16-byte groups per token per pass
  (void)nchunks;
  {
    const int chunk = blockIdx.z;                    // one channel chunk per block: gh * B * nchunks independent blocks
    const int c0 = chunk * kGatherCh;
    // image rows of this channel chunk -> smem, 16 bytes per load (W % 4 == 0 and 16-byte aligned rows: checked by the launcher)
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < kGatherCh * p1 * w4; idx += blockDim.x) {
      const int x4 = idx % w4;
      const int r = idx / w4;           // c*p1 + py
      const int c = c0 + r / p1, py = r % p1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C0) {
        v = __ldg(reinterpret_cast<const float4*>(src0 + ((static_cast<size_t>(b) * C0 + c) * H + gy * p1 + py) * W) + x4);
        v.x *= scale0; v.y *= scale0; v.z *= scale0; v.w *= scale0;
      } else if (c < C) {
        v = __ldg(reinterpret_cast<const float4*>(src1 + ((static_cast<size_t>(b) * C1 + (c - C0)) * H + gy * p1 + py) * W) + x4);
      }
      float* t = tile + r * pitch + 4 * x4;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
    __syncthreads();
    // one 16-byte store (8 consecutive k) per thread: consecutive threads cover consecutive groups of one token
    for (int idx = threadIdx.x; idx < gw * groups; idx += blockDim.x) {
      const int q = idx % groups, gx = idx / groups;
      const int k0 = c0 * pp + q * 8;
      if (k0 >= Kp) continue;                       // Kp is a multiple of 8
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kl = q * 8 + j;
        const int c = kl / pp, r = kl - c * pp;
        const int py = r / p2, px = r - py * p2;
        v[j] = tile[(c * p1 + py) * pitch + gx * p2 + px];
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = pack_act2<F16>(v[2 * j], v[2 * j + 1]);
        if (SPLIT)
          lo[j] = pack_act2<F16>(v[2 * j] - unpack_act1<F16>(static_cast<uint16_t>(hi[j] & 0xffffu)),
                                 v[2 * j + 1] - unpack_act1<F16>(static_cast<uint16_t>(hi[j] >> 16)));
      }
      uint16_t* dst = A + (row_base + gx) * lda + k0;
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (SPLIT) *reinterpret_cast<uint4*>(dst + Kp) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

int launch_patch_gather(const float* src0, int C0, float scale0, const float* src1, int C1, void* A, int lda, int Kp,
                        int split, int act_f16, int B, int H, int W, int p1, int p2, cudaStream_t stream) {
  SWB_REQUIRE(H % p1 == 0 && W % p2 == 0, "patch_gather: image %dx%d not divisible by patch %dx%d", H, W, p1, p2);
  SWB_REQUIRE(Kp >= (C0 + C1) * p1 * p2 && Kp % 8 == 0 && lda % 8 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0,
              "patch_gather: need Kp=%d >= C*p1*p2=%d, Kp and lda multiples of 8, A 16-byte aligned", Kp,
              (C0 + C1) * p1 * p2);
  SWB_REQUIRE((kGatherCh * p1 * p2) % 8 == 0, "patch_gather: patch %dx%d unsupported", p1, p2);
  SWB_REQUIRE(W % 4 == 0 && (reinterpret_cast<uintptr_t>(src0) & 15) == 0 && (reinterpret_cast<uintptr_t>(src1) & 15) == 0,
              "patch_gather: image width %d must be a multiple of 4 and the inputs 16-byte aligned", W);
  const size_t smem = static_cast<size_t>(kGatherCh) * p1 * (W + 2) * sizeof(float);
  SWB_REQUIRE(smem <= 48 * 1024, "patch_gather: image width %d too large for the staging tile", W);
  const int cvirt = (Kp + p1 * p2 - 1) / (p1 * p2);
  dim3 grid(H / p1, B, (cvirt + kGatherCh - 1) / kGatherCh);
  uint16_t* A_ = static_cast<uint16_t*>(A);
#define SWB_GATHER(SP, F) \
  patch_gather_kernel<SP, F><<<grid, 256, smem, stream>>>(src0, C0, scale0, src1, C1, A_, lda, Kp, H, W, p1, p2)
  if (split) {
    if (act_f16) SWB_GATHER(true, true); else SWB_GATHER(true, false);
  } else {
    if (act_f16) SWB_GATHER(false, true); else SWB_GATHER(false, false);
  }
#undef SWB_GATHER
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// =========================================================================================================
// x <- x + LN(branch) * gain[b] + bias[b]
//
// reference: ModulatedNorm (swinv2.py:77-86) applied to the branch output, then the residual add
// (swinv2.py:211-212).  gain = gamma*(1+scale(t)), bias = beta*(1+scale(t)) + shift(t) are folded per sample
// by mod_finalize_kernel.  The residual stream x lives in HBM as a 16-bit pair xhl[M, 2D] = [hi | lo] with
// x = hi + lo (about 22 significant bits in fp16 mode); hi is at the same time the A operand of the next GEMM, so no
// separate 16-bit copy is written and the row costs 2 (branch) + 4 (read) + 4 (write) bytes per element.
// One warp per token row; the row lives in registers (two-pass mean/variance in fp32); every global load is issued
// before the first use; accesses are 16 bytes per lane, lane-strided (512 contiguous bytes per warp request).
// BR16: the branch is stored in the 16-bit operand format (fp16 mode) instead of fp32.
#ifndef SWB_LN_MIN_BLOCKS
#define SWB_LN_MIN_BLOCKS 4
#endif
// PAIR = false: x is the hi half alone (row pitch still 2D), one rounding per update; lo is neither read nor written.
template <int NV8, bool F16, bool BR16, bool PAIR>
__global__ void __launch_bounds__(128, SWB_LN_MIN_BLOCKS) ln_mod_residual_kernel(const void* __restrict__ branch_, uint16_t* __restrict__ xhl,
                                                              const float* __restrict__ gain,
                                                              const float* __restrict__ bias, int M, int D,
                                                              int tokens, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int ng = D >> 3;                                   // groups of 8 elements (16 bytes of 16-bit data) per row
  uint4* xh = reinterpret_cast<uint4*>(xhl + static_cast<size_t>(row) * 2 * D);
  uint4* xl = xh + ng;
  float v[NV8][8];
  uint4 rh[NV8], rl[NV8];
  auto unpack8 = [](uint4 r, float* o) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_act2<F16>(w[j]);
      o[2 * j] = f.x;
      o[2 * j + 1] = f.y;
    }
  };
  // all global loads of the row (branch, hi, lo) are issued before the first use
  if constexpr (BR16) {
    const uint4* br = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(branch_) + static_cast<size_t>(row) * D);
    uint4 raw[NV8];
#pragma unroll
    for (int i = 0; i < NV8; ++i) {
      const int c = i * 32 + lane;
      raw[i] = (c < ng) ? __ldg(br + c) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < NV8; ++i) {
      const int c = i * 32 + lane;
      rh[i] = (c < ng) ? xh[c] : make_uint4(0u, 0u, 0u, 0u);
      rl[i] = (PAIR && c < ng) ? xl[c] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < NV8; ++i) unpack8(raw[i], v[i]);
  } else {
    const float4* br = reinterpret_cast<const float4*>(static_cast<const float*>(branch_) + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV8; ++i) {
      const int c = i * 32 + lane;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (c < ng) {
        a = __ldg(br + 2 * c);
        b = __ldg(br + 2 * c + 1);
      }
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
      v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
    }
#pragma unroll
    for (int i = 0; i < NV8; ++i) {
      const int c = i * 32 + lane;
      rh[i] = (c < ng) ? xh[c] : make_uint4(0u, 0u, 0u, 0u);
      rl[i] = (PAIR && c < ng) ? xl[c] : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += v[i][j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(D);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    if (i * 32 + lane < ng) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        var = fmaf(d, d, var);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / static_cast<float>(D) + eps);
  const int b = row / tokens;
  const float4* g4 = reinterpret_cast<const float4*>(gain + static_cast<size_t>(b) * D);
  const float4* b4 = reinterpret_cast<const float4*>(bias + static_cast<size_t>(b) * D);
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    const int c = i * 32 + lane;
    if (c < ng) {
      const float4 ga = __ldg(g4 + 2 * c), gb = __ldg(g4 + 2 * c + 1), ba = __ldg(b4 + 2 * c), bb = __ldg(b4 + 2 * c + 1);
      const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float be[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
      float xhv[8], xlv[8], o[8];
      unpack8(rh[i], xhv);
      unpack8(rl[i], xlv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (xhv[j] + xlv[j]) + fmaf((v[i][j] - mean) * rstd, g[j], be[j]);
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = pack_act2<F16>(o[2 * j], o[2 * j + 1]);
        if constexpr (PAIR) {
          const float2 back = unpack_act2<F16>(h[j]);
          l[j] = pack_act2<F16>(o[2 * j] - back.x, o[2 * j + 1] - back.y);
        }
      }
      xh[c] = make_uint4(h[0], h[1], h[2], h[3]);
      if constexpr (PAIR) xl[c] = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

int launch_ln_mod_residual(const void* branch, int branch_16bit, void* xhl, const float* gain, const float* bias, int M,
                           int D, int tokens, float eps, int act_f16, cudaStream_t stream) {
  SWB_REQUIRE(D % 8 == 0, "ln_mod_residual: dim %d must be a multiple of 8", D);
  SWB_REQUIRE(((reinterpret_cast<uintptr_t>(branch) | reinterpret_cast<uintptr_t>(gain) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(xhl) & 15) == 0,
              "ln_mod_residual: pointers must be 16-byte aligned");
  const bool x_single = (act_f16 & 2) != 0;                 // format word: bit 0 = fp16 operands, bit 1 = single-value residual stream
  act_f16 &= 1;
  SWB_REQUIRE(!x_single || act_f16, "ln_mod_residual: the single-value residual stream needs fp16 operands");
  const int rows_per_block = 4;
  dim3 grid((M + rows_per_block - 1) / rows_per_block);
  auto x_ = static_cast<uint16_t*>(xhl);
#define SWB_LN3(V, F, R, P) ln_mod_residual_kernel<V, F, R, P><<<grid, 128, 0, stream>>>(branch, x_, gain, bias, M, D, tokens, eps)
#define SWB_LN(V)                                                   \
  do {                                                              \
    if (x_single) {                                                 \
      if (branch_16bit) SWB_LN3(V, true, true, false); else SWB_LN3(V, true, false, false);   \
    } else if (act_f16) {                                           \
      if (branch_16bit) SWB_LN3(V, true, true, true); else SWB_LN3(V, true, false, true);   \
    } else {                                                        \
      if (branch_16bit) SWB_LN3(V, false, true, true); else SWB_LN3(V, false, false, true); \
    }                                                               \
  } while (0)
  const int nv8 = (D / 8 + 31) / 32;
  if (nv8 <= 2) SWB_LN(2);
  else if (nv8 <= 3) SWB_LN(3);
  else if (nv8 <= 5) SWB_LN(5);
  else if (nv8 <= 8) SWB_LN(8);
  else {
    set_error("ln_mod_residual: dim %d > 2048 unsupported", D);
    return SWB_ERR_INVALID;
  }
#undef SWB_LN
#undef SWB_LN3
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// =========================================================================================================
// Conditioning (swinv2.py:44-60, :316-321): emb = [sin(t f) | cos(t f)] + aux_embed(aux * sqrt(aux_dim))
__global__ void cond_embed_kernel(const float* __restrict__ t, const float* __restrict__ aux,
                                  const float* __restrict__ aux_w, const float* __restrict__ aux_b, int aux_dim,
                                  float timestep_weight, int D, float* __restrict__ emb) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  const int half = D / 2;
  float e = 0.f;
  if (i < 2 * half) {
    const int j = (i < half) ? i : i - half;
    const float freq = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half));
    const float arg = (t[b] * timestep_weight) * freq;
    e = (i < half) ? sinf(arg) : cosf(arg);
  }
  if (aux != nullptr && aux_dim > 0) {
    const float s = sqrtf(static_cast<float>(aux_dim));
    float a = aux_b[i];
    for (int j = 0; j < aux_dim; ++j) a = fmaf(aux_w[i * aux_dim + j], aux[b * aux_dim + j] * s, a);
    e += a;
  }
  emb[static_cast<size_t>(b) * D + i] = e;
}

// out[b, n] = act(bias[n] + W[n, :] . in[b, :]) in fp32; one warp per output row n, batch tiled by 8.
template <int ACT>
__global__ void __launch_bounds__(256) gemv_rows_kernel(const float* __restrict__ Wm, const float* __restrict__ bias,
                                                        const float* __restrict__ in, float* __restrict__ out, int N,
                                                        int K, int B) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const float* w = Wm + static_cast<size_t>(n) * K;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(w + k);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (b0 + j < B) acc[j] = fmaf(wv, __ldg(in + static_cast<size_t>(b0 + j) * K + k), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (b0 + j < B) {
          float v = acc[j] + (bias ? bias[n] : 0.f);
          if (ACT == 1) v = v / (1.0f + expf(-v));
          out[static_cast<size_t>(b0 + j) * N + n] = v;
        }
      }
    }
  }
}

// gain[l, b, i] = gamma[l, i] * (1 + scale),  bias[l, b, i] = beta[l, i] * (1 + scale) + shift,
// with [scale | shift] = mod[b, l*2D : (l+1)*2D]   (ModulatedNorm, swinv2.py:83-86)
__global__ void mod_finalize_kernel(const float* __restrict__ mod, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float* __restrict__ gain,
                                    float* __restrict__ bias, int L, int B, int D) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(L) * B * D;
  if (idx >= total) return;
  const int i = idx % D;
  const int b = (idx / D) % B;
  const int l = idx / (static_cast<size_t>(D) * B);
  const float* m = mod + (static_cast<size_t>(b) * L + l) * 2 * D;
  const float sc = 1.0f + m[i];
  gain[idx] = gamma[l * D + i] * sc;
  bias[idx] = fmaf(beta[l * D + i], sc, m[D + i]);
}

int launch_conditioning(const CondWeights& w, const float* t, const float* aux, int B, int D, int L,
                        float timestep_weight, float* scratch, float* gain, float* bias, float* cond_out,
                        cudaStream_t stream) {
  // scratch: emb [B,D] | h1 [B,D] | c [B,D] | mod [B, L*2D]
  float* emb = scratch;
  float* h1 = emb + static_cast<size_t>(B) * D;
  float* c = h1 + static_cast<size_t>(B) * D;
  float* mod = c + static_cast<size_t>(B) * D;
  cond_embed_kernel<<<dim3((D + 127) / 128, B), 128, 0, stream>>>(t, aux, w.aux_w, w.aux_b, w.aux_dim,
                                                                  timestep_weight, D, emb);
  const int wpb = 8;
  gemv_rows_kernel<1><<<(D + wpb - 1) / wpb, 256, 0, stream>>>(w.l1_w, w.l1_b, emb, h1, D, D, B);
  gemv_rows_kernel<1><<<(D + wpb - 1) / wpb, 256, 0, stream>>>(w.l2_w, w.l2_b, h1, c, D, D, B);
  const int NM = L * 2 * D;
  gemv_rows_kernel<0><<<(NM + wpb - 1) / wpb, 256, 0, stream>>>(w.mod_w, w.mod_b, c, mod, NM, D, B);
  const size_t total = static_cast<size_t>(L) * B * D;
  mod_finalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(mod, w.ln_gamma, w.ln_beta,
                                                                                      gain, bias, L, B, D);
  if (cond_out)
    SWB_CHECK_CUDA(cudaMemcpyAsync(cond_out, c, static_cast<size_t>(B) * D * sizeof(float),
                                   cudaMemcpyDeviceToDevice, stream));
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fp16 range diagnostics: every fp16 conversion of an activation saturates at +-65504 (F2FP.SATFINITE, ptx.cuh) instead
// of producing inf.  This scan counts the elements of a 16-bit [rows, cols] tensor (row pitch `pitch` elements) that sit
// exactly at the saturation value -- |x| = 65504 = 0x7BFF -- so a checkpoint whose activations leave the fp16 range is
// reported instead of silently clipped (swb200_debug_saturation; off by default, costs nothing when off).
__global__ void __launch_bounds__(256) count_saturated_f16_kernel(const uint16_t* __restrict__ p, long long rows, int cols8,
                                                                  long long pitch, unsigned long long* __restrict__ counter) {
  unsigned n = 0;
  const long long total = rows * cols8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols8;
    const int c = static_cast<int>(i - r * cols8);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + r * pitch) + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      n += ((w[k] & 0x7fffu) == 0x7bffu) ? 1u : 0u;
      n += (((w[k] >> 16) & 0x7fffu) == 0x7bffu) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, static_cast<unsigned long long>(n));
}

int launch_count_saturated_f16(const void* buf, long long rows, int cols, long long pitch, unsigned long long* counter,
                               cudaStream_t stream) {
  SWB_REQUIRE(cols % 8 == 0 && pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(buf) & 15) == 0,
              "count_saturated: cols / pitch must be multiples of 8 elements and the buffer 16-byte aligned");
  if (rows <= 0 || cols <= 0) return SWB_OK;
  const long long total = rows * (cols / 8);
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * num_sms()));
  count_saturated_f16_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(buf), rows, cols / 8, pitch, counter);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
