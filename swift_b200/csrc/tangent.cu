// Forward-mode tangent (JVP) companions of the non-linear stages of the denoiser (first part).
//
// The sCM training loss (reference training/loss.py:186-260) needs  (F, dF) = jvp(net, (x, t), (v_x, v_t))  with the
// result detached: a pure forward pass that carries a tangent next to every activation.  Every Linear is the SAME tcgen05
// GEMM kernel run over a row-stacked operand [x ; dx] (rows 0..M-1 primal, M..2M-1 tangent: d(W x) = W dx), with the
// plain fp32 epilogue; the kernels here apply the derivative rules of what sits between the GEMMs:
//   ln_dual_kernel         ModulatedNorm + residual (models/swinv2.py:77-86, :211-212) with tangent gain / bias
//   qkv_dual_pack_kernel   F.normalize(q, k) * logit scale (:123-127) and its Jacobian, packed for the attention stage
//   attn_scores_dual / attn_softmax_dual / attn_out_dual   windowed softmax attention (:129-135, :189-209):
//                          S = q k^T, dS = dq k^T + q dk^T,  P = softmax S,  dP = P (dS - sum_j P dS),
//                          O = P v, dO = dP v + P dv   (explicit-softmax path the reference selects with jvp=True)
//   swiglu_dual_kernel     silu(gate) * up (:99-100)
//   cond_* dual kernels    timestep embedding + latent MLP + modulation Linears (:44-60, :67-74, :84) w.r.t. t
// Straightforward fp32 CUDA-core kernels: this path is about coverage and parity first (it runs once per training step
// next to a backward pass that is not part of this library yet); the GEMMs, 97 % of its FLOPs, are on the tensor cores.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

#include <algorithm>

namespace swb {

namespace {

template <bool F16>
__device__ __forceinline__ float pair_load(const uint16_t* row, int D, int i) {
  return unpack_act1<F16>(row[i]) + unpack_act1<F16>(row[D + i]);
}
template <bool F16>
__device__ __forceinline__ void pair_store(uint16_t* row, int D, int i, float x) {
  const uint16_t h = pack_act1<F16>(x);
  row[i] = h;
  row[D + i] = pack_act1<F16>(x - unpack_act1<F16>(h));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// x <- x + LN(b) g + beta ;  dx <- dx + d[LN(b) g + beta]      (one warp per token row)
template <bool F16>
__global__ void __launch_bounds__(128) ln_dual_kernel(const float* __restrict__ branch, uint16_t* __restrict__ xhl,
                                                      const float* __restrict__ gain, const float* __restrict__ bias,
                                                      const float* __restrict__ dgain, const float* __restrict__ dbias,
                                                      int M, int D, int tokens, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* b = branch + static_cast<size_t>(row) * D;
  const float* bd = branch + static_cast<size_t>(M + row) * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += b[i];
  const float mean = warp_sum(s) / D;
  float v = 0.f;
  for (int i = lane; i < D; i += 32) v = fmaf(b[i] - mean, b[i] - mean, v);
  const float rstd = rsqrtf(warp_sum(v) / D + eps);
  float m1 = 0.f, m2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float nh = (b[i] - mean) * rstd;
    m1 += bd[i];
    m2 = fmaf(nh, bd[i], m2);
  }
  m1 = warp_sum(m1) / D;
  m2 = warp_sum(m2) / D;
  const int smp = row / tokens;
  const float *g = gain + static_cast<size_t>(smp) * D, *be = bias + static_cast<size_t>(smp) * D;
  const float *dg = dgain + static_cast<size_t>(smp) * D, *dbe = dbias + static_cast<size_t>(smp) * D;
  uint16_t* xr = xhl + static_cast<size_t>(row) * 2 * D;
  uint16_t* xdr = xhl + static_cast<size_t>(M + row) * 2 * D;
  for (int i = lane; i < D; i += 32) {
    const float nh = (b[i] - mean) * rstd;
    const float dnh = rstd * (bd[i] - m1 - nh * m2);
    pair_store<F16>(xr, D, i, pair_load<F16>(xr, D, i) + fmaf(nh, g[i], be[i]));
    pair_store<F16>(xdr, D, i, pair_load<F16>(xdr, D, i) + fmaf(dnh, g[i], fmaf(nh, dg[i], dbe[i])));
  }
}

int launch_ln_dual(const float* branch2, void* xhl2, const float* gain, const float* bias, const float* dgain,
                   const float* dbias, int M, int D, int tokens, float eps, int act_f16, cudaStream_t stream) {
  dim3 grid((M + 3) / 4);
  if (act_f16) ln_dual_kernel<true><<<grid, 128, 0, stream>>>(branch2, static_cast<uint16_t*>(xhl2), gain, bias, dgain, dbias, M, D, tokens, eps);
  else ln_dual_kernel<false><<<grid, 128, 0, stream>>>(branch2, static_cast<uint16_t*>(xhl2), gain, bias, dgain, dbias, M, D, tokens, eps);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// raw [2M, 3D] fp32 (packed column order part*D + head*hd + d) -> qkv, dqkv : [3][H][M][pad] 16-bit, q / k normalised
template <bool F16>
__global__ void __launch_bounds__(128) qkv_dual_pack_kernel(const float* __restrict__ raw, const float* __restrict__ qscale,
                                                            uint16_t* __restrict__ out, uint16_t* __restrict__ dout, int M,
                                                            int D, int heads, int hd, int pad) {
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* r = raw + static_cast<size_t>(row) * 3 * D;
  const float* rd = raw + static_cast<size_t>(M + row) * 3 * D;
  for (int slot = warp; slot < 3 * heads; slot += blockDim.x >> 5) {
    const int part = slot / heads, head = slot - part * heads;
    const float* v = r + slot * hd;
    const float* dv = rd + slot * hd;
    float inv = 1.f, proj = 0.f, sc = 1.f;
    if (part < 2) {
      float ss = 0.f, dot = 0.f;
      for (int i = lane; i < hd; i += 32) {
        ss = fmaf(v[i], v[i], ss);
        dot = fmaf(v[i], dv[i], dot);
      }
      ss = warp_sum(ss);
      dot = warp_sum(dot);
      const float nrm = sqrtf(ss);
      inv = 1.0f / fmaxf(nrm, 1e-12f);
      proj = nrm > 1e-12f ? dot * inv * inv : 0.f;         // (u . dv) / |v| with u = v / |v|
      if (part == 0) sc = qscale[head];
    }
    uint16_t* o = out + (static_cast<size_t>(slot) * M + row) * pad;
    uint16_t* od = dout + (static_cast<size_t>(slot) * M + row) * pad;
    for (int i = lane; i < pad; i += 32) {
      float a = 0.f, da = 0.f;
      if (i < hd) {
        a = v[i] * inv * sc;
        da = (part < 2) ? (dv[i] - v[i] * proj) * inv * sc : dv[i];
      }
      o[i] = pack_act1<F16>(a);
      od[i] = pack_act1<F16>(da);
    }
  }
}

int launch_qkv_dual_pack(const float* raw2, const float* qscale, void* qkv, void* dqkv, int M, int D, int heads, int hd,
                         int pad, int act_f16, cudaStream_t stream) {
  if (act_f16) qkv_dual_pack_kernel<true><<<M, 128, 0, stream>>>(raw2, qscale, static_cast<uint16_t*>(qkv), static_cast<uint16_t*>(dqkv), M, D, heads, hd, pad);
  else qkv_dual_pack_kernel<false><<<M, 128, 0, stream>>>(raw2, qscale, static_cast<uint16_t*>(qkv), static_cast<uint16_t*>(dqkv), M, D, heads, hd, pad);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Windowed attention with tangents.  item = (sample b, window, head); 256 tokens per window (16 x 16).
struct AttnDualGeom {
  int B, gh, gw, heads, M, shift_h, shift_w, pad, hd;
};
// token row (within the chunk) of window-local token n of window `win` of sample b: roll(-shift) + window_partition as
// index arithmetic (models/swinv2.py:192-197)
__device__ __forceinline__ int window_token_row(const AttnDualGeom& g, int b, int win, int n) {
  const int nwx = g.gw / 16;
  const int wy = win / nwx, wx = win - wy * nwx;
  const int y = (wy * 16 + (n >> 4) + g.shift_h) % g.gh;
  const int x = (wx * 16 + (n & 15) + g.shift_w) % g.gw;
  return (b * g.gh + y) * g.gw + x;
}

// S[i, j] = q_i . k_j ;  dS[i, j] = dq_i . k_j + q_i . dk_j      one block per (item, 64 x 64 tile), 16 x 16 threads
template <bool F16>
__global__ void __launch_bounds__(256) attn_scores_dual_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dqkv,
                                                               float* __restrict__ S, float* __restrict__ dS, AttnDualGeom g) {
  __shared__ float sq[64][33], sdq[64][33], sk[64][33], sdk[64][33];
  const int item = blockIdx.y;
  const int head = item % g.heads;
  const int bw = item / g.heads;
  const int nwin = (g.gh / 16) * (g.gw / 16);
  const int win = bw % nwin, b = bw / nwin;
  const int ti = (blockIdx.x >> 2) * 64, tj = (blockIdx.x & 3) * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const size_t slot_q = static_cast<size_t>(head) * g.M, slot_k = static_cast<size_t>(g.heads + head) * g.M;
  float acc[4][4] = {}, dacc[4][4] = {};
  for (int k0 = 0; k0 < g.hd; k0 += 32) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += 256) {
      const int r = idx >> 5, c = idx & 31;
      const int d = k0 + c;
      const int rq = window_token_row(g, b, win, ti + r), rk = window_token_row(g, b, win, tj + r);
      const bool ok = d < g.hd;
      sq[r][c] = ok ? unpack_act1<F16>(qkv[(slot_q + rq) * g.pad + d]) : 0.f;
      sdq[r][c] = ok ? unpack_act1<F16>(dqkv[(slot_q + rq) * g.pad + d]) : 0.f;
      sk[r][c] = ok ? unpack_act1<F16>(qkv[(slot_k + rk) * g.pad + d]) : 0.f;
      sdk[r][c] = ok ? unpack_act1<F16>(dqkv[(slot_k + rk) * g.pad + d]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      float q[4], dq[4], k[4], dk[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        q[a] = sq[ty * 4 + a][c];
        dq[a] = sdq[ty * 4 + a][c];
        k[a] = sk[tx * 4 + a][c];
        dk[a] = sdk[tx * 4 + a][c];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[a][e] = fmaf(q[a], k[e], acc[a][e]);
          dacc[a][e] = fmaf(dq[a], k[e], fmaf(q[a], dk[e], dacc[a][e]));
        }
    }
    __syncthreads();
  }
  float* So = S + static_cast<size_t>(item) * 65536;
  float* dSo = dS + static_cast<size_t>(item) * 65536;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      So[(ti + ty * 4 + a) * 256 + tj + tx * 4 + e] = acc[a][e];
      dSo[(ti + ty * 4 + a) * 256 + tj + tx * 4 + e] = dacc[a][e];
    }
}

// rows of 256: P = softmax(S), dP = P (dS - sum_j P dS), in place     (one warp per row)
__global__ void __launch_bounds__(256) attn_softmax_dual_kernel(float* __restrict__ S, float* __restrict__ dS, size_t rows) {
  const size_t row = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* s = S + row * 256;
  float* ds = dS + row * 256;
  float v[8], dv[8], mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = s[i * 32 + lane];
    dv[i] = ds[i * 32 + lane];
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float z = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = __expf(v[i] - mx);
    z += v[i];
  }
  const float inv = 1.0f / warp_sum(z);
  float c = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] *= inv;
    c = fmaf(v[i], dv[i], c);
  }
  c = warp_sum(c);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s[i * 32 + lane] = v[i];
    ds[i * 32 + lane] = v[i] * (dv[i] - c);
  }
}

// O = P v ; dO = dP v + P dv   -> attn2 [2M, D] 16-bit at the tokens' own rows, column head*hd + d.
// one block per (item, 64-row tile); 16 x 16 threads, thread = 4 rows x 6 d-columns (covers 96 >= hd)
template <bool F16>
__global__ void __launch_bounds__(256) attn_out_dual_kernel(const float* __restrict__ P, const float* __restrict__ dP,
                                                            const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dqkv,
                                                            uint16_t* __restrict__ attn2, AttnDualGeom g) {
  __shared__ float sp[64][33], sdp[64][33], sv[32][97], sdv[32][97];
  const int item = blockIdx.y;
  const int head = item % g.heads;
  const int bw = item / g.heads;
  const int nwin = (g.gh / 16) * (g.gw / 16);
  const int win = bw % nwin, b = bw / nwin;
  const int ti = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const size_t slot_v = static_cast<size_t>(2 * g.heads + head) * g.M;
  const float* Pi = P + static_cast<size_t>(item) * 65536;
  const float* dPi = dP + static_cast<size_t>(item) * 65536;
  float acc[4][6] = {}, dacc[4][6] = {};
  for (int j0 = 0; j0 < 256; j0 += 32) {
    for (int idx = threadIdx.x; idx < 64 * 32; idx += 256) {
      const int r = idx >> 5, c = idx & 31;
      sp[r][c] = Pi[(ti + r) * 256 + j0 + c];
      sdp[r][c] = dPi[(ti + r) * 256 + j0 + c];
    }
    for (int idx = threadIdx.x; idx < 32 * 96; idx += 256) {
      const int r = idx / 96, d = idx - r * 96;
      const int rv = window_token_row(g, b, win, j0 + r);
      const bool ok = d < g.hd;
      sv[r][d] = ok ? unpack_act1<F16>(qkv[(slot_v + rv) * g.pad + d]) : 0.f;
      sdv[r][d] = ok ? unpack_act1<F16>(dqkv[(slot_v + rv) * g.pad + d]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < 32; ++c) {
      float p[4], dp[4], v[6], dv[6];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        p[a] = sp[ty * 4 + a][c];
        dp[a] = sdp[ty * 4 + a][c];
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        v[e] = sv[c][tx * 6 + e];
        dv[e] = sdv[c][tx * 6 + e];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int e = 0; e < 6; ++e) {
          acc[a][e] = fmaf(p[a], v[e], acc[a][e]);
          dacc[a][e] = fmaf(dp[a], v[e], fmaf(p[a], dv[e], dacc[a][e]));
        }
    }
    __syncthreads();
  }
  const int D = g.heads * g.hd;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = window_token_row(g, b, win, ti + ty * 4 + a);
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const int d = tx * 6 + e;
      if (d < g.hd) {
        attn2[static_cast<size_t>(row) * D + head * g.hd + d] = pack_act1<F16>(acc[a][e]);
        attn2[static_cast<size_t>(g.M + row) * D + head * g.hd + d] = pack_act1<F16>(dacc[a][e]);
      }
    }
  }
}

int launch_attention_dual(const void* qkv, const void* dqkv, float* S, float* dS, void* attn2, int B, int gh, int gw,
                          int heads, int hd, int pad, int shift_h, int shift_w, int act_f16, cudaStream_t stream) {
  SWB_REQUIRE(gh % 16 == 0 && gw % 16 == 0 && hd <= 96, "attention_dual: grid %dx%d / head_dim %d unsupported", gh, gw, hd);
  AttnDualGeom g;
  g.B = B; g.gh = gh; g.gw = gw; g.heads = heads; g.M = B * gh * gw;
  g.shift_h = shift_h; g.shift_w = shift_w; g.pad = pad; g.hd = hd;
  const int items = B * (gh / 16) * (gw / 16) * heads;
  const auto* q = static_cast<const uint16_t*>(qkv);
  const auto* dq = static_cast<const uint16_t*>(dqkv);
  if (act_f16) attn_scores_dual_kernel<true><<<dim3(16, items), 256, 0, stream>>>(q, dq, S, dS, g);
  else attn_scores_dual_kernel<false><<<dim3(16, items), 256, 0, stream>>>(q, dq, S, dS, g);
  const size_t rows = static_cast<size_t>(items) * 256;
  attn_softmax_dual_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(S, dS, rows);
  if (act_f16) attn_out_dual_kernel<true><<<dim3(4, items), 256, 0, stream>>>(S, dS, q, dq, static_cast<uint16_t*>(attn2), g);
  else attn_out_dual_kernel<false><<<dim3(4, items), 256, 0, stream>>>(S, dS, q, dq, static_cast<uint16_t*>(attn2), g);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// raw [2M, 2*Dff] fp32, columns in packed tile order (every `tile` columns: [tile/2 gate | tile/2 up]) -> h2 [2M, Dff] 16-bit
template <bool F16>
__global__ void __launch_bounds__(256) swiglu_dual_kernel(const float* __restrict__ raw, uint16_t* __restrict__ h2, int M,
                                                          int Dff, int tile) {
  const size_t total = static_cast<size_t>(M) * Dff;
  const int half = tile / 2;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = idx / Dff;
    const int j = static_cast<int>(idx - row * Dff);
    const int col = (j / half) * tile + (j % half);
    const float* r = raw + row * 2 * Dff;
    const float* rd = raw + (M + row) * 2 * Dff;
    const float gte = r[col], up = r[col + half], dg = rd[col], du = rd[col + half];
    const float sig = 1.0f / (1.0f + __expf(-gte));
    const float silu = gte * sig;
    const float dsilu = sig * (1.0f + gte * (1.0f - sig));
    h2[row * Dff + j] = pack_act1<F16>(silu * up);
    h2[(M + row) * Dff + j] = pack_act1<F16>(fmaf(dsilu * dg, up, silu * du));
  }
}

int launch_swiglu_dual(const float* raw2, void* h2, int M, int Dff, int tile, int act_f16, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(M) * Dff;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>((total + 255) / 256, 148 * 16));
  if (act_f16) swiglu_dual_kernel<true><<<blocks, 256, 0, stream>>>(raw2, static_cast<uint16_t*>(h2), M, Dff, tile);
  else swiglu_dual_kernel<false><<<blocks, 256, 0, stream>>>(raw2, static_cast<uint16_t*>(h2), M, Dff, tile);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// dx (fp32 NCHW tangent image, zero for the condition channels) -> second half of the patch-gather operand is produced by
// the ordinary patch_gather kernel; conditioning with a tangent in t:

// emb = [sin(t f) | cos(t f)] (+ aux embed);  demb = dt * w * f * [cos | -sin]
__global__ void cond_embed_dual_kernel(const float* __restrict__ t, const float* __restrict__ dt, const float* __restrict__ aux,
                                       const float* __restrict__ aux_w, const float* __restrict__ aux_b, int aux_dim,
                                       float timestep_weight, int D, float* __restrict__ emb, float* __restrict__ demb) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  const int half = D / 2;
  float e = 0.f, de = 0.f;
  if (i < 2 * half) {
    const int j = (i < half) ? i : i - half;
    const float freq = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half));
    const float arg = (t[b] * timestep_weight) * freq;
    const float darg = dt[b] * timestep_weight * freq;
    e = (i < half) ? sinf(arg) : cosf(arg);
    de = (i < half) ? cosf(arg) * darg : -sinf(arg) * darg;
  }
  if (aux != nullptr && aux_dim > 0) {
    const float s = sqrtf(static_cast<float>(aux_dim));
    float a = aux_b[i];
    for (int j = 0; j < aux_dim; ++j) a = fmaf(aux_w[i * aux_dim + j], aux[b * aux_dim + j] * s, a);
    e += a;
  }
  emb[static_cast<size_t>(b) * D + i] = e;
  demb[static_cast<size_t>(b) * D + i] = de;
}

// z = W in + bias ; out = act(z) ; dout = act'(z) * (W din)      (ACT 1 = SiLU, 0 = identity); one warp per output n
template <int ACT>
__global__ void __launch_bounds__(256) gemv_rows_dual_kernel(const float* __restrict__ Wm, const float* __restrict__ bias,
                                                             const float* __restrict__ in, const float* __restrict__ din,
                                                             float* __restrict__ out, float* __restrict__ dout, int N, int K,
                                                             int B) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const float* w = Wm + static_cast<size_t>(n) * K;
  for (int b = 0; b < B; ++b) {
    float acc = 0.f, dacc = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(w + k);
      acc = fmaf(wv, in[static_cast<size_t>(b) * K + k], acc);
      dacc = fmaf(wv, din[static_cast<size_t>(b) * K + k], dacc);
    }
    acc = warp_sum(acc);
    dacc = warp_sum(dacc);
    if (lane == 0) {
      const float z = acc + (bias ? bias[n] : 0.f);
      float o = z, d = dacc;
      if (ACT == 1) {
        const float sig = 1.0f / (1.0f + expf(-z));
        o = z * sig;
        d = dacc * sig * (1.0f + z * (1.0f - sig));
      }
      out[static_cast<size_t>(b) * N + n] = o;
      dout[static_cast<size_t>(b) * N + n] = d;
    }
  }
}

// gain = gamma (1 + scale), bias = beta (1 + scale) + shift and their tangents
__global__ void mod_finalize_dual_kernel(const float* __restrict__ mod, const float* __restrict__ dmod,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ gain, float* __restrict__ bias, float* __restrict__ dgain,
                                         float* __restrict__ dbias, int L, int B, int D) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(L) * B * D;
  if (idx >= total) return;
  const int i = idx % D;
  const int b = (idx / D) % B;
  const int l = idx / (static_cast<size_t>(D) * B);
  const size_t o = (static_cast<size_t>(b) * L + l) * 2 * D;
  const float sc = 1.0f + mod[o + i];
  gain[idx] = gamma[l * D + i] * sc;
  bias[idx] = fmaf(beta[l * D + i], sc, mod[o + D + i]);
  dgain[idx] = gamma[l * D + i] * dmod[o + i];
  dbias[idx] = fmaf(beta[l * D + i], dmod[o + i], dmod[o + D + i]);
}

int launch_conditioning_dual(const CondWeights& w, const float* t, const float* dt, const float* aux, int B, int D, int L,
                             float timestep_weight, float* scratch, float* gain, float* bias, float* dgain, float* dbias,
                             cudaStream_t stream) {
  // scratch: 2 x { emb [B,D] | h1 [B,D] | c [B,D] | mod [B, L*2D] }
  const size_t half = static_cast<size_t>(B) * (3 * D + static_cast<size_t>(L) * 2 * D);
  float *emb = scratch, *h1 = emb + static_cast<size_t>(B) * D, *c = h1 + static_cast<size_t>(B) * D, *mod = c + static_cast<size_t>(B) * D;
  float *demb = scratch + half, *dh1 = demb + static_cast<size_t>(B) * D, *dc = dh1 + static_cast<size_t>(B) * D,
        *dmod = dc + static_cast<size_t>(B) * D;
  cond_embed_dual_kernel<<<dim3((D + 127) / 128, B), 128, 0, stream>>>(t, dt, aux, w.aux_w, w.aux_b, w.aux_dim, timestep_weight,
                                                                       D, emb, demb);
  const int wpb = 8;
  gemv_rows_dual_kernel<1><<<(D + wpb - 1) / wpb, 256, 0, stream>>>(w.l1_w, w.l1_b, emb, demb, h1, dh1, D, D, B);
  gemv_rows_dual_kernel<1><<<(D + wpb - 1) / wpb, 256, 0, stream>>>(w.l2_w, w.l2_b, h1, dh1, c, dc, D, D, B);
  const int NM = L * 2 * D;
  gemv_rows_dual_kernel<0><<<(NM + wpb - 1) / wpb, 256, 0, stream>>>(w.mod_w, w.mod_b, c, dc, mod, dmod, NM, D, B);
  const size_t total = static_cast<size_t>(L) * B * D;
  mod_finalize_dual_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(mod, dmod, w.ln_gamma, w.ln_beta, gain,
                                                                                           bias, dgain, dbias, L, B, D);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
