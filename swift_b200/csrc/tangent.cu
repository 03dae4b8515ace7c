// Forward-mode tangent (JVP) companions of the non-linear stages of the denoiser (first part).
//
// The sCM training loss (reference training/loss.py:186-260) needs  (F, dF) = jvp(net, (x, t), (v_x, v_t))  with the
// result detached: a pure forward pass that carries a tangent next to every activation.  Every Linear is the SAME tcgen05
// GEMM kernel run over a row-stacked operand [x ; dx] (rows 0..M-1 primal, M..2M-1 tangent: d(W x) = W dx), with the
// plain fp32 epilogue; the kernels here apply the derivative rules of what sits between the GEMMs:
//   ln_dual_kernel         ModulatedNorm + residual (models/swinv2.py:77-86, :211-212) with tangent gain / bias
//   qkv_dual_pack_kernel   F.normalize(q, k) * logit scale (:123-127) and its Jacobian, packed for the attention stage
//   attn_scores_dual / attn_out_dual (softmax fused)       windowed softmax attention (:129-135, :189-209):
//                          S = q k^T, dS = dq k^T + q dk^T,  P = softmax S,  dP = P (dS - sum_j P dS),
//                          O = P v, dO = dP v + P dv   (explicit-softmax path the reference selects with jvp=True)
//   swiglu_dual_kernel     silu(gate) * up (:99-100)
//   cond_* dual kernels    timestep embedding + latent MLP + modulation Linears (:44-60, :67-74, :84) w.r.t. t
// The score / output products of the dual attention run on the tensor cores (mma.sync m16n8k16); the other kernels are
// straightforward fp32 CUDA-core code: this path is about coverage and parity first (it runs once per training step next
// to a backward pass that is not part of this library yet).
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace swb {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// x <- x + LN(b) g + beta ;  dx <- dx + d[LN(b) g + beta]      (one warp per token row)
// The primal and tangent branch rows live in registers (NV8 groups of 8 values per lane, 16-byte accesses, every load of
// the row issued before the first use); the residual pairs are read, updated and written one 16-byte group at a time.
namespace {
__device__ __forceinline__ void load8(const float4* p, bool ok, float* o) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (ok) {
    a = __ldg(p);
    b = __ldg(p + 1);
  }
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <bool F16>
__device__ __forceinline__ void pair_update8(uint4* xh, uint4* xl, const float* add) {
  const uint4 rh = *xh, rl = *xl;
  const uint32_t wh[4] = {rh.x, rh.y, rh.z, rh.w}, wl[4] = {rl.x, rl.y, rl.z, rl.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = unpack_act2<F16>(wh[j]), c = unpack_act2<F16>(wl[j]);
    const float o0 = (a.x + c.x) + add[2 * j], o1 = (a.y + c.y) + add[2 * j + 1];
    h[j] = pack_act2<F16>(o0, o1);
    const float2 back = unpack_act2<F16>(h[j]);
    l[j] = pack_act2<F16>(o0 - back.x, o1 - back.y);
  }
  *xh = make_uint4(h[0], h[1], h[2], h[3]);
  *xl = make_uint4(l[0], l[1], l[2], l[3]);
}
}  // namespace

template <int NV8, bool F16>
__global__ void __launch_bounds__(128) ln_dual_kernel(const float* __restrict__ branch, uint16_t* __restrict__ xhl,
                                                      const float* __restrict__ gain, const float* __restrict__ bias,
                                                      const float* __restrict__ dgain, const float* __restrict__ dbias,
                                                      int M, int D, int tokens, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int ng = D >> 3;
  const float4* b4 = reinterpret_cast<const float4*>(branch + static_cast<size_t>(row) * D);
  const float4* bd4 = reinterpret_cast<const float4*>(branch + static_cast<size_t>(M + row) * D);
  float v[NV8][8], dv[NV8][8];
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    const int c = i * 32 + lane;
    load8(b4 + 2 * c, c < ng, v[i]);
    load8(bd4 + 2 * c, c < ng, dv[i]);
  }
  const float inv_d = 1.0f / static_cast<float>(D);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  const float mean = warp_sum(s) * inv_d;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i)
    if (i * 32 + lane < ng) {
#pragma unroll
      for (int j = 0; j < 8; ++j) var = fmaf(v[i][j] - mean, v[i][j] - mean, var);
    }
  const float rstd = rsqrtf(warp_sum(var) * inv_d + eps);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i)
    if (i * 32 + lane < ng) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = (v[i][j] - mean) * rstd;          // n-hat from here on
        m1 += dv[i][j];
        m2 = fmaf(v[i][j], dv[i][j], m2);
      }
    }
  m1 = warp_sum(m1) * inv_d;
  m2 = warp_sum(m2) * inv_d;
  const int smp = row / tokens;
  const float4* g4 = reinterpret_cast<const float4*>(gain + static_cast<size_t>(smp) * D);
  const float4* be4 = reinterpret_cast<const float4*>(bias + static_cast<size_t>(smp) * D);
  const float4* dg4 = reinterpret_cast<const float4*>(dgain + static_cast<size_t>(smp) * D);
  const float4* dbe4 = reinterpret_cast<const float4*>(dbias + static_cast<size_t>(smp) * D);
  uint4* xh = reinterpret_cast<uint4*>(xhl + static_cast<size_t>(row) * 2 * D);
  uint4* xdh = reinterpret_cast<uint4*>(xhl + static_cast<size_t>(M + row) * 2 * D);
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    const int c = i * 32 + lane;
    if (c < ng) {
      float g[8], be[8], dg[8], dbe[8], add[8], dadd[8];
      load8(g4 + 2 * c, true, g);
      load8(be4 + 2 * c, true, be);
      load8(dg4 + 2 * c, true, dg);
      load8(dbe4 + 2 * c, true, dbe);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float nh = v[i][j];
        const float dnh = rstd * (dv[i][j] - m1 - nh * m2);
        add[j] = fmaf(nh, g[j], be[j]);
        dadd[j] = fmaf(dnh, g[j], fmaf(nh, dg[j], dbe[j]));
      }
      pair_update8<F16>(xh + c, xh + ng + c, add);
      pair_update8<F16>(xdh + c, xdh + ng + c, dadd);
    }
  }
}

int launch_ln_dual(const float* branch2, void* xhl2, const float* gain, const float* bias, const float* dgain,
                   const float* dbias, int M, int D, int tokens, float eps, int act_f16, cudaStream_t stream) {
  SWB_REQUIRE(D % 8 == 0 && D <= 2048, "ln_dual: dim %d must be a multiple of 8 and at most 2048", D);
  SWB_REQUIRE(((reinterpret_cast<uintptr_t>(branch2) | reinterpret_cast<uintptr_t>(xhl2) | reinterpret_cast<uintptr_t>(gain) |
                reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(dgain) | reinterpret_cast<uintptr_t>(dbias)) & 15) == 0,
              "ln_dual: pointers must be 16-byte aligned");
  dim3 grid((M + 3) / 4);
  auto x_ = static_cast<uint16_t*>(xhl2);
#define SWB_LND(V)                                                                                                    \
  do {                                                                                                                \
    if (act_f16) ln_dual_kernel<V, true><<<grid, 128, 0, stream>>>(branch2, x_, gain, bias, dgain, dbias, M, D, tokens, eps);  \
    else ln_dual_kernel<V, false><<<grid, 128, 0, stream>>>(branch2, x_, gain, bias, dgain, dbias, M, D, tokens, eps);         \
  } while (0)
  const int nv8 = (D / 8 + 31) / 32;
  if (nv8 <= 2) SWB_LND(2);
  else if (nv8 <= 3) SWB_LND(3);
  else if (nv8 <= 5) SWB_LND(5);
  else SWB_LND(8);
#undef SWB_LND
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// raw [2M, 3D] fp32 (packed column order part*D + head*hd + d) -> qkv, dqkv : [3][H][M][pad] 16-bit, q / k normalised
// one warp per (row, slot): lane l < hd/4 owns elements 4l..4l+3 (float4 loads), lanes < pad/4 store 8 bytes each
template <bool F16>
__global__ void __launch_bounds__(128) qkv_dual_pack_kernel(const float* __restrict__ raw, const float* __restrict__ qscale,
                                                            uint16_t* __restrict__ out, uint16_t* __restrict__ dout, int M,
                                                            int D, int heads, int hd, int pad) {
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* r = raw + static_cast<size_t>(row) * 3 * D;
  const float* rd = raw + static_cast<size_t>(M + row) * 3 * D;
  const bool has = 4 * lane < hd;
  for (int slot = warp; slot < 3 * heads; slot += blockDim.x >> 5) {
    const int part = slot / heads, head = slot - part * heads;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), dv = v;
    if (has) {
      v = __ldg(reinterpret_cast<const float4*>(r + slot * hd) + lane);
      dv = __ldg(reinterpret_cast<const float4*>(rd + slot * hd) + lane);
    }
    float inv = 1.f, proj = 0.f, sc = 1.f;
    if (part < 2) {
      const float ss = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
      const float dot = warp_sum(v.x * dv.x + v.y * dv.y + v.z * dv.z + v.w * dv.w);
      const float nrm = sqrtf(ss);
      inv = 1.0f / fmaxf(nrm, 1e-12f);
      proj = nrm > 1e-12f ? dot * inv * inv : 0.f;         // (u . dv) / |v| with u = v / |v|
      if (part == 0) sc = qscale[head];
    }
    if (4 * lane < pad) {
      const float k = inv * sc;
      float4 a = make_float4(v.x * k, v.y * k, v.z * k, v.w * k), da = dv;     // lanes beyond hd hold zeros
      if (part < 2) da = make_float4((dv.x - v.x * proj) * k, (dv.y - v.y * proj) * k, (dv.z - v.z * proj) * k,
                                     (dv.w - v.w * proj) * k);
      const size_t o = (static_cast<size_t>(slot) * M + row) * pad + 4 * lane;
      *reinterpret_cast<uint2*>(out + o) = make_uint2(pack_act2<F16>(a.x, a.y), pack_act2<F16>(a.z, a.w));
      *reinterpret_cast<uint2*>(dout + o) = make_uint2(pack_act2<F16>(da.x, da.y), pack_act2<F16>(da.z, da.w));
    }
  }
}

int launch_qkv_dual_pack(const float* raw2, const float* qscale, void* qkv, void* dqkv, int M, int D, int heads, int hd,
                         int pad, int act_f16, cudaStream_t stream) {
  SWB_REQUIRE(hd % 4 == 0 && pad % 4 == 0 && hd <= 128 && pad <= 128 && D % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(raw2) | reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0,
              "qkv_dual_pack: head dim %d / pad %d must be multiples of 4 (<= 128) and the buffers 16-byte aligned", hd, pad);
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

// ---- mma.sync m16n8k16 helpers (16-bit operands, fp32 accumulate)
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  if constexpr (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int kDPitch = 104;          // 16-bit elements per smem row of a 96-wide operand (208 B: conflict-free ldmatrix)
constexpr int kPPitch = 264;          // ... of a 256-wide P / dP row (528 B)

// gather `rows` token rows (96 halfs = 12 x 16 B each) of one (part, head) slot into smem with cp.async
__device__ __forceinline__ void gather_rows_async(const uint16_t* src_slot, uint16_t* dst, const AttnDualGeom& g, int b, int win,
                                                  int n0, int rows, int tid, int nthreads) {
  for (int idx = tid; idx < rows * 12; idx += nthreads) {
    const int r = idx / 12, ch = idx - r * 12;
    const uint16_t* src = src_slot + static_cast<size_t>(window_token_row(g, b, win, n0 + r)) * g.pad + ch * 8;
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst + r * kDPitch + ch * 8));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  }
}

// S[i, j] = q_i . k_j ;  dS[i, j] = dq_i . k_j + q_i . dk_j      one block per (item, 64 x 64 tile); 4 warps x 16 rows,
// tensor cores (mma.sync), q / dq / k / dk rows gathered by window index arithmetic
template <bool F16>
__global__ void __launch_bounds__(128) attn_scores_dual_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dqkv,
                                                               float* __restrict__ S, float* __restrict__ dS, AttnDualGeom g) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint16_t* sq = reinterpret_cast<uint16_t*>(smem_dyn);
  uint16_t* sdq = sq + 64 * kDPitch;
  uint16_t* sk = sdq + 64 * kDPitch;
  uint16_t* sdk = sk + 64 * kDPitch;
  const int item = blockIdx.y;
  const int head = item % g.heads;
  const int bw = item / g.heads;
  const int nwin = (g.gh / 16) * (g.gw / 16);
  const int win = bw % nwin, b = bw / nwin;
  const int ti = (blockIdx.x >> 2) * 64, tj = (blockIdx.x & 3) * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t slot_q = static_cast<size_t>(head) * g.M * g.pad, slot_k = static_cast<size_t>(g.heads + head) * g.M * g.pad;
  gather_rows_async(qkv + slot_q, sq, g, b, win, ti, 64, tid, 128);
  gather_rows_async(dqkv + slot_q, sdq, g, b, win, ti, 64, tid, 128);
  gather_rows_async(qkv + slot_k, sk, g, b, win, tj, 64, tid, 128);
  gather_rows_async(dqkv + slot_k, sdk, g, b, win, tj, 64, tid, 128);
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const uint32_t uq = static_cast<uint32_t>(__cvta_generic_to_shared(sq)), udq = static_cast<uint32_t>(__cvta_generic_to_shared(sdq));
  const uint32_t uk = static_cast<uint32_t>(__cvta_generic_to_shared(sk)), udk = static_cast<uint32_t>(__cvta_generic_to_shared(sdk));
  float s[8][4] = {}, ds[8][4] = {};
  const int ar = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ac = (lane >> 4) * 8;
#pragma unroll
  for (int kt = 0; kt < 6; kt += 2) {
    uint32_t qa[2][4], dqa[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ldsm4(uq + (ar * kDPitch + (kt + h) * 16 + ac) * 2, qa[h][0], qa[h][1], qa[h][2], qa[h][3]);
      ldsm4(udq + (ar * kDPitch + (kt + h) * 16 + ac) * 2, dqa[h][0], dqa[h][1], dqa[h][2], dqa[h][3]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = nt * 8 + (lane & 7), c = (lane >> 3) * 8;
      uint32_t b0, b1, b2, b3, d0, d1, d2, d3;
      ldsm4(uk + (key * kDPitch + kt * 16 + c) * 2, b0, b1, b2, b3);
      ldsm4(udk + (key * kDPitch + kt * 16 + c) * 2, d0, d1, d2, d3);
      mma16816<F16>(s[nt], qa[0], b0, b1);
      mma16816<F16>(s[nt], qa[1], b2, b3);
      mma16816<F16>(ds[nt], dqa[0], b0, b1);
      mma16816<F16>(ds[nt], dqa[1], b2, b3);
      mma16816<F16>(ds[nt], qa[0], d0, d1);
      mma16816<F16>(ds[nt], qa[1], d2, d3);
    }
  }
  float* So = S + static_cast<size_t>(item) * 65536;
  float* dSo = dS + static_cast<size_t>(item) * 65536;
  const int gq = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = tj + nt * 8 + 2 * t4;
    const int r0 = ti + warp * 16 + gq;
    *reinterpret_cast<float2*>(So + r0 * 256 + col) = make_float2(s[nt][0], s[nt][1]);
    *reinterpret_cast<float2*>(So + (r0 + 8) * 256 + col) = make_float2(s[nt][2], s[nt][3]);
    *reinterpret_cast<float2*>(dSo + r0 * 256 + col) = make_float2(ds[nt][0], ds[nt][1]);
    *reinterpret_cast<float2*>(dSo + (r0 + 8) * 256 + col) = make_float2(ds[nt][2], ds[nt][3]);
  }
}

// O = P v ; dO = dP v + P dv   -> attn2 [2M, D] 16-bit at the tokens' own rows, column head*hd + d.
// one block per (item, 64-row tile); 4 warps x 16 rows on the tensor cores.  The row softmax and its tangent,
//     P = softmax(S),  dP = P (dS - sum_j P dS),
// happen while the 64 x 256 score rows are staged into shared memory (one warp per row, 8 consecutive keys per lane), so
// P / dP exist only as 16-bit operands in shared memory and S / dS are read from HBM exactly once.
template <bool F16>
__global__ void __launch_bounds__(128) attn_out_dual_kernel(const float* __restrict__ P, const float* __restrict__ dP,
                                                            const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dqkv,
                                                            uint16_t* __restrict__ attn2, AttnDualGeom g) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint16_t* sp = reinterpret_cast<uint16_t*>(smem_dyn);
  uint16_t* sdp = sp + 64 * kPPitch;
  uint16_t* sv = sdp + 64 * kPPitch;
  uint16_t* sdv = sv + 256 * kDPitch;
  const int item = blockIdx.y;
  const int head = item % g.heads;
  const int bw = item / g.heads;
  const int nwin = (g.gh / 16) * (g.gw / 16);
  const int win = bw % nwin, b = bw / nwin;
  const int ti = blockIdx.x * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t slot_v = static_cast<size_t>(2 * g.heads + head) * g.M * g.pad;
  gather_rows_async(qkv + slot_v, sv, g, b, win, 0, 256, tid, 128);
  gather_rows_async(dqkv + slot_v, sdv, g, b, win, 0, 256, tid, 128);
  asm volatile("cp.async.commit_group;" ::: "memory");
  const float* Pi = P + static_cast<size_t>(item) * 65536 + static_cast<size_t>(ti) * 256;
  const float* dPi = dP + static_cast<size_t>(item) * 65536 + static_cast<size_t>(ti) * 256;
#pragma unroll 2
  for (int rr = 0; rr < 16; ++rr) {                                   // this warp's 16 rows, 8 keys per lane
    const int r = warp * 16 + rr;
    const float4 a0 = *reinterpret_cast<const float4*>(Pi + r * 256 + lane * 8), a1 = *reinterpret_cast<const float4*>(Pi + r * 256 + lane * 8 + 4);
    const float4 d0 = *reinterpret_cast<const float4*>(dPi + r * 256 + lane * 8), d1 = *reinterpret_cast<const float4*>(dPi + r * 256 + lane * 8 + 4);
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    float mx = v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[i]);
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
    float dp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i] = v[i] * (dv[i] - c);
    *reinterpret_cast<uint4*>(sp + r * kPPitch + lane * 8) =
        make_uint4(pack_act2<F16>(v[0], v[1]), pack_act2<F16>(v[2], v[3]), pack_act2<F16>(v[4], v[5]), pack_act2<F16>(v[6], v[7]));
    *reinterpret_cast<uint4*>(sdp + r * kPPitch + lane * 8) =
        make_uint4(pack_act2<F16>(dp[0], dp[1]), pack_act2<F16>(dp[2], dp[3]), pack_act2<F16>(dp[4], dp[5]), pack_act2<F16>(dp[6], dp[7]));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const uint32_t up = static_cast<uint32_t>(__cvta_generic_to_shared(sp)), udp = static_cast<uint32_t>(__cvta_generic_to_shared(sdp));
  const uint32_t uv = static_cast<uint32_t>(__cvta_generic_to_shared(sv)), udv = static_cast<uint32_t>(__cvta_generic_to_shared(sdv));
  float o[12][4] = {}, dout[12][4] = {};
  const int ar = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ac = (lane >> 4) * 8;
#pragma unroll 2
  for (int kk = 0; kk < 16; ++kk) {                                   // 16 keys per step
    uint32_t pa[4], dpa[4];
    ldsm4(up + (ar * kPPitch + kk * 16 + ac) * 2, pa[0], pa[1], pa[2], pa[3]);
    ldsm4(udp + (ar * kPPitch + kk * 16 + ac) * 2, dpa[0], dpa[1], dpa[2], dpa[3]);
    const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = (lane >> 4) * 8;
#pragma unroll
    for (int np = 0; np < 6; ++np) {
      uint32_t b0, b1, b2, b3, d0, d1, d2, d3;
      ldsm4_trans(uv + (key * kDPitch + np * 16 + c) * 2, b0, b1, b2, b3);
      ldsm4_trans(udv + (key * kDPitch + np * 16 + c) * 2, d0, d1, d2, d3);
      mma16816<F16>(o[2 * np], pa, b0, b1);
      mma16816<F16>(o[2 * np + 1], pa, b2, b3);
      mma16816<F16>(dout[2 * np], dpa, b0, b1);
      mma16816<F16>(dout[2 * np + 1], dpa, b2, b3);
      mma16816<F16>(dout[2 * np], pa, d0, d1);
      mma16816<F16>(dout[2 * np + 1], pa, d2, d3);
    }
  }
  const int D = g.heads * g.hd;
  const int gq = lane >> 2, t4 = lane & 3;
  const int row0 = window_token_row(g, b, win, ti + warp * 16 + gq), row1 = window_token_row(g, b, win, ti + warp * 16 + gq + 8);
#pragma unroll
  for (int nt = 0; nt < 11; ++nt) {                                   // 11 x 8 = 88 real head-dim columns
    const int col = head * g.hd + nt * 8 + 2 * t4;
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(row0) * D + col) = pack_act2<F16>(o[nt][0], o[nt][1]);
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(row1) * D + col) = pack_act2<F16>(o[nt][2], o[nt][3]);
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(g.M + row0) * D + col) = pack_act2<F16>(dout[nt][0], dout[nt][1]);
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(g.M + row1) * D + col) = pack_act2<F16>(dout[nt][2], dout[nt][3]);
  }
}

// Fused dual attention: scores, softmax, output and their tangents in one pass over the keys, nothing materialised.
// With p~ = exp(S - m) (running row maximum m, flash-attention style) and l = sum p~:
//     O = (sum_j p~_j v_j) / l,     c = (sum_j p~_j dS_j) / l,
//     dO = sum_j dP_j v_j + P_j dv_j  with dP_j = P_j (dS_j - c)   =   (sum_j (p~_j dS_j) v_j + p~_j dv_j) / l - c O,
// so O, the combined tangent accumulator, l and c are all rescaled by exp(m_old - m_new) per 64-key chunk.
// One block per (item, 64-row tile); 4 warps x 16 rows; S / dS live in mma.sync accumulators whose layout is the next
// MMA's A fragment (p~ and p~ dS are packed straight from registers); K / dK / V / dV arrive in 64-key chunks by cp.async.
template <bool F16>
__global__ void __launch_bounds__(128) attn_fused_dual_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dqkv,
                                                              uint16_t* __restrict__ attn2, AttnDualGeom g) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint16_t* sq = reinterpret_cast<uint16_t*>(smem_dyn);
  uint16_t* sdq = sq + 64 * kDPitch;
  uint16_t* sk = sdq + 64 * kDPitch;
  uint16_t* sdk = sk + 64 * kDPitch;
  uint16_t* sv = sdk + 64 * kDPitch;
  uint16_t* sdv = sv + 64 * kDPitch;
  const int item = blockIdx.y;
  const int head = item % g.heads;
  const int bw = item / g.heads;
  const int nwin = (g.gh / 16) * (g.gw / 16);
  const int win = bw % nwin, b = bw / nwin;
  const int ti = blockIdx.x * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t slot_q = static_cast<size_t>(head) * g.M * g.pad, slot_k = static_cast<size_t>(g.heads + head) * g.M * g.pad;
  const size_t slot_v = static_cast<size_t>(2 * g.heads + head) * g.M * g.pad;
  gather_rows_async(qkv + slot_q, sq, g, b, win, ti, 64, tid, 128);
  gather_rows_async(dqkv + slot_q, sdq, g, b, win, ti, 64, tid, 128);
  const uint32_t uq = static_cast<uint32_t>(__cvta_generic_to_shared(sq)), udq = static_cast<uint32_t>(__cvta_generic_to_shared(sdq));
  const uint32_t uk = static_cast<uint32_t>(__cvta_generic_to_shared(sk)), udk = static_cast<uint32_t>(__cvta_generic_to_shared(sdk));
  const uint32_t uv = static_cast<uint32_t>(__cvta_generic_to_shared(sv)), udv = static_cast<uint32_t>(__cvta_generic_to_shared(sdv));
  float o[12][4] = {}, da[12][4] = {};
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f}, crow[2] = {0.f, 0.f};
  const int ar = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ac = (lane >> 4) * 8;
#pragma unroll 1
  for (int kc = 0; kc < 4; ++kc) {
    __syncthreads();                                                  // the previous chunk's operands are consumed
    gather_rows_async(qkv + slot_k, sk, g, b, win, kc * 64, 64, tid, 128);
    gather_rows_async(dqkv + slot_k, sdk, g, b, win, kc * 64, 64, tid, 128);
    gather_rows_async(qkv + slot_v, sv, g, b, win, kc * 64, 64, tid, 128);
    gather_rows_async(dqkv + slot_v, sdv, g, b, win, kc * 64, 64, tid, 128);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- S = q k^T, dS = dq k^T + q dk^T for this warp's 16 rows x 64 keys
    float s[8][4] = {}, ds[8][4] = {};
#pragma unroll
    for (int kt = 0; kt < 6; kt += 2) {
      uint32_t qa[2][4], dqa[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ldsm4(uq + (ar * kDPitch + (kt + h) * 16 + ac) * 2, qa[h][0], qa[h][1], qa[h][2], qa[h][3]);
        ldsm4(udq + (ar * kDPitch + (kt + h) * 16 + ac) * 2, dqa[h][0], dqa[h][1], dqa[h][2], dqa[h][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = nt * 8 + (lane & 7), c = (lane >> 3) * 8;
        uint32_t b0, b1, b2, b3, d0, d1, d2, d3;
        ldsm4(uk + (key * kDPitch + kt * 16 + c) * 2, b0, b1, b2, b3);
        ldsm4(udk + (key * kDPitch + kt * 16 + c) * 2, d0, d1, d2, d3);
        mma16816<F16>(s[nt], qa[0], b0, b1);
        mma16816<F16>(s[nt], qa[1], b2, b3);
        mma16816<F16>(ds[nt], dqa[0], b0, b1);
        mma16816<F16>(ds[nt], dqa[1], b2, b3);
        mma16816<F16>(ds[nt], qa[0], d0, d1);
        mma16816<F16>(ds[nt], qa[1], d2, d3);
      }
    }
    // ---- online softmax: row h = 0 (lane / 4) holds s[nt][0..1], row h = 1 (lane / 4 + 8) holds s[nt][2..3]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float mx = mrow[h];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mx = fmaxf(mx, fmaxf(s[nt][2 * h], s[nt][2 * h + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float scale = __expf(mrow[h] - mx);
      mrow[h] = mx;
      float z = 0.f, e = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float pj = __expf(s[nt][2 * h + j] - mx);
          s[nt][2 * h + j] = pj;                       // p~
          ds[nt][2 * h + j] *= pj;                     // p~ dS
          z += pj;
          e += ds[nt][2 * h + j];
        }
      }
      z += __shfl_xor_sync(0xffffffffu, z, 1);
      z += __shfl_xor_sync(0xffffffffu, z, 2);
      e += __shfl_xor_sync(0xffffffffu, e, 1);
      e += __shfl_xor_sync(0xffffffffu, e, 2);
      lrow[h] = fmaf(lrow[h], scale, z);
      crow[h] = fmaf(crow[h], scale, e);
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) {
        o[nt][2 * h] *= scale; o[nt][2 * h + 1] *= scale;
        da[nt][2 * h] *= scale; da[nt][2 * h + 1] *= scale;
      }
    }
    // ---- O += p~ v ; dacc += (p~ dS) v + p~ dv      (accumulator pairs of two key octets = one k16 A fragment)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4], ea[4];
      pa[0] = pack_act2<F16>(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_act2<F16>(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_act2<F16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_act2<F16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      ea[0] = pack_act2<F16>(ds[2 * kk][0], ds[2 * kk][1]);
      ea[1] = pack_act2<F16>(ds[2 * kk][2], ds[2 * kk][3]);
      ea[2] = pack_act2<F16>(ds[2 * kk + 1][0], ds[2 * kk + 1][1]);
      ea[3] = pack_act2<F16>(ds[2 * kk + 1][2], ds[2 * kk + 1][3]);
      const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = (lane >> 4) * 8;
#pragma unroll
      for (int np = 0; np < 6; ++np) {
        uint32_t b0, b1, b2, b3, d0, d1, d2, d3;
        ldsm4_trans(uv + (key * kDPitch + np * 16 + c) * 2, b0, b1, b2, b3);
        ldsm4_trans(udv + (key * kDPitch + np * 16 + c) * 2, d0, d1, d2, d3);
        mma16816<F16>(o[2 * np], pa, b0, b1);
        mma16816<F16>(o[2 * np + 1], pa, b2, b3);
        mma16816<F16>(da[2 * np], ea, b0, b1);
        mma16816<F16>(da[2 * np + 1], ea, b2, b3);
        mma16816<F16>(da[2 * np], pa, d0, d1);
        mma16816<F16>(da[2 * np + 1], pa, d2, d3);
      }
    }
  }
  const int D = g.heads * g.hd;
  const int gq = lane >> 2, t4 = lane & 3;
  const int row0 = window_token_row(g, b, win, ti + warp * 16 + gq), row1 = window_token_row(g, b, win, ti + warp * 16 + gq + 8);
  const float inv0 = 1.0f / lrow[0], inv1 = 1.0f / lrow[1];
  const float c0 = crow[0] * inv0, c1 = crow[1] * inv1;
#pragma unroll
  for (int nt = 0; nt < 11; ++nt) {                                   // 11 x 8 = 88 real head-dim columns
    const int col = head * g.hd + nt * 8 + 2 * t4;
    const float oa = o[nt][0] * inv0, ob = o[nt][1] * inv0, oc = o[nt][2] * inv1, od = o[nt][3] * inv1;
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(row0) * D + col) = pack_act2<F16>(oa, ob);
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(row1) * D + col) = pack_act2<F16>(oc, od);
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(g.M + row0) * D + col) =
        pack_act2<F16>(fmaf(-c0, oa, da[nt][0] * inv0), fmaf(-c0, ob, da[nt][1] * inv0));
    *reinterpret_cast<uint32_t*>(attn2 + static_cast<size_t>(g.M + row1) * D + col) =
        pack_act2<F16>(fmaf(-c1, oc, da[nt][2] * inv1), fmaf(-c1, od, da[nt][3] * inv1));
  }
}

int launch_attention_dual(const void* qkv, const void* dqkv, float* S, float* dS, void* attn2, int B, int gh, int gw,
                          int heads, int hd, int pad, int shift_h, int shift_w, int act_f16, cudaStream_t stream) {
  SWB_REQUIRE(gh % 16 == 0 && gw % 16 == 0 && hd <= 96, "attention_dual: grid %dx%d / head_dim %d unsupported", gh, gw, hd);
  // the tcgen05 / TMEM kernel (attention_dual_tc.cu) for shifts that are multiples of 8; SWB_DUAL_ATTN_MMA: A/B knob (tools)
  static const bool force_mma = getenv("SWB_DUAL_ATTN_MMA") != nullptr || getenv("SWB_DUAL_ATTN_SPLIT") != nullptr;
  if (!force_mma && hd == 88 && pad == 96 && shift_h % 8 == 0 && shift_w % 8 == 0)
    return launch_attention_dual_tc(qkv, dqkv, attn2, B, gh, gw, heads, shift_h, shift_w, act_f16, stream);
  AttnDualGeom g;
  g.B = B; g.gh = gh; g.gw = gw; g.heads = heads; g.M = B * gh * gw;
  g.shift_h = shift_h; g.shift_w = shift_w; g.pad = pad; g.hd = hd;
  const int items = B * (gh / 16) * (gw / 16) * heads;
  const auto* q = static_cast<const uint16_t*>(qkv);
  const auto* dq = static_cast<const uint16_t*>(dqkv);
  constexpr int kScoresSmem = 4 * 64 * kDPitch * 2;                          // 53 KB
  constexpr int kOutSmem = 2 * 64 * kPPitch * 2 + 2 * 256 * kDPitch * 2;     // 174 KB
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_scores_dual_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScoresSmem));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_scores_dual_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScoresSmem));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_out_dual_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOutSmem));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_out_dual_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOutSmem));
    attr_done.set(true);
  }
  SWB_REQUIRE(hd == 88 && pad == 96, "attention_dual: the tensor-core kernels are specialised for head_dim 88 padded to 96");
  static const bool split = getenv("SWB_DUAL_ATTN_SPLIT") != nullptr;        // tools only: the three-stage A/B path
  if (!split) {
    constexpr int kFusedSmem = 6 * 64 * kDPitch * 2;                          // 78 KB: two blocks per SM
    static PerDevice<bool> fused_attr;
    if (!fused_attr.get()) {
      SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_fused_dual_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
      SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_fused_dual_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
      fused_attr.set(true);
    }
    if (act_f16) attn_fused_dual_kernel<true><<<dim3(4, items), 128, kFusedSmem, stream>>>(q, dq, static_cast<uint16_t*>(attn2), g);
    else attn_fused_dual_kernel<false><<<dim3(4, items), 128, kFusedSmem, stream>>>(q, dq, static_cast<uint16_t*>(attn2), g);
    SWB_CHECK_CUDA(cudaGetLastError());
    return SWB_OK;
  }
  if (act_f16) attn_scores_dual_kernel<true><<<dim3(16, items), 128, kScoresSmem, stream>>>(q, dq, S, dS, g);
  else attn_scores_dual_kernel<false><<<dim3(16, items), 128, kScoresSmem, stream>>>(q, dq, S, dS, g);
  if (act_f16) attn_out_dual_kernel<true><<<dim3(4, items), 128, kOutSmem, stream>>>(S, dS, q, dq, static_cast<uint16_t*>(attn2), g);
  else attn_out_dual_kernel<false><<<dim3(4, items), 128, kOutSmem, stream>>>(S, dS, q, dq, static_cast<uint16_t*>(attn2), g);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// raw [2M, 2*Dff] fp32, columns in packed tile order (every `tile` columns: [tile/2 gate | tile/2 up]) -> h2 [2M, Dff] 16-bit
// four consecutive outputs per thread: float4 loads of gate / up / dgate / dup, one 8-byte store per output row
template <bool F16>
__global__ void __launch_bounds__(256) swiglu_dual_kernel(const float* __restrict__ raw, uint16_t* __restrict__ h2, int M,
                                                          int Dff, int tile) {
  const int q = Dff >> 2;                                    // groups of 4 outputs per row (tile/2 is a multiple of 4)
  const size_t total = static_cast<size_t>(M) * q;
  const int half = tile / 2;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = idx / q;
    const int j = static_cast<int>(idx - row * q) * 4;
    const int col = (j / half) * tile + (j % half);
    const float* r = raw + row * 2 * Dff + col;
    const float* rd = raw + (M + row) * 2 * Dff + col;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(r)), u4 = __ldg(reinterpret_cast<const float4*>(r + half));
    const float4 dg4 = __ldg(reinterpret_cast<const float4*>(rd)), du4 = __ldg(reinterpret_cast<const float4*>(rd + half));
    const float gte[4] = {g4.x, g4.y, g4.z, g4.w}, up[4] = {u4.x, u4.y, u4.z, u4.w};
    const float dg[4] = {dg4.x, dg4.y, dg4.z, dg4.w}, du[4] = {du4.x, du4.y, du4.z, du4.w};
    float o[4], d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sig = 1.0f / (1.0f + __expf(-gte[k]));
      const float silu = gte[k] * sig;
      const float dsilu = sig * (1.0f + gte[k] * (1.0f - sig));
      o[k] = silu * up[k];
      d[k] = fmaf(dsilu * dg[k], up[k], silu * du[k]);
    }
    *reinterpret_cast<uint2*>(h2 + row * Dff + j) = make_uint2(pack_act2<F16>(o[0], o[1]), pack_act2<F16>(o[2], o[3]));
    *reinterpret_cast<uint2*>(h2 + (M + row) * Dff + j) = make_uint2(pack_act2<F16>(d[0], d[1]), pack_act2<F16>(d[2], d[3]));
  }
}

int launch_swiglu_dual(const float* raw2, void* h2, int M, int Dff, int tile, int act_f16, cudaStream_t stream) {
  SWB_REQUIRE(Dff % 4 == 0 && tile % 8 == 0 && ((reinterpret_cast<uintptr_t>(raw2) | reinterpret_cast<uintptr_t>(h2)) & 15) == 0,
              "swiglu_dual: Dff %d / tile %d must be multiples of 4 / 8 and the buffers 16-byte aligned", Dff, tile);
  const size_t total = static_cast<size_t>(M) * (Dff / 4);
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
