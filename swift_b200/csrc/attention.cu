// Shifted-window scaled-cosine attention (reference: models/swinv2.py:118-135 inside the block loop :186-209).
//
// What is fused away relative to the reference:
//   * torch.roll(-s) / window_partition / window_reverse / torch.roll(+s)  (4 full-tensor copies on odd layers, 2 on
//     even ones): a window's 256 tokens are gathered straight from the un-shifted token order by index arithmetic
//     ((wy*16 + iy + s) mod gh, (wx*16 + ix + s) mod gw) and the result is scattered back to the same rows.  The
//     reference applies NO attention mask after the roll, so wrapped tokens attend freely -- same here.
//   * q/k L2-normalisation and the per-head logit scale were applied by the qkv GEMM epilogue (EPI_QKV) before the
//     single bf16 rounding, so this kernel sees q_hat*scale and k_hat; softmax scale is 1.0.
//   * "b h n d -> b n (h d)": the output is written as the [M, heads*88] bf16 operand of the wo GEMM.
//
// This file: the general-shift kernel (any shift; used when the shift is not a multiple of 8 tokens): cp.async gather
// -> smem, mma.sync m16n8k16 with an online softmax over 4 chunks of 64 keys, 16 warps x 16 query rows per
// (window, head) CTA.  The production kernel for Swift's shift of 8 is attention_tc.cu (tcgen05 / TMEM / TMA).
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace swb {

namespace att {
constexpr int kWin = 16;
constexpr int kTok = kWin * kWin;     // 256 tokens per window
constexpr int kHd = 88;
constexpr int kHdPad = 96;
constexpr int kPitch = 104;           // bf16 per smem row (208 B): 8 consecutive rows hit 8 distinct 16-byte bank groups
constexpr int kThreads = 512;
constexpr int kSmemBytes = 3 * kTok * kPitch * 2 + kTok * 4;
}  // namespace att

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  if constexpr (F16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, "
        "%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, "
        "%2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool F16>
__device__ __forceinline__ uint32_t pack2_act(float lo, float hi) {
  if constexpr (F16) {
    __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

template <bool F16, bool OUT_F16>
__global__ void __launch_bounds__(att::kThreads, 1)
window_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int M, int gh, int gw,
                        int heads, int shift_h, int shift_w) {
  using namespace att;
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + kTok * kPitch;
  __nv_bfloat16* sV = sK + kTok * kPitch;
  int* sRow = reinterpret_cast<int*>(sV + kTok * kPitch);

  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nwx = gw / kWin;
  const int wy = win / nwx, wx = win - wy * nwx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid < kTok) {
    const int iy = tid / kWin, ix = tid % kWin;
    const int gy = (wy * kWin + iy + shift_h) % gh;
    const int gx = (wx * kWin + ix + shift_w) % gw;
    sRow[tid] = (b * gh + gy) * gw + gx;
  }
  __syncthreads();

  // ---- gather q, k, v rows (96 bf16 = 12 x 16 B each) with cp.async
  for (int idx = tid; idx < 3 * kTok * 12; idx += kThreads) {
    const int part = idx / (kTok * 12);
    const int rem = idx - part * (kTok * 12);
    const int r = rem / 12, ch = rem - r * 12;
    const __nv_bfloat16* src =
        qkv + (static_cast<size_t>(part * heads + head) * M + sRow[r]) * kHdPad + ch * 8;
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(sQ + (part * kTok + r) * kPitch + ch * 8));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const uint32_t sQ_u = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));
  const uint32_t sK_u = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
  const uint32_t sV_u = static_cast<uint32_t>(__cvta_generic_to_shared(sV));
  const int q0 = warp * 16;
  const int g = lane >> 2, t4 = lane & 3;

  // ---- Q fragments: 6 k-tiles of 16 along the (padded) head dim
  uint32_t qf[6][4];
  {
    const int r = q0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int c = (lane >> 4) * 8;
#pragma unroll
    for (int kt = 0; kt < 6; ++kt)
      ldsm_x4(sQ_u + (r * kPitch + kt * 16 + c) * 2, qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3]);
  }

  float o[12][4];
#pragma unroll
  for (int i = 0; i < 12; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max / sum for rows g and g+8
  constexpr float kLog2e = 1.4426950408889634f;

#pragma unroll 1
  for (int kc = 0; kc < kTok; kc += 64) {
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const int key = kc + nt * 8 + (lane & 7);
      const int c = (lane >> 3) * 8;
#pragma unroll
      for (int kt = 0; kt < 6; kt += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sK_u + (key * kPitch + kt * 16 + c) * 2, b0, b1, b2, b3);
        mma_16816<F16>(s[nt], qf[kt], b0, b1);
        mma_16816<F16>(s[nt], qf[kt + 1], b2, b3);
      }
    }
    // online softmax (logits already carry the learned scale; softmax scale = 1)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float a0 = exp2f((m0 - mx0) * kLog2e), a1 = exp2f((m1 - mx1) * kLog2e);
    m0 = mx0;
    m1 = mx1;
    const float mb0 = mx0 * kLog2e, mb1 = mx1 * kLog2e;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(fmaf(s[nt][0], kLog2e, -mb0));
      s[nt][1] = exp2f(fmaf(s[nt][1], kLog2e, -mb0));
      s[nt][2] = exp2f(fmaf(s[nt][2], kLog2e, -mb1));
      s[nt][3] = exp2f(fmaf(s[nt][3], kLog2e, -mb1));
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      o[i][0] *= a0;
      o[i][1] *= a0;
      o[i][2] *= a1;
      o[i][3] *= a1;
    }
    // O += P V : 4 k-tiles of 16 keys, 12 n-tiles of 8 head-dim columns (cols 88..95 are zero padding)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      pf[0] = pack2_act<F16>(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack2_act<F16>(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack2_act<F16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack2_act<F16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const int key = kc + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int c = (lane >> 4) * 8;
#pragma unroll
      for (int np = 0; np < 6; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(sV_u + (key * kPitch + np * 16 + c) * 2, b0, b1, b2, b3);
        mma_16816<F16>(o[2 * np], pf, b0, b1);
        mma_16816<F16>(o[2 * np + 1], pf, b2, b3);
      }
    }
  }
  // row sums live distributed over the 4 lanes of a quad
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;

  // ---- stage the warp's 16 x 88 output tile in its own (now dead) Q rows, then write full 176-byte rows
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 11; ++nt) {
    const int c = nt * 8 + 2 * t4;
    *reinterpret_cast<uint32_t*>(sQ + (q0 + g) * kPitch + c) = pack2_act<OUT_F16>(o[nt][0] * inv0, o[nt][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + (q0 + g + 8) * kPitch + c) = pack2_act<OUT_F16>(o[nt][2] * inv1, o[nt][3] * inv1);
  }
  __syncwarp();
  const int dmodel = heads * kHd;
  for (int idx = lane; idx < 16 * 11; idx += 32) {
    const int r = idx / 11, ch = idx - r * 11;
    const uint4 v = *reinterpret_cast<const uint4*>(sQ + (q0 + r) * kPitch + ch * 8);
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(sRow[q0 + r]) * dmodel + head * kHd + ch * 8) = v;
  }
}

int launch_window_attention(const void* qkv, void* out, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                            int act_f16, int out_f16, int impl, cudaStream_t stream, float* lse) {
  using namespace att;
  SWB_REQUIRE(act_f16 || !out_f16, "window_attention: bf16 q/k/v with fp16 output is not a supported combination");
  SWB_REQUIRE(impl >= 0 && impl <= 2, "window_attention: impl must be 0 (auto), 1 (mma.sync) or 2 (tcgen05)");
  if (impl == 2 || (impl == 0 && shift_h % 8 == 0 && shift_w % 8 == 0))
    return launch_window_attention_tc(qkv, out, B, gh, gw, heads, shift_h, shift_w, act_f16, out_f16, stream, lse);
  SWB_REQUIRE(lse == nullptr, "window_attention: the log-sum-exp output needs the tcgen05 kernel (shift a multiple of 8)");
  SWB_REQUIRE(gh % kWin == 0 && gw % kWin == 0, "window_attention: token grid %dx%d not divisible by 16x16 windows",
              gh, gw);
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kSmemBytes));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kSmemBytes));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kSmemBytes));
    attr_done.set(true);
  }
  const int M = B * gh * gw;
  dim3 grid((gh / kWin) * (gw / kWin), heads, B);
  auto q_ = static_cast<const __nv_bfloat16*>(qkv);
  auto o_ = static_cast<__nv_bfloat16*>(out);
  if (act_f16 && out_f16)
    window_attention_kernel<true, true><<<grid, kThreads, kSmemBytes, stream>>>(q_, o_, M, gh, gw, heads, shift_h, shift_w);
  else if (act_f16)
    window_attention_kernel<true, false><<<grid, kThreads, kSmemBytes, stream>>>(q_, o_, M, gh, gw, heads, shift_h, shift_w);
  else
    window_attention_kernel<false, false><<<grid, kThreads, kSmemBytes, stream>>>(q_, o_, M, gh, gw, heads, shift_h, shift_w);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
