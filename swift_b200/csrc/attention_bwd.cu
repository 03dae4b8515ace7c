// Backward of the shifted-window scaled-cosine attention (models/swinv2.py:118-135 with the roll / window_partition of
// :186-209 as index arithmetic), flash style on the tensor cores (mma.sync m16n8k16, bf16 operands, fp32 accumulate):
// nothing of size tokens x tokens is ever materialised.
//
// Per item (sample, window, head), with q~ = q_hat * s (s = exp(min(scale, ln 100))), k_hat, v the packed operands the
// forward used, O the attention output and dO its gradient (both [M, heads*hd] bf16, token order):
//     S = q~ k_hat^T,  P = softmax(S),  D_i = sum_d dO_id O_id,
//     dP = dO v^T,  dS = P (dP - D),  dq~ = dS k_hat,  dk_hat = dS^T q~,  dv = P^T dO,
// then the Jacobian of the cosine normalisation (F.normalize, eps 1e-12) and of the logit scale:
//     dq = s |q|^-1 (dq~ - q_hat (q_hat . dq~)),   dk = |k|^-1 (dk_hat - k_hat (k_hat . dk_hat)),   ds = sum_i q_hat_i . dq~_i.
//
// Two kernels, both one block per (item, 64-token tile), 4 warps x 16 rows, K / V / Q / dO in 64-row chunks by cp.async:
//   attn_bwd_dq_kernel : row statistics (log-sum-exp L_i, D_i; also stored for the second kernel), dq;
//   attn_bwd_dkv_kernel: works on S^T = k_hat q~^T directly (rows = keys), so P^T and dS^T come out of the accumulators in
//                        the layout the next MMA's A fragment needs and no transpose exists anywhere.
// Output: dqkv bf16 [M, 3*D] in the reference's to_qkv column order (head, part, d): the A operand of the to_qkv dgrad /
// wgrad GEMMs.  ds partial sums: one fp32 per block (fixed-order reduction by the caller).
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace swb {

namespace {

struct BwdGeom {
  int B, gh, gw, heads, M, shift_h, shift_w, pad, hd;
};

__device__ __forceinline__ int win_row(const BwdGeom& g, int b, int win, int n) {
  const int nwx = g.gw / 16;
  const int wy = win / nwx, wx = win - wy * nwx;
  const int y = (wy * 16 + (n >> 4) + g.shift_h) % g.gh;
  const int x = (wx * 16 + (n & 15) + g.shift_w) % g.gw;
  return (b * g.gh + y) * g.gw + x;
}

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kPitch = 104;          // 16-bit elements per smem row of a 96-wide operand (208 B: conflict-free ldmatrix)
constexpr int kTile = 64;

// rows n0 .. n0+63 of the window from a packed [M][pad] slot (pad halfs = pad/8 16-byte chunks per row)
__device__ __forceinline__ void gather_packed(const uint16_t* slot, uint16_t* dst, const BwdGeom& g, int b, int win, int n0, int tid) {
  const int nch = g.pad >> 3;
  for (int idx = tid; idx < kTile * nch; idx += 128) {
    const int r = idx / nch, ch = idx - r * nch;
    const uint16_t* src = slot + static_cast<size_t>(win_row(g, b, win, n0 + r)) * g.pad + ch * 8;
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst + r * kPitch + ch * 8));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  }
}
// rows of the head's hd-column slice of a token-major [M, heads*hd] tensor; columns hd .. 95 are zeroed (MMA K = 96)
__device__ __forceinline__ void gather_tokens(const uint16_t* base, int head, uint16_t* dst, const BwdGeom& g, int b, int win, int n0,
                                              int tid) {
  const int nch = g.hd >> 3;                                     // 11 for hd = 88
  const int D = g.heads * g.hd;
  for (int idx = tid; idx < kTile * 12; idx += 128) {
    const int r = idx / 12, ch = idx - r * 12;
    uint16_t* d16 = dst + r * kPitch + ch * 8;
    if (ch < nch) {
      const uint16_t* src = base + static_cast<size_t>(win_row(g, b, win, n0 + r)) * D + head * g.hd + ch * 8;
      const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(d16));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    } else {
      *reinterpret_cast<uint4*>(d16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// acc[nt] (16 rows x 64 cols, nt = column octet) = A(16 x 96, rows ar of sa) * B^T (B rows = columns of the product, 64 x 96)
__device__ __forceinline__ void mma_rows_x_rows(float (*acc)[4], uint32_t ua, uint32_t ub, int ar, int ac, int lane) {
#pragma unroll
  for (int kt = 0; kt < 6; kt += 2) {
    uint32_t a[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) ldsm4(ua + (ar * kPitch + (kt + h) * 16 + ac) * 2, a[h][0], a[h][1], a[h][2], a[h][3]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int row = nt * 8 + (lane & 7), c = (lane >> 3) * 8;
      uint32_t b0, b1, b2, b3;
      ldsm4(ub + (row * kPitch + kt * 16 + c) * 2, b0, b1, b2, b3);
      mma_bf16(acc[nt], a[0], b0, b1);
      mma_bf16(acc[nt], a[1], b2, b3);
    }
  }
}

// out[np] (16 rows x 96 cols) += Afrag(16 x 64, from accumulators p[8][4]) * B (64 rows x 96 cols, row-major in smem)
__device__ __forceinline__ void mma_frag_x_cols(float (*out)[4], const float (*p)[4], uint32_t ub, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, c = (lane >> 4) * 8;
#pragma unroll
    for (int np = 0; np < 6; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm4_trans(ub + (row * kPitch + np * 16 + c) * 2, b0, b1, b2, b3);
      mma_bf16(out[2 * np], a, b0, b1);
      mma_bf16(out[2 * np + 1], a, b2, b3);
    }
  }
}

__device__ __forceinline__ float2 ld_bf16x2(const uint16_t* p) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ O,
                                                          const uint16_t* __restrict__ dO, const float* __restrict__ invn,
                                                          const float* __restrict__ qscale, uint16_t* __restrict__ dqkv,
                                                          float* __restrict__ Lbuf, float* __restrict__ Dbuf,
                                                          float* __restrict__ ds_part, BwdGeom g) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint16_t* sq = reinterpret_cast<uint16_t*>(smem_dyn);
  uint16_t* sdo = sq + kTile * kPitch;
  uint16_t* sk = sdo + kTile * kPitch;
  uint16_t* sv = sk + kTile * kPitch;
  float* sD = reinterpret_cast<float*>(sv + kTile * kPitch);        // [64]
  float* sred = sD + kTile;                                          // [4]
  const int item = blockIdx.y;
  const int head = item % g.heads;
  const int bw = item / g.heads;
  const int nwin = (g.gh / 16) * (g.gw / 16);
  const int win = bw % nwin, b = bw / nwin;
  const int ti = blockIdx.x * kTile;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t slot = static_cast<size_t>(g.M) * g.pad;
  const uint16_t* q_slot = qkv + static_cast<size_t>(head) * slot;
  const uint16_t* k_slot = qkv + static_cast<size_t>(g.heads + head) * slot;
  const uint16_t* v_slot = qkv + static_cast<size_t>(2 * g.heads + head) * slot;
  const uint32_t uq = static_cast<uint32_t>(__cvta_generic_to_shared(sq)), udo = static_cast<uint32_t>(__cvta_generic_to_shared(sdo));
  const uint32_t uk = static_cast<uint32_t>(__cvta_generic_to_shared(sk)), uv = static_cast<uint32_t>(__cvta_generic_to_shared(sv));
  const int ar = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ac = (lane >> 4) * 8;
  const int gq = lane >> 2, t4 = lane & 3;

  // ---- D_i = sum_d dO_id O_id (O staged in the K buffer)
  gather_packed(q_slot, sq, g, b, win, ti, tid);
  gather_tokens(dO, head, sdo, g, b, win, ti, tid);
  gather_tokens(O, head, sk, g, b, win, ti, tid);
  cp_wait_all();
  __syncthreads();
  {
    const int r = tid >> 1, half = tid & 1;                         // 2 threads per row, 48 columns each
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < 48; c += 2) {
      const float2 x = ld_bf16x2(sdo + r * kPitch + half * 48 + c), y = ld_bf16x2(sk + r * kPitch + half * 48 + c);
      a = fmaf(x.x, y.x, fmaf(x.y, y.y, a));
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (half == 0) sD[r] = a;
  }
  // ---- pass 1: log-sum-exp of every row
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
#pragma unroll 1
  for (int kc = 0; kc < 4; ++kc) {
    __syncthreads();
    gather_packed(k_slot, sk, g, b, win, kc * kTile, tid);
    cp_wait_all();
    __syncthreads();
    float s[8][4] = {};
    mma_rows_x_rows(s, uq, uk, ar, ac, lane);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float mx = mrow[h];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mx = fmaxf(mx, fmaxf(s[nt][2 * h], s[nt][2 * h + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float z = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) z += __expf(s[nt][2 * h] - mx) + __expf(s[nt][2 * h + 1] - mx);
      z += __shfl_xor_sync(0xffffffffu, z, 1);
      z += __shfl_xor_sync(0xffffffffu, z, 2);
      lrow[h] = fmaf(lrow[h], __expf(mrow[h] - mx), z);
      mrow[h] = mx;
    }
  }
  const float L[2] = {mrow[0] + __logf(lrow[0]), mrow[1] + __logf(lrow[1])};
  const float Dv[2] = {sD[warp * 16 + gq], sD[warp * 16 + gq + 8]};
  const int row_g[2] = {win_row(g, b, win, ti + warp * 16 + gq), win_row(g, b, win, ti + warp * 16 + gq + 8)};
  if (t4 == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      Lbuf[static_cast<size_t>(head) * g.M + row_g[h]] = L[h];
      Dbuf[static_cast<size_t>(head) * g.M + row_g[h]] = Dv[h];
    }
  }
  // ---- pass 2: dq~ = sum over key chunks of dS k_hat
  float dq[12][4] = {};
#pragma unroll 1
  for (int kc = 0; kc < 4; ++kc) {
    __syncthreads();
    gather_packed(k_slot, sk, g, b, win, kc * kTile, tid);
    gather_packed(v_slot, sv, g, b, win, kc * kTile, tid);
    cp_wait_all();
    __syncthreads();
    float s[8][4] = {}, dp[8][4] = {};
    mma_rows_x_rows(s, uq, uk, ar, ac, lane);
    mma_rows_x_rows(dp, udo, uv, ar, ac, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int h = j >> 1;
        const float p = __expf(s[nt][j] - L[h]);
        s[nt][j] = p * (dp[nt][j] - Dv[h]);                       // dS
      }
    mma_frag_x_cols(dq, s, uk, lane);
  }
  // ---- Jacobian of q~ = s q / |q|; ds partial
  const float sc = __ldg(qscale + head);
  float r[2] = {0.f, 0.f};
  float qv[12][4];
#pragma unroll
  for (int nt = 0; nt < 12; ++nt) {
    const int col = nt * 8 + 2 * t4;
    const float2 a = ld_bf16x2(sq + (warp * 16 + gq) * kPitch + col), c = ld_bf16x2(sq + (warp * 16 + gq + 8) * kPitch + col);
    qv[nt][0] = a.x; qv[nt][1] = a.y; qv[nt][2] = c.x; qv[nt][3] = c.y;
    r[0] = fmaf(dq[nt][0], a.x, fmaf(dq[nt][1], a.y, r[0]));
    r[1] = fmaf(dq[nt][2], c.x, fmaf(dq[nt][3], c.y, r[1]));
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    r[h] += __shfl_xor_sync(0xffffffffu, r[h], 1);
    r[h] += __shfl_xor_sync(0xffffffffu, r[h], 2);
  }
  const float inv_s = 1.0f / sc;
  const int D3 = 3 * g.heads * g.hd;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float mul = sc * __ldg(invn + static_cast<size_t>(head) * g.M + row_g[h]);     // part 0 (q) of invn [2][heads][M]
    const float rr = r[h] * inv_s * inv_s;
    uint16_t* dst = dqkv + static_cast<size_t>(row_g[h]) * D3 + (head * 3 + 0) * g.hd;
#pragma unroll
    for (int nt = 0; nt < 11; ++nt) {
      const int col = nt * 8 + 2 * t4;
      const float x0 = mul * (dq[nt][2 * h] - qv[nt][2 * h] * rr), x1 = mul * (dq[nt][2 * h + 1] - qv[nt][2 * h + 1] * rr);
      *reinterpret_cast<uint32_t*>(dst + col) = pack_bf16x2(x0, x1);
    }
  }
  // ds = sum_rows (q_hat . dq~) = sum_rows r / s: one value per block (rows g and g+8 of the lanes with t4 == 0)
  float dsum = (t4 == 0) ? (r[0] + r[1]) * inv_s : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  if (lane == 0) sred[warp] = dsum;
  __syncthreads();
  if (tid == 0) ds_part[static_cast<size_t>(item) * 4 + blockIdx.x] = (sred[0] + sred[1]) + (sred[2] + sred[3]);
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ dO,
                                                           const float* __restrict__ invn, const float* __restrict__ Lbuf,
                                                           const float* __restrict__ Dbuf, uint16_t* __restrict__ dqkv,
                                                           BwdGeom g) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  uint16_t* sk = reinterpret_cast<uint16_t*>(smem_dyn);
  uint16_t* sv = sk + kTile * kPitch;
  uint16_t* sq = sv + kTile * kPitch;
  uint16_t* sdo = sq + kTile * kPitch;
  float* sL = reinterpret_cast<float*>(sdo + kTile * kPitch);       // [64]
  float* sD = sL + kTile;                                            // [64]
  const int item = blockIdx.y;
  const int head = item % g.heads;
  const int bw = item / g.heads;
  const int nwin = (g.gh / 16) * (g.gw / 16);
  const int win = bw % nwin, b = bw / nwin;
  const int ti = blockIdx.x * kTile;                                 // this block's 64 keys
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t slot = static_cast<size_t>(g.M) * g.pad;
  const uint16_t* q_slot = qkv + static_cast<size_t>(head) * slot;
  const uint16_t* k_slot = qkv + static_cast<size_t>(g.heads + head) * slot;
  const uint16_t* v_slot = qkv + static_cast<size_t>(2 * g.heads + head) * slot;
  const uint32_t uk = static_cast<uint32_t>(__cvta_generic_to_shared(sk)), uv = static_cast<uint32_t>(__cvta_generic_to_shared(sv));
  const uint32_t uq = static_cast<uint32_t>(__cvta_generic_to_shared(sq)), udo = static_cast<uint32_t>(__cvta_generic_to_shared(sdo));
  const int ar = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ac = (lane >> 4) * 8;
  const int gq = lane >> 2, t4 = lane & 3;
  gather_packed(k_slot, sk, g, b, win, ti, tid);
  gather_packed(v_slot, sv, g, b, win, ti, tid);
  float dk[12][4] = {}, dv[12][4] = {};
#pragma unroll 1
  for (int qc = 0; qc < 4; ++qc) {
    __syncthreads();
    gather_packed(q_slot, sq, g, b, win, qc * kTile, tid);
    gather_tokens(dO, head, sdo, g, b, win, qc * kTile, tid);
    if (tid < kTile) {
      const int row = win_row(g, b, win, qc * kTile + tid);
      sL[tid] = __ldg(Lbuf + static_cast<size_t>(head) * g.M + row);
      sD[tid] = __ldg(Dbuf + static_cast<size_t>(head) * g.M + row);
    }
    cp_wait_all();
    __syncthreads();
    // S^T = k_hat q~^T and dP^T = v dO^T: rows = this warp's 16 keys, columns = the chunk's 64 queries
    float st[8][4] = {}, dpt[8][4] = {};
    mma_rows_x_rows(st, uk, uq, ar, ac, lane);
    mma_rows_x_rows(dpt, uv, udo, ar, ac, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = nt * 8 + 2 * t4 + (j & 1);
        const float p = __expf(st[nt][j] - sL[col]);
        st[nt][j] = p;                                            // P^T
        dpt[nt][j] = p * (dpt[nt][j] - sD[col]);                  // dS^T
      }
    mma_frag_x_cols(dv, st, udo, lane);                           // dv += P^T dO
    mma_frag_x_cols(dk, dpt, uq, lane);                           // dk_hat += dS^T q~
  }
  // ---- Jacobian of k_hat = k / |k|; v passes straight through
  float r[2] = {0.f, 0.f};
  float kv[12][4];
#pragma unroll
  for (int nt = 0; nt < 12; ++nt) {
    const int col = nt * 8 + 2 * t4;
    const float2 a = ld_bf16x2(sk + (warp * 16 + gq) * kPitch + col), c = ld_bf16x2(sk + (warp * 16 + gq + 8) * kPitch + col);
    kv[nt][0] = a.x; kv[nt][1] = a.y; kv[nt][2] = c.x; kv[nt][3] = c.y;
    r[0] = fmaf(dk[nt][0], a.x, fmaf(dk[nt][1], a.y, r[0]));
    r[1] = fmaf(dk[nt][2], c.x, fmaf(dk[nt][3], c.y, r[1]));
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    r[h] += __shfl_xor_sync(0xffffffffu, r[h], 1);
    r[h] += __shfl_xor_sync(0xffffffffu, r[h], 2);
  }
  const int D3 = 3 * g.heads * g.hd;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = win_row(g, b, win, ti + warp * 16 + gq + 8 * h);
    const float mul = __ldg(invn + (static_cast<size_t>(g.heads) + head) * g.M + row);     // part 1 (k) of invn
    uint16_t* dstk = dqkv + static_cast<size_t>(row) * D3 + (head * 3 + 1) * g.hd;
    uint16_t* dstv = dstk + g.hd;
#pragma unroll
    for (int nt = 0; nt < 11; ++nt) {
      const int col = nt * 8 + 2 * t4;
      const float x0 = mul * (dk[nt][2 * h] - kv[nt][2 * h] * r[h]), x1 = mul * (dk[nt][2 * h + 1] - kv[nt][2 * h + 1] * r[h]);
      *reinterpret_cast<uint32_t*>(dstk + col) = pack_bf16x2(x0, x1);
      *reinterpret_cast<uint32_t*>(dstv + col) = pack_bf16x2(dv[nt][2 * h], dv[nt][2 * h + 1]);
    }
  }
}

// qkv: packed bf16 [3][heads][M][pad] (q~ | k_hat | v); O, dO: bf16 [M, heads*hd]; invn fp32 [2][heads][M];
// Lbuf, Dbuf: fp32 [heads][M] scratch; ds_part: fp32 [items * 4]; dqkv: bf16 [M, 3*heads*hd] (head, part, d).
int launch_attention_bwd(const void* qkv, const void* O, const void* dO, const float* invn, const float* qscale, void* dqkv,
                         float* Lbuf, float* Dbuf, float* ds_part, int B, int gh, int gw, int heads, int hd, int pad,
                         int shift_h, int shift_w, cudaStream_t stream) {
  SWB_REQUIRE(gh % 16 == 0 && gw % 16 == 0 && hd == 88 && pad == 96, "attention_bwd: needs 16x16 windows and head dim 88 (pad 96)");
  BwdGeom g;
  g.B = B; g.gh = gh; g.gw = gw; g.heads = heads; g.M = B * gh * gw;
  g.shift_h = shift_h; g.shift_w = shift_w; g.pad = pad; g.hd = hd;
  const int items = B * (gh / 16) * (gw / 16) * heads;
  constexpr int kSmem = 4 * kTile * kPitch * 2 + 2 * kTile * 4 + 64;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_done.set(true);
  }
  attn_bwd_dq_kernel<<<dim3(4, items), 128, kSmem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<const uint16_t*>(O),
                                                            static_cast<const uint16_t*>(dO), invn, qscale,
                                                            static_cast<uint16_t*>(dqkv), Lbuf, Dbuf, ds_part, g);
  SWB_CHECK_CUDA(cudaGetLastError());
  attn_bwd_dkv_kernel<<<dim3(4, items), 128, kSmem, stream>>>(static_cast<const uint16_t*>(qkv), static_cast<const uint16_t*>(dO),
                                                             invn, Lbuf, Dbuf, static_cast<uint16_t*>(dqkv), g);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
