// Backward of the shifted-window scaled-cosine attention on tcgen05 / TMEM, fed by TMA (reference forward:
// models/swinv2.py:118-135, :186-209; what autograd differentiates in training/loss.py:226-260).
//
// One work item = (sample, window, head), 256 tokens.  q~ = q_hat * s, k_hat, v are the packed bf16 operands of the forward
// ([3][heads][M][96]); O / dO are token-major [M, heads*88]; L (log-sum-exp of every score row, written by the forward
// kernel) and D_i = sum_d dO_id O_id (attn_bwd_prep_kernel) are fp32 [heads][M].  A persistent CTA per SM:
//
//   warp 0      TMA: the window's q~, k_hat, v, dO as four 8x8-token boxes each (d[0,64) SWIZZLE_128B + d[64,96) SWIZZLE_64B);
//               roll / window_partition are box coordinates; the next item is prefetched into L2 meanwhile.
//   warp 1      UMMA issuer.  Four units per item, each [two 128x256x96 products] -> elementwise -> [products from TMEM]:
//                 A_h (query half h):  S = q~_h k^T | dP = dO_h v^T  ->  dS = P (dP - D)  ->  dq~_h = dS k_hat
//                 B_j (key half j):    S^T = k_j q~^T | dP^T = v_j dO^T  ->  P^T, dS^T  ->  dv_j = P^T dO,  dk_j = dS^T q~
//               The transposed products are computed directly (operands swapped), so P^T / dS^T sit in TMEM with lanes =
//               keys -- exactly the A operand the second stage needs; nothing is ever transposed.  Second-stage B operands
//               (k_hat, dO, q~ as [token][d]) are the SAME shared-memory tiles read MN-major.
//   warps 4-11  elementwise: thread = (TMEM lane = row of the unit, column half c): P = exp(S - L), dS = P (dP - D) from two
//               512-column fp32 accumulators, written back as packed bf16 behind the read pointer (the TS-UMMA A operand);
//               then the accumulators are read out, the Jacobian of the cosine normalisation is applied
//               (dq = s |q|^-1 (dq~ - q_hat (q_hat . dq~)), same for k) and rows leave through a shared-memory transpose as the
//               bf16 [M, 3*D] operand (head, q|k|v, d) of the to_qkv dgrad / wgrad GEMMs.
//
// TMEM columns of a unit: first product [0,256), second [256,512); packed outputs: thread column half c writes its own
// input range ([0,64) | [128,192) and [256,320) | [384,448)); second-stage accumulators use what is free by then.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace swb {

namespace abt {
constexpr int kThreads = 384;
constexpr int kHd = 88;
constexpr int kHdPad = 96;
// shared memory map (offsets from a 1024-byte aligned base); every operand: 256 x 128 B (SW128) + 256 x 64 B (SW64)
constexpr uint32_t kOp = 49152;
constexpr uint32_t kQ0 = 0, kQ1 = 32768;
constexpr uint32_t kK0 = kOp, kK1 = kOp + 32768;
constexpr uint32_t kV0 = 2 * kOp, kV1 = 2 * kOp + 32768;
constexpr uint32_t kG0 = 3 * kOp, kG1 = 3 * kOp + 32768;            // dO
constexpr uint32_t kVec = 4 * kOp;                                  // L[2][256], D[2][256] fp32 (double buffered per item)
constexpr uint32_t kScratch = kVec + 4096;                          // 8 warps x 2 KB
constexpr uint32_t kBars = kScratch + 8 * 2048;
constexpr uint32_t kSmemBytes = kBars + 256 + 1024;
constexpr uint32_t kLoadBytes = 4 * kOp;
enum Bar { LOAD_FULL = 0, LOAD_EMPTY, SDP_FULL, DS_FULL, ACC_FULL, ACC_FREE, EPI_DONE, NBARS };
}  // namespace abt

struct AttnBwdTcParams {
  int B, gh, gw, heads, M;
  int shift_by, shift_bx;        // cyclic shift in units of 8 tokens
  const float* L;                // [heads][M]
  const float* D;                // [heads][M]
  const float* invn;             // [2][heads][M]
  const float* qscale;           // [heads]
  uint16_t* dqkv;                // [M, 3*heads*88]
  float* ds_part;                // [items][8]
};

namespace {

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// element (row, d) of an operand tile as the TMA wrote it: d < 64 in the SWIZZLE_128B chunk, else in the SWIZZLE_64B chunk
__device__ __forceinline__ uint32_t op_addr16(uint32_t base0, uint32_t base1, int row, int chunk16) {
  // chunk16: index of the 16-byte chunk along d (0..11)
  if (chunk16 < 8) return base0 + row * 128 + ((chunk16 ^ (row & 7)) << 4);
  return base1 + row * 64 + (((chunk16 - 8) ^ ((row >> 1) & 3)) << 4);
}

// 32 rows (lane = row; the rows are 4 grid lines of 8 tokens) x NCH 16-byte chunks: stage in this warp's 2 KB scratch, then
// full-row-segment global stores (pitch1: bytes between tokens of a line, pitch8: between lines)
template <int NCH>
__device__ __forceinline__ void warp_store_rows(uint8_t* g_row0, size_t pitch1, size_t pitch8, const uint4* v, uint32_t scratch,
                                                int lane) {
  static_assert(NCH <= 4, "2 KB of scratch per warp");
#pragma unroll
  for (int c = 0; c < NCH; ++c) st_shared_v4(scratch + lane * 64 + ((c ^ (lane & 3)) << 4), v[c]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / NCH, c = idx - r * NCH;
    const uint4 q = ld_shared_v4(scratch + r * 64 + ((c ^ (r & 3)) << 4));
    *reinterpret_cast<uint4*>(g_row0 + (r >> 3) * pitch8 + (r & 7) * pitch1 + c * 16) = q;
  }
  __syncwarp();
}

}  // namespace

// D[head][row] = sum_d dO[row, head*88 + d] * O[row, head*88 + d]
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const uint16_t* __restrict__ O, const uint16_t* __restrict__ dO,
                                                            float* __restrict__ D, int M, int heads) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(M) * heads) return;
  const int row = static_cast<int>(idx / heads), head = static_cast<int>(idx - static_cast<long long>(row) * heads);
  const size_t off = (static_cast<size_t>(row) * heads + head) * abt::kHd;
  const uint4* o4 = reinterpret_cast<const uint4*>(O + off);
  const uint4* g4 = reinterpret_cast<const uint4*>(dO + off);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 11; ++c) {
    const uint4 a = __ldg(o4 + c), b = __ldg(g4 + c);
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc = fmaf(__uint_as_float(wa[j] << 16), __uint_as_float(wb[j] << 16), acc);
      acc = fmaf(__uint_as_float(wa[j] & 0xffff0000u), __uint_as_float(wb[j] & 0xffff0000u), acc);
    }
  }
  D[static_cast<size_t>(head) * M + row] = acc;
}

__global__ void __launch_bounds__(abt::kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tq128, const __grid_constant__ CUtensorMap tq64,
                   const __grid_constant__ CUtensorMap tg128, const __grid_constant__ CUtensorMap tg64, const AttnBwdTcParams p) {
  using namespace abt;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sb_ptr = smem_raw + (sb - smem_u32(smem_raw));
  const uint32_t bars = sb + kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * NBARS;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(sb_ptr + kBars + 8u * NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwx = p.gw / 16, nwin = (p.gh / 16) * nwx;
  const int nbx = p.gw / 8, nby = p.gh / 8;
  const int num_items = p.B * nwin * p.heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tq128);
    tma_prefetch_desc(&tq64);
    tma_prefetch_desc(&tg128);
    tma_prefetch_desc(&tg64);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar(LOAD_FULL), 1);
    mbar_init(bar(LOAD_EMPTY), 1);
    mbar_init(bar(SDP_FULL), 1);
    mbar_init(bar(DS_FULL), 8);
    mbar_init(bar(ACC_FULL), 1);
    mbar_init(bar(ACC_FREE), 8);
    mbar_init(bar(EPI_DONE), 8);
    fence_mbar_init_cluster();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  // box coordinates of 8x8-token block `blk` of the item's window
  auto box_xy = [&](int item, int blk, int& x0, int& y0) {
    const int bw = item / p.heads;
    const int win = bw % nwin, b = bw / nwin;
    const int wy = win / nwx, wx = win - wy * nwx;
    x0 = ((2 * wx + (blk & 1) + p.shift_bx) % nbx) * 8;
    y0 = ((2 * wy + (blk >> 1) + p.shift_by) % nby) * 8 + b * p.gh;
  };

  if (warp == 0) {
    // =========================================== TMA loader ===========================================
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int head = item % p.heads;
      mbar_wait(bar(LOAD_EMPTY), par ^ 1u, 11);       // every MMA of the previous item has retired ...
      mbar_wait(bar(EPI_DONE), par ^ 1u, 12);         // ... and its epilogues have read their q~ / k_hat rows
      mbar_arrive_expect_tx_elect(bar(LOAD_FULL), kLoadBytes);
#pragma unroll 1
      for (int part = 0; part < 3; ++part) {
        const int slot = part * p.heads + head;
        const uint32_t d0 = sb + part * kOp, d1 = d0 + 32768;
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          int x0, y0;
          box_xy(item, blk, x0, y0);
          tma_load_4d_elect(d0 + blk * 8192, &tq128, bar(LOAD_FULL), 0, x0, y0, slot);
          tma_load_4d_elect(d1 + blk * 4096, &tq64, bar(LOAD_FULL), 64, x0, y0, slot);
        }
      }
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        int x0, y0;
        box_xy(item, blk, x0, y0);
        tma_load_4d_elect(sb + kG0 + blk * 8192, &tg128, bar(LOAD_FULL), head * kHd, x0, y0, 0);
        tma_load_4d_elect(sb + kG1 + blk * 4096, &tg64, bar(LOAD_FULL), head * kHd + 64, x0, y0, 0);
      }
      // pull the next item's operands into L2 while this one computes (shared memory is single-buffered)
      const int nxt = item + gridDim.x;
      if (nxt < num_items && lane == 0) {
        const int nh = nxt % p.heads;
        for (int blk = 0; blk < 4; ++blk) {
          int x0, y0;
          box_xy(nxt, blk, x0, y0);
          for (int part = 0; part < 3; ++part) {
            tma_prefetch_4d(&tq128, 0, x0, y0, part * p.heads + nh);
            tma_prefetch_4d(&tq64, 64, x0, y0, part * p.heads + nh);
          }
          tma_prefetch_4d(&tg128, nh * kHd, x0, y0, 0);
          tma_prefetch_4d(&tg64, nh * kHd + 64, x0, y0, 0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================================== UMMA issuer ===========================================
    constexpr uint32_t idesc_big = make_idesc_f16(128, 256, false, false, 0, 0);     // both operands K-major
    constexpr uint32_t idesc_n64 = make_idesc_f16(128, 64, false, false, 0, 1);      // A from TMEM, B MN-major
    constexpr uint32_t idesc_n32 = make_idesc_f16(128, 32, false, false, 0, 1);
    const uint64_t hi128 = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint64_t hi64 = make_smem_desc(0, 16, 512, SWZ_64B);
    auto lo = [](uint32_t addr) { return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4); };
    // D[tmem d] = A(128 rows at row offset ar of operand a) * B(all 256 rows of operand b)^T over d = 96
    auto product = [&](uint32_t d, uint32_t a_base, int a_half, uint32_t b_base) {
      const uint64_t a0 = hi128 | lo(sb + a_base + a_half * 16384), b0 = hi128 | lo(sb + b_base);
      const uint64_t a1 = hi64 | lo(sb + a_base + 32768 + a_half * 8192), b1 = hi64 | lo(sb + b_base + 32768);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss_elect<1>(d, a0 + 2u * k, b0 + 2u * k, idesc_big, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_f16_ss_elect<1>(d, a1 + 2u * k, b1 + 2u * k, idesc_big, 1u);
    };
    // D[tmem d64 | d32] = A(packed bf16 in TMEM: 256 K values at columns a_lo.. (k < 128) and a_hi.. (k >= 128)) * B(256 x 96, MN-major)
    auto from_tmem = [&](uint32_t d64, uint32_t d32, uint32_t a_lo, uint32_t a_hi, uint32_t b_base) {
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t a = (ks < 8 ? a_lo + 8u * ks : a_hi + 8u * (ks - 8));
        const uint64_t b0 = hi128 | lo(sb + b_base + 2048u * ks);
        const uint64_t b1 = hi64 | lo(sb + b_base + 32768u + 1024u * ks);
        umma_f16_ts_elect(d64, a, b0, idesc_n64, ks != 0 ? 1u : 0u);
        umma_f16_ts_elect(d32, a, b1, idesc_n32, ks != 0 ? 1u : 0u);
      }
    };
    int it = 0;
    uint32_t n = 0;                                    // units issued so far (4 per item): barrier phase = n & 1
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      mbar_wait(bar(LOAD_FULL), it & 1, 21);
      tcgen05_fence_after();
#pragma unroll 1
      for (int u = 0; u < 4; ++u, ++n) {
        const int half = u & 1;
        mbar_wait(bar(ACC_FREE), (n & 1u) ^ 1u, 22);   // the previous unit's accumulators have been read out
        tcgen05_fence_after();
        if (u < 2) {                                   // A_h: S = q~_h k^T ; dP = dO_h v^T
          product(tmem, kQ0, half, kK0);
          product(tmem + 256, kG0, half, kV0);
        } else {                                       // B_j: S^T = k_j q~^T ; dP^T = v_j dO^T
          product(tmem, kK0, half, kQ0);
          product(tmem + 256, kV0, half, kG0);
        }
        umma_commit_elect<1>(bar(SDP_FULL));
        mbar_wait(bar(DS_FULL), n & 1u, 23);           // packed P / dS are in TMEM
        tcgen05_fence_after();
        if (u < 2) {
          from_tmem(tmem + 256, tmem + 320, tmem, tmem + 128, kK0);                       // dq~_h = dS k_hat
        } else {
          from_tmem(tmem + 64, tmem + 192, tmem, tmem + 128, kG0);                        // dv_j = P^T dO
          from_tmem(tmem + 320, tmem + 448, tmem + 256, tmem + 384, kQ0);                 // dk_j = dS^T q~
        }
        umma_commit_elect<1>(bar(ACC_FULL));
      }
      umma_commit_elect<1>(bar(LOAD_EMPTY));           // operand tiles reusable once every MMA of the item has retired
    }
  } else if (warp >= 4) {
    // =========================================== elementwise / output ===========================================
    const int quad = warp & 3;                         // TMEM lane quadrant of this warp
    const int c = (warp - 4) >> 2;                     // column half handled in the elementwise pass
    const int et = threadIdx.x - 128;                  // 0..255
    const int lrow = quad * 32 + lane;                 // row of the unit (TMEM lane)
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t scratch = sb + kScratch + (warp - 4) * 2048;
    float* vecs = reinterpret_cast<float*>(sb_ptr + kVec);
    constexpr float kLog2e = 1.4426950408889634f;
    const int D3 = 3 * p.heads * kHd;
    const size_t opitch = static_cast<size_t>(D3) * 2;
    int it = 0;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int head = item % p.heads;
      // ---- L, D of the window's 256 tokens in shared-memory row order (block, line, x)
      float* sL = vecs + (it & 1) * 512;
      float* sD = sL + 256;
      {
        const int blk = et >> 6, line = (et >> 3) & 7, x = et & 7;
        int x0, y0;
        box_xy(item, blk, x0, y0);
        const size_t grow = (static_cast<size_t>(y0) + line) * p.gw + x0 + x;
        sL[et] = __ldg(p.L + static_cast<size_t>(head) * p.M + grow) * kLog2e;
        sD[et] = __ldg(p.D + static_cast<size_t>(head) * p.M + grow);
      }
      bar_sync_named(1, 256);
      mbar_wait(bar(LOAD_FULL), it & 1, 31);           // q~ / k_hat rows are read from shared memory in the epilogues
      const float sc = __ldg(p.qscale + head);
      float ds_acc = 0.f;
#pragma unroll 1
      for (int u = 0; u < 4; ++u, ++n) {
        const int half = u & 1;
        const int urow = half * 128 + lrow;            // this thread's row of the unit among the window's 256 tokens
        mbar_wait(bar(SDP_FULL), n & 1u, 32);
        tcgen05_fence_after();
        const uint32_t t_s = tl + c * 128, t_dp = tl + 256 + c * 128;
        if (u < 2) {
          // rows = queries: L, D of this thread's own row
          const float Lr = sL[urow], Dr = sD[urow];
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            float s[32], dp[32];
            __syncwarp();
            tmem_ld_x32(t_s + 32 * i, s);
            tmem_ld_x32(t_dp + 32 * i, dp);
            tmem_ld_wait();
            tmem_ld_fence_regs<32>(s);
            tmem_ld_fence_regs<32>(dp);
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float p0 = ex2f(fmaf(s[2 * j], kLog2e, -Lr)), p1 = ex2f(fmaf(s[2 * j + 1], kLog2e, -Lr));
              w[j] = pack_bf16x2(p0 * (dp[2 * j] - Dr), p1 * (dp[2 * j + 1] - Dr));
            }
            tmem_st_x16(t_s + 16 * i, w);
          }
        } else {
          // rows = keys, columns = queries: L, D per column (broadcast reads)
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            float s[32], dp[32];
            __syncwarp();
            tmem_ld_x32(t_s + 32 * i, s);
            tmem_ld_x32(t_dp + 32 * i, dp);
            tmem_ld_wait();
            tmem_ld_fence_regs<32>(s);
            tmem_ld_fence_regs<32>(dp);
            const float4* L4 = reinterpret_cast<const float4*>(sL + c * 128 + 32 * i);
            const float4* D4 = reinterpret_cast<const float4*>(sD + c * 128 + 32 * i);
            uint32_t wp[16], wd[16];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 l = L4[j4], d = D4[j4];
              const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {d.x, d.y, d.z, d.w};
              float pv[4], gv[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                pv[k] = ex2f(fmaf(s[4 * j4 + k], kLog2e, -lv[k]));
                gv[k] = pv[k] * (dp[4 * j4 + k] - dv[k]);
              }
              wp[2 * j4] = pack_bf16x2(pv[0], pv[1]);
              wp[2 * j4 + 1] = pack_bf16x2(pv[2], pv[3]);
              wd[2 * j4] = pack_bf16x2(gv[0], gv[1]);
              wd[2 * j4 + 1] = pack_bf16x2(gv[2], gv[3]);
            }
            tmem_st_x16(t_s + 16 * i, wp);
            tmem_st_x16(t_dp + 16 * i, wd);
          }
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(DS_FULL));
        // ---- second-stage accumulators
        mbar_wait(bar(ACC_FULL), n & 1u, 33);
        tcgen05_fence_after();
        // what this warp reads out: A units: column half 0 -> dq~ (96 columns at 256), half 1 -> nothing;
        //                           B units: half 0 -> dv (64 at 64 | 32 at 192), half 1 -> dk (64 at 320 | 32 at 448)
        const bool active = (u >= 2) || (c == 0);
        float acc[kHd];
        if (active) {
          const uint32_t a64 = tl + (u < 2 ? 256u : (c == 0 ? 64u : 320u));
          const uint32_t a32 = tl + (u < 2 ? 320u : (c == 0 ? 192u : 448u));
          __syncwarp();
          tmem_ld_x32(a64, acc);
          tmem_ld_x32(a64 + 32, acc + 32);
          tmem_ld_x16(a32, acc + 64);
          tmem_ld_x8(a32 + 16, acc + 80);
          tmem_ld_wait();
          tmem_ld_fence_regs<kHd>(acc);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(ACC_FREE));
        if (!active) continue;
        // ---- Jacobian of the normalisation (q and k), then rows out.  part: 0 = q, 1 = k, 2 = v
        const int part = (u < 2) ? 0 : (c == 0 ? 2 : 1);
        const int blk = urow >> 6;                     // the warp's 32 rows: block blk, lines 4*((urow >> 5) & 1) .. +3
        int x0, y0;
        box_xy(item, blk, x0, y0);
        const size_t grow0 = (static_cast<size_t>(y0) + 4 * ((urow >> 5) & 1)) * p.gw + x0;      // row of lane 0
        const size_t grow = grow0 + (lane >> 3) * p.gw + (lane & 7);
        uint8_t* g = reinterpret_cast<uint8_t*>(p.dqkv) + grow0 * opitch + static_cast<size_t>((head * 3 + part) * kHd) * 2;
        const uint32_t b0 = sb + (part == 0 ? kQ0 : kK0), b1 = b0 + 32768;
        float mul = 1.0f, rr = 0.f;
        if (part != 2) {
          float r = 0.f;                               // r = x . d x~ over the row (x = q~ or k_hat, from shared memory)
#pragma unroll
          for (int ch = 0; ch < 11; ++ch) {
            const uint4 q4 = ld_shared_v4(op_addr16(b0, b1, urow, ch));
            const uint32_t ww[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              r = fmaf(acc[8 * ch + 2 * k], __uint_as_float(ww[k] << 16), r);
              r = fmaf(acc[8 * ch + 2 * k + 1], __uint_as_float(ww[k] & 0xffff0000u), r);
            }
          }
          const float inv = __ldg(p.invn + (static_cast<size_t>(part) * p.heads + head) * p.M + grow);
          if (part == 0) {                             // x = q~ = s q_hat:  dq = s inv (dq~ - q~ r / s^2);  ds += r / s
            mul = sc * inv;
            rr = r / (sc * sc);
            ds_acc += r / sc;
          } else {                                     // x = k_hat:         dk = inv (dk_hat - k_hat r)
            mul = inv;
            rr = r;
          }
        }
        // rows out in three column groups of 4 + 4 + 3 sixteen-byte chunks (the row again from shared memory for q / k)
#pragma unroll
        for (int grp = 0; grp < 3; ++grp) {
          const int nch = grp < 2 ? 4 : 3;
          uint4 w4[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (cc >= nch) continue;
            const int ch = grp * 4 + cc;
            uint32_t ww[4] = {0u, 0u, 0u, 0u};
            if (part != 2) {
              const uint4 q4 = ld_shared_v4(op_addr16(b0, b1, urow, ch));
              ww[0] = q4.x; ww[1] = q4.y; ww[2] = q4.z; ww[3] = q4.w;
            }
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float x0v = __uint_as_float(ww[k] << 16), x1v = __uint_as_float(ww[k] & 0xffff0000u);
              o[k] = pack_bf16x2(mul * (acc[8 * ch + 2 * k] - x0v * rr), mul * (acc[8 * ch + 2 * k + 1] - x1v * rr));
            }
            w4[cc] = make_uint4(o[0], o[1], o[2], o[3]);
          }
          if (grp < 2) warp_store_rows<4>(g + grp * 64, opitch, opitch * p.gw, w4, scratch, lane);
          else warp_store_rows<3>(g + 128, opitch, opitch * p.gw, w4, scratch, lane);
        }
      }
      // ---- d scale: this warp's share of sum_i q_hat_i . dq~_i (only the column-half-0 warps hold any)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
      if (lane == 0) {
        p.ds_part[static_cast<size_t>(item) * 8 + (warp - 4)] = ds_acc;
        mbar_arrive(bar(EPI_DONE));                    // this warp no longer reads the item's operand tiles
      }
    }
  }
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

namespace {
typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encode encode_fn() {
  static PFN_encode fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encode>(ptr);
  }
  return fn;
}
// bf16 tensor viewed as [d3][y][x][d] with box [1][8][8][box_d]
int make_tmap4(CUtensorMap* out, const void* base, uint64_t d_extent, uint64_t gw, uint64_t rows_y, uint64_t d3, uint64_t pitch_tok,
               uint64_t pitch_d3, int box_d, CUtensorMapSwizzle swz) {
  PFN_encode fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return SWB_ERR_DRIVER;
  }
  cuuint64_t gdim[4] = {d_extent, gw, rows_y, d3};
  cuuint64_t gstr[3] = {pitch_tok * 2, gw * pitch_tok * 2, pitch_d3 * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), 8, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (attention backward, box_d=%d) failed (%d)", box_d, (int)r);
    return SWB_ERR_DRIVER;
  }
  return SWB_OK;
}
}  // namespace

// qkv: packed bf16 [3][heads][M][96]; O, dO: bf16 [M, heads*88]; L: fp32 [heads][M] (forward); invn fp32 [2][heads][M];
// Dbuf: fp32 [heads][M] scratch; ds_part: fp32 [items * 8]; dqkv: bf16 [M, 3*heads*88] (head, part, d).
int launch_attention_bwd_tc(const void* qkv, const void* O, const void* dO, const float* L, const float* invn, const float* qscale,
                            void* dqkv, float* Dbuf, float* ds_part, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                            cudaStream_t stream) {
  using namespace abt;
  SWB_REQUIRE(gh % 16 == 0 && gw % 16 == 0 && shift_h % 8 == 0 && shift_w % 8 == 0,
              "attention_bwd_tc: grid %dx%d / shift %d,%d unsupported", gh, gw, shift_h, shift_w);
  const int M = B * gh * gw, D = heads * kHd;
  SWB_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dO) | reinterpret_cast<uintptr_t>(O) |
                reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0, "attention_bwd_tc: buffers must be 16-byte aligned");
  attn_bwd_prep_kernel<<<static_cast<unsigned>((static_cast<long long>(M) * heads + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint16_t*>(O), static_cast<const uint16_t*>(dO), Dbuf, M, heads);
  SWB_CHECK_CUDA(cudaGetLastError());
  CUtensorMap tq128, tq64, tg128, tg64;
  int rc = make_tmap4(&tq128, qkv, kHdPad, gw, static_cast<uint64_t>(B) * gh, 3 * heads, kHdPad, static_cast<uint64_t>(M) * kHdPad, 64,
                      CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap4(&tq64, qkv, kHdPad, gw, static_cast<uint64_t>(B) * gh, 3 * heads, kHdPad, static_cast<uint64_t>(M) * kHdPad, 32,
                  CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_tmap4(&tg128, dO, D, gw, static_cast<uint64_t>(B) * gh, 1, D, static_cast<uint64_t>(M) * D, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap4(&tg64, dO, D, gw, static_cast<uint64_t>(B) * gh, 1, D, static_cast<uint64_t>(M) * D, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done.set(true);
  }
  AttnBwdTcParams p;
  p.B = B;
  p.gh = gh;
  p.gw = gw;
  p.heads = heads;
  p.M = M;
  p.shift_by = shift_h / 8;
  p.shift_bx = shift_w / 8;
  p.L = L;
  p.D = Dbuf;
  p.invn = invn;
  p.qscale = qscale;
  p.dqkv = static_cast<uint16_t*>(dqkv);
  p.ds_part = ds_part;
  const int items = B * (gh / 16) * (gw / 16) * heads;
  const int grid = items < num_sms() ? items : num_sms();
  attn_bwd_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tq128, tq64, tg128, tg64, p);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
