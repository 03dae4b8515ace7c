// Forward-mode tangent of the shifted-window scaled-cosine attention on tcgen05 / TMEM, fed by TMA: what
// `torch.func.jvp` differentiates in models/swinv2.py:129-134 (explicit softmax branch) for the sCM loss
// (training/loss.py:216-225).  Same dual-number algebra as the mma.sync kernel of tangent.cu (kept for shifts that are not
// multiples of 8):
//     S = q k^T,  dS = dq k^T + q dk^T,  p~ = exp(S - max),  l = sum p~,  c = (sum p~ dS) / l,
//     O = (p~ v) / l,     dO = ((p~ dS) v + p~ dv) / l - c O
//
// One work item = (sample, window, head); per query half h (128 rows) one unit:
//   UMMA   S_h = q_h k^T into TMEM columns [0,256);  dS_h = dq_h k^T + q_h dk^T accumulated into [256,512)
//   warps  two threads per row (column halves), row max / l / sum p~ dS exchanged through shared memory; p~ and E = p~ dS are
//          written back as packed 16-bit pairs behind the read pointer (TS-UMMA A operands)
//   UMMA   O~ = p~ v;   T~ = E v + p~ dv   (v, dv MN-major from the same tiles the TMA wrote)
//   warps  column half 0: O = O~ / l -> primal rows;  half 1: dO = T~ / l - c O~ / l -> tangent rows of attn2 [2M, heads*88]
// Shared memory: q_h / dq_h double buffered by half (2 x 48 KB); ONE 96 KB buffer holds (k, dk) for the first stage and is
// refilled with (v, dv) while the warps do the softmax -- six 48 KB operands do not fit next to each other.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace swb {

namespace adt {
constexpr int kThreads = 384;
constexpr int kHd = 88;
constexpr int kHdPad = 96;
constexpr uint32_t kHalf = 24576;                                   // 128 rows: 128 x 128 B + 128 x 64 B
constexpr uint32_t kQD = 0;                                         // [half][q | dq] x 24 KB
constexpr uint32_t kKV = 4 * kHalf;                                 // op0 (k or v) 48 KB | op1 (dk or dv) 48 KB
constexpr uint32_t kOp = 49152;
constexpr uint32_t kXch = kKV + 2 * kOp;                            // exchange: 3 arrays x [2][128] fp32
constexpr uint32_t kScratch = kXch + 4096;                          // 8 warps x 2 KB
constexpr uint32_t kBars = kScratch + 8 * 2048;
constexpr uint32_t kSmemBytes = kBars + 256 + 1024;
enum Bar { QD_FULL0 = 0, QD_FULL1, QD_EMPTY0, QD_EMPTY1, KK_FULL, KK_EMPTY, VV_FULL, VV_EMPTY, S_FULL, P_FULL, ACC_FULL, ACC_FREE, NBARS };
}  // namespace adt

struct AttnDualTcParams {
  int B, gh, gw, heads, M;
  int shift_by, shift_bx;
  uint16_t* attn2;               // [2M, heads*88]
};

namespace {
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_pair(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int NCH>
__device__ __forceinline__ void dual_store_rows(uint8_t* g_row0, size_t pitch1, size_t pitch8, const uint4* v, uint32_t scratch, int lane) {
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

template <bool F16>
__global__ void __launch_bounds__(adt::kThreads, 1)
attn_dual_tc_kernel(const __grid_constant__ CUtensorMap tp128, const __grid_constant__ CUtensorMap tp64,
                    const __grid_constant__ CUtensorMap tt128, const __grid_constant__ CUtensorMap tt64, const AttnDualTcParams p) {
  using namespace adt;
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
    tma_prefetch_desc(&tp128);
    tma_prefetch_desc(&tp64);
    tma_prefetch_desc(&tt128);
    tma_prefetch_desc(&tt64);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NBARS; ++i) mbar_init(bar(i), (i == P_FULL || i == ACC_FREE) ? 8 : 1);
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

  auto box_xy = [&](int item, int blk, int& x0, int& y0) {
    const int bw = item / p.heads;
    const int win = bw % nwin, b = bw / nwin;
    const int wy = win / nwx, wx = win - wy * nwx;
    x0 = ((2 * wx + (blk & 1) + p.shift_bx) % nbx) * 8;
    y0 = ((2 * wy + (blk >> 1) + p.shift_by) % nby) * 8 + b * p.gh;
  };

  if (warp == 0) {
    // =========================================== TMA loader ===========================================
    // rows [r0, r0 + 64 nblk) of (part, head) of the primal (tangent = 1: dqkv) tensor -> dst (SW128 chunk | SW64 chunk at +c1)
    auto load_rows = [&](int item, int tangent, int part, int blk0, int nblk, uint32_t dst, uint32_t c1_off, uint32_t full_bar) {
      const int slot = part * p.heads + item % p.heads;
      for (int i = 0; i < nblk; ++i) {
        int x0, y0;
        box_xy(item, blk0 + i, x0, y0);
        tma_load_4d_elect(dst + i * 8192, tangent ? &tt128 : &tp128, full_bar, 0, x0, y0, slot);
        tma_load_4d_elect(dst + c1_off + i * 4096, tangent ? &tt64 : &tp64, full_bar, 64, x0, y0, slot);
      }
    };
    int it = 0;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h, ++n) {
        // q_h, dq_h
        mbar_wait(bar(QD_EMPTY0 + h), (it & 1) ^ 1u, 11);
        mbar_arrive_expect_tx_elect(bar(QD_FULL0 + h), 2 * kHalf);
        load_rows(item, 0, 0, 2 * h, 2, sb + kQD + (2 * h) * kHalf, 16384, bar(QD_FULL0 + h));
        load_rows(item, 1, 0, 2 * h, 2, sb + kQD + (2 * h + 1) * kHalf, 16384, bar(QD_FULL0 + h));
        // k, dk into the shared buffer (free once the previous unit's second-stage MMAs have retired)
        mbar_wait(bar(VV_EMPTY), (n & 1u) ^ 1u, 12);
        mbar_arrive_expect_tx_elect(bar(KK_FULL), 2 * kOp);
        load_rows(item, 0, 1, 0, 4, sb + kKV, 32768, bar(KK_FULL));
        load_rows(item, 1, 1, 0, 4, sb + kKV + kOp, 32768, bar(KK_FULL));
        // v, dv over them as soon as the first-stage MMAs have retired
        mbar_wait(bar(KK_EMPTY), n & 1u, 13);
        mbar_arrive_expect_tx_elect(bar(VV_FULL), 2 * kOp);
        load_rows(item, 0, 2, 0, 4, sb + kKV, 32768, bar(VV_FULL));
        load_rows(item, 1, 2, 0, 4, sb + kKV + kOp, 32768, bar(VV_FULL));
      }
    }
  } else if (warp == 1) {
    // =========================================== UMMA issuer ===========================================
    constexpr uint32_t idesc_big = make_idesc_f16(128, 256, F16, F16, 0, 0);
    constexpr uint32_t idesc_n64 = make_idesc_f16(128, 64, F16, F16, 0, 1);
    constexpr uint32_t idesc_n32 = make_idesc_f16(128, 32, F16, F16, 0, 1);
    const uint64_t hi128 = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint64_t hi64 = make_smem_desc(0, 16, 512, SWZ_64B);
    auto lo = [](uint32_t addr) { return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4); };
    // D = A(128-row tile at a_base: SW128 chunk | SW64 chunk at +16384) * B(256-row operand at b_base: chunks at +0 | +32768)^T
    auto product = [&](uint32_t d, uint32_t a_base, uint32_t b_base, uint32_t acc_first) {
      const uint64_t a0 = hi128 | lo(sb + a_base), b0 = hi128 | lo(sb + b_base);
      const uint64_t a1 = hi64 | lo(sb + a_base + 16384), b1 = hi64 | lo(sb + b_base + 32768);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss_elect<1>(d, a0 + 2u * k, b0 + 2u * k, idesc_big, (k != 0 || acc_first) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_f16_ss_elect<1>(d, a1 + 2u * k, b1 + 2u * k, idesc_big, 1u);
    };
    auto from_tmem = [&](uint32_t d64, uint32_t d32, uint32_t a_lo, uint32_t a_hi, uint32_t b_base, uint32_t acc_first) {
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t a = (ks < 8 ? a_lo + 8u * ks : a_hi + 8u * (ks - 8));
        const uint64_t b0 = hi128 | lo(sb + b_base + 2048u * ks);
        const uint64_t b1 = hi64 | lo(sb + b_base + 32768u + 1024u * ks);
        umma_f16_ts_elect(d64, a, b0, idesc_n64, (ks != 0 || acc_first) ? 1u : 0u);
        umma_f16_ts_elect(d32, a, b1, idesc_n32, (ks != 0 || acc_first) ? 1u : 0u);
      }
    };
    int it = 0;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h, ++n) {
        const uint32_t qh = kQD + (2 * h) * kHalf, dqh = kQD + (2 * h + 1) * kHalf;
        mbar_wait(bar(QD_FULL0 + h), it & 1, 21);
        mbar_wait(bar(KK_FULL), n & 1u, 22);
        mbar_wait(bar(ACC_FREE), (n & 1u) ^ 1u, 23);
        tcgen05_fence_after();
        product(tmem, qh, kKV, 0u);                    // S  = q_h k^T
        product(tmem + 256, dqh, kKV, 0u);             // dS = dq_h k^T
        product(tmem + 256, qh, kKV + kOp, 1u);        //    + q_h dk^T
        umma_commit_elect<1>(bar(S_FULL));
        umma_commit_elect<1>(bar(KK_EMPTY));
        umma_commit_elect<1>(bar(QD_EMPTY0 + h));
        mbar_wait(bar(VV_FULL), n & 1u, 24);
        mbar_wait(bar(P_FULL), n & 1u, 25);
        tcgen05_fence_after();
        from_tmem(tmem + 64, tmem + 192, tmem, tmem + 128, kKV, 0u);                    // O~ = p~ v
        from_tmem(tmem + 320, tmem + 448, tmem + 256, tmem + 384, kKV, 0u);             // T~ = E v
        from_tmem(tmem + 320, tmem + 448, tmem, tmem + 128, kKV + kOp, 1u);             //    + p~ dv
        umma_commit_elect<1>(bar(ACC_FULL));
        umma_commit_elect<1>(bar(VV_EMPTY));
      }
    }
  } else if (warp >= 4) {
    // =========================================== softmax / output ===========================================
    const int quad = warp & 3;
    const int c = (warp - 4) >> 2;
    const int lrow = quad * 32 + lane;
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t scratch = sb + kScratch + (warp - 4) * 2048;
    float* xmax = reinterpret_cast<float*>(sb_ptr + kXch);           // [2][128]
    float* xsum = xmax + 256;
    float* xe = xsum + 256;
    constexpr float kLog2e = 1.4426950408889634f;
    const int Dm = p.heads * kHd;
    const size_t opitch = static_cast<size_t>(Dm) * 2;
    int it = 0;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int head = item % p.heads;
#pragma unroll 1
      for (int h = 0; h < 2; ++h, ++n) {
        mbar_wait(bar(S_FULL), n & 1u, 31);
        tcgen05_fence_after();
        const uint32_t t_s = tl + c * 128, t_ds = tl + 256 + c * 128;
        // pass 1: maximum of this thread's 128 scores, then of the row
        float mx = -INFINITY;
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          float s[64];
          __syncwarp();
          tmem_ld_x32(t_s + 64 * i, s);
          tmem_ld_x32(t_s + 64 * i + 32, s + 32);
          tmem_ld_wait();
          tmem_ld_fence_regs<64>(s);
#pragma unroll
          for (int j = 0; j < 64; ++j) mx = fmaxf(mx, s[j]);
        }
        xmax[c * 128 + lrow] = mx;
        bar_pair(2 + quad);
        mx = fmaxf(mx, xmax[(c ^ 1) * 128 + lrow]);
        const float mb = mx * kLog2e;
        // pass 2: p~, E = p~ dS, their sums; packed pairs behind the read pointer
        float lsum = 0.f, esum = 0.f;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          float s[32], ds[32];
          __syncwarp();
          tmem_ld_x32(t_s + 32 * i, s);
          tmem_ld_x32(t_ds + 32 * i, ds);
          tmem_ld_wait();
          tmem_ld_fence_regs<32>(s);
          tmem_ld_fence_regs<32>(ds);
          uint32_t wp[16], we[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = ex2a(fmaf(s[2 * j], kLog2e, -mb)), p1 = ex2a(fmaf(s[2 * j + 1], kLog2e, -mb));
            const float e0 = p0 * ds[2 * j], e1 = p1 * ds[2 * j + 1];
            lsum += p0 + p1;
            esum += e0 + e1;
            wp[j] = pack_act2<F16>(p0, p1);
            we[j] = pack_act2<F16>(e0, e1);
          }
          tmem_st_x16(t_s + 16 * i, wp);
          tmem_st_x16(t_ds + 16 * i, we);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(P_FULL));
        xsum[c * 128 + lrow] = lsum;
        xe[c * 128 + lrow] = esum;
        bar_pair(2 + quad);
        lsum += xsum[(c ^ 1) * 128 + lrow];
        esum += xe[(c ^ 1) * 128 + lrow];
        const float inv = 1.0f / lsum;
        const float ctot = esum * inv;
        // ---- outputs
        mbar_wait(bar(ACC_FULL), n & 1u, 32);
        tcgen05_fence_after();
        // this warp's 32 rows: tokens of block blk, lines 4*(quad&1) .. +3
        const int urow = h * 128 + lrow;
        const int blk = urow >> 6;
        int x0, y0;
        box_xy(item, blk, x0, y0);
        const size_t grow0 = (static_cast<size_t>(y0) + 4 * (quad & 1)) * p.gw + x0;
        uint8_t* g = reinterpret_cast<uint8_t*>(p.attn2) + (grow0 + (c ? static_cast<size_t>(p.M) : 0)) * opitch +
                     static_cast<size_t>(head) * kHd * 2;
        uint32_t w[44];
        if (c == 0) {                                  // O = O~ / l
          float o[kHd];
          __syncwarp();
          tmem_ld_x32(tl + 64, o);
          tmem_ld_x32(tl + 96, o + 32);
          tmem_ld_x16(tl + 192, o + 64);
          tmem_ld_x8(tl + 208, o + 80);
          tmem_ld_wait();
          tmem_ld_fence_regs<kHd>(o);
#pragma unroll
          for (int j = 0; j < 44; ++j) w[j] = pack_act2<F16>(o[2 * j] * inv, o[2 * j + 1] * inv);
        } else {                                       // dO = T~ / l - c O~ / l, 32 + 32 + 24 columns at a time
          const float ci = ctot * inv;
          {
            float tv[32], ov[32];
            __syncwarp();
            tmem_ld_x32(tl + 320, tv);
            tmem_ld_x32(tl + 64, ov);
            tmem_ld_wait();
            tmem_ld_fence_regs<32>(tv);
            tmem_ld_fence_regs<32>(ov);
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_act2<F16>(fmaf(-ci, ov[2 * j], tv[2 * j] * inv), fmaf(-ci, ov[2 * j + 1], tv[2 * j + 1] * inv));
          }
          {
            float tv[32], ov[32];
            __syncwarp();
            tmem_ld_x32(tl + 352, tv);
            tmem_ld_x32(tl + 96, ov);
            tmem_ld_wait();
            tmem_ld_fence_regs<32>(tv);
            tmem_ld_fence_regs<32>(ov);
#pragma unroll
            for (int j = 0; j < 16; ++j) w[16 + j] = pack_act2<F16>(fmaf(-ci, ov[2 * j], tv[2 * j] * inv), fmaf(-ci, ov[2 * j + 1], tv[2 * j + 1] * inv));
          }
          {
            float tv[24], ov[24];
            __syncwarp();
            tmem_ld_x16(tl + 448, tv);
            tmem_ld_x8(tl + 464, tv + 16);
            tmem_ld_x16(tl + 192, ov);
            tmem_ld_x8(tl + 208, ov + 16);
            tmem_ld_wait();
            tmem_ld_fence_regs<24>(tv);
            tmem_ld_fence_regs<24>(ov);
#pragma unroll
            for (int j = 0; j < 12; ++j) w[32 + j] = pack_act2<F16>(fmaf(-ci, ov[2 * j], tv[2 * j] * inv), fmaf(-ci, ov[2 * j + 1], tv[2 * j + 1] * inv));
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(ACC_FREE));
        const uint4* w4 = reinterpret_cast<const uint4*>(w);
        dual_store_rows<4>(g, opitch, opitch * p.gw, w4, scratch, lane);
        dual_store_rows<4>(g + 64, opitch, opitch * p.gw, w4 + 4, scratch, lane);
        dual_store_rows<3>(g + 128, opitch, opitch * p.gw, w4 + 8, scratch, lane);
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
typedef CUresult (*PFN_encode2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_tmap_packed(CUtensorMap* out, const void* base, bool f16, int heads, int B, int gh, int gw, int box_d, CUtensorMapSwizzle swz) {
  static PFN_encode2 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled driver entry point not available");
      return SWB_ERR_DRIVER;
    }
    fn = reinterpret_cast<PFN_encode2>(ptr);
  }
  const cuuint64_t M = static_cast<cuuint64_t>(B) * gh * gw;
  cuuint64_t gdim[4] = {adt::kHdPad, static_cast<cuuint64_t>(gw), static_cast<cuuint64_t>(B) * gh, static_cast<cuuint64_t>(3 * heads)};
  cuuint64_t gstr[3] = {adt::kHdPad * 2, static_cast<cuuint64_t>(gw) * adt::kHdPad * 2, M * adt::kHdPad * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), 8, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (dual attention, box_d=%d) failed (%d)", box_d, (int)r);
    return SWB_ERR_DRIVER;
  }
  return SWB_OK;
}
}  // namespace

// qkv / dqkv: packed 16-bit [3][heads][M][96] (primal / tangent); attn2: 16-bit [2M, heads*88] (primal rows, then tangent rows)
int launch_attention_dual_tc(const void* qkv, const void* dqkv, void* attn2, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                             int act_f16, cudaStream_t stream) {
  using namespace adt;
  SWB_REQUIRE(gh % 16 == 0 && gw % 16 == 0 && shift_h % 8 == 0 && shift_w % 8 == 0,
              "attention_dual_tc: grid %dx%d / shift %d,%d unsupported", gh, gw, shift_h, shift_w);
  CUtensorMap tp128, tp64, tt128, tt64;
  int rc;
  if ((rc = make_tmap_packed(&tp128, qkv, act_f16 != 0, heads, B, gh, gw, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_packed(&tp64, qkv, act_f16 != 0, heads, B, gh, gw, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_packed(&tt128, dqkv, act_f16 != 0, heads, B, gh, gw, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_packed(&tt64, dqkv, act_f16 != 0, heads, B, gh, gw, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_dual_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(attn_dual_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done.set(true);
  }
  AttnDualTcParams p;
  p.B = B;
  p.gh = gh;
  p.gw = gw;
  p.heads = heads;
  p.M = B * gh * gw;
  p.shift_by = shift_h / 8;
  p.shift_bx = shift_w / 8;
  p.attn2 = static_cast<uint16_t*>(attn2);
  const int items = B * (gh / 16) * (gw / 16) * heads;
  const int grid = items < num_sms() ? items : num_sms();
  if (act_f16) attn_dual_tc_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(tp128, tp64, tt128, tt64, p);
  else attn_dual_tc_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(tp128, tp64, tt128, tt64, p);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
