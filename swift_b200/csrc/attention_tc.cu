// Shifted-window scaled-cosine attention on tcgen05 / TMEM, fed by TMA  (reference: models/swinv2.py:118-135, :186-209).
//
// One work item = (sample, window, head): Q, K, V are 256 tokens x 88 (stored padded to 96) in the 16-bit operand
// format, already normalised / scaled by the qkv GEMM epilogue.  A persistent CTA per SM walks over items.
//
//   warp 0      TMA loader: the window's 256 tokens are four 8x8-token boxes of the [slot][y][x][d] view of the qkv buffer;
//               cyclic shift (a multiple of 8) and wrap-around are pure box-coordinate arithmetic, so the
//               roll / window_partition / window_reverse copies of the reference never exist.
//   warp 1      UMMA issuer:  S_h = Q_h K^T   (M=128 rows, N=256 keys, K=96: SWIZZLE_128B chunk d[0,64) + SWIZZLE_64B
//               chunk d[64,96)), accumulators S_0 / S_1 fill all 512 TMEM columns;
//               O_h = P_h V  with P read straight from TMEM (A operand in tensor memory) and V as an MN-major smem
//               operand (tokens x d is exactly how it sits in memory, no transpose anywhere).
//   warps 4-7   softmax group for rows 0..127, warps 8-11 for rows 128..255: one thread per query row; two passes over
//               its TMEM row (max, then exp2 / sum), P written back over S as packed 16-bit pairs (tcgen05.st), later
//               O_h is read out of TMEM, scaled by 1/rowsum and stored as the [M, heads*88] operand of the wo GEMM.
//
// TMEM columns: S_h at h*256 .. +255; P_h aliases S_h columns [0,128); O_h lives in S_h columns [128,224).
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace swb {

namespace atc {
constexpr int kThreads = 384;
constexpr int kTok = 256;
constexpr int kHd = 88;
constexpr int kHdPad = 96;
// shared memory map (offsets from a 1024-byte aligned base)
constexpr uint32_t kQ0 = 0;            // 256 x 128 B  d[0,64)   SWIZZLE_128B
constexpr uint32_t kQ1 = 32768;        // 256 x  64 B  d[64,96)  SWIZZLE_64B
constexpr uint32_t kK0 = 49152;
constexpr uint32_t kK1 = 81920;
constexpr uint32_t kV0 = 98304;        // 256 x 128 B  d[0,64)   SWIZZLE_128B
constexpr uint32_t kV1 = 131072;       // 256 x 128 B  d[64,128) SWIZZLE_128B (d >= 96 zero-filled by TMA)
constexpr uint32_t kScratch = 163840;  // 8 epilogue warps x 4 KB
constexpr uint32_t kBars = 196608;
constexpr uint32_t kSmemBytes = kBars + 256 + 1024;
constexpr uint32_t kQKBytes = 2 * (32768 + 16384);
constexpr uint32_t kVBytes = 65536;
enum Bar { QK_FULL = 0, QK_EMPTY, V_FULL, V_EMPTY, S_FULL0, S_FULL1, P_FULL0, P_FULL1, O_FULL0, O_FULL1, O_FREE0, O_FREE1, NBARS };
}  // namespace atc

// 2^x on the MUFU pipe without exp2f()'s denormal-range fix-up (arguments are <= 0; results below 2^-126 flush to 0)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnTcParams {
  int B, gh, gw, heads, M;
  int shift_by, shift_bx;     // cyclic shift in units of 8 tokens
  void* out;                  // [M, heads*88] 16-bit
  float* lse;                 // optional [heads][M]: log-sum-exp of every score row (what the backward needs), or null
};

// rows of a warp are 32 tokens = 4 token-grid lines of 8: row r sits at (r>>3)*pitch8 + (r&7)*pitch1 bytes
template <int NCH>
__device__ __forceinline__ void warp_store_rows_2p(uint8_t* g_row0, size_t pitch1, size_t pitch8, const uint4* v,
                                                   uint32_t scratch, int lane) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) st_shared_v4(scratch + lane * 128 + ((c ^ (lane & 7)) << 4), v[c]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / NCH, c = idx - r * NCH;
    const uint4 q = ld_shared_v4(scratch + r * 128 + ((c ^ (r & 7)) << 4));
    *reinterpret_cast<uint4*>(g_row0 + (r >> 3) * pitch8 + (r & 7) * pitch1 + c * 16) = q;
  }
  __syncwarp();
}

// F16: format of q / k / v and P (the tensor-core operands of this kernel); OUT_F16: format of the output rows (the A
// operand of the wo GEMM, i.e. the model's operand format).  (F16, !OUT_F16) is the bf16 model with fp16 attention
// internals: q_hat * scale <= 100, k_hat <= 1 and P <= 1 are bounded, so fp16's 8x finer rounding is free accuracy.
template <bool F16, bool OUT_F16>
__global__ void __launch_bounds__(atc::kThreads, 1)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap128, const __grid_constant__ CUtensorMap tmap64,
                           const AttnTcParams p) {
  using namespace atc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = sb + kBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * NBARS;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwx = p.gw / 16, nwin = (p.gh / 16) * nwx;
  const int nbx = p.gw / 8, nby = p.gh / 8;
  const int num_items = p.B * nwin * p.heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap128);
    tma_prefetch_desc(&tmap64);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NBARS; ++i) mbar_init(bar(i), (i >= P_FULL0 && i <= P_FULL1) || i >= O_FREE0 ? 4 : 1);
    fence_mbar_init_cluster();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================================== TMA loader ===========================================
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int head = item % p.heads;
      const int bw = item / p.heads;
      const int win = bw % nwin, b = bw / nwin;
      const int wy = win / nwx, wx = win - wy * nwx;
      mbar_wait(bar(QK_EMPTY), par ^ 1u, 11);
      mbar_arrive_expect_tx_elect(bar(QK_FULL), kQKBytes);
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        const int slot = part * p.heads + head;
        const uint32_t d0 = sb + (part ? kK0 : kQ0), d1 = sb + (part ? kK1 : kQ1);
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          const int x0 = ((2 * wx + (blk & 1) + p.shift_bx) % nbx) * 8;
          const int y0 = ((2 * wy + (blk >> 1) + p.shift_by) % nby) * 8 + b * p.gh;
          tma_load_4d_elect(d0 + blk * 8192, &tmap128, bar(QK_FULL), 0, x0, y0, slot);
          tma_load_4d_elect(d1 + blk * 4096, &tmap64, bar(QK_FULL), 64, x0, y0, slot);
        }
      }
      mbar_wait(bar(V_EMPTY), par ^ 1u, 12);
      mbar_arrive_expect_tx_elect(bar(V_FULL), kVBytes);
      {
        const int slot = 2 * p.heads + head;
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          const int x0 = ((2 * wx + (blk & 1) + p.shift_bx) % nbx) * 8;
          const int y0 = ((2 * wy + (blk >> 1) + p.shift_by) % nby) * 8 + b * p.gh;
          tma_load_4d_elect(sb + kV0 + blk * 8192, &tmap128, bar(V_FULL), 0, x0, y0, slot);
          tma_load_4d_elect(sb + kV1 + blk * 8192, &tmap128, bar(V_FULL), 64, x0, y0, slot);
        }
      }
    }
  } else if (warp == 1) {
    // =========================================== UMMA issuer ===========================================
    constexpr uint32_t idesc_s = make_idesc_f16(128, 256, F16, F16, 0, 0);       // S = Q K^T, both K-major
    constexpr uint32_t idesc_o = make_idesc_f16(128, kHdPad, F16, F16, 0, 1);    // O = P V, V is MN-major
    const uint64_t hi128 = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint64_t hi64 = make_smem_desc(0, 16, 512, SWZ_64B);
    const uint64_t hiV = make_smem_desc(0, kV1 - kV0, 1024, SWZ_128B);           // LBO = distance between the two d chunks
    auto lo = [](uint32_t addr) { return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4); };
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      mbar_wait(bar(QK_FULL), par, 21);
      tcgen05_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(bar(O_FREE0 + h), par ^ 1u, 22);         // previous item's O_h has been read out of TMEM
        tcgen05_fence_after();
        const uint32_t d = tmem_base + h * 256;
        const uint64_t a0 = hi128 | lo(sb + kQ0 + h * 16384), b0 = hi128 | lo(sb + kK0);
        const uint64_t a1 = hi64 | lo(sb + kQ1 + h * 8192), b1 = hi64 | lo(sb + kK1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss_elect<1>(d, a0 + 2u * k, b0 + 2u * k, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_ss_elect<1>(d, a1 + 2u * k, b1 + 2u * k, idesc_s, 1u);
        umma_commit_elect<1>(bar(S_FULL0 + h));
      }
      umma_commit_elect<1>(bar(QK_EMPTY));                 // Q, K smem reusable once both S MMAs retire
      mbar_wait(bar(V_FULL), par, 23);
      tcgen05_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(bar(P_FULL0 + h), par, 24);              // softmax group h has written P_h to TMEM
        tcgen05_fence_after();
        const uint32_t d = tmem_base + h * 256 + 128;
        const uint32_t a = tmem_base + h * 256;
        const uint64_t bv = hiV | lo(sb + kV0);
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)                    // 16 keys per step: 8 TMEM columns of P, 2 KB of V rows
          umma_f16_ts_elect(d, a + 8u * ks, bv + 128u * ks, idesc_o, ks != 0 ? 1u : 0u);
        umma_commit_elect<1>(bar(O_FULL0 + h));
      }
      umma_commit_elect<1>(bar(V_EMPTY));
    }
  } else if (warp >= 4) {
    // =========================================== softmax / output ===========================================
    const int h = (warp - 4) >> 2;                          // row half handled by this group
    const int quad = warp & 3;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + h * 256;
    const uint32_t scratch = sb + kScratch + (warp - 4) * 4096;
    constexpr float kLog2e = 1.4426950408889634f;
    const size_t opitch = static_cast<size_t>(p.heads) * kHd * 2;      // bytes per token row of `out`
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int head = item % p.heads;
      const int bw = item / p.heads;
      const int win = bw % nwin, b = bw / nwin;
      const int wy = win / nwx, wx = win - wy * nwx;
      mbar_wait(bar(S_FULL0 + h), par, 31);
      tcgen05_fence_after();
      // pass 1: row max (128 columns per tcgen05.wait::ld: the TMEM read latency is paid twice, not eight times)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float s[128];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_x32(trow + 128 * c + 32 * q, s + 32 * q);
        tmem_ld_wait();
        tmem_ld_fence_regs<128>(s);
#pragma unroll
        for (int j = 0; j < 128; ++j) mx = fmaxf(mx, s[j]);
      }
      // pass 2: p = exp(s - max); P (16-bit pairs) overwrites S columns [0,128) behind the read pointer
      const float mb = mx * kLog2e;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float s[64];
        __syncwarp();
        tmem_ld_x32(trow + 64 * c, s);
        tmem_ld_x32(trow + 64 * c + 32, s + 32);
        tmem_ld_wait();
        tmem_ld_fence_regs<64>(s);
        uint32_t w[32];
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float p0 = ex2_approx(fmaf(s[2 * j], kLog2e, -mb));      // one FFMA + one MUFU.EX2 per score
          const float p1 = ex2_approx(fmaf(s[2 * j + 1], kLog2e, -mb));
          part += p0 + p1;
          w[j] = pack_act2<F16>(p0, p1);
        }
        sum += part;
        tmem_st_x16(trow + 32 * c, w);
        tmem_st_x16(trow + 32 * c + 16, w + 16);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(P_FULL0 + h));
      if (p.lse != nullptr) {                                // training forward: L_i = max_i + ln(sum_j exp(S_ij - max_i))
        const int blk_ = 2 * h + (quad >> 1);
        const int xl = ((2 * wx + (blk_ & 1) + p.shift_bx) % nbx) * 8 + (lane & 7);
        const int yl = ((2 * wy + (blk_ >> 1) + p.shift_by) % nby) * 8 + 4 * (quad & 1) + (lane >> 3);
        p.lse[static_cast<size_t>(head) * p.M + (static_cast<size_t>(b) * p.gh + yl) * p.gw + xl] = mx + __logf(sum);
      }
      // O_h = P_h V
      mbar_wait(bar(O_FULL0 + h), par, 32);
      tcgen05_fence_after();
      float o[kHd];
      __syncwarp();
      tmem_ld_x32(trow + 128, o);
      tmem_ld_x32(trow + 160, o + 32);
      tmem_ld_x16(trow + 192, o + 64);
      tmem_ld_x8(trow + 208, o + 80);
      tmem_ld_wait();
      tmem_ld_fence_regs<kHd>(o);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(O_FREE0 + h));         // S_h / P_h / O_h columns may be overwritten by the next item
      const float inv = 1.0f / sum;
      uint32_t w[44];
#pragma unroll
      for (int j = 0; j < 44; ++j) w[j] = pack_act2<OUT_F16>(o[2 * j] * inv, o[2 * j + 1] * inv);
      // this warp's 32 rows: block blk = 2h + quad/2 (8x8 tokens), lines 4*(quad&1) .. +3
      const int blk = 2 * h + (quad >> 1);
      const int x0 = ((2 * wx + (blk & 1) + p.shift_bx) % nbx) * 8;
      const int y0 = ((2 * wy + (blk >> 1) + p.shift_by) % nby) * 8 + 4 * (quad & 1);
      const size_t row0 = (static_cast<size_t>(b) * p.gh + y0) * p.gw + x0;
      uint8_t* g = static_cast<uint8_t*>(p.out) + row0 * opitch + static_cast<size_t>(head) * kHd * 2;
      warp_store_rows_2p<8>(g, opitch, opitch * p.gw, reinterpret_cast<const uint4*>(w), scratch, lane);
      warp_store_rows_2p<3>(g + 128, opitch, opitch * p.gw, reinterpret_cast<const uint4*>(w + 32), scratch, lane);
    }
  }
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// 4-D view of the packed qkv buffer: [slot = part*heads + head][y (all samples)][x][d] with a [1][8][8][box_d] box
static int make_tmap_qkv(CUtensorMap* out, const void* qkv, bool f16, int heads, int B, int gh, int gw, int box_d,
                         CUtensorMapSwizzle swz) {
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static PFN fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled driver entry point not available");
      return SWB_ERR_DRIVER;
    }
    fn = reinterpret_cast<PFN>(ptr);
  }
  const cuuint64_t M = static_cast<cuuint64_t>(B) * gh * gw;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(atc::kHdPad), static_cast<cuuint64_t>(gw),
                        static_cast<cuuint64_t>(B) * gh, static_cast<cuuint64_t>(3 * heads)};
  cuuint64_t gstr[3] = {atc::kHdPad * 2, static_cast<cuuint64_t>(gw) * atc::kHdPad * 2, M * atc::kHdPad * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), 8, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                  const_cast<void*>(qkv), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (qkv 4-D, box_d=%d) failed (%d)", box_d, (int)r);
    return SWB_ERR_DRIVER;
  }
  return SWB_OK;
}

int launch_window_attention_tc(const void* qkv, void* out, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                               int act_f16, int out_f16, cudaStream_t stream, float* lse) {
  using namespace atc;
  SWB_REQUIRE(act_f16 || !out_f16, "window_attention_tc: bf16 q/k/v with fp16 output is not a supported combination");
  SWB_REQUIRE(gh % 16 == 0 && gw % 16 == 0 && shift_h % 8 == 0 && shift_w % 8 == 0,
              "window_attention_tc: grid %dx%d / shift %d,%d unsupported", gh, gw, shift_h, shift_w);
  SWB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "window_attention_tc: buffers must be 16-byte aligned");
  CUtensorMap t128, t64;
  int rc = make_tmap_qkv(&t128, qkv, act_f16 != 0, heads, B, gh, gw, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_qkv(&t64, qkv, act_f16 != 0, heads, B, gh, gw, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<true, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<true, false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    SWB_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<false, false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done.set(true);
  }
  AttnTcParams p;
  p.B = B;
  p.gh = gh;
  p.gw = gw;
  p.heads = heads;
  p.M = B * gh * gw;
  p.shift_by = shift_h / 8;
  p.shift_bx = shift_w / 8;
  p.out = out;
  p.lse = lse;
  const int items = B * (gh / 16) * (gw / 16) * heads;
  const int grid = items < num_sms() ? items : num_sms();
  if (act_f16 && out_f16)
    window_attention_tc_kernel<true, true><<<grid, kThreads, kSmemBytes, stream>>>(t128, t64, p);
  else if (act_f16)
    window_attention_tc_kernel<true, false><<<grid, kThreads, kSmemBytes, stream>>>(t128, t64, p);
  else
    window_attention_tc_kernel<false, false><<<grid, kThreads, kSmemBytes, stream>>>(t128, t64, p);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
