// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T
// (A and W both fp16 or both bf16, fp32 accumulate in TMEM) with the epilogues the Swift denoiser needs fused in.
// A and W are both K-major (row-major [rows, K]), which is how activations and nn.Linear weights are stored, so
// neither is ever transposed in memory.
//
//   CG = 2 (default): the two CTAs of a cluster form one MMA pair (tcgen05 cta_group::2).  The pair owns a
//        256 x (176*NSUB) output tile: each CTA stages its own 128 rows of A and half of the W tile, the leader CTA
//        issues UMMA 256x176x16 and accumulator rows 0..127 / 128..255 land in each CTA's own TMEM.
//   CG = 1: single-CTA fallback (UMMA M=128, NSUB=1), same pipeline, used for bring-up and as a cross-check.
//
//   NSUB = 1: 256x176 tile, two TMEM accumulator stages (epilogue of tile i overlaps the main loop of tile i+1).
//   NSUB = 2: 256x352 tile = two UMMA N=176 sub-tiles sharing one A stage: 30 % fewer L2->SMEM bytes per FLOP (the
//        176-wide tile is L2-bandwidth bound on B200).  One accumulator set (2 x 176 of the 512 TMEM columns); each
//        sub-tile is drained by its own group of 4 epilogue warps and handed back to the issuer separately.
//
//   warp 0    TMA producer  (whole warp converged, one elected lane issues; SWIZZLE_128B tiles, mbarrier ring)
//   warp 1    UMMA issuer   (leader CTA; tcgen05.commit releases smem stages / publishes accumulators)
//   warp 2    TMEM allocator
//   warp 3    idle
//   warp 4..  epilogue: tcgen05.ld -> registers -> fused math -> per-warp smem transpose -> full-line global stores
#pragma once
#include "ptx.cuh"

#ifndef SWB_A_TMEM
#define SWB_A_TMEM 0     // 1: the 256x352 tile stages its A operand in tensor memory (tcgen05.cp + TMEM-A UMMA)
#endif

namespace swb {

enum GemmEpilogue : int {
  EPI_STORE_F32 = 0,   // out0[M, ldo] fp32
  EPI_STORE_ACT = 1,   // out0[M, ldo] in the 16-bit operand format (fp16 / bf16)
  EPI_EMBED = 2,       // x = acc + bias[n] + pos[row % pos_rows, n];  out0[M, 2N] = [hi | lo] 16-bit pair of x
  EPI_QKV = 3,         // scaled-cosine q/k normalisation fused; out0 = [3][heads][M][96] 16-bit
  EPI_SWIGLU = 4,      // tile = [gate slots | up slots];  out0[M, N/2] 16-bit = silu(gate) * up
  EPI_HEAD = 5,        // pixel-shuffle to NCHW + sampler update: y = alpha*xt + beta*F + gamma*fprev
  EPI_LN_RES = 10,     // x += LayerNorm(acc) * gain[b] + bias[b] on the residual pair xhl (row statistics exchanged between
                       // the CTAs that hold the other column tiles of the same rows)
  EPI_LN_RES1 = 12,    // EPI_LN_RES on the single-value residual stream (x = the hi half of xhl alone): chosen by launch_gemm from
                       // bit 1 of the format word; a compile-time variant because both paths in one kernel spill 130 registers
  EPI_DISCARD = 6,     // profiling: accumulators handed back unread (main-loop rate)
  EPI_DRAIN = 7,       // profiling: accumulators read out of TMEM, nothing stored (main loop + drain)
  EPI_SMEM_ONLY = 8,   // profiling: EPI_STORE_ACT without its global stores (drain + pack + smem transposes)
  EPI_DIRECT = 9,      // profiling: EPI_STORE_ACT with per-thread 16-byte global stores (no smem transpose)
  EPI_BUSY = 11,       // profiling: accumulators discarded, then `heads` x 64 dependent FMAs per epilogue thread and tile
                       // (no memory traffic at all: does ALU work of the epilogue warps slow the UMMA issuer down?)
};

struct GemmParams {
  int M, N, K;
  void* out0;
  void* out1;
  int ldo;                 // row pitch (elements) of out0/out1 for the plain / embed / swiglu epilogues
  // split-K (EPI_STORE_F32 only; the weight-gradient GEMMs contract over tokens and have few output tiles): the K
  // extent of A and W is splits * K, split s reads columns [s*K, (s+1)*K) and stores its partial product at rows
  // [s*M, (s+1)*M) of out0 ([splits * M, ldo]); K % 64 == 0 when splits > 1.  0 / 1 = no split.
  int splits;
  // batched GEMM (EPI_STORE_F32 only; the Newton-Schulz chains of same-shaped matrices): A is [batch * M, K], W is
  // [batch * N, K], problem b multiplies rows [b*M, (b+1)*M) of A with rows [b*N, (b+1)*N) of W and stores at rows
  // [(b * splits + s) * M, ...) of out0.  Rows of a tile beyond M / N belong to the next problem: they only feed
  // accumulator rows / columns that the store masks.  0 / 1 = one problem.
  int batch;
  int tma_store;           // EPI_STORE_ACT / EPI_QKV / EPI_SWIGLU: write the 16-bit output with TMA tile stores (tmap_o0/o1)
  // EPI_EMBED
  const float* bias;       // [N]
  const float* pos;        // [pos_rows, N]
  int pos_rows;
  // EPI_QKV
  const float* qscale;     // [heads]  exp(min(scale, ln 100))
  int heads, dmodel;       // N == 3 * dmodel
  int qkv_f16;             // format of the packed q / k / v output (may be fp16 while the GEMM operands are bf16)
  // EPI_HEAD
  const float* xt;         // [B, C, H, W] or null
  const float* fprev;      // [B, C, H, W] or null
  float* out_f;            // raw network output F (optional)
  float alpha, beta, gamma;
  int C, H, W, p1, p2, gw, tokens;   // tokens = gh*gw per sample
  // EPI_HEAD, rollout mode (generate.py:120-131 folded in): state != null
  float* state;            // [B, state_C, H, W] standardised state (first C channels updated in place)
  int state_C;
  const float* x_std;      // [C] per-channel sigma_x, mu_x, sigma_diff
  const float* x_mean;
  const float* d_std;
  float* phys;             // [B, C, H, W] physical-space state (optional)
  int zero_channel;        // channel forced to 0 (era5.py zero_field) or -1
  // EPI_LN_RES (N == model dim; `tokens` rows per sample)
  uint16_t* xhl;           // [M, 2N] residual stream as a 16-bit [hi | lo] pair, updated in place
  const float* gain;       // [B, N]  gamma * (1 + scale(t))
  const float* lnbias;     // [B, N]  beta * (1 + scale(t)) + shift(t)
  float2* ln_stats;        // [tiles_n * NSUB][ln_stride]  per-row partial (sum, M2) of each 176-column group
  unsigned ln_tag;         // 1..7: generation tag of this launch, carried in the three low mantissa bits of every published M2
  int ln_stride;
  float ln_eps;
  int ln_debug;            // profiling only (SWB_LN_DEBUG): bit 0 = do not wait for the other groups, bit 1 = skip the x update
  // 256x352 tile only: sub-tile 0 runs `skew` k-blocks ahead of sub-tile 1 at both ends of a tile, so its accumulator is
  // published (and drained, and handed back) while the tensor pipe still works on sub-tile 1 -- the hand-over latency of
  // the single accumulator set hides behind the other sub-tile's MMAs.  0 = both sub-tiles in lock step.
  int skew;
  // EPI_EMBED / EPI_LN_RES: the residual stream is ONE 16-bit value per element (the `hi` half of xhl, still at row pitch
  // 2N); the `lo` half is neither read nor written.  fp16 forecast path only (swb200_model.x_single).
  int x_single;
  unsigned long long* prof;   // cycle counters of a -DSWB_PROFILE_EPILOGUES build (null otherwise)
};

constexpr int kBlockM = 128;     // rows of A per CTA
constexpr int kBlockK = 64;      // 16-bit elements per k-block: 128 bytes = one SWIZZLE_128B atom row
constexpr int kUmmaK = 16;
constexpr int kUmmaN = 176;      // UMMA N: 1056 = 6*176, 3168 = 18*176, 5632 = 32*176
constexpr int kSlot = 88;        // a tile is made of 88-column slots: one (q|k|v, head) / one gate or up block
constexpr int kSubStride = 256;  // TMEM columns between accumulators
constexpr int kTmemCols = 512;
constexpr int kHeadDim = 88;     // Swift-B head dim
constexpr int kHeadDimPad = 96;  // q/k/v rows are stored padded to 96 (K of QK^T must be a multiple of 16)

#ifndef SWB_GEMM_MAX_STAGES
#define SWB_GEMM_MAX_STAGES 8
#endif

// LNSTATS: the fused-LayerNorm epilogue keeps 32 x (rstd, shift) per epilogue warp in shared memory; without it the
// 256x352 tile has room for a fifth pipeline stage.
template <int NSUB, int CG, bool LNSTATS = false>
struct GemmCfg {
  static_assert(NSUB == 1 || (NSUB == 2 && CG == 2), "the 352-wide tile needs the CTA pair");
  static constexpr int kTileN = kUmmaN * NSUB;
  static constexpr int kBRows = kTileN / CG;                   // W rows staged per CTA (one TMA box)
  static constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarps = 4 * NSUB;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  // per-warp 32 x 128 B transpose buffer; the fused-LayerNorm epilogue has two (both 64-column parts of the branch are parked
  // there across the statistics exchange) + 32 x (rstd, shift)
  static constexpr int kScratchPerWarp = LNSTATS ? 8192 : 4096;
  static constexpr int kScratchBytes = kEpiWarps * (kScratchPerWarp + (LNSTATS ? 256 : 0));
  static constexpr int kBarBytes = 1024;
  static constexpr int kMaxSmem = 227 * 1024;
  static constexpr int kStagesRaw = (kMaxSmem - kScratchBytes - kBarBytes - 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > SWB_GEMM_MAX_STAGES ? SWB_GEMM_MAX_STAGES : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + kScratchBytes + kBarBytes + 1024;  // +1024 alignment slack
  static_assert(kBBytes % 1024 == 0 && (kUmmaN / CG) * 128 % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
  static_assert(kStages >= 3, "pipeline too shallow");
};

// ---------------------------------------------------------------------------------------------------------
// epilogue helpers

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// silu(g) * u = g u / (1 + 2^(-g log2 e)): 3 FMUL + FADD + MUFU.EX2 + MUFU.RCP (the __fdividef / __expf form spends 9
// instructions per element on range fix-ups this argument range does not need; at the power cap instructions are time)
__device__ __forceinline__ float swiglu_f(float g, float u) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(g * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return (g * u) * r;
}

// Each lane holds NCH 16-byte chunks of ITS OWN row (lane = row of a 32-row block).  Writing them straight to global
// would touch 32 different 128-byte lines per instruction with 16 useful bytes each; instead the warp transposes
// through a 4 KB XOR-swizzled smem buffer (4 wavefronts per instruction both ways = the minimum for 512 bytes) and
// writes NCH*16 contiguous bytes per row with all lanes cooperating on consecutive chunks.
template <int NCH>
__device__ __forceinline__ void warp_store_rows(uint8_t* g_row0, size_t pitch_bytes, const uint4* v, uint32_t scratch,
                                                int lane, int rows_valid, int chunks_valid) {
  static_assert(NCH >= 1 && NCH <= 8, "one 128-byte smem row per lane");
#pragma unroll
  for (int c = 0; c < NCH; ++c) st_shared_v4(scratch + lane * 128 + ((c ^ (lane & 7)) << 4), v[c]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / NCH, c = idx - r * NCH;
    const uint4 q = ld_shared_v4(scratch + r * 128 + ((c ^ (r & 7)) << 4));
    if (r < rows_valid && c < chunks_valid) *reinterpret_cast<uint4*>(g_row0 + r * pitch_bytes + c * 16) = q;
  }
  __syncwarp();
}

// Inverse of warp_store_rows: fetch NCH 16-byte chunks of 32 consecutive rows with full-line global loads and hand
// every lane the chunks of its own row.
template <int NCH>
__device__ __forceinline__ void warp_load_rows(const uint8_t* g_row0, size_t pitch_bytes, uint4* v, uint32_t scratch,
                                               int lane, int rows_valid, int chunks_valid) {
  static_assert(NCH >= 1 && NCH <= 8, "one 128-byte smem row per lane");
  uint4 q[NCH];                                             // every global load is issued before the first smem store
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / NCH, c = idx - r * NCH;
    q[it] = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid && c < chunks_valid) q[it] = __ldg(reinterpret_cast<const uint4*>(g_row0 + r * pitch_bytes + c * 16));
  }
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / NCH, c = idx - r * NCH;
    st_shared_v4(scratch + r * 128 + ((c ^ (r & 7)) << 4), q[it]);
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NCH; ++c) v[c] = ld_shared_v4(scratch + lane * 128 + ((c ^ (lane & 7)) << 4));
  __syncwarp();
}

template <int NCOL>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, float* v) {
  // NCOL multiple of 8: greedy x32, x16, x8 (tcgen05.ld is .sync.aligned: the warp must be converged)
  __syncwarp();
  constexpr int n32 = NCOL / 32;
#pragma unroll
  for (int i = 0; i < n32; ++i) tmem_ld_x32(taddr + 32 * i, v + 32 * i);
  constexpr int r32 = NCOL - 32 * n32;
  if constexpr (r32 >= 16) tmem_ld_x16(taddr + 32 * n32, v + 32 * n32);
  if constexpr (r32 % 16 == 8) tmem_ld_x8(taddr + 32 * n32 + (r32 / 16) * 16, v + 32 * n32 + (r32 / 16) * 16);
  tmem_ld_wait();
  tmem_ld_fence_regs<NCOL>(v);
}

// two NCOL-column blocks with a single tcgen05.wait::ld
template <int NCOL>
__device__ __forceinline__ void tmem_load_cols2(uint32_t ta, float* a, uint32_t tb, float* b) {
  __syncwarp();
  constexpr int n32 = NCOL / 32;
  constexpr int r32 = NCOL - 32 * n32;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t t = h ? tb : ta;
    float* v = h ? b : a;
#pragma unroll
    for (int i = 0; i < n32; ++i) tmem_ld_x32(t + 32 * i, v + 32 * i);
    if constexpr (r32 >= 16) tmem_ld_x16(t + 32 * n32, v + 32 * n32);
    if constexpr (r32 % 16 == 8) tmem_ld_x8(t + 32 * n32 + (r32 / 16) * 16, v + 32 * n32 + (r32 / 16) * 16);
  }
  tmem_ld_wait();
  tmem_ld_fence_regs<NCOL>(a);
  tmem_ld_fence_regs<NCOL>(b);
}

template <bool F16, int NW>
__device__ __forceinline__ void pack_row16(const float* v, uint32_t* w) {   // 2*NW floats -> NW packed words
#pragma unroll
  for (int i = 0; i < NW; ++i) w[i] = pack_act2<F16>(v[2 * i], v[2 * i + 1]);
}

#ifdef SWB_PROFILE_EPILOGUES
// cycle counters of a profiling build (tools/gemm_roles.py): [0] issuer total, [1] issuer waiting for TMA data, [2] issuer
// waiting for a free accumulator, [3] tiles (issuer), [4] producer waiting for a free stage, [5] epilogue warps waiting for
// an accumulator, [6] accumulator published -> handed back, [7] handed back -> end of the tile's epilogue, [8] warp-tiles,
// [9] of [7]: waiting for an earlier TMA store to release the scratch buffer, [10] producer total
#define SWB_PROF(x) x
#else
#define SWB_PROF(x)
#endif

// Per-warp epilogue context
struct EpiCtx {
  SWB_PROF(long long store_wait;)
  int row0;          // first row of this warp's 32-row block
  int rows_valid;    // rows of it inside M
  int lane;
  uint32_t scratch;  // this warp's 4 KB transpose buffer
  const CUtensorMap* to0;   // output tensor maps of the TMA-store epilogues: 64-column box (SWIZZLE_128B) ...
  const CUtensorMap* to1;   // ... and the tail box (24 columns un-swizzled, or 32 columns SWIZZLE_64B for qkv)
  bool pending;      // a TMA store issued by lane 0 may still be reading the scratch buffer
  // EPI_LN_RES1 on the 256x352 tile: x travels by TMA (tile loads into the two 4 KB buffers, updated in place, tile stores)
  uint32_t xbar;     // two mbarriers (8 bytes apart): data of buffer 0 / 1 has landed
  uint32_t xph;      // bit b: parity of buffer b's next completion
};

// ---- TMA-store epilogue pieces.  Every lane owns one row of the warp's 32-row block and writes its 16-byte chunks
// into the scratch buffer in exactly the swizzled image a TMA tile store expects (bank-conflict free), then one
// elected lane issues the store: no LDS, no STG, full-line writes generated by the copy engine.
__device__ __forceinline__ void epi_scratch_acquire(EpiCtx& e) {
  if (e.pending) {
    SWB_PROF(const long long t0 = clock64();)
    if (e.lane == 0) bulk_wait_read_all();
    __syncwarp();
    SWB_PROF(e.store_wait += clock64() - t0;)
    e.pending = false;
  }
}
// 64 columns (8 chunks) of 32 rows -> global (column c0, row r0)
__device__ __forceinline__ void tma_store_part64(EpiCtx& e, const uint32_t* w, int c0, int r0) {
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
  epi_scratch_acquire(e);
#pragma unroll
  for (int c = 0; c < 8; ++c) st_shared_v4(e.scratch + e.lane * 128 + ((c ^ (e.lane & 7)) << 4), w4[c]);
  fence_proxy_async_smem();
  __syncwarp();
  if (e.lane == 0) {
    tma_store_2d(e.to0, e.scratch, c0, r0);
    bulk_commit_group();
  }
  e.pending = true;
}
// tails of up to two slots: NT chunks per row (NT = 3: 24 columns, plain image; NT = 4: 32 columns, SWIZZLE_64B image)
template <int NT>
__device__ __forceinline__ void tma_store_tails(EpiCtx& e, const uint32_t* wa, bool va, int ca, int ra, const uint32_t* wb,
                                                bool vb, int cb, int rb) {
  if (!va && !vb) return;
  epi_scratch_acquire(e);
  const uint32_t sa = e.scratch, sb = e.scratch + 2048;
  const int swz = NT == 4 ? ((e.lane >> 1) & 3) : 0;
#pragma unroll
  for (int c = 0; c < NT; ++c) {
    if (va) st_shared_v4(sa + e.lane * (NT * 16) + ((c ^ swz) << 4), reinterpret_cast<const uint4*>(wa)[c]);
    if (vb) st_shared_v4(sb + e.lane * (NT * 16) + ((c ^ swz) << 4), reinterpret_cast<const uint4*>(wb)[c]);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (e.lane == 0) {
    if (va) tma_store_2d(e.to1, sa, ca, ra);
    if (vb) tma_store_2d(e.to1, sb, cb, rb);
    bulk_commit_group();
  }
  e.pending = true;
}

// ---- one 88-column slot, fp32 / 16-bit / embed outputs ---------------------------------------------------
// vreg == nullptr: the slot is read from TMEM block by block (accumulator still held); else its 88 values are in registers
template <int EPI, bool F16, bool FROM_REGS = false>
__device__ __forceinline__ void epi_slot_store(const GemmParams& p, const EpiCtx& e, uint32_t tslot, int n0,
                                               const float* vreg = nullptr) {
  // blocks of 32, 32, 24 columns
#pragma unroll(FROM_REGS ? 3 : 1)
  for (int blk = 0; blk < 3; ++blk) {
    const int c0 = blk * 32;
    const int n = n0 + c0;
    float v[32];
    if constexpr (FROM_REGS) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (c0 + j < kSlot) ? vreg[(c0 + j < kSlot) ? c0 + j : 0] : 0.f;
    } else if (blk < 2) {
      tmem_load_cols<32>(tslot + c0, v);
    } else {
      tmem_load_cols<24>(tslot + c0, v);
#pragma unroll
      for (int j = 24; j < 32; ++j) v[j] = 0.f;
    }
    const int ncol = (blk < 2) ? 32 : 24;
    int cols_valid = p.N - n;                                 // may be <= 0 for a padded tail slot
    cols_valid = cols_valid < 0 ? 0 : (cols_valid > ncol ? ncol : cols_valid);
    if (cols_valid == 0) continue;
    if constexpr (EPI == EPI_EMBED) {
      // + bias[n] + pos_embed[row % tokens, n]: 32-row blocks never straddle a sample (tokens % 32 == 0), so the
      // warp's pos rows are contiguous and can be fetched with full-line loads
      float pe[32];
      const uint8_t* pg = reinterpret_cast<const uint8_t*>(p.pos + static_cast<size_t>(e.row0 % p.pos_rows) * p.N + n);
      warp_load_rows<8>(pg, static_cast<size_t>(p.N) * 4, reinterpret_cast<uint4*>(pe), e.scratch, e.lane, e.rows_valid,
                        cols_valid >> 2);
      if (p.bias) {                                       // (the model path folds the bias into pos_embed at pack time)
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < cols_valid) v[j] += __ldg(p.bias + n + j);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cols_valid) v[j] += pe[j];
    }
    if constexpr (EPI == EPI_STORE_F32) {
      uint8_t* g = reinterpret_cast<uint8_t*>(static_cast<float*>(p.out0) + static_cast<size_t>(e.row0) * p.ldo + n);
      const size_t pitch = static_cast<size_t>(p.ldo) * 4;
      if (blk < 2) warp_store_rows<8>(g, pitch, reinterpret_cast<const uint4*>(v), e.scratch, e.lane, e.rows_valid, cols_valid >> 2);
      else warp_store_rows<6>(g, pitch, reinterpret_cast<const uint4*>(v), e.scratch, e.lane, e.rows_valid, cols_valid >> 2);
    }
    if constexpr (EPI == EPI_STORE_ACT || EPI == EPI_EMBED) {
      uint32_t w[16];
      pack_row16<F16, 16>(v, w);
      uint8_t* g = reinterpret_cast<uint8_t*>(static_cast<uint16_t*>(p.out0) + static_cast<size_t>(e.row0) * p.ldo + n);
      const size_t pitch = static_cast<size_t>(p.ldo) * 2;
      if (blk < 2) warp_store_rows<4>(g, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, cols_valid >> 3);
      else warp_store_rows<3>(g, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, cols_valid >> 3);
      if (EPI == EPI_EMBED && !p.x_single) {
        // residual stream is kept as a 16-bit [hi | lo] pair (hi doubles as the next GEMM's A operand): lo at column N + n
        uint32_t wl[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          wl[j] = pack_act2<F16>(v[2 * j] - unpack_act1<F16>(static_cast<uint16_t>(w[j] & 0xffffu)),
                                 v[2 * j + 1] - unpack_act1<F16>(static_cast<uint16_t>(w[j] >> 16)));
        uint8_t* gl = g + static_cast<size_t>(p.N) * 2;
        if (blk < 2) warp_store_rows<4>(gl, pitch, reinterpret_cast<const uint4*>(wl), e.scratch, e.lane, e.rows_valid, cols_valid >> 3);
        else warp_store_rows<3>(gl, pitch, reinterpret_cast<const uint4*>(wl), e.scratch, e.lane, e.rows_valid, cols_valid >> 3);
      }
    }
  }
}

// ---- one (part, head) slot of the qkv projection ------------------------------------------------------------
// packed weight rows: n = part*dmodel + head*88 + d  (part 0 = q, 1 = k, 2 = v); slot index = n / 88.
template <bool F16>
__device__ __forceinline__ void epi_slot_qkv(const GemmParams& p, const EpiCtx& e, uint32_t tslot, int n0) {
  float v[kHeadDim];
  tmem_load_cols<kHeadDim>(tslot, v);
  const int slot = n0 / kHeadDim;
  const int part = slot / p.heads;
  const int head = slot - part * p.heads;
  if (part >= 3) return;                                       // padded tail slot (warp-uniform)
  if (part < 2) {
    // F.normalize(x, dim=-1, eps=1e-12) (reference models/swinv2.py:123-127); q additionally * exp(min(scale, ln100))
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) ss = fmaf(v[j], v[j], ss);
    float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (part == 0) inv *= __ldg(p.qscale + head);
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) v[j] *= inv;
  }
  uint32_t w[48];                                              // 96 halfs: 88 values + zero pad
  pack_row16<F16, 44>(v, w);
  w[44] = w[45] = w[46] = w[47] = 0u;
  uint8_t* g = reinterpret_cast<uint8_t*>(static_cast<uint16_t*>(p.out0) +
                                          (static_cast<size_t>(part * p.heads + head) * p.M + e.row0) * kHeadDimPad);
  const size_t pitch = kHeadDimPad * 2;
  warp_store_rows<8>(g, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, 8);
  warp_store_rows<4>(g + 128, pitch, reinterpret_cast<const uint4*>(w + 32), e.scratch, e.lane, e.rows_valid, 4);
}

// ---- SwiGLU: gate slot x up slot -> 88 outputs ---------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ void epi_slot_swiglu(const GemmParams& p, const EpiCtx& e, uint32_t tgate, uint32_t tup,
                                                int out_col0) {
  uint8_t* g0 = reinterpret_cast<uint8_t*>(static_cast<uint16_t*>(p.out0) + static_cast<size_t>(e.row0) * p.ldo + out_col0);
  const size_t pitch = static_cast<size_t>(p.ldo) * 2;
#pragma unroll 1
  for (int blk = 0; blk < 3; ++blk) {
    const int c0 = blk * 32;
    float g[32], u[32];
    __syncwarp();
    if (blk < 2) {
      tmem_ld_x32(tgate + c0, g);
      tmem_ld_x32(tup + c0, u);
    } else {
      tmem_ld_x16(tgate + c0, g);
      tmem_ld_x8(tgate + c0 + 16, g + 16);
      tmem_ld_x16(tup + c0, u);
      tmem_ld_x8(tup + c0 + 16, u + 16);
    }
    tmem_ld_wait();
    tmem_ld_fence_regs<24>(g);
    tmem_ld_fence_regs<24>(u);
    if (blk < 2) {
      tmem_ld_fence_regs<8>(g + 24);
      tmem_ld_fence_regs<8>(u + 24);
    }
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = pack_act2<F16>(silu_f(g[2 * j]) * u[2 * j], silu_f(g[2 * j + 1]) * u[2 * j + 1]);
    if (blk < 2) warp_store_rows<4>(g0 + c0 * 2, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, 4);
    else warp_store_rows<3>(g0 + c0 * 2, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, 3);
  }
}

// ---- output head: packed column order is the reference's "(c p1 p2)" (models/swinv2.py:242); write NCHW directly ----
// NV (1 or 2) horizontally adjacent pixels of one output channel.  F is the raw network output.
template <int NV>
__device__ __forceinline__ void head_apply(const GemmParams& p, int b, int ch, size_t pix, const float* F) {
  const size_t hw = static_cast<size_t>(p.H) * p.W;
  const size_t a = (static_cast<size_t>(b) * p.C + ch) * hw + pix;
  float y[NV], t[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) y[i] = p.beta * F[i];
  auto load = [&](const float* src, size_t off) {
    if constexpr (NV == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src + off));
      t[0] = v.x; t[1] = v.y;
    } else {
      t[0] = __ldg(src + off);
    }
  };
  auto store = [&](float* dst, size_t off, const float* v) {
    if constexpr (NV == 2) *reinterpret_cast<float2*>(dst + off) = make_float2(v[0], v[1]);
    else dst[off] = v[0];
  };
  if (p.xt) {
    load(p.xt, a);
#pragma unroll
    for (int i = 0; i < NV; ++i) y[i] = fmaf(p.alpha, t[i], y[i]);
  }
  if (p.fprev) {
    load(p.fprev, a);
#pragma unroll
    for (int i = 0; i < NV; ++i) y[i] = fmaf(p.gamma, t[i], y[i]);
  }
  if (p.out0) store(static_cast<float*>(p.out0), a, y);
  if (p.out_f) store(p.out_f, a, F);
  if (p.state) {
    // generate.py:120-131 (residual branch): X_phys = unstd_x(X) + Y*sigma_diff;  X <- std_x(X_phys)
    const size_t sa = (static_cast<size_t>(b) * p.state_C + ch) * hw + pix;
    const float xs = __ldg(p.x_std + ch), xm = __ldg(p.x_mean + ch), ds = __ldg(p.d_std + ch);
    const float rxs = 1.0f / xs;
    float ph[NV], sn[NV];
    if constexpr (NV == 2) {
      const float2 v = *reinterpret_cast<const float2*>(p.state + sa);
      t[0] = v.x; t[1] = v.y;
    } else {
      t[0] = p.state[sa];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      ph[i] = fmaf(t[i], xs, xm) + y[i] * ds;
      sn[i] = (ph[i] - xm) * rxs;
      if (ch == p.zero_channel) ph[i] = sn[i] = 0.f;
    }
    store(p.state, sa, sn);
    if (p.phys) store(p.phys, a, ph);
  }
}

__device__ __forceinline__ void epi_slot_head(const GemmParams& p, const EpiCtx& e, uint32_t tslot, int n0) {
  const int row = e.row0 + e.lane;
  const bool row_ok = row < p.M;
  const int b = row / p.tokens;
  const int tok = row - b * p.tokens;
  const int gy = tok / p.gw, gx = tok - gy * p.gw;
  const int pp = p.p1 * p.p2;
  const bool pairs = (p.p2 & 1) == 0;      // px runs fastest: even p2 -> columns (n, n+1) are horizontally adjacent pixels
#pragma unroll 1
  for (int c = 0; c < kSlot; c += 8) {
    float v[8];
    tmem_load_cols<8>(tslot + c, v);
    if (!row_ok) continue;
    // lanes are consecutive gx: every warp access below is one contiguous run of 32*p2 floats per (channel, py)
    if (pairs) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const int n = n0 + c + j;
        if (n < p.N) {
          const int ch = n / pp, r = n - ch * pp;
          const int py = r / p.p2, px = r - py * p.p2;
          head_apply<2>(p, b, ch, static_cast<size_t>(gy * p.p1 + py) * p.W + (gx * p.p2 + px), v + j);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = n0 + c + j;
        if (n < p.N) {
          const int ch = n / pp, r = n - ch * pp;
          const int py = r / p.p2, px = r - py * p.p2;
          head_apply<1>(p, b, ch, static_cast<size_t>(gy * p.p1 + py) * p.W + (gx * p.p2 + px), v + j);
        }
      }
    }
  }
}

// NB horizontally adjacent pixel pairs (columns n0, n0 + 2, ...) of one token: same arithmetic as head_apply<2>, but every
// load of the batch is issued before the first store (the state is updated in place, so the compiler may not hoist a
// load above an earlier store by itself: un-batched, the epilogue pays one DRAM round trip per pair).
template <int NB>
__device__ __forceinline__ void head_apply_pairs(const GemmParams& p, int b, int gy, int gx, int n0, const float* F) {
  const size_t hw = static_cast<size_t>(p.H) * p.W;
  const int pp = p.p1 * p.p2;
  size_t a[NB], sa[NB];
  int ch[NB];
  bool ok[NB];
  float2 xt[NB], fp[NB], st[NB];
  float xs[NB], xm[NB], ds[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int n = n0 + 2 * i;
    ok[i] = n < p.N;
    ch[i] = n / pp;
    const int r = n - ch[i] * pp;
    const int py = r / p.p2, px = r - py * p.p2;
    const size_t pix = static_cast<size_t>(gy * p.p1 + py) * p.W + (gx * p.p2 + px);
    a[i] = (static_cast<size_t>(b) * p.C + ch[i]) * hw + pix;
    sa[i] = (static_cast<size_t>(b) * p.state_C + ch[i]) * hw + pix;
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    xt[i] = fp[i] = st[i] = make_float2(0.f, 0.f);
    xs[i] = 1.f;
    xm[i] = ds[i] = 0.f;
    if (ok[i]) {
      if (p.xt) xt[i] = __ldg(reinterpret_cast<const float2*>(p.xt + a[i]));
      if (p.fprev) fp[i] = __ldg(reinterpret_cast<const float2*>(p.fprev + a[i]));
      if (p.state) {
        st[i] = *reinterpret_cast<const float2*>(p.state + sa[i]);
        xs[i] = __ldg(p.x_std + ch[i]);
        xm[i] = __ldg(p.x_mean + ch[i]);
        ds[i] = __ldg(p.d_std + ch[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    if (!ok[i]) continue;
    float y0 = p.beta * F[2 * i], y1 = p.beta * F[2 * i + 1];
    if (p.xt) { y0 = fmaf(p.alpha, xt[i].x, y0); y1 = fmaf(p.alpha, xt[i].y, y1); }
    if (p.fprev) { y0 = fmaf(p.gamma, fp[i].x, y0); y1 = fmaf(p.gamma, fp[i].y, y1); }
    if (p.out0) *reinterpret_cast<float2*>(static_cast<float*>(p.out0) + a[i]) = make_float2(y0, y1);
    if (p.out_f) *reinterpret_cast<float2*>(p.out_f + a[i]) = make_float2(F[2 * i], F[2 * i + 1]);
    if (p.state) {
      // generate.py:120-131 (residual branch): X_phys = unstd_x(X) + Y*sigma_diff;  X <- std_x(X_phys)
      const float rxs = 1.0f / xs[i];
      float ph0 = fmaf(st[i].x, xs[i], xm[i]) + y0 * ds[i], ph1 = fmaf(st[i].y, xs[i], xm[i]) + y1 * ds[i];
      float sn0 = (ph0 - xm[i]) * rxs, sn1 = (ph1 - xm[i]) * rxs;
      if (ch[i] == p.zero_channel) ph0 = ph1 = sn0 = sn1 = 0.f;
      *reinterpret_cast<float2*>(p.state + sa[i]) = make_float2(sn0, sn1);
      if (p.phys) *reinterpret_cast<float2*>(p.phys + a[i]) = make_float2(ph0, ph1);
    }
  }
}

// head epilogue of one epilogue group: both slots leave TMEM, the accumulator is handed back, then the scatter runs
template <typename Release>
__device__ __forceinline__ void epi_group_head(const GemmParams& p, const EpiCtx& e, uint32_t tacc, int n_lo, int n_hi,
                                               Release&& release) {
  // (the fp32 values of both slots plus a batch of scatter operands do not fit in registers: slot A is scattered while
  // slot B still sits in TMEM; the head GEMM's main loop is K = 2112 long, the hand-over is late by one slot's scatter)
  const int row = e.row0 + e.lane;
  const bool row_ok = row < p.M;
  const int b = row / p.tokens;
  const int tok = row - b * p.tokens;
  const int gy = tok / p.gw, gx = tok - gy * p.gw;
  const int pp = p.p1 * p.p2;
  const bool pairs = (p.p2 & 1) == 0;     // px runs fastest: columns (n, n+1) are horizontally adjacent pixels
#pragma unroll 1
  for (int hf = 0; hf < 2; ++hf) {
    const int n0 = hf ? n_hi : n_lo;
    float v[kSlot];
    tmem_load_cols<kSlot>(tacc + hf * kSlot, v);
    if (hf) release();
    if (!row_ok) continue;
    if (pairs) {
#pragma unroll
      for (int c = 0; c < kSlot; c += 8)
        if (n0 + c < p.N) head_apply_pairs<4>(p, b, gy, gx, n0 + c, v + c);
    } else {
#pragma unroll
      for (int j = 0; j < kSlot; ++j) {
        const int n = n0 + j;
        if (n < p.N) {
          const int ch = n / pp, r = n - ch * pp;
          const int py = r / p.p2, px = r - py * p.p2;
          head_apply<1>(p, b, ch, static_cast<size_t>(gy * p.p1 + py) * p.W + (gx * p.p2 + px), v + j);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// "Drain first" epilogues.  With the 352-wide tile there is a single accumulator set, so the UMMA issuer cannot start the
// next tile on a sub-accumulator until its epilogue group has read it.  These handlers pull both 88-column slots of
// the group out of TMEM (packing the first one to 16 bits on the way so that everything fits in registers), hand the
// accumulator back (`release`) and only then do the math / transposes / global stores, which thereby overlap with the
// next tile's main loop.

// 88 packed 16-bit values of one row -> out[row, n0 .. n0+87] through the coalescing transpose
__device__ __forceinline__ void store_slot16(void* out, int ldo, int N, const EpiCtx& e, int n0, const uint32_t* w) {
  int cv = N - n0;
  cv = cv < 0 ? 0 : (cv > kSlot ? kSlot : cv);
  const int ch = cv >> 3;                                   // valid 16-byte chunks (N % 8 == 0)
  if (ch == 0) return;
  uint8_t* g = reinterpret_cast<uint8_t*>(static_cast<uint16_t*>(out) + static_cast<size_t>(e.row0) * ldo + n0);
  const size_t pitch = static_cast<size_t>(ldo) * 2;
  warp_store_rows<8>(g, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, ch);
  if (ch > 8) warp_store_rows<3>(g + 128, pitch, reinterpret_cast<const uint4*>(w + 32), e.scratch, e.lane, e.rows_valid, ch - 8);
}

// profiling variants of warp_store_rows: MODE 1 = transposes only (the global store is predicated off at run time),
// MODE 2 = every lane stores the chunks of its own row directly
template <int NCH, int MODE>
__device__ __forceinline__ void warp_store_rows_prof(uint8_t* g_row0, size_t pitch_bytes, const uint4* v, uint32_t scratch,
                                                     int lane, int rows_valid, int chunks_valid, int never) {
  if constexpr (MODE == 2) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (lane < rows_valid && c < chunks_valid) *reinterpret_cast<uint4*>(g_row0 + lane * pitch_bytes + c * 16) = v[c];
  } else {
#pragma unroll
    for (int c = 0; c < NCH; ++c) st_shared_v4(scratch + lane * 128 + ((c ^ (lane & 7)) << 4), v[c]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NCH; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx / NCH, c = idx - r * NCH;
      const uint4 q = ld_shared_v4(scratch + r * 128 + ((c ^ (r & 7)) << 4));
      if (r < rows_valid && c < chunks_valid && q.x == static_cast<uint32_t>(never))
        *reinterpret_cast<uint4*>(g_row0 + r * pitch_bytes + c * 16) = q;
    }
    __syncwarp();
  }
}

template <bool F16, int MODE, typename Release>
__device__ __forceinline__ void epi_group_store16_prof(const GemmParams& p, const EpiCtx& e, uint32_t tacc, int n_lo, int n_hi,
                                                       Release&& release) {
  float v[kSlot];
  uint32_t wa[44], wb[44];
  tmem_load_cols<kSlot>(tacc, v);
  pack_row16<F16, 44>(v, wa);
  tmem_load_cols<kSlot>(tacc + kSlot, v);
  release();
  pack_row16<F16, 44>(v, wb);
  const size_t pitch = static_cast<size_t>(p.ldo) * 2;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int n0 = hf ? n_hi : n_lo;
    const uint32_t* w = hf ? wb : wa;
    if (n0 >= p.N) continue;
    uint8_t* g = reinterpret_cast<uint8_t*>(static_cast<uint16_t*>(p.out0) + static_cast<size_t>(e.row0) * p.ldo + n0);
    warp_store_rows_prof<8, MODE>(g, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, 8, p.heads + 0x7fc00000);
    warp_store_rows_prof<3, MODE>(g + 128, pitch, reinterpret_cast<const uint4*>(w + 32), e.scratch, e.lane, e.rows_valid, 3, p.heads + 0x7fc00000);
  }
}

template <bool F16, typename Release>
__device__ __forceinline__ void epi_group_store16(const GemmParams& p, EpiCtx& e, uint32_t tacc, int n_lo, int n_hi,
                                                  Release&& release) {
  float va[kSlot], vb[kSlot];
  uint32_t wa[44], wb[44];
  tmem_load_cols2<kSlot>(tacc, va, tacc + kSlot, vb);       // both slots leave TMEM before anything else happens
  release();
  pack_row16<F16, 44>(va, wa);
  pack_row16<F16, 44>(vb, wb);
  if (p.tma_store) {
    if (e.rows_valid <= 0) return;                          // rows and columns outside the tensor are clipped by TMA
    if (n_lo < p.N) tma_store_part64(e, wa, n_lo, e.row0);
    if (n_hi < p.N) tma_store_part64(e, wb, n_hi, e.row0);
    tma_store_tails<3>(e, wa + 32, n_lo + 64 < p.N, n_lo + 64, e.row0, wb + 32, n_hi + 64 < p.N, n_hi + 64, e.row0);
    return;
  }
  store_slot16(p.out0, p.ldo, p.N, e, n_lo, wa);
  store_slot16(p.out0, p.ldo, p.N, e, n_hi, wb);
}

// normalise (q, k) / scale (q) one (part, head) slot held in v[88] and pack it to 96 halfs (zero padded)
template <bool F16>
__device__ __forceinline__ bool qkv_slot_pack(const GemmParams& p, int n0, float* v, uint32_t* w, int* dst_slot) {
  const int slot = n0 / kHeadDim;
  const int part = slot / p.heads;
  const int head = slot - part * p.heads;
  *dst_slot = slot;
  if (part >= 3) return false;                               // padded tail slot (warp-uniform)
  if (part < 2) {
    // F.normalize(x, dim=-1, eps=1e-12) (reference models/swinv2.py:123-127); q additionally * exp(min(scale, ln100))
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) ss = fmaf(v[j], v[j], ss);
    float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (part == 0) inv *= __ldg(p.qscale + head);
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) v[j] *= inv;
  }
  if constexpr (F16) {
    pack_row16<true, 44>(v, w);
  } else {                                                   // bf16 operands: q / k / v may still be stored as fp16
    if (p.qkv_f16) pack_row16<true, 44>(v, w);
    else pack_row16<false, 44>(v, w);
  }
  w[44] = w[45] = w[46] = w[47] = 0u;
  return true;
}

__device__ __forceinline__ void qkv_slot_store(const GemmParams& p, const EpiCtx& e, int slot, const uint32_t* w) {
  uint8_t* g = reinterpret_cast<uint8_t*>(static_cast<uint16_t*>(p.out0) +
                                          (static_cast<size_t>(slot) * p.M + e.row0) * kHeadDimPad);
  const size_t pitch = kHeadDimPad * 2;
  warp_store_rows<8>(g, pitch, reinterpret_cast<const uint4*>(w), e.scratch, e.lane, e.rows_valid, 8);
  warp_store_rows<4>(g + 128, pitch, reinterpret_cast<const uint4*>(w + 32), e.scratch, e.lane, e.rows_valid, 4);
}

template <bool F16, typename Release>
__device__ __forceinline__ void epi_group_qkv(const GemmParams& p, EpiCtx& e, uint32_t tacc, int n_lo, int n_hi,
                                              Release&& release) {
  float va[kSlot], vb[kSlot];
  int sa, sb;
  tmem_load_cols2<kSlot>(tacc, va, tacc + kSlot, vb);
  release();
  if (p.tma_store) {
    // output viewed as [3 * heads * M, 96]: slot s, row r -> row s * M + r (M % 32 == 0: a box never leaves its slot)
    uint32_t wa[48], wb[48];
    const bool oka = qkv_slot_pack<F16>(p, n_lo, va, wa, &sa) && e.rows_valid > 0;
    if (oka) tma_store_part64(e, wa, 0, sa * p.M + e.row0);
    tmem_ld_fence_regs<kSlot>(vb);                          // keeps slot B's math behind slot A's stores (register pressure)
    const bool okb = qkv_slot_pack<F16>(p, n_hi, vb, wb, &sb) && e.rows_valid > 0;
    if (okb) tma_store_part64(e, wb, 0, sb * p.M + e.row0);
    tma_store_tails<4>(e, wa + 32, oka, 64, sa * p.M + e.row0, wb + 32, okb, 64, sb * p.M + e.row0);
    return;
  }
  uint32_t w[48];
  if (qkv_slot_pack<F16>(p, n_lo, va, w, &sa)) qkv_slot_store(p, e, sa, w);
  if (qkv_slot_pack<F16>(p, n_hi, vb, w, &sb)) qkv_slot_store(p, e, sb, w);
}

template <bool F16, typename Release>
__device__ __forceinline__ void epi_group_swiglu(const GemmParams& p, EpiCtx& e, uint32_t tacc, int out_col0,
                                                 bool valid, Release&& release) {
  // gate slot at columns [0,88), up slot at [88,176) of this group's accumulator
#ifdef SWB_SWIGLU_PACKED_GATE
  // gate parked as 16-bit pairs between the two TMEM loads (one extra 2^-12 rounding in front of a 16-bit result)
  float v[kSlot];
  uint32_t g16[44], w[44];
  tmem_load_cols<kSlot>(tacc, v);
  pack_row16<F16, 44>(v, g16);
  tmem_load_cols<kSlot>(tacc + kSlot, v);
  release();
#pragma unroll
  for (int j = 0; j < 44; ++j) {
    const float g0 = unpack_act1<F16>(static_cast<uint16_t>(g16[j] & 0xffffu));
    const float g1 = unpack_act1<F16>(static_cast<uint16_t>(g16[j] >> 16));
    w[j] = pack_act2<F16>(silu_f(g0) * v[2 * j], silu_f(g1) * v[2 * j + 1]);
  }
#else
  float g[kSlot], v[kSlot];
  uint32_t w[44];
  tmem_load_cols2<kSlot>(tacc, g, tacc + kSlot, v);
  release();
#pragma unroll
  for (int j = 0; j < 44; ++j) w[j] = pack_act2<F16>(swiglu_f(g[2 * j], v[2 * j]), swiglu_f(g[2 * j + 1], v[2 * j + 1]));
#endif
  if (p.tma_store) {
    if (valid && e.rows_valid > 0) {
      tma_store_part64(e, w, out_col0, e.row0);
      tma_store_tails<3>(e, w + 32, true, out_col0 + 64, e.row0, w, false, 0, 0);
    }
    return;
  }
  if (valid) store_slot16(p.out0, p.ldo, p.ldo, e, out_col0, w);
}

// ---------------------------------------------------------------------------------------------------------
// EPI_LN_RES: the post-norm residual update of the reference (models/swinv2.py:83-86, :137-138, :100-101, :211-212)
//     x <- x + LayerNorm(branch) * gain[b] + bias[b],     branch = A W^T  (wo or w2 projection)
// done in the GEMM epilogue, so the branch never exists in HBM.  A LayerNorm row spans every column tile, while one
// epilogue warp only sees its 176 columns of 32 rows: each warp publishes the (sum, M2) of its part of the rows to
// global memory (generation-tagged, see epi_group_lnres) and polls until all tiles_n * NSUB groups of those rows have arrived (they
// run at the same time on neighbouring clusters of this persistent kernel: the launcher keeps the number of clusters a
// multiple of tiles_n, every CTA is resident, and the accumulator has already been handed back, so main loops never
// wait on this exchange).  The partials are merged with Chan's formula; the branch is held in registers as fp16 pairs
// (the same rounding the stand-alone LN kernel sees in fp16 mode) and goes through the coalescing transpose, x hi / lo
// are read and written in place with full-line accesses.
// one 64-bit scalar access per (sum, M2) partial: single-copy atomic, so a reader sees a partial whole or not at all
__device__ __forceinline__ float2 ld_relaxed_gpu_f2(const float2* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return make_float2(__uint_as_float(static_cast<unsigned>(v)), __uint_as_float(static_cast<unsigned>(v >> 32)));
}
__device__ __forceinline__ void st_relaxed_gpu_f2(float2* p, float2 v) {
  const unsigned long long u = static_cast<unsigned long long>(__float_as_uint(v.x)) |
                               (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32);
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(u) : "memory");
}
__device__ __forceinline__ float2 ld_shared_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

// 8 consecutive columns of one row: x (hi + lo) += fma(branch * rstd + shift, gain, bias); re-split into hi / lo
template <bool F16, bool SINGLE>
__device__ __forceinline__ void ln_apply8(uint4 br, float2 st, const float* g, const float* b, uint4& xh, uint4& xl) {
  if constexpr (SINGLE) {                                             // x is the hi half alone: one rounding per update
    const uint32_t bw[4] = {br.x, br.y, br.z, br.w};
    uint32_t hw[4] = {xh.x, xh.y, xh.z, xh.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = h2_to_f2(bw[j]);
      const float2 xhi = unpack_act2<F16>(hw[j]);
      hw[j] = pack_act2<F16>(xhi.x + fmaf(fmaf(v.x, st.x, st.y), g[2 * j], b[2 * j]),
                             xhi.y + fmaf(fmaf(v.y, st.x, st.y), g[2 * j + 1], b[2 * j + 1]));
    }
    xh = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    return;
  }
  const uint32_t bw[4] = {br.x, br.y, br.z, br.w};
  uint32_t hw[4] = {xh.x, xh.y, xh.z, xh.w};
  uint32_t lw[4] = {xl.x, xl.y, xl.z, xl.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 v = h2_to_f2(bw[j]);
    const float2 xhi = unpack_act2<F16>(hw[j]), xlo = unpack_act2<F16>(lw[j]);
    const float o0 = (xhi.x + xlo.x) + fmaf(fmaf(v.x, st.x, st.y), g[2 * j], b[2 * j]);
    const float o1 = (xhi.y + xlo.y) + fmaf(fmaf(v.y, st.x, st.y), g[2 * j + 1], b[2 * j + 1]);
    hw[j] = pack_act2<F16>(o0, o1);
    const float2 back = unpack_act2<F16>(hw[j]);
    lw[j] = pack_act2<F16>(o0 - back.x, o1 - back.y);
  }
  xh = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  xl = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// NCH 16-byte chunks (8 columns each) of this warp's 32 rows, starting at column n0.  In the coalesced orientation lane
// `lane` handles (row r, chunk c) = divmod(it * 32 + lane, NCH) in iteration it: full-line accesses to x hi / lo.
template <int NCH>
__device__ __forceinline__ void ln_part_rc(int it, int lane, int& r, int& c) {
  const int idx = it * 32 + lane;
  r = idx / NCH;
  c = idx - r * NCH;
}
// issue every x load of the part (they do not depend on the statistics: called as early as registers allow)
template <int NCH>
__device__ __forceinline__ void ln_part_load(const GemmParams& p, const EpiCtx& e, int n0, uint4* xh, uint4* xl) {
  const uint32_t pitch = static_cast<uint32_t>(p.N) * 2;               // elements per xhl row
  const uint16_t* x0 = p.xhl + static_cast<size_t>(e.row0) * pitch + n0;
  const uint32_t lo_off = static_cast<uint32_t>(p.N) >> 3;             // uint4s between hi and lo of an element
  if constexpr (NCH == 8) {
    // (row, chunk) = (4 it + lane / 8, lane % 8): one base pointer, a constant stride of 4 rows per iteration
    const int r0 = e.lane >> 3;
    const uint4* px = reinterpret_cast<const uint4*>(x0 + r0 * pitch + (e.lane & 7) * 8);
    const uint32_t step = pitch >> 1;                                  // 4 rows in uint4 units (4 * pitch * 2 B / 16 B)
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      xh[it] = xl[it] = make_uint4(0u, 0u, 0u, 0u);
      if (it * 4 + r0 < e.rows_valid) {
        xh[it] = px[it * step];
        xl[it] = px[it * step + lo_off];
      }
    }
  } else {
#pragma unroll
    for (int it = 0; it < NCH; ++it) {
      int r, c;
      ln_part_rc<NCH>(it, e.lane, r, c);
      xh[it] = xl[it] = make_uint4(0u, 0u, 0u, 0u);
      if (r < e.rows_valid) {
        const uint4* px = reinterpret_cast<const uint4*>(x0 + r * pitch + c * 8);
        xh[it] = *px;
        xl[it] = *(px + lo_off);
      }
    }
  }
}
// transpose the fp16 branch values of the part through smem and update x
template <bool F16, int NCH>
__device__ __forceinline__ void ln_part_apply(const GemmParams& p, const EpiCtx& e, uint32_t stat_smem, const uint32_t* w,
                                              int n0, const float* gn, const float* bs, uint4* xh, uint4* xl) {
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int c = 0; c < NCH; ++c) st_shared_v4(e.scratch + e.lane * 128 + ((c ^ (e.lane & 7)) << 4), w4[c]);
  __syncwarp();
  const size_t pitch = static_cast<size_t>(p.N) * 2;
  uint16_t* x0 = p.xhl + static_cast<size_t>(e.row0) * pitch + n0;
  float g[8], b[8];
  auto load_gb = [&](int c) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gn + n0 + c * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gn + n0 + c * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bs + n0 + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bs + n0 + c * 8 + 4));
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  };
  if constexpr (NCH == 8) load_gb(e.lane & 7);                          // the chunk index is the same in every iteration
  uint4* px8 = reinterpret_cast<uint4*>(x0 + (e.lane >> 3) * pitch + (e.lane & 7) * 8);
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    int r, c;
    ln_part_rc<NCH>(it, e.lane, r, c);
    const uint4 q = ld_shared_v4(e.scratch + r * 128 + ((c ^ (r & 7)) << 4));
    const float2 st = ld_shared_f2(stat_smem + r * 8);
    if constexpr (NCH != 8) load_gb(c);
    ln_apply8<F16, false>(q, st, g, b, xh[it], xl[it]);
    if (r < e.rows_valid) {
      uint4* px;
      if constexpr (NCH == 8) px = px8 + it * (static_cast<uint32_t>(p.N));     // 4 rows further per iteration
      else px = reinterpret_cast<uint4*>(x0 + r * pitch + c * 8);
      *px = xh[it];
      *(px + (static_cast<uint32_t>(p.N) >> 3)) = xl[it];
    }
  }
  __syncwarp();
}

// ---- single-value residual stream (p.x_single): x is the hi half alone.  All x loads of a warp-tile fit in registers
// (22 x 16 bytes per lane), so they are issued in one go BEFORE the statistics exchange and its latency hides theirs.
template <int NCH>
__device__ __forceinline__ void ln_part_load1(const GemmParams& p, const EpiCtx& e, int n0, uint4* xh) {
  const uint32_t pitch = static_cast<uint32_t>(p.N) * 2;               // elements per xhl row
  const uint16_t* x0 = p.xhl + static_cast<size_t>(e.row0) * pitch + n0;
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    int r, c;
    ln_part_rc<NCH>(it, e.lane, r, c);
    xh[it] = make_uint4(0u, 0u, 0u, 0u);
    if (r < e.rows_valid SWB_PROF(&& !(p.ln_debug & 32))) xh[it] = *reinterpret_cast<const uint4*>(x0 + r * pitch + c * 8);
  }
}
// park NCH 16-byte chunks of every lane's row in a 4 KB transpose buffer (read back in the coalesced orientation)
template <int NCH>
__device__ __forceinline__ void ln_stage_branch(uint32_t scratch, int lane, const uint32_t* w) {
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int c = 0; c < NCH; ++c) st_shared_v4(scratch + lane * 128 + ((c ^ (lane & 7)) << 4), w4[c]);
}
template <bool F16, int NCH>
__device__ __forceinline__ void ln_part_apply1(const GemmParams& p, const EpiCtx& e, uint32_t stat_smem, uint32_t scratch,
                                               int n0, const float* gn, const float* bs, const uint4* xh) {
  const size_t pitch = static_cast<size_t>(p.N) * 2;
  uint16_t* x0 = p.xhl + static_cast<size_t>(e.row0) * pitch + n0;
  float g[8], b[8];
  auto load_gb = [&](int c) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gn + n0 + c * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gn + n0 + c * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bs + n0 + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bs + n0 + c * 8 + 4));
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  };
#ifdef SWB_PROFILE_EPILOGUES
  const bool no_gb = (p.ln_debug & 64) != 0, no_st = (p.ln_debug & 16) != 0;
  if (no_gb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { g[j] = 1.f; b[j] = 0.f; }
  }
#else
  constexpr bool no_gb = false, no_st = false;
#endif
  if constexpr (NCH == 8) { if (!no_gb) load_gb(e.lane & 7); }          // the chunk index is the same in every iteration
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    int r, c;
    ln_part_rc<NCH>(it, e.lane, r, c);
    const uint4 q = ld_shared_v4(scratch + r * 128 + ((c ^ (r & 7)) << 4));
    const float2 st = ld_shared_f2(stat_smem + r * 8);
    if constexpr (NCH != 8) { if (!no_gb) load_gb(c); }
    uint4 xv = xh[it], dummy = make_uint4(0u, 0u, 0u, 0u);
    ln_apply8<F16, true>(q, st, g, b, xv, dummy);
    if (r < e.rows_valid && (!no_st || xv.x == 0x7fc01234u)) *reinterpret_cast<uint4*>(x0 + r * pitch + c * 8) = xv;
  }
  __syncwarp();
}

// ---- single-value residual stream on the 256x352 tile: x by TMA.  The two 64-column parts of the warp's rows are loaded
// into its two 4 KB buffers (SWIZZLE_128B image: lane = row) BEFORE the warp waits for the accumulator, updated in place in
// the row-per-lane orientation the accumulator arrives in (no transpose of the branch, statistics stay in registers) and
// stored as tiles; the two 24-column tails follow through the same buffers.  No LSU traffic for x at all.
__device__ __forceinline__ void lnx_load(const EpiCtx& e, int buf, bool tail, int n0) {
  const uint32_t bar = e.xbar + 8u * buf;
  if (e.lane == 0) mbar_arrive_expect_tx(bar, tail ? 32u * 48u : 4096u);
  __syncwarp();
  tma_load_2d_elect(e.scratch + 4096u * buf, tail ? e.to1 : e.to0, bar, n0, e.row0);
}
__device__ __forceinline__ void lnx_begin_tile(const GemmParams& p, EpiCtx& e, int n_lo, int n_hi) {
  if (e.rows_valid <= 0) return;
  epi_scratch_acquire(e);                                     // the previous tile's stores have left the buffers
  if (n_lo < p.N) lnx_load(e, 0, false, n_lo);
  if (n_hi < p.N) lnx_load(e, 1, false, n_hi);
  if (e.lane == 0) {                                          // the tails follow through the same buffers: have them in L2 by then
    if (n_lo < p.N) tma_prefetch_2d(e.to1, n_lo + 64, e.row0);
    if (n_hi < p.N) tma_prefetch_2d(e.to1, n_hi + 64, e.row0);
  }
}
// NCH chunks (8 columns each) of this lane's row: x += (branch * rstd + shift) * gain + bias, in place in the buffer
template <bool F16, int NCH>
__device__ __forceinline__ void lnx_apply(const GemmParams& p, EpiCtx& e, int buf, const uint32_t* w, int n0, float2 st,
                                          const float* gn, const float* bs) {
  const uint32_t bar = e.xbar + 8u * buf;
  SWB_PROF(const long long ta_ = clock64();)
  mbar_wait(bar, (e.xph >> buf) & 1u, 40 + buf);
  SWB_PROF(if (e.lane == 0) atomicAdd(&p.prof[NCH == 8 ? 14 : 15], (unsigned long long)(clock64() - ta_));)
  e.xph ^= 1u << buf;
  const uint32_t base = e.scratch + 4096u * buf;
  // three phases, so that the (volatile) shared-memory accesses do not serialise the chunks: every x chunk of the row is
  // read, then the arithmetic runs with the gain / bias loads (the same address in every lane: one L1 wavefront each)
  // free to be scheduled ahead, then everything is written back
  auto addr = [&](int c) { return NCH == 8 ? base + e.lane * 128 + ((c ^ (e.lane & 7)) << 4) : base + e.lane * 48 + c * 16; };
  uint4 xv[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) xv[c] = ld_shared_v4(addr(c));
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint4 dummy = make_uint4(0u, 0u, 0u, 0u);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gn + n0 + c * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gn + n0 + c * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bs + n0 + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bs + n0 + c * 8 + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    ln_apply8<F16, true>(make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]), st, g, b, xv[c], dummy);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) st_shared_v4(addr(c), xv[c]);
#ifdef SWB_PROFILE_EPILOGUES
  if (!(p.ln_debug & 128)) fence_proxy_async_smem();
  __syncwarp();
  if (e.lane == 0) {
    if (!(p.ln_debug & 256)) tma_store_2d(NCH == 8 ? e.to0 : e.to1, base, n0, e.row0);
    bulk_commit_group();
  }
#else
  fence_proxy_async_smem();
  __syncwarp();
  if (e.lane == 0) {
    tma_store_2d(NCH == 8 ? e.to0 : e.to1, base, n0, e.row0);
    bulk_commit_group();
  }
#endif
  e.pending = true;
}

// L2 prefetch of the x hi / lo segments this warp will update (issued before the warp waits for the accumulator, a whole
// main loop ahead of their use)
__device__ __forceinline__ void ln_prefetch_x(const GemmParams& p, int row0, int rows_valid, int lane, int n_lo, int n_hi) {
  if (lane >= rows_valid) return;
  const uint16_t* xr = p.xhl + static_cast<size_t>(row0 + lane) * (2 * p.N);
#pragma unroll
  for (int k = 0; k < (p.x_single ? 2 : 4); ++k) {
    const int n0 = (k & 1) ? n_hi : n_lo;
    if (n0 < p.N) {
      const uint16_t* a = xr + n0 + ((k & 2) ? p.N : 0);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(kSlot * 2) : "memory");
    }
  }
}

// columns [lo, lo + 88) and [hi, hi + 88) held by statistics group s (s = tile_n * NSUB + epilogue group)
template <int NSUB, int CG>
__device__ __forceinline__ void ln_group_cols(int s, int& lo, int& hi) {
  if constexpr (NSUB == 2) {
    lo = (s >> 1) * (2 * kUmmaN) + (s & 1) * kSlot;
    hi = lo + 2 * kSlot;
  } else {
    lo = s * kUmmaN;
    hi = lo + kSlot;
  }
}

template <int NSUB, int CG, bool F16, bool SINGLE, typename Release>
__device__ __forceinline__ void epi_group_lnres(const GemmParams& p, EpiCtx& e, uint32_t stat_smem, uint32_t tacc, int sgrp,
                                                int nslots, Release&& release) {
  constexpr bool kXTma = SINGLE && NSUB == 2;                 // x by TMA (lnx_*): its loads were issued by lnx_begin_tile
  int n_lo, n_hi;
  ln_group_cols<NSUB, CG>(sgrp, n_lo, n_hi);
  const bool va = n_lo < p.N, vb = n_hi < p.N;               // N % 88 == 0: a slot is entirely inside or outside
  // The branch is parked in registers as fp16 pairs (whatever the operand format).  Row statistics in one pass: sums of
  // (x - K) and (x - K)^2 about a pivot K taken from the row itself (no cancellation problem).
  float v[kSlot];
  uint32_t wa[44], wb[44];
  float s1 = 0.f, s2 = 0.f;
  tmem_load_cols<kSlot>(tacc, v);
  const float K = v[0];
#pragma unroll
  for (int j = 0; j < 44; ++j) {
    wa[j] = pack_act2<true>(v[2 * j], v[2 * j + 1]);
    const float d0 = v[2 * j] - K, d1 = v[2 * j + 1] - K;   // (statistics of the fp32 accumulators: they differ from those of
    s1 += d0 + d1;                                          //  the fp16-rounded values by O(2^-12 / sqrt(N)) relative)
    s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
  }
  tmem_load_cols<kSlot>(tacc + kSlot, v);
  release();
  if (!va) { s1 = s2 = 0.f; }
  {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int j = 0; j < 44; ++j) {
      wb[j] = pack_act2<true>(v[2 * j], v[2 * j + 1]);
      const float d0 = v[2 * j] - K, d1 = v[2 * j + 1] - K;
      t1 += d0 + d1;
      t2 = fmaf(d0, d0, fmaf(d1, d1, t2));
    }
    if (vb) { s1 += t1; s2 += t2; }
  }
  if (e.rows_valid <= 0) return;                             // warp-uniform: none of the groups of these rows takes part
  const int cnt = kSlot * (static_cast<int>(va) + static_cast<int>(vb));
  const float fc = static_cast<float>(cnt);
  const float sum = cnt ? fmaf(fc, K, s1) : 0.f;             // sum x      = s1 + n K
  const float m2 = cnt ? fmaxf(s2 - s1 * s1 / fc, 0.f) : 0.f;   // sum (x - mean_p)^2
  // publish, then wait for the other groups of these 32 rows
  SWB_PROF(long long tq0 = clock64();)
  // The partial IS the message: M2 >= 0 carries this launch's generation tag in its three low mantissa bits (2^-20 relative),
  // written with one 64-bit store -- no counter, no release, no fence.  Every (group, row) slot is rewritten by every launch
  // that shares the workspace, so a reader that finds this launch's tag has this launch's partial (launches are ordered
  // by the stream; the slots start cleared, tag 0, which no launch uses).
  const int row = e.row0 + e.lane;
  st_relaxed_gpu_f2(p.ln_stats + static_cast<size_t>(sgrp) * p.ln_stride + row,
                    make_float2(sum, __uint_as_float((__float_as_uint(m2) & ~7u) | p.ln_tag)));
  SWB_PROF({ const long long t_ = clock64(); if (e.lane == 0) atomicAdd(&p.prof[11], (unsigned long long)(t_ - tq0)); tq0 = t_; })
  constexpr bool single = SINGLE;
  // single-value stream: every x load of this warp-tile goes out before the wait; pair: the loads of the first part
  uint4 xhA[8], xlA[8], xhB[3], xlB[3];
  uint4 xs0[8], xs1[3], xs2[8], xs3[3];
  if constexpr (kXTma) {
  } else if constexpr (single) {
    // the two 64-column parts of the branch wait in shared memory (frees 64 registers for the x loads)
    ln_stage_branch<8>(e.scratch, e.lane, wa);
    ln_stage_branch<8>(e.scratch + 4096, e.lane, wb);
    if (va) { ln_part_load1<8>(p, e, n_lo, xs0); ln_part_load1<3>(p, e, n_lo + 64, xs1); }
    if (vb) { ln_part_load1<8>(p, e, n_hi, xs2); ln_part_load1<3>(p, e, n_hi + 64, xs3); }
  } else {
    ln_part_load<8>(p, e, va ? n_lo : n_hi, xhA, xlA);
  }
  // every lane polls the partials of its own row until all of them carry this launch's tag (gpu-scope loads: L2)
  constexpr int kMaxGroups = 12;
  float2 part[kMaxGroups];
  SWB_PROF({ const long long t_ = clock64(); if (e.lane == 0) atomicAdd(&p.prof[12], (unsigned long long)(t_ - tq0)); tq0 = t_; })
  {
    const long long t0 = clock64();
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int s = 0; s < kMaxGroups; ++s) {
        part[s] = make_float2(0.f, 0.f);
        if (s < nslots) {
          part[s] = ld_relaxed_gpu_f2(p.ln_stats + static_cast<size_t>(s) * p.ln_stride + row);
          ok = ok && ((__float_as_uint(part[s].y) & 7u) == p.ln_tag);
        }
      }
      if (__all_sync(0xffffffffu, ok) || (p.ln_debug & 1)) break;
      __nanosleep(20);
      if (clock64() - t0 > SWB_WATCHDOG_CYCLES) {
        if (e.lane == 0)
          printf("[swift_b200] LayerNorm statistics watchdog: block %d warp %d rows %d.. (tag %u)\n", (int)blockIdx.x,
                 (int)(threadIdx.x >> 5), e.row0, p.ln_tag);
        __trap();
      }
    }
#pragma unroll
    for (int s = 0; s < kMaxGroups; ++s) part[s].y = __uint_as_float(__float_as_uint(part[s].y) & ~7u);
  }
  // merge the partials (Chan et al.): mean, M2 over all N columns
  float tot = 0.f;
#pragma unroll
  for (int s = 0; s < kMaxGroups; ++s) tot += part[s].x;
  const float mean = tot / static_cast<float>(p.N);
  float M2 = 0.f;
#pragma unroll
  for (int s = 0; s < kMaxGroups; ++s) {
    int lo, hi;
    ln_group_cols<NSUB, CG>(s, lo, hi);
    const int c = kSlot * (static_cast<int>(lo < p.N) + static_cast<int>(hi < p.N));
    if (s < nslots && c) {
      const float d = part[s].x / static_cast<float>(c) - mean;
      M2 += part[s].y + static_cast<float>(c) * d * d;
    }
  }
  const float rstd = rsqrtf(M2 / static_cast<float>(p.N) + p.ln_eps);
  const int b = e.row0 / p.tokens;                           // tokens % 32 == 0: the 32 rows belong to one sample
  const float* gn = p.gain + static_cast<size_t>(b) * p.N;
  const float* bs = p.lnbias + static_cast<size_t>(b) * p.N;
  if constexpr (kXTma) {
    const float2 st = make_float2(rstd, -mean * rstd);
    SWB_PROF({ const long long t_ = clock64(); if (e.lane == 0) atomicAdd(&p.prof[13], (unsigned long long)(t_ - tq0)); tq0 = t_; })
    // the 24-column tails re-use the buffers: each is requested as soon as its buffer's store has been read, so that the
    // other slot's arithmetic covers the round trip
    if (va) {
      lnx_apply<F16, 8>(p, e, 0, wa, n_lo, st, gn, bs);
      SWB_PROF(const long long tb_ = clock64();)
      if (e.lane == 0) bulk_wait_read_all();
      __syncwarp();
      SWB_PROF(e.store_wait += clock64() - tb_;)
      lnx_load(e, 0, true, n_lo + 64);
    }
    if (vb) {
      lnx_apply<F16, 8>(p, e, 1, wb, n_hi, st, gn, bs);
      SWB_PROF(const long long tb_ = clock64();)
      if (e.lane == 0) bulk_wait_read_all();
      __syncwarp();
      SWB_PROF(e.store_wait += clock64() - tb_;)
      lnx_load(e, 1, true, n_hi + 64);
    }
    if (va) lnx_apply<F16, 3>(p, e, 0, wa + 32, n_lo + 64, st, gn, bs);
    if (vb) lnx_apply<F16, 3>(p, e, 1, wb + 32, n_hi + 64, st, gn, bs);
    SWB_PROF({ const long long t_ = clock64(); if (e.lane == 0) atomicAdd(&p.prof[16], (unsigned long long)(t_ - tq0)); })
    return;
  }
  st_shared_f2(stat_smem + e.lane * 8, make_float2(rstd, -mean * rstd));
  __syncwarp();
  SWB_PROF({ const long long t_ = clock64(); if (e.lane == 0) atomicAdd(&p.prof[13], (unsigned long long)(t_ - tq0)); tq0 = t_; })
  if (p.ln_debug & 2) return;
  if constexpr (single) {
    // (the __syncwarp after the statistics store also orders the staged branch parts)
    if (va) ln_part_apply1<F16, 8>(p, e, stat_smem, e.scratch, n_lo, gn, bs, xs0);
    if (vb) ln_part_apply1<F16, 8>(p, e, stat_smem, e.scratch + 4096, n_hi, gn, bs, xs2);
    ln_stage_branch<3>(e.scratch, e.lane, wa + 32);             // the 24-column tails take the buffers over
    ln_stage_branch<3>(e.scratch + 4096, e.lane, wb + 32);
    __syncwarp();
    if (va) ln_part_apply1<F16, 3>(p, e, stat_smem, e.scratch, n_lo + 64, gn, bs, xs1);
    if (vb) ln_part_apply1<F16, 3>(p, e, stat_smem, e.scratch + 4096, n_hi + 64, gn, bs, xs3);
    return;
  }
  // parts: (lo, 64 columns) (lo, 24) (hi, 64) (hi, 24); the loads of the next part are in flight while one is applied
  if (va) {
    ln_part_load<3>(p, e, n_lo + 64, xhB, xlB);
    ln_part_apply<F16, 8>(p, e, stat_smem, wa, n_lo, gn, bs, xhA, xlA);
    if (vb) ln_part_load<8>(p, e, n_hi, xhA, xlA);
    ln_part_apply<F16, 3>(p, e, stat_smem, wa + 32, n_lo + 64, gn, bs, xhB, xlB);
  }
  if (vb) {
    ln_part_load<3>(p, e, n_hi + 64, xhB, xlB);
    ln_part_apply<F16, 8>(p, e, stat_smem, wb, n_hi, gn, bs, xhA, xlA);
    ln_part_apply<F16, 3>(p, e, stat_smem, wb + 32, n_hi + 64, gn, bs, xhB, xlB);
  }
}

// ---------------------------------------------------------------------------------------------------------

template <int NSUB, int CG, int EPI, bool F16>
__global__ void __launch_bounds__(GemmCfg<NSUB, CG>::kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_o0, const __grid_constant__ CUtensorMap tmap_o1,
                    const GemmParams p) {
  using S = GemmCfg<NSUB, CG, EPI == EPI_LN_RES || EPI == EPI_LN_RES1>;
  constexpr int kStages = S::kStages;
  constexpr int BN = kUmmaN;
  constexpr int kTileN = S::kTileN;
  constexpr int kKSteps = kBlockK / kUmmaK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStages * S::kABytes;
  const uint32_t scratch0 = smem_base + kStages * S::kStageBytes;
  const uint32_t bars = scratch0 + S::kScratchBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  const int tiles_n = (p.N + kTileN - 1) / kTileN;
  const int tiles_m = (p.M + kBlockM * CG - 1) / (kBlockM * CG);
  const int tiles_mn = tiles_m * tiles_n;
  const int nsplit = p.splits > 1 ? p.splits : 1;
  const int num_tiles = tiles_mn * nsplit * (p.batch > 1 ? p.batch : 1);   // (batch, split)-major tile order
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int tail_k = p.K - (num_kb - 1) * kBlockK;
  const int tail_ksteps = (tail_k + kUmmaK - 1) / kUmmaK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_store) {
      tma_prefetch_desc(&tmap_o0);
      tma_prefetch_desc(&tmap_o1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
#ifdef SWB_PROFILE_EPILOGUES
      mbar_init(full_bar(s), (CG == 2 && (p.ln_debug & 8)) ? 2 : 1);
#else
      mbar_init(full_bar(s), 1);
#endif
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * CG);   // one arrive per epilogue warp (of one group) of every CTA in the pair
    }
    if constexpr (EPI == EPI_LN_RES1 && NSUB == 2) {
      for (int s = 0; s < 2 * S::kEpiWarps; ++s) mbar_init(bars + 512u + 8u * s, 1);   // x tile loads of the epilogue warps
    }
    fence_mbar_init_cluster();
  }
  if (warp == 2) {
    tmem_alloc<CG>(tmem_slot, kTmemCols);
    tmem_relinquish<CG>();
  }
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // 384 threads x 168 registers at launch: warps 0-3 keep 64, the two epilogue warpgroups grow to 216 (64*128 + 216*256 = 384*168: setmaxnreg.inc blocks for ever if the CTA pool is exceeded) (the
  // re-allocation sits at the top of each warpgroup's own branch: ptxas budgets the code it dominates)
  if (warp < 4) {
#ifndef SWB_NO_SETMAXNREG
  if constexpr (S::kEpiWarps == 8) setmaxnreg_dec<64>();
#endif
  if (warp == 0) {
    // ===================================== TMA producer (whole warp converged, one elected lane issues) ==========
    const uint32_t full_leader0 = (CG == 2) ? mapa_u32(full_bar(0), 0) : full_bar(0);
    int stage = 0;
    uint32_t phase = 0;
    SWB_PROF(long long pw = 0; const long long pt0 = clock64();)
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int bs = tile / tiles_mn, tmn = tile - bs * tiles_mn;
      const int bidx = bs / nsplit, split = bs - bidx * nsplit;
      const int tm = tmn / tiles_n, tn = tmn - tm * tiles_n;
      const int row0 = bidx * p.M + tm * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM;
      const int col0 = bidx * p.N + tn * kTileN + static_cast<int>(cta_rank) * S::kBRows;
      const int k0 = split * p.K;
      for (int kb = 0; kb < num_kb; ++kb) {
        SWB_PROF(const long long tw = clock64();)
        mbar_wait(empty_bar(stage), phase ^ 1u, 1);
        SWB_PROF(pw += clock64() - tw;)
#ifdef SWB_PROFILE_EPILOGUES
        if (p.ln_debug & 8) {       // profiling: nothing is loaded -- with bit 2 the epilogue runs alone
          if (cta_rank == 0) mbar_arrive_expect_tx_elect(full_bar(stage), 0);
          else if (lane == 0) mbar_arrive_cluster(full_leader0 + 8u * stage);     // (keeps the two producers in step)
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
          continue;
        }
#endif
        if (cta_rank == 0) mbar_arrive_expect_tx_elect(full_bar(stage), S::kStageBytes * CG);
        const uint32_t a_dst = smem_a + stage * S::kABytes;
        const uint32_t b_dst = smem_b + stage * S::kBBytes;
        if constexpr (CG == 2) {
          const uint32_t bar = full_leader0 + 8u * stage;
          tma_load_2d_pair_elect(a_dst, &tmap_a, bar, k0 + kb * kBlockK, row0);
          tma_load_2d_pair_elect(b_dst, &tmap_b, bar, k0 + kb * kBlockK, col0);
        } else {
          tma_load_2d_elect(a_dst, &tmap_a, full_bar(stage), k0 + kb * kBlockK, row0);
          tma_load_2d_elect(b_dst, &tmap_b, full_bar(stage), k0 + kb * kBlockK, col0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
    SWB_PROF(if (lane == 0) { atomicAdd(&p.prof[4], (unsigned long long)pw); atomicAdd(&p.prof[10], (unsigned long long)(clock64() - pt0)); })
  } else if (warp == 1) {
    // ===================================== UMMA issuer (leader CTA; whole warp converged) =======================
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM * CG, BN, /*A=*/F16, /*B=*/F16);   // mixed A/B formats trap
      // K-major SWIZZLE_128B descriptors: 8-row groups 1024 B apart; only the 14-bit address field varies.
      // +16 elements along K = +32 bytes = +2 in the (address >> 4) field; never carries out of the field.
      const uint64_t desc_hi = make_smem_desc(0, 16, 1024, SWZ_128B);
      constexpr uint32_t kSubDescStep = (BN / CG) * 128 / 16;        // next sub-tile's W rows inside the stage
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      // skew: at most kStages - 2 k-blocks (the stages of a tile end are held while sub-tile 1 catches up) and num_kb / 2
#ifdef SWB_PROFILE_EPILOGUES
      const bool nomma = (p.ln_debug & 4) != 0;     // profiling: no MMA is issued -- the rate at which TMA alone can feed the tiles
#else
      constexpr bool nomma = false;
#endif
      SWB_PROF(long long iw_full = 0; long long iw_acc = 0; const long long it0 = clock64();)
      int skew = (NSUB == 2) ? p.skew : 0;
      skew = skew > kStages - 2 ? kStages - 2 : skew;
      skew = skew > num_kb / 2 ? num_kb / 2 : skew;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const uint32_t par = static_cast<uint32_t>(it) & 1u;
        uint32_t tmem_d;
        if constexpr (NSUB == 1) {
          // double-buffered accumulator: stage it&1, re-used every second tile
          SWB_PROF(const long long tw_ = clock64();)
          mbar_wait(tempty_bar(par), ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u, 2);
          SWB_PROF(iw_acc += clock64() - tw_;)
          tcgen05_fence_after();
          tmem_d = tmem_base + par * kSubStride;
        } else {
          tmem_d = tmem_base;
        }
        int kb_begin = 0, kb_end = num_kb;
#if !SWB_A_TMEM
        if constexpr (NSUB == 2) {
          if (skew > 0) {
            // ---- head of a skewed tile: sub-tile 0 alone over the first `skew` k-blocks (its accumulator was published
            // `skew` k-blocks before the end of the previous tile, so it has been drained by now), then sub-tile 1 over the
            // same stages, which it releases
            int s = stage;
            uint32_t ph = phase;
            SWB_PROF(const long long tw_ = clock64();)
            mbar_wait(tempty_bar(0), par ^ 1u, 2);
            SWB_PROF(iw_acc += clock64() - tw_;)
            tcgen05_fence_after();
            for (int kb = 0; kb < skew; ++kb) {
              SWB_PROF(const long long tf_ = clock64();)
              mbar_wait(full_bar(s), ph, 3);
              SWB_PROF(iw_full += clock64() - tf_;)
              tcgen05_fence_after();
              const uint64_t ad = desc_hi | static_cast<uint64_t>(((smem_a + s * S::kABytes) & 0x3FFFFu) >> 4);
              const uint64_t bd = desc_hi | static_cast<uint64_t>(((smem_b + s * S::kBBytes) & 0x3FFFFu) >> 4);
#pragma unroll
              for (int k = 0; k < kKSteps; ++k) umma_f16_ss_elect<CG>(tmem_d, ad + 2u * k, bd + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              if (++s == kStages) { s = 0; ph ^= 1u; }
            }
            SWB_PROF(const long long tw2_ = clock64();)
            mbar_wait(tempty_bar(1), par ^ 1u, 2);
            SWB_PROF(iw_acc += clock64() - tw2_;)
            tcgen05_fence_after();
            for (int kb = 0; kb < skew; ++kb) {
              const uint64_t ad = desc_hi | static_cast<uint64_t>(((smem_a + stage * S::kABytes) & 0x3FFFFu) >> 4);
              const uint64_t bd = desc_hi | static_cast<uint64_t>(((smem_b + stage * S::kBBytes) & 0x3FFFFu) >> 4);
#pragma unroll
              for (int k = 0; k < kKSteps; ++k)
                umma_f16_ss_elect<CG>(tmem_d + kSubStride, ad + 2u * k, bd + kSubDescStep + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_commit_elect<CG>(empty_bar(stage));
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            kb_begin = skew;
            kb_end = num_kb - skew;
          }
        }
#endif
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          SWB_PROF(const long long tf_ = clock64();)
          mbar_wait(full_bar(stage), phase, 3);             // TMA bytes of this stage have landed (both CTAs)
          SWB_PROF(iw_full += clock64() - tf_;)
          tcgen05_fence_after();
          const uint64_t adesc0 = desc_hi | static_cast<uint64_t>(((smem_a + stage * S::kABytes) & 0x3FFFFu) >> 4);
          const uint64_t bdesc0 = desc_hi | static_cast<uint64_t>(((smem_b + stage * S::kBBytes) & 0x3FFFFu) >> 4);
          const bool full_block = (kb != num_kb - 1) || (tail_ksteps == kKSteps);
          if constexpr (NSUB == 1) {
            if (full_block) {
#pragma unroll
              for (int k = 0; k < kKSteps; ++k)
                umma_f16_ss_elect<CG>(tmem_d, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            } else {
              for (int k = 0; k < tail_ksteps; ++k)
                umma_f16_ss_elect<CG>(tmem_d, adesc0 + 2u * k, bdesc0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          } else {
            // A operand of this k-block: straight from shared memory (read once per sub-tile: SWB_A_TMEM 0), or copied once
            // into one of two 32-column TMEM slots (columns 176..239, free between the accumulators) and read from there
            // by both sub-tiles: shared-memory bandwidth, not the tensor pipe, is what limits this kernel.
            const int nk = full_block ? kKSteps : tail_ksteps;
#if SWB_A_TMEM
            const uint32_t a_t = tmem_base + kUmmaN + (static_cast<uint32_t>(kb) & 1u) * 32u;
            if (full_block) {
#pragma unroll
              for (int k = 0; k < kKSteps; ++k) utccp_128x256b_pair_elect(a_t + 8u * k, adesc0 + 2u * k);
            } else {
              for (int k = 0; k < tail_ksteps; ++k) utccp_128x256b_pair_elect(a_t + 8u * k, adesc0 + 2u * k);
            }
            auto mma = [&](int j, int k, uint32_t acc) {
              umma_f16_ts_pair_elect(tmem_d + j * kSubStride, a_t + 8u * k, bdesc0 + j * kSubDescStep + 2u * k, idesc, acc);
            };
#else
            auto mma = [&](int j, int k, uint32_t acc) {
              if (nomma) return;
              umma_f16_ss_elect<CG>(tmem_d + j * kSubStride, adesc0 + 2u * k, bdesc0 + j * kSubDescStep + 2u * k, idesc, acc);
            };
#endif
            if (kb == 0) {
              // first k-block of a tile: each sub-accumulator becomes writable as soon as its own epilogue group
              // of the previous tile has drained it, so start on sub 0 while sub 1 may still be read
#pragma unroll 1
              for (int j = 0; j < NSUB; ++j) {
                SWB_PROF(const long long tw_ = clock64();)
                mbar_wait(tempty_bar(j), par ^ 1u, 2);
                SWB_PROF(iw_acc += clock64() - tw_;)
                tcgen05_fence_after();
                for (int k = 0; k < nk; ++k) mma(j, k, k != 0 ? 1u : 0u);
              }
            } else if (full_block) {
#if SWB_A_TMEM
#pragma unroll
              for (int k = 0; k < kKSteps; ++k) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) mma(j, k, 1u);
              }
#else
              // steady state: the 8 UMMAs of the k-block behind one elect (issue slots are contended, see ptx.cuh)
              if (!nomma) umma_f16_kblock_pair_elect(tmem_d, tmem_d + kSubStride, static_cast<uint32_t>(adesc0), static_cast<uint32_t>(bdesc0),
                                         kSubDescStep, static_cast<uint32_t>(desc_hi >> 32), idesc);
#endif
            } else {
              for (int k = 0; k < tail_ksteps; ++k)
                for (int j = 0; j < NSUB; ++j) mma(j, k, 1u);
            }
          }
          umma_commit_elect<CG>(empty_bar(stage));          // smem stage reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
#if !SWB_A_TMEM
        if constexpr (NSUB == 2) {
          if (skew > 0) {
            // ---- tail of a skewed tile: sub-tile 0 finishes first and is published while sub-tile 1 still runs
            int s = stage;
            uint32_t ph = phase;
            for (int kb = num_kb - skew; kb < num_kb; ++kb) {
              SWB_PROF(const long long tf_ = clock64();)
              mbar_wait(full_bar(s), ph, 3);
              SWB_PROF(iw_full += clock64() - tf_;)
              tcgen05_fence_after();
              const uint64_t ad = desc_hi | static_cast<uint64_t>(((smem_a + s * S::kABytes) & 0x3FFFFu) >> 4);
              const uint64_t bd = desc_hi | static_cast<uint64_t>(((smem_b + s * S::kBBytes) & 0x3FFFFu) >> 4);
              const int nk = (kb == num_kb - 1) ? tail_ksteps : kKSteps;
              for (int k = 0; k < nk; ++k) umma_f16_ss_elect<CG>(tmem_d, ad + 2u * k, bd + 2u * k, idesc, 1u);
              if (++s == kStages) { s = 0; ph ^= 1u; }
            }
            umma_commit_elect<CG>(tfull_bar(0));
            for (int kb = num_kb - skew; kb < num_kb; ++kb) {
              const uint64_t ad = desc_hi | static_cast<uint64_t>(((smem_a + stage * S::kABytes) & 0x3FFFFu) >> 4);
              const uint64_t bd = desc_hi | static_cast<uint64_t>(((smem_b + stage * S::kBBytes) & 0x3FFFFu) >> 4);
              const int nk = (kb == num_kb - 1) ? tail_ksteps : kKSteps;
              for (int k = 0; k < nk; ++k)
                umma_f16_ss_elect<CG>(tmem_d + kSubStride, ad + 2u * k, bd + kSubDescStep + 2u * k, idesc, 1u);
              umma_commit_elect<CG>(empty_bar(stage));
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            umma_commit_elect<CG>(tfull_bar(1));
            continue;
          }
        }
#endif
        umma_commit_elect<CG>(tfull_bar(NSUB == 1 ? par : 0));   // accumulator(s) complete -> epilogue (both CTAs)
        if constexpr (NSUB == 2) umma_commit_elect<CG>(tfull_bar(1));     // (every epilogue group waits on its own barrier)
      }
      SWB_PROF(if (lane == 0) {
        atomicAdd(&p.prof[0], (unsigned long long)(clock64() - it0));
        atomicAdd(&p.prof[1], (unsigned long long)iw_full);
        atomicAdd(&p.prof[2], (unsigned long long)iw_acc);
        atomicAdd(&p.prof[3], (unsigned long long)it);
      })
    }
  }
  } else {
    // ===================================== epilogue =====================================
#ifndef SWB_NO_SETMAXNREG
    if constexpr (S::kEpiWarps == 8) setmaxnreg_inc<216>();
#endif
    const int quad = warp & 3;                              // TMEM lane quadrant this warp may access
    const int grp = (warp - 4) >> 2;                        // NSUB == 2: epilogue group = sub-tile it drains
    const uint32_t tempty_leader0 = (CG == 2) ? mapa_u32(tempty_bar(0), 0) : tempty_bar(0);
    EpiCtx e;
    e.lane = lane;
    e.scratch = scratch0 + (warp - 4) * S::kScratchPerWarp;
    e.to0 = &tmap_o0;
    e.to1 = &tmap_o1;
    e.pending = false;
    e.xbar = bars + 512u + 16u * (warp - 4);
    e.xph = 0u;
    SWB_PROF(e.store_wait = 0; long long ew_full = 0; long long ew_drain = 0; long long ew_rest = 0; long long t_rel = 0;)
    const uint32_t stat_smem = scratch0 + S::kEpiWarps * S::kScratchPerWarp + (warp - 4) * 256;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const uint32_t par = static_cast<uint32_t>(it) & 1u;
      const int bs = tile / tiles_mn, tmn = tile - bs * tiles_mn;      // bs = batch * splits + split: the output block
      const int tm = tmn / tiles_n, tn = tmn - tm * tiles_n;
      const int n_tile = tn * kTileN;
      if constexpr (EPI == EPI_LN_RES || EPI == EPI_LN_RES1) {
        const int r0 = tm * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM + quad * 32;
        int lo, hi;
        ln_group_cols<NSUB, CG>(tn * NSUB + grp, lo, hi);
        if constexpr (EPI == EPI_LN_RES1 && NSUB == 2) {
          // the x tiles of this warp's rows start their way into shared memory a whole main loop before they are needed
          e.row0 = r0;
          e.rows_valid = p.M - r0 < 0 ? 0 : (p.M - r0 > 32 ? 32 : p.M - r0);
          lnx_begin_tile(p, e, lo, hi);
        } else {
          ln_prefetch_x(p, r0, p.M - r0, lane, lo, hi);
        }
      }
      uint32_t tacc;                                        // column 0 of the accumulator this warp drains
      if constexpr (NSUB == 1) {
        mbar_wait(tfull_bar(par), (static_cast<uint32_t>(it) >> 1) & 1u, 4);
        tacc = tmem_base + par * kSubStride;
      } else {
        SWB_PROF(const long long tw_ = clock64();)
        mbar_wait(tfull_bar(grp), par, 4);
        SWB_PROF(t_rel = clock64(); ew_full += t_rel - tw_;)
        tacc = tmem_base + grp * kSubStride;
      }
      tcgen05_fence_after();
      tacc += static_cast<uint32_t>(quad * 32) << 16;
      e.row0 = tm * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM + quad * 32;
      const int rv = p.M - e.row0;
      e.rows_valid = rv < 0 ? 0 : (rv > 32 ? 32 : rv);
      if constexpr (EPI == EPI_STORE_F32) e.row0 += bs * p.M;          // (batch, split) results are stacked along the rows
      // Accumulator columns [0,88) and [88,176) of sub-tile j hold the global 88-column slots
      //   CG == 2:  s = j and s = NSUB + j   (the pair splits the staged W rows in two contiguous halves)
      //   CG == 1:  s = 0 and s = 1
      const int s_lo = (NSUB == 1) ? 0 : grp;
      const int s_hi = (NSUB == 1) ? 1 : NSUB + grp;
      auto release = [&]() {
        // every TMEM read of this warp for this tile has completed (tcgen05.wait::ld): hand the accumulator back
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          const int bar_idx = (NSUB == 1) ? static_cast<int>(par) : grp;
          if constexpr (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * bar_idx);
          else mbar_arrive(tempty_bar(bar_idx));
        }
        SWB_PROF({ const long long t_ = clock64(); ew_drain += t_ - t_rel; t_rel = t_; })
      };
      const int n_lo = n_tile + s_lo * kSlot, n_hi = n_tile + s_hi * kSlot;
      if constexpr (EPI == EPI_SWIGLU) {
        // packed w1 rows per tile: [NSUB gate slots | NSUB up slots]
        epi_group_swiglu<F16>(p, e, tacc, (tn * NSUB + s_lo) * kSlot, n_tile < p.N, release);
      } else if constexpr (EPI == EPI_QKV) {
        epi_group_qkv<F16>(p, e, tacc, n_lo, n_hi, release);
      } else if constexpr (EPI == EPI_STORE_ACT) {
        epi_group_store16<F16>(p, e, tacc, n_lo, n_hi, release);
      } else if constexpr (EPI == EPI_HEAD) {
        epi_group_head(p, e, tacc, n_lo, n_hi, release);
      } else if constexpr ((EPI == EPI_EMBED || EPI == EPI_STORE_F32) && NSUB == 2) {
        float va[kSlot], vb[kSlot];
        tmem_load_cols2<kSlot>(tacc, va, tacc + kSlot, vb);
        release();
        epi_slot_store<EPI, F16, true>(p, e, 0u, n_lo, va);
        epi_slot_store<EPI, F16, true>(p, e, 0u, n_hi, vb);
      } else if constexpr (EPI == EPI_LN_RES || EPI == EPI_LN_RES1) {
        epi_group_lnres<NSUB, CG, F16, EPI == EPI_LN_RES1>(p, e, stat_smem, tacc, tn * NSUB + grp, tiles_n * NSUB, release);
      } else if constexpr (EPI == EPI_SMEM_ONLY) {
        epi_group_store16_prof<F16, 1>(p, e, tacc, n_lo, n_hi, release);
      } else if constexpr (EPI == EPI_DIRECT) {
        epi_group_store16_prof<F16, 2>(p, e, tacc, n_lo, n_hi, release);
      } else if constexpr (EPI == EPI_BUSY) {
        release();
        // p.dmodel == 0: one dependent chain (a warp issues every 4th cycle); else 8 independent chains (every cycle)
        float a[8], bq = 1.0001f;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = static_cast<float>(lane + c);
        if (p.dmodel == 0) {
          for (int i = 0; i < p.heads; ++i) {
#pragma unroll
            for (int j = 0; j < 64; ++j) a[0] = fmaf(a[0], bq, 0.5f);
          }
        } else {
          for (int i = 0; i < p.heads; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
              for (int c = 0; c < 8; ++c) a[c] = fmaf(a[c], bq, 0.5f);
          }
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += a[c];
        if (acc == 1.2345e-33f && p.out0) static_cast<float*>(p.out0)[0] = acc;
      } else if constexpr (EPI == EPI_DISCARD) {
        release();
      } else if constexpr (EPI == EPI_DRAIN) {
        float va[kSlot], vb[kSlot];
        tmem_load_cols2<kSlot>(tacc, va, tacc + kSlot, vb);
        release();
        float acc = 0.f;                                    // consume the values (ptxas drops tcgen05.ld with dead results)
#pragma unroll
        for (int j = 0; j < kSlot; ++j) acc += va[j] * vb[j];
        if (acc == 1.2345e-33f && p.out0) static_cast<float*>(p.out0)[0] = acc;
      } else {
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          const int n0 = hf ? n_hi : n_lo;
          const uint32_t tslot = tacc + hf * kSlot;
          if constexpr (EPI == EPI_HEAD) epi_slot_head(p, e, tslot, n0);
          else epi_slot_store<EPI, F16>(p, e, tslot, n0);
        }
        release();
      }
      SWB_PROF(ew_rest += clock64() - t_rel;)
    }
    SWB_PROF(if (lane == 0) {
      atomicAdd(&p.prof[5], (unsigned long long)ew_full);
      atomicAdd(&p.prof[6], (unsigned long long)ew_drain);
      atomicAdd(&p.prof[7], (unsigned long long)ew_rest);
      atomicAdd(&p.prof[8], (unsigned long long)it);
      atomicAdd(&p.prof[9], (unsigned long long)e.store_wait);
    })
    epi_scratch_acquire(e);     // outstanding TMA stores have finished reading this warp's smem
  }

  // teardown: nobody may exit (or free TMEM) while the peer can still signal our barriers / read our smem
  __syncwarp();
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

}  // namespace swb
