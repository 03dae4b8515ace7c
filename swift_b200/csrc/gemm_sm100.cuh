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

namespace swb {

enum GemmEpilogue : int {
  EPI_STORE_F32 = 0,   // out0[M, ldo] fp32
  EPI_STORE_ACT = 1,   // out0[M, ldo] in the 16-bit operand format (fp16 / bf16)
  EPI_EMBED = 2,       // x = acc + bias[n] + pos[row % pos_rows, n];  out0[M, 2N] = [hi | lo] 16-bit pair of x
  EPI_QKV = 3,         // scaled-cosine q/k normalisation fused; out0 = [3][heads][M][96] 16-bit
  EPI_SWIGLU = 4,      // tile = [gate slots | up slots];  out0[M, N/2] 16-bit = silu(gate) * up
  EPI_HEAD = 5,        // pixel-shuffle to NCHW + sampler update: y = alpha*xt + beta*F + gamma*fprev
};

struct GemmParams {
  int M, N, K;
  void* out0;
  void* out1;
  int ldo;                 // row pitch (elements) of out0/out1 for the plain / embed / swiglu epilogues
  // EPI_EMBED
  const float* bias;       // [N]
  const float* pos;        // [pos_rows, N]
  int pos_rows;
  // EPI_QKV
  const float* qscale;     // [heads]  exp(min(scale, ln 100))
  int heads, dmodel;       // N == 3 * dmodel
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

template <int NSUB, int CG>
struct GemmCfg {
  static_assert(NSUB == 1 || (NSUB == 2 && CG == 2), "the 352-wide tile needs the CTA pair");
  static constexpr int kTileN = kUmmaN * NSUB;
  static constexpr int kBRows = kTileN / CG;                   // W rows staged per CTA (one TMA box)
  static constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarps = 4 * NSUB;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static constexpr int kScratchBytes = kEpiWarps * 4096;       // per-warp 32 x 128 B transpose buffer
  static constexpr int kBarBytes = 1024;
  static constexpr int kMaxSmem = 227 * 1024;
  static constexpr int kStagesRaw = (kMaxSmem - kScratchBytes - kBarBytes - 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + kScratchBytes + kBarBytes + 1024;  // +1024 alignment slack
  static_assert(kBBytes % 1024 == 0 && (kUmmaN / CG) * 128 % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
  static_assert(kStages >= 3, "pipeline too shallow");
};

// ---------------------------------------------------------------------------------------------------------
// epilogue helpers

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

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
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / NCH, c = idx - r * NCH;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid && c < chunks_valid) q = __ldg(reinterpret_cast<const uint4*>(g_row0 + r * pitch_bytes + c * 16));
    st_shared_v4(scratch + r * 128 + ((c ^ (r & 7)) << 4), q);
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

template <bool F16, int NW>
__device__ __forceinline__ void pack_row16(const float* v, uint32_t* w) {   // 2*NW floats -> NW packed words
#pragma unroll
  for (int i = 0; i < NW; ++i) w[i] = pack_act2<F16>(v[2 * i], v[2 * i + 1]);
}

// Per-warp epilogue context
struct EpiCtx {
  int row0;          // first row of this warp's 32-row block
  int rows_valid;    // rows of it inside M
  int lane;
  uint32_t scratch;  // this warp's 4 KB transpose buffer
};

// ---- one 88-column slot, fp32 / 16-bit / embed outputs ---------------------------------------------------
template <int EPI, bool F16>
__device__ __forceinline__ void epi_slot_store(const GemmParams& p, const EpiCtx& e, uint32_t tslot, int n0) {
  // blocks of 32, 32, 24 columns
#pragma unroll 1
  for (int blk = 0; blk < 3; ++blk) {
    const int c0 = blk * 32;
    const int n = n0 + c0;
    float v[32];
    if (blk < 2) {
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
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cols_valid) v[j] += __ldg(p.bias + n + j) + pe[j];
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
      if constexpr (EPI == EPI_EMBED) {
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

template <bool F16, typename Release>
__device__ __forceinline__ void epi_group_store16(const GemmParams& p, const EpiCtx& e, uint32_t tacc, int n_lo, int n_hi,
                                                  Release&& release) {
  float v[kSlot];
  uint32_t wa[44], wb[44];
  tmem_load_cols<kSlot>(tacc, v);
  pack_row16<F16, 44>(v, wa);
  tmem_load_cols<kSlot>(tacc + kSlot, v);
  release();
  pack_row16<F16, 44>(v, wb);
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
  pack_row16<F16, 44>(v, w);
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
__device__ __forceinline__ void epi_group_qkv(const GemmParams& p, const EpiCtx& e, uint32_t tacc, int n_lo, int n_hi,
                                              Release&& release) {
  float v[kSlot];
  uint32_t wa[48], wb[48];
  int sa, sb;
  tmem_load_cols<kSlot>(tacc, v);
  const bool oka = qkv_slot_pack<F16>(p, n_lo, v, wa, &sa);
  tmem_load_cols<kSlot>(tacc + kSlot, v);
  release();
  const bool okb = qkv_slot_pack<F16>(p, n_hi, v, wb, &sb);
  if (oka) qkv_slot_store(p, e, sa, wa);
  if (okb) qkv_slot_store(p, e, sb, wb);
}

template <bool F16, typename Release>
__device__ __forceinline__ void epi_group_swiglu(const GemmParams& p, const EpiCtx& e, uint32_t tacc, int out_col0,
                                                 bool valid, Release&& release) {
  // gate slot at columns [0,88), up slot at [88,176) of this group's accumulator.  The gate is parked in registers as
  // 16-bit pairs (one extra 2^-12 rounding in front of a result that is itself stored in 16 bits).
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
  if (valid) store_slot16(p.out0, p.ldo, p.ldo, e, out_col0, w);
}

// ---------------------------------------------------------------------------------------------------------

template <int NSUB, int CG, int EPI, bool F16>
__global__ void __launch_bounds__(GemmCfg<NSUB, CG>::kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using S = GemmCfg<NSUB, CG>;
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
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int tail_k = p.K - (num_kb - 1) * kBlockK;
  const int tail_ksteps = (tail_k + kUmmaK - 1) / kUmmaK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * CG);   // one arrive per epilogue warp (of one group) of every CTA in the pair
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

  if (warp == 0) {
    // ===================================== TMA producer (whole warp converged, one elected lane issues) ==========
    const uint32_t full_leader0 = (CG == 2) ? mapa_u32(full_bar(0), 0) : full_bar(0);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      const int row0 = tm * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM;
      const int col0 = tn * kTileN + static_cast<int>(cta_rank) * S::kBRows;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u, 1);
        if (cta_rank == 0) mbar_arrive_expect_tx_elect(full_bar(stage), S::kStageBytes * CG);
        const uint32_t a_dst = smem_a + stage * S::kABytes;
        const uint32_t b_dst = smem_b + stage * S::kBBytes;
        if constexpr (CG == 2) {
          const uint32_t bar = full_leader0 + 8u * stage;
          tma_load_2d_pair_elect(a_dst, &tmap_a, bar, kb * kBlockK, row0);
          tma_load_2d_pair_elect(b_dst, &tmap_b, bar, kb * kBlockK, col0);
        } else {
          tma_load_2d_elect(a_dst, &tmap_a, full_bar(stage), kb * kBlockK, row0);
          tma_load_2d_elect(b_dst, &tmap_b, full_bar(stage), kb * kBlockK, col0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
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
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const uint32_t par = static_cast<uint32_t>(it) & 1u;
        uint32_t tmem_d;
        if constexpr (NSUB == 1) {
          // double-buffered accumulator: stage it&1, re-used every second tile
          mbar_wait(tempty_bar(par), ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u, 2);
          tcgen05_fence_after();
          tmem_d = tmem_base + par * kSubStride;
        } else {
          tmem_d = tmem_base;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 3);             // TMA bytes of this stage have landed (both CTAs)
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
            if (kb == 0) {
              // first k-block of a tile: each sub-accumulator becomes writable as soon as its own epilogue group
              // of the previous tile has drained it, so start on sub 0 while sub 1 may still be read
              const int nk = full_block ? kKSteps : tail_ksteps;
#pragma unroll 1
              for (int j = 0; j < NSUB; ++j) {
                mbar_wait(tempty_bar(j), par ^ 1u, 2);
                tcgen05_fence_after();
                for (int k = 0; k < nk; ++k)
                  umma_f16_ss_elect<CG>(tmem_d + j * kSubStride, adesc0 + 2u * k, bdesc0 + j * kSubDescStep + 2u * k,
                                        idesc, k != 0 ? 1u : 0u);
              }
            } else if (full_block) {
#pragma unroll
              for (int k = 0; k < kKSteps; ++k) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j)
                  umma_f16_ss_elect<CG>(tmem_d + j * kSubStride, adesc0 + 2u * k, bdesc0 + j * kSubDescStep + 2u * k,
                                        idesc, 1u);
              }
            } else {
              for (int k = 0; k < tail_ksteps; ++k)
                for (int j = 0; j < NSUB; ++j)
                  umma_f16_ss_elect<CG>(tmem_d + j * kSubStride, adesc0 + 2u * k, bdesc0 + j * kSubDescStep + 2u * k,
                                        idesc, 1u);
            }
          }
          umma_commit_elect<CG>(empty_bar(stage));          // smem stage reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_elect<CG>(tfull_bar(NSUB == 1 ? par : 0));   // accumulator(s) complete -> epilogue (both CTAs)
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    const int quad = warp & 3;                              // TMEM lane quadrant this warp may access
    const int grp = (warp - 4) >> 2;                        // NSUB == 2: epilogue group = sub-tile it drains
    const uint32_t tempty_leader0 = (CG == 2) ? mapa_u32(tempty_bar(0), 0) : tempty_bar(0);
    EpiCtx e;
    e.lane = lane;
    e.scratch = scratch0 + (warp - 4) * 4096;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const uint32_t par = static_cast<uint32_t>(it) & 1u;
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      const int n_tile = tn * kTileN;
      uint32_t tacc;                                        // column 0 of the accumulator this warp drains
      if constexpr (NSUB == 1) {
        mbar_wait(tfull_bar(par), (static_cast<uint32_t>(it) >> 1) & 1u, 4);
        tacc = tmem_base + par * kSubStride;
      } else {
        mbar_wait(tfull_bar(0), par, 4);
        tacc = tmem_base + grp * kSubStride;
      }
      tcgen05_fence_after();
      tacc += static_cast<uint32_t>(quad * 32) << 16;
      e.row0 = tm * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM + quad * 32;
      const int rv = p.M - e.row0;
      e.rows_valid = rv < 0 ? 0 : (rv > 32 ? 32 : rv);
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
      };
      const int n_lo = n_tile + s_lo * kSlot, n_hi = n_tile + s_hi * kSlot;
      if constexpr (EPI == EPI_SWIGLU) {
        // packed w1 rows per tile: [NSUB gate slots | NSUB up slots]
        epi_group_swiglu<F16>(p, e, tacc, (tn * NSUB + s_lo) * kSlot, n_tile < p.N, release);
      } else if constexpr (EPI == EPI_QKV) {
        epi_group_qkv<F16>(p, e, tacc, n_lo, n_hi, release);
      } else if constexpr (EPI == EPI_STORE_ACT) {
        epi_group_store16<F16>(p, e, tacc, n_lo, n_hi, release);
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
    }
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
