// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (A and W both fp16 or both bf16, fp32 accumulate in TMEM)
// with the epilogues the Swift denoiser needs fused in.  A and W are both K-major (row-major [rows, K]), which is
// how activations and nn.Linear weights are stored, so neither is ever transposed in memory.
//
//   CG = 2 (default): the two CTAs of a cluster form one MMA pair (tcgen05 cta_group::2).  The pair owns a
//        256 x BN output tile: each CTA stages its own 128 rows of A and half (BN/2 rows) of the W tile, the
//        leader CTA issues UMMA M=256 and the accumulator rows 0..127 / 128..255 land in each CTA's own TMEM.
//   CG = 1: single-CTA fallback (UMMA M=128), same pipeline, used for bring-up and as a cross-check.
//
//   warp 0   TMA producer   (one elected lane; STAGES-deep smem ring, full/empty mbarriers)
//   warp 1   UMMA issuer    (leader CTA only, one lane; tcgen05.commit releases smem stages / publishes accumulators)
//   warp 2   TMEM allocator
//   warp 3   idle
//   warp 4-7 epilogue       (tcgen05.ld -> registers -> fused math -> global), double-buffered TMEM accumulators so
//                            the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include "ptx.cuh"

namespace swb {

enum GemmEpilogue : int {
  EPI_STORE_F32 = 0,   // out0[M, ldo] fp32
  EPI_STORE_ACT = 1,   // out0[M, ldo] in the activation format (bf16 / fp16)
  EPI_EMBED = 2,       // x = acc + bias[n] + pos[row % pos_rows, n];  out0 = x (fp32), out1 = bf16(x)
  EPI_QKV = 3,         // scaled-cosine q/k normalisation fused; out0 = [3][heads][M][HD_PAD] bf16
  EPI_SWIGLU = 4,      // tile = [gate(BN/2) | up(BN/2)];  out0[M, N/2] bf16 = silu(gate) * up
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
};

constexpr int kBlockM = 128;     // rows of A per CTA
constexpr int kBlockK = 64;      // bf16 per k-block: 128 bytes = one SWIZZLE_128B atom row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;
constexpr int kHeadDim = 88;     // Swift-B head dim; EPI_QKV/EPI_SWIGLU tile = 2 x 88 = 176 columns
constexpr int kHeadDimPad = 96;  // q/k/v rows are stored padded to 96 (K of QK^T must be a multiple of 16)

template <int BN, int CG>
struct GemmSmem {
  static constexpr int kBRows = BN / CG;                       // W rows staged per CTA
  static constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kBarBytes = 1024;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024 alignment slack
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024-byte alignment for SWIZZLE_128B");
};

// ---------------------------------------------------------------------------------------------------------
// epilogues: one thread owns one accumulator row (TMEM lane); `tacc` addresses column 0 of that row's tile.

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

template <int NCOL>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, float* v) {
  // NCOL in {8, 16, 24, ..}: greedy x16 then x8
  constexpr int n16 = NCOL / 16;
  __syncwarp();   // tcgen05.ld is .sync.aligned: the warp must be converged (epilogue guards may diverge lanes)
#pragma unroll
  for (int i = 0; i < n16; ++i) tmem_ld_x16(taddr + 16 * i, v + 16 * i);
  if constexpr (NCOL % 16 == 8) tmem_ld_x8(taddr + 16 * n16, v + 16 * n16);
  tmem_ld_wait();
  tmem_ld_fence_regs<NCOL>(v);
}

template <int BN, int EPI, bool F16>
__device__ __forceinline__ void gemm_epilogue_row(const GemmParams& p, uint32_t tacc, int row, int n0) {
  const bool row_ok = row < p.M;
  if constexpr (EPI == EPI_STORE_F32 || EPI == EPI_STORE_ACT || EPI == EPI_EMBED) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_load_cols<16>(tacc + c, v);
      const int n = n0 + c;
      if (!row_ok || n >= p.N) continue;
      if constexpr (EPI == EPI_EMBED) {
        const float* pr = p.pos + static_cast<size_t>(row % p.pos_rows) * p.N + n;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n + j < p.N) v[j] += __ldg(p.bias + n + j) + __ldg(pr + j);
      }
      const size_t off = static_cast<size_t>(row) * p.ldo + n;
      if (n + 16 <= p.N) {
        if constexpr (EPI != EPI_STORE_ACT) {
          float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out0) + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if constexpr (EPI != EPI_STORE_F32) {
          void* dst = (EPI == EPI_EMBED) ? p.out1 : p.out0;
          uint4* o = reinterpret_cast<uint4*>(static_cast<uint16_t*>(dst) + off);
          o[0] = make_uint4(pack_act2<F16>(v[0], v[1]), pack_act2<F16>(v[2], v[3]), pack_act2<F16>(v[4], v[5]),
                            pack_act2<F16>(v[6], v[7]));
          o[1] = make_uint4(pack_act2<F16>(v[8], v[9]), pack_act2<F16>(v[10], v[11]), pack_act2<F16>(v[12], v[13]),
                            pack_act2<F16>(v[14], v[15]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (n + j < p.N) {
            if constexpr (EPI != EPI_STORE_ACT) static_cast<float*>(p.out0)[off + j] = v[j];
            if constexpr (EPI != EPI_STORE_F32) {
              void* dst = (EPI == EPI_EMBED) ? p.out1 : p.out0;
              static_cast<uint16_t*>(dst)[off + j] = pack_act1<F16>(v[j]);
            }
          }
        }
      }
    }
  } else if constexpr (EPI == EPI_QKV) {
    // packed weight rows: n = part*dmodel + head*88 + d  (part 0 = q, 1 = k, 2 = v); a tile holds two 88-wide
    // (part, head) slots, which may straddle a part boundary when the head count is odd.
    static_assert(BN == 2 * kHeadDim, "EPI_QKV needs a 176-column tile");
    const int slot0 = n0 / kHeadDim;                       // (part, head) slots are 88 columns wide
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      float v[kHeadDim];
      tmem_load_cols<kHeadDim>(tacc + hh * kHeadDim, v);
      const int slot = slot0 + hh;
      const int part = slot / p.heads;
      const int head = slot - part * p.heads;
      if (!row_ok || part >= 3) continue;
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
      uint16_t* dst = static_cast<uint16_t*>(p.out0) +
                           (static_cast<size_t>(part * p.heads + head) * p.M + row) * kHeadDimPad;
      uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int j = 0; j < kHeadDim / 8; ++j)
        o[j] = make_uint4(pack_act2<F16>(v[8 * j], v[8 * j + 1]), pack_act2<F16>(v[8 * j + 2], v[8 * j + 3]),
                          pack_act2<F16>(v[8 * j + 4], v[8 * j + 5]), pack_act2<F16>(v[8 * j + 6], v[8 * j + 7]));
      o[kHeadDim / 8] = make_uint4(0u, 0u, 0u, 0u);   // zero pad 88..95
    }
  } else if constexpr (EPI == EPI_SWIGLU) {
    // packed weight rows: tile j = [gate rows j*88..j*88+87 | up rows j*88..]; N = 2*Dff.
    static_assert(BN == 2 * kHeadDim, "EPI_SWIGLU needs a 176-column tile");
    constexpr int HB = BN / 2;
    const int o0 = (n0 / BN) * HB;
    uint16_t* dst = static_cast<uint16_t*>(p.out0) + static_cast<size_t>(row) * p.ldo + o0;
#pragma unroll 1
    for (int c = 0; c < HB; c += 16) {
      float g[16], u[16];
      __syncwarp();
      if (c + 16 <= HB) {
        tmem_ld_x16(tacc + c, g);
        tmem_ld_x16(tacc + HB + c, u);
        tmem_ld_wait();
        tmem_ld_fence_regs<16>(g);
        tmem_ld_fence_regs<16>(u);
        if (row_ok && n0 < p.N) {
          uint4* o = reinterpret_cast<uint4*>(dst + c);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            w[j] = pack_act2<F16>(silu_f(g[2 * j]) * u[2 * j], silu_f(g[2 * j + 1]) * u[2 * j + 1]);
          o[0] = make_uint4(w[0], w[1], w[2], w[3]);
          o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      } else {
        tmem_ld_x8(tacc + c, g);
        tmem_ld_x8(tacc + HB + c, u);
        tmem_ld_wait();
        tmem_ld_fence_regs<8>(g);
        tmem_ld_fence_regs<8>(u);
        if (row_ok && n0 < p.N) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            w[j] = pack_act2<F16>(silu_f(g[2 * j]) * u[2 * j], silu_f(g[2 * j + 1]) * u[2 * j + 1]);
          *reinterpret_cast<uint4*>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  } else if constexpr (EPI == EPI_HEAD) {
    // packed column order is the reference's "(c p1 p2)" (models/swinv2.py:242); write NCHW directly.
    const int b = row / p.tokens;
    const int tok = row - b * p.tokens;
    const int gy = tok / p.gw, gx = tok - gy * p.gw;
    const int pp = p.p1 * p.p2;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_load_cols<16>(tacc + c, v);
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c + j;
        if (n < p.N) {
          const int ch = n / pp, r = n - ch * pp;
          const int py = r / p.p2, px = r - py * p.p2;
          const size_t a =
              ((static_cast<size_t>(b) * p.C + ch) * p.H + (gy * p.p1 + py)) * p.W + (gx * p.p2 + px);
          float y = p.beta * v[j];
          if (p.xt) y = fmaf(p.alpha, __ldg(p.xt + a), y);
          if (p.fprev) y = fmaf(p.gamma, __ldg(p.fprev + a), y);
          static_cast<float*>(p.out0)[a] = y;
          if (p.out_f) p.out_f[a] = v[j];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------

template <int BN, int CG, int EPI, bool F16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using S = GemmSmem<BN, CG>;
  constexpr int kStages = S::kStages;
  static_assert(BN % 16 == 0 && (BN / CG) % 8 == 0 && BN <= 256, "UMMA N constraint");
  static_assert(BN <= kAccStride, "accumulator stage stride");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStages * S::kABytes;
  const uint32_t bars = smem_base + kStages * S::kStageBytes;
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

  const int tiles_n = (p.N + BN - 1) / BN;
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
      mbar_init(tempty_bar(s), 4 * CG);   // one arrive per epilogue warp of every CTA in the pair
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
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      const uint32_t full_leader0 = (CG == 2) ? mapa_u32(full_bar(0), 0) : full_bar(0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
        const int row0 = tm * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM;
        const int col0 = tn * BN + static_cast<int>(cta_rank) * S::kBRows;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 1);
          if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), S::kStageBytes * CG);
          const uint32_t a_dst = smem_a + stage * S::kABytes;
          const uint32_t b_dst = smem_b + stage * S::kBBytes;
          if constexpr (CG == 2) {
            const uint32_t bar = full_leader0 + 8u * stage;
            tma_load_2d_pair(a_dst, &tmap_a, bar, kb * kBlockK, row0);
            tma_load_2d_pair(b_dst, &tmap_b, bar, kb * kBlockK, col0);
          } else {
            tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * kBlockK, row0);
            tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * kBlockK, col0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================== UMMA issuer (leader CTA) =====================================
    if (cta_rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM * CG, BN, /*A=*/F16, /*B=*/F16);   // mixed A/B formats trap
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);     // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 3);             // TMA bytes of this stage have landed (both CTAs)
          tcgen05_fence_after();
          const uint32_t a_addr = smem_a + stage * S::kABytes;
          const uint32_t b_addr = smem_b + stage * S::kBBytes;
          const int nk = (kb == num_kb - 1) ? tail_ksteps : (kBlockK / kUmmaK);
          for (int k = 0; k < nk; ++k) {
            // K-major SWIZZLE_128B: 8-row groups are 1024 B apart; +16 elements along K = +32 bytes
            const uint64_t adesc = make_smem_desc(a_addr + k * (kUmmaK * 2), 16, 1024, SWZ_128B);
            const uint64_t bdesc = make_smem_desc(b_addr + k * (kUmmaK * 2), 16, 1024, SWZ_128B);
            umma_bf16_ss<CG>(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<CG>(empty_bar(stage));                // smem stage reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit<CG>(tfull_bar(acc));                    // accumulator complete -> epilogue (both CTAs)
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================== epilogue =====================================
    const int quad = warp & 3;                              // TMEM lane quadrant this warp may access
    const uint32_t tempty_leader0 = (CG == 2) ? mapa_u32(tempty_bar(0), 0) : tempty_bar(0);
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      mbar_wait(tfull_bar(acc), acc_phase, 4);
      tcgen05_fence_after();
      const int row = tm * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM + quad * 32 + lane;
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride;
      gemm_epilogue_row<BN, EPI, F16>(p, tacc, row, tn * BN);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc);
        else mbar_arrive(tempty_bar(acc));
      }
    }
  }

  // teardown: nobody may exit (or free TMEM) while the peer can still signal our barriers / read our smem
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

}  // namespace swb
