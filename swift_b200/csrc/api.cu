// extern "C" layer: argument validation, workspace carving and the kernel sequence of one denoiser forward.
#include "../../include/swift_b200.h"

#include "common.h"
#include "gemm_sm100.cuh"
#include "kernels.h"

#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

using namespace swb;

namespace {

// ------------------------------------------------------------------------------------------------
// Optional in-situ tracing (swb200_trace_enable): CUDA events around every kernel of swb200_forward, so per-kernel
// durations can be read at the clocks of the real, power-capped step instead of under a profiler.  Off by default
// (and illegal during graph capture).
enum TraceSlot { T_GATHER = 0, T_EMBED, T_QKV, T_ATTN, T_WO, T_LN, T_W1, T_W2, T_HEAD, T_NSLOTS };
const char* const kTraceNames[T_NSLOTS] = {"patch_gather", "gemm_embed", "gemm_qkv", "window_attention", "gemm_wo",
                                           "ln_mod_residual", "gemm_w1_swiglu", "gemm_w2", "gemm_head"};
struct Trace {
  bool on = false;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev;
  double ms[T_NSLOTS] = {0};
  long n[T_NSLOTS] = {0};
} g_trace;

struct TraceScope {
  int slot;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  TraceScope(int s, cudaStream_t stream) : slot(s), st(stream) {
    if (g_trace.on) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, st);
    }
  }
  ~TraceScope() {
    if (e0) {
      cudaEventRecord(e1, st);
      g_trace.ev.push_back({slot, {e0, e1}});
    }
  }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// fp16 range diagnostics (swb200_debug_saturation): device counters, one per tensor class of swb200_forward
enum SatSlot { SAT_X_HI = 0, SAT_X_LO, SAT_QKV, SAT_ATTN, SAT_BRANCH, SAT_H, SAT_NSLOTS };
unsigned long long* g_sat_counters = nullptr;

struct Geom {
  int gh, gw, tokens, pp, k_embed_total, k_head_total;
};
Geom geom(const swb200_model* m) {
  Geom g;
  g.gh = m->img_h / m->patch_h;
  g.gw = m->img_w / m->patch_w;
  g.tokens = g.gh * g.gw;
  g.pp = m->patch_h * m->patch_w;
  g.k_embed_total = m->k_embed * (1 + (m->split_embed ? 1 : 0));
  g.k_head_total = m->dim * (1 + (m->split_head ? 1 : 0));
  return g;
}

// per-sample workspace layout (offsets in bytes, every buffer 1024-byte aligned per chunk)
struct Workspace {
  size_t xhl, qkv, attn, branch, h, lnws, total;   // sizes for `chunk` samples; a_embed aliases qkv
};

// LayerNorm statistics exchange of the fused wo / w2 epilogue: [counters: Mpad/32 u32, padded to 1 KB][stats: groups x Mpad float2]
struct LnWs {
  size_t counters_bytes, total;
  int stride, groups;
};
LnWs ln_ws_layout(int M, int dim) {
  LnWs l;
  l.stride = (M + 255) / 256 * 256;
  l.groups = (dim + kUmmaN - 1) / kUmmaN + 1;             // enough for every tile configuration
  l.counters_bytes = align_up(static_cast<size_t>(l.stride / 32) * sizeof(unsigned), 1024);
  l.total = l.counters_bytes + static_cast<size_t>(l.groups) * l.stride * sizeof(float2);
  return l;
}
Workspace carve(const swb200_model* m, int chunk) {
  const Geom g = geom(m);
  const size_t M = static_cast<size_t>(chunk) * g.tokens;
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  w.xhl = take(M * m->dim * 2 * 2);          // residual stream as a 16-bit [hi | lo] pair
  size_t qkv_bytes = static_cast<size_t>(3) * m->heads * M * kHeadDimPad * 2;
  size_t emb_bytes = M * g.k_embed_total * 2;
  w.qkv = take(qkv_bytes > emb_bytes ? qkv_bytes : emb_bytes);
  w.attn = take(M * m->dim * 2);
  w.branch = take(M * m->dim * 4);
  w.h = take(M * static_cast<size_t>(m->dff) * 2);
  w.lnws = take(ln_ws_layout(static_cast<int>(M), m->dim).total);
  w.total = off;
  return w;
}

int validate(const swb200_model* m) {
  SWB_REQUIRE(m != nullptr, "model is NULL");
  SWB_REQUIRE(m->patch_h > 0 && m->patch_w > 0 && m->img_h % m->patch_h == 0 && m->img_w % m->patch_w == 0,
              "image %dx%d not divisible by patch %dx%d", m->img_h, m->img_w, m->patch_h, m->patch_w);
  const Geom g = geom(m);
  SWB_REQUIRE(m->win_h == 16 && m->win_w == 16,
              "only 16x16 windows are implemented (got %dx%d); Swift-B uses 16x16", m->win_h, m->win_w);
  SWB_REQUIRE(g.gh % 16 == 0 && g.gw % 16 == 0, "token grid %dx%d is not a multiple of the 16x16 window", g.gh, g.gw);
  SWB_REQUIRE(m->heads > 0 && m->dim == m->heads * kHeadDim,
              "only head_dim 88 is implemented (dim=%d heads=%d); Swift-B uses 1056/12", m->dim, m->heads);
  SWB_REQUIRE(m->dim % 8 == 0 && m->dff > 0 && m->dff % kHeadDim == 0, "mlp dim %d must be a multiple of 88", m->dff);
  SWB_REQUIRE(m->k_embed % 8 == 0 && m->k_embed >= m->in_channels * g.pp, "k_embed=%d invalid for %d input features",
              m->k_embed, m->in_channels * g.pp);
  SWB_REQUIRE(m->shift_h >= 0 && m->shift_w >= 0 && m->shift_h < 16 && m->shift_w < 16, "bad shift %d,%d", m->shift_h,
              m->shift_w);
  SWB_REQUIRE(m->depth > 0 && m->aux_dim >= 0, "bad depth/aux_dim");
  SWB_REQUIRE(m->gemm_tile >= 1 && m->gemm_tile <= 3, "gemm_tile must be 1, 2 or 3 (got %d)", m->gemm_tile);
  SWB_REQUIRE(m->attn_impl >= 0 && m->attn_impl <= 2, "attn_impl must be 0, 1 or 2 (got %d)", m->attn_impl);
  SWB_REQUIRE(m->fuse_ln >= 0 && m->fuse_ln <= 3 && (m->fuse_ln == 0 || m->dim <= 12 * kUmmaN),
              "fuse_ln must be 0..3 (bit 0: wo, bit 1: w2) and needs dim <= 2112 (got %d, dim %d)", m->fuse_ln, m->dim);
  SWB_REQUIRE(m->gemm_tile != 3 || m->dff % (2 * kHeadDim) == 0,
              "gemm_tile 3 (256x352) needs mlp dim %d to be a multiple of 176", m->dff);
  return SWB_OK;
}

GemmParams base_params(int M, int N, int K) {
  GemmParams p = {};
  p.M = M;
  p.N = N;
  p.K = K;
  return p;
}

}  // namespace

extern "C" {

SWB200_API int swb200_abi_version(void) { return SWB200_ABI_VERSION; }
SWB200_API const char* swb200_last_error(void) { return get_error(); }
SWB200_API int swb200_validate(const swb200_model* m) { return validate(m); }

SWB200_API size_t swb200_workspace_bytes(const swb200_model* m, int chunk) {
  if (validate(m) != SWB_OK || chunk <= 0) return 0;
  return carve(m, chunk).total;
}

SWB200_API size_t swb200_conditioning_scratch_bytes(const swb200_model* m, int B) {
  if (m == nullptr || B <= 0) return 0;
  const size_t L = 2 * static_cast<size_t>(m->depth);
  return static_cast<size_t>(B) * (3 * m->dim + L * 2 * m->dim) * sizeof(float);
}

SWB200_API int swb200_conditioning(const swb200_model* m, const float* t, const float* aux, int B, float* gain, float* bias,
                        float* cond_out, void* scratch, size_t scratch_bytes, void* stream) {
  int rc = validate(m);
  if (rc) return rc;
  SWB_REQUIRE(B > 0 && t && gain && bias && scratch, "conditioning: NULL argument or B=%d", B);
  SWB_REQUIRE(scratch_bytes >= swb200_conditioning_scratch_bytes(m, B), "conditioning: scratch too small");
  CondWeights w;
  w.aux_w = m->aux_w;
  w.aux_b = m->aux_b;
  w.aux_dim = m->aux_dim;
  w.l1_w = m->l1_w;
  w.l1_b = m->l1_b;
  w.l2_w = m->l2_w;
  w.l2_b = m->l2_b;
  w.mod_w = m->mod_w;
  w.mod_b = m->mod_b;
  w.ln_gamma = m->ln_gamma;
  w.ln_beta = m->ln_beta;
  const float* aux_eff = (m->aux_dim > 0 && m->aux_w != nullptr) ? aux : nullptr;
  return launch_conditioning(w, t, aux_eff, B, m->dim, 2 * m->depth, m->timestep_weight,
                             static_cast<float*>(scratch), gain, bias, cond_out, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_forward(const swb200_model* m, const float* x0, int c0, float scale0, const float* x1, int c1, int B,
                   const float* gain, const float* bias, const swb200_update* upd, float* y, void* workspace,
                   size_t workspace_bytes, void* stream_) {
  int rc = validate(m);
  if (rc) return rc;
  SWB_REQUIRE(B > 0 && x0 && gain && bias && upd && workspace, "forward: NULL argument or B=%d", B);
  SWB_REQUIRE(y != nullptr || upd->state != nullptr, "forward: y may only be NULL in rollout mode (upd->state set)");
  SWB_REQUIRE(upd->state == nullptr || (upd->x_std && upd->x_mean && upd->d_std && upd->state_channels >= m->out_channels),
              "forward: rollout mode needs x_std / x_mean / d_std and state_channels >= out_channels");
  SWB_REQUIRE(c0 + c1 == m->in_channels && (c1 == 0 || x1 != nullptr), "forward: c0+c1=%d != in_channels=%d", c0 + c1,
              m->in_channels);
  SWB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "forward: workspace must be 1024-byte aligned");
  const size_t per1 = carve(m, 1).total;
  int chunk = static_cast<int>(workspace_bytes / per1);
  SWB_REQUIRE(chunk >= 1, "forward: workspace of %zu bytes cannot hold one sample (%zu bytes)", workspace_bytes, per1);
  if (chunk > B) chunk = B;
  while (carve(m, chunk).total > workspace_bytes) --chunk;   // alignment slack
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, H = m->heads, Dff = m->dff;
  const int F16 = m->act_fp16 ? 1 : 0;
  // q / k / v and P in fp16 even when the GEMM operands are bf16: the attention kernel's operand format is independent of
  // the GEMMs' (its inputs come out of an epilogue, its output goes into one), and q_hat*scale <= 100, k_hat <= 1, P <= 1
  const int AF16 = (F16 || m->attn_fp16) ? 1 : 0;
  const int kDefaultCG = m->gemm_tile;          // tile config of every GEMM (the w1 packing depends on it)
  const size_t img_in0 = static_cast<size_t>(c0) * m->img_h * m->img_w;
  const size_t img_in1 = static_cast<size_t>(c1) * m->img_h * m->img_w;
  const size_t img_out = static_cast<size_t>(m->out_channels) * m->img_h * m->img_w;
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int bc = (B - b0 < chunk) ? (B - b0) : chunk;
    const int M = bc * g.tokens;
    const Workspace w = carve(m, bc);
    void* xhl = ws + w.xhl;                    // [M, 2D]: hi (= GEMM A operand, row pitch 2D) | lo
    void* qkv = ws + w.qkv;
    void* a_emb = qkv;
    void* attn = ws + w.attn;
    void* branch = ws + w.branch;
    // fp16 mode: the wo / w2 branch outputs are stored in fp16 (saturated); bf16 mode keeps them in fp32 because the
    // extra bf16 rounding in front of the LayerNorm costs ~10 % accuracy
    const int BR16 = F16;
    const int epi_branch = BR16 ? EPI_STORE_ACT : EPI_STORE_F32;
    void* hbuf = ws + w.h;
    // the fused epilogue spins on statistics published by other CTAs of its grid: if the device cannot keep the whole grid
    // resident (reported before anything is launched) the same update runs as GEMM + LayerNorm kernel instead
    bool fuse_wo = (m->fuse_ln & 1) != 0, fuse_w2 = (m->fuse_ln & 2) != 0;
    // debug only: count fp16 values sitting at the saturation value after the kernel that produced them
    auto sat = [&](int slot, const void* buf, long long rows, int cols, long long pitch) -> int {
      if (g_sat_counters == nullptr) return SWB_OK;
      return launch_count_saturated_f16(buf, rows, cols, pitch, g_sat_counters + slot, stream);
    };
    auto sat_x = [&]() -> int {
      if (g_sat_counters == nullptr || !F16) return SWB_OK;
      int r = sat(SAT_X_HI, xhl, M, D, 2 * D);
      return r ? r : sat(SAT_X_LO, static_cast<const uint16_t*>(xhl) + D, M, D, 2 * D);
    };
    void* lnws = ws + w.lnws;
    int ln_gen = 0;                             // fused launches of this chunk so far (the first one clears the counters)

    // 1. concat + patchify + cast
    { TraceScope ts_(T_GATHER, stream);
    rc = launch_patch_gather(x0 + b0 * img_in0, c0, scale0, x1 ? x1 + b0 * img_in1 : nullptr, c1, a_emb,
                             g.k_embed_total, m->k_embed, m->split_embed, F16, bc, m->img_h, m->img_w, m->patch_h,
                             m->patch_w, stream); }
    if (rc) return rc;
    // 2. patch-embed GEMM (+bias +pos_embed)
    {
      GemmParams p = base_params(M, D, g.k_embed_total);
      p.out0 = xhl;
      p.ldo = 2 * D;
      p.bias = m->b_embed;
      p.pos = m->pos_embed;
      p.pos_rows = g.tokens;
      { TraceScope ts_(T_EMBED, stream);
      rc = launch_gemm(EPI_EMBED, kDefaultCG, F16, a_emb, g.k_embed_total, m->w_embed, g.k_embed_total, p, stream); }
      if (rc) return rc;
      if ((rc = sat_x())) return rc;
    }
    // 3. transformer blocks
    for (int l = 0; l < m->depth; ++l) {
      const bool shifted = (m->shift_h || m->shift_w) && (l & 1);
      {
        GemmParams p = base_params(M, 3 * D, D);
        p.out0 = qkv;
        p.qscale = m->qscale + static_cast<size_t>(l) * H;
        p.heads = H;
        p.dmodel = D;
        p.qkv_f16 = AF16;
        const auto* wq = static_cast<const __nv_bfloat16*>(m->w_qkv) + static_cast<size_t>(l) * 3 * D * D;
        { TraceScope ts_(T_QKV, stream);
        rc = launch_gemm(EPI_QKV, kDefaultCG, F16, xhl, 2 * D, wq, D, p, stream); }
        if (rc) return rc;
        if (AF16 && (rc = sat(SAT_QKV, qkv, 3LL * H * M, kHeadDimPad, kHeadDimPad))) return rc;
      }
      { TraceScope ts_(T_ATTN, stream);
      rc = launch_window_attention(qkv, attn, bc, g.gh, g.gw, H, shifted ? m->shift_h : 0, shifted ? m->shift_w : 0,
                                   AF16, F16, m->attn_impl, stream); }
      if (rc) return rc;
      if (F16 && (rc = sat(SAT_ATTN, attn, M, D, D))) return rc;
      const auto* wo = static_cast<const __nv_bfloat16*>(m->w_o) + static_cast<size_t>(l) * D * D;
      const float* gain_a = gain + (static_cast<size_t>(2 * l) * B + b0) * D;
      const float* bias_a = bias + (static_cast<size_t>(2 * l) * B + b0) * D;
      if (fuse_wo) {
        // wo projection + LayerNorm + modulation + residual add in one kernel (no branch buffer)
        { TraceScope ts_(T_WO, stream);
        rc = swb200_gemm_ln_residual(kDefaultCG, F16, attn, D, wo, D, xhl, gain_a, bias_a, M, D, g.tokens, lnws, ln_gen,
                                     stream_); }
        if (rc == SWB_ERR_RESIDENCY) fuse_wo = fuse_w2 = false;
        else if (rc) return rc;
        else ++ln_gen;
      }
      if (!fuse_wo) {
        GemmParams p = base_params(M, D, D);
        p.out0 = branch;
        p.ldo = D;
        { TraceScope ts_(T_WO, stream);
        rc = launch_gemm(epi_branch, kDefaultCG, F16, attn, D, wo, D, p, stream); }
        if (rc) return rc;
        if (BR16 && (rc = sat(SAT_BRANCH, branch, M, D, D))) return rc;
        { TraceScope ts_(T_LN, stream);
        rc = launch_ln_mod_residual(branch, BR16, xhl, gain_a, bias_a, M, D, g.tokens, 1e-6f, F16, stream); }
        if (rc) return rc;
      }
      if ((rc = sat_x())) return rc;
      {
        GemmParams p = base_params(M, 2 * Dff, D);
        p.out0 = hbuf;
        p.ldo = Dff;
        const auto* w1 = static_cast<const __nv_bfloat16*>(m->w_1) + static_cast<size_t>(l) * 2 * Dff * D;
        { TraceScope ts_(T_W1, stream);
        rc = launch_gemm(EPI_SWIGLU, kDefaultCG, F16, xhl, 2 * D, w1, D, p, stream); }
        if (rc) return rc;
        if (F16 && (rc = sat(SAT_H, hbuf, M, Dff, Dff))) return rc;
      }
      const auto* w2 = static_cast<const __nv_bfloat16*>(m->w_2) + static_cast<size_t>(l) * D * Dff;
      const float* gain_f = gain + (static_cast<size_t>(2 * l + 1) * B + b0) * D;
      const float* bias_f = bias + (static_cast<size_t>(2 * l + 1) * B + b0) * D;
      if (fuse_w2) {
        { TraceScope ts_(T_W2, stream);
        rc = swb200_gemm_ln_residual(kDefaultCG, F16, hbuf, Dff, w2, Dff, xhl, gain_f, bias_f, M, D, g.tokens, lnws,
                                     ln_gen, stream_); }
        if (rc == SWB_ERR_RESIDENCY) fuse_wo = fuse_w2 = false;
        else if (rc) return rc;
        else ++ln_gen;
      }
      if (!fuse_w2) {
        GemmParams p = base_params(M, D, Dff);
        p.out0 = branch;
        p.ldo = D;
        { TraceScope ts_(T_W2, stream);
        rc = launch_gemm(epi_branch, kDefaultCG, F16, hbuf, Dff, w2, Dff, p, stream); }
        if (rc) return rc;
        if (BR16 && (rc = sat(SAT_BRANCH, branch, M, D, D))) return rc;
        { TraceScope ts_(T_LN, stream);
        rc = launch_ln_mod_residual(branch, BR16, xhl, gain_f, bias_f, M, D, g.tokens, 1e-6f, F16, stream); }
        if (rc) return rc;
      }
      if ((rc = sat_x())) return rc;
    }
    // 4. head GEMM + pixel shuffle + sampler update
    {
      swb200_update u = *upd;
      if (u.xt) u.xt += b0 * img_out;
      if (u.fprev) u.fprev += b0 * img_out;
      if (u.out_f) u.out_f += b0 * img_out;
      if (u.state) u.state += b0 * static_cast<size_t>(u.state_channels) * m->img_h * m->img_w;
      if (u.phys) u.phys += b0 * img_out;
      // the head reads the residual pair directly: K = 2D ([hi | lo] x [W | W]) or K = D (hi only), row pitch 2D
      { TraceScope ts_(T_HEAD, stream);
      rc = swb200_gemm_head(kDefaultCG, m, xhl, 2 * D, g.k_head_total, bc, &u, y ? y + b0 * img_out : nullptr,
                            stream_); }
      if (rc) return rc;
    }
  }
  return SWB_OK;
}

// ------------------------------------------------------------------------------------------------ single kernels

SWB200_API int swb200_gemm(int epi, int tile, int act_fp16, const void* A, int lda, const void* W, int ldw, void* out,
                int ldo, int M, int N, int K, void* stream) {
  SWB_REQUIRE(epi == EPI_STORE_F32 || epi == EPI_STORE_ACT || (epi >= EPI_DISCARD && epi <= EPI_DIRECT) || epi >= 1000,
              "swb200_gemm: epi must be 0 (fp32), 1 (activation format), 6..9 (profiling variants)");
  SWB_REQUIRE(A && W && out, "swb200_gemm: NULL pointer");
  SWB_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && N % (epi == EPI_STORE_F32 ? 4 : 8) == 0,
              "swb200_gemm: out must be 16-byte aligned, ldo %% 8 == 0, N %% 4 (fp32) / 8 (16-bit) == 0");
  GemmParams p = base_params(M, N, K);
  p.out0 = out;
  p.ldo = ldo;
  if (epi >= 1000) {                          // profiling: EPI_BUSY with (epi - 1000) x 64 FMAs per epilogue thread and tile
    p.heads = (epi - 1000) % 1000;
    p.dmodel = (epi - 1000) / 1000;             // 1: independent FMA chains (full issue rate)
    epi = EPI_BUSY;
  }
  return launch_gemm(epi, tile, act_fp16, A, lda, W, ldw, p, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_gemm_qkv(int tile, int act_fp16, int qkv_fp16, const void* A, int lda, const void* W, const float* qscale,
                    void* out, int M, int dim, int heads, void* stream) {
  SWB_REQUIRE(!act_fp16 || qkv_fp16, "swb200_gemm_qkv: fp16 operands with a bf16 q/k/v output is not a supported combination");
  SWB_REQUIRE(A && W && qscale && out, "swb200_gemm_qkv: NULL pointer");
  SWB_REQUIRE(dim == heads * kHeadDim, "swb200_gemm_qkv: need head_dim 88 (dim=%d heads=%d)", dim, heads);
  GemmParams p = base_params(M, 3 * dim, dim);
  p.out0 = out;
  p.qscale = qscale;
  p.heads = heads;
  p.dmodel = dim;
  p.qkv_f16 = qkv_fp16 ? 1 : 0;
  return launch_gemm(EPI_QKV, tile, act_fp16, A, lda, W, dim, p, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_gemm_swiglu(int tile, int act_fp16, const void* A, int lda, const void* W, void* out, int M, int dim, int dff,
                       void* stream) {
  SWB_REQUIRE(A && W && out, "swb200_gemm_swiglu: NULL pointer");
  SWB_REQUIRE(dff % kHeadDim == 0 && (tile != 3 || dff % (2 * kHeadDim) == 0),
              "swb200_gemm_swiglu: dff must be a multiple of 88 (176 for tile 3)");
  GemmParams p = base_params(M, 2 * dff, dim);
  p.out0 = out;
  p.ldo = dff;
  return launch_gemm(EPI_SWIGLU, tile, act_fp16, A, lda, W, dim, p, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_gemm_embed(int tile, int act_fp16, const void* A, int lda, const void* W, int K, const float* bias,
                      const float* pos, int tokens, void* xhl, int M, int dim, void* stream) {
  SWB_REQUIRE(A && W && pos && xhl, "swb200_gemm_embed: NULL pointer");
  GemmParams p = base_params(M, dim, K);
  p.out0 = xhl;
  p.ldo = 2 * dim;
  p.bias = bias;
  p.pos = pos;
  p.pos_rows = tokens;
  return launch_gemm(EPI_EMBED, tile, act_fp16, A, lda, W, K, p, static_cast<cudaStream_t>(stream));
}

SWB200_API size_t swb200_ln_workspace_bytes(int M, int dim) {
  if (M <= 0 || dim <= 0) return 0;
  return ln_ws_layout(M, dim).total;
}

SWB200_API int swb200_gemm_ln_residual(int tile, int act_fp16, const void* A, int lda, const void* W, int K, void* xhl,
                                       const float* gain, const float* bias, int M, int dim, int tokens, void* ln_ws, int gen,
                                       void* stream_) {
  SWB_REQUIRE(A && W && xhl && gain && bias && ln_ws, "swb200_gemm_ln_residual: NULL pointer");
  SWB_REQUIRE(dim % kSlot == 0 && tokens > 0 && tokens % 32 == 0 && gen >= 0,
              "swb200_gemm_ln_residual: dim %d must be a multiple of 88, tokens %d a multiple of 32, gen >= 0", dim, tokens);
  SWB_REQUIRE(((reinterpret_cast<uintptr_t>(xhl) | reinterpret_cast<uintptr_t>(gain) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(ln_ws) & 255) == 0,
              "swb200_gemm_ln_residual: xhl / gain / bias must be 16-byte aligned, ln_ws 256-byte aligned");
  SWB_REQUIRE(tile >= 1 && tile <= 3, "swb200_gemm_ln_residual: tile config must be 1, 2 or 3 (got %d)", tile);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const LnWs l = ln_ws_layout(M, dim);
  if (gen == 0) SWB_CHECK_CUDA(cudaMemsetAsync(ln_ws, 0, l.counters_bytes, stream));
  const int tile_n = tile == 3 ? 2 * kUmmaN : kUmmaN;
  const int groups = (dim + tile_n - 1) / tile_n * (tile == 3 ? 2 : 1);
  SWB_REQUIRE(groups <= 12, "swb200_gemm_ln_residual: dim %d needs %d statistics groups (at most 12 are supported)", dim, groups);
  GemmParams p = base_params(M, dim, K);
  p.xhl = static_cast<uint16_t*>(xhl);
  p.gain = gain;
  p.lnbias = bias;
  p.tokens = tokens;
  p.ln_counter = static_cast<unsigned*>(ln_ws);
  p.ln_stats = reinterpret_cast<float2*>(static_cast<uint8_t*>(ln_ws) + l.counters_bytes);
  p.ln_stride = l.stride;
  p.ln_target = static_cast<unsigned>(groups) * static_cast<unsigned>(gen + 1);
  p.ln_eps = 1e-6f;
  static const int dbg = getenv("SWB_LN_DEBUG") ? atoi(getenv("SWB_LN_DEBUG")) : 0;   // profiling knob, see GemmParams
  p.ln_debug = dbg;
  return launch_gemm(EPI_LN_RES, tile, act_fp16, A, lda, W, K, p, stream);
}

SWB200_API int swb200_gemm_head(int tile, const swb200_model* m, const void* A, int lda, int K, int B,
                     const swb200_update* upd, float* y, void* stream) {
  SWB_REQUIRE(m && A && upd && (y || upd->state), "swb200_gemm_head: NULL pointer");
  const Geom g = geom(m);
  GemmParams p = base_params(B * g.tokens, m->out_channels * g.pp, K);
  p.out0 = y;
  p.xt = upd->xt;
  p.fprev = upd->fprev;
  p.out_f = upd->out_f;
  p.alpha = upd->alpha;
  p.beta = upd->beta;
  p.gamma = upd->gamma;
  p.C = m->out_channels;
  p.H = m->img_h;
  p.W = m->img_w;
  p.p1 = m->patch_h;
  p.p2 = m->patch_w;
  p.gw = g.gw;
  p.tokens = g.tokens;
  p.state = upd->state;
  p.state_C = upd->state_channels;
  p.x_std = upd->x_std;
  p.x_mean = upd->x_mean;
  p.d_std = upd->d_std;
  p.phys = upd->phys;
  p.zero_channel = upd->state ? upd->zero_channel : -1;
  return launch_gemm(EPI_HEAD, tile, m->act_fp16 ? 1 : 0, A, lda, m->w_head, K, p, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_patch_gather(const swb200_model* m, const float* x0, int c0, float scale0, const float* x1, int c1, int B,
                        void* A, int lda, void* stream) {
  SWB_REQUIRE(m && x0 && A, "swb200_patch_gather: NULL pointer");
  SWB_REQUIRE(c0 + c1 == m->in_channels, "swb200_patch_gather: c0+c1 != in_channels");
  return launch_patch_gather(x0, c0, scale0, x1, c1, A, lda, m->k_embed, m->split_embed, m->act_fp16 ? 1 : 0, B, m->img_h, m->img_w,
                             m->patch_h, m->patch_w, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_ln_mod_residual(const void* branch, int branch_16bit, void* xhl, const float* gain, const float* bias,
                           int M, int dim, int tokens, int act_fp16, void* stream) {
  SWB_REQUIRE(branch && xhl && gain && bias, "swb200_ln_mod_residual: NULL pointer");
  return launch_ln_mod_residual(branch, branch_16bit, xhl, gain, bias, M, dim, tokens, 1e-6f, act_fp16,
                                static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_window_attention(const void* qkv, void* out, int B, int grid_h, int grid_w, int heads, int shift_h,
                            int shift_w, int qkv_fp16, int out_fp16, int impl, void* stream) {
  SWB_REQUIRE(qkv && out, "swb200_window_attention: NULL pointer");
  return launch_window_attention(qkv, out, B, grid_h, grid_w, heads, shift_h, shift_w, qkv_fp16, out_fp16, impl,
                                 static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ forward-mode tangent

namespace {
struct JvpWs {
  size_t a2, xhl2, raw, qkvp, dqkvp, S, dS, attn2, branch2, h2, poszero, total;
};
JvpWs carve_jvp(const swb200_model* m) {
  const Geom g = geom(m);
  const size_t M = g.tokens;                                 // one sample at a time
  JvpWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  const size_t nraw = static_cast<size_t>(std::max(3 * m->dim, 2 * m->dff));
  const size_t items = static_cast<size_t>(g.gh / 16) * (g.gw / 16) * m->heads;
  w.a2 = take(2 * M * g.k_embed_total * 2);
  w.xhl2 = take(2 * M * m->dim * 2 * 2);
  w.raw = take(2 * M * nraw * 4);
  w.qkvp = take(static_cast<size_t>(3) * m->heads * M * kHeadDimPad * 2);
  w.dqkvp = take(static_cast<size_t>(3) * m->heads * M * kHeadDimPad * 2);
  w.S = take(items * 65536 * 4);
  w.dS = take(items * 65536 * 4);
  w.attn2 = take(2 * M * m->dim * 2);
  w.branch2 = take(2 * M * m->dim * 4);
  w.h2 = take(2 * M * static_cast<size_t>(m->dff) * 2);
  w.poszero = take(M * m->dim * 4);
  w.total = off;
  return w;
}
}  // namespace

extern "C" {

SWB200_API size_t swb200_jvp_workspace_bytes(const swb200_model* m) {
  if (validate(m) != SWB_OK) return 0;
  return carve_jvp(m).total;
}

SWB200_API size_t swb200_conditioning_jvp_scratch_bytes(const swb200_model* m, int B) {
  return 2 * swb200_conditioning_scratch_bytes(m, B);
}

SWB200_API int swb200_conditioning_jvp(const swb200_model* m, const float* t, const float* dt, const float* aux, int B, float* gain,
                            float* bias, float* dgain, float* dbias, void* scratch, size_t scratch_bytes, void* stream) {
  int rc = validate(m);
  if (rc) return rc;
  SWB_REQUIRE(B > 0 && t && dt && gain && bias && dgain && dbias && scratch, "conditioning_jvp: NULL argument or B=%d", B);
  SWB_REQUIRE(scratch_bytes >= swb200_conditioning_jvp_scratch_bytes(m, B), "conditioning_jvp: scratch too small");
  CondWeights w;
  w.aux_w = m->aux_w; w.aux_b = m->aux_b; w.aux_dim = m->aux_dim;
  w.l1_w = m->l1_w; w.l1_b = m->l1_b; w.l2_w = m->l2_w; w.l2_b = m->l2_b;
  w.mod_w = m->mod_w; w.mod_b = m->mod_b; w.ln_gamma = m->ln_gamma; w.ln_beta = m->ln_beta;
  const float* aux_eff = (m->aux_dim > 0 && m->aux_w != nullptr) ? aux : nullptr;
  return launch_conditioning_dual(w, t, dt, aux_eff, B, m->dim, 2 * m->depth, m->timestep_weight, static_cast<float*>(scratch),
                                  gain, bias, dgain, dbias, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_forward_jvp(const swb200_model* m, const float* x0, int c0, float scale0, const float* x1, int c1,
                       const float* dx0, int B, const float* gain, const float* bias, const float* dgain, const float* dbias,
                       float* y, float* dy, void* workspace, size_t workspace_bytes, void* stream_) {
  int rc = validate(m);
  if (rc) return rc;
  SWB_REQUIRE(B > 0 && x0 && dx0 && gain && bias && dgain && dbias && y && dy && workspace, "forward_jvp: NULL argument or B=%d", B);
  SWB_REQUIRE(c0 + c1 == m->in_channels && (c1 == 0 || x1 != nullptr), "forward_jvp: c0+c1=%d != in_channels=%d", c0 + c1,
              m->in_channels);
  SWB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "forward_jvp: workspace must be 1024-byte aligned");
  const JvpWs w = carve_jvp(m);
  SWB_REQUIRE(workspace_bytes >= w.total, "forward_jvp: workspace of %zu bytes < %zu needed", workspace_bytes, w.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, H = m->heads, Dff = m->dff, M = g.tokens;
  const int F16 = m->act_fp16 ? 1 : 0;
  const int tile = m->gemm_tile;
  const size_t img_in0 = static_cast<size_t>(c0) * m->img_h * m->img_w;
  const size_t img_in1 = static_cast<size_t>(c1) * m->img_h * m->img_w;
  const size_t img_out = static_cast<size_t>(m->out_channels) * m->img_h * m->img_w;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint16_t* a2 = reinterpret_cast<uint16_t*>(ws + w.a2);
  uint16_t* xhl2 = reinterpret_cast<uint16_t*>(ws + w.xhl2);
  float* raw = reinterpret_cast<float*>(ws + w.raw);
  void* qkvp = ws + w.qkvp;
  void* dqkvp = ws + w.dqkvp;
  float* S = reinterpret_cast<float*>(ws + w.S);
  float* dS = reinterpret_cast<float*>(ws + w.dS);
  uint16_t* attn2 = reinterpret_cast<uint16_t*>(ws + w.attn2);
  float* branch2 = reinterpret_cast<float*>(ws + w.branch2);
  uint16_t* h2 = reinterpret_cast<uint16_t*>(ws + w.h2);
  float* poszero = reinterpret_cast<float*>(ws + w.poszero);
  SWB_CHECK_CUDA(cudaMemsetAsync(poszero, 0, static_cast<size_t>(M) * D * 4, stream));
  const int KE = g.k_embed_total;

  for (int b = 0; b < B; ++b) {
    // 1. patch operands of the primal input cat([x0*scale0, x1]) and of its tangent cat([dx0*scale0, 0])
    rc = launch_patch_gather(x0 + b * img_in0, c0, scale0, x1 ? x1 + b * img_in1 : nullptr, c1, a2, KE, m->k_embed,
                             m->split_embed, F16, 1, m->img_h, m->img_w, m->patch_h, m->patch_w, stream);
    if (rc) return rc;
    rc = launch_patch_gather(dx0 + b * img_in0, c0, scale0, nullptr, 0, a2 + static_cast<size_t>(M) * KE, KE, m->k_embed,
                             m->split_embed, F16, 1, m->img_h, m->img_w, m->patch_h, m->patch_w, stream);
    if (rc) return rc;
    // 2. patch embed: the primal rows get bias + pos_embed, the tangent rows nothing
    for (int half = 0; half < 2; ++half) {
      GemmParams p = base_params(M, D, KE);
      p.out0 = xhl2 + static_cast<size_t>(half) * M * 2 * D;
      p.ldo = 2 * D;
      p.bias = half ? nullptr : m->b_embed;
      p.pos = half ? poszero : m->pos_embed;
      p.pos_rows = g.tokens;
      rc = launch_gemm(EPI_EMBED, tile, F16, a2 + static_cast<size_t>(half) * M * KE, KE, m->w_embed, KE, p, stream);
      if (rc) return rc;
    }
    // 3. transformer blocks on the row-stacked operand [x ; dx]
    for (int l = 0; l < m->depth; ++l) {
      const bool shifted = (m->shift_h || m->shift_w) && (l & 1);
      const size_t ca = (static_cast<size_t>(2 * l) * B + b) * D, cf = (static_cast<size_t>(2 * l + 1) * B + b) * D;
      {
        GemmParams p = base_params(2 * M, 3 * D, D);
        p.out0 = raw;
        p.ldo = 3 * D;
        const auto* wq = static_cast<const __nv_bfloat16*>(m->w_qkv) + static_cast<size_t>(l) * 3 * D * D;
        rc = launch_gemm(EPI_STORE_F32, tile, F16, xhl2, 2 * D, wq, D, p, stream);
        if (rc) return rc;
      }
      rc = launch_qkv_dual_pack(raw, m->qscale + static_cast<size_t>(l) * H, qkvp, dqkvp, M, D, H, kHeadDim, kHeadDimPad, F16, stream);
      if (rc) return rc;
      rc = launch_attention_dual(qkvp, dqkvp, S, dS, attn2, 1, g.gh, g.gw, H, kHeadDim, kHeadDimPad, shifted ? m->shift_h : 0,
                                 shifted ? m->shift_w : 0, F16, stream);
      if (rc) return rc;
      {
        GemmParams p = base_params(2 * M, D, D);
        p.out0 = branch2;
        p.ldo = D;
        const auto* wo = static_cast<const __nv_bfloat16*>(m->w_o) + static_cast<size_t>(l) * D * D;
        rc = launch_gemm(EPI_STORE_F32, tile, F16, attn2, D, wo, D, p, stream);
        if (rc) return rc;
      }
      rc = launch_ln_dual(branch2, xhl2, gain + ca, bias + ca, dgain + ca, dbias + ca, M, D, M, 1e-6f, F16, stream);
      if (rc) return rc;
      {
        GemmParams p = base_params(2 * M, 2 * Dff, D);
        p.out0 = raw;
        p.ldo = 2 * Dff;
        const auto* w1 = static_cast<const __nv_bfloat16*>(m->w_1) + static_cast<size_t>(l) * 2 * Dff * D;
        rc = launch_gemm(EPI_STORE_F32, tile, F16, xhl2, 2 * D, w1, D, p, stream);
        if (rc) return rc;
      }
      rc = launch_swiglu_dual(raw, h2, M, Dff, tile == 3 ? 2 * kUmmaN : kUmmaN, F16, stream);
      if (rc) return rc;
      {
        GemmParams p = base_params(2 * M, D, Dff);
        p.out0 = branch2;
        p.ldo = D;
        const auto* w2 = static_cast<const __nv_bfloat16*>(m->w_2) + static_cast<size_t>(l) * D * Dff;
        rc = launch_gemm(EPI_STORE_F32, tile, F16, h2, Dff, w2, Dff, p, stream);
        if (rc) return rc;
      }
      rc = launch_ln_dual(branch2, xhl2, gain + cf, bias + cf, dgain + cf, dbias + cf, M, D, M, 1e-6f, F16, stream);
      if (rc) return rc;
    }
    // 4. output head on the primal and on the tangent rows (plain F, no sampler update)
    swb200_update u = {};
    u.beta = 1.0f;
    rc = swb200_gemm_head(tile, m, xhl2, 2 * D, g.k_head_total, 1, &u, y + b * img_out, stream_);
    if (rc) return rc;
    rc = swb200_gemm_head(tile, m, xhl2 + static_cast<size_t>(M) * 2 * D, 2 * D, g.k_head_total, 1, &u, dy + b * img_out, stream_);
    if (rc) return rc;
  }
  return SWB_OK;
}

}  // extern "C"

extern "C" {

// ------------------------------------------------------------------------------------------------ tracing

SWB200_API int swb200_trace_enable(int on) {
  g_trace.on = on != 0;
  if (!on) return SWB_OK;
  for (int i = 0; i < T_NSLOTS; ++i) { g_trace.ms[i] = 0; g_trace.n[i] = 0; }
  return SWB_OK;
}

SWB200_API int swb200_trace_report(char* buf, size_t buf_bytes) {
  SWB_REQUIRE(buf && buf_bytes > 0, "swb200_trace_report: NULL buffer");
  SWB_CHECK_CUDA(cudaDeviceSynchronize());
  for (auto& e : g_trace.ev) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.second.first, e.second.second) == cudaSuccess) {
      g_trace.ms[e.first] += ms;
      g_trace.n[e.first] += 1;
    }
    cudaEventDestroy(e.second.first);
    cudaEventDestroy(e.second.second);
  }
  g_trace.ev.clear();
  std::string out = "{";
  char tmp[160];
  for (int i = 0; i < T_NSLOTS; ++i) {
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.4f, \"launches\": %ld}", i ? ", " : "", kTraceNames[i],
             g_trace.ms[i], g_trace.n[i]);
    out += tmp;
  }
  out += "}";
  SWB_REQUIRE(out.size() + 1 <= buf_bytes, "swb200_trace_report: buffer too small (%zu needed)", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return SWB_OK;
}

// ------------------------------------------------------------------------------------------------ rollout glue

SWB200_API int swb200_rollout_noise(float* latents, const uint64_t* seeds, const int32_t* step, int B, int64_t n_per_sample,
                         void* stream) {
  SWB_REQUIRE(latents && seeds && step && B > 0, "swb200_rollout_noise: NULL pointer or B=%d", B);
  return launch_rollout_noise(latents, reinterpret_cast<const unsigned long long*>(seeds), step, B, n_per_sample,
                              static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_rollout_forcings(float* cond, int total_channels, int state_channels, const float* table, int n_forcings,
                            int n_times, const int32_t* base, int stride, const int32_t* step, int B, int hw,
                            void* stream) {
  SWB_REQUIRE(cond && table && step && B > 0, "swb200_rollout_forcings: NULL pointer or B=%d", B);
  return launch_rollout_forcings(cond, total_channels, state_channels, table, n_forcings, n_times, base, stride, step, B, hw,
                                 static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_ensemble_stats(const float* phys, const float* truth, const float* w_lat, int n_ic, int members,
                                     int n_var, int H, int W, const int32_t* step, int n_steps, int out_stride, double* out,
                                     void* stream) {
  SWB_REQUIRE(phys && truth && w_lat && out, "swb200_ensemble_stats: NULL pointer");
  SWB_REQUIRE(step == nullptr || n_steps > 0, "swb200_ensemble_stats: n_steps %d with a device step counter", n_steps);
  SWB_REQUIRE(step == nullptr || out_stride >= n_ic * n_var * 4, "swb200_ensemble_stats: out_stride %d < n_ic*n_var*4", out_stride);
  return launch_ensemble_stats(phys, truth, w_lat, n_ic, members, n_var, H, W, step, n_steps, out_stride, out,
                               static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_scm_noised_inputs(const float* x, const float* z, const float* t, int B, int C, int H, int W, float* x_t,
                                        float* dxt, float* vx, float* vt, void* stream) {
  SWB_REQUIRE(x && z && t && x_t && dxt && vx && vt, "swb200_scm_noised_inputs: NULL pointer");
  return launch_scm_noised_inputs(x, z, t, B, C, H, W, x_t, dxt, vx, vt, static_cast<cudaStream_t>(stream));
}

SWB200_API size_t swb200_scm_target_scratch_bytes(int B) { return B > 0 ? scm_target_scratch_bytes(B) : 0; }

SWB200_API int swb200_scm_tangent_target(const float* F, const float* dF, const float* x_t, const float* dxt, const float* t,
                                         float r, float sigma_data, const float* w_var, const float* w_lat, int B, int C, int H,
                                         int W, float* g, float* cot, float* loss, void* scratch, size_t scratch_bytes,
                                         void* stream) {
  SWB_REQUIRE(F && dF && x_t && dxt && t && g && cot && loss && scratch, "swb200_scm_tangent_target: NULL pointer");
  return launch_scm_tangent_target(F, dF, x_t, dxt, t, r, sigma_data, w_var, w_lat, B, C, H, W, g, cot, loss, scratch,
                                   scratch_bytes, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_debug_saturation(uint64_t* counters) {
  g_sat_counters = reinterpret_cast<unsigned long long*>(counters);
  return SWB_OK;
}

SWB200_API int swb200_rollout_advance(int32_t* step, void* stream) {
  SWB_REQUIRE(step, "swb200_rollout_advance: NULL pointer");
  return launch_rollout_advance(step, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
