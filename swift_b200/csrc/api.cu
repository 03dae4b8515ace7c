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
static int gemm_head_impl(int tile, const swb200_model* m, const void* A, int lda, int K, int ldw, int B, const swb200_update* upd,
                          float* y, void* stream);


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
  // fp16 forecast path: the residual stream is ONE fp16 value per element (the hi half of xhl; section 2 of DESIGN.md)
  const int XS = (F16 && m->x_single) ? 1 : 0;
  const int FX = F16 | (XS << 1);             // format word of the kernels that touch the residual stream
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
      if (XS) return r;
      return r ? r : sat(SAT_X_LO, static_cast<const uint16_t*>(xhl) + D, M, D, 2 * D);
    };
    void* lnws = ws + w.lnws;
    int ln_gen = 0;                             // fused launches of this chunk so far (the first one clears the counters)

    // 1. concat + patchify + cast
    // single-value stream: the embed output is rounded to fp16 anyway, so the patch operand is its hi half alone (the lo half is
    // not written, K = k_embed on the first half of the packed [W | W]; row pitches unchanged): Swift-B one step 1.93e-3 ->
    // 2.05e-3.  The tangent / training paths keep the split operand.
    const int ES = (XS || !m->split_embed) ? 0 : 1;
    { TraceScope ts_(T_GATHER, stream);
    rc = launch_patch_gather(x0 + b0 * img_in0, c0, scale0, x1 ? x1 + b0 * img_in1 : nullptr, c1, a_emb,
                             g.k_embed_total, m->k_embed, ES, F16, bc, m->img_h, m->img_w, m->patch_h,
                             m->patch_w, stream); }
    if (rc) return rc;
    // 2. patch-embed GEMM (+bias +pos_embed)
    {
      GemmParams p = base_params(M, D, ES ? g.k_embed_total : m->k_embed);
      p.out0 = xhl;
      p.ldo = 2 * D;
      p.bias = m->b_embed;
      p.pos = m->pos_embed;
      p.pos_rows = g.tokens;
      { TraceScope ts_(T_EMBED, stream);
      rc = launch_gemm(EPI_EMBED, kDefaultCG, FX, a_emb, g.k_embed_total, m->w_embed, g.k_embed_total, p, stream); }
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
        rc = swb200_gemm_ln_residual(kDefaultCG, FX, attn, D, wo, D, xhl, gain_a, bias_a, M, D, g.tokens, lnws, ln_gen,
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
        rc = launch_ln_mod_residual(branch, BR16, xhl, gain_a, bias_a, M, D, g.tokens, 1e-6f, FX, stream); }
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
        rc = swb200_gemm_ln_residual(kDefaultCG, FX, hbuf, Dff, w2, Dff, xhl, gain_f, bias_f, M, D, g.tokens, lnws,
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
        rc = launch_ln_mod_residual(branch, BR16, xhl, gain_f, bias_f, M, D, g.tokens, 1e-6f, FX, stream); }
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
      rc = gemm_head_impl(kDefaultCG, m, xhl, 2 * D, XS ? D : g.k_head_total, g.k_head_total, bc, &u,
                          y ? y + b0 * img_out : nullptr, stream_); }
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
  // launch 0 clears every (group, row) slot (tag 0 = "nothing published"); launch g tags its partials 1 + g % 7: a slot is
  // rewritten by every launch, so the tag it holds when launch g starts is launch g-1's, never launch g's
  if (gen == 0) SWB_CHECK_CUDA(cudaMemsetAsync(ln_ws, 0, l.total, stream));
  const int tile_n = tile == 3 ? 2 * kUmmaN : kUmmaN;
  const int groups = (dim + tile_n - 1) / tile_n * (tile == 3 ? 2 : 1);
  SWB_REQUIRE(groups <= 12, "swb200_gemm_ln_residual: dim %d needs %d statistics groups (at most 12 are supported)", dim, groups);
  GemmParams p = base_params(M, dim, K);
  p.xhl = static_cast<uint16_t*>(xhl);
  p.gain = gain;
  p.lnbias = bias;
  p.tokens = tokens;
  p.ln_stats = reinterpret_cast<float2*>(static_cast<uint8_t*>(ln_ws) + l.counters_bytes);
  p.ln_stride = l.stride;
  p.ln_tag = 1u + static_cast<unsigned>(gen) % 7u;
  p.ln_eps = 1e-6f;
  static const int dbg = getenv("SWB_LN_DEBUG") ? atoi(getenv("SWB_LN_DEBUG")) : 0;   // profiling knob, see GemmParams
  p.ln_debug = dbg;
  return launch_gemm(EPI_LN_RES, tile, act_fp16, A, lda, W, K, p, stream);
}

SWB200_API int swb200_gemm_head(int tile, const swb200_model* m, const void* A, int lda, int K, int B,
                     const swb200_update* upd, float* y, void* stream) {
  return gemm_head_impl(tile, m, A, lda, K, K, B, upd, y, stream);
}

// K < ldw: only the first K columns of the packed head weight [W | W] are used (single-value residual stream)
static int gemm_head_impl(int tile, const swb200_model* m, const void* A, int lda, int K, int ldw, int B, const swb200_update* upd,
                   float* y, void* stream) {
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
  return launch_gemm(EPI_HEAD, tile, m->act_fp16 ? 1 : 0, A, lda, m->w_head, ldw, p, static_cast<cudaStream_t>(stream));
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
                            int shift_w, int qkv_fp16, int out_fp16, int impl, float* lse, void* stream) {
  SWB_REQUIRE(qkv && out, "swb200_window_attention: NULL pointer");
  return launch_window_attention(qkv, out, B, grid_h, grid_w, heads, shift_h, shift_w, qkv_fp16, out_fp16, impl,
                                 static_cast<cudaStream_t>(stream), lse);
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

SWB200_API int swb200_scm_tangent_target_logvar(const float* F, const float* dF, const float* x_t, const float* dxt,
                                                const float* t, float r, float sigma_data, const float* w_var,
                                                const float* w_lat, int B, int C, int H, int W, const float* logvar, float* g,
                                                float* cot, float* loss, float* dlogvar, void* scratch, size_t scratch_bytes,
                                                void* stream) {
  SWB_REQUIRE(F && dF && x_t && dxt && t && g && cot && loss && scratch && logvar && dlogvar,
              "swb200_scm_tangent_target_logvar: NULL pointer");
  return launch_scm_tangent_target(F, dF, x_t, dxt, t, r, sigma_data, w_var, w_lat, B, C, H, W, g, cot, loss, scratch,
                                   scratch_bytes, static_cast<cudaStream_t>(stream), logvar, dlogvar);
}

SWB200_API int swb200_scm_distill_direction(const float* F_teacher, const float* t, float sigma_data, int B, int C, int H, int W,
                                            float* dxt, float* vx, void* stream) {
  SWB_REQUIRE(F_teacher && t && dxt && vx, "swb200_scm_distill_direction: NULL pointer");
  return launch_scm_distill_direction(F_teacher, t, sigma_data, B, C, H, W, dxt, vx, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_logvar_head(const swb200_model* m, const void* fwd_scratch, const float* lv_w, const float* lv_b, int B,
                                  float* logvar, void* stream) {
  SWB_REQUIRE(m && fwd_scratch && lv_w && lv_b && logvar && B > 0, "swb200_logvar_head: NULL pointer");
  return launch_logvar_head(static_cast<const float*>(fwd_scratch), lv_w, lv_b, B, m->dim, logvar, static_cast<cudaStream_t>(stream));
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

// ------------------------------------------------------------------------------------------------ reverse mode
// The grad-enabled forward and the backward of the sCM training step (training/loss.py:226-260, trainer.py:199-219):
// see include/swift_b200.h for the contract, train.cu / attention_bwd.cu for the kernels between the GEMMs.

namespace {

struct Tape {
  size_t a_emb, x, x_stride, layers, layer_stride;
  size_t qkv, invn, lse, attn, b1, gu, h, b2;   // offsets inside one layer block
  size_t total;
};
Tape carve_tape(const swb200_model* m, int B) {
  const Geom g = geom(m);
  const size_t M = static_cast<size_t>(B) * g.tokens, D = m->dim, Dff = m->dff;
  Tape t;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  t.a_emb = take(M * g.k_embed_total * 2);
  t.x_stride = align_up(M * D * 2 * 2, 1024);
  t.x = take(t.x_stride * (2 * static_cast<size_t>(m->depth) + 1));
  t.layers = off;
  size_t loff = 0;
  auto ltake = [&](size_t bytes) {
    size_t o = loff;
    loff = align_up(loff + bytes, 1024);
    return o;
  };
  t.qkv = ltake(static_cast<size_t>(3) * m->heads * M * kHeadDimPad * 2);
  t.invn = ltake(static_cast<size_t>(2) * m->heads * M * 4);
  t.lse = ltake(static_cast<size_t>(m->heads) * M * 4);
  t.attn = ltake(M * D * 2);
  t.b1 = ltake(M * D * 4);
  t.gu = ltake(M * 2 * Dff * 2);
  t.h = ltake(M * Dff * 2);
  t.b2 = ltake(M * D * 4);
  t.layer_stride = loff;
  t.total = t.layers + t.layer_stride * m->depth;
  return t;
}

// split-K factor of a weight-gradient GEMM [n_out, k_in] contracting over `tokens` rows: enough tiles for ~2 waves of the
// 74 CTA pairs, at least 512 tokens per split
int choose_splits(int n_out, int k_in, int tokens) {
  const int tiles = ((n_out + 255) / 256) * ((k_in + 351) / 352);
  int s = 1;
  while (s < 16 && tiles * s < 128 && tokens % (2 * s * kBlockK) == 0 && tokens / (2 * s) >= 512) s *= 2;
  return s;
}

struct TrainWs {
  size_t dx, tmp, dy16, da16, dyT, xT, part, lnpart, Lbuf, Dbuf, dspart, total;
};
TrainWs carve_train_ws(const swb200_train_model* tm, int B) {
  const swb200_model* m = &tm->base;
  const Geom g = geom(m);
  const size_t M = static_cast<size_t>(B) * g.tokens, D = m->dim, Dff = m->dff, Kp = tm->kp_head;
  const size_t Ke = m->k_embed;
  TrainWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  w.dx = take(M * D * 4);
  w.tmp = take(M * std::max(Dff, 3 * D) * 4);
  const size_t wide = std::max(std::max(3 * D, 2 * Dff), Kp);
  w.dy16 = take(M * wide * 2);
  w.da16 = take(M * D * 2);
  w.dyT = take(wide * M * 2);
  w.xT = take(std::max(std::max(Dff, D), Ke) * M * 2);
  const int Mi = static_cast<int>(M);
  size_t pf = 0;
  auto need = [&](size_t n_out, size_t k_in) {
    pf = std::max(pf, static_cast<size_t>(choose_splits(static_cast<int>(n_out), static_cast<int>(k_in), Mi)) * n_out * k_in);
  };
  need(3 * D, D);
  need(D, D);
  need(2 * Dff, D);
  need(D, Dff);
  need(Kp, D);
  need(Ke, D);
  w.part = take(pf * 4);
  w.lnpart = take((ln_bwd_partial_floats(Mi, m->dim) + 2 * static_cast<size_t>(B) * D + colsum_partial_floats(Mi, m->dim)) * 4);
  w.Lbuf = take(static_cast<size_t>(m->heads) * M * 4);
  w.Dbuf = take(static_cast<size_t>(m->heads) * M * 4);
  const size_t items = static_cast<size_t>(B) * (g.gh / 16) * (g.gw / 16) * m->heads;
  w.dspart = take((items * 8 + static_cast<size_t>(m->heads) * 8) * 4);
  w.total = off;
  return w;
}

// the tcgen05 attention kernels (forward with log-sum-exp output, backward) handle shifts that are multiples of 8
bool attn_tc_path(const swb200_model* m, bool shifted) {
  if (m->attn_impl == 1) return false;
  return !shifted || (m->shift_h % 8 == 0 && m->shift_w % 8 == 0);
}

int validate_train(const swb200_train_model* tm) {
  SWB_REQUIRE(tm != nullptr, "train model is NULL");
  int rc = validate(&tm->base);
  if (rc) return rc;
  SWB_REQUIRE(tm->base.act_fp16 == 0, "the training path computes with bf16 operands: base.act_fp16 must be 0");
  SWB_REQUIRE(tm->w_qkv && tm->w_o && tm->w_1 && tm->w_2 && tm->wt_qkv && tm->wt_o && tm->wt_1 && tm->wt_2 && tm->wt_head,
              "train model: NULL weight pointer");
  const Geom g = geom(&tm->base);
  SWB_REQUIRE(tm->kp_head % 8 == 0 && tm->kp_head >= tm->base.out_channels * g.pp, "kp_head=%d invalid", tm->kp_head);
  return SWB_OK;
}

// dW [n_out, k_in] (+)= dyT [n_out, tokens] * xT [k_in, tokens]^T, split-K over the tokens; `out_rows` <= n_out rows are kept
int wgrad(const swb200_train_model* tm, const void* dyT, int n_out, const void* xT, int k_in, int tokens, float* part,
          float* out, int out_rows, int accumulate, cudaStream_t stream) {
  const int S = choose_splits(n_out, k_in, tokens);
  GemmParams p = base_params(n_out, k_in, tokens / S);
  p.out0 = part;
  p.ldo = k_in;
  p.splits = S;
  int rc = launch_gemm(EPI_STORE_F32, tm->base.gemm_tile, 0, dyT, tokens, xT, tokens, p, stream);
  if (rc) return rc;
  return launch_splitk_reduce(part, S, static_cast<long long>(n_out) * k_in, out, static_cast<long long>(out_rows) * k_in,
                              accumulate, stream);
}

// out [tokens, n_in] = dy16 [tokens, k] * wt [n_in, k]^T
int dgrad(const swb200_train_model* tm, const void* dy16, int k, const void* wt, int n_in, int tokens, void* out, int epi,
          cudaStream_t stream) {
  GemmParams p = base_params(tokens, n_in, k);
  p.out0 = out;
  p.ldo = n_in;
  return launch_gemm(epi, tm->base.gemm_tile, 0, dy16, k, wt, k, p, stream);
}

}  // namespace

extern "C" {

SWB200_API size_t swb200_train_tape_bytes(const swb200_train_model* tm, int B) {
  if (validate_train(tm) != SWB_OK || B <= 0) return 0;
  return carve_tape(&tm->base, B).total;
}

SWB200_API size_t swb200_train_workspace_bytes(const swb200_train_model* tm, int B) {
  if (validate_train(tm) != SWB_OK || B <= 0) return 0;
  return carve_train_ws(tm, B).total;
}

SWB200_API int swb200_train_forward(const swb200_train_model* tm, const float* x0, int c0, float scale0, const float* x1, int c1,
                                    int B, const float* gain, const float* bias, float* y, void* tape_, size_t tape_bytes,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  int rc = validate_train(tm);
  if (rc) return rc;
  const swb200_model* m = &tm->base;
  SWB_REQUIRE(B > 0 && x0 && gain && bias && y && tape_ && workspace, "train_forward: NULL argument or B=%d", B);
  SWB_REQUIRE(c0 + c1 == m->in_channels && (c1 == 0 || x1 != nullptr), "train_forward: c0+c1=%d != in_channels=%d", c0 + c1,
              m->in_channels);
  const Tape t = carve_tape(m, B);
  const TrainWs w = carve_train_ws(tm, B);
  SWB_REQUIRE(tape_bytes >= t.total && workspace_bytes >= w.total, "train_forward: tape (%zu < %zu) or workspace (%zu < %zu) too small",
              tape_bytes, t.total, workspace_bytes, w.total);
  SWB_REQUIRE(((reinterpret_cast<uintptr_t>(tape_) | reinterpret_cast<uintptr_t>(workspace)) & 1023) == 0,
              "train_forward: tape and workspace must be 1024-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, H = m->heads, Dff = m->dff, tile = m->gemm_tile;
  const int M = B * g.tokens;
  uint8_t* tp = static_cast<uint8_t*>(tape_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* raw = reinterpret_cast<float*>(ws + w.tmp);
  auto xbuf = [&](int i) { return tp + t.x + t.x_stride * static_cast<size_t>(i); };
  void* a_emb = tp + t.a_emb;
  rc = launch_patch_gather(x0, c0, scale0, x1, c1, a_emb, g.k_embed_total, m->k_embed, m->split_embed, 0, B, m->img_h, m->img_w,
                           m->patch_h, m->patch_w, stream);
  if (rc) return rc;
  {
    GemmParams p = base_params(M, D, g.k_embed_total);
    p.out0 = xbuf(0);
    p.ldo = 2 * D;
    p.bias = m->b_embed;
    p.pos = m->pos_embed;
    p.pos_rows = g.tokens;
    rc = launch_gemm(EPI_EMBED, tile, 0, a_emb, g.k_embed_total, m->w_embed, g.k_embed_total, p, stream);
    if (rc) return rc;
  }
  const size_t xbytes = static_cast<size_t>(M) * D * 2 * 2;
  for (int l = 0; l < m->depth; ++l) {
    uint8_t* lb = tp + t.layers + t.layer_stride * static_cast<size_t>(l);
    const bool shifted = (m->shift_h || m->shift_w) && (l & 1);
    void* x0p = xbuf(2 * l);
    void* x1p = xbuf(2 * l + 1);
    void* x2p = xbuf(2 * l + 2);
    {   // to_qkv (reference row order) -> raw fp32 -> normalise / scale / pack
      GemmParams p = base_params(M, 3 * D, D);
      p.out0 = raw;
      p.ldo = 3 * D;
      const auto* wq = static_cast<const __nv_bfloat16*>(tm->w_qkv) + static_cast<size_t>(l) * 3 * D * D;
      rc = launch_gemm(EPI_STORE_F32, tile, 0, x0p, 2 * D, wq, D, p, stream);
      if (rc) return rc;
      rc = launch_qkv_pack_train(raw, m->qscale + static_cast<size_t>(l) * H, lb + t.qkv, reinterpret_cast<float*>(lb + t.invn), M,
                                 H, kHeadDim, kHeadDimPad, stream);
      if (rc) return rc;
    }
    rc = launch_window_attention(lb + t.qkv, lb + t.attn, B, g.gh, g.gw, H, shifted ? m->shift_h : 0, shifted ? m->shift_w : 0, 0, 0,
                                 m->attn_impl, stream, attn_tc_path(m, shifted) ? reinterpret_cast<float*>(lb + t.lse) : nullptr);
    if (rc) return rc;
    {
      GemmParams p = base_params(M, D, D);
      p.out0 = lb + t.b1;
      p.ldo = D;
      const auto* wo = static_cast<const __nv_bfloat16*>(tm->w_o) + static_cast<size_t>(l) * D * D;
      rc = launch_gemm(EPI_STORE_F32, tile, 0, lb + t.attn, D, wo, D, p, stream);
      if (rc) return rc;
    }
    SWB_CHECK_CUDA(cudaMemcpyAsync(x1p, x0p, xbytes, cudaMemcpyDeviceToDevice, stream));
    rc = launch_ln_mod_residual(lb + t.b1, 0, x1p, gain + static_cast<size_t>(2 * l) * B * D, bias + static_cast<size_t>(2 * l) * B * D,
                                M, D, g.tokens, 1e-6f, 0, stream);
    if (rc) return rc;
    {
      GemmParams p = base_params(M, 2 * Dff, D);
      p.out0 = lb + t.gu;
      p.ldo = 2 * Dff;
      const auto* w1 = static_cast<const __nv_bfloat16*>(tm->w_1) + static_cast<size_t>(l) * 2 * Dff * D;
      rc = launch_gemm(EPI_STORE_ACT, tile, 0, x1p, 2 * D, w1, D, p, stream);
      if (rc) return rc;
      rc = launch_swiglu_fwd_train(lb + t.gu, lb + t.h, M, Dff, stream);
      if (rc) return rc;
    }
    {
      GemmParams p = base_params(M, D, Dff);
      p.out0 = lb + t.b2;
      p.ldo = D;
      const auto* w2 = static_cast<const __nv_bfloat16*>(tm->w_2) + static_cast<size_t>(l) * D * Dff;
      rc = launch_gemm(EPI_STORE_F32, tile, 0, lb + t.h, Dff, w2, Dff, p, stream);
      if (rc) return rc;
    }
    SWB_CHECK_CUDA(cudaMemcpyAsync(x2p, x1p, xbytes, cudaMemcpyDeviceToDevice, stream));
    rc = launch_ln_mod_residual(lb + t.b2, 0, x2p, gain + static_cast<size_t>(2 * l + 1) * B * D,
                                bias + static_cast<size_t>(2 * l + 1) * B * D, M, D, g.tokens, 1e-6f, 0, stream);
    if (rc) return rc;
  }
  swb200_update u = {};
  u.beta = 1.0f;
  u.zero_channel = -1;
  return swb200_gemm_head(tile, m, xbuf(2 * m->depth), 2 * D, g.k_head_total, B, &u, y, stream_);
}

SWB200_API int swb200_train_backward_head(const swb200_train_model* tm, int B, const float* cot, const void* tape_,
                                          void* workspace, size_t workspace_bytes, const swb200_train_grads* gr, void* stream_) {
  int rc = validate_train(tm);
  if (rc) return rc;
  const swb200_model* m = &tm->base;
  SWB_REQUIRE(B > 0 && cot && tape_ && workspace && gr && gr->w_head, "train_backward_head: NULL argument");
  const Tape t = carve_tape(m, B);
  const TrainWs w = carve_train_ws(tm, B);
  SWB_REQUIRE(workspace_bytes >= w.total, "train_backward_head: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, M = B * g.tokens, Kp = tm->kp_head, Nh = m->out_channels * g.pp;
  const uint8_t* tp = static_cast<const uint8_t*>(tape_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* dF16 = ws + w.dy16;
  rc = launch_cot_patchify(cot, dF16, B, m->out_channels, m->img_h, m->img_w, m->patch_h, m->patch_w, Kp, stream);
  if (rc) return rc;
  rc = dgrad(tm, dF16, Kp, tm->wt_head, D, M, ws + w.dx, EPI_STORE_F32, stream);
  if (rc) return rc;
  rc = launch_transpose16(dF16, M, Kp, Kp, ws + w.dyT, M, stream);
  if (rc) return rc;
  const void* xL = tp + t.x + t.x_stride * static_cast<size_t>(2 * m->depth);
  rc = launch_transpose16(xL, M, D, 2 * D, ws + w.xT, M, stream);
  if (rc) return rc;
  return wgrad(tm, ws + w.dyT, Kp, ws + w.xT, D, M, reinterpret_cast<float*>(ws + w.part), gr->w_head, Nh, gr->accumulate, stream);
}

SWB200_API int swb200_train_backward_layer(const swb200_train_model* tm, int l, int B, const float* gain, const void* tape_,
                                           void* workspace, size_t workspace_bytes, const swb200_train_grads* gr, void* stream_) {
  int rc = validate_train(tm);
  if (rc) return rc;
  const swb200_model* m = &tm->base;
  SWB_REQUIRE(B > 0 && gain && tape_ && workspace && gr && l >= 0 && l < m->depth, "train_backward_layer: bad argument (layer %d)", l);
  SWB_REQUIRE(gr->w_qkv && gr->w_o && gr->w_1 && gr->w_2 && gr->dscale && gr->dgain && gr->dbias, "train_backward_layer: NULL gradient buffer");
  const Tape t = carve_tape(m, B);
  const TrainWs w = carve_train_ws(tm, B);
  SWB_REQUIRE(workspace_bytes >= w.total, "train_backward_layer: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, H = m->heads, Dff = m->dff, M = B * g.tokens, acc = gr->accumulate;
  const uint8_t* tp = static_cast<const uint8_t*>(tape_);
  const uint8_t* lb = tp + t.layers + t.layer_stride * static_cast<size_t>(l);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* dx = reinterpret_cast<float*>(ws + w.dx);
  float* tmp = reinterpret_cast<float*>(ws + w.tmp);
  void* dy16 = ws + w.dy16;
  void* dyT = ws + w.dyT;
  void* xT = ws + w.xT;
  float* part = reinterpret_cast<float*>(ws + w.part);
  float* lnpart = reinterpret_cast<float*>(ws + w.lnpart);
  const void* x0p = tp + t.x + t.x_stride * static_cast<size_t>(2 * l);
  const void* x1p = tp + t.x + t.x_stride * static_cast<size_t>(2 * l + 1);
  const bool shifted = (m->shift_h || m->shift_w) && (l & 1);
  const size_t vec = static_cast<size_t>(B) * D;
  using bf = __nv_bfloat16;

  // ---- feed-forward branch: x2 = x1 + LNmod(w2 (silu(g) u)),  [g | u] = w1 x1
  rc = launch_ln_bwd(dx, nullptr, reinterpret_cast<const float*>(lb + t.b2), gain + (2 * l + 1) * vec, dy16, lnpart,
                     gr->dgain + (2 * l + 1) * vec, gr->dbias + (2 * l + 1) * vec, M, D, g.tokens, 1e-6f, acc, stream);
  if (rc) return rc;
  if ((rc = launch_transpose16(dy16, M, D, D, dyT, M, stream))) return rc;
  if ((rc = launch_transpose16(lb + t.h, M, Dff, Dff, xT, M, stream))) return rc;
  if ((rc = wgrad(tm, dyT, D, xT, Dff, M, part, gr->w_2 + static_cast<size_t>(l) * D * Dff, D, acc, stream))) return rc;
  if ((rc = dgrad(tm, dy16, D, static_cast<const bf*>(tm->wt_2) + static_cast<size_t>(l) * Dff * D, Dff, M, tmp, EPI_STORE_F32, stream)))
    return rc;
  if ((rc = launch_swiglu_bwd(tmp, lb + t.gu, dy16, M, Dff, stream))) return rc;
  if ((rc = launch_transpose16(dy16, M, 2 * Dff, 2 * Dff, dyT, M, stream))) return rc;
  if ((rc = launch_transpose16(x1p, M, D, 2 * D, xT, M, stream))) return rc;
  if ((rc = wgrad(tm, dyT, 2 * Dff, xT, D, M, part, gr->w_1 + static_cast<size_t>(l) * 2 * Dff * D, 2 * Dff, acc, stream))) return rc;
  if ((rc = dgrad(tm, dy16, 2 * Dff, static_cast<const bf*>(tm->wt_1) + static_cast<size_t>(l) * D * 2 * Dff, D, M, tmp, EPI_STORE_F32,
                  stream)))
    return rc;
  // ---- attention branch: x1 = x0 + LNmod(wo attention(qkv x0));  dx <- dx + (w1 dgrad) on the way
  rc = launch_ln_bwd(dx, tmp, reinterpret_cast<const float*>(lb + t.b1), gain + (2 * l) * vec, dy16, lnpart, gr->dgain + (2 * l) * vec,
                     gr->dbias + (2 * l) * vec, M, D, g.tokens, 1e-6f, acc, stream);
  if (rc) return rc;
  if ((rc = launch_transpose16(dy16, M, D, D, dyT, M, stream))) return rc;
  if ((rc = launch_transpose16(lb + t.attn, M, D, D, xT, M, stream))) return rc;
  if ((rc = wgrad(tm, dyT, D, xT, D, M, part, gr->w_o + static_cast<size_t>(l) * D * D, D, acc, stream))) return rc;
  if ((rc = dgrad(tm, dy16, D, static_cast<const bf*>(tm->wt_o) + static_cast<size_t>(l) * D * D, D, M, ws + w.da16, EPI_STORE_ACT, stream)))
    return rc;
  float* dspart = reinterpret_cast<float*>(ws + w.dspart);
  const bool tc = attn_tc_path(m, shifted);
  const int nper = tc ? 8 : 4;                  // partial sums of the logit-scale gradient per (sample, window, head)
  if (tc)
    rc = launch_attention_bwd_tc(lb + t.qkv, lb + t.attn, ws + w.da16, reinterpret_cast<const float*>(lb + t.lse),
                                 reinterpret_cast<const float*>(lb + t.invn), m->qscale + static_cast<size_t>(l) * H, dy16,
                                 reinterpret_cast<float*>(ws + w.Dbuf), dspart, B, g.gh, g.gw, H, shifted ? m->shift_h : 0,
                                 shifted ? m->shift_w : 0, stream);
  else
    rc = launch_attention_bwd(lb + t.qkv, lb + t.attn, ws + w.da16, reinterpret_cast<const float*>(lb + t.invn),
                              m->qscale + static_cast<size_t>(l) * H, dy16, reinterpret_cast<float*>(ws + w.Lbuf),
                              reinterpret_cast<float*>(ws + w.Dbuf), dspart, B, g.gh, g.gw, H, kHeadDim, kHeadDimPad,
                              shifted ? m->shift_h : 0, shifted ? m->shift_w : 0, stream);
  if (rc) return rc;
  {   // ds[head] = sum over (sample, window) and the kernel's partials per item: [bw][head][nper] -> [head][nper] -> [head]
    const int per = B * (g.gh / 16) * (g.gw / 16);
    float* small = dspart + static_cast<size_t>(per) * H * nper;
    if ((rc = launch_reduce_partials(dspart, per, H * nper, small, 1, 0, stream))) return rc;
    if ((rc = launch_reduce_partials(small, nper, 1, gr->dscale + static_cast<size_t>(l) * H, H, acc, stream))) return rc;
  }
  if ((rc = launch_transpose16(dy16, M, 3 * D, 3 * D, dyT, M, stream))) return rc;
  if ((rc = launch_transpose16(x0p, M, D, 2 * D, xT, M, stream))) return rc;
  if ((rc = wgrad(tm, dyT, 3 * D, xT, D, M, part, gr->w_qkv + static_cast<size_t>(l) * 3 * D * D, 3 * D, acc, stream))) return rc;
  if ((rc = dgrad(tm, dy16, 3 * D, static_cast<const bf*>(tm->wt_qkv) + static_cast<size_t>(l) * D * 3 * D, D, M, tmp, EPI_STORE_F32,
                  stream)))
    return rc;
  return launch_add_f32(dx, tmp, dx, nullptr, static_cast<long long>(M) * D, stream);
}

SWB200_API int swb200_train_backward_embed(const swb200_train_model* tm, int B, const void* tape_, void* workspace,
                                           size_t workspace_bytes, const swb200_train_grads* gr, void* stream_) {
  int rc = validate_train(tm);
  if (rc) return rc;
  const swb200_model* m = &tm->base;
  SWB_REQUIRE(B > 0 && tape_ && workspace && gr && gr->w_embed_t && gr->b_embed && gr->pos_embed, "train_backward_embed: NULL argument");
  const Tape t = carve_tape(m, B);
  const TrainWs w = carve_train_ws(tm, B);
  SWB_REQUIRE(workspace_bytes >= w.total, "train_backward_embed: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, M = B * g.tokens, Ke = m->k_embed, acc = gr->accumulate;
  const uint8_t* tp = static_cast<const uint8_t*>(tape_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* dx = reinterpret_cast<float*>(ws + w.dx);
  float* lnpart = reinterpret_cast<float*>(ws + w.lnpart);
  if ((rc = launch_add_f32(dx, nullptr, nullptr, ws + w.dy16, static_cast<long long>(M) * D, stream))) return rc;
  if ((rc = launch_sum_over_samples(dx, B, static_cast<long long>(g.tokens) * D, gr->pos_embed, acc, stream))) return rc;
  if ((rc = launch_colsum(dx, M, D, lnpart, gr->b_embed, acc, stream))) return rc;
  if ((rc = launch_transpose16(ws + w.dy16, M, D, D, ws + w.dyT, M, stream))) return rc;
  if ((rc = launch_transpose16(tp + t.a_emb, M, Ke, g.k_embed_total, ws + w.xT, M, stream))) return rc;
  return wgrad(tm, ws + w.xT, Ke, ws + w.dyT, D, M, reinterpret_cast<float*>(ws + w.part), gr->w_embed_t, Ke, acc, stream);
}

SWB200_API size_t swb200_conditioning_backward_scratch_bytes(const swb200_model* m, int B) {
  if (m == nullptr || B <= 0) return 0;
  return conditioning_bwd_scratch_floats(B, m->dim, 2 * m->depth) * sizeof(float);
}

SWB200_API int swb200_conditioning_backward(const swb200_model* m, const float* aux, int B, const void* fwd_scratch,
                                            const float* dgain, const float* dbias, const swb200_cond_grads* gr, int accumulate,
                                            void* scratch, size_t scratch_bytes, void* stream) {
  return swb200_conditioning_backward_logvar(m, aux, B, fwd_scratch, dgain, dbias, gr, accumulate, scratch, scratch_bytes,
                                             nullptr, nullptr, nullptr, nullptr, stream);
}

SWB200_API int swb200_conditioning_backward_logvar(const swb200_model* m, const float* aux, int B, const void* fwd_scratch,
                                                   const float* dgain, const float* dbias, const swb200_cond_grads* gr,
                                                   int accumulate, void* scratch, size_t scratch_bytes, const float* lv_w,
                                                   const float* dlogvar, float* g_lv_w, float* g_lv_b, void* stream) {
  int rc = validate(m);
  SWB_REQUIRE((lv_w == nullptr) == (dlogvar == nullptr), "conditioning_backward: lv_w and dlogvar come together");
  if (rc) return rc;
  SWB_REQUIRE(B > 0 && fwd_scratch && dgain && dbias && gr && scratch, "conditioning_backward: NULL argument");
  SWB_REQUIRE(gr->l1_w && gr->l1_b && gr->l2_w && gr->l2_b && gr->mod_w && gr->mod_b && gr->ln_gamma && gr->ln_beta,
              "conditioning_backward: NULL gradient buffer");
  SWB_REQUIRE(scratch_bytes >= swb200_conditioning_backward_scratch_bytes(m, B), "conditioning_backward: scratch too small");
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
  CondGrads g;
  g.aux_w = gr->aux_w;
  g.aux_b = gr->aux_b;
  g.l1_w = gr->l1_w;
  g.l1_b = gr->l1_b;
  g.l2_w = gr->l2_w;
  g.l2_b = gr->l2_b;
  g.mod_w = gr->mod_w;
  g.mod_b = gr->mod_b;
  g.ln_gamma = gr->ln_gamma;
  g.ln_beta = gr->ln_beta;
  const float* aux_eff = (m->aux_dim > 0 && m->aux_w != nullptr) ? aux : nullptr;
  return launch_conditioning_bwd(w, g, aux_eff, static_cast<const float*>(fwd_scratch), dgain, dbias, B, m->dim, 2 * m->depth,
                                 static_cast<float*>(scratch), accumulate, static_cast<cudaStream_t>(stream), lv_w, dlogvar,
                                 g_lv_w, g_lv_b);
}

// ---- unit-test entry points of the reverse-mode kernels
SWB200_API int swb200_gemm_splitk(int tile, const void* A, int lda, const void* W, int ldw, float* partials, int ldo, int M, int N,
                                  int K_per_split, int splits, void* stream) {
  SWB_REQUIRE(A && W && partials, "swb200_gemm_splitk: NULL pointer");
  GemmParams p = base_params(M, N, K_per_split);
  p.out0 = partials;
  p.ldo = ldo;
  p.splits = splits;
  return launch_gemm(EPI_STORE_F32, tile, 0, A, lda, W, ldw, p, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_transpose16(const void* in, int rows, int cols, int64_t ldi, void* out, int64_t ldo, void* stream) {
  SWB_REQUIRE(in && out, "swb200_transpose16: NULL pointer");
  return launch_transpose16(in, rows, cols, ldi, out, ldo, static_cast<cudaStream_t>(stream));
}

SWB200_API size_t swb200_ln_backward_scratch_bytes(int M, int dim, int tokens) {
  if (M <= 0 || dim <= 0 || tokens <= 0) return 0;
  return (ln_bwd_partial_floats(M, dim) + 2 * static_cast<size_t>(M / tokens) * dim) * sizeof(float);
}

SWB200_API int swb200_ln_backward(float* dx, const float* add, const float* branch, const float* gain, void* db16, float* dgain,
                                  float* dbias, int M, int dim, int tokens, int accumulate, void* scratch, size_t scratch_bytes,
                                  void* stream) {
  SWB_REQUIRE(dx && branch && gain && db16 && dgain && dbias && scratch, "swb200_ln_backward: NULL pointer");
  SWB_REQUIRE(scratch_bytes >= swb200_ln_backward_scratch_bytes(M, dim, tokens), "swb200_ln_backward: scratch too small");
  return launch_ln_bwd(dx, add, branch, gain, db16, static_cast<float*>(scratch), dgain, dbias, M, dim, tokens, 1e-6f, accumulate,
                       static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_swiglu_backward(const float* dh, const void* gu, void* dgu, int M, int dff, void* stream) {
  SWB_REQUIRE(dh && gu && dgu, "swb200_swiglu_backward: NULL pointer");
  return launch_swiglu_bwd(dh, gu, dgu, M, dff, static_cast<cudaStream_t>(stream));
}

SWB200_API size_t swb200_attention_backward_scratch_bytes(int B, int grid_h, int grid_w, int heads) {
  if (B <= 0 || grid_h <= 0 || grid_w <= 0 || heads <= 0) return 0;
  const size_t M = static_cast<size_t>(B) * grid_h * grid_w;
  const size_t items = static_cast<size_t>(B) * (grid_h / 16) * (grid_w / 16) * heads;
  return (2 * heads * M + items * 8 + static_cast<size_t>(heads) * 8) * sizeof(float);
}

SWB200_API int swb200_attention_backward(const void* qkv, const void* O, const void* dO, const float* invn, const float* qscale,
                                         const float* lse, int impl, void* dqkv, float* dscale, int B, int grid_h, int grid_w,
                                         int heads, int shift_h, int shift_w, int accumulate, void* scratch, size_t scratch_bytes,
                                         void* stream_) {
  SWB_REQUIRE(qkv && O && dO && invn && qscale && dqkv && dscale && scratch, "swb200_attention_backward: NULL pointer");
  SWB_REQUIRE(impl == 1 || impl == 2, "swb200_attention_backward: impl must be 1 (mma.sync) or 2 (tcgen05)");
  SWB_REQUIRE(impl == 1 || lse != nullptr, "swb200_attention_backward: the tcgen05 kernel needs the forward's log-sum-exp");
  SWB_REQUIRE(scratch_bytes >= swb200_attention_backward_scratch_bytes(B, grid_h, grid_w, heads),
              "swb200_attention_backward: scratch too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t M = static_cast<size_t>(B) * grid_h * grid_w;
  float* Lbuf = static_cast<float*>(scratch);
  float* Dbuf = Lbuf + heads * M;
  float* dspart = Dbuf + heads * M;
  const int nper = impl == 2 ? 8 : 4;
  int rc = impl == 2 ? launch_attention_bwd_tc(qkv, O, dO, lse, invn, qscale, dqkv, Dbuf, dspart, B, grid_h, grid_w, heads, shift_h,
                                               shift_w, stream)
                     : launch_attention_bwd(qkv, O, dO, invn, qscale, dqkv, Lbuf, Dbuf, dspart, B, grid_h, grid_w, heads, kHeadDim,
                                            kHeadDimPad, shift_h, shift_w, stream);
  if (rc) return rc;
  const int per = B * (grid_h / 16) * (grid_w / 16);
  float* small = dspart + static_cast<size_t>(per) * heads * nper;
  if ((rc = launch_reduce_partials(dspart, per, heads * nper, small, 1, 0, stream))) return rc;
  return launch_reduce_partials(small, nper, 1, dscale, heads, accumulate, stream);
}

SWB200_API int swb200_qkv_pack_train(const float* raw, const float* qscale, void* packed, float* invn, int M, int heads,
                                     void* stream) {
  SWB_REQUIRE(raw && qscale && packed && invn, "swb200_qkv_pack_train: NULL pointer");
  return launch_qkv_pack_train(raw, qscale, packed, invn, M, heads, kHeadDim, kHeadDimPad, static_cast<cudaStream_t>(stream));
}

SWB200_API size_t swb200_muon_workspace_bytes(int rows, int cols, int batch) {
  return (rows >= 8 && cols >= 8 && batch >= 1) ? muon_workspace_bytes(rows, cols, batch) : 0;
}

SWB200_API int swb200_muon_step(float* const* params, const float* const* grads, float* const* momenta, int batch, int rows, int cols,
                                float lr, float weight_decay, float beta, int nesterov, int ns_steps, void* workspace,
                                size_t workspace_bytes, void* stream) {
  SWB_REQUIRE(params && grads && momenta && workspace, "swb200_muon_step: NULL pointer");
  return launch_muon_step(params, grads, momenta, batch, rows, cols, lr, weight_decay, beta, nesterov, ns_steps, workspace,
                          workspace_bytes, static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_muon_vector_step(float* param, const float* grad, float* momentum, int rows, int cols, float lr,
                                       float weight_decay, float beta, int nesterov, int ns_steps, void* stream) {
  SWB_REQUIRE(param && grad && momentum, "swb200_muon_vector_step: NULL pointer");
  return launch_muon_vector(param, grad, momentum, rows, cols, lr, weight_decay, beta, nesterov, ns_steps,
                            static_cast<cudaStream_t>(stream));
}

SWB200_API int swb200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, void* stream) {
  SWB_REQUIRE(param && grad && exp_avg && exp_avg_sq, "swb200_adam_step: NULL pointer");
  return launch_adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step,
                          static_cast<cudaStream_t>(stream));
}

// ---- checkpoint packing
namespace {
struct PackLayout {
  size_t w_embed, pos, aux_w, aux_b, l1_w, l1_b, l2_w, l2_b, mod_w, mod_b, ln_gamma, ln_beta, qscale, w_qkv, w_o, w_1, w_2, w_head,
      total;
};
PackLayout pack_layout(const swb200_model* m) {
  const Geom g = geom(m);
  const size_t D = m->dim, L = m->depth, Dff = m->dff, H = m->heads;
  PackLayout p;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  p.w_embed = take(D * g.k_embed_total * 2);
  p.pos = take(static_cast<size_t>(g.tokens) * D * 4);
  p.aux_w = take(D * std::max(1, m->aux_dim) * 4);
  p.aux_b = take(D * 4);
  p.l1_w = take(D * D * 4);
  p.l1_b = take(D * 4);
  p.l2_w = take(D * D * 4);
  p.l2_b = take(D * 4);
  p.mod_w = take(2 * L * 2 * D * D * 4);
  p.mod_b = take(2 * L * 2 * D * 4);
  p.ln_gamma = take(2 * L * D * 4);
  p.ln_beta = take(2 * L * D * 4);
  p.qscale = take(L * H * 4);
  p.w_qkv = take(L * 3 * D * D * 2);
  p.w_o = take(L * D * D * 2);
  p.w_1 = take(L * 2 * Dff * D * 2);
  p.w_2 = take(L * D * Dff * 2);
  p.w_head = take(static_cast<size_t>(m->out_channels) * g.pp * g.k_head_total * 2);
  p.total = off;
  return p;
}
}  // namespace

SWB200_API size_t swb200_packed_bytes(const swb200_model* m) {
  if (validate(m) != SWB_OK) return 0;
  return pack_layout(m).total;
}

static int pack_weights_impl(swb200_model* m, const swb200_ref_params* r, void* packed, size_t packed_bytes, void* stream_,
                             bool layer_matrices) {
  int rc = validate(m);
  if (rc) return rc;
  SWB_REQUIRE(r && packed, "pack_weights: NULL argument");
  SWB_REQUIRE(r->pos_embed && r->patch_w && r->patch_b && r->l1_w && r->l1_b && r->l2_w && r->l2_b && r->head_w && r->scale &&
                  r->attn_ln_w && r->attn_ln_b && r->attn_mod_w && r->attn_mod_b && r->to_qkv && r->wo && r->ff_ln_w && r->ff_ln_b &&
                  r->ff_mod_w && r->ff_mod_b && r->w1 && r->w2,
              "pack_weights: NULL parameter pointer");
  SWB_REQUIRE(m->aux_dim == 0 || (r->aux_w && r->aux_b), "pack_weights: aux_dim=%d but no auxiliary_embed parameters", m->aux_dim);
  const PackLayout p = pack_layout(m);
  SWB_REQUIRE(packed_bytes >= p.total && (reinterpret_cast<uintptr_t>(packed) & 255) == 0,
              "pack_weights: buffer too small (%zu < %zu) or not 256-byte aligned", packed_bytes, p.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, L = m->depth, Dff = m->dff, H = m->heads, F16 = m->act_fp16 ? 1 : 0;
  uint8_t* base = static_cast<uint8_t*>(packed);
  auto f32p = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  auto copy = [&](size_t off, const float* src, size_t n) -> int {
    SWB_CHECK_CUDA(cudaMemcpyAsync(base + off, src, n * 4, cudaMemcpyDeviceToDevice, st));
    return SWB_OK;
  };
  if ((rc = launch_pack_embed(r->patch_w, base + p.w_embed, D, m->in_channels, g.pp, m->k_embed, m->split_embed, F16, st))) return rc;
  if ((rc = launch_pack_pos(r->pos_embed, r->patch_b, f32p(p.pos), static_cast<long long>(g.tokens) * D, D, st))) return rc;
  if (m->aux_dim > 0) {
    if ((rc = copy(p.aux_w, r->aux_w, static_cast<size_t>(D) * m->aux_dim))) return rc;
    if ((rc = copy(p.aux_b, r->aux_b, D))) return rc;
  }
  if ((rc = copy(p.l1_w, r->l1_w, static_cast<size_t>(D) * D))) return rc;
  if ((rc = copy(p.l1_b, r->l1_b, D))) return rc;
  if ((rc = copy(p.l2_w, r->l2_w, static_cast<size_t>(D) * D))) return rc;
  if ((rc = copy(p.l2_b, r->l2_b, D))) return rc;
  const int half = kHeadDim * (m->gemm_tile == 3 ? 2 : 1);
  for (int l = 0; l < L; ++l) {
    const float* mw[2] = {r->attn_mod_w[l], r->ff_mod_w[l]};
    const float* mb[2] = {r->attn_mod_b[l], r->ff_mod_b[l]};
    const float* gw[2] = {r->attn_ln_w[l], r->ff_ln_w[l]};
    const float* gb[2] = {r->attn_ln_b[l], r->ff_ln_b[l]};
    for (int k = 0; k < 2; ++k) {
      const size_t i = 2 * static_cast<size_t>(l) + k;
      SWB_REQUIRE(mw[k] && mb[k] && gw[k] && gb[k], "pack_weights: NULL norm parameter in layer %d", l);
      if ((rc = copy(p.mod_w + i * 2 * D * D * 4, mw[k], static_cast<size_t>(2) * D * D))) return rc;
      if ((rc = copy(p.mod_b + i * 2 * D * 4, mb[k], static_cast<size_t>(2) * D))) return rc;
      if ((rc = copy(p.ln_gamma + i * D * 4, gw[k], D))) return rc;
      if ((rc = copy(p.ln_beta + i * D * 4, gb[k], D))) return rc;
    }
    SWB_REQUIRE(r->scale[l] && r->to_qkv[l] && r->wo[l] && r->w1[l] && r->w2[l], "pack_weights: NULL weight in layer %d", l);
    if ((rc = launch_pack_qscale(r->scale[l], f32p(p.qscale) + static_cast<size_t>(l) * H, H, st))) return rc;
    if (!layer_matrices) continue;               // the training model keeps its own (reference-order) copies of these four
    if ((rc = launch_pack_rows(r->to_qkv[l], base + p.w_qkv + static_cast<size_t>(l) * 3 * D * D * 2, 3 * D, D, D, 0, 1, H, kHeadDim, F16, st)))
      return rc;
    if ((rc = launch_pack_rows(r->wo[l], base + p.w_o + static_cast<size_t>(l) * D * D * 2, D, D, D, 0, 0, 0, 0, F16, st))) return rc;
    if ((rc = launch_pack_rows(r->w1[l], base + p.w_1 + static_cast<size_t>(l) * 2 * Dff * D * 2, 2 * Dff, D, D, 0, 2, half, Dff, F16, st)))
      return rc;
    if ((rc = launch_pack_rows(r->w2[l], base + p.w_2 + static_cast<size_t>(l) * D * Dff * 2, D, Dff, Dff, 0, 0, 0, 0, F16, st))) return rc;
  }
  if ((rc = launch_pack_rows(r->head_w, base + p.w_head, m->out_channels * g.pp, D, g.k_head_total, m->split_head ? D : 0, 0, 0, 0, F16, st)))
    return rc;
  m->w_embed = base + p.w_embed;
  m->b_embed = nullptr;
  m->pos_embed = f32p(p.pos);
  m->aux_w = m->aux_dim > 0 ? f32p(p.aux_w) : nullptr;
  m->aux_b = m->aux_dim > 0 ? f32p(p.aux_b) : nullptr;
  m->l1_w = f32p(p.l1_w);
  m->l1_b = f32p(p.l1_b);
  m->l2_w = f32p(p.l2_w);
  m->l2_b = f32p(p.l2_b);
  m->mod_w = f32p(p.mod_w);
  m->mod_b = f32p(p.mod_b);
  m->ln_gamma = f32p(p.ln_gamma);
  m->ln_beta = f32p(p.ln_beta);
  m->qscale = f32p(p.qscale);
  m->w_qkv = base + p.w_qkv;
  m->w_o = base + p.w_o;
  m->w_1 = base + p.w_1;
  m->w_2 = base + p.w_2;
  m->w_head = base + p.w_head;
  return SWB_OK;
}

SWB200_API int swb200_pack_weights(swb200_model* m, const swb200_ref_params* r, void* packed, size_t packed_bytes, void* stream) {
  return pack_weights_impl(m, r, packed, packed_bytes, stream, true);
}

// ---- checkpoint packing for the training path: the base model (bf16) + plain / transposed bf16 copies
namespace {
struct TrainPackLayout {
  size_t base, w_qkv, w_o, w_1, w_2, wt_qkv, wt_o, wt_1, wt_2, wt_head, total;
  int kp_head;
};
TrainPackLayout train_pack_layout(const swb200_model* m) {
  const Geom g = geom(m);
  const size_t D = m->dim, L = m->depth, Dff = m->dff;
  TrainPackLayout p;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  p.kp_head = (m->out_channels * g.pp + 7) / 8 * 8;
  p.base = take(pack_layout(m).total);
  p.w_qkv = take(L * 3 * D * D * 2);
  p.w_o = take(L * D * D * 2);
  p.w_1 = take(L * 2 * Dff * D * 2);
  p.w_2 = take(L * D * Dff * 2);
  p.wt_qkv = take(L * D * 3 * D * 2);
  p.wt_o = take(L * D * D * 2);
  p.wt_1 = take(L * D * 2 * Dff * 2);
  p.wt_2 = take(L * Dff * D * 2);
  p.wt_head = take(D * static_cast<size_t>(p.kp_head) * 2);
  p.total = off;
  return p;
}
}  // namespace

SWB200_API size_t swb200_train_packed_bytes(const swb200_train_model* tm) {
  if (tm == nullptr || validate(&tm->base) != SWB_OK) return 0;
  return train_pack_layout(&tm->base).total;
}

SWB200_API int swb200_pack_train_weights(swb200_train_model* tm, const swb200_ref_params* r, void* packed, size_t packed_bytes,
                                         void* stream_) {
  SWB_REQUIRE(tm && r && packed, "pack_train_weights: NULL argument");
  swb200_model* m = &tm->base;
  int rc = validate(m);
  if (rc) return rc;
  SWB_REQUIRE(m->act_fp16 == 0 && m->split_embed == 1 && m->split_head == 1,
              "pack_train_weights: the training path uses bf16 operands with split embed / head operands");
  const TrainPackLayout p = train_pack_layout(m);
  SWB_REQUIRE(packed_bytes >= p.total && (reinterpret_cast<uintptr_t>(packed) & 255) == 0,
              "pack_train_weights: buffer too small (%zu < %zu) or not 256-byte aligned", packed_bytes, p.total);
  uint8_t* base = static_cast<uint8_t*>(packed);
  if ((rc = pack_weights_impl(m, r, base + p.base, pack_layout(m).total, stream_, false))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const Geom g = geom(m);
  const int D = m->dim, L = m->depth, Dff = m->dff, Nh = m->out_channels * g.pp;
  for (int l = 0; l < L; ++l) {
    const size_t ls = static_cast<size_t>(l);
    if ((rc = launch_pack_rows(r->to_qkv[l], base + p.w_qkv + ls * 3 * D * D * 2, 3 * D, D, D, 0, 0, 0, 0, 0, st))) return rc;
    if ((rc = launch_pack_rows(r->wo[l], base + p.w_o + ls * D * D * 2, D, D, D, 0, 0, 0, 0, 0, st))) return rc;
    if ((rc = launch_pack_rows(r->w1[l], base + p.w_1 + ls * 2 * Dff * D * 2, 2 * Dff, D, D, 0, 0, 0, 0, 0, st))) return rc;
    if ((rc = launch_pack_rows(r->w2[l], base + p.w_2 + ls * D * Dff * 2, D, Dff, Dff, 0, 0, 0, 0, 0, st))) return rc;
    if ((rc = launch_pack_transposed(r->to_qkv[l], base + p.wt_qkv + ls * D * 3 * D * 2, 3 * D, D, 3 * D, 0, st))) return rc;
    if ((rc = launch_pack_transposed(r->wo[l], base + p.wt_o + ls * D * D * 2, D, D, D, 0, st))) return rc;
    if ((rc = launch_pack_transposed(r->w1[l], base + p.wt_1 + ls * D * 2 * Dff * 2, 2 * Dff, D, 2 * Dff, 0, st))) return rc;
    if ((rc = launch_pack_transposed(r->w2[l], base + p.wt_2 + ls * Dff * D * 2, D, Dff, D, 0, st))) return rc;
  }
  SWB_CHECK_CUDA(cudaMemsetAsync(base + p.wt_head, 0, static_cast<size_t>(D) * p.kp_head * 2, st));
  if ((rc = launch_pack_transposed(r->head_w, base + p.wt_head, Nh, D, p.kp_head, 0, st))) return rc;
  tm->w_qkv = base + p.w_qkv;
  tm->w_o = base + p.w_o;
  tm->w_1 = base + p.w_1;
  tm->w_2 = base + p.w_2;
  tm->wt_qkv = base + p.wt_qkv;
  tm->wt_o = base + p.wt_o;
  tm->wt_1 = base + p.wt_1;
  tm->wt_2 = base + p.wt_2;
  tm->wt_head = base + p.wt_head;
  tm->kp_head = p.kp_head;
  return SWB_OK;
}

}  // extern "C"
