/* swift_b200 -- C ABI of the B200 (sm_100a) implementation of the Swift forecast hot path.
 *
 * The reference (stockeh/swift) is pure Python/PyTorch and has no FFI of its own; its
 * plugin boundary for this path is the hydra `_target_` class `swift.models.swinv2.SwinV2`
 * (src/swift/configs/model/swinv2.yaml:1, instantiated at src/swift/models/precond.py:123-131) whose
 * `forward(x, t, auxiliary)` (src/swift/models/swinv2.py:305-330) is called by
 * `DiffusionSampler.scm_solver` / `dpm_solver_2s` (src/swift/generating/diffusion.py:417-461 / :355-415).
 * The entry points below are what a binding for that boundary needs: plain pointers and sizes, an opaque
 * cudaStream_t passed as void*, int return codes (0 = ok), no torch types.  The Python host layer
 * (swift_b200/swinv2.py, loaded through ctypes) is the binding used in this repo; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Ownership: the caller owns every buffer.  All pointers are DEVICE pointers unless stated otherwise and are
 * only borrowed for the duration of the call (the work is enqueued on `stream`; buffers must stay alive until
 * the stream reaches that point).  No function synchronises the device or allocates memory, so every call is
 * legal inside CUDA-graph stream capture.
 */
#ifndef SWIFT_B200_H_
#define SWIFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB200_ABI_VERSION 14
#if defined(__GNUC__)
#define SWB200_API __attribute__((visibility("default")))
#else
#define SWB200_API
#endif

/* Geometry + packed parameters of one SwinV2 denoiser (constructor arguments of swinv2.py:255-270).
 * Packed layouts are produced by swift_b200/packing.py (documented in DESIGN.md section 3):
 * ("h16" = fp16 when act_fp16 else bf16)
 *   b_embed : fp32 [dim] or NULL when the bias has been folded into pos_embed (what packing.py does)
 *   w_embed : h16 [dim, k_embed * (1 + split_embed)]   columns in "(c p1 p2)" order, zero padded to k_embed,
 *             duplicated when split_embed (the A operand is then [hi | lo], see swb200_forward)
 *   w_qkv   : h16  [depth][3*dim, dim]   rows reordered to  part*dim + head*88 + d   (part = q,k,v)
 *   w_o     : h16  [depth][dim, dim]
 *   w_1     : h16  [depth][2*dff, dim]   rows reordered per GEMM tile of T = 176 (gemm_tile 1, 2) or 352 (gemm_tile 3)
 *             rows: [T/2 gate rows | T/2 up rows], so gate and up of an output column share an accumulator row
 *   w_2     : h16  [depth][dim, dff]
 *   w_head  : h16  [out_channels*p1*p2, dim * (1 + split_head)]   rows in the reference "(c p1 p2)" order
 *   mod_w/b : fp32 [2*depth*2*dim, dim] / [2*depth*2*dim]   ModulatedNorm.modulation of layer l attention
 *             (index 2l) and feed-forward (index 2l+1), each [scale(dim) | shift(dim)]
 *   ln_gamma/ln_beta : fp32 [2*depth, dim]   LayerNorm affine, same order
 *   qscale  : fp32 [depth, heads]   exp(min(scale, ln 100))  (swinv2.py:125-126)                              */
typedef struct swb200_model {
  int32_t img_h, img_w, patch_h, patch_w, win_h, win_w, shift_h, shift_w;
  int32_t in_channels, out_channels, depth, dim, heads, dff, aux_dim;
  int32_t k_embed, split_embed, split_head;
  int32_t gemm_tile;          /* GEMM tile: 1 = 128x176 single CTA, 2 = 256x176 CTA pair, 3 = 256x352 CTA pair (default) */
  int32_t attn_impl;          /* window attention: 0 = auto (tcgen05 kernel for shifts that are multiples of 8), 1 = mma.sync, 2 = tcgen05 */
  int32_t act_fp16;           /* 16-bit tensor-core operand format of activations AND packed weights: 1 = fp16, 0 = bf16 */
  int32_t fuse_ln;            /* LayerNorm + modulation + residual add in the GEMM epilogue: bit 0 = wo, bit 1 = w2 (host default 3); 0 = separate kernel */
  int32_t attn_fp16;          /* with act_fp16 = 0: keep q / k / v and P in fp16 inside the attention (bounded by construction); ignored when act_fp16 */
  int32_t x_single;           /* with act_fp16 = 1: the forecast path keeps the residual stream as ONE fp16 value per element (the hi half of
                                 the [hi | lo] pair; lo is neither read nor written) -- 40 % less residual traffic for one extra 2^-11 rounding
                                 per update (Swift-B one step: 1.4e-3 -> 2.0e-3 per-field rel-L2); ignored when act_fp16 = 0 */
  float timestep_weight;
  const void* w_embed;
  const float* b_embed;
  const float* pos_embed;     /* [tokens, dim] */
  const float* aux_w;         /* [dim, aux_dim] or NULL */
  const float* aux_b;
  const float* l1_w;
  const float* l1_b;
  const float* l2_w;
  const float* l2_b;
  const float* mod_w;
  const float* mod_b;
  const float* ln_gamma;
  const float* ln_beta;
  const float* qscale;
  const void* w_qkv;
  const void* w_o;
  const void* w_1;
  const void* w_2;
  const void* w_head;
} swb200_model;

/* Output combination applied in the head epilogue (all NCHW fp32 [B, out_channels, img_h, img_w]):
 *     y = alpha * xt + beta * F + gamma * fprev          (terms with a NULL pointer are dropped)
 * and, when out_f != NULL, the raw network output F is stored there as well.
 *   module forward (swinv2.py:324):                  xt = NULL, beta = 1
 *   sCM step (diffusion.py:459):                     alpha = cos t, beta = -sin t * sigma_d, xt = x_t
 *   2S Euler (diffusion.py:402) / Heun (:410):       alpha = 1, beta = delta*sigma_d [*0.5], gamma = beta, fprev = F_s */
typedef struct swb200_update {
  const float* xt;
  const float* fprev;
  float* out_f;
  float alpha, beta, gamma;
  /* Rollout mode (state != NULL): the per-step glue of generate.py:120-131 (residual branch) is applied to y in the
   * same epilogue:  X_phys = X_std*x_std + x_mean + y*d_std;  X_std <- (X_phys - x_mean)/x_std  written IN PLACE into
   * the first out_channels channels of `state` [B, state_channels, H, W] (the condition buffer of the next step);
   * phys (optional) receives X_phys; channel `zero_channel` (>= 0) is forced to 0 (data/era5.py zero_field).
   * In this mode `y` of swb200_forward may be NULL. */
  float* state;
  int32_t state_channels;
  int32_t zero_channel;
  const float* x_std;
  const float* x_mean;
  const float* d_std;
  float* phys;
} swb200_update;

/* DEVICE pointers to the fp32 parameters of a reference checkpoint, by their `SwinV2.state_dict()` names
 * (models/swinv2.py:278-292): what swb200_pack_weights turns into the packed layouts above.
 * The per-layer members are HOST arrays of `depth` device pointers. */
typedef struct swb200_ref_params {
  const float* pos_embed;              /* pos_embed [1, tokens, dim] */
  const float* patch_w;                /* patch_embed.emb.weight [dim, p1*p2*in_channels], feature order (p1 p2 c) */
  const float* patch_b;                /* patch_embed.emb.bias [dim] */
  const float* aux_w;                  /* auxiliary_embed.weight [dim, aux_dim] or NULL */
  const float* aux_b;
  const float* l1_w;                   /* latent_embed.l1.weight [dim, dim] ... */
  const float* l1_b;
  const float* l2_w;
  const float* l2_b;
  const float* head_w;                 /* head.head.0.weight [out_channels*p1*p2, dim] */
  const float* const* scale;           /* transformer.layers.{l}.0.scale [1, heads, 1, 1] */
  const float* const* attn_ln_w;       /* ...{l}.0.norm.norm.weight [dim] */
  const float* const* attn_ln_b;
  const float* const* attn_mod_w;      /* ...{l}.0.norm.modulation.weight [2*dim, dim] */
  const float* const* attn_mod_b;
  const float* const* to_qkv;          /* ...{l}.0.to_qkv.weight [3*dim, dim] */
  const float* const* wo;              /* ...{l}.0.wo.weight [dim, dim] */
  const float* const* ff_ln_w;         /* ...{l}.1.norm.norm.weight */
  const float* const* ff_ln_b;
  const float* const* ff_mod_w;
  const float* const* ff_mod_b;
  const float* const* w1;              /* ...{l}.1.w1.weight [2*dff, dim] */
  const float* const* w2;              /* ...{l}.1.w2.weight [dim, dff] */
} swb200_ref_params;

/* ---- library ------------------------------------------------------------------------------------------ */
SWB200_API int swb200_abi_version(void);
SWB200_API const char* swb200_last_error(void);          /* message for the last non-zero return code on this thread */

/* Check that `m` is a configuration the kernels support (16x16 windows, head_dim 88, mlp dim % 88 == 0, ...).
 * Returns 0 or SWB200 error with swb200_last_error() set.  Host only. */
SWB200_API int swb200_validate(const swb200_model* m);

/* Build the packed model from a reference checkpoint.  In: the geometry and option fields of `m` (img_*, patch_*, win_*,
 * shift_*, in_channels, out_channels, depth, dim, heads, dff, aux_dim, k_embed, split_embed, split_head, gemm_tile,
 * attn_impl, act_fp16, fuse_ln, attn_fp16, timestep_weight); `ref`: the fp32 parameters on the device; `packed`: a device
 * buffer of swb200_packed_bytes(m) bytes, 256-byte aligned, that must outlive every use of `m`.  Out: every weight
 * pointer of `m` points into `packed` (b_embed = NULL: the bias is folded into the position table).  The conversions are
 * enqueued on `stream`. */
SWB200_API size_t swb200_packed_bytes(const swb200_model* m);
SWB200_API int swb200_pack_weights(swb200_model* m, const swb200_ref_params* ref, void* packed, size_t packed_bytes, void* stream);

/* Bytes of device workspace needed to push `chunk` samples through swb200_forward at once. Host only. */
SWB200_API size_t swb200_workspace_bytes(const swb200_model* m, int chunk);
/* Bytes of scratch for swb200_conditioning with batch B. Host only. */
SWB200_API size_t swb200_conditioning_scratch_bytes(const swb200_model* m, int B);

/* ---- the hot path ------------------------------------------------------------------------------------- */

/* Conditioning vectors (swinv2.py:316-321 + every ModulatedNorm.modulation, :84):
 *   t [B], aux [B, aux_dim] or NULL  ->  gain, bias : fp32 [2*depth, B, dim]  with
 *   gain = gamma*(1+scale(t)), bias = beta*(1+scale(t)) + shift(t).  cond_out (optional) = latent_embed output
 *   [B, dim].  For the 1-step sCM sampler (t = pi/2, aux = 0.6 fixed) this is computed once per rollout. */
SWB200_API int swb200_conditioning(const swb200_model* m, const float* t, const float* aux, int B, float* gain, float* bias,
                        float* cond_out, void* scratch, size_t scratch_bytes, void* stream);

/* One denoiser forward, SwinV2.forward (swinv2.py:305-330) with PassPrecond's channel concat
 * (precond.py:139-141) and the sampler update fused:
 *   network input = cat([x0 * scale0 (c0 channels), x1 (c1 channels)], dim=1), c0 + c1 == in_channels
 *   (x1 may be NULL when c1 == 0); gain/bias from swb200_conditioning with the same B;
 *   y = upd-combination of F (see swb200_update).  Samples are processed in chunks that fit `workspace`. */
SWB200_API int swb200_forward(const swb200_model* m, const float* x0, int c0, float scale0, const float* x1, int c1, int B,
                   const float* gain, const float* bias, const swb200_update* upd, float* y, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- individual kernels (unit tests, profiling) -------------------------------------------------------- */

/* D[M,N] = A[M,K] (row pitch lda) * W[N,K]^T (row pitch ldw), both fp16 if act_fp16 else both bf16, fp32 accumulate
 * on tcgen05.  epi: 0 store fp32 out[M,ldo], 1 store out[M,ldo] in the 16-bit operand format; profiling only:
 * 6 = accumulators discarded (main-loop rate), 7 = accumulators read out of TMEM but not stored, 8 = epilogue 1 without
 * its global stores, 9 = epilogue 1 with per-thread stores instead of the shared-memory transpose.
 * tile: 1 = 128x176 single CTA, 2 = 256x176 CTA pair (cta_group::2), 3 = 256x352 CTA pair. */
SWB200_API int swb200_gemm(int epi, int tile, int act_fp16, const void* A, int lda, const void* W, int ldw, void* out,
                int ldo, int M, int N, int K, void* stream);
/* qkv projection with fused scaled-cosine normalisation: out = 16-bit [3][heads][M][96] in fp16 when qkv_fp16 else bf16
 * (bf16 operands with an fp16 q/k/v output is the bf16 model with fp16 attention internals); W packed as w_qkv. */
SWB200_API int swb200_gemm_qkv(int tile, int act_fp16, int qkv_fp16, const void* A, int lda, const void* W, const float* qscale,
                    void* out, int M, int dim, int heads, void* stream);
/* SwiGLU up-projection: out[M, dff] = silu(gate) * up; W packed as w_1. */
SWB200_API int swb200_gemm_swiglu(int tile, int act_fp16, const void* A, int lda, const void* W, void* out, int M, int dim, int dff,
                       void* stream);
/* Patch-embed: x[M,dim] = A*W^T + bias + pos[row % tokens] (bias may be NULL), written as the 16-bit residual pair xhl[M, 2*dim] =
 * [hi | lo] with x = hi + lo (hi is the A operand of the next GEMM, row pitch 2*dim).
 * FORMAT WORD: in swb200_gemm_embed, swb200_gemm_ln_residual and swb200_ln_mod_residual the `act_fp16` argument is a bit set:
 * bit 0 = fp16 operands (else bf16), bit 1 (value 2, needs bit 0) = single-value residual stream: x is the hi half of xhl alone
 * (still at row pitch 2*dim), rounded to fp16 once per update; the lo half is neither read nor written (swb200_model.x_single). */
SWB200_API int swb200_gemm_embed(int tile, int act_fp16, const void* A, int lda, const void* W, int K, const float* bias,
                      const float* pos, int tokens, void* xhl, int M, int dim, void* stream);
/* Fused post-norm residual update (swinv2.py:83-86, :137-138, :100-101, :211-212):
 *     x <- x + LayerNorm(A W^T) * gain[b] + bias[b]     on the residual pair xhl[M, 2*dim] (in place), b = row / tokens,
 * A [M, K] (row pitch lda), W [dim, K]; gain / bias fp32 [B, dim].  The row statistics are exchanged between the CTAs
 * that hold the column tiles of a row through `ln_ws` (swb200_ln_workspace_bytes(M, dim) bytes, 256-byte aligned,
 * private to the stream).  `gen` numbers the launches that share ln_ws since it was last cleared: launch 0 clears it (a
 * memset node on `stream`); every launch tags the partials it publishes with 1 + gen % 7 (three low mantissa bits of M2), so
 * consecutive launches on one workspace must use consecutive `gen` values (a launch must never find its own tag left
 * behind by an earlier one).
 * Every CTA of the grid must be resident: the launcher checks the occupancy and returns -4 (nothing launched) when the
 * device cannot hold the grid -- swb200_forward then runs the same update as GEMM + swb200_ln_mod_residual; do not run
 * other kernels on the device concurrently. */
SWB200_API size_t swb200_ln_workspace_bytes(int M, int dim);
SWB200_API int swb200_gemm_ln_residual(int tile, int act_fp16, const void* A, int lda, const void* W, int K, void* xhl,
                            const float* gain, const float* bias, int M, int dim, int tokens, void* ln_ws, int gen,
                            void* stream);
/* Output head with pixel-shuffle + update; A is [M, K] with K = dim*(1+split). */
SWB200_API int swb200_gemm_head(int tile, const swb200_model* m, const void* A, int lda, int K, int B,
                     const swb200_update* upd, float* y, void* stream);
/* cat + patchify + bf16 cast (+ hi/lo split): A[B*tokens, lda] */
SWB200_API int swb200_patch_gather(const swb200_model* m, const float* x0, int c0, float scale0, const float* x1, int c1, int B,
                        void* A, int lda, void* stream);
/* x += LN(branch)*gain[b] + bias[b] on the residual pair xhl[M, 2*dim] = [hi | lo] (in place).
 * branch is fp32 [M, dim], or the 16-bit operand format when branch_16bit. */
SWB200_API int swb200_ln_mod_residual(const void* branch, int branch_16bit, void* xhl, const float* gain, const float* bias,
                           int M, int dim, int tokens, int act_fp16, void* stream);
/* shifted-window cosine attention on the packed qkv buffer (fp16 when qkv_fp16 else bf16; P uses the same format);
 * out 16-bit [M, heads*88], fp16 when out_fp16 else bf16 (out_fp16 needs qkv_fp16).
 * impl: 0 auto, 1 general-shift mma.sync kernel, 2 tcgen05/TMEM/TMA kernel (shift must be a multiple of 8).
 * lse (optional, tcgen05 kernel only): fp32 [heads][M], the log-sum-exp of every score row (saved for the backward). */
SWB200_API int swb200_window_attention(const void* qkv, void* out, int B, int grid_h, int grid_w, int heads, int shift_h,
                            int shift_w, int qkv_fp16, int out_fp16, int impl, float* lse, void* stream);

/* ---- tracing ---------------------------------------------------------------------------------------------- */

/* In-situ per-kernel timing of swb200_forward (CUDA events around every launch; not usable during graph capture).
 * The reference's counterpart is the optional torch.profiler block of training/trainer.py:155-177.
 * swb200_trace_enable(1) resets and starts; swb200_trace_report synchronises the device and writes a JSON object
 * {"kernel": {"ms": total, "launches": n}, ...} into buf. */
SWB200_API int swb200_trace_enable(int on);
SWB200_API int swb200_trace_report(char* buf, size_t buf_bytes);

/* fp16 range diagnostics.  Every fp16 conversion of an activation saturates at +-65504 instead of producing inf; with
 * `counters` set (DEVICE pointer to 6 zero-initialised uint64, NULL switches the diagnostics off again; process-global,
 * not thread-safe) every later swb200_forward adds, after the kernel that produced a tensor, the number of its elements
 * sitting at the saturation value:  [0] residual stream hi  [1] residual stream lo  [2] packed q/k/v  [3] attention output
 * [4] wo / w2 branch outputs (un-fused LayerNorm path only; the fused epilogue never stores them)  [5] SwiGLU hidden.
 * Non-zero counts mean the checkpoint's activations leave the fp16 range: run that model with act_fp16 = 0 (bf16). */
SWB200_API int swb200_debug_saturation(uint64_t* counters);

/* ---- rollout glue around the sampler (generate.py:97-118), graph-capturable ------------------------------ */

/* latents[b, i] ~ N(0,1): Philox4x32-10 keyed by seeds[b], counter (i/4, *step), Box-Muller.  Replaces the per-member
 * torch.Generator of generate.py:83 / factory.py:52-56 with a stream that depends only on (seed, step, element). */
SWB200_API int swb200_rollout_noise(float* latents, const uint64_t* seeds, const int32_t* step, int B, int64_t n_per_sample,
                         void* stream);
/* cond[b, state_channels + f, :] = table[base[b] + *step * stride, f, :] for f < n_forcings: the standardised forcings the
 * reference fetches per sample and step as `get_forcings(j + i*interval//6) for j in idx` (generate.py:100-117).  table:
 * fp32 [n_times, n_forcings, hw], indexed by time (6 h file index); base: int32 [B], the table row of every trajectory's
 * initial time (NULL = 0 for all); stride = interval // 6.  Rows outside the table are written as NaN (never read). */
SWB200_API int swb200_rollout_forcings(float* cond, int total_channels, int state_channels, const float* table, int n_forcings,
                            int n_times, const int32_t* base, int stride, const int32_t* step, int B, int hw,
                            void* stream);
/* *step += 1 (device-side, so a captured graph of one step can be replayed for the whole rollout) */
SWB200_API int swb200_rollout_advance(int32_t* step, void* stream);

/* ---- forward-mode tangent of the denoiser (sCM training loss, training/loss.py:216-225) ------------------- */

/* (F, dF) = jvp(SwinV2.forward, (x, t), (dx, dt)) with x = cat([x0*scale0, x1]) and dx = cat([dx0*scale0, 0]): what
 * `torch.func.jvp(lambda x, t: net(x, t, condition, auxiliary, jvp=True), (x_t/sigma_d, t), (v_x, v_t))` evaluates.  Every
 * Linear is the tcgen05 GEMM of the forecast path run over the row-stacked operand [x ; dx]; the derivative rules of
 * LayerNorm-modulation, q/k normalisation, windowed softmax attention and SwiGLU run in swift_b200/csrc/tangent.cu.
 * One sample at a time; workspace: swb200_jvp_workspace_bytes(m) bytes, 1024-byte aligned.
 * gain / bias / dgain / dbias: fp32 [2*depth, B, dim] from swb200_conditioning_jvp (t [B], dt [B] tangent of t). */
SWB200_API size_t swb200_jvp_workspace_bytes(const swb200_model* m);
SWB200_API size_t swb200_conditioning_jvp_scratch_bytes(const swb200_model* m, int B);
SWB200_API int swb200_conditioning_jvp(const swb200_model* m, const float* t, const float* dt, const float* aux, int B, float* gain,
                            float* bias, float* dgain, float* dbias, void* scratch, size_t scratch_bytes, void* stream);
SWB200_API int swb200_forward_jvp(const swb200_model* m, const float* x0, int c0, float scale0, const float* x1, int c1,
                       const float* dx0, int B, const float* gain, const float* bias, const float* dgain, const float* dbias,
                       float* y, float* dy, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the sCM training loss around the denoiser (training/loss.py:196-260; replaces its ~25 PyTorch elementwise ops) ---- */

/* Before the network: x_t = cos t x + sin t z (:203), dxt = cos t z - sin t x (:211), vx = cos t sin t dxt (:216 without
 * the 1/sigma_d, which swb200_forward_jvp applies as scale0), vt = cos t sin t (:217).  x, z, x_t, dxt, vx: fp32
 * [B, C, H, W] (W a multiple of 4, 16-byte aligned); t, vt: fp32 [B]. */
SWB200_API int swb200_scm_noised_inputs(const float* x, const float* z, const float* t, int B, int C, int H, int W, float* x_t,
                             float* dxt, float* vx, float* vt, void* stream);
/* After the network: g = -cos^2 (sd F - dxt) - r (cos sin x_t + sd dF) (:241-243), normalised per sample by
 * |g| sqrt(1/CHW) + 0.1 (:246-248); cot = dL/dF_x = -2 w_var[c] w_lat[h] g / (B H W) and loss = sum w g^2 / (B H W) -- the
 * value and output gradient of :253-260 with logvar = 0.  w_var [C] / w_lat [H] may be NULL (= 1).  loss: one fp32 on
 * the device.  scratch: swb200_scm_target_scratch_bytes(B) bytes, 8-byte aligned (fixed-order fp64 partial sums). */
SWB200_API size_t swb200_scm_target_scratch_bytes(int B);
SWB200_API int swb200_scm_tangent_target(const float* F, const float* dF, const float* x_t, const float* dxt, const float* t, float r,
                              float sigma_data, const float* w_var, const float* w_lat, int B, int C, int H, int W,
                              float* g, float* cot, float* loss, void* scratch, size_t scratch_bytes, void* stream);
/* Distillation (training/loss.py:205-210): with a pretrained v-prediction teacher, dx_t/dt = sigma_d * F_teacher(x_t / sigma_d, t)
 * replaces cos(t) z - sin(t) x; overwrites dxt and v_x = cos(t) sin(t) dx_t/dt / sigma_d of swb200_scm_noised_inputs. */
SWB200_API int swb200_scm_distill_direction(const float* F_teacher, const float* t, float sigma_data, int B, int C, int H, int W,
                                 float* dxt, float* vx, void* stream);
/* SCMLoss with a logvar head (training/loss.py:227-232, :252-258): L = mean_{b,h,w} sum_c [exp(-logvar_b) w g^2 + logvar_b];
 * cot = exp(-logvar_b) * (-2 w g) / (B H W), dlogvar [B] = dL/dlogvar_b = (C H W - exp(-logvar_b) sum_{c,h,w} w g^2) / (B H W). */
SWB200_API int swb200_scm_tangent_target_logvar(const float* F, const float* dF, const float* x_t, const float* dxt, const float* t,
                                     float r, float sigma_data, const float* w_var, const float* w_lat, int B, int C, int H,
                                     int W, const float* logvar, float* g, float* cot, float* loss, float* dlogvar,
                                     void* scratch, size_t scratch_bytes, void* stream);

/* ---- reverse mode: the grad-enabled forward + backward of the sCM training step ---------------------------
 * Reference: `F_x = net(x_t / sigma_d, t, condition, auxiliary)` under autograd (training/loss.py:226-232) followed by
 * `loss.backward()` (training/trainer.py:199-219).  The training path computes with bf16 tensor-core operands and fp32
 * accumulation / residual stream / gradients (the reference trains under bf16 autocast); `base.act_fp16` must be 0.
 * Every Linear's dgrad (dX = dY W) and wgrad (dW = dY^T X, split-K over the tokens) is the tcgen05 GEMM of the forecast
 * path; swift_b200/csrc/train.cu and attention_bwd.cu hold the derivative rules in between.
 *
 * Weights: `base` (geometry, fp32 conditioning parameters, bf16 w_embed / w_head packs as for swb200_forward) plus plain
 * bf16 copies of the four per-layer matrices in the REFERENCE's row order and their transposes (the dgrad operands):
 *   w_qkv [depth][3*dim, dim] rows (head, q|k|v, d) as `rearrange(.., "b n (h d) -> b h n d").chunk(3)` reads them
 *   (models/swinv2.py:119-121); w_1 [depth][2*dff, dim] rows [gate | up] (:99); w_o [depth][dim, dim]; w_2 [depth][dim, dff];
 *   wt_* = the same matrices transposed ([depth][K, N]); wt_head [dim, kp_head] = head weight^T, kp_head =
 *   out_channels*p1*p2 rounded up to 8 (zero padded). */
typedef struct swb200_train_model {
  swb200_model base;
  const void* w_qkv;
  const void* w_o;
  const void* w_1;
  const void* w_2;
  const void* wt_qkv;
  const void* wt_o;
  const void* wt_1;
  const void* wt_2;
  const void* wt_head;
  int32_t kp_head;
} swb200_train_model;

/* fp32 gradient buffers (DEVICE).  accumulate != 0: add to what the buffers hold (gradient accumulation over micro
 * batches), else overwrite.  Layouts:
 *   w_qkv / w_o / w_1 / w_2 : as the corresponding swb200_train_model matrices ([depth][N, K], reference row order)
 *   w_head  [out_channels*p1*p2, dim] (reference row order "(c p1 p2)")
 *   w_embed_t [k_embed, dim]: TRANSPOSED patch-embed weight gradient, rows in the packed "(c p1 p2)" input-feature order
 *             (rows >= in_channels*p1*p2 are padding); b_embed [dim]; pos_embed [tokens, dim]
 *   dscale [depth, heads]: gradient with respect to s = exp(min(scale, ln 100)) (the host applies ds/dscale)
 *   dgain / dbias [2*depth, B, dim]: gradients of the per-sample LayerNorm gain / bias vectors of swb200_conditioning,
 *             consumed by swb200_conditioning_backward. */
typedef struct swb200_train_grads {
  float* w_qkv;
  float* w_o;
  float* w_1;
  float* w_2;
  float* w_head;
  float* w_embed_t;
  float* b_embed;
  float* pos_embed;
  float* dscale;
  float* dgain;
  float* dbias;
  int32_t accumulate;
} swb200_train_grads;

/* fp32 gradients of the conditioning parameters (shapes of the swb200_model fields of the same names; mod_w / mod_b /
 * ln_gamma / ln_beta stacked over the 2*depth ModulatedNorms). */
typedef struct swb200_cond_grads {
  float* aux_w;
  float* aux_b;
  float* l1_w;
  float* l1_b;
  float* l2_w;
  float* l2_b;
  float* mod_w;
  float* mod_b;
  float* ln_gamma;
  float* ln_beta;
} swb200_cond_grads;

/* Build the training model from a reference checkpoint on the device (like swb200_pack_weights): in: the geometry / option
 * fields of m->base (act_fp16 = 0, split_embed = split_head = 1); out: every pointer of *m and kp_head, pointing into `packed`
 * (swb200_train_packed_bytes(m) bytes, 256-byte aligned, must outlive every use of *m). */
SWB200_API size_t swb200_train_packed_bytes(const swb200_train_model* m);
SWB200_API int swb200_pack_train_weights(swb200_train_model* m, const swb200_ref_params* ref, void* packed, size_t packed_bytes,
                              void* stream);
/* Bytes of the activation tape one grad-enabled forward of B samples writes (and the backward reads), and of the scratch
 * workspace shared by forward and backward; both 1024-byte aligned.  Host only. */
SWB200_API size_t swb200_train_tape_bytes(const swb200_train_model* m, int B);
SWB200_API size_t swb200_train_workspace_bytes(const swb200_train_model* m, int B);
/* F = SwinV2(cat([x0*scale0, x1])) like swb200_forward (y: NCHW fp32 [B, out_channels, H, W]) on the un-fused training
 * path, saving per layer the residual pairs, packed q/k/v + inverse norms, attention output, pre-LayerNorm branches,
 * gate / up and the SwiGLU hidden into `tape`. */
SWB200_API int swb200_train_forward(const swb200_train_model* m, const float* x0, int c0, float scale0, const float* x1, int c1,
                         int B, const float* gain, const float* bias, float* y, void* tape, size_t tape_bytes,
                         void* workspace, size_t workspace_bytes, void* stream);
/* The backward, in the order a caller overlaps it with gradient communication (trainer.py:76-84 wraps the net in DDP):
 *   head:  cot = dL/dF (NCHW fp32) -> w_head gradient, the gradient of the residual stream enters `workspace`;
 *   layer l = depth-1 .. 0: gradients of w_2, w_1, w_o, w_qkv, dscale, dgain / dbias rows 2l+1 and 2l;
 *   embed: w_embed_t, b_embed, pos_embed.
 * `workspace` carries the residual-stream gradient from call to call and must not be touched in between. */
SWB200_API int swb200_train_backward_head(const swb200_train_model* m, int B, const float* cot, const void* tape,
                               void* workspace, size_t workspace_bytes, const swb200_train_grads* grads, void* stream);
SWB200_API int swb200_train_backward_layer(const swb200_train_model* m, int layer, int B, const float* gain, const void* tape,
                                void* workspace, size_t workspace_bytes, const swb200_train_grads* grads, void* stream);
SWB200_API int swb200_train_backward_embed(const swb200_train_model* m, int B, const void* tape, void* workspace,
                                size_t workspace_bytes, const swb200_train_grads* grads, void* stream);
/* Conditioning backward: dgain / dbias [2*depth, B, dim] -> gradients of the LayerNorm affines, the 2*depth modulation
 * Linears, the latent MLP and the auxiliary embedding.  fwd_scratch: the scratch buffer swb200_conditioning was given for
 * the same (t, aux) batch, untouched since (it holds the embedding, the MLP activations and the modulation vectors).
 * scratch: swb200_conditioning_backward_scratch_bytes(m, B) bytes. */
SWB200_API size_t swb200_conditioning_backward_scratch_bytes(const swb200_model* m, int B);
SWB200_API int swb200_conditioning_backward(const swb200_model* m, const float* aux, int B, const void* fwd_scratch,
                                 const float* dgain, const float* dbias, const swb200_cond_grads* grads, int accumulate,
                                 void* scratch, size_t scratch_bytes, void* stream);
/* The same with a logvar head (models/swinv2.py:281, :326-327: logvar = logvar_embed(c), c the conditioning vector): lv_w [dim]
 * its weight, dlogvar [B] = dL/dlogvar; adds dlogvar_b * lv_w to the gradient of c before the latent MLP is differentiated and
 * writes (or accumulates) the head's own gradients g_lv_w [dim], g_lv_b [1].  lv_w = dlogvar = NULL: no head. */
SWB200_API int swb200_conditioning_backward_logvar(const swb200_model* m, const float* aux, int B, const void* fwd_scratch,
                                        const float* dgain, const float* dbias, const swb200_cond_grads* gr, int accumulate,
                                        void* scratch, size_t scratch_bytes, const float* lv_w, const float* dlogvar,
                                        float* g_lv_w, float* g_lv_b, void* stream);
/* logvar [B] = lv_w . c_b + lv_b from the forward scratch of swb200_conditioning (same (t, aux) batch, untouched since). */
SWB200_API int swb200_logvar_head(const swb200_model* m, const void* fwd_scratch, const float* lv_w, const float* lv_b, int B,
                       float* logvar, void* stream);
/* Unit-test entry points of the reverse-mode kernels (see swift_b200/csrc/kernels.h for the argument meaning). */
SWB200_API int swb200_gemm_splitk(int tile, const void* A, int lda, const void* W, int ldw, float* partials, int ldo, int M,
                       int N, int K_per_split, int splits, void* stream);
SWB200_API int swb200_transpose16(const void* in, int rows, int cols, int64_t ldi, void* out, int64_t ldo, void* stream);
SWB200_API size_t swb200_ln_backward_scratch_bytes(int M, int dim, int tokens);
SWB200_API int swb200_ln_backward(float* dx, const float* add, const float* branch, const float* gain, void* db16, float* dgain,
                       float* dbias, int M, int dim, int tokens, int accumulate, void* scratch, size_t scratch_bytes,
                       void* stream);
SWB200_API int swb200_swiglu_backward(const float* dh, const void* gu, void* dgu, int M, int dff, void* stream);
SWB200_API size_t swb200_attention_backward_scratch_bytes(int B, int grid_h, int grid_w, int heads);
/* impl 1: flash-style mma.sync kernels (any shift, recompute the row statistics); impl 2: the tcgen05 / TMEM kernel (shift a
 * multiple of 8; lse = the log-sum-exp rows the tcgen05 forward wrote). */
SWB200_API int swb200_attention_backward(const void* qkv, const void* O, const void* dO, const float* invn, const float* qscale,
                              const float* lse, int impl, void* dqkv, float* dscale, int B, int grid_h, int grid_w, int heads,
                              int shift_h, int shift_w, int accumulate, void* scratch, size_t scratch_bytes, void* stream);
SWB200_API int swb200_qkv_pack_train(const float* raw, const float* qscale, void* packed, float* invn, int M, int heads,
                          void* stream);

/* ---- optimiser step of the sCM training configuration (training/optimizers/muon.py: MuonWithAuxAdam) ------
 * Muon for `batch` 2-D parameters of ONE shape [rows, cols] (fp32, both extents multiples of 8; params / grads / momenta are
 * HOST arrays of `batch` device pointers): momentum <- lerp(momentum, grad, 1 - beta); update = nesterov ? lerp(grad,
 * momentum, beta) : momentum (muon.py:36-45); five (ns_steps) quintic Newton-Schulz steps in bf16 on the tcgen05 GEMM, all
 * matrices of the stack in one batched launch per product (muon.py:5-33); param <- param (1 - lr wd) -
 * lr max(1, rows/cols)^0.5 update (muon.py:42, :232-233).  grads are not modified.
 * workspace: swb200_muon_workspace_bytes(rows, cols, batch) bytes, 1024-byte aligned. */
SWB200_API size_t swb200_muon_workspace_bytes(int rows, int cols, int batch);
SWB200_API int swb200_muon_step(float* const* params, const float* const* grads, float* const* momenta, int batch, int rows, int cols,
                     float lr, float weight_decay, float beta, int nesterov, int ns_steps, void* workspace,
                     size_t workspace_bytes, void* stream);
/* The same update for a vector-shaped parameter (rows == 1 or cols == 1, at most 4096 elements; Swift-B: the [1, heads, 1, 1]
 * logit scales, which train.py:289 also hands to Muon): the Newton-Schulz products collapse to scalars. */
SWB200_API int swb200_muon_vector_step(float* param, const float* grad, float* momentum, int rows, int cols, float lr,
                            float weight_decay, float beta, int nesterov, int ns_steps, void* stream);
/* AuxAdam for one tensor of n elements (muon.py:147-152, :261-266); step counts from 1. */
SWB200_API int swb200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, void* stream);

/* ---- ensemble verification statistics on resident trajectories (eval/metrics.py:39-134) ------------------ */

/* phys [n_ic * members, n_var, H, W] physical-space forecasts, member m of initial condition j at row j*members + m
 * (the IC-major trajectory order of the rollout); truth [n_ic, n_var, H, W]; w_lat [H] = cos(lat) / mean(cos(lat)).
 * Writes, per (ic, var), the four latitude-weighted sums from which the reference's scores follow
 *   [0] sum w (mean_n p - y)^2   [1] sum_n sum w |p_n - y|   [2] sum_{i<j} sum w |p_i - p_j|   [3] sum w var_n(p)
 * as fp64 at out[(step ? *step : 0) * out_stride + (ic * n_var + var) * 4 + k]; `step` is a DEVICE pointer (the
 * rollout's step counter) so the call can sit inside the replayed CUDA graph of a 6 h step; `out` then has n_steps rows
 * and a counter outside [0, n_steps) writes nothing.
 *   rmse = mean_ic sqrt([0] / HW)    crps = mean_ic([1] / (N HW) - [2] / (N (N-1) HW))    ssr = mean_ic sqrt([3] / HW) / rmse */
SWB200_API int swb200_ensemble_stats(const float* phys, const float* truth, const float* w_lat, int n_ic, int members,
                          int n_var, int H, int W, const int32_t* step, int n_steps, int out_stride, double* out,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWIFT_B200_H_ */
