// Internal launcher declarations shared by the .cu files and the C-ABI layer (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace swb {

struct CondWeights {
  const float* aux_w;   // [D, aux_dim] or null
  const float* aux_b;   // [D]
  int aux_dim;
  const float* l1_w;    // [D, D]
  const float* l1_b;
  const float* l2_w;
  const float* l2_b;
  const float* mod_w;   // [L*2D, D]  all ModulatedNorm.modulation weights stacked in layer order
  const float* mod_b;   // [L*2D]
  const float* ln_gamma;  // [L, D]
  const float* ln_beta;   // [L, D]
};

int launch_patch_gather(const float* src0, int C0, float scale0, const float* src1, int C1, void* A, int lda, int Kp,
                        int split, int act_f16, int B, int H, int W, int p1, int p2, cudaStream_t stream);

// xhl: residual stream as a 16-bit [hi | lo] pair, [M, 2*D]; branch: fp32 or 16-bit [M, D]
int launch_ln_mod_residual(const void* branch, int branch_16bit, void* xhl, const float* gain, const float* bias, int M,
                           int D, int tokens, float eps, int act_f16, cudaStream_t stream);

// scratch needs B*(3*D + L*2*D) floats; gain/bias are [L, B, D]
int launch_conditioning(const CondWeights& w, const float* t, const float* aux, int B, int D, int L,
                        float timestep_weight, float* scratch, float* gain, float* bias, float* cond_out,
                        cudaStream_t stream);

// Shifted-window scaled-cosine attention on the packed qkv buffer [3][heads][M][96] (fp16/bf16, q/k already
// normalised and q scaled by the GEMM epilogue); out is [M, heads*88] in the same 16-bit format in un-shifted token order.
// impl: 0 = auto (tcgen05 kernel when the shift is a multiple of 8, else the mma.sync kernel), 1 = mma.sync, 2 = tcgen05
// act_f16: format of q / k / v (and of P inside the kernel); out_f16: format of `out` (fp16 out needs fp16 q/k/v)
// lse (optional, tcgen05 kernel only): fp32 [heads][M] log-sum-exp of every score row, saved for the backward
int launch_window_attention(const void* qkv, void* out, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                            int act_f16, int out_f16, int impl, cudaStream_t stream, float* lse = nullptr);
int launch_window_attention_tc(const void* qkv, void* out, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                               int act_f16, int out_f16, cudaStream_t stream, float* lse = nullptr);
int launch_attention_bwd_tc(const void* qkv, const void* O, const void* dO, const float* L, const float* invn, const float* qscale,
                            void* dqkv, float* Dbuf, float* ds_part, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                            cudaStream_t stream);

// elements of a fp16 [rows, cols] tensor (row pitch in elements) with |x| == 65504 are added to *counter
int launch_count_saturated_f16(const void* buf, long long rows, int cols, long long pitch, unsigned long long* counter,
                               cudaStream_t stream);

// rollout glue (rollout.cu)
int launch_rollout_noise(float* latents, const unsigned long long* seeds, const int* step, int B,
                         long long n_per_sample, cudaStream_t stream);
int launch_rollout_forcings(float* cond, int total_ch, int state_ch, const float* table, int n_forc, int n_times,
                            const int* base, int stride, const int* step, int B, int hw, cudaStream_t stream);
int launch_rollout_advance(int* step, cudaStream_t stream);

// forward-mode tangent companions (tangent.cu); "2" buffers hold primal rows 0..M-1 followed by tangent rows M..2M-1
int launch_ln_dual(const float* branch2, void* xhl2, const float* gain, const float* bias, const float* dgain,
                   const float* dbias, int M, int D, int tokens, float eps, int act_f16, cudaStream_t stream);
int launch_qkv_dual_pack(const float* raw2, const float* qscale, void* qkv, void* dqkv, int M, int D, int heads, int hd,
                         int pad, int act_f16, cudaStream_t stream);
int launch_attention_dual(const void* qkv, const void* dqkv, float* S, float* dS, void* attn2, int B, int gh, int gw,
                          int heads, int hd, int pad, int shift_h, int shift_w, int act_f16, cudaStream_t stream);
int launch_attention_dual_tc(const void* qkv, const void* dqkv, void* attn2, int B, int gh, int gw, int heads, int shift_h, int shift_w,
                             int act_f16, cudaStream_t stream);
int launch_swiglu_dual(const float* raw2, void* h2, int M, int Dff, int tile, int act_f16, cudaStream_t stream);
int launch_conditioning_dual(const CondWeights& w, const float* t, const float* dt, const float* aux, int B, int D, int L,
                             float timestep_weight, float* scratch, float* gain, float* bias, float* dgain, float* dbias,
                             cudaStream_t stream);

// ---- reverse mode (train.cu, attention_bwd.cu): see the headers of those files
struct CondGrads {          // fp32 gradients of the conditioning parameters, same shapes as CondWeights
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
};
int launch_swiglu_fwd_train(const void* gu, void* h, int M, int Dff, cudaStream_t stream);
int launch_swiglu_bwd(const float* dh, const void* gu, void* dgu, int M, int Dff, cudaStream_t stream);
int launch_qkv_pack_train(const float* raw, const float* qscale, void* packed, float* invn, int M, int heads, int hd,
                          int pad, cudaStream_t stream);
int launch_transpose16(const void* in, int R, int C, long long ldi, void* out, long long ldo, cudaStream_t stream);
size_t ln_bwd_partial_floats(int M, int D);       // + 2*B*D floats of reduction scratch behind it
int launch_ln_bwd(float* dx, const float* add, const float* branch, const float* gain, void* db16, float* part,
                  float* dgain, float* dbias, int M, int D, int tokens, float eps, int accumulate, cudaStream_t stream);
int launch_reduce_partials(const float* part, int per, int width, float* out, int groups, int accumulate,
                           cudaStream_t stream);
size_t colsum_partial_floats(int R, int W);
int launch_colsum(const float* in, int R, int W, float* part, float* out, int accumulate, cudaStream_t stream);
int launch_splitk_reduce(const float* part, int splits, long long stride, float* out, long long n, int accumulate,
                         cudaStream_t stream);
int launch_add_f32(const float* a, const float* b, float* dst, void* dst16, long long n, cudaStream_t stream);
int launch_cot_patchify(const float* cot, void* dF, int B, int C, int H, int W, int p1, int p2, int Kp, cudaStream_t stream);
int launch_sum_over_samples(const float* dx, int B, long long per, float* out, int accumulate, cudaStream_t stream);
size_t conditioning_bwd_scratch_floats(int B, int D, int L);
int launch_conditioning_bwd(const CondWeights& w, const CondGrads& g, const float* aux, const float* fwd_scratch,
                            const float* dgain, const float* dbias, int B, int D, int L, float* scratch, int accumulate,
                            cudaStream_t stream, const float* lv_w = nullptr, const float* dlogvar = nullptr,
                            float* g_lv_w = nullptr, float* g_lv_b = nullptr);
// logvar[b] = lv_w . c_b + lv_b on the conditioning vector c kept in the forward scratch of launch_conditioning
int launch_logvar_head(const float* fwd_scratch, const float* lv_w, const float* lv_b, int B, int D, float* logvar,
                       cudaStream_t stream);
int launch_attention_bwd(const void* qkv, const void* O, const void* dO, const float* invn, const float* qscale, void* dqkv,
                         float* Lbuf, float* Dbuf, float* ds_part, int B, int gh, int gw, int heads, int hd, int pad,
                         int shift_h, int shift_w, cudaStream_t stream);

// checkpoint packing (pack.cu); mode 0 plain, 1 qkv row order (a = heads, b = head dim), 2 w1 tile interleave (a = half, b = dff)
int launch_pack_rows(const float* src, void* dst, int rows, int K, int ldd, int dup, int mode, int a, int b, int f16,
                     cudaStream_t st);
int launch_pack_transposed(const float* src, void* dst, int N, int K, int ldd, int f16, cudaStream_t st);
int launch_pack_embed(const float* src, void* dst, int D, int C, int pp, int k_embed, int split, int f16, cudaStream_t st);
int launch_pack_pos(const float* pos, const float* bias, float* out, long long n, int D, cudaStream_t st);
int launch_pack_qscale(const float* scale, float* out, int heads, cudaStream_t st);

// optimiser step (muon.cu): Muon's momentum + Newton-Schulz orthogonalisation + parameter update for one matrix; AuxAdam
size_t muon_workspace_bytes(int rows, int cols, int batch);
int launch_muon_step(float* const* param, const float* const* grad, float* const* momentum, int batch, int rows, int cols, float lr,
                     float weight_decay, float beta, int nesterov, int ns_steps, void* workspace, size_t ws_bytes, cudaStream_t st);
int launch_muon_vector(float* param, const float* grad, float* momentum, int rows, int cols, float lr, float weight_decay, float beta,
                       int nesterov, int ns_steps, cudaStream_t st);
int launch_transpose16_batched(const void* in, int R, int C, void* out, int batch, cudaStream_t stream);
int launch_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
                     int step, cudaStream_t st);

// ensemble verification statistics (ensemble.cu): out[(*step) * out_stride + (ic * V + v) * 4 + k]
int launch_ensemble_stats(const float* phys, const float* truth, const float* w_lat, int n_ic, int members, int V, int H,
                          int W, const int* step, int n_steps, int out_stride, double* out, cudaStream_t stream);

// sCM training-loss glue (scm_target.cu)
int launch_scm_noised_inputs(const float* x, const float* z, const float* t, int B, int C, int H, int W, float* x_t,
                             float* dxt, float* vx, float* vt, cudaStream_t stream);
size_t scm_target_scratch_bytes(int B);
int launch_scm_distill_direction(const float* F_teacher, const float* t, float sigma_data, int B, int C, int H, int W, float* dxt,
                                 float* vx, cudaStream_t stream);
int launch_scm_tangent_target(const float* F, const float* dF, const float* x_t, const float* dxt, const float* t, float r,
                              float sigma_data, const float* w_var, const float* w_lat, int B, int C, int H, int W, float* g,
                              float* cot, float* loss, void* scratch, size_t scratch_bytes, cudaStream_t stream,
                              const float* logvar = nullptr, float* dlogvar = nullptr);

}  // namespace swb
