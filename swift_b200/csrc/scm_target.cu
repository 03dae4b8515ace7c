// The elementwise / reduction half of the sCM training loss around the denoiser (reference: training/loss.py:196-260),
// as two entry points on either side of swb200_forward_jvp:
//
//   scm_noised_inputs   x_t = cos t x + sin t z          (:203)      network input (scaled by 1/sigma_d in the patch gather)
//                       dxt = cos t z - sin t x          (:211)
//                       vx  = cos t sin t dxt            (:216, the 1/sigma_d again folded into the gather's scale0)
//                       vt  = cos t sin t                (:217)
//   scm_tangent_target  g   = -cos^2 (sd F - dxt) - r (cos sin x_t + sd dF)                        (:241-243)
//                       g  /= |g|_b sqrt(1 / CHW) + 0.1                                             (:246-248)
//                       cot = dL/dF = -2 w_var w_lat g / (B H W);   loss = sum w g^2 / (B H W)      (:253-260, logvar = 0)
//
// HBM-bound streaming kernels (float4 accesses, 5 reads + 2 writes per element in the target pass); the per-sample norm
// and the loss are reduced in a fixed order through fp64 partial sums (bit-reproducible, no atomics).
#include "common.h"
#include "kernels.h"

namespace swb {

constexpr int kScmThreads = 256;
constexpr int kScmMaxBlocks = 128;      // blocks per sample (partials reduced in order by the finishing kernels)

__global__ void __launch_bounds__(kScmThreads) scm_noised_inputs_kernel(const float4* __restrict__ x, const float4* __restrict__ z,
                                                                        const float* __restrict__ t, size_t n4,
                                                                        float4* __restrict__ x_t, float4* __restrict__ dxt,
                                                                        float4* __restrict__ vx, float* __restrict__ vt) {
  const int b = blockIdx.y;
  float s, c;
  sincosf(t[b], &s, &c);
  const float cs = c * s;
  if (blockIdx.x == 0 && threadIdx.x == 0) vt[b] = cs;
  const size_t base = static_cast<size_t>(b) * n4;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 xv = __ldg(x + base + i), zv = __ldg(z + base + i);
    float4 a, d, v;
    a.x = c * xv.x + s * zv.x; a.y = c * xv.y + s * zv.y; a.z = c * xv.z + s * zv.z; a.w = c * xv.w + s * zv.w;
    d.x = c * zv.x - s * xv.x; d.y = c * zv.y - s * xv.y; d.z = c * zv.z - s * xv.z; d.w = c * zv.w - s * xv.w;
    v.x = cs * d.x; v.y = cs * d.y; v.z = cs * d.z; v.w = cs * d.w;
    x_t[base + i] = a;
    dxt[base + i] = d;
    vx[base + i] = v;
  }
}

__device__ __forceinline__ double block_sum(double a, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) red[warp] = a;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kScmThreads / 32; ++w) r += red[w];
  __syncthreads();
  return r;      // valid in thread 0
}

// pass 1: g_raw and the per-(sample, block) partial sums of g_raw^2
__global__ void __launch_bounds__(kScmThreads) scm_target_raw_kernel(const float4* __restrict__ F, const float4* __restrict__ dF,
                                                                     const float4* __restrict__ x_t, const float4* __restrict__ dxt,
                                                                     const float* __restrict__ t, float r, float sd, size_t n4,
                                                                     float4* __restrict__ g, double* __restrict__ part) {
  __shared__ double red[kScmThreads / 32];
  const int b = blockIdx.y;
  float s, c;
  sincosf(t[b], &s, &c);
  const float c2 = c * c, cs = c * s;
  const size_t base = static_cast<size_t>(b) * n4;
  double acc = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 f = __ldg(F + base + i), df = __ldg(dF + base + i), xt = __ldg(x_t + base + i), d = __ldg(dxt + base + i);
    float4 o;
    o.x = -c2 * (sd * f.x - d.x) - r * (cs * xt.x + sd * df.x);
    o.y = -c2 * (sd * f.y - d.y) - r * (cs * xt.y + sd * df.y);
    o.z = -c2 * (sd * f.z - d.z) - r * (cs * xt.z + sd * df.z);
    o.w = -c2 * (sd * f.w - d.w) - r * (cs * xt.w + sd * df.w);
    g[base + i] = o;
    acc += static_cast<double>(o.x * o.x + o.y * o.y) + static_cast<double>(o.z * o.z + o.w * o.w);
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) part[static_cast<size_t>(b) * gridDim.x + blockIdx.x] = tot;
}

// pass 2: normalise, cotangent, per-(sample, block) partial sums of w g^2
__global__ void __launch_bounds__(kScmThreads) scm_target_finish_kernel(float4* __restrict__ g, float4* __restrict__ cot,
                                                                        const double* __restrict__ part, int nparts,
                                                                        const float* __restrict__ w_var,
                                                                        const float* __restrict__ w_lat, int C, int H, int W4,
                                                                        float inv_bhw, const float* __restrict__ logvar,
                                                                        double* __restrict__ loss_part) {
  __shared__ double red[kScmThreads / 32];
  __shared__ float inv_den;
  const int b = blockIdx.y;
  const size_t n4 = static_cast<size_t>(C) * H * W4;
  if (threadIdx.x == 0) {
    double ss = 0.0;
    for (int k = 0; k < nparts; ++k) ss += part[static_cast<size_t>(b) * nparts + k];
    const float gn = static_cast<float>(sqrt(ss));
    // gn.numel() / g.numel() = B / (B C H W): the norm is made invariant to the field size (loss.py:247)
    inv_den = 1.0f / (gn * sqrtf(1.0f / static_cast<float>(n4 * 4)) + 0.1f);
  }
  __syncthreads();
  const float inv = inv_den;
  // logvar head (loss.py:227-232, :252-258): the squared term of sample b is weighted by exp(-logvar_b)
  const float elv = logvar ? expf(-__ldg(logvar + b)) : 1.0f;
  const size_t base = static_cast<size_t>(b) * n4;
  const size_t hw4 = static_cast<size_t>(H) * W4;
  double acc = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i / hw4);
    const int h = static_cast<int>((i - static_cast<size_t>(ch) * hw4) / W4);
    const float w = (w_var ? __ldg(w_var + ch) : 1.0f) * (w_lat ? __ldg(w_lat + h) : 1.0f);
    float4 v = g[base + i];
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    g[base + i] = v;
    const float k = -2.0f * w * inv_bhw * elv;
    cot[base + i] = make_float4(k * v.x, k * v.y, k * v.z, k * v.w);
    acc += static_cast<double>(w) * (static_cast<double>(v.x * v.x + v.y * v.y) + static_cast<double>(v.z * v.z + v.w * v.w));
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) loss_part[static_cast<size_t>(b) * gridDim.x + blockIdx.x] = tot;
}

// loss = mean_{b,h,w} sum_c [exp(-lv_b) w g^2 + lv_b];  dL/dlv_b = (-exp(-lv_b) S_b + C H W) / (B H W),  S_b = sum_{c,h,w} w g^2
// (fixed summation order; without a logvar head lv = 0 and the second term vanishes)
__global__ void scm_loss_sum_kernel(const double* __restrict__ loss_part, int B, int nb, float inv_bhw, double chw,
                                    const float* __restrict__ logvar, float* __restrict__ dlogvar, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double tot = 0.0;
    for (int b = 0; b < B; ++b) {
      double sb = 0.0;
      for (int k = 0; k < nb; ++k) sb += loss_part[static_cast<size_t>(b) * nb + k];
      if (logvar) {
        const double lv = static_cast<double>(logvar[b]), e = exp(-lv);
        tot += e * sb + chw * lv;
        if (dlogvar) dlogvar[b] = static_cast<float>((chw - e * sb) * static_cast<double>(inv_bhw));
      } else {
        tot += sb;
      }
    }
    *loss = static_cast<float>(tot * static_cast<double>(inv_bhw));
  }
}

// distillation (loss.py:205-210): dx_t/dt = sigma_d * F_teacher replaces cos z - sin x, and with it v_x = cos sin dx_t/dt / sigma_d
__global__ void __launch_bounds__(kScmThreads) scm_distill_direction_kernel(const float4* __restrict__ Ft, const float* __restrict__ t,
                                                                            float sd, size_t n4, float4* __restrict__ dxt,
                                                                            float4* __restrict__ vx) {
  const int b = blockIdx.y;
  float s, c;
  sincosf(t[b], &s, &c);
  const float cs = c * s;
  const size_t base = static_cast<size_t>(b) * n4;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 f = __ldg(Ft + base + i);
    dxt[base + i] = make_float4(sd * f.x, sd * f.y, sd * f.z, sd * f.w);
    vx[base + i] = make_float4(cs * f.x, cs * f.y, cs * f.z, cs * f.w);
  }
}

static int blocks_per_sample(size_t n4) {
  const size_t want = (n4 + kScmThreads * 4 - 1) / (kScmThreads * 4);
  return static_cast<int>(want < 1 ? 1 : (want > kScmMaxBlocks ? kScmMaxBlocks : want));
}

int launch_scm_noised_inputs(const float* x, const float* z, const float* t, int B, int C, int H, int W, float* x_t,
                             float* dxt, float* vx, float* vt, cudaStream_t stream) {
  SWB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0, "scm_noised_inputs: need B, C, H > 0 and W a multiple of 4 (W=%d)", W);
  const size_t n4 = static_cast<size_t>(C) * H * (W / 4);
  dim3 grid(blocks_per_sample(n4), B);
  scm_noised_inputs_kernel<<<grid, kScmThreads, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(z),
                                                            t, n4, reinterpret_cast<float4*>(x_t), reinterpret_cast<float4*>(dxt),
                                                            reinterpret_cast<float4*>(vx), vt);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

int launch_scm_distill_direction(const float* F_teacher, const float* t, float sigma_data, int B, int C, int H, int W, float* dxt,
                                 float* vx, cudaStream_t stream) {
  SWB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0, "scm_distill_direction: need B, C, H > 0 and W a multiple of 4 (W=%d)", W);
  const size_t n4 = static_cast<size_t>(C) * H * (W / 4);
  dim3 grid(blocks_per_sample(n4), B);
  scm_distill_direction_kernel<<<grid, kScmThreads, 0, stream>>>(reinterpret_cast<const float4*>(F_teacher), t, sigma_data, n4,
                                                                reinterpret_cast<float4*>(dxt), reinterpret_cast<float4*>(vx));
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

size_t scm_target_scratch_bytes(int B) { return static_cast<size_t>(2) * B * kScmMaxBlocks * sizeof(double); }

int launch_scm_tangent_target(const float* F, const float* dF, const float* x_t, const float* dxt, const float* t, float r,
                              float sigma_data, const float* w_var, const float* w_lat, int B, int C, int H, int W, float* g,
                              float* cot, float* loss, void* scratch, size_t scratch_bytes, cudaStream_t stream,
                              const float* logvar, float* dlogvar) {
  SWB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0, "scm_tangent_target: need B, C, H > 0 and W a multiple of 4 (W=%d)", W);
  SWB_REQUIRE(scratch_bytes >= scm_target_scratch_bytes(B), "scm_tangent_target: scratch of %zu bytes, need %zu", scratch_bytes,
              scm_target_scratch_bytes(B));
  const size_t n4 = static_cast<size_t>(C) * H * (W / 4);
  const int nb = blocks_per_sample(n4);
  double* part = static_cast<double*>(scratch);
  double* loss_part = part + static_cast<size_t>(B) * kScmMaxBlocks;
  dim3 grid(nb, B);
  const float inv_bhw = 1.0f / (static_cast<float>(B) * H * W);
  scm_target_raw_kernel<<<grid, kScmThreads, 0, stream>>>(reinterpret_cast<const float4*>(F), reinterpret_cast<const float4*>(dF),
                                                         reinterpret_cast<const float4*>(x_t), reinterpret_cast<const float4*>(dxt),
                                                         t, r, sigma_data, n4, reinterpret_cast<float4*>(g), part);
  SWB_CHECK_CUDA(cudaGetLastError());
  scm_target_finish_kernel<<<grid, kScmThreads, 0, stream>>>(reinterpret_cast<float4*>(g), reinterpret_cast<float4*>(cot), part, nb,
                                                            w_var, w_lat, C, H, W / 4, inv_bhw, logvar, loss_part);
  SWB_CHECK_CUDA(cudaGetLastError());
  scm_loss_sum_kernel<<<1, 32, 0, stream>>>(loss_part, B, nb, inv_bhw, static_cast<double>(C) * H * W, logvar, dlogvar, loss);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
