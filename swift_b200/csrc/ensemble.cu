// Ensemble verification statistics on resident trajectories: the per-(initial condition,
// variable) sufficient statistics of the reference's scores (eval/metrics.py:39-134) computed from the physical state
// the rollout step has just written, so forecasts never have to leave the GPU to be scored.
//
//   out[ic, v, 0] = sum_hw w_h (mean_n p - y)^2          -> lat-weighted RMSE of the ensemble mean   (metrics.py:48-65)
//   out[ic, v, 1] = sum_n sum_hw w_h |p_n - y|           -> CRPS error term                            (:86-89)
//   out[ic, v, 2] = sum_{i<j} sum_hw w_h |p_i - p_j|     -> CRPS spread term (half of the ordered sum) (:92-98)
//   out[ic, v, 3] = sum_hw w_h var_n(p)  (unbiased)      -> spread of the spread / skill ratio         (:125-127)
//
// Trajectories are laid out IC-major (rollout.shard_trajectories): member m of local IC j is row j*N + m of `phys`.
// One block per (variable, IC); every thread walks pixels with coalesced loads of the N member planes; fp64 partial sums,
// fixed-order block reduction (bit-reproducible).  HBM-bound: N*V*H*W*4 bytes per IC, read once.
#include "common.h"
#include "kernels.h"

namespace swb {

constexpr int kMaxMembers = 32;

template <int NMAX>
__global__ void __launch_bounds__(256) ensemble_stats_kernel(const float* __restrict__ phys, const float* __restrict__ truth,
                                                             const float* __restrict__ w_lat, int N, int V, int H, int W,
                                                             const int* __restrict__ step, int n_steps,
                                                             int out_stride, double* __restrict__ out) {
  const int v = blockIdx.x, ic = blockIdx.y;
  if (step && (*step < 0 || *step >= n_steps)) return;      // a row outside `out` (host bookkeeping error): never write
  const size_t hw = static_cast<size_t>(H) * W;
  const float* p0 = phys + (static_cast<size_t>(ic) * N * V + v) * hw;
  const size_t mstride = static_cast<size_t>(V) * hw;
  const float* y0 = truth + (static_cast<size_t>(ic) * V + v) * hw;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const float inv_n = 1.0f / static_cast<float>(N);
  const float inv_nm1 = N > 1 ? 1.0f / static_cast<float>(N - 1) : 0.f;
  for (size_t idx = threadIdx.x; idx < hw; idx += blockDim.x) {
    float p[NMAX];
#pragma unroll
    for (int m = 0; m < NMAX; ++m) p[m] = m < N ? __ldg(p0 + m * mstride + idx) : 0.f;
    const float y = __ldg(y0 + idx);
    const float w = __ldg(w_lat + idx / W);
    float s = 0.f, ae = 0.f;
#pragma unroll
    for (int m = 0; m < NMAX; ++m)
      if (m < N) {
        s += p[m];
        ae += fabsf(p[m] - y);
      }
    const float mean = s * inv_n;
    float ss = 0.f, pd = 0.f;
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
      if (i < N) {
        const float d = p[i] - mean;
        ss = fmaf(d, d, ss);
#pragma unroll
        for (int j = i + 1; j < NMAX; ++j)
          if (j < N) pd += fabsf(p[i] - p[j]);
      }
    acc[0] += static_cast<double>(w * (mean - y) * (mean - y));
    acc[1] += static_cast<double>(w * ae);
    acc[2] += static_cast<double>(w * pd);
    acc[3] += static_cast<double>(w * ss * inv_nm1);
  }
  __shared__ double red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double a = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[k][warp] = a;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int wdx = 0; wdx < 8; ++wdx) a += red[threadIdx.x][wdx];
    const size_t base = step ? static_cast<size_t>(*step) * out_stride : 0;
    out[base + (static_cast<size_t>(ic) * V + v) * 4 + threadIdx.x] = a;
  }
}

int launch_ensemble_stats(const float* phys, const float* truth, const float* w_lat, int n_ic, int members, int V, int H,
                          int W, const int* step, int n_steps, int out_stride, double* out, cudaStream_t stream) {
  SWB_REQUIRE(members >= 1 && members <= kMaxMembers, "ensemble_stats: %d members unsupported (1..%d)", members, kMaxMembers);
  SWB_REQUIRE(n_ic > 0 && V > 0 && H > 0 && W > 0, "ensemble_stats: empty problem");
  dim3 grid(V, n_ic);
  if (members <= 16)
    ensemble_stats_kernel<16><<<grid, 256, 0, stream>>>(phys, truth, w_lat, members, V, H, W, step, n_steps, out_stride, out);
  else
    ensemble_stats_kernel<32><<<grid, 256, 0, stream>>>(phys, truth, w_lat, members, V, H, W, step, n_steps, out_stride, out);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
