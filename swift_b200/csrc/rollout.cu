// Device-side pieces of the autoregressive rollout step that sit around the denoiser (generate.py:97-118):
// per-trajectory N(0,1) latents, the per-step forcings channels of the condition, and a device step counter so that
// one captured CUDA graph can be replayed for every 6 h step.
#include "common.h"
#include "kernels.h"

namespace swb {

// ---------------------------------------------------------------------------------------------------------
// Counter-based noise: latents[b, i] = N(0,1) from Philox4x32-10 keyed by the trajectory seed, counter = (i/4, step).
// The value of element i of trajectory (member, ic) at step s therefore depends on nothing else: not on the batch
// composition, the chunking or the number of GPUs (the reference's torch.Generator stream is consumed in batch
// order, generate.py:83-118).  oracle/philox_oracle.py restates the same function in numpy.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return (static_cast<float>(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(256) rollout_noise_kernel(float* __restrict__ latents,
                                                            const unsigned long long* __restrict__ seeds,
                                                            const int* __restrict__ step, long long n4_per_sample) {
  const int b = blockIdx.y;
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;     // group of 4 elements
  if (q >= n4_per_sample) return;
  const unsigned long long seed = seeds[b];
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), static_cast<uint32_t>(*step), 0u,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  // Box-Muller on two uniform pairs
  const float r0 = sqrtf(-2.0f * logf(u01(r[0]))), r1 = sqrtf(-2.0f * logf(u01(r[2])));
  const float t0 = 6.283185307179586f * u01(r[1]), t1 = 6.283185307179586f * u01(r[3]);
  float4 z = make_float4(r0 * cosf(t0), r0 * sinf(t0), r1 * cosf(t1), r1 * sinf(t1));
  reinterpret_cast<float4*>(latents + static_cast<size_t>(b) * n4_per_sample * 4)[q] = z;
}

int launch_rollout_noise(float* latents, const unsigned long long* seeds, const int* step, int B,
                         long long n_per_sample, cudaStream_t stream) {
  SWB_REQUIRE(n_per_sample % 4 == 0 && (reinterpret_cast<uintptr_t>(latents) & 15) == 0,
              "rollout_noise: elements per sample must be a multiple of 4 and the buffer 16-byte aligned");
  const long long n4 = n_per_sample / 4;
  dim3 grid(static_cast<unsigned>((n4 + 255) / 256), B);
  rollout_noise_kernel<<<grid, 256, 0, stream>>>(latents, seeds, step, n4);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

// cond[b, state_ch + f, :] = table[base[b] + step * stride, f, :]
// The reference appends, every step i and for every sample of the batch, the standardised forcings of the file
// `j + i * interval // 6` where j is the dataset (time) index of that sample's initial condition (generate.py:100-117):
// ICs have different valid times, so the table is indexed by TIME (6 h file index relative to its first row) and every
// trajectory carries the row of its initial time in `base` (NULL: every trajectory starts at row 0).  A row outside the
// table is a host-side bookkeeping error (the host checks before launching); the kernel then writes NaN, never reads
// out of bounds.
__global__ void __launch_bounds__(256) rollout_forcings_kernel(float* __restrict__ cond, int total_ch, int state_ch,
                                                               const float* __restrict__ table, int n_forc, int n_times,
                                                               const int* __restrict__ base, int stride,
                                                               const int* __restrict__ step, int hw4) {
  const int b = blockIdx.z, f = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw4) return;
  const long long row = static_cast<long long>(base ? base[b] : 0) + static_cast<long long>(*step) * stride;
  float4 v;
  if (row < 0 || row >= n_times) {
    const float nan = __int_as_float(0x7fc00000);
    v = make_float4(nan, nan, nan, nan);
  } else {
    v = __ldg(reinterpret_cast<const float4*>(table + (static_cast<size_t>(row) * n_forc + f) * hw4 * 4) + i);
  }
  reinterpret_cast<float4*>(cond + (static_cast<size_t>(b) * total_ch + state_ch + f) * hw4 * 4)[i] = v;
}

int launch_rollout_forcings(float* cond, int total_ch, int state_ch, const float* table, int n_forc, int n_times,
                            const int* base, int stride, const int* step, int B, int hw, cudaStream_t stream) {
  SWB_REQUIRE(hw % 4 == 0 && state_ch + n_forc <= total_ch && n_times > 0 && stride >= 0,
              "rollout_forcings: bad channel layout / table (hw=%d, channels %d+%d of %d, %d rows, stride %d)", hw,
              state_ch, n_forc, total_ch, n_times, stride);
  if (n_forc == 0) return SWB_OK;
  dim3 grid((hw / 4 + 255) / 256, n_forc, B);
  rollout_forcings_kernel<<<grid, 256, 0, stream>>>(cond, total_ch, state_ch, table, n_forc, n_times, base, stride, step,
                                                    hw / 4);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

__global__ void rollout_advance_kernel(int* step) { *step += 1; }

int launch_rollout_advance(int* step, cudaStream_t stream) {
  rollout_advance_kernel<<<1, 1, 0, stream>>>(step);
  SWB_CHECK_CUDA(cudaGetLastError());
  return SWB_OK;
}

}  // namespace swb
