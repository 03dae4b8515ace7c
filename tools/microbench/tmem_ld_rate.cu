// TMEM read-rate microbenchmark: bytes per cycle per SM of tcgen05.ld (32x32b.x32) when 4 / 8 warps (one or two per
// lane quadrant) stream 176-column accumulators, i.e. what the GEMM epilogue's drain can reach at best.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I swift_b200/csrc tools/microbench/tmem_ld_rate.cu -o tools/microbench/tmem_ld_rate
#include "ptx.cuh"
#include <cstdio>
#include <vector>
using namespace swb;

__global__ void __launch_bounds__(320, 1) tmem_ld_kernel(int nwarps, int iters, int wait_every, long long* out_cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&slot), 512); tmem_relinquish<1>(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp >= 2 && warp < 2 + nwarps) {
    const int w = warp - 2;
    const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (w >> 2) * 256;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      float v[32];
#pragma unroll
      for (int c = 0; c < 5; ++c) {                 // 5 x 32 columns = 160 of a 176-column accumulator
        tmem_ld_x32(taddr + 32 * c, v);
        if (wait_every == 1 || c == 4) {
          tmem_ld_wait();
          tmem_ld_fence_regs<32>(v);
          acc += v[0] + v[31];
        }
      }
    }
    t1 = clock64();
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && warp >= 2 && warp < 2 + nwarps) {
    out_cycles[blockIdx.x * 8 + warp - 2] = t1 - t0;
    sink[blockIdx.x * 8 + warp - 2] = acc;
  }
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  float* s;
  cudaMalloc(&d, sms * 8 * sizeof(long long));
  cudaMalloc(&s, sms * 8 * sizeof(float));
  for (int wait_every : {1, 0})
    for (int nw : {1, 4, 8}) {
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        tmem_ld_kernel<<<sms, 320>>>(nw, iters, wait_every, d, s);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      std::vector<long long> h(sms * 8);
      cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int b = 0; b < sms; ++b)
        for (int w = 0; w < nw; ++w) mx = std::max(mx, (double)h[b * 8 + w]);
      const double bytes = (double)nw * 32 * 160 * 4 * iters;
      printf("tcgen05.ld.32x32b.x32, %d warp(s)/SM, wait %s: %8.1f cycles per 160-column row block, %6.1f B/cycle/SM\n", nw,
             wait_every ? "after every ld" : "once per 5 lds", mx / iters, bytes / mx);
    }
  return 0;
}
