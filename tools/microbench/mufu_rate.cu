// MUFU throughput on sm_100a: ex2.approx.ftz.f32 vs ex2.approx.ftz.f16x2 (two results per instruction) vs ex2.approx.f16,
// and tanh.approx.f32 / .f16x2.  8 warps per SM sub-partition... 1024 threads per CTA, one CTA per SM, 8 independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench/mufu_rate.cu -o tools/microbench/mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, int iters, long long* cyc) {
  uint32_t a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = 0x38003800u + threadIdx.x * 8 + c;    // ~0.5 in both halves / a small float
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[c]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[c]));
      if (MODE == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(a[c]));
      if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[c]));
      if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+r"(a[c]));
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s ^= a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int results_per_instr) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, 1024>>>(out, iters, cyc);
  k<MODE><<<148, 1024>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double instr = 1024.0 * iters * 8;
  printf("%-22s %8.2f thread-instr/clk/SM  -> %6.2f results/clk/SM  (%s)\n", name, instr / c, results_per_instr * instr / c,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 2);
  run<2>("tanh.approx.f32", 1);
  run<3>("tanh.approx.f16x2", 2);
  run<4>("rcp.approx.ftz.f32", 1);
  return 0;
}
