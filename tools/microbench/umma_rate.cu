// Tensor-core issue-rate microbenchmark: cycles per tcgen05.mma (kind::f16, fp16 in, fp32 accumulate) as a function of
// N, for cta_group::1 (M=128) and cta_group::2 (M=256).  Operands are whatever is in shared memory (zeros); one CTA
// (pair) per SM issues `iters` MMAs back to back into one accumulator and waits for the final commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I swift_b200/csrc tools/microbench/umma_rate.cu -o /tmp/umma_rate
#include "ptx.cuh"
#include <cstdio>
#include <vector>
using namespace swb;

template <int CG>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int n, int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = sb + 96 * 1024;
  const uint32_t slot = bar + 16;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (sb - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init_cluster(); }
  if (warp == 1) { tmem_alloc<CG>(slot, 512); tmem_relinquish<CG>(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(128 * CG, n, true, true);
    const uint64_t hi = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint64_t a = hi | ((sb & 0x3FFFFu) >> 4), b = hi | (((sb + 32768) & 0x3FFFFu) >> 4);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss_elect<CG>(tmem, a + 2u * k, b + 2u * k, idesc, 1u);
    }
    umma_commit_elect<CG>(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x / CG] = t1 - t0;
  }
  __syncwarp();
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<CG>(tmem, 512); }
}

// Same, but with the bookkeeping of the real GEMM main loop around every 8 MMAs: wait on a (completed) mbarrier,
// tcgen05 fence, descriptors derived from a runtime stage index, two alternating accumulators, commit to an mbarrier.
template <int CG>
__global__ void __launch_bounds__(128, 1) umma_loop_kernel(int n, int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = sb + 96 * 1024;
  const uint32_t slot = bar + 64;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (sb - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);            // final
    mbar_init(bar + 8, 1);        // "full": never armed, waited with the parity that passes immediately
    mbar_init(bar + 16, 1 << 19); // "empty": absorbs the per-k-block commits
    fence_mbar_init_cluster();
  }
  if (warp == 1) { tmem_alloc<CG>(slot, 512); tmem_relinquish<CG>(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(128 * CG, n, true, true);
    const uint64_t hi = make_smem_desc(0, 16, 1024, SWZ_128B);
    int stage = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bar + 8, 1u, 7);
      tcgen05_fence_after();
      const uint64_t a = hi | (((sb + stage * 4096) & 0x3FFFFu) >> 4), b = hi | (((sb + 32768 + stage * 2048) & 0x3FFFFu) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          umma_f16_ss_elect<CG>(tmem + j * 256, a + 2u * k, b + j * 704u + 2u * k, idesc, 1u);
      umma_commit_elect<CG>(bar + 16);
      if (++stage == 5) stage = 0;
    }
    umma_commit_elect<CG>(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x / CG] = t1 - t0;
  }
  __syncwarp();
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<CG>(tmem, 512); }
}

template <int CG>
void run_loop(int n, int iters) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(umma_loop_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_loop_kernel<CG>, n, iters, d);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("loop CG=%d N=%d: %s\n", CG, n, cudaGetErrorString(e)); return; }
  }
  std::vector<long long> h(sms / CG);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto v : h) avg += v;
  avg /= h.size();
  const double per = avg / (8.0 * iters);
  printf("GEMM-like issue loop CG=%d N=%3d: %7.1f cycles/MMA (MMA itself: %d) -> issue keeps the pipe %5.1f%% busy\n", CG, n,
         per, n / 2, 100.0 * (n / 2) / per);
  cudaFree(d);
}

template <int CG>
void run(int n, int iters) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(umma_rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_rate_kernel<CG>, n, iters, d);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("CG=%d N=%d: %s\n", CG, n, cudaGetErrorString(e)); return; }
  }
  std::vector<long long> h(sms / CG);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto v : h) avg += v;
  avg /= h.size();
  const double per = avg / (4.0 * iters);
  const double flop_per_cyc_sm = 2.0 * 128 * CG * n * 16 / per / CG;
  printf("CG=%d M=%d N=%3d: %7.1f cycles/MMA  %7.0f FLOP/cycle/SM  (ideal 8192: %5.1f%%)\n", CG, 128 * CG, n, per,
         flop_per_cyc_sm, 100.0 * flop_per_cyc_sm / 8192.0);
  cudaFree(d);
}

int main() {
  for (int n : {64, 96, 128, 160, 176, 192, 208, 224, 240, 256}) run<1>(n, 2000);
  for (int n : {64, 96, 128, 160, 176, 192, 208, 224, 240, 256}) run<2>(n, 2000);
  run_loop<2>(176, 2000);
  run_loop<2>(256, 2000);
  run_loop<1>(176, 2000);
  return 0;
}
