// Does staging the A operand in tensor memory pay?  Per k-block (64 K-elements) the 256x352 GEMM tile issues 8 UMMAs
// (2 sub-tiles x 4 k-steps, M=256 cta_group::2, N=176).  Variant SS reads A (16 KB per CTA and stage) from shared memory
// twice, once per sub-tile; variant TS copies it once into TMEM with 4 x tcgen05.cp.128x256b and both sub-tiles read it
// from there.  Prints cycles per k-block for both (704 = tensor pipe saturated).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I swift_b200/csrc tools/microbench/utccp_ts_rate.cu -o tools/microbench/utccp_ts_rate
#include "ptx.cuh"
#include <cstdio>
#include <vector>
using namespace swb;

__device__ __forceinline__ void utccp_128x256b_cg2_elect(uint32_t taddr, uint64_t sdesc) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.cp.cta_group::2.128x256b [%0], %1;\n\t}" ::"r"(taddr), "l"(sdesc)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts_cg2_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                      uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int MODE>   // 0 = SS, 1 = TS with tcgen05.cp, 2 = TS without the copies (upper bound)
__global__ void __launch_bounds__(128, 1) kblock_kernel(int iters, int aslots, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = sb + 200 * 1024;
  const uint32_t slot = bar + 64;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (sb - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    mbar_init(bar + 16, 1 << 19);
    fence_mbar_init_cluster();
  }
  if (warp == 1) { tmem_alloc<2>(slot, 512); tmem_relinquish<2>(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, 176, true, true);
    const uint64_t hi = make_smem_desc(0, 16, 1024, SWZ_128B);
    int stage = 0, as = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_wait(bar + 8, 1u, 7);
      tcgen05_fence_after();
      const uint64_t a = hi | (((sb + stage * 16384) & 0x3FFFFu) >> 4);
      const uint64_t b = hi | (((sb + 5 * 16384 + stage * 22528) & 0x3FFFFu) >> 4);
      const uint32_t ta = tmem + 352 + as * 32;
      if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) utccp_128x256b_cg2_elect(ta + 8 * k, a + 2u * k);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (MODE == 0) umma_f16_ss_elect<2>(tmem + j * 176, a + 2u * k, b + j * 704u + 2u * k, idesc, 1u);
          else umma_f16_ts_cg2_elect(tmem + j * 176, ta + 8 * k, b + j * 704u + 2u * k, idesc, 1u);
        }
      umma_commit_elect<2>(bar + 16);
      if (++stage == 5) stage = 0;
      if (++as == aslots) as = 0;
    }
    umma_commit_elect<2>(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x / 2] = t1 - t0;
  }
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<2>(tmem, 512); }
}

template <int MODE>
void run(const char* name, int aslots) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  const int smem = 204 * 1024;
  cudaFuncSetAttribute(kblock_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kblock_kernel<MODE>, iters, aslots, d);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  std::vector<long long> h(sms / 2);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto v : h) avg += v;
  avg /= h.size();
  printf("%-44s A slots %d: %7.1f cycles per k-block (8 UMMAs = 704 at full rate)\n", name, aslots, avg / iters);
  cudaFree(d);
}

int main() {
  run<0>("SS  (A and B from shared memory)", 1);
  run<2>("TS  (A from TMEM, no copies: upper bound)", 1);
  run<1>("TS + 4 x tcgen05.cp.128x256b per k-block", 1);
  run<1>("TS + 4 x tcgen05.cp.128x256b per k-block", 2);
  run<1>("TS + 4 x tcgen05.cp.128x256b per k-block", 4);
  return 0;
}
