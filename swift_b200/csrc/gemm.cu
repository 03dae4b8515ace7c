// Host launcher for the tcgen05 GEMM (gemm_sm100.cuh): builds the TMA descriptors and dispatches the
// (tile, cta_group, epilogue) instantiation.
#include "common.h"
#include "gemm_sm100.cuh"

#include <mutex>
#include <stdlib.h>

namespace swb {

// ----------------------------------------------------------------------------- error string
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int num_sms() {
  static PerDevice<int> n;
  if (n.get() == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, current_device());
    n.set(v);
  }
  return n.get();
}

// ----------------------------------------------------------------------------- TMA descriptors
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  });
  return fn;
}

static int make_tmap_16bit_2d_swz(CUtensorMap* out, const void* ptr, bool is_f16, uint64_t rows, uint64_t cols,
                                  uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz) {
  PFN_tmapEncodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return SWB_ERR_DRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((pitch_elems * 2) & 15)) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ptr=%p pitch=%llu elems)", ptr,
              (unsigned long long)pitch_elems);
    return SWB_ERR_INVALID;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box=%ux%u", (int)r, (unsigned long long)rows,
              (unsigned long long)cols, box_rows, box_cols);
    return SWB_ERR_DRIVER;
  }
  return SWB_OK;
}

int make_tmap_16bit_2d(CUtensorMap* out, const void* ptr, bool is_f16, uint64_t rows, uint64_t cols,
                       uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_16bit_2d_swz(out, ptr, is_f16, rows, cols, pitch_elems, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_128B);
}

#ifdef SWB_PROFILE_EPILOGUES
__device__ unsigned long long g_gemm_prof[24];
extern "C" __attribute__((visibility("default"))) int swb200_debug_gemm_prof(unsigned long long* out24, int reset) {
  if (out24 && cudaMemcpyFromSymbol(out24, g_gemm_prof, sizeof(unsigned long long) * 24) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[24] = {};
    if (cudaMemcpyToSymbol(g_gemm_prof, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif

// ----------------------------------------------------------------------------- launch
template <int NSUB, int CG, int EPI, bool F16>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to0, const CUtensorMap& to1,
                       const GemmParams& p, cudaStream_t stream) {
  constexpr bool kLn = EPI == EPI_LN_RES || EPI == EPI_LN_RES1;
  using S = GemmCfg<NSUB, CG, kLn>;
  auto kern = gemm_tcgen05_kernel<NSUB, CG, EPI, F16>;
  static PerDevice<bool> attr_done;
  if (!attr_done.get()) {
    SWB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_done.set(true);
  }
  const int tiles_n = (p.N + S::kTileN - 1) / S::kTileN;
  const int tiles_m = (p.M + kBlockM * CG - 1) / (kBlockM * CG);
  const int num_tiles = tiles_m * tiles_n * (p.splits > 1 ? p.splits : 1) * (p.batch > 1 ? p.batch : 1);
  int clusters = num_sms() / CG;
  if (clusters > num_tiles) clusters = num_tiles;
#ifdef SWB_PROFILE_EPILOGUES
  static const int max_cl = getenv("SWB_GEMM_MAX_CLUSTERS") ? atoi(getenv("SWB_GEMM_MAX_CLUSTERS")) : 0;   // part of the chip only
  if (max_cl > 0 && clusters > max_cl) clusters = max_cl;
#endif
  if constexpr (kLn) {
    // the groups that exchange LayerNorm statistics (the tiles_n column tiles of one row block) must run at the same
    // time: keep the cluster count a multiple of tiles_n so that they sit on neighbouring clusters in every wave
    SWB_REQUIRE(clusters >= tiles_n, "gemm_ln_residual: %d column tiles need at least as many CTA clusters (%d)", tiles_n,
                clusters);
    clusters = clusters / tiles_n * tiles_n;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(S::kThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if constexpr (kLn) {
    // The statistics exchange spins on other CTAs of this grid: every cluster has to be resident AT THE SAME TIME.  An
    // occupancy query only speaks for an idle device, so the kernel is launched COOPERATIVELY: the driver then starts
    // the grid only when all of it fits next to whatever else is running (other streams, other processes under MPS)
    // and fails the launch -- instead of deadlocking -- when it never can.  Either failure is reported as
    // SWB_ERR_RESIDENCY before anything ran, and swb200_forward falls back to GEMM + LayerNorm kernel.
    static PerDevice<int> max_clusters;                    // 0 = not queried yet
    if (max_clusters.get() == 0) {
      int mc = 0;
      SWB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&mc, kern, &cfg));
      max_clusters.set(mc > 0 ? mc : -1);
    }
    if (max_clusters.get() < clusters) {
      set_error("gemm_ln_residual: only %d of %d CTA clusters can be co-resident", max_clusters.get(), clusters);
      return SWB_ERR_RESIDENCY;
    }
    // Nsight Compute cannot replay a cooperative cluster launch (it aborts the process with "LaunchFailed"), and it
    // serialises kernels, so under its injection (the variables ncu exports to the target) the occupancy check above does
    // speak for the device: launch normally there.  SWB_LN_NO_COOPERATIVE: A/B knob (tools only).
    static const bool cooperative = getenv("SWB_LN_NO_COOPERATIVE") == nullptr &&
                                    getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") == nullptr &&
                                    getenv("NV_NSIGHT_INJECTION_PORT_BASE") == nullptr;
    if (cooperative) {
      attr[1].id = cudaLaunchAttributeCooperative;
      attr[1].val.cooperative = 1;
      cfg.numAttrs = 2;
    }
    const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, ta, tb, to0, to1, p);
    if (err == cudaErrorCooperativeLaunchTooLarge || err == cudaErrorLaunchOutOfResources) {
      cudaGetLastError();                                   // clear the sticky-less launch error
      set_error("gemm_ln_residual: cooperative launch of %d clusters refused (%s)", clusters, cudaGetErrorString(err));
      return SWB_ERR_RESIDENCY;
    }
    SWB_CHECK_CUDA(err);
    return SWB_OK;
  }
  SWB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to0, to1, p));
  return SWB_OK;
}

template <int NSUB, int CG, bool F16>
static int launch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to0,
                      const CUtensorMap& to1, const GemmParams& p, cudaStream_t stream) {
  switch (epi) {
    case EPI_STORE_F32: return launch_inst<NSUB, CG, EPI_STORE_F32, F16>(ta, tb, to0, to1, p, stream);
    case EPI_STORE_ACT: return launch_inst<NSUB, CG, EPI_STORE_ACT, F16>(ta, tb, to0, to1, p, stream);
    case EPI_EMBED: return launch_inst<NSUB, CG, EPI_EMBED, F16>(ta, tb, to0, to1, p, stream);
    case EPI_QKV: return launch_inst<NSUB, CG, EPI_QKV, F16>(ta, tb, to0, to1, p, stream);
    case EPI_SWIGLU: return launch_inst<NSUB, CG, EPI_SWIGLU, F16>(ta, tb, to0, to1, p, stream);
    case EPI_HEAD: return launch_inst<NSUB, CG, EPI_HEAD, F16>(ta, tb, to0, to1, p, stream);
    case EPI_LN_RES: return launch_inst<NSUB, CG, EPI_LN_RES, F16>(ta, tb, to0, to1, p, stream);
    case EPI_LN_RES1:
      if constexpr (F16) return launch_inst<NSUB, CG, EPI_LN_RES1, true>(ta, tb, to0, to1, p, stream);
      break;
#ifdef SWB_PROFILE_EPILOGUES      // measurement-only epilogues (tools/gemm_phases.py): build with SWB_NVCC_DEFINES=-DSWB_PROFILE_EPILOGUES
    case EPI_BUSY: return launch_inst<NSUB, CG, EPI_BUSY, F16>(ta, tb, to0, to1, p, stream);
    case EPI_DISCARD: return launch_inst<NSUB, CG, EPI_DISCARD, F16>(ta, tb, to0, to1, p, stream);
    case EPI_DRAIN: return launch_inst<NSUB, CG, EPI_DRAIN, F16>(ta, tb, to0, to1, p, stream);
    case EPI_SMEM_ONLY: return launch_inst<NSUB, CG, EPI_SMEM_ONLY, F16>(ta, tb, to0, to1, p, stream);
    case EPI_DIRECT: return launch_inst<NSUB, CG, EPI_DIRECT, F16>(ta, tb, to0, to1, p, stream);
#endif
  }
  set_error("unknown GEMM epilogue %d (the profiling epilogues 6..10 need a -DSWB_PROFILE_EPILOGUES build)", epi);
  return SWB_ERR_INVALID;
}

// A: [M, K] with row pitch lda; W: [N, K] with row pitch ldw (nn.Linear layout); both fp16 (act_f16) or both bf16.
// tile: 1 = single CTA 128x176, 2 = CTA pair 256x176 (double-buffered accumulator), 3 = CTA pair 256x352.
int launch_gemm(int epi, int tile, int act_f16, const void* A, int lda, const void* W, int ldw, const GemmParams& p_in,
                cudaStream_t stream) {
  GemmParams p = p_in;
  SWB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  SWB_REQUIRE(p.K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K and row pitches must be multiples of 8 (K=%d)",
              p.K);
  SWB_REQUIRE(p.splits <= 1 || (epi == EPI_STORE_F32 && p.K % kBlockK == 0),
              "gemm: split-K needs the fp32 store epilogue and K %% 64 == 0 (epi=%d K=%d splits=%d)", epi, p.K, p.splits);
  SWB_REQUIRE(p.batch <= 1 || epi == EPI_STORE_F32, "gemm: batched problems need the fp32 store epilogue (epi=%d)", epi);
  SWB_REQUIRE(tile >= 1 && tile <= 3, "gemm: tile config must be 1 (128x176), 2 (256x176) or 3 (256x352), got %d", tile);
  // sub-tile 0 of the 256x352 tile runs three k-blocks ahead of sub-tile 1 (the kernel clamps to its pipeline depth - 2: two
  // with the fused LayerNorm's four stages; gemm_sm100.cuh); SWB_GEMM_SKEW: A/B knob (tools only)
  static const int skew_env = getenv("SWB_GEMM_SKEW") ? atoi(getenv("SWB_GEMM_SKEW")) : 3;
  p.skew = skew_env;
  // the fused-LayerNorm GEMMs: one k-block (their two epilogue groups meet again in the statistics exchange, so a group that
  // starts early only waits longer there: 384.9 vs 383.2 member-steps/s with 2); SWB_GEMM_SKEW_LN: A/B knob (tools only)
  static const int skew_ln_env = getenv("SWB_GEMM_SKEW_LN") ? atoi(getenv("SWB_GEMM_SKEW_LN")) : 1;
  if (epi == EPI_LN_RES || epi == EPI_LN_RES1) p.skew = skew_ln_env;
#ifdef SWB_PROFILE_EPILOGUES
  SWB_CHECK_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&p.prof), g_gemm_prof));
  static const bool nomma = getenv("SWB_GEMM_NOMMA") != nullptr;       // TMA feed rate alone (see gemm_sm100.cuh)
  if (nomma) p.ln_debug |= 4;
  static const bool noload = getenv("SWB_GEMM_NOLOAD") != nullptr;     // no TMA loads either: the epilogue alone
  if (noload) p.ln_debug |= 8;
#endif
  const int cg = tile == 1 ? 1 : 2;
  const int nsub = tile == 3 ? 2 : 1;
  const bool f16 = (act_f16 & 1) != 0;        // bit 1 of the format word: single-value residual stream (EPI_EMBED / EPI_LN_RES)
  p.x_single = (act_f16 >> 1) & 1;
  SWB_REQUIRE(!p.x_single || f16, "gemm: the single-value residual stream needs fp16 operands");
  if (epi == EPI_LN_RES && p.x_single) epi = EPI_LN_RES1;
  CUtensorMap ta, tb, to0, to1;
  const uint64_t k_total = static_cast<uint64_t>(p.K) * (p.splits > 1 ? p.splits : 1);
  const uint64_t nb = p.batch > 1 ? p.batch : 1;
  int rc = make_tmap_16bit_2d(&ta, A, f16, p.M * nb, k_total, lda, kBlockM, kBlockK);
  if (rc) return rc;
  rc = make_tmap_16bit_2d(&tb, W, f16, p.N * nb, k_total, ldw, kUmmaN * nsub / cg, kBlockK);
  if (rc) return rc;
  // 16-bit outputs leave through TMA tile stores: a 64-column box (SWIZZLE_128B image) plus the 24-column rest of an
  // 88-column slot (plain image), or the 32-column rest of a padded 96-column q/k/v row (SWIZZLE_64B image)
  static const bool no_tma_store = getenv("SWB_NO_TMA_STORE") != nullptr;      // profiling knob: LDS + STG path
  p.tma_store = 0;
  to0 = ta;
  to1 = ta;
  if (!no_tma_store && (epi == EPI_STORE_ACT || epi == EPI_SWIGLU)) {
    const uint64_t cols = epi == EPI_SWIGLU ? static_cast<uint64_t>(p.N / 2) : static_cast<uint64_t>(p.N);
    rc = make_tmap_16bit_2d_swz(&to0, p.out0, f16, p.M, cols, p.ldo, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_16bit_2d_swz(&to1, p.out0, f16, p.M, cols, p.ldo, 32, 24, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    p.tma_store = 1;
  } else if (!no_tma_store && epi == EPI_QKV && p.M % 32 == 0) {
    const uint64_t rows = static_cast<uint64_t>(3) * p.heads * p.M;
    rc = make_tmap_16bit_2d_swz(&to0, p.out0, f16, rows, kHeadDimPad, kHeadDimPad, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_16bit_2d_swz(&to1, p.out0, f16, rows, kHeadDimPad, kHeadDimPad, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    p.tma_store = 1;
  }
  if (epi == EPI_LN_RES1 && tile == 3) {
    // the residual stream (hi half of xhl: [M, N] at row pitch 2N) travels by TMA: 64-column SWIZZLE_128B boxes and the
    // 24-column rest of an 88-column slot, 32 rows each, for loads and stores alike
    rc = make_tmap_16bit_2d_swz(&to0, p.xhl, true, p.M, p.N, 2 * static_cast<uint64_t>(p.N), 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_16bit_2d_swz(&to1, p.xhl, true, p.M, p.N, 2 * static_cast<uint64_t>(p.N), 32, 24, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    p.tma_store = 1;
  }
  if (tile == 3)
    return f16 ? launch_epi<2, 2, true>(epi, ta, tb, to0, to1, p, stream) : launch_epi<2, 2, false>(epi, ta, tb, to0, to1, p, stream);
  if (tile == 2)
    return f16 ? launch_epi<1, 2, true>(epi, ta, tb, to0, to1, p, stream) : launch_epi<1, 2, false>(epi, ta, tb, to0, to1, p, stream);
  return f16 ? launch_epi<1, 1, true>(epi, ta, tb, to0, to1, p, stream) : launch_epi<1, 1, false>(epi, ta, tb, to0, to1, p, stream);
}

}  // namespace swb
