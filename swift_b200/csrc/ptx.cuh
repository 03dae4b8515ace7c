// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM), clusters.
// Everything here is hand-written against the PTX ISA; no CUTLASS/CuTe types are used.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace swb {

#ifndef SWB_WATCHDOG_CYCLES
#define SWB_WATCHDOG_CYCLES (4000000000ll)   // ~2 s at 1.9 GHz: a legitimate wait is microseconds
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ----------------------------------------------------------------------------- cluster
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// Register re-allocation between warpgroups (4 consecutive warps execute the same instruction): the data-movement warps
// give registers back to the CTA pool, the epilogue warpgroups take them.
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init_cluster() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin on a phase parity.  A watchdog turns a protocol bug into a trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > SWB_WATCHDOG_CYCLES) {
      printf("[swift_b200] mbarrier watchdog: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier that lives in another CTA of the cluster (address from mapa_u32)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // relaxed: the barrier only hands a TMEM accumulator back to the MMA issuer (ordered by the tcgen05 fences);
  // a release here would make the epilogue wait for all of its global stores to drain first (MEMBAR + ERRBAR).
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tile load, completion on an mbarrier of the executing CTA.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// 2-D tile load issued by either CTA of an MMA pair; `cluster_bar` may address the peer CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tmap, uint32_t cluster_bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}

// L2 prefetch of a 2-D tile
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1)
               : "memory");
}
// 2-D tile store smem -> global (bulk async group of the issuing thread); the smem image uses the tensor map's swizzle
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores committed by this thread have finished READING shared memory (the buffer may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent group
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// ----------------------------------------------------------------------------- tcgen05 / TMEM
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
  if constexpr (CG == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// Shared-memory matrix descriptor (tcgen05 "smem descriptor"), canonical layouts only.
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4   bits [46,48) version = 1 (Blackwell)
//   bits [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t swizzle_code) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(swizzle_code) << 61;
  return d;
}
constexpr uint32_t SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6;

// Instruction descriptor for kind::f16 (A, B each fp16 or bf16) and fp32 accumulate.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (0 = f16, 1 = bf16)  [10,13) B fmt  [15] A major (0 = K)  [16] B major
//   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, bool a_is_f16, bool b_is_f16, int a_mn_major = 0,
                                                      int b_mn_major = 0) {
  return (1u << 4) | ((a_is_f16 ? 0u : 1u) << 7) | ((b_is_f16 ? 0u : 1u) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA (pair).
template <int CG>
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Make an mbarrier track completion of all tcgen05 ops issued so far by this thread.
// CG == 2: the arrive is multicast to the same barrier offset in both CTAs of the pair.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            bar),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// ----------------------------------------------------------------------------- warp-collective single-issuer forms
// The producer / MMA warps run their loops with all 32 lanes converged (warp-uniform control flow and operands, which
// lets ptxas keep descriptors, coordinates and barrier addresses in uniform registers); exactly one lane, chosen by
// elect.sync inside the asm block, issues the instruction.  (Issuing from inside an `if (lane == 0)` branch instead
// makes ptxas wrap every UTCHMMA / UTMALDG in a vote + R2UR "waterfall" sequence of ~25 instructions.)
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_elect(uint32_t dst, const void* tmap, uint32_t cluster_bar, int c0,
                                                       int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];\n\t}"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_f16_ss_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One full k-block of the 256x352 tile in a single asm statement: 4 k-steps x 2 sub-tiles = 8 accumulating UMMAs behind ONE
// elect.  Descriptors are passed as their low words (start address | LBO: only the 14-bit address field changes, by +2 per
// k-step of 32 bytes) plus the constant high word, so the issuing warp spends ~6 instead of ~14 instructions per UMMA: it
// shares its scheduler with two epilogue warps and every issue slot it needs is contended (tools/issue_contention.py).
__device__ __forceinline__ void umma_f16_kblock_pair_elect(uint32_t tmem_d0, uint32_t tmem_d1, uint32_t a_lo, uint32_t b_lo,
                                                           uint32_t b_sub_step, uint32_t desc_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred q, t;\n\t.reg .b32 a1, a2, a3, b1, b2, b3, c0, c1, c2, c3;\n\t"
      ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3, dc0, dc1, dc2, dc3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "add.u32 a1, %2, 2;\n\tadd.u32 a2, %2, 4;\n\tadd.u32 a3, %2, 6;\n\t"
      "add.u32 b1, %3, 2;\n\tadd.u32 b2, %3, 4;\n\tadd.u32 b3, %3, 6;\n\t"
      "add.u32 c0, %3, %4;\n\tadd.u32 c1, b1, %4;\n\tadd.u32 c2, b2, %4;\n\tadd.u32 c3, b3, %4;\n\t"
      "mov.b64 da0, {%2, %5};\n\tmov.b64 da1, {a1, %5};\n\tmov.b64 da2, {a2, %5};\n\tmov.b64 da3, {a3, %5};\n\t"
      "mov.b64 db0, {%3, %5};\n\tmov.b64 db1, {b1, %5};\n\tmov.b64 db2, {b2, %5};\n\tmov.b64 db3, {b3, %5};\n\t"
      "mov.b64 dc0, {c0, %5};\n\tmov.b64 dc1, {c1, %5};\n\tmov.b64 dc2, {c2, %5};\n\tmov.b64 dc3, {c3, %5};\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da0, db0, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%1], da0, dc0, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da1, db1, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%1], da1, dc1, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da2, db2, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%1], da2, dc2, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da3, db3, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%1], da3, dc3, %6, t;\n\t}"
      :
      : "r"(tmem_d0), "r"(tmem_d1), "r"(a_lo), "r"(b_lo), "r"(b_sub_step), "r"(desc_hi), "r"(idesc)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::
            "r"(bar),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]  (A operand read from tensor memory: lanes = rows, 16-bit pairs packed along columns)
__device__ __forceinline__ void umma_f16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA-pair form (M = 256: every CTA reads its 128 rows of A from its own tensor memory)
__device__ __forceinline__ void umma_f16_ts_pair_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// smem -> TMEM copy of 128 rows x 32 bytes (one K-step of a K-major 16-bit operand) in both CTAs of the pair; executes in
// issue order with the tcgen05.mma instructions of the same thread
__device__ __forceinline__ void utccp_128x256b_pair_elect(uint32_t taddr, uint64_t sdesc) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.cp.cta_group::2.128x256b [%0], %1;\n\t}" ::"r"(taddr),
      "l"(sdesc)
      : "memory");
}
// registers -> TMEM: each lane writes 16 consecutive 32-bit columns of its own TMEM lane
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
      "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 4-D tile load (box fixed by the tensor map), completion on an mbarrier of the executing CTA; one elected lane issues
__device__ __forceinline__ void tma_load_4d_elect(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];\n\t}"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMEM -> registers: each lane of the warp reads its own TMEM lane, N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.ld is asynchronous: its destination registers may only be read after tcgen05.wait::ld.  The wait
// has no data dependence on them, so pin every consumer behind it with an empty volatile asm per register
// (volatile asm statements keep their program order).
template <int N>
__device__ __forceinline__ void tmem_ld_fence_regs(float* v) {
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+f"(v[i]));
}

// ----------------------------------------------------------------------------- small math / packing helpers
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Tensor-core operands (every GEMM A operand, the packed weights, q/k/v, P) use one 16-bit format chosen per model:
//   F16 = true : IEEE half (11-bit significand; activations saturate to +-65504 so no inf is ever produced)
//   F16 = false: bfloat16
// tcgen05 kind::f16 encodes the A and B formats separately, but mixing them (fp16 x bf16) raises an illegal-instruction
// trap on sm_100a (measured), so the weights are packed in the same format as the activations.
template <bool F16>
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
  if constexpr (F16) {
    uint32_t r;                                   // one F2FP.SATFINITE: |x| > 65504 (and inf) -> +-65504
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  } else {
    return pack_bf16x2(lo, hi);
  }
}
template <bool F16>
__device__ __forceinline__ uint16_t pack_act1(float x) {
  if constexpr (F16) {
    uint16_t r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return r;
  } else {
    __nv_bfloat16 v = __float2bfloat16_rn(x);
    return *reinterpret_cast<uint16_t*>(&v);
  }
}
// both halves of a packed 16-bit pair -> float2 (two HADD2.F32 / two integer ops, no extraction of the halves first)
template <bool F16>
__device__ __forceinline__ float2 unpack_act2(uint32_t w) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  else return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <bool F16>
__device__ __forceinline__ float unpack_act1(uint16_t u) {
  if constexpr (F16) return __half2float(*reinterpret_cast<__half*>(&u));
  else return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
}

}  // namespace swb
