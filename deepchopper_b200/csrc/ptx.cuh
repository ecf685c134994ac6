// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / st / fences) and the UMMA shared-memory + instruction descriptors.
// Bit layouts follow the PTX ISA "tcgen05" chapter (cross-checked against CUTLASS's
// cute/arch/mma_sm100_desc.hpp, used as documentation only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "elect.sync %%rx|%%px, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, %%px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA -----------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// Multicast load: the box lands at the same shared-memory offset in every CTA of `mask`, and each of those CTAs gets the
// complete_tx on its own barrier at the same offset: one L2 read feeds the whole cluster.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 1-D bulk copy global -> shared (bytes multiple of 16, both addresses 16-byte aligned), mbarrier completion
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
// Elected-lane forms for warp-converged producers (whole warp runs the loop; see umma_bf16_x4_e): a producer inside an
// `if (lane == 0)` branch pays an R2UR / ELECT waterfall per TMA instruction (~800 cycles per stage in the Toeplitz kernel).
__device__ __forceinline__ void mbar_arrive_expect_tx_e(uint32_t bar, uint32_t bytes, uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "setp.ne.b32 pe, %2, 0;\n"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
      "}\n" ::"r"(bar),
      "r"(bytes), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_e(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                              uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "setp.ne.b32 pe, %6, 0;\n"
      "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
      "}\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d_e(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "setp.ne.b32 pe, %4, 0;\n"
      "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      "}\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar), "r"(elected)
      : "memory");
}
// TMA store shared -> global (bulk-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// TMA prefetch of a tile into L2 (no smem destination, no completion)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {  // <= N groups still reading their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, one CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask` (a stage shared by a cluster through multicast
// loads is free once every CTA's MMAs have read it)
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp gets lane (32*(warp%4)+i), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// Register reallocation between warpgroups (4 consecutive warps execute it together).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): one MMA spans the two SMs of a TPC ----------------------------------
// Each CTA holds its own 128 rows of A and HALF of the B tile's rows, so per-SM operand traffic halves.
// The leader (cluster rank 0) issues the MMAs; completion is multicast to the same barrier offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nclusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// same, default (release.cta) semantics: what CUTLASS's ClusterBarrier::arrive(cta_id) emits; avoids the
// MEMBAR.ALL.GPU that .release.cluster costs on every arrival
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquire at cluster scope
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA load into MY smem whose completion bytes are credited to a barrier given by its shared::cluster address
// (the leader's barrier): both CTAs of the pair fill one pipeline stage
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {  // one warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have retired) on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// Four consecutive K = 16 MMAs of one 64-wide K block in ONE asm block: the descriptors advance by a constant number
// of 16-byte units (2 = 32 bytes inside a 128B-swizzled K-major tile), the accumulate predicate is evaluated once.
// The issuing thread is a single lane running dependent scalar code: rebuilding two 64-bit descriptors and a
// predicate per MMA (~15 dependent instructions) made small-N MMAs issue-bound at ~200 cycles apiece.
template <int kCtaGroup>
__device__ __forceinline__ void umma_bf16_x4(uint32_t tmem_d, uint64_t a0, uint32_t a_inc, uint64_t b0, uint32_t b_inc,
                                             uint32_t idesc, uint32_t accumulate_first) {
  const uint64_t a1 = a0 + a_inc, a2 = a0 + 2 * a_inc, a3 = a0 + 3 * a_inc;
  const uint64_t b1 = b0 + b_inc, b2 = b0 + 2 * b_inc, b3 = b0 + 3 * b_inc;
  if (kCtaGroup == 1) {
    asm volatile(
        "{\n"
        ".reg .pred p0, p1;\n"
        "setp.ne.b32 p0, %10, 0;\n"
        "setp.eq.b32 p1, %10, %10;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %5, %9, p0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %6, %9, p1;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %7, %9, p1;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %4, %8, %9, p1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(accumulate_first)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p0, p1;\n"
        "setp.ne.b32 p0, %10, 0;\n"
        "setp.eq.b32 p1, %10, %10;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %5, %9, p0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %2, %6, %9, p1;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %3, %7, %9, p1;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %4, %8, %9, p1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(accumulate_first)
        : "memory");
  }
}

// ---- warp-converged issue: the whole MMA warp runs the issue loop (so ptxas keeps barrier addresses and descriptors in
// uniform registers instead of moving them there with R2UR / ELECT waterfall loops inside a one-lane branch: ~70 dependent
// instructions per 4 MMAs, ~120 cycles per MMA, which bounds MMAs smaller than N = 256) and only the tcgen05 instructions
// themselves are predicated on the elected lane.
template <int kCtaGroup>
__device__ __forceinline__ void umma_bf16_x4_e(uint32_t tmem_d, uint64_t a0, uint32_t a_inc, uint64_t b0, uint32_t b_inc,
                                               uint32_t idesc, uint32_t accumulate_first, uint32_t elected) {
  const uint64_t a1 = a0 + a_inc, a2 = a0 + 2 * a_inc, a3 = a0 + 3 * a_inc;
  const uint64_t b1 = b0 + b_inc, b2 = b0 + 2 * b_inc, b3 = b0 + 3 * b_inc;
  if (kCtaGroup == 1) {
    asm volatile(
        "{\n"
        ".reg .pred p0, p1, pe;\n"
        "setp.ne.b32 p0, %10, 0;\n"
        "setp.eq.b32 p1, %10, %10;\n"
        "setp.ne.b32 pe, %11, 0;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %5, %9, p0;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %6, %9, p1;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %7, %9, p1;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %4, %8, %9, p1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(accumulate_first), "r"(elected)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p0, p1, pe;\n"
        "setp.ne.b32 p0, %10, 0;\n"
        "setp.eq.b32 p1, %10, %10;\n"
        "setp.ne.b32 pe, %11, 0;\n"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %5, %9, p0;\n"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %2, %6, %9, p1;\n"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %3, %7, %9, p1;\n"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %4, %8, %9, p1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(accumulate_first), "r"(elected)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_e(uint32_t bar, uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "setp.ne.b32 pe, %1, 0;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar),
      "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc_e(uint32_t bar, uint16_t mask, uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "setp.ne.b32 pe, %2, 0;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(bar),
      "h"(mask), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_e(uint32_t bar, uint16_t mask, uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "setp.ne.b32 pe, %2, 0;\n"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(bar),
      "h"(mask), "r"(elected)
      : "memory");
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: half the issue slots of scalar fp32) -------
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// ---- descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle.  Fields in 16-byte units.
//   [0,14) start address  [16,30) leading-dim byte offset  [32,46) stride-dim byte offset
//   [46,48) version = 1 (Blackwell)  [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// No swizzle ("interleave"), K-major: 8-row x 16-byte core matrices (128 contiguous bytes each);
// LBO = byte distance between the two core matrices of one K=16 slice, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D (dense, no negate, no saturate).
//   [4,6) D format: 1 = F32   [7,10) A format: 1 = BF16   [10,13) B format: 1 = BF16
//   [15] A major (0 = K, 1 = MN)   [16] B major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace ptx

// Optional timeline trace (DCB200_TRACE=mlp|inproj): (tag, SM clock) records of one MMA thread / one epilogue warp / one producer,
// read back with dcb200_ctx_read_workspace("trace").  Costs one predictable branch when off.
constexpr int kTraceCap = 4096;
struct Tracer {
  long long* buf;
  int n;
  __device__ __forceinline__ void operator()(int tag) {
    if (buf && n < kTraceCap) {
      buf[2 * n] = tag;
      buf[2 * n + 1] = clock64();
      ++n;
    }
  }
};
// Compile-time switch: kernels instantiated with On = false carry no trace code at all.
template <bool On>
struct TracerT {
  long long* buf;
  int n;
  __device__ __forceinline__ void operator()(int tag) {
    if constexpr (On) {
      if (buf && n < kTraceCap) {
        buf[2 * n] = tag;
        buf[2 * n + 1] = clock64();
        ++n;
      }
    }
  }
};


}  // namespace dcb
