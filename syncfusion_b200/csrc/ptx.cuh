// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
// Hand-written; descriptor bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace sfb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.b32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait with a post-mortem record.  A waiter polls the hinted try_wait below; when its poll count passes
// 2^SFB_WAIT_BOUND_LOG2 (default 2^21: >= 30 ms of continuously woken polling - 30x any healthy wait of these kernels
// - or ~8 s of parked polling; a parked poll lasts ~4 us on B200, tools/trywait_bench.cu) it leaves a record in a
// host-mapped log and traps, so a pipeline bug ends in seconds with a CUDA error that names the kernel, the plan op and
// the wait site instead of hanging the GPU (host side: sfb.cu wait_log_text -> sfb_last_error()).
// Two record levels (SFB_WAIT_LOG, build time), because the cold code at ~25 wait sites per kernel is not free:
//   1 (default)  the waiter stores {call site, CTA, thread} to fixed words (last writer wins) and traps: ~8
//                instructions per site, nothing live across them.  The plan op comes from the progress marker every
//                kernel writes after griddepcontrol.wait (mark_progress: the last op that started is the stuck one).
//   2            full per-waiter records {site, barrier address + parity, CTA, thread, op, raw barrier word} through ONE
//                out-of-line noreturn function, then ~20 ms of lingering so the other stuck roles record too.
//                (+10 % SASS per kernel, measured -4 % end to end; inlining the record cost 11 % on sk_kernel, a
//                returning call 11 % on the attention kernel.)  python -m syncfusion_b200.build --wait-log 2
//   0            trap only.
#ifndef SFB_WAIT_BOUND_LOG2
#define SFB_WAIT_BOUND_LOG2 21
#endif
#ifndef SFB_WAIT_LOG
#define SFB_WAIT_LOG 1
#endif
struct WaitRecord { uint32_t site, bar, cta_x, cta_yz, thread, tag, state_lo, state_hi; };
constexpr int kWaitRecMax = 120;
struct WaitLog {
  uint32_t count;                 // level 2: number of records
  uint32_t cur_tag;               // progress marker: plan op of the most recently started kernel (0xFFFFFFFF: none)
  uint32_t light_site, light_cta, light_thread, light_flag;      // level 1: last stuck waiter (flag = 1 once written)
  uint32_t pad[2];
  WaitRecord rec[kWaitRecMax];
};
__device__ WaitLog* g_wait_log = nullptr;            // device pointer of the host-mapped log (sfb_create)

// Every kernel of the chain calls this right after griddepcontrol.wait (the previous grid has completed): one
// fire-and-forget store per launch.
__device__ __forceinline__ void mark_progress(int tag) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    WaitLog* lg = g_wait_log;
    if (lg != nullptr) *reinterpret_cast<volatile uint32_t*>(&lg->cur_tag) = (uint32_t)tag;
  }
}

__device__ __noinline__ __attribute__((noreturn)) void wait_timeout_trap(uint32_t bar, uint32_t parity, uint32_t site, uint32_t tag) {
  WaitLog* lg = g_wait_log;
  if (lg != nullptr && lane_id() == 0) {
    const uint32_t slot = atomicAdd(&lg->count, 1u);
    if (slot < (uint32_t)kWaitRecMax) {
      uint32_t lo, hi;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(bar));
      volatile uint32_t* r = reinterpret_cast<volatile uint32_t*>(&lg->rec[slot]);
      r[0] = site; r[1] = bar | (parity << 31); r[2] = blockIdx.x; r[3] = blockIdx.y | (blockIdx.z << 16);
      r[4] = threadIdx.x | (blockDim.x << 16); r[5] = tag; r[6] = lo; r[7] = hi;
    }
    __threadfence_system();
  }
  for (int i = 0; i < 200; ++i) __nanosleep(100000);     // ~20 ms: let the other stuck waiters record before the grid dies
  __threadfence_system();
  __trap();
  while (true) {}
}
// Level 1: out of line, noreturn, ONE immediate argument - a wait site costs {mov, call} and pins no register.  (Written
// inline with blockIdx / threadIdx the compiler hoisted the packed words to the kernel entry and held two registers for
// the whole 128-register sk kernel: -3 % end to end.)
__device__ __noinline__ __attribute__((noreturn)) void wait_trap_light(uint32_t site) {
  WaitLog* lg = g_wait_log;
  if (lg != nullptr) {
    volatile uint32_t* w = reinterpret_cast<volatile uint32_t*>(&lg->light_site);
    w[0] = site; w[1] = blockIdx.x | (blockIdx.y << 16) | (blockIdx.z << 24); w[2] = threadIdx.x | (blockDim.x << 16); w[3] = 1u;
    __threadfence_system();
  }
  __trap();
  while (true) {}
}
// try_wait with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes or the
// hint elapses, so a waiting warp issues almost nothing (it does not compete with the working warps of its
// sub-partition).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n"
      "selp.b32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_at(uint64_t* bar, uint32_t parity, uint32_t site, uint32_t tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity)) {
    if (__builtin_expect(++spins > (1u << SFB_WAIT_BOUND_LOG2), 0)) {      // cold: keep it out of the loop's cache lines
#if SFB_WAIT_LOG >= 2
      wait_timeout_trap(smem_u32(bar), parity, site, tag);
#elif SFB_WAIT_LOG == 1
      wait_trap_light(site);
#else
      __trap();
#endif
    }
  }
  (void)site; (void)tag;
}
// Call-site form: `p` is the kernel's __grid_constant__ parameter struct (its `tag` = plan op index lives in the
// constant bank, so naming the op costs no register), SFB_FILE_ID is set by each kernel header.
#define mbar_wait(bar, parity) ::sfb::mbar_wait_at((bar), (parity), (uint32_t)((SFB_FILE_ID << 16) | __LINE__), (uint32_t)p.tag)

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Every kernel of the per-step chain calls pdl_trigger() first (the NEXT kernel's CTAs may be scheduled as soon as
// SMs free up and run their prologue) and pdl_wait() before touching global memory (blocks until the PREVIOUS kernel
// has completed and flushed).  Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* g) { asm volatile("prefetch.global.L1 [%0];" ::"l"(g)); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("SFB_NO_PDL") ? 0 : 1;
  return on != 0;
}
// Set while the sampling loop is being CAPTURED into a CUDA graph: inside a graph the kernel-to-kernel hand-off is already
// cheap and the early-resident CTAs of the next node only take SMs away from the current one (measured: 44.4 -> 44.9
// clips/s without the attribute under the graph, 42.6 -> 41.4 without it on plain stream launches).
inline thread_local bool g_pdl_suppress = false;
// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute.
// ONLY for kernels that call pdl_wait() before their first dependent global access.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && !g_pdl_suppress) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], one CTA, bf16/fp16 operands (kind::f16) or tf32 (kind::tf32), fp32 accumulate.
template <bool kTF32>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows = lanes, K along the columns, 32 bits per column: two bf16 or
// one tf32 element) is read from tensor memory.
template <bool kTF32>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2, cluster of two)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// TMA loads whose completion is signalled on an mbarrier of the pair's LEADER CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]: M = 256 over the pair, issued by the leader.
__device__ __forceinline__ void umma_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the mbarrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have completed.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
// one fp32 column of this warp's 32 lanes (cross-warp exchange of per-row scalars through spare TMEM columns)
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit): [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
// [61,64) layout (2 = SWIZZLE_128B).
// layout: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks; the only MN-major tf32 layout)
__host__ __device__ constexpr uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t((smem_addr >> 4) & 0x3FFF)) | (uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32) | (uint64_t(1) << 46) | (uint64_t(layout) << 61);
}
__host__ __device__ constexpr uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_smem_desc(smem_addr, lbo_bytes, sbo_bytes, 2);
}
// Instruction descriptor (32 bit): [4,6) D fmt (1 = f32), [7,10) A fmt, [10,13) B fmt (0 f16, 1 bf16, 2 tf32),
// [15] A major, [16] B major (0 = K-major, 1 = MN-major), [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t m, uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <typename T> struct ElemTraits;
template <> struct ElemTraits<__nv_bfloat16> {
  static constexpr bool kTF32 = false;
  static constexpr uint32_t kFmt = 1;          // BF16
  static constexpr int kAtomElems = 64;        // elements per 128-byte swizzle row
  static constexpr int kUmmaK = 16;
};
template <> struct ElemTraits<float> {
  static constexpr bool kTF32 = true;
  static constexpr uint32_t kFmt = 2;          // TF32
  static constexpr int kAtomElems = 32;
  static constexpr int kUmmaK = 8;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
// fp32-mode operands feed kind::tf32 MMAs, which truncate the low 13 mantissa bits: round to nearest here so the
// operand error is unbiased (half an ulp of tf32) instead of a systematic truncation.
template <> __device__ __forceinline__ float from_f32<float>(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace sfb
