// utils.cuh — device-side helpers for the B200 (sm_100a) attention forward path.
//
// Takes the place of the reference's kernels/utils.cuh (reference: /root/reference/kernels/utils.cuh:12-113).
// There the helpers were cooperative-groups dot products, shuffle reductions and the FA-1 style
// per-tile-normalised softmax state.  Here the helpers are the raw sm_100a building blocks the
// stages in loaders.cuh / computers.cuh are made of:
//   * mbarrier (init / expect_tx / arrive / parity wait with a hang guard)
//   * TMA bulk-tensor loads (cp.async.bulk.tensor.4d) with L2 eviction hints
//   * tcgen05: TMEM alloc/dealloc, mma (SS and TS forms), commit, ld/st, fences
//   * UMMA shared-memory / instruction descriptor builders
//   * packed fp32x2 math, exp2, bf16/fp16 packing
// WARP and FLOAT_SIZE keep the reference's names (utils.cuh:12-13).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>

#define WARP 32
#define FLOAT_SIZE 4

#ifndef FA_HANG_GUARD_CYCLES
// Any mbarrier wait that spins longer than this many SM cycles traps instead of hanging the GPU.
#define FA_HANG_GUARD_CYCLES (4000000000LL)
#endif

// Phase profiling (debug builds with -DFA_PHASE_PROFILE): cycle counters per pipeline phase, summed with atomics.
#ifdef FA_PHASE_PROFILE
#define FA_PROF_DECL(n) long long prof_acc_[n] = {}; long long prof_t_ = clock64()
#define FA_PROF_MARK(i) do { const long long t_ = clock64(); prof_acc_[i] += t_ - prof_t_; prof_t_ = t_; } while (0)
#define FA_PROF_FLUSH(ptr, base, n) do { if (ptr) for (int i_ = 0; i_ < (n); ++i_) atomicAdd((ptr) + (base) + i_, (unsigned long long)prof_acc_[i_]); } while (0)
#else
#define FA_PROF_DECL(n)
#define FA_PROF_MARK(i)
#define FA_PROF_FLUSH(ptr, base, n)
#endif

// FA_TRACE builds: CTA 0 timestamps the hand-offs of a few 128-key steps of its second work item into the debug profile
// buffer (scripts/trace_steps.py prints the timeline): slot = 64 + ((tile * 2 + role) * 16 + (j - 8)) * 8 + event, role 0 =
// softmax warpgroup (events: got S, released S, exponentials start, first / second P half delivered, ready for S),
// role 1 = MMA issue (Q K^T issue begins / ends, first P V half begins / ends, second half begins / ends)
#ifdef FA_TRACE
#define FA_TRACE_EV(prof, k, tile, role, j, ev)                                                                  \
    do {                                                                                                            \
        if ((prof) && blockIdx.x == 0 && (k) == 1 && (j) >= 8 && (j) < 24 && (threadIdx.x & 31) == 0)               \
            (prof)[64 + (((tile) * 2 + (role)) * 16 + ((j) - 8)) * 8 + (ev)] = (unsigned long long)clock64();       \
    } while (0)
#else
#define FA_TRACE_EV(prof, k, tile, role, j, ev)
#endif

// FA_TRACE2 builds: an event log of one CTA (FA_T2_CTA, default 0) in the debug profile buffer, for launch-bound shapes
// (scripts/trace_cta.py).  Every traced role keeps its own event counter in a register and writes (code << 48) | clock into
// its own region of 128 slots (slot 64 + 128 * role + i): no atomics, a few cycles per event.  Roles: 0 thread 0 (1 kernel
// entry, 2 set-up done, 50 CTA done), 1 producer (3 item published, 4 Q loads issued, 5 K/V tile issued), 2 / 3 MMA issuer of
// slot 0 / 1 (6 / 7 Q K^T issued), 4 / 5 softmax warpgroup 0 / 1 (10 / 11 S tile taken, 20 / 21 P delivered, 30 / 31 epilogue
// begins, 40 / 41 epilogue done).
#ifdef FA_TRACE2
#ifndef FA_T2_CTA
#define FA_T2_CTA 0
#endif
#define FA_T2_DECL int t2_cnt_ = 0
#define FA_T2(prof, role, code)                                                                                     \
    do {                                                                                                            \
        if ((prof) && blockIdx.x == FA_T2_CTA && (threadIdx.x & 31) == 0 && t2_cnt_ < 127)                            \
            (prof)[64 + 128 * (role) + t2_cnt_++] = ((unsigned long long)(code) << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); \
    } while (0)
#else
#define FA_T2_DECL
#define FA_T2(prof, role, code)
#endif

namespace fa {

enum DType : int { kF32 = 0, kF16 = 1, kBF16 = 2 };

// ------------------------------------------------------------------------------------------
// addresses
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp (the lowest active one) gets true.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
#ifndef FA_TRYWAIT_HINT_NS
#define FA_TRYWAIT_HINT_NS 0
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if FA_TRYWAIT_HINT_NS > 0
    // with a suspend-time hint: the thread may sleep in hardware up to that long before try_wait returns false
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(FA_TRYWAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
    return ok;
}
// Parity wait.  The slow path carries a cycle-count guard so a protocol bug traps (a CUDA error the
// host sees) instead of hanging the device.
__device__ __forceinline__ void mbar_hang_report(uint32_t bar, uint32_t parity) {
#ifdef FA_DEBUG_HANG   // names the barrier; off by default because the printf call costs registers in every role
    printf("fa_b200: mbarrier wait timed out (block %d,%d,%d thread %d bar 0x%x parity %u)\n",
           blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
#endif
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > FA_HANG_GUARD_CYCLES) mbar_hang_report(bar, parity);
    }
}

// ------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load: coordinates are (innermost .. outermost) = (c0 = column in d, c1 = row in N, c2 = head, c3 = batch)
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint32_t dst_smem, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_hint(const CUtensorMap* m, uint32_t dst_smem, uint32_t bar,
                                                 int c0, int c1, int c2, int c3, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
          "l"(policy)
        : "memory");
}
// L2 prefetch of one box of a tiled tensor (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 4-D tiled STORE shared -> global (bulk async-group completion); rows / columns outside the tensor are clipped by the hardware
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every bulk group of this thread has finished READING its shared-memory source (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// every bulk group of this thread has completed (its global writes are performed)
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA) after this fence + a barrier
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// L2 eviction policies (same encodings createpolicy would return; used as cache hints)
static constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
static constexpr uint64_t kEvictLast  = 0x14F0000000000000ull;

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM management, fences, commit
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {   // whole warp
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Arrive (count 1) on an mbarrier once every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Register re-allocation between warpgroups (all 128 threads of a warpgroup must execute it).
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ------------------------------------------------------------------------------------------
// tcgen05.mma  (kind::f16: bf16 or fp16 operands, fp32 accumulate in TMEM)
// ------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Instruction descriptor for kind::f16 (bit layout: c_format[4,6) a_format[7,10) b_format[10,13)
// a_major[15] b_major[16] N>>3 [17,23) M>>4 [24,29)).  ab_fmt: 0 = fp16, 1 = bf16.  *_mn: 1 = MN-major.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int ab_fmt, int a_mn, int b_mn) {
    return (1u << 4) | (uint32_t(ab_fmt) << 7) | (uint32_t(ab_fmt) << 10) | (uint32_t(a_mn) << 15) |
           (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// Shared-memory matrix descriptor, 128-byte swizzle, sm_100 version bits.
//   start address  bits [0,14)   (>>4)
//   leading  byte offset bits [16,30) (>>4)
//   stride   byte offset bits [32,46) (>>4)
//   version        bits [46,48) = 1
//   layout type    bits [61,64) = 2 (SWIZZLE_128B)
// K-major operand (rows of 128 B = 64 halfs of K): SBO = 1024 (8 rows), LBO unused.
// MN-major operand (rows of 128 B = 64 halfs of MN, one row per K index): SBO = 1024 (8 K rows),
// LBO = distance between consecutive 64-element MN blocks.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3FFFu);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}

// ------------------------------------------------------------------------------------------
// CTA pair (cluster of 2, tcgen05 cta_group::2): the two CTAs of a pair run on the two SMs of one TPC.  One MMA instruction,
// issued by the even ("leader") CTA, multiplies a 256-row A (128 rows from each CTA's own shared memory / TMEM) with a B whose N
// dimension is split over the two CTAs' shared memories, and leaves each CTA its own 128 rows of D in its own TMEM — every CTA
// loads and reads only HALF of each K / V tile.  Shared-window addresses of a CTA are valid shared::cluster addresses of its
// own memory; mapa gives the address of the same offset in another CTA of the cluster.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
// arrive / expect_tx on an mbarrier of ANY CTA of the cluster (address from mapa_shared).  Default semantics (release at CTA
// scope) for the TMEM hand-offs, whose ordering comes from the tcgen05 fences: a cluster-scope release costs the arriving
// warp several hundred cycles each time (measured: three of them per step made the softmax 1.5x slower).  The cluster-scope
// form is for the one hand-off that publishes ordinary shared-memory DATA to the other CTA (the work-item mailbox).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t caddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(caddr), "r"(bytes) : "memory");
}
// parity wait on a LOCAL mbarrier whose arrivals come from the other CTA (acquire at cluster scope); same hang guard
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > FA_HANG_GUARD_CYCLES) mbar_hang_report(bar, parity);
    }
}
__device__ __forceinline__ void st_shared_cluster_v4(uint32_t caddr, int a, int b, int c, int d) {
    asm volatile("st.shared::cluster.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// TMA load issued by either CTA of a pair: data into the issuing CTA's shared memory, completion bytes onto the mbarrier at
// `bar_caddr` (a shared::cluster address: the LEADER's barrier, whichever CTA issues)
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* m, uint32_t dst_smem, uint32_t bar_caddr,
                                                 int c0, int c1, int c2, int c3, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_caddr), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
          "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {   // one whole warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive (count 1) on the mbarrier at the same offset in every CTA of `mask` once all MMAs issued so far by this thread retire
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask = 3) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05.ld / tcgen05.st, shape 32x32b: thread i of the warp <-> TMEM lane (lane_base + i),
// N consecutive 32-bit columns per thread.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05.ld / tcgen05.st, 16-lane shapes: the warp addresses 16 TMEM lanes (the lane field of taddr is the first one and must
// lie inside the warp's own 32-lane quarter, so two warps of the same quarter can each own 16 rows of a tile).
//   16x256b.xN : thread t holds, per repetition g (8 columns): regs 4g+{0,1} = lane t/4,     columns 8g + 2(t%4) + {0,1}
//                                                              regs 4g+{2,3} = lane t/4 + 8, same columns
//                (the accumulator fragment of mma.sync: a row is spread over the 4 threads of a quad)
//   16x128b.xN : per repetition g (4 columns): reg 2g = lane t/4, column 4g + t%4; reg 2g+1 = lane t/4 + 8, same column
//                (what the packed 16-bit pairs of a 16x256b fragment become: column 4g + t%4 = keys 8g + 2(t%4) + {0,1})
// ------------------------------------------------------------------------------------------
#define FA_R8(r, o) "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define FA_W8(r, o) "r"(r[o]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {   // 64 columns -> 32 registers
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : FA_R8(r, 0), FA_R8(r, 8), FA_R8(r, 16), FA_R8(r, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {   // 32 columns -> 16 registers
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : FA_R8(r, 0), FA_R8(r, 8)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), FA_W8(r, 0), FA_W8(r, 8)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t* r) {   // 16 columns -> 8 registers
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : FA_R8(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), FA_W8(r, 0) : "memory");
}
// 32x32b.x16: thread i <-> lane (lane_base + i), 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : FA_R8(r, 0), FA_R8(r, 8)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t* r) {   // 32 columns <- 16 registers
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), FA_W8(r, 0), FA_W8(r, 8)
        : "memory");
}
// named barrier over `threads` threads (a multiple of 32) of the CTA; id 0 is __syncthreads
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (a.x*b.x + c.x, a.y*b.y + c.y) in one instruction (sm_100 packed fp32)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    uint64_t ua, ub, uc, ud;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    uint64_t ua, ub, ud;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
    return d;
}
// v if idx <= lim else -inf; written as one predicated select so the compiler keeps the score row in registers
__device__ __forceinline__ uint32_t mask_gt(uint32_t v, int idx, int lim) {
    uint32_t d;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, %3;\n\tselp.b32 %0, 0xff800000, %1, p;\n\t}"
        : "=r"(d) : "r"(v), "r"(idx), "r"(lim));
    return d;
}
// exp2 on the FMA/ALU pipes for a pair of values (Cody-Waite split + degree-3 polynomial), used for a fraction of the
// softmax exponentials so that MUFU.EX2 (16/clk/SM) is not the only pipe doing them:
//   x = n + f, n = floor(x) via the 1.5*2^23 magic add in round-down mode, f in [0,1)
//   2^f ~= 1 + f*(c1 + f*(c2 + f*c3))   (max relative error 8.6e-5, far below the 2^-9 rounding of a bf16 P)
//   2^x = 2^f with n added to the exponent field (integer add of the magic sum's low bits shifted by 23)
// Inputs are clamped at -127 (=> 2^-127, flushed to 0), which also makes -inf (masked scores) safe.
__device__ __forceinline__ float2 ex2_emu2(float2 x) {
    // NaN-propagating clamp (max.NaN): a NaN score must come out as NaN here exactly as it does from MUFU.EX2, not as 2^-127
    asm("max.NaN.f32 %0, %0, %1;" : "+f"(x.x) : "f"(-127.f));
    asm("max.NaN.f32 %0, %0, %1;" : "+f"(x.y) : "f"(-127.f));
    uint64_t ux, ut, un, uf, up;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ux) : "f"(x.x), "f"(x.y));
    asm("{\n\t.reg .b64 m;\n\tmov.b64 m, {%2, %2};\n\tadd.rm.ftz.f32x2 %0, %1, m;\n\t}" : "=l"(ut) : "l"(ux), "f"(12582912.f));
    asm("{\n\t.reg .b64 m;\n\tmov.b64 m, {%2, %2};\n\tadd.rn.ftz.f32x2 %0, %1, m;\n\t}" : "=l"(un) : "l"(ut), "f"(-12582912.f));
    asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(uf) : "l"(ux), "l"(un));
    asm("{\n\t.reg .b64 c1, c2, c3, one, t;\n\t"
        "mov.b64 c3, {%2, %2};\n\tmov.b64 c2, {%3, %3};\n\tmov.b64 c1, {%4, %4};\n\tmov.b64 one, {%5, %5};\n\t"
        "fma.rn.ftz.f32x2 t, %1, c3, c2;\n\t"
        "fma.rn.ftz.f32x2 t, t, %1, c1;\n\t"
        "fma.rn.ftz.f32x2 %0, t, %1, one;\n\t}"
        : "=l"(up) : "l"(uf), "f"(0.07706617563962936f), "f"(0.22764593362808228f), "f"(0.6951165795326233f), "f"(1.0f));
    uint32_t p0, p1, t0, t1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(up));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(ut));
    float2 r;
    r.x = __uint_as_float(p0 + (t0 << 23));
    r.y = __uint_as_float(p1 + (t1 << 23));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// pack two fp32 into one 32-bit word of 16-bit floats: low half = lo, high half = hi
template <int DT>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) {
    uint32_t d;
    if constexpr (DT == kBF16) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    } else {
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    }
    return d;
}

__device__ __forceinline__ float4 ld_global_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_global_f4(float* p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 256-bit store (sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

}  // namespace fa
