// loaders.cuh — tile loaders of the B200 attention forward path.
//
// Takes the place of the reference's kernels/loaders.cuh:
//   * shared-memory carve-up      (reference: loaders.cuh:23-52,  [Q0|Q1|K0|K1|V0|V1|O] in floats)
//   * asyncBufferLoad / loader warps (reference: loaders.cuh:55-83, 114-203, per-lane cp.async pieces
//     issued by two producer warps that walk every pipeline phase in lock-step with the math warps)
// Here one elected thread issues TMA bulk-tensor copies (cp.async.bulk.tensor.4d) of whole
// 128-row tiles into 128B-swizzled shared memory; completion is signalled on mbarriers
// (expect_tx / complete_tx), and slots are handed back by tcgen05.commit from the MMA issuer.
//
// The kernel is persistent: one CTA per SM walks a queue of work items (256-row query blocks).
// Shared memory map (all tile buffers 1024-B aligned, required by SWIZZLE_128B):
//   [ Q tile 0 | Q tile 1 | KV ring slot 0 .. slot S-1 | mbarriers | work-item slots | tmem base ]
// A 128 x D tile of 16-bit elements is stored as D/64 "halves"; half h holds columns [64h, 64h+64)
// as 128 rows of 128 bytes (row r at byte r*128, 16-byte chunks XOR-swizzled by r%8) — exactly
// what a TMA box of {64, 128} with CU_TENSOR_MAP_SWIZZLE_128B writes and what the UMMA
// descriptors in computers.cuh describe.  K and V tiles alternate through one ring: K0 V0 K1 V1 ...
#pragma once

#include "utils.cuh"

namespace fa {

constexpr int kBlockM = 128;       // query rows per MMA tile (TMEM lanes)
constexpr int kTilesPerCta = 2;    // two query tiles per CTA, ping-ponged through the tensor pipe
constexpr int kBlockN = 128;       // key/value rows per tile
constexpr int kHalfCols = 64;      // columns per swizzle-128B half (64 x 2 B = 128 B)
constexpr int kHalfBytes = kBlockN * 128;   // 16 KiB: one half of a 128-row tile

// Warp roles.  The softmax side comes in two layouts, chosen per head dim at compile time (KCfg<D>):
//   8 softmax warps  (384 threads): warps 0-3 query tile 0, warps 4-7 query tile 1; one full score row (128 columns) per
//                    thread (tcgen05.ld 32x32b), row max / row sum need no shuffles (softmaxWarpgroup)
//   16 softmax warps (640 threads): 8 warps per query tile; the two warps that share a TMEM lane quarter split its 32 rows
//                    16 / 16 (tcgen05.ld 16x256b: a row is spread over the 4 threads of a quad, 2 x 32 scores per thread),
//                    so every SM sub-partition runs FOUR softmax warps and the MUFU time of one overlaps the FMA / ALU / TMEM
//                    instructions of the others (softmaxRows16).  A lone warp cannot overlap its own MUFU.EX2 with anything
//                    (8.14 clk per MUFU + its other instructions, scripts/microbench/pipes.cu), which is what bounds the
//                    8-warp layout at ~2,600 clk per pair of 128-key tiles against 2,048 clk of MMAs.
// Both layouts are compiled (template parameter SW of the kernel); the launcher picks one per problem from the measured
// tile table (FlashAttention.cu: kTileTable).
template <int SW>
struct KCfg {
    static constexpr int kSoftmaxWarps = SW;
    static_assert(kSoftmaxWarps == 8 || kSoftmaxWarps == 16, "8 or 16 softmax warps");
    static constexpr bool kRows16 = kSoftmaxWarps == 16;
    static constexpr int kSoftmaxThreadsPerTile = kSoftmaxWarps * 32 / kTilesPerCta;   // arrivals per query tile on s_free / p_full / o_free
    static constexpr int kMmaWarp0 = kSoftmaxWarps;           // MMA issuer: every Q K^T (by type) / query tile 0 (by tile)
    static constexpr int kLoadWarp = kSoftmaxWarps + 1;       // TMA producer + scheduler (one thread)
    static constexpr int kMmaWarp1 = kSoftmaxWarps + 2;       // MMA issuer: every P V (by type) / query tile 1 (by tile); also allocates / frees TMEM
    static constexpr int kTmemWarp = kMmaWarp1;
    static constexpr int kNumThreads = (kSoftmaxWarps + 4) * 32;
    // Register split (setmaxnreg): what __launch_bounds__(kNumThreads, 1) gives every thread at launch is re-divided between
    // the softmax warps and the rest; the CTA may not hold more than it was launched with.
    //   8 warps : 168 at launch -> 256 x 216 + 128 x 72;   16 warps: 96 at launch -> 512 x 104 + 128 x 64
    // (A fourth warpgroup that only stores O — 256 x 208 + 128 x 56 + 128 x 40 — was built and measured in round 2: with 40
    // registers it drains TMEM 16 columns at a time, the next item's first P V waits for it, and causal N <= 2K lost 11-27 %.)
    static constexpr int kLaunchRegs = (65536 / kNumThreads) / 8 * 8;
#ifdef FA_SOFTMAX_REGS
    static constexpr int kSoftmaxRegs = FA_SOFTMAX_REGS;
    static constexpr int kOtherRegs = FA_OTHER_REGS;
#else
    static constexpr int kSoftmaxRegs = kRows16 ? 104 : 216;
    static constexpr int kOtherRegs = kRows16 ? 64 : 72;
#endif
    static_assert(kSoftmaxWarps * 32 * kSoftmaxRegs + 128 * kOtherRegs <= kNumThreads * kLaunchRegs,
                  "register split exceeds what the CTA owns at launch");
};

// FMA-pipe exp2 share (template parameter EMU of the kernel): of every 8 consecutive score pairs, EMU take ex2_emu2 instead of
// MUFU.EX2.  Pays only where the MUFU is the binding pipe AND other warps can overlap the extra FMA work: d = 64 with 16
// softmax warps (-2.4 % cycles at N = 8K); at d = 128 it costs 3-10 % in every layout measured.

// Split wait for the previous P V (d = 128 only): the softmax warpgroup waits for the FIRST half of P_t V_{j-1} (o_half, an
// extra tcgen05.commit) before it overwrites the first half of P_t, and for the second half (o_full) only right before it
// overwrites the second half.  Measured on the shipped (un-instrumented) build: 2,712 -> 2,671 clk per step causal 8K,
// 2,670 -> 2,614 non-causal; wall clock +1 .. +4 % in bench.py, +0.9 .. +1.4 % held for 0.5 s under the power cap.  (The
// FA_PHASE_PROFILE build shows the opposite, 2,813 -> 2,940: its clock reads perturb exactly this hand-off; trust
// scripts/cycles.py.)  At d = 64, where P V is half as long, it loses 3.6 % and stays off.
#ifndef FA_SPLIT_OFULL
#define FA_SPLIT_OFULL 1
#endif
template <int D>
constexpr bool kSplitOFull = (FA_SPLIT_OFULL != 0) && D == 128;

// The two MMA issuer warps are split by type at d = 128 (all Q K^T / all P V, mmaTypeIssuerWarp) and by query tile at
// d = 64 (mmaIssuerWarp).  By type, measured against by tile on the shipped build: 2,672 -> 2,645 clk per step causal 8K,
// 2,613 -> 2,598 non-causal, 3,055 -> 2,949 causal 2K; +0.1 .. +0.9 % wall clock held under the power cap; at d = 64
// 2,559 -> 2,612 (worse: the P V there is too short to be worth a warp of its own).
#ifndef FA_ISSUER_BY_TYPE
#define FA_ISSUER_BY_TYPE 1
#endif
template <int D>
constexpr bool kIssuerByType = (FA_ISSUER_BY_TYPE != 0) && D == 128;

struct FwdParams {
    void* O;                 // output, same dtype as Q
    float* lse;              // optional [B, Hq, Nq] log-sum-exp (natural log), may be null
    float* acc_o;            // carry mode (ring-KV steps): fp32 [B, Hq, Nq, d] running output, updated in place; O is not written
    float* acc_lse;          // carry mode: fp32 [B, Hq, Nq] running log-sum-exp, updated in place
    int acc_rows, acc_off;   // carry mode: the running pair has acc_rows (>= Nq) rows per (batch, head); this call's query row i is its row acc_off + i
    int B, Hq, Hkv, Nq, Nk;
    long long o_stride_b, o_stride_h, o_stride_n;   // in elements; innermost (d) stride is 1
    float scale;             // softmax scale (1/sqrt(d) by default)
    float scale_log2;        // scale * log2(e)
    int causal;              // 0/1
    int causal_off;          // Nk - Nq: key j visible to query i iff j <= i + causal_off
    int q_heads_per_kv;      // Hq / Hkv
    // Exact division of the work-item index by the three run-time divisors above with one multiply + shift each (host-made
    // magic numbers, FastDiv): every role decodes every item, and three hardware integer divisions on one thread cost ~1,400
    // clk at each item boundary (scripts/trace_cta.py) — 10 % of a five-step item.
    unsigned div_qblocks_mul, div_qblocks_shr, div_hq_mul, div_hq_shr, div_group_mul, div_group_shr;
    int num_q_blocks;        // 256-row query blocks per (batch, head)
    int total_items;         // work items of the launch: n_full_items 256-row items, then two 128-row items for every remaining query block
    int n_full_items;        // the first n_full_items query blocks (in queue order) are one 256-row item each; the rest are split in halves
    int split_half;          // 1: a half item runs on BOTH query-tile slots — slot t takes key tiles t, t+2, ... of the same 128 rows and
                             //    the two partial results are merged in the epilogue (8-warp layouts, plain mode); 0: slot 0 alone
    int pair_heads;          // CTA-pair kernel: query heads per work item — 1: pairs cut by rows (512 rows of one head), 2: by two heads
                             //    of a kv group (256 rows each), 4: by four heads (128 rows each); see decode_pair_item
    int* sched_counter;      // device int, zero at launch and left zero by the launch: next work item = gridDim.x + atomicAdd(counter, 1)
    unsigned long long* prof; // FA_PHASE_PROFILE builds only: per-phase cycle counters (see scripts/phase_profile.py)
};

// CG = 2 (CTA-pair kernel, tcgen05 cta_group::2): a ring slot holds this CTA's HALF of a K tile (64 of the 128 keys, all d
// columns) or of a V tile (all 128 keys, 64 of the d columns) — 16 KiB at d = 128; everything else is laid out as for CG = 1.
template <int D, int STAGES, int CG = 1>
struct SmemLayout {
    static constexpr int kQTileBytes = kBlockM * D * 2;
    static constexpr int kKVTileBytes = kBlockN * D * 2 / CG;
    static constexpr int kQOff = 0;
    static constexpr int kKVOff = kTilesPerCta * kQTileBytes;
    static constexpr int kBarOff = kKVOff + STAGES * kKVTileBytes;
    // barrier indices
    static constexpr int kBarQFull = 0;                        //        TMA -> MMA     : both query tiles landed
    static constexpr int kBarQEmpty = 1;                       //        MMA -> TMA     : last Q K^T of the item retired (both issuers)
    static constexpr int kBarKVFull = 2;
    static constexpr int kBarKVEmpty = kBarKVFull + STAGES;
    static constexpr int kBarSFull = kBarKVEmpty + STAGES;     // [2]    MMA -> softmax : S tile ready in TMEM
    static constexpr int kBarPFull = kBarSFull + 2;            // [2][2] softmax -> MMA : first / second 64 keys of P written (O rescaled)
    static constexpr int kBarOFull = kBarPFull + 4;            // [2]    MMA -> softmax : P*V of this step retired
    static constexpr int kBarOFree = kBarOFull + 2;            // [2]    softmax -> MMA : epilogue has read O out of TMEM
    static constexpr int kBarSchedFull = kBarOFree + 2;        // [2]    TMA -> all     : next work item published
    static constexpr int kBarSchedEmpty = kBarSchedFull + 2;   // [2]    all -> TMA     : work item slot consumed
    static constexpr int kBarSFree = kBarSchedEmpty + 2;       // [2]    softmax -> MMA : S tile of query tile t copied into registers
    static constexpr int kBarOHalf = kBarSFree + 2;            // [2]    MMA -> softmax : first half (keys 0..63) of P*V of this step retired
    static constexpr int kNumBars = kBarOHalf + 2;
    // work-item mailbox: two slots of 16 ints — the item index and its DECODED form (WorkItem), written by the producer one
    // item ahead.  Decoding is ~1,000 clk of dependent single-thread arithmetic; done once by the producer, off the
    // critical path, instead of by every role at every item boundary.
    static constexpr int kSchedItemOff = (kBarOff + kNumBars * 8 + 15) & ~15;
    static constexpr int kSchedSlotBytes = 64;
    static constexpr int kTmemPtrOff = kSchedItemOff + 2 * kSchedSlotBytes;
    // 16-softmax-warp layout: per query tile and row, (1 / row sum, log-sum-exp) handed from the warp that owns the row in the
    // 16-lane layout to the warp that stores it in the epilogue (float2[2][128])
    static constexpr int kExchOff = kTmemPtrOff + 16;
    static constexpr int kExchBytes = kTilesPerCta * kBlockM * 8;
    static constexpr int kBytes = kExchOff + kExchBytes;
    // staged epilogue (template parameter ST of the kernel): one 128-row x 64-column output piece per query tile, in the
    // 128B-swizzled layout a TMA store reads (16 KiB each); d = 128 pays for it with a 4-slot K/V ring
    static constexpr int kStageOff = (kBytes + 1023) & ~1023;
    static constexpr int kStageTileBytes = kBlockM * 128;
    static constexpr int kBytesStaged = kStageOff + kTilesPerCta * kStageTileBytes;
    static_assert(kBytes <= 232448, "more than 227 KB of shared memory");
    // The dynamic shared-memory window of a kernel without static shared memory starts 1024-B aligned (the kernel traps if
    // it ever does not), so no alignment slack is reserved.
    static constexpr int kDynamicBytes = kBytes;
};

// n / d for 0 <= n < 2^31 with the (mul, shr) pair the host made for d (FlashAttention.cu: make_fast_div): exact
// (host-callable too: tests/test_abi.py replays the item decode on the CPU through fa_debug_decode_items)
__host__ __device__ __forceinline__ int fast_div(int n, unsigned mul, unsigned shr) {
#ifdef __CUDA_ARCH__
    return shr >= 32u ? n : int(__umulhi(unsigned(n), mul) >> shr);      // shr = 32 marks d == 1
#else
    return shr >= 32u ? n : int(unsigned(((unsigned long long)unsigned(n) * mul) >> 32) >> shr);
#endif
}

// One work item: a 256-row query block of one (batch, head) — or, for the tail of a small launch, one 128-row half of
// such a block (then only query-tile slot 0 of the CTA works; a half item costs ~0.6 of a full one, so splitting the last
// partial wave of blocks into halves lets all SMs finish together: BASELINE configs[1] is 192 blocks on 148 SMs).
struct WorkItem {
    int b, h, h_kv;
    int q0;        // first query row of the item
    int rows;      // 256, or 128 for a half item
    int split;     // half item in split-KV mode: both slots work on rows [q0, q0 + 128), slot t on key tiles t, t+2, ...
    int n_kv;      // number of 128-row key/value tiles the item loads (that of its busiest query tile)
    int n_steps;   // steps of the item on the shared score buffer: n_kv, or ceil(n_kv / 2) in split-KV mode
    int n_tile0, n_tile1;   // steps each query-tile slot takes part in (causal: the early tile stops one sooner; split: ceil / floor of n_kv / 2)
    int hstep;     // CTA-pair kernel, pairs cut by four heads: slot t of a CTA works on head h + t (same rows); 0 everywhere else
    __host__ __device__ __forceinline__ int n_tile(int t) const { return t == 0 ? n_tile0 : n_tile1; }
    // key tile slot t multiplies with at its step s, and the first row of slot t's query tile
    __host__ __device__ __forceinline__ int kv_tile(int t, int s) const { return split ? 2 * s + t : s; }
    __host__ __device__ __forceinline__ int tile_row0(int t) const { return split ? q0 : q0 + t * kBlockM; }
};

// Items are numbered (batch, head)-major so that CTAs running at the same time share K/V through L2; inside a head
// the heavy (late) causal query blocks come first, which makes the dynamic scheduler an LPT queue.
__host__ __device__ __forceinline__ WorkItem decode_item(const FwdParams& p, int item) {
    WorkItem w;
    int blk = item, half = 0;
    w.rows = kTilesPerCta * kBlockM;
    w.split = 0;
    w.hstep = 0;
    if (item >= p.n_full_items) {
        const int r = item - p.n_full_items;
        blk = p.n_full_items + (r >> 1);
        half = r & 1;
        w.rows = kBlockM;
        w.split = p.split_half;
    }
    const int bh = fast_div(blk, p.div_qblocks_mul, p.div_qblocks_shr);
    const int r = blk - bh * p.num_q_blocks;
    const int qb = p.causal ? (p.num_q_blocks - 1 - r) : r;
    w.b = fast_div(bh, p.div_hq_mul, p.div_hq_shr);
    w.h = bh - w.b * p.Hq;
    w.h_kv = fast_div(w.h, p.div_group_mul, p.div_group_shr);
    w.q0 = qb * (kTilesPerCta * kBlockM) + half * kBlockM;
    const int n_all = (p.Nk + kBlockN - 1) / kBlockN;
    w.n_kv = 0;
#pragma unroll
    for (int t = 0; t < kTilesPerCta; ++t) {
        int n = n_all;
        if (p.causal) {
            const int last_col = w.q0 + (t + 1) * kBlockM - 1 + p.causal_off;   // last key any row of the tile may see
            const int n_c = last_col < 0 ? 0 : last_col / kBlockN + 1;
            n = n_c < n ? n_c : n;
        }
        if (w.q0 + t * kBlockM >= p.Nq) n = 0;    // tile entirely past the end of the sequence
        if (t * kBlockM >= w.rows) n = 0;         // half item: query-tile slot 1 has no rows
        if (t == 0) w.n_tile0 = n; else w.n_tile1 = n;
        w.n_kv = n > w.n_kv ? n : w.n_kv;
    }
    w.n_steps = w.n_kv;
    if (w.split) {      // slot 0's tile count is the item's (slot 1 was given no rows above): deal the key tiles out alternately
        w.n_tile1 = w.n_kv / 2;
        w.n_tile0 = w.n_kv - w.n_tile1;
        w.n_steps = w.n_tile0;
    }
    return w;
}

// Consumer side of the work-item hand-off (whole warp): returns the item index (-1 when the queue is drained) and the decoded
// item the producer left in the mailbox slot (12 ints, three 128-bit shared-memory loads from a warp-uniform address).
template <int D, int STAGES, int CG = 1>
__device__ __forceinline__ int fetch_item(uint32_t smem_base, int k, WorkItem& w) {
    using L = SmemLayout<D, STAGES, CG>;
    const uint32_t bar0 = smem_base + L::kBarOff;
    const int slot = k & 1;
    // CTA pair: the leader's producer writes both CTAs' mailboxes (remote stores + a remote arrive), and every consumer of
    // either CTA hands the slot back on the LEADER's sched_empty barrier
    if constexpr (CG == 2) mbar_wait_cluster(bar0 + 8 * (L::kBarSchedFull + slot), (k >> 1) & 1);
    else mbar_wait(bar0 + 8 * (L::kBarSchedFull + slot), (k >> 1) & 1);
    const uint32_t a = smem_base + L::kSchedItemOff + L::kSchedSlotBytes * slot;
    int item, pad;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(item), "=r"(w.b), "=r"(w.h), "=r"(w.h_kv) : "r"(a) : "memory");
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(w.q0), "=r"(w.rows), "=r"(w.split), "=r"(w.n_kv) : "r"(a + 16) : "memory");
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(w.n_steps), "=r"(w.n_tile0), "=r"(w.n_tile1), "=r"(pad) : "r"(a + 32) : "memory");
    w.hstep = CG == 2 ? pad : 0;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(bar0 + 8 * (L::kBarSchedEmpty + slot), 0));
        else mbar_arrive(bar0 + 8 * (L::kBarSchedEmpty + slot));
    }
    return item;
}


// Producer: a single thread.  It is also the scheduler: it claims the CTA's next work item (the first one is
// blockIdx.x, later ones come from a global atomic counter), publishes it to the other roles one item ahead, then
// streams that item's tiles: Q tiles once, then K_j, V_j for j = 0..n_kv-1 through the ring.
template <int D, int STAGES>
__device__ __forceinline__ void tmaLoaderThread(const CUtensorMap* tmQ, const CUtensorMap* tmK,
                                                const CUtensorMap* tmV, uint32_t smem_base, const FwdParams& p) {
    using L = SmemLayout<D, STAGES>;
    constexpr int kHalves = D / kHalfCols;
    const uint32_t bar0 = smem_base + L::kBarOff;
    const uint32_t q_full = bar0 + 8 * L::kBarQFull;
    const uint32_t q_empty = bar0 + 8 * L::kBarQEmpty;

    FA_T2_DECL;
    FA_T2(p.prof, 1, 60);      // producer running
    int it = 0;      // K/V ring fills so far
    int kq = 0;      // Q loads so far
    int item = blockIdx.x;
    for (int k = 0;; ++k) {
        const int slot = k & 1;
        mbar_wait(bar0 + 8 * (L::kBarSchedEmpty + slot), ((k >> 1) & 1) ^ 1);
        const int pub = item < p.total_items ? item : -1;
        WorkItem w{};
        if (pub >= 0) w = decode_item(p, item);
        {
            const uint32_t a = smem_base + L::kSchedItemOff + L::kSchedSlotBytes * slot;
            asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pub), "r"(w.b), "r"(w.h), "r"(w.h_kv) : "memory");
            asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16), "r"(w.q0), "r"(w.rows), "r"(w.split), "r"(w.n_kv) : "memory");
            asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a + 32), "r"(w.n_steps), "r"(w.n_tile0), "r"(w.n_tile1), "r"(0) : "memory");
        }
        mbar_arrive(bar0 + 8 * (L::kBarSchedFull + slot));
        FA_T2(p.prof, 1, 3);
        if (pub < 0) break;
        if (w.n_kv > 0) {
#ifndef FA_NO_Q_PREFETCH
            // The Q tiles can only land once the previous item's last Q K^T has retired, and they come from HBM (each is read
            // exactly once): start them towards L2 now, two or three key tiles early, so that the load below is an L2 hit and
            // the tensor pipe's bubble at an item boundary shrinks by most of a DRAM round trip.
            if (k > 0)
                for (int t = 0; t < w.rows / kBlockM; ++t)
#pragma unroll
                    for (int hf = 0; hf < kHalves; ++hf) tma_prefetch_l2_4d(tmQ, hf * kHalfCols, w.q0 + t * kBlockM, w.h, w.b);
#endif
            FA_T2(p.prof, 1, 61);      // item decoded
            mbar_wait(q_empty, (kq & 1) ^ 1);      // the previous item's last Q K^T has retired
            FA_T2(p.prof, 1, 62);      // Q buffers free
            ++kq;
            const int q_tiles = w.split ? kTilesPerCta : w.rows / kBlockM;    // split-KV: the same 128 rows into both Q buffers
            mbar_expect_tx(q_full, q_tiles * L::kQTileBytes);
            for (int t = 0; t < q_tiles; ++t)
#pragma unroll
                for (int hf = 0; hf < kHalves; ++hf)
                    tma_load_4d_hint(tmQ, smem_base + L::kQOff + t * L::kQTileBytes + hf * kHalfBytes, q_full,
                                     hf * kHalfCols, w.tile_row0(t), w.h, w.b, kEvictFirst);
            FA_T2(p.prof, 1, 4);
            for (int j = 0; j < w.n_kv; ++j) {
#pragma unroll
                for (int kv = 0; kv < 2; ++kv, ++it) {
                    const int s = it % STAGES;
                    const uint32_t parity = (it / STAGES) & 1;
                    const uint32_t full = bar0 + 8 * (L::kBarKVFull + s);
                    mbar_wait(bar0 + 8 * (L::kBarKVEmpty + s), parity ^ 1);
#ifdef FA_EXP_NOKV
                    // Energy experiment, WRONG RESULTS (scripts/README.md): once the ring has been filled, K/V tiles are not
                    // loaded at all (1) or only their first 64 columns are (2) — the same instruction stream and MMA work with
                    // none / half of the L2 -> shared-memory traffic.  Sizes what K/V multicast across a CTA pair could save.
                    if (it >= STAGES) {
                        if (FA_EXP_NOKV == 1) { mbar_arrive(full); continue; }
                        mbar_expect_tx(full, L::kKVTileBytes / kHalves);
                        tma_load_4d_hint(kv == 0 ? tmK : tmV, smem_base + L::kKVOff + s * L::kKVTileBytes, full, 0, j * kBlockN, w.h_kv, w.b, kEvictLast);
                        continue;
                    }
#endif
                    mbar_expect_tx(full, L::kKVTileBytes);
                    FA_T2(p.prof, 1, 5);
                    const CUtensorMap* tm = kv == 0 ? tmK : tmV;
#pragma unroll
                    for (int hf = 0; hf < kHalves; ++hf)
                        tma_load_4d_hint(tm, smem_base + L::kKVOff + s * L::kKVTileBytes + hf * kHalfBytes, full,
                                         hf * kHalfCols, j * kBlockN, w.h_kv, w.b, kEvictLast);
                }
            }
        }
        // Claim the next item.  Every CTA keeps claiming until a claim fails, so a launch makes exactly total_items claims;
        // whoever makes the last one (no other claim can follow) puts the counter back to zero for the next launch that uses
        // it — the host never has to clear it (no memset node per launch).
        const int claimed = atomicAdd(p.sched_counter, 1);
        if (claimed == p.total_items - 1) atomicExch(p.sched_counter, 0);
        item = int(gridDim.x) + claimed;
    }
    // Tail: wait until the consumer has handed back the last fills, so that no tcgen05.commit arrive is still in
    // flight towards this CTA's shared memory when the CTA exits.
    const int total = it;
    for (int i = (total > STAGES ? total - STAGES : 0); i < total; ++i)
        mbar_wait(bar0 + 8 * (L::kBarKVEmpty + i % STAGES), (i / STAGES) & 1);
    if (kq > 0) mbar_wait(q_empty, (kq - 1) & 1);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel (d = 128).  Every MMA covers 256 query rows, 128 from each CTA, against ONE K / V tile — the two CTAs must
// want the same keys at the same time.  Three ways to cut the work (FwdParams::pair_heads = heads per item: 1, 2, 4):
//   by rows    (any head layout): a work item is a 512-row query block of one (batch, head); MMA tile t is the rows
//              [q0 + 256 t, q0 + 256 t + 256), the leader CTA owns the first 128 of them, the peer the last 128.  Tile counts are
//              those of the whole 256-row tile: on a causal diagonal the leader's half of the tile's last key tile is fully
//              masked (its softmax sees -inf and writes P = 0) — one step per item more than two 1-CTA items would take.
//   by 2 heads (Hq / Hkv even): a work item is a 256-row block of TWO query heads of one kv group; the leader takes head 2 hp,
//              the peer head 2 hp + 1, both the same rows; slot t of each CTA is the rows [q0 + 128 t, +128).  Same keys, same
//              diagonal, same mask in both CTAs: nothing is lost on causal problems, items as fine as the 1-CTA kernel's.
//   by 4 heads (Hq / Hkv a multiple of 4): a work item is a 128-row block of FOUR query heads of one kv group; the leader's
//              slots take heads 4 hq and 4 hq + 1, the peer's 4 hq + 2 and 4 hq + 3, all four the same 128 rows.  Now the two
//              slots of a CTA reach the diagonal together as well: the step in which only the later row block has work (one of
//              every item's n + 1 steps on a causal problem) is gone.
// The mailbox carries, per CTA, its own first head and first row; `rows` tells the row stride between the CTA's two tiles
// (rows / 2, or 0 with `hstep` = 1: then slot t works on head h + t).
// ------------------------------------------------------------------------------------------------
constexpr int kPairRows = 2 * kTilesPerCta * kBlockM;      // 512

__host__ __device__ __forceinline__ WorkItem decode_pair_item(const FwdParams& p, int item) {
    WorkItem w;
    const int hpi = p.pair_heads;                       // heads per item: 1, 2 or 4
    w.rows = hpi == 1 ? kPairRows : (hpi == 2 ? kTilesPerCta * kBlockM : kBlockM);
    w.hstep = hpi == 4 ? 1 : 0;
    const int tile_rows = hpi == 1 ? 2 * kBlockM : kBlockM;      // rows of one MMA tile that decide its key range: 256 (both CTAs' halves) / 128
    const int tile_step = hpi == 4 ? 0 : tile_rows;              // rows between the item's two MMA tiles
    w.split = 0;
    const int bh = fast_div(item, p.div_qblocks_mul, p.div_qblocks_shr);
    const int r = item - bh * p.num_q_blocks;
    const int qb = p.causal ? (p.num_q_blocks - 1 - r) : r;
    // (batch, head GROUP of hpi heads) indexes the items: the host made div_hq for Hq / hpi
    const int groups = p.Hq / hpi;
    w.b = fast_div(bh, p.div_hq_mul, p.div_hq_shr);
    w.h = (bh - w.b * groups) * hpi;                    // the leader's (first) head
    w.h_kv = fast_div(w.h, p.div_group_mul, p.div_group_shr);
    w.q0 = qb * w.rows;
    const int n_all = (p.Nk + kBlockN - 1) / kBlockN;
    w.n_kv = 0;
#pragma unroll
    for (int t = 0; t < kTilesPerCta; ++t) {
        int n = n_all;
        if (p.causal) {
            const int last_col = w.q0 + t * tile_step + tile_rows - 1 + p.causal_off;
            const int n_c = last_col < 0 ? 0 : last_col / kBlockN + 1;
            n = n_c < n ? n_c : n;
        }
        if (w.q0 + t * tile_step >= p.Nq) n = 0;
        if (t == 0) w.n_tile0 = n; else w.n_tile1 = n;
        w.n_kv = n > w.n_kv ? n : w.n_kv;
    }
    w.n_steps = w.n_kv;
    return w;
}

// Producer thread of one CTA of a pair.  The leader's is also the pair's scheduler: it claims items, decodes them and writes
// BOTH mailboxes (the peer's through distributed shared memory; q0 there is the peer's own first row).  Each producer loads
// its own CTA's Q rows and its half of every K / V tile; all load completions land on the LEADER's full barriers (the MMA
// issuers live there), slots come back through multicast tcgen05.commit arrivals on each CTA's own empty barriers.
template <int D, int STAGES>
__device__ __forceinline__ void tmaPairLoaderThread(const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                                    uint32_t smem_base, const FwdParams& p) {
    using L = SmemLayout<D, STAGES, 2>;
    static_assert(D == 128 && STAGES % 2 == 0, "pair kernel: d = 128, K in the even ring slots and V in the odd ones");
    const uint32_t rank = cluster_ctarank();
    const uint32_t bar0 = smem_base + L::kBarOff;
    const uint32_t lead0 = mapa_shared(bar0, 0);
    const uint32_t q_full_l = lead0 + 8 * L::kBarQFull;
    const uint32_t q_empty = bar0 + 8 * L::kBarQEmpty;
    constexpr int kKHalfBytes = (kBlockN / 2) * 128;      // 64 keys x 64 columns

    int it = 0, kq = 0;
    int item = int(cluster_id_x());
    for (int k = 0;; ++k) {
        const int slot = k & 1;
        WorkItem w{};
        int pub;
        if (rank == 0) {
            mbar_wait_cluster(bar0 + 8 * (L::kBarSchedEmpty + slot), ((k >> 1) & 1) ^ 1);
            pub = item < p.total_items ? item : -1;
            if (pub >= 0) w = decode_pair_item(p, item);
#pragma unroll
            for (uint32_t c = 0; c < 2; ++c) {
                const uint32_t a = mapa_shared(smem_base + L::kSchedItemOff + L::kSchedSlotBytes * slot, c);
                // the peer's own first head (pairs by heads: + 1 or + 2) or own first row (pairs by rows)
                st_shared_cluster_v4(a, pub, w.b, w.h + int(c) * (p.pair_heads >> 1), w.h_kv);
                st_shared_cluster_v4(a + 16, w.q0 + (p.pair_heads == 1 ? int(c) * kBlockM : 0), w.rows, 0, w.n_kv);
                st_shared_cluster_v4(a + 32, w.n_steps, w.n_tile0, w.n_tile1, w.hstep);
                mbar_arrive_cluster_release(mapa_shared(bar0 + 8 * (L::kBarSchedFull + slot), c));
            }
        } else {
            mbar_wait_cluster(bar0 + 8 * (L::kBarSchedFull + slot), (k >> 1) & 1);
            const uint32_t a = smem_base + L::kSchedItemOff + L::kSchedSlotBytes * slot;
            int pad;
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(pub), "=r"(w.b), "=r"(w.h), "=r"(w.h_kv) : "r"(a) : "memory");
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(w.q0), "=r"(w.rows), "=r"(w.split), "=r"(w.n_kv) : "r"(a + 16) : "memory");
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(w.n_steps), "=r"(w.n_tile0), "=r"(w.n_tile1), "=r"(pad) : "r"(a + 32) : "memory");
            w.hstep = pad;
            mbar_arrive_cluster(lead0 + 8 * (L::kBarSchedEmpty + slot));
        }
        if (pub < 0) break;
        if (w.n_kv > 0) {
            mbar_wait(q_empty, (kq & 1) ^ 1);      // the previous item's last Q K^T has retired (multicast commit)
            ++kq;
            mbar_expect_tx_cluster(q_full_l, kTilesPerCta * L::kQTileBytes);
#pragma unroll
            for (int t = 0; t < kTilesPerCta; ++t)
#pragma unroll
                for (int hf = 0; hf < D / kHalfCols; ++hf)
                    tma_load_4d_pair(tmQ, smem_base + L::kQOff + t * L::kQTileBytes + hf * kHalfBytes, q_full_l,
                                     hf * kHalfCols, w.q0 + (w.hstep ? 0 : t * (w.rows >> 1)), w.h + t * w.hstep, w.b, kEvictFirst);
            for (int j = 0; j < w.n_kv; ++j) {
                {   // this CTA's 64 keys of K_j, both 64-column halves
                    const int s = it % STAGES;
                    mbar_wait(bar0 + 8 * (L::kBarKVEmpty + s), ((it / STAGES) & 1) ^ 1);
                    const uint32_t full_l = lead0 + 8 * (L::kBarKVFull + s);
                    mbar_expect_tx_cluster(full_l, L::kKVTileBytes);
#pragma unroll
                    for (int hf = 0; hf < D / kHalfCols; ++hf)
                        tma_load_4d_pair(tmK, smem_base + L::kKVOff + s * L::kKVTileBytes + hf * kKHalfBytes, full_l,
                                         hf * kHalfCols, j * kBlockN + int(rank) * (kBlockN / 2), w.h_kv, w.b, kEvictLast);
                    ++it;
                }
                {   // this CTA's 64 columns of V_j, all 128 keys
                    const int s = it % STAGES;
                    mbar_wait(bar0 + 8 * (L::kBarKVEmpty + s), ((it / STAGES) & 1) ^ 1);
                    const uint32_t full_l = lead0 + 8 * (L::kBarKVFull + s);
                    mbar_expect_tx_cluster(full_l, L::kKVTileBytes);
                    tma_load_4d_pair(tmV, smem_base + L::kKVOff + s * L::kKVTileBytes, full_l,
                                     int(rank) * kHalfCols, j * kBlockN, w.h_kv, w.b, kEvictLast);
                    ++it;
                }
            }
        }
        if (rank == 0) {      // claim the pair's next item (same self-resetting counter as the 1-CTA kernel: one claim per cluster)
            const int claimed = atomicAdd(p.sched_counter, 1);
            if (claimed == p.total_items - 1) atomicExch(p.sched_counter, 0);
            item = int(cluster_count_x()) + claimed;
        }
    }
    const int total = it;
    for (int i = (total > STAGES ? total - STAGES : 0); i < total; ++i)
        mbar_wait(bar0 + 8 * (L::kBarKVEmpty + i % STAGES), (i / STAGES) & 1);
    if (kq > 0) mbar_wait(q_empty, (kq - 1) & 1);
}

}  // namespace fa
