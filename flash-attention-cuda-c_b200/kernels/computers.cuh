// computers.cuh — the math stage of the B200 attention forward path.
//
// Takes the place of the reference's kernels/computers.cuh (reference: computers.cuh:5-69, one warp
// per query row, one lane group per key row, scalar FFMA dot products + shuffle reductions, smem O
// read-modify-write) and of the per-tile helpers it calls (reference: utils.cuh:17-113).
//
// Here the two contractions are tcgen05 MMAs accumulating in tensor memory:
//   S   = Q_t K_j^T   (A = Q tile, B = K tile, both K-major in 128B-swizzled smem)    -> the shared S buffer
//   O_t += P_t V_j    (A = P_t in TMEM as packed 16-bit, B = V tile MN-major in smem) -> O_t
// issued by TWO issuer warps (one elected lane per batch of MMAs): at d = 128 one warp issues every Q K^T and the other
// every P V (mmaTypeIssuerWarp; mmaPairIssuerWarp issues the same schedule as 256-row cta_group::2 MMAs for a CTA pair,
// fwdSm100PairKernel), at d = 64 there is one issuer per query tile (mmaIssuerWarp).  Two softmax
// warpgroups (one per query tile) read S with tcgen05.ld in the 32x32b shape — thread i of warp w owns TMEM
// lane 32*(w%4)+i, i.e. one whole score row, so row max / row sum need no shuffles — and keep the
// online-softmax state (running max m, running sum l) in registers:
//   m' = max(m, rowmax(S));  P = exp2(S*c - m*c)  (c = scale*log2 e);  l += rowsum(P)
// Normalisation by 1/l is deferred to the epilogue (the reference normalises every tile,
// utils.cuh:79-80).  O is rescaled lazily: only when a row's max grew by more than 2^8 since the max
// in use (then O *= exp2((m_used - m')c) through a tcgen05.ld / tcgen05.st round trip).
//
// TMEM map (512 columns):  S 0..127 | P_0 128..191 | P_1 192..255 | O_0 256..383 | O_1 384..511.
// ONE score buffer is shared by both query tiles: a tile's softmax warpgroup copies its S row into registers
// within ~100 clk of the MMA retiring and hands the buffer back (s_free), so the buffer is free long before the
// tensor pipe needs it again; P_t (two 16-bit values per 32-bit column) has columns of its own.  With two S tiles
// and P_t aliased over the head of S_t (the first design of this kernel) "Q_t K_{j+1}^T" had to queue behind
// "P_t V_j" and the tensor pipe idled for a third of every step; here the next score tile of a query tile is
// computed while its softmax warpgroup is still exponentiating the current one.
#pragma once

#include "loaders.cuh"

namespace fa {

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemS = 0;       // the shared S buffer
constexpr uint32_t kTmemO0 = 256;    // O tile t at columns 256 + 128*t
__host__ __device__ constexpr uint32_t tmem_p_col(int t) { return 128u + 64u * uint32_t(t); }   // P tile of query tile t
constexpr float kRescaleThreshold = 8.0f;   // log2 units
constexpr int kNoRow = 0x3fff0000;          // first row of a query-tile slot that has no rows (beyond any Nq)

// ------------------------------------------------------------------------------------------------
// MMA issuers: one warp per query tile t (warps kMmaWarp0 / kMmaWarp1).  The whole warp walks the schedule (so that
// addresses and descriptors stay warp-uniform); one elected lane issues each batch of tcgen05.mma and its
// tcgen05.commit.  Persistent: loops over the work items the producer publishes.
// tcgen05.mma issue blocks while the tensor pipe's short queue is full, so an issuing warp runs in lock-step with the
// pipe and every mbarrier wait it makes (~100 clk even when the barrier has already completed) is a bubble in the
// pipe.  With two issuers the bubbles of one are filled by the MMAs of the other, and the order in which the two query
// tiles' MMAs reach the pipe follows readiness instead of a fixed program order.
// Per key tile j issuer t does:   Q_tK_{j+1} (once the previous S tile has been copied out of the shared buffer), P_tV_j
// K/V slots and the Q tiles are handed back to the producer by BOTH issuers (barrier count 2): a tcgen05.commit when the
// issuer multiplied with the tile, a plain arrive when it did not (causal blocks: the early query tile stops one key
// tile sooner) — in both cases only after it has seen the slot full, which keeps the phases in step.
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT, int SW, int HS>
__device__ __forceinline__ void mmaIssuerWarp(uint32_t smem_base, uint32_t tmem_base_in, const FwdParams& p, const int t) {
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);   // tell the compiler it is warp-uniform
    using L = SmemLayout<D, STAGES>;
    constexpr uint32_t kFmt = (DT == kBF16) ? 1u : 0u;
    constexpr uint32_t idesc_qk = umma_idesc(kBlockM, kBlockN, kFmt, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc(kBlockM, D, kFmt, 0, 1);
    const uint32_t bar0 = smem_base + L::kBarOff;
    auto bar = [&](int i) { return bar0 + 8u * uint32_t(i); };

    // Descriptor templates with a zero start address; the 14-bit address field (bytes >> 4) is added per MMA.
    const uint64_t desc_k_major = umma_desc_sw128(0, 16, 1024);             // Q and K tiles (K-major)
    const uint64_t desc_mn_major = umma_desc_sw128(0, kHalfBytes, 1024);    // V tile (MN-major), 64-column halves 16 KiB apart
    const uint64_t q_desc = desc_k_major + ((smem_base + L::kQOff + t * L::kQTileBytes) >> 4);
    const uint32_t s_tmem = tmem_base + kTmemS;
    const uint32_t p_tmem = tmem_base + tmem_p_col(t);
    const uint32_t o_tmem = tmem_base + kTmemO0 + 128u * t;
    FA_PROF_DECL(6);
    FA_T2_DECL;

    // S_t = Q_t K^T, then s_full[t]; `then_release` != 0: also hand the K slot / the Q tiles back (commit = arrive on completion)
    auto issue_qk = [&](uint32_t k_smem, uint32_t release_bar, bool release_q) {
        if (elect_one_sync()) {
            const uint64_t b0 = desc_k_major + (k_smem >> 4);
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
                // 16 halfs = 32 B inside the 128-B swizzle row; the second 64 columns live one half (16 KiB) further
                const uint32_t off = ((ks / 4) * kHalfBytes + (ks % 4) * 32) >> 4;
                umma_ss(s_tmem, q_desc + off, b0 + off, idesc_qk, ks > 0);
            }
            tc_commit(bar(L::kBarSFull + t));
            tc_commit(release_bar);
            if (release_q) tc_commit(bar(L::kBarQEmpty));
        }
        __syncwarp();
        FA_T2(p.prof, 2 + t, 6 + t);
    };
    // O_t (+)= P_t V_j in two halves of 4 k-steps (64 keys each): the first half can start while the softmax warpgroup
    // is still producing the second half of P.  The second half ends with o_full[t] and the release of the V slot.
    auto issue_pv_half = [&](uint32_t v_smem, bool accumulate, int half, uint32_t release_bar) {
        if (elect_one_sync()) {
            const uint64_t b0 = desc_mn_major + (v_smem >> 4);
#pragma unroll
            for (int kk = 0; kk < kBlockN / 32; ++kk) {
                const int ks = half * (kBlockN / 32) + kk;
                // 16 key rows = 2 swizzle atoms of 8 rows x 128 B = 2048 B
                umma_ts(o_tmem, p_tmem + 8u * ks, b0 + ((ks * 2048) >> 4), idesc_pv, (accumulate || ks > 0) ? 1u : 0u);
            }
            if (half == 1) {
                tc_commit(bar(L::kBarOFull + t));
                tc_commit(release_bar);
            } else if (kSplitOFull<D>) {
                tc_commit(bar(L::kBarOHalf + t));
            }
        }
        __syncwarp();
    };
    auto slot_addr = [&](int it) { return smem_base + L::kKVOff + (it % STAGES) * L::kKVTileBytes; };
    auto empty_bar = [&](int it) { return bar(L::kBarKVEmpty + it % STAGES); };
    auto wait_full = [&](int it) {
        mbar_wait(bar(L::kBarKVFull + it % STAGES), (it / STAGES) & 1);
        tc_fence_after();
    };
    auto arrive = [&](uint32_t b) { if (elect_one_sync()) mbar_arrive(b); __syncwarp(); };

    int it0 = 0;              // ring position of this item's K_0  (K_j = it0 + 2j, V_j = it0 + 2j + 1)
    int kq = 0;               // items with work so far (Q loads consumed)
    int st = 0;               // key tiles this issuer's query tile has processed so far (barrier phase bookkeeping)
    int sq = 0;               // score-buffer steps so far (sum of n_kv over items; phase bookkeeping of s_free)
    int ko = 0;               // items in which this issuer's query tile had work (O hand-back bookkeeping)
    for (int k = 0;; ++k) {
        WorkItem w;
        const int item = fetch_item<D, STAGES>(smem_base, k, w);
        if (item < 0) break;
        const int n = w.n_kv;
        if (n <= 0) continue;
        const int nt = w.n_tile(t);

        // Shared S buffer: the CTA-wide order of the score tiles of an item is S_0(0), S_1(0), S_0(1), S_1(1), ... for ALL
        // j < n; a tile a query tile does not take part in is a virtual step (its issuer passes the buffer on without
        // an MMA), so both s_free barriers advance exactly n phases per item and every wait below is on the phase right
        // after the one this warp waited on before.  S_t(j) may be written once its predecessor has been copied out.
        auto wait_s_buffer = [&](int j) {
            const int idx = (t == 0) ? sq + j - 1 : sq + j;     // predecessor: S_1(j-1) for tile 0, S_0(j) for tile 1
            if (idx >= 0) {
                mbar_wait(bar(L::kBarSFree + (1 - t)), idx & 1);
                tc_fence_after();
            }
        };
        auto virtual_qk = [&](int j) {
            wait_s_buffer(j);
            if (elect_one_sync()) mbar_arrive_n(bar(L::kBarSFree + t), KCfg<SW>::kSoftmaxThreadsPerTile);
            __syncwarp();
        };
        auto pv = [&](int j, int it_v) {          // step j of this slot against the V tile in ring entry it_v
            const uint32_t ph = (st + j) & 1;
            // the previous item's epilogue must have read O_t out of TMEM before this item overwrites it
            if (j == 0 && ko > 0) mbar_wait(bar(L::kBarOFree + t), (ko - 1) & 1);
            FA_PROF_MARK(3);
            mbar_wait(bar(L::kBarPFull + 2 * t), ph);
            tc_fence_after();
            FA_PROF_MARK(2);             // waiting for P
            FA_TRACE_EV(p.prof, k, t, 1, j, 2);
            issue_pv_half(slot_addr(it_v), j > 0, 0, 0);
            FA_TRACE_EV(p.prof, k, t, 1, j, 3);
            FA_PROF_MARK(3);             // issue + bookkeeping
            mbar_wait(bar(L::kBarPFull + 2 * t + 1), ph);
            tc_fence_after();
            FA_PROF_MARK(2);
            FA_TRACE_EV(p.prof, k, t, 1, j, 4);
            issue_pv_half(slot_addr(it_v), j > 0, 1, empty_bar(it_v));
            FA_TRACE_EV(p.prof, k, t, 1, j, 5);
            FA_PROF_MARK(3);
        };
        auto qk = [&](int j) {
            FA_PROF_MARK(3);
            wait_s_buffer(j);
            FA_PROF_MARK(4);             // waiting for the shared S buffer
            FA_TRACE_EV(p.prof, k, t, 1, j, 0);
            issue_qk(slot_addr(it0 + 2 * j), empty_bar(it0 + 2 * j), j + 1 == nt);
            FA_TRACE_EV(p.prof, k, t, 1, j, 1);
            FA_PROF_MARK(3);
        };

        mbar_wait(bar(L::kBarQFull), kq & 1);
        ++kq;
        if (HS != 0 && w.split) {
            // Split-KV half item: both slots hold the same 128 query rows; at step s slot t multiplies with key tile
            // jm = 2s + t and only hands the other slot's tile jo = 2s + 1 - t back.  Ring entries in order: K_jm / K_jo, V_...
            // As in the regular path the score tile of step s + 1 is issued BEFORE P V of step s, so that it is ready when the
            // softmax warpgroup comes back for it (issued after, the warpgroup idled ~900 clk per step: scripts/trace_cta.py).
            const int steps = w.n_steps;
            auto split_qk = [&](int sidx) {
                if (sidx < nt) {
                    const int jm = 2 * sidx + t;
                    wait_full(it0 + 2 * jm);
                    wait_s_buffer(sidx);
                    issue_qk(slot_addr(it0 + 2 * jm), empty_bar(it0 + 2 * jm), sidx + 1 == nt);
                } else {
                    virtual_qk(sidx);
                    if (nt == 0) arrive(bar(L::kBarQEmpty));      // this slot never multiplies with Q (one key tile in all)
                }
            };
            split_qk(0);
            for (int sidx = 0; sidx < steps; ++sidx) {
                const int jm = 2 * sidx + t, jo = 2 * sidx + 1 - t;
                if (sidx + 1 < steps) split_qk(sidx + 1);
                if (jo < n) {
                    wait_full(it0 + 2 * jo);
                    arrive(empty_bar(it0 + 2 * jo));
                    wait_full(it0 + 2 * jo + 1);
                    arrive(empty_bar(it0 + 2 * jo + 1));
                }
                if (sidx < nt) {
                    wait_full(it0 + 2 * jm + 1);
                    pv(sidx, it0 + 2 * jm + 1);
                }
            }
            st += nt;
            sq += steps;
            ko += nt > 0 ? 1 : 0;
            it0 += 2 * n;
            continue;
        }
        wait_full(it0);
        FA_PROF_MARK(0);                 // Q + K0 arrival
        if (nt > 0) qk(0);
        else {
            virtual_qk(0);
            arrive(empty_bar(it0));
            arrive(bar(L::kBarQEmpty));
        }
        for (int j = 0; j < n; ++j) {
            const int it_v = it0 + 2 * j + 1, it_k = it0 + 2 * j + 2;
            const bool has_next = j + 1 < n;
            if (has_next) {
                wait_full(it_k);
                FA_PROF_MARK(1);         // waiting for K/V tiles
                if (j + 1 < nt) qk(j + 1);
                else {
                    virtual_qk(j + 1);
                    arrive(empty_bar(it_k));
                }
            }
            wait_full(it_v);
            FA_PROF_MARK(1);
            if (j < nt) pv(j, it_v); else arrive(empty_bar(it_v));
        }
        st += nt;
        sq += n;
        ko += nt > 0 ? 1 : 0;
        it0 += 2 * n;
    }
    FA_PROF_MARK(3);
    if ((threadIdx.x & 31) == 0) FA_PROF_FLUSH(p.prof, 8 + 8 * t, 6);
}

// ------------------------------------------------------------------------------------------------
// MMA issuers split by TYPE instead of by query tile (FA_ISSUER_BY_TYPE): warp kMmaWarp0 issues every Q K^T (both query
// tiles), warp kMmaWarp1 every P V.  tcgen05.mma issue blocks while the pipe's queue is full, so with one issuer per
// query tile a P V half whose P has long been ready sits behind the same warp's Q K^T (measured: ~1,000 clk from "P half
// ready" to its issue) and the score tile is interleaved with the other tile's P V; here a ready MMA of one kind never
// waits behind the issue of the other kind.  Same barriers and phases as mmaIssuerWarp; both roles see every K/V slot
// full before they hand it back (commit by the role that multiplied with it, plain arrive by the other).
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT, int SW, int HS>
__device__ __forceinline__ void mmaTypeIssuerWarp(uint32_t smem_base, uint32_t tmem_base_in, const FwdParams& p, const int role) {
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);
    using L = SmemLayout<D, STAGES>;
    constexpr uint32_t kFmt = (DT == kBF16) ? 1u : 0u;
    constexpr uint32_t idesc_qk = umma_idesc(kBlockM, kBlockN, kFmt, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc(kBlockM, D, kFmt, 0, 1);
    const uint32_t bar0 = smem_base + L::kBarOff;
    auto bar = [&](int i) { return bar0 + 8u * uint32_t(i); };
    const uint64_t desc_k_major = umma_desc_sw128(0, 16, 1024);
    const uint64_t desc_mn_major = umma_desc_sw128(0, kHalfBytes, 1024);
    const uint32_t s_tmem = tmem_base + kTmemS;

    auto slot_addr = [&](int it) { return smem_base + L::kKVOff + (it % STAGES) * L::kKVTileBytes; };
    auto empty_bar = [&](int it) { return bar(L::kBarKVEmpty + it % STAGES); };
    auto wait_full = [&](int it) {
        mbar_wait(bar(L::kBarKVFull + it % STAGES), (it / STAGES) & 1);
        tc_fence_after();
    };
    auto arrive = [&](uint32_t b) { if (elect_one_sync()) mbar_arrive(b); __syncwarp(); };

    int it0 = 0, kq = 0, sq = 0;
    int st[2] = {0, 0}, ko[2] = {0, 0};
    for (int k = 0;; ++k) {
        WorkItem w;
        const int item = fetch_item<D, STAGES>(smem_base, k, w);
        if (item < 0) break;
        const int n = w.n_kv;
        if (n <= 0) continue;
        const int nts[2] = {w.n_tile0, w.n_tile1};

        if (HS != 0 && w.split) {
            // Split-KV half item (see mmaIssuerWarp): step s, slot t <-> key tile jt = 2s + t; every ring entry is used by one
            // slot only, multiplied with by one role and handed back by both.
            const int steps = w.n_steps;
            if (role == 0) {
                mbar_wait(bar(L::kBarQFull), kq & 1);
                ++kq;
            }
            for (int sidx = 0; sidx < steps; ++sidx) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int jt = 2 * sidx + t;
                    const bool act = sidx < nts[t];
                    if (role == 0) {
                        if (act) wait_full(it0 + 2 * jt);
                        const int idx = (t == 0) ? sq + sidx - 1 : sq + sidx;
                        if (idx >= 0) {
                            mbar_wait(bar(L::kBarSFree + (1 - t)), idx & 1);
                            tc_fence_after();
                        }
                        if (elect_one_sync()) {
                            if (act) {
                                const uint64_t a0 = desc_k_major + ((smem_base + L::kQOff + t * L::kQTileBytes) >> 4);
                                const uint64_t b0 = desc_k_major + (slot_addr(it0 + 2 * jt) >> 4);
#pragma unroll
                                for (int ks = 0; ks < D / 16; ++ks) {
                                    const uint32_t off = ((ks / 4) * kHalfBytes + (ks % 4) * 32) >> 4;
                                    umma_ss(s_tmem, a0 + off, b0 + off, idesc_qk, ks > 0);
                                }
                                tc_commit(bar(L::kBarSFull + t));
                                tc_commit(empty_bar(it0 + 2 * jt));
                            } else {
                                mbar_arrive_n(bar(L::kBarSFree + t), KCfg<SW>::kSoftmaxThreadsPerTile);
                            }
                            if (sidx + 1 == steps && t == 1) tc_commit(bar(L::kBarQEmpty));    // every Q K^T of the item has been issued
                        }
                        __syncwarp();
                        if (act) {
                            wait_full(it0 + 2 * jt + 1);
                            arrive(empty_bar(it0 + 2 * jt + 1));
                        }
                    } else if (act) {
                        wait_full(it0 + 2 * jt);
                        arrive(empty_bar(it0 + 2 * jt));
                        wait_full(it0 + 2 * jt + 1);
                        const uint64_t b0 = desc_mn_major + (slot_addr(it0 + 2 * jt + 1) >> 4);
                        const uint32_t ph = (st[t] + sidx) & 1;
                        const uint32_t p_tmem = tmem_base + tmem_p_col(t);
                        const uint32_t o_tmem = tmem_base + kTmemO0 + 128u * t;
                        if (sidx == 0 && ko[t] > 0) mbar_wait(bar(L::kBarOFree + t), (ko[t] - 1) & 1);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            mbar_wait(bar(L::kBarPFull + 2 * t + half), ph);
                            tc_fence_after();
                            if (elect_one_sync()) {
#pragma unroll
                                for (int kk = 0; kk < kBlockN / 32; ++kk) {
                                    const int ks = half * (kBlockN / 32) + kk;
                                    umma_ts(o_tmem, p_tmem + 8u * ks, b0 + ((ks * 2048) >> 4), idesc_pv, (sidx > 0 || ks > 0) ? 1u : 0u);
                                }
                                if (half == 1) {
                                    tc_commit(bar(L::kBarOFull + t));
                                    tc_commit(empty_bar(it0 + 2 * jt + 1));
                                } else if (kSplitOFull<D>) {
                                    tc_commit(bar(L::kBarOHalf + t));
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                st[t] += nts[t];
                ko[t] += nts[t] > 0 ? 1 : 0;
            }
            sq += steps;
            it0 += 2 * n;
            continue;
        }

        if (role == 0) {
            // ---- every Q_t K_j^T, in the order of the shared S buffer: S_0(0), S_1(0), S_0(1), ...
            mbar_wait(bar(L::kBarQFull), kq & 1);
            ++kq;
            for (int j = 0; j < n; ++j) {
                wait_full(it0 + 2 * j);
                const uint32_t k_smem = slot_addr(it0 + 2 * j);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int idx = (t == 0) ? sq + j - 1 : sq + j;     // predecessor in the buffer: S_1(j-1) / S_0(j)
                    if (idx >= 0) {
                        mbar_wait(bar(L::kBarSFree + (1 - t)), idx & 1);
                        tc_fence_after();
                    }
                    FA_TRACE_EV(p.prof, k, t, 1, j, 0);
                    if (elect_one_sync()) {
                        if (j < nts[t]) {
                            const uint64_t a0 = desc_k_major + ((smem_base + L::kQOff + t * L::kQTileBytes) >> 4);
                            const uint64_t b0 = desc_k_major + (k_smem >> 4);
#pragma unroll
                            for (int ks = 0; ks < D / 16; ++ks) {
                                const uint32_t off = ((ks / 4) * kHalfBytes + (ks % 4) * 32) >> 4;
                                umma_ss(s_tmem, a0 + off, b0 + off, idesc_qk, ks > 0);
                            }
                            tc_commit(bar(L::kBarSFull + t));
                        } else {
                            mbar_arrive_n(bar(L::kBarSFree + t), KCfg<SW>::kSoftmaxThreadsPerTile);   // virtual step: pass the buffer on
                        }
                    }
                    __syncwarp();
                    FA_TRACE_EV(p.prof, k, t, 1, j, 1);
                }
                // K_j and (after the item's last step) the Q tiles go back once every Q K^T issued so far has retired
                if (elect_one_sync()) {
                    tc_commit(empty_bar(it0 + 2 * j));
                    if (j + 1 == n) tc_commit(bar(L::kBarQEmpty));
                }
                __syncwarp();
                wait_full(it0 + 2 * j + 1);                  // V_j: not ours, but the slot's phases must be seen in order
                arrive(empty_bar(it0 + 2 * j + 1));
            }
        } else {
            // ---- every P_t V_j, in the order the P halves arrive: P_0 first half, P_0 second half, P_1 first, P_1 second
            for (int j = 0; j < n; ++j) {
                wait_full(it0 + 2 * j);                      // K_j: not ours
                arrive(empty_bar(it0 + 2 * j));
                wait_full(it0 + 2 * j + 1);
                const uint64_t b0 = desc_mn_major + (slot_addr(it0 + 2 * j + 1) >> 4);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (j >= nts[t]) continue;
                    const uint32_t ph = (st[t] + j) & 1;
                    const uint32_t p_tmem = tmem_base + tmem_p_col(t);
                    const uint32_t o_tmem = tmem_base + kTmemO0 + 128u * t;
                    if (j == 0 && ko[t] > 0) mbar_wait(bar(L::kBarOFree + t), (ko[t] - 1) & 1);   // previous item's epilogue has read O_t
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        mbar_wait(bar(L::kBarPFull + 2 * t + half), ph);
                        tc_fence_after();
                        FA_TRACE_EV(p.prof, k, t, 1, j, 2 + 2 * half);
                        if (elect_one_sync()) {
#pragma unroll
                            for (int kk = 0; kk < kBlockN / 32; ++kk) {
                                const int ks = half * (kBlockN / 32) + kk;
                                umma_ts(o_tmem, p_tmem + 8u * ks, b0 + ((ks * 2048) >> 4), idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
                            }
                            if (half == 1) tc_commit(bar(L::kBarOFull + t));
                            else if (kSplitOFull<D>) tc_commit(bar(L::kBarOHalf + t));
                        }
                        __syncwarp();
                        FA_TRACE_EV(p.prof, k, t, 1, j, 3 + 2 * half);
                    }
                }
                if (elect_one_sync()) tc_commit(empty_bar(it0 + 2 * j + 1));   // V_j back once every P V issued so far has retired
                __syncwarp();
            }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            st[t] += nts[t];
            ko[t] += nts[t] > 0 ? 1 : 0;
        }
        sq += n;
        it0 += 2 * n;
    }
}

// ------------------------------------------------------------------------------------------------
// MMA issuers of a CTA pair (cta_group::2, d = 128, split by type like mmaTypeIssuerWarp).  They run in the LEADER CTA only:
// one instruction multiplies the 256-row tile (128 rows of Q / P from each CTA) with a K / V tile of which each CTA's shared
// memory holds half, and writes each CTA's 128 rows of S / O into its own TMEM.  Completion barriers that both CTAs wait on
// (s_full, o_half, o_full, the ring's empty barriers, q_empty) get multicast commits; barriers the issuers wait on (q_full,
// kv_full, s_free, p_full, o_free) live in the leader and collect arrivals from both CTAs — one per softmax WARP (8 in all)
// instead of one per thread, so that the peer's arrivals are 4 remote operations per hand-off, not 128.
// With an even number of ring slots K tiles always sit in even slots and V tiles in odd ones: each issuer only ever touches
// its own slots.
// ------------------------------------------------------------------------------------------------
constexpr int kPairWarpArrivals = 2 * 4;      // softmax warps per query tile in the pair

template <int D, int STAGES, int DT>
__device__ __forceinline__ void mmaPairIssuerWarp(uint32_t smem_base, uint32_t tmem_base_in, const FwdParams& p, const int role) {
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);
    using L = SmemLayout<D, STAGES, 2>;
    static_assert(D == 128 && STAGES % 2 == 0, "pair kernel: d = 128, even ring");
    constexpr uint32_t kFmt = (DT == kBF16) ? 1u : 0u;
    constexpr uint32_t idesc_qk = umma_idesc(2 * kBlockM, kBlockN, kFmt, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc(2 * kBlockM, D, kFmt, 0, 1);
    constexpr int kKHalfBytes = (kBlockN / 2) * 128;
    const uint32_t bar0 = smem_base + L::kBarOff;
    auto bar = [&](int i) { return bar0 + 8u * uint32_t(i); };
    const uint64_t desc_k_major = umma_desc_sw128(0, 16, 1024);
    const uint64_t desc_mn_major = umma_desc_sw128(0, kHalfBytes, 1024);
    const uint32_t s_tmem = tmem_base + kTmemS;
    auto slot_addr = [&](int it) { return smem_base + L::kKVOff + (it % STAGES) * L::kKVTileBytes; };
    auto empty_bar = [&](int it) { return bar(L::kBarKVEmpty + it % STAGES); };
    auto wait_full = [&](int it) {
        mbar_wait(bar(L::kBarKVFull + it % STAGES), (it / STAGES) & 1);
        tc_fence_after();
    };

    FA_T2_DECL;
    int it0 = 0, kq = 0, sq = 0;
    int st[2] = {0, 0}, ko[2] = {0, 0};
    for (int k = 0;; ++k) {
        WorkItem w;
        const int item = fetch_item<D, STAGES, 2>(smem_base, k, w);
        if (item < 0) break;
        const int n = w.n_kv;
        if (n <= 0) continue;
        const int nts[2] = {w.n_tile0, w.n_tile1};
        if (role == 0) {
            mbar_wait(bar(L::kBarQFull), kq & 1);
            ++kq;
            for (int j = 0; j < n; ++j) {
                wait_full(it0 + 2 * j);
                const uint32_t k_smem = slot_addr(it0 + 2 * j);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int idx = (t == 0) ? sq + j - 1 : sq + j;
                    if (idx >= 0) {
                        mbar_wait(bar(L::kBarSFree + (1 - t)), idx & 1);
                        tc_fence_after();
                    }
                    FA_T2(p.prof, 2, 70 + t);      // S buffer free for Q_t K_j^T
                    if (elect_one_sync()) {
                        if (j < nts[t]) {
                            const uint64_t a0 = desc_k_major + ((smem_base + L::kQOff + t * L::kQTileBytes) >> 4);
                            const uint64_t b0 = desc_k_major + (k_smem >> 4);
#pragma unroll
                            for (int ks = 0; ks < D / 16; ++ks) {
                                const uint32_t off_a = ((ks / 4) * kHalfBytes + (ks % 4) * 32) >> 4;
                                const uint32_t off_b = ((ks / 4) * kKHalfBytes + (ks % 4) * 32) >> 4;
                                umma_ss_pair(s_tmem, a0 + off_a, b0 + off_b, idesc_qk, ks > 0);
                            }
                            tc_commit_pair(bar(L::kBarSFull + t));
                        } else {
                            mbar_arrive_n(bar(L::kBarSFree + t), kPairWarpArrivals);      // virtual step: pass the buffer on
                        }
                    }
                    __syncwarp();
                    FA_T2(p.prof, 2, 6 + t);       // Q_t K_j^T issued
                }
                if (elect_one_sync()) {
                    tc_commit_pair(empty_bar(it0 + 2 * j));
                    if (j + 1 == n) tc_commit_pair(bar(L::kBarQEmpty));
                }
                __syncwarp();
            }
        } else {
            for (int j = 0; j < n; ++j) {
                wait_full(it0 + 2 * j + 1);
                const uint64_t b0 = desc_mn_major + (slot_addr(it0 + 2 * j + 1) >> 4);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (j >= nts[t]) continue;
                    const uint32_t ph = (st[t] + j) & 1;
                    const uint32_t p_tmem = tmem_base + tmem_p_col(t);
                    const uint32_t o_tmem = tmem_base + kTmemO0 + 128u * t;
                    if (j == 0 && ko[t] > 0) mbar_wait(bar(L::kBarOFree + t), (ko[t] - 1) & 1);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        mbar_wait(bar(L::kBarPFull + 2 * t + half), ph);
                        tc_fence_after();
                        FA_T2(p.prof, 3, 74 + 2 * t + half);      // P half ready
                        if (elect_one_sync()) {
#pragma unroll
                            for (int kk = 0; kk < kBlockN / 32; ++kk) {
                                const int ks = half * (kBlockN / 32) + kk;
                                umma_ts_pair(o_tmem, p_tmem + 8u * ks, b0 + ((ks * 2048) >> 4), idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
                            }
                            if (half == 1) tc_commit_pair(bar(L::kBarOFull + t));
                            else if (kSplitOFull<D>) tc_commit_pair(bar(L::kBarOHalf + t));
                        }
                        __syncwarp();
                        FA_T2(p.prof, 3, 78 + 2 * t + half);      // P V half issued
                    }
                }
                if (elect_one_sync()) tc_commit_pair(empty_bar(it0 + 2 * j + 1));
                __syncwarp();
            }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            st[t] += nts[t];
            ko[t] += nts[t] > 0 ? 1 : 0;
        }
        sq += n;
        it0 += 2 * n;
    }
}

// ------------------------------------------------------------------------------------------------
// Softmax warpgroup for query tile t (128 threads, one score row each), including the lazy O rescale and the
// epilogue (O/l -> global, optional LSE).  Persistent: loops over the published work items; the epilogue of one item
// overlaps the next item's first Q K^T.
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT, bool OVEC32, int EMU, int ST, int HS, int CG = 1>
__device__ __forceinline__ void softmaxWarpgroup(uint32_t smem_base, uint32_t tmem_base, const FwdParams& p, int t, const CUtensorMap* tmO) {
    using L = SmemLayout<D, STAGES, CG>;
    static_assert(CG == 1 || (CG == 2 && HS == 0), "the pair kernel has no half items");
    uint32_t bar0 = smem_base + L::kBarOff;
    asm volatile("" : "+r"(bar0));     // keep in a register (see below)
    // CTA pair: the barriers this warpgroup ARRIVES on live in the leader CTA (distance `to_lead` in the shared::cluster
    // window; 0 in the leader itself) and take one arrival per warp
    uint32_t to_lead = 0;
    if constexpr (CG == 2) to_lead = mapa_shared(bar0, 0) - bar0;
    auto sm_arrive = [&](uint32_t b) {
        if constexpr (CG == 2) {
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive_cluster(b + to_lead);
        } else {
            mbar_arrive(b);
        }
    };
    const uint32_t s_full = bar0 + 8u * (L::kBarSFull + t);
    const uint32_t p_full0 = bar0 + 8u * (L::kBarPFull + 2 * t);
    const uint32_t p_full1 = p_full0 + 8u;
    const uint32_t o_full = bar0 + 8u * (L::kBarOFull + t);
    const uint32_t o_free = bar0 + 8u * (L::kBarOFree + t);
    const uint32_t s_free = bar0 + 8u * (L::kBarSFree + t);
    const uint32_t o_half = bar0 + 8u * (L::kBarOHalf + t);

    const int warp_in_wg = (threadIdx.x / 32) & 3;
    const int lane = threadIdx.x & 31;
    const uint32_t lane_base = uint32_t(warp_in_wg * 32) << 16;
    uint32_t tS = tmem_base + lane_base + kTmemS;
    uint32_t tP = tmem_base + lane_base + tmem_p_col(t);
    uint32_t tO = tmem_base + lane_base + kTmemO0 + 128u * t;
    const float c = p.scale_log2;
    // Keep the loop's addresses in registers: left alone, the compiler re-derives them from %tid / the shared-window
    // base (S2R, ~25 clk each) at the top of every pass, right on the path between two score tiles.
    asm volatile("" : "+r"(tS), "+r"(tP), "+r"(tO));

    FA_PROF_DECL(6);
    FA_T2_DECL;
    int st = 0;      // key tiles this warpgroup has processed so far (barrier phase bookkeeping)
    for (int k = 0;; ++k) {
        if (k > 0 && warp_in_wg == 0) FA_T2(p.prof, 4 + t, 40 + t);      // the previous item's epilogue is done
        WorkItem w;
        const int item = fetch_item<D, STAGES, CG>(smem_base, k, w);
        if (item < 0) break;
        const int n = w.n_tile(t);
        // a half item has no rows for query-tile slot 1: n == 0 (no barrier traffic) and its row numbers lie past every
        // sequence, so the epilogue below writes nothing for it
        // (Split-KV half item, HS = 1 kernels: both slots hold rows [q0, q0 + 128), slot t takes key tiles t, t+2, ... and slot 0
        // writes the merged result.  The launcher only plans such items when no step needs a mask — non-causal, Nk a multiple
        // of 128 — so the key loop below is untouched: the 216-register loop has no room for a second key-tile numbering.)
        // (CTA pair: w.q0 / w.h are this CTA's own first row / head of the item; its tile t starts rows / 2 further — 256 when the
        // pair cuts a 512-row block by rows, 128 when it cuts by two heads — or, cut by four heads, is the same rows of head h + t)
        if constexpr (CG == 2) w.h += t * w.hstep;
        const int tile_row0 = CG == 2 ? w.q0 + (w.hstep ? 0 : t * (w.rows >> 1)) : ((t * kBlockM < w.rows) ? w.q0 + t * kBlockM : kNoRow);
        const int row = tile_row0 + warp_in_wg * 32 + lane;

        float m_run = -INFINITY;   // max in use, in raw (unscaled) score units
        float l_run = 0.f;

        for (int j = 0; j < n; ++j) {
            const uint32_t ph = (st + j) & 1;
            FA_PROF_MARK(5);             // loop overhead / l update / epilogue
            FA_TRACE_EV(p.prof, k, t, 0, j, 5);
            mbar_wait(s_full, ph);
            tc_fence_after();
            FA_PROF_MARK(0);             // waiting for S
            FA_TRACE_EV(p.prof, k, t, 0, j, 0);

            uint32_t r[kBlockN];
            const int kv0 = j * kBlockN;
            const bool need_mask = (kv0 + kBlockN > p.Nk) || (p.causal && (kv0 + kBlockN - 1 > tile_row0 + p.causal_off));
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
            // first half of the row, then the second half in flight while the first half's row max is computed (-0.4 % cycles)
            tmem_ld32(tS, r);
            tmem_ld32(tS + 32u, r + 32);
            tc_wait_ld();
            tmem_ld32(tS + 64u, r + 64);
            tmem_ld32(tS + 96u, r + 96);
            if (!need_mask) {
#pragma unroll
                for (int cc = 0; cc < kBlockN / 2; cc += 8) {
                    mx0 = max3(mx0, __uint_as_float(r[cc + 0]), __uint_as_float(r[cc + 1]));
                    mx1 = max3(mx1, __uint_as_float(r[cc + 2]), __uint_as_float(r[cc + 3]));
                    mx2 = max3(mx2, __uint_as_float(r[cc + 4]), __uint_as_float(r[cc + 5]));
                    mx3 = max3(mx3, __uint_as_float(r[cc + 6]), __uint_as_float(r[cc + 7]));
                }
            }
            tc_wait_ld();
            tc_fence_before();
            sm_arrive(s_free);           // the score row is in registers: the shared S buffer may be overwritten
            if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 10 + t);
            FA_TRACE_EV(p.prof, k, t, 0, j, 1);
#ifdef FA_PHASE_PROFILE
            // lag between the two warpgroups: clocks since the OTHER warpgroup last took an S tile (warp 0 of each reports)
            if ((threadIdx.x & 127) == 0) {
                uint32_t now, other;
                asm volatile("mov.u32 %0, %%clock;" : "=r"(now));
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(other) : "r"(smem_base + L::kTmemPtrOff + 8u + 4u * (1 - t)) : "memory");
                asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_base + L::kTmemPtrOff + 8u + 4u * t), "r"(now) : "memory");
                if (p.prof && j > 0) atomicAdd(p.prof + 24 + t, (unsigned long long)(now - other));
            }
#endif
            FA_PROF_MARK(1);             // tcgen05.ld of the score row

            // (A warp-uniform fast path for the exactly diagonal tile — compares only in the one 32-column quarter that is
            // lane-dependent, -inf fills for the quarters above it — was measured in round 2: the third code path over the
            // 128-register score row costs the COMMON path 4 % at N = 8K and 14 % at 1K through register allocation, with or
            // without skipping the masked quarters' exponentials.  Not kept; profiles/r2_cycles_diag_fastpath.jsonl.)
            if (need_mask) {
                const int lim_c = p.causal ? (row + p.causal_off) : 0x7fffffff;
                const int lim = min(lim_c, p.Nk - 1) - kv0;   // columns c > lim are masked
#pragma unroll
                for (int cc = 0; cc < kBlockN; ++cc) r[cc] = mask_gt(r[cc], cc, lim);   // -inf where cc > lim
            }

            if (need_mask) {             // the first half's maxima were not taken before the mask was applied
#pragma unroll
                for (int cc = 0; cc < kBlockN / 2; cc += 8) {
                    mx0 = max3(mx0, __uint_as_float(r[cc + 0]), __uint_as_float(r[cc + 1]));
                    mx1 = max3(mx1, __uint_as_float(r[cc + 2]), __uint_as_float(r[cc + 3]));
                    mx2 = max3(mx2, __uint_as_float(r[cc + 4]), __uint_as_float(r[cc + 5]));
                    mx3 = max3(mx3, __uint_as_float(r[cc + 6]), __uint_as_float(r[cc + 7]));
                }
            }
#pragma unroll
            for (int cc = kBlockN / 2; cc < kBlockN; cc += 8) {
                mx0 = max3(mx0, __uint_as_float(r[cc + 0]), __uint_as_float(r[cc + 1]));
                mx1 = max3(mx1, __uint_as_float(r[cc + 2]), __uint_as_float(r[cc + 3]));
                mx2 = max3(mx2, __uint_as_float(r[cc + 4]), __uint_as_float(r[cc + 5]));
                mx3 = max3(mx3, __uint_as_float(r[cc + 6]), __uint_as_float(r[cc + 7]));
            }
            const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(fmaxf(mx2, mx3), m_run));

            if (j == 0) {
                m_run = m_new;
            } else {
                // lazy rescale: (m_new - m_run) is NaN when both are -inf -> compares false
                const bool grow = (m_new - m_run) * c > kRescaleThreshold;
                if (__any_sync(0xffffffffu, grow)) {
                    const float f = grow ? ex2_approx((m_run - m_new) * c) : 1.0f;
                    mbar_wait(o_full, ph ^ 1);      // P V of the previous key tile has retired
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < D / 32; ++q) {
                        uint32_t o[32];
                        tmem_ld32(tO + 32u * q, o);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                        tmem_st32(tO + 32u * q, o);
                    }
                    l_run *= f;
                    if (grow) m_run = m_new;
                }
            }
            FA_PROF_MARK(2);             // mask + row max + (rare) O rescale
            const float m_safe = (m_run == -INFINITY) ? 0.f : m_run;
            const float2 c2 = make_float2(c, c);
            const float2 nm2 = make_float2(-m_safe * c, -m_safe * c);

            float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
            // exp2 of one score pair: MUFU.EX2 for most pairs, FMA-pipe emulation for EMU of every 8
            auto exp_pair = [&](int col) -> float2 {
                float2 x = fma2(make_float2(__uint_as_float(r[col]), __uint_as_float(r[col + 1])), c2, nm2);
                if (((col / 2) % 8) * 3 % 8 < EMU) {     // spread the emulated pairs evenly over the group of 8
                    x = ex2_emu2(x);
                } else {
                    x.x = ex2_approx(x.x);
                    x.y = ex2_approx(x.y);
                }
                return x;
            };
            // 32 scores -> 16 packed columns
            auto exp_quarter = [&](int qt, uint32_t* pk) {
#pragma unroll
                for (int cc = 0; cc < 32; cc += 4) {
                    const float2 x0 = exp_pair(32 * qt + cc), x1 = exp_pair(32 * qt + cc + 2);
                    s0 = add2(s0, x0);
                    s1 = add2(s1, x1);
                    pk[cc / 2 + 0] = pack16<DT>(x0.x, x0.y);
                    pk[cc / 2 + 1] = pack16<DT>(x1.x, x1.y);
                }
            };
            // P_t V_{j-1} must have retired before the P columns are overwritten.  Checked here, before the exponentials,
            // rather than right before the first store: measured 3-6 % faster (a warpgroup that has to wait leaves the
            // MUFU unit to the other one, and the stores stay where the scheduler wants them).
            if (j > 0) {
                mbar_wait(kSplitOFull<D> ? o_half : o_full, ph ^ 1);
                tc_fence_after();
            }
            FA_TRACE_EV(p.prof, k, t, 0, j, 2);
            {
                uint32_t pk[32];
                exp_quarter(0, pk);
                exp_quarter(1, pk + 16);
                tmem_st32(tP, pk);                 // keys 0..63 of P
            }
            {
                uint32_t pk[32];
                exp_quarter(2, pk);
                tc_wait_st();                      // first half landed while quarter 2 was computed
                tc_fence_before();
                sm_arrive(p_full0);                // MMA may start P V on keys 0..63
                FA_TRACE_EV(p.prof, k, t, 0, j, 3);
                exp_quarter(3, pk + 16);
                if (kSplitOFull<D> && j > 0) {
                    mbar_wait(o_full, ph ^ 1);
                    tc_fence_after();
                }
                tmem_st32(tP + 32u, pk);           // keys 64..127 of P
            }
            FA_PROF_MARK(3);             // exp2 / pack / tcgen05.st issue
            tc_wait_st();
            tc_fence_before();
            sm_arrive(p_full1);
            if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 20 + t);
            FA_TRACE_EV(p.prof, k, t, 0, j, 4);
            FA_PROF_MARK(4);             // store drain + arrive

            l_run += (s0.x + s0.y) + (s1.x + s1.y);
        }

        if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 30 + t);
        if (HS != 0 && w.split) {
            // ---- split-KV half item: slot 1 hands (max, sum) to slot 0 and leaves; slot 0 merges both partial results ----
            //   m = max(m0, m1);  l = l0 2^((m0-m)c) + l1 2^((m1-m)c);  O = (O_0 2^((m0-m)c) + O_1 2^((m1-m)c)) / l
            // O_1 lives in the same TMEM lanes as O_0 (128 columns further), so slot 0's thread for a row reads both itself;
            // only the two scalars cross through shared memory (double-buffered by the CTA's item counter).
            const uint32_t xaddr = smem_base + L::kExchOff + uint32_t((k & 1) * kBlockM + warp_in_wg * 32 + lane) * 8u;
            if (n > 0) {
                mbar_wait(o_full, (st + n - 1) & 1);      // this slot's last P V has retired
                tc_fence_after();
            }
            if (t == 1) {
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xaddr), "f"(m_run), "f"(l_run) : "memory");
                tc_fence_before();
                named_bar_sync(3u, 256u);
                st += n;
                continue;
            }
            named_bar_sync(3u, 256u);
            tc_fence_after();
            float m1, l1;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m1), "=f"(l1) : "r"(xaddr) : "memory");
            const int n1 = w.n_tile1;
            const float m = fmaxf(m_run, m1);
            const float a0 = (m_run == -INFINITY) ? 0.f : ex2_approx((m_run - m) * c);
            const float a1 = (m1 == -INFINITY || n1 == 0) ? 0.f : ex2_approx((m1 - m) * c);
            const float l = l_run * a0 + l1 * a1;
            const float inv = (l > 0.f) ? 1.0f / l : 0.f;
            const float w0 = a0 * inv, w1 = a1 * inv;
            const int orow_i = w.q0 + warp_in_wg * 32 + lane;
            const bool row_ok = orow_i < p.Nq;
            uint16_t* orow = reinterpret_cast<uint16_t*>(p.O) + (long long)w.b * p.o_stride_b + (long long)w.h * p.o_stride_h +
                             (long long)orow_i * p.o_stride_n;
#pragma unroll 1
            for (int q = 0; q < D / 16; ++q) {            // 16 columns = one 32-byte sector at a time (rare path: keep it small)
                uint32_t o0[16], o1[16];
                tmem_ld16(tO + 16u * q, o0);
                tmem_ld16(tO + 128u + 16u * q, o1);       // with n1 == 0 the columns hold stale data; w1 == 0 discards it below
                tc_wait_ld();
                uint32_t hh[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float x = __uint_as_float(o0[2 * i]) * w0, y = __uint_as_float(o0[2 * i + 1]) * w0;
                    if (n1 > 0) {
                        x = fmaf(__uint_as_float(o1[2 * i]), w1, x);
                        y = fmaf(__uint_as_float(o1[2 * i + 1]), w1, y);
                    }
                    hh[i] = pack16<DT>(x, y);
                }
                if (row_ok) {
                    if constexpr (OVEC32) {
                        st_global_v8(orow + 16 * q, hh);
                    } else {
                        st_global_v4(orow + 16 * q, hh[0], hh[1], hh[2], hh[3]);
                        st_global_v4(orow + 16 * q + 8, hh[4], hh[5], hh[6], hh[7]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(o_free);                                         // both O tiles may be overwritten by the next item
            if (n1 > 0) mbar_arrive(bar0 + 8u * (L::kBarOFree + 1));
            if (p.lse != nullptr && row_ok)
                p.lse[((long long)w.b * p.Hq + w.h) * p.Nq + orow_i] = (l > 0.f) ? (m * p.scale + logf(l)) : -INFINITY;
            st += n;
            continue;
        }
        // ---- epilogue ----
        const bool row_ok = row < p.Nq;
        const long long row_lin = ((long long)w.b * p.Hq + w.h) * p.Nq + row;
        const float m_fin = (m_run == -INFINITY) ? 0.f : m_run;
        const float lse_part = (l_run > 0.f) ? (m_fin * p.scale + logf(l_run)) : -INFINITY;
        if (p.acc_o == nullptr) {
            // plain mode: O / l -> global in the I/O dtype
            const float inv_l = (l_run > 0.f) ? 1.0f / l_run : 0.f;
            uint16_t* orow = reinterpret_cast<uint16_t*>(p.O) + (long long)w.b * p.o_stride_b + (long long)w.h * p.o_stride_h +
                             (long long)row * p.o_stride_n;
            if (ST != 0 && n > 0) {
                // Staged epilogue, one warp at a time: every warp packs its 32 rows x 64 columns of O / l, writes them into
                // ITS 4 KiB slice of the tile's staging piece (128B-swizzled, the layout of a TMA box of 64 columns x 32 rows)
                // and lane 0 issues the TMA store of that slice — only warp-level synchronisation, no CTA barrier.  Rows
                // past Nq are clipped by the tensor map.  Why: a row-per-lane st.global touches 32 different lines per
                // instruction, 1,024 sector writes per tile, which the SM's load/store path accepts at ~2 clk each: the
                // softmax warpgroup stood 2,200-2,900 clk in its epilogue at every item boundary (scripts/trace_cta.py).
                // The first version of this path (one 16 KiB piece per tile, CTA-level named barriers, one TMA store per 64
                // columns) needed 2,300-2,750 clk at d = 128: ~1,000 of them waiting for the first half's store to finish
                // reading the piece.  Here the second half is loaded and packed BEFORE that wait, and a 4 KiB store drains fast.
                mbar_wait(o_full, (st + n - 1) & 1);
                if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 35);      // the last P V has retired
                tc_fence_after();
                const uint32_t stg = smem_base + L::kStageOff + uint32_t(t) * L::kStageTileBytes;
                const int r_in = warp_in_wg * 32 + lane;
                const uint32_t srow = stg + uint32_t(r_in) * 128u;
                const uint32_t slice = stg + uint32_t(warp_in_wg) * (32u * 128u);
                // TMEM loads run one 32-column chunk ahead: the load of the next chunk is in flight while this one is packed,
                // written to the slice, fenced and handed to TMA.
                constexpr int kChunks = D / 32;
                uint32_t o[32];
                tmem_ld32(tO, o);
#pragma unroll
                for (int hf = 0; hf < D / kHalfCols; ++hf) {
                    uint32_t pkd[32];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            pkd[16 * q + i] = pack16<DT>(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
                        if (2 * hf + q + 1 < kChunks) tmem_ld32(tO + 32u * uint32_t(2 * hf + q + 1), o);
                    }
                    if (hf == D / kHalfCols - 1) {                    // O is out of TMEM: the next item's first P V may overwrite it
                        tc_fence_before();
                        sm_arrive(o_free);
                    }
                    if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 36);      // half packed in registers
                    if (lane == 0) bulk_wait_group_read0();           // this warp's previous store has read the slice
                    __syncwarp();
                    if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 37);      // slice free
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx)              // 16-byte chunk cidx of the row, XOR-swizzled by row % 8
                        st_shared_v4(srow + (uint32_t(cidx ^ (r_in & 7)) << 4), pkd[4 * cidx], pkd[4 * cidx + 1], pkd[4 * cidx + 2], pkd[4 * cidx + 3]);
                    if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 38);      // slice written
                    fence_proxy_async_shared();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(tmO, slice, hf * kHalfCols, tile_row0 + warp_in_wg * 32, w.h, w.b);
                        bulk_commit_group();
                    }
                    if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 39);      // slice handed to TMA
                }
            } else if (n > 0) {
                mbar_wait(o_full, (st + n - 1) & 1);
                if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 35);      // the last P V has retired
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < D / 32; ++q) {
                    uint32_t o[32];
                    tmem_ld32(tO + 32u * q, o);
                    tc_wait_ld();
                    if (warp_in_wg == 0) FA_T2(p.prof, 4 + t, 36);      // chunk in registers
                    uint32_t h[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        h[i] = pack16<DT>(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
                    // Every lane writes its own row (rows are o_stride_n apart), so a store instruction touches 32 different
                    // lines whatever its width: 256-bit stores (one full 32-byte sector per lane) halve the number of such
                    // instructions; measured -0.9 % cycles at N = 8K, -3 % at N <= 2K against 128-bit stores.  The choice is a
                    // template parameter: as a run-time branch on a kernel parameter it gave most of that back at N <= 2K.
                    if (row_ok) {
                        if constexpr (OVEC32) {
#pragma unroll
                            for (int i = 0; i < 2; ++i) st_global_v8(orow + 32 * q + 16 * i, h + 8 * i);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                st_global_v4(orow + 32 * q + 8 * i, h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                        }
                    }
                }
                tc_fence_before();
                sm_arrive(o_free);           // O columns may be overwritten by the next item's first P V
            } else if (row_ok) {
#pragma unroll
                for (int i = 0; i < D / 8; ++i) st_global_v4(orow + 8 * i, 0u, 0u, 0u, 0u);
            }
            if (p.lse != nullptr && row_ok) p.lse[row_lin] = lse_part;
        } else {
            // carry mode (ring-KV step): fold this call's partial result over its key range into the fp32 running
            // (output, log-sum-exp) pair:  lse' = log(e^lse_acc + e^lse_part),  O' = O_acc e^(lse_acc-lse') + (O/l) e^(lse_part-lse')
            float wa = 1.f, wp = 0.f, lse_new = -INFINITY;
            const long long acc_lin = ((long long)w.b * p.Hq + w.h) * p.acc_rows + p.acc_off + row;
            if (row_ok) {
                const float lse_acc = p.acc_lse[acc_lin];
                const float mx = fmaxf(lse_acc, lse_part);
                lse_new = mx;
                if (mx != -INFINITY) {
                    const float ea = expf(lse_acc - mx), ep = expf(lse_part - mx);
                    const float inv = 1.0f / (ea + ep);
                    wa = ea * inv;
                    wp = (l_run > 0.f) ? ep * inv / l_run : 0.f;
                    lse_new = mx + logf(ea + ep);
                }
            }
            if (n > 0) {
                mbar_wait(o_full, (st + n - 1) & 1);
                tc_fence_after();
                float* arow = p.acc_o + acc_lin * D;
#pragma unroll
                for (int q = 0; q < D / 32; ++q) {
                    uint32_t o[32];
                    tmem_ld32(tO + 32u * q, o);
                    tc_wait_ld();
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 a = ld_global_f4(arow + 32 * q + 4 * i);
                            a.x = a.x * wa + __uint_as_float(o[4 * i + 0]) * wp;
                            a.y = a.y * wa + __uint_as_float(o[4 * i + 1]) * wp;
                            a.z = a.z * wa + __uint_as_float(o[4 * i + 2]) * wp;
                            a.w = a.w * wa + __uint_as_float(o[4 * i + 3]) * wp;
                            st_global_f4(arow + 32 * q + 4 * i, a);
                        }
                    }
                }
                tc_fence_before();
                sm_arrive(o_free);
            }
            if (row_ok && n > 0) p.acc_lse[acc_lin] = lse_new;
        }
        st += n;
    }
    if ((threadIdx.x & 31) == 0) FA_PROF_FLUSH(p.prof, 0, 6);
    if (ST != 0 && lane == 0) bulk_wait_group0();     // this warp's TMA stores have landed before the CTA exits
    tc_fence_before();
}

// ------------------------------------------------------------------------------------------------
// Softmax, 16-warp layout (KCfg<D>::kRows16).  Warp (t, h, q) serves query tile t, TMEM lane quarter q (= its SM
// sub-partition), rows [32q + 16h, 32q + 16h + 16) of the tile: the two warps of a quarter split its rows, so nothing but the
// epilogue's (1/l, lse) pair ever crosses between warps.  Scores are read with tcgen05.ld 16x256b — thread (quad rr = lane/4,
// c4 = lane%4) holds rows rr and rr + 8 of the warp's 16, columns 8g + 2 c4 + {0,1} of every 8-column group g: 2 x 32 scores
// per thread and tile instead of 1 x 128 — so a row max costs two quad shuffles, the row sum stays a per-thread partial until
// the epilogue, and the packed P pairs of a thread are exactly one tcgen05.st 16x128b fragment.  Each SM sub-partition then
// runs FOUR softmax warps (two per query tile): while one sits in a MUFU.EX2 (8 clk, during which a warp issues nothing else)
// the other three issue their FFMA2 / FADD2 / F2FP / FMNMX3 / TMEM traffic.  Barriers, phases and the hand-offs with the MMA
// issuers are those of softmaxWarpgroup (arrival counts 256 per query tile instead of 128).
// Epilogue: the same two warps switch to the 32x32b view (one row per lane) and split the COLUMNS of the quarter's 32 rows,
// so every lane still stores contiguous 64- or 128-byte pieces of one output row; 1/l and the log-sum-exp of a row come from
// the thread that owned it during the key loop through 2 KB of shared memory and a 64-thread named barrier.
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT, bool OVEC32, int EMU>
__device__ __forceinline__ void softmaxRows16(uint32_t smem_base, uint32_t tmem_base, const FwdParams& p, const int t, const int h) {
    using L = SmemLayout<D, STAGES>;
    uint32_t bar0 = smem_base + L::kBarOff;
    asm volatile("" : "+r"(bar0));     // keep in a register
    const uint32_t s_full = bar0 + 8u * (L::kBarSFull + t);
    const uint32_t p_full0 = bar0 + 8u * (L::kBarPFull + 2 * t);
    const uint32_t p_full1 = p_full0 + 8u;
    const uint32_t o_full = bar0 + 8u * (L::kBarOFull + t);
    const uint32_t o_free = bar0 + 8u * (L::kBarOFree + t);
    const uint32_t s_free = bar0 + 8u * (L::kBarSFree + t);
    const uint32_t o_half = bar0 + 8u * (L::kBarOHalf + t);

    const int q4 = (threadIdx.x / 32) & 3;     // TMEM lane quarter (and SM sub-partition) of this warp
    const int lane = threadIdx.x & 31;
    const int c4 = lane & 3, rr = lane >> 2;
    const int r_in_tile = q4 * 32 + h * 16 + rr;               // first of this thread's two rows inside the query tile; the other is + 8
    const uint32_t lane16 = uint32_t(q4 * 32 + h * 16) << 16;  // 16-lane window of the key loop
    const uint32_t lane32 = uint32_t(q4 * 32) << 16;           // 32-lane window of the epilogue
    uint32_t tS = tmem_base + lane16 + kTmemS;
    uint32_t tP = tmem_base + lane16 + tmem_p_col(t);
    uint32_t tO = tmem_base + lane16 + kTmemO0 + 128u * t;
    const float c = p.scale_log2;
    asm volatile("" : "+r"(tS), "+r"(tP), "+r"(tO));
    const uint32_t exch = smem_base + L::kExchOff + uint32_t(t * kBlockM) * 8u;

    int st = 0;      // key tiles this query tile has processed so far (barrier phase bookkeeping)
    for (int k = 0;; ++k) {
        int item, n, j_mask;     // key tiles of this query tile; first key tile that needs the causal / tail mask (kept instead of the item's coordinates)
        {
            WorkItem w;          // the decoded item from the mailbox; the two later uses (mask rows, epilogue) re-derive it from `item`
            item = fetch_item<D, STAGES>(smem_base, k, w);
            if (item < 0) break;
            n = w.n_tile(t);
            const int tile_row0 = w.q0 + t * kBlockM;
            // tail: kv0 + 128 > Nk  <=>  j >= Nk / 128;   causal: kv0 + 127 > tile_row0 + off  <=>  128 j > tile_row0 + off - 127
            j_mask = p.Nk / kBlockN;
            if (p.causal) {
                const int x = tile_row0 + p.causal_off - (kBlockN - 1);
                const int jc = x < 0 ? 0 : x / kBlockN + 1;
                j_mask = jc < j_mask ? jc : j_mask;
            }
        }

        float mA = -INFINITY, mB = -INFINITY;            // max in use per row, raw (unscaled) score units
        float2 lA = make_float2(0.f, 0.f), lB = make_float2(0.f, 0.f);   // this thread's share of the two row sums

        for (int j = 0; j < n; ++j) {
            const uint32_t ph = (st + j) & 1;
            mbar_wait(s_full, ph);
            tc_fence_after();

            uint32_t r[kBlockN / 2];
            const int kv0 = j * kBlockN;
            const bool need_mask = j >= j_mask;
            float a0 = -INFINITY, a1 = -INFINITY, b0 = -INFINITY, b1 = -INFINITY;
            tmem_ld_16x256b_x8(tS, r);                   // keys 0..63
            tc_wait_ld();
            tmem_ld_16x256b_x8(tS + 64u, r + 32);        // keys 64..127 in flight while the first half's maxima are taken
            if (!need_mask) {
#pragma unroll
                for (int g = 0; g < 8; g += 2) {
                    a0 = max3(a0, __uint_as_float(r[4 * g + 0]), __uint_as_float(r[4 * g + 1]));
                    b0 = max3(b0, __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                    a1 = max3(a1, __uint_as_float(r[4 * g + 4]), __uint_as_float(r[4 * g + 5]));
                    b1 = max3(b1, __uint_as_float(r[4 * g + 6]), __uint_as_float(r[4 * g + 7]));
                }
            }
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(s_free);         // the scores are in registers: the shared S buffer may be overwritten

            if (need_mask) {
                // column of r[4g + 2*row + e] is 8g + 2 c4 + e: masked iff 8g + e > lim(row) - 2 c4
                const int rowA = decode_item(p, item).q0 + t * kBlockM + r_in_tile;
                const int limA_c = p.causal ? (rowA + p.causal_off) : 0x7fffffff;
                const int limB_c = p.causal ? (rowA + 8 + p.causal_off) : 0x7fffffff;
                const int limA = min(limA_c, p.Nk - 1) - kv0 - 2 * c4;
                const int limB = min(limB_c, p.Nk - 1) - kv0 - 2 * c4;
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    r[4 * g + 0] = mask_gt(r[4 * g + 0], 8 * g, limA);
                    r[4 * g + 1] = mask_gt(r[4 * g + 1], 8 * g + 1, limA);
                    r[4 * g + 2] = mask_gt(r[4 * g + 2], 8 * g, limB);
                    r[4 * g + 3] = mask_gt(r[4 * g + 3], 8 * g + 1, limB);
                }
#pragma unroll
                for (int g = 0; g < 8; g += 2) {
                    a0 = max3(a0, __uint_as_float(r[4 * g + 0]), __uint_as_float(r[4 * g + 1]));
                    b0 = max3(b0, __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                    a1 = max3(a1, __uint_as_float(r[4 * g + 4]), __uint_as_float(r[4 * g + 5]));
                    b1 = max3(b1, __uint_as_float(r[4 * g + 6]), __uint_as_float(r[4 * g + 7]));
                }
            }
#pragma unroll
            for (int g = 8; g < 16; g += 2) {
                a0 = max3(a0, __uint_as_float(r[4 * g + 0]), __uint_as_float(r[4 * g + 1]));
                b0 = max3(b0, __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                a1 = max3(a1, __uint_as_float(r[4 * g + 4]), __uint_as_float(r[4 * g + 5]));
                b1 = max3(b1, __uint_as_float(r[4 * g + 6]), __uint_as_float(r[4 * g + 7]));
            }
            // a row lives in the 4 threads of a quad
            float mxA = fmaxf(a0, a1), mxB = fmaxf(b0, b1);
            mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 1));
            mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 1));
            mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 2));
            mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 2));
            const float nA = fmaxf(mxA, mA), nB = fmaxf(mxB, mB);

            if (j == 0) {
                mA = nA;
                mB = nB;
            } else {
                // lazy rescale: (new - old) is NaN when both are -inf -> compares false
                const bool gA = (nA - mA) * c > kRescaleThreshold, gB = (nB - mB) * c > kRescaleThreshold;
                if (__any_sync(0xffffffffu, gA || gB)) {
                    const float fA = gA ? ex2_approx((mA - nA) * c) : 1.0f;
                    const float fB = gB ? ex2_approx((mB - nB) * c) : 1.0f;
                    mbar_wait(o_full, ph ^ 1);      // P V of the previous key tile has retired
                    tc_fence_after();
                    // 16 columns at a time: the path is rare and must not cost the key loop any registers
#pragma unroll 1
                    for (int ch = 0; ch < D / 16; ++ch) {
                        uint32_t o[8];
                        tmem_ld_16x256b_x2(tO + 16u * ch, o);
                        tc_wait_ld();
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            o[4 * g + 0] = __float_as_uint(__uint_as_float(o[4 * g + 0]) * fA);
                            o[4 * g + 1] = __float_as_uint(__uint_as_float(o[4 * g + 1]) * fA);
                            o[4 * g + 2] = __float_as_uint(__uint_as_float(o[4 * g + 2]) * fB);
                            o[4 * g + 3] = __float_as_uint(__uint_as_float(o[4 * g + 3]) * fB);
                        }
                        tmem_st_16x256b_x2(tO + 16u * ch, o);
                    }
                    lA.x *= fA; lA.y *= fA;
                    lB.x *= fB; lB.y *= fB;
                    if (gA) mA = nA;
                    if (gB) mB = nB;
                }
            }
            const float msA = (mA == -INFINITY) ? 0.f : mA, msB = (mB == -INFINITY) ? 0.f : mB;
            const float2 c2 = make_float2(c, c);
            const float2 nmA = make_float2(-msA * c, -msA * c), nmB = make_float2(-msB * c, -msB * c);

            // exp2 of one score pair (both of the same row): MUFU.EX2 for most pairs, FMA-pipe emulation for EMU of 8
            auto exp_pair = [&](int i, const float2& nm) -> float2 {
                float2 x = fma2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2, nm);
                if (((i / 2) % 8) * 3 % 8 < EMU) {
                    x = ex2_emu2(x);
                } else {
                    x.x = ex2_approx(x.x);
                    x.y = ex2_approx(x.y);
                }
                return x;
            };
            // groups [g0, g0 + ng) of 8 keys -> 2 packed columns each, written IN PLACE over the first half of the registers
            // the scores came from (packed column 2g + e replaces r[2g + e], which group g / 2 has already consumed), so
            // that the block the tcgen05.st reads needs no registers of its own: at 104 registers per thread there are none
            auto exp_groups = [&](int g0, int ng) {
#pragma unroll
                for (int g = g0; g < g0 + ng; ++g) {
                    const float2 xA = exp_pair(4 * g, nmA), xB = exp_pair(4 * g + 2, nmB);
                    lA = add2(lA, xA);
                    lB = add2(lB, xB);
                    const int o = (g < 8) ? 2 * g : 32 + 2 * (g - 8);
                    r[o + 0] = pack16<DT>(xA.x, xA.y);
                    r[o + 1] = pack16<DT>(xB.x, xB.y);
                }
            };
            // P_t V_{j-1} (its first half, with the split wait) must have retired before the P columns are overwritten
            if (j > 0) {
                mbar_wait(kSplitOFull<D> ? o_half : o_full, ph ^ 1);
                tc_fence_after();
            }
            exp_groups(0, 8);
            tmem_st_16x128b_x8(tP, r);             // keys 0..63 of P
            exp_groups(8, 4);
            tc_wait_st();                          // first half landed while the third quarter was computed
            tc_fence_before();
            mbar_arrive(p_full0);                  // MMA may start P V on keys 0..63
            exp_groups(12, 4);
            if (kSplitOFull<D> && j > 0) {
                mbar_wait(o_full, ph ^ 1);
                tc_fence_after();
            }
            tmem_st_16x128b_x8(tP + 32u, r + 32);  // keys 64..127 of P
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(p_full1);
        }

        // ---- epilogue: 32x32b view, this warp stores columns [h D/2, (h+1) D/2) of the quarter's 32 rows ----
        constexpr int HC = D / 2;
        const WorkItem w = decode_item(p, item);
        const int tile_row0 = (t * kBlockM < w.rows) ? w.q0 + t * kBlockM : kNoRow;   // half item: slot 1 has no rows, nothing is written
        const int rowA = tile_row0 + r_in_tile;
        const int row = tile_row0 + q4 * 32 + lane;
        const bool row_ok = row < p.Nq;
        const long long bh_lin = (long long)w.b * p.Hq + w.h;
        const long long row_lin = bh_lin * p.Nq + row;
        const bool carry = p.acc_o != nullptr;
        float2 mine = make_float2(0.f, -INFINITY);     // plain: (1/l, lse) of `row`; carry: (weight of the running O, weight of this call's unnormalised O)
        if (n > 0) {
            float sA = lA.x + lA.y, sB = lB.x + lB.y;
            sA += __shfl_xor_sync(0xffffffffu, sA, 1);
            sB += __shfl_xor_sync(0xffffffffu, sB, 1);
            sA += __shfl_xor_sync(0xffffffffu, sA, 2);
            sB += __shfl_xor_sync(0xffffffffu, sB, 2);
            if (c4 == 0) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float l_run = e ? sB : sA, m_run = e ? mB : mA;
                    const float m_fin = (m_run == -INFINITY) ? 0.f : m_run;
                    const float lse_part = (l_run > 0.f) ? (m_fin * p.scale + logf(l_run)) : -INFINITY;
                    float2 v = make_float2((l_run > 0.f) ? 1.0f / l_run : 0.f, lse_part);
                    if (carry) {
                        // fold into the running log-sum-exp here, by the one thread that owns the row:
                        //   lse' = log(e^lse_acc + e^lse_part),  O' = O_acc e^(lse_acc-lse') + (O/l) e^(lse_part-lse')
                        const int orow_e = rowA + 8 * e;
                        float wa = 1.f, wp = 0.f;
                        if (orow_e < p.Nq) {
                            const long long lin = bh_lin * p.acc_rows + p.acc_off + orow_e;
                            const float lse_acc = p.acc_lse[lin];
                            const float mx = fmaxf(lse_acc, lse_part);
                            if (mx != -INFINITY) {
                                const float ea = expf(lse_acc - mx), ep = expf(lse_part - mx);
                                const float inv = 1.0f / (ea + ep);
                                wa = ea * inv;
                                wp = ep * inv * v.x;
                                p.acc_lse[lin] = mx + logf(ea + ep);
                            }
                        }
                        v = make_float2(wa, wp);
                    }
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(exch + uint32_t(r_in_tile + 8 * e) * 8u), "f"(v.x), "f"(v.y) : "memory");
                }
            }
            named_bar_sync(1u + uint32_t(t * 4 + q4), 64u);     // the two warps of this (tile, quarter)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(mine.x), "=f"(mine.y) : "r"(exch + uint32_t(q4 * 32 + lane) * 8u) : "memory");
        }
        const uint32_t tO32 = tmem_base + lane32 + kTmemO0 + 128u * t + uint32_t(HC * h);
        if (!carry) {
            const float inv_l = mine.x;
            uint16_t* orow = reinterpret_cast<uint16_t*>(p.O) + (long long)w.b * p.o_stride_b + (long long)w.h * p.o_stride_h +
                             (long long)row * p.o_stride_n + HC * h;
            if (n > 0) {
                mbar_wait(o_full, (st + n - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int qq = 0; qq < HC / 16; ++qq) {       // 16 columns = one 32-byte sector of the row at a time
                    uint32_t o[16];
                    tmem_ld16(tO32 + 16u * qq, o);
                    tc_wait_ld();
                    uint32_t hh[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        hh[i] = pack16<DT>(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
                    if (row_ok) {
                        if constexpr (OVEC32) {
                            st_global_v8(orow + 16 * qq, hh);
                        } else {
                            st_global_v4(orow + 16 * qq, hh[0], hh[1], hh[2], hh[3]);
                            st_global_v4(orow + 16 * qq + 8, hh[4], hh[5], hh[6], hh[7]);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(o_free);         // O columns may be overwritten by the next item's first P V
            } else if (row_ok) {
#pragma unroll
                for (int i = 0; i < HC / 8; ++i) st_global_v4(orow + 8 * i, 0u, 0u, 0u, 0u);
            }
            if (p.lse != nullptr && row_ok && h == 0) p.lse[row_lin] = mine.y;
        } else if (n > 0) {
            // carry mode (ring-KV step): O_acc <- O_acc * wa + O * wp on this warp's half of the columns
            const float wa = mine.x, wp = mine.y;
            mbar_wait(o_full, (st + n - 1) & 1);
            tc_fence_after();
            float* arow = p.acc_o + (bh_lin * p.acc_rows + p.acc_off + row) * D + HC * h;
#pragma unroll
            for (int qq = 0; qq < HC / 16; ++qq) {
                uint32_t o[16];
                tmem_ld16(tO32 + 16u * qq, o);
                tc_wait_ld();
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 a = ld_global_f4(arow + 16 * qq + 4 * i);
                        a.x = a.x * wa + __uint_as_float(o[4 * i + 0]) * wp;
                        a.y = a.y * wa + __uint_as_float(o[4 * i + 1]) * wp;
                        a.z = a.z * wa + __uint_as_float(o[4 * i + 2]) * wp;
                        a.w = a.w * wa + __uint_as_float(o[4 * i + 3]) * wp;
                        st_global_f4(arow + 16 * qq + 4 * i, a);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(o_free);
        }
        st += n;
    }
    tc_fence_before();
}

}  // namespace fa
