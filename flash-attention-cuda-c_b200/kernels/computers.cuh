// computers.cuh — the math stage of the B200 attention forward path.
//
// Takes the place of the reference's kernels/computers.cuh (reference: computers.cuh:5-69, one warp
// per query row, one lane group per key row, scalar FFMA dot products + shuffle reductions, smem O
// read-modify-write) and of the per-tile helpers it calls (reference: utils.cuh:17-113).
//
// Here the two contractions are tcgen05 MMAs accumulating in tensor memory:
//   S_t = Q_t K_j^T   (A = Q tile, B = K tile, both K-major in 128B-swizzled smem)   -> TMEM cols [128t, 128t+128)
//   O_t += P_t V_j    (A = P_t in TMEM as packed 16-bit, B = V tile MN-major in smem) -> TMEM cols [256+128t, ..+D)
// issued by ONE thread (mmaIssuerThread).  Two softmax warpgroups (one per query tile t) read S
// with tcgen05.ld in the 32x32b shape — thread i of warp w owns TMEM lane 32*(w%4)+i, i.e. one whole
// score row, so row max / row sum need no shuffles — and keep the online-softmax state
// (running max m, running sum l) in registers:
//   m' = max(m, rowmax(S));  P = exp2(S*c - m*c)  (c = scale*log2 e);  l += rowsum(P)
// Normalisation by 1/l is deferred to the epilogue (the reference normalises every tile,
// utils.cuh:79-80).  O is rescaled lazily: only when a row's max grew by more than 2^8 since the max
// in use (then O *= exp2((m_used - m')c) through a tcgen05.ld / tcgen05.st round trip).
// P overwrites the first 64 columns of its own S tile (two 16-bit values per 32-bit column); the
// tensor pipe executes MMAs in issue order, so "P_t V_j" followed by "Q_t K_{j+1}^T -> S_t" is safe.
#pragma once

#include "loaders.cuh"

namespace fa {

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemS0 = 0;      // S tile t at columns 128*t
constexpr uint32_t kTmemO0 = 256;    // O tile t at columns 256 + 128*t
constexpr float kRescaleThreshold = 8.0f;   // log2 units

// ------------------------------------------------------------------------------------------------
// MMA issuer: the whole warp walks the schedule (so that addresses and descriptors stay warp-uniform and live in
// uniform registers); one elected lane issues each tcgen05.mma / tcgen05.commit.
// Issue order per key tile j (t = query tile):  P_0V_j, Q_0K_{j+1}, P_1V_j, Q_1K_{j+1}
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT>
__device__ __forceinline__ void mmaIssuerWarp(uint32_t smem_base, uint32_t tmem_base_in, const WorkItem& w,
                                              unsigned long long* prof = nullptr) {
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);   // tell the compiler it is warp-uniform
    using L = SmemLayout<D, STAGES>;
    constexpr uint32_t kFmt = (DT == kBF16) ? 1u : 0u;
    constexpr uint32_t idesc_qk = umma_idesc(kBlockM, kBlockN, kFmt, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc(kBlockM, D, kFmt, 0, 1);
    const uint32_t bar0 = smem_base + L::kBarOff;
    auto bar = [&](int i) { return bar0 + 8u * uint32_t(i); };

    if (w.n_kv <= 0) return;

    // Descriptor templates with a zero start address; the 14-bit address field (bytes >> 4) is added per MMA.
    const uint64_t desc_k_major = umma_desc_sw128(0, 16, 1024);             // Q and K tiles (K-major)
    const uint64_t desc_mn_major = umma_desc_sw128(0, kHalfBytes, 1024);    // V tile (MN-major), 64-column halves 16 KiB apart
    auto issue_qk = [&](int t, uint32_t k_smem) {
        const uint64_t a0 = desc_k_major + ((smem_base + L::kQOff + t * L::kQTileBytes) >> 4);
        const uint64_t b0 = desc_k_major + (k_smem >> 4);
        const uint32_t d_tmem = tmem_base + kTmemS0 + 128u * t;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            // 16 halfs = 32 B inside the 128-B swizzle row; the second 64 columns live one half (16 KiB) further
            const uint32_t off = ((ks / 4) * kHalfBytes + (ks % 4) * 32) >> 4;
            if (elect_one_sync()) umma_ss(d_tmem, a0 + off, b0 + off, idesc_qk, ks > 0);
        }
    };
    // P_t V_j in two halves of 4 k-steps (64 keys each): the first half can start while the softmax warpgroup is still
    // producing the second half of P.
    auto issue_pv_half = [&](int t, uint32_t v_smem, bool accumulate, int half) {
        const uint32_t p_tmem = tmem_base + kTmemS0 + 128u * t;    // P aliases the head of S_t
        const uint32_t d_tmem = tmem_base + kTmemO0 + 128u * t;
        const uint64_t b0 = desc_mn_major + (v_smem >> 4);
#pragma unroll
        for (int kk = 0; kk < kBlockN / 32; ++kk) {
            const int ks = half * (kBlockN / 32) + kk;
            // 16 key rows = 2 swizzle atoms of 8 rows x 128 B = 2048 B
            if (elect_one_sync()) umma_ts(d_tmem, p_tmem + 8u * ks, b0 + ((ks * 2048) >> 4), idesc_pv, (accumulate || ks > 0) ? 1u : 0u);
        }
    };
    auto slot_addr = [&](int it) { return smem_base + L::kKVOff + (it % STAGES) * L::kKVTileBytes; };
    auto wait_full = [&](int it) { mbar_wait(bar(L::kBarKVFull + it % STAGES), (it / STAGES) & 1); };
    auto commit = [&](uint32_t b) { if (elect_one_sync()) tc_commit(b); };
    auto release = [&](int it) { commit(bar(L::kBarKVEmpty + it % STAGES)); };

    FA_PROF_DECL(4);
    mbar_wait(bar(L::kBarQFull), 0);
    wait_full(0);
    tc_fence_after();
    FA_PROF_MARK(0);                 // prologue: Q + K0 arrival
    issue_qk(0, slot_addr(0));
    commit(bar(L::kBarSFull + 0));
    issue_qk(1, slot_addr(0));
    commit(bar(L::kBarSFull + 1));
    release(0);

    for (int j = 0; j < w.n_kv; ++j) {
        const int it_v = 2 * j + 1, it_k = 2 * j + 2;
        const bool has_next = j + 1 < w.n_kv;
        FA_PROF_MARK(3);             // issue + bookkeeping
        wait_full(it_v);
        FA_PROF_MARK(1);             // waiting for V/K tiles
#pragma unroll
        for (int t = 0; t < kTilesPerCta; ++t) {
            mbar_wait(bar(L::kBarPFull + 2 * t), j & 1);
            tc_fence_after();
            FA_PROF_MARK(2);         // waiting for P
            issue_pv_half(t, slot_addr(it_v), j > 0, 0);
            FA_PROF_MARK(3);
            mbar_wait(bar(L::kBarPFull + 2 * t + 1), j & 1);
            tc_fence_after();
            FA_PROF_MARK(2);
            issue_pv_half(t, slot_addr(it_v), j > 0, 1);
            commit(bar(L::kBarOFull + t));
            if (has_next) {
                if (t == 0) {
                    FA_PROF_MARK(3);
                    wait_full(it_k);
                    tc_fence_after();
                    FA_PROF_MARK(1);
                }
                issue_qk(t, slot_addr(it_k));
                commit(bar(L::kBarSFull + t));
            }
        }
        release(it_v);
        if (has_next) release(it_k);
    }
    FA_PROF_MARK(3);
    if ((threadIdx.x & 31) == 0) FA_PROF_FLUSH(prof, 8, 4);
}

// ------------------------------------------------------------------------------------------------
// Softmax warpgroup for query tile t (128 threads, one score row each), including the lazy O
// rescale and the epilogue (O/l -> global, optional LSE).
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT>
__device__ __forceinline__ void softmaxWarpgroup(uint32_t smem_base, uint32_t tmem_base, const WorkItem& w,
                                                 const FwdParams& p, int t) {
    using L = SmemLayout<D, STAGES>;
    const uint32_t bar0 = smem_base + L::kBarOff;
    const uint32_t s_full = bar0 + 8u * (L::kBarSFull + t);
    const uint32_t p_full0 = bar0 + 8u * (L::kBarPFull + 2 * t);
    const uint32_t p_full1 = p_full0 + 8u;
    const uint32_t o_full = bar0 + 8u * (L::kBarOFull + t);

    const int warp_in_wg = (threadIdx.x / 32) & 3;
    const int lane = threadIdx.x & 31;
    const uint32_t lane_base = uint32_t(warp_in_wg * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + kTmemS0 + 128u * t;
    const uint32_t tO = tmem_base + lane_base + kTmemO0 + 128u * t;

    const int tile_row0 = w.q0 + t * kBlockM;
    const int row = tile_row0 + warp_in_wg * 32 + lane;

    const float c = p.scale_log2;
    float m_run = -INFINITY;   // max in use, in raw (unscaled) score units
    float l_run = 0.f;

    FA_PROF_DECL(6);
    for (int j = 0; j < w.n_kv; ++j) {
        FA_PROF_MARK(5);             // loop overhead / l update
        mbar_wait(s_full, j & 1);
        tc_fence_after();
        FA_PROF_MARK(0);             // waiting for S

        uint32_t r[kBlockN];
#pragma unroll
        for (int q = 0; q < kBlockN / 32; ++q) tmem_ld32(tS + 32u * q, r + 32 * q);
        tc_wait_ld();
        FA_PROF_MARK(1);             // tcgen05.ld of the score row

        const int kv0 = j * kBlockN;
        const bool need_mask = (kv0 + kBlockN > p.Nk) || (p.causal && (kv0 + kBlockN - 1 > tile_row0 + p.causal_off));
        if (need_mask) {
            const int lim_c = p.causal ? (row + p.causal_off) : 0x7fffffff;
            const int lim = min(lim_c, p.Nk - 1) - kv0;   // columns c > lim are masked
#pragma unroll
            for (int cc = 0; cc < kBlockN; ++cc) r[cc] = mask_gt(r[cc], cc, lim);   // -inf where cc > lim
        }

        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int cc = 0; cc < kBlockN; cc += 8) {
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
                mbar_wait(o_full, (j - 1) & 1);
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
        // exp2 of one score pair: MUFU.EX2 for most pairs, FMA-pipe emulation for kEmuPairsPer8 of every 8
        auto exp_pair = [&](int col) -> float2 {
            float2 x = fma2(make_float2(__uint_as_float(r[col]), __uint_as_float(r[col + 1])), c2, nm2);
            if (((col / 2) % 8) * 3 % 8 < kEmuPairsPer8) {     // spread the emulated pairs evenly over the group of 8
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
        {
            uint32_t pk[32];
            exp_quarter(0, pk);
            exp_quarter(1, pk + 16);
            tmem_st32(tS, pk);                 // keys 0..63 of P
        }
        {
            uint32_t pk[32];
            exp_quarter(2, pk);
            tc_wait_st();                      // first half landed while quarter 2 was computed
            tc_fence_before();
            mbar_arrive(p_full0);              // MMA may start P V on keys 0..63
            exp_quarter(3, pk + 16);
            tmem_st32(tS + 32u, pk);           // keys 64..127 of P
        }
        FA_PROF_MARK(3);             // exp2 / pack / tcgen05.st issue
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(p_full1);
        FA_PROF_MARK(4);             // store drain + arrive

        l_run += (s0.x + s0.y) + (s1.x + s1.y);
    }
    if ((threadIdx.x & 31) == 0) FA_PROF_FLUSH(p.prof, 0, 6);

    // ---- epilogue: O / l -> global ----
    const bool row_ok = row < p.Nq;
    const float inv_l = (l_run > 0.f) ? 1.0f / l_run : 0.f;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.O) + (long long)w.b * p.o_stride_b + (long long)w.h * p.o_stride_h +
                     (long long)row * p.o_stride_n;
    if (w.n_kv > 0) {
        mbar_wait(o_full, (w.n_kv - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < D / 32; ++q) {
            uint32_t o[32];
            tmem_ld32(tO + 32u * q, o);
            tc_wait_ld();
            uint32_t h[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
                h[i] = pack16<DT>(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_global_v4(orow + 32 * q + 8 * i, h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
            }
        }
    } else if (row_ok) {
#pragma unroll
        for (int i = 0; i < D / 8; ++i) st_global_v4(orow + 8 * i, 0u, 0u, 0u, 0u);
    }
    if (p.lse != nullptr && row_ok) {
        const float m_safe = (m_run == -INFINITY) ? 0.f : m_run;
        p.lse[((long long)w.b * p.Hq + w.h) * p.Nq + row] = (l_run > 0.f) ? (m_safe * p.scale + logf(l_run)) : -INFINITY;
    }
    tc_fence_before();
}

}  // namespace fa
