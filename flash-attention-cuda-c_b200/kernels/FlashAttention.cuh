// FlashAttention.cuh — launcher API of the attention forward path.
//
// Keeps the reference's include path and its public entry point
//   template<int D_HEAD, int Q_TILE_ROWS, int KV_TILE_ROWS>
//   __global__ void twoLoaderMhaFlashAttentionKernel(const float* Q, const float* K, const float* V, float* O,
//                                                    int batchSize, int numHeads, int seqLen, float scale, bool is_causal)
// (reference: kernels/FlashAttention.cuh:59-63), and adds the B200-native kernel the C-ABI launcher in
// FlashAttention.cu dispatches to:
//   fa::fwdSm100Kernel<D, STAGES, DT, OVEC32, SW, EMU, ST, HS> — warp-specialised TMA + tcgen05/TMEM kernel (bf16 / fp16; OVEC32: O is
//                                        32-byte aligned, so the epilogue may use 256-bit stores; SW: 8 or 16 softmax warps;
//                                        EMU: share of the exponentials on the FMA pipe; ST: epilogue staged
//                                        through shared memory and written with TMA stores; HS: half items run split-KV
//                                        on both query-tile slots — the build small launches with a short last wave get)
//   fa::fwdSm100PairKernel<D, STAGES, DT, OVEC32, ST> — the same kernel for clusters of two CTAs (d = 128): tcgen05
//                                        cta_group::2 MMAs over 256 query rows, each CTA loading and reading half of
//                                        every K/V tile; pairs cut by rows (512-row items) or by the two heads of a kv group
//   fa::fwdFp32Kernel<D>                — exact-fp32 CUDA-core kernel for fp32 I/O
// The compat template is launched by the *caller* with a grid/block/shared-memory size of its own choosing
// (reference: tests/main.cu:51-61 uses grid 1, (QT+2)*32 threads, (3QT+4R)*D*4 bytes), so it cannot take TMA
// descriptors; it is a self-contained fp32 kernel that is correct for any launch shape and — unlike the
// reference — honours batch/head boundaries, uses K for the scores and handles
// causal masking without NaNs.
#pragma once

#include "utils.cuh"
#include "loaders.cuh"
#include "computers.cuh"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cuda_fp16.h>

namespace fa {

// ------------------------------------------------------------------------------------------------
// B200 kernel: persistent, grid = min(#SMs, work items), 1 CTA / SM; a work item is a 256-row query block of one
// (batch, head), claimed from a global atomic counter (LPT order inside a head, (batch, head)-major overall).
// With S = SW (8 or 16) softmax warps the CTA has S + 4 warps:
//   warps 0 .. S/2-1   softmax + correction + epilogue for query tile 0
//   warps S/2 .. S-1   softmax + correction + epilogue for query tile 1
//   warp  S      MMA issuer: every Q K^T (d = 128) / everything of query tile 0 (d = 64)
//   warp  S+1    TMA producer + scheduler (one thread)
//   warp  S+2    TMEM allocator, then MMA issuer: every P V (d = 128) / everything of query tile 1 (d = 64)
//   warp  S+3    idle (times the CTA for scripts/cycles.py when a debug profile buffer is set)
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT, bool OVEC32, int SW, int EMU, int ST, int HS>
__global__ void __launch_bounds__(KCfg<SW>::kNumThreads, 1)
fwdSm100Kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
    using L = SmemLayout<D, STAGES>;
    using C = KCfg<SW>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) __trap();        // SWIZZLE_128B tiles need 1024-B alignment; see SmemLayout::kDynamicBytes
    uint8_t* smem_gen = smem_raw;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kTmemPtrOff);

    const int warp = threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    FA_T2_DECL;
    if (threadIdx.x == 0) FA_T2(p.prof, 0, 1);

    // Set-up in two phases, so that the producer does not wait for the slow parts (TMEM allocation, the register re-split):
    //   A  the producer thread initialises every mbarrier and prefetches the tensor maps; one CTA-wide barrier publishes them;
    //   B  the producer warp goes straight on — decodes the CTA's first item and issues its Q / K / V loads — while the TMEM
    //      warp allocates and all OTHER warps meet at a named barrier for the TMEM base address.
    // The first loads leave ~2,000 clk earlier than with a single barrier after everything (scripts/trace_cta.py): nothing for
    // a launch of many items per CTA, 5 % of BASELINE configs[1].
    if (warp == C::kLoadWarp && lane == 0) {
        const uint32_t bar0 = smem_base + L::kBarOff;
        mbar_init(bar0 + 8 * L::kBarQFull, 1);
        mbar_init(bar0 + 8 * L::kBarQEmpty, kIssuerByType<D> ? 1 : 2); // both MMA issuers (split by type: the Q K^T issuer alone)
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar0 + 8 * (L::kBarKVFull + s), 1);
            mbar_init(bar0 + 8 * (L::kBarKVEmpty + s), 2);          // both MMA issuers
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar0 + 8 * (L::kBarSFull + t), 1);
            mbar_init(bar0 + 8 * (L::kBarPFull + 2 * t), C::kSoftmaxThreadsPerTile);
            mbar_init(bar0 + 8 * (L::kBarPFull + 2 * t + 1), C::kSoftmaxThreadsPerTile);
            mbar_init(bar0 + 8 * (L::kBarOFull + t), 1);
            mbar_init(bar0 + 8 * (L::kBarOFree + t), C::kSoftmaxThreadsPerTile);
            mbar_init(bar0 + 8 * (L::kBarSchedFull + t), 1);
            mbar_init(bar0 + 8 * (L::kBarSchedEmpty + t), 2 + C::kSoftmaxWarps);   // both MMA issuers + every softmax warp
            mbar_init(bar0 + 8 * (L::kBarSFree + t), C::kSoftmaxThreadsPerTile);
            mbar_init(bar0 + 8 * (L::kBarOHalf + t), 1);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        if (ST != 0) tma_prefetch_desc(&tmO);
    }
    __syncthreads();                                   // phase A done: every barrier exists
    if (threadIdx.x == 0) FA_T2(p.prof, 0, 2);

    uint32_t tmem_base = 0;
    long long cta_t0 = 0;
    if (warp == C::kLoadWarp) {
        reg_dec<C::kOtherRegs>();
        if (lane == 0) tmaLoaderThread<D, STAGES>(&tmQ, &tmK, &tmV, smem_base, p);
    } else {
        if (warp == C::kTmemWarp) {
            tmem_alloc(smem_base + L::kTmemPtrOff, kTmemCols);
            tmem_relinquish();
        }
        tc_fence_before();
        named_bar_sync(9u, uint32_t(C::kNumThreads - 32));      // phase B: everyone but the producer warp
        tc_fence_after();
        tmem_base = *tmem_ptr;
        // Cycle counter for A/B work (fa_debug_set_profile_buffer): the otherwise idle last warp times the CTA from here to the
        // final barrier; prof[30] = max over CTAs, prof[31] = sum.  Wall-clock A/B on a power-capped GPU is too noisy.
        if (p.prof != nullptr && threadIdx.x == C::kNumThreads - 32) cta_t0 = clock64();

        if (warp < C::kSoftmaxWarps) {
            reg_inc<C::kSoftmaxRegs>();
            if constexpr (C::kRows16) softmaxRows16<D, STAGES, DT, OVEC32, EMU>(smem_base, tmem_base, p, warp / 8, (warp / 4) & 1);
            else softmaxWarpgroup<D, STAGES, DT, OVEC32, EMU, ST, HS>(smem_base, tmem_base, p, warp / 4, &tmO);
        } else {
            reg_dec<C::kOtherRegs>();
            if (warp == C::kMmaWarp0) {
                if constexpr (kIssuerByType<D>) mmaTypeIssuerWarp<D, STAGES, DT, SW, HS>(smem_base, tmem_base, p, 0);
                else mmaIssuerWarp<D, STAGES, DT, SW, HS>(smem_base, tmem_base, p, 0);
            } else if (warp == C::kMmaWarp1) {
                if constexpr (kIssuerByType<D>) mmaTypeIssuerWarp<D, STAGES, DT, SW, HS>(smem_base, tmem_base, p, 1);
                else mmaIssuerWarp<D, STAGES, DT, SW, HS>(smem_base, tmem_base, p, 1);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) FA_T2(p.prof, 0, 50);
    if (warp == C::kTmemWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (p.prof != nullptr && threadIdx.x == C::kNumThreads - 32) {
        const unsigned long long dt = (unsigned long long)(clock64() - cta_t0);
        atomicMax(p.prof + 30, dt);
        atomicAdd(p.prof + 31, dt);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel (d = 128): clusters of two CTAs on the two SMs of a TPC, tcgen05 cta_group::2.  A pair works on 512-row
// query blocks; every MMA covers 256 query rows (128 in each CTA's TMEM) against a 128-key tile of which each CTA loads, holds
// and reads only half (K: 64 of the keys; V: 64 of the d columns).  Per CTA and step that is 32 KiB instead of 64 KiB of
// L2 -> shared-memory traffic and 128 KiB instead of 192 KiB of operand reads by the tensor core; the warp roles, the TMEM map
// and the softmax are those of fwdSm100Kernel.  The MMA issuer warps work in the leader (even) CTA only.
//   Why: on a power-capped B200 the step time in CYCLES does not depend on the K/V traffic at all, but the clock does — with
//   the K/V loads removed (wrong results, same instruction stream) the same kernel sustains 7 % more TFLOP/s, with half of
//   them 2.5 % more (profiles/r2_energy_kv_traffic.log).
// ------------------------------------------------------------------------------------------------
template <int D, int STAGES, int DT, bool OVEC32, int ST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(KCfg<8>::kNumThreads, 1)
fwdSm100PairKernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FwdParams p) {
    using L = SmemLayout<D, STAGES, 2>;
    using C = KCfg<8>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) __trap();
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + L::kTmemPtrOff);

    const int warp = threadIdx.x / 32;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    if (warp == C::kLoadWarp && lane == 0) {
        // arrival counts: both producers on the leader's full barriers; one multicast commit on everything the MMAs signal;
        // one arrival per softmax warp of BOTH CTAs on what the issuers wait for (the peer's copies of those are never used)
        const uint32_t bar0 = smem_base + L::kBarOff;
        mbar_init(bar0 + 8 * L::kBarQFull, 2);
        mbar_init(bar0 + 8 * L::kBarQEmpty, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar0 + 8 * (L::kBarKVFull + s), 2);
            mbar_init(bar0 + 8 * (L::kBarKVEmpty + s), 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar0 + 8 * (L::kBarSFull + t), 1);
            mbar_init(bar0 + 8 * (L::kBarPFull + 2 * t), kPairWarpArrivals);
            mbar_init(bar0 + 8 * (L::kBarPFull + 2 * t + 1), kPairWarpArrivals);
            mbar_init(bar0 + 8 * (L::kBarOFull + t), 1);
            mbar_init(bar0 + 8 * (L::kBarOFree + t), kPairWarpArrivals);
            mbar_init(bar0 + 8 * (L::kBarSchedFull + t), 1);
            // leader: its 2 issuers + 8 softmax warps, and the peer's 8 softmax warps + producer
            mbar_init(bar0 + 8 * (L::kBarSchedEmpty + t), 2 + C::kSoftmaxWarps + C::kSoftmaxWarps + 1);
            mbar_init(bar0 + 8 * (L::kBarSFree + t), kPairWarpArrivals);
            mbar_init(bar0 + 8 * (L::kBarOHalf + t), 1);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        if (ST != 0) tma_prefetch_desc(&tmO);
    }
    __syncwarp();
    cluster_sync_all();                                // every barrier of BOTH CTAs exists before anything signals across

    uint32_t tmem_base = 0;
    long long cta_t0 = 0;
    if (warp == C::kLoadWarp) {
        reg_dec<C::kOtherRegs>();
        if (lane == 0) tmaPairLoaderThread<D, STAGES>(&tmQ, &tmK, &tmV, smem_base, p);
    } else {
        if (warp == C::kTmemWarp) {
            tmem_alloc_pair(smem_base + L::kTmemPtrOff, kTmemCols);      // the same warp of both CTAs: 512 columns in each TMEM
            tmem_relinquish_pair();
        }
        tc_fence_before();
        named_bar_sync(9u, uint32_t(C::kNumThreads - 32));
        tc_fence_after();
        tmem_base = *tmem_ptr;
        if (p.prof != nullptr && threadIdx.x == C::kNumThreads - 32) cta_t0 = clock64();

        if (warp < C::kSoftmaxWarps) {
            reg_inc<C::kSoftmaxRegs>();
            softmaxWarpgroup<D, STAGES, DT, OVEC32, 0, ST, 0, 2>(smem_base, tmem_base, p, warp / 4, &tmO);
        } else {
            reg_dec<C::kOtherRegs>();
            if (rank == 0) {
                if (warp == C::kMmaWarp0) mmaPairIssuerWarp<D, STAGES, DT>(smem_base, tmem_base, p, 0);
                else if (warp == C::kMmaWarp1) mmaPairIssuerWarp<D, STAGES, DT>(smem_base, tmem_base, p, 1);
            }
        }
    }

    // Nothing of either CTA may still be in flight towards the other's shared memory or TMEM when one of them exits.
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();
    if (warp == C::kTmemWarp) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
    if (p.prof != nullptr && threadIdx.x == C::kNumThreads - 32) {
        const unsigned long long dt = (unsigned long long)(clock64() - cta_t0);
        atomicMax(p.prof + 30, dt);
        atomicAdd(p.prof + 31, dt);
    }
}

// ------------------------------------------------------------------------------------------------
// Exact-fp32 kernel (fp32 in, fp32 out, fp32 accumulate on the FMA pipe).  grid = (ceil(Nq/64), Hq, B),
// 256 threads.  A 64-row query tile against 64-row key/value tiles staged in shared memory; each thread owns
// a 4x4 patch of S and a 4 x (D/16) patch of O.  Used for fp32 I/O where TF32/bf16 tensor-core products would
// miss the 1e-4 relative tolerance.
// ------------------------------------------------------------------------------------------------
constexpr int kF32Rows = 64;
constexpr int kF32Threads = 256;

template <int D>
struct Fp32Smem {
    // Qt[d][64] (transposed), Kt[d][64] (transposed), V[64][D], P[64][65]
    static constexpr int kQt = 0;
    static constexpr int kKt = kQt + D * kF32Rows;
    static constexpr int kV = kKt + D * kF32Rows;
    static constexpr int kP = kV + kF32Rows * D;
    static constexpr int kFloats = kP + kF32Rows * (kF32Rows + 1);
    static constexpr int kBytes = kFloats * 4;
};

struct Fp32Params {
    const float* Q;
    const float* K;
    const float* V;
    float* O;
    float* lse;
    int B, Hq, Hkv, Nq, Nk;
    long long q_sb, q_sh, q_sn, k_sb, k_sh, k_sn, v_sb, v_sh, v_sn, o_sb, o_sh, o_sn;   // element strides
    float scale;
    int causal, causal_off, q_heads_per_kv;
};

template <int D>
__global__ void __launch_bounds__(kF32Threads) fwdFp32Kernel(const Fp32Params p) {
    static_assert(D % 16 == 0 && D <= 256, "head dim must be a multiple of 16");
    using S = Fp32Smem<D>;
    extern __shared__ float sm[];
    float* Qt = sm + S::kQt;
    float* Kt = sm + S::kKt;
    float* Vs = sm + S::kV;
    float* Ps = sm + S::kP;
    constexpr int DO = D / 16;   // O columns per thread

    const int tid = threadIdx.x;
    const int tx = tid & 15;     // column group
    const int ty = tid >> 4;     // row group
    const int qb = p.causal ? (int(gridDim.x) - 1 - int(blockIdx.x)) : int(blockIdx.x);
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / p.q_heads_per_kv;
    const int q0 = qb * kF32Rows;

    const float* Qg = p.Q + b * p.q_sb + h * p.q_sh;
    const float* Kg = p.K + b * p.k_sb + hk * p.k_sh;
    const float* Vg = p.V + b * p.v_sb + hk * p.v_sh;

    // Q tile -> smem, transposed (zero rows past Nq)
    for (int i = tid; i < kF32Rows * D; i += kF32Threads) {
        const int r = i / D, d = i % D;
        Qt[d * kF32Rows + r] = (q0 + r < p.Nq) ? Qg[(long long)(q0 + r) * p.q_sn + d] : 0.f;
    }

    float m[4], l[4], acc[4][DO];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY;
        l[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DO; ++d) acc[i][d] = 0.f;
    }

    int n_kv = (p.Nk + kF32Rows - 1) / kF32Rows;
    if (p.causal) {
        const int last_col = q0 + kF32Rows - 1 + p.causal_off;
        const int nc = last_col < 0 ? 0 : last_col / kF32Rows + 1;
        n_kv = nc < n_kv ? nc : n_kv;
    }

    for (int j = 0; j < n_kv; ++j) {
        const int kv0 = j * kF32Rows;
        __syncthreads();   // previous tile fully consumed (also covers the Q staging on j == 0)
        for (int i = tid; i < kF32Rows * D; i += kF32Threads) {
            const int r = i / D, d = i % D;
            const bool ok = kv0 + r < p.Nk;
            Kt[d * kF32Rows + r] = ok ? Kg[(long long)(kv0 + r) * p.k_sn + d] : 0.f;
            Vs[r * D + d] = ok ? Vg[(long long)(kv0 + r) * p.v_sn + d] : 0.f;
        }
        __syncthreads();

        // S patch: rows 4*ty..+3, cols 4*tx..+3
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) s[i][k] = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; ++d) {
            const float4 qv = *reinterpret_cast<const float4*>(Qt + d * kF32Rows + 4 * ty);
            const float4 kv = *reinterpret_cast<const float4*>(Kt + d * kF32Rows + 4 * tx);
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
            const float ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) s[i][k] = fmaf(qa[i], ka[k], s[i][k]);
        }
        // scale + mask, row max over the 16 threads sharing a row group
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = q0 + 4 * ty + i;
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = kv0 + 4 * tx + k;
                const bool masked = col >= p.Nk || (p.causal && col > row + p.causal_off);
                s[i][k] = masked ? -INFINITY : s[i][k] * p.scale;
                mx = fmaxf(mx, s[i][k]);
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m[i], mx);
            const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
            const float f = (m[i] == -INFINITY) ? 0.f : expf(m[i] - m_safe);
            float rs = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float e = expf(s[i][k] - m_safe);   // exp(-inf) = 0 for masked entries
                Ps[(4 * ty + i) * (kF32Rows + 1) + 4 * tx + k] = e;
                rs += e;
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            l[i] = l[i] * f + rs;
            m[i] = m_new;
#pragma unroll
            for (int d = 0; d < DO; ++d) acc[i][d] *= f;
        }
        __syncthreads();
        // O patch: rows 4*ty..+3, cols tx + 16*d
#pragma unroll 4
        for (int k = 0; k < kF32Rows; ++k) {
            float pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = Ps[(4 * ty + i) * (kF32Rows + 1) + k];
#pragma unroll
            for (int d = 0; d < DO; ++d) {
                const float v = Vs[k * D + tx + 16 * d];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][d] = fmaf(pv[i], v, acc[i][d]);
            }
        }
    }

    float* Og = p.O + b * p.o_sb + h * p.o_sh;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = q0 + 4 * ty + i;
        if (row >= p.Nq) continue;
        const float inv = l[i] > 0.f ? 1.0f / l[i] : 0.f;
#pragma unroll
        for (int d = 0; d < DO; ++d) Og[(long long)row * p.o_sn + tx + 16 * d] = acc[i][d] * inv;
        if (p.lse != nullptr && tx == 0)
            p.lse[((long long)b * p.Hq + h) * p.Nq + row] = l[i] > 0.f ? m[i] + logf(l[i]) : -INFINITY;
    }
}

// ------------------------------------------------------------------------------------------------
// Combine two partial attention results over disjoint key ranges (ring-KV step):
//   lse = log(exp(lse_a) + exp(lse_b));  O = O_a * exp(lse_a - lse) + O_b * exp(lse_b - lse)
// acc (fp32 O_a, lse_a) is updated in place with a 16-bit partial (O_b, lse_b).  One warp per row.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void mergePartialKernel(float* __restrict__ acc_o, float* __restrict__ acc_lse,
                                   const uint16_t* __restrict__ part_o, const float* __restrict__ part_lse,
                                   long long rows, int d) {
    const long long row = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float la = acc_lse[row], lb = part_lse[row];
    const float mx = fmaxf(la, lb);
    float wa, wb, lnew;
    if (mx == -INFINITY) {
        wa = 0.f; wb = 0.f; lnew = -INFINITY;
    } else {
        const float ea = expf(la - mx), eb = expf(lb - mx);
        const float inv = 1.0f / (ea + eb);
        wa = ea * inv; wb = eb * inv;
        lnew = mx + logf(ea + eb);
    }
    for (int c = lane * 2; c < d; c += 64) {
        const uint32_t pb = *reinterpret_cast<const uint32_t*>(part_o + row * d + c);
        float b0, b1;
        if constexpr (DT == kBF16) {
            b0 = __uint_as_float(pb << 16);
            b1 = __uint_as_float(pb & 0xffff0000u);
        } else {
            const __half2 hh = *reinterpret_cast<const __half2*>(&pb);
            b0 = __low2float(hh);
            b1 = __high2float(hh);
        }
        float2 a = *reinterpret_cast<float2*>(acc_o + row * d + c);
        a.x = a.x * wa + b0 * wb;
        a.y = a.y * wa + b1 * wb;
        *reinterpret_cast<float2*>(acc_o + row * d + c) = a;
    }
    if (lane == 0) acc_lse[row] = lnew;
}

// fp32 accumulator -> 16-bit output
template <int DT>
__global__ void castOutKernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const float2 v = reinterpret_cast<const float2*>(src)[i];
    reinterpret_cast<uint32_t*>(dst)[i] = pack16<DT>(v.x, v.y);
}

}  // namespace fa

// ------------------------------------------------------------------------------------------------
// Compat entry point: same name, template parameters and argument list as the reference
// (reference: kernels/FlashAttention.cuh:59-63).  Correct for any grid / block / dynamic-smem choice of the
// caller: query rows of all (batch, head) pairs are distributed warp-by-warp over the whole grid; each
// warp streams the keys of *its own* (batch, head) in chunks of 32, lane j scoring key j, then accumulates
// O with lanes owning output columns.  Q_TILE_ROWS / KV_TILE_ROWS are accepted for source compatibility;
// the result does not depend on them.
// ------------------------------------------------------------------------------------------------
template <int D_HEAD, int Q_TILE_ROWS, int KV_TILE_ROWS>
__global__ void twoLoaderMhaFlashAttentionKernel(
    const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, float* __restrict__ O,
    int batchSize, int numHeads, int seqLen, float scale, bool is_causal)
{
    constexpr int DPL = (D_HEAD + WARP - 1) / WARP;   // output columns per lane
    const int lane = threadIdx.x % WARP;
    const int warps_per_block = blockDim.x / WARP;
    const long long total_rows = (long long)batchSize * numHeads * seqLen;
    const long long warp_global = (long long)blockIdx.x * warps_per_block + threadIdx.x / WARP;
    const long long warp_stride = (long long)gridDim.x * warps_per_block;

    for (long long grow = warp_global; grow < total_rows; grow += warp_stride) {
        const long long bh = grow / seqLen;
        const int i = int(grow % seqLen);
        const float* q = Q + grow * D_HEAD;
        const float* kbase = K + bh * seqLen * D_HEAD;
        const float* vbase = V + bh * seqLen * D_HEAD;
        const int n_keys = is_causal ? (i + 1) : seqLen;

        float m = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
        for (int c = 0; c < DPL; ++c) acc[c] = 0.f;

        for (int j0 = 0; j0 < n_keys; j0 += WARP) {
            const int j = j0 + lane;
            float s = -INFINITY;
            if (j < n_keys) {
                const float* kr = kbase + (long long)j * D_HEAD;
                float dot = 0.f;
#pragma unroll 8
                for (int d = 0; d < D_HEAD; ++d) dot = fmaf(q[d], kr[d], dot);
                s = dot * scale;
            }
            float mx = s;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m, mx);          // finite: lane 0 of every chunk is a valid key
            const float f = expf(m - m_new);           // exp(-inf) = 0 on the first chunk
            const float e = expf(s - m_new);
            float rs = e;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            l = l * f + rs;
            m = m_new;
#pragma unroll
            for (int c = 0; c < DPL; ++c) acc[c] *= f;
            const int cnt = min(WARP, n_keys - j0);
            for (int jj = 0; jj < cnt; ++jj) {
                const float pj = __shfl_sync(0xffffffffu, e, jj);
                const float* vr = vbase + (long long)(j0 + jj) * D_HEAD;
#pragma unroll
                for (int c = 0; c < DPL; ++c) {
                    const int d = lane + c * WARP;
                    if (d < D_HEAD) acc[c] = fmaf(pj, vr[d], acc[c]);
                }
            }
        }
        const float inv = 1.0f / l;
#pragma unroll
        for (int c = 0; c < DPL; ++c) {
            const int d = lane + c * WARP;
            if (d < D_HEAD) O[grow * D_HEAD + d] = acc[c] * inv;
        }
    }
}
