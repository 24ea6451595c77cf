/* fa_b200.h — C ABI of the B200-native fused attention forward path.
 *
 * This is the drop-in boundary for the one hot path of GMichailov/Flash-Attention-CUDA-C:
 * the fused attention forward kernel behind kernels/FlashAttention.cuh.  The reference has no C ABI of
 * its own — its interface is a header-only __global__ template that the caller launches itself
 * (reference: kernels/FlashAttention.cuh:59-63, only caller tests/main.cu:60-61).  Each entry point below
 * names the reference interface it stands in for.  Plain pointers and sizes only; no C++/torch types.
 *
 * Semantics (reference: check.py:4-25, tests/main.cu:73-91):
 *   for batch b, query head h (kv head h / (Hq/Hkv)), query row i:
 *     s_ij = scale * sum_d Q[b,h,i,d] * K[b,hkv,j,d];   causal: s_ij = -inf for j > i + (Nk - Nq)
 *     O[b,h,i,:] = sum_j softmax_j(s_ij) * V[b,hkv,j,:]          (softmax and accumulation in fp32)
 *     LSE[b,h,i] = log sum_j exp(s_ij)                           (optional)
 * Default layout is the reference's: contiguous [B, H, N, d] (reference: loaders.cuh:57 row*D_HEAD addressing).
 *
 * All functions return FA_OK (0) or a negative error code; none of them calls exit() or assert()
 * (the reference's caller used CUDA_CHECK -> exit(1), tests/main.cu:12-19, and helpers.hpp:34 asserts).
 * All device work is enqueued on the given stream (NULL = default stream); nothing synchronises unless
 * stated.  Threading: every entry point may be called from any number of host threads, on any device and any
 * stream, concurrently.  The state the library keeps is listed here in full:
 *   - per device, behind a mutex: the property cache, and one 4-byte work-item counter per stream that has
 *     launched the persistent kernel (per thread for cudaStreamPerThread; one per launch recorded into a CUDA
 *     graph) — two launches that can be in flight together never share a counter, and a launch leaves its
 *     counter zero, so there is no per-launch memset;
 *   - per device, behind its own mutex: the three streams and staging buffers of fa_fwd_host (calls on one
 *     device take turns, calls on different devices run side by side);
 *   - per host thread: the last-error string and a small cache of TMA descriptors keyed by
 *     (pointer, dtype, shape, strides);
 *   - process-wide atomics: the launch counter and the fa_set_sm_reserve value.
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of Q, K, V, O */
enum { FA_DTYPE_F32 = 0, FA_DTYPE_F16 = 1, FA_DTYPE_BF16 = 2 };

/* return codes */
enum {
    FA_OK = 0,
    FA_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, Hq % Hkv != 0, misaligned pointer/stride */
    FA_ERR_UNSUPPORTED = -2,      /* head dim / dtype combination without a kernel */
    FA_ERR_CUDA = -3,             /* a CUDA runtime / driver call failed; see fa_last_error() */
    FA_ERR_NOT_B200 = -4          /* the current device is not sm_100 (no fallback path exists) */
};

/* ---- the hot path -------------------------------------------------------------------------------------------
 * fa_fwd: fused attention forward on device buffers, contiguous [B,Hq,Nq,d] Q/O and [B,Hkv,Nk,d] K/V.
 * Replaces: launching twoLoaderMhaFlashAttentionKernel<D,QT,KVT><<<grid,block,smem>>>(Q,K,V,O,B,H,N,scale,causal)
 *           (reference: kernels/FlashAttention.cuh:59-63; launch contract tests/main.cu:51-61).
 *   dtype FA_DTYPE_BF16 / FA_DTYPE_F16: d in {64, 128}  -> TMA + tcgen05/TMEM kernel
 *   dtype FA_DTYPE_F32:                 d % 16 == 0, d <= 128 -> exact-fp32 kernel
 *   lse: optional device pointer to B*Hq*Nq floats, or NULL (the reference's intended signature carried
 *        L / M pointers: kernels/FlashAttention.cuh:21,36,50).
 *   scale <= 0 selects 1/sqrt(d) (reference: tests/main.cu:27, check.py:19).
 */
int fa_fwd(const void* Q, const void* K, const void* V, void* O, float* lse,
           int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, float scale, int causal, void* stream);

/* fa_fwd_strided: same, with explicit element strides {batch, head, row} per tensor (innermost stride 1).
 * Replaces: the reference's intended strided signature (kernels/FlashAttention.cuh:23-25, commented).
 * strides = {q_b,q_h,q_n, k_b,k_h,k_n, v_b,v_h,v_n, o_b,o_h,o_n}.  Enables check.py's [B,N,H*d] layout
 * (reference: check.py:14-16) without a transpose.  For 16-bit dtypes strides must be multiples of 8.
 */
int fa_fwd_strided(const void* Q, const void* K, const void* V, void* O, float* lse,
                   int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, float scale, int causal,
                   const long long* strides, void* stream);

/* fa_mha_fwd_f32: the reference kernel's argument list as a C function; the template parameter D_HEAD becomes
 * the runtime argument d_head (Q_TILE_ROWS / KV_TILE_ROWS are chosen by the library).
 * Replaces: twoLoaderMhaFlashAttentionKernel<D_HEAD,QT,KVT>(Q,K,V,O,batchSize,numHeads,seqLen,scale,is_causal)
 *           (reference: kernels/FlashAttention.cuh:59-63) for callers that do not want to pick a launch shape.
 */
int fa_mha_fwd_f32(const float* Q, const float* K, const float* V, float* O,
                   int batchSize, int numHeads, int seqLen, int d_head, float scale, int is_causal, void* stream);

/* fa_fwd_host: end-to-end call on HOST buffers (same layout as fa_fwd): H2D copies, kernel, D2H copy,
 * pipelined over batch*head chunks on internal streams; synchronises before returning.
 * Replaces: the cudaMalloc / cudaMemcpy H2D / launch / cudaDeviceSynchronize / cudaMemcpy D2H sequence of the
 *           reference's driver (reference: tests/main.cu:39-66; main.cpp:30-33 is the intended home).
 * Pinned host memory gives full PCIe bandwidth; pageable memory works but copies serialise.
 */
int fa_fwd_host(const void* hQ, const void* hK, const void* hV, void* hO, float* hlse,
                int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, float scale, int causal);

/* ---- ring-KV building blocks (sequence-sharded long context) ---------------------------------
 * fa_merge_partial: acc (fp32 O [rows,d], lse [rows]) <- combine(acc, partial (16-bit O, fp32 lse)) over
 * disjoint key ranges.  Initialise acc_lse to -inf and acc_o to 0.  No reference counterpart (the reference
 * has no multi-device code); the carry it implements is the (running_max, running_l) recurrence of
 * reference utils.cuh:63-80 applied across ring steps instead of across tiles.
 */
int fa_merge_partial(float* acc_o, float* acc_lse, const void* part_o, const float* part_lse,
                     long long rows, int d, int dtype, void* stream);
/* fa_fwd_carry: one ring step with the merge fused into the attention kernel's epilogue: the partial result of
 * attention(Q, K, V) over this call's key range is folded directly into the running pair (acc_o fp32 [B,Hq,Nq,d]
 * contiguous, acc_lse fp32 [B,Hq,Nq]); no 16-bit O is written.  Start from acc_o = 0, acc_lse = -inf; finish with
 * fa_cast_out.  16-bit dtypes only.  qkv_strides = {q_b,q_h,q_n, k_b,k_h,k_n, v_b,v_h,v_n} or NULL for contiguous.
 * This is the "carry-in / carry-out (O, m, l)" kernel mode the reference's intended signature hinted at with its
 * L / M pointers (reference: kernels/FlashAttention.cuh:21,36,50; archive/archive.cu:34-42). */
int fa_fwd_carry(const void* Q, const void* K, const void* V, float* acc_o, float* acc_lse,
                 int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, float scale, int causal,
                 const long long* qkv_strides, void* stream);
/* fa_fwd_carry_window: fa_fwd_carry where this call's Nq query rows are rows [acc_row_offset, acc_row_offset + Nq) of a
 * running pair that holds acc_rows rows per (batch, head) (acc_o fp32 [B,Hq,acc_rows,d], acc_lse fp32 [B,Hq,acc_rows]).
 * A ring step in which only part of the rank's query rows see the arriving keys (zig-zag causal layout: the late half
 * only) is then ONE launch into the rank's single accumulator — no per-part accumulators, no second launch. */
int fa_fwd_carry_window(const void* Q, const void* K, const void* V, float* acc_o, float* acc_lse, int acc_rows, int acc_row_offset,
                        int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, float scale, int causal,
                        const long long* qkv_strides, void* stream);
/* fa_cast_out: fp32 accumulator -> 16-bit output tensor (n elements, n even). */
int fa_cast_out(const float* src, void* dst, long long n, int dtype, void* stream);

/* fa_workspace_bytes: device scratch a call of this shape needs FROM THE CALLER — always 0.  The kernels keep their state in
 * shared memory and TMEM, the TMA descriptors travel as kernel parameters, the 4-byte work-item counters are the library's
 * own (one per device and stream), and ring steps accumulate into the caller's (acc_o, acc_lse) pair.  Exists so that a host
 * driver written "ask, allocate, call" (the reference's main.cpp:30-33 was to size its buffers before the launch) needs no
 * special case.  Returns 0, or a negative FA_ERR_* code for a shape fa_fwd would reject. */
int fa_workspace_bytes(int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype);

/* ---- host helpers (reference: helpers.hpp:8-36, main.cpp:5-26) ----------------------------------------------*/
typedef struct {
    int cc_major, cc_minor, sm_count;
    size_t global_mem_bytes, smem_per_block_optin, smem_per_sm;
    int regs_per_sm, warp_size, l2_bytes, max_threads_per_sm;
} fa_device_info_t;
/* Replaces check_gpu_props() (reference: main.cpp:5-26) — fills a struct instead of printing. */
int fa_device_info(int device, fa_device_info_t* out);
/* Replace calculateSizeBlockQ / calculateSizeBlockKV (reference: helpers.hpp:8-30): rows per CTA and per KV tile. */
int fa_block_q(int d, int dtype);
int fa_block_kv(int d, int dtype);
/* The measured tile table behind them (the reference sketches register- and L2-driven formulas and returns 64,
 * helpers.hpp:8-30): one row per (head dim, causal, key-length bucket) with the kernel variant that measured fastest on
 * B200.  A row applies to Nk >= n_min; of the matching rows the one with the largest n_min wins.  The launcher uses it. */
typedef struct {
    int d, causal, n_min;      /* key */
    int block_q, block_kv;     /* query rows per CTA and work item (2 MMA tiles of 128), key rows per pipeline stage */
    int stages;                /* K/V ring slots in shared memory (cta_group 2: slots of half a tile) */
    int softmax_warps;         /* 8: one score row per thread; 16: 16-lane TMEM fragments, 4 softmax warps per SM sub-partition */
    int emu_pairs_per_8;       /* of every 8 score pairs, this many take exp2 on the FMA pipe instead of MUFU.EX2 */
    int staged_epilogue;       /* 1: O leaves through shared memory and TMA stores (d = 128: `stages` is 4 then), 0: row-per-lane st.global */
    int issuer_by_type;        /* MMA issuer warps split by type (all QK^T / all PV) instead of by query tile */
    int cta_group;             /* 1: single-CTA tcgen05.mma; 2: CTA pairs (clusters of 2, tcgen05 cta_group::2, d = 128): the two CTAs of a
                                * pair share each K/V tile half and half and work on 2 x block_q query rows per item */
    float tflops;              /* measured for the bucket's representative shape (profiles/r2_tile_sweep.jsonl); 0 = not measured */
} fa_tile_choice_t;
int fa_tile_table(const fa_tile_choice_t** rows);                                              /* returns the row count */
/* The row fa_fwd runs for a LARGE launch with Hq == Hkv.  Two launcher rules sit on top of the table (both measured, DESIGN.md
 * §3.1b): with an even number of query heads per kv group the d = 128 pair kernel is used at every length (its work items are
 * then two or four heads of one kv group over the same rows, so nothing is lost on causal diagonals), and a small launch whose tail the cta_group 1 kernel would smooth with
 * 128-row half items (at most three waves of 256-row blocks, last wave under half full) stays on the cta_group 1 kernel of
 * the same row. */
int fa_choose_tile(int d, int dtype, int causal, int nq, int nk, fa_tile_choice_t* out);
/* fa_choose_kernel: exactly what fa_fwd launches for this problem on the current device — the table row with the two rules
 * applied (and the reserve set by fa_set_sm_reserve): `tile` as it runs (cta_group, stages, epilogue), the query heads that
 * share one work item of the pair kernel (1, 2 or 4; 1 for cta_group 1) and the number of work items in the queue.  Pure host
 * arithmetic; without a GPU it assumes a B200's 148 SMs. */
typedef struct {
    fa_tile_choice_t tile;
    int heads_per_item;
    long long work_items;
} fa_kernel_choice_t;
int fa_choose_kernel(int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, int causal, fa_kernel_choice_t* out);
/* Replaces getNumCta (reference: helpers.hpp:33-36): CTAs along the query axis; ragged sizes round up, no assert. */
int fa_num_cta(int q_dim, int q_block_size);

/* fa_set_sm_reserve: the persistent attention kernel normally occupies every SM (one CTA per SM, all of its registers
 * and shared memory).  A ring step overlaps it with an NCCL send/recv kernel that needs a few SMs of its own; reserving
 * `sms` SMs (process-wide, default 0) lets that kernel run beside the MMAs instead of behind them. */
int fa_set_sm_reserve(int sms);

/* ---- diagnostics -------------------------------------------------------------------------------------------*/
const char* fa_last_error(void);          /* thread-local text of the last failure, "" if none */
long long fa_launch_count(void);          /* kernels this library has launched since load (bench: gpu_launches) */
const char* fa_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
