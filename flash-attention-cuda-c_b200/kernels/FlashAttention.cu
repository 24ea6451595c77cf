// FlashAttention.cu — the launcher: C-ABI implementation of include/fa_b200.h.
//
// The reference's kernels/FlashAttention.cu is a translation unit that only includes the header
// (reference: kernels/FlashAttention.cu:1, everything else commented out) and leaves grid/block/smem
// to the caller (reference: tests/main.cu:51-61).  Here the launcher owns argument validation, TMA
// descriptor construction, dynamic shared-memory opt-in, grid shape and kernel dispatch.
// There is no CPU fallback: on a device that is not sm_100 every entry point fails with FA_ERR_NOT_B200.
#include "FlashAttention.cuh"
#include "../../include/fa_b200.h"

#include <cuda_fp16.h>
#include <atomic>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_sm_reserve{0};       // fa_set_sm_reserve: SMs the persistent kernel leaves unoccupied
unsigned long long* g_prof = nullptr;   // device buffer of phase counters; only set by fa_debug_set_profile_buffer

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define FA_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(FA_ERR_CUDA, "%s -> %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ---- device checks -------------------------------------------------------------------------------
int check_device() {
    int dev = 0;
    FA_CUDA(cudaGetDevice(&dev));
    static std::mutex mu;
    static int cached[64];   // 0 unknown, 1 ok, -1 not sm_100
    if (dev < 0 || dev >= 64) return fail(FA_ERR_INVALID_ARGUMENT, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);
    if (cached[dev] == 0) {
        int major = 0;
        FA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        cached[dev] = (major == 10) ? 1 : -1;
    }
    if (cached[dev] < 0)
        return fail(FA_ERR_NOT_B200, "device %d is not compute capability 10.x; this library has no other path", dev);
    return FA_OK;
}

// ---- TMA descriptors -------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [B, H, N, d] tensor with element strides (sb, sh, sn, 1); box = 64 columns x box_rows rows (128 for the tile loads, 32 for the
// per-warp stores of the staged epilogue), 128B swizzle.
// Encoding a map is a pure function of (base, dtype, shape, strides), so the last few are kept per host thread: a caller
// that launches the same tensors again (a decode loop, a benchmark, a CUDA-graph-less training step) pays the three driver
// calls once.  Launch-bound shapes (BASELINE configs[1]: one 20 us kernel) are where this shows.
struct MapKey {
    const void* base; int dtype, B, H, N, d, box_rows; long long sb, sh, sn;
    bool operator==(const MapKey& o) const {
        return base == o.base && dtype == o.dtype && B == o.B && H == o.H && N == o.N && d == o.d && box_rows == o.box_rows && sb == o.sb && sh == o.sh && sn == o.sn;
    }
};
struct MapCache {
    static constexpr int kEntries = 24;
    MapKey key[kEntries] = {};
    CUtensorMap map[kEntries];
    bool valid[kEntries] = {};
    unsigned next = 0;
};
thread_local MapCache g_maps;

int make_tile_map(CUtensorMap* m, const void* base, int dtype, int B, int H, int N, int d, long long sb, long long sh,
                  long long sn, int box_rows = fa::kBlockN) {
    const MapKey key{base, dtype, B, H, N, d, box_rows, sb, sh, sn};
    for (int i = 0; i < MapCache::kEntries; ++i)
        if (g_maps.valid[i] && g_maps.key[i] == key) { *m = g_maps.map[i]; return FA_OK; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(FA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const CUtensorMapDataType dt = dtype == FA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sn * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    // A size-1 dimension's stride is never used for addressing but must still be a legal value.
    if (H == 1) strides[1] = strides[0] * (cuuint64_t)N;
    if (B == 1) strides[2] = strides[1] * (cuuint64_t)H;
    cuuint32_t box[4] = {(cuuint32_t)fa::kHalfCols, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    const unsigned slot = g_maps.next++ % MapCache::kEntries;
    g_maps.key[slot] = key; g_maps.map[slot] = *m; g_maps.valid[slot] = true;
    return FA_OK;
}

// ---- work-item counters for the persistent kernel ------------------------------------------------------
// The kernel's dynamic scheduler needs one int that is zero when the launch starts.  The kernel leaves it zero itself (the
// CTA that makes the launch's last claim resets it, loaders.cuh), so there is no per-launch memset; what the host has to
// guarantee is that two launches that may be in flight at the same time never share a counter:
//   * eager launches: one counter per (device, stream) — launches on one stream are serialised by the stream;
//     cudaStreamPerThread names a different stream in every host thread, so it gets a counter per thread;
//   * launches recorded into a CUDA graph: a counter of their own each (a replay runs on whatever stream the graph is
//     launched on, possibly beside eager launches on the stream it was captured from), never reused.
// Counters come from zero-filled chunks that are never freed (4 bytes per stream / captured launch).
struct CounterPool {
    static constexpr int kChunk = 4096;
    std::mutex mu;
    std::unordered_map<cudaStream_t, int*> by_stream;
    int* chunk = nullptr;
    int used = kChunk;
    int sms = 0;
    cudaStream_t helper = nullptr;
    // next zeroed int of the pool; mu held.  Safe while some stream of this thread is capturing (relaxed capture mode for
    // the allocation; the fill runs on a private non-blocking stream).
    int* take() {
        if (used == kChunk) {
            cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
            cudaThreadExchangeStreamCaptureMode(&mode);
            int* c = nullptr;
            bool ok = cudaMalloc(&c, kChunk * sizeof(int)) == cudaSuccess;
            if (ok && !helper) ok = cudaStreamCreateWithFlags(&helper, cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaMemsetAsync(c, 0, kChunk * sizeof(int), helper) == cudaSuccess && cudaStreamSynchronize(helper) == cudaSuccess;
            cudaThreadExchangeStreamCaptureMode(&mode);
            if (!ok) return nullptr;
            chunk = c;
            used = 0;
        }
        return chunk + used++;
    }
};
CounterPool g_counters[64];

int* next_counter(cudaStream_t st, int* sm_count_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    CounterPool& pool = g_counters[dev];
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(pool.mu);
    if (!pool.sms) cudaDeviceGetAttribute(&pool.sms, cudaDevAttrMultiProcessorCount, dev);
    *sm_count_out = pool.sms;
    if (cap != cudaStreamCaptureStatusNone) return pool.take();
    if (st == cudaStreamPerThread) {
        thread_local int* mine[64] = {};
        if (!mine[dev]) mine[dev] = pool.take();
        return mine[dev];
    }
    if (st == cudaStreamLegacy) st = nullptr;
    auto it = pool.by_stream.find(st);
    if (it != pool.by_stream.end()) return it->second;
    int* c = pool.take();
    if (c) pool.by_stream.emplace(st, c);
    return c;
}

// ---- measured tile table -----------------------------------------------------------------------------
// Stands where the reference's calculateSizeBlockQ / calculateSizeBlockKV sketch register- and L2-driven formulas and then
// return 64 (reference: helpers.hpp:8-30): per (head dim, causal, key-length bucket) the kernel variant that measured
// fastest on B200 (scripts/tile_sweep.py -> profiles/r2_tile_sweep.jsonl: 28 shapes x the compiled variants, four
// interleaved rounds of 200 ms of back-to-back launches each under the power cap; `tflops` is that run's figure).  fp16 takes the bf16 rows (same cycle counts).
// Tile geometry is the same in every row — 256 query rows per work item (2 x 128-row MMA tiles ping-ponged through the tensor
// pipe), 128 key rows per pipeline stage, the whole TMEM and shared memory of an SM — because the sweeps that varied it
// lost: 64-key steps run the SS MMA at half rate.  What varies is the softmax layout (8 warps x one row per thread / 16 warps
// x 16-lane fragments), the share of exponentials moved to the FMA pipe, how O leaves the SM (row-per-lane st.global, or
// staged per warp through shared memory + TMA stores, which at d = 128 costs the 1-CTA K/V ring its fifth slot and still
// wins), and whether the MMAs are 1-CTA or CTA-pair (cta_group 2: fwdSm100PairKernel — two CTAs on the SMs of one TPC share
// every K/V tile half and half, 512 query rows per pair item, six 16 KiB ring slots; block_q stays the rows per CTA).
const fa_tile_choice_t kTileTable[] = {
    //  d  causal n_min  block_q block_kv stages sm_warps emu staged issuer cta  tflops (first bucket of the row: N = n_min, or 512)
    {128, 0,     0,  256, 128, 6,  8, 0, 1, 1, 2,  910.6f},   // CTA pairs at every length: +3.8 % over the staged 1-CTA kernel at N = 512, +4.0 % at 2K, +4.2 % at 8K, +3.0 % at 32K
    {128, 1,     0,  256, 128, 4,  8, 0, 1, 1, 1,  537.4f},   // causal below 8K: 512-row pair items are too coarse (-7 % at 512, -4 % at 2K, -0.5 % at 4K)
    {128, 1,  8192,  256, 128, 6,  8, 0, 1, 1, 2, 1196.5f},   // +1.7 % at 8K, +3.2 % at 16K, +3.5 % at 32K
    { 64, 0,     0,  256, 128, 8,  8, 0, 1, 0, 1,  652.2f},   // staged epilogue: +5.7 % at N = 512, +3.4 % at 1024 (BASELINE configs[1]), +1.9 % at 2K
    { 64, 0,  4096,  256, 128, 8, 16, 1, 0, 0, 1,  816.6f},   // MUFU-bound: 16 softmax warps + 1/8 of the exponentials on the FMA pipe, +0.4 % at 4K .. +2.0 % at 32K
    { 64, 1,     0,  256, 128, 8,  8, 0, 0, 0, 1,  396.9f},   // staged: -1.3 % at 512, equal from 1K to 4K
    { 64, 1,  8192,  256, 128, 8, 16, 1, 0, 0, 1,  790.0f},   // +0.6 % at 8K, +1.7 % at 16K, +2.2 % at 32K
};
constexpr int kTileRows = (int)(sizeof(kTileTable) / sizeof(kTileTable[0]));
std::atomic<int> g_force_sw{0}, g_force_emu{0}, g_force_stg{0}, g_force_cg{0}, g_pair_heads{4}, g_half_items{1}, g_split_half{1};   // fa_debug_force_variant / fa_debug_half_items (A/B tooling)

const fa_tile_choice_t* choose_tile(int d, int causal, int nk) {
    const fa_tile_choice_t* best = nullptr;
    for (int i = 0; i < kTileRows; ++i) {
        const fa_tile_choice_t& r = kTileTable[i];
        if (r.d == d && r.causal == (causal ? 1 : 0) && nk >= r.n_min && (!best || r.n_min >= best->n_min)) best = &r;
    }
    return best;
}

int device_sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    static std::atomic<int> sms[64];
    int v = sms[dev & 63].load();
    if (!v) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) {
            cudaGetLastError();
            return 148;      // no device to ask (fa_choose_kernel on a CPU-only box): a B200's count
        }
        sms[dev & 63].store(v);
    }
    return v;
}

// ---- tcgen05 path ----------------------------------------------------------------------------------
// Work-item plan of a launch: query blocks are queued as 256-row items in whole waves of max_ctas; when the remainder would
// leave more than half of the SMs idle for a whole item, it is queued as 128-row half items instead.  Launches of many
// waves are left alone: their tail is already short against the rest.
struct ItemPlan { int num_q_blocks, n_full, total, max_ctas; int* counter; };

// (mul, shr) with  n / d == umulhi(n, mul) >> shr  for every 0 <= n < 2^31 and 2 <= d < 2^31 (round-up method:
// s = ceil(log2 d), mul = ceil(2^(31+s) / d) < 2^32, shr = s - 1); d == 1 is marked with shr = 32 (fast_div returns n).
// tests/test_abi.py checks it against integer division through fa_debug_fast_div.
void make_fast_div(unsigned d, unsigned* mul, unsigned* shr) {
    if (d <= 1) { *mul = 0; *shr = 32; return; }
    unsigned s = 0;
    while ((1ull << s) < d) ++s;
    *mul = (unsigned)(((1ull << (31 + s)) + d - 1) / d);
    *shr = s - 1;
}

// the arithmetic of the plan (host only; fa_debug_plan_counts exposes it to the CPU tests)
void plan_counts(long long blocks, long long max_ctas, bool half_items, long long* n_full, long long* total) {
    *n_full = blocks;
    const long long waves = blocks / max_ctas, rem = blocks - waves * max_ctas;
    if (half_items && waves <= 3 && rem > 0 && 2 * rem <= max_ctas) *n_full = waves * max_ctas;
    *total = *n_full + 2 * (blocks - *n_full);
}

int plan_items(fa::FwdParams& p, cudaStream_t st, ItemPlan* plan) {
    const int rows_per_item = fa::kTilesPerCta * fa::kBlockM;
    plan->num_q_blocks = (p.Nq + rows_per_item - 1) / rows_per_item;
    const long long blocks = (long long)plan->num_q_blocks * p.Hq * p.B;
    if (blocks > 0x3fffffffLL - 4096) return fail(FA_ERR_INVALID_ARGUMENT, "too many work items (%lld)", blocks);
    int sm_count = 0;
    plan->counter = next_counter(st, &sm_count);
    if (!plan->counter) return fail(FA_ERR_CUDA, "work-item counter allocation failed");
    int max_ctas = sm_count - g_sm_reserve.load();      // SMs left free for a concurrent communication kernel
    if (max_ctas < 1) max_ctas = 1;
    long long n_full = 0, total = 0;
    plan_counts(blocks, max_ctas, g_half_items.load() != 0, &n_full, &total);
    plan->n_full = (int)n_full;
    plan->total = (int)total;
    plan->max_ctas = max_ctas;
    return FA_OK;
}

template <int D, int STAGES, int DT, bool OVEC32, int SW, int EMU, int ST, int HS>
int launch_sm100(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, fa::FwdParams p,
                 const ItemPlan& plan, cudaStream_t st) {
    using L = fa::SmemLayout<D, STAGES>;
    auto kern = fa::fwdSm100Kernel<D, STAGES, DT, OVEC32, SW, EMU, ST, HS>;
    constexpr int kSmem = ST ? L::kBytesStaged : L::kDynamicBytes;
    // the dynamic shared-memory opt-in is per device
    int dev = 0;
    cudaGetDevice(&dev);
    static std::atomic<unsigned long long> dev_mask{0};
    if (!(dev_mask.load() & (1ull << dev))) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return fail(FA_ERR_CUDA, "cudaFuncSetAttribute(smem=%d) -> %s", kSmem, cudaGetErrorString(e));
        dev_mask.fetch_or(1ull << dev);
    }
    p.num_q_blocks = plan.num_q_blocks;
    make_fast_div((unsigned)p.num_q_blocks, &p.div_qblocks_mul, &p.div_qblocks_shr);
    make_fast_div((unsigned)p.Hq, &p.div_hq_mul, &p.div_hq_shr);
    make_fast_div((unsigned)p.q_heads_per_kv, &p.div_group_mul, &p.div_group_shr);
    p.sched_counter = plan.counter;
    p.n_full_items = plan.n_full;
    p.total_items = plan.total;
    p.split_half = HS;      // half items on both query-tile slots (split-KV) or on slot 0 alone
    const int grid = p.total_items < plan.max_ctas ? p.total_items : plan.max_ctas;   // persistent: one CTA per SM
    kern<<<grid, fa::KCfg<SW>::kNumThreads, kSmem, st>>>(tq, tk, tv, to, p);
    g_launches.fetch_add(1);
    FA_CUDA(cudaGetLastError());
    return FA_OK;
}

// compiled variants: (softmax warps, exp2 share on the FMA pipe, staged TMA-store epilogue) = (8,0,0) (8,0,1) (16,1,0); the staged
// epilogue takes its 32 KiB of shared memory from the K/V ring at d = 128 (4 slots instead of 5) and from spare room at d = 64.
// The two 8-warp variants exist a second time with the split-KV half-item code (HS = 1): launches whose plan has half items
// get that build (plain mode only), all others the build without it.
template <int D, int DT, bool OVEC32>
int launch_variant(int sw, int emu, int stg, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                   fa::FwdParams& p, cudaStream_t st) {
    (void)emu;
    ItemPlan plan;
    if (int rc = plan_items(p, st, &plan)) return rc;
    // split-KV half items: plain mode, and only when no step of the launch needs a mask (the HS kernels keep the key loop of
    // the plain ones: a slot's key-tile numbering is not known to its softmax warps)
    const bool hs = plan.total > plan.n_full && p.acc_o == nullptr && !p.causal && p.Nk % fa::kBlockN == 0 && g_split_half.load() != 0;
    constexpr int kStages = D == 128 ? 5 : 8;
    constexpr int kStagesStaged = D == 128 ? 4 : 8;
    if (sw == 16) return launch_sm100<D, kStages, DT, OVEC32, 16, 1, 0, 0>(tq, tk, tv, to, p, plan, st);
    if (stg) return hs ? launch_sm100<D, kStagesStaged, DT, OVEC32, 8, 0, 1, 1>(tq, tk, tv, to, p, plan, st)
                       : launch_sm100<D, kStagesStaged, DT, OVEC32, 8, 0, 1, 0>(tq, tk, tv, to, p, plan, st);
    return hs ? launch_sm100<D, kStages, DT, OVEC32, 8, 0, 0, 1>(tq, tk, tv, to, p, plan, st)
              : launch_sm100<D, kStages, DT, OVEC32, 8, 0, 0, 0>(tq, tk, tv, to, p, plan, st);
}

// query heads per pair item: 4 / 2 / 1 (= pairs cut by rows), capped by fa_debug_force_cta_group's A/B settings
int pair_heads_for(int q_heads_per_kv) {
    const int cap = g_pair_heads.load();
    if (q_heads_per_kv % 4 == 0 && cap >= 4) return 4;
    if (q_heads_per_kv % 2 == 0 && cap >= 2) return 2;
    return 1;
}

// What a 16-bit launch runs: the tile-table row for (d, causal, Nk) and, on top of it, the launcher's rules.  Pure host
// arithmetic (fa_choose_kernel exposes it; tests/test_abi.py checks it on the CPU).
struct KernelChoice { int sw, emu, stg, cg, heads_per_item; const fa_tile_choice_t* row; };
KernelChoice choose_kernel(int B, int Hq, int Hkv, int Nq, int Nk, int d, int causal, bool carry) {
    KernelChoice k;
    k.row = choose_tile(d, causal, Nk);
    k.sw = k.row ? k.row->softmax_warps : 8;
    k.emu = k.row ? k.row->emu_pairs_per_8 : 0;
    k.stg = k.row ? k.row->staged_epilogue : 0;
    if (g_force_sw.load()) { k.sw = g_force_sw.load(); k.emu = g_force_emu.load(); k.stg = g_force_stg.load(); }
    if (carry) k.stg = 0;      // carry mode folds into an fp32 accumulator in place: there is no 16-bit O to stage
    k.cg = k.row ? k.row->cta_group : 1;
    const bool forced = g_force_cg.load() != 0;
    const int hpi = pair_heads_for(Hq / Hkv);
    if (!forced && d == 128 && k.sw == 8) {
        // GQA with an even number of query heads per kv group: the pair kernel cuts its pairs by HEADS (same rows, same diagonal),
        // so the causal loss that keeps the table's causal rows below 8K on 1-CTA kernels does not exist: +4.4 .. +12 % at causal
        // 1K .. 32K (Hq/Hkv = 32/8; profiles/r2_sustained_gqa_pairs_by_four_heads.log)
        if (hpi > 1) k.cg = 2;
        // Small launches: whenever the 1-CTA plan would smooth its tail with half items (at most three waves of 256-row blocks and
        // a last wave that leaves more than half of the SMs idle, plan_counts), the 1-CTA kernel keeps the launch — the pair
        // kernel has no half items, and such launches gain nothing from pairing (96 pair items: -6 %, 256: -2 %;
        // profiles/r2_sustained_small_launch_pairs.log).  Everything else that the table or the head rule gives to pairs runs
        // on pairs (128 pair items without a half-item tail: +-0).
        if (k.cg == 2) {
            const long long blocks = (long long)((Nq + fa::kTilesPerCta * fa::kBlockM - 1) / (fa::kTilesPerCta * fa::kBlockM)) * Hq * B;
            int ctas = device_sm_count() - g_sm_reserve.load();
            if (ctas < 1) ctas = 1;
            long long n_full = 0, total = 0;
            plan_counts(blocks, ctas, g_half_items.load() != 0, &n_full, &total);
            if (n_full < blocks) k.cg = 1;
        }
    }
    if (forced) k.cg = g_force_cg.load();
    // A reserve of 8 or more SMs means a communication KERNEL runs beside the attention launch (NCCL send/recv in the ring's
    // p2p transport).  Its CTAs land on SMs of different TPCs, a TPC with one SM taken cannot hold a pair, and the pairs that
    // do not fit queue behind the communication kernel: measured on 2 GPUs, 15.4-16.1 ms per ring pass with pairs against
    // 13.7 ms with 1-CTA kernels.  So the pair kernel is only used with small reserves (copy-engine transports).
    if (g_sm_reserve.load() >= 8 && !forced) k.cg = 1;
    if (!(d == 128 && k.sw == 8)) k.cg = 1;      // the pair kernel exists for d = 128 with 8 softmax warps
    k.heads_per_item = k.cg == 2 ? hpi : 1;
    return k;
}

// CTA-pair kernel (d = 128, 8 softmax warps): clusters of two CTAs, 512-row work items, one claim per pair.
constexpr int kPairUnavailable = 1;      // launch_pair's answer on a device that cannot hold a single cluster of two CTAs
#ifndef FA_PAIR_STAGES
#define FA_PAIR_STAGES 6
#endif
constexpr int kPairStages = FA_PAIR_STAGES;      // 16 KiB ring slots: three K halves + three V halves in flight (8 fit and measure the same)
template <int DT, bool OVEC32, int ST>
int launch_pair(const CUtensorMap& tq, const CUtensorMap& tk64, const CUtensorMap& tv, const CUtensorMap& to, fa::FwdParams p, cudaStream_t st) {
    constexpr int D = 128;
    using L = fa::SmemLayout<D, kPairStages, 2>;
    auto kern = fa::fwdSm100PairKernel<D, kPairStages, DT, OVEC32, ST>;
    constexpr int kSmem = ST ? L::kBytesStaged : L::kDynamicBytes;
    int dev = 0;
    cudaGetDevice(&dev);
    static std::atomic<unsigned long long> dev_mask{0};
    static std::atomic<int> max_pairs_dev[64];
    if (!(dev_mask.load() & (1ull << dev))) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return fail(FA_ERR_CUDA, "cudaFuncSetAttribute(pair kernel, smem=%d) -> %s", kSmem, cudaGetErrorString(e));
        // how many pairs the device can hold at once (an SM whose TPC partner is unavailable cannot take one)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 128);
        cfg.blockDim = dim3(fa::KCfg<8>::kNumThreads);
        cfg.dynamicSmemBytes = kSmem;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) { cudaGetLastError(); n = 0; }      // e.g. a partition of the GPU on which no TPC has both SMs: no pairs here
        max_pairs_dev[dev & 63].store(n);
        if (getenv("FA_DEBUG_PAIRS")) fprintf(stderr, "fa_b200: device %d holds %d CTA pairs at once\n", dev, n);
        dev_mask.fetch_or(1ull << dev);
    }
    // pairs cut by four heads of a kv group where the group size allows it (no causal loss and no "later row block alone" step),
    // by two heads where it is even (no causal loss), else by rows: loaders.cuh, decode_pair_item
    const int hpi = pair_heads_for(p.q_heads_per_kv);
    const int item_rows = fa::kPairRows / hpi;
    const int item_heads = p.Hq / hpi;
    const int num_q_blocks = (p.Nq + item_rows - 1) / item_rows;
    const long long items = (long long)num_q_blocks * item_heads * p.B;
    if (items > 0x3fffffffLL - 4096) return fail(FA_ERR_INVALID_ARGUMENT, "too many work items (%lld)", items);
    int sm_count = 0;
    int* counter = next_counter(st, &sm_count);
    if (!counter) return fail(FA_ERR_CUDA, "work-item counter allocation failed");
    int max_pairs = max_pairs_dev[dev & 63].load();
    if (max_pairs < 1) return kPairUnavailable;      // the caller launches the 1-CTA kernel of the same tile-table row instead
    const int by_reserve = (sm_count - g_sm_reserve.load()) / 2;      // fa_set_sm_reserve: SMs left free (small reserves only, see fwd_impl)
    if (by_reserve < max_pairs) max_pairs = by_reserve;
    if (max_pairs < 1) max_pairs = 1;
    p.num_q_blocks = num_q_blocks;
    p.pair_heads = hpi;
    make_fast_div((unsigned)p.num_q_blocks, &p.div_qblocks_mul, &p.div_qblocks_shr);
    make_fast_div((unsigned)item_heads, &p.div_hq_mul, &p.div_hq_shr);
    make_fast_div((unsigned)p.q_heads_per_kv, &p.div_group_mul, &p.div_group_shr);
    p.sched_counter = counter;
    p.n_full_items = p.total_items = (int)items;
    p.split_half = 0;
    const int pairs = items < max_pairs ? (int)items : max_pairs;
    kern<<<2 * pairs, fa::KCfg<8>::kNumThreads, kSmem, st>>>(tq, tk64, tv, to, p);
    g_launches.fetch_add(1);
    FA_CUDA(cudaGetLastError());
    return FA_OK;
}

template <int D>
int launch_fp32(const fa::Fp32Params& p, cudaStream_t st) {
    using S = fa::Fp32Smem<D>;
    auto kern = fa::fwdFp32Kernel<D>;
    int dev = 0;
    cudaGetDevice(&dev);
    static std::atomic<unsigned long long> dev_mask{0};
    if (!(dev_mask.load() & (1ull << dev))) {
        FA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes));
        dev_mask.fetch_or(1ull << dev);
    }
    dim3 grid((p.Nq + fa::kF32Rows - 1) / fa::kF32Rows, p.Hq, p.B);
    kern<<<grid, fa::kF32Threads, S::kBytes, st>>>(p);
    g_launches.fetch_add(1);
    FA_CUDA(cudaGetLastError());
    return FA_OK;
}

int fwd_impl(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int Hq, int Hkv, int Nq, int Nk,
             int d, int dtype, float scale, int causal, const long long* s, cudaStream_t st,
             float* acc_o = nullptr, float* acc_lse = nullptr, int acc_rows = 0, int acc_off = 0) {
    const bool carry = acc_o != nullptr;
    if (carry && (!acc_lse || dtype == FA_DTYPE_F32)) return fail(FA_ERR_INVALID_ARGUMENT, "carry mode needs acc_lse and a 16-bit dtype");
    if (carry) O = acc_o;   // only used for the null / alignment checks below
    if (carry && acc_rows == 0) acc_rows = Nq;
    if (carry && (acc_off < 0 || acc_rows < Nq || acc_off > acc_rows - Nq))
        return fail(FA_ERR_INVALID_ARGUMENT, "carry window [%d, %d) does not fit %d accumulator rows", acc_off, acc_off + Nq, acc_rows);
    if (!Q || !K || !V || !O) return fail(FA_ERR_INVALID_ARGUMENT, "null tensor pointer");
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || Nk <= 0 || d <= 0)
        return fail(FA_ERR_INVALID_ARGUMENT, "non-positive size (B=%d Hq=%d Hkv=%d Nq=%d Nk=%d d=%d)", B, Hq, Hkv, Nq, Nk, d);
    if (Hq % Hkv != 0) return fail(FA_ERR_INVALID_ARGUMENT, "Hq=%d is not a multiple of Hkv=%d", Hq, Hkv);
    if (B > 65535 || Hq > 65535) return fail(FA_ERR_INVALID_ARGUMENT, "B and Hq must be <= 65535");
    if (dtype != FA_DTYPE_F32 && dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16)
        return fail(FA_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if (int rc = check_device()) return rc;
    const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)d);

    long long def[12];
    if (!s) {
        const long long qs[3] = {(long long)Hq * Nq * d, (long long)Nq * d, d};
        const long long ks[3] = {(long long)Hkv * Nk * d, (long long)Nk * d, d};
        for (int i = 0; i < 3; ++i) { def[i] = qs[i]; def[3 + i] = ks[i]; def[6 + i] = ks[i]; def[9 + i] = qs[i]; }
        s = def;
    }
    for (int i = 0; i < 12; ++i)
        if (s[i] <= 0) return fail(FA_ERR_INVALID_ARGUMENT, "stride %d must be positive", i);

    if (dtype == FA_DTYPE_F32) {
        if (d % 16 != 0 || d > 128) return fail(FA_ERR_UNSUPPORTED, "fp32 path needs d %% 16 == 0 and d <= 128 (got %d)", d);
        fa::Fp32Params p;
        p.Q = (const float*)Q; p.K = (const float*)K; p.V = (const float*)V; p.O = (float*)O; p.lse = lse;
        p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.Nq = Nq; p.Nk = Nk;
        p.q_sb = s[0]; p.q_sh = s[1]; p.q_sn = s[2]; p.k_sb = s[3]; p.k_sh = s[4]; p.k_sn = s[5];
        p.v_sb = s[6]; p.v_sh = s[7]; p.v_sn = s[8]; p.o_sb = s[9]; p.o_sh = s[10]; p.o_sn = s[11];
        p.scale = sc; p.causal = causal ? 1 : 0; p.causal_off = Nk - Nq; p.q_heads_per_kv = Hq / Hkv;
        switch (d) {
            case 16: return launch_fp32<16>(p, st);
            case 32: return launch_fp32<32>(p, st);
            case 48: return launch_fp32<48>(p, st);
            case 64: return launch_fp32<64>(p, st);
            case 80: return launch_fp32<80>(p, st);
            case 96: return launch_fp32<96>(p, st);
            case 112: return launch_fp32<112>(p, st);
            case 128: return launch_fp32<128>(p, st);
        }
        return fail(FA_ERR_UNSUPPORTED, "fp32 path: unsupported d=%d", d);
    }

    if (d != 64 && d != 128) return fail(FA_ERR_UNSUPPORTED, "16-bit path supports d in {64,128} (got %d)", d);
    const void* ptrs[4] = {Q, K, V, O};
    for (int i = 0; i < 4; ++i)
        if (reinterpret_cast<uintptr_t>(ptrs[i]) % 16 != 0) return fail(FA_ERR_INVALID_ARGUMENT, "tensor %d is not 16-byte aligned", i);
    for (int i = 0; i < 12; ++i)
        if (s[i] % 8 != 0) return fail(FA_ERR_INVALID_ARGUMENT, "stride %d (=%lld elements) must be a multiple of 8", i, s[i]);

    CUtensorMap tq, tk, tv;
    if (int rc = make_tile_map(&tq, Q, dtype, B, Hq, Nq, d, s[0], s[1], s[2])) return rc;
    if (int rc = make_tile_map(&tk, K, dtype, B, Hkv, Nk, d, s[3], s[4], s[5])) return rc;
    if (int rc = make_tile_map(&tv, V, dtype, B, Hkv, Nk, d, s[6], s[7], s[8])) return rc;

    fa::FwdParams p;
    p.O = O; p.lse = lse; p.acc_o = acc_o; p.acc_lse = acc_lse; p.acc_rows = acc_rows; p.acc_off = acc_off; p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.Nq = Nq; p.Nk = Nk;
    p.o_stride_b = s[9]; p.o_stride_h = s[10]; p.o_stride_n = s[11];
    p.scale = sc; p.scale_log2 = sc * 1.4426950408889634f;
    p.causal = causal ? 1 : 0; p.causal_off = Nk - Nq; p.q_heads_per_kv = Hq / Hkv;
    p.prof = g_prof;
    p.pair_heads = 0;

    // kernel variant: the measured tile table + the launcher's rules (choose_kernel), unless an A/B tool forces one
    const KernelChoice kc = choose_kernel(B, Hq, Hkv, Nq, Nk, d, causal, carry);
    const int sw = kc.sw, emu = kc.emu, stg = kc.stg, cg = kc.cg;

    // the staged epilogue writes O with TMA stores (16-byte alignment, checked above); otherwise 256-bit epilogue stores need
    // every output row to start 32-byte aligned (carry mode does not write O at all)
    CUtensorMap to = tq;
    if (stg && sw == 8) {
        if (int rc = make_tile_map(&to, O, dtype, B, Hq, Nq, d, s[9], s[10], s[11], 32)) return rc;      // one store per warp: 64 columns x 32 rows
    }
    const bool v32 = !carry && reinterpret_cast<uintptr_t>(O) % 32 == 0 && s[9] % 16 == 0 && s[10] % 16 == 0 && s[11] % 16 == 0;
    const bool bf = dtype == FA_DTYPE_BF16;
    if (cg == 2 && d == 128 && sw == 8) {
        // CTA pairs: each CTA loads 64 of a K tile's 128 keys (its own tensor map: 64-row boxes) and 64 of a V tile's columns
        CUtensorMap tk64;
        if (int rc = make_tile_map(&tk64, K, dtype, B, Hkv, Nk, d, s[3], s[4], s[5], 64)) return rc;
        int rc;
        if (stg) {
            if (v32) rc = bf ? launch_pair<fa::kBF16, true, 1>(tq, tk64, tv, to, p, st) : launch_pair<fa::kF16, true, 1>(tq, tk64, tv, to, p, st);
            else rc = bf ? launch_pair<fa::kBF16, false, 1>(tq, tk64, tv, to, p, st) : launch_pair<fa::kF16, false, 1>(tq, tk64, tv, to, p, st);
        } else {
            if (v32) rc = bf ? launch_pair<fa::kBF16, true, 0>(tq, tk64, tv, to, p, st) : launch_pair<fa::kF16, true, 0>(tq, tk64, tv, to, p, st);
            else rc = bf ? launch_pair<fa::kBF16, false, 0>(tq, tk64, tv, to, p, st) : launch_pair<fa::kF16, false, 0>(tq, tk64, tv, to, p, st);
        }
        if (rc != kPairUnavailable) return rc;
    }
    if (d == 128) {
        if (v32) return bf ? launch_variant<128, fa::kBF16, true>(sw, emu, stg, tq, tk, tv, to, p, st) : launch_variant<128, fa::kF16, true>(sw, emu, stg, tq, tk, tv, to, p, st);
        return bf ? launch_variant<128, fa::kBF16, false>(sw, emu, stg, tq, tk, tv, to, p, st) : launch_variant<128, fa::kF16, false>(sw, emu, stg, tq, tk, tv, to, p, st);
    }
    if (v32) return bf ? launch_variant<64, fa::kBF16, true>(sw, emu, stg, tq, tk, tv, to, p, st) : launch_variant<64, fa::kF16, true>(sw, emu, stg, tq, tk, tv, to, p, st);
    return bf ? launch_variant<64, fa::kBF16, false>(sw, emu, stg, tq, tk, tv, to, p, st) : launch_variant<64, fa::kF16, false>(sw, emu, stg, tq, tk, tv, to, p, st);
}

// ---- host-buffer pipeline state: one per device, each behind its own lock ------------------------------
struct HostPipe {
    static constexpr int kSlots = 3;
    std::mutex mu;
    cudaStream_t streams[kSlots] = {nullptr, nullptr, nullptr};
    void* buf[kSlots] = {nullptr, nullptr, nullptr};
    size_t cap[kSlots] = {0, 0, 0};
    bool ready = false;
};
HostPipe g_pipes[64];

}  // namespace

extern "C" {

// Debug hook (not in include/fa_b200.h): phase-counter buffer for FA_PHASE_PROFILE builds, 16 x u64 on the device.
void fa_debug_set_profile_buffer(void* dev_ptr) { g_prof = (unsigned long long*)dev_ptr; }

int fa_set_sm_reserve(int sms) {
    if (sms < 0) return fail(FA_ERR_INVALID_ARGUMENT, "sms must be >= 0");
    g_sm_reserve.store(sms);
    return FA_OK;
}

const char* fa_last_error(void) { return g_err; }
long long fa_launch_count(void) { return g_launches.load(); }
const char* fa_version(void) { return "fa_b200 0.1 (sm_100a, tcgen05/TMEM/TMA)"; }

int fa_fwd(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int Hq, int Hkv, int Nq, int Nk,
           int d, int dtype, float scale, int causal, void* stream) {
    g_err[0] = 0;
    return fwd_impl(Q, K, V, O, lse, B, Hq, Hkv, Nq, Nk, d, dtype, scale, causal, nullptr, (cudaStream_t)stream);
}

int fa_fwd_strided(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int Hq, int Hkv, int Nq,
                   int Nk, int d, int dtype, float scale, int causal, const long long* strides, void* stream) {
    g_err[0] = 0;
    if (!strides) return fail(FA_ERR_INVALID_ARGUMENT, "strides is null");
    return fwd_impl(Q, K, V, O, lse, B, Hq, Hkv, Nq, Nk, d, dtype, scale, causal, strides, (cudaStream_t)stream);
}

int fa_fwd_carry(const void* Q, const void* K, const void* V, float* acc_o, float* acc_lse, int B, int Hq, int Hkv,
                 int Nq, int Nk, int d, int dtype, float scale, int causal, const long long* qkv_strides, void* stream) {
    g_err[0] = 0;
    if (!acc_o || !acc_lse) return fail(FA_ERR_INVALID_ARGUMENT, "acc_o / acc_lse is null");
    long long s[12];
    const long long qs[3] = {(long long)Hq * Nq * d, (long long)Nq * d, d};
    const long long ks[3] = {(long long)Hkv * Nk * d, (long long)Nk * d, d};
    for (int i = 0; i < 3; ++i) { s[i] = qs[i]; s[3 + i] = ks[i]; s[6 + i] = ks[i]; s[9 + i] = qs[i]; }
    if (qkv_strides) for (int i = 0; i < 9; ++i) s[i] = qkv_strides[i];
    return fwd_impl(Q, K, V, nullptr, nullptr, B, Hq, Hkv, Nq, Nk, d, dtype, scale, causal, s, (cudaStream_t)stream, acc_o, acc_lse);
}

int fa_fwd_carry_window(const void* Q, const void* K, const void* V, float* acc_o, float* acc_lse, int acc_rows, int acc_row_offset,
                        int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, float scale, int causal,
                        const long long* qkv_strides, void* stream) {
    g_err[0] = 0;
    if (!acc_o || !acc_lse) return fail(FA_ERR_INVALID_ARGUMENT, "acc_o / acc_lse is null");
    long long s[12];
    const long long qs[3] = {(long long)Hq * Nq * d, (long long)Nq * d, d};
    const long long ks[3] = {(long long)Hkv * Nk * d, (long long)Nk * d, d};
    for (int i = 0; i < 3; ++i) { s[i] = qs[i]; s[3 + i] = ks[i]; s[6 + i] = ks[i]; s[9 + i] = qs[i]; }
    if (qkv_strides) for (int i = 0; i < 9; ++i) s[i] = qkv_strides[i];
    return fwd_impl(Q, K, V, nullptr, nullptr, B, Hq, Hkv, Nq, Nk, d, dtype, scale, causal, s, (cudaStream_t)stream, acc_o, acc_lse,
                    acc_rows, acc_row_offset);
}

int fa_mha_fwd_f32(const float* Q, const float* K, const float* V, float* O, int batchSize, int numHeads, int seqLen,
                   int d_head, float scale, int is_causal, void* stream) {
    g_err[0] = 0;
    return fwd_impl(Q, K, V, O, nullptr, batchSize, numHeads, numHeads, seqLen, seqLen, d_head, FA_DTYPE_F32, scale,
                    is_causal, nullptr, (cudaStream_t)stream);
}

int fa_fwd_host(const void* hQ, const void* hK, const void* hV, void* hO, float* hlse, int B, int Hq, int Hkv, int Nq,
                int Nk, int d, int dtype, float scale, int causal) {
    g_err[0] = 0;
    if (!hQ || !hK || !hV || !hO) return fail(FA_ERR_INVALID_ARGUMENT, "null host pointer");
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || Nk <= 0 || d <= 0 || Hq % Hkv != 0)
        return fail(FA_ERR_INVALID_ARGUMENT, "bad sizes");
    if (int rc = check_device()) return rc;
    const size_t es = dtype == FA_DTYPE_F32 ? 4 : 2;
    const int g = Hq / Hkv;
    // unit = one (batch, kv head): g query heads + 1 K head + 1 V head; units are contiguous in all four tensors
    const long long units = (long long)B * Hkv;
    const size_t q_unit = (size_t)g * Nq * d * es, kv_unit = (size_t)Nk * d * es, lse_unit = (size_t)g * Nq * 4;
    const size_t unit_bytes = 2 * q_unit + 2 * kv_unit + lse_unit;
    const size_t target = 96ull << 20;   // bytes per pipeline chunk
    long long upc = (long long)(target / unit_bytes);
    if (upc < 1) upc = 1;
    if (upc > units) upc = units;
    if (upc > 65535) upc = 65535;
    const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)d);

    int dev = 0;
    FA_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(FA_ERR_INVALID_ARGUMENT, "device index %d out of range", dev);
    HostPipe& pipe = g_pipes[dev];
    // one call at a time per device (the staging buffers are the device's); calls on different devices run side by side
    std::lock_guard<std::mutex> lk(pipe.mu);
    if (!pipe.ready) {
        for (int i = 0; i < HostPipe::kSlots; ++i) FA_CUDA(cudaStreamCreateWithFlags(&pipe.streams[i], cudaStreamNonBlocking));
        pipe.ready = true;
    }
    const size_t need = (size_t)upc * unit_bytes + 1024;
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        if (pipe.cap[i] < need) {
            if (pipe.buf[i]) FA_CUDA(cudaFree(pipe.buf[i]));
            pipe.buf[i] = nullptr; pipe.cap[i] = 0;
            FA_CUDA(cudaMalloc(&pipe.buf[i], need));
            pipe.cap[i] = need;
        }
    }
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    // one chunk: H2D of its Q/K/V units, the kernel, D2H of O (and LSE), all on the chunk's stream
    auto run_chunk = [&](long long u0, long long nu, int slot) -> int {
        cudaStream_t st = pipe.streams[slot];
        char* base = (char*)pipe.buf[slot];
        char* dQ = base;
        char* dK = dQ + al(nu * q_unit);
        char* dV = dK + al(nu * kv_unit);
        char* dO = dV + al(nu * kv_unit);
        float* dL = hlse ? (float*)(dO + al(nu * q_unit)) : nullptr;
        FA_CUDA(cudaMemcpyAsync(dQ, (const char*)hQ + u0 * q_unit, nu * q_unit, cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaMemcpyAsync(dK, (const char*)hK + u0 * kv_unit, nu * kv_unit, cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaMemcpyAsync(dV, (const char*)hV + u0 * kv_unit, nu * kv_unit, cudaMemcpyHostToDevice, st));
        // each unit is presented to the kernel as one "batch" entry with g query heads and one kv head
        if (int rc = fwd_impl(dQ, dK, dV, dO, dL, (int)nu, g, 1, Nq, Nk, d, dtype, sc, causal, nullptr, st)) return rc;
        FA_CUDA(cudaMemcpyAsync((char*)hO + u0 * q_unit, dO, nu * q_unit, cudaMemcpyDeviceToHost, st));
        if (hlse) FA_CUDA(cudaMemcpyAsync((char*)hlse + u0 * lse_unit, dL, nu * lse_unit, cudaMemcpyDeviceToHost, st));
        return FA_OK;
    };
    int rc = FA_OK, slot = 0;
    for (long long u0 = 0; u0 < units && rc == FA_OK; u0 += upc, slot = (slot + 1) % HostPipe::kSlots)
        rc = run_chunk(u0, (units - u0 < upc) ? (units - u0) : upc, slot);
    // drain every stream before returning, also on failure: earlier chunks may still be copying into the caller's buffers
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(pipe.streams[i]);
        if (e != cudaSuccess && rc == FA_OK) rc = fail(FA_ERR_CUDA, "cudaStreamSynchronize -> %s", cudaGetErrorString(e));
    }
    return rc;
}

int fa_merge_partial(float* acc_o, float* acc_lse, const void* part_o, const float* part_lse, long long rows, int d,
                     int dtype, void* stream) {
    g_err[0] = 0;
    if (!acc_o || !acc_lse || !part_o || !part_lse || rows <= 0 || d <= 0 || d % 2)
        return fail(FA_ERR_INVALID_ARGUMENT, "bad merge arguments");
    if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16) return fail(FA_ERR_UNSUPPORTED, "merge takes 16-bit partials");
    if (int rc = check_device()) return rc;
    const int wpb = 8;
    const unsigned blocks = (unsigned)((rows + wpb - 1) / wpb);
    if (dtype == FA_DTYPE_BF16)
        fa::mergePartialKernel<fa::kBF16><<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(acc_o, acc_lse, (const uint16_t*)part_o, part_lse, rows, d);
    else
        fa::mergePartialKernel<fa::kF16><<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(acc_o, acc_lse, (const uint16_t*)part_o, part_lse, rows, d);
    g_launches.fetch_add(1);
    FA_CUDA(cudaGetLastError());
    return FA_OK;
}

int fa_cast_out(const float* src, void* dst, long long n, int dtype, void* stream) {
    g_err[0] = 0;
    if (!src || !dst || n <= 0 || n % 2) return fail(FA_ERR_INVALID_ARGUMENT, "bad cast arguments");
    if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16) return fail(FA_ERR_UNSUPPORTED, "cast target must be 16-bit");
    if (int rc = check_device()) return rc;
    const long long n2 = n / 2;
    const unsigned blocks = (unsigned)((n2 + 255) / 256);
    if (dtype == FA_DTYPE_BF16)
        fa::castOutKernel<fa::kBF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, n2);
    else
        fa::castOutKernel<fa::kF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, n2);
    g_launches.fetch_add(1);
    FA_CUDA(cudaGetLastError());
    return FA_OK;
}

int fa_device_info(int device, fa_device_info_t* out) {
    g_err[0] = 0;
    if (!out) return fail(FA_ERR_INVALID_ARGUMENT, "out is null");
    cudaDeviceProp prop;
    FA_CUDA(cudaGetDeviceProperties(&prop, device));
    out->cc_major = prop.major; out->cc_minor = prop.minor; out->sm_count = prop.multiProcessorCount;
    out->global_mem_bytes = prop.totalGlobalMem; out->smem_per_block_optin = prop.sharedMemPerBlockOptin;
    out->smem_per_sm = prop.sharedMemPerMultiprocessor; out->regs_per_sm = prop.regsPerMultiprocessor;
    out->warp_size = prop.warpSize; out->l2_bytes = prop.l2CacheSize; out->max_threads_per_sm = prop.maxThreadsPerMultiProcessor;
    return FA_OK;
}

int fa_block_q(int d, int dtype) {
    if (dtype == FA_DTYPE_F32) return fa::kF32Rows;
    const fa_tile_choice_t* tc = choose_tile(d, 0, 0);
    return tc ? tc->block_q : fa::kTilesPerCta * fa::kBlockM;
}
int fa_block_kv(int d, int dtype) {
    if (dtype == FA_DTYPE_F32) return fa::kF32Rows;
    const fa_tile_choice_t* tc = choose_tile(d, 0, 0);
    return tc ? tc->block_kv : fa::kBlockN;
}
int fa_tile_table(const fa_tile_choice_t** rows) {
    if (rows) *rows = kTileTable;
    return kTileRows;
}
int fa_choose_tile(int d, int dtype, int causal, int nq, int nk, fa_tile_choice_t* out) {
    g_err[0] = 0;
    (void)nq;
    if (!out) return fail(FA_ERR_INVALID_ARGUMENT, "out is null");
    if (dtype == FA_DTYPE_F32) {      // CUDA-core kernel: one geometry
        *out = fa_tile_choice_t{d, causal ? 1 : 0, 0, fa::kF32Rows, fa::kF32Rows, 1, 0, 0, 0, 0, 1, 0.f};
        return (d % 16 == 0 && d <= 128) ? FA_OK : fail(FA_ERR_UNSUPPORTED, "fp32 path needs d %% 16 == 0 and d <= 128 (got %d)", d);
    }
    const fa_tile_choice_t* tc = choose_tile(d, causal, nk);
    if (!tc) return fail(FA_ERR_UNSUPPORTED, "16-bit path supports d in {64,128} (got %d)", d);
    *out = *tc;
    return FA_OK;
}
int fa_choose_kernel(int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype, int causal, fa_kernel_choice_t* out) {
    g_err[0] = 0;
    if (!out) return fail(FA_ERR_INVALID_ARGUMENT, "out is null");
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || Nk <= 0 || Hq % Hkv != 0) return fail(FA_ERR_INVALID_ARGUMENT, "bad sizes");
    if (int rc = fa_choose_tile(d, dtype, causal, Nq, Nk, &out->tile)) return rc;
    out->heads_per_item = 1;
    if (dtype == FA_DTYPE_F32) {
        out->work_items = (long long)((Nq + fa::kF32Rows - 1) / fa::kF32Rows) * Hq * B;
        return FA_OK;
    }
    const KernelChoice k = choose_kernel(B, Hq, Hkv, Nq, Nk, d, causal, false);
    out->tile.softmax_warps = k.sw;
    out->tile.emu_pairs_per_8 = k.emu;
    out->tile.staged_epilogue = k.stg;
    out->tile.cta_group = k.cg;
    out->tile.stages = k.cg == 2 ? kPairStages : (d == 64 ? 8 : (k.stg && k.sw == 8 ? 4 : 5));
    out->heads_per_item = k.heads_per_item;
    if (k.cg == 2) {
        const int rows = fa::kPairRows / k.heads_per_item;
        out->work_items = (long long)((Nq + rows - 1) / rows) * (Hq / k.heads_per_item) * B;
    } else {
        const long long blocks = (long long)((Nq + fa::kTilesPerCta * fa::kBlockM - 1) / (fa::kTilesPerCta * fa::kBlockM)) * Hq * B;
        int ctas = device_sm_count() - g_sm_reserve.load();
        if (ctas < 1) ctas = 1;
        long long n_full = 0, total = 0;
        plan_counts(blocks, ctas, g_half_items.load() != 0, &n_full, &total);
        out->work_items = total;
    }
    return FA_OK;
}
// A/B tooling (not in include/fa_b200.h): force a kernel variant for every following launch (0, 0 = back to the table);
// switch the half-item tail schedule off / on.
int fa_debug_force_variant(int softmax_warps, int emu, int staged) {
    if (softmax_warps != 0 && softmax_warps != 8 && softmax_warps != 16) return FA_ERR_INVALID_ARGUMENT;
    g_force_sw.store(softmax_warps);
    g_force_emu.store(emu);
    g_force_stg.store(staged);
    return FA_OK;
}
// 0 = as the tile table says, 1 = 1-CTA kernels, 2 = the CTA-pair kernel wherever it exists (d = 128, 8 softmax warps);
// 3 / 4 = the pair kernel with its pairs cut by ROWS / by at most TWO heads even where more heads could share an item (GQA):
// the A/B of the pairings
int fa_debug_force_cta_group(int cta_group) {
    if (cta_group < 0 || cta_group > 4) return FA_ERR_INVALID_ARGUMENT;
    g_force_cg.store(cta_group >= 3 ? 2 : cta_group);
    g_pair_heads.store(cta_group == 3 ? 1 : (cta_group == 4 ? 2 : 4));
    return FA_OK;
}
int fa_debug_plan_counts(long long blocks, int max_ctas, long long* n_full, long long* total) {
    if (blocks < 0 || max_ctas < 1 || !n_full || !total) return FA_ERR_INVALID_ARGUMENT;
    plan_counts(blocks, max_ctas, true, n_full, total);
    return FA_OK;
}
// The work-item decode of the kernels (loaders.cuh: decode_item / decode_pair_item, the same functions compiled for the host)
// over a whole launch, for the CPU tests.  mode 0: 1-CTA kernels (256-row items + the half-item tail, split-KV or not), 1: CTA pairs
// cut by rows, 2: by two heads, 3: by four heads.  out[i] = {b, h, h_kv, q0, rows, split, n_kv, n_steps, n_tile0, n_tile1, tile stride,
// cta rank}; pair modes emit one record per CTA of the pair (what the leader writes into that CTA's mailbox).  Returns the
// number of records, or a negative error code; nothing touches the GPU.
int fa_debug_decode_items(int mode, int B, int Hq, int Hkv, int Nq, int Nk, int causal, int max_ctas, int split_half, int* out, int cap) {
    if (mode < 0 || mode > 3 || B <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv || Nq <= 0 || Nk <= 0 || max_ctas < 1 || !out) return FA_ERR_INVALID_ARGUMENT;
    if ((mode == 2 && (Hq / Hkv) % 2) || (mode == 3 && (Hq / Hkv) % 4)) return FA_ERR_INVALID_ARGUMENT;
    const int hpi = mode == 3 ? 4 : (mode == 2 ? 2 : 1);
    fa::FwdParams p = {};
    p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.Nq = Nq; p.Nk = Nk; p.causal = causal ? 1 : 0; p.causal_off = Nk - Nq; p.q_heads_per_kv = Hq / Hkv;
    const int item_rows = mode == 0 ? fa::kTilesPerCta * fa::kBlockM : fa::kPairRows / hpi;
    const int item_heads = Hq / hpi;
    p.num_q_blocks = (Nq + item_rows - 1) / item_rows;
    p.pair_heads = hpi;
    make_fast_div((unsigned)p.num_q_blocks, &p.div_qblocks_mul, &p.div_qblocks_shr);
    make_fast_div((unsigned)item_heads, &p.div_hq_mul, &p.div_hq_shr);
    make_fast_div((unsigned)p.q_heads_per_kv, &p.div_group_mul, &p.div_group_shr);
    const long long blocks = (long long)p.num_q_blocks * item_heads * B;
    long long n_full = blocks, total = blocks;
    if (mode == 0) plan_counts(blocks, max_ctas, true, &n_full, &total);
    p.n_full_items = (int)n_full; p.total_items = (int)total; p.split_half = mode == 0 ? (split_half ? 1 : 0) : 0;
    int n = 0;
    for (int item = 0; item < p.total_items; ++item) {
        const fa::WorkItem w = mode == 0 ? fa::decode_item(p, item) : fa::decode_pair_item(p, item);
        for (int c = 0; c < (mode == 0 ? 1 : 2); ++c, ++n) {
            if (n >= cap) return FA_ERR_INVALID_ARGUMENT;
            int* r = out + 12 * n;
            r[0] = w.b; r[1] = w.h + (mode ? c * (hpi >> 1) : 0); r[2] = w.h_kv; r[3] = w.q0 + (mode == 1 ? c * fa::kBlockM : 0); r[4] = w.rows; r[5] = w.split;
            r[6] = w.n_kv; r[7] = w.n_steps; r[8] = w.n_tile0; r[9] = w.n_tile1;
            r[10] = mode == 0 ? fa::kBlockM : (w.hstep ? 0 : (w.rows >> 1));      // rows between the CTA's two query tiles (0: two heads instead)
            r[11] = c;
        }
    }
    return n;
}
int fa_debug_fast_div(unsigned d, unsigned n) {      // what the kernels compute for n / d (host replica of loaders.cuh: fast_div)
    unsigned mul = 0, shr = 0;
    make_fast_div(d, &mul, &shr);
    return shr >= 32u ? (int)n : (int)((((unsigned long long)n * mul) >> 32) >> shr);
}
int fa_debug_half_items(int on) { g_half_items.store(on ? 1 : 0); g_split_half.store(on > 1 ? 0 : 1); return FA_OK; }   // 0: off, 1: on (split-KV), 2: on, slot 0 alone
int fa_workspace_bytes(int B, int Hq, int Hkv, int Nq, int Nk, int d, int dtype) {
    g_err[0] = 0;
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || Nk <= 0 || d <= 0) return fail(FA_ERR_INVALID_ARGUMENT, "non-positive size");
    if (Hq % Hkv != 0) return fail(FA_ERR_INVALID_ARGUMENT, "Hq=%d is not a multiple of Hkv=%d", Hq, Hkv);
    if (dtype == FA_DTYPE_F32) {
        if (d % 16 != 0 || d > 128) return fail(FA_ERR_UNSUPPORTED, "fp32 path needs d %% 16 == 0 and d <= 128 (got %d)", d);
    } else if (dtype == FA_DTYPE_F16 || dtype == FA_DTYPE_BF16) {
        if (d != 64 && d != 128) return fail(FA_ERR_UNSUPPORTED, "16-bit path supports d in {64,128} (got %d)", d);
    } else {
        return fail(FA_ERR_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    }
    return 0;      // nothing: no workspace, no padding, no transposed copies
}
int fa_num_cta(int q_dim, int q_block_size) {
    if (q_dim <= 0 || q_block_size <= 0) return 0;
    return (q_dim + q_block_size - 1) / q_block_size;
}

}  // extern "C"
