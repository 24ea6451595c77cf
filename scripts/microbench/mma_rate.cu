// mma_rate.cu — issue rate of tcgen05.mma kind::f16 (bf16) for the shapes the attention kernels use, 1-CTA and CTA-pair:
//   SS  D[tmem] = A[smem] B[smem]   (Q K^T)      TS  D[tmem] = A[tmem] B[smem]   (P V)
//   cta_group::1  M = 128, N = 128       cta_group::2  M = 256 (128 per CTA), N = 128 (64 per CTA's shared memory)
// One thread of the (leader) CTA issues `n` MMAs back to back (operands: whatever is in shared memory / TMEM — timing only),
// one commit, and waits for it.  Prints clk per instruction.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../flash-attention-cuda-c_b200/kernels
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "utils.cuh"
using namespace fa;
constexpr int kOps = 65536 + 98304;      // A: two Q tiles; B: three 32 KiB ring slots

template <int CG, int TS>
__global__ void __launch_bounds__(128, 1) rate(int n, int N, int stream, int mix, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t bar = base + kOps, tptr = base + kOps + 16;
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CG == 2) rank = cluster_ctarank();
    for (int i = threadIdx.x; i < kOps / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // finite values
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if (CG == 2) { tmem_alloc_pair(tptr, 512); tmem_relinquish_pair(); } else { tmem_alloc(tptr, 512); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem + kOps + 16);
    if (CG == 2) cluster_sync_all();
    if (rank == 0 && warp == 1) {
        const uint32_t idesc = umma_idesc(CG == 2 ? 256 : 128, N, 1, 0, TS ? 1 : 0);
        const uint64_t da = umma_desc_sw128(0, 16, 1024) + (base >> 4);
        const uint64_t db = (TS ? umma_desc_sw128(0, 16384, 1024) : umma_desc_sw128(0, 16, 1024)) + ((base + 65536) >> 4);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            t0 = clock64();
            if (elect_one_sync()) {
                for (int i = 0; i < n; ++i) {
                    // stream == 1: operands walk through the tiles of the attention kernels (A: two 32 KiB Q tiles, B: a ring of
                    // 16 KiB (pair) / 32 KiB slots), k-step by k-step, instead of re-reading the same 8 KiB
                    const int ks = i & 7, tile = i >> 3;
                    const uint32_t half_b = (CG == 2) ? 8192u : 16384u;
                    uint32_t off_a = ((i & 3) * 32) >> 4, off_b = off_a, off_v = ((i & 7) * 2048) >> 4;
                    if (stream) {
                        off_a = ((tile & 1) * 32768 + (ks / 4) * 16384 + (ks % 4) * 32) >> 4;
                        off_b = ((tile % 3) * 2 * half_b + (ks / 4) * half_b + (ks % 4) * 32) >> 4;
                        off_v = ((tile % 3) * 16384 * (CG == 2 ? 1 : 2) + ks * 2048) >> 4;
                    }
                    if (TS) {
                        if (CG == 2) umma_ts_pair(tm + 256, tm + 128 + 8 * (i & 7), db + off_v, idesc, 1);
                        else umma_ts(tm + 256, tm + 128 + 8 * (i & 7), db + off_v, idesc, 1);
                    } else {
                        if (CG == 2) umma_ss_pair(tm, da + off_a, db + off_b, idesc, 1);
                        else umma_ss(tm, da + off_a, db + off_b, idesc, 1);
                    }
                    if (mix && (i & 7) == 7) {      // a P V half (4 TS MMAs) after every Q K^T (8 SS MMAs), as in the kernels
                        const uint32_t ipv = umma_idesc(CG == 2 ? 256 : 128, 128, 1, 0, 1);
                        const uint64_t dv = umma_desc_sw128(0, 16384, 1024) + ((base + 65536) >> 4);
                        for (int h = 0; h < 4; ++h) {
                            if (CG == 2) umma_ts_pair(tm + 256, tm + 128 + 8 * h, dv + ((h * 2048) >> 4), ipv, 1);
                            else umma_ts(tm + 256, tm + 128 + 8 * h, dv + ((h * 2048) >> 4), ipv, 1);
                        }
                    }
                }
                if (CG == 2) tc_commit_pair(bar, 1); else tc_commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, rep & 1);
            t1 = clock64();
        }
        if (lane == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) { tc_fence_after(); if (CG == 2) tmem_dealloc_pair(tm, 512); else tmem_dealloc(tm, 512); }
}

template <int CG, int TS>
void run(const char* name, int N, int stream = 0, int mix = 0) {
    long long* d; cudaMalloc(&d, 8);
    auto k = rate<CG, TS>;
    const int smem = kOps + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n = 2048;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, n, N, stream, mix, d);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d %s%s %s  %.1f clk per MMA (K=16)\n", name, N, stream ? "streaming operands " : "", mix ? "+ 4 TS per 8 SS" : "", e == cudaSuccess ? "ok" : cudaGetErrorString(e), double(h) / (mix ? n + n / 2 : n));
    cudaFree(d);
}

int main() {
    for (int N : {64, 128, 256}) {
        run<1, 0>("SS cta_group::1 M=128", N);
        run<2, 0>("SS cta_group::2 M=256", N);
        if (N <= 128) { run<1, 1>("TS cta_group::1 M=128", N); run<2, 1>("TS cta_group::2 M=256", N); }
    }
    run<1, 0>("SS cta_group::1 M=128", 128, 1);
    run<2, 0>("SS cta_group::2 M=256", 128, 1);
    run<1, 1>("TS cta_group::1 M=128", 128, 1);
    run<2, 1>("TS cta_group::2 M=256", 128, 1);
    run<1, 0>("SS+TS cta_group::1", 128, 1, 1);
    run<2, 0>("SS+TS cta_group::2", 128, 1, 1);
    return 0;
}
