// probe_umma.cu — stage-level unit test of the building blocks in kernels/utils.cuh, in the spirit of the
// reference's (stale) loader round-trip test (reference: tests/test_loaders.cu:47-110): one CTA, one tile.
//   mode 0 (QK):  S[128x128] = A[128xKD] * B[128xKD]^T   A, B K-major in 128B-swizzled smem via TMA (SS MMA)
//   mode 1 (PV):  O[128xKD]  = P[128x128] * V[128xKD]    P packed 16-bit in TMEM via tcgen05.st (TS MMA), V MN-major
// The shared-memory descriptor fields (LBO / SBO) are runtime arguments so that one GPU run can tell which
// encoding the hardware accepts; the production values are checked first and the exit code reflects only them.
#include "../kernels/utils.cuh"

#include <cuda_bf16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace fa;

struct ProbeArgs {
    int mode;          // 0 QK, 1 PV
    int kd;            // head dim: 64 or 128
    uint32_t lbo, sbo; // descriptor byte offsets for the B operand (and A in mode 0)
    int pack_swap;     // mode 1: swap the two halves when packing P
    const __nv_bfloat16* A;   // mode 1: P values, row-major [128][128]
    float* out;        // [128][128] (mode 0) or [128][kd] (mode 1)
};

__global__ void __launch_bounds__(128, 1)
probeKernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ProbeArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + 32768, bars = base + 65536;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(gen + 65536 + 64);
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
    const int halves = a.kd / 64;

    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(base + 65536 + 64, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (threadIdx.x == 0) {
        const uint32_t bytes = (a.mode == 0 ? 2u : 1u) * 128u * a.kd * 2u;
        mbar_expect_tx(bars, bytes);
        for (int h = 0; h < halves; ++h) {
            if (a.mode == 0) tma_load_4d(&tmA, sA + h * 16384, bars, h * 64, 0, 0, 0);
            tma_load_4d(&tmB, sB + h * 16384, bars, h * 64, 0, 0, 0);
        }
    }
    if (a.mode == 1) {
        // P row for this thread -> 64 packed columns at TMEM column 0 (the softmax store path)
        const int row = threadIdx.x;
        uint32_t pk[64];
        for (int c = 0; c < 64; ++c) {
            float lo = __bfloat162float(a.A[row * 128 + 2 * c]), hi = __bfloat162float(a.A[row * 128 + 2 * c + 1]);
            pk[c] = a.pack_swap ? pack16<kBF16>(hi, lo) : pack16<kBF16>(lo, hi);
        }
        const uint32_t t = tmem + (uint32_t(warp * 32) << 16);
        tmem_st32(t, pk);
        tmem_st32(t + 32, pk + 32);
        tc_wait_st();
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_wait(bars, 0);
        tc_fence_after();
        if (a.mode == 0) {
            const uint32_t idesc = umma_idesc(128, 128, 1, 0, 0);
            for (int ks = 0; ks < a.kd / 16; ++ks) {
                const uint32_t off = (ks / 4) * 16384 + (ks % 4) * 32;
                umma_ss(tmem + 128, umma_desc_sw128(sA + off, a.lbo, a.sbo), umma_desc_sw128(sB + off, a.lbo, a.sbo),
                        idesc, ks > 0);
            }
        } else {
            const uint32_t idesc = umma_idesc(128, a.kd, 1, 0, 1);
            for (int ks = 0; ks < 8; ++ks)
                umma_ts(tmem + 128, tmem + 8 * ks, umma_desc_sw128(sB + ks * 2048, a.lbo, a.sbo), idesc, ks > 0);
        }
        tc_commit(bars + 8);
    }
    mbar_wait(bars + 8, 0);
    tc_fence_after();
    const int ncols = a.mode == 0 ? 128 : a.kd;
    const uint32_t t = tmem + (uint32_t(warp * 32) << 16) + 128;
    for (int q = 0; q < ncols / 32; ++q) {
        uint32_t r[32];
        tmem_ld32(t + 32 * q, r);
        tc_wait_ld();
        for (int i = 0; i < 32; ++i) a.out[(warp * 32 + lane) * ncols + 32 * q + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

// ---- host -------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_enc;

static CUtensorMap make_map(const void* p, int rows, int cols) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, 1, 1};
    cuuint64_t strides[3] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * rows, (cuuint64_t)cols * 2 * rows};
    cuuint32_t box[4] = {64, 128, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
    return m;
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(3); } } while (0)

static double run(int mode, int kd, uint32_t lbo, uint32_t sbo, int swap, const std::vector<__nv_bfloat16>& hA,
                  const std::vector<__nv_bfloat16>& hB) {
    // mode 0: A [128][kd], B [128][kd];  mode 1: A = P [128][128], B = V [128][kd]
    const int acols = mode == 0 ? kd : 128;
    __nv_bfloat16 *dA, *dB; float* dO;
    const int ocols = mode == 0 ? 128 : kd;
    CK(cudaMalloc(&dA, 128 * acols * 2)); CK(cudaMalloc(&dB, 128 * kd * 2)); CK(cudaMalloc(&dO, 128 * ocols * 4));
    CK(cudaMemcpy(dA, hA.data(), 128 * acols * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), 128 * kd * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0, 128 * ocols * 4));
    CUtensorMap tA = make_map(dA, 128, acols >= 64 ? acols : 64), tB = make_map(dB, 128, kd);
    ProbeArgs a{mode, kd, lbo, sbo, swap, dA, dO};
    const int smem = 65536 + 128 + 1024;
    CK(cudaFuncSetAttribute(probeKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probeKernel<<<1, 128, smem>>>(tA, tB, a);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * ocols);
    CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < ocols; ++j) {
            double ref = 0;
            if (mode == 0) for (int k = 0; k < kd; ++k) ref += (double)__bfloat162float(hA[i * kd + k]) * __bfloat162float(hB[j * kd + k]);
            else for (int k = 0; k < 128; ++k) ref += (double)__bfloat162float(hA[i * 128 + k]) * __bfloat162float(hB[k * kd + j]);
            maxerr = fmax(maxerr, fabs(ref - out[i * ocols + j]));
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return maxerr;
}

int main(int argc, char** argv) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    g_enc = (EncodeTiledFn)p;
    const bool scan = argc > 1 && atoi(argv[1]) != 0;
    int bad = 0;
    for (int kd : {128, 64}) {
        srand(1234 + kd);
        std::vector<__nv_bfloat16> A0(128 * kd), B0(128 * kd), P(128 * 128), V(128 * kd);
        auto rnd = [] { return __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); };
        for (auto& x : A0) x = rnd(); for (auto& x : B0) x = rnd(); for (auto& x : P) x = rnd(); for (auto& x : V) x = rnd();
        double e0 = run(0, kd, 16, 1024, 0, A0, B0);
        printf("probe QK kd=%d lbo=16 sbo=1024      max_abs_err=%.4g %s\n", kd, e0, e0 < 0.05 ? "OK" : "MISMATCH");
        double e1 = run(1, kd, 16384, 1024, 0, P, V);
        printf("probe PV kd=%d lbo=16384 sbo=1024  max_abs_err=%.4g %s\n", kd, e1, e1 < 0.05 ? "OK" : "MISMATCH");
        bad += (e0 >= 0.05) + (e1 >= 0.05);
        if (scan || e0 >= 0.05 || e1 >= 0.05) {
            for (uint32_t lbo : {0u, 16u, 1024u, 16384u})
                for (uint32_t sbo : {1024u, 16384u, 128u})
                    printf("  scan QK kd=%d lbo=%u sbo=%u err=%.4g\n", kd, lbo, sbo, run(0, kd, lbo, sbo, 0, A0, B0));
            for (uint32_t lbo : {0u, 16u, 1024u, 16384u, 2048u})
                for (uint32_t sbo : {1024u, 16384u, 128u, 2048u})
                    for (int sw : {0, 1})
                        printf("  scan PV kd=%d lbo=%u sbo=%u swap=%d err=%.4g\n", kd, lbo, sbo, sw, run(1, kd, lbo, sbo, sw, P, V));
        }
    }
    printf(bad ? "PROBE FAILED (%d)\n" : "PROBE PASSED\n", bad);
    return bad ? 1 : 0;
}
