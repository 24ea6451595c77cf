// probe_softmax.cu — stage-level unit test of the softmax step of kernels/computers.cuh (softmaxRows16), the slot the
// reference left empty (reference: tests/test_computers.cu, 0 bytes): one CTA, one 128 x 128 score tile.
//   1. four warps write a known fp32 score tile S into TMEM columns [0,128) in the 32x32b view (thread = row),
//   2. eight warps run ONE online-softmax step on it in the 16-lane view the kernel uses (tcgen05.ld 16x256b: a row is
//      spread over the 4 threads of a quad, two rows per thread; warps w and w+4 split the 32 rows of a lane quarter):
//      row max with two quad shuffles, P = exp2(S*c - m*c) through fma.rn.f32x2 + ex2.approx, packed to bf16 pairs and
//      stored with tcgen05.st 16x128b into TMEM columns [128,192) — the A operand of the P V MMA,
//   3. four warps read P back in the 32x32b view and everything is compared on the host with
//      m = rowmax(S), P = bf16(2^((S - m) c)), l = sum_j 2^((S - m) c):
// a wrong register <-> (lane, column) assumption for either 16-lane shape shows up as a permuted / misplaced P.
// Also checks the causal-diagonal mask of the fragment layout (mode 1: key j of row i masked iff j > i).
#include "../kernels/utils.cuh"

#include <cuda_bf16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace fa;

struct SmArgs {
    const float* S;        // [128][128]
    uint32_t* P;           // [128][64] packed bf16 pairs read back from TMEM
    float* m;              // [128]
    float* l;              // [128]
    float c;               // scale * log2(e)
    int causal;            // 1: mask j > i
};

__global__ void __launch_bounds__(256, 1) probeSoftmaxKernel(const SmArgs a) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp < 4) {     // S -> TMEM, one row per thread
        const int row = warp * 32 + lane;
        const uint32_t t = tmem + (uint32_t(warp * 32) << 16);
        for (int q = 0; q < 4; ++q) {
            uint32_t r[32];
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(a.S[row * 128 + 32 * q + i]);
            tmem_st32(t + 32 * q, r);
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    {   // the softmax step, 16-lane view: warp (h, q4) owns rows [32 q4 + 16 h, +16)
        const int q4 = warp & 3, h = warp >> 2;
        const int c4 = lane & 3, rr = lane >> 2;
        const int rowA = q4 * 32 + h * 16 + rr;            // rowB = rowA + 8
        const uint32_t lane16 = uint32_t(q4 * 32 + h * 16) << 16;
        uint32_t r[64];
        tmem_ld_16x256b_x8(tmem + lane16, r);
        tmem_ld_16x256b_x8(tmem + lane16 + 64u, r + 32);
        tc_wait_ld();
        if (a.causal) {
            const int limA = rowA - 2 * c4, limB = rowA + 8 - 2 * c4;
            for (int g = 0; g < 16; ++g) {
                r[4 * g + 0] = mask_gt(r[4 * g + 0], 8 * g, limA);
                r[4 * g + 1] = mask_gt(r[4 * g + 1], 8 * g + 1, limA);
                r[4 * g + 2] = mask_gt(r[4 * g + 2], 8 * g, limB);
                r[4 * g + 3] = mask_gt(r[4 * g + 3], 8 * g + 1, limB);
            }
        }
        float mA = -INFINITY, mB = -INFINITY;
        for (int g = 0; g < 16; ++g) {
            mA = max3(mA, __uint_as_float(r[4 * g + 0]), __uint_as_float(r[4 * g + 1]));
            mB = max3(mB, __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
        }
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
        const float2 c2 = make_float2(a.c, a.c);
        const float2 nmA = make_float2(-mA * a.c, -mA * a.c), nmB = make_float2(-mB * a.c, -mB * a.c);
        float2 lA = make_float2(0.f, 0.f), lB = make_float2(0.f, 0.f);
        for (int g = 0; g < 16; ++g) {
            float2 xA = fma2(make_float2(__uint_as_float(r[4 * g + 0]), __uint_as_float(r[4 * g + 1])), c2, nmA);
            float2 xB = fma2(make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])), c2, nmB);
            xA.x = ex2_approx(xA.x); xA.y = ex2_approx(xA.y);
            xB.x = ex2_approx(xB.x); xB.y = ex2_approx(xB.y);
            lA = add2(lA, xA);
            lB = add2(lB, xB);
            const int o = (g < 8) ? 2 * g : 32 + 2 * (g - 8);
            r[o + 0] = pack16<kBF16>(xA.x, xA.y);
            r[o + 1] = pack16<kBF16>(xB.x, xB.y);
        }
        tmem_st_16x128b_x8(tmem + lane16 + 128u, r);
        tmem_st_16x128b_x8(tmem + lane16 + 128u + 32u, r + 32);
        tc_wait_st();
        float sA = lA.x + lA.y, sB = lB.x + lB.y;
        sA += __shfl_xor_sync(0xffffffffu, sA, 1);
        sB += __shfl_xor_sync(0xffffffffu, sB, 1);
        sA += __shfl_xor_sync(0xffffffffu, sA, 2);
        sB += __shfl_xor_sync(0xffffffffu, sB, 2);
        if (c4 == 0) {
            a.m[rowA] = mA; a.m[rowA + 8] = mB;
            a.l[rowA] = sA; a.l[rowA + 8] = sB;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 4) {     // P back in the 32x32b view: 64 packed columns per row
        const int row = warp * 32 + lane;
        const uint32_t t = tmem + (uint32_t(warp * 32) << 16) + 128u;
        for (int q = 0; q < 2; ++q) {
            uint32_t r[32];
            tmem_ld32(t + 32 * q, r);
            tc_wait_ld();
            for (int i = 0; i < 32; ++i) a.P[row * 64 + 32 * q + i] = r[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(3); } } while (0)

static float bf16_to_f(uint16_t h) { uint32_t u = uint32_t(h) << 16; float f; memcpy(&f, &u, 4); return f; }

static int run(int causal) {
    std::vector<float> S(128 * 128);
    srand(77 + causal);
    for (auto& x : S) x = (rand() % 20001 - 10000) / 1000.0f;     // scores in [-10, 10]
    const float c = 0.0883883f * 1.4426950408889634f;             // 1/sqrt(128) * log2(e)
    float *dS, *dm, *dl; uint32_t* dP;
    CK(cudaMalloc(&dS, S.size() * 4)); CK(cudaMalloc(&dP, 128 * 64 * 4)); CK(cudaMalloc(&dm, 512)); CK(cudaMalloc(&dl, 512));
    CK(cudaMemcpy(dS, S.data(), S.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dP, 0xff, 128 * 64 * 4));
    SmArgs a{dS, dP, dm, dl, c, causal};
    probeSoftmaxKernel<<<1, 256>>>(a);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> P(128 * 64); std::vector<float> m(128), l(128);
    CK(cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(m.data(), dm, 512, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(l.data(), dl, 512, cudaMemcpyDeviceToHost));
    double e_m = 0, e_p = 0, e_l = 0;
    for (int i = 0; i < 128; ++i) {
        const int nk = causal ? i + 1 : 128;
        float mx = -INFINITY;
        for (int j = 0; j < nk; ++j) mx = fmaxf(mx, S[i * 128 + j]);
        double sum = 0;
        for (int j = 0; j < 128; ++j) {
            const double p = j < nk ? exp2((double(S[i * 128 + j]) - mx) * c) : 0.0;
            sum += p;
            const uint32_t w = P[i * 64 + j / 2];
            const float got = bf16_to_f(uint16_t(j & 1 ? (w >> 16) : (w & 0xffffu)));      // low half = even key
            e_p = fmax(e_p, fabs(got - p));
        }
        e_m = fmax(e_m, fabs(m[i] - mx));
        e_l = fmax(e_l, fabs(l[i] - sum) / sum);
    }
    const bool ok = e_m == 0 && e_p <= 4e-3 && e_l <= 1e-3;      // bf16 rounding of P <= 2^-9; ex2.approx ~2^-22 relative
    printf("probe softmax step (16x256b ld / 16x128b st) causal=%d: max|m-ref|=%.3g  max|P-ref|=%.3g  max rel|l-ref|=%.3g  %s\n",
           causal, e_m, e_p, e_l, ok ? "OK" : "MISMATCH");
    cudaFree(dS); cudaFree(dP); cudaFree(dm); cudaFree(dl);
    return ok ? 0 : 1;
}

int main() {
    int bad = run(0) + run(1);
    printf(bad ? "SOFTMAX PROBE FAILED (%d)\n" : "SOFTMAX PROBE PASSED\n", bad);
    return bad ? 1 : 0;
}
