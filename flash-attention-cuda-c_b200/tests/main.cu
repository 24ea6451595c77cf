// tests/main.cu — end-to-end check of the compat entry point, launched the way the reference's only caller
// launches it (reference: tests/main.cu:51-61: grid 1, (QT+2) warps, dynamic smem (2QT+4R+QT)*D*4, template
// instantiated with QT-2), plus what that test lacks: random inputs, causal, several (batch,
// head) pairs, a tolerance and a non-zero exit code on failure.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a tests/main.cu -o tests/compat_main && tests/compat_main
#include "../kernels/FlashAttention.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// straightforward host attention over [B*H, N, D] fp32 (contract: SURVEY.md App. B)
static void host_attention(const std::vector<float>& q, const std::vector<float>& k, const std::vector<float>& v,
                           std::vector<float>& o, int BH, int N, int D, float scale, bool causal) {
    std::vector<double> p(N);
    for (int bh = 0; bh < BH; ++bh)
        for (int i = 0; i < N; ++i) {
            const int nk = causal ? i + 1 : N;
            double mx = -1e300, sum = 0;
            for (int j = 0; j < nk; ++j) {
                double dot = 0;
                for (int c = 0; c < D; ++c) dot += (double)q[((size_t)bh * N + i) * D + c] * k[((size_t)bh * N + j) * D + c];
                p[j] = dot * scale;
                mx = fmax(mx, p[j]);
            }
            for (int j = 0; j < nk; ++j) { p[j] = exp(p[j] - mx); sum += p[j]; }
            for (int c = 0; c < D; ++c) {
                double acc = 0;
                for (int j = 0; j < nk; ++j) acc += p[j] * v[((size_t)bh * N + j) * D + c];
                o[((size_t)bh * N + i) * D + c] = (float)(acc / sum);
            }
        }
}

template <int D, int QT, int R>
static int run_case(const char* name, int B, int H, int N, bool ones, bool causal, int grid) {
    const size_t n = (size_t)B * H * N * D;
    std::vector<float> q(n), k(n), v(n), o(n, 0.f), ref(n);
    srand(7);
    for (size_t i = 0; i < n; ++i) {
        q[i] = ones ? 1.f : (rand() % 2001 - 1000) / 500.f;
        k[i] = ones ? 1.f : (rand() % 2001 - 1000) / 500.f;
        v[i] = ones ? 1.f : (rand() % 2001 - 1000) / 500.f;
    }
    const float scale = 1.0f / sqrtf((float)D);
    float *dq, *dk, *dv, *dO;
    CK(cudaMalloc(&dq, n * 4)); CK(cudaMalloc(&dk, n * 4)); CK(cudaMalloc(&dv, n * 4)); CK(cudaMalloc(&dO, n * 4));
    CK(cudaMemcpy(dq, q.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dk, k.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0, n * 4));
    // the reference caller's launch shape
    const int threads = (QT + 2) * WARP;
    const size_t smem = (size_t)(2 * QT + 4 * R + QT) * D * sizeof(float);
    twoLoaderMhaFlashAttentionKernel<D, (QT > 2 ? QT - 2 : QT), R><<<grid, threads, smem>>>(dq, dk, dv, dO, B, H, N, scale, causal);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o.data(), dO, n * 4, cudaMemcpyDeviceToHost));
    host_attention(q, k, v, ref, B * H, N, D, scale, causal);
    double maxdiff = 0, maxref = 0;
    for (size_t i = 0; i < n; ++i) { maxdiff = fmax(maxdiff, fabs((double)o[i] - ref[i])); maxref = fmax(maxref, fabs((double)ref[i])); }
    const bool ok = maxdiff <= 1e-4 * fmax(maxref, 1.0);
    printf("%-44s max|diff| = %.3e (max|ref| %.3f)  %s\n", name, maxdiff, maxref, ok ? "PASS" : "FAIL");
    cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dO);
    return ok ? 0 : 1;
}

int main() {
    int fails = 0;
    // the reference's known-answer test, launched exactly as the reference launches it (tests/main.cu:24-36,105-107)
    fails += run_case<16, 4, 4>("ones  B1 H1 N16  D16 <16,4-2,4> grid1", 1, 1, 16, true, false, 1);
    fails += run_case<16, 4, 4>("rand  B1 H1 N16  D16 grid1", 1, 1, 16, false, false, 1);
    fails += run_case<16, 4, 4>("rand  B1 H1 N16  D16 causal grid1", 1, 1, 16, false, true, 1);
    fails += run_case<64, 8, 8>("rand  B1 H1 N256 D64 grid1 (config 1)", 1, 1, 256, false, false, 1);
    fails += run_case<64, 8, 8>("rand  B2 H3 N100 D64 causal grid7", 2, 3, 100, false, true, 7);
    fails += run_case<128, 8, 8>("rand  B1 H2 N77  D128 grid148", 1, 2, 77, false, false, 148);
    printf(fails ? "COMPAT FAILED (%d)\n" : "COMPAT PASSED\n", fails);
    return fails ? 1 : 0;
}
