// pipes.cu — per-warp issue cost (cycles per warp-instruction) of the instructions the softmax loop is made of,
// measured with one warp per SM sub-partition (4 warps per CTA, 1 CTA per SM) so that a lone warp owns its pipes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 64
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                if (MODE == 1) { uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(a[i]), "f"(a[(i + 1) & 15])); acc ^= d; }
                if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
                if (MODE == 3) { if (i % 2 == 0) { uint64_t u; asm volatile("mov.b64 %0, {%1,%2};" : "=l"(u) : "f"(a[i]), "f"(a[i + 1]));
                                 asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(u)); asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(u)); } }
                if (MODE == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 3) & 15]), "f"(a[(i + 7) & 15]));
                if (MODE == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(a[(i + 5) & 15]), "f"(a[(i + 9) & 15])); acc ^= d; }
                if (MODE == 6) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 8) & 15]) : "f"(1.0001f), "f"(0.5f));
                                 asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 4) & 15]) : "f"(1.0001f), "f"(0.5f)); }
                if (MODE == 7) { uint32_t u = __float_as_uint(a[i]); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u) : "r"(0x800000u), "r"(acc)); a[i] = __uint_as_float(u); }
                if (MODE == 8) { asm volatile("add.rm.ftz.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(12582912.f)); }
                if (MODE == 9) { uint32_t u = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
                if (MODE == 10) { uint32_t u = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
                if (MODE == 11) { uint32_t u = __float_as_uint(a[i]); uint32_t lo, hi; asm volatile("shl.b32 %0, %1, 16;" : "=r"(lo) : "r"(u)); asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(hi) : "r"(u)); acc ^= lo + hi; }
                if (MODE == 12) { uint32_t u = __float_as_uint(a[i]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i])); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 4); cudaMallocManaged(&cyc, 16 * 8);
    const char* names[] = {"MUFU.EX2", "F2FP.BF16 pack", "FFMA", "FFMA2 (per packed instr)", "FMNMX3", "MUFU + F2FP interleaved (per pair)",
                           "MUFU + 2 FFMA interleaved (per triple)", "IMAD", "FADD.RM", "ex2.bf16x2 (2 exps)", "ex2.f16x2 (2 exps)", "bf16x2 unpack (shl+and)", "tanh.f32"};
    for (int warps : {4, 8}) {
        printf("-- %d warps per CTA (%d per sub-partition), 148 CTAs\n", warps, warps / 4);
#define RUN(M) k<M><<<148, warps * 32>>>(out, cyc, 0.5f); cudaDeviceSynchronize(); \
        printf("%-44s %.2f cycles per warp-instruction (group)\n", names[M], double(cyc[M]) / (256.0 * REP / (M == 3 ? 2 : 1)));
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    }
    return 0;
}
