// Microbenchmark: issue/pipe rate of packed fp32x2 ops (FFMA2/FADD2/FMUL2) vs scalar FFMA/FADD and MUFU on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu ; run on the B200.
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE> __global__ void k(float* out, int iters, float seed) {
    float2 a[8];
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i, seed - i); s[i] = seed * i; }
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, -0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { s[i] = fmaf(s[i], 1.0001f, 0.5f); a[i].x = fmaf(a[i].x, 0.9999f, -0.5f); }        // 2 scalar FFMA (imm form)
            if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);                                                       // 1 FFMA2 (3-reg)
            if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); s[i] = fmaf(s[i], m.x, c.x); }                      // FFMA2 + FFMA
            if (MODE == 3) { s[i] = fmaf(s[i], m.x, c.x); a[i].x = fmaf(a[i].x, m.y, c.y); }                    // 2 scalar FFMA (3-reg)
            if (MODE == 4) a[i] = __fadd2_rn(a[i], c);                                                          // FADD2
            if (MODE == 5) { s[i] = s[i] + c.x; a[i].x = a[i].x + c.y; }                                        // 2 scalar FADD
            if (MODE == 6) { a[i] = __ffma2_rn(a[i], m, c); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i])); }  // FFMA2 + MUFU
            if (MODE == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i])); }                         // MUFU only
            if (MODE == 8) { a[i] = __ffma2_rn(a[i], m, c); s[i] = __int_as_float(__float_as_int(s[i]) + 3); }  // FFMA2 + IADD
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, int per_iter_instr) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    // warp-instructions per SM: 8 blocks * 8 warps * iters * per_iter_instr
    double winstr = 64.0 * iters * per_iter_instr;
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM (at nominal %d MHz)  [%d instr/iter]\n", name, ms, winstr / cycles, clk_khz / 1000, per_iter_instr);
    cudaFree(out);
}

int main() {
    run<0>("2x scalar FFMA (imm)", 16);
    run<3>("2x scalar FFMA (3-reg)", 16);
    run<1>("1x FFMA2", 8);
    run<2>("FFMA2 + FFMA", 16);
    run<4>("1x FADD2", 8);
    run<5>("2x scalar FADD", 16);
    run<7>("MUFU.EX2 only", 8);
    run<6>("FFMA2 + MUFU.EX2", 16);
    run<8>("FFMA2 + IADD", 16);
    return 0;
}
