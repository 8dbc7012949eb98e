// Microbenchmark: what warp-instruction issue rate can sm_100a sustain for the INSTRUCTION MIX of the scatter march
// (csrc/atmo_device.cuh scatter_v2: per step 39 FMA-pipe ops, 3 MUFU.EX2 + 1 MUFU.RSQ, 3 integer ops, 1 LDG.128 that hits L1),
// when nothing else is in the way — no dependent-load latency, no loop-carried chains beyond 12 independent accumulators,
// same launch shape as the kernel (128-thread blocks, 10 blocks = 40 warps per SM, <= 48 registers)?
// The ratio kernel / this ceiling is what tuning of the real loop can still win; the gap ceiling / 4.0 is the machine's.
// Every operation is an `asm volatile`, so ptxas keeps the count and the order.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu ; run on the B200.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define FMA(d, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d) : "f"(a), "f"(b))
#define FMA3(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define ADD(d, a) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(d) : "f"(a))
#define MUL(d, a) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(d) : "f"(a))
#define MULD(d, a, b) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define EX2(d, a) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a))
#define RSQ(d, a) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a))
#define IMAD(d, a, b) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b))
#define IMIN(d, a) asm volatile("min.u32 %0, %0, %1;" : "+r"(d) : "r"(a))

// MODE 0: the full mix.  1: MUFU replaced by FMUL (same count).  2: no LDG (replaced by FMA).  3: 46 FMA-pipe ops only.
// 4: the full mix with the four MUFU ops issued back to back (MODE 0 spreads them over the step).
template <int MODE> __global__ void __launch_bounds__(128) mix(float* out, const float4* __restrict__ tab, int iters, float seed) {
    float p0 = seed, p1 = seed + 1, p2 = seed + 2, r0, r1, r2, d2 = 1.0f, sd = 0.5f, inv = 1.0f, dist = 1.0f, y = 0.5f;
    float S = 0.0f, L0 = 0.0f, L1 = 0.0f, L2 = 0.0f, xm, ym, g, h, od, t0 = 0.f, t1 = 0.f, t2 = 0.f, e0 = 1.f, e1 = 1.f, e2 = 1.f, y3;
    const float c0 = 0.999f, c1 = 1e-3f, k0 = -0.1f, k1 = -0.2f, k2 = -0.3f, s0 = 0.3f, s1 = 0.4f, s2 = 0.5f;
    unsigned off = threadIdx.x;
    const unsigned n = 4095u, stride = 257u;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // rel = pos - C (3), |rel|^2 and rel.sun (6)
            r0 = p0; r1 = p1; r2 = p2;
            ADD(r0, c1); ADD(r1, c1); ADD(r2, c1);
            MULD(d2, r0, r0); FMA(d2, r1, r1); FMA(d2, r2, r2);
            MULD(sd, r0, s0); FMA(sd, r1, s1); FMA(sd, r2, s2);
            if (MODE == 1 || MODE == 3) MULD(inv, d2, c0); else if (MODE != 4) RSQ(inv, d2);
            if (MODE == 4) { RSQ(inv, d2); EX2(e0, t0); EX2(e1, t1); EX2(e2, t2); }
            // sqrt_refined (4) + dist - R (1) + y (1)
            MULD(dist, d2, inv); FMA3(y, dist, dist, d2); FMA(dist, y, inv); ADD(dist, c1); FMA3(y, dist, k0, c0);
            // cell coordinates (4) + index (3) + load (1)
            FMA3(xm, sd, inv, c0); FMA3(ym, y, k1, c0); FMA3(g, sd, inv, xm); FMA3(h, y, k1, ym);
            off = __float_as_uint(ym); IMAD(off, stride, __float_as_uint(xm)); IMIN(off, n);
            if (MODE == 2 || MODE == 3) { FMA(q.x, g, h); } else { q = __ldg(tab + off); }
            // interpolation (3), y^3 (2), S (1), od (1), od*k (3)
            FMA3(od, q.w, g, q.z); FMA(od, h, q.x); FMA(od, q.y, g);
            MULD(y3, y, y); MUL(y3, y); ADD(S, y3); FMA(od, S, c1);
            MULD(t0, od, k0); MULD(t1, od, k1); MULD(t2, od, k2);
            if (MODE == 1 || MODE == 3) { MUL(e0, t0); MUL(e1, t1); MUL(e2, t2); }
            else if (MODE != 4) { EX2(e0, t0); EX2(e1, t1); EX2(e2, t2); }
            // accumulators (3), pos += dstep (3), + 5 to reach the kernel's 39 FMA-pipe ops per step
            FMA(L0, y3, e0); FMA(L1, y3, e1); FMA(L2, y3, e2);
            ADD(p0, c1); ADD(p1, c1); ADD(p2, c1);
            FMA(L0, r0, c1); FMA(L1, r1, c1); FMA(L2, r2, c1); MUL(S, c0); ADD(S, c1);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = L0 + L1 + L2 + S + p0 + p1 + p2 + __uint_as_float(off);
}

template <int MODE> void run(const char* name, double per_step) {
    float* out;
    float4* tab;
    const int blocks = 148 * 10;
    cudaMalloc(&out, blocks * 128 * sizeof(float));
    cudaMalloc(&tab, 4096 * sizeof(float4));
    cudaMemset(tab, 0, 4096 * sizeof(float4));
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<MODE><<<blocks, 128>>>(out, tab, 50, 1.0f);
    cudaEventRecord(e0);
    mix<MODE><<<blocks, 128>>>(out, tab, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double winstr = 40.0 * iters * 4.0 * per_step;   // warp-instructions per SM: 10 blocks * 4 warps * iters * 4 steps * per_step
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-44s %8.3f ms  %6.3f warp-instr/clk/SM = %5.1f %% of 4.0  [%.2f SASS instr/step]\n", name, ms, winstr / cycles, 25.0 * winstr / cycles, per_step);
    cudaFree(out); cudaFree(tab);
}

int main(int argc, char** argv) {
    // argv[1..5]: SASS instructions per step of modes 0..4, counted from the binary by issue_mix.py (the loop body of the
    // disassembly / steps per iteration)
    double n[5] = {47, 47, 47, 47, 47};
    for (int i = 0; i < 5 && i + 1 < argc; ++i) n[i] = atof(argv[i + 1]);
    run<0>("scatter mix (39 FP + 4 MUFU + 3 INT + LDG)", n[0]);
    run<4>("same, 4 MUFU back to back", n[4]);
    run<1>("MUFU -> FMUL", n[1]);
    run<2>("LDG -> FMA", n[2]);
    run<3>("FMA pipe + INT only", n[3]);
    return 0;
}
