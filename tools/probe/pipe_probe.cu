// Integer-pipe throughput probe for sm_100a: warp-instructions per cycle per SM sub-partition for the opcodes the
// blind-rotation butterflies are made of.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;
#define CH 8
template <int OP> __global__ void __launch_bounds__(1024) k(u32 *sink, int iters, u32 a, u32 b)
{
    u32 x[CH]; u64 y[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = threadIdx.x + i; y[i] = x[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[i]) : "r"(a), "r"(b));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(a), "r"(b));   // IADD3
            if (OP == 4) asm volatile("{.reg .u32 t; sub.u32 t, %0, %1; min.u32 %0, %0, t;}" : "+r"(x[i]) : "r"(a));   // VIADDMNMX?
            if (OP == 5) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(a), "r"(b)); }
            if (OP == 6) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)); asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(a), "r"(b)); }
            if (OP == 7) { // one Shoup butterfly-like group: IMAD, IMAD.HI, IMAD + 3 ALU
                u32 q, v;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x[i]), "r"(b));
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(v) : "r"(x[i]), "r"(a));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(v) : "r"(q), "r"(b));
                asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(v), "r"(b));
                asm volatile("{.reg .u32 t; sub.u32 t, %0, %1; min.u32 %0, %0, t;}" : "+r"(x[i]) : "r"(a));
                asm volatile("sub.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(v), "r"(b));
            }
            if (OP == 9) { // Montgomery butterfly-like group: IMAD.WIDE, IMAD, IMAD.WIDE + 3 ALU (run-time modulus)
                u64 T, R; u32 tl, th, m, rl, rh;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(T) : "r"(x[i]), "r"(x[(i + 1) % CH]));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(tl), "=r"(th) : "l"(T));
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"(tl), "r"(a));
                asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(R) : "r"(m), "n"(0x3FFE8001), "l"(T));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(rl), "=r"(rh) : "l"(R));
                asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(rh), "r"(rl));
                asm volatile("{.reg .u32 t; sub.u32 t, %0, %1; min.u32 %0, %0, t;}" : "+r"(x[i]) : "r"(a));
                asm volatile("sub.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(rh), "r"(b));
            }
            if (OP == 8) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "n"(12345), "r"(b));   // immediate form
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i] + (u32)y[i] + (u32)(y[i] >> 32);
    if (s == 0x12345678u) sink[0] = s;
}
template <int OP> void run(const char *name, int per_iter, int warps_per_smsp)
{
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    u32 *d; cudaMalloc(&d, 4);
    const int iters = 1 << 13, threads = 32 * 4 * warps_per_smsp;   // one CTA per SM
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<OP><<<pr.multiProcessorCount, threads>>>(d, iters, 3u + r, 5u); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * clk * 1e3;
    const double winst = (double)iters * CH * per_iter * warps_per_smsp;      // warp-instructions per SMSP
    printf("%-28s warps/SMSP=%d  %.3f warp-inst/clk/SMSP  (%.2f clk per warp-inst)\n", name, warps_per_smsp, winst / cycles, cycles / winst);
    cudaFree(d);
}
int main()
{
    for (int w : {2, 4, 8}) {
        run<0>("IMAD lo", 1, w); run<8>("IMAD lo imm", 1, w); run<1>("IMAD.HI", 1, w); run<2>("IMAD.WIDE", 1, w); run<3>("IADD3", 1, w); run<4>("VIADDMNMX", 1, w);
        run<5>("IMAD + IADD3", 2, w); run<6>("IMAD.HI + IADD3", 2, w); run<7>("Shoup butterfly mix (6)", 6, w); run<9>("Montgomery bfly mix (6)", 6, w);
    }
    return 0;
}
