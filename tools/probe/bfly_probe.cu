// Butterfly throughput probe for sm_100a, compiled from C as the NTT passes are: a thread holds 16 values and runs radix-2
// stages on them; one twiddle load serves 4 butterflies (2 bootstraps x 2 primes in the real kernel).
// Reports clocks per warp-butterfly per SM sub-partition (the multiplier-pipe bound of a Shoup butterfly is 8).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;
#define P1 0x3FFE8001u
__device__ __forceinline__ u32 fold(u32 x) { u32 y = x - 2 * P1; return y < x ? y : x; }
// V: 0 Shoup butterfly; 1 Montgomery (WIDE, IMAD, WIDE); 2 Shoup without the fold (small-prime lazy ranges);
//    3 Shoup with the quotient taken from a full-rate IMAD.WIDE (low word kept alive); 4 multiplies only; 5 adds only
template <int V> __global__ void __launch_bounds__(1024) k(u32 *x, const uint4 *w, u32 pinv, int n)
{
    u32 y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) y[i] = x[threadIdx.x + blockDim.x * i];
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (i & (1 << s)) continue;
            const int j = i | (1 << s);
            const uint4 t4 = w[(it & 63) * 8 + s * 2 + (i >> 3)];
            const u32 tw = (i & 2) ? t4.z : t4.x, tws = (i & 2) ? t4.w : t4.y;
            u32 v = y[j], z = 0;
            if (V == 0 || V == 2 || V == 4) v = tw * y[j] - __umulhi(tws, y[j]) * P1;
            if (V == 1) { u64 T = (u64)y[j] * tw; u32 m = (u32)T * pinv; u64 R = (u64)m * P1 + T; v = (u32)(R >> 32); z = (u32)R; }
            if (V == 3) { u64 T = (u64)y[j] * tws; v = tw * y[j] - (u32)(T >> 32) * P1; z = (u32)T; }
            if (V == 6) { u64 T = (u64)y[j] * tws; u32 lo = (u32)T; asm volatile("" :: "r"(lo)); v = tw * y[j] - (u32)(T >> 32) * P1; }
            if (V == 7) { v = tw * y[j] - (u32)__mulhi((int)tws, (int)y[j]) * P1; }
            if (V == 8) { u32 q; asm("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(y[j]), "r"(tws)); v = tw * y[j] - q * P1; }
            if (V == 4) { y[j] = v; continue; }
            const u32 u = (V == 2) ? y[i] : fold(y[i]);
            y[i] = u + v + z; y[j] = u - v + 2 * P1;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) x[threadIdx.x + blockDim.x * i] = y[i];
}
template <int V> void run(const char *name, int warps)
{
    int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const int threads = 128 * warps, iters = 4096;
    u32 *d; uint4 *w; cudaMalloc(&d, 4 * 16 * threads); cudaMalloc(&w, 16 * 64 * 8); cudaMemset(d, 1, 4 * 16 * threads); cudaMemset(w, 3, 16 * 64 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; r++) { cudaEventRecord(e0); k<V><<<pr.multiProcessorCount, threads>>>(d, w, 0xFFFE7FFFu, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    const double cycles = best * 1e-3 * clk * 1e3, bf = (double)iters * 32 * warps;
    printf("%-28s warps/SMSP=%d  %.2f clk per warp-butterfly per SMSP\n", name, warps, cycles / bf);
    cudaFree(d); cudaFree(w);
}
int main()
{
    for (int w : {4}) { run<0>("Shoup", w); run<1>("Montgomery", w); run<2>("Shoup, no fold", w); run<3>("Shoup, quotient via WIDE", w); run<4>("multiplies only", w); run<5>("adds only", w); run<6>("quotient WIDE, lo kept alive", w); run<7>("signed mulhi", w); run<8>("mulhi operands swapped", w); }
    return 0;
}
