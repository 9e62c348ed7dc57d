// Butterfly throughput probe: Shoup (IMAD, IMAD.HI, IMAD) against Montgomery (IMAD.WIDE, IMAD, IMAD.WIDE) lazy butterflies,
// compiled from C as the NTT passes are.  Reports clocks per butterfly per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;
#define P1 0x3FFE8001u
__device__ __forceinline__ u32 fold(u32 x) { u32 y = x - 2 * P1; return y < x ? y : x; }
template <int V> __global__ void __launch_bounds__(1024) k(u32 *x, const uint2 *w, u32 pinv, int n)
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
            const uint2 tw = w[(it & 63) * 32 + s * 8 + (i >> 1)];
            u32 v, z = 0;
            if (V == 0) { v = tw.x * y[j] - __umulhi(tw.y, y[j]) * P1; }
            else { u64 T = (u64)y[j] * tw.x; u32 m = (u32)T * pinv; u64 R = (u64)m * P1 + T; v = (u32)(R >> 32); z = (u32)R; }
            const u32 u = fold(y[i]);
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
    u32 *d; uint2 *w; cudaMalloc(&d, 4 * 16 * threads); cudaMalloc(&w, 8 * 64 * 32); cudaMemset(d, 1, 4 * 16 * threads); cudaMemset(w, 3, 8 * 64 * 32);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; r++) { cudaEventRecord(e0); k<V><<<pr.multiProcessorCount, threads>>>(d, w, 0xFFFE7FFFu, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    const double cycles = best * 1e-3 * clk * 1e3, bf = (double)iters * 32 * warps;
    printf("%-12s warps/SMSP=%d  %.2f clk per warp-butterfly per SMSP\n", name, warps, cycles / bf);
}
int main() { for (int w : {2, 4, 8}) { run<0>("Shoup", w); run<1>("Montgomery", w); } return 0; }
