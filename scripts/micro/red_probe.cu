// L2 fp32 reduction throughput as the generic LSTMP backward uses it (csrc/lstmp_sm100.cu: every CTA adds a
// [256 rows x NB utterances] fp32 partial tile into dmt[t-1] with coalesced red.global.add.f32; KS CTAs hit each
// element).  Variants: scalar red, red.v4.f32 (one 16-byte op per 4 rows), and plain stores as the bandwidth floor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/red_probe scripts/micro/red_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// grid = groups * KS * MS CTAs of 128 threads; CTA (ks, ms) adds rows [256 ms, 256 ms + 256) x NB utterances, `iters` times
template <int MODE>
__global__ void __launch_bounds__(128) probe(float* dmt, int Cp, int KS, int MS, int NB, int iters) {
    const int per = KS * MS, grp = blockIdx.x / per, rem = blockIdx.x % per, ms = rem % MS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        float* base = dmt + ((size_t)(it & 1) * 4096 + (size_t)grp * NB) * Cp;
        for (int mt = 0; mt < 2; ++mt) {
            if (MODE == 0 || MODE == 2) {
                const int crow = 256 * ms + 128 * mt + warp * 32 + lane;
                for (int n = 0; n < NB; ++n) {
                    if (MODE == 0) red1(base + (size_t)n * Cp + crow, 1.0f);
                    else base[(size_t)n * Cp + crow] = 1.0f;
                }
            } else {
                // after a 4 x 4 transpose among lane quads a thread holds 4 consecutive rows of NB / 4 utterances
                const int crow = 256 * ms + 128 * mt + warp * 32 + (lane >> 2) * 4;
                for (int n = (lane & 3); n < NB; n += 4) red4(base + (size_t)n * Cp + crow, 1.f, 1.f, 1.f, 1.f);
            }
        }
    }
}

int main() {
    float* d;
    cudaMalloc(&d, sizeof(float) * 2 * 4096 * 1024);
    cudaMemset(d, 0, sizeof(float) * 2 * 4096 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 200;
    printf("%-8s %5s %4s %4s %4s %6s | %10s %12s\n", "mode", "Cp", "KS", "MS", "NB", "CTAs", "us/iter", "Gadd/s");
    struct { int Cp, KS, MS, NB, groups; } cases[] = {{1024, 16, 4, 32, 2}, {1024, 16, 4, 32, 1}, {1024, 16, 4, 16, 2}, {512, 8, 2, 32, 4},
                                                     {1024, 8, 8, 32, 1}, {768, 12, 3, 16, 1}};
    for (auto c : cases) {
        for (int mode = 0; mode < 3; ++mode) {
            const int grid = c.groups * c.KS * c.MS;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) probe<0><<<grid, 128>>>(d, c.Cp, c.KS, c.MS, c.NB, iters);
                else if (mode == 1) probe<1><<<grid, 128>>>(d, c.Cp, c.KS, c.MS, c.NB, iters);
                else probe<2><<<grid, 128>>>(d, c.Cp, c.KS, c.MS, c.NB, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double us = best * 1e3 / iters, adds = (double)grid * 256 * c.NB;
            printf("%-8s %5d %4d %4d %4d %6d | %10.2f %12.1f\n", mode == 0 ? "red" : mode == 1 ? "red.v4" : "store", c.Cp, c.KS,
                   c.MS, c.NB, grid, us, adds / us * 1e-3);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
