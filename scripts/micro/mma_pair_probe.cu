// Micro-benchmark + known-answer test for the weight-stationary MMA the recurrence kernels are built on:
//   D[tmem] = A[tmem, K-major, resident] * B[smem, K-major, un-swizzled core-matrix layout [k/8][row][8]]
// issued (a) by one CTA (cta_group::1, M = 128) and (b) by a CTA PAIR (cta_group::2, M = 256: each CTA holds
// its own 128 rows of A in its TMEM and HALF of the B rows in its shared memory).  Prints the cycles of one
// dependent chain of K/16 MMAs + commit + wait for N = 16 / 32 / 64, and checks the accumulator against an
// exact integer reference (operands are small multiples of 1/8: every product and sum is exact in fp32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_pair_probe mma_pair_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rsrgan_b200/csrc/common.cuh"
using namespace rsr;

__device__ __forceinline__ void tc_mma_f16_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__host__ __device__ inline float a_val(int r, int k) { return (float)((r * 7 + k * 3) % 13 - 6) * 0.125f; }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k) % 11 - 5) * 0.25f; }

template <int CG>
__global__ void __launch_bounds__(128, 1) probe(int N, int K, int reps, long long* out_cycles, int* out_bad) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const int R = N / CG;                                  // B rows held by this CTA
    const uint32_t sB = base;                              // [K/8][R][8] 16-bit
    const uint32_t sB_bytes = (uint32_t)K * R * 2u;
    const uint32_t barM = base + sB_bytes, tslot = barM + 8;
    if (tid == 0) { mbar_init(barM, 1); fence_mbar_init(); }
    if (warp == 0) { if (CG == 2) tmem_alloc_2cta(tslot, 512); else tmem_alloc(tslot, 512); }
    // B operand: row n (global utterance index), hypothesis: CTA `rank` holds rows [rank * R, rank * R + R)
    for (int i = tid; i < K * R; i += 128) {
        const int k = i / R, n = i % R;                    // iterate so that consecutive threads take consecutive rows
        const uint32_t off = (uint32_t)(k >> 3) * (uint32_t)(R * 16) + (uint32_t)n * 16u + (uint32_t)(k & 7) * 2u;
        *reinterpret_cast<__half*>(base_ptr + off) = __float2half_rn(b_val((int)rank * R + n, k));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    {   // A operand -> TMEM: thread <-> row, 2 k per 32-bit column
        const int r = (int)rank * 128 + tid;
        for (int cb = 0; cb < K / 64; ++cb) {
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int k = cb * 64 + 2 * c;
                const __half2 h2 = __floats2half2_rn(a_val(r, k), a_val(r, k + 1));
                v[c] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb * 32u, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t acc = tmem + (uint32_t)K / 2u;
    const uint32_t idesc = umma_idesc(128 * CG, N, 0, 0, 0);
    const uint64_t db0 = umma_desc_nosw(sB, (uint32_t)R * 16u, 128u);
    long long cyc = 0;
    for (int rep = 0; rep < reps; ++rep) {
        const long long ta_ = clock64();
        if (warp == 0 && rank == 0) {
            if (elect_one_sync()) {
                uint64_t db = db0;
                uint32_t ta = tmem;
                for (int kk = 0; kk < K / 16; ++kk) {
                    if (CG == 2) tc_mma_f16_ts_2cta(acc, ta, db, idesc, kk ? 1u : 0u);
                    else tc_mma_f16_ts(acc, ta, db, idesc, kk ? 1u : 0u);
                    ta += 8u;
                    db += (uint64_t)((2u * (uint32_t)R * 16u) >> 4);
                }
                if (CG == 2) tc_commit_2cta_mc(barM, (uint16_t)3); else tc_commit(barM);
            }
        }
        __syncwarp();
        mbar_wait(barM, (uint32_t)(rep & 1));
        tc_fence_after();
        if (rep >= 2) cyc += clock64() - ta_;              // issue + execution + commit -> wake-up, without the sync below
        if (CG == 2) cluster_sync_all();                   // the leader must not re-arm the barrier before the peer has seen the phase
        else __syncthreads();
    }
    // verify
    int bad = 0;
    {
        const int r = (int)rank * 128 + tid;
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tmem_ld16(acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int j = 0; j < 16; ++j) {
                float ref = 0.f;
                for (int k = 0; k < K; ++k) ref += a_val(r, k) * b_val(c0 + j, k);
                if (v[j] != ref) ++bad;
            }
        }
    }
    atomicAdd(out_bad, bad);
    if (tid == 0 && blockIdx.x == 0) out_cycles[0] = cyc / (reps - 2);
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) { if (CG == 2) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int CG>
void run(int N, int K) {
    long long* d; int* b;
    cudaMalloc(&d, 8); cudaMalloc(&b, 4);
    cudaMemset(d, 0, 8); cudaMemset(b, 0, 4);
    size_t smem = 1024 + (size_t)K * (N / CG) * 2 + 64;
    cudaFuncSetAttribute(probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG>, N, K, 52, d, b);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h = 0; int hb = -1;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hb, b, 4, cudaMemcpyDeviceToHost);
    printf("cta_group::%d M=%3d N=%3d K=%4d: %6lld cycles per chain of %2d MMAs (%5.1f / MMA), mismatches %d  [%s %s]\n", CG,
           128 * CG, N, K, h, K / 16, (double)h / (K / 16), hb, cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(d); cudaFree(b);
}

int main() {
    for (int K : {256, 512, 768}) {
        if (K > 768) continue;
        for (int N : {16, 32, 64}) { run<1>(N, K); run<2>(N, K); }
    }
    return 0;
}
