// Micro-benchmark: all-gather of a SLICE-byte block per CTA to every CTA of a 16- (or 8-) CTA cluster,
// (a) with st.async 16-byte remote stores from 128 threads, (b) with cp.async.bulk shared::cta ->
// shared::cluster copies issued by 16 threads (one per destination), completing on the receiver's mbarrier.
// Prints cycles per all-gather round (steady state).   nvcc -arch=sm_100a -o dsmem_bw dsmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rsrgan_b200/csrc/common.cuh"
using namespace rsr;

__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t rbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(rbar) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int G, int slice, int rounds, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x;
    const uint32_t j = cluster_ctarank();
    const uint32_t recv0 = base, recv_bytes = (uint32_t)G * slice;       // two receive buffers
    const uint32_t stage = base + 2 * recv_bytes;                        // local staging slice
    const uint32_t bars = stage + slice;                                 // full0, full1
    if (tid == 0) {
        mbar_init(bars, 1); mbar_init(bars + 8, 1);
        fence_mbar_init();
        mbar_expect_tx(bars, recv_bytes); mbar_expect_tx(bars + 8, recv_bytes);
    }
    __syncthreads();
    cluster_sync_all();
    long long t0 = 0;
    for (int r = 0; r < rounds; ++r) {
        if (r == 4 && tid == 0) t0 = clock64();
        const int buf = r & 1;
        const uint32_t dst = recv0 + buf * recv_bytes + j * slice, dbar = bars + 8 * buf;
        if (MODE == 0) {
            // each thread sends 16-byte pieces: piece index p = tid + 128*i over (slice/16) pieces x G destinations
            const int pieces = slice / 16;
            for (int q = tid; q < pieces * G; q += 128) {
                const int d = q / pieces, pc = q % pieces;
                const uint32_t delta = mapa_u32(base, (uint32_t)d) - base;
                st_async_v4(dst + pc * 16 + delta, r, q, 3u, 4u, dbar + delta);
            }
        } else {
            // staging is "written" by all threads, then one thread per destination issues a bulk copy
            for (int q = tid; q < slice / 16; q += 128) st_shared_v4(stage + q * 16, r, q, 3u, 4u);
            fence_proxy_async_smem();
            __syncthreads();
            if (tid < G) {
                const uint32_t delta = mapa_u32(base, (uint32_t)tid) - base;
                bulk_s2c(dst + delta, stage, (uint32_t)slice, dbar + delta);
            }
        }
        // wait for everyone's slice of this round
        mbar_wait(dbar, (uint32_t)((r >> 1) & 1));
        __syncthreads();
        if (tid == 0 && r + 2 < rounds) mbar_expect_tx(dbar, recv_bytes);
        if (MODE == 1) {   // staging may be rewritten only after our bulk copies have read it: they completed remotely,
                           // which every peer confirmed by finishing its wait -> cheap cluster-wide ordering via next round's data
        }
    }
    if (tid == 0 && blockIdx.x == 0) out[0] = (clock64() - t0) / (rounds - 4);
    cluster_sync_all();
}

template <int MODE>
void run(int G, int slice, const char* name) {
    long long* d; cudaMalloc(&d, 8);
    cudaMemset(d, 0, 8);
    size_t smem = 1024 + 2 * (size_t)G * slice + slice + 64;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G * 4); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, G, slice, 204, d);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-10s G=%2d slice=%5d B (all-gather %6d B/CTA): %6lld cycles/round  -> %.1f B/cycle in  [%s %s]\n", name, G, slice,
           G * slice, h, h ? (double)G * slice / h : 0.0, cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(d);
}

int main() {
    for (int G : {8, 16}) for (int slice : {512, 1024, 2048, 4096}) { run<0>(G, slice, "st.async"); run<1>(G, slice, "bulk"); }
    return 0;
}
