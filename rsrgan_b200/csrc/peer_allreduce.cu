// average_gradients (utils/ops.py:343-376: the tower-mean of every gradient, "a sync point across all towers")
// as ONE kernel over NVLink peer memory: every rank's flat gradient buffer lives in a cudaMalloc'ed block that
// the other ranks of the node have opened through CUDA IPC, so a rank reads its 1/N slice of all N buffers with
// plain loads (NVSwitch: full bandwidth to every peer), sums it in rank order and stores the sum into all N
// buffers.  The 1/N of the mean stays folded into the update kernel.  Per rank (N - 1)/N of the buffer crosses
// the links in each direction -- the two-shot all-reduce's minimum -- and the whole exchange is two flag
// barriers plus one pass, which replaces NCCL's ring/NVLS kernel and its launch latency (measured numbers:
// DESIGN.md section 5).  Every slice is reduced by exactly one rank in the fixed order 0..N-1, so all replicas
// hold bit-identical sums.
//
// Memory layout of a rank's block:  [ header 16 KB | data ... ]
//   header  uint32 flags[PEER_MAX_BLOCKS][PEER_MAX]   written by the PEERS (slot [block][writer rank])
//           uint32 count[PEER_MAX_BLOCKS]             local: barriers this block has passed (monotonic)
//           uint32 error                              local: a barrier wait timed out
// Barrier (per block index, between the same-index blocks of all ranks): thread p stores the block's next count
// into rank p's flags[block][me] with release.sys and spins on its own flags[block][p] with acquire.sys; counts
// only grow, so a rank that is one barrier ahead can never be mistaken for a late one and nothing is ever reset
// (CUDA-graph replays need no host involvement).  A kernel starts only after the producer kernels of its stream
// have finished, so "block b of rank p has arrived" implies rank p's gradients are complete.
#include <cstring>

#include "common.cuh"
#include "handle.h"

namespace {

constexpr int PEER_MAX = RSR_PEER_MAX_RANKS;
constexpr int PEER_MAX_BLOCKS = 148;
constexpr int PEER_THREADS = 512;
constexpr long long PEER_TIMEOUT_CYCLES = 60000000000LL;    // ~30 s (ranks may be seconds apart around a graph capture);
                                                            // a dead peer must not hang the GPU for good

struct PeerHeader {
    uint32_t flags[PEER_MAX_BLOCKS][PEER_MAX];
    uint32_t count[PEER_MAX_BLOCKS];
    uint32_t error;
};
static_assert(sizeof(PeerHeader) <= RSR_PEER_HEADER_BYTES, "header does not fit");

struct PeerParams {
    char* base[PEER_MAX];       // every rank's block (own at [rank])
    long long data_off;         // byte offset of the buffer inside each block
    long long n4;               // float4 elements
    int rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer data: written by kernels of OTHER devices; read once, straight from the home L2
__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float4* p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// P.base[] indexed by a run-time value would put the parameter struct on the local stack: select with constant indices
template <int N>
__device__ __forceinline__ char* base_of(const PeerParams& P, int r) {
    char* b = P.base[0];
#pragma unroll
    for (int q = 1; q < N; ++q) b = r == q ? P.base[q] : b;
    return b;
}

template <int N>
__device__ __forceinline__ void peer_barrier(const PeerParams& P, PeerHeader* mine, uint32_t target) {
    __syncthreads();                                   // this block's loads / stores precede the release below
    const int p = threadIdx.x;
    if (p < N) {
        PeerHeader* theirs = reinterpret_cast<PeerHeader*>(base_of<N>(P, p));
        st_release_sys(&theirs->flags[blockIdx.x][P.rank], target);
        const uint32_t* w = &mine->flags[blockIdx.x][p];
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(w) - target) < 0) {
            if (clock64() - t0 > PEER_TIMEOUT_CYCLES) { mine->error = 1u; break; }
        }
    }
    __syncthreads();
}

// U float4 per thread and pass, all N * U loads issued before the first add: the pass is one NVLink round trip deep
// (N * U = 8 -> 64 KB in flight per SM, ~9 MB per GPU against the ~3 MB that 900 GB/s x 3 us need)
template <int N, int U>
__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(PeerParams P) {
    PeerHeader* mine = reinterpret_cast<PeerHeader*>(base_of<N>(P, P.rank));
    const uint32_t c0 = mine->count[blockIdx.x];
    peer_barrier<N>(P, mine, c0 + 1);                  // every rank's gradients are complete
    const long long per = (P.n4 + N - 1) / N;
    const long long lo = (long long)P.rank * per;
    const long long hi = lo + per < P.n4 ? lo + per : P.n4;
    const long long stride = (long long)gridDim.x * PEER_THREADS;
    for (long long i0 = lo + (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i0 < hi; i0 += stride * U) {
        float4 v[U][N];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
#pragma unroll
            for (int p = 0; p < N; ++p)
                if (i < hi) v[u][p] = ld_peer(reinterpret_cast<const float4*>(P.base[p] + P.data_off) + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i >= hi) break;
            float4 a = v[u][0];
#pragma unroll
            for (int p = 1; p < N; ++p) { a.x += v[u][p].x; a.y += v[u][p].y; a.z += v[u][p].z; a.w += v[u][p].w; }
#pragma unroll
            for (int p = 0; p < N; ++p) st_peer(reinterpret_cast<float4*>(P.base[p] + P.data_off) + i, a);
        }
    }
    peer_barrier<N>(P, mine, c0 + 2);                  // every rank's sums have landed in every buffer
    if (threadIdx.x == 0) mine->count[blockIdx.x] = c0 + 2;
}

}  // namespace

extern "C" int rsr_peer_alloc(rsr_handle* h, long long data_bytes, void** block, unsigned char* ipc_handle) {
    if (!h || !block || !ipc_handle || data_bytes <= 0) return RSR_E_ARG;
    void* p = nullptr;
    const size_t total = (size_t)RSR_PEER_HEADER_BYTES + (size_t)data_bytes;
    cudaError_t e = cudaMalloc(&p, total);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, total);
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    cudaIpcMemHandle_t hd;
    e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    static_assert(sizeof(hd) == RSR_PEER_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(ipc_handle, &hd, sizeof(hd));
    *block = p;
    return 0;
}

extern "C" int rsr_peer_open(rsr_handle* h, const unsigned char* ipc_handle, void** block) {
    if (!h || !ipc_handle || !block) return RSR_E_ARG;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle, sizeof(hd));
    const cudaError_t e = cudaIpcOpenMemHandle(block, hd, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int rsr_peer_close(rsr_handle* h, void* block) {
    if (!h || !block) return RSR_E_ARG;
    const cudaError_t e = cudaIpcCloseMemHandle(block);
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int rsr_peer_free(rsr_handle* h, void* block) {
    if (!h || !block) return RSR_E_ARG;
    const cudaError_t e = cudaFree(block);
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int rsr_peer_error(rsr_handle* h, const void* block, int* error) {
    if (!h || !block || !error) return RSR_E_ARG;
    uint32_t v = 0;
    const cudaError_t e = cudaMemcpy(&v, (const char*)block + offsetof(PeerHeader, error), sizeof(v), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    *error = (int)v;
    return 0;
}

extern "C" int rsr_peer_allreduce(rsr_handle* h, void* stream, void* const* blocks, int rank, int world,
                                  long long data_off_bytes, long long n_floats, int max_blocks) {
    if (!h || !blocks || world < 1 || world > PEER_MAX || rank < 0 || rank >= world || n_floats <= 0) return RSR_E_ARG;
    if (world != 1 && world != 2 && world != 4 && world != 8) return RSR_E_SHAPE;
    if ((n_floats & 3) || (data_off_bytes & 15) || data_off_bytes < RSR_PEER_HEADER_BYTES) return RSR_E_SHAPE;
    if (world == 1) return 0;
    PeerParams P;
    for (int p = 0; p < PEER_MAX; ++p) P.base[p] = p < world ? (char*)blocks[p] : nullptr;
    for (int p = 0; p < world; ++p)
        if (!P.base[p]) return RSR_E_ARG;
    P.data_off = data_off_bytes; P.n4 = n_floats / 4; P.rank = rank;
    // the SAME grid on every rank (flags are indexed by block): a function of the size and world only
    const long long per = (P.n4 + world - 1) / world;
    const int unroll = 8 / world;
    long long grid = (per + (long long)PEER_THREADS * unroll - 1) / ((long long)PEER_THREADS * unroll);
    const int cap = max_blocks > 0 && max_blocks < PEER_MAX_BLOCKS ? max_blocks : PEER_MAX_BLOCKS;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (world == 2) peer_allreduce_kernel<2, 4><<<(int)grid, PEER_THREADS, 0, st>>>(P);
    else if (world == 4) peer_allreduce_kernel<4, 2><<<(int)grid, PEER_THREADS, 0, st>>>(P);
    else peer_allreduce_kernel<8, 1><<<(int)grid, PEER_THREADS, 0, st>>>(P);
    RSR_LAUNCH_CHECK();
    return 0;
}
