// Per-rank library handle: device info, driver entry point for TMA descriptor
// encoding, a cache of encoded tensor maps, and the small device workspace the
// persistent recurrent kernels use for their group barriers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <unordered_map>
#include <string>

#include "../../include/rsrgan_b200.h"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapKey {
    const void* ptr; uint64_t d0, d1, stride1; uint32_t b0, b1, kind;   // kind = element bytes | swizzle bytes << 8
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && stride1 == o.stride1 && b0 == o.b0 && b1 == o.b1 &&
               kind == o.kind;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
        h ^= k.d0 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= k.d1 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= k.stride1 + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= ((uint64_t)k.b0 << 32 | k.b1) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h ^= (uint64_t)k.kind + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

struct rsr_handle {
    int device = 0;
    int dtype = RSR_DTYPE_F16;   // 16-bit operand type
    int num_sms = 0;
    int max_smem = 0;
    PFN_encodeTiled encode = nullptr;
    std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> tmaps;
    std::mutex mu;
    unsigned int* flags = nullptr;   // device: group-barrier counters, RSR_FLAG_WORDS words
    int flag_cursor = 0;
    float* partials = nullptr;       // device: per-1024-block partial sums of rsr_seg_sumsq, RSR_PARTIAL_WORDS floats
    // co-resident clusters of the cluster recurrence kernels: [fwd|bwd][Cp/256 - 1][NB 16|32]; -1 = not queried yet
    int gemm2_pairs = -1;            // co-resident CTA pairs of the two-CTA GEMM (-1 = not queried yet)
    int fused_ik[2] = {-1, -1};      // Ik the fused-forward capacity entry was computed for
    int pair_cap[2][2] = {{-1, -1}, {-1, -1}};   // CTA-pair kernels (lstmp_pair_sm100.cu): [fwd|bwd][Cp/256 - 1]
    int pair_ik[2] = {-1, -1};
    // layer-wavefront kernels: [fwd NBP 32 | fwd 48 | bwd 32 | bwd 48][Cp/256 - 1]
    int wave_cap[4][2] = {{-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}};
    int wave_key[4][2] = {{-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}};
    int cluster_cap[3][2][2] = {{{-1, -1}, {-1, -1}}, {{-1, -1}, {-1, -1}}, {{-1, -1}, {-1, -1}}};   // [2] = fused fwd
};

#define RSR_FLAG_WORDS 4096
#define RSR_PARTIAL_WORDS (1 << 17)   // 128 Mi parameters per network

// `n` zero-initialised-by-the-caller counters of the group-barrier workspace (ring allocation: a launch's counters are
// dead long before the cursor wraps)
inline unsigned int* rsr_take_flags(rsr_handle* h, int n) {
    std::lock_guard<std::mutex> g(h->mu);
    if (h->flag_cursor + n > RSR_FLAG_WORDS) h->flag_cursor = 0;
    unsigned int* f = h->flags + h->flag_cursor;
    h->flag_cursor += n;
    return f;
}

// zeroes `n` counters with a one-block kernel (lstmp_sm100.cu).  A cudaMemsetAsync node in front of a kernel costs ~13 us of
// dependency latency inside a captured graph (profiles/r2_timeline_graph_cfg2_v2.txt), a kernel node ~0.5 us.
int rsr_zero_u32(unsigned int* p, int n, cudaStream_t stream);

// Dynamic shared memory floors that keep TMEM owners apart when kernels of different streams overlap:
// two recurrence CTAs (each allocates up to all 512 TMEM columns and waits on its cluster) must never share
// an SM (cross-cluster alloc waits could deadlock), and a GEMM CTA must not share one with either.
#define RSR_EXCLUSIVE_SMEM_REC (120 * 1024)
#define RSR_EXCLUSIVE_SMEM_GEMM (160 * 1024)

// 2-D tensor map over a 16-bit row-major matrix [d1 rows, d0 cols] with row pitch `ld` elements,
// box [b1 rows, b0 cols], 128-byte swizzle (b0 * 2 bytes must be 128).
int rsr_get_tmap(rsr_handle* h, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld,
                 uint32_t b0, uint32_t b1, CUtensorMap* out);
// general form: elem_bytes in {2, 4}; swizzle in {64, 128} bytes and b0 * elem_bytes == swizzle
// (TMA store targets of the GEMM epilogue: fp32 boxes of 32 columns / 16-bit boxes of 32 columns).
int rsr_get_tmap_ex(rsr_handle* h, const void* ptr, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t ld,
                    uint32_t b0, uint32_t b1, int swizzle, CUtensorMap* out);

// cluster (DSMEM) variants of the recurrence, lstmp_cluster_sm100.cu; RSR_E_RESIDENT = not applicable
int rsr_lstmp_fwd_cluster(rsr_handle* h, void* stream, int B, int T, int Cp, const float* zx, const void* wcT,
                          const float* w_i, const float* w_f, const float* w_o, float forget_bias,
                          const int* lengths, void* mt_seq, float* save);
int rsr_lstmp_fused_fwd_cluster(rsr_handle* h, void* stream, int B, int T, int I, int Cp, const void* x16, int ldx,
                                const void* kxT, const float* bias, const void* wcT, const float* w_i,
                                const float* w_f, const float* w_o, float forget_bias, const int* lengths,
                                void* mt_seq, float* save);
// CTA-pair (cta_group::2) variants, lstmp_pair_sm100.cu; RSR_E_RESIDENT = not applicable
int rsr_lstmp_fused_fwd_pair(rsr_handle* h, void* stream, int B, int T, int I, int Cp, const void* x16, int ldx,
                             const void* kxT, const float* bias, const void* wcT, const float* w_i,
                             const float* w_f, const float* w_o, float forget_bias, const int* lengths,
                             void* mt_seq, float* save);
int rsr_lstmp_bwd_pair(rsr_handle* h, void* stream, int B, int T, int Cp, const float* dmt, const void* wc,
                       const float* w_i, const float* w_f, const float* w_o, const int* lengths,
                       const float* save, void* dz16, float* dbias, float* dw_i, float* dw_f, float* dw_o);
int rsr_lstmp_bwd_cluster(rsr_handle* h, void* stream, int B, int T, int Cp, const float* dmt, const void* wc,
                          const float* w_i, const float* w_f, const float* w_o, const int* lengths,
                          const float* save, void* dz16, float* dbias, float* dw_i, float* dw_f, float* dw_o);

#define RSR_CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define RSR_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)
