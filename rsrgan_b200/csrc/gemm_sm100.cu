// tcgen05 GEMM with fused epilogue for sm_100a.
//
//   D[M,N] = epi(alpha * A[M,K] * B[K,N])        16-bit operands, fp32 accumulate in TMEM
//
// One CTA computes one 128 x BN output tile.  Warp 0 / lane 0 is the TMA producer
// (cp.async.bulk.tensor into a ring of 128B-swizzled stages), warp 1 / lane 0 issues
// tcgen05.mma (UMMA 128 x BN x 16, cta_group::1) and releases stages with tcgen05.commit,
// then all four warps drain the accumulator with tcgen05.ld (thread <-> row) and apply the
// bias / residual / activation / activation-gradient epilogue in registers.
// Operands may be K-major or MN-major (UMMA descriptors do the transposition), so the same
// kernel serves Y = X W (B MN-major), dX = dY W^T (both K-major) and dW = X^T dY (both
// MN-major) without materialising transposes.
//
// Replaces: tf.contrib.layers.fully_connected (models/lstm.py:82-87,121-124;
// models/discriminator_dnn.py:61-93; models/discriminator_lstm.py:100-104), the x_t half of
// LSTMCell's _Linear (models/lstm.py:90-96) hoisted over all frames, the projection
// (models/BNLSTMCell.py:207-213) hoisted over all frames, and tf.gradients of all of them.
#include "common.cuh"
#include "handle.h"

using namespace rsr;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;

struct GemmKParams {
    int M, N, K;
    int bn, stages, a_mn, b_mn, bf;
    float alpha, beta;
    const float* bias;
    const float* resid; int ldr;
    int act;
    const uint16_t* dsrc; int ldd; int dact;
    float* out32; int ldc32;
    uint16_t* out16; int ldc16;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case RSR_ACT_RELU: return fmaxf(v, 0.0f);
        case RSR_ACT_LRELU: return fmaxf(v, 0.3f * v);
        case RSR_ACT_CLIP: return fminf(fmaxf(v, -0.5f), 1.5f);
        default: return v;
    }
}
// derivative of act evaluated from the activation OUTPUT y (relu / lrelu are sign-preserving)
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    switch (act) {
        case RSR_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
        case RSR_ACT_LRELU: return y > 0.0f ? 1.0f : 0.3f;
        default: return 1.0f;
    }
}

__global__ void __launch_bounds__(128)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int a_stage_bytes = BM * BK * 2;                          // 16 KB
    const int b_boxes = p.b_mn ? (p.bn + 63) / 64 : 1;
    const int b_stage_bytes = p.b_mn ? b_boxes * BK * 128 : p.bn * 128;
    const int stage_bytes = a_stage_bytes + ((b_stage_bytes + 1023) & ~1023);
    const uint32_t bars = smem_base + p.stages * stage_bytes;       // full[s], empty[s], accfull, tmem slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
    const uint32_t acc_bar = bars + 16u * p.stages;
    const uint32_t tmem_slot = acc_bar + 8u;

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.bn) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_acc;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * p.bn;
    const int nkb = (p.K + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            const uint32_t tx = (uint32_t)(a_stage_bytes + b_stage_bytes);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)((kb / p.stages) & 1);
                mbar_wait(empty_bar(s), ph ^ 1u);
                const uint32_t sa = smem_base + s * stage_bytes;
                const uint32_t sb = sa + a_stage_bytes;
                mbar_expect_tx(full_bar(s), tx);
                const int k0 = kb * BK;
                if (!p.a_mn) {
                    tma_load_2d(sa, &tmA, full_bar(s), k0, m0);                 // box [128 rows][64 k]
                } else {
                    tma_load_2d(sa, &tmA, full_bar(s), m0, k0);                 // box [64 k][64 m]
                    tma_load_2d(sa + BK * 128, &tmA, full_bar(s), m0 + 64, k0);
                }
                if (!p.b_mn) {
                    tma_load_2d(sb, &tmB, full_bar(s), k0, n0);                 // box [bn rows][64 k]
                } else {
                    for (int j = 0; j < b_boxes; ++j)
                        tma_load_2d(sb + j * BK * 128, &tmB, full_bar(s), n0 + 64 * j, k0);  // box [64 k][64 n]
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint32_t idesc = umma_idesc(BM, p.bn, p.bf, p.a_mn, p.b_mn);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)((kb / p.stages) & 1);
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = smem_base + s * stage_bytes;
                const uint32_t sb = sa + a_stage_bytes;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t da = p.a_mn ? umma_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                               : umma_desc_sw128(sa + k * 32, 16, 1024);
                    const uint64_t db = p.b_mn ? umma_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                               : umma_desc_sw128(sb + k * 32, 16, 1024);
                    tc_mma_f16(tmem_acc, da, db, idesc, (kb | k) ? 1u : 0u);
                }
                tc_commit(empty_bar(s));       // frees the stage when these MMAs retire
            }
            tc_commit(acc_bar);                // accumulator complete
        }
        __syncwarp();
    }

    // ---------------- epilogue: all 4 warps ----------------
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    const bool row_ok = row < p.M;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.bn; c0 += 16) {
        float v[16];
        tmem_ld16(lane_addr + (uint32_t)c0, v);
        const int n = n0 + c0;
        if (!row_ok || n >= p.N) continue;
        const bool full = (n + 16 <= p.N);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= p.alpha;
        if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (full || n + j < p.N) v[j] += __ldg(p.bias + n + j);
        }
        if (p.resid) {
            const float* r = p.resid + (size_t)row * p.ldr + n;
            if (full) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(r + j);
                    v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                }
            } else {
                for (int j = 0; j < 16; ++j) if (n + j < p.N) v[j] += r[j];
            }
        }
        if (p.act != RSR_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
        }
        if (p.dsrc) {
            const uint16_t* d = p.dsrc + (size_t)row * p.ldd + n;
            if (full) {
                uint4 q0 = *reinterpret_cast<const uint4*>(d);
                uint4 q1 = *reinterpret_cast<const uint4*>(d + 8);
                const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[2 * j] *= act_grad_from_out(h2f((uint16_t)(w[j] & 0xFFFF), p.bf), p.dact);
                    v[2 * j + 1] *= act_grad_from_out(h2f((uint16_t)(w[j] >> 16), p.bf), p.dact);
                }
            } else {
                for (int j = 0; j < 16; ++j)
                    if (n + j < p.N) v[j] *= act_grad_from_out(h2f(d[j], p.bf), p.dact);
            }
        }
        if (p.out32) {   // fp32 output carries the beta accumulation; the 16-bit output below does not
            float* o = p.out32 + (size_t)row * p.ldc32 + n;
            if (full) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 w = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (p.beta != 0.0f) {
                        const float4 q = *reinterpret_cast<const float4*>(o + j);
                        w.x += p.beta * q.x; w.y += p.beta * q.y; w.z += p.beta * q.z; w.w += p.beta * q.w;
                    }
                    *reinterpret_cast<float4*>(o + j) = w;
                }
            } else {
                for (int j = 0; j < 16; ++j)
                    if (n + j < p.N) o[j] = p.beta != 0.0f ? v[j] + p.beta * o[j] : v[j];
            }
        }
        if (p.out16) {
            uint16_t* o = p.out16 + (size_t)row * p.ldc16 + n;
            if (full) {
                uint4 q0, q1;
                q0.x = pack2(v[0], v[1], p.bf);  q0.y = pack2(v[2], v[3], p.bf);
                q0.z = pack2(v[4], v[5], p.bf);  q0.w = pack2(v[6], v[7], p.bf);
                q1.x = pack2(v[8], v[9], p.bf);  q1.y = pack2(v[10], v[11], p.bf);
                q1.z = pack2(v[12], v[13], p.bf); q1.w = pack2(v[14], v[15], p.bf);
                *reinterpret_cast<uint4*>(o) = q0;
                *reinterpret_cast<uint4*>(o + 8) = q1;
            } else {
                for (int j = 0; j < 16; ++j) if (n + j < p.N) o[j] = f2h(v[j], p.bf);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_acc, tmem_cols);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
extern "C" int rsr_version(void) { return 100; }

extern "C" int rsr_create(rsr_handle** out, int device, int dtype) {
    if (!out || (dtype != RSR_DTYPE_F16 && dtype != RSR_DTYPE_BF16)) return RSR_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return RSR_E_NODEV;
    cudaDeviceProp prop;
    RSR_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return RSR_E_NODEV;   // sm_100a binary only; no fallback path exists
    RSR_CHECK_CUDA(cudaSetDevice(device));
    rsr_handle* h = new rsr_handle();
    h->device = device;
    h->dtype = dtype;
    h->num_sms = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { delete h; return RSR_E_NODEV; }
    h->encode = (PFN_encodeTiled)fn;
    e = cudaMalloc(&h->flags, RSR_FLAG_WORDS * sizeof(unsigned int));
    if (e != cudaSuccess) { delete h; return (int)e; }
    cudaMemset(h->flags, 0, RSR_FLAG_WORDS * sizeof(unsigned int));
    *out = h;
    return 0;
}

extern "C" int rsr_destroy(rsr_handle* h) {
    if (!h) return RSR_E_ARG;
    if (h->flags) cudaFree(h->flags);
    delete h;
    return 0;
}

extern "C" int rsr_num_sms(rsr_handle* h) { return h ? h->num_sms : RSR_E_ARG; }

int rsr_get_tmap(rsr_handle* h, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld,
                 uint32_t b0, uint32_t b1, CUtensorMap* out) {
    if (((uintptr_t)ptr & 15) || (ld * 2) % 16 || b0 * 2 != 128 || b1 > 256 || d0 == 0 || d1 == 0) return RSR_E_ARG;
    TmapKey key{ptr, d0, d1, ld * 2, b0, b1};
    {
        std::lock_guard<std::mutex> g(h->mu);
        auto it = h->tmaps.find(key);
        if (it != h->tmaps.end()) { *out = it->second; return 0; }
    }
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = h->encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return RSR_E_ARG;
    {
        std::lock_guard<std::mutex> g(h->mu);
        if (h->tmaps.size() > 4096) h->tmaps.clear();
        h->tmaps[key] = m;
    }
    *out = m;
    return 0;
}

// ---------------------------------------------------------------------------------------
// rsr_gemm
// ---------------------------------------------------------------------------------------
extern "C" int rsr_gemm(rsr_handle* h, void* stream, const rsr_gemm_args* a) {
    if (!h || !a || !a->A || !a->B) return RSR_E_ARG;
    if (a->M <= 0 || a->N <= 0 || a->K <= 0) return RSR_E_ARG;
    if (!a->out32 && !a->out16) return RSR_E_ARG;
    if ((a->lda & 7) || (a->ldb & 7)) return RSR_E_SHAPE;
    if (a->out32 && ((a->ldc32 & 3) || ((uintptr_t)a->out32 & 15))) return RSR_E_SHAPE;
    if (a->out16 && ((a->ldc16 & 7) || ((uintptr_t)a->out16 & 15))) return RSR_E_SHAPE;
    if (a->resid && ((a->ldr & 3) || ((uintptr_t)a->resid & 15))) return RSR_E_SHAPE;
    if (a->dact_src && ((a->ldd & 7) || ((uintptr_t)a->dact_src & 15))) return RSR_E_SHAPE;

    GemmKParams p;
    p.M = a->M; p.N = a->N; p.K = a->K;
    int bn = a->tile_n;
    if (bn <= 0) {
        // wide tiles for big problems; for small N take N rounded up to 16
        bn = 128;
        if (a->N < 128) bn = (a->N + 15) & ~15;
    }
    if (bn < 16 || bn > 256 || (bn & 15)) return RSR_E_SHAPE;
    p.bn = bn;
    p.a_mn = a->a_mn ? 1 : 0; p.b_mn = a->b_mn ? 1 : 0; p.bf = h->dtype == RSR_DTYPE_BF16;
    p.alpha = a->alpha; p.beta = a->beta;
    p.bias = a->bias; p.resid = a->resid; p.ldr = a->ldr; p.act = a->act;
    p.dsrc = (const uint16_t*)a->dact_src; p.ldd = a->ldd; p.dact = a->dact;
    p.out32 = a->out32; p.ldc32 = a->ldc32; p.out16 = (uint16_t*)a->out16; p.ldc16 = a->ldc16;

    const int a_stage = BM * BK * 2;
    const int b_boxes = p.b_mn ? (bn + 63) / 64 : 1;
    const int b_stage = p.b_mn ? b_boxes * BK * 128 : bn * 128;
    const int stage_bytes = a_stage + ((b_stage + 1023) & ~1023);
    const int nkb = (a->K + BK - 1) / BK;
    int stages = nkb < 4 ? nkb : 4;
    // keep two CTAs per SM resident when the tile allows it (epilogue of one overlaps mainloop of the other)
    while (stages > 2 && stages * stage_bytes > 100 * 1024) --stages;
    p.stages = stages;
    const int smem = stages * stage_bytes + 1024 /*align slack*/ + 16 * stages + 32;
    if (smem > h->max_smem) return RSR_E_SHAPE;

    CUtensorMap tmA, tmB;
    int rc;
    if (!p.a_mn) rc = rsr_get_tmap(h, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, 64, BM, &tmA);
    else         rc = rsr_get_tmap(h, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK, &tmA);
    if (rc) return rc;
    if (!p.b_mn) rc = rsr_get_tmap(h, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, 64, (uint32_t)bn, &tmB);
    else         rc = rsr_get_tmap(h, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK, &tmB);
    if (rc) return rc;

    static bool attr_set = false;
    if (!attr_set) {
        RSR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_set = true;
    }
    dim3 grid((a->N + bn - 1) / bn, (a->M + BM - 1) / BM);
    gemm_tcgen05_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
    RSR_LAUNCH_CHECK();
    return 0;
}
