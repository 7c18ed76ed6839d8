// tcgen05 GEMM with fused epilogue for sm_100a.
//
//   D[M,N] = epi(alpha * A[M,K] * B[K,N])        16-bit operands, fp32 accumulate in TMEM
//
// Persistent kernel, one CTA per SM, 128 x BN output tiles (BN up to 256).  Warp 0 / lane 0 is the
// TMA producer (cp.async.bulk.tensor into a ring of 128B-swizzled stages), warp 1 / lane 0 issues
// tcgen05.mma (UMMA 128 x BN x 16, cta_group::1) into one of two TMEM accumulators and releases
// stages with tcgen05.commit, warps 2-5 drain the other accumulator with tcgen05.ld, transpose it
// through shared memory so that every global access of the bias / residual / activation /
// activation-gradient epilogue is coalesced, and hand the accumulator back.  Weight-gradient
// products (K = all frames of the minibatch, few output tiles) are split along K across CTAs.
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
constexpr int GEMM_THREADS = 192;      // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int EPI_PITCH = 33;          // fp32 words per row of the per-warp transpose tile

struct GemmKParams {
    int M, N, K;
    int bn, stages, a_mn, b_mn, bf;
    int m_tiles, n_tiles, splits, acc_stride, tmem_cols;
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

// Persistent, warp-specialised GEMM.  Each CTA (one per SM) walks tiles tile = blockIdx.x + i*gridDim.x of
// the (k-split, m, n) tile space.  Three pipelines run concurrently: TMA -> smem ring (full/empty
// mbarriers), tcgen05.mma -> double-buffered TMEM accumulator (tfull/tempty), and the epilogue warps,
// which drain accumulator `a` while the MMA warp already fills accumulator `a^1` for the next tile.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B-swizzle atoms need 1024 B alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int a_stage_bytes = BM * BK * 2;                          // 16 KB
    const int b_boxes = p.b_mn ? (p.bn + 63) / 64 : 1;
    const int b_stage_bytes = p.b_mn ? b_boxes * BK * 128 : p.bn * 128;
    const int stage_bytes = a_stage_bytes + ((b_stage_bytes + 1023) & ~1023);
    const uint32_t epi_off = (uint32_t)(p.stages * stage_bytes);
    const uint32_t bars = smem_base + epi_off + 4u * 32u * EPI_PITCH * 4u;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
    auto tfull_bar = [&](int a) { return bars + 16u * p.stages + 8u * a; };
    auto tempty_bar = [&](int a) { return bars + 16u * p.stages + 16u + 8u * a; };
    const uint32_t tmem_slot = bars + 16u * p.stages + 32u;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int nkb = (p.K + BK - 1) / BK;
    const int tiles_mn = p.m_tiles * p.n_tiles;
    const int total = tiles_mn * p.splits;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            const uint32_t tx = (uint32_t)(a_stage_bytes + b_stage_bytes);
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int ks = tile / tiles_mn, mn = tile % tiles_mn;
                const int m0 = (mn / p.n_tiles) * BM, n0 = (mn % p.n_tiles) * p.bn;
                const int kb0 = (int)((long long)ks * nkb / p.splits), kb1 = (int)((long long)(ks + 1) * nkb / p.splits);
                for (int kb = kb0; kb < kb1; ++kb) {
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
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint32_t idesc = umma_idesc(BM, p.bn, p.bf, p.a_mn, p.b_mn);
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int ks = tile / tiles_mn;
                const int kb0 = (int)((long long)ks * nkb / p.splits), kb1 = (int)((long long)(ks + 1) * nkb / p.splits);
                mbar_wait(tempty_bar(acc), aph ^ 1u);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * p.acc_stride);
                for (int kb = kb0; kb < kb1; ++kb) {
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
                        tc_mma_f16(tacc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(empty_bar(s));       // frees the stage when these MMAs retire
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                tc_commit(tfull_bar(acc));         // accumulator complete
                if (++acc == 2) { acc = 0; aph ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue warps (TMEM -> registers -> smem transpose -> coalesced global) --------
        const int q = warp & 3;                                // TMEM lane quadrant this warp may access
        float* st = reinterpret_cast<float*>(smem_gen + epi_off) + (warp - 2) * 32 * EPI_PITCH;
        const int cpair = 2 * (lane & 15), rsub = lane >> 4;   // read-back mapping: 2 columns x (row parity)
        int acc = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int mn = tile % tiles_mn;
            const int m0 = (mn / p.n_tiles) * BM, n0 = (mn % p.n_tiles) * p.bn;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(acc * p.acc_stride) + ((uint32_t)(q * 32) << 16);
            const int rbase = m0 + q * 32;
            for (int c0 = 0; c0 < p.bn; c0 += 32) {
                float v[32];
                __syncwarp();                                  // reconverge: tcgen05.ld is .sync.aligned
                if (c0 + 32 <= p.bn) {
                    tmem_ld32(taddr + (uint32_t)c0, v);
                } else {                                       // bn is a multiple of 16
                    tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 16; j < 32; ++j) v[j] = 0.f;
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; ++j) st[lane * EPI_PITCH + j] = v[j];
                __syncwarp();
                const int col = n0 + c0 + cpair;
                const bool live = (col < p.N) && (c0 + cpair < p.bn);
                const bool two = (col + 1 < p.N);
                float b0 = 0.f, b1 = 0.f;
                if (p.bias && live) { b0 = __ldg(p.bias + col); if (two) b1 = __ldg(p.bias + col + 1); }
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int r = 2 * i + rsub;
                    const int row = rbase + r;
                    if (row >= p.M || !live) break;
                    float x0 = st[r * EPI_PITCH + cpair] * p.alpha + b0;
                    float x1 = st[r * EPI_PITCH + cpair + 1] * p.alpha + b1;
                    if (p.resid) {
                        const float* rp = p.resid + (size_t)row * p.ldr + col;
                        if (two) { const float2 t = *reinterpret_cast<const float2*>(rp); x0 += t.x; x1 += t.y; }
                        else x0 += rp[0];
                    }
                    if (p.act != RSR_ACT_NONE) { x0 = apply_act(x0, p.act); x1 = apply_act(x1, p.act); }
                    if (p.dsrc) {
                        const uint16_t* dp = p.dsrc + (size_t)row * p.ldd + col;
                        if (two) {
                            const uint32_t w = *reinterpret_cast<const uint32_t*>(dp);
                            x0 *= act_grad_from_out(h2f((uint16_t)(w & 0xFFFF), p.bf), p.dact);
                            x1 *= act_grad_from_out(h2f((uint16_t)(w >> 16), p.bf), p.dact);
                        } else {
                            x0 *= act_grad_from_out(h2f(dp[0], p.bf), p.dact);
                        }
                    }
                    if (p.out16) {
                        uint16_t* o = p.out16 + (size_t)row * p.ldc16 + col;
                        if (two) *reinterpret_cast<uint32_t*>(o) = pack2(x0, x1, p.bf);
                        else o[0] = f2h(x0, p.bf);
                    }
                    if (p.out32) {
                        float* o = p.out32 + (size_t)row * p.ldc32 + col;
                        if (p.splits > 1) {                    // split-K partial sums meet in L2 (out32 += ...)
                            red_add_f32(o, x0);
                            if (two) red_add_f32(o + 1, x1);
                        } else if (two) {
                            float2 w = make_float2(x0, x1);
                            if (p.beta != 0.0f) { const float2 t = *reinterpret_cast<const float2*>(o); w.x += p.beta * t.x; w.y += p.beta * t.y; }
                            *reinterpret_cast<float2*>(o) = w;
                        } else {
                            o[0] = p.beta != 0.0f ? x0 + p.beta * o[0] : x0;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));       // accumulator may be overwritten
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

    const bool plain_accumulate = a->out32 && !a->out16 && !a->bias && !a->resid && a->act == RSR_ACT_NONE &&
                                  !a->dact_src && a->beta == 1.0f;
    GemmKParams p;
    p.M = a->M; p.N = a->N; p.K = a->K;
    const int m_tiles = (a->M + BM - 1) / BM;
    int bn = a->tile_n;
    if (bn <= 0) {
        bn = 128;
        if (a->N < 128) bn = (a->N + 15) & ~15;
        // wide tiles when the problem still fills the machine twice over
        else if (a->N % 256 == 0 && (long long)m_tiles * (a->N / 256) >= 2LL * h->num_sms) bn = 256;
    }
    if (bn < 16 || bn > 256 || (bn & 15)) return RSR_E_SHAPE;
    p.bn = bn;
    p.a_mn = a->a_mn ? 1 : 0; p.b_mn = a->b_mn ? 1 : 0; p.bf = h->dtype == RSR_DTYPE_BF16;
    p.alpha = a->alpha; p.beta = a->beta;
    p.bias = a->bias; p.resid = a->resid; p.ldr = a->ldr; p.act = a->act;
    p.dsrc = (const uint16_t*)a->dact_src; p.ldd = a->ldd; p.dact = a->dact;
    p.out32 = a->out32; p.ldc32 = a->ldc32; p.out16 = (uint16_t*)a->out16; p.ldc16 = a->ldc16;
    p.m_tiles = m_tiles;
    p.n_tiles = (a->N + bn - 1) / bn;

    const int a_stage = BM * BK * 2;
    const int b_boxes = p.b_mn ? (bn + 63) / 64 : 1;
    const int b_stage = p.b_mn ? b_boxes * BK * 128 : bn * 128;
    const int stage_bytes = a_stage + ((b_stage + 1023) & ~1023);
    const int nkb = (a->K + BK - 1) / BK;
    // split-K: only for "out32 += A B" (weight gradients: few output tiles, K = all frames); partial
    // sums are added into out32 with red.global.add.f32, which is the beta = 1 semantics
    int splits = a->split_k;
    const int tiles_mn = p.m_tiles * p.n_tiles;
    if (splits <= 0) {
        splits = 1;
        if (plain_accumulate && tiles_mn < h->num_sms && nkb >= 8) {
            splits = (h->num_sms + tiles_mn - 1) / tiles_mn;
            if (splits > nkb / 4) splits = nkb / 4;
            if (splits < 1) splits = 1;
        }
    }
    if (splits > 1 && !plain_accumulate) return RSR_E_ARG;
    if (splits > nkb) splits = nkb;
    p.splits = splits;
    const int fixed = 1024 /*align slack*/ + 4 * 32 * EPI_PITCH * 4 + 64 + 16 * 8;
    int stages = (h->max_smem - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    const int kb_per_tile = (nkb + splits - 1) / splits;
    const long long total_tiles = (long long)tiles_mn * splits;
    const int grid = (int)(total_tiles < h->num_sms ? total_tiles : h->num_sms);
    const long long kb_per_cta = (long long)kb_per_tile * ((total_tiles + grid - 1) / grid);
    if (stages > kb_per_cta) stages = (int)kb_per_cta;
    if (stages < 1) return RSR_E_SHAPE;
    p.stages = stages;
    const int smem = stages * stage_bytes + fixed + 16 * stages;
    if (smem > h->max_smem) return RSR_E_SHAPE;
    int acc_stride = 32;
    while (acc_stride < bn) acc_stride <<= 1;
    p.acc_stride = acc_stride;
    p.tmem_cols = 2 * acc_stride;

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
    gemm_tcgen05_kernel<<<grid, GEMM_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
    RSR_LAUNCH_CHECK();
    return 0;
}
