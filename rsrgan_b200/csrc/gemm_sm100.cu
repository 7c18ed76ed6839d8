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
#include <stdlib.h>

#include "common.cuh"
#include "handle.h"

using namespace rsr;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 8;                                    // two warps per TMEM lane quadrant
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;               // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue

struct GemmKParams {
    int M, N, K;
    int bn, stages, a_mn, b_mn, bf;
    int m_tiles, n_tiles, splits, acc_stride, tmem_cols;
    int st_stride;            // bytes of epilogue staging per warp (fp32 box 4 KB and/or 16-bit box 4 KB)
    int w16;                  // columns per 16-bit store box: 64 (128-byte rows, SW128) or 32 (ragged N; 64-byte rows, SW64)
    int reduce32;             // out32 is accumulated with TMA reduce-add (split-K partials, or beta == 1)
    float alpha, beta;
    const float* bias;
    const float* resid; int ldr;
    int act;
    const uint16_t* dsrc; int ldd; int dact;
    float* out32; int ldc32;         // direct access: previous contents when beta is neither 0 nor 1, ragged N edge
    uint16_t* out16; int ldc16;
    int has32, has16;
    int fast16;               // 16-bit-only output in 64-column boxes, no residual / alpha: epilogue_fast16
    float* stats;             // STATS instances: per 128-row block (count, mean, M2) of every output column, [block][3][N]
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

// Epilogue of ONE 128 x bn accumulator tile, executed by one epilogue warp: thread <-> accumulator row
// (TMEM lane), 32 columns per tcgen05.ld; bias / residual / activation / activation-gradient in registers;
// results go through a swizzled per-warp staging box and leave with one TMA store (or reduce-add) per box.
// STATS (batch_norm statistics in the epilogue, models/dnn.py:56-62, models/discriminator_dnn.py:36-46): the fp32 chunk that
// is about to leave through the staging box is read back column-wise -- lane <-> column, 32 conflict-free shared loads --
// and its (mean, M2) over the live rows of this warp's 32-row quadrant go to sstat[quadrant][column]; the kernel merges the
// four quadrants of a tile in a fixed order (Chan) and writes one (count, mean, M2) partial per 128-row block and column,
// the layout rsr_bn_train_stats' finish kernel merges -- so the statistics kernel never re-reads the pre-activation.
template <bool STATS>
__device__ __forceinline__ void epilogue_tile(const GemmKParams& p, const CUtensorMap* tmC32, const CUtensorMap* tmC16,
                                              uint32_t taddr, int row0, int n0, int lane, int ew, uint32_t st32,
                                              uint32_t st16, bool leader, uint32_t sstat) {
    const int span = (p.has16 && p.w16 == 64) ? 64 : 32;   // the two warps of a quadrant interleave column spans
    const uint32_t sw32 = (uint32_t)(lane & 7), sw16 = (uint32_t)((lane >> 1) & 3);
    const bool general_beta = p.has32 && !p.reduce32 && p.beta != 0.0f;
    const int row = row0 + lane;
    const bool row_ok = row < p.M;
    for (int s0 = (ew >> 2) * span; s0 < p.bn; s0 += 2 * span) {
      bool wrote16 = false;
      for (int c0 = s0; c0 < s0 + span && c0 < p.bn; c0 += 32) {
        const int col0 = n0 + c0;
        if (col0 >= p.N) break;                        // warp-uniform
        // epilogue operands of this row (independent of the accumulator: issue first)
        float bj = 0.f;
        if (p.bias && col0 + lane < p.N) bj = __ldg(p.bias + col0 + lane);
        float4 rs[8];
        uint4 dq[4];
        if (p.resid) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                rs[c] = (row_ok && col0 + 4 * c < p.N)
                            ? __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ldr + col0) + c)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (p.dsrc) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                dq[c] = (row_ok && col0 + 8 * c < p.N)
                            ? __ldg(reinterpret_cast<const uint4*>(p.dsrc + (size_t)row * p.ldd + col0) + c)
                            : make_uint4(0u, 0u, 0u, 0u);
        }
        float v[32];
        __syncwarp();                                  // reconverge: tcgen05.ld is .sync.aligned
        if (c0 + 32 <= p.bn) {
            tmem_ld32(taddr + (uint32_t)c0, v);
        } else {                                       // bn is a multiple of 16 (only when n_tiles == 1)
            tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
            for (int j = 16; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], p.alpha, __shfl_sync(0xffffffffu, bj, j));
        if (p.resid) {
#pragma unroll
            for (int c = 0; c < 8; ++c) { v[4 * c] += rs[c].x; v[4 * c + 1] += rs[c].y; v[4 * c + 2] += rs[c].z; v[4 * c + 3] += rs[c].w; }
        }
        // activation / activation-gradient: the selector is hoisted out of the per-element loops
        if (p.act == RSR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (p.act == RSR_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.3f * v[j]);
        } else if (p.act == RSR_ACT_CLIP) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], -0.5f), 1.5f);
        }
        if (p.dsrc && p.dact != RSR_ACT_NONE) {
            // relu' / lrelu' from the sign of the stored 16-bit activation OUTPUT (fp16 and bf16 alike:
            // y > 0  <=>  sign bit clear and magnitude bits non-zero)
            const float neg = p.dact == RSR_ACT_LRELU ? 0.3f : 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t w[4] = {dq[c].x, dq[c].y, dq[c].z, dq[c].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool pos_lo = (int16_t)(w[k] & 0xFFFFu) > 0, pos_hi = (int32_t)w[k] >= 0x10000;
                    v[8 * c + 2 * k] *= pos_lo ? 1.0f : neg;
                    v[8 * c + 2 * k + 1] *= pos_hi ? 1.0f : neg;
                }
            }
        }
        if (col0 + 32 > p.N && (p.N & 7)) {
            // ragged right edge (N not a multiple of the 16-byte TMA store granule): plain stores
            if (row_ok && p.has16) {                   // out16 holds the value without the beta term
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.N) p.out16[(size_t)row * p.ldc16 + col0 + j] = f2h(v[j], p.bf);
            }
            if (row_ok && p.has32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (col0 + j < p.N) {
                        float* o = p.out32 + (size_t)row * p.ldc32 + col0 + j;
                        if (p.reduce32) red_add_f32(o, v[j]);
                        else *o = general_beta ? v[j] + p.beta * *o : v[j];
                    }
                }
            }
            continue;
        }
        // the previous boxes of this warp must have been read out of the staging buffers
        if (leader) tma_wait_group_read<0>();
        __syncwarp();
        if (p.has16) {
            const uint32_t cc0 = (uint32_t)(c0 - s0) >> 3;       // first 16-byte chunk of this half inside the box row
#pragma unroll
            uint32_t pk[16];
            if (p.bf) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                    pk[k] = *reinterpret_cast<const uint32_t*>(&t);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const __half2 t = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                    pk[k] = *reinterpret_cast<const uint32_t*>(&t);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t dst = p.w16 == 64 ? st16 + (uint32_t)lane * 128u + (((cc0 + (uint32_t)c) ^ sw32) << 4)
                                                 : st16 + (uint32_t)lane * 64u + (((uint32_t)c ^ sw16) << 4);
                st_shared_v4(dst, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
            wrote16 = true;
        }
        if (p.has32) {
            if (general_beta) {                        // out32 = v + beta * out32_old  (beta not in {0, 1})
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (row_ok && col0 + 4 * c < p.N) {
                        const float4 o = *(reinterpret_cast<const float4*>(p.out32 + (size_t)row * p.ldc32 + col0) + c);
                        v[4 * c] += p.beta * o.x; v[4 * c + 1] += p.beta * o.y; v[4 * c + 2] += p.beta * o.z; v[4 * c + 3] += p.beta * o.w;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
                st_shared_v4(st32 + (uint32_t)lane * 128u + (((uint32_t)c ^ sw32) << 4),
                             __float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]),
                             __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (leader && row0 < p.M) {
            if (p.has16 && p.w16 == 32) tma_store_2d(tmC16, st16, col0, row0);
            if (p.has32) {
                if (p.reduce32) tma_reduce_add_2d(tmC32, st32, col0, row0);
                else tma_store_2d(tmC32, st32, col0, row0);
            }
            tma_commit_group();
        }
        if constexpr (STATS) {
            int live = p.M - row0;
            live = live < 0 ? 0 : (live > 32 ? 32 : live);
            float sh = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x)
                             : "r"(st32 + (uint32_t)r * 128u + ((((uint32_t)(lane >> 2)) ^ (uint32_t)(r & 7)) << 4) + (uint32_t)(lane & 3) * 4u));
                if (r == 0) sh = x;
                if (r < live) { const float d = x - sh; a0 += d; a1 = fmaf(d, d, a1); }
            }
            float mean = 0.f, m2 = 0.f;
            if (live > 0) {
                const float inv = 1.0f / (float)live;
                mean = sh + a0 * inv;
                m2 = fmaxf(a1 - a0 * a0 * inv, 0.0f);
            }
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sstat + (uint32_t)(c0 + lane) * 8u), "f"(mean), "f"(m2) : "memory");
        }
      }
      if (wrote16 && p.w16 == 64 && leader && row0 < p.M) {   // one 64-column (128-byte rows) box per span
          tma_store_2d(tmC16, st16, n0 + s0, row0);
          tma_commit_group();
      }
    }
}

// Fast path of the epilogue for the shape most of the step's GEMMs have (fully_connected forward and data gradient:
// 16-bit output only, alpha = 1, no residual, 64-column store boxes): one warp drains a 32-row x 64-column span per
// iteration with BOTH tcgen05.ld in flight before a single wait, the bias as uniform 16-byte loads (no shuffles), one
// proxy fence and one TMA store per span -- half the fixed per-chunk cost of the general path below, which was the
// pacing stage of these GEMMs (ncu: epilogue warps busy 2/3 of the kernel, tensor pipe 41 %).
// (Measured and not kept: loading the activation-gradient source of a warp's first span before it waits for the
//  accumulator, and the next span's while the current one is drained -- 32 more live registers at the 168-register cap of
//  this 320-thread kernel: 180 instead of 44 bytes of spills in EVERY instance, cfg-2 3.101 against 3.057 ms per schedule.)
template <bool HAS_D>
__device__ __forceinline__ void epilogue_fast16(const GemmKParams& p, const CUtensorMap* tmC16, uint32_t taddr, int row0,
                                                int n0, int lane, int ew, uint32_t st16, bool leader) {
    const uint32_t sw = (uint32_t)(lane & 7);
    const int row = row0 + lane;
    const bool row_ok = row < p.M;
    for (int s0 = (ew >> 2) * 64; s0 < p.bn; s0 += 128) {
        const int col0 = n0 + s0;
        if (col0 >= p.N) break;                        // warp-uniform
        uint4 dq[8];
        if (HAS_D) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                dq[c] = (row_ok && col0 + 8 * c < p.N)
                            ? __ldg(reinterpret_cast<const uint4*>(p.dsrc + (size_t)row * p.ldd + col0) + c)
                            : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t r[64];
        __syncwarp();                                  // reconverge: tcgen05.ld is .sync.aligned
        tmem_ld32_nowait(taddr + (uint32_t)s0, r);
        tmem_ld32_nowait(taddr + (uint32_t)s0 + 32u, r + 32);
        tmem_ld_wait();
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {                                  // same address in every lane: one broadcast transaction each
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (col0 + 4 * c < p.N) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + c);
                    v[4 * c] += b.x; v[4 * c + 1] += b.y; v[4 * c + 2] += b.z; v[4 * c + 3] += b.w;
                }
            }
        }
        if (p.act == RSR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (p.act == RSR_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.3f * v[j]);
        } else if (p.act == RSR_ACT_CLIP) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = fminf(fmaxf(v[j], -0.5f), 1.5f);
        }
        if (HAS_D && p.dact != RSR_ACT_NONE) {
            const float neg = p.dact == RSR_ACT_LRELU ? 0.3f : 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t w[4] = {dq[c].x, dq[c].y, dq[c].z, dq[c].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool pos_lo = (int16_t)(w[k] & 0xFFFFu) > 0, pos_hi = (int32_t)w[k] >= 0x10000;
                    v[8 * c + 2 * k] *= pos_lo ? 1.0f : neg;
                    v[8 * c + 2 * k + 1] *= pos_hi ? 1.0f : neg;
                }
            }
        }
        uint32_t pk[32];
        if (p.bf) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                pk[k] = *reinterpret_cast<const uint32_t*>(&t);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const __half2 t = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                pk[k] = *reinterpret_cast<const uint32_t*>(&t);
            }
        }
        if (leader) tma_wait_group_read<0>();          // the previous box of this warp has left the staging buffer
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
            st_shared_v4(st16 + (uint32_t)lane * 128u + (((uint32_t)c ^ sw) << 4), pk[4 * c], pk[4 * c + 1],
                         pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (leader && row0 < p.M) {
            tma_store_2d(tmC16, st16, col0, row0);
            tma_commit_group();
        }
    }
}

// STATS: column `col` (< bn) of a finished tile -- the four quadrants' (mean, M2) merged in a fixed order
__device__ __forceinline__ void stats_merge_tile(const GemmKParams& p, uint32_t sstat_tile, int m0, int n0, int col) {
    if (col >= p.bn || n0 + col >= p.N) return;
    float na = 0.f, ma = 0.f, qa = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int live = p.M - (m0 + 32 * q);
        live = live < 0 ? 0 : (live > 32 ? 32 : live);
        if (live == 0) continue;
        float mb, qb;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(mb), "=f"(qb) : "r"(sstat_tile + (uint32_t)(q * p.bn + col) * 8u));
        const float nb = (float)live, nab = na + nb, d = mb - ma;
        ma += d * (nb / nab);
        qa += qb + d * d * (na * nb / nab);
        na = nab;
    }
    float* o = p.stats + (size_t)(m0 / BM) * 3 * p.N + n0 + col;
    o[0] = na; o[p.N] = ma; o[2 * (size_t)p.N] = qa;
}

// Persistent, warp-specialised GEMM.  Each CTA (one per SM) walks tiles tile = blockIdx.x + i*gridDim.x of
// the (k-split, m, n) tile space.  Three pipelines run concurrently: TMA -> smem ring (full/empty
// mbarriers), tcgen05.mma -> double-buffered TMEM accumulator (tfull/tempty), and the epilogue warps,
// which drain accumulator `a` while the MMA warp already fills accumulator `a^1` for the next tile.
// Epilogue: thread <-> accumulator row, 32 columns at a time (tcgen05.ld 32x32b.x32); epilogue operands
// (residual, activation-gradient source) are read as 16-byte vectors of that row; results go to a
// swizzled per-warp staging box and leave with ONE TMA store per box (cp.async.bulk.tensor, or
// cp.reduce...add for accumulated fp32 outputs), so HBM sees full 128-byte lines and the M / N edges are
// clipped by the tensor map.
template <bool STATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC32, const __grid_constant__ CUtensorMap tmC16,
                    const GemmKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B-swizzle atoms need 1024 B alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int a_stage_bytes = BM * BK * 2;                          // 16 KB
    const int b_boxes = p.b_mn ? (p.bn + 63) / 64 : 1;
    const int b_stage_bytes = p.b_mn ? b_boxes * BK * 128 : p.bn * 128;
    const int stage_bytes = a_stage_bytes + ((b_stage_bytes + 1023) & ~1023);
    const uint32_t epi_off = (uint32_t)(p.stages * stage_bytes);
    const uint32_t bars = smem_base + epi_off + (uint32_t)(EPI_WARPS * p.st_stride);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
    auto tfull_bar = [&](int a) { return bars + 16u * p.stages + 8u * a; };
    auto tempty_bar = [&](int a) { return bars + 16u * p.stages + 16u + 8u * a; };
    const uint32_t tmem_slot = bars + 16u * p.stages + 32u;
    const uint32_t sstat0 = (tmem_slot + 16u + 15u) & ~15u;           // STATS: float2 [2 accumulators][4 quadrants][bn]

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.has32) tma_prefetch_desc(&tmC32);
        if (p.has16) tma_prefetch_desc(&tmC16);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
        fence_mbar_init();
    }
    pdl_launch_dependents();                                         // the next kernel's prologue may overlap this one
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                                                      // predecessors' global writes are visible from here on
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int nkb = (p.K + BK - 1) / BK;
    const int tiles_mn = p.m_tiles * p.n_tiles;
    const int total = tiles_mn * p.splits;

    if (warp == 0) {
        if (elect_one_sync()) {
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
        if (elect_one_sync()) {
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
        // ---------------- epilogue warps ----------------
        const int ew = warp - 2;
        const int q = warp & 3;                                // TMEM lane quadrant this warp may access
        const uint32_t st32 = smem_base + epi_off + (uint32_t)(ew * p.st_stride);     // [32 rows][128 B], SW128
        const uint32_t st16 = st32 + (p.has32 ? 4096u : 0u);                           // [32 rows][128 B] SW128 | [32][64 B] SW64
        const bool leader = elect_one_sync();                  // issues (and waits for) this warp's TMA stores
        int acc = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int mn = tile % tiles_mn;
            const int m0 = (mn / p.n_tiles) * BM, n0 = (mn % p.n_tiles) * p.bn;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(acc * p.acc_stride) + ((uint32_t)(q * 32) << 16);
            const int row0 = m0 + q * 32;
            if (p.fast16) {
                if (p.dsrc) epilogue_fast16<true>(p, &tmC16, taddr, row0, n0, lane, ew, st16, leader);
                else epilogue_fast16<false>(p, &tmC16, taddr, row0, n0, lane, ew, st16, leader);
            } else {
                epilogue_tile<STATS>(p, &tmC32, &tmC16, taddr, row0, n0, lane, ew, st32, st16, leader,
                                     sstat0 + (uint32_t)((acc * 4 + q) * p.bn) * 8u);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));       // accumulator may be overwritten
            if constexpr (STATS) {      // all eight epilogue warps have written their quadrant / column spans of this tile
                named_bar_sync(1u, 32 * EPI_WARPS);
                stats_merge_tile(p, sstat0 + (uint32_t)(acc * 4 * p.bn) * 8u, m0, n0, ew * 32 + lane);
            }
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
        if (leader) tma_wait_group<0>();                    // stores complete before the CTA retires
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}


// Two-CTA variant (cluster of 2, tcgen05 cta_group::2): a pair of CTAs computes one 256 x bn tile.  Each CTA
// loads its own 128 rows of A and only HALF of the B tile (bn/2 columns); the leader issues UMMA 256 x bn x 16
// instructions that run on both SMs, each SM reading the other's B half through the pair's datapath.  Per CTA
// and k-block the operand traffic drops from 16 KB + 32 KB to 16 KB + 16 KB (bn = 256), which is what limits
// the one-CTA kernel on the large products (L2 -> SM operand bandwidth).  Barriers: TMA loads of both CTAs
// complete on the LEADER's full barrier; the leader's commits are multicast to both CTAs' empty / tfull
// barriers; both CTAs' epilogue warps arrive on the leader's tempty barrier.
template <bool STATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC32, const __grid_constant__ CUtensorMap tmC16,
                     const GemmKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                         // 0 = leader
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int hn = p.bn / 2;                                         // B columns held by each CTA

    const int a_stage_bytes = BM * BK * 2;                           // 16 KB
    const int b_boxes = p.b_mn ? hn / 64 : 1;
    const int b_stage_bytes = p.b_mn ? b_boxes * BK * 128 : hn * 128;
    const int stage_bytes = a_stage_bytes + ((b_stage_bytes + 1023) & ~1023);
    const uint32_t epi_off = (uint32_t)(p.stages * stage_bytes);
    const uint32_t bars = smem_base + epi_off + (uint32_t)(EPI_WARPS * p.st_stride);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
    auto tfull_bar = [&](int a) { return bars + 16u * p.stages + 8u * a; };
    auto tempty_bar = [&](int a) { return bars + 16u * p.stages + 16u + 8u * a; };
    const uint32_t tmem_slot = bars + 16u * p.stages + 32u;
    const uint32_t sstat0 = (tmem_slot + 16u + 15u) & ~15u;           // STATS: float2 [2 accumulators][4 quadrants][bn]
    const uint32_t to_leader = mapa_u32(smem_base, 0u) - smem_base;  // shared::cta -> shared::cluster address in the leader

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.has32) tma_prefetch_desc(&tmC32);
        if (p.has16) tma_prefetch_desc(&tmC16);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPI_WARPS); }
        fence_mbar_init();
    }
    pdl_launch_dependents();
    if (warp == 1) tmem_alloc_2cta(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                              // both CTAs' barriers exist before any remote arrive
    tc_fence_after();
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int nkb = (p.K + BK - 1) / BK;
    const int tiles_mn = p.m_tiles * p.n_tiles;                      // m_tiles counts 256-row tiles here
    const int total = tiles_mn * p.splits;

    if (warp == 0) {
        if (elect_one_sync()) {
            // ---------------- TMA producer (both CTAs) ----------------
            const uint32_t tx2 = 2u * (uint32_t)(a_stage_bytes + b_stage_bytes);    // bytes of BOTH CTAs land on the leader's barrier
            int s = 0; uint32_t ph = 0;
            for (int tile = pair; tile < total; tile += npairs) {
                const int ks = tile / tiles_mn, mn = tile % tiles_mn;
                const int m0 = (mn / p.n_tiles) * 2 * BM + (int)rank * BM, n0 = (mn % p.n_tiles) * p.bn + (int)rank * hn;
                const int kb0 = (int)((long long)ks * nkb / p.splits), kb1 = (int)((long long)(ks + 1) * nkb / p.splits);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = smem_base + s * stage_bytes;
                    const uint32_t sb = sa + a_stage_bytes;
                    const uint32_t lead_full = full_bar(s) + to_leader;
                    if (rank == 0) mbar_expect_tx(full_bar(s), tx2);
                    const int k0 = kb * BK;
                    if (!p.a_mn) {
                        tma_load_2d_2cta(sa, &tmA, lead_full, k0, m0);
                    } else {
                        tma_load_2d_2cta(sa, &tmA, lead_full, m0, k0);
                        tma_load_2d_2cta(sa + BK * 128, &tmA, lead_full, m0 + 64, k0);
                    }
                    if (!p.b_mn) {
                        tma_load_2d_2cta(sb, &tmB, lead_full, k0, n0);
                    } else {
                        for (int j = 0; j < b_boxes; ++j)
                            tma_load_2d_2cta(sb + j * BK * 128, &tmB, lead_full, n0 + 64 * j, k0);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (rank == 0 && elect_one_sync()) {
            // ---------------- MMA issuer (leader CTA only) ----------------
            const uint32_t idesc = umma_idesc(2 * BM, p.bn, p.bf, p.a_mn, p.b_mn);
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int tile = pair; tile < total; tile += npairs) {
                const int ks = tile / tiles_mn;
                const int kb0 = (int)((long long)ks * nkb / p.splits), kb1 = (int)((long long)(ks + 1) * nkb / p.splits);
                mbar_wait(tempty_bar(acc), aph ^ 1u);          // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * p.acc_stride);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(s), ph);                // both CTAs' stage s has landed
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * stage_bytes;
                    const uint32_t sb = sa + a_stage_bytes;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = p.a_mn ? umma_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                                   : umma_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = p.b_mn ? umma_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                                   : umma_desc_sw128(sb + k * 32, 16, 1024);
                        tc_mma_f16_2cta(tacc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit_2cta_mc(empty_bar(s), (uint16_t)3);   // frees stage s in both CTAs
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                tc_commit_2cta_mc(tfull_bar(acc), (uint16_t)3);     // accumulator complete in both CTAs
                if (++acc == 2) { acc = 0; aph ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue warps (both CTAs, own 128 rows) ----------------
        const int ew = warp - 2;
        const int q = warp & 3;
        const uint32_t st32 = smem_base + epi_off + (uint32_t)(ew * p.st_stride);
        const uint32_t st16 = st32 + (p.has32 ? 4096u : 0u);
        const bool leader = elect_one_sync();
        int acc = 0; uint32_t aph = 0;
        for (int tile = pair; tile < total; tile += npairs) {
            const int mn = tile % tiles_mn;
            const int m0 = (mn / p.n_tiles) * 2 * BM + (int)rank * BM, n0 = (mn % p.n_tiles) * p.bn;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(acc * p.acc_stride) + ((uint32_t)(q * 32) << 16);
            if (p.fast16) {
                if (p.dsrc) epilogue_fast16<true>(p, &tmC16, taddr, m0 + q * 32, n0, lane, ew, st16, leader);
                else epilogue_fast16<false>(p, &tmC16, taddr, m0 + q * 32, n0, lane, ew, st16, leader);
            } else {
                epilogue_tile<STATS>(p, &tmC32, &tmC16, taddr, m0 + q * 32, n0, lane, ew, st32, st16, leader,
                                     sstat0 + (uint32_t)((acc * 4 + q) * p.bn) * 8u);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(acc) + to_leader);
            if constexpr (STATS) {
                named_bar_sync(1u, 32 * EPI_WARPS);
                stats_merge_tile(p, sstat0 + (uint32_t)(acc * 4 * p.bn) * 8u, m0, n0, ew * 32 + lane);
            }
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
        if (leader) tma_wait_group<0>();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                              // the peer may still read this CTA's operands / signal its barriers
    if (warp == 1) tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols);
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
    if (cudaMalloc(&h->partials, RSR_PARTIAL_WORDS * sizeof(float)) != cudaSuccess) { cudaFree(h->flags); delete h; return RSR_E_NODEV; }
    *out = h;
    return 0;
}

extern "C" int rsr_destroy(rsr_handle* h) {
    if (!h) return RSR_E_ARG;
    if (h->flags) cudaFree(h->flags);
    if (h->partials) cudaFree(h->partials);
    delete h;
    return 0;
}

extern "C" int rsr_num_sms(rsr_handle* h) { return h ? h->num_sms : RSR_E_ARG; }

int rsr_get_tmap_ex(rsr_handle* h, const void* ptr, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t ld,
                    uint32_t b0, uint32_t b1, int swizzle, CUtensorMap* out) {
    if ((elem_bytes != 2 && elem_bytes != 4) || (swizzle != 64 && swizzle != 128)) return RSR_E_ARG;
    if (((uintptr_t)ptr & 15) || (ld * elem_bytes) % 16 || (int)(b0 * elem_bytes) != swizzle || b1 > 256 || d0 == 0 || d1 == 0)
        return RSR_E_ARG;
    TmapKey key{ptr, d0, d1, ld * elem_bytes, b0, b1, (uint32_t)elem_bytes | ((uint32_t)swizzle << 8)};
    {
        std::lock_guard<std::mutex> g(h->mu);
        auto it = h->tmaps.find(key);
        if (it != h->tmaps.end()) { *out = it->second; return 0; }
    }
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * elem_bytes};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = h->encode(&m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                           const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
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

int rsr_get_tmap(rsr_handle* h, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld,
                 uint32_t b0, uint32_t b1, CUtensorMap* out) {
    return rsr_get_tmap_ex(h, ptr, 2, d0, d1, ld, b0, b1, 128, out);
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
    // Two-CTA (cta_group::2) variant for the large products: 256 x bn tiles, half the B traffic per CTA.
    int two = 0, bn2 = (a->N % 256 == 0) ? 256 : ((a->N % 128 == 0) ? 128 : 0);
    // (measured on cfg-2 shapes: pays off for K >= 512 and >= 1e10 flop; below that the cluster launch and the two
    //  cluster barriers cost more than the saved operand traffic)
    if (a->tile_n <= 0 && bn2 && a->M >= 512 && a->K >= 512 && 2.0 * a->M * a->N * a->K >= 1.0e10 && !getenv("RSR_NO_2CTA")) {
        if (h->gemm2_pairs < 0) {   // co-resident CTA pairs, queried once
            cudaFuncSetAttribute(gemm2_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2, 1, 1); cfg.blockDim = dim3(GEMM_THREADS, 1, 1); cfg.dynamicSmemBytes = h->max_smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, gemm2_tcgen05_kernel<false>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
            h->gemm2_pairs = n;
        }
        const long long pair_tiles = (long long)((a->M + 2 * BM - 1) / (2 * BM)) * (a->N / bn2);
        // worth it when the pairs are kept busy (split-K multiplies the tile count further down)
        if (h->gemm2_pairs > 0 && (pair_tiles >= h->gemm2_pairs / 2 || plain_accumulate)) two = 1;
    }
    const int units = two ? h->gemm2_pairs : h->num_sms;          // concurrently running tile workers
    const int m_tiles = two ? (a->M + 2 * BM - 1) / (2 * BM) : (a->M + BM - 1) / BM;
    int bn = two ? bn2 : a->tile_n;
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
    p.has32 = a->out32 ? 1 : 0; p.has16 = a->out16 ? 1 : 0;
    p.stats = a->stats;
    // statistics in the epilogue: plain fp32 pre-activation output, whole 32-column chunks, at most 256 row blocks (the
    // partial buffer of rsr_bn_train_stats); anything else -> RSR_E_SHAPE, the caller runs the statistics kernel instead
    if (a->stats && (!a->out32 || a->out16 || a->bias || a->resid || a->act != RSR_ACT_NONE || a->dact_src || a->beta != 0.0f ||
                     a->alpha != 1.0f || (a->N & 31) || (a->M + BM - 1) / BM > 256 || a->split_k > 1))
        return RSR_E_SHAPE;
    p.st_stride = (p.has32 ? 4096 : 0) + (p.has16 ? 4096 : 0);
    p.m_tiles = m_tiles;
    p.n_tiles = (a->N + bn - 1) / bn;

    const int a_stage = BM * BK * 2;
    const int bcols = two ? bn / 2 : bn;                           // B columns each CTA holds
    const int b_boxes = p.b_mn ? (bcols + 63) / 64 : 1;
    const int b_stage = p.b_mn ? b_boxes * BK * 128 : bcols * 128;
    const int stage_bytes = a_stage + ((b_stage + 1023) & ~1023);
    const int nkb = (a->K + BK - 1) / BK;
    // split-K: only for "out32 += A B" (weight gradients: few output tiles, K = all frames); partial
    // sums are added into out32 with red.global.add.f32, which is the beta = 1 semantics
    int splits = a->split_k;
    const int tiles_mn = p.m_tiles * p.n_tiles;
    if (splits <= 0) {
        splits = 1;
        if (plain_accumulate && tiles_mn < units && nkb >= 8) {
            // cost of a candidate = rounds of the persistent grid x (k-blocks per tile + a per-tile epilogue / pipeline-fill
            // charge of ~6 k-blocks); more splits also mean more reduce-add traffic, so ties go to the smaller count
            long long best = -1;
            const int max_splits = nkb / 4 < 64 ? nkb / 4 : 64;
            for (int s = 1; s <= max_splits; ++s) {
                const long long rounds = ((long long)tiles_mn * s + units - 1) / units;
                const long long cost = rounds * ((nkb + s - 1) / s + 6);
                if (best < 0 || cost < best) { best = cost; splits = s; }
            }
        }
    }
    if (splits > 1 && !plain_accumulate) return RSR_E_ARG;
    if (splits > nkb) splits = nkb;
    p.splits = splits;
    // accumulated fp32 outputs (out32 += ..., i.e. beta == 1, and all split-K partials) leave through TMA reduce-add
    p.reduce32 = (a->out32 && (splits > 1 || a->beta == 1.0f)) ? 1 : 0;
    if (p.n_tiles > 1 && (bn & 31)) return RSR_E_SHAPE;   // a 32-column store box must not reach into the next tile
    // 16-bit outputs leave in 64-column boxes (128-byte rows: 64-byte rows reach only a fraction of the TMA store
    // rate) unless N is ragged against the 16-byte store granule or a box would reach into the next tile
    p.w16 = (a->out16 && (a->N & 7) == 0 && (p.n_tiles == 1 || (bn & 63) == 0)) ? 64 : 32;
    p.fast16 = (p.w16 == 64 && !a->out32 && !a->resid && a->alpha == 1.0f && (bn & 63) == 0 &&
                (!a->bias || ((uintptr_t)a->bias & 15) == 0) && !getenv("RSR_NO_FAST_EPI")) ? 1 : 0;
    const int fixed = 1024 /*align slack*/ + EPI_WARPS * p.st_stride + 64 + 16 * 8 + (a->stats ? 2 * 4 * 256 * 8 + 32 : 0);
    int stages = (h->max_smem - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    const int kb_per_tile = (nkb + splits - 1) / splits;
    const long long total_tiles = (long long)tiles_mn * splits;
    const int grid = (int)(total_tiles < units ? total_tiles : units);   // CTAs, or CTA pairs
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
    if (!p.b_mn) rc = rsr_get_tmap(h, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, 64, (uint32_t)bcols, &tmB);
    else         rc = rsr_get_tmap(h, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK, &tmB);
    if (rc) return rc;
    CUtensorMap tmC32 = tmA, tmC16 = tmA;   // placeholders when an output is absent (never dereferenced)
    if (a->out32) {
        rc = rsr_get_tmap_ex(h, a->out32, 4, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc32, 32, 32, 128, &tmC32);
        if (rc) return rc;
    }
    if (a->out16) {
        rc = rsr_get_tmap_ex(h, a->out16, 2, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc16, (uint32_t)p.w16, 32,
                             p.w16 == 64 ? 128 : 64, &tmC16);
        if (rc) return rc;
    }

    static bool attr_set = false;
    if (!attr_set) {
        RSR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        RSR_CHECK_CUDA(cudaFuncSetAttribute(gemm2_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        RSR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        RSR_CHECK_CUDA(cudaFuncSetAttribute(gemm2_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_set = true;
    }
    // One CTA per SM, and never next to a CTA of a recurrence kernel (those hold the whole TMEM of their SM for
    // hundreds of microseconds; a GEMM CTA sharing the SM would sit in tcgen05.alloc until they retire).
    const int smem_launch = smem < RSR_EXCLUSIVE_SMEM_GEMM ? RSR_EXCLUSIVE_SMEM_GEMM : smem;
    // programmatic dependent launch (opt-in, RSR_PDL=1): this kernel may begin (prologue only) before its predecessor in
    // the stream ends.  Measured on cfg-2 inside the CUDA graph: no gain (4.12 ms with, 4.08 ms without) -- the graph's
    // kernel-to-kernel gaps are already short and the exclusive-SM shared-memory floors leave little to overlap.
    static const bool pdl = getenv("RSR_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_launch; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (two) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    if (two) {
        cfg.gridDim = dim3(2 * grid, 1, 1);
        if (p.stats) RSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_tcgen05_kernel<true>, tmA, tmB, tmC32, tmC16, p));
        else RSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_tcgen05_kernel<false>, tmA, tmB, tmC32, tmC16, p));
        return 0;
    }
    cfg.gridDim = dim3(grid, 1, 1);
    if (p.stats) RSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<true>, tmA, tmB, tmC32, tmC16, p));
    else RSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<false>, tmA, tmB, tmC32, tmC16, p));
    return 0;
}
