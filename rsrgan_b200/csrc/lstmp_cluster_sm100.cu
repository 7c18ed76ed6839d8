// LSTMP recurrence on thread-block clusters (sm_100a): the fused LSTM-gate kernels for Cp <= 512.
//
// Same math and the same C-ABI contract as lstmp_sm100.cu (see there for the reference citations:
// models/lstm.py:89-112, models/BNLSTMCell.py:176-213), but the CTAs that share an utterance slice
// form ONE CLUSTER (Cp/32 = 8 or 16 CTAs) and exchange the recurrent state through distributed shared
// memory instead of L2:
//
//   forward   z_t = Zx_t + mt_{t-1} Wc.  CTA j keeps the 128 packed gate rows of its 32 cells of Wc^T in
//             shared memory for all T steps (TMA, SW128), runs one swap-AB tcgen05 tile per step
//             (M = 128 gate rows, N = NB utterances, K = Cp), applies the sigmoid/tanh/peephole/cell
//             update in registers and pushes its 32 x NB slice of mt_t straight from registers into the
//             B-operand buffer of EVERY CTA of the cluster with st.async (16-byte remote stores that
//             complete_tx on the receiver's mbarrier: data and signal travel together).  The B operand
//             uses the un-swizzled canonical UMMA layout [k/8][row][8] so a CTA's slice is contiguous.
//   backward  dmt_{t-1} = dOut_{t-1} Wp^T + dz_t Wc^T, split along K: CTA j owns the 4 gates of its 32
//             cells (it computes that part of dz_t locally, no exchange of the B operand), multiplies by
//             its K-slice of Wc (all Cp output rows, Cp/128 tcgen05 tiles) and reduce-scatters the
//             partial rows to their owner CTAs, 16-bit, again with st.async + mbarrier.
//
// A time step therefore costs one DSMEM hop (~0.2 us) instead of a release/acquire round trip through
// L2 plus a TMA load (~4 us measured with the L2 variant), and groups are independent clusters: no
// co-residency requirement between groups, any batch size runs (in waves if need be).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "handle.h"
#include "cluster_util.h"

using namespace rsr;

// Optional in-kernel phase timing (build with -DRSR_TRACE; read back with rsr_debug_trace): thread 0 of
// CTA 0 records %clock64 at fixed points of each time step.
#ifdef RSR_TRACE
__device__ unsigned long long g_rsr_trace[8192];
#define TRACE(slot) do { if (blockIdx.x == 0 && tid == 0 && t < 64) g_rsr_trace[t * 8 + (slot)] = clock64(); } while (0)
#define TRACE_T(tt, slot) do { if (blockIdx.x == 0 && tid == 0 && (tt) < 64) g_rsr_trace[(tt) * 8 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#define TRACE_T(tt, slot) do { } while (0)
#endif

namespace {

constexpr int NB = 16;          // utterances per half (UMMA N)

// =========================================================================================
// forward
// =========================================================================================
struct CFwdParams {
    int B, T, Cp, bf;
    float forget_bias;
    const float* zx;            // [T*B, 4Cp] packed gate columns
    const uint16_t* wcT;        // [4Cp, Cp] packed gate rows
    const float* w_i; const float* w_f; const float* w_o;   // [Cp]
    const int* lengths;         // [B]
    uint16_t* mt_seq;           // [(T+1)*B, Cp]
    float* save;                // [T*B, 5, Cp] or null
    // fused input half (FUSEX): z_t = x_t K_x + b + mt_{t-1} Wc entirely in this kernel
    const uint16_t* kxT;        // [4Cp, Ik] packed gate rows of K_x^T, Ik = input width padded to 16, zero padded
    const float* bias;          // [4Cp] packed
    int Ik;
    int xmode;                  // exchange of mt_t: 0 = st.async all-gather (DSMEM), 1 = via L2 + TMA multicast (tmM)
};

// One cluster = one utterance group of NHALF x 16 utterances; CTA j owns cells [32j, 32j+32).  The CTA's
// 128 gate rows of Wc^T live in TMEM (A operand) for the whole sequence.  Each HALF is an independent
// recurrence over 16 utterances run by its own 4 warps (own accumulator, buffers, barriers); two halves
// share the resident weights and overlap each other's DSMEM exchange with MMA + gate math.
// FUSEX: the input half x_t K_x is computed here too (K_x^T slice resident in TMEM next to Wc^T, x_t tiles
// prefetched by TMA one step ahead, its MMAs issued BEFORE the wait for mt_{t-1} so they hide behind the
// exchange); otherwise Zx = X K_x + b arrives precomputed (fp32, from rsr_gemm).
template <int NHALF, bool FUSEX>
__global__ void __launch_bounds__(128 * NHALF, 1)
lstmp_fwd_cluster_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmM,
                         const CFwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, hh = warp >> 2;     // TMEM lane quadrant, half
    const int htid = tid & 127;
    const int G = p.Cp / 32;                    // CTAs per utterance group == cluster size
    const int HG = G / 2;
    const uint32_t j = cluster_ctarank();       // cell block: cells [32j, 32j+32)
    const int grp = blockIdx.x / G;
    const int b0 = grp * (NB * NHALF) + hh * NB;

    constexpr int XP = NB + 1;                                       // xchg pitch in floats
    const uint32_t sB_bytes = (uint32_t)p.Cp * NB * 2u;              // one B-operand buffer [Cp/8][NB][8] 16-bit
    const int KBX = FUSEX ? (p.Ik + 63) / 64 : 0;                   // 64-wide k sub-tiles of the x_t tile
    const uint32_t xt_bytes = (uint32_t)KBX * NB * 128u;             // one x_t tile: KBX x [NB x 64] 16-bit, SW128
    const uint32_t half_bytes = 2u * xt_bytes + 2u * sB_bytes + 2u * 128u * XP * 4u;   // x tiles (1024-aligned) | B buffers | xchg
    const uint32_t sXt0 = base + (uint32_t)hh * half_bytes;
    const uint32_t sB0 = sXt0 + 2u * xt_bytes;
    const uint32_t sX = sB0 + 2u * sB_bytes;                         // float xchg[2][128][XP]
    const uint32_t sBar = base + (uint32_t)NHALF * half_bytes + (uint32_t)hh * 64u;
    const uint32_t barM = sBar, full0 = sBar + 8, full1 = sBar + 16, xfull0 = sBar + 24, xfull1 = sBar + 32;
    const uint32_t tslot = base + (uint32_t)NHALF * half_bytes + (uint32_t)NHALF * 64u;
    float* xchg = reinterpret_cast<float*>(base_ptr + (sX - base));

    // TMEM: columns [0, Cp/2) = this CTA's 128 gate rows of Wc^T (A operand, 2 x 16 bit per column),
    //       columns [Cp/2 + 16 hh, +16) = accumulator of half hh
    //       (FUSEX: K_x^T slice in columns [Cp/2, Cp/2 + Ik/2) in between)
    const uint32_t ax_cols = FUSEX ? (uint32_t)p.Ik / 2u : 0u;
    const uint32_t a_cols = (uint32_t)p.Cp / 2u + ax_cols;
    uint32_t tcols = 32;
    while (tcols < a_cols + NB * NHALF) tcols <<= 1;
    if (htid == 0) {
        if (FUSEX) tma_prefetch_desc(&tmX);
        if (p.xmode) tma_prefetch_desc(&tmM);
        mbar_init(barM, 1); mbar_init(full0, 1); mbar_init(full1, 1); mbar_init(xfull0, 1); mbar_init(xfull1, 1);
        fence_mbar_init();
        mbar_expect_tx(full1, sB_bytes);        // armed for step 1 (mt_0 of every CTA of the cluster)
        mbar_expect_tx(full0, sB_bytes);        // armed for step 2
    }
    if (warp == 0) tmem_alloc(tslot, tcols);
    for (uint32_t i = htid; i < sB_bytes / 16u; i += 128u) st_shared_v4(sB0 + i * 16u, 0u, 0u, 0u, 0u);   // m_{-1} = 0
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    const uint32_t tmem_acc = tmem + a_cols + (uint32_t)(hh * NB);

    {   // weight slab -> TMEM, resident for the whole sequence: thread <-> gate row, 64 k (32 columns) per store
        const uint16_t* wrow = p.wcT + (size_t)(128 * j + 32 * q + lane) * p.Cp;
        for (int cb = hh; cb < p.Cp / 64; cb += NHALF) {
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + cb * 64) + c);
                r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
            }
            tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cb * 32u, r);
        }
        if (FUSEX) {   // K_x^T slice, 16 k (8 columns) per store; rows are zero padded to Ik
            const uint16_t* xrow = p.kxT + (size_t)(128 * j + 32 * q + lane) * p.Ik;
            for (int cb = hh; cb < p.Ik / 16; cb += NHALF) {
                const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(xrow + cb * 16));
                const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(xrow + cb * 16) + 1);
                tmem_st8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)p.Cp / 2u + (uint32_t)cb * 8u,
                         v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // every CTA's barriers are initialised and armed before a remote store can arrive

    // gate-math ownership: thread <-> (cells 4a..4a+3 of the block, utterance n = 4 q + nsub of this half)
    const int a = lane & 7, nsub = lane >> 3;
    const int n_own = 4 * q + nsub;
    const int cell0 = 32 * (int)j + 4 * a;
    const float4 wi4 = *reinterpret_cast<const float4*>(p.w_i + cell0);
    const float4 wf4 = *reinterpret_cast<const float4*>(p.w_f + cell0);
    const float4 wo4 = *reinterpret_cast<const float4*>(p.w_o + cell0);
    const float wi[4] = {wi4.x, wi4.y, wi4.z, wi4.w}, wf[4] = {wf4.x, wf4.y, wf4.z, wf4.w},
                wo[4] = {wo4.x, wo4.y, wo4.z, wo4.w};
    float creg[4] = {0.f, 0.f, 0.f, 0.f};
    const int b_own = b0 + n_own;
    const int len = b_own < p.B ? p.lengths[b_own] : 0;
    // destinations of this thread's remote stores: even quads serve ranks [0, G/2), odd quads [G/2, G)
    uint32_t rdelta[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rdelta[i] = i < HG ? mapa_u32(base, (uint32_t)((a & 1) * HG + i)) - base : 0u;

    const uint32_t idesc = umma_idesc(128, NB, p.bf, 0, 0);
    const size_t zx_ld = (size_t)4 * p.Cp;
    const int grow = 32 * q + lane;             // gate row of this thread in the tile (TMEM lane)
    const float* zx_col = FUSEX ? nullptr : p.zx + 128 * j + grow;
    const float bias_r = FUSEX ? __ldg(p.bias + 128 * j + grow) : 0.f;
    constexpr int NZ = FUSEX ? 1 : NB;
    float zx_cur[NZ], zx_nxt[NZ];
#pragma unroll
    for (int n = 0; n < NZ; ++n) {
        const int b = b0 + n;
        zx_cur[n] = (!FUSEX && b < p.B) ? __ldg(zx_col + (size_t)b * zx_ld) : 0.f;
        zx_nxt[n] = 0.f;
    }
    const uint64_t db_base = umma_desc_nosw(sB0, NB * 16u, 128u);
    const int KK = p.Cp / 16;
    const int KKX = FUSEX ? p.Ik / 16 : 0;
    const uint32_t bar_id = 1u + (uint32_t)hh;
    if (FUSEX && q == 0 && elect_one_sync()) {   // x_0 tile
        mbar_expect_tx(xfull0, xt_bytes);
        for (int kb = 0; kb < KBX; ++kb) tma_load_2d(sXt0 + (uint32_t)kb * (NB * 128u), &tmX, xfull0, kb * 64, b0);
    }

    for (int t = 0; t < p.T; ++t) {
        const int buf = t & 1;
        TRACE(0);
        if (!FUSEX && t + 1 < p.T) {   // Zx of the next step does not depend on the recurrence: in flight during this step
#pragma unroll
            for (int n = 0; n < NZ; ++n) {
                const int b = b0 + n;
                zx_nxt[n] = b < p.B ? __ldg(zx_col + ((size_t)(t + 1) * p.B + b) * zx_ld) : 0.f;
            }
        }
        if (q == 0) {   // first warp of the half, warp-uniform: the elected lane issues, operands stay in uniform registers
            const uint32_t fb = buf ? full1 : full0;
            if (FUSEX) {
                // input half first: it does not depend on mt_{t-1}, so it runs while the exchange is still in flight
                const uint32_t xb = buf ? xfull1 : xfull0;
                if (elect_one_sync()) {
                    if (t + 1 < p.T) {   // prefetch x_{t+1}; its buffer was last read by the MMAs of step t-1 (retired: barM)
                        const uint32_t xn = buf ? xfull0 : xfull1;
                        const uint32_t dstx = sXt0 + (uint32_t)(buf ^ 1) * xt_bytes;
                        mbar_expect_tx(xn, xt_bytes);
                        for (int kb = 0; kb < KBX; ++kb)
                            tma_load_2d(dstx + (uint32_t)kb * (NB * 128u), &tmX, xn, kb * 64, (t + 1) * p.B + b0);
                    }
                    mbar_wait(xb, (uint32_t)((t >> 1) & 1));
                    tc_fence_after();
                    const uint32_t sxt = sXt0 + (uint32_t)buf * xt_bytes;
                    for (int kk = 0; kk < KKX; ++kk) {
                        const uint64_t dx = umma_desc_sw128(sxt + (uint32_t)(kk >> 2) * (NB * 128u) + (uint32_t)(kk & 3) * 32u, 16, 1024);
                        tc_mma_f16_ts(tmem_acc, tmem + (uint32_t)p.Cp / 2u + (uint32_t)kk * 8u, dx, idesc, kk ? 1u : 0u);
                    }
                }
                __syncwarp();
            }
            if (t > 0) mbar_wait(fb, (uint32_t)(((t - 1) >> 1) & 1));     // all G slices of mt_{t-1} have landed
            TRACE(1);
            fence_proxy_async_smem();
            tc_fence_after();
            if (elect_one_sync()) {
                uint64_t db = db_base + (uint64_t)(((uint32_t)buf * sB_bytes) >> 4);
                uint32_t ta = tmem;
#pragma unroll 8
                for (int kk = 0; kk < KK; ++kk) {
                    tc_mma_f16_ts(tmem_acc, ta, db, idesc, (FUSEX || kk) ? 1u : 0u);
                    ta += 8u;                               // 16 k = 8 columns
                    db += (uint64_t)((2u * NB * 16u) >> 4); // two k-chunks of [NB rows][16 B]
                }
                tc_commit(barM);
                if (t > 0 && t + 2 < p.T) mbar_expect_tx(fb, sB_bytes);   // re-arm this buffer for step t + 2
            }
            TRACE(2);
        }
        __syncwarp();
        mbar_wait(barM, (uint32_t)(t & 1));
        tc_fence_after();
        TRACE(3);
        float acc[NB];
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16), acc);
        float* xc = xchg + buf * 128 * XP;          // quadrant q holds gate q (i, j, f, o) of cells 32j + lane
#pragma unroll
        for (int n = 0; n < NB; ++n) xc[grow * XP + n] = acc[n] + (FUSEX ? bias_r : zx_cur[FUSEX ? 0 : n]);
        tc_fence_before();
        named_bar_sync(bar_id, 128);
        TRACE(4);
        {
            const int n = n_own;
            const bool active = t < len;
            float mtv[4], sv[5][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cl = 4 * a + c;
                const float zi = xc[(0 * 32 + cl) * XP + n];
                const float zj = xc[(1 * 32 + cl) * XP + n];
                const float zf = xc[(2 * 32 + cl) * XP + n];
                const float zo = xc[(3 * 32 + cl) * XP + n];
                const float cp = creg[c];
                const float ig = sigmoidf_(zi + wi[c] * cp);
                const float fg = sigmoidf_(zf + p.forget_bias + wf[c] * cp);
                const float jg = tanhf_(zj);
                const float cn = fg * cp + ig * jg;
                const float og = sigmoidf_(zo + wo[c] * cn);
                const float mt = og * tanhf_(cn);
                sv[0][c] = ig; sv[1][c] = fg; sv[2][c] = og; sv[3][c] = jg; sv[4][c] = cn;
                mtv[c] = active ? mt : 0.f;
                if (active) creg[c] = cn;
            }
            const uint32_t lo = pack2(mtv[0], mtv[1], p.bf), hi = pack2(mtv[2], mtv[3], p.bf);
            TRACE(5);
            const size_t row = (size_t)t * p.B + b_own;
            if (p.xmode == 0) {
                if (t + 1 < p.T) {
                    // pair (quad 2k, quad 2k+1) -> one 16-byte k-chunk (8 cells) of row n
                    const uint32_t plo = __shfl_xor_sync(0xffffffffu, lo, 1), phi = __shfl_xor_sync(0xffffffffu, hi, 1);
                    const bool odd = a & 1;
                    const uint32_t w0 = odd ? plo : lo, w1 = odd ? phi : hi, w2 = odd ? lo : plo, w3 = odd ? hi : phi;
                    const uint32_t off = (uint32_t)(4 * j + (a >> 1)) * (NB * 16u) + (uint32_t)n * 16u;
                    const uint32_t dst = sB0 + (uint32_t)(buf ^ 1) * sB_bytes + off;
                    const uint32_t dbar = buf ? full0 : full1;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < HG) st_async_v4(dst + rdelta[k], w0, w1, w2, w3, dbar + rdelta[k]);
                }
                TRACE(6);
                // off the critical path: operand of the hoisted projection GEMM and of the backward pass
                if (b_own < p.B) *reinterpret_cast<uint2*>(p.mt_seq + (row + p.B) * p.Cp + cell0) = make_uint2(lo, hi);
            } else {
                // The global copy of mt_t (needed anyway) is the exchange medium: each CTA stores its 32-cell slice
                // (1 KB per half) and one elected thread multicasts that slice from L2 into the B-operand buffer of
                // every CTA of the cluster with a single 3-D TMA load -- 1 KB leaves the SM instead of 16 KB of
                // st.async traffic (same step time in practice, see fwd_xmode).
                if (b_own < p.B) *reinterpret_cast<uint2*>(p.mt_seq + (row + p.B) * p.Cp + cell0) = make_uint2(lo, hi);
                if (t + 1 < p.T) {
                    fence_proxy_async_all();                 // generic-proxy stores -> visible to the async proxy (TMA)
                    named_bar_sync(bar_id, 128);
                    if (q == 0 && elect_one_sync()) {
                        const uint32_t dst = sB0 + (uint32_t)(buf ^ 1) * sB_bytes + (uint32_t)(4 * j) * (NB * 16u);
                        const uint32_t dbar = buf ? full0 : full1;
                        tma_load_3d_multicast(dst, &tmM, dbar, 0, (t + 1) * p.B + b0, 4 * (int)j,
                                              (uint16_t)((1u << G) - 1u));
                    }
                }
                TRACE(6);
            }
            if (b_own < p.B && p.save) {
                float* s = p.save + row * 5 * p.Cp + cell0;
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    *reinterpret_cast<float4*>(s + (size_t)k * p.Cp) = make_float4(sv[k][0], sv[k][1], sv[k][2], sv[k][3]);
            }
        }
        TRACE(7);
#pragma unroll
        for (int n = 0; n < NZ; ++n) zx_cur[n] = zx_nxt[n];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tcols);
    cluster_sync_all();   // nobody leaves while a peer may still address its shared memory
}

// =========================================================================================
// backward
// =========================================================================================
struct CBwdParams {
    int B, T, Cp, bf;
    const float* dmt;           // [T*B, Cp] dOut W_p^T (read only)
    const uint16_t* wc;         // [Cp, 4Cp] packed gate columns
    const float* w_i; const float* w_f; const float* w_o;
    const int* lengths;
    const float* save;          // [T*B, 5, Cp]
    uint16_t* dz16;             // [T*B, 4Cp] packed
    float* dbias; float* dw_i; float* dw_f; float* dw_o;
};

template <int NHALF>
__global__ void __launch_bounds__(128 * NHALF, 1)
lstmp_bwd_cluster_kernel(const CBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, hh = warp >> 2;
    const int htid = tid & 127;
    const int MT = p.Cp / 128;                  // output-row tiles (cells) of the K-slice product: 2 or 4
    const int G = p.Cp / 32;
    const uint32_t j = cluster_ctarank();
    const int grp = blockIdx.x / G;
    const int b0 = grp * (NB * NHALF) + hh * NB;

    constexpr uint32_t SLOT = NB * 64u;                              // bytes one source CTA sends per step: [2][32 cells][8] 16-bit
    const uint32_t sR_bytes = (uint32_t)G * SLOT;
    const uint32_t half_bytes = 2u * NB * 128u + 2u * sR_bytes;      // per half: dz tile + two receive buffers
    const uint32_t sB = base + (uint32_t)hh * half_bytes;            // 2 x [NB x 64] 16-bit, SW128 (dz of my 128 gate columns)
    const uint32_t sR0 = sB + 2u * NB * 128u;
    const uint32_t sBar = base + (uint32_t)NHALF * half_bytes + (uint32_t)hh * 32u;
    const uint32_t barM = sBar, full0 = sBar + 8, full1 = sBar + 16;
    const uint32_t tslot = base + (uint32_t)NHALF * half_bytes + (uint32_t)NHALF * 32u;
    uint8_t* sB_ptr = base_ptr + (sB - base);

    // TMEM: tile mt of the A operand (rows = cells 128 mt.., K = my 128 gate columns) at columns [64 mt, 64 mt + 64),
    //       accumulator (half hh, tile mt) at columns 64 MT + (hh MT + mt) * 16
    const uint32_t a_cols = 64u * (uint32_t)MT;
    uint32_t tcols = 32;
    while (tcols < a_cols + (uint32_t)(MT * NB * NHALF)) tcols <<= 1;
    if (htid == 0) {
        mbar_init(barM, 1); mbar_init(full0, 1); mbar_init(full1, 1);
        fence_mbar_init();
        mbar_expect_tx(full1, sR_bytes);
        mbar_expect_tx(full0, sR_bytes);
    }
    if (warp == 0) tmem_alloc(tslot, tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    const uint32_t tmem_acc = tmem + a_cols + (uint32_t)(hh * MT * NB);

    for (int it = hh; it < 2 * MT; it += NHALF) {   // weight slab -> TMEM: thread <-> output row, 64 k (32 columns) per store
        const int mt = it >> 1, kh = it & 1;
        const uint16_t* wrow = p.wc + (size_t)(128 * mt + 32 * q + lane) * 4 * p.Cp + 128 * j + 64 * kh;
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow) + c);
            r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
        }
        tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * mt + 32 * kh), r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();

    // gate-backward ownership: thread <-> (cell 32j + lane, utterances n = 4 q + u of this half)
    constexpr int UPT = 4;
    const int cell = 32 * (int)j + lane;
    const float wi = p.w_i[cell], wf = p.w_f[cell], wo = p.w_o[cell];
    float dcar[UPT];
    int len[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) {
        dcar[u] = 0.f;
        const int b = b0 + UPT * q + u;
        len[u] = b < p.B ? p.lengths[b] : 0;
    }
    float a_dwi = 0.f, a_dwf = 0.f, a_dwo = 0.f, a_db[4] = {0.f, 0.f, 0.f, 0.f};
    // rows 128 mt + 32 q + lane of the product belong to CTA 4 mt + q
    uint32_t rd[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) rd[mt] = mt < MT ? mapa_u32(base, (uint32_t)(4 * mt + q)) - base : 0u;
    const uint32_t idesc = umma_idesc(128, NB, p.bf, 0, 0);
    const size_t Cp = (size_t)p.Cp;
    // my 4 utterances inside the 16-byte chunk [h][cell][8 utterances] of a source slot
    const uint32_t roff = (uint32_t)(q >> 1) * 512u + (uint32_t)lane * 16u + (uint32_t)(q & 1) * 8u;
    const uint32_t bar_id = 1u + (uint32_t)hh;

    // Saved forward activations (i, f, o, tanh j, c) and the projection term dmt do not depend on the recurrence, but
    // they come from HBM (the save buffer of one layer is larger than L2): they are loaded ONE STEP AHEAD into
    // registers, so their latency hides behind the previous step instead of sitting in the dependent chain
    // (measured: 1100 of 4000 cycles per step were spent waiting for these loads).  c_{t-1} of step t is c_t of step
    // t-1, so the cell state is fetched two steps ahead and handed down.
    float s_i[UPT], s_f[UPT], s_o[UPT], s_j[UPT], s_c[UPT], s_cp[UPT], dm[UPT];
    float n_i[UPT], n_f[UPT], n_o[UPT], n_j[UPT], n_cp[UPT], n_dm[UPT];
#pragma unroll
    for (int u = 0; u < UPT; ++u) {
        const int b = b0 + UPT * q + u;
        s_i[u] = s_f[u] = s_o[u] = s_j[u] = s_c[u] = s_cp[u] = dm[u] = 0.f;
        if (b < p.B) {
            const size_t row = (size_t)(p.T - 1) * p.B + b;
            const float* s = p.save + row * 5 * Cp + cell;
            s_i[u] = __ldg(s); s_f[u] = __ldg(s + Cp); s_o[u] = __ldg(s + 2 * Cp); s_j[u] = __ldg(s + 3 * Cp);
            s_c[u] = __ldg(s + 4 * Cp);
            if (p.T > 1) s_cp[u] = __ldg(s - (size_t)p.B * 5 * Cp + 4 * Cp);
            dm[u] = __ldg(p.dmt + row * Cp + cell);
        }
    }

    for (int step = 0; step < p.T; ++step) {
        const int t = p.T - 1 - step;
        const int buf = step & 1;
        TRACE_T(step, 0);
#pragma unroll
        for (int u = 0; u < UPT; ++u) {            // operands of step t-1 (and c_{t-2}): in flight during this step
            const int b = b0 + UPT * q + u;
            n_i[u] = n_f[u] = n_o[u] = n_j[u] = n_cp[u] = n_dm[u] = 0.f;
            if (b < p.B && t > 0) {
                const size_t row = (size_t)(t - 1) * p.B + b;
                const float* s = p.save + row * 5 * Cp + cell;
                n_i[u] = __ldg(s); n_f[u] = __ldg(s + Cp); n_o[u] = __ldg(s + 2 * Cp); n_j[u] = __ldg(s + 3 * Cp);
                if (t > 1) n_cp[u] = __ldg(s - (size_t)p.B * 5 * Cp + 4 * Cp);
                n_dm[u] = __ldg(p.dmt + row * Cp + cell);
            }
        }
        if (step > 0) {
            const uint32_t fb = buf ? full1 : full0;
            mbar_wait(fb, (uint32_t)(((step - 1) >> 1) & 1));   // partial rows of dz_{t+1} Wc^T from all G CTAs
            TRACE_T(step, 1);
            const uint32_t rbase = sR0 + (uint32_t)buf * sR_bytes + roff;
            for (int s0 = 0; s0 < G; s0 += 8) {       // G is 8 or 16: eight loads in flight, then a short add tree
                uint32_t x[8], y[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x[k]), "=r"(y[k]) : "r"(rbase + (uint32_t)(s0 + k) * SLOT));
                float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    a0[0] += h2f((uint16_t)(x[k] & 0xFFFFu), p.bf); a0[1] += h2f((uint16_t)(x[k] >> 16), p.bf);
                    a0[2] += h2f((uint16_t)(y[k] & 0xFFFFu), p.bf); a0[3] += h2f((uint16_t)(y[k] >> 16), p.bf);
                    a1[0] += h2f((uint16_t)(x[k + 1] & 0xFFFFu), p.bf); a1[1] += h2f((uint16_t)(x[k + 1] >> 16), p.bf);
                    a1[2] += h2f((uint16_t)(y[k + 1] & 0xFFFFu), p.bf); a1[3] += h2f((uint16_t)(y[k + 1] >> 16), p.bf);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) dm[u] += a0[u] + a1[u];
            }
            __syncwarp();
            if (htid == 0 && step + 2 < p.T) mbar_expect_tx(fb, sR_bytes);   // re-arm for step + 2
        }
        TRACE_T(step, 2);
        uint16_t hz[UPT][4];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            const int b = b0 + UPT * q + u;
            // branch-free (a per-utterance `if` would serialise the four dependent chains): frozen steps and rows
            // past the batch contribute exact zeros through the mask (their operands are finite or zero)
            const float m = ((b < p.B) && (t < len[u])) ? 1.f : 0.f;
            const float tc = tanhf_(s_c[u]);
            const float dz_o = m * dm[u] * tc * s_o[u] * (1.f - s_o[u]);
            const float dc = dcar[u] + dm[u] * s_o[u] * (1.f - tc * tc) + dz_o * wo;
            const float dz_f = m * dc * s_cp[u] * s_f[u] * (1.f - s_f[u]);
            const float dz_i = m * dc * s_j[u] * s_i[u] * (1.f - s_i[u]);
            const float dz_j = m * dc * s_i[u] * (1.f - s_j[u] * s_j[u]);
            dcar[u] = m * (dc * s_f[u] + dz_f * wf + dz_i * wi);
            a_dwo += dz_o * s_c[u]; a_dwf += dz_f * s_cp[u]; a_dwi += dz_i * s_cp[u];
            a_db[0] += dz_i; a_db[1] += dz_j; a_db[2] += dz_f; a_db[3] += dz_o;
            hz[u][0] = f2h(dz_i, p.bf); hz[u][1] = f2h(dz_j, p.bf); hz[u][2] = f2h(dz_f, p.bf); hz[u][3] = f2h(dz_o, p.bf);
        }
        TRACE_T(step, 3);
        // B operand of the recurrent product first (it is what the next MMA waits for), the global copy after
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            const int n = UPT * q + u;
            // K-major, SW128: row n, local packed gate column g*32 + lane -> k-subtile g/2
            *reinterpret_cast<uint16_t*>(sB_ptr + sw128_off(n, lane)) = hz[u][0];                        // g = 0
            *reinterpret_cast<uint16_t*>(sB_ptr + sw128_off(n, 32 + lane)) = hz[u][1];                   // g = 1
            *reinterpret_cast<uint16_t*>(sB_ptr + NB * 128 + sw128_off(n, lane)) = hz[u][2];             // g = 2
            *reinterpret_cast<uint16_t*>(sB_ptr + NB * 128 + sw128_off(n, 32 + lane)) = hz[u][3];        // g = 3
        }
        if (t == 0) {                          // last step: only the global copy of dz is left
#pragma unroll
            for (int u = 0; u < UPT; ++u) {
                const int b = b0 + UPT * q + u;
                if (b < p.B) {
                    uint16_t* d = p.dz16 + ((size_t)t * p.B + b) * 4 * Cp + 128 * j + lane;
                    d[0] = hz[u][0]; d[32] = hz[u][1]; d[64] = hz[u][2]; d[96] = hz[u][3];
                }
            }
        }
        if (t == 0) break;                     // no earlier step to feed
        TRACE_T(step, 4);
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (q == 0) {
            tc_fence_after();
            if (elect_one_sync()) {
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint64_t db = umma_desc_sw128(sB + (uint32_t)(kk >> 2) * (NB * 128u) + (uint32_t)(kk & 3) * 32u, 16, 1024);
                        tc_mma_f16_ts(tmem_acc + (uint32_t)(mt * NB), tmem + (uint32_t)(64 * mt + 8 * kk), db, idesc, kk ? 1u : 0u);
                    }
                }
                tc_commit(barM);
            }
        }
        __syncwarp();
        TRACE_T(step, 5);
        // global copy of dz (operand of the weight-gradient GEMMs): off the dependent chain, behind the MMAs
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            const int b = b0 + UPT * q + u;
            if (b < p.B) {
                uint16_t* d = p.dz16 + ((size_t)t * p.B + b) * 4 * Cp + 128 * j + lane;
                d[0] = hz[u][0]; d[32] = hz[u][1]; d[64] = hz[u][2]; d[96] = hz[u][3];
            }
        }
        mbar_wait(barM, (uint32_t)(step & 1));
        tc_fence_after();
        TRACE_T(step, 6);
        const uint32_t dst0 = sR0 + (uint32_t)(buf ^ 1) * sR_bytes + j * SLOT + (uint32_t)lane * 16u;
        const uint32_t dbar = buf ? full0 : full1;
        uint32_t acc[4][NB];                   // all MT accumulator tiles in flight, one wait
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
            if (mt < MT) tmem_ld16_nowait(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NB), acc[mt]);
        tmem_ld_wait();
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            if (mt < MT) {
                // row 128 mt + 32 q + lane = cell `lane` of CTA 4 mt + q
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t* a = acc[mt] + 8 * c;
                    st_async_v4(dst0 + (uint32_t)c * 512u + rd[mt],
                                pack2(__uint_as_float(a[0]), __uint_as_float(a[1]), p.bf),
                                pack2(__uint_as_float(a[2]), __uint_as_float(a[3]), p.bf),
                                pack2(__uint_as_float(a[4]), __uint_as_float(a[5]), p.bf),
                                pack2(__uint_as_float(a[6]), __uint_as_float(a[7]), p.bf), dbar + rd[mt]);
                }
            }
        }
        tc_fence_before();
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            s_c[u] = s_cp[u]; s_cp[u] = n_cp[u];
            s_i[u] = n_i[u]; s_f[u] = n_f[u]; s_o[u] = n_o[u]; s_j[u] = n_j[u]; dm[u] = n_dm[u];
        }
        TRACE_T(step, 7);
    }
    atomicAdd(p.dw_i + cell, a_dwi); atomicAdd(p.dw_f + cell, a_dwf); atomicAdd(p.dw_o + cell, a_dwo);
    {
        float* db = p.dbias + 128 * j + lane;
        atomicAdd(db, a_db[0]); atomicAdd(db + 32, a_db[1]); atomicAdd(db + 64, a_db[2]); atomicAdd(db + 96, a_db[3]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tcols);
    cluster_sync_all();
}

size_t cfwd_smem(int Cp, int nhalf, int Ik = 0) {
    const size_t xt = (size_t)((Ik + 63) / 64) * NB * 128;
    const size_t need = 1024 + (size_t)nhalf * (2 * xt + 2 * (size_t)Cp * NB * 2 + 2 * 128 * (size_t)(NB + 1) * 4) + (size_t)nhalf * 64 + 64;
    return need < RSR_EXCLUSIVE_SMEM_REC ? RSR_EXCLUSIVE_SMEM_REC : need;
}
size_t cbwd_smem(int Cp, int nhalf) {
    const size_t need = 1024 + (size_t)nhalf * (2 * (size_t)NB * 128 + 2 * (size_t)(Cp / 32) * NB * 64) + (size_t)nhalf * 32 + 64;
    return need < RSR_EXCLUSIVE_SMEM_REC ? RSR_EXCLUSIVE_SMEM_REC : need;
}

// halves (independent 16-utterance recurrences) per cluster: 1 when that many clusters run concurrently
// (shortest step), else 2 (the halves overlap each other's exchange)
template <typename K1, typename K2>
int pick_cluster_halves(rsr_handle* h, int which, K1 k1, K2 k2, int B, int Cp, size_t s1, size_t s2) {
    const int G = Cp / 32;
    int* cap = h->cluster_cap[which][Cp / 256 - 1];
    if (cap[0] < 0) {
        std::lock_guard<std::mutex> g(h->mu);
        cap[0] = s1 <= (size_t)h->max_smem ? cluster_capacity(k1, G, 128, s1) : 0;
        cap[1] = s2 <= (size_t)h->max_smem ? cluster_capacity(k2, G, 256, s2) : 0;
        if (getenv("RSR_DEBUG"))
            fprintf(stderr, "[rsr] %s cluster kernel Cp=%d: %d-CTA clusters co-resident: 1 half %d, 2 halves %d\n",
                    which == 1 ? "bwd" : which == 2 ? "fused fwd" : "fwd", Cp, G, cap[0], cap[1]);
    }
    const int g1 = (B + NB - 1) / NB;
    if (cap[0] > 0 && (g1 <= cap[0] || cap[1] <= 0)) return 1;
    if (cap[1] > 0) return 2;
    return 0;
}

}  // namespace

// 3-D view of mt_seq [(T+1)*B, Cp] for the multicast exchange: (8 cells = 16 B | row | k-chunk of 8 cells), box
// (8, NB, 4) = one CTA's 32-cell slice of NB utterances, landing in shared memory as [k-chunk][row][16 B] -- the
// no-swizzle K-major core-matrix layout of the recurrent B operand.
int make_mt_tmap(rsr_handle* h, const void* mt_seq, int B, int T, int Cp, CUtensorMap* out) {
    cuuint64_t dims[3] = {8, (cuuint64_t)(T + 1) * B, (cuuint64_t)Cp / 8};
    cuuint64_t strides[2] = {(cuuint64_t)Cp * 2, 16};
    cuuint32_t box[3] = {8, (cuuint32_t)NB, 4};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(mt_seq), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RSR_E_ARG;
}

// Exchange mode of the forward kernels: st.async all-gather over DSMEM (default) or, with RSR_FWD_XCHG=l2mc, through
// the global copy of mt_t + one multicast TMA load per CTA and step.  Measured on B200 (Cp = 512): the same step time
// at B = 128 (2.87 vs 2.96 us) and at one half per cluster (2.27 vs 2.33), slower at Cp = 256 (2.24 vs 1.97) -- the
// exchange costs ~2300 cycles of latency either way (st.async issue + landing vs store + proxy fence + L2 read), so
// the 16 KB/step of DSMEM traffic is not what paces the kernel.  Kept as a tested alternative.
int fwd_xmode(int nhalf) {
    (void)nhalf;
    const char* e = getenv("RSR_FWD_XCHG");
    return (e && e[0] == 'l') ? 1 : 0;
}

// Returns 0 when launched, RSR_E_RESIDENT when the cluster variant does not apply (caller falls back to
// the L2-exchange kernels of lstmp_sm100.cu), other values on error.
int rsr_lstmp_fwd_cluster(rsr_handle* h, void* stream, int B, int T, int Cp, const float* zx, const void* wcT,
                          const float* w_i, const float* w_f, const float* w_o, float forget_bias,
                          const int* lengths, void* mt_seq, float* save) {
    if (Cp > 512) return RSR_E_RESIDENT;
    const int G = Cp / 32;
    const size_t s1 = cfwd_smem(Cp, 1), s2 = cfwd_smem(Cp, 2);
    const int nh = pick_cluster_halves(h, 0, lstmp_fwd_cluster_kernel<1, false>, lstmp_fwd_cluster_kernel<2, false>, B, Cp, s1, s2);
    if (!nh) return RSR_E_RESIDENT;
    const int groups = (B + NB * nh - 1) / (NB * nh);
    CFwdParams p;
    CUtensorMap tmX = {}, tmM = {};   // tmX unused by the unfused variant
    p.kxT = nullptr; p.bias = nullptr; p.Ik = 0;
    p.xmode = fwd_xmode(nh);
    if (p.xmode) { const int rc = make_mt_tmap(h, mt_seq, B, T, Cp, &tmM); if (rc) return rc; }
    p.wcT = (const uint16_t*)wcT;
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16; p.forget_bias = forget_bias;
    p.zx = zx; p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths;
    p.mt_seq = (uint16_t*)mt_seq; p.save = save;
    if (nh == 1) return cluster_launch(lstmp_fwd_cluster_kernel<1, false>, groups, G, 128, s1, (cudaStream_t)stream, tmX, tmM, p);
    return cluster_launch(lstmp_fwd_cluster_kernel<2, false>, groups, G, 256, s2, (cudaStream_t)stream, tmX, tmM, p);
}

// Fully fused forward: input GEMM + recurrent GEMM + gate epilogue.  RSR_E_RESIDENT when K_x^T does not fit in
// TMEM next to Wc^T (caller then runs rsr_gemm for Zx and the unfused kernel).
int rsr_lstmp_fused_fwd_cluster(rsr_handle* h, void* stream, int B, int T, int I, int Cp, const void* x16, int ldx,
                                const void* kxT, const float* bias, const void* wcT, const float* w_i,
                                const float* w_f, const float* w_o, float forget_bias, const int* lengths,
                                void* mt_seq, float* save) {
    if (Cp > 512) return RSR_E_RESIDENT;
    const int Ik = (I + 15) & ~15;
    if (Cp / 2 + Ik / 2 + 2 * NB > 512) return RSR_E_RESIDENT;
    const int G = Cp / 32;
    const size_t s1 = cfwd_smem(Cp, 1, Ik), s2 = cfwd_smem(Cp, 2, Ik);
    // capacity is cached per Cp for the fused variant too (its shared memory grows with Ik: re-query when it changes)
    if (h->fused_ik[Cp / 256 - 1] != Ik) { h->cluster_cap[2][Cp / 256 - 1][0] = h->cluster_cap[2][Cp / 256 - 1][1] = -1; h->fused_ik[Cp / 256 - 1] = Ik; }
    const int nh = pick_cluster_halves(h, 2, lstmp_fwd_cluster_kernel<1, true>, lstmp_fwd_cluster_kernel<2, true>, B, Cp, s1, s2);
    if (!nh) return RSR_E_RESIDENT;
    const int groups = (B + NB * nh - 1) / (NB * nh);
    CUtensorMap tmX;
    int rc = rsr_get_tmap(h, x16, (uint64_t)ldx, (uint64_t)T * B, (uint64_t)ldx, 64, NB, &tmX);
    if (rc) return rc;
    CFwdParams p;
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16; p.forget_bias = forget_bias;
    p.zx = nullptr; p.wcT = (const uint16_t*)wcT; p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths;
    p.mt_seq = (uint16_t*)mt_seq; p.save = save;
    p.kxT = (const uint16_t*)kxT; p.bias = bias; p.Ik = Ik;
    CUtensorMap tmM = {};
    p.xmode = fwd_xmode(nh);
    if (p.xmode) { rc = make_mt_tmap(h, mt_seq, B, T, Cp, &tmM); if (rc) return rc; }
    if (nh == 1) return cluster_launch(lstmp_fwd_cluster_kernel<1, true>, groups, G, 128, s1, (cudaStream_t)stream, tmX, tmM, p);
    return cluster_launch(lstmp_fwd_cluster_kernel<2, true>, groups, G, 256, s2, (cudaStream_t)stream, tmX, tmM, p);
}

int rsr_lstmp_bwd_cluster(rsr_handle* h, void* stream, int B, int T, int Cp, const float* dmt, const void* wc,
                          const float* w_i, const float* w_f, const float* w_o, const int* lengths,
                          const float* save, void* dz16, float* dbias, float* dw_i, float* dw_f, float* dw_o) {
    if (Cp > 512) return RSR_E_RESIDENT;
    const int G = Cp / 32;
    const size_t s1 = cbwd_smem(Cp, 1), s2 = cbwd_smem(Cp, 2);
    const int nh = pick_cluster_halves(h, 1, lstmp_bwd_cluster_kernel<1>, lstmp_bwd_cluster_kernel<2>, B, Cp, s1, s2);
    if (!nh) return RSR_E_RESIDENT;
    const int groups = (B + NB * nh - 1) / (NB * nh);
    CBwdParams p;
    p.wc = (const uint16_t*)wc;
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16;
    p.dmt = dmt; p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths; p.save = save;
    p.dz16 = (uint16_t*)dz16; p.dbias = dbias; p.dw_i = dw_i; p.dw_f = dw_f; p.dw_o = dw_o;
    if (nh == 1) return cluster_launch(lstmp_bwd_cluster_kernel<1>, groups, G, 128, s1, (cudaStream_t)stream, p);
    return cluster_launch(lstmp_bwd_cluster_kernel<2>, groups, G, 256, s2, (cudaStream_t)stream, p);
}

// debug: copies the phase-timing trace (all zeros unless built with -DRSR_TRACE) to the host
extern "C" int rsr_debug_trace(unsigned long long* host_out, int n) {
#ifdef RSR_TRACE
    if (!host_out || n <= 0 || n > 8192) return RSR_E_ARG;
    RSR_CHECK_CUDA(cudaDeviceSynchronize());
    RSR_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_rsr_trace, sizeof(unsigned long long) * n));
    return 0;
#else
    (void)host_out; (void)n;
    return RSR_E_ARG;
#endif
}
