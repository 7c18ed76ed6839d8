// Persistent LSTMP (LSTM with peepholes + projection) recurrence kernels for sm_100a.
//
// Reference semantics: tf.contrib.rnn.LSTMCell(use_peepholes=True, num_proj=P, forget_bias=1.0)
// under tf.nn.dynamic_rnn(sequence_length=lengths, zero initial state) -- call sites
// models/lstm.py:89-112, models/res_lstm_l.py:86-138, models/discriminator_lstm.py:70-91; the
// gate math (order i, j, f, o; peepholes; bias-free projection) is restated in the reference at
// models/BNLSTMCell.py:176-213.
//
// B200 design.  The recurrence is folded: with mt_t = o (.) tanh(c_t) (the pre-projection
// output) and Wc = W_proj K_h (C x 4C),
//        z_t = Zx_t + mt_{t-1} Wc                 (Zx = X K_x + b hoisted into one big GEMM)
//        out_t = mt_t W_proj                      (hoisted into one big GEMM after the loop)
// so a time step is ONE dependent matmul.  The forward kernel keeps its slice of Wc resident in
// shared memory for all T steps (weight-stationary), computes the four gate pre-activations of
// 32 cells x NB utterances per CTA as one tcgen05 tile (swap-AB: 128 gate rows = UMMA M,
// utterances = UMMA N, K = C) whose B operand (mt_{t-1}) arrives by TMA, adds Zx, and applies
// the sigmoid/tanh/peephole/cell update in registers.  CTAs that share an utterance slice form
// a group that exchanges mt_t through L2 and synchronises with one release/acquire counter per
// step; different utterance slices never synchronise.  The backward kernel runs the reversed
// recurrence dmt_{t-1} += dz_t Wc^T the same way, split 2-D (gate-column slices x output-row
// slices) so that every CTA's Wc slab fits in shared memory; partial sums meet in L2 via
// coalesced red.global.add.f32.
#include <stdlib.h>

#include "common.cuh"
#include "handle.h"

using namespace rsr;

namespace {

// =========================================================================================
// forward
// =========================================================================================
struct FwdParams {
    int B, T, Cp, bf;
    float forget_bias;
    const float* zx;            // [T*B, 4Cp] packed gate columns
    const float* w_i; const float* w_f; const float* w_o;   // [Cp]
    const int* lengths;         // [B]
    uint16_t* mt_seq;           // [(T+1)*B, Cp]
    float* save;                // [T*B, 5, Cp] or null
    unsigned int* flags;        // one counter per utterance group
    int kb_smem;                // K sub-tiles of the weight slab held in shared memory; the rest live in TMEM
    const uint16_t* wcT;        // [4Cp, Cp] (source of the TMEM-resident part)
};

template <int NB>
__global__ void __launch_bounds__(128, 1)
lstmp_rec_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmM,
                     const FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = p.Cp / 64;                 // 64-wide K sub-tiles
    const int G = p.Cp / 32;                  // CTAs per utterance group
    const int j = blockIdx.x % G;             // cell block: cells [32j, 32j+32)
    const int grp = blockIdx.x / G;
    const int b0 = grp * NB;

    // Weight slab of this CTA: 128 packed gate rows x Cp.  K sub-tiles [0, KS) are resident in shared memory (TMA,
    // SW128); for Cp > 768 the slab (256 KB at Cp = 1024) does not fit, so sub-tiles [KS, KB) are resident in TMEM
    // instead and enter the same accumulation as the tcgen05 A-from-TMEM operand.
    const int KS = p.kb_smem;
    const uint32_t sA = base;                                // KS x [128 x 64] 16-bit, SW128
    const uint32_t sB = sA + (uint32_t)KS * 16384u;          // KB x [NB x 64]
    const uint32_t sX = sB + (uint32_t)KB * NB * 128u;       // float xchg[4][32][NB+1]
    const uint32_t sBar = sX + 4u * 32u * (NB + 1) * 4u;
    const uint32_t barA = (sBar + 7u) & ~7u, barB = barA + 8, barM = barA + 16, tslot = barA + 24;
    float* xchg = reinterpret_cast<float*>(smem_raw + (sX - smem_u32(smem_raw)));

    constexpr uint32_t ACOLS = NB < 32 ? 32 : NB;            // accumulator columns
    uint32_t TCOLS = ACOLS;
    while (TCOLS < ACOLS + 32u * (uint32_t)(KB - KS)) TCOLS <<= 1;
    if (tid == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmM);
        mbar_init(barA, 1); mbar_init(barB, 1); mbar_init(barM, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tslot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    const uint32_t tmem_a = tmem + ACOLS;

    if (warp == 0 && elect_one_sync()) {   // weight slab: resident for the whole sequence
        mbar_expect_tx(barA, (uint32_t)KS * 16384u);
        for (int kb = 0; kb < KS; ++kb) tma_load_2d(sA + kb * 16384u, &tmW, barA, kb * 64, 128 * j);
    }
    if (KS < KB) {   // thread <-> gate row, 64 k (32 columns of packed 16-bit pairs) per store
        const uint16_t* wrow = p.wcT + (size_t)(128 * j + tid) * p.Cp;
        for (int kb = KS; kb < KB; ++kb) {
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + kb * 64) + c);
                r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
            }
            tmem_st32(tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kb - KS) * 32u, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    // phase-2 ownership: thread <-> (cell cl, utterances n = q*NQ .. q*NQ+NQ-1)
    constexpr int NQ = NB / 4;
    const int cl = lane, q = warp;
    const int cell = 32 * j + cl;
    const float wi = p.w_i[cell], wf = p.w_f[cell], wo = p.w_o[cell];
    float creg[NQ];
    int len[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        creg[i] = 0.f;
        const int b = b0 + q * NQ + i;
        len[i] = b < p.B ? p.lengths[b] : 0;
    }
    const uint32_t idesc = umma_idesc(128, NB, p.bf, 0, 0);
    const size_t zx_ld = (size_t)4 * p.Cp;
    unsigned int* flag = p.flags + grp;

    for (int t = 0; t < p.T; ++t) {
        // Zx for gate row `tid` (independent of the recurrence: issue early)
        float zxv[NB];
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int b = b0 + n;
            zxv[n] = b < p.B ? __ldg(p.zx + ((size_t)t * p.B + b) * zx_ld + 128 * j + tid) : 0.f;
        }
        if (warp == 0) {   // warp-uniform; the elected lane issues TMA / MMA (operands in uniform registers)
            if (t > 0) {
                const unsigned int want = (unsigned int)G * (unsigned int)t;
                spin_until_ge(flag, want);
            }
            fence_proxy_async_all();
            if (elect_one_sync()) {
                // (measured and not kept, profiles/r2_rec_steps_big_v2.txt / _v3.txt: one barrier per K sub-tile so that its
                //  MMAs start as soon as it lands -- 5.5 instead of 4.6 us per step at Cp = 1024, sixteen try_wait round
                //  trips cost more than the overlap gains; the saved activations stored behind the release -- no change)
                mbar_expect_tx(barB, (uint32_t)KB * NB * 128u);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(sB + kb * NB * 128u, &tmM, barB, kb * 64, t * p.B + b0);
                if (t == 0) mbar_wait(barA, 0);
                mbar_wait(barB, (uint32_t)(t & 1));
                tc_fence_after();
                for (int kb = 0; kb < KS; ++kb) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = umma_desc_sw128(sA + kb * 16384u + k * 32u, 16, 1024);
                        const uint64_t db = umma_desc_sw128(sB + kb * NB * 128u + k * 32u, 16, 1024);
                        tc_mma_f16(tmem, da, db, idesc, (kb | k) ? 1u : 0u);
                    }
                }
                for (int kb = KS; kb < KB; ++kb) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t db = umma_desc_sw128(sB + kb * NB * 128u + k * 32u, 16, 1024);
                        tc_mma_f16_ts(tmem, tmem_a + (uint32_t)((kb - KS) * 32 + k * 8), db, idesc, 1u);
                    }
                }
                tc_commit(barM);
            }
        }
        __syncwarp();
        mbar_wait(barM, (uint32_t)(t & 1));
        tc_fence_after();
        float acc[NB];
        if constexpr (NB == 16) tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), acc);
        else tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), acc);
        // warp w holds gate w (i, j, f, o) of cells 32j + lane
#pragma unroll
        for (int n = 0; n < NB; ++n) xchg[(warp * 32 + lane) * (NB + 1) + n] = acc[n] + zxv[n];
        tc_fence_before();
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const int n = q * NQ + i;
            const int b = b0 + n;
            const float zi = xchg[(0 * 32 + cl) * (NB + 1) + n];
            const float zj = xchg[(1 * 32 + cl) * (NB + 1) + n];
            const float zf = xchg[(2 * 32 + cl) * (NB + 1) + n];
            const float zo = xchg[(3 * 32 + cl) * (NB + 1) + n];
            const float cp = creg[i];
            const float ig = sigmoidf_(zi + wi * cp);
            const float fg = sigmoidf_(zf + p.forget_bias + wf * cp);
            const float jg = tanhf_(zj);
            const float cn = fg * cp + ig * jg;
            const float og = sigmoidf_(zo + wo * cn);
            const float mt = og * tanhf_(cn);
            const bool active = t < len[i];
            if (b < p.B) {
                const size_t row = (size_t)t * p.B + b;
                if (p.save) {
                    float* s = p.save + row * 5 * p.Cp + cell;
                    s[0] = ig; s[p.Cp] = fg; s[2 * (size_t)p.Cp] = og; s[3 * (size_t)p.Cp] = jg; s[4 * (size_t)p.Cp] = cn;
                }
                p.mt_seq[(row + p.B) * p.Cp + cell] = f2h(active ? mt : 0.f, p.bf);
            }
            if (active) creg[i] = cn;
        }
        __syncthreads();
        if (tid == 0) red_release_add(flag, 1u);   // release: orders the CTA's writes observed via bar.sync
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// =========================================================================================
// backward
// =========================================================================================
struct BwdParams {
    int B, T, Cp, bf;
    float* dmt;                 // [T*B, Cp] in: dOut W_p^T ; kernel adds recurrent term
    const float* w_i; const float* w_f; const float* w_o;
    const int* lengths;
    const float* save;          // [T*B, 5, Cp]
    uint16_t* dz16;             // [T*B, 4Cp] packed
    float* dbias; float* dw_i; float* dw_f; float* dw_o;
    unsigned int* flags;
};

// CTA (ks, ms): gate-column slice ks = 64 cells (256 packed gate columns), output rows
// [256 ms, 256 ms + 256) as two UMMA M=128 tiles.  The block has 8 NB threads (128 for 16 utterances, 256 for 32): every
// thread differentiates ONE cell for 8 utterances and issues 32 reductions per step whatever NB is -- with 128 threads
// the NB = 32 slice (what Cp = 1024, B = 64 needs to stay co-resident) ran 7.7 us per step against 4.4 for NB = 16
// (profiles/r2_rec_steps_big_v0.txt), all of it the doubled per-thread gate math and red.global issue.
template <int NB>
__global__ void __launch_bounds__(8 * NB, 1)
lstmp_rec_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const BwdParams p) {
    constexpr int NPART = NB / 8;                   // utterance lanes of the gate math: threads / 64
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KS = p.Cp / 64, MS = p.Cp / 256;
    const int per_grp = KS * MS;
    const int grp = blockIdx.x / per_grp;
    const int rem = blockIdx.x % per_grp;
    const int ks = rem / MS, ms = rem % MS;
    const int b0 = grp * NB;

    const uint32_t sA = base;                       // 2 m-tiles x 4 k-subtiles x [128 x 64]
    const uint32_t sB = sA + 8u * 16384u;           // 4 k-subtiles x [NB x 64]
    const uint32_t sBar = sB + 4u * NB * 128u;
    const uint32_t barA = sBar, barM = sBar + 8, tslot = sBar + 16;
    uint8_t* sB_ptr = base_ptr + 8 * 16384;

    constexpr uint32_t TCOLS = 2 * NB < 32 ? 32 : 2 * NB;
    if (tid == 0) {
        tma_prefetch_desc(&tmW);
        mbar_init(barA, 1); mbar_init(barM, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tslot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));

    if (warp == 0 && elect_one_sync()) {
        mbar_expect_tx(barA, 8u * 16384u);
        for (int mt = 0; mt < 2; ++mt)
            for (int kb = 0; kb < 4; ++kb)
                tma_load_2d(sA + (mt * 4 + kb) * 16384u, &tmW, barA, 256 * ks + 64 * kb, 256 * ms + 128 * mt);
    }

    // gate-backward ownership: thread <-> (cell c of the slice, utterances n = half, half+2, ...)
    constexpr int NH = NB / NPART;                  // = 8
    const int c = tid & 63, half = tid >> 6;
    const int cell = 64 * ks + c;
    const int jl = c >> 5, c32 = c & 31;
    const float wi = p.w_i[cell], wf = p.w_f[cell], wo = p.w_o[cell];
    float dcar[NH];
    int len[NH];
#pragma unroll
    for (int i = 0; i < NH; ++i) {
        dcar[i] = 0.f;
        const int b = b0 + half + NPART * i;
        len[i] = b < p.B ? p.lengths[b] : 0;
    }
    float a_dwi = 0.f, a_dwf = 0.f, a_dwo = 0.f, a_db[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t idesc = umma_idesc(128, NB, p.bf, 0, 0);
    unsigned int* flag = p.flags + grp;
    const size_t Cp = (size_t)p.Cp;
    if (tid == 0) mbar_wait(barA, 0);      // weight slab landed (also guarantees no TMA is in flight at exit)

    for (int step = 0; step < p.T; ++step) {
        const int t = p.T - 1 - step;
        // saved forward activations do not depend on the recurrence: load before the group wait
        float s_i[NH], s_f[NH], s_o[NH], s_j[NH], s_c[NH], s_cp[NH];
#pragma unroll
        for (int i = 0; i < NH; ++i) {
            const int b = b0 + half + NPART * i;
            if (b < p.B) {
                const float* s = p.save + ((size_t)t * p.B + b) * 5 * Cp + cell;
                s_i[i] = __ldg(s); s_f[i] = __ldg(s + Cp); s_o[i] = __ldg(s + 2 * Cp); s_j[i] = __ldg(s + 3 * Cp);
                s_c[i] = __ldg(s + 4 * Cp);
                s_cp[i] = t > 0 ? __ldg(s - (size_t)p.B * 5 * Cp + 4 * Cp) : 0.f;
            } else {
                s_i[i] = s_f[i] = s_o[i] = s_j[i] = s_c[i] = s_cp[i] = 0.f;
            }
        }
        if (step > 0) {
            if (tid == 0) {
                const unsigned int want = (unsigned int)per_grp * (unsigned int)step;
                spin_until_ge(flag, want);
            }
            __syncthreads();
        }
        // every dmt value of this thread in flight at once (L2: other CTAs' atomics landed there), then branch-free
        // gate math -- a per-utterance `if` with the load inside serialised NH L2 round trips per step
        float dmv[NH];
#pragma unroll
        for (int i = 0; i < NH; ++i) {
            const int b = b0 + half + NPART * i;
            dmv[i] = b < p.B ? __ldcg(p.dmt + ((size_t)t * p.B + b) * Cp + cell) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NH; ++i) {
            const int n = half + NPART * i;
            const int b = b0 + n;
            const float m = ((b < p.B) && (t < len[i])) ? 1.f : 0.f;     // frozen steps / rows past the batch: exact zeros
            const float dm = dmv[i];
            const float tc = tanhf_(s_c[i]);
            const float dz_o = m * dm * tc * s_o[i] * (1.f - s_o[i]);
            const float dc = dcar[i] + dm * s_o[i] * (1.f - tc * tc) + dz_o * wo;
            const float dz_f = m * dc * s_cp[i] * s_f[i] * (1.f - s_f[i]);
            const float dz_i = m * dc * s_j[i] * s_i[i] * (1.f - s_i[i]);
            const float dz_j = m * dc * s_i[i] * (1.f - s_j[i] * s_j[i]);
            dcar[i] = m * (dc * s_f[i] + dz_f * wf + dz_i * wi);
            a_dwo += dz_o * s_c[i]; a_dwf += dz_f * s_cp[i]; a_dwi += dz_i * s_cp[i];
            a_db[0] += dz_i; a_db[1] += dz_j; a_db[2] += dz_f; a_db[3] += dz_o;
            const uint16_t h_i = f2h(dz_i, p.bf), h_j = f2h(dz_j, p.bf), h_f = f2h(dz_f, p.bf), h_o = f2h(dz_o, p.bf);
            // B operand tile (K-major, SW128): row n, local packed column jl*128 + g*32 + c32
            const int cb = (jl * 2) * NB * 128;
            *reinterpret_cast<uint16_t*>(sB_ptr + cb + sw128_off(n, c32)) = h_i;                       // g = 0
            *reinterpret_cast<uint16_t*>(sB_ptr + cb + sw128_off(n, 32 + c32)) = h_j;                  // g = 1
            *reinterpret_cast<uint16_t*>(sB_ptr + cb + NB * 128 + sw128_off(n, c32)) = h_f;            // g = 2
            *reinterpret_cast<uint16_t*>(sB_ptr + cb + NB * 128 + sw128_off(n, 32 + c32)) = h_o;       // g = 3
            if (ms == 0 && b < p.B) {
                uint16_t* d = p.dz16 + ((size_t)t * p.B + b) * 4 * Cp + 256 * ks + jl * 128 + c32;
                d[0] = h_i; d[32] = h_j; d[64] = h_f; d[96] = h_o;
            }
        }
        if (t == 0) break;                     // no earlier step to feed
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one_sync()) {
                for (int mt = 0; mt < 2; ++mt) {
                    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = umma_desc_sw128(sA + (mt * 4 + kb) * 16384u + k * 32u, 16, 1024);
                            const uint64_t db = umma_desc_sw128(sB + kb * NB * 128u + k * 32u, 16, 1024);
                            tc_mma_f16(tmem + (uint32_t)(mt * NB), da, db, idesc, (kb | k) ? 1u : 0u);
                        }
                    }
                }
                tc_commit(barM);
            }
        }
        __syncwarp();
        mbar_wait(barM, (uint32_t)(step & 1));
        tc_fence_after();
        // warp w reads TMEM lane quadrant w % 4: four warps drain both m-tiles (NB = 16), eight warps one tile each (NB = 32)
#pragma unroll
        for (int mi = 0; mi < (NB == 16 ? 2 : 1); ++mi) {
            const int mt = NB == 16 ? mi : (warp >> 2);
            const int wq = warp & 3;
            float acc[NB];
            if constexpr (NB == 16) tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * NB), acc);
            else tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * NB), acc);
            const int crow = 256 * ms + 128 * mt + wq * 32 + lane;
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const int b = b0 + n;
                if (b < p.B) red_add_f32(p.dmt + ((size_t)(t - 1) * p.B + b) * Cp + crow, acc[n]);
            }
        }
        tc_fence_before();
        __syncthreads();
        if (tid == 0) red_release_add(flag, 1u);   // release: orders the CTA's writes observed via bar.sync
    }
    if (ms == 0) {
        atomicAdd(p.dw_i + cell, a_dwi); atomicAdd(p.dw_f + cell, a_dwf); atomicAdd(p.dw_o + cell, a_dwo);
        float* db = p.dbias + 256 * ks + jl * 128 + c32;
        atomicAdd(db, a_db[0]); atomicAdd(db + 32, a_db[1]); atomicAdd(db + 64, a_db[2]); atomicAdd(db + 96, a_db[3]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// K sub-tiles of the forward weight slab kept in shared memory (the rest are TMEM-resident): everything up to
// Cp = 768 (reference-native C = 760), 8 of them (128 KB) beyond that
int fwd_kb_smem(int Cp) { return Cp <= 768 ? Cp / 64 : 8; }
size_t fwd_smem(int Cp, int nb) {
    const int KB = Cp / 64;
    return 1024 + (size_t)fwd_kb_smem(Cp) * 16384 + (size_t)KB * nb * 128 + 4 * 32 * (nb + 1) * 4 + 64;
}
size_t bwd_smem(int nb) { return 1024 + 8 * 16384 + (size_t)4 * nb * 128 + 64; }

// smallest utterance slice whose grid is co-resident (1 CTA / SM) and whose smem fits
int pick_nb(int B, int ctas_per_group, int num_sms, int max_smem, int Cp, bool fwd) {
    for (int nb = 16; nb <= 32; nb *= 2) {
        const int groups = (B + nb - 1) / nb;
        const size_t smem = fwd ? fwd_smem(Cp, nb) : bwd_smem(nb);
        if (groups * ctas_per_group <= num_sms && (int)smem <= max_smem) return nb;
    }
    return 0;
}

unsigned int* take_flags(rsr_handle* h, int n) { return rsr_take_flags(h, n); }

__global__ void zero_u32_kernel(unsigned int* p, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0u;
}

}  // namespace

int rsr_zero_u32(unsigned int* p, int n, cudaStream_t stream) {
    zero_u32_kernel<<<1, 128, 0, stream>>>(p, n);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_lstmp_rec_fwd(rsr_handle* h, void* stream, int B, int T, int Cp, const float* zx,
                                 const void* wcT, const float* w_i, const float* w_f, const float* w_o,
                                 float forget_bias, const int* lengths, void* mt_seq, float* save) {
    if (!h || !zx || !wcT || !w_i || !w_f || !w_o || !lengths || !mt_seq) return RSR_E_ARG;
    if (B <= 0 || T <= 0 || Cp <= 0 || (Cp & 255)) return RSR_E_SHAPE;
    if (((uintptr_t)zx | (uintptr_t)wcT | (uintptr_t)w_i | (uintptr_t)w_f | (uintptr_t)w_o | (uintptr_t)mt_seq | (uintptr_t)save) & 15)
        return RSR_E_ARG;
    if (!getenv("RSR_NO_CLUSTER")) {   // cluster / DSMEM kernel when the cell block count fits one cluster
        const int rcc = rsr_lstmp_fwd_cluster(h, stream, B, T, Cp, zx, wcT, w_i, w_f, w_o, forget_bias, lengths, mt_seq, save);
        if (rcc != RSR_E_RESIDENT) return rcc;
    }
    const int G = Cp / 32;
    if (32 + 32 * (Cp / 64 - fwd_kb_smem(Cp)) > 512) return RSR_E_SHAPE;   // TMEM-resident part + accumulator: Cp <= 1728
    const int nb = pick_nb(B, G, h->num_sms, h->max_smem, Cp, true);
    if (!nb) return RSR_E_RESIDENT;
    const int groups = (B + nb - 1) / nb;
    const size_t smem = fwd_smem(Cp, nb);
    CUtensorMap tmW, tmM;
    int rc = rsr_get_tmap(h, wcT, (uint64_t)Cp, (uint64_t)4 * Cp, (uint64_t)Cp, 64, 128, &tmW);
    if (rc) return rc;
    rc = rsr_get_tmap(h, mt_seq, (uint64_t)Cp, (uint64_t)(T + 1) * B, (uint64_t)Cp, 64, (uint32_t)nb, &tmM);
    if (rc) return rc;
    FwdParams p;
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16; p.forget_bias = forget_bias;
    p.zx = zx; p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths;
    p.mt_seq = (uint16_t*)mt_seq; p.save = save;
    p.kb_smem = fwd_kb_smem(Cp); p.wcT = (const uint16_t*)wcT;
    p.flags = take_flags(h, groups);
    { const int rz = rsr_zero_u32(p.flags, groups, (cudaStream_t)stream); if (rz) return rz; }
    if (nb == 16) {
        RSR_CHECK_CUDA(cudaFuncSetAttribute(lstmp_rec_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        lstmp_rec_fwd_kernel<16><<<groups * G, 128, smem, (cudaStream_t)stream>>>(tmW, tmM, p);
    } else {
        RSR_CHECK_CUDA(cudaFuncSetAttribute(lstmp_rec_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        lstmp_rec_fwd_kernel<32><<<groups * G, 128, smem, (cudaStream_t)stream>>>(tmW, tmM, p);
    }
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_lstmp_fused_fwd(rsr_handle* h, void* stream, int B, int T, int I, int Cp, const void* x16, int ldx,
                                   const void* kxT, const float* bias, const void* wcT, const float* w_i,
                                   const float* w_f, const float* w_o, float forget_bias, const int* lengths,
                                   void* mt_seq, float* save) {
    if (!h || !x16 || !kxT || !bias || !wcT || !w_i || !w_f || !w_o || !lengths || !mt_seq) return RSR_E_ARG;
    if (B <= 0 || T <= 0 || I <= 0 || Cp <= 0 || (Cp & 255) || ldx < I || (ldx & 7)) return RSR_E_SHAPE;
    if (((uintptr_t)x16 | (uintptr_t)kxT | (uintptr_t)wcT | (uintptr_t)w_i | (uintptr_t)w_f | (uintptr_t)w_o |
         (uintptr_t)mt_seq | (uintptr_t)save) & 15)
        return RSR_E_ARG;
    if (getenv("RSR_NO_CLUSTER") || getenv("RSR_NO_FUSEX")) return RSR_E_RESIDENT;
    if (!getenv("RSR_NO_PAIR")) {   // CTA-pair kernel first (half the exchange traffic and MMA issues per utterance)
        const int rcp = rsr_lstmp_fused_fwd_pair(h, stream, B, T, I, Cp, x16, ldx, kxT, bias, wcT, w_i, w_f, w_o, forget_bias,
                                                 lengths, mt_seq, save);
        if (rcp != RSR_E_RESIDENT) return rcp;
    }
    return rsr_lstmp_fused_fwd_cluster(h, stream, B, T, I, Cp, x16, ldx, kxT, bias, wcT, w_i, w_f, w_o, forget_bias,
                                       lengths, mt_seq, save);
}

extern "C" int rsr_lstmp_rec_bwd(rsr_handle* h, void* stream, int B, int T, int Cp, float* dmt,
                                 const void* wc, const float* w_i, const float* w_f, const float* w_o,
                                 const int* lengths, const float* save, void* dz16,
                                 float* dbias, float* dw_i, float* dw_f, float* dw_o) {
    if (!h || !dmt || !wc || !w_i || !w_f || !w_o || !lengths || !save || !dz16 || !dbias || !dw_i || !dw_f || !dw_o)
        return RSR_E_ARG;
    if (B <= 0 || T <= 0 || Cp <= 0 || (Cp & 255)) return RSR_E_SHAPE;
    if (((uintptr_t)dmt | (uintptr_t)wc | (uintptr_t)save | (uintptr_t)dz16) & 15) return RSR_E_ARG;
    if (!getenv("RSR_NO_CLUSTER") && !getenv("RSR_NO_PAIR") && B > 16) {   // CTA-pair kernel (32 utterances per cluster)
        const int rcp = rsr_lstmp_bwd_pair(h, stream, B, T, Cp, dmt, wc, w_i, w_f, w_o, lengths, save, dz16,
                                           dbias, dw_i, dw_f, dw_o);
        if (rcp != RSR_E_RESIDENT) return rcp;
    }
    if (!getenv("RSR_NO_CLUSTER")) {
        const int rcc = rsr_lstmp_bwd_cluster(h, stream, B, T, Cp, dmt, wc, w_i, w_f, w_o, lengths, save, dz16,
                                              dbias, dw_i, dw_f, dw_o);
        if (rcc != RSR_E_RESIDENT) return rcc;
    }
    const int per_grp = (Cp / 64) * (Cp / 256);
    const int nb = pick_nb(B, per_grp, h->num_sms, h->max_smem, Cp, false);
    if (!nb) return RSR_E_RESIDENT;
    const int groups = (B + nb - 1) / nb;
    const size_t smem = bwd_smem(nb);
    CUtensorMap tmW;
    int rc = rsr_get_tmap(h, wc, (uint64_t)4 * Cp, (uint64_t)Cp, (uint64_t)4 * Cp, 64, 128, &tmW);
    if (rc) return rc;
    BwdParams p;
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16;
    p.dmt = dmt; p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths; p.save = save;
    p.dz16 = (uint16_t*)dz16; p.dbias = dbias; p.dw_i = dw_i; p.dw_f = dw_f; p.dw_o = dw_o;
    p.flags = take_flags(h, groups);
    { const int rz = rsr_zero_u32(p.flags, groups, (cudaStream_t)stream); if (rz) return rz; }
    if (nb == 16) {
        RSR_CHECK_CUDA(cudaFuncSetAttribute(lstmp_rec_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        lstmp_rec_bwd_kernel<16><<<groups * per_grp, 128, smem, (cudaStream_t)stream>>>(tmW, p);
    } else {
        RSR_CHECK_CUDA(cudaFuncSetAttribute(lstmp_rec_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        lstmp_rec_bwd_kernel<32><<<groups * per_grp, 256, smem, (cudaStream_t)stream>>>(tmW, p);
    }
    RSR_LAUNCH_CHECK();
    return 0;
}
