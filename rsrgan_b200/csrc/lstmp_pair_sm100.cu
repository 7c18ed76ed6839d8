// LSTMP recurrence on thread-block clusters, CTA-PAIR variant (sm_100a, Cp <= 512): the fused LSTM-gate kernels.
//
// Same math and C-ABI contract as lstmp_cluster_sm100.cu / lstmp_sm100.cu (reference: models/lstm.py:89-112,
// models/BNLSTMCell.py:176-213).  What changes is who holds the B operand of the recurrent product.  Measured on
// B200 (profiles/r2_trace_v0.txt, r2_dsmem_bw.txt): the single-CTA kernel spends 1000 of its 3460 cycles per step
// ISSUING the st.async all-gather of mt_t (16 KB leave every SM per 16 utterances: 15 B/clk, the DSMEM rate) and
// another 600 waiting for the slowest peer's copy.  Here two CTAs of the cluster form a PAIR that runs ONE
// tcgen05.mma.cta_group::2 (M = 256 gate rows: 128 per CTA, N = 32 utterances): the pair's B operand is split
// along N, so each CTA receives -- and every sender ships to it -- only the 16 utterances it holds.  Per SM and
// step that is 16 KB for 32 utterances instead of 32 KB, half as many MMA instructions (one CTA issues for two;
// an N = 32 MMA costs the same 22 cycles as an N = 16 one, profiles/r2_mma_pair_probe.txt), and the MMAs are
// issued by a dedicated warp that never does gate math, so step t+1's input half (x_{t+1} K_x, which does not
// depend on the recurrence) is already in the accumulator when the exchange lands.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "handle.h"
#include "cluster_util.h"

using namespace rsr;

// Optional in-kernel phase timing (build with -DRSR_TRACE; read back with rsr_debug_trace_pair): gate thread 0 and the
// issuer lane of CTA 0 record %clock64 at fixed points of each time step.
#ifdef RSR_TRACE
__device__ unsigned long long g_rsr_ptrace[8192];
#define PTRACE(cond, tt, slot) do { if (blockIdx.x == 0 && (cond) && (tt) < 64) g_rsr_ptrace[(tt) * 8 + (slot)] = clock64(); } while (0)
#else
#define PTRACE(cond, tt, slot) do { } while (0)
#endif

namespace {

constexpr int NBP = 32;         // utterances per cluster
constexpr int GATE_THREADS = 256;
constexpr int PAIR_THREADS = GATE_THREADS + 32;   // + the issuer / relay warp

__device__ __forceinline__ void tc_mma_f16_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct PFwdParams {
    int B, T, Cp, bf;
    float forget_bias;
    const uint16_t* wcT;        // [4Cp, Cp] packed gate rows of Wc^T
    const uint16_t* kxT;        // [4Cp, Ik] packed gate rows of K_x^T, zero padded to Ik
    const float* bias;          // [4Cp] packed
    const float* w_i; const float* w_f; const float* w_o;   // [Cp]
    const int* lengths;         // [B]
    uint16_t* mt_seq;           // [(T+1)*B, Cp]
    float* save;                // [T*B, 5, Cp] or null
    int Ik;
};

// One cluster (G = Cp/32 CTAs = G/2 pairs) = one group of 32 utterances, run as NCH independent recurrences
// ("chains") of NBC = 32/NCH utterances: NCH = 1 -> one pair MMA of N = 32 per step (the one that is launched: two
// N = 16 chains behind one issuer fall into lockstep and gain nothing, profiles/r2_pair_fwd_variants.txt).
// CTA j owns cells [32j, 32j+32): its 128 packed gate rows of Wc^T and K_x^T are resident in its TMEM (A operand) for
// the whole sequence, and it holds the B-operand rows (mt_{t-1}, x_t) of NBR = NBC/2 utterances of every chain (the
// even CTA of a pair the first half, the odd CTA the second).
// Warps 0-7: gate math (thread <-> 4 cells x 1 utterance; 8/NCH warps per chain); warp 8: in the even CTA of each
// pair the MMA issuer, in the odd CTA the relay that tells the issuer when the odd half of a B operand has landed.
// Tried and measured slower (profiles/r2_pair_fwd_variants_*.txt): shipping every 8-cell k-chunk as soon as its cell
// index is done, either with st.async from the gate warps (they stall on the 15 B/clk DSMEM port: 2766 cycles for gate
// math + sends instead of 927 + 1093) or with 256-byte bulk copies from a sender warp (per-copy overhead and a proxy
// fence per chunk: 5579 cycles per step instead of 3763); sharing reciprocals between gates (7 MUFU operations per cell
// instead of 10: the gate phase is latency-, not MUFU-bound).
template <int NCH>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
lstmp_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmX, const PFwdParams p) {
    constexpr int NBC = NBP / NCH;              // utterances per chain = UMMA N
    constexpr int NBR = NBC / 2;                // B-operand rows per CTA and chain
    constexpr int CW = 8 / NCH;                 // gate warps per chain
    constexpr int XP = NBC + 1;                 // xchg pitch in floats
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = p.Cp / 32;                    // cluster size
    const uint32_t j = cluster_ctarank();       // cell block
    const uint32_t e = j & 1u;                  // which half of every chain's utterances this CTA holds
    const int grp = blockIdx.x / G;
    const int b0 = grp * NBP;

    const uint32_t sB_bytes = (uint32_t)p.Cp * NBR * 2u;             // one B-operand buffer [Cp/8][NBR][8] 16-bit
    const int KBX = (p.Ik + 63) / 64;                                // 64-wide k sub-tiles of the x_t tile
    const uint32_t xt_bytes = (uint32_t)KBX * NBR * 128u;            // this CTA's rows of one x_t tile, SW128 (1024-aligned)
    const uint32_t chain_bytes = 2u * xt_bytes + 2u * sB_bytes;      // per chain: 2 x tiles | 2 B buffers
    const uint32_t sX0 = base + (uint32_t)NCH * chain_bytes;         // float xchg[NCH][128][XP]
    const uint32_t sBar0 = sX0 + (uint32_t)NCH * 128u * XP * 4u;     // per chain 64 B: barM, full0/1, xfull0/1, peer0/1
    const uint32_t tslot = sBar0 + (uint32_t)NCH * 64u;
    auto sXt = [&](int hc, int b) { return base + (uint32_t)hc * chain_bytes + (uint32_t)b * xt_bytes; };
    auto sB = [&](int hc, int b) { return base + (uint32_t)hc * chain_bytes + 2u * xt_bytes + (uint32_t)b * sB_bytes; };
    auto barM = [&](int hc) { return sBar0 + (uint32_t)hc * 64u; };
    auto full = [&](int hc, int b) { return sBar0 + (uint32_t)hc * 64u + 8u + 8u * (uint32_t)b; };
    auto xfull = [&](int hc, int b) { return sBar0 + (uint32_t)hc * 64u + 24u + 8u * (uint32_t)b; };
    auto peer = [&](int hc, int b) { return sBar0 + (uint32_t)hc * 64u + 40u + 8u * (uint32_t)b; };

    // TMEM: [0, Cp/2) Wc^T slice | [Cp/2, Cp/2 + Ik/2) K_x^T slice | accumulators [chain][step parity] of NBC columns
    const uint32_t a_cols = (uint32_t)p.Cp / 2u + (uint32_t)p.Ik / 2u;
    uint32_t tcols = 32;
    while (tcols < a_cols + 2u * NBP) tcols <<= 1;
    if (tid == GATE_THREADS) {
        tma_prefetch_desc(&tmX);
        for (int hc = 0; hc < NCH; ++hc) {
            mbar_init(barM(hc), 1);
            for (int b = 0; b < 2; ++b) { mbar_init(full(hc, b), 1); mbar_init(xfull(hc, b), 1); mbar_init(peer(hc, b), 1); }
        }
        fence_mbar_init();
        for (int hc = 0; hc < NCH; ++hc) {
            mbar_expect_tx(full(hc, 1), sB_bytes);      // armed for step 1 (mt_0 from every CTA of the cluster)
            mbar_expect_tx(full(hc, 0), sB_bytes);      // armed for step 2
        }
    }
    if (warp == 8) tmem_alloc_2cta(tslot, tcols);
    if (tid < GATE_THREADS)                             // m_{-1} = 0
        for (int hc = 0; hc < NCH; ++hc)
            for (uint32_t i = tid; i < sB_bytes / 16u; i += GATE_THREADS) st_shared_v4(sB(hc, 0) + i * 16u, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    const uint32_t tmem_acc = tmem + a_cols;

    if (warp < 8) {   // weight slab -> TMEM, resident for the whole sequence: thread <-> gate row
        const int q = warp & 3, hh = warp >> 2;
        const uint16_t* wrow = p.wcT + (size_t)(128 * j + 32 * q + lane) * p.Cp;
        for (int cb = hh; cb < p.Cp / 64; cb += 2) {          // 64 k (32 columns) per store
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + cb * 64) + c);
                r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
            }
            tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cb * 32u, r);
        }
        const uint16_t* xrow = p.kxT + (size_t)(128 * j + 32 * q + lane) * p.Ik;
        for (int cb = hh; cb < p.Ik / 16; cb += 2) {          // 16 k (8 columns) per store
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(xrow + cb * 16));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(xrow + cb * 16) + 1);
            tmem_st8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)p.Cp / 2u + (uint32_t)cb * 8u,
                     v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // every CTA's barriers are initialised and armed before a remote store / arrive can reach them

    const int KK = p.Cp / 16, KKX = p.Ik / 16;
    const uint32_t lead = mapa_u32(base, j & ~1u) - base;            // shared::cta -> shared::cluster offset of the pair's even CTA

    if (warp == 8) {
        // ------------------------------------------------------------------------------------------------
        // issuer (even CTA) / relay (odd CTA): one elected lane
        // ------------------------------------------------------------------------------------------------
        if (elect_one_sync()) {
            // x_0 tiles; both CTAs' rows complete on the issuer's barrier
            for (int hc = 0; hc < NCH; ++hc) {
                if (e == 0) mbar_expect_tx(xfull(hc, 0), 2u * xt_bytes);
                for (int kb = 0; kb < KBX; ++kb)
                    tma_load_2d_2cta(sXt(hc, 0) + (uint32_t)kb * (NBR * 128u), &tmX, xfull(hc, 0) + lead, kb * 64,
                                     b0 + hc * NBC + (int)e * NBR);
            }
            const uint32_t idesc = umma_idesc(256, NBC, p.bf, 0, 0);
            const uint16_t pair_mask = (uint16_t)(3u << (j & ~1u));
            for (int t = 0; t < p.T; ++t) {
                const int buf = t & 1;
                for (int hc = 0; hc < NCH; ++hc) {
                    if (t + 1 < p.T) {   // prefetch x_{t+1}; its buffer was last read by the MMAs of step t-1 (retired: barM)
                        if (t > 0) mbar_wait(barM(hc), (uint32_t)((t - 1) & 1));
                        if (e == 0) mbar_expect_tx(xfull(hc, buf ^ 1), 2u * xt_bytes);
                        for (int kb = 0; kb < KBX; ++kb)
                            tma_load_2d_2cta(sXt(hc, buf ^ 1) + (uint32_t)kb * (NBR * 128u), &tmX, xfull(hc, buf ^ 1) + lead, kb * 64,
                                             (t + 1) * p.B + b0 + hc * NBC + (int)e * NBR);
                    }
                    if (e == 0) {
                        const uint32_t acc = tmem_acc + (uint32_t)(hc * 2 + buf) * NBC;
                        // input half first: it does not depend on mt_{t-1}, so it runs while the exchange is still in flight
                        mbar_wait(xfull(hc, buf), (uint32_t)((t >> 1) & 1));
                        tc_fence_after();
                        const uint32_t sxt = sXt(hc, buf);
                        for (int kk = 0; kk < KKX; ++kk) {
                            const uint64_t dx = umma_desc_sw128(sxt + (uint32_t)(kk >> 2) * (NBR * 128u) + (uint32_t)(kk & 3) * 32u, 16, 1024);
                            tc_mma_f16_ts_2cta(acc, tmem + (uint32_t)p.Cp / 2u + (uint32_t)kk * 8u, dx, idesc, kk ? 1u : 0u);
                        }
                        if (t > 0) {
                            const uint32_t ph = (uint32_t)(((t - 1) >> 1) & 1);
                            mbar_wait(full(hc, buf), ph);            // my rows of mt_{t-1}: all G slices have landed
                            mbar_wait(peer(hc, buf), ph);            // ... and the odd CTA's rows
                        }
                        PTRACE(hc == 0, t, 1);
                        fence_proxy_async_smem();
                        tc_fence_after();
                        uint64_t db = umma_desc_nosw(sB(hc, buf), NBR * 16u, 128u);
                        uint32_t ta = tmem;
#pragma unroll 8
                        for (int kk = 0; kk < KK; ++kk) {
                            tc_mma_f16_ts_2cta(acc, ta, db, idesc, 1u);
                            ta += 8u;                                    // 16 k = 8 columns
                            db += (uint64_t)((2u * NBR * 16u) >> 4);     // two k-chunks of [NBR rows][16 B]
                        }
                        tc_commit_2cta_mc(barM(hc), pair_mask);
                        if (t > 0 && t + 2 < p.T) mbar_expect_tx(full(hc, buf), sB_bytes);   // re-arm for step t + 2
                        PTRACE(hc == 0, t, 2);
                    } else if (t > 0) {
                        mbar_wait(full(hc, buf), (uint32_t)(((t - 1) >> 1) & 1));
                        fence_proxy_async_smem();
                        mbar_arrive_cluster(peer(hc, buf) + lead);
                        if (t + 2 < p.T) mbar_expect_tx(full(hc, buf), sB_bytes);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------------------------------------
        // gate warps: thread <-> (cells 4a..4a+3 of the block, utterance nl of chain hc)
        // ------------------------------------------------------------------------------------------------
        const int hc = warp / CW, wc = warp % CW;
        const int q = warp & 3, chh = wc >> 2;      // TMEM lane quadrant, 16-column half of the accumulator (NCH = 1)
        const int a = lane & 7;
        const int nl = 4 * wc + (lane >> 3);        // utterance within the chain
        const int cell0 = 32 * (int)j + 4 * a;
        const float4 wi4 = *reinterpret_cast<const float4*>(p.w_i + cell0);
        const float4 wf4 = *reinterpret_cast<const float4*>(p.w_f + cell0);
        const float4 wo4 = *reinterpret_cast<const float4*>(p.w_o + cell0);
        const float wi[4] = {wi4.x, wi4.y, wi4.z, wi4.w}, wf[4] = {wf4.x, wf4.y, wf4.z, wf4.w},
                    wo[4] = {wo4.x, wo4.y, wo4.z, wo4.w};
        float creg[4] = {0.f, 0.f, 0.f, 0.f};
        const int b_own = b0 + hc * NBC + nl;
        const int len = b_own < p.B ? p.lengths[b_own] : 0;
        // destinations of this thread's remote stores: the G/2 CTAs that hold this utterance's row (parity nl / NBR); the
        // two lanes that build one 16-byte chunk (8 cells) split them
        const int HD = G / 4;
        uint32_t rdelta[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            rdelta[i] = i < HD ? mapa_u32(base, (uint32_t)(2 * ((a & 1) * HD + i) + nl / NBR)) - base : 0u;
        const int grow = 32 * q + lane;             // gate row of this thread in the tile (TMEM lane)
        const float bias_r = __ldg(p.bias + 128 * j + grow);
        float* xchg = reinterpret_cast<float*>(base_ptr + (sX0 - base)) + hc * 128 * XP;
        const uint32_t my_barM = barM(hc);
        const uint32_t send_off = (uint32_t)(4 * j + (a >> 1)) * (NBR * 16u) + (uint32_t)(nl % NBR) * 16u;

        for (int t = 0; t < p.T; ++t) {
            const int buf = t & 1;
            PTRACE(tid == 0, t, 0);
            mbar_wait(my_barM, (uint32_t)(t & 1));
            tc_fence_after();
            PTRACE(tid == 0, t, 3);
            float acc[16];
            tmem_ld16(tmem_acc + (uint32_t)(hc * 2 + buf) * NBC + ((uint32_t)(q * 32) << 16) + (uint32_t)chh * 16u, acc);
#pragma unroll
            for (int k = 0; k < 16; ++k) xchg[grow * XP + chh * 16 + k] = acc[k] + bias_r;
            tc_fence_before();
            named_bar_sync(1u + (uint32_t)hc, 32 * CW);
            PTRACE(tid == 0, t, 4);
            const bool active = t < len;
            float mtv[4], sv[5][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cl = 4 * a + c;
                const float zi = xchg[(0 * 32 + cl) * XP + nl];
                const float zj = xchg[(1 * 32 + cl) * XP + nl];
                const float zf = xchg[(2 * 32 + cl) * XP + nl];
                const float zo = xchg[(3 * 32 + cl) * XP + nl];
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
            PTRACE(tid == 0, t, 5);
            // (the exchange buffer is rewritten at step t+1 only after the NEXT barM, which needs every warp's sends below)
            if (t + 1 < p.T) {
                // lanes (a = 2k, 2k+1) -> one 16-byte k-chunk (8 cells) of this utterance's row
                const uint32_t plo = __shfl_xor_sync(0xffffffffu, lo, 1), phi = __shfl_xor_sync(0xffffffffu, hi, 1);
                const bool odd = a & 1;
                const uint32_t w0 = odd ? plo : lo, w1 = odd ? phi : hi, w2 = odd ? lo : plo, w3 = odd ? hi : phi;
                const uint32_t dst = sB(hc, buf ^ 1) + send_off;
                const uint32_t dbar = full(hc, buf ^ 1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < HD) st_async_v4(dst + rdelta[k], w0, w1, w2, w3, dbar + rdelta[k]);
            }
            PTRACE(tid == 0, t, 6);
            // off the critical path: operand of the hoisted projection GEMM and of the backward pass
            const size_t row = (size_t)t * p.B + b_own;
            if (b_own < p.B) {
                *reinterpret_cast<uint2*>(p.mt_seq + (row + p.B) * p.Cp + cell0) = make_uint2(lo, hi);
                if (p.save) {
                    float* s = p.save + row * 5 * p.Cp + cell0;
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                        *reinterpret_cast<float4*>(s + (size_t)k * p.Cp) = make_float4(sv[k][0], sv[k][1], sv[k][2], sv[k][3]);
                }
            }
            PTRACE(tid == 0, t, 7);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // nobody leaves (or frees TMEM the pair's MMAs address) while a peer may still use it
    if (warp == 8) tmem_dealloc_2cta(tmem, tcols);
}

size_t pfwd_smem(int Cp, int Ik, int nch) {
    const int nbc = NBP / nch, nbr = nbc / 2;
    const size_t xt = (size_t)((Ik + 63) / 64) * nbr * 128;
    const size_t need = 1024 + (size_t)nch * (2 * xt + 2 * (size_t)Cp * nbr * 2 + 128 * (size_t)(nbc + 1) * 4 + 64) + 64;
    return need < RSR_EXCLUSIVE_SMEM_REC ? RSR_EXCLUSIVE_SMEM_REC : need;
}

// chains per cluster: 1 (RSR_PAIR_CHAINS=2 selects the two-chain experiment)
int pair_chains() {
    const char* ev = getenv("RSR_PAIR_CHAINS");
    return (ev && ev[0] == '2') ? 2 : 1;
}

}  // namespace

// Returns 0 when launched, RSR_E_RESIDENT when the pair variant does not apply (the caller then tries the
// single-CTA cluster kernels of lstmp_cluster_sm100.cu).
int rsr_lstmp_fused_fwd_pair(rsr_handle* h, void* stream, int B, int T, int I, int Cp, const void* x16, int ldx,
                             const void* kxT, const float* bias, const void* wcT, const float* w_i,
                             const float* w_f, const float* w_o, float forget_bias, const int* lengths,
                             void* mt_seq, float* save) {
    if (Cp > 512) return RSR_E_RESIDENT;
    const int Ik = (I + 15) & ~15;
    if (Cp / 2 + Ik / 2 + 2 * NBP > 512) return RSR_E_RESIDENT;
    const int G = Cp / 32;
    const int nch = pair_chains();
    const size_t smem = pfwd_smem(Cp, Ik, nch);
    if (smem > (size_t)h->max_smem) return RSR_E_RESIDENT;
    int& cap = h->pair_cap[0][Cp / 256 - 1];
    if (cap < 0 || h->pair_ik[Cp / 256 - 1] != Ik * 4 + nch) {
        std::lock_guard<std::mutex> g(h->mu);
        cap = nch == 1 ? cluster_capacity(lstmp_fwd_pair_kernel<1>, G, PAIR_THREADS, smem)
                       : cluster_capacity(lstmp_fwd_pair_kernel<2>, G, PAIR_THREADS, smem);
        h->pair_ik[Cp / 256 - 1] = Ik * 4 + nch;
        if (getenv("RSR_DEBUG")) fprintf(stderr, "[rsr] fused fwd pair kernel Cp=%d Ik=%d: %d-CTA clusters co-resident: %d\n", Cp, Ik, G, cap);
    }
    if (cap <= 0) return RSR_E_RESIDENT;
    const int groups = (B + NBP - 1) / NBP;
    CUtensorMap tmX;
    int rc = rsr_get_tmap(h, x16, (uint64_t)ldx, (uint64_t)T * B, (uint64_t)ldx, 64, (uint32_t)(NBP / nch / 2), &tmX);
    if (rc) return rc;
    PFwdParams p;
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16; p.forget_bias = forget_bias;
    p.wcT = (const uint16_t*)wcT; p.kxT = (const uint16_t*)kxT; p.bias = bias;
    p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths;
    p.mt_seq = (uint16_t*)mt_seq; p.save = save; p.Ik = Ik;
    if (nch == 1) return cluster_launch(lstmp_fwd_pair_kernel<1>, groups, G, PAIR_THREADS, smem, (cudaStream_t)stream, tmX, p);
    return cluster_launch(lstmp_fwd_pair_kernel<2>, groups, G, PAIR_THREADS, smem, (cudaStream_t)stream, tmX, p);
}

// debug: copies the phase-timing trace of the pair kernels (all zeros unless built with -DRSR_TRACE) to the host
#ifdef RSR_TRACE
extern "C" int rsr_debug_trace_pair(unsigned long long* host_out, int n) {
    if (!host_out || n <= 0 || n > 8192) return RSR_E_ARG;
    RSR_CHECK_CUDA(cudaDeviceSynchronize());
    RSR_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_rsr_ptrace, sizeof(unsigned long long) * n));
    return 0;
}
#endif
