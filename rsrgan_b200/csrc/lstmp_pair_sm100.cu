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
#include <type_traits>

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

constexpr int NBP_BWD = 32;     // utterances per cluster of the backward kernel
constexpr int GATE_THREADS = 256;                 // backward kernel: gate-backward warps
constexpr int PAIR_THREADS = GATE_THREADS + 32;   // + the issuer warp

// Geometry of the forward kernel for NBP utterances per cluster (32: the default; 48: the layer-wavefront launch,
// which has to seat two layers' clusters plus the projection cluster in the 7 sixteen-CTA clusters this part places).
template <int NBP> struct PairGeom {
    static constexpr int NBR = NBP / 2;                 // B-operand rows (utterances) per CTA
    static constexpr int CW = NBP / 4;                  // gate warps: thread <-> 4 cells x 1 utterance
    static constexpr int GT = 32 * CW;                  // gate threads
    static constexpr int THREADS = GT + 64;             // + the issuer / relay warp + the poster warp (layer wavefront)
    static constexpr int ACCS = NBP <= 32 ? 32 : 64;    // TMEM column stride between the two accumulators
    static constexpr int XP = NBP + 1;                  // xchg pitch in floats
};

// MUFU.TANH forms of the gate non-linearities (default; RSR_FAST_GATES=0 selects the ex2 + rcp forms of common.cuh):
// one special-function instruction per gate instead of two and ~6 FP32 instructions.  Relative error 2^-11 -- the size
// of the rounding every 16-bit MMA operand already carries -- and no measurable effect on parity at the benchmarked
// lengths (profiles/r2_parity_measured_v0.jsonl: generator output RMS vs the float64 oracle at B = 128, T = 100
// 5.79e-5 with, 5.84e-5 without; gradients and post-schedule output likewise), for a gate phase of 582 instead of 976
// cycles per step (profiles/r2_pair_kernels_v6.txt)
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__device__ __forceinline__ void tc_mma_f16_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {   // non-blocking arrival
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 4-byte / 16-byte asynchronous global -> shared copies (LDGSTS): a prefetch that holds no registers.  src_bytes = 0
// zero-fills the destination without touching the source.
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float ld_shared_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

// named barrier the gate threads ARRIVE at and the poster warp syncs on, step by step: four ids in rotation, so that
// arrivals of a later step can only collide with an unconsumed earlier one if the poster fell four steps behind
__device__ __forceinline__ uint32_t post_bar_id(int step) { return 2u + (uint32_t)(step & 3) + ((step & 3) ? 2u : 0u); }   // 2, 5, 6, 7

// monotonic step counter of another cluster (layer wavefront): returns the value seen once it is >= want; bounded
__device__ __forceinline__ unsigned int spin_get_ge(const unsigned int* p, unsigned int want) {
    unsigned int v = ld_acquire(p);
    if (v >= want) return v;
    const uint64_t t0 = global_ns();
    uint32_t spins = 0;
    while ((v = ld_acquire(p)) < want) {
        if ((++spins & 1023u) == 0 && global_ns() - t0 > RSR_SPIN_LIMIT_NS) __trap();
    }
    return v;
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
    // layer wavefront (null / 0 otherwise): post[grp] += 1 per CTA once the CTA's rows of mt_t are in global memory;
    // the x tile of step t may be fetched once wait[grp] >= wait_per_step * (t + 1)
    unsigned int* post;
    const unsigned int* wait;
    unsigned int wait_per_step;
};

// One cluster (G = Cp/32 CTAs = G/2 pairs) = one group of NBP utterances: one pair MMA of N = NBP per k-step.
// CTA j owns cells [32j, 32j+32): its 128 packed gate rows of Wc^T and K_x^T are resident in its TMEM (A operand) for
// the whole sequence, and it holds the B-operand rows (mt_{t-1}, x_t) of NBR = NBP/2 utterances (the even CTA of a
// pair the first half, the odd CTA the second).
// Warps 0..CW-1: gate math (thread <-> 4 cells x 1 utterance); warp CW: in the even CTA of each pair the MMA issuer,
// in the odd CTA the relay that tells the issuer when the odd half of a B operand has landed; warp CW+1: the poster of
// the layer wavefront (idle otherwise).
// Tried and measured slower (profiles/r2_pair_fwd_variants_*.txt): shipping every 8-cell k-chunk as soon as its cell
// index is done, either with st.async from the gate warps (they stall on the 15 B/clk DSMEM port: 2766 cycles for gate
// math + sends instead of 927 + 1093) or with 256-byte bulk copies from a sender warp (per-copy overhead and a proxy
// fence per chunk: 5579 cycles per step instead of 3763); sharing reciprocals between gates (7 MUFU operations per cell
// instead of 10: the gate phase is latency-, not MUFU-bound); two N = 16 recurrences per cluster behind one issuer (they
// fall into lockstep: 3959 vs 3763 cycles per step).
template <int NBP, int BF, int FAST>
__device__ __forceinline__ void pair_fwd_body(const CUtensorMap* tmX, const PFwdParams& p, const int grp) {
    using GM = PairGeom<NBP>;
    constexpr int NBR = GM::NBR, CW = GM::CW, GT = GM::GT, XP = GM::XP;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = p.Cp / 32;                    // cluster size
    const uint32_t j = cluster_ctarank();       // cell block
    const uint32_t e = j & 1u;                  // which half of the group's utterances this CTA holds
    const int b0 = grp * NBP;

    const uint32_t sB_bytes = (uint32_t)p.Cp * NBR * 2u;             // one B-operand buffer [Cp/8][NBR][8] 16-bit
    const int KBX = (p.Ik + 63) / 64;                                // 64-wide k sub-tiles of the x_t tile
    const uint32_t xt_bytes = (uint32_t)KBX * NBR * 128u;            // this CTA's rows of one x_t tile, SW128 (1024-aligned)
    const uint32_t sX0 = base + 2u * xt_bytes + 2u * sB_bytes;       // float xchg[128][XP]
    const uint32_t sBar0 = (sX0 + 128u * XP * 4u + 7u) & ~7u;        // barM, full0/1, xfull0/1, peer0/1
    const uint32_t tslot = sBar0 + 64u;
    auto sXt = [&](int b) { return base + (uint32_t)b * xt_bytes; };
    auto sB = [&](int b) { return base + 2u * xt_bytes + (uint32_t)b * sB_bytes; };
    const uint32_t barM = sBar0;
    auto full = [&](int b) { return sBar0 + 8u + 8u * (uint32_t)b; };
    auto xfull = [&](int b) { return sBar0 + 24u + 8u * (uint32_t)b; };
    auto peer = [&](int b) { return sBar0 + 40u + 8u * (uint32_t)b; };

    // TMEM: [0, Cp/2) Wc^T slice | [Cp/2, Cp/2 + Ik/2) K_x^T slice | two accumulators (step parity), ACCS columns apart
    const uint32_t a_cols = (uint32_t)p.Cp / 2u + (uint32_t)p.Ik / 2u;
    uint32_t tcols = 32;
    while (tcols < a_cols + 2u * GM::ACCS) tcols <<= 1;
    if (tid == GT) {
        tma_prefetch_desc(tmX);
        mbar_init(barM, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(full(b), 1); mbar_init(xfull(b), 1); mbar_init(peer(b), 1); }
        fence_mbar_init();
        mbar_expect_tx(full(1), sB_bytes);      // armed for step 1 (mt_0 from every CTA of the cluster)
        mbar_expect_tx(full(0), sB_bytes);      // armed for step 2
    }
    if (warp == CW) tmem_alloc_2cta(tslot, tcols);
    if (tid < GT)                               // m_{-1} = 0
        for (uint32_t i = tid; i < sB_bytes / 16u; i += GT) st_shared_v4(sB(0) + i * 16u, 0u, 0u, 0u, 0u);
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

    if (warp == CW) {
        // ------------------------------------------------------------------------------------------------
        // issuer (even CTA) / relay (odd CTA): one elected lane
        // ------------------------------------------------------------------------------------------------
        if (elect_one_sync()) {
            // layer wavefront: the rows of x_t come from another cluster of this grid; `seen` caches the producer's counter,
            // so that a consumer that trails by several steps polls once per several steps
            unsigned int seen = 0;
            auto x_ready = [&](int t) {
                if (p.wait) {
                    const unsigned int need = p.wait_per_step * (unsigned int)(t + 1);
                    if (seen < need) { seen = spin_get_ge(p.wait + grp, need); fence_proxy_async_all(); }
                }
            };
            // x_0 tile; both CTAs' rows complete on the issuer's barrier
            x_ready(0);
            if (e == 0) mbar_expect_tx(xfull(0), 2u * xt_bytes);
            for (int kb = 0; kb < KBX; ++kb)
                tma_load_2d_2cta(sXt(0) + (uint32_t)kb * (NBR * 128u), tmX, xfull(0) + lead, kb * 64, b0 + (int)e * NBR);
            const uint32_t idesc = umma_idesc(256, NBP, BF, 0, 0);
            const uint16_t pair_mask = (uint16_t)(3u << (j & ~1u));
            for (int t = 0; t < p.T; ++t) {
                const int buf = t & 1;
                if (t + 1 < p.T) {   // prefetch x_{t+1}; its buffer was last read by the MMAs of step t-1 (retired: barM)
                    if (t > 0) mbar_wait(barM, (uint32_t)((t - 1) & 1));
                    x_ready(t + 1);
                    if (e == 0) mbar_expect_tx(xfull(buf ^ 1), 2u * xt_bytes);
                    for (int kb = 0; kb < KBX; ++kb)
                        tma_load_2d_2cta(sXt(buf ^ 1) + (uint32_t)kb * (NBR * 128u), tmX, xfull(buf ^ 1) + lead, kb * 64,
                                         (t + 1) * p.B + b0 + (int)e * NBR);
                }
                if (e == 0) {
                    const uint32_t acc = tmem_acc + (uint32_t)buf * GM::ACCS;
                    // input half first: it does not depend on mt_{t-1}, so it runs while the exchange is still in flight
                    mbar_wait(xfull(buf), (uint32_t)((t >> 1) & 1));
                    tc_fence_after();
                    const uint32_t sxt = sXt(buf);
                    for (int kk = 0; kk < KKX; ++kk) {
                        const uint64_t dx = umma_desc_sw128(sxt + (uint32_t)(kk >> 2) * (NBR * 128u) + (uint32_t)(kk & 3) * 32u, 16, 1024);
                        tc_mma_f16_ts_2cta(acc, tmem + (uint32_t)p.Cp / 2u + (uint32_t)kk * 8u, dx, idesc, kk ? 1u : 0u);
                    }
                    if (t > 0) {
                        const uint32_t ph = (uint32_t)(((t - 1) >> 1) & 1);
                        mbar_wait(full(buf), ph);            // my rows of mt_{t-1}: all G slices have landed
                        mbar_wait(peer(buf), ph);            // ... and the odd CTA's rows
                    }
                    PTRACE(true, t, 1);
                    fence_proxy_async_smem();
                    tc_fence_after();
                    uint64_t db = umma_desc_nosw(sB(buf), NBR * 16u, 128u);
                    uint32_t ta = tmem;
#pragma unroll 8
                    for (int kk = 0; kk < KK; ++kk) {
                        tc_mma_f16_ts_2cta(acc, ta, db, idesc, 1u);
                        ta += 8u;                                    // 16 k = 8 columns
                        db += (uint64_t)((2u * NBR * 16u) >> 4);     // two k-chunks of [NBR rows][16 B]
                    }
                    tc_commit_2cta_mc(barM, pair_mask);
                    if (t > 0 && t + 2 < p.T) mbar_expect_tx(full(buf), sB_bytes);   // re-arm for step t + 2
                    PTRACE(true, t, 2);
                } else if (t > 0) {
                    mbar_wait(full(buf), (uint32_t)(((t - 1) >> 1) & 1));
                    fence_proxy_async_smem();
                    mbar_arrive_cluster(peer(buf) + lead);
                    if (t + 2 < p.T) mbar_expect_tx(full(buf), sB_bytes);
                }
            }
        }
        __syncwarp();
    } else if (warp == CW + 1) {
        // poster (layer wavefront): once every gate thread has stored its piece of mt_t, one release per CTA and step
        if (p.post)
            for (int t = 0; t < p.T; ++t) {
                named_bar_sync(post_bar_id(t), GT + 32);
                if (lane == 0) red_release_add(p.post + grp, 1u);
            }
    } else {
        // ------------------------------------------------------------------------------------------------
        // gate warps: thread <-> (cells 4a..4a+3 of the block, utterance nl of the group)
        // ------------------------------------------------------------------------------------------------
        const int q = warp & 3, chh = warp >> 2;    // TMEM lane quadrant, 16-column piece of the accumulator
        const int a = lane & 7;
        const int nl = 4 * warp + (lane >> 3);      // utterance within the group
        const int cell0 = 32 * (int)j + 4 * a;
        const float4 wi4 = *reinterpret_cast<const float4*>(p.w_i + cell0);
        const float4 wf4 = *reinterpret_cast<const float4*>(p.w_f + cell0);
        const float4 wo4 = *reinterpret_cast<const float4*>(p.w_o + cell0);
        const float wi[4] = {wi4.x, wi4.y, wi4.z, wi4.w}, wf[4] = {wf4.x, wf4.y, wf4.z, wf4.w},
                    wo[4] = {wo4.x, wo4.y, wo4.z, wo4.w};
        float creg[4] = {0.f, 0.f, 0.f, 0.f};
        const int b_own = b0 + nl;
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
        float* xchg = reinterpret_cast<float*>(base_ptr + (sX0 - base));
        const uint32_t send_off = (uint32_t)(4 * j + (a >> 1)) * (NBR * 16u) + (uint32_t)(nl % NBR) * 16u;

        for (int t = 0; t < p.T; ++t) {
            const int buf = t & 1;
            PTRACE(tid == 0, t, 0);
            mbar_wait(barM, (uint32_t)(t & 1));
            tc_fence_after();
            PTRACE(tid == 0, t, 3);
            float acc[16];
            tmem_ld16(tmem_acc + (uint32_t)buf * GM::ACCS + ((uint32_t)(q * 32) << 16) + (uint32_t)chh * 16u, acc);
#pragma unroll
            for (int k = 0; k < 16; ++k) xchg[grow * XP + chh * 16 + k] = acc[k] + bias_r;
            tc_fence_before();
            named_bar_sync(1u, GT);
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
                const float ig = FAST ? sigmoid_fast(zi + wi[c] * cp) : sigmoidf_(zi + wi[c] * cp);
                const float fg = FAST ? sigmoid_fast(zf + p.forget_bias + wf[c] * cp) : sigmoidf_(zf + p.forget_bias + wf[c] * cp);
                const float jg = FAST ? tanh_fast(zj) : tanhf_(zj);
                const float cn = fg * cp + ig * jg;
                const float og = FAST ? sigmoid_fast(zo + wo[c] * cn) : sigmoidf_(zo + wo[c] * cn);
                const float mt = og * (FAST ? tanh_fast(cn) : tanhf_(cn));
                sv[0][c] = ig; sv[1][c] = fg; sv[2][c] = og; sv[3][c] = jg; sv[4][c] = cn;
                mtv[c] = active ? mt : 0.f;
                if (active) creg[c] = cn;
            }
            const uint32_t lo = pack2(mtv[0], mtv[1], BF), hi = pack2(mtv[2], mtv[3], BF);
            PTRACE(tid == 0, t, 5);
            // (the exchange buffer is rewritten at step t+1 only after the NEXT barM, which needs every warp's sends below)
            if (t + 1 < p.T) {
                // lanes (a = 2k, 2k+1) -> one 16-byte k-chunk (8 cells) of this utterance's row
                const uint32_t plo = __shfl_xor_sync(0xffffffffu, lo, 1), phi = __shfl_xor_sync(0xffffffffu, hi, 1);
                const bool odd = a & 1;
                const uint32_t w0 = odd ? plo : lo, w1 = odd ? phi : hi, w2 = odd ? lo : plo, w3 = odd ? hi : phi;
                const uint32_t dst = sB(buf ^ 1) + send_off;
                const uint32_t dbar = full(buf ^ 1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < HD) st_async_v4(dst + rdelta[k], w0, w1, w2, w3, dbar + rdelta[k]);
            }
            PTRACE(tid == 0, t, 6);
            // off the critical path: operand of the hoisted projection GEMM and of the backward pass
            const size_t row = (size_t)t * p.B + b_own;
            if (b_own < p.B) *reinterpret_cast<uint2*>(p.mt_seq + (row + p.B) * p.Cp + cell0) = make_uint2(lo, hi);
            // layer wavefront: this CTA's rows of mt_t are on their way; the poster warp releases them to the consumer (the
            // gate warps only ARRIVE: a release fence here costs them ~0.5 us per step, profiles/r2_wave_steps_v0.txt)
            if (p.post) named_bar_arrive(post_bar_id(t), GT + 32);
            if (b_own < p.B && p.save) {
                float* s = p.save + row * 5 * p.Cp + cell0;
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    *reinterpret_cast<float4*>(s + (size_t)k * p.Cp) = make_float4(sv[k][0], sv[k][1], sv[k][2], sv[k][3]);
            }
            PTRACE(tid == 0, t, 7);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // nobody leaves (or frees TMEM the pair's MMAs address) while a peer may still use it
    if (warp == CW) tmem_dealloc_2cta(tmem, tcols);
}

template <int NBP, int BF, int FAST>
__global__ void __launch_bounds__(PairGeom<NBP>::THREADS, 1)
lstmp_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmX, const PFwdParams p) {
    pair_fwd_body<NBP, BF, FAST>(&tmX, p, (int)blockIdx.x / (p.Cp / 32));
}

// =========================================================================================
// layer wavefront: two stacked LSTMP layers in ONE launch (models/lstm.py:89-112 builds them as a MultiRNNCell, i.e.
// layer 2 consumes layer 1's output of the same time step).  Clusters [0, groups) run layer 1, clusters
// [groups, 2 groups) layer 2 -- one to a few steps behind, gated by per-group counters in global memory -- and the
// last cluster streams the projection out1_t = mt1_t W_p1 between them (a few of its CTAs; the others exit at once):
//   layer 1, step t : mt1_t -> global, post flag1[g]            (16 CTAs -> 16 per step)
//   projection      : wait flag1[g] >= 16 (t+1); TMA mt1_t (B operand, N = NBP), A = W_p1^T slice resident in TMEM
//                     (128 features per CTA), D -> 16-bit rows of out1_t -> global, post flag2[g]
//   layer 2, step t : its x tile is out1_t: wait flag2[g] >= NFT (t+1) before the TMA
// Nobody waits on a LATER cluster of the grid, clusters are dispatched in index order, and the host launches only if
// all 2 groups + 1 clusters are co-resident -- so the spins cannot deadlock; they are bounded anyway (RSR_SPIN_LIMIT_NS).
// =========================================================================================
struct WaveProj {
    int groups, P, Pp, ldo;
    const uint16_t* wpT;        // [Pp, Cp] = W_p1^T (16-bit, rows >= P zero)
    uint16_t* out;              // [(T+1)*B, ldo] layer-1 output sequence (slot 0 = initial state, not written)
    unsigned int* flag1;        // [groups] posted by layer 1
    unsigned int* flag2;        // [groups] posted by the projection
};

template <int NBP, int BF>
__device__ __forceinline__ void wave_proj_body(const CUtensorMap* tmM, const PFwdParams& p, const WaveProj& w) {
    constexpr int ACCS = PairGeom<NBP>::ACCS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NFT = (w.Pp + 127) / 128;
    const int c = (int)cluster_ctarank();
    if (c >= w.groups * NFT) return;            // whole CTA: nothing of this cluster is shared
    const int g = c / NFT, ft = c % NFT;
    const int b0 = g * NBP;
    const int KB = p.Cp / 64;
    const uint32_t bt_bytes = (uint32_t)KB * NBP * 128u;             // one mt1_t tile: KB sub-tiles [NBP rows x 64 k], SW128
    const uint32_t sBar = base + 2u * bt_bytes;
    auto sBt = [&](int b) { return base + (uint32_t)b * bt_bytes; };
    auto fullB = [&](int b) { return sBar + 8u * (uint32_t)b; };
    auto emptyB = [&](int b) { return sBar + 16u + 8u * (uint32_t)b; };
    auto accfull = [&](int b) { return sBar + 32u + 8u * (uint32_t)b; };
    auto accfree = [&](int b) { return sBar + 48u + 8u * (uint32_t)b; };
    const uint32_t tslot = sBar + 64u;
    const uint32_t a_cols = (uint32_t)p.Cp / 2u;
    uint32_t tcols = 32;
    while (tcols < a_cols + 2u * ACCS) tcols <<= 1;
    if (tid == 0) {
        tma_prefetch_desc(tmM);
        for (int b = 0; b < 2; ++b) { mbar_init(fullB(b), 1); mbar_init(emptyB(b), 1); mbar_init(accfull(b), 1); mbar_init(accfree(b), 1); }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tslot, tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    const uint32_t tmem_acc = tmem + a_cols;
    if (warp < 4) {   // W_p1^T slice -> TMEM: thread <-> feature row, 64 k (32 columns) per store
        const int f = 128 * ft + 32 * warp + lane;
        const uint16_t* wrow = w.wpT + (size_t)f * p.Cp;
        for (int cb = 0; cb < KB; ++cb) {
            uint32_t r[32];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (f < w.Pp) v = __ldg(reinterpret_cast<const uint4*>(wrow + cb * 64) + k);
                r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
            }
            tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb * 32u, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp >= 6) return;                      // (only mbarriers and named barriers 3, 4 below)

    const int G = p.Cp / 32;
    if (warp == 0) {
        if (elect_one_sync()) {                 // producer: mt1_t tiles as layer 1 releases them
            unsigned int seen = 0;
            for (int t = 0; t < p.T; ++t) {
                const int buf = t & 1;
                if (t >= 2) mbar_wait(emptyB(buf), (uint32_t)(((t - 2) >> 1) & 1));
                const unsigned int need = (unsigned int)G * (unsigned int)(t + 1);
                if (seen < need) { seen = spin_get_ge(w.flag1 + g, need); fence_proxy_async_all(); }
                mbar_expect_tx(fullB(buf), bt_bytes);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(sBt(buf) + (uint32_t)kb * (NBP * 128u), tmM, fullB(buf), kb * 64, (t + 1) * p.B + b0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one_sync()) {                 // D[128 features, NBP utterances] = W_p1^T (TMEM) x mt1_t^T
            const uint32_t idesc = umma_idesc(128, NBP, BF, 0, 0);
            for (int t = 0; t < p.T; ++t) {
                const int buf = t & 1;
                mbar_wait(fullB(buf), (uint32_t)((t >> 1) & 1));
                if (t >= 2) mbar_wait(accfree(buf), (uint32_t)(((t - 2) >> 1) & 1));
                tc_fence_after();
                const uint32_t sb = sBt(buf);
                for (int kk = 0; kk < p.Cp / 16; ++kk) {
                    const uint64_t db = umma_desc_sw128(sb + (uint32_t)(kk >> 2) * (NBP * 128u) + (uint32_t)(kk & 3) * 32u, 16, 1024);
                    tc_mma_f16_ts(tmem_acc + (uint32_t)buf * ACCS, tmem + (uint32_t)kk * 8u, db, idesc, kk ? 1u : 0u);
                }
                tc_commit(emptyB(buf));
                tc_commit(accfull(buf));
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                 // warps 2..5 <-> TMEM lane quadrants 2, 3, 0, 1
        const int f = 128 * ft + 32 * q + lane;
        for (int t = 0; t < p.T; ++t) {
            const int buf = t & 1;
            mbar_wait(accfull(buf), (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            uint16_t* orow = w.out + ((size_t)(t + 1) * p.B + b0) * w.ldo + f;
#pragma unroll
            for (int c16 = 0; c16 < NBP / 16; ++c16) {
                float acc[16];
                tmem_ld16(tmem_acc + (uint32_t)buf * ACCS + ((uint32_t)(q * 32) << 16) + (uint32_t)c16 * 16u, acc);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int n = c16 * 16 + k;
                    if (b0 + n < p.B && f < w.Pp) orow[(size_t)n * w.ldo] = f2h(acc[k], BF);
                }
            }
            tc_fence_before();
            named_bar_sync(3u, 128);
            if (warp == 2 && lane == 0) {
                mbar_arrive(accfree(buf));
                red_release_add(w.flag2 + g, 1u);
            }
        }
    }
    named_bar_sync(4u, 192);
    if (warp == 0) tmem_dealloc(tmem, tcols);
}

template <int NBP, int BF, int FAST>
__global__ void __launch_bounds__(PairGeom<NBP>::THREADS, 1)
lstmp_wave_fwd_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                      const __grid_constant__ CUtensorMap tmM, const PFwdParams p1, const PFwdParams p2, const WaveProj w) {
    const int cid = (int)blockIdx.x / (p1.Cp / 32);
    if (cid < w.groups) pair_fwd_body<NBP, BF, FAST>(&tmX1, p1, cid);
    else if (cid < 2 * w.groups) pair_fwd_body<NBP, BF, FAST>(&tmX2, p2, cid - w.groups);
    else wave_proj_body<NBP, BF>(&tmM, p1, w);
}

size_t pfwd_smem(int Cp, int Ik, int nbp) {
    const int nbr = nbp / 2;
    const size_t xt = (size_t)((Ik + 63) / 64) * nbr * 128;
    const size_t need = 1024 + 2 * xt + 2 * (size_t)Cp * nbr * 2 + 128 * (size_t)(nbp + 1) * 4 + 8 + 64 + 16;
    return need < RSR_EXCLUSIVE_SMEM_REC ? RSR_EXCLUSIVE_SMEM_REC : need;
}
size_t wproj_smem(int Cp, int nbp, int stages = 2) { return 1024 + (size_t)stages * (Cp / 64) * nbp * 128 + 128 + 16; }

bool fast_gates() {
    static const int fast = getenv("RSR_FAST_GATES") ? (atoi(getenv("RSR_FAST_GATES")) ? 1 : 0) : 1;
    return fast != 0;
}

void fill_fwd_params(PFwdParams& p, rsr_handle* h, int B, int T, int Cp, int Ik, const void* kxT, const float* bias,
                     const void* wcT, const float* w_i, const float* w_f, const float* w_o, float forget_bias,
                     const int* lengths, void* mt_seq, float* save) {
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16; p.forget_bias = forget_bias;
    p.wcT = (const uint16_t*)wcT; p.kxT = (const uint16_t*)kxT; p.bias = bias;
    p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths;
    p.mt_seq = (uint16_t*)mt_seq; p.save = save; p.Ik = Ik;
    p.post = nullptr; p.wait = nullptr; p.wait_per_step = 0;
}

}  // namespace

// Returns 0 when launched, RSR_E_RESIDENT when the pair variant does not apply (the caller then tries the
// single-CTA cluster kernels of lstmp_cluster_sm100.cu).
int rsr_lstmp_fused_fwd_pair(rsr_handle* h, void* stream, int B, int T, int I, int Cp, const void* x16, int ldx,
                             const void* kxT, const float* bias, const void* wcT, const float* w_i,
                             const float* w_f, const float* w_o, float forget_bias, const int* lengths,
                             void* mt_seq, float* save) {
    constexpr int NBP = 32;
    if (Cp > 512) return RSR_E_RESIDENT;
    const int Ik = (I + 15) & ~15;
    if (Cp / 2 + Ik / 2 + 2 * PairGeom<NBP>::ACCS > 512) return RSR_E_RESIDENT;
    const int G = Cp / 32;
    const size_t smem = pfwd_smem(Cp, Ik, NBP);
    if (smem > (size_t)h->max_smem) return RSR_E_RESIDENT;
    const int bf = h->dtype == RSR_DTYPE_BF16;
    const int fast = fast_gates();
    auto run = [&](auto kernel) -> int {
        int& cap = h->pair_cap[0][Cp / 256 - 1];
        const int key = (Ik * 2 + bf) * 2 + fast;
        if (cap < 0 || h->pair_ik[Cp / 256 - 1] != key) {
            std::lock_guard<std::mutex> g(h->mu);
            cap = cluster_capacity(kernel, G, PairGeom<NBP>::THREADS, smem);
            h->pair_ik[Cp / 256 - 1] = key;
            if (getenv("RSR_DEBUG")) fprintf(stderr, "[rsr] fused fwd pair kernel Cp=%d Ik=%d: %d-CTA clusters co-resident: %d\n", Cp, Ik, G, cap);
        }
        if (cap <= 0) return RSR_E_RESIDENT;
        const int groups = (B + NBP - 1) / NBP;
        CUtensorMap tmX;
        int rc = rsr_get_tmap(h, x16, (uint64_t)ldx, (uint64_t)T * B, (uint64_t)ldx, 64, (uint32_t)(NBP / 2), &tmX);
        if (rc) return rc;
        PFwdParams p;
        fill_fwd_params(p, h, B, T, Cp, Ik, kxT, bias, wcT, w_i, w_f, w_o, forget_bias, lengths, mt_seq, save);
        return cluster_launch(kernel, groups, G, PairGeom<NBP>::THREADS, smem, (cudaStream_t)stream, tmX, p);
    };
    if (fast) return bf ? run(lstmp_fwd_pair_kernel<NBP, 1, 1>) : run(lstmp_fwd_pair_kernel<NBP, 0, 1>);
    return bf ? run(lstmp_fwd_pair_kernel<NBP, 1, 0>) : run(lstmp_fwd_pair_kernel<NBP, 0, 0>);
}

// Two stacked LSTMP layers of equal cell count as one wavefront launch (see lstmp_wave_fwd_kernel).  Layer 2's input
// is layer 1's projected output: I2 = P1, x2 rows = out1 rows.  Returns RSR_E_RESIDENT when the shape does not fit
// (the caller then runs the layers one after the other).
extern "C" int rsr_lstmp_wave_fwd(rsr_handle* h, void* stream, const rsr_wave_args* a) {
    if (!h || !a) return RSR_E_ARG;
    const int B = a->B, T = a->T, Cp = a->Cp, I1 = a->I1, P1 = a->P1;
    if (!a->x16 || !a->kxT1 || !a->bias1 || !a->wcT1 || !a->w_i1 || !a->w_f1 || !a->w_o1 || !a->mt1 || !a->wpT1 || !a->out1 ||
        !a->kxT2 || !a->bias2 || !a->wcT2 || !a->w_i2 || !a->w_f2 || !a->w_o2 || !a->mt2 || !a->lengths)
        return RSR_E_ARG;
    if (B <= 0 || T <= 0 || I1 <= 0 || P1 <= 0 || Cp <= 0 || (Cp & 255) || a->ldx < I1 || (a->ldx & 7) || a->ldo1 < P1 || (a->ldo1 & 7))
        return RSR_E_SHAPE;
    if (((uintptr_t)a->x16 | (uintptr_t)a->kxT1 | (uintptr_t)a->wcT1 | (uintptr_t)a->mt1 | (uintptr_t)a->save1 | (uintptr_t)a->wpT1 |
         (uintptr_t)a->out1 | (uintptr_t)a->kxT2 | (uintptr_t)a->wcT2 | (uintptr_t)a->mt2 | (uintptr_t)a->save2) & 15)
        return RSR_E_ARG;
    if (getenv("RSR_NO_CLUSTER") || getenv("RSR_NO_PAIR") || getenv("RSR_NO_WAVE") || Cp > 512) return RSR_E_RESIDENT;
    const int Ik1 = (I1 + 15) & ~15, Ik2 = (P1 + 15) & ~15, Ikm = Ik1 > Ik2 ? Ik1 : Ik2;
    const int Pp = (P1 + 7) & ~7;
    const int G = Cp / 32, NFT = (Pp + 127) / 128;
    const int bf = h->dtype == RSR_DTYPE_BF16, fast = fast_gates();
    auto run = [&](auto kernel, auto nbp_tag) -> int {
        constexpr int NBP = decltype(nbp_tag)::value;
        using GM = PairGeom<NBP>;
        const int groups = (B + NBP - 1) / NBP;
        if (Cp / 2 + Ikm / 2 + 2 * GM::ACCS > 512 || groups * NFT > G) return RSR_E_RESIDENT;
        size_t smem = pfwd_smem(Cp, Ikm, NBP);
        if (wproj_smem(Cp, NBP) > smem) smem = wproj_smem(Cp, NBP);
        if (smem > (size_t)h->max_smem) return RSR_E_RESIDENT;
        int cap;
        {
            std::lock_guard<std::mutex> g(h->mu);
            int& c = h->wave_cap[NBP == 32 ? 0 : 1][Cp / 256 - 1];
            const int key = ((Ikm * 2 + bf) * 2 + fast);
            if (c < 0 || h->wave_key[NBP == 32 ? 0 : 1][Cp / 256 - 1] != key) {
                c = cluster_capacity(kernel, G, GM::THREADS, smem);
                h->wave_key[NBP == 32 ? 0 : 1][Cp / 256 - 1] = key;
                if (getenv("RSR_DEBUG")) fprintf(stderr, "[rsr] wave fwd kernel Cp=%d NBP=%d: %d-CTA clusters co-resident: %d\n", Cp, NBP, G, c);
            }
            cap = c;
        }
        if (cap < 2 * groups + 1) return RSR_E_RESIDENT;
        CUtensorMap tmX1, tmX2, tmM;
        int rc = rsr_get_tmap(h, a->x16, (uint64_t)a->ldx, (uint64_t)T * B, (uint64_t)a->ldx, 64, (uint32_t)(NBP / 2), &tmX1);
        if (rc) return rc;
        // layer 2 reads rows [B, (T+1) B) of out1: its step t is out1 row (t+1) B + b
        rc = rsr_get_tmap(h, (const uint16_t*)a->out1 + (size_t)B * a->ldo1, (uint64_t)a->ldo1, (uint64_t)T * B, (uint64_t)a->ldo1, 64,
                          (uint32_t)(NBP / 2), &tmX2);
        if (rc) return rc;
        rc = rsr_get_tmap(h, a->mt1, (uint64_t)Cp, (uint64_t)(T + 1) * B, (uint64_t)Cp, 64, (uint32_t)NBP, &tmM);
        if (rc) return rc;
        unsigned int* flags = rsr_take_flags(h, 2 * groups);
        { const int rz = rsr_zero_u32(flags, 2 * groups, (cudaStream_t)stream); if (rz) return rz; }
        PFwdParams p1, p2;
        fill_fwd_params(p1, h, B, T, Cp, Ik1, a->kxT1, a->bias1, a->wcT1, a->w_i1, a->w_f1, a->w_o1, a->forget_bias, a->lengths, a->mt1, a->save1);
        fill_fwd_params(p2, h, B, T, Cp, Ik2, a->kxT2, a->bias2, a->wcT2, a->w_i2, a->w_f2, a->w_o2, a->forget_bias, a->lengths, a->mt2, a->save2);
        p1.post = flags;
        p2.wait = flags + groups; p2.wait_per_step = (unsigned int)NFT;
        WaveProj w;
        w.groups = groups; w.P = P1; w.Pp = Pp; w.ldo = a->ldo1; w.wpT = (const uint16_t*)a->wpT1; w.out = (uint16_t*)a->out1;
        w.flag1 = flags; w.flag2 = flags + groups;
        return cluster_launch(kernel, 2 * groups + 1, G, GM::THREADS, smem, (cudaStream_t)stream, tmX1, tmX2, tmM, p1, p2, w);
    };
    auto pick = [&](auto nbp_tag) -> int {
        constexpr int NBP = decltype(nbp_tag)::value;
        if (fast) return bf ? run(lstmp_wave_fwd_kernel<NBP, 1, 1>, nbp_tag) : run(lstmp_wave_fwd_kernel<NBP, 0, 1>, nbp_tag);
        return bf ? run(lstmp_wave_fwd_kernel<NBP, 1, 0>, nbp_tag) : run(lstmp_wave_fwd_kernel<NBP, 0, 0>, nbp_tag);
    };
    // 32 utterances per cluster when the batch seats that way (the faster step), else 48
    const int force = getenv("RSR_WAVE_NBP") ? atoi(getenv("RSR_WAVE_NBP")) : 0;
    if (force != 48) {
        const int rc = pick(std::integral_constant<int, 32>());
        if (rc != RSR_E_RESIDENT || force == 32) return rc;
    }
    return pick(std::integral_constant<int, 48>());
}

// =========================================================================================
// backward, CTA-pair variant
// =========================================================================================
namespace {

template <int NBP> struct PairGeomB {
    static constexpr int NBR = NBP / 2;                 // utterances this CTA differentiates (B-operand rows it produces)
    static constexpr int GW = NBR / 2;                  // gate-backward warps: thread <-> one cell x 4 utterances
    static constexpr int GT = 32 * GW;                  // 64 cells x NBR / 4 utterance quads
    static constexpr int THREADS = GT + 64;             // + the issuer warp + the poster warp (layer wavefront)
    static constexpr int ACCS = NBP <= 32 ? 32 : 64;    // TMEM column stride between the output tiles' accumulators
    static constexpr uint32_t SLOT = NBR * 128u;        // bytes one source pair sends per step: [NBR/8][64 cells][8 utt] 16-bit
};

struct PBwdParams {
    int B, T, Cp, bf;
    const float* dmt;           // [T*B, Cp] dOut W_p^T, read only -- or, with `grouped`,
    int grouped;                // layer wavefront: [T][groups][Cp][NBP] (utterances of a group contiguous), accumulated by the
    int groups;                 //   producer stage; every element is reset to 0 once read
    const uint16_t* wc;         // [Cp, 4Cp] packed gate columns
    const float* w_i; const float* w_f; const float* w_o;
    const int* lengths;
    const float* save;          // [T*B, 5, Cp]
    uint16_t* dz16;             // [T*B, 4Cp] packed
    float* dbias; float* dw_i; float* dw_f; float* dw_o;
    // layer wavefront (null / 0 otherwise): post[grp] += 1 per CTA once its piece of dz_t is in global memory; the dmt rows
    // of step s (t = T-1-s) may be read once wait[grp] >= wait_per_step * (s + 1)
    unsigned int* post;
    const unsigned int* wait;
    unsigned int wait_per_step;
};

// dmt_{t-1} = dOut_{t-1} W_p^T + dz_t Wc^T, split along K over the G/2 PAIRS of the cluster: pair p owns the four gates
// of cells [64p, 64p+64) (256 packed gate columns) for the NBP utterances of the group -- CTA (p, e) computes dz_t of
// those cells for utterances [NBR e, NBR e + NBR) (no exchange of the B operand: its rows of the pair's B tile are
// produced locally), the pair runs ONE tcgen05.mma.cta_group::2 per k-step (M = 256 output cells: 128 per CTA, N = NBP,
// K = 256; Wc slices resident in TMEM) and every CTA reduce-scatters its partial rows, 16-bit, to the CTA that owns (cell
// block, utterance half).  Per SM and step 16 KB leave for 32 utterances (the single-CTA kernel ships 16 KB per 16), and
// one CTA issues the MMAs of two.  Warps 0..GW-1: gate backward (thread <-> one cell x 4 utterances) and the sends; warp
// GW: MMA issuer (even CTA); warp GW+1: in the layer wavefront, the poster of this CTA's dz_t.
template <int NBP, int BF, int FAST>
__device__ __forceinline__ void pair_bwd_body(const PBwdParams& p, const int grp) {
    using GM = PairGeomB<NBP>;
    constexpr int NBR = GM::NBR, GW = GM::GW, GT = GM::GT, ACCS = GM::ACCS;
    constexpr uint32_t SLOT = GM::SLOT;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = p.Cp / 32, NP = G / 2;        // cluster size, pairs
    const int MT2 = p.Cp / 256;                 // 256-row output tiles of the pair product: 1 or 2
    const uint32_t j = cluster_ctarank();
    const uint32_t e = j & 1u, pr = j >> 1;     // utterance half, pair index
    const int b0 = grp * NBP + (int)e * NBR;    // first utterance this CTA differentiates

    const uint32_t sR_bytes = (uint32_t)NP * SLOT;
    const uint32_t sBt = base;                                       // dz tile: 4 k-subtiles [NBR rows x 64 k] 16-bit, SW128
    const uint32_t sR0 = sBt + 4u * NBR * 128u;                      // two receive buffers
    const uint32_t sBar = sR0 + 2u * sR_bytes;
    const uint32_t barM = sBar, full0 = sBar + 8, full1 = sBar + 16, dzr = sBar + 24;
    const uint32_t tslot = sBar + 32;
    // 48 utterances per cluster: the operands of the next step are prefetched into shared memory, not registers (below)
    const uint32_t sPre = sBar + 64;                                 // float pre[20][GT] | float4 predm[GT]
    const uint32_t sPreDm = sPre + 20u * GT * 4u;
    uint8_t* sB_ptr = base_ptr + (sBt - base);

    // TMEM: tile mt of the A operand (rows = output cells 256 mt + 128 e + lane, K = the pair's 256 gate columns) at columns
    //       [128 mt, 128 mt + 128); accumulator of tile mt at columns 128 MT2 + ACCS mt
    const uint32_t a_cols = 128u * (uint32_t)MT2;
    uint32_t tcols = 32;
    while (tcols < a_cols + (uint32_t)(MT2 * ACCS)) tcols <<= 1;
    if (tid == GT) {
        mbar_init(barM, 1); mbar_init(full0, 1); mbar_init(full1, 1); mbar_init(dzr, 2);
        fence_mbar_init();
        mbar_expect_tx(full1, sR_bytes);
        mbar_expect_tx(full0, sR_bytes);
    }
    if (warp == GW) tmem_alloc_2cta(tslot, tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    const uint32_t tmem_acc = tmem + a_cols;

    if (warp < 8) {   // weight slab -> TMEM: thread <-> output row, 64 k (32 columns) per store
        const int q = warp & 3;
        for (int it = warp >> 2; it < 4 * MT2; it += 2) {
            const int mt = it >> 2, kq = it & 3;
            const uint16_t* wrow = p.wc + (size_t)(256 * mt + 128 * (int)e + 32 * q + lane) * 4 * p.Cp + 256 * pr + 64 * kq;
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow) + c);
                r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
            }
            tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(128 * mt + 32 * kq), r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();

    const uint32_t lead = mapa_u32(base, j & ~1u) - base;

    if (warp == GW) {
        if (e == 0 && elect_one_sync()) {
            const uint32_t idesc = umma_idesc(256, NBP, BF, 0, 0);
            const uint16_t pair_mask = (uint16_t)(3u << (j & ~1u));
            for (int step = 0; step + 1 < p.T; ++step) {
                mbar_wait(dzr, (uint32_t)(step & 1));      // both CTAs' rows of the dz tile are written
                tc_fence_after();
                for (int mt = 0; mt < MT2; ++mt) {
#pragma unroll
                    for (int kk = 0; kk < 16; ++kk) {
                        const uint64_t db = umma_desc_sw128(sBt + (uint32_t)(kk >> 2) * (NBR * 128u) + (uint32_t)(kk & 3) * 32u, 16, 1024);
                        tc_mma_f16_ts_2cta(tmem_acc + (uint32_t)(mt * ACCS), tmem + (uint32_t)(128 * mt + 8 * kk), db, idesc, kk ? 1u : 0u);
                    }
                }
                tc_commit_2cta_mc(barM, pair_mask);
                PTRACE(true, step, 5);
            }
        }
        __syncwarp();
    } else if (warp == GW + 1) {
        // poster (layer wavefront): every gate thread has stored its piece of dz_t (they only arrive); one release per CTA
        // and step -- the fence behind it waits for those scattered 2-byte stores, which is why no busy warp does it
        if (p.post)
            for (int step = 0; step < p.T; ++step) {
                named_bar_sync(post_bar_id(step), GT + 32);
                if (lane == 0) red_release_add(p.post + grp, 1u);
            }
    } else {
        // gate-backward ownership: thread <-> (cell 64 pr + cc, utterances 4 uq + u of this CTA's NBR); a warp holds 16
        // cells x 2 utterance quads so that its reads of a received slot are 256 contiguous bytes (no bank conflicts)
        constexpr int UPT = 4;
        const int cc = 16 * (warp & 3) + (lane & 15), uq = 2 * (warp >> 2) + (lane >> 4);
        const int cl = cc & 31;                     // column of this cell inside a 32-wide gate block
        const int cell = 64 * (int)pr + cc;
        const float wi = p.w_i[cell], wf = p.w_f[cell], wo = p.w_o[cell];
        float dcar[UPT];
        int len[UPT];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            dcar[u] = 0.f;
            const int b = b0 + UPT * uq + u;
            len[u] = b < p.B ? p.lengths[b] : 0;
        }
        float a_dwi = 0.f, a_dwf = 0.f, a_dwo = 0.f, a_db[4] = {0.f, 0.f, 0.f, 0.f};
        // sends: this thread drains accumulator rows 32 q + lane, columns [16 piece, +16) of every tile; row 256 mt + 128 e +
        // 32 q + lane is cell 32 (q & 1) + lane of pair 4 mt + 2 e + q / 2; each 8-column unit (8 utterances) goes to the CTA
        // of that pair which holds the unit's utterance half
        const int q = warp & 3, piece = warp >> 2;
        uint32_t rd[2][2], soff[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int u8 = 2 * piece + c, par = u8 / (NBR / 8), hh = u8 % (NBR / 8);
            soff[c] = pr * SLOT + (uint32_t)hh * 1024u + (uint32_t)(32 * (q & 1) + lane) * 16u;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                rd[mt][c] = mt < MT2 ? mapa_u32(base, (uint32_t)(2 * (4 * mt + 2 * (int)e + (q >> 1)) + par)) - base : 0u;
        }
        const size_t Cp = (size_t)p.Cp;
        // my 4 utterances inside the 16-byte chunk [h][cell][8 utterances] of a source slot
        const uint32_t roff = (uint32_t)(uq >> 1) * 1024u + (uint32_t)cc * 16u + (uint32_t)(uq & 1) * 8u;
        // dz tile: row n, pair-local packed gate column 128 (cc / 32) + 32 g + cc % 32 -> k-subtile 2 (cc / 32) + g / 2
        const uint32_t zoff = (uint32_t)(2 * (cc >> 5)) * (NBR * 128u);
        uint16_t* dzg = p.dz16 + 256 * pr + 128 * (cc >> 5) + (cc & 31);

        // layer wavefront: the dmt rows come from another cluster of this grid
        unsigned int seen = 0;
        auto dm_ready = [&](int s) {
            if (p.wait) {
                const unsigned int need = p.wait_per_step * (unsigned int)(s + 1);
                if (seen < need) {
                    unsigned int v = 0;
                    if (lane == 0) v = spin_get_ge(p.wait + grp, need);
                    seen = __shfl_sync(0xffffffffu, v, 0);
                }
            }
        };
        // dmt of time step t for this thread's four utterances
        auto load_dm4 = [&](int t, float* out) {
            if (p.grouped) {
                float4* a = reinterpret_cast<float4*>(const_cast<float*>(p.dmt)) +
                            ((((size_t)t * p.groups + grp) * Cp + cell) * NBP + e * NBR + UPT * uq) / 4;
                const float4 v = __ldcg(a);
                __stcg(a, make_float4(0.f, 0.f, 0.f, 0.f));
                out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
            } else {
#pragma unroll
                for (int u = 0; u < UPT; ++u) {
                    const int b = b0 + UPT * uq + u;
                    out[u] = b < p.B ? __ldg(p.dmt + ((size_t)t * p.B + b) * Cp + cell) : 0.f;
                }
            }
        };

        // saved activations / dmt of step t-1 are loaded one step ahead (see lstmp_bwd_cluster_kernel)
        // 48 utterances per cluster = 14 warps, four of them on one SM sub-partition, i.e. at most 128 registers per thread
        // where this loop wants 168: there (SPRE) the one-step-ahead prefetch goes to shared memory with cp.async -- no
        // registers -- and is picked up at the end of the step, the two output tiles are drained one after the other and
        // the 16-bit dz values stay packed.
        constexpr bool SPRE = NBP > 32;
        float s_i[UPT], s_f[UPT], s_o[UPT], s_j[UPT], s_c[UPT], s_cp[UPT], dm[UPT];
        float n_i[SPRE ? 1 : UPT], n_f[SPRE ? 1 : UPT], n_o[SPRE ? 1 : UPT], n_j[SPRE ? 1 : UPT], n_cp[SPRE ? 1 : UPT],
              n_dm[SPRE ? 1 : UPT];
        const uint32_t pre_me = sPre + (uint32_t)tid * 4u, predm_me = sPreDm + (uint32_t)tid * 16u;
        dm_ready(0);
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            const int b = b0 + UPT * uq + u;
            s_i[u] = s_f[u] = s_o[u] = s_j[u] = s_c[u] = s_cp[u] = dm[u] = 0.f;
            if (b < p.B) {
                const size_t row = (size_t)(p.T - 1) * p.B + b;
                const float* s = p.save + row * 5 * Cp + cell;
                s_i[u] = __ldg(s); s_f[u] = __ldg(s + Cp); s_o[u] = __ldg(s + 2 * Cp); s_j[u] = __ldg(s + 3 * Cp);
                s_c[u] = __ldg(s + 4 * Cp);
                if (p.T > 1) s_cp[u] = __ldg(s - (size_t)p.B * 5 * Cp + 4 * Cp);
            }
        }
        load_dm4(p.T - 1, dm);

        for (int step = 0; step < p.T; ++step) {
            const int t = p.T - 1 - step;
            const int buf = step & 1;
            PTRACE(tid == 0, step, 0);
            if (t > 0) dm_ready(step + 1);
            if constexpr (SPRE) {
                if (t > 0) {
#pragma unroll
                    for (int u = 0; u < UPT; ++u) {        // operands of step t-1 (and c_{t-2}) -> shared memory, asynchronously
                        const int b = b0 + UPT * uq + u;
                        const bool ok = b < p.B;
                        const float* s = p.save + ((size_t)(t - 1) * p.B + (ok ? b : 0)) * 5 * Cp + cell;
                        const uint32_t d = pre_me + (uint32_t)(5 * u) * (GT * 4u);
                        cp_async4(d, s, ok ? 4u : 0u);
                        cp_async4(d + GT * 4u, s + Cp, ok ? 4u : 0u);
                        cp_async4(d + 2u * GT * 4u, s + 2 * Cp, ok ? 4u : 0u);
                        cp_async4(d + 3u * GT * 4u, s + 3 * Cp, ok ? 4u : 0u);
                        cp_async4(d + 4u * GT * 4u, t > 1 ? s - (size_t)p.B * 5 * Cp + 4 * Cp : s, (ok && t > 1) ? 4u : 0u);
                        if (!p.grouped)
                            cp_async4(predm_me + 4u * u, p.dmt + ((size_t)(t - 1) * p.B + (ok ? b : 0)) * Cp + cell, ok ? 4u : 0u);
                    }
                    if (p.grouped)
                        cp_async16_cg(predm_me, p.dmt + ((((size_t)(t - 1) * p.groups + grp) * Cp + cell) * NBP + e * NBR + UPT * uq));
                }
            } else {
#pragma unroll
                for (int u = 0; u < UPT; ++u) {            // operands of step t-1 (and c_{t-2}): in flight during this step
                    const int b = b0 + UPT * uq + u;
                    n_i[u] = n_f[u] = n_o[u] = n_j[u] = n_cp[u] = n_dm[u] = 0.f;
                    if (b < p.B && t > 0) {
                        const size_t row = (size_t)(t - 1) * p.B + b;
                        const float* s = p.save + row * 5 * Cp + cell;
                        n_i[u] = __ldg(s); n_f[u] = __ldg(s + Cp); n_o[u] = __ldg(s + 2 * Cp); n_j[u] = __ldg(s + 3 * Cp);
                        if (t > 1) n_cp[u] = __ldg(s - (size_t)p.B * 5 * Cp + 4 * Cp);
                    }
                }
                if (t > 0) load_dm4(t - 1, n_dm);
            }
            if (step > 0) {
                const uint32_t fb = buf ? full1 : full0;
                mbar_wait(fb, (uint32_t)(((step - 1) >> 1) & 1));   // partial rows of dz_{t+1} Wc^T from all NP pairs
                PTRACE(tid == 0, step, 1);
                const uint32_t rbase = sR0 + (uint32_t)buf * sR_bytes + roff;
#pragma unroll
                for (int s0 = 0; s0 < 8; s0 += 4) {
                    if (s0 >= NP) break;
                    uint32_t x[4], y[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x[k]), "=r"(y[k]) : "r"(rbase + (uint32_t)(s0 + k) * SLOT));
                    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k = 0; k < 4; k += 2) {
                        a0[0] += h2f((uint16_t)(x[k] & 0xFFFFu), BF); a0[1] += h2f((uint16_t)(x[k] >> 16), BF);
                        a0[2] += h2f((uint16_t)(y[k] & 0xFFFFu), BF); a0[3] += h2f((uint16_t)(y[k] >> 16), BF);
                        a1[0] += h2f((uint16_t)(x[k + 1] & 0xFFFFu), BF); a1[1] += h2f((uint16_t)(x[k + 1] >> 16), BF);
                        a1[2] += h2f((uint16_t)(y[k + 1] & 0xFFFFu), BF); a1[3] += h2f((uint16_t)(y[k + 1] >> 16), BF);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) dm[u] += a0[u] + a1[u];
                }
                __syncwarp();
                if (tid == 0 && step + 2 < p.T) mbar_expect_tx(fb, sR_bytes);   // re-arm for step + 2
            }
            PTRACE(tid == 0, step, 2);
            // dz of (utterance, gate): 16-bit values, one per register -- or (SPRE) packed in pairs (i, j), (f, o)
            uint32_t hz[UPT][SPRE ? 2 : 4];
            auto hz_get = [&](int u, int g) -> uint16_t {
                if constexpr (SPRE) return (uint16_t)((g & 1) ? (hz[u][g >> 1] >> 16) : (hz[u][g >> 1] & 0xFFFFu));
                else return (uint16_t)hz[u][g];
            };
#pragma unroll
            for (int u = 0; u < UPT; ++u) {
                const int b = b0 + UPT * uq + u;
                // branch-free: frozen steps and rows past the batch contribute exact zeros through the mask
                const float m = ((b < p.B) && (t < len[u])) ? 1.f : 0.f;
                const float tc = FAST ? tanh_fast(s_c[u]) : tanhf_(s_c[u]);
                const float dz_o = m * dm[u] * tc * s_o[u] * (1.f - s_o[u]);
                const float dc = dcar[u] + dm[u] * s_o[u] * (1.f - tc * tc) + dz_o * wo;
                const float dz_f = m * dc * s_cp[u] * s_f[u] * (1.f - s_f[u]);
                const float dz_i = m * dc * s_j[u] * s_i[u] * (1.f - s_i[u]);
                const float dz_j = m * dc * s_i[u] * (1.f - s_j[u] * s_j[u]);
                dcar[u] = m * (dc * s_f[u] + dz_f * wf + dz_i * wi);
                a_dwo += dz_o * s_c[u]; a_dwf += dz_f * s_cp[u]; a_dwi += dz_i * s_cp[u];
                a_db[0] += dz_i; a_db[1] += dz_j; a_db[2] += dz_f; a_db[3] += dz_o;
                if constexpr (SPRE) { hz[u][0] = pack2(dz_i, dz_j, BF); hz[u][1] = pack2(dz_f, dz_o, BF); }
                else { hz[u][0] = f2h(dz_i, BF); hz[u][1] = f2h(dz_j, BF); hz[u][2] = f2h(dz_f, BF); hz[u][3] = f2h(dz_o, BF); }
            }
            PTRACE(tid == 0, step, 3);
            if (t > 0) {
                // B operand of the recurrent product first (it is what the next MMA waits for), the global copy after
#pragma unroll
                for (int u = 0; u < UPT; ++u) {
                    const int n = UPT * uq + u;
                    *reinterpret_cast<uint16_t*>(sB_ptr + zoff + sw128_off(n, cl)) = hz_get(u, 0);
                    *reinterpret_cast<uint16_t*>(sB_ptr + zoff + sw128_off(n, 32 + cl)) = hz_get(u, 1);
                    *reinterpret_cast<uint16_t*>(sB_ptr + zoff + NBR * 128 + sw128_off(n, cl)) = hz_get(u, 2);
                    *reinterpret_cast<uint16_t*>(sB_ptr + zoff + NBR * 128 + sw128_off(n, 32 + cl)) = hz_get(u, 3);
                }
                fence_proxy_async_smem();
                named_bar_sync(1, GT);
                if (tid == 0) mbar_arrive_cluster(dzr + lead);
                PTRACE(tid == 0, step, 4);
            }
            // global copy of dz (operand of the weight-gradient GEMMs): off the dependent chain, behind the MMAs
#pragma unroll
            for (int u = 0; u < UPT; ++u) {
                const int b = b0 + UPT * uq + u;
                if (b < p.B) {
                    uint16_t* d = dzg + ((size_t)t * p.B + b) * 4 * Cp;
                    d[0] = hz_get(u, 0); d[32] = hz_get(u, 1); d[64] = hz_get(u, 2); d[96] = hz_get(u, 3);
                }
            }
            if (p.post) named_bar_arrive(post_bar_id(step), GT + 32);
            if (t == 0) break;                     // no earlier step to feed
            mbar_wait(barM, (uint32_t)(step & 1));
            tc_fence_after();
            PTRACE(tid == 0, step, 6);
            const uint32_t dst0 = sR0 + (uint32_t)(buf ^ 1) * sR_bytes;
            const uint32_t dbar = buf ? full0 : full1;
            uint32_t acc[SPRE ? 1 : 2][16];        // both tiles in flight, one wait (SPRE: one after the other)
            if constexpr (!SPRE) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    if (mt < MT2) tmem_ld16_nowait(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * ACCS + piece * 16), acc[mt]);
                tmem_ld_wait();
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (mt < MT2) {
                    if constexpr (SPRE) {
                        tmem_ld16_nowait(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * ACCS + piece * 16), acc[0]);
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t* a = acc[SPRE ? 0 : mt] + 8 * c;
                        st_async_v4(dst0 + soff[c] + rd[mt][c],
                                    pack2(__uint_as_float(a[0]), __uint_as_float(a[1]), BF),
                                    pack2(__uint_as_float(a[2]), __uint_as_float(a[3]), BF),
                                    pack2(__uint_as_float(a[4]), __uint_as_float(a[5]), BF),
                                    pack2(__uint_as_float(a[6]), __uint_as_float(a[7]), BF), dbar + rd[mt][c]);
                    }
                }
            }
            tc_fence_before();
            if constexpr (SPRE) {
                cp_async_wait_all();                   // (each thread reads only what it copied itself)
#pragma unroll
                for (int u = 0; u < UPT; ++u) {
                    const uint32_t d = pre_me + (uint32_t)(5 * u) * (GT * 4u);
                    s_c[u] = s_cp[u];
                    s_i[u] = ld_shared_f32(d); s_f[u] = ld_shared_f32(d + GT * 4u); s_o[u] = ld_shared_f32(d + 2u * GT * 4u);
                    s_j[u] = ld_shared_f32(d + 3u * GT * 4u); s_cp[u] = ld_shared_f32(d + 4u * GT * 4u);
                    dm[u] = ld_shared_f32(predm_me + 4u * u);
                }
                if (p.grouped)
                    __stcg(reinterpret_cast<float4*>(const_cast<float*>(p.dmt)) +
                               ((((size_t)(t - 1) * p.groups + grp) * Cp + cell) * NBP + e * NBR + UPT * uq) / 4,
                           make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
#pragma unroll
                for (int u = 0; u < UPT; ++u) {
                    s_c[u] = s_cp[u]; s_cp[u] = n_cp[u];
                    s_i[u] = n_i[u]; s_f[u] = n_f[u]; s_o[u] = n_o[u]; s_j[u] = n_j[u]; dm[u] = n_dm[u];
                }
            }
            PTRACE(tid == 0, step, 7);
        }
        atomicAdd(p.dw_i + cell, a_dwi); atomicAdd(p.dw_f + cell, a_dwf); atomicAdd(p.dw_o + cell, a_dwo);
        {
            float* db = p.dbias + 256 * pr + 128 * (cc >> 5) + (cc & 31);
            atomicAdd(db, a_db[0]); atomicAdd(db + 32, a_db[1]); atomicAdd(db + 64, a_db[2]); atomicAdd(db + 96, a_db[3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == GW) tmem_dealloc_2cta(tmem, tcols);
}

template <int NBP, int BF, int FAST>
__global__ void __launch_bounds__(PairGeomB<NBP>::THREADS, 1)
lstmp_bwd_pair_kernel(const PBwdParams p) {
    pair_bwd_body<NBP, BF, FAST>(p, (int)blockIdx.x / (p.Cp / 32));
}

// =========================================================================================
// layer wavefront, backward: layer 2's clusters [0, groups), layer 1's [groups, 2 groups) -- one to a few steps behind --
// and between them the last cluster, which turns every dz2_t into layer 1's incoming gradient
//     dmt1_t = (dz2_t K_x2^T) W_p1^T = dz2_t F^T,   F = W_p1 K_x2  [Cp, 4Cp]  (refreshed with the weights, like Wc)
// CTA (mi, ki) of that cluster keeps rows [128 mi, +128) x columns [Cp ki, +Cp) of F resident in TMEM (A operand, 4 K
// slices) and per (group, step) runs D[128 cells, NBP utterances] = F slice x dz2_t slice^T (dz2_t by TMA from the global
// copy layer 2 writes anyway) and adds its K-slice partial into the dmt1 scratch with fire-and-forget 16-byte fp32
// reductions in L2 (the scratch is laid out [t][group][cell][utterance], so a CTA's tile is one contiguous stream and a
// layer-1 thread reads its four utterances as one float4); layer 1 resets every element to zero as it reads it, so the
// scratch is all zeros again when the launch ends (four separate partial buffers summed by the reader were measured
// first: sixteen dependent L2 loads per thread and step, 7 us per step).  Counters as in the forward wavefront.
// =========================================================================================
struct WaveProjB {
    int groups;
    const uint16_t* fT;         // [Cp, 4Cp] F (16-bit)
    float* part;                // dmt1 scratch [T][groups][Cp][NBP], all zeros on entry
    unsigned int* flagA;        // [groups] posted by layer 2 (G per step)
    unsigned int* flagB;        // [groups] posted by this stage (G per step)
};

template <int NBP, int BF>
__device__ __forceinline__ void wave_projb_body(const CUtensorMap* tmZ, const PBwdParams& p, const WaveProjB& w) {
    constexpr int ACCS = PairGeomB<NBP>::ACCS;
    constexpr int MAXG = 8;
    constexpr int NST = 4;                      // dz2 slices in flight: three groups per step, ~2 us of TMA latency each
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = p.Cp / 32, MT = p.Cp / 128;   // K slices = G / MT = 4, each Cp wide
    const int c = (int)cluster_ctarank();
    const int mi = c % MT, ki = c / MT;
    const int KB = p.Cp / 64;
    const uint32_t bt_bytes = (uint32_t)KB * NBP * 128u;             // one dz2_t slice: KB sub-tiles [NBP rows x 64 k], SW128
    constexpr int SP = NBP + 4;                 // staging pitch in floats: 16-byte rows, conflict-free for v4 stores
    const uint32_t sStage = base + (uint32_t)NST * bt_bytes;         // float stage[128][SP]: the D tile, cell-major
    const uint32_t sBar = sStage + 128u * SP * 4u;
    auto sBt = [&](int b) { return base + (uint32_t)b * bt_bytes; };
    auto fullB = [&](int b) { return sBar + 8u * (uint32_t)b; };
    auto emptyB = [&](int b) { return sBar + 32u + 8u * (uint32_t)b; };
    auto accfull = [&](int b) { return sBar + 64u + 8u * (uint32_t)b; };
    auto accfree = [&](int b) { return sBar + 80u + 8u * (uint32_t)b; };
    auto done = [&](int b) { return sBar + 96u + 8u * (uint32_t)b; };
    const uint32_t tslot = sBar + 112u;
    const uint32_t a_cols = (uint32_t)p.Cp / 2u;
    uint32_t tcols = 32;
    while (tcols < a_cols + 2u * ACCS) tcols <<= 1;
    if (tid == 0) {
        tma_prefetch_desc(tmZ);
        for (int b = 0; b < NST; ++b) { mbar_init(fullB(b), 1); mbar_init(emptyB(b), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(accfull(b), 1); mbar_init(accfree(b), 4); }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(done(0)), "r"(0u) : "memory");
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tslot, tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tslot));
    const uint32_t tmem_acc = tmem + a_cols;
    if (warp < 4) {   // F slice -> TMEM: thread <-> layer-1 cell, 64 k (32 columns) per store
        const uint16_t* wrow = w.fT + (size_t)(128 * mi + 32 * warp + lane) * 4 * p.Cp + (size_t)ki * p.Cp;
        for (int cb = 0; cb < KB; ++cb) {
            uint32_t r[32];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + cb * 64) + k);
                r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
            }
            tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb * 32u, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp >= 7) return;                      // (only mbarriers and named barriers below)

    const int items = p.T * w.groups;           // item = (step, group), step-major
    if (warp == 0) {
        if (elect_one_sync()) {                 // producer: dz2_t slices as layer 2 releases them
            unsigned int seen[MAXG];
#pragma unroll
            for (int g = 0; g < MAXG; ++g) seen[g] = 0;
            for (int it = 0; it < items; ++it) {
                const int s = it / w.groups, g = it % w.groups, buf = it % NST;
                if (it >= NST) mbar_wait(emptyB(buf), (uint32_t)((it / NST - 1) & 1));
                const unsigned int need = (unsigned int)G * (unsigned int)(s + 1);
                if (seen[g] < need) { seen[g] = spin_get_ge(w.flagA + g, need); fence_proxy_async_all(); }
                mbar_expect_tx(fullB(buf), bt_bytes);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(sBt(buf) + (uint32_t)kb * (NBP * 128u), tmZ, fullB(buf), ki * p.Cp + kb * 64,
                                (p.T - 1 - s) * p.B + g * NBP);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc(128, NBP, BF, 0, 0);
            for (int it = 0; it < items; ++it) {
                const int buf = it % NST, ab = it & 1;
                mbar_wait(fullB(buf), (uint32_t)((it / NST) & 1));
                if (it >= 2) mbar_wait(accfree(ab), (uint32_t)(((it - 2) >> 1) & 1));
                tc_fence_after();
                const uint32_t sb = sBt(buf);
                for (int kk = 0; kk < p.Cp / 16; ++kk) {
                    const uint64_t db = umma_desc_sw128(sb + (uint32_t)(kk >> 2) * (NBP * 128u) + (uint32_t)(kk & 3) * 32u, 16, 1024);
                    tc_mma_f16_ts(tmem_acc + (uint32_t)ab * ACCS, tmem + (uint32_t)kk * 8u, db, idesc, kk ? 1u : 0u);
                }
                tc_commit(emptyB(buf));
                tc_commit(accfull(ab));
            }
        }
        __syncwarp();
    } else if (warp == 6) {
        // poster: releases a whole step's rows (all groups) with ONE fence -- a fence per group would cost more than the
        // step.  `done` is a monotonic count of (epilogue warp, item) completions in shared memory (no phase to lap).
        if (elect_one_sync()) {
            for (int s1 = 1; s1 <= p.T; ++s1) {
                const unsigned int need = 4u * (unsigned int)(s1 * w.groups);
                unsigned int v;
                do {
                    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(done(0)) : "memory");
                } while (v < need);
                __threadfence();
                for (int g = 0; g < w.groups; ++g)
                    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(w.flagB + g), "r"(1u) : "memory");
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                 // warps 2..5 <-> TMEM lane quadrants 2, 3, 0, 1
        const int et = tid - 64;                // 0..127
        const uint32_t srow = sStage + (uint32_t)(32 * q + lane) * (SP * 4u);
        for (int it = 0; it < items; ++it) {
            const int s = it / w.groups, g = it % w.groups, buf = it & 1;
            mbar_wait(accfull(buf), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            uint32_t acc[NBP];
#pragma unroll
            for (int c16 = 0; c16 < NBP / 16; ++c16)
                tmem_ld16_nowait(tmem_acc + (uint32_t)buf * ACCS + ((uint32_t)(q * 32) << 16) + (uint32_t)c16 * 16u, acc + 16 * c16);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accfree(buf));           // this warp's quadrant of the accumulator is drained
            // D tile -> staging (thread <-> cell row), then the tile leaves as one linear stream of 16-byte reductions:
            // dmt1 scratch [t][group][cell][NBP], so the 128 x NBP tile of this CTA is NBP * 512 contiguous bytes
#pragma unroll
            for (int k = 0; k < NBP / 4; ++k)
                st_shared_v4(srow + 16u * (uint32_t)k, acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
            named_bar_sync(3u, 128);
            float* tile = w.part + ((((size_t)(p.T - 1 - s) * w.groups + g) * p.Cp) + 128 * mi) * NBP;
#pragma unroll
            for (int k = 0; k < NBP / 4; ++k) {
                const int f = et + 128 * k, cell = f / (NBP / 4), part = f % (NBP / 4);
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(sStage + (uint32_t)cell * (SP * 4u) + 16u * (uint32_t)part));
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(tile + 4 * (size_t)f), "f"(v.x), "f"(v.y), "f"(v.z),
                             "f"(v.w) : "memory");
            }
            named_bar_sync(5u, 128);                            // staging may be rewritten; all four warps' rows are on their way
            if (lane == 0) asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(done(0)), "r"(1u) : "memory");
        }
    }
    named_bar_sync(4u, 224);
    if (warp == 0) tmem_dealloc(tmem, tcols);
}

template <int NBP, int BF, int FAST>
__global__ void __launch_bounds__(PairGeomB<NBP>::THREADS, 1)
lstmp_wave_bwd_kernel(const __grid_constant__ CUtensorMap tmZ, const PBwdParams p2, const PBwdParams p1, const WaveProjB w) {
    const int cid = (int)blockIdx.x / (p2.Cp / 32);
    if (cid < w.groups) pair_bwd_body<NBP, BF, FAST>(p2, cid);
    else if (cid < 2 * w.groups) pair_bwd_body<NBP, BF, FAST>(p1, cid - w.groups);
    else wave_projb_body<NBP, BF>(&tmZ, p2, w);
}

size_t pbwd_smem(int Cp, int nbp) {
    const size_t gt = 16 * (size_t)(nbp / 2);       // gate threads
    const size_t need = 1024 + 4 * (size_t)(nbp / 2) * 128 + 2 * (size_t)(Cp / 64) * (nbp / 2) * 128 + 64 +
                        (nbp > 32 ? 20 * gt * 4 + gt * 16 : 0);
    return need < RSR_EXCLUSIVE_SMEM_REC ? RSR_EXCLUSIVE_SMEM_REC : need;
}

void fill_bwd_params(PBwdParams& p, rsr_handle* h, int B, int T, int Cp, const float* dmt, const void* wc, const float* w_i,
                     const float* w_f, const float* w_o, const int* lengths, const float* save, void* dz16, float* dbias,
                     float* dw_i, float* dw_f, float* dw_o) {
    p.wc = (const uint16_t*)wc;
    p.B = B; p.T = T; p.Cp = Cp; p.bf = h->dtype == RSR_DTYPE_BF16;
    p.dmt = dmt; p.grouped = 0; p.groups = 0;
    p.w_i = w_i; p.w_f = w_f; p.w_o = w_o; p.lengths = lengths; p.save = save;
    p.dz16 = (uint16_t*)dz16; p.dbias = dbias; p.dw_i = dw_i; p.dw_f = dw_f; p.dw_o = dw_o;
    p.post = nullptr; p.wait = nullptr; p.wait_per_step = 0;
}

}  // namespace

int rsr_lstmp_bwd_pair(rsr_handle* h, void* stream, int B, int T, int Cp, const float* dmt, const void* wc,
                       const float* w_i, const float* w_f, const float* w_o, const int* lengths,
                       const float* save, void* dz16, float* dbias, float* dw_i, float* dw_f, float* dw_o) {
    constexpr int NBP = 32;
    using GM = PairGeomB<NBP>;
    if (Cp > 512) return RSR_E_RESIDENT;
    const int G = Cp / 32;
    const size_t smem = pbwd_smem(Cp, NBP);
    int& cap = h->pair_cap[1][Cp / 256 - 1];
    if (cap < 0) {
        std::lock_guard<std::mutex> g(h->mu);
        cap = cluster_capacity(lstmp_bwd_pair_kernel<NBP, 0, 0>, G, GM::THREADS, smem);
        if (cap > 0) {   // sets the launch attributes of the other instances
            cluster_capacity(lstmp_bwd_pair_kernel<NBP, 1, 0>, G, GM::THREADS, smem);
            cluster_capacity(lstmp_bwd_pair_kernel<NBP, 0, 1>, G, GM::THREADS, smem);
            cluster_capacity(lstmp_bwd_pair_kernel<NBP, 1, 1>, G, GM::THREADS, smem);
        }
        if (getenv("RSR_DEBUG")) fprintf(stderr, "[rsr] bwd pair kernel Cp=%d: %d-CTA clusters co-resident: %d\n", Cp, G, cap);
    }
    if (cap <= 0) return RSR_E_RESIDENT;
    const int groups = (B + NBP - 1) / NBP;
    PBwdParams p;
    fill_bwd_params(p, h, B, T, Cp, dmt, wc, w_i, w_f, w_o, lengths, save, dz16, dbias, dw_i, dw_f, dw_o);
    if (fast_gates()) {
        if (p.bf) return cluster_launch(lstmp_bwd_pair_kernel<NBP, 1, 1>, groups, G, GM::THREADS, smem, (cudaStream_t)stream, p);
        return cluster_launch(lstmp_bwd_pair_kernel<NBP, 0, 1>, groups, G, GM::THREADS, smem, (cudaStream_t)stream, p);
    }
    if (p.bf) return cluster_launch(lstmp_bwd_pair_kernel<NBP, 1, 0>, groups, G, GM::THREADS, smem, (cudaStream_t)stream, p);
    return cluster_launch(lstmp_bwd_pair_kernel<NBP, 0, 0>, groups, G, GM::THREADS, smem, (cudaStream_t)stream, p);
}

// Backward of two stacked LSTMP layers of equal cell count as one wavefront launch (see lstmp_wave_bwd_kernel).
extern "C" int rsr_lstmp_wave_bwd(rsr_handle* h, void* stream, const rsr_wave_bwd_args* a) {
    if (!h || !a) return RSR_E_ARG;
    const int B = a->B, T = a->T, Cp = a->Cp;
    if (!a->lengths || !a->dmt2 || !a->wc2 || !a->w_i2 || !a->w_f2 || !a->w_o2 || !a->save2 || !a->dz2 || !a->dbias2 || !a->dw_i2 ||
        !a->dw_f2 || !a->dw_o2 || !a->fT || !a->part || !a->wc1 || !a->w_i1 || !a->w_f1 || !a->w_o1 || !a->save1 || !a->dz1 ||
        !a->dbias1 || !a->dw_i1 || !a->dw_f1 || !a->dw_o1)
        return RSR_E_ARG;
    if (B <= 0 || T <= 0 || Cp <= 0 || (Cp & 255)) return RSR_E_SHAPE;
    if (((uintptr_t)a->dmt2 | (uintptr_t)a->wc2 | (uintptr_t)a->save2 | (uintptr_t)a->dz2 | (uintptr_t)a->fT | (uintptr_t)a->part |
         (uintptr_t)a->wc1 | (uintptr_t)a->save1 | (uintptr_t)a->dz1) & 15)
        return RSR_E_ARG;
    if (getenv("RSR_NO_CLUSTER") || getenv("RSR_NO_PAIR") || getenv("RSR_NO_WAVE") || getenv("RSR_NO_WAVE_BWD") || Cp > 512)
        return RSR_E_RESIDENT;
    const int G = Cp / 32;
    const int bf = h->dtype == RSR_DTYPE_BF16, fast = fast_gates();
    auto run = [&](auto kernel, auto nbp_tag) -> int {
        constexpr int NBP = decltype(nbp_tag)::value;
        using GM = PairGeomB<NBP>;
        const int groups = (B + NBP - 1) / NBP;
        if (groups > 8) return RSR_E_RESIDENT;
        size_t smem = pbwd_smem(Cp, NBP);
        if (wproj_smem(Cp, NBP, 4) + 128 * (NBP + 4) * 4 > smem) smem = wproj_smem(Cp, NBP, 4) + 128 * (NBP + 4) * 4;
        if (smem > (size_t)h->max_smem) return RSR_E_RESIDENT;
        int cap;
        {
            std::lock_guard<std::mutex> g(h->mu);
            int& c = h->wave_cap[2 + (NBP == 32 ? 0 : 1)][Cp / 256 - 1];
            const int key = bf * 2 + fast;
            if (c < 0 || h->wave_key[2 + (NBP == 32 ? 0 : 1)][Cp / 256 - 1] != key) {
                c = cluster_capacity(kernel, G, GM::THREADS, smem);
                h->wave_key[2 + (NBP == 32 ? 0 : 1)][Cp / 256 - 1] = key;
                if (getenv("RSR_DEBUG")) fprintf(stderr, "[rsr] wave bwd kernel Cp=%d NBP=%d: %d-CTA clusters co-resident: %d\n", Cp, NBP, G, c);
            }
            cap = c;
        }
        if (cap < 2 * groups + 1) return RSR_E_RESIDENT;
        CUtensorMap tmZ;
        int rc = rsr_get_tmap(h, a->dz2, (uint64_t)4 * Cp, (uint64_t)T * B, (uint64_t)4 * Cp, 64, (uint32_t)NBP, &tmZ);
        if (rc) return rc;
        unsigned int* flags = rsr_take_flags(h, 2 * groups);
        { const int rz = rsr_zero_u32(flags, 2 * groups, (cudaStream_t)stream); if (rz) return rz; }
        PBwdParams p2, p1;
        fill_bwd_params(p2, h, B, T, Cp, a->dmt2, a->wc2, a->w_i2, a->w_f2, a->w_o2, a->lengths, a->save2, a->dz2, a->dbias2,
                        a->dw_i2, a->dw_f2, a->dw_o2);
        fill_bwd_params(p1, h, B, T, Cp, a->part, a->wc1, a->w_i1, a->w_f1, a->w_o1, a->lengths, a->save1, a->dz1, a->dbias1,
                        a->dw_i1, a->dw_f1, a->dw_o1);
        p2.post = flags;
        p1.grouped = 1; p1.groups = groups;
        p1.wait = flags + groups; p1.wait_per_step = (unsigned int)G;
        WaveProjB w;
        w.groups = groups; w.fT = (const uint16_t*)a->fT; w.part = a->part;
        w.flagA = flags; w.flagB = flags + groups;
        return cluster_launch(kernel, 2 * groups + 1, G, GM::THREADS, smem, (cudaStream_t)stream, tmZ, p2, p1, w);
    };
    auto pick = [&](auto nbp_tag) -> int {
        constexpr int NBP = decltype(nbp_tag)::value;
        if (fast) return bf ? run(lstmp_wave_bwd_kernel<NBP, 1, 1>, nbp_tag) : run(lstmp_wave_bwd_kernel<NBP, 0, 1>, nbp_tag);
        return bf ? run(lstmp_wave_bwd_kernel<NBP, 1, 0>, nbp_tag) : run(lstmp_wave_bwd_kernel<NBP, 0, 0>, nbp_tag);
    };
    // 32 utterances per cluster when the batch seats that way (2.7 us per step), else 48 (3.9 us; B = 128 at Cp = 512: the 7
    // placeable clusters are 3 + 3 + 1).  The 48-utterance variant has 14 warps, i.e. four on one SM sub-partition and at
    // most 128 registers per thread where the gate-backward loop wants 168: with the operand prefetch in registers it
    // spilled and took 5.9 us per step, slower than the two launches one after the other (2 x 2.4 + two GEMMs;
    // profiles/r2_wave_steps_v5.txt, r2_wave_trace_bwd_v0.txt).  Tried and worse: setmaxnreg (ptxas kept allocating for 128
    // and spilled 2.4 KB per thread); no prefetch at all (8.5 us: the saved activations stream from HBM, their latency does
    // not hide behind the wait for the partial rows, r2_wave_steps_v6.txt).  What works is the prefetch through shared
    // memory with cp.async (no registers), tiles drained one after the other, packed dz: r2_wave_steps_v7.txt.
    const int force = getenv("RSR_WAVE_NBP") ? atoi(getenv("RSR_WAVE_NBP")) : 0;
    if (force != 48) {
        const int rc = pick(std::integral_constant<int, 32>());
        if (rc != RSR_E_RESIDENT || force == 32 || (a->max_nbp > 0 && a->max_nbp < 48)) return rc;
    }
    return pick(std::integral_constant<int, 48>());
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
