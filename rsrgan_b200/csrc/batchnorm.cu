// tf.contrib.layers.batch_norm(is_training, scale=True, renorm=True) between a bias-free
// fully_connected and its activation, and tf.nn.dropout behind it -- the optional normalizer /
// dropout of models/dnn.py:56-62,79-110, models/discriminator_dnn.py:36-46,61-90 and
// models/lstm.py:61-67,82-87.  All kernels are HBM streams over the fp32 pre-activation z
// [rows, N] that the tensor-core GEMM wrote:
//   statistics      one read of z  (4 B / element), per-thread Welford, fixed-order Chan merges
//   normalise       one read of z + one 16-bit write (6 B / element), activation and dropout fused
//   backward        two reads of z and of the incoming 16-bit gradient + one 16-bit write
//                   (2 x 6 + 2 B / element): column sums first, then dz
// Reductions never use atomics: every column's partials are merged in one fixed order, so
// data-parallel replicas that see the same rows produce bit-identical statistics.
#include <cstdlib>

#include "common.cuh"
#include "handle.h"

using namespace rsr;

namespace {

// reduction kernels: 256-thread blocks of TX column groups (float4 each) x 256 / TX row lanes; a wider TX reads longer
// contiguous runs of every row (TX * 16 bytes) at the price of more row splits to merge
constexpr int BN_MAX_SPLITS = 256;   // RSR_BN_SCRATCH_FLOATS(N) = 3 * BN_MAX_SPLITS * N
constexpr int FIN_COLS = 32;         // finish kernels: 32 columns x 16 split lanes per block
constexpr int FIN_LANES = 16;

// Row splits of the two reduction kernels: one wave of 3 resident 256-thread blocks per SM over all column tiles
// (fat blocks: no wave quantisation, few partials to merge), at least 8 rows per split.
inline int bn_tx(int N) {
    static const int forced = getenv("RSR_BN_TX") ? atoi(getenv("RSR_BN_TX")) : 0;      // tuning aid: 32 | 64 | 128
    if (forced == 32 || forced == 64 || forced == 128) return forced;
    return N >= 512 ? 64 : 32;
}
inline int bn_splits(long long rows, int N, int num_sms) {
    const int cols = bn_tx(N) * 4;
    const int tiles = (N + cols - 1) / cols;
    long long s = (3LL * num_sms) / tiles;
    if (s > rows / 8) s = rows / 8;
    if (s > BN_MAX_SPLITS) s = BN_MAX_SPLITS;
    if (s < 1) s = 1;
    return (int)s;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
// counter-based dropout mask: the backward pass regenerates it from (seed, tick, salt, element index)
__device__ __forceinline__ uint64_t drop_key(const unsigned long long* rng, unsigned salt) {
    return splitmix64(rng[0] + 0x9E3779B97F4A7C15ull * (rng[1] * 65536ull + salt));
}
// one hash serves the two elements of an (even, odd) column pair: bits 63..40 and bits 39..16
__device__ __forceinline__ void drop_keep2(uint64_t key, uint64_t idx_even, uint32_t thr24, bool& k0, bool& k1) {
    const uint64_t hsh = splitmix64(key ^ (idx_even >> 1));
    k0 = (uint32_t)(hsh >> 40) < thr24;
    k1 = ((uint32_t)(hsh >> 16) & 0xffffffu) < thr24;
}

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == RSR_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == RSR_ACT_LRELU) return fmaxf(v, 0.3f * v);
    return v;
}
__device__ __forceinline__ float act_slope(float y, int act) {
    if (act == RSR_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
    if (act == RSR_ACT_LRELU) return y > 0.0f ? 1.0f : 0.3f;
    return 1.0f;
}

struct BnBwdIn {
    const uint16_t* da; int ldda;   // gradient wrt the layer output (after activation and dropout)
    const float* A; const float* Bc; const float* mean; const float* inv_std;
    int act; uint32_t thr24; float inv_keep; const unsigned long long* rng; unsigned salt; int bf;
    int S, L;                       // frame layout of the convolutional generator (S = 0: every row counts): row r is data
                                    // iff r % S < L; the rows behind are the SAME padding shared with the next frame
};
// (the frame layout has fewer than 2^31 rows -- lines_check -- so the remainder is a 32-bit one)
__device__ __forceinline__ bool row_live(long long r, int S, int L) {
    return S == 0 || (unsigned)r % (unsigned)S < (unsigned)L;
}

// per-thread column coefficients (4 consecutive columns), loaded once
struct Col4 { float A[4], B[4], mean[4], istd[4]; };
__device__ __forceinline__ Col4 load_col4(const BnBwdIn& p, int c) {
    Col4 k;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        k.A[j] = p.A ? p.A[c + j] : 1.0f;
        k.B[j] = p.Bc[c + j];
        k.mean[j] = p.mean ? p.mean[c + j] : 0.0f;
        k.istd[j] = p.mean ? p.inv_std[c + j] : 0.0f;
    }
    return k;
}

// gradient wrt the normalised value y = z A + B of 4 consecutive columns, and x_hat
__device__ __forceinline__ void bwd_elem4(const BnBwdIn& p, const Col4& k, const float4 z, const uint2 raw, long long r,
                                          int c, int N, uint64_t key, float g[4], float xh[4]) {
    const uint16_t hv[4] = {(uint16_t)(raw.x & 0xffff), (uint16_t)(raw.x >> 16), (uint16_t)(raw.y & 0xffff),
                            (uint16_t)(raw.y >> 16)};
    const float zz[4] = {z.x, z.y, z.z, z.w};
    bool keep[4] = {true, true, true, true};
    if (p.thr24 < (1u << 24)) {
        const uint64_t idx = (uint64_t)r * (uint64_t)N + (uint64_t)c;
        drop_keep2(key, idx, p.thr24, keep[0], keep[1]);
        drop_keep2(key, idx + 2, p.thr24, keep[2], keep[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float y = fmaf(zz[j], k.A[j], k.B[j]);
        const float gv = h2f(hv[j], p.bf) * act_slope(y, p.act);
        g[j] = keep[j] ? gv * p.inv_keep : 0.0f;
        xh[j] = (zz[j] - k.mean[j]) * k.istd[j];
    }
}

// MODE 0: per-split (count, mean, M2) of z.   MODE 1: per-split (sum g, sum g x_hat).
// grid (ceil(N / (4 TX)), splits), block (TX, 256 / TX); partial layout [split][3 | 2][N]
// MODE 0 uses shifted sums (shift = the first value the thread sees: no cancellation, no per-element division, no
// dependency chain between the loads), converted to (n, mean, M2) per thread and merged pairwise (Chan) in a fixed order.
template <int MODE, int TX>
__global__ void __launch_bounds__(256) bn_partial_kernel(const float* __restrict__ z, int ldz, long long rows, int N,
                                                         BnBwdIn p, float* __restrict__ partial) {
    constexpr int BN_COLS = TX * 4, BN_LANES = 256 / TX;
    __shared__ float sm[BN_LANES][3][BN_COLS];
    const int c = blockIdx.x * BN_COLS + threadIdx.x * 4;
    const int splits = gridDim.y;
    const long long chunk = (rows + splits - 1) / splits;
    const long long r0 = (long long)blockIdx.y * chunk;
    const long long r1 = r0 + chunk < rows ? r0 + chunk : rows;
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    float n = 0.f;
    if (c < N) {
        if (MODE == 0) {
            float sh[4] = {0.f, 0.f, 0.f, 0.f};
            const long long rf = r0 + threadIdx.y;
            if (rf < r1) {
                const float4 v = *reinterpret_cast<const float4*>(z + rf * ldz + c);
                sh[0] = v.x; sh[1] = v.y; sh[2] = v.z; sh[3] = v.w;
            }
            auto acc = [&](const float4 v) {
                const float vv[4] = {v.x, v.y, v.z, v.w};
                n += 1.0f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d = vv[k] - sh[k];
                    a0[k] += d;
                    a1[k] = fmaf(d, d, a1[k]);
                }
            };
            long long r = rf;
            for (; r + 3 * BN_LANES < r1; r += 4 * BN_LANES) {       // four independent 16-byte loads in flight
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(z + (r + u * BN_LANES) * ldz + c);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (row_live(r + u * BN_LANES, p.S, p.L)) acc(v[u]);
            }
            for (; r < r1; r += BN_LANES)
                if (row_live(r, p.S, p.L)) acc(*reinterpret_cast<const float4*>(z + r * ldz + c));
            if (n > 0.f) {
                const float inv_n = 1.0f / n;
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // -> a0 = mean, a1 = M2 of this thread's rows
                    const float s = a0[k];
                    a0[k] = sh[k] + s * inv_n;
                    a1[k] = fmaxf(a1[k] - s * s * inv_n, 0.0f);
                }
            }
        } else {
            uint64_t key = 0;
            if (p.thr24 < (1u << 24)) key = drop_key(p.rng, p.salt);
            const Col4 k4 = load_col4(p, c);
            long long r = r0 + threadIdx.y;
            for (; r + BN_LANES < r1; r += 2 * BN_LANES) {           // two rows (4 loads) in flight
                float4 v[2];
                uint2 raw[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    v[u] = *reinterpret_cast<const float4*>(z + (r + u * BN_LANES) * ldz + c);
                    raw[u] = *reinterpret_cast<const uint2*>(p.da + (r + u * BN_LANES) * p.ldda + c);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (!row_live(r + u * BN_LANES, p.S, p.L)) continue;
                    float g[4], xh[4];
                    bwd_elem4(p, k4, v[u], raw[u], r + u * BN_LANES, c, N, key, g, xh);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { a0[k] += g[k]; a1[k] = fmaf(g[k], xh[k], a1[k]); }
                }
            }
            for (; r < r1; r += BN_LANES) {
                if (!row_live(r, p.S, p.L)) continue;
                const float4 v = *reinterpret_cast<const float4*>(z + r * ldz + c);
                const uint2 raw = *reinterpret_cast<const uint2*>(p.da + r * p.ldda + c);
                float g[4], xh[4];
                bwd_elem4(p, k4, v, raw, r, c, N, key, g, xh);
#pragma unroll
                for (int k = 0; k < 4; ++k) { a0[k] += g[k]; a1[k] = fmaf(g[k], xh[k], a1[k]); }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sm[threadIdx.y][0][threadIdx.x * 4 + k] = a0[k];
        sm[threadIdx.y][1][threadIdx.x * 4 + k] = a1[k];
    }
    if (MODE == 0) sm[threadIdx.y][2][threadIdx.x * 4] = n;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cc = threadIdx.x * 4 + k;
            if (MODE == 0) {
                float na = 0.f, ma = 0.f, qa = 0.f;
                for (int l = 0; l < BN_LANES; ++l) {       // Chan merge, fixed order
                    const float nb = sm[l][2][threadIdx.x * 4];
                    if (nb == 0.f) continue;
                    const float mb = sm[l][0][cc], qb = sm[l][1][cc];
                    const float nab = na + nb, d = mb - ma;
                    ma += d * (nb / nab);
                    qa += qb + d * d * (na * nb / nab);
                    na = nab;
                }
                partial[((long long)blockIdx.y * 3 + 0) * N + c + k] = na;
                partial[((long long)blockIdx.y * 3 + 1) * N + c + k] = ma;
                partial[((long long)blockIdx.y * 3 + 2) * N + c + k] = qa;
            } else {
                float s1 = 0.f, s2 = 0.f;
                for (int l = 0; l < BN_LANES; ++l) { s1 += sm[l][0][cc]; s2 += sm[l][1][cc]; }
                partial[((long long)blockIdx.y * 2 + 0) * N + c + k] = s1;
                partial[((long long)blockIdx.y * 2 + 1) * N + c + k] = s2;
            }
        }
    }
}

// Merges the per-split partials of one column: block (32 columns, 16 split lanes); lane y merges splits y, y + 16, ...
// in order, then lane 0 merges the 16 lane results in order -- coalesced loads, a chain of splits/16 + 16 merges
// instead of `splits`, and still one fixed order (bit-identical across replicas).
__device__ __forceinline__ void chan_merge(float& na, float& ma, float& qa, float nb, float mb, float qb) {
    if (nb == 0.f) return;
    const float nab = na + nb, d = mb - ma;
    ma += d * (nb / nab);
    qa += qb + d * d * (na * nb / nab);
    na = nab;
}
constexpr int FIN_PER_LANE = BN_MAX_SPLITS / FIN_LANES;
__device__ __forceinline__ bool merge_moments(const float* __restrict__ partial, int splits, int N, int c, float& n,
                                              float& m, float& q) {
    __shared__ float sm[FIN_LANES][3][FIN_COLS];
    float pn[FIN_PER_LANE], pm[FIN_PER_LANE], pq[FIN_PER_LANE];
#pragma unroll
    for (int i = 0; i < FIN_PER_LANE; ++i) {               // every load in flight before the first merge
        const int s = threadIdx.y + i * FIN_LANES;
        const bool ok = c < N && s < splits;
        pn[i] = ok ? partial[((long long)s * 3 + 0) * N + c] : 0.f;
        pm[i] = ok ? partial[((long long)s * 3 + 1) * N + c] : 0.f;
        pq[i] = ok ? partial[((long long)s * 3 + 2) * N + c] : 0.f;
    }
    float na = 0.f, ma = 0.f, qa = 0.f;
#pragma unroll
    for (int i = 0; i < FIN_PER_LANE; ++i) chan_merge(na, ma, qa, pn[i], pm[i], pq[i]);
    sm[threadIdx.y][0][threadIdx.x] = na; sm[threadIdx.y][1][threadIdx.x] = ma; sm[threadIdx.y][2][threadIdx.x] = qa;
    __syncthreads();
    if (threadIdx.y != 0 || c >= N) return false;
    for (int l = 1; l < FIN_LANES; ++l) chan_merge(na, ma, qa, sm[l][0][threadIdx.x], sm[l][1][threadIdx.x], sm[l][2][threadIdx.x]);
    n = na; m = ma; q = qa;
    return true;
}
__device__ __forceinline__ bool merge_sums(const float* __restrict__ partial, int splits, int N, int c, float& s1,
                                           float& s2) {
    __shared__ float sm[FIN_LANES][2][FIN_COLS];
    float pa[FIN_PER_LANE], pb[FIN_PER_LANE];
#pragma unroll
    for (int i = 0; i < FIN_PER_LANE; ++i) {
        const int s = threadIdx.y + i * FIN_LANES;
        const bool ok = c < N && s < splits;
        pa[i] = ok ? partial[((long long)s * 2 + 0) * N + c] : 0.f;
        pb[i] = ok ? partial[((long long)s * 2 + 1) * N + c] : 0.f;
    }
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < FIN_PER_LANE; ++i) { a += pa[i]; b += pb[i]; }
    sm[threadIdx.y][0][threadIdx.x] = a; sm[threadIdx.y][1][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y != 0 || c >= N) return false;
    for (int l = 1; l < FIN_LANES; ++l) { a += sm[l][0][threadIdx.x]; b += sm[l][1][threadIdx.x]; }
    s1 = a; s2 = b;
    return true;
}

// state rows: 0 moving_mean 1 moving_variance 2 renorm_mean 3 renorm_stddev 4 renorm_mean_weight 5 renorm_stddev_weight
// coef  rows: 0 A = scale / stddev  1 B = offset - mean A  2 mean  3 1 / stddev  4 r  5 d  6 mean(g)  7 mean(g x_hat)
// Renorm corrections, coefficients and (optionally) the UPDATE_OPS of ONE channel from its batch moments.  `sc` = the
// channel's index in the state rows (pitch lds); returns coef rows 0-5 in k[6].
__device__ __forceinline__ void renorm_channel(float mean, float var, float gamma, float beta, float eps,
                                               float* __restrict__ state, long long lds, int sc, float momentum,
                                               float renorm_momentum, int update_state, float k6[6]) {
    const float stddev = sqrtf(var + eps);
    float* mm = state + 0 * lds; float* mv = state + 1 * lds;
    float* rm = state + 2 * lds; float* rs = state + 3 * lds;
    float* rmw = state + 4 * lds; float* rsw = state + 5 * lds;
    const int c = sc;
    // corrections from the PRE-update renorm averages, "as if they were initialised with this batch's moments"
    const float mixed_mean = rm[c] + (1.0f - rmw[c]) * mean;
    const float mixed_std = rs[c] + (1.0f - rsw[c]) * stddev;
    const float r = stddev / mixed_std;
    const float d = (mean - mixed_mean) / mixed_std;
    const float scale = r * gamma, offset = fmaf(d, gamma, beta);
    const float A = scale / stddev;
    k6[0] = A; k6[1] = offset - mean * A; k6[2] = mean; k6[3] = 1.0f / stddev; k6[4] = r; k6[5] = d;
    if (update_state) {
        const float k = 1.0f - renorm_momentum;
        const float rm_n = rm[c] - (rm[c] - mean) * k, rmw_n = rmw[c] - (rmw[c] - 1.0f) * k;
        const float rs_n = rs[c] - (rs[c] - stddev) * k, rsw_n = rsw[c] - (rsw[c] - 1.0f) * k;
        const float new_mean = rm_n / rmw_n, new_std = rs_n / rsw_n;
        const float new_var = new_std * new_std - eps;
        rm[c] = rm_n; rmw[c] = rmw_n; rs[c] = rs_n; rsw[c] = rsw_n;
        mm[c] -= (mm[c] - new_mean) * (1.0f - momentum);
        mv[c] -= (mv[c] - new_var) * (1.0f - momentum);
    }
}

__global__ void bn_finish_train_kernel(const float* __restrict__ partial, int splits, long long rows, int N,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                       float* __restrict__ state, float momentum, float renorm_momentum,
                                       int update_state, float* __restrict__ coef) {
    const int c = blockIdx.x * FIN_COLS + threadIdx.x;
    float na, ma, qa;
    if (!merge_moments(partial, splits, N, c, na, ma, qa)) return;
    float k6[6];
    renorm_channel(ma, qa / (float)rows, gamma[c], beta[c], eps, state, N, c, momentum, renorm_momentum, update_state, k6);
#pragma unroll
    for (int j = 0; j < 6; ++j) coef[j * (long long)N + c] = k6[j];
}

// batch_norm behind a [splice, w] convolution (models/rced.py:63-71,94-97): the normalised axis is the conv2d channel,
// pooled over (frame, line, position).  In the channels-last frame buffer the H lines of channel ch are the columns
// line * C + ch, and only rows r % S < L are data (the partial kernels skip the rest): one block per channel merges the
// splits * H column partials -- thread-strided in order, then a fixed binary tree -- and writes the channel's
// coefficients to its H columns, so the normalise / backward kernels stay per-column.  Columns >= H * C (padding of the
// GEMM N) get zero coefficients.  state rows have pitch lds (>= C).
constexpr int LINES_THREADS = 256;
__global__ void __launch_bounds__(LINES_THREADS) bn_finish_train_lines_kernel(
    const float* __restrict__ partial, int splits, float count, int N, int H, int C, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, float* __restrict__ state, int lds, float momentum, float renorm_momentum,
    int update_state, float* __restrict__ coef) {
    __shared__ float sn[LINES_THREADS], smn[LINES_THREADS], sq[LINES_THREADS];
    __shared__ float k6s[6];
    const int ch = blockIdx.x, t = threadIdx.x;
    float na = 0.f, ma = 0.f, qa = 0.f;
    for (int i = t; i < splits * H; i += LINES_THREADS) {
        const int split = i / H, col = (i % H) * C + ch;
        chan_merge(na, ma, qa, partial[((long long)split * 3 + 0) * N + col], partial[((long long)split * 3 + 1) * N + col],
                   partial[((long long)split * 3 + 2) * N + col]);
    }
    sn[t] = na; smn[t] = ma; sq[t] = qa;
    __syncthreads();
    for (int s = LINES_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) {
            chan_merge(na, ma, qa, sn[t + s], smn[t + s], sq[t + s]);
            sn[t] = na; smn[t] = ma; sq[t] = qa;
        }
        __syncthreads();
    }
    if (t == 0) {
        float k6[6];
        renorm_channel(ma, qa / count, gamma[ch], beta[ch], eps, state, lds, ch, momentum, renorm_momentum, update_state,
                       k6);
#pragma unroll
        for (int j = 0; j < 6; ++j) k6s[j] = k6[j];
    }
    __syncthreads();
    for (int i = t; i < H * 6; i += LINES_THREADS) coef[(i % 6) * (long long)N + (i / 6) * C + ch] = k6s[i % 6];
    if (ch == 0)
        for (int i = t; i < (N - H * C) * 8; i += LINES_THREADS) coef[(i % 8) * (long long)N + H * C + i / 8] = 0.0f;
}

__global__ void bn_eval_coef_lines_kernel(int N, int H, int C, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float eps, const float* __restrict__ state,
                                          int lds, float* __restrict__ coef) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    const bool live = col < H * C;
    const int ch = col % C;
    const float inv = live ? rsqrtf(state[1 * (long long)lds + ch] + eps) : 0.0f;
    const float A = live ? gamma[ch] * inv : 0.0f;
    coef[0 * (long long)N + col] = A;
    coef[1 * (long long)N + col] = live ? beta[ch] - state[ch] * A : 0.0f;
    coef[2 * (long long)N + col] = live ? state[ch] : 0.0f;
    coef[3 * (long long)N + col] = inv;
    coef[4 * (long long)N + col] = live ? 1.0f : 0.0f;
    coef[5 * (long long)N + col] = 0.0f;
}

// backward totals of one channel over its H columns: dgamma, dbeta (per channel) and the two means of the dz kernel
__global__ void __launch_bounds__(LINES_THREADS) bn_bwd_finish_lines_kernel(
    const float* __restrict__ partial, int splits, float count, int N, int H, int C, float* __restrict__ coef,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float s1s[LINES_THREADS], s2s[LINES_THREADS];
    const int ch = blockIdx.x, t = threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    for (int i = t; i < splits * H; i += LINES_THREADS) {
        const int split = i / H, col = (i % H) * C + ch;
        s1 += partial[((long long)split * 2 + 0) * N + col];
        s2 += partial[((long long)split * 2 + 1) * N + col];
    }
    s1s[t] = s1; s2s[t] = s2;
    __syncthreads();
    for (int s = LINES_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) { s1s[t] += s1s[t + s]; s2s[t] += s2s[t + s]; }
        __syncthreads();
    }
    s1 = s1s[0]; s2 = s2s[0];
    if (t == 0) {
        if (dgamma) atomicAdd(dgamma + ch, coef[4 * (long long)N + ch] * s2 + coef[5 * (long long)N + ch] * s1);
        if (dbeta) atomicAdd(dbeta + ch, s1);
    }
    for (int i = t; i < H; i += LINES_THREADS) {
        coef[6 * (long long)N + i * C + ch] = s1 / count;
        coef[7 * (long long)N + i * C + ch] = s2 / count;
    }
}

// Virtual batch normalisation (utils/bnorm.py:11-69): statistics are the per-channel mean and mean of squares over
// (batch, time).  Reference pass (ref NULL, weight 1): the batch's own; live pass: blended with the reference batch's,
//   mean = w mean_B + (1 - w) mean_ref,  mean_sq = w msq_B + (1 - w) msq_ref,  w = 1 / (reference batch size + 1)
//   std = sqrt(eps + mean_sq - mean^2),  y = (x - mean) / std * gamma + beta
// coef rows as for batch_norm: A = gamma / std, B = beta - mean A, mean, 1 / std, r = 1, d = 0.  stats_out (optional)
// receives (mean_B, msq_B) -- what a reference pass hands to the live passes.
__global__ void vbn_finish_kernel(const float* __restrict__ partial, int splits, long long rows, int N,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float w,
                                  const float* __restrict__ ref, float* __restrict__ stats_out, float* __restrict__ coef) {
    const int c = blockIdx.x * FIN_COLS + threadIdx.x;
    float na, ma, qa;
    if (!merge_moments(partial, splits, N, c, na, ma, qa)) return;
    const float mean_b = ma, msq_b = qa / (float)rows + ma * ma;
    float mean = mean_b, msq = msq_b;
    if (ref) {
        mean = w * mean_b + (1.0f - w) * ref[c];
        msq = w * msq_b + (1.0f - w) * ref[(long long)N + c];
    }
    const float stddev = sqrtf(eps + msq - mean * mean);
    const float A = gamma[c] / stddev;
    coef[0 * (long long)N + c] = A;
    coef[1 * (long long)N + c] = beta[c] - mean * A;
    coef[2 * (long long)N + c] = mean;
    coef[3 * (long long)N + c] = 1.0f / stddev;
    coef[4 * (long long)N + c] = 1.0f;
    coef[5 * (long long)N + c] = 0.0f;
    if (stats_out) { stats_out[c] = mean_b; stats_out[(long long)N + c] = msq_b; }
}

__global__ void bn_eval_coef_kernel(int N, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                    const float* __restrict__ state, float* __restrict__ coef) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const float inv = rsqrtf(state[1 * (long long)N + c] + eps);
    const float A = gamma[c] * inv;
    coef[0 * (long long)N + c] = A;
    coef[1 * (long long)N + c] = beta[c] - state[c] * A;
    coef[2 * (long long)N + c] = state[c];
    coef[3 * (long long)N + c] = inv;
    coef[4 * (long long)N + c] = 1.0f;
    coef[5 * (long long)N + c] = 0.0f;
}

// Every thread owns one group of 4 columns for the whole launch (the grid is sized so that the thread count is a
// multiple of N / 4): coefficients live in registers, no per-element index division, rows advance by a constant.
template <bool O16, bool O32>
__global__ void __launch_bounds__(256) affine_act_drop_kernel(const float* __restrict__ z, int ldz, long long rows,
                                                              int N, const float* __restrict__ A,
                                                              const float* __restrict__ Bc, int act, uint32_t thr24,
                                                              float inv_keep, const unsigned long long* __restrict__ rng,
                                                              unsigned salt, uint16_t* __restrict__ out, int ldo,
                                                              float* __restrict__ out32, int ldo32, int bf, int S,
                                                              int L) {
    const int n4 = N >> 2;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long rstep = ((long long)gridDim.x * blockDim.x) / n4;
    const int c = (int)(tid % n4) * 4;
    uint64_t key = 0;
    if (thr24 < (1u << 24)) key = drop_key(rng, salt);
    const float4 a = A ? *reinterpret_cast<const float4*>(A + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b = *reinterpret_cast<const float4*>(Bc + c);
#pragma unroll 4
    for (long long r = tid / n4; r < rows; r += rstep) {
        const float4 v = *reinterpret_cast<const float4*>(z + r * ldz + c);
        float y[4] = {fmaf(v.x, a.x, b.x), fmaf(v.y, a.y, b.y), fmaf(v.z, a.z, b.z), fmaf(v.w, a.w, b.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = row_live(r, S, L) ? act_apply(y[k], act) : 0.0f;
        if (thr24 < (1u << 24)) {
            bool keep[4];
            const uint64_t idx = (uint64_t)r * (uint64_t)N + (uint64_t)c;
            drop_keep2(key, idx, thr24, keep[0], keep[1]);
            drop_keep2(key, idx + 2, thr24, keep[2], keep[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = keep[k] ? y[k] * inv_keep : 0.0f;
        }
        if (O16) {
            uint2 o;
            o.x = pack2(y[0], y[1], bf);
            o.y = pack2(y[2], y[3], bf);
            *reinterpret_cast<uint2*>(out + r * ldo + c) = o;
        }
        if (O32) *reinterpret_cast<float4*>(out32 + r * ldo32 + c) = make_float4(y[0], y[1], y[2], y[3]);
    }
}

// column totals of the backward partials; parameter gradients; means for the dz kernel
// stat_w: weight of THIS batch in the statistics the normalisation differentiates through (1 for batch_norm; the live
// pass of virtual batch norm blends the batch with a fixed reference: utils/bnorm.py:42-47)
__global__ void bn_bwd_finish_kernel(const float* __restrict__ partial, int splits, long long rows, int N, int bn,
                                     float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     float stat_w) {
    const int c = blockIdx.x * FIN_COLS + threadIdx.x;
    float s1, s2;
    if (!merge_sums(partial, splits, N, c, s1, s2)) return;
    if (bn) {
        // y = (x_hat r + d) gamma + beta with r, d under stop_gradient
        // atomics: the D(labels) and D(G(x)) backward passes of one update run on two streams (two addends: order-free)
        if (dgamma) atomicAdd(dgamma + c, coef[4 * (long long)N + c] * s2 + coef[5 * (long long)N + c] * s1);
        coef[6 * (long long)N + c] = stat_w * s1 / (float)rows;
        coef[7 * (long long)N + c] = stat_w * s2 / (float)rows;
    }
    if (dbeta) atomicAdd(dbeta + c, s1);
}

template <bool O16, bool O32>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ z, int ldz, long long rows, int N,
                                                           BnBwdIn p, int bn, const float* __restrict__ m1,
                                                           const float* __restrict__ m2, uint16_t* __restrict__ dz,
                                                           int lddz, float* __restrict__ dz32, int lddz32) {
    const int n4 = N >> 2;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long rstep = ((long long)gridDim.x * blockDim.x) / n4;
    const int c = (int)(tid % n4) * 4;
    uint64_t key = 0;
    if (p.thr24 < (1u << 24)) key = drop_key(p.rng, p.salt);
    const Col4 k4 = load_col4(p, c);
    float u1[4] = {0.f, 0.f, 0.f, 0.f}, u2[4] = {0.f, 0.f, 0.f, 0.f};
    if (bn) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { u1[k] = m1[c + k]; u2[k] = m2[c + k]; }
    }
#pragma unroll 4
    for (long long r = tid / n4; r < rows; r += rstep) {
        const float4 v = *reinterpret_cast<const float4*>(z + r * ldz + c);
        const uint2 raw = *reinterpret_cast<const uint2*>(p.da + r * p.ldda + c);
        float g[4], xh[4];
        bwd_elem4(p, k4, v, raw, r, c, N, key, g, xh);
        if (bn) {
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] = k4.A[k] * (g[k] - u1[k] - xh[k] * u2[k]);
        }
        if (!row_live(r, p.S, p.L)) g[0] = g[1] = g[2] = g[3] = 0.0f;
        if (O16) {
            uint2 o;
            o.x = pack2(g[0], g[1], p.bf);
            o.y = pack2(g[2], g[3], p.bf);
            *reinterpret_cast<uint2*>(dz + r * lddz + c) = o;
        }
        if (O32) *reinterpret_cast<float4*>(dz32 + r * lddz32 + c) = make_float4(g[0], g[1], g[2], g[3]);
    }
}

__global__ void rng_tick_kernel(unsigned long long* rng) { rng[1] += 1ull; }

template <int MODE>
void launch_partial(int tx, int splits, cudaStream_t st, const float* z, int ldz, long long rows, int N,
                    const BnBwdIn& p, float* scratch) {
    const dim3 grid((N + tx * 4 - 1) / (tx * 4), splits), block(tx, 256 / tx);
    if (tx == 128) bn_partial_kernel<MODE, 128><<<grid, block, 0, st>>>(z, ldz, rows, N, p, scratch);
    else if (tx == 64) bn_partial_kernel<MODE, 64><<<grid, block, 0, st>>>(z, ldz, rows, N, p, scratch);
    else bn_partial_kernel<MODE, 32><<<grid, block, 0, st>>>(z, ldz, rows, N, p, scratch);
}

inline uint32_t keep_threshold(float keep_prob) {
    if (!(keep_prob < 1.0f)) return 1u << 24;
    double t = (double)keep_prob * 16777216.0;
    if (t < 0) t = 0;
    return (uint32_t)t;
}

inline long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }

// blocks of 256 threads whose total is a multiple of n4 (= N / 4), about 8 per SM, at most one thread per item
inline int ew_grid(long long rows, int n4, int num_sms) {
    const long long unit = n4 / gcd_ll(n4, 256);            // blocks per whole number of rows
    long long want = (rows * n4 + 255) / 256;
    const long long cap = (long long)num_sms * 8;
    if (want > cap) want = cap;
    long long blocks = (want + unit - 1) / unit * unit;
    if (blocks < unit) blocks = unit;
    return (int)blocks;
}

}  // namespace

extern "C" int rsr_bn_train_stats(rsr_handle* h, void* stream, const float* z, int ldz, long long rows, int N,
                                  const float* gamma, const float* beta, float eps, float* state, float momentum,
                                  float renorm_momentum, int update_state, float* coef, float* scratch) {
    if (!h || !z || !gamma || !beta || !state || !coef || !scratch || rows <= 0 || N <= 0) return RSR_E_ARG;
    if ((N & 3) || (ldz & 3)) return RSR_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = bn_splits(rows, N, h->num_sms);
    BnBwdIn none = {};
    launch_partial<0>(bn_tx(N), splits, st, z, ldz, rows, N, none, scratch);
    RSR_LAUNCH_CHECK();
    bn_finish_train_kernel<<<(N + FIN_COLS - 1) / FIN_COLS, dim3(FIN_COLS, FIN_LANES), 0, st>>>(scratch, splits, rows, N, gamma, beta, eps, state, momentum,
                                                            renorm_momentum, update_state, coef);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_bn_train_finish(rsr_handle* h, void* stream, int splits, long long rows, int N, const float* gamma,
                                   const float* beta, float eps, float* state, float momentum, float renorm_momentum,
                                   int update_state, float* coef, const float* scratch) {
    if (!h || !gamma || !beta || !state || !coef || !scratch || rows <= 0 || N <= 0) return RSR_E_ARG;
    if ((N & 3) || splits <= 0 || splits > BN_MAX_SPLITS) return RSR_E_SHAPE;
    bn_finish_train_kernel<<<(N + FIN_COLS - 1) / FIN_COLS, dim3(FIN_COLS, FIN_LANES), 0, (cudaStream_t)stream>>>(
        scratch, splits, rows, N, gamma, beta, eps, state, momentum, renorm_momentum, update_state, coef);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_bn_eval_coef(rsr_handle* h, void* stream, int N, const float* gamma, const float* beta, float eps,
                                const float* state, float* coef) {
    if (!h || !gamma || !beta || !state || !coef || N <= 0) return RSR_E_ARG;
    bn_eval_coef_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, gamma, beta, eps, state, coef);
    RSR_LAUNCH_CHECK();
    return 0;
}

static int affine_impl(rsr_handle* h, void* stream, const float* z, int ldz, long long rows, int N, const float* A,
                       const float* Bc, int act, float keep_prob, const unsigned long long* rng, unsigned salt, void* out16,
                       int ld16, float* out32, int ld32, int S, int L) {
    if (!h || !z || !Bc || (!out16 && !out32) || rows <= 0 || N <= 0) return RSR_E_ARG;
    if ((N & 3) || (ldz & 3) || (out16 && (ld16 & 3)) || (out32 && (ld32 & 3))) return RSR_E_SHAPE;
    if (act != RSR_ACT_NONE && act != RSR_ACT_RELU && act != RSR_ACT_LRELU) return RSR_E_SHAPE;
    const uint32_t thr = keep_threshold(keep_prob);
    if (thr < (1u << 24) && (!rng || thr == 0)) return RSR_E_ARG;
    const int grid = ew_grid(rows, N >> 2, h->num_sms);
    const float ik = thr < (1u << 24) ? 1.0f / keep_prob : 1.0f;
    const int bf = h->dtype == RSR_DTYPE_BF16;
    cudaStream_t st = (cudaStream_t)stream;
#define RSR_AFFINE(O16, O32) affine_act_drop_kernel<O16, O32><<<grid, 256, 0, st>>>( \
        z, ldz, rows, N, A, Bc, act, thr, ik, rng, salt, (uint16_t*)out16, ld16, out32, ld32, bf, S, L)
    if (out16 && out32) RSR_AFFINE(true, true);
    else if (out16) RSR_AFFINE(true, false);
    else RSR_AFFINE(false, true);
#undef RSR_AFFINE
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_affine_act_drop(rsr_handle* h, void* stream, const float* z, int ldz, long long rows, int N,
                                   const float* A, const float* Bc, int act, float keep_prob,
                                   const unsigned long long* rng, unsigned salt, void* out16, int ld16, float* out32,
                                   int ld32) {
    return affine_impl(h, stream, z, ldz, rows, N, A, Bc, act, keep_prob, rng, salt, out16, ld16, out32, ld32, 0, 0);
}

static int bn_bwd_impl(rsr_handle* h, void* stream, const void* da16, int ldda, const float* z, int ldz,
                       long long rows, int N, int act, float keep_prob, const unsigned long long* rng, unsigned salt,
                       int bn, float* coef, const float* bias, float* dgamma, float* dbeta, void* dz16, int lddz,
                       float* dz32, int lddz32, float* scratch, float stat_w, int S = 0, int L = 0, int H = 0, int C = 0) {
    if (!h || !da16 || !z || !scratch || rows <= 0 || N <= 0) return RSR_E_ARG;
    if (bn ? !coef : !bias) return RSR_E_ARG;
    if ((N & 3) || (ldz & 3) || (ldda & 3) || (dz16 && (lddz & 3)) || (dz32 && (lddz32 & 3))) return RSR_E_SHAPE;
    if (act != RSR_ACT_NONE && act != RSR_ACT_RELU && act != RSR_ACT_LRELU) return RSR_E_SHAPE;
    const uint32_t thr = keep_threshold(keep_prob);
    if (thr < (1u << 24) && (!rng || thr == 0)) return RSR_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    BnBwdIn p;
    p.da = (const uint16_t*)da16; p.ldda = ldda;
    if (bn) { p.A = coef; p.Bc = coef + (long long)N; p.mean = coef + 2ll * N; p.inv_std = coef + 3ll * N; }
    else { p.A = nullptr; p.Bc = bias; p.mean = nullptr; p.inv_std = nullptr; }
    p.act = act; p.thr24 = thr; p.inv_keep = thr < (1u << 24) ? 1.0f / keep_prob : 1.0f; p.rng = rng; p.salt = salt;
    p.bf = h->dtype == RSR_DTYPE_BF16;
    p.S = S; p.L = L;
    const int splits = bn_splits(rows, N, h->num_sms);
    if (bn || dbeta) {
        launch_partial<1>(bn_tx(N), splits, st, z, ldz, rows, N, p, scratch);
        RSR_LAUNCH_CHECK();
        if (H > 0)
            bn_bwd_finish_lines_kernel<<<C, LINES_THREADS, 0, st>>>(scratch, splits, (float)(rows / S) * (float)(L * H), N, H, C,
                                                                    coef, dgamma, dbeta);
        else
            bn_bwd_finish_kernel<<<(N + FIN_COLS - 1) / FIN_COLS, dim3(FIN_COLS, FIN_LANES), 0, st>>>(scratch, splits, rows, N, bn, coef, dgamma, dbeta, stat_w);
        RSR_LAUNCH_CHECK();
    }
    if (dz16 || dz32) {
        const int grid = ew_grid(rows, N >> 2, h->num_sms);
#define RSR_APPLY(O16, O32) bn_bwd_apply_kernel<O16, O32><<<grid, 256, 0, st>>>( \
        z, ldz, rows, N, p, bn, bn ? coef + 6ll * N : nullptr, bn ? coef + 7ll * N : nullptr, (uint16_t*)dz16, lddz, \
        dz32, lddz32)
        if (dz16 && dz32) RSR_APPLY(true, true);
        else if (dz16) RSR_APPLY(true, false);
        else RSR_APPLY(false, true);
#undef RSR_APPLY
        RSR_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int rsr_bn_bwd(rsr_handle* h, void* stream, const void* da16, int ldda, const float* z, int ldz,
                          long long rows, int N, int act, float keep_prob, const unsigned long long* rng, unsigned salt,
                          int bn, float* coef, const float* bias, float* dgamma, float* dbeta, void* dz16, int lddz,
                          float* dz32, int lddz32, float* scratch) {
    return bn_bwd_impl(h, stream, da16, ldda, z, ldz, rows, N, act, keep_prob, rng, salt, bn, coef, bias, dgamma, dbeta, dz16,
                       lddz, dz32, lddz32, scratch, 1.0f);
}

// ---- batch_norm behind the convolutions of the frame layout (see bn_finish_train_lines_kernel) ----
static int lines_check(long long frames, int S, int L, int H, int C, int N) {
    if (frames <= 0 || S <= 0 || L <= 0 || H <= 0 || C <= 0) return RSR_E_ARG;
    if (S < L || N < H * C || (N & 3) || frames * S >= (1LL << 31)) return RSR_E_SHAPE;
    return 0;
}

extern "C" int rsr_bn_train_stats_lines(rsr_handle* h, void* stream, const float* z, int ldz, long long frames, int S, int L,
                                        int H, int C, int N, const float* gamma, const float* beta, float eps, float* state,
                                        int lds, float momentum, float renorm_momentum, int update_state, float* coef,
                                        float* scratch) {
    if (!h || !z || !gamma || !beta || !state || !coef || !scratch) return RSR_E_ARG;
    if (int e = lines_check(frames, S, L, H, C, N)) return e;
    if ((ldz & 3) || lds < C) return RSR_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = frames * S;
    const int splits = bn_splits(rows, N, h->num_sms);
    BnBwdIn rows_only = {};
    rows_only.S = S; rows_only.L = L;
    launch_partial<0>(bn_tx(N), splits, st, z, ldz, rows, N, rows_only, scratch);
    RSR_LAUNCH_CHECK();
    bn_finish_train_lines_kernel<<<C, LINES_THREADS, 0, st>>>(scratch, splits, (float)frames * (float)(L * H), N, H, C, gamma,
                                                              beta, eps, state, lds, momentum, renorm_momentum, update_state,
                                                              coef);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_bn_eval_coef_lines(rsr_handle* h, void* stream, int N, int H, int C, const float* gamma, const float* beta,
                                      float eps, const float* state, int lds, float* coef) {
    if (!h || !gamma || !beta || !state || !coef || N <= 0 || H <= 0 || C <= 0) return RSR_E_ARG;
    if (N < H * C || lds < C) return RSR_E_SHAPE;
    bn_eval_coef_lines_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, H, C, gamma, beta, eps, state, lds, coef);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_affine_act_lines(rsr_handle* h, void* stream, const float* z, int ldz, long long frames, int S, int L,
                                    int N, const float* A, const float* Bc, int act, void* out16, int ld16) {
    if (frames <= 0 || S <= 0 || L <= 0 || S < L || frames * S >= (1LL << 31)) return RSR_E_ARG;
    return affine_impl(h, stream, z, ldz, frames * S, N, A, Bc, act, 1.0f, nullptr, 0u, out16, ld16, nullptr, 0, S, L);
}

extern "C" int rsr_bn_bwd_lines(rsr_handle* h, void* stream, const void* da16, int ldda, const float* z, int ldz,
                                long long frames, int S, int L, int H, int C, int N, int act, float* coef, float* dgamma,
                                float* dbeta, void* dz16, int lddz, float* scratch) {
    if (int e = lines_check(frames, S, L, H, C, N)) return e;
    if (!dz16) return RSR_E_ARG;
    return bn_bwd_impl(h, stream, da16, ldda, z, ldz, frames * S, N, act, 1.0f, nullptr, 0u, 1, coef, nullptr, dgamma, dbeta,
                       dz16, lddz, nullptr, 0, scratch, 1.0f, S, L, H, C);
}

extern "C" int rsr_vbn_stats(rsr_handle* h, void* stream, const float* z, int ldz, long long rows, int N,
                             const float* gamma, const float* beta, float eps, float batch_weight, const float* ref_stats,
                             float* stats_out, float* coef, float* scratch) {
    if (!h || !z || !gamma || !beta || !coef || !scratch || rows <= 0 || N <= 0) return RSR_E_ARG;
    if ((N & 3) || (ldz & 3)) return RSR_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = bn_splits(rows, N, h->num_sms);
    BnBwdIn none = {};
    launch_partial<0>(bn_tx(N), splits, st, z, ldz, rows, N, none, scratch);
    RSR_LAUNCH_CHECK();
    vbn_finish_kernel<<<(N + FIN_COLS - 1) / FIN_COLS, dim3(FIN_COLS, FIN_LANES), 0, st>>>(scratch, splits, rows, N, gamma, beta, eps,
                                                                                         batch_weight, ref_stats, stats_out, coef);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_vbn_bwd(rsr_handle* h, void* stream, const void* da16, int ldda, const float* z, int ldz,
                           long long rows, int N, int act, float batch_weight, float* coef, float* dgamma, float* dbeta,
                           void* dz16, int lddz, float* dz32, int lddz32, float* scratch) {
    return bn_bwd_impl(h, stream, da16, ldda, z, ldz, rows, N, act, 1.0f, nullptr, 0u, 1, coef, nullptr, dgamma, dbeta, dz16,
                       lddz, dz32, lddz32, scratch, batch_weight);
}

// Gaussian draws from the same counter-based stream (the discriminator's input noise, utils/ops.py:19-30): element i of
// the draw named (seed, tick, salt) is Box-Muller of the two 24-bit fields of splitmix64(key ^ i)
__global__ void __launch_bounds__(256) gauss_noise_kernel(const unsigned long long* __restrict__ rng, unsigned salt,
                                                          float* __restrict__ out, long long n, float stddev) {
    const uint64_t key = drop_key(rng, salt);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint64_t hsh = splitmix64(key ^ (uint64_t)i);
        const float u1 = ((float)(uint32_t)(hsh >> 40) + 1.0f) * (1.0f / 16777216.0f);          // (0, 1]
        const float u2 = (float)((uint32_t)(hsh >> 16) & 0xffffffu) * (1.0f / 16777216.0f);     // [0, 1)
        out[i] = stddev * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
}

extern "C" int rsr_gauss_noise(rsr_handle* h, void* stream, const unsigned long long* rng, unsigned salt, float* out,
                               long long n, float stddev) {
    if (!h || !rng || !out || n <= 0) return RSR_E_ARG;
    long long blocks = (n + 255) / 256;
    if (blocks > 4LL * h->num_sms) blocks = 4LL * h->num_sms;
    gauss_noise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rng, salt, out, n, stddev);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_rng_tick(rsr_handle* h, void* stream, unsigned long long* rng) {
    if (!h || !rng) return RSR_E_ARG;
    rng_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng);
    RSR_LAUNCH_CHECK();
    return 0;
}
