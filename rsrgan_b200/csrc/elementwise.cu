// HBM-bound kernels of the GAN step: input staging (CMVN + noise + time-major 16-bit
// conversion), LSGAN / MSE losses with their gradients (warp-shuffle reductions), bias-gradient
// column sums, and the fused clip_by_norm + SGD|Adam + EMA + 16-bit re-pack sweep.
// All are coalesced, vectorised where the layout allows, and sized in multiples of the SM count.
#include <cstring>

#include "common.cuh"
#include "handle.h"

using namespace rsr;

namespace {

inline int grid_for(long long work_items, int threads, int num_sms, int per_sm = 8) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)num_sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------------------
// staging: (B,T,D) fp32 batch-major -> [T*B, ld] time-major, 16-bit (+ fp32 copy)
// ---------------------------------------------------------------------------------------
// one output row (= one frame) per warp pass: no per-element division, coalesced 4-byte reads / 2-byte writes
__global__ void stage_input_kernel(const float* __restrict__ x, int ldx, int time_major_in, int B, int T, int D,
                                   const float* __restrict__ mean, const float* __restrict__ istd,
                                   const float* __restrict__ noise, uint16_t* __restrict__ out16, int ld16,
                                   float* __restrict__ out32, int ld32, int bf) {
    const long long rows = (long long)B * T;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < rows; r += nwarps) {          // output row = t*B + b
        const int b = (int)(r % B);
        const int t = (int)(r / B);
        const float* src = time_major_in ? x + r * ldx : x + ((long long)b * T + t) * ldx;
        const float* nz = noise ? noise + (long long)b * D : nullptr;
        for (int d = lane; d < D; d += 32) {
            float v = src[d];
            if (mean) v = (v - mean[d]) * istd[d];
            if (nz) v += nz[d];
            if (out16) out16[r * ld16 + d] = f2h(v, bf);
            if (out32) out32[r * ld32 + d] = v;
        }
    }
}

__global__ void unstage_output_kernel(const float* __restrict__ y, int ld, int B, int T, int D,
                                      const float* __restrict__ mean, const float* __restrict__ std,
                                      float* __restrict__ out) {
    const long long total = (long long)B * T * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const long long r = i / D;          // batch-major output row = b*T + t
        const int t = (int)(r % T);
        const int b = (int)(r / T);
        float v = y[((long long)t * B + b) * ld + d];
        if (mean) v = v * std[d] + mean[d];
        out[i] = v;
    }
}

// CMVN on (N, D) fp32: mode 0: (x - mean) / std ; mode 1: y * std + mean.  An HBM stream: the matrix is read and written
// as 16-byte vectors of the FLAT array (D = 257 or 40: rows are not 16-byte aligned, the flat array is), the feature
// index of a vector's first element comes from one division and is carried with a wrap for the other three, and the
// statistics sit in shared memory.  The tail (n % 4 elements) is handled by the last thread.
__global__ void __launch_bounds__(256) cmvn_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                   const float* __restrict__ std, long long n, int D, int mode,
                                                   float* __restrict__ out) {
    extern __shared__ float cm_sh[];
    float* s_mean = cm_sh;
    float* s_std = cm_sh + D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { s_mean[d] = mean[d]; s_std[d] = std[d]; }
    __syncthreads();
    const long long nv = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4* xv = reinterpret_cast<const float4*>(x);
    float4* ov = reinterpret_cast<float4*>(out);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 v = xv[i];
        int d = (int)((i << 2) % D);
        float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k] = mode ? fmaf(r[k], s_std[d], s_mean[d]) : (r[k] - s_mean[d]) / s_std[d];
            if (++d == D) d = 0;
        }
        ov[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = nv << 2; i < n; ++i) {
            const int d = (int)(i % D);
            out[i] = mode ? fmaf(x[i], s_std[d], s_mean[d]) : (x[i] - s_mean[d]) / s_std[d];
        }
}

// The loader's CMVN on a PADDED minibatch (B, T, D) fp32, as the reference applies it per utterance BEFORE padding
// (io_funcs/make_tfrecords.py:84-87: float64 (x - mean) / stddev, then astype(float32); the zero padding of
// io_funcs/tfrecords_dataset.py:149-152 comes after, so padded frames are exact zeros of the NORMALISED domain):
//   out[b, t, d] = t < length[b] ? float((double) x[b, t, d] - mean[d]) / std[d]) : 0
// float64 arithmetic with the float64 statistics of train_cmvn.npz: bit-identical to the reference's numpy.  Same
// vector stream as above.
__global__ void __launch_bounds__(256) cmvn_padded_kernel(const float* __restrict__ x, const int* __restrict__ lengths,
                                                          const double* __restrict__ mean, const double* __restrict__ std,
                                                          long long n, int T, int D, float* __restrict__ out) {
    extern __shared__ double cmd_sh[];
    double* s_mean = cmd_sh;
    double* s_std = cmd_sh + D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) { s_mean[d] = mean[d]; s_std[d] = std[d]; }
    __syncthreads();
    const long long nv = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4* xv = reinterpret_cast<const float4*>(x);
    float4* ov = reinterpret_cast<float4*>(out);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 v = xv[i];
        long long row = (i << 2) / D;                  // b * T + t
        int d = (int)((i << 2) - row * D);
        int t = (int)(row % T);
        int len = lengths[row / T];
        float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k] = t < len ? (float)(((double)r[k] - s_mean[d]) / s_std[d]) : 0.0f;
            if (++d == D) {
                d = 0; ++row;
                if (++t == T) { t = 0; if (row < n / D) len = lengths[row / T]; }
            }
        }
        ov[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = nv << 2; i < n; ++i) {
            const long long row = i / D;
            const int d = (int)(i - row * D);
            out[i] = (int)(row % T) < lengths[row / T] ? (float)(((double)x[i] - s_mean[d]) / s_std[d]) : 0.0f;
        }
}

// ---------------------------------------------------------------------------------------
// LSGAN + MSE losses and gradients
// ---------------------------------------------------------------------------------------
struct LossParams {
    const float* rl; const float* fk; int ldl; long long n_logit; int clip;
    const float* g; int ldg; const float* y; int ldy; long long n_frames; int D;
    float d_real, d_fake, lambda, gscale;
    float* losses;
    uint16_t* d_rl_grad; uint16_t* d_fk_grad; uint16_t* g_adv_grad; int ldgrad;
    float* dg_mse; int lddg;
    int bf;
};

__device__ __forceinline__ void block_atomic_add(float v, float* dst, float* sh) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float s = lane < (blockDim.x >> 5) ? sh[lane] : 0.0f;
        s = warp_sum(s);
        if (lane == 0) atomicAdd(dst, s);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) lsgan_mse_kernel(const LossParams p) {
    __shared__ float sh[8];
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    float s_rl = 0.f, s_fk = 0.f, s_adv = 0.f, s_mse = 0.f;
    const float inv_nl = p.n_logit > 0 ? 1.0f / (float)p.n_logit : 0.f;
    for (long long i = tid; i < p.n_logit; i += nth) {
        if (p.rl) {
            const float u = p.rl[i * p.ldl];
            const float l = p.clip ? fminf(fmaxf(u, -0.5f), 1.5f) : u;
            const float in = (!p.clip || (u >= -0.5f && u <= 1.5f)) ? 1.f : 0.f;
            const float e = l - p.d_real;
            s_rl += e * e;
            if (p.d_rl_grad) p.d_rl_grad[i * p.ldgrad] = f2h(p.gscale * 2.f * e * inv_nl * in, p.bf);
        }
        if (p.fk) {
            const float u = p.fk[i * p.ldl];
            const float l = p.clip ? fminf(fmaxf(u, -0.5f), 1.5f) : u;
            const float in = (!p.clip || (u >= -0.5f && u <= 1.5f)) ? 1.f : 0.f;
            const float ef = l - p.d_fake, er = l - p.d_real;
            s_fk += ef * ef;
            s_adv += er * er;
            if (p.d_fk_grad) p.d_fk_grad[i * p.ldgrad] = f2h(p.gscale * 2.f * ef * inv_nl * in, p.bf);
            if (p.g_adv_grad) p.g_adv_grad[i * p.ldgrad] = f2h(p.gscale * 2.f * er * inv_nl * in, p.bf);
        }
    }
    if (p.g && p.y) {
        // g_mse = 0.5 * D * mean((g-y)^2)  ->  d/dg = lambda * 0.5 * D * 2 (g-y) / (n_frames*D) = lambda (g-y)/n_frames
        const long long total = p.n_frames * p.D;
        const float cg = p.gscale * p.lambda / (float)p.n_frames;
        const bool flat = p.ldg == p.D && p.ldy == p.D && (!p.dg_mse || p.lddg == p.D) && (total & 3) == 0 &&
                          (((uintptr_t)p.g | (uintptr_t)p.y | (uintptr_t)p.dg_mse) & 15) == 0;
        if (flat) {
            // dense rows (the generator output is stored at its own width): one 16-byte vector stream over g, y and dg,
            // no index arithmetic per element
            const float4* gv = reinterpret_cast<const float4*>(p.g);
            const float4* yv = reinterpret_cast<const float4*>(p.y);
            float4* dv = reinterpret_cast<float4*>(p.dg_mse);
            for (long long i = tid; i < (total >> 2); i += nth) {
                const float4 a = gv[i], b = yv[i];
                const float e0 = a.x - b.x, e1 = a.y - b.y, e2 = a.z - b.z, e3 = a.w - b.w;
                s_mse += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
                if (dv) dv[i] = make_float4(cg * e0, cg * e1, cg * e2, cg * e3);
            }
        } else
        for (long long i = tid; i < total; i += nth) {
            const int d = (int)(i % p.D);
            const long long r = i / p.D;
            const float e = p.g[r * p.ldg + d] - p.y[r * p.ldy + d];
            s_mse += e * e;
            if (p.dg_mse) p.dg_mse[r * p.lddg + d] = cg * e;
        }
    }
    if (p.rl) block_atomic_add(s_rl * inv_nl, p.losses + 0, sh);
    if (p.fk) { block_atomic_add(s_fk * inv_nl, p.losses + 1, sh); block_atomic_add(s_adv * inv_nl, p.losses + 2, sh); }
    if (p.g && p.y) block_atomic_add(s_mse * 0.5f / (float)p.n_frames, p.losses + 3, sh);
}

// ---------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] (+)= sum_m x[m, n]
// block = 32 x 8 threads handles 32 columns over a slab of rows; atomics combine slabs.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int ld, long long M, int N,
                                                     float* __restrict__ out, int bf, long long rows_per_block) {
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const long long r0 = blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    float s = 0.f;
    if (n < N) {
        for (long long r = r0 + ty; r < r1; r += 8) {
            if constexpr (sizeof(T) == 2) s += h2f(x[r * ld + n], bf);
            else s += x[r * ld + n];
        }
    }
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][tx];
        atomicAdd(out + n, t);
    }
}

// 16-bit, 16-byte vectorised variant (N and ld multiples of 8, 16-byte aligned base): thread <-> 8 consecutive
// columns, 32 row lanes per block; block = 256 columns x a slab of rows.
__global__ void __launch_bounds__(256) colsum16_vec_kernel(const uint16_t* __restrict__ x, int ld, long long M, int N,
                                                           float* __restrict__ out, int bf, long long rows_per_block) {
    __shared__ float sh[8][256 + 8];
    const int cg = threadIdx.x & 31, ry = threadIdx.x >> 5;       // column group (8 columns), row lane
    const int n0 = blockIdx.x * 256 + cg * 8;
    const long long r0 = blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n0 < N) {
        for (long long r = r0 + ry; r < r1; r += 8) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld + n0));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s[2 * k] += h2f((uint16_t)(w[k] & 0xFFFFu), bf);
                s[2 * k + 1] += h2f((uint16_t)(w[k] >> 16), bf);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sh[ry][cg * 8 + k] = s[k];
    __syncthreads();
    const int c = threadIdx.x;                                     // one column per thread
    if (blockIdx.x * 256 + c < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][c];
        atomicAdd(out + blockIdx.x * 256 + c, t);
    }
}

__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ src, int ld_src, int rows, int cols,
                                                          uint16_t* __restrict__ dst, int ld_dst) {
    __shared__ uint16_t tile[32][34];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = ty; i < 32; i += 8)
        tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? src[(size_t)(r0 + i) * ld_src + c0 + tx] : (uint16_t)0;
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) dst[(size_t)(c0 + i) * ld_dst + r0 + tx] = tile[tx][i];
}

__global__ void fill32_kernel(float* x, long long n, float v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void cast16_kernel(const float* __restrict__ x, long long n, uint16_t* __restrict__ o, int bf) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = f2h(x[i], bf);
}

// residual add: out32 = a + b, out16 = h16(a + b)   (models/res_lstm_l.py:116,127,138,187)
__global__ void add_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4,
                                float* __restrict__ o32, uint16_t* __restrict__ o16, int bf) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[i];
        const float4 y = reinterpret_cast<const float4*>(b)[i];
        const float4 v = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        if (o32) reinterpret_cast<float4*>(o32)[i] = v;
        if (o16) {
            uint2 o; o.x = pack2(v.x, v.y, bf); o.y = pack2(v.z, v.w, bf);
            reinterpret_cast<uint2*>(o16)[i] = o;
        }
    }
}

// L2 regulariser gradient on the flat buffer: grad += scale * theta on the 1024-blocks whose
// segment is flagged (non-bias tensors; models/gan_rnn_placeholder.py:253-258)
__global__ void __launch_bounds__(256) l2_grad_kernel(float* __restrict__ g, const float* __restrict__ theta,
                                                      const int* __restrict__ seg_id, const int* __restrict__ seg_flag,
                                                      float scale) {
    if (!seg_flag[seg_id[blockIdx.x]]) return;
    const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
    float4 q = *reinterpret_cast<float4*>(g + base);
    const float4 t = *reinterpret_cast<const float4*>(theta + base);
    q.x += scale * t.x; q.y += scale * t.y; q.z += scale * t.z; q.w += scale * t.w;
    *reinterpret_cast<float4*>(g + base) = q;
}

// ---------------------------------------------------------------------------------------
// update sweep. One block = 1024 contiguous elements of exactly one segment (tensor).
// ---------------------------------------------------------------------------------------
// Deterministic (no atomics): pass 1 leaves one partial sum per 1024-element block, pass 2 adds the partials of each
// segment in a fixed order.  Every data-parallel rank therefore derives bit-identical clip scales from the all-reduced
// gradients and the replicas' weights stay bit-identical (an atomicAdd version let them drift apart by ulps whenever a
// tensor's norm exceeded the clip threshold).
__global__ void __launch_bounds__(256) seg_sumsq_kernel(const float* __restrict__ g, float gmul,
                                                        float* __restrict__ partials) {
    __shared__ float sh[8];
    const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
    const float4 q = *reinterpret_cast<const float4*>(g + base);
    const float a = q.x * gmul, b = q.y * gmul, c = q.z * gmul, d = q.w * gmul;
    float s = a * a + b * b + c * c + d * d;
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) partials[blockIdx.x] = t;
    }
}
// one block per segment: its 1024-blocks are contiguous in seg_id (non-decreasing)
__global__ void __launch_bounds__(256) seg_sumsq_finish_kernel(const float* __restrict__ partials,
                                                               const int* __restrict__ seg_id, int n_blocks,
                                                               float* __restrict__ sumsq) {
    __shared__ float sh[256];
    __shared__ int range[2];
    const int seg = blockIdx.x;
    if (threadIdx.x < 2) {                       // first block with seg_id >= seg + threadIdx.x
        const int want = seg + (int)threadIdx.x;
        int lo = 0, hi = n_blocks;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (seg_id[mid] < want) lo = mid + 1; else hi = mid;
        }
        range[threadIdx.x] = lo;
    }
    __syncthreads();
    float s = 0.f;
    for (int i = range[0] + (int)threadIdx.x; i < range[1]; i += 256) s += partials[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sumsq[seg] = sh[0];
}

// Overflow guard of the 16-bit backward pass (fp16 operands under a static loss scale): a gradient tensor that picked
// up an inf / NaN has a non-finite norm.  The update sweep and the Adam tick then leave EVERYTHING untouched (weights,
// slots, shadows, beta powers -- the fp32 reference would have taken a finite, clipped step; skipping one update is the
// closest finite behaviour) and count the event in hyper[7]; the host halves the loss scale when the counter moves.
__device__ __forceinline__ bool any_nonfinite_norm(const float* __restrict__ sumsq, int n_seg) {
    bool bad = false;
    for (int s = (int)(threadIdx.x & 31); s < n_seg; s += 32) bad |= !isfinite(sumsq[s]);
    return __any_sync(0xffffffffu, bad);
}

__global__ void adam_tock_kernel(float* hyper, const float* __restrict__ sumsq, int n_seg) {
    if (any_nonfinite_norm(sumsq, n_seg)) return;
    if (threadIdx.x) return;
    hyper[6] = hyper[0] * sqrtf(1.f - hyper[5]) / (1.f - hyper[4]);   // lr_t of the step just applied (diagnostic)
    hyper[4] *= hyper[1];
    hyper[5] *= hyper[2];
}

template <int ADAM>
__global__ void __launch_bounds__(256) clip_update_kernel(const float* __restrict__ g, float gmul,
                                                          const int* __restrict__ seg_id,
                                                          const float* __restrict__ sumsq, float max_norm,
                                                          float* __restrict__ hyper, float ema_decay,
                                                          float* __restrict__ theta, float* __restrict__ m,
                                                          float* __restrict__ v, float* __restrict__ ema,
                                                          uint16_t* __restrict__ theta16, int bf, int n_seg) {
    if (any_nonfinite_norm(sumsq, n_seg)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) hyper[7] += 1.0f;
        return;
    }
    const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
    const float nrm = sqrtf(sumsq[seg_id[blockIdx.x]]);
    const float sc = gmul * (max_norm / fmaxf(nrm, max_norm));   // tf.clip_by_norm: g * clip / max(norm, clip)
    const float4 gq = *reinterpret_cast<const float4*>(g + base);
    float4 th = *reinterpret_cast<const float4*>(theta + base);
    float gg[4] = {gq.x * sc, gq.y * sc, gq.z * sc, gq.w * sc};
    float tt[4] = {th.x, th.y, th.z, th.w};
    if (ADAM) {
        // tf.train.AdamOptimizer keeps beta{1,2}_power as fp32 variables; lr_t is formed from the powers BEFORE they
        // are multiplied (they start at beta1 / beta2); adam_tock_kernel advances them after this sweep
        const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
        const float lr_t = hyper[0] * sqrtf(1.f - hyper[5]) / (1.f - hyper[4]);
        float4 mq = *reinterpret_cast<const float4*>(m + base);
        float4 vq = *reinterpret_cast<const float4*>(v + base);
        float mm[4] = {mq.x, mq.y, mq.z, mq.w}, vv[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mm[k] = b1 * mm[k] + (1.f - b1) * gg[k];
            vv[k] = b2 * vv[k] + (1.f - b2) * gg[k] * gg[k];
            tt[k] -= lr_t * mm[k] / (sqrtf(vv[k]) + eps);
        }
        *reinterpret_cast<float4*>(m + base) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(v + base) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
        const float lr = hyper[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) tt[k] -= lr * gg[k];
    }
    *reinterpret_cast<float4*>(theta + base) = make_float4(tt[0], tt[1], tt[2], tt[3]);
    if (ema) {
        float4 eq = *reinterpret_cast<const float4*>(ema + base);
        float ee[4] = {eq.x, eq.y, eq.z, eq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) ee[k] -= (1.f - ema_decay) * (ee[k] - tt[k]);
        *reinterpret_cast<float4*>(ema + base) = make_float4(ee[0], ee[1], ee[2], ee[3]);
    }
    if (theta16) {
        uint2 o;
        o.x = pack2(tt[0], tt[1], bf);
        o.y = pack2(tt[2], tt[3], bf);
        *reinterpret_cast<uint2*>(theta16 + base) = o;
    }
}


// ---------------------------------------------------------------------------------------
// 1-D convolution family (models/rced.py:90-101): the convolutions themselves run on the tcgen05
// GEMM over an OVERLAPPED strided view of the channels-last frame buffer (row m of the view =
// taps m-w/2 .. m+w/2, row pitch = one position), so no im2col matrix is ever written.  What
// remains here is the layout glue: frames -> padded channels-last rows, re-zeroing the SAME
// padding rows after each layer, and the flipped/transposed taps for the data gradient.
// Frame layout: frame r occupies rows [r*S, r*S+S) of a [*, Cp] 16-bit buffer, positions
// 0..L-1 hold data, rows L..S-1 are the zero padding shared by this frame's right edge and the
// next frame's left edge (S - L >= w/2).
// ---------------------------------------------------------------------------------------
__global__ void conv_stage_frames_kernel(const float* __restrict__ x, int ldx, int time_major_in, int B, int T, int L,
                                         int S, int Cp, const float* __restrict__ mean, const float* __restrict__ istd,
                                         uint16_t* __restrict__ out, int bf) {
    const int chunks = Cp / 8;
    const long long total = (long long)B * T * S * chunks;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % chunks);
        const long long row = i / chunks;        // r*S + p
        const int p = (int)(row % S);
        const long long r = row / S;             // output frame = t*B + b
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (c == 0 && p < L) {
            const int b = (int)(r % B);
            const int t = (int)(r / B);
            float v = time_major_in ? x[r * ldx + p] : x[((long long)b * T + t) * ldx + p];
            if (mean) v = (v - mean[p]) * istd[p];
            o.x = (uint32_t)f2h(v, bf);          // channel 0 = the spectrum bin, channels 1.. are zero padding
        }
        *reinterpret_cast<uint4*>(out + row * Cp + 8 * c) = o;
    }
}

__global__ void conv_mask_rows_kernel(uint16_t* __restrict__ buf, long long frames, int S, int L, int Cp) {
    const int chunks = Cp / 8, pad = S - L;
    const long long total = frames * pad * chunks;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % chunks);
        const long long q = i / chunks;
        const int p = L + (int)(q % pad);
        const long long r = q / pad;
        *reinterpret_cast<uint4*>(buf + (r * S + p) * Cp + 8 * c) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// out[k][co][ci] = w[W-1-k][ci][co]   (w: [W, Cin_p, Cout_p], out: [W, Cout_p, Cin_p])
__global__ void conv_w_flip_kernel(const uint16_t* __restrict__ w, int W, int Cin_p, int Cout_p,
                                   uint16_t* __restrict__ out) {
    const int total = W * Cin_p * Cout_p;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % Cin_p;
        const int co = (i / Cin_p) % Cout_p;
        const int k = i / (Cin_p * Cout_p);
        out[i] = w[((W - 1 - k) * Cin_p + ci) * Cout_p + co];
    }
}

// [splice, w] SAME convolution over H stacked lines (models/rced.py:94-101 with splice > 1) as the 1-D overlapped-view
// GEMM: the H lines of a frame are CHANNELS of one position (channel = line * C + c), and the compact TensorFlow filter
// w[kh][W][ci][co] expands to block-Toeplitz taps
//   out[k][h_in * ci + a][h_out * co + b] = w[h_in - h_out + kh / 2][k][a][b]      (zero outside the filter / the padding)
// (out: [W, cip, cop], cip / cop = H * ci / H * co padded to multiples of 8).  The expansion is a derived operand,
// rebuilt after every update like Wc; conv_toeplitz_fold_kernel sums the gradient of the expansion over the tied copies,
// one thread per compact element in a fixed order (deterministic).
__global__ void conv_toeplitz_expand_kernel(const uint16_t* __restrict__ w, int kh, int W, int ci, int co, int H,
                                            int cip, int cop, uint16_t* __restrict__ out) {
    const long long total = (long long)W * cip * cop;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % cop);
        const int a = (int)((i / cop) % cip);
        const int k = (int)(i / ((long long)cop * cip));
        uint16_t v = 0;
        if (a < H * ci && col < H * co) {
            const int h_in = a / ci, h_out = col / co;
            const int r = h_in - h_out + kh / 2;
            if (r >= 0 && r < kh) v = w[(((long long)r * W + k) * ci + (a - h_in * ci)) * co + (col - h_out * co)];
        }
        out[i] = v;
    }
}

__global__ void conv_toeplitz_fold_kernel(const float* __restrict__ dw2, int kh, int W, int ci, int co, int H, int cip,
                                          int cop, float* __restrict__ grad) {
    const long long total = (long long)kh * W * ci * co;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i % co);
        const int a = (int)((i / co) % ci);
        const int k = (int)((i / ((long long)co * ci)) % W);
        const int r = (int)(i / ((long long)co * ci * W));
        float s = 0.f;
        for (int h_out = 0; h_out < H; ++h_out) {
            const int h_in = h_out + r - kh / 2;
            if (h_in >= 0 && h_in < H) s += dw2[((long long)k * cip + h_in * ci + a) * cop + h_out * co + b];
        }
        grad[i] += s;
    }
}

// per-channel vectors of those layers (bias): tile[h * co + b] = v[b]; fold: grad[b] += sum_h t[h * co + b]
__global__ void vec_tile_kernel(const float* __restrict__ v, int co, int H, int cop, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cop; i += gridDim.x * blockDim.x)
        out[i] = i < H * co ? v[i % co] : 0.f;
}
__global__ void vec_fold_kernel(const float* __restrict__ t, int co, int H, float* __restrict__ grad) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < co; b += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += t[h * co + b];
        grad[b] += s;
    }
}

// x fp32 frames of H stacked lines ((B, T, H*L) batch-major or [T*B, ldx] time-major) -> out16 [T*B*S, Cp]:
// row (t*B+b)*S + p, channel h = (x[.., h*L + p] - mean[h*L + p]) * istd[h*L + p]; channels >= H and rows p >= L zero.
__global__ void conv_stage_lines_kernel(const float* __restrict__ x, int ldx, int time_major_in, int B, int T, int H,
                                        int L, int S, int Cp, const float* __restrict__ mean,
                                        const float* __restrict__ istd, uint16_t* __restrict__ out, int bf) {
    const long long total = (long long)B * T * S * Cp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        const long long row = i / Cp;
        const int p = (int)(row % S);
        const long long r = row / S;
        float v = 0.f;
        if (c < H && p < L) {
            const int b = (int)(r % B), t = (int)(r / B);
            const int f = c * L + p;
            v = time_major_in ? x[r * ldx + f] : x[((long long)b * T + t) * ldx + f];
            if (mean) v = (v - mean[f]) * istd[f];
        }
        out[i] = f2h(v, bf);
    }
}

// Taps of one PHASE of a stride-`step` transposed convolution (data gradient of utils/ops.py `downconv`, forward of
// `deconv`): out[q][b][a] = w[step * (nj - 1 - q) + c][a][b]   (w: [W, Ap, Bp], out: [nj, Bp, Ap]; taps with
// index = c mod step, in reverse order, channel axes transposed).  step = 1, c = 0, nj = W is conv_w_flip.
__global__ void conv_w_phase_kernel(const uint16_t* __restrict__ w, int Ap, int Bp, int step, int c, int nj,
                                    uint16_t* __restrict__ out) {
    const int total = nj * Ap * Bp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int a = i % Ap;
        const int b = (i / Ap) % Bp;
        const int q = i / (Ap * Bp);
        out[i] = w[((step * (nj - 1 - q) + c) * Ap + a) * Bp + b];
    }
}

// ---------------------------------------------------------------------------------------
// fully_connected with ONE output unit (the discriminator heads, models/discriminator_dnn.py:90-92,
// models/discriminator_lstm.py:100-104): a 128 x 16 tensor-core tile is almost all padding there, and both
// directions are pure HBM streams -- forward = one dot product per row, data gradient = an outer product masked by
// the producer's relu'.  One warp per row, 16-byte chunks per lane, the weight column staged in shared memory.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fc1_fwd_kernel(const uint16_t* __restrict__ x, int ldx, long long rows, int K,
                                                      const uint16_t* __restrict__ w, int ldw, const float* __restrict__ bias,
                                                      float* __restrict__ out, int ldo, int bf) {
    extern __shared__ float fc1_w[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) fc1_w[k] = h2f(w[(size_t)k * ldw], bf);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float b0 = bias ? bias[0] : 0.f;
    // two rows per warp and iteration: twice the loads in flight (a 1024-wide row is only four 16-byte loads per lane)
    for (long long r = warp0; r < rows; r += 2 * nwarps) {
        const long long r2 = r + nwarps;
        const bool two = r2 < rows;
        const uint16_t* xa = x + r * ldx;
        const uint16_t* xb = x + (two ? r2 : r) * ldx;
        float acc = 0.f, acc2 = 0.f;
        for (int k0 = lane * 8; k0 < K; k0 += 256) {       // K is a multiple of 8 (zero-padded activations)
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(xa + k0));
            const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(xb + k0));
            const uint32_t u[4] = {q.x, q.y, q.z, q.w}, u2[4] = {q2.x, q2.y, q2.z, q2.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w0 = fc1_w[k0 + 2 * i], w1 = fc1_w[k0 + 2 * i + 1];
                acc = fmaf(h2f((uint16_t)(u[i] & 0xFFFFu), bf), w0, acc);
                acc = fmaf(h2f((uint16_t)(u[i] >> 16), bf), w1, acc);
                acc2 = fmaf(h2f((uint16_t)(u2[i] & 0xFFFFu), bf), w0, acc2);
                acc2 = fmaf(h2f((uint16_t)(u2[i] >> 16), bf), w1, acc2);
            }
        }
        acc = warp_sum(acc);
        acc2 = warp_sum(acc2);
        if (lane == 0) {
            out[r * ldo] = acc + b0;
            if (two) out[r2 * ldo] = acc2 + b0;
        }
    }
}

// dx[r, k] = dy[r] * w[k] * act'(y[r, k])   (y = the producer's activation OUTPUT; dact NONE: no mask)
__global__ void __launch_bounds__(256) fc1_bwd_dx_kernel(const uint16_t* __restrict__ dy, int ldy, long long rows, int K,
                                                         const uint16_t* __restrict__ w, int ldw,
                                                         const uint16_t* __restrict__ y, int ldd, int dact,
                                                         uint16_t* __restrict__ dx, int ldo, int bf) {
    extern __shared__ float fc1_w[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) fc1_w[k] = h2f(w[(size_t)k * ldw], bf);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float neg = dact == RSR_ACT_LRELU ? 0.3f : 0.0f;
    const bool mask = y != nullptr && dact != RSR_ACT_NONE;
    for (long long r = warp0; r < rows; r += nwarps) {
        const float g = h2f(dy[r * ldy], bf);
        for (int k0 = lane * 8; k0 < K; k0 += 256) {
            uint4 q = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);   // any positive pattern
            if (mask) q = __ldg(reinterpret_cast<const uint4*>(y + r * ldd + k0));
            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool pos_lo = !mask || (int16_t)(u[i] & 0xFFFFu) > 0, pos_hi = !mask || (int32_t)u[i] >= 0x10000;
                o[i] = pack2(g * fc1_w[k0 + 2 * i] * (pos_lo ? 1.0f : neg), g * fc1_w[k0 + 2 * i + 1] * (pos_hi ? 1.0f : neg), bf);
            }
            *reinterpret_cast<uint4*>(dx + r * ldo + k0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// The discriminator head in ONE pass over the last hidden activation: logit = x . w + b (models/discriminator_dnn.py:90-93),
// the LSGAN loss terms and their gradient at the logit (models/gan_rnn_placeholder.py:244-252, the same arithmetic as
// lsgan_mse_kernel, clip_by_value mask included), and the head's data gradient dx = dlogit w (.) act'(x) -- what
// fc1_fwd + lsgan_mse + fc1_bwd_dx do in three launches and two more passes over the activation.  One warp per row; the
// row is read twice back to back (second time from L1).  which = 0: D(labels) pass, target d_real; which = 1: D(G(x))
// pass, losses against d_fake and d_real, gradient against grad_target (d_fake in the D update, d_real in the G update).
struct HeadParams {
    const uint16_t* x; int ldx; long long rows; int K;
    const uint16_t* w; int ldw; const float* bias;
    int which, clip; float d_real, d_fake, grad_target, gscale;
    float* losses;
    float* logit; int ldl;
    uint16_t* dlogit; int ldg;
    int dact;                   // RSR_ACT_*: the activation that produced x (mask of the data gradient), NONE = no mask
    uint16_t* dx; int ldo;
    int bf;
};

__global__ void __launch_bounds__(256) fc1_head_kernel(const HeadParams p) {
    extern __shared__ float fc1_w[];
    __shared__ float sh[8];
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) fc1_w[k] = h2f(p.w[(size_t)k * p.ldw], p.bf);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float b0 = p.bias ? p.bias[0] : 0.f;
    const float inv_nl = 1.0f / (float)p.rows;
    const float neg = p.dact == RSR_ACT_LRELU ? 0.3f : 0.0f;
    const bool mask = p.dact != RSR_ACT_NONE;
    float s_a = 0.f, s_b = 0.f;
    for (long long r = warp0; r < p.rows; r += nwarps) {
        const uint16_t* xr = p.x + r * p.ldx;
        float acc = 0.f;
        for (int k0 = lane * 8; k0 < p.K; k0 += 256) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(xr + k0));
            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc = fmaf(h2f((uint16_t)(u[i] & 0xFFFFu), p.bf), fc1_w[k0 + 2 * i], acc);
                acc = fmaf(h2f((uint16_t)(u[i] >> 16), p.bf), fc1_w[k0 + 2 * i + 1], acc);
            }
        }
        const float uu = warp_sum(acc) + b0;
        const float l = p.clip ? fminf(fmaxf(uu, -0.5f), 1.5f) : uu;
        const float in = (!p.clip || (uu >= -0.5f && uu <= 1.5f)) ? 1.f : 0.f;
        const float e_real = l - p.d_real, e_fake = l - p.d_fake;
        float e_grad;
        if (p.which == 0) { e_grad = e_real; if (lane == 0) s_a += e_real * e_real; }
        else { e_grad = l - p.grad_target; if (lane == 0) { s_a += e_fake * e_fake; s_b += e_real * e_real; } }
        const uint16_t g16 = f2h(p.gscale * 2.f * e_grad * inv_nl * in, p.bf);
        if (lane == 0) {
            if (p.logit) p.logit[r * p.ldl] = uu;
            if (p.dlogit) p.dlogit[r * p.ldg] = g16;
        }
        if (p.dx) {
            const float g = h2f(g16, p.bf);
            for (int k0 = lane * 8; k0 < p.K; k0 += 256) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(xr + k0));
                const uint32_t u[4] = {q.x, q.y, q.z, q.w};
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool pos_lo = !mask || (int16_t)(u[i] & 0xFFFFu) > 0, pos_hi = !mask || (int32_t)u[i] >= 0x10000;
                    o[i] = pack2(g * fc1_w[k0 + 2 * i] * (pos_lo ? 1.0f : neg), g * fc1_w[k0 + 2 * i + 1] * (pos_hi ? 1.0f : neg), p.bf);
                }
                *reinterpret_cast<uint4*>(p.dx + r * p.ldo + k0) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    if (p.losses) {
        block_atomic_add(s_a * inv_nl, p.losses + (p.which == 0 ? 0 : 1), sh);
        if (p.which != 0) block_atomic_add(s_b * inv_nl, p.losses + 2, sh);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" int rsr_stage_input(rsr_handle* h, void* stream, const float* x, int ldx, int time_major_in,
                               int B, int T, int D, const float* mean, const float* istd, const float* noise,
                               void* out16, int ld16, float* out32, int ld32) {
    if (!h || !x || B <= 0 || T <= 0 || D <= 0 || (!out16 && !out32)) return RSR_E_ARG;
    if ((mean == nullptr) != (istd == nullptr)) return RSR_E_ARG;
    if ((out16 && ld16 < D) || (out32 && ld32 < D) || ldx < D) return RSR_E_SHAPE;
    const long long total = (long long)B * T * 32;              // one warp per row
    stage_input_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, time_major_in, B, T, D, mean, istd, noise, (uint16_t*)out16, ld16, out32, ld32,
        h->dtype == RSR_DTYPE_BF16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_unstage_output(rsr_handle* h, void* stream, const float* y_tm, int ld, int B, int T, int D,
                                  const float* mean, const float* std, float* out_bm) {
    if (!h || !y_tm || !out_bm || B <= 0 || T <= 0 || D <= 0 || ld < D) return RSR_E_ARG;
    if ((mean == nullptr) != (std == nullptr)) return RSR_E_ARG;
    const long long total = (long long)B * T * D;
    unstage_output_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        y_tm, ld, B, T, D, mean, std, out_bm);
    RSR_LAUNCH_CHECK();
    return 0;
}

static int cmvn_launch(rsr_handle* h, void* stream, const float* x, const float* mean, const float* std,
                       long long N, int D, float* out, int mode) {
    if (!h || !x || !mean || !std || !out || N < 0 || D <= 0) return RSR_E_ARG;
    if (N == 0) return 0;
    const long long total = N * D;
    if (D > 4096 || (((uintptr_t)x | (uintptr_t)out) & 15)) return RSR_E_ARG;
    cmvn_kernel<<<grid_for((total + 3) / 4, 256, h->num_sms), 256, 2 * D * sizeof(float), (cudaStream_t)stream>>>(
        x, mean, std, total, D, mode, out);
    RSR_LAUNCH_CHECK();
    return 0;
}
extern "C" int rsr_cmvn_apply_padded(rsr_handle* h, void* stream, const float* x, const int* lengths, const double* mean,
                                     const double* std, int B, int T, int D, float* out) {
    if (!h || !x || !lengths || !mean || !std || !out || B < 0 || T <= 0 || D <= 0 || D > 2048) return RSR_E_ARG;
    if (((uintptr_t)x | (uintptr_t)out) & 15) return RSR_E_ARG;
    if (B == 0) return 0;
    const long long total = (long long)B * T * D;
    cmvn_padded_kernel<<<grid_for((total + 3) / 4, 256, h->num_sms), 256, 2 * D * sizeof(double), (cudaStream_t)stream>>>(
        x, lengths, mean, std, total, T, D, out);
    RSR_LAUNCH_CHECK();
    return 0;
}
extern "C" int rsr_cmvn_apply(rsr_handle* h, void* stream, const float* x, const float* mean, const float* std,
                              long long N, int D, float* out) { return cmvn_launch(h, stream, x, mean, std, N, D, out, 0); }
extern "C" int rsr_cmvn_invert(rsr_handle* h, void* stream, const float* y, const float* mean, const float* std,
                               long long N, int D, float* out) { return cmvn_launch(h, stream, y, mean, std, N, D, out, 1); }

extern "C" int rsr_lsgan_mse_losses(rsr_handle* h, void* stream, const float* d_rl_logit, const float* d_fk_logit,
                                    int ld_logit, long long n_logit, int clip, const float* g, int ldg,
                                    const float* y, int ldy, long long n_frames, int D_out, float d_real,
                                    float d_fake, float lambda, float gscale, float* losses, void* d_rl_grad,
                                    void* d_fk_grad, void* g_adv_grad, int ld_grad, float* dg_mse, int lddg) {
    if (!h || !losses) return RSR_E_ARG;
    if ((d_rl_logit || d_fk_logit) && (n_logit <= 0 || ld_logit <= 0)) return RSR_E_ARG;
    if ((g != nullptr) != (y != nullptr)) return RSR_E_ARG;
    if (g && (n_frames <= 0 || D_out <= 0 || ldg < D_out || ldy < D_out)) return RSR_E_ARG;
    if ((d_rl_grad || d_fk_grad || g_adv_grad) && ld_grad <= 0) return RSR_E_ARG;
    LossParams p;
    p.rl = d_rl_logit; p.fk = d_fk_logit; p.ldl = ld_logit; p.n_logit = (d_rl_logit || d_fk_logit) ? n_logit : 0; p.clip = clip;
    p.g = g; p.ldg = ldg; p.y = y; p.ldy = ldy; p.n_frames = n_frames; p.D = D_out;
    p.d_real = d_real; p.d_fake = d_fake; p.lambda = lambda; p.gscale = gscale; p.losses = losses;
    p.d_rl_grad = (uint16_t*)d_rl_grad; p.d_fk_grad = (uint16_t*)d_fk_grad; p.g_adv_grad = (uint16_t*)g_adv_grad;
    p.ldgrad = ld_grad; p.dg_mse = dg_mse; p.lddg = lddg; p.bf = h->dtype == RSR_DTYPE_BF16;
    long long work = p.n_logit;
    if (g && n_frames * D_out > work) work = n_frames * D_out;
    lsgan_mse_kernel<<<grid_for(work, 256, h->num_sms, 4), 256, 0, (cudaStream_t)stream>>>(p);
    RSR_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int colsum_launch(rsr_handle* h, void* stream, const T* x, int ld, long long M, int N, float* out, int accumulate) {
    if (!h || !x || !out || M <= 0 || N <= 0 || ld < N) return RSR_E_ARG;
    if (!accumulate) RSR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, (cudaStream_t)stream));
    if constexpr (sizeof(T) == 2) {
        if ((N & 7) == 0 && (ld & 7) == 0 && ((uintptr_t)x & 15) == 0) {
            const int gxv = (N + 255) / 256;
            long long gyv = (4LL * h->num_sms + gxv - 1) / gxv;
            if (gyv > (M + 63) / 64) gyv = (M + 63) / 64;
            if (gyv < 1) gyv = 1;
            const long long rpbv = (M + gyv - 1) / gyv;
            colsum16_vec_kernel<<<dim3(gxv, (unsigned)gyv), 256, 0, (cudaStream_t)stream>>>(
                (const uint16_t*)x, ld, M, N, out, h->dtype == RSR_DTYPE_BF16, rpbv);
            RSR_LAUNCH_CHECK();
            return 0;
        }
    }
    const int gx = (N + 31) / 32;
    long long gy = (2LL * h->num_sms + gx - 1) / gx;
    if (gy > (M + 63) / 64) gy = (M + 63) / 64;
    if (gy < 1) gy = 1;
    const long long rpb = (M + gy - 1) / gy;
    colsum_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(x, ld, M, N, out, h->dtype == RSR_DTYPE_BF16, rpb);
    RSR_LAUNCH_CHECK();
    return 0;
}
extern "C" int rsr_colsum16(rsr_handle* h, void* stream, const void* x16, int ld, long long M, int N, float* out, int accumulate) {
    return colsum_launch<uint16_t>(h, stream, (const uint16_t*)x16, ld, M, N, out, accumulate);
}
extern "C" int rsr_colsum32(rsr_handle* h, void* stream, const float* x32, int ld, long long M, int N, float* out, int accumulate) {
    return colsum_launch<float>(h, stream, x32, ld, M, N, out, accumulate);
}

extern "C" int rsr_seg_sumsq(rsr_handle* h, void* stream, const float* grad, float gmul, const int* seg_id,
                             long long n_elems, int n_seg, float* sumsq) {
    if (!h || !grad || !seg_id || !sumsq || n_elems <= 0 || (n_elems & 1023) || n_seg <= 0) return RSR_E_ARG;
    const long long n_blocks = n_elems / 1024;
    if (n_blocks > RSR_PARTIAL_WORDS) return RSR_E_SHAPE;
    // (calls that share a handle must be stream-ordered: the per-block partials live in the handle's workspace)
    seg_sumsq_kernel<<<(unsigned)n_blocks, 256, 0, (cudaStream_t)stream>>>(grad, gmul, h->partials);
    seg_sumsq_finish_kernel<<<n_seg, 256, 0, (cudaStream_t)stream>>>(h->partials, seg_id, (int)n_blocks, sumsq);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_clip_sgd_ema(rsr_handle* h, void* stream, const float* grad, float gmul, const int* seg_id,
                                const float* sumsq, int n_seg, float max_norm, float* hyper, float ema_decay,
                                long long n_elems, float* theta, float* ema, void* theta16) {
    if (!h || !grad || !seg_id || !sumsq || !hyper || !theta || n_elems <= 0 || (n_elems & 1023) || n_seg < 0) return RSR_E_ARG;
    clip_update_kernel<0><<<(unsigned)(n_elems / 1024), 256, 0, (cudaStream_t)stream>>>(
        grad, gmul, seg_id, sumsq, max_norm, hyper, ema_decay, theta, nullptr, nullptr, ema, (uint16_t*)theta16,
        h->dtype == RSR_DTYPE_BF16, n_seg);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_clip_adam_ema(rsr_handle* h, void* stream, const float* grad, float gmul, const int* seg_id,
                                 const float* sumsq, int n_seg, float max_norm, float* hyper, float ema_decay,
                                 long long n_elems, float* theta, float* m, float* v, float* ema, void* theta16) {
    if (!h || !grad || !seg_id || !sumsq || !hyper || !theta || !m || !v || n_elems <= 0 || (n_elems & 1023) || n_seg < 0)
        return RSR_E_ARG;
    clip_update_kernel<1><<<(unsigned)(n_elems / 1024), 256, 0, (cudaStream_t)stream>>>(
        grad, gmul, seg_id, sumsq, max_norm, hyper, ema_decay, theta, m, v, ema, (uint16_t*)theta16,
        h->dtype == RSR_DTYPE_BF16, n_seg);
    adam_tock_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hyper, sumsq, n_seg);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_add_cast(rsr_handle* h, void* stream, const float* a, const float* b, long long n,
                            float* out32, void* out16) {
    if (!h || !a || !b || (!out32 && !out16) || n < 0 || (n & 3)) return RSR_E_ARG;
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out32) & 15 || ((uintptr_t)out16 & 7)) return RSR_E_ARG;
    if (n == 0) return 0;
    add_cast_kernel<<<grid_for(n / 4, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        a, b, n / 4, out32, (uint16_t*)out16, h->dtype == RSR_DTYPE_BF16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_l2_grad(rsr_handle* h, void* stream, float* grad, const float* theta, const int* seg_id,
                           const int* seg_flag, float scale, long long n_elems) {
    if (!h || !grad || !theta || !seg_id || !seg_flag || n_elems <= 0 || (n_elems & 1023)) return RSR_E_ARG;
    l2_grad_kernel<<<(unsigned)(n_elems / 1024), 256, 0, (cudaStream_t)stream>>>(grad, theta, seg_id, seg_flag, scale);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_transpose16(rsr_handle* h, void* stream, const void* src, int ld_src, int rows, int cols, void* dst,
                               int ld_dst) {
    if (!h || !src || !dst || rows <= 0 || cols <= 0 || ld_src < cols || ld_dst < rows) return RSR_E_ARG;
    transpose16_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)src, ld_src, rows, cols, (uint16_t*)dst, ld_dst);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_cast16(rsr_handle* h, void* stream, const float* x, long long n, void* out16) {
    if (!h || !x || !out16 || n < 0) return RSR_E_ARG;
    if (n == 0) return 0;
    cast16_kernel<<<grid_for(n, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(x, n, (uint16_t*)out16, h->dtype == RSR_DTYPE_BF16);
    RSR_LAUNCH_CHECK();
    return 0;
}
extern "C" int rsr_fill32(rsr_handle* h, void* stream, float* x, long long n, float v) {
    if (!h || !x || n < 0) return RSR_E_ARG;
    if (n == 0) return 0;
    fill32_kernel<<<grid_for(n, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(x, n, v);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_conv_stage_frames(rsr_handle* h, void* stream, const float* x, int ldx, int time_major_in,
                                     int B, int T, int L, int S, int Cp, const float* mean, const float* istd,
                                     void* out16) {
    if (!h || !x || !out16 || B <= 0 || T <= 0 || L <= 0) return RSR_E_ARG;
    if ((mean == nullptr) != (istd == nullptr)) return RSR_E_ARG;
    if (S < L || Cp < 8 || (Cp & 7) || ldx < L || ((uintptr_t)out16 & 15)) return RSR_E_SHAPE;
    const long long total = (long long)B * T * S * (Cp / 8);
    conv_stage_frames_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, time_major_in, B, T, L, S, Cp, mean, istd, (uint16_t*)out16, h->dtype == RSR_DTYPE_BF16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_conv_mask_rows(rsr_handle* h, void* stream, void* buf16, long long frames, int S, int L, int Cp) {
    if (!h || !buf16 || frames <= 0) return RSR_E_ARG;
    if (S < L || L <= 0 || Cp < 8 || (Cp & 7) || ((uintptr_t)buf16 & 15)) return RSR_E_SHAPE;
    if (S == L) return 0;
    const long long total = frames * (S - L) * (Cp / 8);
    conv_mask_rows_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        (uint16_t*)buf16, frames, S, L, Cp);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_conv_w_flip(rsr_handle* h, void* stream, const void* w16, int W, int Cin_p, int Cout_p, void* out16) {
    if (!h || !w16 || !out16 || W <= 0 || Cin_p <= 0 || Cout_p <= 0) return RSR_E_ARG;
    conv_w_flip_kernel<<<grid_for((long long)W * Cin_p * Cout_p, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)w16, W, Cin_p, Cout_p, (uint16_t*)out16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_conv_toeplitz_expand(rsr_handle* h, void* stream, const void* w16, int kh, int W, int ci, int co, int H,
                                        int cip, int cop, void* out16) {
    if (!h || !w16 || !out16 || kh <= 0 || !(kh & 1) || W <= 0 || ci <= 0 || co <= 0 || H <= 0) return RSR_E_ARG;
    if (cip < H * ci || cop < H * co || (cip & 7) || (cop & 7)) return RSR_E_SHAPE;
    conv_toeplitz_expand_kernel<<<grid_for((long long)W * cip * cop, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)w16, kh, W, ci, co, H, cip, cop, (uint16_t*)out16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_conv_toeplitz_fold(rsr_handle* h, void* stream, const float* dw2, int kh, int W, int ci, int co, int H,
                                      int cip, int cop, float* grad) {
    if (!h || !dw2 || !grad || kh <= 0 || !(kh & 1) || W <= 0 || ci <= 0 || co <= 0 || H <= 0) return RSR_E_ARG;
    if (cip < H * ci || cop < H * co) return RSR_E_SHAPE;
    conv_toeplitz_fold_kernel<<<grid_for((long long)kh * W * ci * co, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        dw2, kh, W, ci, co, H, cip, cop, grad);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_vec_tile(rsr_handle* h, void* stream, const float* v, int co, int H, int cop, float* out) {
    if (!h || !v || !out || co <= 0 || H <= 0 || cop < H * co) return RSR_E_ARG;
    vec_tile_kernel<<<(cop + 255) / 256, 256, 0, (cudaStream_t)stream>>>(v, co, H, cop, out);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_vec_fold(rsr_handle* h, void* stream, const float* t, int co, int H, float* grad) {
    if (!h || !t || !grad || co <= 0 || H <= 0) return RSR_E_ARG;
    vec_fold_kernel<<<(co + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, co, H, grad);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_conv_stage_lines(rsr_handle* h, void* stream, const float* x, int ldx, int time_major_in, int B, int T,
                                    int H, int L, int S, int Cp, const float* mean, const float* istd, void* out16) {
    if (!h || !x || !out16 || B <= 0 || T <= 0 || H <= 0 || L <= 0) return RSR_E_ARG;
    if ((mean == nullptr) != (istd == nullptr)) return RSR_E_ARG;
    if (S < L || Cp < H || (Cp & 7) || ldx < H * L || ((uintptr_t)out16 & 15)) return RSR_E_SHAPE;
    const long long total = (long long)B * T * S * Cp;
    conv_stage_lines_kernel<<<grid_for(total, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, time_major_in, B, T, H, L, S, Cp, mean, istd, (uint16_t*)out16, h->dtype == RSR_DTYPE_BF16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_conv_w_phase(rsr_handle* h, void* stream, const void* w16, int W, int Ap, int Bp, int step, int phase,
                                void* out16) {
    if (!h || !w16 || !out16 || W <= 0 || Ap <= 0 || Bp <= 0 || step <= 0 || phase < 0 || phase >= step) return RSR_E_ARG;
    const int nj = (W - phase + step - 1) / step;          // taps with index = phase mod step
    if (nj <= 0) return RSR_E_ARG;
    conv_w_phase_kernel<<<grid_for((long long)nj * Ap * Bp, 256, h->num_sms), 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)w16, Ap, Bp, step, phase, nj, (uint16_t*)out16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_fc1_fwd(rsr_handle* h, void* stream, const void* x16, int ldx, long long rows, int K,
                           const void* w16, int ldw, const float* bias, float* out32, int ldo) {
    if (!h || !x16 || !w16 || !out32 || rows <= 0 || K <= 0 || ldo <= 0 || ldw <= 0) return RSR_E_ARG;
    if ((K & 7) || (ldx & 7) || ldx < K || ((uintptr_t)x16 & 15) || K > 8192) return RSR_E_SHAPE;
    // few, long-lived blocks: every block stages the strided weight column once
    fc1_fwd_kernel<<<grid_for(rows * 16, 256, h->num_sms, 4), 256, (size_t)K * 4, (cudaStream_t)stream>>>(
        (const uint16_t*)x16, ldx, rows, K, (const uint16_t*)w16, ldw, bias, out32, ldo, h->dtype == RSR_DTYPE_BF16);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_fc1_head(rsr_handle* h, void* stream, const void* x16, int ldx, long long rows, int K,
                            const void* w16, int ldw, const float* bias, int which, int clip, float d_real, float d_fake,
                            float grad_target, float gscale, float* losses, float* logit32, int ldl, void* dlogit16, int ldg,
                            int dact, void* dx16, int ldo) {
    if (!h || !x16 || !w16 || rows <= 0 || K <= 0 || ldw <= 0 || (which != 0 && which != 1)) return RSR_E_ARG;
    if ((K & 7) || (ldx & 7) || ldx < K || ((uintptr_t)x16 & 15) || K > 8192) return RSR_E_SHAPE;
    if (dx16 && ((ldo & 7) || ldo < K || ((uintptr_t)dx16 & 15))) return RSR_E_SHAPE;
    if ((logit32 && ldl <= 0) || (dlogit16 && ldg <= 0)) return RSR_E_ARG;
    HeadParams p;
    p.x = (const uint16_t*)x16; p.ldx = ldx; p.rows = rows; p.K = K; p.w = (const uint16_t*)w16; p.ldw = ldw; p.bias = bias;
    p.which = which; p.clip = clip; p.d_real = d_real; p.d_fake = d_fake; p.grad_target = grad_target; p.gscale = gscale;
    p.losses = losses; p.logit = logit32; p.ldl = ldl; p.dlogit = (uint16_t*)dlogit16; p.ldg = ldg;
    p.dact = dact; p.dx = (uint16_t*)dx16; p.ldo = ldo; p.bf = h->dtype == RSR_DTYPE_BF16;
    fc1_head_kernel<<<grid_for(rows * 32, 256, h->num_sms, 8), 256, (size_t)K * 4, (cudaStream_t)stream>>>(p);
    RSR_LAUNCH_CHECK();
    return 0;
}

extern "C" int rsr_fc1_bwd_dx(rsr_handle* h, void* stream, const void* dy16, int ldy, long long rows, int K,
                              const void* w16, int ldw, const void* dact_src, int ldd, int dact, void* dx16, int ldo) {
    if (!h || !dy16 || !w16 || !dx16 || rows <= 0 || K <= 0 || ldy <= 0 || ldw <= 0) return RSR_E_ARG;
    if ((K & 7) || (ldo & 7) || ldo < K || ((uintptr_t)dx16 & 15) || K > 8192) return RSR_E_SHAPE;
    if (dact_src && ((ldd & 7) || ldd < K || ((uintptr_t)dact_src & 15))) return RSR_E_SHAPE;
    fc1_bwd_dx_kernel<<<grid_for(rows * 32, 256, h->num_sms, 2), 256, (size_t)K * 4, (cudaStream_t)stream>>>(
        (const uint16_t*)dy16, ldy, rows, K, (const uint16_t*)w16, ldw, (const uint16_t*)dact_src, ldd, dact,
        (uint16_t*)dx16, ldo, h->dtype == RSR_DTYPE_BF16);
    RSR_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------
// Kaldi compressed matrix ("CM") decode, io_funcs/kaldi_io.py:121-161, fused with the CMVN of
// io_funcs/make_tfrecords.py:84-87.  Bytes are column-major on disk, the result is row-major: a
// 32-column x 128-row byte tile goes through shared memory (byte reads coalesced along rows, fp32 / fp64 writes
// coalesced along columns).  All arithmetic is float64 with the reference's operation order and explicit
// round-to-nearest intrinsics (no FMA contraction), so the result is bit-identical to the Python reader.
// ---------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ double u16_to_double(double minv, double range, unsigned v) {
    // min_value + range * 1.52590218966964e-05 * value  (left to right)
    return __dadd_rn(minv, __dmul_rn(__dmul_rn(range, 1.52590218966964e-05), (double)v));
}

__global__ void __launch_bounds__(256) ark_decompress_kernel(const uint16_t* __restrict__ hdr,
                                                             const uint8_t* __restrict__ data, double minv, double range,
                                                             int rows, int cols, double* __restrict__ out64, int ld64,
                                                             const double* __restrict__ mean,
                                                             const double* __restrict__ stdv, float* __restrict__ out32,
                                                             int ld32) {
    __shared__ uint8_t tile[32][132];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 128;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int ci = ty; ci < 32; ci += 8) {
        const int c = c0 + ci;
        if (c < cols) {
            const uint8_t* src = data + (long long)c * rows;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = r0 + k * 32 + tx;
                if (r < rows) tile[ci][k * 32 + tx] = src[r];
            }
        }
    }
    __syncthreads();
    const int c = c0 + tx;
    if (c >= cols) return;
    const double p0 = u16_to_double(minv, range, hdr[c * 4 + 0]), p25 = u16_to_double(minv, range, hdr[c * 4 + 1]);
    const double p75 = u16_to_double(minv, range, hdr[c * 4 + 2]), p100 = u16_to_double(minv, range, hdr[c * 4 + 3]);
    const double d_lo = __dsub_rn(p25, p0), d_mid = __dsub_rn(p75, p25), d_hi = __dsub_rn(p100, p75);
    const double mu = mean ? mean[c] : 0.0, sd = stdv ? stdv[c] : 1.0;
    for (int rr = ty; rr < 128; rr += 8) {
        const int r = r0 + rr;
        if (r >= rows) break;
        const int v = tile[tx][rr];
        double x;
        if (v < 64) x = __dadd_rn(p0, __dmul_rn(__dmul_rn(d_lo, (double)v), 1.0 / 64.0));
        else if (v <= 192) x = __dadd_rn(p25, __dmul_rn(__dmul_rn(d_mid, (double)(v - 64)), 1.0 / 128.0));
        else x = __dadd_rn(p75, __dmul_rn(__dmul_rn(d_hi, (double)(v - 192)), 1.0 / 63.0));
        if (out64) out64[(long long)r * ld64 + c] = x;
        if (out32) out32[(long long)r * ld32 + c] = (float)(mean ? __ddiv_rn(__dsub_rn(x, mu), sd) : x);
    }
}
}  // namespace

extern "C" int rsr_ark_decompress(rsr_handle* h, void* stream, const void* col_hdr, const void* data, float min_value,
                                  float range, int rows, int cols, double* out64, int ld64, const double* mean,
                                  const double* std, float* out32, int ld32) {
    if (!h || !col_hdr || !data || (!out64 && !out32) || rows <= 0 || cols <= 0) return RSR_E_ARG;
    if ((mean == nullptr) != (std == nullptr)) return RSR_E_ARG;
    if ((out64 && ld64 < cols) || (out32 && ld32 < cols)) return RSR_E_SHAPE;
    dim3 grid((cols + 31) / 32, (rows + 127) / 128), block(32, 8);
    ark_decompress_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint16_t*)col_hdr, (const uint8_t*)data,
                                                                   (double)min_value, (double)range, rows, cols, out64,
                                                                   ld64, mean, std, out32, ld32);
    RSR_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------
// CRC-32C (Castagnoli), host side: the checksum of TensorFlow's checkpoint-V2 "tensor bundle" files
// (block trailers of the .index table and the per-tensor crc32c of BundleEntryProto), used by
// rsrgan_b200/tf_checkpoint.py to read / write the reference's `GAN_RNN-<step>` checkpoints
// (models/gan_rnn_placeholder.py:26-60).  Slicing-by-8 table lookup, ~1 GB/s: a 30 MB checkpoint in 30 ms.
// ---------------------------------------------------------------------------------------
namespace {
struct Crc32cTables {
    uint32_t t[8][256];
    Crc32cTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
    }
};
}  // namespace

extern "C" unsigned int rsr_crc32c_host(const void* data_host, unsigned long long n, unsigned int crc) {
    static const Crc32cTables T;
    const unsigned char* p = (const unsigned char*)data_host;
    uint32_t c = ~crc;
    while (n >= 8) {
        uint32_t lo, hi;
        memcpy(&lo, p, 4);
        memcpy(&hi, p + 4, 4);
        lo ^= c;
        c = T.t[7][lo & 0xff] ^ T.t[6][(lo >> 8) & 0xff] ^ T.t[5][(lo >> 16) & 0xff] ^ T.t[4][lo >> 24] ^
            T.t[3][hi & 0xff] ^ T.t[2][(hi >> 8) & 0xff] ^ T.t[1][(hi >> 16) & 0xff] ^ T.t[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xff];
    return ~c;
}
