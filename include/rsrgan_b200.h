/* rsrgan_b200 -- C ABI of the B200-native (sm_100a) RSRGAN GAN-training hot path.
 *
 * The reference (wangkenpu/rsrgan) has no FFI: its hot path is inline TensorFlow-1 ops.
 * Each entry point below replaces one group of those TF call sites (cited as
 * reference file:line, paths relative to the reference root) and is what a ctypes / cffi /
 * pybind stub on the reference side would bind.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     immediately; nothing synchronises, allocates persistent memory or throws;
 *   - return value: 0 = ok, > 0 = cudaError_t, < 0 = argument / shape error (RSR_E_*);
 *   - 16-bit tensors ("h16") hold IEEE fp16 or bf16 according to the handle's dtype;
 *   - internal sequence tensors are TIME-major: row index = t * B + b;
 *   - there is no CPU fallback: without an sm_100 device every compute call fails.
 */
#ifndef RSRGAN_B200_H_
#define RSRGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSR_E_ARG (-1)      /* null / misaligned pointer, non-positive size           */
#define RSR_E_SHAPE (-2)    /* shape the kernels do not support (see each function)  */
#define RSR_E_NODEV (-3)    /* no sm_100 device / driver entry point missing         */
#define RSR_E_RESIDENT (-4) /* persistent kernel would not be co-resident            */

#define RSR_DTYPE_F16 0
#define RSR_DTYPE_BF16 1

/* activations of tf.contrib.layers.fully_connected call sites */
#define RSR_ACT_NONE 0
#define RSR_ACT_RELU 1  /* models/discriminator_dnn.py:61-90, models/dnn.py:79-110 */
#define RSR_ACT_LRELU 2 /* utils/ops.py:120-121 leakyrelu(alpha=0.3), models/lstm.py:82-87 */
#define RSR_ACT_CLIP 3  /* models/discriminator_dnn.py:93 clip_by_value(-0.5, 1.5) */

typedef struct rsr_handle rsr_handle;

/* library / handle ------------------------------------------------------------------ */
int rsr_version(void);
/* Creates the per-rank handle (tensor-map cache, barrier workspace). dtype = RSR_DTYPE_*. */
int rsr_create(rsr_handle** out, int device, int dtype);
int rsr_destroy(rsr_handle* h);
int rsr_num_sms(rsr_handle* h);

/* GEMM with fused epilogue on the tcgen05 tensor cores ------------------------------- */
/* D[M,N] = epi(alpha * A[M,K] * B[K,N]) ; A and B are h16 operands fed by TMA.
 *   a_mn = 0: A stored row-major [M, lda] (K contiguous);  a_mn = 1: A stored [K, lda] (M contiguous)
 *   b_mn = 0: B stored [N, ldb] (K contiguous);             b_mn = 1: B stored row-major [K, ldb] (N contiguous)
 * epi(v): v += bias[n]; v += resid[m,n]; v = act(v); v *= act'(dact_src[m,n]) ; out16 = h16(v) ; out32 = v + beta*out32_old
 * Replaces tf.contrib.layers.fully_connected (models/lstm.py:82-87,121-124,
 * models/discriminator_dnn.py:61-93, models/discriminator_lstm.py:100-104), the input half of
 * LSTMCell's _Linear (models/lstm.py:90-96) hoisted over all frames, and their gradients. */
typedef struct rsr_gemm_args {
    int M, N, K;
    const void* A; int lda; int a_mn;
    const void* B; int ldb; int b_mn;
    float alpha, beta;
    const float* bias;            /* [N] or NULL */
    const float* resid; int ldr;  /* fp32 [M, ldr] or NULL */
    int act;                      /* RSR_ACT_* applied to the value */
    const void* dact_src; int ldd; int dact; /* h16 [M, ldd]: multiply by derivative of RSR_ACT_{RELU,LRELU} at that OUTPUT value */
    float* out32; int ldc32;      /* fp32 output or NULL */
    void* out16; int ldc16;       /* h16 output or NULL  */
    int tile_n;                   /* 0 = auto */
    int split_k;                  /* 0 = auto; > 1 only for "out32 += alpha A B" (beta = 1, no other epilogue term):
                                     partial sums are added into out32 with fp32 atomics */
    float* stats;                 /* NULL, or batch_norm statistics out of the epilogue (models/dnn.py:56-62,
                                     models/discriminator_dnn.py:36-46): for every block of 128 output rows and every
                                     column the (count, mean, M2) of out32, laid out [ceil(M / 128)][3][N] -- the partials
                                     rsr_bn_train_finish merges.  Only for a plain fp32 output (no bias / activation /
                                     16-bit copy, alpha = 1, beta = 0), N a multiple of 32, M <= 32768; else RSR_E_SHAPE */
} rsr_gemm_args;
int rsr_gemm(rsr_handle* h, void* stream, const rsr_gemm_args* a);

/* fully_connected with ONE output unit -- the discriminator heads (models/discriminator_dnn.py:90-92,
 * models/discriminator_lstm.py:100-104).  Same arithmetic as rsr_gemm (16-bit operands, fp32 accumulate) without
 * the padded tensor-core tile: both directions are HBM streams.
 *   forward        out32[r * ldo] = sum_k x16[r, k] * w16[k * ldw] + bias[0]          (K multiple of 8)
 *   data gradient  dx16[r, k] = dy16[r * ldy] * w16[k * ldw] * act'(dact_src[r, k])   (dact_src NULL: no mask;
 *                  act' from the producer's OUTPUT as in rsr_gemm)
 * The weight gradient stays rsr_gemm (x^T dy). */
int rsr_fc1_fwd(rsr_handle* h, void* stream, const void* x16, int ldx, long long rows, int K,
                const void* w16, int ldw, const float* bias, float* out32, int ldo);
int rsr_fc1_bwd_dx(rsr_handle* h, void* stream, const void* dy16, int ldy, long long rows, int K,
                   const void* w16, int ldw, const void* dact_src, int ldd, int dact, void* dx16, int ldo);
/* The whole discriminator head in one pass over the last hidden activation x16 (models/discriminator_dnn.py:90-93 +
 * the logit terms of models/gan_rnn_placeholder.py:244-252): rsr_fc1_fwd, the logit half of rsr_lsgan_mse_losses and
 * rsr_fc1_bwd_dx with identical arithmetic, three launches and two passes over the activation fewer.
 *   which = 0  D(labels):  losses[0] += mean((l - d_real)^2),                         gradient target d_real
 *   which = 1  D(G(x)):    losses[1] += mean((l - d_fake)^2), losses[2] += mean((l - d_real)^2), gradient target
 *              grad_target (d_fake in the discriminator update, d_real in the generator update)
 *   l = clip ? clip_by_value(logit, -0.5, 1.5) : logit (gradient zero outside the range);
 *   dlogit16[r * ldg] = h16(gscale * 2 (l - target) / rows);  dx16[r, k] = h16(dlogit16[r] * w16[k] * act'(x16[r, k])) with
 *   act = dact (the activation that produced x16; RSR_ACT_NONE: no mask).  logit32 / dlogit16 / dx16 / losses may be NULL. */
int rsr_fc1_head(rsr_handle* h, void* stream, const void* x16, int ldx, long long rows, int K,
                 const void* w16, int ldw, const float* bias, int which, int clip, float d_real, float d_fake,
                 float grad_target, float gscale, float* losses, float* logit32, int ldl, void* dlogit16, int ldg,
                 int dact, void* dx16, int ldo);

/* input staging ---------------------------------------------------------------------- */
/* x fp32 batch-major (B, T, D) -> h16 time-major [T*B, ld16] (and optional fp32 copy [T*B, ld32]):
 *   v = (x - mean[d]) * istd[d]   (mean/istd NULL = identity; io_funcs/make_tfrecords.py:84-87)
 *   v += noise[b, d]              (noise NULL = none; utils/ops.py:19-30 draws ONE (B,1,D) vector
 *                                  per utterance, broadcast over time; models/discriminator_lstm.py:60)
 * time_major_in != 0: x is already [T*B, ldx] fp32 time-major (used for G outputs fed to D). */
int rsr_stage_input(rsr_handle* h, void* stream, const float* x, int ldx, int time_major_in,
                    int B, int T, int D, const float* mean, const float* istd, const float* noise,
                    void* out16, int ld16, float* out32, int ld32);
/* fp32 time-major [T*B, ld] -> fp32 batch-major (B, T, D), optional decode CMVN y*std+mean
 * (scripts/train_gan_rnn_placeholder.py:286-287). */
int rsr_unstage_output(rsr_handle* h, void* stream, const float* y_tm, int ld, int B, int T, int D,
                       const float* mean, const float* std, float* out_bm);
/* CMVN on flat (N, D) matrices: out = (x - mean) / std, io_funcs/make_tfrecords.py:84-87;
 * inverse: out = y * std + mean, scripts/train_gan_rnn_placeholder.py:286-287.  x and out 16-byte aligned
 * (the kernel streams 16-byte vectors of the flat array), D <= 4096. */
int rsr_cmvn_apply(rsr_handle* h, void* stream, const float* x, const float* mean, const float* std,
                   long long N, int D, float* out);
int rsr_cmvn_invert(rsr_handle* h, void* stream, const float* y, const float* mean, const float* std,
                    long long N, int D, float* out);
/* The loader's CMVN on a zero-padded minibatch (B, T, D) fp32 with per-utterance lengths: frames t < lengths[b] become
 * float((double)x - mean[d]) / std[d]) -- float64 arithmetic on the float64 statistics of train_cmvn.npz, bit-identical
 * to io_funcs/make_tfrecords.py:84-87 -- and frames past the length are exact zeros, which is what padding AFTER the
 * normalisation gives (io_funcs/tfrecords_dataset.py:149-152).  out may alias x. */
int rsr_cmvn_apply_padded(rsr_handle* h, void* stream, const float* x, const int* lengths, const double* mean,
                          const double* std, int B, int T, int D, float* out);

/* LSTMP (tf.contrib.rnn.LSTMCell(use_peepholes, num_proj, forget_bias=1) under
 * tf.nn.dynamic_rnn(sequence_length)) ------------------------------------------------- */
/* Call sites: models/lstm.py:89-112, models/res_lstm_l.py:86-138, models/discriminator_lstm.py:70-91;
 * gate math: models/BNLSTMCell.py:176-213 (order i, j, f, o).
 *
 * The recurrence is run in the algebraically folded form
 *     z_t = Zx_t + mt_{t-1} * Wc ,  Wc = W_proj * K_h   (C x 4C),  out_t = mt_t * W_proj
 * so one persistent kernel does every time step with Wc resident on chip (TMEM for Cp <= 512, where the
 * CTAs of an utterance group form a thread-block cluster and exchange mt_t through distributed shared
 * memory; shared memory + L2 exchange for Cp > 512) and the projection becomes a batched GEMM outside the loop.  Columns of every 4C-wide tensor are in
 * PACKED gate order: col = (cell/32)*128 + gate*32 + cell%32, gate in (i, j, f, o); Cp = C padded
 * to a multiple of 256 with zero weights (padded cells stay exactly 0).
 *
 *   zx      fp32 [T*B, 4Cp]   x_t * K_x + bias, packed columns (from rsr_gemm)
 *   wcT     h16  [4Cp, Cp]    Wc transposed, packed rows (forward MMA A operand)
 *   w_i,w_f,w_o fp32 [Cp]     peepholes
 *   lengths int32 [B]         frames t >= lengths[b] are frozen: mt written as 0
 *   mt_seq  h16  [(T+1)*B, Cp] slot 0 must be zero on entry; slot t+1 receives mt_t
 *   save    fp32 [T*B, 5, Cp]  (i, f, o, tanh j, c_new) for the backward pass, or NULL
 */
int rsr_lstmp_rec_fwd(rsr_handle* h, void* stream, int B, int T, int Cp, const float* zx,
                      const void* wcT, const float* w_i, const float* w_f, const float* w_o,
                      float forget_bias, const int* lengths, void* mt_seq, float* save);
/* Fully fused LSTMP forward -- "the LSTM cell's gate GEMMs in one tcgen05 tile with the sigmoid/tanh/
 * elementwise epilogue in registers": z_t = x_t K_x + b + mt_{t-1} Wc computed entirely inside the persistent
 * cluster kernel (K_x^T and Wc^T slices resident in TMEM as the MMA A operand, x_t tiles prefetched by TMA),
 * no Zx round trip through HBM.  Same outputs as rsr_gemm(Zx) + rsr_lstmp_rec_fwd.
 *   x16     h16  [T*B, ldx]   layer input, time-major, zero padded to ldx (multiple of 8) columns
 *   kxT     h16  [4Cp, Ik]    K_x^T, packed gate rows, Ik = I rounded up to 16, zero padded
 *   bias    fp32 [4Cp]        packed
 * Returns RSR_E_RESIDENT (nothing launched) when the variant does not apply -- Cp > 512 or K_x^T does not fit
 * in TMEM beside Wc^T; the caller then falls back to rsr_gemm + rsr_lstmp_rec_fwd. */
int rsr_lstmp_fused_fwd(rsr_handle* h, void* stream, int B, int T, int I, int Cp, const void* x16, int ldx,
                        const void* kxT, const float* bias, const void* wcT, const float* w_i,
                        const float* w_f, const float* w_o, float forget_bias, const int* lengths,
                        void* mt_seq, float* save);
/* Layer wavefront: two stacked LSTMP layers of equal cell count (tf.contrib.rnn.MultiRNNCell under ONE dynamic_rnn,
 * models/lstm.py:89-112 -- layer 2 consumes layer 1's output of the same time step) in one launch.  Layer 1's clusters,
 * a projection stage (out1_t = mt1_t W_p1 per step, W_p1^T resident in TMEM) and layer 2's clusters run side by side,
 * layer 2 a few steps behind, handing rows over through global memory with one release / acquire counter per group of
 * utterances and step.  Same outputs as rsr_lstmp_fused_fwd(layer 1) + rsr_gemm(projection) + rsr_lstmp_fused_fwd
 * (layer 2): mt1, save1, out1 (16-bit rows [B, (T+1) B) of the layer-1 output sequence), mt2, save2.
 *   wpT1    h16 [Pp1, Cp]   W_p1^T (rows >= P1 zero); the other operands as for rsr_lstmp_fused_fwd, per layer;
 *                           layer 2's input width is P1, its K_x^T is [4Cp, round_up(P1, 16)].
 * Returns RSR_E_RESIDENT (nothing launched) when the shape does not apply: Cp > 512, the operands do not fit in TMEM,
 * or the 2 * groups + 1 clusters are not co-resident -- the caller then runs the layers one after the other. */
typedef struct rsr_wave_args {
    int B, T, Cp;
    int I1, P1;                   /* layer-1 input width and projection width (= layer-2 input width) */
    float forget_bias;
    const int* lengths;
    const void* x16; int ldx;     /* layer-1 input, time-major [T*B, ldx] */
    const void* kxT1; const float* bias1; const void* wcT1;
    const float* w_i1; const float* w_f1; const float* w_o1;
    void* mt1; float* save1;
    const void* wpT1;
    void* out1; int ldo1;         /* [(T+1)*B, ldo1]; rows [0, B) (initial state) are not written */
    const void* kxT2; const float* bias2; const void* wcT2;
    const float* w_i2; const float* w_f2; const float* w_o2;
    void* mt2; float* save2;
} rsr_wave_args;
int rsr_lstmp_wave_fwd(rsr_handle* h, void* stream, const rsr_wave_args* a);
/* Layer wavefront, backward (tf.gradients through the MultiRNNCell while-loop of models/lstm.py:89-112): the reversed
 * recurrences of layer 2 and layer 1 in one launch, layer 1 a few steps behind.  Between them a stage turns every dz2_t
 * into layer 1's incoming gradient  dmt1_t = (dz2_t K_x2^T) W_p1^T = dz2_t F^T  with  F = W_p1 K_x2  resident in TMEM.
 *   dmt2    fp32 [T*B, Cp]        dOut2 W_p2^T (from rsr_gemm), read only
 *   fT      h16  [Cp, 4Cp]        F: row = layer-1 cell, columns = layer-2 packed gate columns (refreshed with the weights)
 *   part    fp32 T*(B+48)*Cp      scratch for dmt1 (laid out per group of utterances): ALL ZEROS on entry, all zeros again
 *                                 on return (the stage accumulates its four K-slice partial products into it, layer 1
 *                                 resets every element as it reads it)
 * Outputs as two rsr_lstmp_rec_bwd calls: dz2, dz1 (h16 [T*B, 4Cp]) and the bias / peephole gradients (accumulated).
 * The gradient wrt layer 1's OUTPUT (dz2 K_x2^T), which the weight gradients of layer 1 need, stays an rsr_gemm.
 * Returns RSR_E_RESIDENT (nothing launched) when the shape does not apply, as rsr_lstmp_wave_fwd.
 * max_nbp = 32: a caller that hides weight-gradient GEMMs behind the per-layer launches (which leave 84 SMs free) says so
 * -- the 48-utterance launch (B = 128 at Cp = 512: all 7 placeable clusters) is 19 % faster than the two launches it
 * replaces but leaves 36 SMs, and the un-hidden GEMMs cost what it gains. */
typedef struct rsr_wave_bwd_args {
    int B, T, Cp;
    int max_nbp;                  /* 0: any; 32: decline rather than seat 48 utterances per cluster (see below) */
    const int* lengths;
    const float* dmt2; const void* wc2; const float* w_i2; const float* w_f2; const float* w_o2; const float* save2;
    void* dz2; float* dbias2; float* dw_i2; float* dw_f2; float* dw_o2;
    const void* fT; float* part;
    const void* wc1; const float* w_i1; const float* w_f1; const float* w_o1; const float* save1;
    void* dz1; float* dbias1; float* dw_i1; float* dw_f1; float* dw_o1;
} rsr_wave_bwd_args;
int rsr_lstmp_wave_bwd(rsr_handle* h, void* stream, const rsr_wave_bwd_args* a);
/*   dmt     fp32 [T*B, Cp]    dOut_t * W_proj^T (from rsr_gemm); read only on the cluster path (Cp <= 512), the
 *                              L2-exchange path (Cp > 512) adds the recurrent term dz_{t+1} * Wc^T in place
 *   wc      h16  [Cp, 4Cp]    Wc, packed columns (backward MMA A operand)
 *   dz16    h16  [T*B, 4Cp]   gate pre-activation gradients, packed columns (output)
 *   dbias   fp32 [4Cp] packed, dw_i, dw_f, dw_o fp32 [Cp]: ACCUMULATED into (caller zeroes)
 */
int rsr_lstmp_rec_bwd(rsr_handle* h, void* stream, int B, int T, int Cp, float* dmt,
                      const void* wc, const float* w_i, const float* w_f, const float* w_o,
                      const int* lengths, const float* save, void* dz16,
                      float* dbias, float* dw_i, float* dw_f, float* dw_o);

/* losses (models/gan_rnn_placeholder.py:244-260; same formulas models/gan.py:200-208) ---- */
/* All means run over EVERY element including padded frames (zero-padding of
 * io_funcs/tfrecords_dataset.py:149-152 is part of the mean).
 *   d_rl_logit / d_fk_logit : fp32, element i at [i * ld_logit] (D(labels), D(G(x))); either may be NULL
 *   clip != 0 : the values are PRE-clip outputs of discriminator_dnn; the kernel applies
 *               clip_by_value(-0.5, 1.5) (models/discriminator_dnn.py:93) and zeroes the gradient outside
 *   g, y      : fp32 [n_frames, ldg|ldy] generator output and labels (same row order); both or neither
 *   losses[0..3] += d_rl_loss, d_fk_loss, g_adv_loss, g_mse_loss  (caller zeroes; atomics)
 *       d_rl = mean((D(y)-d_real)^2)  d_fk = mean((D(g)-d_fake)^2)  g_adv = mean((D(g)-d_real)^2)
 *       g_mse = 0.5 * D_out * mean((g-y)^2)
 *   gradients (each may be NULL), h16 element i at [i * ld_grad], times the static loss scale gscale:
 *       d_rl_grad = 2 (D(y)-d_real)/n   d_fk_grad = 2 (D(g)-d_fake)/n   g_adv_grad = 2 (D(g)-d_real)/n
 *   dg_mse fp32 [n_frames, lddg] = gscale * lambda * (g - y) / n_frames     (d(lambda*g_mse)/dg) */
int rsr_lsgan_mse_losses(rsr_handle* h, void* stream, const float* d_rl_logit, const float* d_fk_logit,
                         int ld_logit, long long n_logit, int clip, const float* g, int ldg,
                         const float* y, int ldy, long long n_frames, int D_out, float d_real,
                         float d_fake, float lambda, float gscale, float* losses, void* d_rl_grad,
                         void* d_fk_grad, void* g_adv_grad, int ld_grad, float* dg_mse, int lddg);

/* column sums: out[n] (+)= sum_m X16[m, n]  (bias gradients) */
int rsr_colsum16(rsr_handle* h, void* stream, const void* x16, int ld, long long M, int N,
                 float* out, int accumulate);
int rsr_colsum32(rsr_handle* h, void* stream, const float* x32, int ld, long long M, int N,
                 float* out, int accumulate);

/* update (models/gan_rnn_placeholder.py:144-150,177-189; utils/ops.py:343-376) ---------- */
/* One flat fp32 parameter buffer per network; every tensor (TF variable) is a segment padded to a
 * multiple of 1024 elements, seg_id[k] = segment of the k-th 1024-element block (device int32).
 * Gradients arrive summed over ranks (ncclAllReduce replaces utils/ops.py:343-376
 * average_gradients) and multiplied by the static loss scale: gmul = 1 / (world_size * loss_scale).
 *   pass 1  rsr_seg_sumsq : sumsq[s] = sum over segment s of (gmul*g)^2        (deterministic: per-block partials, then
 *           a fixed-order sum per segment -- no atomics, so every data-parallel rank derives bit-identical clip scales;
 *           calls on one handle must be stream-ordered, the partials live in the handle's workspace)
 *   pass 2  ghat = gmul*g * max_norm / max(sqrt(sumsq[s]), max_norm)           (tf.clip_by_norm per tensor, :177-182)
 *           SGD : theta -= lr * ghat                                            (:144)
 *           Adam: m = b1 m + (1-b1) ghat ; v = b2 v + (1-b2) ghat^2 ;
 *                 theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)       (:147, TF-1 form)
 *           EMA : ema -= (1 - decay) * (ema - theta_new)                        (:149-150,185-189; ema NULL = skip)
 *           theta16 = h16(theta_new)      operand copy for the tensor cores (NULL = skip)
 * hyper (device fp32[8]): [0]=lr [1]=beta1 [2]=beta2 [3]=eps [4]=beta1_power [5]=beta2_power
 * [6]=lr_t scratch [7]=number of updates SKIPPED because a gradient norm was not finite (overflow
 * guard of the 16-bit backward pass: with n_seg > 0 both calls scan sumsq[0..n_seg) first and leave
 * theta / m / v / ema / theta16 / the beta powers untouched when any entry is inf or NaN; n_seg = 0
 * switches the guard off).  Learning rates live on the device so a captured CUDA graph stays valid when
 * the host decays them (scripts/train_gan_rnn_placeholder.py:525-533).  The Adam call advances
 * the beta powers exactly as TF's _finish does (initialise [4]=beta1, [5]=beta2). */
int rsr_seg_sumsq(rsr_handle* h, void* stream, const float* grad, float gmul, const int* seg_id,
                  long long n_elems, int n_seg, float* sumsq);
int rsr_clip_sgd_ema(rsr_handle* h, void* stream, const float* grad, float gmul, const int* seg_id,
                     const float* sumsq, int n_seg, float max_norm, float* hyper, float ema_decay,
                     long long n_elems, float* theta, float* ema, void* theta16);
int rsr_clip_adam_ema(rsr_handle* h, void* stream, const float* grad, float gmul, const int* seg_id,
                      const float* sumsq, int n_seg, float max_norm, float* hyper, float ema_decay,
                      long long n_elems, float* theta, float* m, float* v, float* ema, void* theta16);

/* L2 regulariser of G (models/gan_rnn_placeholder.py:253-258): grad += scale * theta on every
 * 1024-block whose segment has seg_flag != 0 (tensors whose name does not contain "bias"). */
int rsr_l2_grad(rsr_handle* h, void* stream, float* grad, const float* theta, const int* seg_id,
                const int* seg_flag, float scale, long long n_elems);

/* residual connection x_{l+1} = out_l + x_l (models/res_lstm_l.py:116,127,138,187):
 * out32 = a + b, out16 = h16(a + b); n multiple of 4; either output may be NULL. */
int rsr_add_cast(rsr_handle* h, void* stream, const float* a, const float* b, long long n,
                 float* out32, void* out16);

/* 1-D convolution family -------------------------------------------------------------- */
/* tf.contrib.layers.conv2d(inputs, C_out, [splice, w], SAME, relu) of models/rced.py:90-101 with
 * splice = 1 (a 1-D convolution over the 257 spectrum bins of one frame) is computed by rsr_gemm
 * itself: activations are channels-last 16-bit rows [frames*S, Cp] (frame r = rows r*S .. r*S+S-1,
 * positions 0..L-1 data, rows L..S-1 zero = the SAME padding shared with the next frame, S-L >= w/2,
 * plus >= w/2 zero guard rows before the first and after the last frame), and the GEMM A operand is the
 * OVERLAPPED view   A[m, k*Cp + c] = act[m - w/2 + k, c]   i.e. pointer act - (w/2)*Cp, lda = Cp,
 * K = w*Cp; B = taps [w*Cin_p, Cout_p].  No im2col matrix is materialised; TMA reads the window.
 *   forward        Y = relu(A(X) W + b)              then rsr_conv_mask_rows(Y)
 *   weight grad    dW += A(X)^T dY  (a_mn = 1)       bias grad = rsr_colsum16(dY)
 *   data grad      dX = A(dY) Wflip, times relu'(X)  Wflip from rsr_conv_w_flip
 * The three helpers below are the layout glue around those GEMMs. */
/* x fp32 frames ((B, T, L) batch-major, or [T*B, ldx] time-major) -> out16 [T*B*S, Cp]:
 * row (t*B+b)*S + p, channel 0 = (x[b,t,p] - mean[p]) * istd[p] (mean/istd NULL = identity), every other
 * element (channels 1.., rows p >= L) zero.  models/rced.py:46-57 (reshape to [batch, splice, input_dim, 1]). */
int rsr_conv_stage_frames(rsr_handle* h, void* stream, const float* x, int ldx, int time_major_in,
                          int B, int T, int L, int S, int Cp, const float* mean, const float* istd,
                          void* out16);
/* zeroes rows p in [L, S) of every frame (the GEMM computes garbage there). */
int rsr_conv_mask_rows(rsr_handle* h, void* stream, void* buf16, long long frames, int S, int L, int Cp);
/* out[k][co][ci] = w[W-1-k][ci][co]: taps of the transposed convolution (tf.gradients of conv2d wrt its input). */
int rsr_conv_w_flip(rsr_handle* h, void* stream, const void* w16, int W, int Cin_p, int Cout_p, void* out16);

/* [splice, w] convolutions (models/rced.py:94-101 with splice = left_context + 1 + right_context > 1, the shape
 * run_dnn.sh:129-140 trains: 40 bins x 11 lines): the H = splice lines of a frame are stored as CHANNELS of one position
 * (channel = line * C + c), so the 2-D SAME convolution IS the 1-D overlapped-view GEMM above with the block-Toeplitz
 * taps  W2[k][h_in*ci + a][h_out*co + b] = w[h_in - h_out + kh/2][k][a][b]  (zero outside the filter: the SAME padding
 * along the lines).  The parameter stays TensorFlow's compact filter [kh, W, ci, co]:
 *   rsr_conv_toeplitz_expand   compact 16-bit filter -> taps16 [W*cip, cop]  (after every update, like Wc)
 *   rsr_conv_toeplitz_fold     grad[kh, W, ci, co] += sum over the tied copies of dW2 [W*cip, cop] (fixed order)
 *   rsr_vec_tile / _fold       per-channel bias <-> its per-(line, channel) tiling and the gradient back
 *   rsr_conv_stage_lines       fp32 frames (.., H*L) -> channels-last 16-bit rows [frames*S, Cp], channel h = line h */
int rsr_conv_toeplitz_expand(rsr_handle* h, void* stream, const void* w16, int kh, int W, int ci, int co, int H,
                             int cip, int cop, void* out16);
int rsr_conv_toeplitz_fold(rsr_handle* h, void* stream, const float* dw2, int kh, int W, int ci, int co, int H,
                           int cip, int cop, float* grad);
int rsr_vec_tile(rsr_handle* h, void* stream, const float* v, int co, int H, int cop, float* out);
int rsr_vec_fold(rsr_handle* h, void* stream, const float* t, int co, int H, float* grad);
int rsr_conv_stage_lines(rsr_handle* h, void* stream, const float* x, int ldx, int time_major_in, int B, int T,
                         int H, int L, int S, int Cp, const float* mean, const float* istd, void* out16);

/* Strided members of the family (utils/ops.py:78-98 `downconv`: conv2d, strides [1, 2, 1, 1], SAME, kwidth 31;
 * utils/ops.py:277-310 `deconv`: conv2d_transpose with the same geometry; consumer models/discriminator.py:38-90).
 * Same channels-last sequence layout, sequence pitch S even, pl = SAME pad before (oracle/rsr_oracle.py same_pad):
 *   downconv forward   Y[o] = sum_k X[2o + k - pl] W[k]: rsr_gemm over the view A[m, k*Cp + c] = x[2m - pl + k, c]
 *                      (pointer x - pl*Cp, lda = 2*Cp, K = kwidth*Cp): TMA reads every second window
 *   weight gradient    dW += A(X)^T dY  (a_mn = 1) over the same view
 *   data gradient      dX[2u + r] = sum_j dY[u + e - j] W[2j + c]^T, one rsr_gemm per output parity r over a stride-1
 *                      window of dY with the taps of that phase (rsr_conv_w_phase), rows written with ldc = 2*Cp
 *   deconv             forward = that two-phase product, its gradients = the strided-window products.
 * rsr_conv_w_phase: out[q][b][a] = w[step*(nj-1-q) + phase][a][b] for the nj taps of index = phase mod step
 * (w16 [W, Ap, Bp] -> out16 [nj, Bp, Ap]). */
int rsr_conv_w_phase(rsr_handle* h, void* stream, const void* w16, int W, int Ap, int Bp, int step, int phase,
                     void* out16);

/* Virtual batch normalisation (utils/bnorm.py:11-69) on an fp32 pre-activation z [rows, N] (rows = batch x time):
 *   mean = w mean_B + (1-w) mean_ref, mean_sq likewise, std = sqrt(eps + mean_sq - mean^2), y = (z-mean)/std*gamma+beta
 * rsr_vbn_stats fills coef [8, N] like rsr_bn_train_stats (rows 0-3: A, B, mean, 1/std; then apply with
 * rsr_affine_act_drop).  ref_stats NULL (batch_weight 1) = the reference pass (:31-35); stats_out [2, N] (optional)
 * receives (mean_B, mean_sq_B), the reference statistics of later live passes (batch_weight = 1 / (batch + 1), :42-47).
 * rsr_vbn_bwd = rsr_bn_bwd through those statistics, the batch counted with batch_weight:
 *   dz = A (g - w mean(g) - w x_hat mean(g x_hat)), dgamma += sum g x_hat, dbeta += sum g,  g = da act'(y). */
int rsr_vbn_stats(rsr_handle* h, void* stream, const float* z, int ldz, long long rows, int N, const float* gamma,
                  const float* beta, float eps, float batch_weight, const float* ref_stats, float* stats_out,
                  float* coef, float* scratch);
int rsr_vbn_bwd(rsr_handle* h, void* stream, const void* da16, int ldda, const float* z, int ldz, long long rows, int N,
                int act, float batch_weight, float* coef, float* dgamma, float* dbeta, void* dz16, int lddz,
                float* dz32, int lddz32, float* scratch);

/* batch_norm(renorm) / dropout behind a fully_connected ----------------------------------- */
/* tf.contrib.layers.fully_connected(..., normalizer_fn=batch_norm, normalizer_params={is_training, scale=True,
 * renorm=True}) followed by tf.nn.dropout -- models/dnn.py:56-62,79-94, models/discriminator_dnn.py:36-46,61-83,
 * models/lstm.py:61-67,82-87.  With a normalizer the layer has no bias; contrib batch_norm defaults: decay 0.999,
 * epsilon 1e-3, center=True, renorm_decay 0.99, no renorm clipping; reduction over every row (all axes but the last).
 * The GEMM writes the fp32 pre-activation z [rows, N] (rsr_gemm, out32, no bias / activation); these entries do the
 * rest as HBM streams.  N and every leading dimension are multiples of 4.
 *
 *   state fp32 [6, N]: moving_mean, moving_variance, renorm_mean, renorm_stddev, renorm_mean_weight,
 *                      renorm_stddev_weight (the two weights are TF scalars, replicated per column)
 *   coef  fp32 [8, N]: A, B (y = z A + B), mean, 1/stddev, r, d  (written by the forward entries) and the two column
 *                      means the backward pass derives (rows 6, 7)
 *   scratch fp32 [RSR_BN_SCRATCH_FLOATS(N)]: per-split partial moments; calls sharing one scratch must be stream-ordered
 *
 * rsr_bn_train_stats (is_training=True): mean/variance of z over rows (Welford per thread, fixed-order merges -- no
 *   atomics), stddev = sqrt(var + eps); r = stddev / (renorm_stddev + (1 - renorm_stddev_weight) stddev),
 *   d = (mean - (renorm_mean + (1 - renorm_mean_weight) mean)) / (same denominator), both constants for the gradient;
 *   A = r gamma / stddev, B = d gamma + beta - mean A.  update_state != 0 also runs the UPDATE_OPS: renorm averages with
 *   renorm_momentum, then moving_mean / moving_variance with `momentum` towards the de-biased renorm values
 *   (models/dnn_trainer_single_gpu.py:101-104 runs them; models/gan_rnn_placeholder.py:169-175 does not).
 * rsr_bn_eval_coef (is_training=False): A = gamma rsqrt(moving_variance + eps), B = beta - moving_mean A.
 * rsr_affine_act_drop: out16 / out32 (either may be NULL) = dropout(act(z A + B)); A NULL = 1 (plain bias in B).  With
 *   A NULL, B = 0 and no activation it is tf.contrib.rnn.DropoutWrapper(output_keep_prob) on an LSTM layer's output
 *   (models/lstm.py:99-102, models/res_lstm_l.py:96-99), and rsr_bn_bwd(bn = 0) its gradient.  keep_prob >= 1: no dropout;
 *   otherwise one hash serves each (even, odd) column pair: with i = r N + c, h = splitmix64(key ^ (i >> 1)) and
 *   key = splitmix64(rng[0] + 0x9E3779B97F4A7C15 (rng[1] 65536 + salt)), element (r, c) is kept iff
 *   (c even ? h >> 40 : (h >> 16) & 0xffffff) < floor(keep_prob 2^24) (keep_prob as the float32 passed); kept values are
 *   divided by keep_prob (tf.nn.dropout).  rng = device {seed, tick}; rsr_rng_tick advances tick (once per update), salt
 *   names the layer call.
 * rsr_bn_bwd: da16 = gradient wrt the layer OUTPUT.  g = da act'(y) [kept / keep_prob];  dbeta += sum g;
 *   bn != 0: dgamma += r sum(g x_hat) + d sum(g),  dz = A (g - mean(g) - x_hat mean(g x_hat));   bn == 0 (bias + activation
 *   + dropout only; `bias` replaces coef): dz = g.  The mask is regenerated from the same (rng, salt).
 *   dgamma / dbeta / dz16 / dz32 (an fp32 copy of dz) may be NULL. */
#define RSR_BN_SCRATCH_FLOATS(N) (768LL * (N))
int rsr_bn_train_stats(rsr_handle* h, void* stream, const float* z, int ldz, long long rows, int N,
                       const float* gamma, const float* beta, float eps, float* state, float momentum,
                       float renorm_momentum, int update_state, float* coef, float* scratch);
/* rsr_bn_train_stats without its first pass: `scratch` already holds `splits` row-block partials [split][3][N] of
 * (count, mean, M2), written by the epilogue of the rsr_gemm that produced z (rsr_gemm_args.stats) */
int rsr_bn_train_finish(rsr_handle* h, void* stream, int splits, long long rows, int N, const float* gamma,
                        const float* beta, float eps, float* state, float momentum, float renorm_momentum,
                        int update_state, float* coef, const float* scratch);
int rsr_bn_eval_coef(rsr_handle* h, void* stream, int N, const float* gamma, const float* beta, float eps,
                     const float* state, float* coef);
int rsr_affine_act_drop(rsr_handle* h, void* stream, const float* z, int ldz, long long rows, int N,
                        const float* A, const float* Bc, int act, float keep_prob,
                        const unsigned long long* rng, unsigned salt, void* out16, int ld16, float* out32, int ld32);
int rsr_bn_bwd(rsr_handle* h, void* stream, const void* da16, int ldda, const float* z, int ldz,
               long long rows, int N, int act, float keep_prob, const unsigned long long* rng, unsigned salt,
               int bn, float* coef, const float* bias, float* dgamma, float* dbeta, void* dz16, int lddz,
               float* dz32, int lddz32, float* scratch);
int rsr_rng_tick(rsr_handle* h, void* stream, unsigned long long* rng);
/* The discriminator's input noise, tf.random_normal(shape = (B, 1, D), stddev) broadcast over time (utils/ops.py:19-30,
 * models/discriminator_lstm.py:54-60): out[i] = stddev * sqrt(-2 ln u1) cos(2 pi u2), i < n, with u1 = (h >> 40 + 1) /
 * 2^24 and u2 = ((h >> 16) & 0xffffff) / 2^24 of h = splitmix64(key ^ i), key as for the dropout mask above (TensorFlow's
 * random stream cannot be reproduced; parity tests feed the noise explicitly).  The caller ticks rng after each draw. */
int rsr_gauss_noise(rsr_handle* h, void* stream, const unsigned long long* rng, unsigned salt, float* out,
                    long long n, float stddev);

/* batch_norm behind the convolutions -- tf.contrib.layers.conv2d(..., normalizer_fn=batch_norm, normalizer_params=
 * {is_training, scale=True, renorm=True}), models/rced.py:63-71,94-97: no bias, the normalised axis is the conv2d
 * channel and the moments pool every (frame, line, position).  In the frame layout of the convolution family the
 * pre-activation z is [frames * S, N] fp32 with channel ch of line l in column l * C + ch (l < H; H = 1 for the
 * [1, w] convolutions) and only rows r % S < L are data; gamma, beta [C] and state [6, lds >= C] are per channel,
 * coef [8, N] is per column (the H columns of a channel hold the same values, columns >= H * C hold zeros) so the
 * apply kernels stay per column.  Same arithmetic as the fully-connected entries above; count = frames * H * L.
 *   rsr_bn_train_stats_lines / rsr_bn_eval_coef_lines   coefficients (+ UPDATE_OPS)
 *   rsr_affine_act_lines   out16 = act(z A + B) on data rows, 0 on the padding rows (replaces rsr_conv_mask_rows)
 *   rsr_bn_bwd_lines       da16 = gradient wrt the layer OUTPUT (any values on padding rows); dz16 = gradient wrt z,
 *                          0 on padding rows; dgamma, dbeta [C] accumulate */
int rsr_bn_train_stats_lines(rsr_handle* h, void* stream, const float* z, int ldz, long long frames, int S, int L, int H,
                             int C, int N, const float* gamma, const float* beta, float eps, float* state, int lds,
                             float momentum, float renorm_momentum, int update_state, float* coef, float* scratch);
int rsr_bn_eval_coef_lines(rsr_handle* h, void* stream, int N, int H, int C, const float* gamma, const float* beta,
                           float eps, const float* state, int lds, float* coef);
int rsr_affine_act_lines(rsr_handle* h, void* stream, const float* z, int ldz, long long frames, int S, int L, int N,
                         const float* A, const float* Bc, int act, void* out16, int ld16);
int rsr_bn_bwd_lines(rsr_handle* h, void* stream, const void* da16, int ldda, const float* z, int ldz, long long frames,
                     int S, int L, int H, int C, int N, int act, float* coef, float* dgamma, float* dbeta, void* dz16,
                     int lddz, float* scratch);

/* Kaldi compressed matrix ("CM" ark entries) decoded on the device -- io_funcs/kaldi_io.py:121-161 (uint16_to_float,
 * char_to_float, read_compress) fused with the CMVN of io_funcs/make_tfrecords.py:84-87.
 *   col_hdr  u16 [cols, 4]    PerColHeader percentiles (0, 25, 75, 100), as on disk (little endian)
 *   data     u8  [cols, rows] the byte matrix, column-major as on disk
 *   min_value, range          GlobalHeader floats
 *   out64    f64 [rows, ld64] the reader's float64 matrix, or NULL
 *   out32    f32 [rows, ld32] float32((x - mean[c]) / std[c]) evaluated in float64 (mean/std f64 [cols]), or float32(x)
 *                             when mean/std are NULL; or NULL
 * float64 arithmetic in the reference's operation order with round-to-nearest intrinsics: BIT-IDENTICAL to the Python
 * reader (tests/golden/kaldi_small_expected.npz was produced by the reference's own reader). */
int rsr_ark_decompress(rsr_handle* h, void* stream, const void* col_hdr, const void* data, float min_value, float range,
                       int rows, int cols, double* out64, int ld64, const double* mean, const double* std,
                       float* out32, int ld32);

/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum of TensorFlow checkpoint-V2
 * tensor-bundle files, used by rsrgan_b200/tf_checkpoint.py to read / write the reference's `GAN_RNN-<step>`
 * checkpoints (models/gan_rnn_placeholder.py:26-60).  Needs no device.  Returns the checksum (not a status). */
unsigned int rsr_crc32c_host(const void* data_host, unsigned long long n, unsigned int crc);

/* average_gradients over NVLink peer memory ------------------------------------------------
 * utils/ops.py:343-376 (the tower mean behind both optimizers, models/gan_rnn_placeholder.py:144-160) for one rank per
 * GPU on one node: the flat gradient buffers live in blocks the other ranks have opened through CUDA IPC, and ONE
 * kernel per update sums them in place over NVLink -- rank r reads slice r of all `world` buffers, adds them in rank
 * order 0..world-1 and stores the sum into all of them, between two flag barriers (csrc/peer_allreduce.cu).  The result
 * is the SUM (the 1/world of the mean is the gmul of rsr_seg_sumsq / rsr_clip_*_ema) and is bit-identical on every rank.
 * Stream-ordered like every other entry and capturable into a CUDA graph; every rank must issue the same calls in the
 * same order.  A barrier that waits longer than ~30 s gives up and raises the block's error flag (rsr_peer_error).
 *   rsr_peer_alloc   cudaMalloc of RSR_PEER_HEADER_BYTES + data_bytes (zeroed) + its 64-byte IPC handle
 *   rsr_peer_open    maps another rank's block from its handle;  rsr_peer_close unmaps it;  rsr_peer_free releases one's own
 *   rsr_peer_allreduce  blocks[world] = every rank's block (own at [rank]); the buffer is n_floats (multiple of 4) at byte
 *                    offset data_off_bytes (multiple of 16, >= RSR_PEER_HEADER_BYTES) of EVERY block; world in {1, 2, 4, 8};
 *                    max_blocks caps the grid (0 = 148; identical on all ranks) */
#define RSR_PEER_MAX_RANKS 8
#define RSR_PEER_HEADER_BYTES 16384
#define RSR_PEER_IPC_HANDLE_BYTES 64
int rsr_peer_alloc(rsr_handle* h, long long data_bytes, void** block, unsigned char* ipc_handle);
int rsr_peer_open(rsr_handle* h, const unsigned char* ipc_handle, void** block);
int rsr_peer_close(rsr_handle* h, void* block);
int rsr_peer_free(rsr_handle* h, void* block);
int rsr_peer_error(rsr_handle* h, const void* block, int* error);
int rsr_peer_allreduce(rsr_handle* h, void* stream, void* const* blocks, int rank, int world, long long data_off_bytes,
                       long long n_floats, int max_blocks);

/* misc ---------------------------------------------------------------------------------- */
/* dst[c, r] = src[r, c] for r < rows, c < cols (16-bit elements; other elements of dst untouched):
 * keeps the K_x^T operand of rsr_lstmp_fused_fwd in step with the updated weights. */
int rsr_transpose16(rsr_handle* h, void* stream, const void* src, int ld_src, int rows, int cols, void* dst,
                    int ld_dst);
int rsr_cast16(rsr_handle* h, void* stream, const float* x, long long n, void* out16);
int rsr_fill32(rsr_handle* h, void* stream, float* x, long long n, float v);

#ifdef __cplusplus
}
#endif
#endif /* RSRGAN_B200_H_ */
