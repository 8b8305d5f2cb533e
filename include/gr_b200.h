/*
 * gr_b200.h -- C ABI of the B200-native BLSTM -> softmax -> CTC -> decode hot path.
 *
 * The reference (AlexGidiotis/Multimodal-Gesture-Recognition-with-LSTMs-and-CTC) has no FFI
 * layer: its boundary is the Python operator surface (Keras callables).  Every entry point
 * below is the device-side replacement of one such callable; the Python host package mirrors
 * the callable itself (same names / argument meaning / error behaviour) and forwards raw device
 * pointers here.  See INTEGRATION.md for the reference-side binding.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no framework types.  `stream` is a cudaStream_t
 *     passed as void*; all work is enqueued on it, nothing synchronises the device.
 *   - Every pointer is a DEVICE pointer unless the name ends in `_host`.  The caller owns every
 *     buffer (including workspaces, sized by the *_workspace_bytes queries); the library never
 *     allocates or frees device memory and keeps no global state => re-entrant, stream-ordered,
 *     callable from several host threads on different streams.
 *   - Return value: GR_OK (0) or a negative GR_E* code for argument errors (checked on the host,
 *     synchronously).  Per-sequence DATA errors (the ones TensorFlow raises at session.run) are
 *     written to a device `status` array and raised by the Python wrapper.
 *   - Row-major, batch-major tensors: (B, T, C) means ((b*T)+t)*C + c.  fp32 unless stated.
 */
#ifndef GR_B200_H_
#define GR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR_OK 0
#define GR_EINVAL (-1)      /* bad argument (null pointer, non-positive size, unsupported shape) */
#define GR_EWORKSPACE (-2)  /* workspace too small */
#define GR_ECUDA (-3)       /* CUDA runtime error at launch; see gr_last_error() */
#define GR_EUNSUPPORTED (-4)

/* per-sequence status codes written by gr_ctc_loss_grad_f32 (mirror TF CTCLossOp errors) */
#define GR_CTC_OK 0
#define GR_CTC_ZERO_LABELS 1     /* "Labels length is zero in batch b" */
#define GR_CTC_NOT_ENOUGH_TIME 2 /* "Not enough time for target transition sequence" */
#define GR_CTC_NONNULL_AFTER_NULL 3 /* "Saw a non-null label ... following a null label" */
#define GR_CTC_NO_VALID_PATH 4   /* loss = +inf, grad = softmax (TF warning, not an error) */
#define GR_CTC_BAD_INPUT_LENGTH 5 /* input_length > T - drop_frames, or < 1 */

int gr_version(void);
/* last CUDA error string seen by this thread's most recent failing call ("" if none). */
const char* gr_last_error(void);

/* ------------------------------------------------------------------------------------------
 * CTC loss + gradient.  Replaces `ctc_lambda_func(args)` -> `K.ctc_batch_cost`
 * (/root/reference/audio_network/losses.py:4-15, multimodal_fusion/losses.py:4-15; Lambda call
 * sites speech_lstm_ctc_words.py:107-109, multimodal.py:197-199).
 *
 *   x            (B, T, C)   input_is_logits == 0: softmax PROBABILITIES y_pred (the reference
 *                            contract); grad_out is d/d y_pred.
 *                            input_is_logits == 1: head logits (pre-softmax Dense output); the
 *                            model softmax is fused and grad_out is d/d logits.
 *   drop_frames              leading frames ignored (the reference's `y_pred[:, 2:, :]`); their
 *                            gradient rows are written as zero.
 *   eps                      Keras ctc_batch_cost epsilon (1e-8 in the pinned Keras 2.1.4).
 *   labels       (B, Lmax)   int32; entries >= label_len[b] are never read.  blank = C-1; a
 *                            label >= C-1 terminates the sequence (TF rule).
 *   label_len, input_len (B) int32; input_len counts frames AFTER the drop.
 *   upstream     (B) or NULL per-sequence multiplier of the gradient (NULL = 1).
 *   loss         (B)         -log p(l|x)
 *   grad_out     (B, T, C) or NULL (loss only)
 *   status       (B) int32   GR_CTC_* code per sequence (may be NULL)
 */
int gr_ctc_workspace_bytes(int B, int T, int C, int Lmax, size_t* bytes_out);
int gr_ctc_loss_grad_f32(const float* x, int input_is_logits, int B, int T, int C,
                         int drop_frames, float eps, const int32_t* labels, int Lmax,
                         const int32_t* label_len, const int32_t* input_len,
                         const float* upstream, float* loss, float* grad_out, int32_t* status,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decoders.  `gr_ctc_bestpath_ref_f32` replaces the numeric part of `decode_batch`
 * (/root/reference/multimodal_fusion/sequence_decoding.py:38-50; audio_network/
 * sequence_decoding.py:38-50): per-frame argmax/max on frames >= drop_frames, the COUNT-based
 * confidence filter of the Python-2 loop at :45-48, and the adjacent-repeat collapse at :50.
 * Blank is kept.  out_ids (N, T) int32 padded with -1, out_len (N).
 * threshold is compared in double precision, as Python does (0.97 is not an fp32 number).
 *
 * `gr_ctc_greedy_f32` / `gr_ctc_beam_f32`: `K.ctc_decode` -> TF ctc_greedy_decoder /
 * CTCBeamSearchDecoder semantics (no call site in the reference; BASELINE.json config 5).
 * probs (N, T, C) are softmax probabilities, inputs = log(p + eps).  seq_len (N) or NULL (= T).
 * out_score: greedy -> -sum_t max_c log(p+eps); beam -> log-probability of each returned path.
 */
int gr_ctc_bestpath_ref_f32(const float* probs, int N, int T, int C, int drop_frames,
                            double threshold, int32_t* out_ids, int32_t* out_len, void* stream);
int gr_ctc_greedy_f32(const float* probs, int N, int T, int C, const int32_t* seq_len, float eps,
                      int32_t* out_ids, int32_t* out_len, float* out_score, void* stream);
int gr_ctc_beam_workspace_bytes(int N, int T, int C, int beam_width, size_t* bytes_out);
int gr_ctc_beam_f32(const float* probs, int N, int T, int C, const int32_t* seq_len, float eps,
                    int beam_width, int top_paths, int merge_repeated,
                    int32_t* out_ids /* (N, top_paths, T) */, int32_t* out_len /* (N, top_paths) */,
                    float* out_logp /* (N, top_paths) */, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Bidirectional LSTM.  Replaces `Bidirectional(LSTM(H, activation='tanh',
 * recurrent_activation='hard_sigmoid', return_sequences=True), merge_mode='concat')(x)`
 * (speech_lstm_ctc_words.py:56-77, skeletal_lstm_ctc.py:309-331, multimodal.py:159-168).
 *
 * Column order of every "8H" axis: dir*4H + gate*H + j, dir 0 = forward layer, 1 = backward
 * layer, gate order i,f,c,o (Keras).  Weights are the Keras arrays themselves:
 *   W  (F, 8H)  = [fwd kernel | bwd kernel]      (each (F,4H))
 *   U  (2, H, 4H) = fwd recurrent_kernel, bwd recurrent_kernel
 *   b  (8H)     = [fwd bias | bwd bias]
 *
 * The op is split in the two stages the hardware wants:
 *   (1) gr_gemm_*: the hoisted input projection P = X W + b over all B*T rows (tensor cores);
 *   (2) gr_lstm_recurrence_fwd_f32: the serial part.  `gates` (B, T, 8H) holds P on entry and
 *       the post-activation gates i,f,g,o on exit (in place); `cell` (B, T, 2H) receives c_t;
 *       y (B, T, 2H) receives [h_fwd | h_bwd] (backward direction already re-reversed in time).
 *   Backward: gr_lstm_recurrence_bwd_f32 consumes dy (B,T,2H), the saved gates/cell and y, and
 *   overwrites `gates` with dP (gradient wrt the pre-activations, (B,T,8H)); the weight/input
 *   gradients are then plain GEMMs on dP (dW = X^T dP, dU = Hprev^T dP, db = colsum dP,
 *   dX = dP W^T), issued by the host through gr_gemm_*.
 *   workspace: >= gr_lstm_workspace_bytes(B, H) bytes of device memory (exchange buffers,
 *   carried cell state, barrier counters; initialised by the call itself).
 */
int gr_lstm_workspace_bytes(int B, int H, size_t* bytes_out);
int gr_lstm_recurrence_fwd_f32(float* gates, const float* U, int B, int T, int H, float* y,
                               float* cell /* may be NULL: inference, c_t not kept */,
                               void* workspace, size_t workspace_bytes, void* stream);
/* The same forward recurrence with an AUXILIARY destination for h: `aux` points at the first of 2H consecutive columns of a
 * (B*T, ld_aux) fp32 matrix -- e.g. this tower's block of the fusion model's Merge(concat) buffer (multimodal.py:155-156).
 * accumulate == 0: h is stored there as well; accumulate != 0: h is ADDED to what is there (TMA reduce-add), which is how the
 * residual `add([blstm_1, blstm_2])` (speech_lstm_ctc_words.py:79, multimodal.py:119-133) lands in the concat buffer with no
 * separate pass: layer 1 stores, layer 2 accumulates.  y may be NULL (h is then written to aux only); cell as above.
 * Only the tensor-memory kernel of the wide layers has this output: gr_lstm_recurrence_aux_supported(B, H) != 0,
 * else GR_EUNSUPPORTED.  aux, ld_aux and H must keep 16-byte alignment. */
int gr_lstm_recurrence_aux_supported(int B, int H);
/* CTAs (= SMs held for the whole launch) of the tensor-memory recurrence kernels for this shape, 0 when another
 * kernel would run (narrow layers, unsupported H).  Lets the host decide whether two half-batch launches fit side by side. */
int gr_lstm_recurrence_grid(int B, int H);
int gr_lstm_recurrence_fwd_aux_f32(float* gates, const float* U, int B, int T, int H, float* y, float* cell,
                                   float* aux, int ld_aux, int accumulate, void* workspace,
                                   size_t workspace_bytes, void* stream);
int gr_lstm_recurrence_bwd_f32(float* gates /* in: i,f,g,o ; out: dP */, const float* cell,
                               const float* dy, const float* U, int B, int T, int H,
                               void* workspace, size_t workspace_bytes, void* stream);

/* C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N]) in fp32-faithful arithmetic on tcgen05 tensor cores:
 * each fp32 operand is split into bf16 hi + bf16 lo and three MMAs (hi*hi + hi*lo + lo*hi)
 * accumulate in fp32 TMEM ("bf16x3").  passes = 1 uses hi*hi only (plain bf16).
 * A and B are given pre-split: a_hi/a_lo (M, K) bf16 row-major (K contiguous), b_hi/b_lo (N, K).
 * lda/ldb = row strides in elements; K, lda, ldb multiples of 8 (16-byte TMA rows); M, N
 * arbitrary.  ldc = row stride of C in floats.  accumulate != 0: C += result.  Split-K (fp32
 * red.add into C) is chosen internally when the tile count cannot fill the SMs. */
int gr_gemm_bf16x3_f32(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                       const float* bias, float* C, int ldc, int M, int N, int K, int lda, int ldb,
                       int passes, int accumulate, void* stream);
/* Projection GEMM with the fp32 prologue FUSED (no bf16 copy of the activations in HBM):
 *   for v in [0, nvar):  C[:, v*Nv:(v+1)*Nv] (+)= op(A, mask_v) * B[v*Nv:(v+1)*Nv, :]^T + bias
 * A fp32: transA == 0 -> (M, K) row-major, mask_v[(m / rows_per_seq) * K + k];
 *         transA != 0 -> (K, M) row-major, element (m,k) = A[(k + row_shift) * lda + m] when row
 *         k + row_shift lies in the same sequence of rows_per_seq rows (else 0),
 *         mask_v[(k / rows_per_seq) * M + m].   mask: (nvar, n_sequences, K or M) or NULL --
 *         the per-gate LSTM input-dropout masks (speech_lstm_ctc_words.py:61,73).
 * B: pre-split bf16 hi/lo, (nvar*Nv, ldb) K-major, zero padded to ldb (multiple of 8) columns. */
int gr_gemm_a32_f32(const float* A, int lda, int transA, int row_shift, const float* mask,
                    int rows_per_seq, int nvar, const void* b_hi, const void* b_lo, int ldb,
                    const float* bias, float* C, int ldc, int M, int Nv, int K, int accumulate,
                    void* stream);
/* Same contraction for masks that are Keras DROPOUT masks: every mask element is either 0 or mask_scale
 * (= 1/(1-rate), what K.dropout multiplies kept inputs by; speech_lstm_ctc_words.py:61,73 `dropout=`).  The kernel
 * then splits each loaded fp32 tile once for four variants, keeps/zeroes the packed bf16 words per variant and
 * applies mask_scale to the accumulator, i.e. it returns mask_scale * ((A o [mask != 0]) B^T) + bias, which equals
 * the generic result up to fp32 rounding of the products.  A mask element that is neither 0 nor mask_scale is
 * treated as mask_scale.  mask == NULL: identical to gr_gemm_a32_f32. */
int gr_gemm_a32_dropout_f32(const float* A, int lda, int transA, int row_shift, const float* mask,
                            float mask_scale, int rows_per_seq, int nvar, const void* b_hi, const void* b_lo,
                            int ldb, const float* bias, float* C, int ldc, int M, int Nv, int K,
                            int accumulate, void* stream);
/* plain fp32 CUDA-core GEMM with the same contract on unsplit operands (cross-check only). */
int gr_gemm_simt_f32(const float* A, const float* B, const float* bias, float* C, int M, int N,
                     int K, int lda, int ldb, int ldc, int accumulate, void* stream);

/* fp32 -> (bf16 hi, bf16 lo) split with optional fused prologue, the A-operand producer:
 *   v = x[r, k]; if add: v += add[r, k] (residual `layers.add`, speech:79); if noise: v += noise[r,k]
 *   (GaussianNoise, speech:53); if mask: v *= mask[(r / rows_per_seq), k] (LSTM input dropout,
 *   one mask row per sequence, constant over time).  transpose != 0 writes (K, R) instead of
 *   (R, K); with transpose, row_shift s makes output column r read input row r+s when that row
 *   is in the same sequence, else 0 (pairs h_{t-1}/h_{t+1} with dP_t for the dU contraction).
 *   ld_out = row stride of the outputs in elements (>= padded width, pad written as zero). */
int gr_split_bf16_f32(const float* x, const float* add, const float* noise, const float* mask,
                      int rows_per_seq, int R, int K, int ldx, int transpose, int row_shift,
                      void* out_hi, void* out_lo, int ld_out, void* stream);
/* out[r,k] (+)= tmp[r,k] * mask[r / rows_per_seq, k]: the input-dropout mask applied to dX. */
int gr_mask_mul_acc_f32(float* out, const float* tmp, const float* mask, int rows_per_seq, int R,
                        int K, int accumulate, void* stream);

/* Dense(C) + softmax head (speech:86-90, multimodal:175-179): logits = x Wd + bd over R rows,
 * probs = softmax(logits).  x (R, Fin) with optional dropout mask (R, Fin) (Dropout layer,
 * speech:82).  Wd (Fin, C).  Either output may be NULL.  Backward: given g_logits (R, C):
 * dWd (Fin, C), dbd (C), dx (R, Fin) (NULL to skip). */
int gr_dense_softmax_fwd_f32(const float* x, const float* drop_mask, const float* Wd,
                             const float* bd, int R, int Fin, int C, float* logits, float* probs,
                             void* stream);
int gr_dense_bwd_f32(const float* x, const float* drop_mask, const float* Wd,
                     const float* g_logits, int R, int Fin, int C, float* dWd, float* dbd,
                     float* dx, void* stream);

/* column sums: out[n] = sum_r a[r, n]  (bias gradients). */
int gr_colsum_f32(const float* a, int R, int N, int lda, float* out, void* stream);

/* Elementwise helpers on the path: out = a + b (residual), concat along the last axis
 * (`Merge(mode='concat')`, multimodal.py:155-156: speech first). */
int gr_add_f32(const float* a, const float* b, float* out, size_t n, void* stream);
/* out[r, c] = a[r, c] + b[r, c] for c < cols, out with row stride ldo (floats): layers.add written into its
 * column block of the Merge(concat) buffer (multimodal_fusion/multimodal.py:111,117,155-156). */
int gr_add_into_f32(const float* a, const float* b, float* out, size_t rows, int cols, int ldo, void* stream);
int gr_concat2_f32(const float* a, int Fa, const float* b, int Fb, float* out, size_t rows,
                   void* stream);

/* Optimiser step of the reference (multimodal.py:206-208; speech:115-116): Adam(lr, beta1 .9,
 * beta2 .999, eps 1e-7, decay) with element-wise clipvalue, then maxnorm(max_norm) over axis 0
 * for `kernel` tensors (rows x cols, constraint applied per column) when max_norm > 0.
 * `step` is the 0-based iteration.  grad is read-only. */
int gr_adam_step_f32(float* param, const float* grad, float* m, float* v, size_t n, int rows,
                     int cols, float lr, float beta1, float beta2, float eps, float decay,
                     float clipvalue, float max_norm, int64_t step, void* stream);

/* Multi-tensor optimiser epilogue (one launch over the flat gradient bucket behind the all-reduce; the reference's
 * optimiser updates every trainable weight in one session.run: multimodal.py:206-208).  `srcs` / `params` / `sizes`
 * are HOST arrays of n_tensors (<= GR_MT_MAX) device pointers / element counts; flat buffers hold the tensors back to
 * back in that order.  gr_pack_f32: flat <- concat(srcs).  gr_adam_flat_f32: clipvalue + Keras Adam (lr decay as in
 * gr_adam_step_f32) on every tensor.  gr_maxnorm_f32: kernel_constraint=maxnorm (multimodal.py:165) on one (rows, cols). */
#define GR_MT_MAX 8
int gr_pack_f32(const float* const* srcs, const size_t* sizes, int n_tensors, float* flat, void* stream);
int gr_adam_flat_f32(float* const* params, const size_t* sizes, int n_tensors, const float* flat_grad,
                     float* flat_m, float* flat_v, float lr, float beta1, float beta2, float eps,
                     float decay, float clipvalue, int64_t step, void* stream);
int gr_maxnorm_f32(float* w, int rows, int cols, float max_norm, void* stream);

/* Philox-based regularisers (on-device RNG; the reference's GaussianNoise / Dropout /
 * LSTM input-dropout masks).  out[i] = keep ? 1/(1-p) : 0 ; noise[i] ~ N(0, stddev). */
int gr_dropout_mask_f32(float* out, size_t n, float p, uint64_t seed, uint64_t offset,
                        void* stream);
int gr_gaussian_noise_f32(float* out, size_t n, float stddev, uint64_t seed, uint64_t offset,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GR_B200_H_ */
