/*
 * b200quant.h — C ABI of libb200quant.so, the sm_100a implementation of the
 * per-layer weight-quantization arithmetic of vimarsh244/llm-quantization.
 *
 * The reference has no FFI of its own: its hot path is a set of plain Python
 * functions built from torch ops.  Every entry point below replaces one such
 * expression; the `ref:` tag names the reference file:line it stands in for.
 * The Python modules in llm-quantization_b200/ (same names and signatures as
 * the reference's) bind these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - matrices are row-major [N rows = out_features, K cols = in_features];
 *   - `dtype` is the storage type of the weights/activations (B200Q_F32 /
 *     B200Q_F16 / B200Q_BF16).  Arithmetic follows torch's eager semantics for
 *     that type: every elementwise op is evaluated in fp32 and rounded to the
 *     storage type, so fp32 inputs give results bit-identical to torch's;
 *   - outputs are caller-allocated; optional outputs may be NULL;
 *   - `stream` is a cudaStream_t passed as void*; no call synchronises;
 *   - return value: 0 on success, negative B200Q_E* otherwise, with a
 *     thread-local message from b200q_last_error();
 *   - there is no CPU path: every entry point launches sm_100a kernels.
 */
#ifndef B200QUANT_H_
#define B200QUANT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q_F32 0
#define B200Q_F16 1
#define B200Q_BF16 2

#define B200Q_OK 0
#define B200Q_EINVAL (-1)   /* bad argument (shape, alignment, dtype) */
#define B200Q_ECUDA (-2)    /* CUDA runtime / launch error */
#define B200Q_EUNSUPPORTED (-3)

/* pre-/post-operation applied per input column by b200q_group_fakequant */
#define B200Q_COLOP_NONE 0
#define B200Q_COLOP_MUL_DIV 1 /* w*=m[k] before, /=m[k] after   ref: awq_quantizer.py:70,81 */
#define B200Q_COLOP_DIV 2     /* w/=m[k] before, nothing after  ref: smooth_quant_quantizer.py:170 */

const char* b200q_last_error(void);
int b200q_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t b200q_launch_count(void);

/* ---- in-library kernel timing ---------------------------------------------------------
 * b200q_profile_enable(1) clears old records and makes every following entry point bracket
 * its launches with two CUDA events on the launching stream; (0) stops recording.
 * b200q_profile_query sums, for the entry point called `name` ("group_fakequant", "pot_quant",
 * ...; NULL = all), the device time, launch count and the ALGORITHMIC bytes / flops the calls
 * declared.  It waits for the recorded events. */
void b200q_profile_enable(int on);
int b200q_profile_query(const char* name, double* total_ms, int64_t* launches, double* bytes,
                        double* flops);

/* Device self-test of the exact reused-divisor division the fake-quant kernels use (reciprocal
 * + two Markstein FMA corrections) against __fdiv_rn on ~n_quotients random and adversarial
 * pairs.  *mismatches must come back 0.  Synchronises the stream. */
int b200q_selftest_div(int64_t n_quotients, uint64_t seed, int64_t* mismatches, void* stream);

/* ---- packed integer export (the reference stores no codes) -----
 * Each row of the uint8 code plane [N,K] (values < 2^n_bit, n_bit in [1,8]) becomes a
 * little-endian bit stream: code k occupies bits [k*n_bit, (k+1)*n_bit), stored as
 * b200q_packed_words_per_row(K, n_bit) = ceil(K*n_bit/32) uint32 words, zero padded. */
int64_t b200q_packed_words_per_row(int64_t K, int n_bit);
int b200q_pack_codes(const uint8_t* codes, int64_t N, int64_t K, int n_bit, uint32_t* packed,
                     void* stream);
int b200q_unpack_codes(const uint32_t* packed, int64_t N, int64_t K, int n_bit, uint8_t* codes,
                       void* stream);

/* ---- W4A16 GEMM on a packed record -----------------------------
 * Y[M, N] = X[M, K] * dequant(qweight)[N, K]^T: the nn.Linear forward of the reference's
 * perplexity loop (ref: quantization_utils.py:269-322) evaluated on the PACKED weight the export
 * produces -- 4-bit codes (eight per uint32, b200q_pack_codes layout), fp32 scale / zero point per
 * group of `group` input channels ((q - zero) * scale, ref: quantization_utils.py:405).  The
 * dequantisation runs inside the tcgen05 GEMM's operand pipeline; no fp16 copy of W is made.
 *   X: 16-bit activations (act_dtype = B200Q_F16 / B200Q_BF16), row-major, K % 8 == 0
 *   rec_dtype: dtype the weight was quantised in (its scale arithmetic is reproduced);
 *   Y: act_dtype, or fp32 when out_f32 != 0. */
int b200q_w4a16_gemm(const void* X, int64_t M, int64_t K, int act_dtype, const uint32_t* qweight,
                     const float* scales, const float* zeros, int64_t N, int64_t group,
                     int rec_dtype, void* Y, int out_f32, void* stream);

/* ---- torch-CPU log2 semantics, exported for the CPU test-suite -------------------
 * rne(log2f(r)) and floor(log2f(m)) as torch's CPU kernel evaluates them are step
 * functions of r; the library tabulates the step positions on the host at load
 * time (double log2 rounded to float).  These return the bit pattern of the smallest
 * positive float whose result is >= e+1 (round) / >= e (floor).
 * ref: pot_apot_quantizer.py:66,88,105 */
uint32_t b200q_log2_round_threshold_bits(int e);  /* e in [-127,127] */
uint32_t b200q_log2_floor_threshold_bits(int e);  /* e in [-149,127] */
/* the same for fp16 / bf16 tensors (dtype = B200Q_F16 / B200Q_BF16): torch rounds log2's result to
 * the tensor's type first, and the argument is a 16-bit value; 0x7f800000 = never reached */
uint32_t b200q_log2_round_threshold_bits_dt(int e, int dtype);
uint32_t b200q_log2_floor_threshold_bits_dt(int e, int dtype);

/* ---- column statistics -------------------------------------------------------------
 * colmax[k] = max_i |W[i,k]|, as fp32.   ref: gptq_quantizer.py:182 (per column over
 * all rows), smooth_quant_quantizer.py:156.  `accumulate`!=0 keeps the values already
 * in colmax (running max across row shards / calls); 0 overwrites. */
int b200q_col_absmax(const void* W, int64_t N, int64_t K, int64_t ld, int dtype,
                     float* colmax, int accumulate, void* stream);

/* ---- GPTQ column stage, reference-parity semantics ---------------------------------
 * s[k] = clamp(colmax[k]/(2^b-1), 1e-5); q = clamp(rne(W/s), -2^b, 2^b-1); out = q*s.
 * ref: gptq_quantizer.py:167-206 (closed form of the column loop; the reference applies
 * no error compensation, so perm/blocksize/H do not reach the output).
 * codes (int8, optional) and scales (fp32 [K], optional) expose the integers. */
int b200q_gptq_parity_quant(const void* W, void* out, int8_t* codes, const float* colmax,
                            float* scales, int64_t N, int64_t K, int64_t ld, int n_bit,
                            int dtype, void* stream);

/* ---- uniform group fake-quant ------------------------------------------------------
 * asymmetric (symmetric=0): ref quantization_utils.py:362-413 (pseudo_quantize_tensor)
 * symmetric  (symmetric=1): ref gptq_quantizer.py:79-108 (_simple_quantize_layer)
 * W is [N,K]; groups are `group` consecutive elements of a row (K % group == 0);
 * group <= 0 means one group per row.  colop/colvec: see B200Q_COLOP_*.
 * codes: uint8 (asym, [0,2^b-1]) or int8 (sym) when n_bit<=8; scales/zeros fp32 per group. */
int b200q_group_fakequant(const void* W, void* out, void* codes, float* scales, float* zeros,
                          int64_t N, int64_t K, int64_t group, int n_bit, int symmetric,
                          int colop, const float* colvec, int dtype, void* stream);

/* ---- SmoothQuant --------------------------------------------------------------------
 * s[k] = clamp( clamp(a[k],1e-5)^alpha / clamp(wmax[k],1e-5)^(1-alpha), 1e-5 )
 * ref: smooth_quant_quantizer.py:159-166.  act_dtype / w_dtype give the precision each
 * pow is rounded to (torch evaluates them in the tensors' own types). */
int b200q_smooth_scale(const float* act_scale, const float* wmax, float* s, int64_t K,
                       float alpha, int act_dtype, int w_dtype, void* stream);
/* out[i,k] = W[i,k] / s[k]   ref: smooth_quant_quantizer.py:170 ; mul!=0: W*s (:251) */
int b200q_col_scale(const void* W, void* out, const float* s, int64_t N, int64_t K, int mul,
                    int dtype, void* stream);

/* ---- activation statistics ---------------------------------------------------------
 * X is [T,K].  meanabs: ref quantization_utils.py:231 ; maxabs: ref
 * smooth_quant_quantizer.py:68 (+ running max :74 when accumulate!=0). fp32 out. */
int64_t b200q_act_stat_workspace(int64_t T, int64_t K); /* bytes of device scratch for meanabs */
int b200q_act_meanabs(const void* X, int64_t T, int64_t K, int dtype, float* out, void* work,
                      void* stream);
/* the same for a batch of n_samples equal-length samples stacked along the rows: out is
 * [n_samples, K]; work needs n_samples * b200q_act_stat_workspace(rows_per_sample, K) bytes */
int b200q_act_meanabs_batched(const void* X, int n_samples, int64_t rows_per_sample, int64_t K,
                              int dtype, float* out, void* work, void* stream);
int b200q_act_maxabs(const void* X, int64_t T, int64_t K, int dtype, float* out, int accumulate,
                     void* stream);
/* out[k] = ((0 + V[0,k]) + V[1,k]) + ... sequential, each partial sum rounded to V's dtype: the
 * order and precision of Python's sum() over a list of [K] tensors.   ref: awq_quantizer.py:57 */
int b200q_seq_sum_rows(const void* V, int64_t n, int64_t K, int dtype, float* out, void* stream);

/* colmul[i] = factor for the k largest entries of importance[K] (ties at the k-th value taken
 * in index order), 1 elsewhere; mask (uint8 [K], optional) flags them.
 * ref: awq_quantizer.py:60-61 (torch.topk) feeding :70,:81 */
int b200q_topk_colmul(const float* importance, int64_t K, int64_t k, float factor, float* colmul,
                      uint8_t* mask, void* stream);

/* ---- whole-layer entry points --------------------------------------------------------
 * One host call per nn.Linear: the per-layer launch sequence of a model walker, so that
 * host overhead stays below the kernels' HBM time.  `work` is device scratch of 3*K floats.
 *   awq_layer:        feats [n_feats,K] -> importance -> top n_protect -> fused scale/quant/unscale
 *                     ref: awq_quantizer.py:56-84
 *   gptq_parity_layer: column |max| (into colmax[K]) -> column quantisation
 *                     ref: gptq_quantizer.py:167-206 (single GPU; row shards all-reduce colmax
 *                     between b200q_col_absmax and b200q_gptq_parity_quant instead)
 *   smoothquant_layer: column |max| -> s[K] -> fused W/s + group fake-quant
 *                     ref: smooth_quant_quantizer.py:150-170,313 */
int b200q_awq_layer(const void* W, void* out, int64_t N, int64_t K, int64_t group, int n_bit,
                    const void* feats, int64_t n_feats, int feat_dtype, int64_t n_protect,
                    float scale_factor, float* work, uint8_t* salient_mask, int dtype,
                    void* stream);
int b200q_gptq_parity_layer(const void* W, void* out, int64_t N, int64_t K, int n_bit,
                            float* colmax, int dtype, void* stream);
int b200q_smoothquant_layer(const void* W, void* out, int64_t N, int64_t K, int64_t group,
                            int n_bit, const float* act_scale, float alpha, int act_dtype,
                            float* s, float* work, int dtype, void* stream);

/* ---- SmoothQuant alpha sweep (ref: smooth_quant_quantizer.py:327-371, a stub in the reference) --
 * err[a] (+)= sum_{i,k} ((Q(W[i,k] / S[a,k]) * S[a,k] - W[i,k]) * act_weight[k])^2 for a < n_alpha,
 * Q = the asymmetric group quantizer of b200q_group_fakequant.  S is [n_alpha, K] fp32 (one
 * smoothing-scale vector per candidate alpha, from b200q_smooth_scale), err is fp64 [n_alpha] on
 * the device (accumulate != 0 adds to it, so a model walk needs no host synchronisation).
 * W is read from HBM once for all alphas; no N x K temporary is written. */
int64_t b200q_smooth_alpha_workspace(int64_t N, int64_t K, int64_t group, int n_alpha);
int b200q_smooth_alpha_errors(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                              const float* S, int n_alpha, const float* act_weight, int dtype,
                              void* work, double* err, int accumulate, void* stream);

/* ---- POT ------------------------------------------------------------------------------
 * ref: pot_apot_quantizer.py:25-115.  w is [n_groups, group] contiguous; grid_host is the
 * HOST array torch.arange(0.01, 2.01, 0.01) materialised by the caller (n_grid floats).
 * Outputs: out (same dtype), exps (uint8 exponent code E, optional), best_scale (fp32 per
 * group, optional), best_idx (int32 per group, -1 = no candidate beat +inf, optional). */
int b200q_pot_quant(const void* w, void* out, uint8_t* exps, float* best_scale, int32_t* best_idx,
                    int64_t n_groups, int64_t group, int n_bit, const float* grid_host,
                    int n_grid, int dtype, void* stream);

/* ---- APOT -----------------------------------------------------------------------------
 * ref: pot_apot_quantizer.py:192-351.  levels_host: the signed, normalised, sorted level
 * set (<= 32 entries) built by the caller exactly as :227-247; grid_host as :262.
 * Outputs: out, level index per element (uint8, optional), best_scale, best_idx. */
int b200q_apot_quant(const void* w, void* out, uint8_t* level_idx, float* best_scale,
                     int32_t* best_idx, int64_t n_groups, int64_t group,
                     const float* levels_host, int n_levels, const float* grid_host, int n_grid,
                     int dtype, void* stream);

/* ---- GPTQ Hessian (tcgen05 tensor cores, TMA-fed) -------------------------------------------
 * H (+)= sum_i a_i^2 * X_i^T X_i  with a_i = 1/(||X_i||_F + 1e-5); X is [n_samples *
 * rows_per_sample, K] row-major and sample i is rows [i*rows_per_sample, (i+1)*rows_per_sample).
 * ref: gptq_quantizer.py:137-144 (1-D features are samples of one row, :140-141).
 * H is fp32 [K,K], fully written (both triangles); accumulate != 0 adds to its contents (ragged
 * sample lists are fed as several calls).  norms_out (optional, fp32 [n_samples]) receives
 * ||X_i||_F.  work: b200q_hessian_workspace(T, K, n_samples) bytes of device scratch.
 * K must be a multiple of 8.  normalize = 0 drops the per-sample factor (plain X^T X, the
 * Gram matrix the AWQ search measures its reconstruction error with).  fp16 / bf16 activations
 * are read in place by TMA (no staging copy) for normalize = 0, and for normalize = 1 when
 * rows_per_sample is a multiple of 64 and >= 512; fp32 input is staged as scaled fp16. */
int64_t b200q_hessian_workspace(int64_t T, int64_t K, int n_samples);
int b200q_hessian_accum(const void* X, int n_samples, int64_t rows_per_sample, int64_t K, int dtype,
                        int normalize, float* H, int accumulate, float* norms_out, void* work,
                        void* stream);
/* H = H * scale + damp * I     ref: gptq_quantizer.py:150 (scale = 1/len(input_feat), damp =
 * perp_damp) and :160 (scale = 1, damp = 1e-6) */
int b200q_hessian_finalize(float* H, int64_t K, float scale, float damp, void* stream);

/* ---- damped SPD inverse ------------------------------------------------------------------
 * For symmetric positive definite H (fp32 [K,K]): Hinv = inv(H) and/or U = the upper Cholesky
 * factor of inv(H) (U^T U = inv(H)), via blocked Cholesky, triangular inverse and L^-T L^-1.
 * ref: gptq_quantizer.py:160-165 (torch.linalg.inv(H + 1e-6 I); the caller adds the ridge with
 * b200q_hessian_finalize).  U is what the error-compensated column loop consumes.
 * Either of Hinv / U may be NULL.  work: b200q_spd_inverse_workspace(K) bytes.  info (device
 * int, optional, zero it first): 0 = ok, j > 0 = pivot j was not positive. */
int64_t b200q_spd_inverse_workspace(int64_t K);
int b200q_spd_inverse(const float* H, float* Hinv, float* U, int64_t K, void* work, int* info,
                      void* stream);

/* ---- GPTQ with error compensation (opt-in; NOT what the reference computes) ---------------
 * The blockwise column loop gptq_quantizer.py:173-197 sketches and then skips ("we skip error
 * compensation"), as in Frantar et al. 2022 Alg. 1: per column quantise with the asymmetric group
 * grid of pseudo_quantize_tensor, propagate e = (w - q)/U[j,j] into the remaining columns of the
 * 128-column block, then push the block's errors into all later columns with one rank-128 GEMM.
 * W: fp32 [N,K], overwritten; Q: fp32 [N,K] result; U: upper Cholesky factor of inv(H) from
 * b200q_spd_inverse.  group: 128, a multiple of 128, or <= 0 (per row).  blocksize must be 128.
 * work: b200q_gptq_compensated_workspace(N, K) bytes. */
int64_t b200q_gptq_compensated_workspace(int64_t N, int64_t K);
int b200q_gptq_compensated(float* W, float* Q, const float* U, int64_t N, int64_t K, int64_t group,
                           int n_bit, int blocksize, void* work, void* stream);

/* ---- AWQ scale search (tcgen05) -----------------------------------------------------------
 * loss[c] += sum_rows dW_c H dW_c^T = ||(Q_c(W) - W) X^T||_F^2 for H = X^T X, for every candidate
 * scale factor sf[c]; Q_c is awq_quantize_model_weight's arithmetic with scale_factor = sf[c] on
 * the columns flagged in `salient` (uint8 [K]).
 * ref: awq_quantizer.py:88-126 — a stub returning the midpoint; its docstring (:116-119) states
 * this search.  PARITY UNPINNED.  Stage 1 writes all candidates' dW as bf16 from one read of W,
 * stage 2 is one K-major bf16 GEMM over the stacked candidates with <dW H, dW> fused into the
 * epilogue.  group must be 128 and K a multiple of 128.  loss is accumulated into (zero it; row
 * shards on several GPUs all-reduce it).  sf_host: n_cand <= 32 floats on the host. */
int64_t b200q_awq_search_workspace(int64_t N, int64_t K, int n_cand);
int b200q_awq_search_loss(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                          const uint8_t* salient, const float* sf_host, int n_cand, const float* H,
                          int dtype, void* work, float* loss, void* stream);
/* same, with the folded operand Hb[n][k] = bf16(H[n][k] + H[k][n]) (k < n), bf16(H[n][n]) on the
 * diagonal, 0 above it, supplied by the caller as a [K,K] bf16 matrix (multi-GPU runs build it from
 * the packed exchange below and skip the fold pass). */
int b200q_awq_search_loss_folded(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                                 const uint8_t* salient, const float* sf_host, int n_cand,
                                 const void* Hb_folded, int dtype, void* work, float* loss,
                                 void* stream);
/* the two halves of b200q_awq_search_loss as separate calls, so that a caller can run the
 * candidate quantisation (which needs only W and the salient mask) on another stream WHILE the Gram
 * matrix is still being computed: _delta fills the workspace with dW_c = Q_c(W) - W (bf16) for all
 * candidates; _loss_prepared folds H (or takes Hb_folded; pass exactly one of the two), runs the
 * loss GEMM on the workspace and adds the per-candidate losses.  Same workspace, same N, K, n_cand
 * in both calls; the caller orders them. */
int b200q_awq_search_delta(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                           const uint8_t* salient, const float* sf_host, int n_cand, int dtype,
                           void* work, void* stream);
int b200q_awq_search_loss_prepared(int64_t N, int64_t K, int n_cand, const float* H,
                                   const void* Hb_folded, void* work, float* loss, void* stream);

/* ---- symmetric [K,K] matrices between GPUs: the packed lower triangle (new) -----
 * X^T X partial sums (ref: gptq_quantizer.py:144 accumulated over calibration samples dealt to
 * the ranks) are exactly symmetric, so they are exchanged as row n = columns 0..n at offset
 * n(n+1)/2: half the bytes of the square.  sym_fold_packed_bf16 turns packed elements [e0, e1)
 * (a rank's reduce-scattered slice, P_slice pointing at element e0; elements past the triangle's
 * end are padding and give 0) into the bf16 folded values of the AWQ search operand. */
int64_t b200q_sym_packed_len(int64_t K);
int b200q_sym_pack_lower(const float* H, int64_t K, float* P, void* stream);
int b200q_sym_unpack_lower(const float* P, int64_t K, float* H, void* stream);
int b200q_sym_fold_packed_bf16(const float* P_slice, int64_t K, int64_t e0, int64_t e1, void* out_bf16,
                               void* stream);
int b200q_sym_unpack_folded_bf16(const void* Pb, int64_t K, void* Hb, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200QUANT_H_ */
