/*
 * mmgclip_b200 -- C ABI of the B200-native contrastive hot path of abdel-habib/mmg-clip.
 *
 * The reference (pure Python / eager PyTorch) has no FFI of its own; its "plugin API" for this path is the set of
 * nn.Module / loss callables resolved by name in mmgclip/networks/projection_controller.py:3-24 and
 * mmgclip/loss/loss_controller.py:3-23.  The Python mirror of those classes lives in mmgclip_b200/ and reaches the
 * GPU only through the entry points below (ctypes).  Each entry point cites the reference lines
 * whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C: raw device pointers, explicit sizes/strides, no torch types, no exceptions across the boundary;
 *   - every call is asynchronous on the cudaStream_t passed as `stream` (an opaque void*), never synchronises the
 *     device and never allocates or frees user-visible memory (the caller owns all buffers, including workspaces);
 *   - return value: 0 = ok, negative = MMG_ERR_*; mmg_last_error_string() gives the thread-local reason;
 *   - there is NO CPU fallback: host pointers or a missing GPU are errors;
 *   - `prec` selects the arithmetic: MMG_PREC_FP32 = fp32 operands, fp32 FFMA accumulation (reference-faithful,
 *     parity 1e-5); MMG_PREC_BF16 = bf16 operands on tcgen05 tensor cores with fp32 accumulation in TMEM (2e-3);
 *     MMG_PREC_F16 (fused InfoNCE entry points only) = the same tensor-core path with the L2-NORMALISED EMBEDDING operands
 *     a_hat / b_hat stored as IEEE fp16 instead of bf16 (|x| <= 1, so the 3 extra mantissa bits cost no range and the
 *     tensor-core rate is the same) and the gradient coefficients as fp16 in 2^14-scaled units (see
 *     mmg_infonce_bwd_prep; tcgen05.mma kind::f16 cannot mix bf16 and fp16 operands in one instruction).  Embedding /
 *     weight gradients ~8x closer to float64 than with bf16 operands.
 */
#ifndef MMGCLIP_B200_H_
#define MMGCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMG_OK 0
#define MMG_ERR_BAD_ARG (-1)
#define MMG_ERR_BAD_ALIGN (-2)
#define MMG_ERR_UNSUPPORTED_SHAPE (-3)
#define MMG_ERR_CUDA (-4)
#define MMG_ERR_NO_DEVICE (-5)

#define MMG_PREC_FP32 0
#define MMG_PREC_BF16 1
#define MMG_PREC_F16 2

/* C store modes of the dense contraction */
#define MMG_STORE 0      /* C  = v                      */
#define MMG_ACCUMULATE 1 /* C += v  (caller guarantees exclusive ownership of C) */
#define MMG_ATOMIC_ADD 2 /* C += v  via red.global.add (needed when k_splits > 1) */

typedef void* mmg_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------------------- */
int mmg_version(void);                    /* 10000*major + 100*minor + patch */
const char* mmg_last_error_string(void);  /* thread-local, never NULL */
int mmg_device_info(int* sm_count, int* cc_major, int* cc_minor);
long long mmg_kernel_launch_count(void);   /* kernels launched by this library since load (process-wide) */

/* ---- dense contraction: C[M,N] (op)= alpha * A . B^T (+bias) (ReLU) ------------------------------------- */
/* Replaces nn.Linear forward / backward: mmgclip/networks/projection.py:17,33 (LinearProjectionLayer),
 * :45-59 (MultiLinearHead), :88-97 (MLPProjectionHead) and their autograd transposes.
 *   A: a_mn == 0 -> row-major [M, K] (lda = row pitch in elements), a_mn == 1 -> row-major [K, M];
 *   B: b_mn == 0 -> row-major [N, K],                               b_mn == 1 -> row-major [K, N];
 *   C: fp32 row-major [M, N]; bias: fp32 [N] or NULL; alpha_dev: optional DEVICE scalar multiplied into alpha
 *   (used for logits = s * A.B^T with s living on the device, mmgclip_model.py:132-136).
 * prec BF16: A, B are bf16 (pitches multiple of 8 elements, 16-byte aligned pointers), tcgen05/TMA kernel.
 * prec FP32: A, B are fp32, SIMT FFMA kernel.  k_splits > 1 requires mode == MMG_ATOMIC_ADD.
 * With MMG_ACCUMULATE the result is act(C_old + alpha*A.B^T + bias), i.e. bias/ReLU belong on the LAST pass. */
int mmg_gemm(int prec, const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, float* C,
             long long ldc, int M, int N, int K, float alpha, const float* alpha_dev, const float* bias, int relu,
             int mode, int k_splits, mmg_stream_t stream);

/* Split-precision bf16 contraction in ONE launch (what the projection heads use in bf16 mode):
 *   C (op)= alpha * ( A_hi.B_hi^T + A_hi.B_lo^T + A_lo.B_hi^T )        (A ~ A_hi + A_lo, B ~ B_hi + B_lo, all bf16)
 * run as a single contraction over the concatenated K range (3K), one TMEM accumulator, one epilogue -- instead of
 * three accumulate passes over C.  A_lo and/or B_lo may be NULL (their term is dropped).  Same layouts, modes and
 * split-K rules as mmg_gemm (k_splits divides the concatenated range). */
int mmg_gemm_split(const void* A_hi, const void* A_lo, long long lda, int a_mn, const void* B_hi, const void* B_lo,
                   long long ldb, int b_mn, float* C, long long ldc, int M, int N, int K, float alpha, const float* bias,
                   int relu, int mode, int k_splits, mmg_stream_t stream);

/* ---- element-wise helpers -------------------------------------------------------------------------------- */
int mmg_cast_f32_to_bf16(const float* x, void* y_bf16, long long n, mmg_stream_t stream);
int mmg_cast_f32_to_f16(const float* x, void* y_f16, long long n, mmg_stream_t stream);
/* hi = bf16(x), lo = bf16(x - hi): operands of the three-pass "bf16x3" contraction A_hi.B_hi + A_hi.B_lo + A_lo.B_hi that
 * keeps the projection heads fp32-faithful (~2^-17 per operand) on the bf16 tensor pipe. */
int mmg_cast_f32_to_bf16_split(const float* x, void* hi_bf16, void* lo_bf16, long long n, mmg_stream_t stream);

/* Push all-gather (multi-GPU exchange step 1): copy `bytes` bytes from `src` to dst_ptrs[i] + dst_offset_bytes
 * for every i < n_dst (<= 8).  dst_ptrs is a HOST array of device pointers -- this rank's and its peers' symmetric buffers
 * mapped into this process (NVLink peer memory); with dst_offset_bytes = rank * bytes and a cross-rank barrier afterwards
 * every rank holds all shards.  src, offset and length are multiples of 16 bytes.  The reference has no multi-GPU code. */
int mmg_push_rows(const void* src, long long bytes, void* const* dst_ptrs, int n_dst, long long dst_offset_bytes,
                  mmg_stream_t stream);

/* Row-wise L2 normalisation y = u / ||u||_2, no epsilon (mmgclip/networks/mmgclip_model.py:128-129,163;
 * mmgclip/evaluator.py:79,86).  inv_norm[B] is kept for the backward.  y_16 (nullable) receives a 16-bit operand copy for
 * the fused InfoNCE: fp16 when y16_is_f16 != 0 (MMG_PREC_F16), else bf16. */
int mmg_l2norm_fwd(const float* u, int B, int D, float* y, float* inv_norm, void* y_16, int y16_is_f16,
                   mmg_stream_t stream);
/* du = (dy - y * <y, dy>) * inv_norm   (autograd of the line above).  Outputs (each nullable, at least one): du fp32,
 * du_bf16 = bf16(du), du_bf16_lo = bf16(du - du_bf16) for the bf16x3 weight-gradient contraction.
 * zero_fill (nullable; 16-byte aligned, zero_floats % 4 == 0): a buffer the same launch clears -- the split-K output of the
 * weight-gradient contraction that follows accumulates with reduce-adds and would otherwise need a fill launch. */
int mmg_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int B, int D, float* du, void* du_bf16,
                   void* du_bf16_lo, float* zero_fill, long long zero_floats, mmg_stream_t stream);

/* Hidden-layer pieces of MultiLinearHead (projection.py:54-61): ReLU/inverted-dropout backward and bias gradient.
 *   dz[i] = dy[i] * (y ? y[i] > 0 : 1) * (mask ? mask[i] * keep_scale : 1)      db[n] = sum_rows dz[:, n]
 * (y = layer output after ReLU+dropout, NULL = no ReLU; mask NULL = no dropout) */
int mmg_dropout_apply(float* y, const uint8_t* mask, float keep_scale, long long n, mmg_stream_t stream);
/* nn.Dropout(p) in training mode (projection.py:51,59,92,98) as ONE launch that draws, applies and records the keep mask:
 *   keep[i] = philox4x32_10(counter = state[1] + i/4, key = state[0]).word[i % 4] >= p * 2^32
 *   y[i] = keep[i] ? y[i] / (1 - p) : 0        mask_out[i] = keep[i]
 * state = three DEVICE uint64 {seed, offset, 0}: the launch advances `offset` by ceil(n/4) itself (last CTA to retire), so
 * replayed CUDA graphs draw fresh masks.  Bit-for-bit the masks differ from torch's (different stream layout), as they do
 * between any two torch versions; the distribution is the same. */
int mmg_dropout_draw_apply(float* y, uint8_t* mask_out, float p, long long n, unsigned long long* state,
                           mmg_stream_t stream);
int mmg_relu_dropout_bwd(const float* dy, const float* y, const uint8_t* mask, float keep_scale, float* dz,
                         long long n, mmg_stream_t stream);
int mmg_colsum(const float* x, int rows, int cols, float* out, mmg_stream_t stream);

/* MLPProjectionHead pieces (projection.py:85-101): exact-erf GELU and LayerNorm(eps) over the last dim. */
int mmg_add(const float* x, const float* y, float* out, long long n, mmg_stream_t stream); /* residual x + y */
int mmg_gelu_fwd(const float* x, float* y, long long n, mmg_stream_t stream);
int mmg_gelu_bwd(const float* dy, const float* x, float* dx, long long n, mmg_stream_t stream);
int mmg_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int cols, float eps, float* y,
                      float* mean, float* rstd, mmg_stream_t stream);
int mmg_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      int rows, int cols, float* dx, float* dgamma, float* dbeta, mmg_stream_t stream);

/* ---- fused symmetric InfoNCE over normalised embeddings --------------------------------------------------- */
/* Replaces: logits_per_image / logits_per_text (mmgclip_model.py:135-136; losses.py:73-74,85-86) and the two
 * F.cross_entropy calls with labels = arange(n) (losses.py:39-43,78-82,88-91), forward and backward, WITHOUT
 * materialising the [rows x cols] logit matrix.  a_hat: [rows, D] (the local row shard), b_hat: [cols, D] (all
 * columns), fp32 or bf16 per `prec`, row-major, D contiguous.  `diag_offset` is the global column index paired with
 * local row 0 (0 on one GPU, rank*rows in the sharded case).  `scale` is a DEVICE pointer to s = exp(logit_scale).
 *
 * Forward accumulates (caller zero-fills first):
 *     rowsum[r] += sum_c exp(s*cos[r,c] - s)     colsum[c] += sum_r exp(s*cos[r,c] - s)     diag[r] = s*cos[r,r']
 * The fixed shift m = s is valid because |cos| <= 1, so partial sums from different tiles and GPUs add. */
size_t mmg_infonce_workspace_bytes(int prec, int rows, int cols, int D);
int mmg_infonce_fwd(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                    const float* scale, float* rowsum, float* colsum, float* diag, void* workspace,
                    size_t workspace_bytes, mmg_stream_t stream);

/* loss_out[0] = inv_two_b * sum_{i<n} ( log rowsum[i] + log colsum[i] + 2*s - 2*diag[i] )
 * = (CE_rows + CE_cols)/2 of losses.py:40-43 when n == B and inv_two_b == 1/(2B).  Deterministic reduction. */
int mmg_infonce_loss(const float* rowsum, const float* colsum, const float* diag, int n, const float* scale,
                     float inv_two_b, float* loss_out, mmg_stream_t stream);

/* The same loss in two parts for the row-sharded case (one cross-rank sum then serves the column sums AND the loss):
 *     part_out[0] = sum_{r < rows} ( log rowsum[r] - 2*diag[r] )            -- local rows, before the exchange
 *     loss_out[0] = inv_two_b * ( row_part[0] + sum_{c < cols} log colsum[c] + 2*cols*s )
 * with row_part = the sum of every rank's part and colsum the global column sums; equals mmg_infonce_loss over the whole
 * batch (losses.py:40-43).  Deterministic reductions; a vanished / overflowed sum gives NaN. */
int mmg_infonce_row_part(const float* rowsum, const float* diag, int rows, float* part_out, mmg_stream_t stream);
int mmg_infonce_loss_cols(const float* colsum, int cols, const float* scale, const float* row_part, float inv_two_b,
                          float* loss_out, mmg_stream_t stream);

/* rinv[r] = s*gl*inv_two_b / rowsum[r], cinv[c] = s*gl*inv_two_b / colsum[c]; scal (4 floats): [1] = dcoef =
 * 2*s*gl*inv_two_b; [0] = the diagonal coefficient mmg_infonce_bwd itself subtracts (dcoef, or 0 when diag_in_fp32);
 * [2] = diag_in_fp32 flag: mmg_infonce_bwd then ZEROES the matching-pair element of g and the caller applies it with
 * mmg_infonce_bwd_diag (what the bf16 path does).  grad_loss is a DEVICE scalar (1 for a bare loss.backward()).
 * prec = the precision of the mmg_infonce_bwd* call that follows.  With MMG_PREC_F16 the coefficients are stored as fp16 in
 * scaled units: rinv = 2^14 / rowsum, cinv = 2^14 / colsum (so g' = E*(rinv + cinv) lies in [0, 2^15]), scal[0] in the
 * same units and scal[3] = s*gl*inv_two_b * 2^-14, the factor the gradient epilogues multiply back in (scal[3] = 1
 * otherwise). */
int mmg_infonce_bwd_prep(int prec, const float* rowsum, int rows, const float* colsum, int cols, const float* scale,
                         const float* grad_loss, float inv_two_b, int diag_in_fp32, float* rinv, float* cinv,
                         float* scal, mmg_stream_t stream);

/* Matching-pair element of the gradient applied from the fp32 embeddings (keeps the bf16 operand rounding out of the
 * dominant, heavily cancelling term).  b32, dB and cinv_paired point at the `rows` column-side entries paired with the
 * local rows (column diag_offset + r); diag is the forward's diag[] (the pair's logit):
 *   g = exp(diag[r] - s)*(rinv[r] + cinv_paired[r]) - dcoef
 *   dA[r,:] += g*b32[r,:],   dB[r,:] += g*a32[r,:],   dlogscale_acc += g * diag[r]/s   (dlogscale_acc may be NULL).
 * init != 0: the rows are WRITTEN (dA[r,:] = g*b32[r,:], dB[r,:] = g*a32[r,:]) instead of accumulated -- call it before
 * mmg_infonce_bwd on uninitialised dA / dB rows and skip their zero-fill (column rows of dB that are not paired with a
 * local row still have to be zeroed by the caller). */
int mmg_infonce_bwd_diag(const float* a32, const float* b32, int rows, int D, const float* diag, const float* scale,
                         const float* rinv, const float* cinv_paired, const float* scal, float* dA, float* dB,
                         float* dlogscale_acc, int init, mmg_stream_t stream);

/* mmg_infonce_bwd_prep(diag_in_fp32 = 1) and mmg_infonce_bwd_diag(init = 1) as ONE launch (what the bf16 path runs before
 * its contraction kernel): writes rinv[rows], cinv[cols], scal[4] and the matching-pair rows dA[r,:] = g*b32[r,:],
 * dB_matching[r,:] = g*a32[r,:] (g as above, formed from rowsum[r] and colsum[diag_offset + r] directly).  b32 and
 * dB_matching point at the `rows` column-side rows paired with the local rows. */
int mmg_infonce_bwd_prep_diag(int prec, const float* rowsum, int rows, const float* colsum, int cols, int diag_offset,
                              const float* scale, const float* grad_loss, float inv_two_b, float* rinv, float* cinv,
                              float* scal, const float* a32, const float* b32, int D, const float* diag, float* dA,
                              float* dB_matching, float* dlogscale_acc, mmg_stream_t stream);

/* dA[rows,D] += g . b_hat,  dB[cols,D] += g^T . a_hat,  dlogscale_acc[0] += sum g*cos   with
 *     g = exp(s*cos - s) * (rinv[r] + cinv[c]) - scal[0]*[c == r + diag_offset]    ( = s * dloss/dlogit; the
 *         matching-pair element is zeroed instead when scal[2] != 0, see mmg_infonce_bwd_diag )
 * dA, dB, dlogscale_acc are fp32 and must be zero-filled (or hold a running sum) by the caller; dlogscale_acc may be
 * NULL when d loss / d logit_scale is not wanted (the reference's logit_scale is not a Parameter on CUDA).
 * Works block by block (block_rows x block_cols, 0 = library default) through `workspace`. */
int mmg_infonce_bwd(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                    const float* scale, const float* rinv, const float* cinv, const float* scal, float* dA, float* dB,
                    float* dlogscale_acc, int block_rows, int block_cols, void* workspace, size_t workspace_bytes,
                    mmg_stream_t stream);

/* The same backward with the column-side gradient DISTRIBUTED over `n_owners` buffers (<= 8): column c accumulates into
 * dB_owners[c / (cols/n_owners)] at row c % (cols/n_owners).  dB_owners is a HOST array of device pointers, each an fp32
 * [cols/n_owners, D] buffer -- in the row-sharded multi-GPU run the ranks' own gradient buffers mapped into this process
 * (NVLink peer memory, e.g. torch symmetric memory), so the gradient GEMM's TMA reduce-add IS the reduce-scatter: no
 * [cols, D] staging buffer and no separate collective.  The caller initialises the owners' buffers (zero, or the
 * matching-pair term via mmg_infonce_bwd_diag(init=1)) and synchronises the ranks before and after the call.
 * The owners may also be slices of a local staging buffer that a reduce-scatter then sends home; with n_parts > 1 one
 * call covers only part `part` (0-based) of EVERY owner's columns -- dB_owners[i] is then the [cols/n_owners/n_parts, D]
 * buffer of that part, dA accumulates across the calls -- so the reduce-scatter of one part travels while the next part
 * is computed.  prec = MMG_PREC_BF16 or MMG_PREC_F16 (type of a_hat / b_hat); runs as the one fused persistent launch or
 * returns MMG_ERR_UNSUPPORTED_SHAPE. */
int mmg_infonce_bwd_owners(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                           const float* scale, const float* rinv, const float* cinv, const float* scal, float* dA,
                           float* const* dB_owners, int n_owners, int n_parts, int part, float* dlogscale_acc,
                           void* workspace, size_t workspace_bytes, mmg_stream_t stream);

/* ---- stored-E variant of the fused InfoNCE (bf16 path): trade 2*rows*cols bytes of HBM for a third of the backward ----
 * The forward additionally keeps E = exp(logit - s) as bf16 [rows, lde] (same reference lines as mmg_infonce_fwd:
 * losses.py:28-44 via mmgclip_model.py:135-136); the backward then turns E into the gradient coefficients with a streaming
 * transform instead of recomputing the cosines on the tensor cores (4 instead of 6 rows*cols*D FLOPs).  No
 * d/d logit_scale in this mode (use mmg_infonce_bwd when logit_scale is trained).  mmg_infonce_stored_supported tells
 * whether mmg_infonce_bwd_stored covers a shape (rows, cols/owners/parts and D multiples of 256).  prec = MMG_PREC_BF16 or
 * MMG_PREC_F16 for the forward (type of a_hat / b_hat; E itself is always bf16); the stored-E backward takes
 * MMG_PREC_BF16 operands only. */
int mmg_infonce_stored_supported(int rows, int cols, int D, int n_owners, int n_parts);
int mmg_infonce_fwd_store(int prec, const void* a_hat, const void* b_hat, int rows, int cols, int D, int diag_offset,
                          const float* scale, float* rowsum, float* colsum, float* diag, void* e_out, long long lde,
                          mmg_stream_t stream);
int mmg_infonce_bwd_stored(int prec, const void* a_hat, const void* b_hat, const void* e_stored, long long lde, int rows, int cols,
                           int D, int diag_offset, const float* scale, const float* rinv, const float* cinv,
                           const float* scal, float* dA, float* const* dB_owners, int n_owners, int n_parts, int part,
                           void* workspace, size_t workspace_bytes, mmg_stream_t stream);

/* ---- SURVEY s8(f) "next" rows: the callers either side of the path ---- */

/* 'eos' text pooling, mmgclip/networks/mmgclip_model.py:108-111: idx[r] = attention_mask[r,:].sum() - 1 (negative wraps
 * to seq-1 like Python indexing); out[r,:] = hidden[r, idx[r], :].  hidden [n, seq, H] fp32, attention_mask [n, seq]
 * int64, out [n, H] fp32, idx_out [n] int64 or NULL (needed by the backward). */
int mmg_eos_pool(const float* hidden, const long long* attention_mask, int n, int seq, int H, float* out,
                 long long* idx_out, mmg_stream_t stream);
/* dhidden [n, seq, H] = 0 except dhidden[r, idx[r], :] = dout[r, :]. */
int mmg_eos_pool_bwd(const float* dout, const long long* idx, int n, int seq, int H, float* dhidden,
                     mmg_stream_t stream);

/* Multi-tensor AdamW step, replacing `self.optimizer.step()` of mmgclip/experiments/ClassifierExperiment.py:74,118
 * (torch.optim.AdamW(model.parameters(), lr, weight_decay); betas (0.9, 0.999), eps 1e-8 by default):
 *   p *= 1 - lr*wd;  m += (1-b1)(g-m);  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
 * params/grads/exp_avg/exp_avg_sq/numel are HOST arrays of n_tensors device pointers / element counts (fp32 tensors).
 * step_state: device int64[2] = {t (steps taken so far), 0}; the launch reads t+1 and stores it back, so a captured
 * CUDA graph advances it on every replay.  lr_dev: device fp32 scalar read at run time (NULL: use `lr`). */
int mmg_adamw_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const long long* numel, int n_tensors, float lr, const float* lr_dev, float beta1, float beta2,
                   float eps, float weight_decay, long long* step_state, mmg_stream_t stream);

/* Plan knobs of the fused InfoNCE backward, process-wide: key in {"fused" (0 = block loop), "fused_rb", "fused_cb" (block
 * shape), "fused_nbuf" (scratch buffers), "fused_ksl", "fused_ksl_t" (64-wide K blocks per dA / dB gradient slice),
 * "fused_sr", "fused_sc" (block-order super-tile)}; value < 0 unsets the knob, key "reset" unsets all.  Every setting gives
 * the same results (the tests walk several schedules with it); the defaults are the measured best.  The product library
 * reads no environment variable -- the measurement build (make measure, -DMMG_MEASURE) maps MMG_* variables onto these. */
int mmg_tune(const char* key, int value);

/* Debug timeline of the fused backward (measurement builds only: -DMMG_MEASURE and MMG_FUSED_TRACE=1 at launch time; the
 * product library always returns 0 here).  mmg_infonce_workspace_bytes then
 * includes the region): per CTA and role (0 TMA producer, 1 MMA issuer, 2 epilogue warp 0, 3 transform warp 0)
 * `records_per_role` records {globaltimer ns, tag} of 16 bytes at workspace + *offset; tag = role<<60 | event<<56 |
 * type<<52 | block<<32 | tm<<16 | tn (events: 0 item picked up, 1 dependency wait over, 2 item finished).  Returns 1 when
 * tracing is enabled and the shape is covered by the fused launch, else 0. */
int mmg_debug_fused_trace_region(int rows, int cols, int D, size_t* offset, size_t* bytes, int* records_per_role,
                                 int* roles);

/* Introspection (host only, no GPU): the static work-item schedule of the fused backward for CTA pair `pair` of `pairs`
 * -- rows of items[] are 8 ints {type (0 coefficient tile, 1 dA slice, 2 dB slice), block, tm, tn, kb0, nkb, global
 * column block, row block}; info[8] = {Rb, Cb, nbuf, nA, nB, nblk, kslI, kslT}.  Returns the pair's item count (0 = shape not covered).
 * tests/test_fused_schedule_cpu.py uses it to check the schedule's ordering / dead-lock-freedom invariants. */
int mmg_fused_bwd_schedule(int rows, int cols, int D, int n_owners, int n_parts, int part, int pairs, int pair, int* items,
                           int max_items, int* info);

/* ---- literal cross-entropy on materialised logits (losses.py:28-44, 207-212) ----------------------------- */
/* F.cross_entropy(logits[n, m], labels) pieces; labels: int64 DEVICE array [n] or NULL = arange(n) (needs n <= m).
 * lse[n] is kept for the backward.   loss_out[0] += coef * sum_r (lse[r] - logits[r, labels[r]]). */
int mmg_ce_fwd(const float* logits, long long ld, int n, int m, const long long* labels, float coef, float* lse,
               float* loss_out, mmg_stream_t stream);
/* dlogits[r,c] = coef * gl * (exp(logits[r,c] - lse[r]) - [c == labels[r]]) */
int mmg_ce_bwd(const float* logits, long long ld, int n, int m, const long long* labels, const float* lse,
               const float* grad_loss, float coef, float* dlogits, long long ldd, mmg_stream_t stream);
/* out[0] = sum_i x[i]*y[i] (deterministic); d logit_scale of materialised logits s*(A.B^T). */
int mmg_dot_sum(const float* x, const float* y, long long n, float* out, mmg_stream_t stream);

/* ---- zero-shot prompt scoring (mmgclip_model.py:201-209; evaluator.py:182-188,282-299,354-368) ------------- */
/* logits[n,c] = (s*img[n,:]) . txt[c,:] in fp32 (scale-then-multiply, as the reference's operator precedence does),
 * probs = softmax over c, argmax with ties -> lowest index, top-k ordered (value desc, index asc).
 * img: [N, D], txt: [C, D] fp32 row-major, k <= 8.  Any output pointer may be NULL when C <= 64; with more prompts a
 * one-warp-per-row kernel runs instead of the tiled one and logits_out is required (it parks the row's logits there). */
int mmg_zeroshot_score(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                       float* probs_out, long long* argmax_out, int k, long long* topk_idx_out, float* topk_val_out,
                       mmg_stream_t stream);

/* Same contract on the tensor pipe: both operands are split into two TF32 terms inside the kernel (a ~ a_hi + a_lo) and
 * a_hi.b_hi + a_lo.b_hi + a_hi.b_lo is accumulated in fp32 (tcgen05.mma kind::tf32), which keeps the logits within
 * ~1e-6 (max-abs / max-abs) of the fp32 FFMA evaluation above while the embeddings stream from HBM exactly once.
 * Needs D % 4 == 0, 16-byte aligned img, and a 256-byte aligned workspace of mmg_zeroshot_workspace_bytes(C, D) bytes
 * (the prompts' TF32 terms); returns MMG_ERR_UNSUPPORTED_SHAPE otherwise. */
size_t mmg_zeroshot_workspace_bytes(int C, int D);
int mmg_zeroshot_score_tc(const float* img, const float* txt, int N, int C, int D, const float* scale, float* logits_out,
                          float* probs_out, long long* argmax_out, int k, long long* topk_idx_out, float* topk_val_out,
                          void* workspace, size_t workspace_bytes, mmg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMGCLIP_B200_H_ */
