/* maest_b200.h — C ABI of libmaest_b200.so: the B200 (sm_100a) hot path of palonso/MAEST.
 *
 * The reference (pure Python) has no FFI layer; its boundary for this path is the Python API
 * (get_maest / MAEST.forward / predict_labels / Module.training_step).  This header is the C boundary we
 * introduce UNDER that API: every entry point names the reference code it replaces (file:line relative to
 * the reference repo root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (e.g. torch tensors); nothing is allocated,
 *     freed or retained; kernels are enqueued on `stream` (a cudaStream_t passed as void*); no host sync.
 *   - return 0 on success, negative on error; maest_last_error() returns a thread-local message.
 *   - "op16" tensors are 16-bit GEMM operands: dtype MAEST_F16 or MAEST_BF16 (same tensor-core rate).
 *   - row-major everywhere; [M, K] means M rows of K contiguous elements.
 */
#ifndef MAEST_B200_H_
#define MAEST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MAEST_F16 = 0, MAEST_BF16 = 1, MAEST_F32 = 2 };

/* GEMM epilogues (maest_linear_fwd) */
enum {
  MAEST_EPI_STORE16 = 0, /* out16 = A W^T + bias                                   (qkv: models/maest.py:361)            */
  MAEST_EPI_GELU16 = 1,  /* out16 = gelu_erf(A W^T + bias)                         (fc1+GELU: models/maest.py:203-204)   */
  MAEST_EPI_RESID32 = 2, /* out32 = resid32 + A W^T + bias                         (proj/fc2 + residual: :376,:206,:418-419) */
  MAEST_EPI_STORE32 = 3, /* out32 = A W^T + bias (+ addend table, + row remap)     (patch-embed + pos-embed: :250,:670-675) */
  MAEST_EPI_GELUBWD16 = 4, /* out16 = (A B^T) * gelu'(aux16)    (autograd of mlp.act + mlp.fc2, models/maest.py:204-206)           */
  MAEST_EPI_ATOMIC32 = 5,  /* out32 += A B^T (split-K, atomics)  (weight gradients; autograd of every nn.Linear / the conv)        */
  /* LayerNorm folded into the GEMMs around it (maest_linear_ln_fwd; Block.forward, models/maest.py:418-419 with :395,:405):
   *   LN(x) W^T + b  =  rstd * ((x*gamma) W^T)  -  rstd * mean * (W gamma)  +  (W beta + b)                                        */
  MAEST_EPI_STORE16_LN = 7, /* consumer (qkv): A = x*gamma op16; out16 = rstd*acc - rstd*mean*ln_vec[n] + bias[n]                   */
  MAEST_EPI_GELU16_LN = 8,  /* consumer (fc1): gelu_erf of the same                                                                 */
  MAEST_EPI_RESID32_LN = 9  /* producer (proj, fc2): RESID32, plus out16b = x*ln_vec[n] (op16) and per-chunk row statistics      */
};

/* pooling modes (maest_pool_head_fwd) */
enum { MAEST_HEAD_MEAN = 0, MAEST_HEAD_SEPARATED = 1 };

/* Per-block weights for maest_encoder_fwd (state-dict names: blocks.<i>.*; models/maest.py:381-420). */
typedef struct MaestBlockWeights {
  const float* ln1_w; const float* ln1_b;      /* norm1.{weight,bias}        [768]                 */
  const void* qkv_w;  const float* qkv_b;      /* attn.qkv.weight op16 [2304,768], bias fp32 [2304] */
  const void* proj_w; const float* proj_b;     /* attn.proj.weight op16 [768,768], bias [768]      */
  const float* ln2_w; const float* ln2_b;      /* norm2.*                                         */
  const void* fc1_w;  const float* fc1_b;      /* mlp.fc1.weight op16 [3072,768], bias [3072]      */
  const void* fc2_w;  const float* fc2_b;      /* mlp.fc2.weight op16 [768,3072], bias [768]       */
  /* optional (all four or none; NULL = every LayerNorm runs as its own kernel): vectors made by maest_ln_fold */
  const float* qkv_wg; const float* qkv_bf;    /* qkv_w  . norm1.weight [2304];  qkv_b + qkv_w . norm1.bias [2304] */
  const float* fc1_wg; const float* fc1_bf;    /* fc1_w  . norm2.weight [3072];  fc1_b + fc1_w . norm2.bias [3072] */
} MaestBlockWeights;

const char* maest_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int32_t maest_abi_version(void);
/* Select device, resolve the driver's tensor-map encoder, opt kernels into >48 KB dynamic smem, upload K1 tables.
 * Must be called once per process per device before any other call. */
int32_t maest_init(int32_t device);

/* K1 — waveform -> normalised log-mel.  Replaces MelSpectrogram.forward, models/helpers/melspectrogram.py:47-60
 * (torchaudio Spectrogram + MelScale + log10(1+1e4 x) + z-norm) and helpers/melspectrogram_extractor.py:15-48.
 * wav fp32 [B, S] with row stride wav_stride (elements); mel fp32 [B, 96, T], T = 1 + S/256.  Requires S > 256. */
int32_t maest_logmel_fwd(const float* wav, int32_t B, int32_t S, int64_t wav_stride, float* mel, void* stream);

/* K1, dataset-file flavour.  Same STFT + mel + log compression, but the output is what the
 * reference's offline extractor stores: UN-normalised log10(1 + 1e4 mel) as float16, time-major [B, T, 96]
 * (helpers/melspectrogram_extractor.py:15-48 writes [frames, 96] float16 .mmap files; the z-norm is applied later by the
 * data module).  framing 0: the torchaudio framing of K1 (reflect padding, periodic Hann, T = 1 + S/256).  framing 1: the
 * framing of the reference's extractor, restated from Essentia's published algorithms (essentia is a third-party dependency,
 * unpinned in pyproject.toml and absent from the reference tree): FrameCutter(frameSize 512, hopSize 256, startFromZero=false)
 * = frames centred on sample 256 t with ZERO padding, T = ceil(S/256); Windowing(type='hann', normalized=false) = symmetric
 * Hann; Spectrum + MelBands(slaneyMel, unit_tri, power) + log10(1 + 1e4 x) as in framing 0. */
int32_t maest_logmel_raw16_fwd(const float* wav, int32_t B, int32_t S, int64_t wav_stride, void* raw_tm16, int32_t framing, void* stream);

/* Fused AdamW (+ SWA running average) over all parameters in one launch.  Replaces
 * torch.optim.AdamW as built by Module.get_optimizer (models/module.py:237-243) and the running average kept by the SWA
 * callback (helpers/swa_callback.py:11-15 -> torch.optim.swa_utils avg_fn).
 *   tensor_table: device array of { float* p; const float* g; float* m; float* v; float* swa (or NULL); int64 n; }
 *   chunk_table:  device array of { int32 tensor; int32 pad; int64 start; }, one entry per 8192 elements of every tensor
 *   step: 1-based step count (bias corrections);  grad_scale: multiplies the gradients first (1/loss_scale, or 1)
 *   swa_inv: 1/(n_averaged+1) to also update the running average of every tensor with swa != NULL, or 0 */
int32_t maest_adamw_step(const void* tensor_table, const void* chunk_table, int32_t n_chunks, float lr, float beta1, float beta2,
                         float eps, float weight_decay, int32_t step, float grad_scale, float swa_inv, void* stream);

/* Epoch-end weight averaging without a second model copy: for every tensor of the table (same layout as above; only p, swa and
 * n are read) swa += (p - swa) * swa_inv.  Replaces AveragedModel.update_parameters as driven by Lightning's
 * StochasticWeightAveraging / helpers/swa_callback.py:11-17 (the averaged model IS net_swa here). */
int32_t maest_swa_fold(const void* tensor_table, const void* chunk_table, int32_t n_chunks, float swa_inv, void* stream);

/* Validation metrics on the device.  Replaces sklearn's average_precision_score / roc_auc_score
 * (average=None) as called by Module.on_test_validation_epoch_end, models/module.py:189-190, on host copies of the gathered
 * predictions.  score_sorted / label_sorted: fp32 [n, C], every class column ordered by descending score (labels re-ordered
 * with the same permutation).  ap, auc: fp64 [C]; n_pos: int32 [C] positives per class.  A class without positives or
 * without negatives gets auc = NaN and ap = 0 / 1 (what scikit-learn >= 1.6 returns, with a warning). */
int32_t maest_ap_roc_fwd(const float* score_sorted, const float* label_sorted, int32_t n, int32_t C, double* ap, double* auc,
                         int32_t* n_pos, void* stream);

/* Loader -> device ingest.  Replaces, per batch, DiscogsDataset.load_melspectrogram's
 * zero-pad + centring np.roll + transpose (discogs/dataset.py:120-139), DiscogsDataModule's norm_func
 * ((x - mean) / (2 std) in float16 arithmetic, discogs/datamodule.py:126-137) and roll_func (torch.roll along time, :111-123).
 *   raw_tm16    fp16 [B, T, 96]: the window bytes of each clip exactly as read from its .mmap file (time-major); rows
 *               >= frames_read[b] are never read
 *   frames_read int32 [B] (device) frames read from the file, <= T; NULL = T everywhere
 *   roll_shift  int32 [B] (device) time shift per clip, NULL = none
 *   out         fp16 [B, 1, 96, T]
 * Bit-exact with the reference (same float16 rounding points). */
int32_t maest_mel_ingest_fwd(const void* raw_tm16, const int32_t* frames_read, const int32_t* roll_shift, int32_t B, int32_t T,
                             int32_t do_norm, float norm_mean, float norm_std, void* out, void* stream);

/* K2 — mel -> packed token buffer.  Replaces PatchEmbed.forward (models/maest.py:243-256) and the pre-block part
 * of MAEST.forward_features (:645-800): + time/freq pos-embed, structured/unstructured patchout, flatten,
 * CLS/DIST rows.
 *   mel        [B, 96, T] of mel_dtype (MAEST_F32 or MAEST_F16)
 *   w_pe       patch_embed.proj.weight viewed [768, 256] as op16; conv_bias fp32 [768]
 *   freq_pe    freq_new_pos_embed viewed [768, Fp];  time_pe  time_new_pos_embed viewed [768, Wt]
 *   cls_token, dist_token [768]; new_pos_embed [2, 768]
 *   keep_ft    device int32 [P]: kept grid cells (f << 16 | t) in sequence order, or NULL = all Fp*Tp cells
 *   t_offset   column offset into time_pe (0 in eval; random in training, :647-657)
 *   tokens     fp32 [B, 2 + P, 768] (out)
 *   workspace  >= maest_patch_workspace_bytes(B, P) bytes
 * Returns -2 if Tp + t_offset exceeds Wt (the reference raises, :664-668). */
size_t maest_patch_workspace_bytes(int32_t B, int32_t P);
int32_t maest_patch_tokens_fwd(const void* mel, int32_t mel_dtype, int32_t B, int32_t T, const void* w_pe,
                               int32_t op_dtype, const float* conv_bias, const float* freq_pe, int32_t Fp,
                               const float* time_pe, int32_t Wt, const float* cls_token, const float* dist_token,
                               const float* new_pos_embed, const int32_t* keep_ft, int32_t P, int32_t t_offset,
                               float* tokens, void* workspace, size_t workspace_bytes, void* stream);

/* K1 + K2 in one call: waveform in, packed token buffer out (SURVEY.md section 8(b) `maest_wave_tokens_fwd`).  Replaces
 * MelSpectrogram.forward + PatchEmbed.forward + the pre-block part of forward_features for the waveform path of
 * MAEST.forward (models/maest.py:862-903 -> :634-800).  Two kernel sequences on `stream` (log-mel, then gather / pos table /
 * patch GEMM); the fp32 mel [B, 96, T] lives in the first 4*B*96*T bytes (rounded up to 256) of `workspace`, followed by the
 * K2 workspace -- it never goes back to the caller.  T = 1 + S/256 (the batched waveform path, which the reference does not
 * trim, :890-892; the 1-D chunking path that trims 626 -> 625 frames keeps the two separate calls).
 * workspace_bytes >= maest_wave_tokens_workspace_bytes(B, S, P). */
size_t maest_wave_tokens_workspace_bytes(int32_t B, int32_t S, int32_t P);
int32_t maest_wave_tokens_fwd(const float* wav, int32_t B, int32_t S, int64_t wav_stride, const void* w_pe,
                              int32_t op_dtype, const float* conv_bias, const float* freq_pe, int32_t Fp, const float* time_pe,
                              int32_t Wt, const float* cls_token, const float* dist_token, const float* new_pos_embed,
                              const int32_t* keep_ft, int32_t P, int32_t t_offset, float* tokens, void* workspace,
                              size_t workspace_bytes, void* stream);

/* LayerNorm over 768-wide rows, fp32 in -> op16 out (GEMM operand).  Replaces norm1/norm2, models/maest.py:395,405,418-419.
 * mean/rstd (fp32 [rows]) are optional saves for the backward pass (may be NULL). */
int32_t maest_layernorm_fwd(const float* x, const float* w, const float* b, void* y16, int32_t op_dtype, int32_t rows,
                            float eps, float* mean, float* rstd, void* stream);

/* out = epilogue(A[M,K] W[N,K]^T): tcgen05 GEMM, A and W op16 (K contiguous, lda / ldw elements).  Replaces the nn.Linear /
 * Conv2d-as-GEMM calls listed at the MAEST_EPI_* enum.  Output row of GEMM row m:
 * (m / rows_per_group) * group_stride + row_offset + m % rows_per_group  (rows_per_group = 0 -> identity).
 * addend: optional fp32 [rows_per_group, N] table added in MAEST_EPI_STORE32; for MAEST_EPI_GELUBWD16 (maest_gemm) an optional fp32
 * [N] vector the column sums of the OUTPUT are accumulated into (the fc1 bias gradient: saves a maest_colsum pass).
 * MAEST_EPI_RESID32 with out == resid (x += A W^T + b in place, blocks.N.attn.proj / mlp.fc2 of the inference encoder): the
 * residual is not read by the SMs at all, acc + bias is added into out by the TMA engine (cp.reduce.async.bulk.tensor .add.f32,
 * one add per element: deterministic); with out != resid the epilogue loads, adds and stores.  Both forms give
 * resid + (acc + bias) up to the order of the two fp32 additions.  Environment MAEST_RESID_REDUCE=0 forces the second form.
 * N % 32 == 0, K % 8 == 0. */
int32_t maest_linear_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias, int32_t M,
                         int32_t N, int32_t K, int32_t op_dtype, int32_t epilogue, void* out, int64_t ld_out,
                         const float* resid, const float* addend, int32_t rows_per_group, int32_t group_stride,
                         int32_t row_offset, void* stream);

/* General tcgen05 GEMM used by the training step: out = epilogue(sum_k A(m,k) B(n,k)).
 *   a_mn = 0: A is [M, K] with K contiguous;  a_mn = 1: A is stored [K, M] with M contiguous (e.g. dY for a weight gradient)
 *   b_mn = 0: B is [N, K] with K contiguous;  b_mn = 1: B is stored [K, N] with N contiguous (e.g. W for an input gradient)
 * so forward (0,0), input-gradient (0,1) and weight-gradient (1,1) GEMMs all read activations, gradients and weights in their
 * natural layouts.  aux16: MAEST_EPI_GELU16 optional 2nd output (pre-activation); MAEST_EPI_GELUBWD16 input (an input gradient
 * has no bias term: bias must be NULL there, error -1 otherwise).
 * k_splits > 1 splits the reduction across CTAs (MAEST_EPI_ATOMIC32 only). */
int32_t maest_gemm(const void* a, int64_t lda, int32_t a_mn, const void* b, int64_t ldb, int32_t b_mn, const float* bias,
                   int32_t M, int32_t N, int32_t K, int32_t op_dtype, int32_t epilogue, void* out, int64_t ld_out,
                   const float* resid, const float* addend, int32_t rows_per_group, int32_t group_stride,
                   int32_t row_offset, void* aux16, int32_t k_splits, void* stream);

/* Tile shape of the forward (K-major) GEMMs: 0 = one CTA per 128x256 tile (cta_group::1), 1 = a CTA pair per 256x256 tile
 * (cta_group::2, W tile shared by the two SMs of a TPC), 2 = per-shape choice (default).  Process-wide tuning switch;
 * results are identical. */
int32_t maest_set_gemm_mode(int32_t pair_mode);

/* Host-side TMA descriptor cache (api.cu make_tmap): lookups served from the calling thread's cache / descriptors encoded
 * through cuTensorMapEncodeTiled since the library was loaded.  Diagnostic only; either pointer may be NULL. */
int32_t maest_tmap_cache_stats(uint64_t* hits, uint64_t* misses);

/* LayerNorm folding, weight side (once per weight version): wg[n] = sum_k gamma[k] W[n,k], bf[n] = bias[n] + sum_k beta[k] W[n,k]
 * from the 16-bit operand copy w16 [N, K] that the GEMM multiplies.  Replaces nothing by itself: it moves norm1 / norm2
 * (models/maest.py:395,405) into the qkv / fc1 GEMM epilogues. */
int32_t maest_ln_fold(const void* w16, const float* gamma, const float* beta, const float* bias, int32_t N, int32_t K,
                      int32_t op_dtype, float* wg, float* bf, void* stream);

/* Linear layers with the neighbouring LayerNorm folded in (epilogue = MAEST_EPI_*_LN).
 *   producer (RESID32_LN): out32 = resid + A W^T + bias;  out16b = out32 * ln_vec (gamma of the next LayerNorm);
 *                          ln_stats = partials fp32 [N/128, M, 4]: (pivot, sum (x - pivot), sum (x - pivot)^2, feature count) of every
 *                          32-column chunk of the new rows -- plain stores, no atomics, bit-reproducible
 *   maest_ln_finalize:     partials -> stats fp32 [M, 2] = (rstd, -mean * rstd)
 *   consumer (STORE16_LN / GELU16_LN): A = the producer's out16b; ln_stats = stats; ln_vec = wg, bias = bf from maest_ln_fold */
int32_t maest_linear_ln_fwd(const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias, int32_t M, int32_t N,
                            int32_t K, int32_t op_dtype, int32_t epilogue, void* out, int64_t ld_out, const float* resid,
                            float* ln_stats, const float* ln_vec, void* out16b, void* stream);
int32_t maest_ln_finalize(const float* partials, int32_t rows, int32_t n_features, float eps, float* stats, void* stream);

/* Fused multi-head attention, d_head 64.  Replaces Attention.forward lines models/maest.py:362-375.
 * qkv op16 [B*N, 3*H*64] as written by the qkv linear (columns = [q|k|v][head][64]); out op16 [B*N, H*64].
 * variant: 0 = default (P in TMEM, exponentials taken against the running max of the previous KV tiles with a verified
 *   redo when a tile raises the max by more than 2^8, last KV tile narrowed to the real keys); 1 = max-first, P staged through
 *   shared memory; 2 = max-first, P in TMEM.  All variants compute the same function (tests hold them to the same tolerance).
 * lse: optional fp32 [B, H, N] (may be NULL): per-row max + log2(sum) in the scaled log2 domain, saved for the backward pass. */
int32_t maest_attention_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t N, int32_t H, int32_t op_dtype,
                            int32_t variant, void* stream);

/* The 12-block encoder on the fp32 residual stream x [B*N, 768], in place.  Replaces Block.forward x n_blocks
 * (models/maest.py:414-420, driven from :804-820).  If last_attn_only != 0 the last block writes attn(norm1(x))
 * WITHOUT residual into x (return_self_attention, :415-416).  workspace >= maest_encoder_workspace_bytes(B*N). */
size_t maest_encoder_workspace_bytes(int64_t rows);
int32_t maest_encoder_fwd(float* x, int32_t B, int32_t N, const MaestBlockWeights* blocks, int32_t n_blocks,
                          int32_t last_attn_only, int32_t op_dtype, int32_t attn_variant, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Pooling + head.  Replaces models/maest.py:806-810 (final LN, only rows 0/1 are normalised) and :905-925.
 * x fp32 [B, N, 768]; logits [B, C]; logits_dist [B, C] (MAEST_HEAD_SEPARATED only, else NULL); feats [B, 768].
 * ln_cls / ln_dist: optional [B,768] saves of the normalised cls / dist rows (NULL in inference). */
int32_t maest_pool_head_fwd(const float* x, int32_t B, int32_t N, const float* norm_w, const float* norm_b,
                            const float* head_ln_w, const float* head_ln_b, const float* head_w, const float* head_b,
                            const float* head_dist_w, const float* head_dist_b, int32_t C, int32_t mode,
                            float* logits, float* logits_dist, float* feats, float* ln_cls, float* ln_dist,
                            void* stream);

/* Block-k embedding: emb[b] = cat(x[b,0], x[b,1], mean(x[b,2:], 0)), fp32 [B, 2304].  Replaces models/maest.py:825-829. */
int32_t maest_block_embedding_fwd(const float* x, int32_t B, int32_t N, float* emb, void* stream);

/* fp32 -> op16 cast (weight staging). */
int32_t maest_cast_to16(const float* src, void* dst, int64_t n, int32_t op_dtype, void* stream);

/* ---- training step: Module.training_step (models/module.py:73-102) and the autograd mirrors of the forward path ---- */

/* Attention backward (autograd of models/maest.py:362-375).  qkv / o / d_o op16 as in the forward; lse from the forward.
 * delta: fp32 [B,H,N] scratch; dq32: fp32 [B*N, H*64] scratch (atomically accumulated dQ); dqkv: op16 [B*N, 3*H*64] out. */
int32_t maest_attention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* delta, float* dq32,
                            void* dqkv, int32_t B, int32_t N, int32_t H, int32_t op_dtype, void* stream);

/* mixup blend of a batch with a permutation of itself (models/module.py:77-86, helpers/mixup.py:5-12):
 * out[b,:] = x[b,:]*lam[b] + x[perm[b],:]*(1-lam[b]).  x: [B, L] of x_dtype (MAEST_F16 | MAEST_F32); out fp32. */
int32_t maest_mixup_fwd(const void* x, int32_t x_dtype, const int32_t* perm, const float* lam, float* out, int32_t B,
                        int64_t L, void* stream);

/* F.binary_cross_entropy_with_logits(logits, targets), mean over n elements (models/module.py:90).
 * loss: fp32 scalar; dlogits[n] = (sigmoid(z) - y) / n  (the gradient for d(loss) = 1). */
int32_t maest_bce_logits_fwd(const float* logits, const float* targets, int32_t n, float* loss, float* dlogits, void* stream);

/* Backward of pooling + head ("mean" mode): final LayerNorm on rows 0/1, (cls+dist)/2, head LN + Linear
 * (autograd of models/maest.py:806-810, :906-909).  x: saved residual stream [B,N,768]; gscale: device scalar d(loss).
 * Writes rows 0/1 of every clip in dx [B,N,768]; accumulates the six parameter gradients; hz_ws: fp32 [B,768] scratch. */
int32_t maest_head_bwd(const float* x, int32_t B, int32_t N, const float* dlogits, const float* gscale, const float* norm_w,
                       const float* norm_b, const float* head_ln_w, const float* head_ln_b, const float* head_w, int32_t C,
                       float* dx, float* hz_ws, float* d_norm_w, float* d_norm_b, float* d_head_ln_w, float* d_head_ln_b,
                       float* d_head_w, float* d_head_b, void* stream);

/* The same for distilled_type = "separated" (models/maest.py:914-925: logits = head(cls), logits_dist = head_dist(dist); the
 * teacher-student step of models/module.py:279-313 trains both): dlogits / dlogits_dist are the two BCE gradients, z1_ws is a
 * second fp32 [B,768] scratch, head_dist is a bare Linear. */
int32_t maest_head_bwd_separated(const float* x, int32_t B, int32_t N, const float* dlogits, const float* dlogits_dist,
                                 const float* gscale, const float* norm_w, const float* norm_b, const float* head_ln_w,
                                 const float* head_ln_b, const float* head_w, const float* head_dist_w, int32_t C, float* dx,
                                 float* hz_ws, float* z1_ws, float* d_norm_w, float* d_norm_b, float* d_head_ln_w,
                                 float* d_head_ln_b, float* d_head_w, float* d_head_b, float* d_head_dist_w, float* d_head_dist_b,
                                 void* stream);

/* LayerNorm backward (autograd of norm1 / norm2): dx += LN'(dy); dgamma/dbeta accumulated; optional op16 copy of the updated dx;
 * optional dx_colsum[768] += column sums of the UPDATED dx (the bias gradient of the linear layer that produced the normalised
 * tensor's input: models/maest.py:418-419 -- saves a separate maest_colsum pass). */
int32_t maest_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                            float* dx, void* dx16, int32_t op_dtype, float* dgamma, float* dbeta, float* dx_colsum, int32_t rows,
                            void* stream);

/* out[n] += sum_m in[m,n]  (bias gradients); in: [M, N] of in_dtype with row stride ld. */
int32_t maest_colsum(const void* in, int32_t in_dtype, int64_t ld, int32_t M, int32_t N, float* out, void* stream);

/* dst16[m, 0:768] = src32[remap(m), 0:768] with the row remap of maest_linear_fwd (rows_per_group = 0 -> identity); dst row stride dst_ld. */
int32_t maest_cast_rows16(const float* src, void* dst, int64_t dst_ld, int32_t rows, int32_t rows_per_group, int32_t group_stride,
                          int32_t row_offset, int32_t op_dtype, void* stream);

/* Gradients of the token-assembly stage (autograd of models/maest.py:670-675, :785-796) from the gradient stream dx [B,N,768]:
 * cls/dist tokens, new_pos_embed [2,768], conv bias [768], freq_new_pos_embed [768,Fp], time_new_pos_embed [768,Wt]. Accumulates. */
int32_t maest_token_grad(const float* dx, int32_t B, int32_t N, int32_t P, int32_t Tp, int32_t Fp, int32_t Wt, int32_t t_offset,
                         const int32_t* keep_ft, float* d_cls, float* d_dist, float* d_new_pos, float* d_conv_bias,
                         float* d_freq, float* d_time, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAEST_B200_H_ */
