/*
 * manner_b200.h -- C ABI of libmanner_b200.so: the B200 (sm_100a) scoring / ensemble / metrics hot
 * path of MANNeR (andreeaiana/manner) as a drop-in for the reference's evaluation path.
 *
 * The reference is pure Python and has NO native operator interface; its plugin seam
 * is Hydra's `_target_` (configs/model/*.yaml:1).  This header is therefore the interface a
 * maintainer binds with ctypes (INTEGRATION.md shows the stub); each entry point cites the
 * reference lines whose arithmetic it replaces.
 *
 * Conventions
 *  - plain C: pointers and sizes only.  Every buffer is CALLER-OWNED DEVICE memory unless marked
 *    HOST; the library never allocates, frees or retains caller memory.
 *  - every call is asynchronous on `stream` (a cudaStream_t / CUstream passed as void*, NULL = the
 *    legacy default stream) and runs on the device that owns `tables[0]` / `preds`.
 *  - return value: MB200_OK or an MB200_ERR_* code; never throws, never exits.  Data-dependent
 *    problems found by a kernel (bad row id, impression longer than max_cand) are OR-ed into the
 *    caller's device `flags` word (MB200_FLAG_*), to be read after the stream is synchronised.
 *  - re-entrant; one call per (device, stream) at a time.  NCCL stays outside this ABI: the sums
 *    and rank statistics it returns are plain additive integers / fp64 that the host layer
 *    all-reduces (manner_b200/dist.py).
 */
#ifndef MANNER_B200_H
#define MANNER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MB200_API __attribute__((visibility("default")))
#else
#define MB200_API
#endif

#define MB200_ABI_VERSION 4
#define MB200_MAX_MODULES 4 /* CR-Module + up to 3 A-Modules (reference: CR, category, sentiment) */
#define MB200_MAX_K 31      /* largest ranking cut-off k for nDCG / diversity / personalization */
#define MB200_MAX_CLASSES 64
#define MB200_MAX_TABLE_SHARDS 8 /* GPUs of one NVSwitch box a row-sharded embedding table may be spread over */
#define MB200_MAX_UPLOAD_SEGMENTS 32 /* segments of a pipelined behaviour upload (mb200_upload_begin) */

/* ---- status codes ------------------------------------------------------------------------------ */
#define MB200_OK 0
#define MB200_ERR_INVALID_ARG 1 /* NULL / negative / inconsistent argument                      */
#define MB200_ERR_UNSUPPORTED 2 /* shape outside what the kernels cover (dim, max_cand, ...)     */
#define MB200_ERR_CUDA 3        /* a CUDA runtime call failed (mb200_last_cuda_error() has text) */
#define MB200_ERR_WORKSPACE 4   /* workspace NULL or smaller than mb200_*_workspace_bytes()      */

/* ---- bits of the device `flags` word -------------------------------------------------------------- */
#define MB200_FLAG_BAD_ID 1          /* a hist/cand id was outside [0, n_news): the row was read as 0 */
#define MB200_FLAG_CAND_OVERFLOW 2   /* an impression had more than max_cand candidates: skipped      */
#define MB200_FLAG_OUTSIDE_UNIT 4    /* some materialised score was not in [0,1] (AUROC sigmoid rule) */
#define MB200_FLAG_BAD_ASPECT 8      /* an aspect label was outside [0, num_classes)                  */
#define MB200_FLAG_EXCHANGE_TIMEOUT 16 /* mb200_exchange_finish: a peer GPU's stores did not arrive within 4 s     */
#define MB200_FLAG_POS_OVERFLOW 32   /* mb200_exchange_finish: a rank had more positives than pos_capacity: AUROC invalid */
#define MB200_FLAG_UPLOAD_TIMEOUT 64 /* mb200_score_eval with `ready`: a segment of the pipelined upload did not arrive within 4 s */

/* ---- embedding table element types ----------------------------------------------------------------- */
#define MB200_F32 0
#define MB200_BF16 1 /* rows stored as bf16, all arithmetic in fp32 */

/* ---- metric slots: layout of `sums` [n_weightings][MB200_NUM_METRICS] (fp64) and of
 *      `per_impression` [n_weightings][n_impressions][MB200_NUM_METRICS] (fp32) ------------------------ */
#define MB200_M_MRR 0        /* RetrievalMRR: first-hit reciprocal rank      (cr_module.py:82)          */
#define MB200_M_NDCG_K0 1    /* RetrievalNormalizedDCG(k=k0)                 (cr_module.py:83)          */
#define MB200_M_NDCG_K1 2    /* RetrievalNormalizedDCG(k=k1)                 (cr_module.py:84)          */
#define MB200_M_GAUC 3       /* per-impression AUC, 0 where undefined        (new)           */
#define MB200_M_GAUC_VALID 4 /* 1 where 0 < positives < candidates                                      */
#define MB200_M_CATEG_DIV_K0 5  /* Diversity(num_categ_classes, k0)          (ensemble_module.py:56-61) */
#define MB200_M_CATEG_DIV_K1 6
#define MB200_M_SENT_DIV_K0 7   /* Diversity(num_sent_classes, k)            (ensemble_module.py:62-67) */
#define MB200_M_SENT_DIV_K1 8
#define MB200_M_CATEG_PERS_K0 9 /* Personalization(num_categ_classes, k)     (ensemble_module.py:68-73) */
#define MB200_M_CATEG_PERS_K1 10
#define MB200_M_SENT_PERS_K0 11 /* Personalization(num_sent_classes, k)      (ensemble_module.py:74-79) */
#define MB200_M_SENT_PERS_K1 12
#define MB200_M_LOSS 13         /* per-impression loss of `loss_kind` on the scores of `scores_weighting` (cr_module.py:140-171) */
#define MB200_M_LOSS_NONZERO 14 /* 1 where that loss is > 0 (the AvgNonZero reduction of the SupCon loss)                       */
#define MB200_NUM_METRICS 15
#define MB200_PAYLOAD_TAIL 6

/* ---- loss kinds (mb200_eval_desc.loss_kind) ---------------------------------------------------------- */
#define MB200_LOSS_NONE 0
#define MB200_LOSS_CE 1     /* torch CrossEntropyLoss()(scores [B, Cmax], y_true [B, Cmax]) with the 0/1 labels as class
                               probabilities; padded columns (score 0) take part in the log-softmax  (cr_module.py:75-76,171) */
#define MB200_LOSS_SUPCON 2 /* components/losses.py:6-40 on pytorch_metric_learning's SupConLoss, similarity = scores
                               (cr_module.py:144-169): -mean over positives of log softmax(s / T) over the real candidates */

/*
 * One evaluation call = one pass over a set of impressions: what the reference spreads over
 * CRModule.forward (cr_module.py:105-131), model_step's flattening (:173-182), test_step (:253-264)
 * and on_test_epoch_end (:266-274); with zscore != 0 and weights, EnsembleModule.forward /
 * _submodel_forward (ensemble_module.py:95-151) and its epoch end (:214-238).
 *
 *   for every impression i and every active module m:
 *       u     = (sum of table_m[hist_ids[h]] over the impression's history) / H_i   cr_module.py:116-123
 *       s^m_j = dot(u, table_m[cand_ids[j]])                                          click_predictors.py:12
 *       zscore: s^m = (s^m - sum(s^m)/C_i) / std_unbiased(s^m)                        ensemble_module.py:137-149
 *   for every weighting w:  s = w[0]*s^0 (+ w[m]*s^m for m >= 1 when w[m] != 0)       ensemble_module.py:97-107
 *       rank by s descending, lower position first on ties; per-impression metrics; fp64 sums.
 */
typedef struct mb200_eval_desc {
  uint32_t struct_size; /* = sizeof(mb200_eval_desc), checked */

  /* embedding tables: n_modules row-major [n_news, dim] arrays of `dtype`, rows `row_stride` elements apart
     (row_stride * element size must be a multiple of 16 bytes; table base 16-byte aligned) */
  int32_t n_modules; /* 1..MB200_MAX_MODULES; module 0 is the CR-Module */
  int32_t dtype;     /* MB200_F32 | MB200_BF16 */
  int32_t dim;       /* embedding width D (text_embedding_dim, configs/model/cr_module.yaml:23) */
  int32_t active_modules_mask; /* bit m set = module m is gathered; bit 0 must be set.  A module whose
                                  weight is 0 in every weighting need not be active
                                  (ensemble_module.py:37-46 does not even load it) */
  int64_t n_news;
  int64_t row_stride;
  const void* tables[MB200_MAX_MODULES];

  /* behaviours in CSR form (the reference's sorted segment-id vectors, mind_rec_dataset.py:114-132,171-174) */
  int64_t n_impressions;
  const int32_t* hist_offsets; /* [n_impressions + 1] */
  const int32_t* hist_ids;     /* [hist_offsets[n_impressions]] table rows, already cut to the first 50 clicks */
  const int32_t* cand_offsets; /* [n_impressions + 1] */
  const int32_t* cand_ids;     /* [cand_offsets[n_impressions]] */
  const uint8_t* labels;       /* [same] 0 / 1 */
  int32_t max_cand;            /* upper bound on candidates per impression (sizes shared memory) */

  /* ensemble */
  int32_t zscore;         /* 0: raw dot products (CRModule); 1: per-impression z-score per module (EnsembleModule) */
  int32_t n_weightings;   /* W >= 1 */
  const float* weights;   /* [W, n_modules] fp32 DEVICE; NULL = a single weighting of all ones */

  /* metric cut-offs (reference: 5 and 10), 1..MB200_MAX_K */
  int32_t k0, k1;

  /* optional aspect labels per news row for Diversity / Personalization; NULL skips slots 5..12 */
  const int32_t* news_category;  /* [n_news] in [0, num_categ_classes) */
  const int32_t* news_sentiment; /* [n_news] in [0, num_sent_classes)  */
  int32_t num_categ_classes;     /* <= MB200_MAX_CLASSES (configs/model/ensemble_module.yaml:8 -> 19) */
  int32_t num_sent_classes;      /* (ensemble_module.yaml:9 -> 4) */

  /* outputs (any may be NULL except sums) */
  float* scores;            /* [sum C] combined scores of weighting `scores_weighting` == the reference's flat `preds` */
  int32_t scores_weighting;
  int32_t pack_payload;     /* != 0: `sums` has MB200_PAYLOAD_TAIL more doubles behind the [W, MB200_NUM_METRICS] block:
                               n_impressions, then one 0/1 double per MB200_FLAG_* bit (1, 2, 4, 8, 64) -- everything
                               a multi-GPU caller sum-reduces, in one buffer, with no host-side packing */
  float* per_impression;    /* [W, n_impressions, MB200_NUM_METRICS] */
  double* sums;             /* [W, MB200_NUM_METRICS] sums over impressions (means = sums / n_impressions) */
  int32_t* flags;           /* one int32, OR-ed with MB200_FLAG_*; the caller zeroes it */

  void* workspace; /* >= mb200_eval_workspace_bytes(desc) bytes, 256-byte aligned */
  size_t workspace_bytes;

  /* early fusion (cr_module.py:124-125 with late_fusion=False: NAMLUserEncoder -> AdditiveAttention, user_encoder.py:9-21,
     attention.py:15-29).  attn_logits[m] = [n_news + 1] fp32 from mb200_attention_logits: query . tanh(W x_n + b) per news
     row, and in the last slot the logit of an all-zero row.  The user vector of module m becomes
     sum_h softmax(logits of the history rows and of hist_pad[i] zero rows)_h * row_h -- the reference does NOT mask the
     rows its dense batch pads with (attention.py:23), so they take softmax mass.  NULL = late fusion (mean). */
  const float* attn_logits[MB200_MAX_MODULES];
  const int32_t* hist_pad; /* optional [n_impressions]: zero rows the reference's step batch appends to impression i's
                              history = (longest history of its step) - H_i; NULL = none */

  /* loss on the scores of `scores_weighting` (slots MB200_M_LOSS / _NONZERO; 0 in the aspect-weight sweep mode) */
  int32_t loss_kind;        /* MB200_LOSS_* */
  float loss_temperature;   /* SupCon temperature (configs/model/cr_module.yaml:6) */
  const int32_t* cand_pad;  /* optional [n_impressions]: zero columns the step batch appends to impression i's candidates
                               (cross entropy only); NULL = none */
  float* loss_per_impression; /* optional [n_impressions] */

  /* Row-sharded embedding tables, for catalogues too large to replicate on every GPU: news row n of module m lives
     in table_shards[m][n >> table_shard_shift] at local row n & ((1 << table_shard_shift) - 1).  The pointers are device pointers
     valid in THIS process: the local shard and peer GPUs' shards opened through CUDA IPC / peer access; the kernel reads a row
     where it lives (plain loads over NVLink / NVSwitch for remote shards), there is no separate exchange step.  n_table_shards
     <= 1: `tables[]` is used (replicated table).  Reference width (768), late fusion. */
  int32_t n_table_shards;
  int32_t table_shard_shift;
  const void* table_shards[MB200_MAX_MODULES][MB200_MAX_TABLE_SHARDS];

  /* Pipelined upload (mb200_upload_begin / _finish): `ready` = DEVICE uint32 that the copy stream raises to the number of leading
     impressions whose hist_ids / cand_ids / labels are resident; the offsets (and pads) are resident when the call is made.  The
     persistent grid then walks `ready_segments` work-balanced chunks per warp, in upload order, and waits on `ready` before it
     touches a chunk (bounded: 4 s -> MB200_FLAG_UPLOAD_TIMEOUT, remaining impressions skipped).  NULL = everything is resident. */
  const uint32_t* ready;
  int32_t ready_segments; /* 1..MB200_MAX_UPLOAD_SEGMENTS when `ready` is set */
  int32_t zero_flags;     /* != 0: the call clears *flags on the stream before it launches (a caller that reuses one flags word per
                             pass need not queue a fill of its own) */
} mb200_eval_desc;

MB200_API int mb200_abi_version(void);
MB200_API const char* mb200_status_str(int status);
MB200_API const char* mb200_last_cuda_error(void); /* HOST string, thread-local, "" if none */

MB200_API size_t mb200_eval_workspace_bytes(const mb200_eval_desc* desc);
MB200_API int mb200_score_eval(const mb200_eval_desc* desc, void* stream);

/*
 * Pipelined host -> device upload of one behaviour set (the CSR arrays of mb200_eval_desc).  The reference moves each batch to
 * the device in front of its forward pass (Lightning's batch transfer before cr_module.py:105 / ensemble_module.py:95); here a
 * pass has ONE CSR set of ~20 MB, and the copy is overlapped with the fused kernel instead of preceding it:
 *   mb200_upload_begin  (compute stream S, copy stream C != S):  S: ready = 0;  C waits for S, copies the two offset arrays (+ the
 *                        optional pads), S waits for that;  C: the first `segments_first` segments of hist_ids / cand_ids /
 *                        labels, each followed by a 4-byte copy that raises `ready` to the segment's last impression + 1.  The
 *                        remaining segments are queued on C by a library thread, concurrently with the caller.
 *   mb200_score_eval    on S with desc.ready / desc.ready_segments = n_segments: runs while the copies are still arriving
 *   mb200_upload_finish  waits (host side) until the library thread has queued every segment and returns its status; call it
 *                        before the host arrays or the descriptor go away.
 * segments_first = n_segments queues every copy on the caller's thread in front of the launch (the Python layer's default: ~2 us
 * per copy).  With fewer, the library thread overlaps the rest with the caller's launch; plain blocking launches
 * (CUDA_LAUNCH_BLOCKING=1) are fine with that, but a tool that serialises ALL CUDA calls behind a running kernel (ncu) can keep
 * the thread's copies from being issued while the kernel waits for them (-> MB200_FLAG_UPLOAD_TIMEOUT after 4 s).
 * Segments grow geometrically (the first holds 1/2^(n_segments-1) of the rows gathered, each later one as much as all before it),
 * so little is exposed in front of the kernel's first chunk.  Copies cover disjoint 128-byte aligned ranges.  All HOST arrays must be page-locked and stay untouched until C has run the copies.
 */
typedef struct mb200_upload_desc {
  uint32_t struct_size;   /* = sizeof(mb200_upload_desc) */
  int32_t n_segments;     /* 1..MB200_MAX_UPLOAD_SEGMENTS */
  int32_t segments_first; /* 0..n_segments: queued by _begin itself; the rest by the library thread */
  int32_t reserved;
  int64_t n_impressions;  /* >= 1 */
  const int32_t* h_hist_offsets; /* HOST [n_impressions + 1] */
  const int32_t* h_hist_ids;     /* HOST */
  const int32_t* h_cand_offsets; /* HOST [n_impressions + 1] */
  const int32_t* h_cand_ids;     /* HOST */
  const uint8_t* h_labels;       /* HOST */
  int32_t* d_hist_offsets;
  int32_t* d_hist_ids;  /* 128-byte aligned, like d_cand_ids and d_labels */
  int32_t* d_cand_offsets;
  int32_t* d_cand_ids;
  uint8_t* d_labels;
  const int32_t* h_hist_pad; /* optional HOST [n_impressions] (early fusion / cross entropy), with its destination */
  int32_t* d_hist_pad;
  const int32_t* h_cand_pad;
  int32_t* d_cand_pad;
  uint32_t* ready;   /* DEVICE, one uint32 */
  uint32_t* h_marks; /* HOST page-locked scratch [n_segments], written by these calls */
  void* copy_stream;
} mb200_upload_desc;

MB200_API int mb200_upload_begin(const mb200_upload_desc* desc, void* compute_stream);
MB200_API int mb200_upload_finish(const mb200_upload_desc* desc);

/*
 * Per-news additive-attention logits for early fusion: out[n] = query . tanh(weight x_n + bias) for the n_rows rows of
 * `table` (fp32 arithmetic), out[n_rows] = query . tanh(bias) (the logit of a padded, all-zero history row).  weight
 * [q_dim, dim], bias [q_dim], query [q_dim] fp32 = NAMLUserEncoder.additive_attention.{linear.weight, linear.bias, query}
 * (attention.py:9-13).  The logit depends on the news row only, so it is cached next to the embedding table instead of
 * being recomputed per impression (attention.py:20-24 runs the linear layer on every padded history row of every step).
 * dim % 4 == 0, dim <= 1024, q_dim <= 1024.
 */
MB200_API int mb200_attention_logits(const void* table, int dtype, int dim, int64_t row_stride, int64_t n_rows, const float* weight,
                                     const float* bias, const float* query, int q_dim, float* out /* [n_rows + 1] */, void* stream);

/*
 * The reference logs `test/loss` / `val/loss` as a MeanMetric over its steps (cr_module.py:214-225,253-259): every step of
 * `step` consecutive impressions contributes one value -- the mean of its impressions' losses (cross entropy), or the mean
 * of the losses that are > 0 (SupCon, AvgNonZeroReducer).  SupCon additionally has two STEP-level guards
 * (components/losses.py:15-16 `all(len(x) <= 1 for x in indices_tuple)` and :22 `pos_mask.any() and neg_mask.any()`): a step
 * with at most one positive and at most one negative candidate in total, or without any positive or any negative, is worth 0;
 * they need `cand_offsets` [n_impressions + 1] and `labels` (both NULL: guards skipped).  out (device, 2 doubles) = {sum of the
 * step values, number of steps}: both additive over ranks when every rank's shard starts on a step boundary.
 */
MB200_API int mb200_step_loss(const float* loss_per_impression, int64_t n_impressions, int step, int loss_kind, const int32_t* cand_offsets,
                              const uint8_t* labels, double* out, void* stream);

/*
 * Ranking metrics on scores that already exist: the seam of torchmetrics' `update(preds, target, indexes)` / `compute()` that
 * every module of the reference -- CRModule, EnsembleModule and the nine baseline recommenders -- ends its epoch with
 * (cr_module.py:266-274, ensemble_module.py:214-256, nrms_plm_module.py:275-313), and of the in-repo Diversity /
 * Personalization (metrics/diversity.py, metrics/personalization.py, metrics/base.py:62-129).  Same ranking rule, same
 * per-impression arithmetic and the same `sums` layout as mb200_score_eval, without the gather: `preds` is the flat [sum C]
 * prediction vector with every impression's candidates contiguous (`cand_offsets` = prefix sums of `cand_news_size`).
 * Aspect metrics take per-ROW labels as the reference's metric objects do (`target_categories`, `hist_categories`, ...).
 * The pooled AUROC of the same vector is mb200_pooled_auc.
 */
typedef struct mb200_metrics_desc {
  uint32_t struct_size; /* = sizeof(mb200_metrics_desc) */
  int32_t k0, k1;       /* 1..MB200_MAX_K */
  int32_t max_cand;     /* upper bound on candidates per impression */
  int64_t n_impressions;
  const float* preds;          /* [sum C] */
  const uint8_t* labels;       /* [sum C] 0 / 1 */
  const int32_t* cand_offsets; /* [n_impressions + 1] */
  /* optional: all five or none */
  const int32_t* cand_category;  /* [sum C] in [0, num_categ_classes) */
  const int32_t* cand_sentiment; /* [sum C] */
  const int32_t* hist_offsets;   /* [n_impressions + 1] */
  const int32_t* hist_category;  /* [sum H] */
  const int32_t* hist_sentiment; /* [sum H] */
  int32_t num_categ_classes, num_sent_classes;
  float* per_impression; /* optional [n_impressions, MB200_NUM_METRICS] */
  double* sums;          /* [MB200_NUM_METRICS] */
  int32_t* flags;        /* optional; the caller zeroes it */
  void* workspace;       /* >= mb200_metrics_workspace_bytes(desc), 256-byte aligned */
  size_t workspace_bytes;
} mb200_metrics_desc;

MB200_API size_t mb200_metrics_workspace_bytes(const mb200_metrics_desc* desc);
MB200_API int mb200_rank_metrics(const mb200_metrics_desc* desc, void* stream);

/*
 * Pooled AUROC exactly as torchmetrics 0.11.4 `AUROC(task="binary")` defines it (cr_module.py:81,273;
 * SURVEY a12/A6): all candidate rows of the epoch in one pool, fp32 sigmoid applied iff some pred is
 * outside [0,1], ties get half credit.  Evaluated as the exact rank statistic
 *     auc = sum over positives (#neg below + #neg not above) / (2 * P * N)
 * in integer arithmetic, in three stages so that a multi-GPU caller can exchange the (few) positive
 * keys between stages 2 and 3:
 *   1. build_keys : order-preserving uint32 key per row; negatives stay in place (positives become
 *                   0xFFFFFFFF), positives are appended to `pos_keys`; counts[0] = P
 *   2. sort_keys  : ascending radix sort of the n keys (the negatives end up in [0, n - P))
 *   3. rank_sum   : for every positive key, lower_bound + upper_bound in the sorted negatives, added
 *                   into *sum2 (uint64)
 * sigmoid_mode: 0 never, 1 always, 2 = iff (*flags & MB200_FLAG_OUTSIDE_UNIT).
 */
MB200_API int mb200_auc_build_keys(const float* preds, const uint8_t* labels, int64_t n, int sigmoid_mode,
                         const int32_t* flags, uint32_t* neg_keys, uint32_t* pos_keys,
                         int64_t* n_pos /* device, zeroed by this call */, void* stream);
MB200_API size_t mb200_auc_sort_workspace_bytes(int64_t n);
MB200_API int mb200_auc_sort_keys(const uint32_t* keys_in, uint32_t* keys_out, int64_t n, void* workspace,
                        size_t workspace_bytes, void* stream);
MB200_API int mb200_auc_rank_sum(const uint32_t* sorted_keys, int64_t n_sorted, const int64_t* n_pos_local /* device: sorted negatives = n_sorted - *n_pos_local */,
                       const uint32_t* pos_keys, int64_t pos_capacity, const int64_t* n_pos /* device: entries of pos_keys to use */,
                       uint64_t* sum2 /* device, accumulated (caller zeroes) */, void* stream);

/* Single-GPU convenience: the three stages + the division.  out (device, 4 doubles) = {auc, P, N, sum2}.
 * auc = 0 when P == 0 or N == 0 (torchmetrics returns 0 with a warning). */
MB200_API size_t mb200_pooled_auc_workspace_bytes(int64_t n);
MB200_API int mb200_pooled_auc(const float* preds, const uint8_t* labels, int64_t n, int sigmoid_mode,
                     const int32_t* flags, void* workspace, size_t workspace_bytes, double* out, void* stream);

/* The same statistic when the caller knows an upper bound on the number of positives (`pos_capacity`; the labels are host data
 * wherever behaviours are uploaded from) and they are few -- a click log has ~4 %: by symmetry
 *     sum over positives (#neg below + #neg not above)  ==  sum over negatives (#pos above + #pos not below),
 * so only the POSITIVES are sorted (0.02 ms instead of 0.12 ms for MIND-small) and the negatives are ranked against them in one
 * streaming pass.  Same `out`; if the rows hold more than pos_capacity positives, out[0] is NaN and out[1] the true count. */
MB200_API size_t mb200_pooled_auc_bounded_workspace_bytes(int64_t n, int64_t pos_capacity);
MB200_API int mb200_pooled_auc_bounded(const float* preds, const uint8_t* labels, int64_t n, int64_t pos_capacity, int sigmoid_mode,
                                       const int32_t* flags, void* workspace, size_t workspace_bytes, double* out, void* stream);

/*
 * Full-catalog retrieval (BASELINE.json configs[4]; no reference counterpart -- the reference scores only an
 * impression's candidates, cr_module.py:105-131): scores = users [n_users, dim] x catalog [n_catalog, dim]^T in
 * bf16 on the tcgen05 tensor cores (fp32 accumulation in TMEM), per-user top-k fused into the epilogue.
 * Output lists are sorted by score descending, catalogue id ascending on ties; ids are int64 =
 * catalogue row + catalog_id_offset (row-sharded catalogues); unused slots hold -inf / -1.
 */
typedef struct mb200_retrieval_desc {
  uint32_t struct_size;  /* = sizeof(mb200_retrieval_desc) */
  int32_t dim;           /* multiple of 64 (768 in the reference's configs) */
  int32_t k;             /* 1..128 */
  int32_t reserved;
  int64_t n_users;
  int64_t n_catalog;
  int64_t catalog_id_offset;
  const void* users;    /* bf16 [n_users, dim] row-major, 16-byte aligned (mb200_pool_users makes it) */
  const void* catalog;  /* bf16 [n_catalog, dim] row-major, 16-byte aligned */
  float* out_scores;    /* [n_users, k] */
  int64_t* out_ids;     /* [n_users, k] */
  float* debug_scores;  /* optional [n_users, n_catalog] fp32: the full score matrix (tests only) */
  void* workspace;      /* >= mb200_retrieval_workspace_bytes(desc), 256-byte aligned */
  size_t workspace_bytes;

  /* Fused exchange for a row-sharded catalogue (replaces the NCCL all-gather of the per-shard lists): when n_peers > 1 every
     finished user row is also STORED into the gather buffers of the other GPUs over NVLink / NVSwitch peer memory, while
     the kernel keeps scoring the next user tiles.  peer_scores[r] / peer_ids[r] = GPU r's gather buffers
     [n_peers, peer_rows, k] as seen from this process (mb200_ipc_open); this rank writes slot `my_rank`, rows
     [0, n_users).  out_scores / out_ids must then be this rank's own slot of its own gather buffer
     (peer_scores[my_rank] + my_rank * peer_rows * k).  After a cross-GPU completion signal (any collective on the same
     stream) each GPU merges its buffer with mb200_merge_topk. */
  int32_t n_peers; /* 0 / 1 = no fused exchange */
  int32_t my_rank;
  int64_t peer_rows;
  float* peer_scores[MB200_MAX_TABLE_SHARDS];
  int64_t* peer_ids[MB200_MAX_TABLE_SHARDS];
} mb200_retrieval_desc;

MB200_API size_t mb200_retrieval_workspace_bytes(const mb200_retrieval_desc* desc);
MB200_API int mb200_retrieve_topk(const mb200_retrieval_desc* desc, void* stream);
/* user matrix for retrieval: out[u] = bf16(mean of table[hist_ids[h]] over user u's history), the late-fusion
 * user vector of cr_module.py:116-123.  dtype MB200_F32 | MB200_BF16, dim even and <= 1024. */
MB200_API int mb200_pool_users(const void* table, int dtype, int dim, int64_t row_stride, int64_t n_news, const int32_t* hist_offsets,
                               const int32_t* hist_ids, int64_t n_users, void* out_bf16, int32_t* flags, void* stream);
/* merge of per-shard top-k lists [shards, n_users, k] (each sorted as above) into the global top-k (shards <= 32) */
MB200_API int mb200_merge_topk(const float* scores, const int64_t* ids, int shards, int64_t n_users, int k, float* out_scores,
                               int64_t* out_ids, void* stream);

/*
 * Fused multi-GPU exchange of one evaluation (the reference is single-device, configs/trainer/default.yaml:9, so
 * the contract is "same numbers as one GPU").  Replaces the NCCL all-reduce of the metric payload, the all-gather of the
 * positive keys and the all-reduce of the AUROC statistics: `mb200_exchange_post` STORES this rank's payload (the packed `sums` of
 * mb200_score_eval with pack_payload) and the raw keys of its positives into slot `my_rank` of EVERY rank's mailbox over
 * NVLink / NVSwitch peer memory (and, while those stores fly, prepares the sigmoid keys of its sorted negatives);
 * `mb200_exchange_finish` (same stream, after it) waits for all ranks' stores, sums the payloads in rank order into
 * `out_payload` (bit-identical on every rank), ranks all ranks' positives against this rank's sorted negatives, exchanges the
 * three additive integers the same way and writes their sums to `out_stats` = {sum2, P, N} (auc = sum2 / (2 P N)).
 * mailbox[r] = rank r's mailbox as mapped in THIS process (mb200_ipc_open; own = local memory),
 * mb200_exchange_mailbox_bytes() bytes each, zero-filled once before the first exchange.  `epoch` = 1, 2, 3, ... must advance by
 * one per exchange on every rank.  Keys: mb200_auc_build_keys with sigmoid_mode 0 + mb200_auc_sort_keys (RAW score order: no
 * dependency on the other ranks, so the sort overlaps their kernels); whether torchmetrics' AUROC applies the sigmoid is only
 * known from the reduced payload (entry `outside_index` > 0 on any rank) -- the fp32 sigmoid is monotone, so the raw order is a
 * valid order of the sigmoid keys too (ties only merge) and the search simply runs on the prepared sigmoid keys.
 * Waits are bounded (4 s -> MB200_FLAG_EXCHANGE_TIMEOUT).  One process per GPU: the kernels of different ranks run on different GPUs.
 */
typedef struct mb200_exchange_desc {
  uint32_t struct_size; /* = sizeof(mb200_exchange_desc) */
  int32_t n_ranks;      /* 1..MB200_MAX_TABLE_SHARDS */
  int32_t my_rank;
  uint32_t epoch;       /* >= 1 */
  int32_t n_payload;    /* doubles in the payload */
  int32_t outside_index; /* payload entry that is > 0 when a score of that rank was outside [0,1]; -1 = never apply the sigmoid */
  int64_t pos_capacity; /* positive keys a mailbox slot can hold (same on every rank) */
  void* mailbox[MB200_MAX_TABLE_SHARDS];
  const double* payload;      /* [n_payload] this rank's */
  const uint32_t* pos_keys;   /* this rank's positive keys (mb200_auc_build_keys, sigmoid_mode 0) */
  const int64_t* n_pos;       /* device: how many */
  const uint32_t* sorted_neg; /* [n_rows] mb200_auc_sort_keys output: negatives in [0, n_rows - *n_pos) */
  int64_t n_rows;             /* 0 = no pooled AUROC wanted */
  double* out_payload;        /* [n_payload] sums over ranks */
  int64_t* out_stats;         /* [3] */
  int32_t* flags;             /* optional: TWO int32 (an int64 slot), zeroed by _post, then OR-ed with MB200_FLAG_EXCHANGE_TIMEOUT /
                                 MB200_FLAG_POS_OVERFLOW by _finish */
  void* workspace;            /* >= mb200_exchange_workspace_bytes(n_rows), 256-byte aligned, this rank's own */
  size_t workspace_bytes;
} mb200_exchange_desc;

MB200_API size_t mb200_exchange_mailbox_bytes(int n_ranks, int n_payload, int64_t pos_capacity);
MB200_API size_t mb200_exchange_workspace_bytes(int64_t n_rows);
MB200_API int mb200_exchange_post(const mb200_exchange_desc* desc, void* stream);
MB200_API int mb200_exchange_finish(const mb200_exchange_desc* desc, void* stream);

/* Read-bandwidth probe for the roofline denominators bench.py reports: every warp streams 3 KB rows of `buf` (the access shape
 * of the row gather: six 16-byte loads per lane and row; rows of a batch far apart) for `repeats` passes.  bytes / 3072 must be
 * a power of two.  A buffer that fits the 126 MB L2 measures the L2 -> SM read bandwidth a gather can reach at best; a larger
 * one the HBM read bandwidth.  mode 0 / 1 / 2 = 16 warps x 4 rows in flight / 32 x 2 / 64 x 1 per SM; the caller times it with
 * events and takes the best.  `sink`: 4 bytes of device memory. */
MB200_API int mb200_read_probe(const void* buf, size_t bytes, int repeats, int mode, void* sink, void* stream);

/* Lets kernels running on `device` load from memory that lives on `peer` (cudaDeviceEnablePeerAccess; "already enabled" is
 * not an error): needed once per pair of GPUs before row-sharded tables (mb200_eval_desc.table_shards) are used. */
MB200_API int mb200_enable_peer_access(int device, int peer);

/* CUDA IPC for row-sharded tables: `mb200_ipc_export` describes the allocation that contains `ptr` (64-byte IPC handle of its
 * base + the offset of `ptr` in it); a PEER PROCESS passes both to `mb200_ipc_open`, which maps the allocation with
 * cudaIpcMemLazyEnablePeerAccess while `device` (the GPU whose kernels will read it) is current and returns the address of
 * `ptr` in the calling process.  The exporter keeps the memory alive; a mapping lives until mb200_ipc_close or the end of the process. */
MB200_API int mb200_ipc_export(const void* ptr, unsigned char handle[64], int64_t* offset);
MB200_API int mb200_ipc_open(const unsigned char handle[64], int64_t offset, int device, void** out_ptr);
/* Unmaps an allocation mapped by mb200_ipc_open: `mapped_base` = the returned pointer minus the offset that was passed in.  Every
 * pointer into the mapping is dead afterwards; the caller makes sure no kernel that uses it is still running. */
MB200_API int mb200_ipc_close(void* mapped_base, int device);

/* ---- introspection ------------------------------------------------------------------------------- */
/* 1 / log2(rank + 1) as fp32, rank = 1..MB200_MAX_K: the discount table the kernels use for
 * torchmetrics' `_dcg` (HOST function; lets CPU tests pin it against torch.log2). */
MB200_API float mb200_dcg_discount(int rank);
/* number of kernels this library has launched in this process (own kernels; CUB sort passes counted
 * separately by mb200_library_launch_count). */
MB200_API int64_t mb200_launch_count(void);
MB200_API int64_t mb200_library_launch_count(void);
/* tuning knobs; returns the previous value.  key 0 = impression chunks per resident warp of the fused kernel (0 = default: 1 when the
 * behaviours are resident, 16 handed out dynamically in impression order under a pipelined upload; key 7 = 1 / 2 forces the
 * static / dynamic schedule); 1 = variant of the reference-width fused
 * kernel (-1 = default: 2 for fp32 rows, 3 for bf16 rows; 0/2/3/5/6 = 4-warp CTAs, rows in flight x resident CTAs per SM:
 * 4x3, 3x4, 2x5, 3x6, 2x7 for fp32 rows, twice the rows for bf16; 1 = 4x3 with L1::no_allocate loads; measured and not shipped:
 * 9 = one 16-warp CTA per SM with a hot-row cache in shared memory, 10 = that CTA shape without the cache, 7 / 8 = rotating row
 * pipeline without / with the cache, 11 = bf16 candidates on mma.sync);
 * 2 = cap on CTAs per SM; 3 = time the fused kernel with
 * CUDA events; 4 = retrieval diagnostics (1, 2: parts of the epilogue disabled, RESULTS INVALID; 4: cycle counters in the
 * workspace header, results valid); 5 = retrieval pipeline (1 = CTA pairs / tcgen05 cta_group::2 [default], 0 = one CTA per tile); 6 = cap in KB on the
 * hot-row cache of variants 8 / 9 (0 = no cap); 7 = chunk schedule of the fused kernel (0 auto, 1 static, 2 dynamic); 8 = retrieval sweep
 * throttle: catalogue tiles a CTA may run ahead of the slowest one (default 48, 0 = off) */
MB200_API int mb200_set_tuning(int key, int value);

/* duration in ms of the most recent fused score/eval kernel launched while tuning key 3 was on
 * (synchronises on its end event; -1 if none).  This is the kernel bench.py reports a roofline for. */
MB200_API float mb200_last_score_kernel_ms(void);
/* ms from `event` (a cudaEvent_t recorded with timing by the caller) to the begin of that kernel: how long the stream waited for
 * the host to get the kernel launched (-1 if unavailable; synchronises). */
MB200_API float mb200_last_score_kernel_begin_after(void* event);

/* Hot-row cache of the most recent mb200_score_eval launch in this process (HOST int32 out[4]; synchronises the device):
 * {rows cached per module (0 = cache off for that behaviour set), sampled row reads, sampled row reads that hit the cached rows,
 * slots per module the launch had room for}.  covered / sampled = the fraction of the gather served from shared memory
 * instead of the L2 -- bench.py reports it beside the roofline. */
MB200_API int mb200_last_hot_stats(int32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MANNER_B200_H */
