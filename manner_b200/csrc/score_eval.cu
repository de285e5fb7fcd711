// Fused gather -> late-fusion mean pool -> dot-product scores -> z-score -> aspect-weighted ensemble ->
// per-impression ranking metrics, one warp per impression, for sm_100a.
//
// Replaces (reference = andreeaiana/manner, file:line under /root/reference/manner):
//   news_encoder(x_hist / x_cand) on cached embeddings ........ models/cr_module.py:107,113 (row gather)
//   to_dense_batch x2, hist_size loop, sum(dim=1) / hist_size .. models/cr_module.py:108-123
//   DotProduct.forward (bmm) ................................... models/components/click_predictors.py:9-12
//   z-score + `scores += w * z` ................................ models/ensemble_module.py:95-109,137-149
//   mask-flatten loops, epoch-end cat / indexes ................ models/cr_module.py:173-182,266-271
//   RetrievalMRR / RetrievalNormalizedDCG(k) per impression ..... models/cr_module.py:82-84 (torchmetrics 0.11.4)
//   Diversity / Personalization per impression .................. metrics/functional.py:8-70, metrics/base.py:92-129
//
// Data layout: tables row-major [n_news, dim] (fp32 or bf16) resident in HBM; behaviours as CSR
// (int32 offsets / ids, uint8 labels).  Nothing is padded.  A warp owns one impression at a time:
// each lane keeps dim/32 elements of the user vector in registers, rows are fetched as coalesced
// 16-byte vectors (lane l reads vectors l, l+32, ...), R rows in flight per warp, dot products are
// finished with warp shuffles, scores of the impression live in shared memory for ranking.
// HBM-bound: 0.5 flop per byte -- tensor cores are deliberately not used here.

#include <float.h>
#include <math.h>

#include "common.cuh"

namespace mb200 {

__constant__ float c_inv_disc[32] = {MB200_INV_DISC_TABLE};
static const float h_inv_disc[32] = {MB200_INV_DISC_TABLE};
float host_dcg_discount(int rank) { return (rank >= 1 && rank <= MB200_MAX_K) ? h_inv_disc[rank - 1] : 0.0f; }

struct EvalParams {
  const void* tables[MB200_MAX_MODULES];
  const void* shard_base[MB200_MAX_MODULES][MB200_MAX_TABLE_SHARDS];  // row-sharded tables: shard s of module m (own or peer-GPU memory)
  int shard_shift;  // rows per shard = 1 << shard_shift
  const int32_t* hist_offsets;
  const int32_t* hist_ids;
  const int32_t* cand_offsets;
  const int32_t* cand_ids;
  const uint8_t* labels;
  const float* weights;
  const int32_t* news_category;
  const int32_t* news_sentiment;
  const int32_t* flat_cand_aspect[2];  // rank_metrics: per-row category / sentiment labels instead of per-news tables
  const int32_t* flat_hist_aspect[2];
  const float* scores_in;  // rank_metrics: flat predictions
  const float* attn_logits[MB200_MAX_MODULES];  // [n_news + 1] per module, or null (late fusion)
  const int32_t* hist_pad;
  const int32_t* cand_pad;
  float* loss_per_impr;
  float* scores;
  float* per_impr;
  double* partials;  // [W][MB200_NUM_METRICS][n_partials]: one slot per chunk (score_eval_kernel) or per warp (the other kernels), slot
                     // index innermost so that the reduction reads every (weighting, metric) row contiguously
  int n_partials;
  const int32_t* bounds;  // [n_chunks + 1] impression boundaries of the work-balanced chunks
  int* chunk_counter;     // score_eval_kernel: next chunk to hand out (zeroed by partition_kernel); null = static round-robin
  int32_t* flags;
  long long n_news;
  long long row_stride;  // elements
  int n_impr;
  int n_modules;
  int active_mask;
  int vec_per_row;  // 16-byte vectors per row
  int zscore;
  int n_weightings;
  int scores_weighting;
  int k0, k1;
  int cpad;           // max_cand rounded up to 32
  int max_cand;
  int n_chunks;
  int num_categ, num_sent;
  int smem_per_warp;  // bytes
  int acc_bytes;      // bytes of the per-warp fp64 accumulators (16-byte multiple)
  int acc_stride;     // doubles per weighting in the accumulators: MB200_NUM_METRICS, or kSweepSlots in the aspect-weight sweep mode
  int loss_kind;
  float loss_temperature;
  // hot-row cache of the streaming kernel (score_eval_stream_kernel): the rows gathered most often in THIS behaviour set are
  // kept in shared memory for the whole launch
  const struct HotDir* hot_dir;   // counters written by hot_select_kernel
  const int32_t* hot_ids;         // [hot_cap] news id cached in slot s
  const uint8_t* slot_of;         // [n_news] slot of a news id, kHotCold = not cached
  int hot_cap;                    // slots per module the launch reserved shared memory for
  int hot_bytes;                  // bytes of the cache at the front of the CTA's shared memory
  int upart_off;                  // byte offset, in the per-warp area, of the three bf16 parts of the user vector (tensor-core path)
  int has_aspects;                // the per-warp area holds the aspect buffers
  int comb_alias;                 // one weighting: the combined scores overwrite module 0's (no separate buffer)
  const unsigned int* ready;      // pipelined upload: leading impressions whose ids / labels are resident (null: all)
};

struct HotDir {
  int32_t n_hot;    // rows cached this launch (0: cache off, e.g. uniform ids)
  int32_t total;    // sampled row reads
  int32_t covered;  // of which go to the cached rows
  int32_t reserved;
};
constexpr int kHotCold = 255;
constexpr int kHotMaxSlots = 254;

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&f)[4]) {
    f[0] = __uint_as_float(r.x), f[1] = __uint_as_float(r.y), f[2] = __uint_as_float(r.z), f[3] = __uint_as_float(r.w);
  }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&f)[8]) {
    f[0] = __uint_as_float(r.x << 16), f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16), f[3] = __uint_as_float(r.y & 0xffff0000u);
    f[4] = __uint_as_float(r.z << 16), f[5] = __uint_as_float(r.z & 0xffff0000u);
    f[6] = __uint_as_float(r.w << 16), f[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};

template <int POLICY>
__device__ __forceinline__ uint4 load_row_vec(const uint4* p) {
  if constexpr (POLICY == 1) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
  } else {
    return __ldg(p);
  }
}

// torch sort semantics: NaN is the largest value; equal NaNs tie.
__device__ __forceinline__ bool ranks_before(float a, float b) { return (a > b) || (a != a && b == b); }
__device__ __forceinline__ bool ranks_equal(float a, float b) { return (a == b) || (a != a && b != b); }
// The same order as one unsigned integer: key(a) > key(b) <=> ranks_before(a, b), key(a) == key(b) <=> ranks_equal(a, b).  Every NaN
// maps to the largest key, -0 to +0's; the smallest key of a real float (-inf) is 0x007fffff, so 0 is free as "none".
__device__ __forceinline__ unsigned rank_key(float x) {
  if (x != x) return 0xffffffffu;
  if (x == 0.0f) x = 0.0f;
  const unsigned b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Sum of x_j = (bit j of mask) ? 1/log2(j+2) : 0 over j < L in the fp32 summation order of ATen's
// CPU `sum` over a contiguous vector (SumKernel.cpp: 8-lane vectors, scalar tail first, then the
// lanes in order; L < 8 takes the 4-way ILP scalar path) -- what torchmetrics' `_dcg` evaluates.
__device__ __forceinline__ float dcg_term(unsigned mask, int j) { return ((mask >> j) & 1u) ? c_inv_disc[j] : 0.0f; }
__device__ __noinline__ float dcg_sum_aten_order(unsigned mask, int L) {
  if (L < 8) {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    const int r = L >> 2;
    if (r) p0 = dcg_term(mask, 0), p1 = dcg_term(mask, 1), p2 = dcg_term(mask, 2), p3 = dcg_term(mask, 3);
#pragma unroll 1
    for (int j = 4 * r; j < L; ++j) p0 = __fadd_rn(p0, dcg_term(mask, j));
    p0 = __fadd_rn(p0, p1), p0 = __fadd_rn(p0, p2), p0 = __fadd_rn(p0, p3);
    return p0;
  }
  const int vs = L >> 3;  // 1..3 because L <= MB200_MAX_K
  float fin = 0.f;
#pragma unroll 1
  for (int j = 8 * vs; j < L; ++j) fin = __fadd_rn(fin, dcg_term(mask, j));
#pragma unroll 1
  for (int l8 = 0; l8 < 8; ++l8) {
    float a = dcg_term(mask, l8);
#pragma unroll 1
    for (int v = 1; v < vs; ++v) a = __fadd_rn(a, dcg_term(mask, 8 * v + l8));
    fin = __fadd_rn(fin, a);
  }
  return fin;
}

__device__ __noinline__ float ndcg_at(unsigned hit_mask, int n_pos, int n_cand, int k) {
  const int L = min(k, n_cand);
  const int ideal_hits = min(n_pos, L);
  if (ideal_hits == 0) return 0.f;
  const float idcg = dcg_sum_aten_order((ideal_hits >= 32) ? 0xffffffffu : ((1u << ideal_hits) - 1u), L);
  const unsigned keep = (L >= 32) ? 0xffffffffu : ((1u << L) - 1u);
  const float dcg = dcg_sum_aten_order(hit_mask & keep, L);
  return __fdiv_rn(dcg, idcg);
}

// ensemble_module.py:137-149.  mean = fp32 row sum / C; std = unbiased, accumulated in fp64 and
// rounded to fp32 (ATen's CPU std of a float tensor accumulates in double).  C == 1 gives NaN.
__device__ __noinline__ void zscore_inplace(float* s, int C, int lane) {
  __builtin_assume(__isShared(s));
  float part = 0.f;
  double dpart = 0.0;
#pragma unroll 1
  for (int j = lane; j < C; j += 32) {
    const float v = s[j];
    part += v;
    dpart += (double)v;
  }
  const float mean = __fdiv_rn(warp_sum(part), (float)C);
  const double dmean = warp_sum(dpart) / (double)C;
  double q = 0.0;
#pragma unroll 1
  for (int j = lane; j < C; j += 32) {
    const double d = (double)s[j] - dmean;
    q += d * d;
  }
  const float sd = (float)sqrt(warp_sum(q) / (double)(C - 1));
#pragma unroll 1
  for (int j = lane; j < C; j += 32) s[j] = __fdiv_rn(__fsub_rn(s[j], mean), sd);
}

// Entropy-based Diversity@k of one aspect (metrics/functional.py:8-28): lanes own classes.
__device__ __noinline__ float diversity_value(const uint8_t* top, int kk, int num_classes, int lane) {
  __builtin_assume(__isShared(top));
  float prob[2];
  float total = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int c = lane + 32 * t;
    int cnt = 0;
    if (c < num_classes)
#pragma unroll 1
      for (int r = 0; r < kk; ++r) cnt += (top[r] == c);
    prob[t] = __fdiv_rn((float)cnt, (float)num_classes);
    total += prob[t];
  }
  total = warp_sum(total);
  float ent = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int c = lane + 32 * t;
    if (c < num_classes) {
      const float p = __fdiv_rn(prob[t], total);
      const float pc = fminf(fmaxf(p, FLT_EPSILON), 1.0f - FLT_EPSILON);
      ent += __fmul_rn(logf(pc), p);
    }
  }
  ent = -warp_sum(ent);
  return __fdiv_rn(ent, logf((float)num_classes));
}

// Generalised Jaccard of top-k candidate aspect counts vs history aspect counts (functional.py:31-70).
__device__ __noinline__ float personalization_value(const uint8_t* top, int kk, const int* hist_count, int num_classes, int lane) {
  __builtin_assume(__isShared(top));
  __builtin_assume(__isShared(hist_count));
  int mn = 0, mx = 0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int c = lane + 32 * t;
    if (c < num_classes) {
      int cnt = 0;
#pragma unroll 1
      for (int r = 0; r < kk; ++r) cnt += (top[r] == c);
      const int hc = hist_count[c];
      mn += min(cnt, hc);
      mx += max(cnt, hc);
    }
  }
  mn = __reduce_add_sync(kFull, mn);
  mx = __reduce_add_sync(kFull, mx);
  return __fdiv_rn((float)mn, (float)mx);
}

// Diversity@k0, Diversity@k1, Personalization@k0, Personalization@k1 of ONE aspect in one go: the class counts of both cut-offs come
// from a single scan of the top-k labels, and the two entropy chains (division, log, two warp reductions each) run interleaved
// instead of back to back -- with the aspect metrics on, these dependent chains, eight calls per impression, were the largest
// stall source after the row gathers.  Every value is computed with the operation order of diversity_value / personalization_value
// (bit-identical results; those two stay as the single-value reference forms).
struct AspectValues {
  float div0, div1, pers0, pers1;
};
__device__ __noinline__ AspectValues aspect_values(const uint8_t* top, int kk0, int kk1, const int* hist_count, int num_classes, int lane) {
  __builtin_assume(__isShared(top));
  __builtin_assume(__isShared(hist_count));
  const int kmax = max(kk0, kk1);
  int cnt0[2], cnt1[2];
  int mn0 = 0, mx0 = 0, mn1 = 0, mx1 = 0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int c = lane + 32 * t;
    cnt0[t] = cnt1[t] = 0;
    if (c < num_classes) {
#pragma unroll 1
      for (int r = 0; r < kmax; ++r) {
        const int m = (top[r] == c) ? 1 : 0;
        cnt0[t] += (r < kk0) ? m : 0;
        cnt1[t] += (r < kk1) ? m : 0;
      }
      const int hc = hist_count[c];
      mn0 += min(cnt0[t], hc), mx0 += max(cnt0[t], hc);
      mn1 += min(cnt1[t], hc), mx1 += max(cnt1[t], hc);
    }
  }
  float prob0[2], prob1[2], total0 = 0.f, total1 = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    prob0[t] = __fdiv_rn((float)cnt0[t], (float)num_classes);
    prob1[t] = __fdiv_rn((float)cnt1[t], (float)num_classes);
    total0 += prob0[t], total1 += prob1[t];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    total0 += __shfl_xor_sync(kFull, total0, o);
    total1 += __shfl_xor_sync(kFull, total1, o);
  }
  float ent0 = 0.f, ent1 = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int c = lane + 32 * t;
    if (c < num_classes) {
      const float p0 = __fdiv_rn(prob0[t], total0), p1 = __fdiv_rn(prob1[t], total1);
      const float pc0 = fminf(fmaxf(p0, FLT_EPSILON), 1.0f - FLT_EPSILON), pc1 = fminf(fmaxf(p1, FLT_EPSILON), 1.0f - FLT_EPSILON);
      ent0 += __fmul_rn(logf(pc0), p0);
      ent1 += __fmul_rn(logf(pc1), p1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ent0 += __shfl_xor_sync(kFull, ent0, o);
    ent1 += __shfl_xor_sync(kFull, ent1, o);
  }
  const float lognc = logf((float)num_classes);
  AspectValues v;
  v.div0 = __fdiv_rn(-ent0, lognc), v.div1 = __fdiv_rn(-ent1, lognc);
  mn0 = __reduce_add_sync(kFull, mn0), mx0 = __reduce_add_sync(kFull, mx0);
  mn1 = __reduce_add_sync(kFull, mn1), mx1 = __reduce_add_sync(kFull, mx1);
  v.pers0 = __fdiv_rn((float)mn0, (float)mx0), v.pers1 = __fdiv_rn((float)mn1, (float)mx1);
  return v;
}

// One module of one impression: gather the history rows and mean-pool them (cr_module.py:107-123),
// then gather the candidate rows and dot them with the pooled user vector (click_predictors.py:12).
// The loops are branch-free on purpose: a batch always loads R rows (slots past the end re-read the
// batch's first row and are masked out), so the R*NV 16-byte loads stay in registers and are all in
// flight before the first use.  Returns MB200_FLAG_* bits.
//
// ATTN (early fusion, cr_module.py:124-125): when `logits` is given the history rows are combined with the
// additive-attention weights softmax(logits of the H history rows and of n_pad zero rows) instead of 1/H
// (attention.py:20-27; the reference's softmax runs over the PADDED history, so the pad rows take mass).
//
// HOT: rows whose id has a slot in the CTA's shared-memory hot-row cache (`slot_of[id] != kHotCold`; see score_eval_stream_kernel)
// are read from there with LDS.128 instead of through the L2 -> SM crossbar; the arithmetic is the same either way.
//
// MMA (bf16 rows, reference width, contiguous rows): the candidate dot products run on the tensor cores (mma.sync m16n8k16,
// fp32 accumulation) instead of 24 unpack + 24 FMA instructions per lane and row -- the bf16 kernel is issue-bound, not
// bandwidth-bound (profiles/r2_score_eval_bf16_ncu.json: 65 % issue-active at 41 % of the L2 roof).  The fp32 user vector is
// split exactly into three bf16 vectors u = hi + mid + lo (8 significand bits each) that occupy three of the eight MMA
// columns; 16 candidate rows are the M dimension.  Products of bf16 pairs are exact in fp32, sums are fp32: the same
// "bf16 storage, fp32 arithmetic" contract, in tensor-core summation order.
template <typename T, int NV, int R, bool EXACT, int POLICY, bool ATTN, bool HOT = false, bool MMA = false>
__device__ __forceinline__ int gather_pool_score(const T* __restrict__ table, long long row_stride, int vec_per_row, long long n_news,
                                                 const int32_t* __restrict__ hist_ids, int H, const int32_t* __restrict__ cand_ids, int C,
                                                 float* __restrict__ s_out, const float* __restrict__ logits, int n_pad,
                                                 const void* const* __restrict__ shard_base, int shard_shift,
                                                 const unsigned char* hot_rows = nullptr, const uint8_t* __restrict__ slot_of = nullptr,
                                                 unsigned char* upart = nullptr) {
  constexpr int E = Elem<T>::E;
  constexpr int kRowBytes = NV * 32 * 16;
  const int lane = threadIdx.x & 31;
  int flags = 0;
  // per-lane vector slots: slot v covers 16-byte vector lane + 32 v of the row; for widths that do not
  // fill the last slot the address is clamped and the slot's contribution multiplied by 0
  int voff[NV];
  float vmask[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const bool ok = EXACT || (lane + 32 * v < vec_per_row);
    voff[v] = ok ? lane + 32 * v : 0;
    vmask[v] = ok ? 1.0f : 0.0f;
  }

  float u[NV * E];
#pragma unroll
  for (int t = 0; t < NV * E; ++t) u[t] = 0.f;

  // softmax statistics of the attention logits: maximum and sum of exp over history + pad rows
  float att_max = 0.f, att_sum = 1.f;
  const bool attn = ATTN && logits != nullptr;
  if (ATTN && attn) {
    const float pad_logit = logits[n_news];
    float mx = n_pad > 0 ? pad_logit : -INFINITY;
#pragma unroll 1
    for (int h = lane; h < H; h += 32) {
      int id = hist_ids[h];
      if ((unsigned long long)(long long)id >= (unsigned long long)n_news) id = 0;
      mx = fmaxf(mx, logits[id]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    float se = 0.f;
#pragma unroll 1
    for (int h = lane; h < H; h += 32) {
      int id = hist_ids[h];
      if ((unsigned long long)(long long)id >= (unsigned long long)n_news) id = 0;
      se += expf(logits[id] - mx);
    }
    se = warp_sum(se);
    if (n_pad > 0) se = __fadd_rn(se, __fmul_rn((float)n_pad, expf(pad_logit - mx)));
    att_max = mx, att_sum = se;
  }

#pragma unroll 1
  for (int b0 = 0; b0 < H; b0 += 32) {
    const int cnt = min(32, H - b0);
    int my_id = hist_ids[b0 + ((lane < cnt) ? lane : 0)];
    if ((unsigned long long)(long long)my_id >= (unsigned long long)n_news) my_id = 0, flags |= MB200_FLAG_BAD_ID;
    float my_w = 1.0f;  // weight of this lane's history row: softmax weight (early fusion) or 1 (late fusion divides by H below)
    if (ATTN && attn) my_w = __fdiv_rn(expf(logits[my_id] - att_max), att_sum);
    if (HOT && slot_of != nullptr) {
      const int sl = slot_of[my_id];
      if (sl != kHotCold) my_id = ~sl;  // negative: slot of the hot-row cache
    }
#pragma unroll 1
    for (int r0 = 0; r0 < cnt; r0 += R) {
      uint4 buf[R][NV];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int src = (r0 + r < cnt) ? (r0 + r) : 0;
        const int id = __shfl_sync(kFull, my_id, src);
        if (HOT && id < 0) {
          const uint4* hrow = reinterpret_cast<const uint4*>(hot_rows + (unsigned)(~id) * (unsigned)kRowBytes) + lane;
#pragma unroll
          for (int v = 0; v < NV; ++v) buf[r][v] = hrow[32 * v];
          continue;
        }
        // POLICY 3: the table is row-sharded over the GPUs of the box; the row is read where it lives -- a plain load from the
        // owner's memory over NVLink / NVSwitch peer access when it is not this GPU's shard
        const T* tbase = POLICY == 3 ? reinterpret_cast<const T*>(shard_base[id >> shard_shift]) : table;
        const long long lrow = POLICY == 3 ? (long long)(id & ((1 << shard_shift) - 1)) : (long long)id;
        const uint4* row = reinterpret_cast<const uint4*>(tbase + lrow * row_stride);
#pragma unroll
        for (int v = 0; v < NV; ++v) buf[r][v] = load_row_vec<POLICY>(row + (EXACT ? lane + 32 * v : voff[v]));
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float keep = (r0 + r < cnt) ? 1.0f : 0.0f;
        if (ATTN) keep *= __shfl_sync(kFull, my_w, (r0 + r < cnt) ? (r0 + r) : 0);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float f[E];
          Elem<T>::unpack(buf[r][v], f);
          const float k = EXACT ? keep : keep * vmask[v];
#pragma unroll
          for (int e = 0; e < E; ++e) u[v * E + e] = fmaf(f[e], k, u[v * E + e]);  // late fusion: k is 1 or 0, an exact add or a no-op
        }
      }
    }
  }
  // true division by the history length, as torch.div(sum, hist_size) (cr_module.py:121-123).  The
  // correctly rounded quotient x / h is produced without the generic division subroutine: with
  // rh = RN(1/h), q = RN(x rh), r = x - q h (exact, one FMA), RN(q + r rh) is the IEEE quotient for
  // integer-valued h < 2^23 (Markstein; checked against true fp32 division on 15 M cases).
  if (!(ATTN && attn)) {
    const float hf = (float)H;
    const float rh = __frcp_rn(hf);
#pragma unroll
    for (int t = 0; t < NV * E; ++t) {
      const float q = __fmul_rn(u[t], rh);
      u[t] = __fmaf_rn(__fmaf_rn(-q, hf, u[t]), rh, q);
    }
  }
  if (!EXACT) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < E; ++e) u[v * E + e] = (vmask[v] != 0.0f) ? u[v * E + e] : 0.0f;
  }

  if constexpr (MMA) {
    static_assert(Elem<T>::E == 8 && NV == 3 && EXACT && POLICY == 0 && !HOT, "tensor-core candidate scoring: bf16 rows of the reference width");
    __builtin_assume(__isShared(upart));
    constexpr int kRow = 1536;  // bytes per row / per u part
    // u = hi + mid + lo, each a bf16 vector in the rows' own memory layout
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      uint4 part[3];
      unsigned* w[3] = {&part[0].x, &part[1].x, &part[2].x};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unsigned packed[3] = {0u, 0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float x = u[v * 8 + 2 * q + h];
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const __nv_bfloat16 b = __float2bfloat16_rn(x);
            packed[t] |= (unsigned)__bfloat16_as_ushort(b) << (16 * h);
            x = __fsub_rn(x, __bfloat162float(b));  // exact: the remainder has at most 16 significant bits
          }
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) w[t][q] = packed[t];
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4*>(upart + t * kRow + (256 * v + 8 * lane) * 2) = part[t];
    }
    __syncwarp();
    const int g = lane >> 2, j = lane & 3;
    const unsigned char* tbytes = reinterpret_cast<const unsigned char*>(table) + 16 * j;
    const unsigned char* ub = upart + (g < 3 ? g : 0) * kRow + 16 * j;
#pragma unroll 1
    for (int b0 = 0; b0 < C; b0 += 32) {
      const int cnt = min(32, C - b0);
      int my_id = cand_ids[b0 + ((lane < cnt) ? lane : 0)];
      if ((unsigned long long)(long long)my_id >= (unsigned long long)n_news) my_id = 0, flags |= MB200_FLAG_BAD_ID;
#pragma unroll 1
      for (int h0 = 0; h0 < cnt; h0 += 16) {
        const int ra = h0 + g, rb = h0 + g + 8;  // MMA rows g and g + 8 of this lane
        const int ida = __shfl_sync(kFull, my_id, ra < cnt ? ra : 0), idb = __shfl_sync(kFull, my_id, rb < cnt ? rb : 0);
        const unsigned char* pa = tbytes + (unsigned long long)(unsigned)ida * kRow;
        const unsigned char* pb = tbytes + (unsigned long long)(unsigned)idb * kRow;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
        // 6 chunks of 4 blocks of 32 dims (64 B per row and block: lanes j = 0..3 of a row read consecutive 16 B), double buffered
        uint4 xa[2][4], xb[2][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xa[0][k] = __ldg(reinterpret_cast<const uint4*>(pa + 64 * k)), xb[0][k] = __ldg(reinterpret_cast<const uint4*>(pb + 64 * k));
#pragma unroll
        for (int ch = 0; ch < 6; ++ch) {
          if (ch + 1 < 6) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              xa[(ch + 1) & 1][k] = __ldg(reinterpret_cast<const uint4*>(pa + 256 * (ch + 1) + 64 * k));
              xb[(ch + 1) & 1][k] = __ldg(reinterpret_cast<const uint4*>(pb + 256 * (ch + 1) + 64 * k));
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 uv = make_uint4(0u, 0u, 0u, 0u);
            if (g < 3) uv = *reinterpret_cast<const uint4*>(ub + 256 * ch + 64 * k);  // columns 0..2 = hi, mid, lo; 3..7 = 0
            const uint4 a = xa[ch & 1][k], b = xb[ch & 1][k];
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a.x), "r"(b.x), "r"(a.y), "r"(b.y), "r"(uv.x), "r"(uv.y));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a.z), "r"(b.z), "r"(a.w), "r"(b.w), "r"(uv.z), "r"(uv.w));
          }
        }
        // lane (g, j) holds columns 2j, 2j + 1 of rows g (c0, c1) and g + 8 (c2, c3): score = (hi + mid) + lo
        float sa = __fadd_rn(__fadd_rn(c[0], d[0]), __fadd_rn(c[1], d[1]));
        float sb = __fadd_rn(__fadd_rn(c[2], d[2]), __fadd_rn(c[3], d[3]));
        sa = __fadd_rn(sa, __shfl_xor_sync(kFull, sa, 1));
        sb = __fadd_rn(sb, __shfl_xor_sync(kFull, sb, 1));
        if (j == 0) {
          if (ra < cnt) s_out[b0 + ra] = sa;
          if (rb < cnt) s_out[b0 + rb] = sb;
        }
      }
    }
    return flags;
  }

#pragma unroll 1
  for (int b0 = 0; b0 < C; b0 += 32) {
    const int cnt = min(32, C - b0);
    int my_id = cand_ids[b0 + ((lane < cnt) ? lane : 0)];
    if ((unsigned long long)(long long)my_id >= (unsigned long long)n_news) my_id = 0, flags |= MB200_FLAG_BAD_ID;
    if (HOT && slot_of != nullptr) {
      const int sl = slot_of[my_id];
      if (sl != kHotCold) my_id = ~sl;
    }
#pragma unroll 1
    for (int r0 = 0; r0 < cnt; r0 += R) {
      uint4 buf[R][NV];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int src = (r0 + r < cnt) ? (r0 + r) : 0;
        const int id = __shfl_sync(kFull, my_id, src);
        if (HOT && id < 0) {
          const uint4* hrow = reinterpret_cast<const uint4*>(hot_rows + (unsigned)(~id) * (unsigned)kRowBytes) + lane;
#pragma unroll
          for (int v = 0; v < NV; ++v) buf[r][v] = hrow[32 * v];
          continue;
        }
        // POLICY 3: the table is row-sharded over the GPUs of the box; the row is read where it lives -- a plain load from the
        // owner's memory over NVLink / NVSwitch peer access when it is not this GPU's shard
        const T* tbase = POLICY == 3 ? reinterpret_cast<const T*>(shard_base[id >> shard_shift]) : table;
        const long long lrow = POLICY == 3 ? (long long)(id & ((1 << shard_shift) - 1)) : (long long)id;
        const uint4* row = reinterpret_cast<const uint4*>(tbase + lrow * row_stride);
#pragma unroll
        for (int v = 0; v < NV; ++v) buf[r][v] = load_row_vec<POLICY>(row + (EXACT ? lane + 32 * v : voff[v]));
      }
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float part = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float f[E];
          Elem<T>::unpack(buf[r][v], f);
#pragma unroll
          for (int e = 0; e < E; ++e) part = fmaf(u[v * E + e], f[e], part);  // u is 0 in masked slots
        }
        part = warp_sum(part);
        mine = (lane == r) ? part : mine;
      }
      if (lane < R && r0 + lane < cnt) s_out[b0 + r0 + lane] = mine;
    }
  }
  return flags;
}

struct WarpSmem {
  double* acc;        // [W][NUM_METRICS] fp64 sums of this warp
  float* sc;          // [n_active][cpad] per-module scores of the current impression
  float* comb;        // [cpad] combined scores of one weighting
  uint8_t* lab;       // [cpad] labels
  uint8_t* ccat;      // [cpad] candidate category labels
  uint8_t* csent;     // [cpad] candidate sentiment labels
  int* hist_cat;      // [MB200_MAX_CLASSES] history category histogram
  int* hist_sent;     // [MB200_MAX_CLASSES]
  uint8_t* top_cat;   // [32] category of the candidate at rank r
  uint8_t* top_sent;  // [32]
};

constexpr int kSweepMinWeightings = 16;  // from here on lanes own weightings instead of candidates
constexpr int kSweepSlots = MB200_M_GAUC_VALID + 1;  // metric slots the sweep mode fills: its accumulators are packed to these,
                                                      // which keeps 121 weightings at 4 resident CTAs per SM instead of 2
constexpr int kSweepPos = 2;             // positives ranked per pass over the candidates (91 % of MIND impressions have <= 2)

// s = w0 * z0 ; s += w_m * z_m for m >= 1 when w_m != 0  (ensemble_module.py:97-107): separate multiply
// and add in fp32 -- no fused multiply-add -- like the reference's two torch ops.
__device__ __forceinline__ float combine_at(const float* sc, int cpad, int j, const float (&wt)[MB200_MAX_MODULES], int n_modules,
                                            int active_mask) {
  float s = (wt[0] == 1.0f) ? sc[j] : __fmul_rn(wt[0], sc[j]);
  int sl = 1;
#pragma unroll
  for (int m = 1; m < MB200_MAX_MODULES; ++m) {
    if (m < n_modules && ((active_mask >> m) & 1)) {
      if (wt[m] != 0.0f) s = __fadd_rn(s, __fmul_rn(wt[m], sc[(size_t)sl * cpad + j]));
      ++sl;
    }
  }
  return s;
}

// Aspect-weight sweep (BASELINE.json configs[3]): the z-scored per-module scores of the impression are
// already in shared memory; every lane re-scores the impression under its own weightings (w = lane,
// lane + 32, ...) and ranks only the positives, reading the module scores as warp-wide broadcasts.
// W weightings cost one gather.
__device__ __noinline__ int sweep_weightings(const EvalParams& p, const WarpSmem& sm, int i, int c0, int C) {
  const int lane = threadIdx.x & 31;
  const int W = p.n_weightings;
  const float* sc = sm.sc;
  const uint8_t* lab = sm.lab;
  int* pos_list = reinterpret_cast<int*>(sm.comb);  // the combined-score buffer is not needed in this mode
  __builtin_assume(__isShared(sc));
  __builtin_assume(__isShared(lab));
  __builtin_assume(__isShared(pos_list));
  __builtin_assume(__isShared(sm.acc));
  int flags = 0;

  int n_pos = 0;
#pragma unroll 1
  for (int b0 = 0; b0 < C; b0 += 32) {
    const int j = b0 + lane;
    const bool is_pos = j < C && lab[j] != 0;
    const unsigned m = __ballot_sync(kFull, is_pos);
    if (is_pos) pos_list[n_pos + __popc(m & ((1u << lane) - 1u))] = j;
    n_pos += __popc(m);
  }
  __syncwarp();

  if (p.scores != nullptr) {
    float wt[MB200_MAX_MODULES];
#pragma unroll
    for (int m = 0; m < MB200_MAX_MODULES; ++m) wt[m] = (m < p.n_modules) ? p.weights[(size_t)p.scores_weighting * p.n_modules + m] : 0.0f;
    bool outside = false;
#pragma unroll 1
    for (int j = lane; j < C; j += 32) {
      const float s = combine_at(sc, p.cpad, j, wt, p.n_modules, p.active_mask);
      p.scores[c0 + j] = s;
      outside |= !(s >= 0.0f && s <= 1.0f);
    }
    if (outside) flags |= MB200_FLAG_OUTSIDE_UNIT;
  }

#pragma unroll 1
  for (int w = lane; w < W; w += 32) {
    float wt[MB200_MAX_MODULES];
#pragma unroll
    for (int m = 0; m < MB200_MAX_MODULES; ++m) wt[m] = (m < p.n_modules) ? p.weights[(size_t)w * p.n_modules + m] : 0.0f;
    int min_rank = 0x7fffffff;
    unsigned hit_mask = 0;
    long long gauc2 = 0;
#pragma unroll 1
    for (int p0 = 0; p0 < n_pos; p0 += kSweepPos) {
      int pj[kSweepPos], before[kSweepPos], nlt[kSweepPos], neq[kSweepPos];
      float sp[kSweepPos];
#pragma unroll
      for (int q = 0; q < kSweepPos; ++q) {
        pj[q] = (p0 + q < n_pos) ? pos_list[p0 + q] : -1;
        sp[q] = (pj[q] >= 0) ? combine_at(sc, p.cpad, pj[q], wt, p.n_modules, p.active_mask) : 0.0f;
        before[q] = nlt[q] = neq[q] = 0;
      }
#pragma unroll 4
      for (int k = 0; k < C; ++k) {
        const float sk = combine_at(sc, p.cpad, k, wt, p.n_modules, p.active_mask);
        const bool neg = lab[k] == 0;
#pragma unroll
        for (int q = 0; q < kSweepPos; ++q) {
          before[q] += (ranks_before(sk, sp[q]) || (k < pj[q] && ranks_equal(sk, sp[q]))) ? 1 : 0;
          nlt[q] += (neg && sk < sp[q]) ? 1 : 0;
          neq[q] += (neg && sk == sp[q]) ? 1 : 0;
        }
      }
#pragma unroll
      for (int q = 0; q < kSweepPos; ++q) {
        if (pj[q] >= 0) {
          const int rank = 1 + before[q];
          min_rank = min(min_rank, rank);
          if (rank <= 32) hit_mask |= 1u << (rank - 1);
          gauc2 += 2ll * nlt[q] + neq[q];
        }
      }
    }
    const float mrr = n_pos ? __fdiv_rn(1.0f, (float)min_rank) : 0.f;
    const float nd0 = ndcg_at(hit_mask, n_pos, C, p.k0);
    const float nd1 = ndcg_at(hit_mask, n_pos, C, p.k1);
    const bool gvalid = n_pos > 0 && n_pos < C;
    const float g = gvalid ? (float)((double)gauc2 / (2.0 * (double)n_pos * (double)(C - n_pos))) : 0.f;
    double* a = sm.acc + (size_t)w * p.acc_stride;
    a[MB200_M_MRR] += (double)mrr, a[MB200_M_NDCG_K0] += (double)nd0, a[MB200_M_NDCG_K1] += (double)nd1;
    a[MB200_M_GAUC] += (double)g, a[MB200_M_GAUC_VALID] += gvalid ? 1.0 : 0.0;
    if (p.per_impr) {
      float* o = p.per_impr + ((size_t)w * p.n_impr + i) * MB200_NUM_METRICS;
      o[MB200_M_MRR] = mrr, o[MB200_M_NDCG_K0] = nd0, o[MB200_M_NDCG_K1] = nd1, o[MB200_M_GAUC] = g, o[MB200_M_GAUC_VALID] = gvalid ? 1.f : 0.f;
#pragma unroll 1
      for (int t = MB200_M_GAUC_VALID + 1; t < MB200_NUM_METRICS; ++t) o[t] = 0.f;
    }
  }
  __syncwarp();
  return flags;
}

// Loss of one impression on its final scores (cr_module.py:140-171), fp32 terms, fp64 accumulation.
//   cross entropy: -sum over positives of log_softmax(s)_p, the softmax running over the C real columns and the
//                  n_pad zero columns the reference's dense batch appends (torch CrossEntropyLoss with the 0/1 label
//                  matrix as class probabilities; the batch mean is taken by mb200_step_loss)
//   SupCon:        -(sum over positives of (s_p / T - logsumexp over the real candidates of s / T)) / P, 0 when P = 0
//                  (components/losses.py:19-40)
__device__ __noinline__ float impression_loss(const EvalParams& p, const float* s, const uint8_t* lab, int C, int i, int lane) {
  __builtin_assume(__isShared(s));
  __builtin_assume(__isShared(lab));
  const bool ce = p.loss_kind == MB200_LOSS_CE;
  // zero columns the step's dense batch appends: part of the cross entropy's softmax; for SupCon only of the row maximum that
  // is subtracted for stability (losses.py:24-25 takes mat.max over the dense row, the log-sum-exp keeps pos + neg only)
  const int n_pad = p.cand_pad ? p.cand_pad[i] : 0;
  const float T = ce ? 1.0f : p.loss_temperature;
  float mx = n_pad > 0 ? 0.f : -INFINITY;
#pragma unroll 1
  for (int j = lane; j < C; j += 32) mx = fmaxf(mx, ce ? s[j] : __fdiv_rn(s[j], T));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
  double se = 0.0, sp = 0.0;
  int n_pos = 0;
#pragma unroll 1
  for (int j = lane; j < C; j += 32) {
    const float x = __fsub_rn(ce ? s[j] : __fdiv_rn(s[j], T), mx);
    se += (double)expf(x);
    if (lab[j] != 0) sp += (double)x, ++n_pos;
  }
  se = warp_sum(se), sp = warp_sum(sp);
  n_pos = __reduce_add_sync(kFull, n_pos);
  if (ce && n_pad > 0) se += (double)n_pad * (double)expf(-mx);
  float loss = 0.f;
  if (n_pos > 0) {
    const double lse = (double)logf((float)se);
    const double total = (double)n_pos * lse - sp;  // -sum over positives of (x_p - lse)
    loss = ce ? (float)total : (float)(total / (double)n_pos);
  }
  if (p.loss_per_impr && lane == 0) p.loss_per_impr[i] = loss;
  return loss;
}

// Everything after the per-module scores of impression i sit in shared memory: z-scores are already
// applied; combine per weighting, rank, metrics, accumulate.  Not inlined (see gather_pool_score).
__device__ __noinline__ int rank_and_metrics(const EvalParams& p, const WarpSmem& sm, int i, int h0, int H, int c0, int C) {
  const int lane = threadIdx.x & 31;
  const int W = p.n_weightings;
  const int n_active = __popc(p.active_mask);
  const int kmax = max(p.k0, p.k1);
  const bool flat = p.flat_cand_aspect[0] != nullptr;  // labels per row (mb200_rank_metrics) instead of per news id
  const bool aspects = p.news_category != nullptr || flat;
  float* sc = sm.sc;
  float* comb = sm.comb;
  const uint8_t* lab = sm.lab;
  __builtin_assume(__isShared(sc));
  __builtin_assume(__isShared(comb));
  __builtin_assume(__isShared(lab));
  __builtin_assume(__isShared(sm.acc));
  __builtin_assume(__isShared(sm.ccat));
  __builtin_assume(__isShared(sm.csent));
  __builtin_assume(__isShared(sm.hist_cat));
  __builtin_assume(__isShared(sm.hist_sent));
  __builtin_assume(__isShared(sm.top_cat));
  __builtin_assume(__isShared(sm.top_sent));
  int warp_flags = 0;

  // ---- aspects of this impression (once, independent of the weighting) -----------------------------
  bool categ_group_ok = false, sent_group_ok = false;
  if (aspects) {
#pragma unroll 1
    for (int t = lane; t < MB200_MAX_CLASSES; t += 32) sm.hist_cat[t] = 0, sm.hist_sent[t] = 0;
    __syncwarp();
    int cat_sum = 0, sent_sum = 0;
#pragma unroll 1
    for (int j = lane; j < C; j += 32) {
      int a, b;
      if (flat) {
        a = p.flat_cand_aspect[0][c0 + j], b = p.flat_cand_aspect[1][c0 + j];
      } else {
        int id = p.cand_ids[c0 + j];
        if ((unsigned long long)(long long)id >= (unsigned long long)p.n_news) id = 0;
        a = p.news_category[id], b = p.news_sentiment[id];
      }
      if ((unsigned)a >= (unsigned)p.num_categ) a = 0, warp_flags |= MB200_FLAG_BAD_ASPECT;
      if ((unsigned)b >= (unsigned)p.num_sent) b = 0, warp_flags |= MB200_FLAG_BAD_ASPECT;
      sm.ccat[j] = (uint8_t)a, sm.csent[j] = (uint8_t)b;
      cat_sum += a, sent_sum += b;
    }
#pragma unroll 1
    for (int h = lane; h < H; h += 32) {
      int a, b;
      if (flat) {
        a = p.flat_hist_aspect[0][h0 + h], b = p.flat_hist_aspect[1][h0 + h];
      } else {
        int id = p.hist_ids[h0 + h];
        if ((unsigned long long)(long long)id >= (unsigned long long)p.n_news) id = 0;
        a = p.news_category[id], b = p.news_sentiment[id];
      }
      if ((unsigned)a >= (unsigned)p.num_categ) a = 0, warp_flags |= MB200_FLAG_BAD_ASPECT;
      if ((unsigned)b >= (unsigned)p.num_sent) b = 0, warp_flags |= MB200_FLAG_BAD_ASPECT;
      atomicAdd(&sm.hist_cat[a], 1);
      atomicAdd(&sm.hist_sent[b], 1);
    }
    // `if not mini_target.sum()` -> 0.0 (metrics/base.py:114-122 and torchmetrics RetrievalMetric.compute)
    categ_group_ok = __reduce_add_sync(kFull, cat_sum) != 0;
    sent_group_ok = __reduce_add_sync(kFull, sent_sum) != 0;
    __syncwarp();
  }

  if (!aspects && p.weights != nullptr && W >= kSweepMinWeightings) return warp_flags | sweep_weightings(p, sm, i, c0, C);

#pragma unroll 1
  for (int w = 0; w < W; ++w) {
    const float* scores_w = sc;
    if (p.weights != nullptr || n_active > 1) {
      // s = w0 * z0 ; s += w_m * z_m for m >= 1 when w_m != 0  (ensemble_module.py:97-107; separate
      // multiply and add in fp32, no fused multiply-add, like the reference's two torch ops)
      float wt[MB200_MAX_MODULES];
#pragma unroll
      for (int m = 0; m < MB200_MAX_MODULES; ++m)
        wt[m] = (m < p.n_modules) ? (p.weights ? p.weights[(size_t)w * p.n_modules + m] : 1.0f) : 0.0f;
#pragma unroll 1
      for (int j = lane; j < C; j += 32) {
        float s = (wt[0] == 1.0f) ? sc[j] : __fmul_rn(wt[0], sc[j]);
        int sl = 1;
#pragma unroll
        for (int m = 1; m < MB200_MAX_MODULES; ++m) {
          if (m < p.n_modules && ((p.active_mask >> m) & 1)) {
            if (wt[m] != 0.0f) s = __fadd_rn(s, __fmul_rn(wt[m], sc[(size_t)sl * p.cpad + j]));
            ++sl;
          }
        }
        comb[j] = s;
      }
      __syncwarp();
      scores_w = comb;
    }

    if (p.scores != nullptr && w == p.scores_weighting) {
      bool outside = false;
#pragma unroll 1
      for (int j = lane; j < C; j += 32) {
        const float s = scores_w[j];
        p.scores[c0 + j] = s;
        outside |= !(s >= 0.0f && s <= 1.0f);
      }
      if (outside) warp_flags |= MB200_FLAG_OUTSIDE_UNIT;
    }

    float loss = 0.f;
    if (p.loss_kind != MB200_LOSS_NONE && w == p.scores_weighting) loss = impression_loss(p, scores_w, lab, C, i, lane);

    int n_pos = 0, min_rank = 0x7fffffff;
    unsigned hit_mask = 0;
    long long gauc2 = 0;  // sum over positives of 2 * (#neg below) + (#neg equal)

    // the positives' ranks: lanes split the comparison partners (all the ranking metrics need)
#pragma unroll 1
    for (int b0 = 0; b0 < C; b0 += 32) {
      const int j = b0 + lane;
      unsigned pm = __ballot_sync(kFull, j < C && lab[j] != 0);
      n_pos += __popc(pm);
#pragma unroll 1
      while (pm) {
        const int pj = b0 + __ffs(pm) - 1;
        pm &= pm - 1;
        const float sp = scores_w[pj];
        int before = 0, nlt = 0, neq = 0;
#pragma unroll 1
        for (int k = lane; k < C; k += 32) {
          const float sk = scores_w[k];
          before += (ranks_before(sk, sp) || (k < pj && ranks_equal(sk, sp))) ? 1 : 0;
          if (lab[k] == 0) nlt += (sk < sp) ? 1 : 0, neq += (sk == sp) ? 1 : 0;
        }
        const int rank = 1 + __reduce_add_sync(kFull, before);
        gauc2 += 2ll * __reduce_add_sync(kFull, nlt) + __reduce_add_sync(kFull, neq);
        min_rank = min(min_rank, rank);
        if (rank <= 32) hit_mask |= 1u << (rank - 1);
      }
    }
    if (aspects) {
      // Identity of the top-kmax candidates (Diversity / Personalization look at their aspect labels): kmax rounds of a warp-wide
      // arg-max over the candidates that rank after the previous pick -- score descending, lower position first among equals, NaN
      // above everything and equal to itself (torch.sort).  O(kmax C / 32) instead of ranking every candidate against every other
      // (O(C^2 / 32): 27 % of the kernel's instructions with aspect metrics on, profiles/r2_ac_score_eval_aspects_ncu.json).
      const int kk = min(kmax, C);
      unsigned prev_key = 0xffffffffu;
      int prev_idx = -1;
#pragma unroll 1
      for (int r = 0; r < kk; ++r) {
        unsigned best_key = 0u;
        int best_idx = 0x7fffffff;
#pragma unroll 1
        for (int j = lane; j < C; j += 32) {
          const unsigned key = rank_key(scores_w[j]);
          const bool after = key < prev_key || (key == prev_key && j > prev_idx);  // ranks strictly after the previous pick
          if (after && (key > best_key || best_idx == 0x7fffffff)) best_key = key, best_idx = j;  // ascending j: the first of equals stays
        }
        const unsigned top = __reduce_max_sync(kFull, best_idx == 0x7fffffff ? 0u : best_key);
        const int pick = __reduce_min_sync(kFull, (best_idx != 0x7fffffff && best_key == top) ? best_idx : 0x7fffffff);
        if (lane == 0) sm.top_cat[r] = sm.ccat[pick], sm.top_sent[r] = sm.csent[pick];
        prev_key = top, prev_idx = pick;
      }
      __syncwarp();
    }

    // lane t computes metric slot t (uniform inputs), so there is no per-lane register array
    float mine = 0.f;
    // RetrievalMRR: 1 / (first hit position + 1); impressions without a positive count as 0
    if (lane == MB200_M_MRR) mine = n_pos ? __fdiv_rn(1.0f, (float)min_rank) : 0.f;
    if (lane == MB200_M_NDCG_K0) mine = ndcg_at(hit_mask, n_pos, C, p.k0);
    if (lane == MB200_M_NDCG_K1) mine = ndcg_at(hit_mask, n_pos, C, p.k1);
    if (n_pos > 0 && n_pos < C) {
      if (lane == MB200_M_GAUC) mine = (float)((double)gauc2 / (2.0 * (double)n_pos * (double)(C - n_pos)));
      if (lane == MB200_M_GAUC_VALID) mine = 1.f;
    }
    if (lane == MB200_M_LOSS) mine = loss;
    if (lane == MB200_M_LOSS_NONZERO) mine = loss > 0.f ? 1.f : 0.f;
    if (aspects) {
      const int kk0 = min(p.k0, C), kk1 = min(p.k1, C);
      if (categ_group_ok) {
        const AspectValues v = aspect_values(sm.top_cat, kk0, kk1, sm.hist_cat, p.num_categ, lane);
        if (lane == MB200_M_CATEG_DIV_K0) mine = v.div0;
        if (lane == MB200_M_CATEG_DIV_K1) mine = v.div1;
        if (lane == MB200_M_CATEG_PERS_K0) mine = v.pers0;
        if (lane == MB200_M_CATEG_PERS_K1) mine = v.pers1;
      }
      if (sent_group_ok) {
        const AspectValues v = aspect_values(sm.top_sent, kk0, kk1, sm.hist_sent, p.num_sent, lane);
        if (lane == MB200_M_SENT_DIV_K0) mine = v.div0;
        if (lane == MB200_M_SENT_DIV_K1) mine = v.div1;
        if (lane == MB200_M_SENT_PERS_K0) mine = v.pers0;
        if (lane == MB200_M_SENT_PERS_K1) mine = v.pers1;
      }
      __syncwarp();
    }
    if (lane < MB200_NUM_METRICS) {
      sm.acc[w * p.acc_stride + lane] += (double)mine;
      if (p.per_impr) p.per_impr[((size_t)w * p.n_impr + i) * MB200_NUM_METRICS + lane] = mine;
    }
    __syncwarp();
  }
  return warp_flags;
}

// Pipelined upload: wait (lane 0 polls, bounded to 4 s) until the copy stream has raised *ready to at least `need`; returns the
// value seen.  The poll itself is a relaxed system-scope load -- an acquire load would invalidate the SM's L1 (CCTL.IVALL) on
// every iteration, under the feet of the 15 other warps -- and ONE acquire fence follows once the word is high enough: the ids
// and labels behind it were written by the copy engine before the word.
__device__ __noinline__ unsigned wait_ready(const unsigned int* ready, unsigned need) {
  unsigned v = 0;
  if ((threadIdx.x & 31) == 0) {
    unsigned long long t0 = 0;
    while (true) {
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ready) : "memory");
      if (v >= need) break;
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) break;
      // thousands of warps may be waiting on this one word: poll slowly enough that they do not saturate its L2 slice
      // (16 polling warps per SM at 400 ns stretched the copies themselves)
      __nanosleep(now - t0 < 20000ull ? 1000 : 4000);
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  v = __shfl_sync(kFull, v, 0);
  __syncwarp();
  return v;
}

template <typename T, int NV, int R, bool EXACT, int POLICY, int MINB, bool ATTN = false, bool MMA = false>
__global__ void __launch_bounds__(kThreads, MINB) score_eval_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kWarpsPerCta + warp;
  const int total_warps = gridDim.x * kWarpsPerCta;
  const int W = p.n_weightings;
  const int n_active = __popc(p.active_mask);

  unsigned char* base = smem + (size_t)warp * p.smem_per_warp;
  WarpSmem sm;
  sm.acc = reinterpret_cast<double*>(base);
  sm.sc = reinterpret_cast<float*>(base + p.acc_bytes);
  sm.comb = sm.sc + (size_t)n_active * p.cpad;
  sm.lab = reinterpret_cast<uint8_t*>(sm.comb + p.cpad);
  sm.ccat = sm.lab + p.cpad;
  sm.csent = sm.ccat + p.cpad;
  sm.hist_cat = reinterpret_cast<int*>(sm.csent + p.cpad);
  sm.hist_sent = sm.hist_cat + MB200_MAX_CLASSES;
  sm.top_cat = reinterpret_cast<uint8_t*>(sm.hist_sent + MB200_MAX_CLASSES);
  sm.top_sent = sm.top_cat + 32;

  for (int t = lane; t < W * p.acc_stride; t += 32) sm.acc[t] = 0.0;
  __syncwarp();

  // Chunks are contiguous impression ranges balanced by rows gathered (partition_kernel).  Behaviours resident: one chunk per
  // warp.  Pipelined upload: 16 small chunks per warp handed out from a counter in impression order, so the grid consumes the
  // set in the order its segments arrive and no warp sits on a late range.  Every chunk's sums go to the chunk's own slot of
  // `partials`: the reduction order -- every bit of the fp64 sums -- does not depend on which warp ran which chunk.
  int warp_flags = 0;
  unsigned ready_seen = p.ready ? 0u : 0xffffffffu;
  int chunk = gw - total_warps;
  while (true) {
    if (p.chunk_counter != nullptr) {
      if (lane == 0) chunk = atomicAdd(p.chunk_counter, 1);
      chunk = __shfl_sync(kFull, chunk, 0);
    } else {
      chunk += total_warps;  // static: chunk gw, gw + total_warps, ...
    }
    if (chunk >= p.n_chunks) break;
    const int i_begin = p.bounds[chunk], i_end = p.bounds[chunk + 1];
    if (ready_seen < (unsigned)i_end) {
      // pipelined upload: the ids / labels of this chunk may still be on their way (upload.cu)
      ready_seen = wait_ready(p.ready, (unsigned)i_end);
      if (ready_seen < (unsigned)i_end) {
        warp_flags |= MB200_FLAG_UPLOAD_TIMEOUT;
        break;
      }
    }
    for (int i = i_begin; i < i_end; ++i) {
      const int h0 = p.hist_offsets[i], h1 = p.hist_offsets[i + 1];
      const int c0 = p.cand_offsets[i], c1 = p.cand_offsets[i + 1];
      const int H = h1 - h0, C = c1 - c0;
      if (C > p.max_cand || C <= 0 || H < 0) {
        warp_flags |= MB200_FLAG_CAND_OVERFLOW;
        if (p.per_impr)
          for (int t = lane; t < W * MB200_NUM_METRICS; t += 32)
            p.per_impr[((size_t)(t / MB200_NUM_METRICS) * p.n_impr + i) * MB200_NUM_METRICS + t % MB200_NUM_METRICS] = 0.f;
        continue;
      }
      __syncwarp();
      for (int j = lane; j < C; j += 32) sm.lab[j] = p.labels[c0 + j];

      int slot = 0;
      for (int m = 0; m < p.n_modules; ++m) {
        if (!((p.active_mask >> m) & 1)) continue;
        float* s_m = sm.sc + (size_t)slot * p.cpad;
        warp_flags |= gather_pool_score<T, NV, R, EXACT, POLICY, ATTN, false, MMA>(reinterpret_cast<const T*>(p.tables[m]), p.row_stride, p.vec_per_row,
                                                                                   p.n_news, p.hist_ids + h0, H, p.cand_ids + c0, C, s_m,
                                                                                   ATTN ? p.attn_logits[m] : nullptr,
                                                                                   (ATTN && p.hist_pad) ? p.hist_pad[i] : 0, p.shard_base[m], p.shard_shift,
                                                                                   nullptr, nullptr, MMA ? base + p.upart_off : nullptr);
        __syncwarp();
        if (p.zscore) {
          zscore_inplace(s_m, C, lane);
          __syncwarp();
        }
        ++slot;
      }
      warp_flags |= rank_and_metrics(p, sm, i, h0, H, c0, C);
    }
    __syncwarp();
    for (int t = lane; t < W * MB200_NUM_METRICS; t += 32) {
      const int w = t / MB200_NUM_METRICS, k = t % MB200_NUM_METRICS;
      p.partials[(size_t)t * p.n_partials + chunk] = k < p.acc_stride ? sm.acc[w * p.acc_stride + k] : 0.0;
    }
    __syncwarp();
    for (int t = lane; t < W * p.acc_stride; t += 32) sm.acc[t] = 0.0;
    __syncwarp();
  }

  warp_flags = __reduce_or_sync(kFull, warp_flags);
  if (lane == 0 && warp_flags && p.flags) atomicOr(p.flags, warp_flags);
}

// ------------------------------------------------------------------------------------------------------
// Streaming variant for the reference width (dim 768), late fusion, replicated tables: the product path.
//
// (1) Hot-row cache.  News popularity is heavy-tailed (a few dozen rows carry a quarter of all row reads of a
//     MIND-shaped behaviour set), but the L1 cannot hold on to them: every SM streams ~180 MB through it per launch.  One
//     CTA of 16 warps per SM therefore keeps the most frequently gathered rows of THIS behaviour set (found per launch by
//     hot_count_kernel / hot_select_kernel) in shared memory -- everything the per-warp score areas leave of the 227 KB --
//     and a gathered row is read from there (LDS.128) instead of through the L2 -> SM crossbar, which is the binding
//     resource on Zipf-shaped ids.  The arithmetic does not depend on where a row is read from:
//     outputs are bit-identical with the cache on or off.
// (2) Rotating row pipeline.  History and candidate rows of a module form ONE stream of slots; R rows are in flight per warp
//     and a row's registers are refilled with the row R slots ahead as soon as it is consumed, so candidate rows are already
//     arriving while the user vector is finished and there is no issue-all / wait-all bubble per batch.  Ids (and their
//     cache slots) are fetched 32 at a time one chunk ahead of the rows that need them.
// ------------------------------------------------------------------------------------------------------

constexpr int kStreamWarps = 16;

struct StreamSrc {
  const int32_t* hist_ids;  // of this impression
  const int32_t* cand_ids;
  int H, Hp, C, n_slots;    // Hp = H rounded up to R: candidate slots start R-aligned; slots [H, Hp) are empty
  long long n_news;
  const uint8_t* slot_of;   // null = cache off
};

// ref of stream slot base + lane: news id (>= 0), ~cache slot (< 0), or 0 for an empty slot; OR-s MB200_FLAG_BAD_ID into flags
__device__ __noinline__ int stream_fetch_refs(const StreamSrc& s, int base, int& flags) {
  const int t = base + (threadIdx.x & 31);
  int id = 0;
  bool valid = false;
  if (t < s.H) id = s.hist_ids[t], valid = true;
  else if (t >= s.Hp && t < s.n_slots) id = s.cand_ids[t - s.Hp], valid = true;
  if ((unsigned long long)(long long)id >= (unsigned long long)s.n_news) id = 0, flags |= MB200_FLAG_BAD_ID;
  int ref = id;
  if (s.slot_of != nullptr && valid) {
    const int sl = s.slot_of[id];
    if (sl != kHotCold) ref = ~sl;
  }
  return ref;
}

template <typename T, int NV, int R>
__device__ __forceinline__ int stream_pool_score(const T* __restrict__ table, const unsigned char* hot_rows, const StreamSrc& src,
                                                 int ref_first, float* __restrict__ s_out) {
  constexpr int E = Elem<T>::E;
  constexpr int kRowBytes = NV * 32 * 16;  // rows are contiguous (row stride == dim): a row's address is one multiply-add
  const int lane = threadIdx.x & 31;
  int flags = 0;
  const int H = src.H, Hp = src.Hp, n_slots = src.n_slots;
  const unsigned char* gbase = reinterpret_cast<const unsigned char*>(table) + lane * 16;  // this lane's first vector of row 0
  const unsigned char* sbase = hot_rows + lane * 16;

  float u[NV * E];
#pragma unroll
  for (int t = 0; t < NV * E; ++t) u[t] = 0.f;

  uint4 buf[R][NV];
  int ref_cur = 0, ref_nxt = ref_first;

  // issue the loads of slot sn (if it holds a row) into b; advances the id window when sn enters a new chunk of 32
  auto refill = [&](uint4 (&b)[NV], int sn) {
    if (sn < n_slots) {
      if ((sn & 31) == 0) {
        ref_cur = ref_nxt;
        if (sn + 32 < n_slots) ref_nxt = stream_fetch_refs(src, sn + 32, flags);
      }
      if (sn < H || sn >= Hp) {
        const int ref = __shfl_sync(kFull, ref_cur, sn & 31);
        if (ref < 0) {
          const uint4* row = reinterpret_cast<const uint4*>(sbase + (unsigned)(~ref) * (unsigned)kRowBytes);
#pragma unroll
          for (int v = 0; v < NV; ++v) b[v] = row[32 * v];
        } else {
          const uint4* row = reinterpret_cast<const uint4*>(gbase + (unsigned long long)(unsigned)ref * (unsigned long long)kRowBytes);
#pragma unroll
          for (int v = 0; v < NV; ++v) b[v] = __ldg(row + 32 * v);
        }
      }
    }
  };

#pragma unroll
  for (int r = 0; r < R; ++r) refill(buf[r], r);

  // history: u = sum of the rows (cr_module.py:116-123)
#pragma unroll 1
  for (int t0 = 0; t0 < Hp; t0 += R) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < H) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float f[E];
          Elem<T>::unpack(buf[r][v], f);
#pragma unroll
          for (int e = 0; e < E; ++e) u[v * E + e] = __fadd_rn(u[v * E + e], f[e]);
        }
      }
      refill(buf[r], t + R);
    }
  }
  // true division by the history length (see gather_pool_score)
  {
    const float hf = (float)H;
    const float rh = __frcp_rn(hf);
#pragma unroll
    for (int t = 0; t < NV * E; ++t) {
      const float q = __fmul_rn(u[t], rh);
      u[t] = __fmaf_rn(__fmaf_rn(-q, hf, u[t]), rh, q);
    }
  }
  // candidates: s_j = u . row_j (click_predictors.py:12)
#pragma unroll 1
  for (int t0 = Hp; t0 < n_slots; t0 += R) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + r;
      if (t < n_slots) {
        float part = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float f[E];
          Elem<T>::unpack(buf[r][v], f);
#pragma unroll
          for (int e = 0; e < E; ++e) part = fmaf(u[v * E + e], f[e], part);
        }
        part = warp_sum(part);
        if (lane == 0) s_out[t - Hp] = part;
        refill(buf[r], t + R);
      }
    }
  }
  return flags;
}

__device__ __forceinline__ WarpSmem warp_smem_of(unsigned char* base, const EvalParams& p, int n_active) {
  WarpSmem sm;
  sm.acc = reinterpret_cast<double*>(base);
  sm.sc = reinterpret_cast<float*>(base + p.acc_bytes);
  sm.comb = p.comb_alias ? sm.sc : sm.sc + (size_t)n_active * p.cpad;
  sm.lab = reinterpret_cast<uint8_t*>(sm.sc + (size_t)(n_active + (p.comb_alias ? 0 : 1)) * p.cpad);
  sm.ccat = sm.lab + p.cpad;
  sm.csent = sm.ccat + p.cpad;
  sm.hist_cat = reinterpret_cast<int*>(sm.csent + p.cpad);  // cpad is a multiple of 32: 4-byte aligned
  sm.hist_sent = sm.hist_cat + MB200_MAX_CLASSES;
  sm.top_cat = reinterpret_cast<uint8_t*>(sm.hist_sent + MB200_MAX_CLASSES);
  sm.top_sent = sm.top_cat + 32;
  if (!p.has_aspects) sm.ccat = sm.csent = sm.top_cat = sm.top_sent = sm.lab, sm.hist_cat = sm.hist_sent = reinterpret_cast<int*>(sm.sc);  // never touched
  return sm;
}

// PIPE 0: register batches (gather_pool_score: R rows issued, then consumed) -- the shipped path; PIPE 1: rotating row pipeline
// (stream_pool_score) -- measured slower (profiles/r2_stream_experiment.md: per-row refills defeat the load scoreboards), kept
// selectable for the record.
template <typename T, int NV, int R, bool HOT, int PIPE>
__global__ void __launch_bounds__(kStreamWarps * 32, 1) score_eval_stream_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kVecPerRow = NV * 32;
  constexpr int kRowBytes = kVecPerRow * 16;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kStreamWarps + warp;
  const int total_warps = gridDim.x * kStreamWarps;
  const int W = p.n_weightings;
  const int n_active = __popc(p.active_mask);

  // ---- fill the hot-row cache: slot s of active module a at smem[(a * hot_cap + s) * row bytes] --------------------
  int n_hot = 0;
  if (HOT) {
    n_hot = min(p.hot_dir->n_hot, p.hot_cap);
    const int per_module = n_hot * kVecPerRow;
    int a = 0;
    for (int m = 0; m < p.n_modules; ++m) {
      if (!((p.active_mask >> m) & 1)) continue;
      const T* table = reinterpret_cast<const T*>(p.tables[m]);
      uint4* dst = reinterpret_cast<uint4*>(smem + (size_t)a * p.hot_cap * kRowBytes);
      for (int idx = threadIdx.x; idx < per_module; idx += kStreamWarps * 32) {
        const int s = idx / kVecPerRow, v = idx - s * kVecPerRow;
        const int id = p.hot_ids[s];
        dst[idx] = __ldg(reinterpret_cast<const uint4*>(table + (long long)id * p.row_stride) + v);
      }
      ++a;
    }
    __syncthreads();
  }

  WarpSmem sm = warp_smem_of(smem + p.hot_bytes + (size_t)warp * p.smem_per_warp, p, n_active);
  for (int t = lane; t < W * p.acc_stride; t += 32) sm.acc[t] = 0.0;
  __syncwarp();

  int warp_flags = 0;
  for (int chunk = gw; chunk < p.n_chunks; chunk += total_warps) {
    const int i_begin = p.bounds[chunk], i_end = p.bounds[chunk + 1];
    if (i_begin >= i_end) continue;
    // offsets one impression ahead: (h0, c0) of impression i and the ends of i and i + 1 are in registers when i starts
    int h0 = p.hist_offsets[i_begin], c0 = p.cand_offsets[i_begin];
    int h1 = p.hist_offsets[i_begin + 1], c1 = p.cand_offsets[i_begin + 1];
    for (int i = i_begin; i < i_end; ++i) {
      const int nxt = min(i + 2, p.n_impr);
      const int h2 = p.hist_offsets[nxt], c2 = p.cand_offsets[nxt];  // consumed at the end of this iteration
      const int H = h1 - h0, C = c1 - c0;
      if (C > p.max_cand || C <= 0 || H < 0) {
        warp_flags |= MB200_FLAG_CAND_OVERFLOW;
        if (p.per_impr)
          for (int t = lane; t < W * MB200_NUM_METRICS; t += 32)
            p.per_impr[((size_t)(t / MB200_NUM_METRICS) * p.n_impr + i) * MB200_NUM_METRICS + t % MB200_NUM_METRICS] = 0.f;
      } else {
        StreamSrc src;
        int ref_first = 0;
        if (PIPE == 1) {
          src.hist_ids = p.hist_ids + h0, src.cand_ids = p.cand_ids + c0;
          src.H = H, src.Hp = (H + R - 1) / R * R, src.C = C, src.n_slots = src.Hp + C;
          src.n_news = p.n_news;
          src.slot_of = (HOT && n_hot > 0) ? p.slot_of : nullptr;
          ref_first = stream_fetch_refs(src, 0, warp_flags);  // the same ids serve every module
        }
        __syncwarp();
        for (int j = lane; j < C; j += 32) sm.lab[j] = p.labels[c0 + j];
        int slot = 0;
        for (int m = 0; m < p.n_modules; ++m) {
          if (!((p.active_mask >> m) & 1)) continue;
          float* s_m = sm.sc + (size_t)slot * p.cpad;
          if (PIPE == 1)
            warp_flags |= stream_pool_score<T, NV, R>(reinterpret_cast<const T*>(p.tables[m]), smem + (size_t)slot * p.hot_cap * kRowBytes, src,
                                                      ref_first, s_m);
          else
            warp_flags |= gather_pool_score<T, NV, R, true, 0, false, HOT>(reinterpret_cast<const T*>(p.tables[m]), p.row_stride, p.vec_per_row, p.n_news,
                                                                           p.hist_ids + h0, H, p.cand_ids + c0, C, s_m, nullptr, 0, nullptr, 0,
                                                                           smem + (size_t)slot * p.hot_cap * kRowBytes,
                                                                           (HOT && n_hot > 0) ? p.slot_of : nullptr);
          __syncwarp();
          if (p.zscore) {
            zscore_inplace(s_m, C, lane);
            __syncwarp();
          }
          ++slot;
        }
        warp_flags |= rank_and_metrics(p, sm, i, h0, H, c0, C);
      }
      h0 = h1, c0 = c1, h1 = h2, c1 = c2;
    }
  }

  __syncwarp();
  for (int t = lane; t < W * MB200_NUM_METRICS; t += 32) {
    const int w = t / MB200_NUM_METRICS, k = t % MB200_NUM_METRICS;
    p.partials[(size_t)t * p.n_partials + gw] = k < p.acc_stride ? sm.acc[w * p.acc_stride + k] : 0.0;
  }
  warp_flags = __reduce_or_sync(kFull, warp_flags);
  if (lane == 0 && warp_flags && p.flags) atomicOr(p.flags, warp_flags);
}

// ---- hot-row directory: which rows does this behaviour set gather most often? -------------------------------------
// counts[id] += 1 for the ids of every `stride`-th block of 1024 consecutive hist / cand ids (an unbiased sample: the
// impressions are exchangeable); REDs on distinct ids run in parallel in the L2, the hottest id serialises ~n/stride/B of them.
__global__ void __launch_bounds__(256) hot_count_kernel(const int32_t* __restrict__ hist_offsets, const int32_t* __restrict__ hist_ids,
                                                        const int32_t* __restrict__ cand_offsets, const int32_t* __restrict__ cand_ids, int n_impr,
                                                        long long n_news, int32_t* __restrict__ counts) {
  const long long n_hist = hist_offsets[n_impr], n_cand = cand_offsets[n_impr];
  const long long hist_blocks = (n_hist + 1023) / 1024, blocks = hist_blocks + (n_cand + 1023) / 1024;
  const long long stride = max(1ll, (blocks * 1024 + (1ll << 21) - 1) >> 21);  // sample <= 2^21 ids: counts stay below 2^22
  for (long long blk = (long long)blockIdx.x * stride; blk < blocks; blk += (long long)gridDim.x * stride) {
    const int32_t* ids = hist_ids;
    long long n = n_hist, first = blk * 1024;
    if (blk >= hist_blocks) ids = cand_ids, n = n_cand, first = (blk - hist_blocks) * 1024;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long t = first + k * 256 + threadIdx.x;
      if (t < n) {
        const int id = ids[t];
        if ((unsigned long long)(long long)id < (unsigned long long)n_news) atomicAdd(&counts[id], 1);
      }
    }
  }
}

// The `cap` most frequently counted ids get cache slots (two-level radix select on the counts, one CTA), every other id
// kHotCold; the cache is switched off (n_hot = 0) when the selected rows cover less than 1/32 of the sampled reads.
__global__ void __launch_bounds__(1024) hot_select_kernel(const int32_t* __restrict__ counts, int n_news, int cap, HotDir* __restrict__ dir,
                                                          int32_t* __restrict__ hot_ids, uint8_t* __restrict__ slot_of) {
  __shared__ int hist[2048];
  __shared__ int s_bin, s_above, s_next, s_ties_left, s_total, s_covered;
  const int t = threadIdx.x;
  // level 1: counts >> 11 (counts < 2^22: the sample is at most 2^21 ids)
  for (int b = t; b < 2048; b += 1024) hist[b] = 0;
  if (t == 0) s_total = 0, s_covered = 0, s_next = 0;
  __syncthreads();
  int local_total = 0;
  for (int i = t; i < n_news; i += 1024) {
    const int c = counts[i];
    local_total += c;
    if (c > 0) atomicAdd(&hist[min(c >> 11, 2047)], 1);
  }
  local_total = __reduce_add_sync(kFull, local_total);
  if ((t & 31) == 0 && local_total) atomicAdd(&s_total, local_total);
  __syncthreads();
  if (t == 0) {
    int above = 0, b = 2047;
    for (; b > 0 && above + hist[b] < cap; --b) above += hist[b];
    s_bin = b, s_above = above;  // the cap-th largest count lies in bin b (or there are fewer than cap non-zero counts: b = 0)
  }
  __syncthreads();
  const int bin1 = s_bin, above1 = s_above;
  __syncthreads();
  // level 2: low 11 bits of the counts in bin1
  for (int b = t; b < 2048; b += 1024) hist[b] = 0;
  __syncthreads();
  for (int i = t; i < n_news; i += 1024) {
    const int c = counts[i];
    if (c > 0 && min(c >> 11, 2047) == bin1) atomicAdd(&hist[c & 2047], 1);
  }
  __syncthreads();
  if (t == 0) {
    int above = above1, b = 2047;
    for (; b > 0 && above + hist[b] < cap; --b) above += hist[b];
    s_bin = b, s_above = above;
    s_ties_left = cap - above;  // slots left for ids whose count equals the threshold
  }
  __syncthreads();
  const int thresh = (bin1 << 11) | s_bin;  // ids with count > thresh are hot, ties fill what is left; 0 never is
  for (int i = t; i < n_news; i += 1024) {
    const int c = counts[i];
    int slot = kHotCold;
    if (c > 0 && c >= thresh) {
      bool take = c > thresh;
      if (!take) take = atomicSub(&s_ties_left, 1) > 0;
      if (take) {
        slot = atomicAdd(&s_next, 1);
        if (slot < cap) hot_ids[slot] = i, atomicAdd(&s_covered, c);
        else slot = kHotCold;
      }
    }
    slot_of[i] = (uint8_t)slot;
  }
  __syncthreads();
  if (t == 0) {
    const int n = min(s_next, cap);
    dir->total = s_total, dir->covered = s_covered;
    dir->n_hot = ((long long)s_covered * 32 >= (long long)s_total && s_total > 0) ? n : 0;
    dir->reserved = n;
  }
}

// mb200_rank_metrics: the ranking / metrics half of the fused kernel on predictions that already exist (one warp per
// impression, same shared-memory layout, rank_and_metrics unchanged).
__global__ void __launch_bounds__(kThreads) rank_metrics_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * kWarpsPerCta + warp;
  const int total_warps = gridDim.x * kWarpsPerCta;
  unsigned char* base = smem + (size_t)warp * p.smem_per_warp;
  WarpSmem sm;
  sm.acc = reinterpret_cast<double*>(base);
  sm.sc = reinterpret_cast<float*>(base + p.acc_bytes);
  sm.comb = sm.sc + p.cpad;
  sm.lab = reinterpret_cast<uint8_t*>(sm.comb + p.cpad);
  sm.ccat = sm.lab + p.cpad;
  sm.csent = sm.ccat + p.cpad;
  sm.hist_cat = reinterpret_cast<int*>(sm.csent + p.cpad);
  sm.hist_sent = sm.hist_cat + MB200_MAX_CLASSES;
  sm.top_cat = reinterpret_cast<uint8_t*>(sm.hist_sent + MB200_MAX_CLASSES);
  sm.top_sent = sm.top_cat + 32;
  for (int t = lane; t < MB200_NUM_METRICS; t += 32) sm.acc[t] = 0.0;
  __syncwarp();
  int warp_flags = 0;
  for (int chunk = gw; chunk < p.n_chunks; chunk += total_warps) {
    for (int i = p.bounds[chunk]; i < p.bounds[chunk + 1]; ++i) {
      const int c0 = p.cand_offsets[i], C = p.cand_offsets[i + 1] - c0;
      const int h0 = p.flat_hist_aspect[0] ? p.hist_offsets[i] : 0;
      const int H = p.flat_hist_aspect[0] ? p.hist_offsets[i + 1] - h0 : 0;
      if (C > p.max_cand || C <= 0 || H < 0) {
        warp_flags |= MB200_FLAG_CAND_OVERFLOW;
        if (p.per_impr)
          for (int t = lane; t < MB200_NUM_METRICS; t += 32) p.per_impr[(size_t)i * MB200_NUM_METRICS + t] = 0.f;
        continue;
      }
      __syncwarp();
      for (int j = lane; j < C; j += 32) sm.sc[j] = p.scores_in[c0 + j], sm.lab[j] = p.labels[c0 + j];
      __syncwarp();
      warp_flags |= rank_and_metrics(p, sm, i, h0, H, c0, C);
    }
  }
  __syncwarp();
  for (int t = lane; t < MB200_NUM_METRICS; t += 32) p.partials[(size_t)t * p.n_partials + gw] = sm.acc[t];
  warp_flags = __reduce_or_sync(kFull, warp_flags);
  if (lane == 0 && warp_flags && p.flags) atomicOr(p.flags, warp_flags);
}

// bounds[c] = first impression whose work prefix reaches c/n_chunks of the total; work = rows gathered
// (+ a per-impression constant), so chunks are balanced by sum(H_i + C_i), not by count.
__global__ void partition_kernel(const int32_t* __restrict__ hist_offsets, const int32_t* __restrict__ cand_offsets, int n_impr,
                                 int n_chunks, int32_t* __restrict__ bounds) {
  constexpr long long kPerImpression = 4;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_chunks) return;
  if (c == n_chunks) {
    bounds[c] = n_impr;
    bounds[c + 1] = 0;  // the dynamic chunk counter of score_eval_kernel lives behind the bounds
    return;
  }
  const long long total = (long long)hist_offsets[n_impr] + cand_offsets[n_impr] + kPerImpression * n_impr;
  const long long target = total * c / n_chunks;  // total < 2^35, c < 2^16: no overflow
  int lo = 0, hi = n_impr;  // smallest i with work(i) >= target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const long long wk = (long long)hist_offsets[mid] + cand_offsets[mid] + kPerImpression * mid;
    if (wk >= target) hi = mid; else lo = mid + 1;
  }
  bounds[c] = lo;
}

// Deterministic second stage: sums[w][k] = sum over warps of partials, fixed order (strided per thread,
// then a fixed shared-memory tree).  One block per (weighting, metric slot): the slots reduce side by side
// instead of one 256-thread block walking all of them (16 us of a 1.7 ms step).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partials, int total_warps, int W,
                                                              double* __restrict__ sums, const int32_t* __restrict__ flags, int n_impr,
                                                              int pack_payload) {
  __shared__ double sh[256];
  const int w = blockIdx.x, k = blockIdx.y, t = threadIdx.x;
  double a = 0.0;
  const double* row = partials + ((size_t)w * MB200_NUM_METRICS + k) * total_warps;  // this (weighting, metric)'s slots, contiguous
  for (int g = t; g < total_warps; g += 256) a += row[g];
  sh[t] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) sh[t] += sh[t + s];
    __syncthreads();
  }
  if (t == 0) sums[(size_t)w * MB200_NUM_METRICS + k] = sh[0];
  if (pack_payload && w == 0 && k == 0 && t >= 32 && t < 32 + MB200_PAYLOAD_TAIL) {
    // additive tail for a multi-GPU sum-reduction: impression count, then the flag word one bit per double
    const int j = t - 32;
    const int f = flags ? *flags : 0;
    // bits 1, 2, 4, 8 and 64 (MB200_FLAG_UPLOAD_TIMEOUT); 16 / 32 belong to the exchange kernels
    sums[(size_t)W * MB200_NUM_METRICS + j] = (j == 0) ? (double)n_impr : (double)((f >> (j == 5 ? 6 : j - 1)) & 1);
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------

struct LaunchPlan {
  int grid = 0;
  int warps_per_cta = kWarpsPerCta;
  int total_warps = 0;
  int n_chunks = 0;
  int cpad = 0;
  int acc_bytes = 0;
  int acc_stride = MB200_NUM_METRICS;
  int smem_per_warp = 0;
  size_t smem_per_cta = 0;
  size_t bounds_bytes = 0, partials_bytes = 0;
  int n_partials = 0;  // slots of `partials`: one per chunk (score_eval_kernel) or per warp (the other kernels)
  // streaming kernel only
  int has_aspects = 1, comb_alias = 0;
  int hot_cap = 0, hot_bytes = 0;
  int upart_off = 0;  // tensor-core bf16 path: where the per-warp area keeps u = hi + mid + lo
};

constexpr int kUpartBytes = 3 * 768 * 2;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr size_t kMaxSmemPerCta = 227 * 1024;

// tuning key 7: 0 = dynamic hand-out exactly when the behaviours arrive through a pipelined upload, 1 = always static, 2 = always dynamic
static bool dynamic_schedule(const mb200_eval_desc* d) {
  const int mode = tuning().static_chunks;
  return mode == 2 || (mode == 0 && d->ready != nullptr);
}

static size_t hot_region_bytes(long long n_news) {
  // HotDir, hot_ids[256], counts[n_news], slot_of[n_news]
  return 256 + 1024 + align_up((size_t)n_news * sizeof(int32_t), 256) + align_up((size_t)n_news, 256);
}

// `stream`: plan for score_eval_stream_kernel (one CTA of kStreamWarps warps per SM; per-warp area trimmed to what the call
// uses; the rest of the shared memory becomes the hot-row cache, `row_bytes` per row and active module).
static int make_plan(const mb200_eval_desc* d, int sm_count, int ctas, LaunchPlan* plan, bool stream = false, int row_bytes = 0, bool hot = false,
                     bool mma = false, bool per_chunk_slots = false) {
  const int n_active = __builtin_popcount((unsigned)d->active_modules_mask);
  plan->cpad = (int)align_up((size_t)(d->max_cand > 0 ? d->max_cand : 1), 32);
  // the aspect-weight sweep (lane per weighting: rank_and_metrics -> sweep_weightings) fills only the first kSweepSlots slots
  const bool sweep = d->weights != nullptr && d->n_weightings >= kSweepMinWeightings && d->news_category == nullptr;
  plan->acc_stride = sweep ? kSweepSlots : MB200_NUM_METRICS;
  plan->acc_bytes = (int)align_up((size_t)d->n_weightings * plan->acc_stride * sizeof(double), 16);
  plan->warps_per_cta = stream ? kStreamWarps : kWarpsPerCta;
  size_t per_warp;
  if (stream) {
    plan->has_aspects = d->news_category != nullptr;
    plan->comb_alias = d->n_weightings == 1;
    per_warp = (size_t)plan->acc_bytes + (size_t)(n_active + (plan->comb_alias ? 0 : 1)) * plan->cpad * sizeof(float) + (size_t)plan->cpad;
    if (plan->has_aspects) per_warp += 2 * (size_t)plan->cpad + 2 * MB200_MAX_CLASSES * sizeof(int) + 64;
    ctas = 1;
  } else {
    per_warp = (size_t)plan->acc_bytes + (size_t)(n_active + 1) * plan->cpad * sizeof(float) + 3 * (size_t)plan->cpad +
               2 * MB200_MAX_CLASSES * sizeof(int) + 64;
  }
  per_warp = align_up(per_warp, 16);
  if (mma) plan->upart_off = (int)per_warp, per_warp += kUpartBytes;
  plan->smem_per_warp = (int)per_warp;
  plan->smem_per_cta = per_warp * plan->warps_per_cta;
  if (plan->smem_per_cta > kMaxSmemPerCta) return MB200_ERR_UNSUPPORTED;
  plan->hot_cap = plan->hot_bytes = 0;
  if (stream && hot) {
    long long cap = (long long)(kMaxSmemPerCta - plan->smem_per_cta) / ((long long)n_active * row_bytes);
    if (cap > kHotMaxSlots) cap = kHotMaxSlots;
    if (tuning().hot_kb_cap > 0) {
      const long long lim = (long long)tuning().hot_kb_cap * 1024 / ((long long)n_active * row_bytes);
      if (cap > lim) cap = lim;
    }
    if (cap >= 4) {
      plan->hot_cap = (int)cap;
      plan->hot_bytes = (int)(cap * n_active * row_bytes);
      plan->smem_per_cta += plan->hot_bytes;
    }
  }
  if (ctas < 1) ctas = 1;
  plan->grid = sm_count * ctas;
  plan->total_warps = plan->grid * plan->warps_per_cta;
  // score_eval_kernel: resident behaviours -> one static chunk per warp (measured best: 1.484 ms against 1.505 ms for 16 dynamic
  // chunks per warp, profiles/r2_i_schedule.log); pipelined upload -> 16 chunks per warp handed out in impression order, so the
  // grid consumes the set in the order the segments arrive.  The other kernels walk one static chunk per warp.
  const int cpw = tuning().chunks_per_warp;
  const int per_warp_chunks = per_chunk_slots ? (cpw > 0 ? cpw : (dynamic_schedule(d) ? 16 : 1)) : 1;
  long long chunks = (long long)plan->total_warps * per_warp_chunks;
  if (chunks > d->n_impressions) chunks = d->n_impressions;
  if (chunks < 1) chunks = 1;
  plan->n_chunks = (int)chunks;
  plan->n_partials = per_chunk_slots ? plan->n_chunks : plan->total_warps;
  plan->bounds_bytes = align_up((size_t)(plan->n_chunks + 2) * sizeof(int32_t), 256);
  plan->partials_bytes = align_up((size_t)plan->n_partials * d->n_weightings * MB200_NUM_METRICS * sizeof(double), 256);
  return MB200_OK;
}

static int validate(const mb200_eval_desc* d) {
  if (d == nullptr || d->struct_size != sizeof(mb200_eval_desc)) return MB200_ERR_INVALID_ARG;
  if (d->n_modules < 1 || d->n_modules > MB200_MAX_MODULES) return MB200_ERR_INVALID_ARG;
  if (!(d->active_modules_mask & 1) || (d->active_modules_mask >> d->n_modules) != 0) return MB200_ERR_INVALID_ARG;
  if (d->dtype != MB200_F32 && d->dtype != MB200_BF16) return MB200_ERR_INVALID_ARG;
  if (d->n_news <= 0 || d->dim <= 0 || d->row_stride < d->dim) return MB200_ERR_INVALID_ARG;
  if (d->n_impressions < 0 || d->n_impressions > 0x7ffffff0ll) return MB200_ERR_INVALID_ARG;
  if (!d->hist_offsets || !d->cand_offsets || !d->sums) return MB200_ERR_INVALID_ARG;
  if (d->n_impressions > 0 && (!d->hist_ids || !d->cand_ids || !d->labels)) return MB200_ERR_INVALID_ARG;
  if (d->n_weightings < 1 || d->max_cand < 1) return MB200_ERR_INVALID_ARG;
  if (d->k0 < 1 || d->k0 > MB200_MAX_K || d->k1 < 1 || d->k1 > MB200_MAX_K) return MB200_ERR_INVALID_ARG;
  if (d->scores && (d->scores_weighting < 0 || d->scores_weighting >= d->n_weightings)) return MB200_ERR_INVALID_ARG;
  if ((d->news_category == nullptr) != (d->news_sentiment == nullptr)) return MB200_ERR_INVALID_ARG;
  if (d->news_category &&
      (d->num_categ_classes < 1 || d->num_categ_classes > MB200_MAX_CLASSES || d->num_sent_classes < 1 || d->num_sent_classes > MB200_MAX_CLASSES))
    return MB200_ERR_INVALID_ARG;
  if (d->n_table_shards < 0 || d->n_table_shards > MB200_MAX_TABLE_SHARDS) return MB200_ERR_INVALID_ARG;
  if (d->n_table_shards > 1) {
    if (d->table_shard_shift < 1 || d->table_shard_shift > 30 || ((long long)d->n_table_shards << d->table_shard_shift) < d->n_news) return MB200_ERR_INVALID_ARG;
    for (int m = 0; m < d->n_modules; ++m)
      for (int sh = 0; sh < d->n_table_shards; ++sh)
        if (((d->active_modules_mask >> m) & 1) && (d->table_shards[m][sh] == nullptr || ((uintptr_t)d->table_shards[m][sh] & 15))) return MB200_ERR_INVALID_ARG;
  }
  if (d->ready != nullptr && (d->ready_segments < 1 || d->ready_segments > MB200_MAX_UPLOAD_SEGMENTS || ((uintptr_t)d->ready & 3))) return MB200_ERR_INVALID_ARG;
  if (d->loss_kind < MB200_LOSS_NONE || d->loss_kind > MB200_LOSS_SUPCON) return MB200_ERR_INVALID_ARG;
  if (d->loss_kind == MB200_LOSS_SUPCON && !(d->loss_temperature > 0.f)) return MB200_ERR_INVALID_ARG;
  if (d->loss_kind != MB200_LOSS_NONE && (d->scores_weighting < 0 || d->scores_weighting >= d->n_weightings)) return MB200_ERR_INVALID_ARG;
  const int esz = d->dtype == MB200_F32 ? 4 : 2;
  for (int m = 0; m < d->n_modules; ++m) {
    if (d->n_table_shards <= 1 && ((d->active_modules_mask >> m) & 1) && (d->tables[m] == nullptr || ((uintptr_t)d->tables[m] & 15))) return MB200_ERR_INVALID_ARG;
  }
  if ((d->row_stride * esz) % 16 != 0 || (d->dim * esz) % 16 != 0) return MB200_ERR_UNSUPPORTED;
  const int vec_per_row = d->dim * esz / 16;
  if (vec_per_row > 8 * 32) return MB200_ERR_UNSUPPORTED;  // dim <= 1024 (fp32) / 2048 (bf16)
  return MB200_OK;
}

using KernelFn = void (*)(const EvalParams);

// Instantiations: the reference width (dim 768: 6 fp32 / 3 bf16 16-byte vectors per lane) gets exact,
// predicate-free kernels in a few rows-in-flight / occupancy trade-offs (mb200_set_tuning key 1); every
// other width runs a predicated kernel whose per-lane vector count is rounded up to 1, 2, 4 or 8.
template <typename T, int NV>
static KernelFn generic_kernel(bool attn) {
  constexpr int R = (NV <= 2) ? 8 : (NV <= 4) ? 6 : 3;
  if (attn) return score_eval_kernel<T, NV, R, false, 0, 3, true>;
  return score_eval_kernel<T, NV, R, false, 0, 3>;
}

template <typename T>
static KernelFn select_kernel(int vec_per_row, bool attn, bool sharded) {
  constexpr int kRefNV = 768 / Elem<T>::E / 32;  // 6 (fp32) or 3 (bf16)
  constexpr int R = (kRefNV == 6) ? 4 : 8;
  constexpr int R3 = (R * 3 + 3) / 4, R2 = (R + 1) / 2;
  if (sharded) return (vec_per_row == kRefNV * 32 && !attn) ? score_eval_kernel<T, kRefNV, R3, true, 3, 4> : nullptr;
  if (vec_per_row == kRefNV * 32 && attn) return score_eval_kernel<T, kRefNV, R3, true, 0, 4, true>;  // early fusion (separate code: the late-fusion kernels keep their tuning)
  if (vec_per_row == kRefNV * 32) {
    int variant = tuning().variant;
    // default: fp32 rows sit at the L2 -> SM limit with 3 rows x 4 CTAs; bf16 rows (half the bytes) are latency bound and
    // want more resident warps: 4 rows x 5 CTAs (measured: profiles/r1_v6_bench_bf16_variants.log)
    if (variant < 0) variant = (Elem<T>::E == 8) ? 3 : 2;
    switch (variant) {
      case 0: return score_eval_kernel<T, kRefNV, R, true, 0, 3>;
      case 1: return score_eval_kernel<T, kRefNV, R, true, 1, 3>;
      case 3: return score_eval_kernel<T, kRefNV, R2, true, 0, 5>;
      case 5: return score_eval_kernel<T, kRefNV, 3, true, 0, 6>;
      case 6: return score_eval_kernel<T, kRefNV, 2, true, 0, 7>;
      default: return score_eval_kernel<T, kRefNV, R3, true, 0, 4>;
    }
  }
  const int nv = (vec_per_row + 31) / 32;
  if (nv <= 1) return generic_kernel<T, 1>(attn);
  if (nv <= 2) return generic_kernel<T, 2>(attn);
  if (nv <= 4) return generic_kernel<T, 4>(attn);
  return generic_kernel<T, 8>(attn);
}

static int sm_count_of(int device, int* out) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) {
    *out = cached[device];
    return MB200_OK;
  }
  int n = 0;
  int st = cuda_status(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute");
  if (st != MB200_OK) return st;
  if (device >= 0 && device < 64) cached[device] = n;
  *out = n;
  return MB200_OK;
}

// Optional CUDA-event bracket around the fused kernel alone (bench.py's roofline figure needs the
// kernel's own duration, not the duration of the partition + kernel + reduction sequence).
struct KernelTimer {
  cudaEvent_t begin = nullptr, end = nullptr;
  bool armed = false;
};
static KernelTimer g_timers[64];
static KernelTimer* g_last_timer = nullptr;

static KernelTimer* timer_for(int device) {
  if (device < 0 || device >= 64) return nullptr;
  KernelTimer* t = &g_timers[device];
  if (t->begin == nullptr) {
    if (cudaEventCreate(&t->begin) != cudaSuccess || cudaEventCreate(&t->end) != cudaSuccess) return nullptr;
  }
  return t;
}

// introspection for bench.py: the hot-row directory of the most recent launch that used the cache
static const HotDir* g_last_hot_dir = nullptr;
static int g_last_hot_cap = 0;

int last_hot_stats(int32_t out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0;
  if (g_last_hot_dir == nullptr) return MB200_OK;
  HotDir h;
  int st = cuda_status(cudaMemcpy(&h, g_last_hot_dir, sizeof(h), cudaMemcpyDeviceToHost), "cudaMemcpy(hot dir)");  // synchronises
  if (st != MB200_OK) return st;
  out[0] = h.n_hot, out[1] = h.total, out[2] = h.covered, out[3] = g_last_hot_cap;
  return MB200_OK;
}

float last_score_kernel_ms() {
  KernelTimer* t = g_last_timer;
  if (t == nullptr || !t->armed) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(t->end) != cudaSuccess || cudaEventElapsedTime(&ms, t->begin, t->end) != cudaSuccess) return -1.0f;
  return ms;
}

// ms from a caller's event to the begin of the most recent timed fused kernel (tools/host_overhead.py: how long the GPU waited
// for the host to get the kernel launched)
float last_score_kernel_begin_after(cudaEvent_t since) {
  KernelTimer* t = g_last_timer;
  if (t == nullptr || !t->armed || since == nullptr) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(t->begin) != cudaSuccess || cudaEventElapsedTime(&ms, since, t->begin) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1.0f;
  }
  return ms;
}

size_t eval_workspace_bytes(const mb200_eval_desc* d) {
  if (validate(d) != MB200_OK) return 0;
  LaunchPlan plan;
  // the SM count is not known without a device; size for the largest part this library targets (148 SMs, <= 160)
  if (make_plan(d, 160, 8, &plan, false, 0, false, false, true) != MB200_OK) return 0;  // upper bound: 160 SMs x 8 CTAs, dynamic chunks (the streaming kernel runs 16 warps per SM: less)
  return plan.bounds_bytes + plan.partials_bytes + 256 + hot_region_bytes(d->n_news);
}

template <typename T>
static KernelFn select_stream_kernel(bool hot, bool rotating) {
  constexpr int kRefNV = 768 / Elem<T>::E / 32;  // 6 (fp32) or 3 (bf16)
  constexpr int R = (kRefNV == 6) ? 3 : 6;       // rows in flight per warp: 72 registers of landing buffers either way
  if (rotating) return hot ? score_eval_stream_kernel<T, kRefNV, R, true, 1> : score_eval_stream_kernel<T, kRefNV, R, false, 1>;
  return hot ? score_eval_stream_kernel<T, kRefNV, R, true, 0> : score_eval_stream_kernel<T, kRefNV, R, false, 0>;
}

int score_eval(const mb200_eval_desc* d, cudaStream_t stream) {
  int st = validate(d);
  if (st != MB200_OK) return st;
  int device = 0;
  st = use_device_of(d->n_table_shards > 1 ? (const void*)d->sums : d->tables[0], &device);  // sharded tables: peers own most shards
  if (st != MB200_OK) return st;
  int sms = 0;
  st = sm_count_of(device, &sms);
  if (st != MB200_OK) return st;
  if (sms > 160) return MB200_ERR_UNSUPPORTED;
  const int esz = d->dtype == MB200_F32 ? 4 : 2;
  const int vec_per_row = d->dim * esz / 16;
  bool attn = false;
  for (int m = 0; m < d->n_modules; ++m) attn |= ((d->active_modules_mask >> m) & 1) && d->attn_logits[m] != nullptr;
  const bool sharded = d->n_table_shards > 1;

  // Round-2 experiments for the reference width with late fusion and replicated tables, selectable but NOT the default
  // (profiles/r2_stream_experiment.md): tuning variant 9 = one 16-warp CTA per SM with the hot-row cache in shared memory (key 6
  // caps its size), 10 = that CTA shape without the cache, 7 / 8 = rotating row pipeline without / with the cache.  Default and
  // 0..6 = the register-batch kernels (4-warp CTAs): the L1 is the landing buffer of the gathers in flight, and every KB of
  // shared memory the cache takes from it costs more bandwidth than the cache's hits return.
  const int variant = tuning().variant;
  LaunchPlan plan;
  KernelFn kern = nullptr;
  bool stream_path = d->dim == 768 && d->row_stride == 768 && !attn && !sharded && (variant >= 7 && variant <= 10) && d->n_news < (1ll << 31) &&
                     d->ready == nullptr;  // the experimental kernels do not gate on a pipelined upload
  bool hot = false, mma = false;
  if (stream_path) {
    if (make_plan(d, sms, 1, &plan, true, vec_per_row * 16, variant != 7 && variant != 10) != MB200_OK) stream_path = false;  // per-warp areas too large for 16 warps
  }
  if (stream_path) {
    hot = plan.hot_cap > 0;
    const bool rotating = variant == 7 || variant == 8;
    kern = (d->dtype == MB200_F32) ? select_stream_kernel<float>(hot, rotating) : select_stream_kernel<__nv_bfloat16>(hot, rotating);
    st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_per_cta), "cudaFuncSetAttribute");
    if (st != MB200_OK) return st;
  } else {
    // bf16 rows of the reference width: tuning variant 11 runs the candidate dot products on the tensor cores (mma.sync).  Measured
    // slower than the FMA kernel (1.36 vs 1.23 ms on Zipf ids, 2.64 vs 1.68 ms on uniform ids: the fragment layout makes every
    // load instruction touch 8 rows x 64 B instead of 512 contiguous bytes of one row; profiles/r2_stream_experiment.md) -- not the default.
    mma = d->dtype == MB200_BF16 && d->dim == 768 && d->row_stride == 768 && !attn && !sharded && variant == 11;
    if (mma) kern = score_eval_kernel<__nv_bfloat16, 3, 4, true, 0, 5, false, true>;
    else kern = (d->dtype == MB200_F32) ? select_kernel<float>(vec_per_row, attn, sharded) : select_kernel<__nv_bfloat16>(vec_per_row, attn, sharded);
    if (kern == nullptr) return MB200_ERR_UNSUPPORTED;  // row-sharded tables: reference width, late fusion only
    st = make_plan(d, sms, 1, &plan, false, 0, false, mma, true);  // shared-memory sizes first: they decide how many CTAs fit
    if (st != MB200_OK) return st;
    if (plan.smem_per_cta > 48 * 1024) {
      st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_per_cta), "cudaFuncSetAttribute");
      if (st != MB200_OK) return st;
    }
    // persistent grid: exactly as many CTAs as are resident at once (registers and shared memory decide)
    int resident = 0;
    st = cuda_status(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kThreads, plan.smem_per_cta), "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (st != MB200_OK) return st;
    if (resident < 1) return MB200_ERR_UNSUPPORTED;
    const int want = tuning().ctas_per_sm > 0 ? tuning().ctas_per_sm : resident;
    st = make_plan(d, sms, want < resident ? want : resident, &plan, false, 0, false, mma, true);
    if (st != MB200_OK) return st;
  }
  const size_t need = plan.bounds_bytes + plan.partials_bytes + (hot ? hot_region_bytes(d->n_news) : 0);
  if (d->workspace == nullptr || ((uintptr_t)d->workspace & 255) || d->workspace_bytes < need) return MB200_ERR_WORKSPACE;

  EvalParams p{};
  for (int m = 0; m < MB200_MAX_MODULES; ++m) p.tables[m] = d->tables[m];
  if (sharded) {
    for (int m = 0; m < d->n_modules; ++m)
      for (int sh = 0; sh < d->n_table_shards; ++sh) p.shard_base[m][sh] = d->table_shards[m][sh];
    p.shard_shift = d->table_shard_shift;
  }
  p.hist_offsets = d->hist_offsets, p.hist_ids = d->hist_ids, p.cand_offsets = d->cand_offsets, p.cand_ids = d->cand_ids;
  p.labels = d->labels, p.weights = d->weights;
  p.news_category = d->news_category, p.news_sentiment = d->news_sentiment;
  p.scores = d->scores, p.per_impr = d->per_impression, p.flags = d->flags;
  unsigned char* ws = reinterpret_cast<unsigned char*>(d->workspace);
  p.bounds = reinterpret_cast<int32_t*>(ws);
  p.chunk_counter = dynamic_schedule(d) ? reinterpret_cast<int*>(ws) + plan.n_chunks + 1 : nullptr;
  p.partials = reinterpret_cast<double*>(ws + plan.bounds_bytes);
  p.n_partials = plan.n_partials;
  p.n_news = d->n_news, p.row_stride = d->row_stride;
  p.n_impr = (int)d->n_impressions, p.n_modules = d->n_modules, p.active_mask = d->active_modules_mask;
  p.vec_per_row = vec_per_row;
  p.zscore = d->zscore, p.n_weightings = d->n_weightings, p.scores_weighting = d->scores_weighting;
  p.k0 = d->k0, p.k1 = d->k1, p.cpad = plan.cpad, p.max_cand = d->max_cand, p.n_chunks = plan.n_chunks;
  p.num_categ = d->num_categ_classes, p.num_sent = d->num_sent_classes;
  p.smem_per_warp = plan.smem_per_warp, p.acc_bytes = plan.acc_bytes, p.acc_stride = plan.acc_stride;
  for (int m = 0; m < MB200_MAX_MODULES; ++m) p.attn_logits[m] = (m < d->n_modules) ? d->attn_logits[m] : nullptr;
  p.hist_pad = d->hist_pad, p.cand_pad = d->cand_pad, p.loss_per_impr = d->loss_per_impression;
  p.loss_kind = d->loss_kind, p.loss_temperature = d->loss_temperature;
  p.has_aspects = plan.has_aspects, p.comb_alias = plan.comb_alias;
  p.hot_cap = plan.hot_cap, p.hot_bytes = plan.hot_bytes, p.upart_off = plan.upart_off;
  p.ready = d->ready;

  KernelTimer* timer = tuning().time_kernel ? timer_for(device) : nullptr;
  int launches = 3;
  if (!hot) g_last_hot_dir = nullptr;
  if (hot && d->n_impressions > 0) {
    // hot-row directory of this behaviour set: sampled id histogram -> the hot_cap most gathered rows -> id -> slot map
    unsigned char* hr = ws + plan.bounds_bytes + plan.partials_bytes;
    HotDir* dir = reinterpret_cast<HotDir*>(hr);
    int32_t* hot_ids = reinterpret_cast<int32_t*>(hr + 256);
    int32_t* counts = reinterpret_cast<int32_t*>(hr + 256 + 1024);
    uint8_t* slot_of = hr + 256 + 1024 + align_up((size_t)d->n_news * sizeof(int32_t), 256);
    st = cuda_status(cudaMemsetAsync(counts, 0, (size_t)d->n_news * sizeof(int32_t), stream), "cudaMemsetAsync(hot counts)");
    if (st != MB200_OK) return st;
    hot_count_kernel<<<sms * 8, 256, 0, stream>>>(d->hist_offsets, d->hist_ids, d->cand_offsets, d->cand_ids, p.n_impr, d->n_news, counts);
    if ((st = cuda_status(cudaGetLastError(), "hot_count_kernel")) != MB200_OK) return st;
    hot_select_kernel<<<1, 1024, 0, stream>>>(counts, (int)d->n_news, plan.hot_cap, dir, hot_ids, slot_of);
    if ((st = cuda_status(cudaGetLastError(), "hot_select_kernel")) != MB200_OK) return st;
    p.hot_dir = dir, p.hot_ids = hot_ids, p.slot_of = slot_of;
    g_last_hot_dir = dir, g_last_hot_cap = plan.hot_cap;
    launches += 2;
  }
  if (d->zero_flags && d->flags) {
    st = cuda_status(cudaMemsetAsync(d->flags, 0, sizeof(int32_t), stream), "cudaMemsetAsync(flags)");
    if (st != MB200_OK) return st;
  }
  partition_kernel<<<(plan.n_chunks + 1 + 255) / 256, 256, 0, stream>>>(d->hist_offsets, d->cand_offsets, p.n_impr, plan.n_chunks,
                                                                       reinterpret_cast<int32_t*>(d->workspace));
  st = cuda_status(cudaGetLastError(), "partition_kernel");
  if (st != MB200_OK) return st;
  if (timer) cudaEventRecord(timer->begin, stream);
  kern<<<plan.grid, plan.warps_per_cta * 32, plan.smem_per_cta, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (timer) cudaEventRecord(timer->end, stream), timer->armed = true, g_last_timer = timer;
  st = cuda_status(e, "score_eval_kernel");
  if (st != MB200_OK) return st;
  reduce_partials_kernel<<<dim3(d->n_weightings, MB200_NUM_METRICS), 256, 0, stream>>>(p.partials, plan.n_partials, d->n_weightings, d->sums, d->flags, p.n_impr,
                                                              d->pack_payload);
  st = cuda_status(cudaGetLastError(), "reduce_partials_kernel");
  if (st != MB200_OK) return st;
  note_launch(launches);
  return MB200_OK;
}

// CUDA loads kernels lazily, at their first launch, and that load can stall behind a kernel that is running and waiting for work
// the host has yet to queue -- exactly what the fused kernel does under a pipelined upload.  Everything that is launched behind
// it is therefore loaded ahead of time (cudaFuncGetAttributes forces the load), once per process, by mb200_upload_begin.
int force_load_eval_kernels() {
  cudaFuncAttributes a;
  int st = cuda_status(cudaFuncGetAttributes(&a, reduce_partials_kernel), "cudaFuncGetAttributes(reduce_partials_kernel)");
  if (st == MB200_OK) st = cuda_status(cudaFuncGetAttributes(&a, partition_kernel), "cudaFuncGetAttributes(partition_kernel)");
  return st;
}

// ---- mb200_rank_metrics ---------------------------------------------------------------------------------
static int metrics_plan(const mb200_metrics_desc* d, int sms, LaunchPlan* plan) {
  if (d == nullptr || d->struct_size != sizeof(mb200_metrics_desc)) return MB200_ERR_INVALID_ARG;
  if (d->n_impressions < 0 || d->n_impressions > 0x7ffffff0ll || d->max_cand < 1 || !d->cand_offsets || !d->sums) return MB200_ERR_INVALID_ARG;
  if (d->n_impressions > 0 && (!d->preds || !d->labels)) return MB200_ERR_INVALID_ARG;
  if (d->k0 < 1 || d->k0 > MB200_MAX_K || d->k1 < 1 || d->k1 > MB200_MAX_K) return MB200_ERR_INVALID_ARG;
  const int n_aspect = (d->cand_category != nullptr) + (d->cand_sentiment != nullptr) + (d->hist_offsets != nullptr) + (d->hist_category != nullptr) +
                       (d->hist_sentiment != nullptr);
  if (n_aspect != 0 && n_aspect != 5) return MB200_ERR_INVALID_ARG;
  if (n_aspect == 5 && (d->num_categ_classes < 1 || d->num_categ_classes > MB200_MAX_CLASSES || d->num_sent_classes < 1 ||
                        d->num_sent_classes > MB200_MAX_CLASSES))
    return MB200_ERR_INVALID_ARG;
  mb200_eval_desc e{};  // the plan only looks at these fields
  e.active_modules_mask = 1, e.max_cand = d->max_cand, e.n_weightings = 1, e.n_impressions = d->n_impressions;
  return make_plan(&e, sms, 4, plan);
}

size_t metrics_workspace_bytes(const mb200_metrics_desc* d) {
  LaunchPlan plan;
  if (metrics_plan(d, 160, &plan) != MB200_OK) return 0;
  return plan.bounds_bytes + plan.partials_bytes + 256;
}

int rank_metrics(const mb200_metrics_desc* d, cudaStream_t stream) {
  LaunchPlan plan;
  int st = metrics_plan(d, 160, &plan);
  if (st != MB200_OK) return st;
  int device = 0;
  if ((st = use_device_of(d->sums, &device)) != MB200_OK) return st;
  int sms = 0;
  if ((st = sm_count_of(device, &sms)) != MB200_OK) return st;
  if (sms > 160) return MB200_ERR_UNSUPPORTED;
  if ((st = metrics_plan(d, sms, &plan)) != MB200_OK) return st;
  if (plan.smem_per_cta > 48 * 1024) {
    st = cuda_status(cudaFuncSetAttribute(rank_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_per_cta), "cudaFuncSetAttribute");
    if (st != MB200_OK) return st;
  }
  if (d->workspace == nullptr || ((uintptr_t)d->workspace & 255) || d->workspace_bytes < plan.bounds_bytes + plan.partials_bytes) return MB200_ERR_WORKSPACE;
  EvalParams p{};
  p.cand_offsets = d->cand_offsets, p.hist_offsets = d->hist_offsets ? d->hist_offsets : d->cand_offsets;
  p.labels = d->labels, p.scores_in = d->preds;
  p.flat_cand_aspect[0] = d->cand_category, p.flat_cand_aspect[1] = d->cand_sentiment;
  p.flat_hist_aspect[0] = d->hist_category, p.flat_hist_aspect[1] = d->hist_sentiment;
  p.per_impr = d->per_impression, p.flags = d->flags;
  p.bounds = reinterpret_cast<int32_t*>(d->workspace);
  p.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d->workspace) + plan.bounds_bytes);
  p.n_partials = plan.total_warps;
  p.n_impr = (int)d->n_impressions, p.n_modules = 1, p.active_mask = 1, p.n_weightings = 1;
  p.k0 = d->k0, p.k1 = d->k1, p.cpad = plan.cpad, p.max_cand = d->max_cand, p.n_chunks = plan.n_chunks;
  p.num_categ = d->num_categ_classes, p.num_sent = d->num_sent_classes;
  p.smem_per_warp = plan.smem_per_warp, p.acc_bytes = plan.acc_bytes, p.acc_stride = plan.acc_stride;
  partition_kernel<<<(plan.n_chunks + 1 + 255) / 256, 256, 0, stream>>>(p.hist_offsets, d->cand_offsets, p.n_impr, plan.n_chunks,
                                                                       reinterpret_cast<int32_t*>(d->workspace));
  if ((st = cuda_status(cudaGetLastError(), "partition_kernel")) != MB200_OK) return st;
  rank_metrics_kernel<<<plan.grid, kThreads, plan.smem_per_cta, stream>>>(p);
  if ((st = cuda_status(cudaGetLastError(), "rank_metrics_kernel")) != MB200_OK) return st;
  reduce_partials_kernel<<<dim3(1, MB200_NUM_METRICS), 256, 0, stream>>>(p.partials, plan.total_warps, 1, d->sums, d->flags, p.n_impr, 0);
  if ((st = cuda_status(cudaGetLastError(), "reduce_partials_kernel")) != MB200_OK) return st;
  note_launch(3);
  return MB200_OK;
}

}  // namespace mb200
