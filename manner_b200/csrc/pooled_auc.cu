// Pooled AUROC over all candidate rows of an epoch, as torchmetrics 0.11.4 `AUROC(task="binary")`
// (reference call sites manner/models/cr_module.py:81,273), evaluated as the exact
// rank statistic in integer arithmetic instead of a global argsort + cumsum + fp32 trapezoid:
//
//   auc = sum over positives p of ( #neg with key < key_p  +  #neg with key <= key_p ) / (2 P N)
//
// Stage 1 turns every prediction into an order-preserving uint32 key (fp32 sigmoid first when the
// reference would apply it), leaves the negatives in place and appends the (few) positives to a side
// list.  Stage 2 radix-sorts the keys (cub::DeviceRadixSort -- library code, the one non-hand-written
// kernel family in this library).  Stage 3 binary-searches each positive in the sorted negatives.
// All three are additive over shards: a multi-GPU caller all-gathers only the positive keys and
// all-reduces one uint64 (manner_b200/dist.py).

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mb200 {

constexpr uint32_t kPositiveSentinel = 0xffffffffu;

__device__ __forceinline__ uint32_t orderable_key(float x) {
  if (x == 0.0f) x = 0.0f;  // -0.0 and +0.0 are the same threshold for torchmetrics (preds[1:] - preds[:-1] == 0)
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float key_to_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

// fp32 sigmoid with the reference's operation sequence (ATen sigmoid on a float tensor:
// e = exp(-x); t = 1 + e; s = 1 / t, each rounded to fp32), so that saturation -- s == 1 for x > ~16.6,
// s == 0 once exp overflows fp32 -- produces the same ties torchmetrics sees.  exp itself is evaluated in
// fp64 and rounded once (ATen's vectorised fp32 exp is within 2 ulp of that).
__device__ __forceinline__ float sigmoid_f32(float x) {
  const float e = (float)exp(-(double)x);
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

// One atomic per 2 048-row tile (not per warp: ~60 k same-address atomics serialised the first version at 46 us for 2.7 M rows):
// every thread classifies 8 rows, the block scans the per-thread positive counts, thread 0 reserves the tile's slots.
constexpr int kBuildItems = 8;
// neg_keys == nullptr: only the positives are collected (pooled_auc_bounded); at most pos_capacity of them are stored, *n_pos counts all
__global__ void __launch_bounds__(256) auc_build_keys_kernel(const float* __restrict__ preds, const uint8_t* __restrict__ labels,
                                                             long long n, int sigmoid_mode, const int32_t* __restrict__ flags,
                                                             uint32_t* __restrict__ neg_keys, uint32_t* __restrict__ pos_keys,
                                                             unsigned long long pos_capacity, unsigned long long* __restrict__ n_pos) {
  const bool sig = sigmoid_mode == 1 || (sigmoid_mode == 2 && flags != nullptr && (*flags & MB200_FLAG_OUTSIDE_UNIT));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ unsigned int warp_count[8];
  __shared__ unsigned long long tile_slot;
  constexpr long long kTile = 256 * kBuildItems;
  for (long long tile = (long long)blockIdx.x * kTile; tile < n; tile += (long long)gridDim.x * kTile) {
    uint32_t key[kBuildItems];
    unsigned pos_bits = 0;
#pragma unroll
    for (int u = 0; u < kBuildItems; ++u) {
      const long long i = tile + u * 256 + threadIdx.x;
      key[u] = kPositiveSentinel;
      if (i < n) {
        float x = preds[i];
        if (sig) x = sigmoid_f32(x);
        key[u] = orderable_key(x);
        const bool pos = labels[i] != 0;
        pos_bits |= (pos ? 1u : 0u) << u;
        if (neg_keys) neg_keys[i] = pos ? kPositiveSentinel : key[u];
      }
    }
    // exclusive scan of the per-thread positive counts over the block
    const unsigned mine = __popc(pos_bits);
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_count[warp] = incl;
    __syncthreads();
    unsigned before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const unsigned c = warp_count[w];
      before += (w < warp) ? c : 0u;
      total += c;
    }
    if (threadIdx.x == 0 && total) tile_slot = atomicAdd(n_pos, (unsigned long long)total);
    __syncthreads();
    if (mine) {
      unsigned long long slot = tile_slot + before + (incl - mine);
#pragma unroll
      for (int u = 0; u < kBuildItems; ++u)
        if ((pos_bits >> u) & 1u) {
          if (slot < pos_capacity) pos_keys[slot] = key[u];
          ++slot;
        }
    }
    __syncthreads();  // tile_slot / warp_count are reused by the next tile
  }
}

// ---- rank search: positives against the sorted negatives ---------------------------------------------------------------
// lower_bound + upper_bound of every positive key.  A plain binary search is ~22 dependent L2 round trips per bound; with R ranks'
// positives to rank (multi-GPU: every rank ranks ALL positives against its own negatives) that was the longest kernel after the
// fused one.  Here the first 10 levels run on kSplitters evenly spaced keys staged in shared memory, the remaining levels are a
// branch-free search over a window of uniform length (kSearchIlp keys per thread advance in lock step, so their loads are in flight
// together), and the upper bound gallops from the lower bound (ties between fp32 scores are short runs unless the sigmoid saturates).
constexpr int kSplitters = 1024;      // searches in the sorted negatives (few keys to rank, long array)
constexpr int kPosSplitters = 4096;   // searches in the sorted positives (millions of keys to rank, short array: 2 sectors per key left)
constexpr int kSearchIlp = 4;

struct SortedKeys {
  const uint32_t* keys;
  long long n;       // searchable keys (sentinels behind them are not searched)
  long long window;  // uniform length of the second-level search
  bool split;        // splitters staged (n large enough)
};

template <int S>
__device__ __forceinline__ long long splitter_pos(long long n, int j) { return (long long)(j + 1) * n / (S + 1); }

// block-wide: stage S evenly spaced keys of `keys[0, n)` into sh[S]
template <int S>
__device__ __forceinline__ SortedKeys stage_splitters(const uint32_t* __restrict__ keys, long long n, uint32_t* sh) {
  SortedKeys s;
  s.keys = keys, s.n = n, s.split = n >= 4 * (S + 1);
  s.window = s.split ? n / (S + 1) + 2 : n;
  if (s.split)
    for (int j = threadIdx.x; j < S; j += blockDim.x) sh[j] = __ldg(keys + splitter_pos<S>(n, j));
  __syncthreads();
  return s;
}

__device__ __forceinline__ uint32_t key_at(const SortedKeys& s, long long idx) { return idx < s.n ? __ldg(s.keys + idx) : 0xffffffffu; }

// sum over the keys whose bit is set in `valid` of lower_bound + upper_bound in s
template <int S>
__device__ __forceinline__ unsigned long long bounds_sum(const SortedKeys& s, const uint32_t* sh, const uint32_t (&key)[kSearchIlp], unsigned valid) {
  long long base[kSearchIlp];
#pragma unroll
  for (int k = 0; k < kSearchIlp; ++k) {
    base[k] = 0;
    if (s.split) {
      int j = 0;  // first splitter >= key
#pragma unroll
      for (int half = S / 2; half > 0; half >>= 1) j += (sh[j + half - 1] < key[k]) ? half : 0;
      j += (sh[j] < key[k]) ? 1 : 0;
      base[k] = j ? splitter_pos<S>(s.n, j - 1) + 1 : 0;
    }
  }
  // the answer lies in [base, base + window]; slots beyond the array count as +inf
  long long len = s.window;
  while (len > 1) {
    const long long half = len >> 1;
    uint32_t v[kSearchIlp];
#pragma unroll
    for (int k = 0; k < kSearchIlp; ++k) v[k] = key_at(s, base[k] + half - 1);
#pragma unroll
    for (int k = 0; k < kSearchIlp; ++k) base[k] += (v[k] < key[k]) ? half : 0;
    len -= half;
  }
  unsigned long long sum = 0;
#pragma unroll
  for (int k = 0; k < kSearchIlp; ++k) {
    if (!((valid >> k) & 1u)) continue;
    long long lb = base[k];
    if (len == 1) lb += (key_at(s, lb) < key[k]) ? 1 : 0;
    // upper bound: gallop over the run of keys equal to this one
    long long ub = lb;
    if (ub < s.n && __ldg(s.keys + ub) <= key[k]) {
      long long step = 1;
      while (ub + step < s.n && __ldg(s.keys + ub + step) <= key[k]) step <<= 1;
      long long lo = ub + (step >> 1) + 1, hi = min(ub + step, s.n);  // first index with a key above: in [lo, hi]
      while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(s.keys + mid) <= key[k]) lo = mid + 1; else hi = mid;
      }
      ub = lo;
    }
    sum += (unsigned long long)(lb + ub);
  }
  return sum;
}

// `count` keys of `pos_keys` (optionally mapped through the fp32 sigmoid first), strided over the grid
template <bool SIGMOID>
__device__ __forceinline__ unsigned long long rank_sum_span(const SortedKeys& s, const uint32_t* sh, const uint32_t* pos_keys,
                                                            long long count) {
  unsigned long long local = 0;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += nthreads * kSearchIlp) {
    uint32_t key[kSearchIlp];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < kSearchIlp; ++k) {
      const long long ii = i + k * nthreads;
      key[k] = 0u;
      if (ii < count) {
        key[k] = pos_keys[ii];
        if (SIGMOID) key[k] = orderable_key(sigmoid_f32(key_to_float(key[k])));
        valid |= 1u << k;
      }
    }
    local += bounds_sum<kSplitters>(s, sh, key, valid);
  }
  return local;
}

__global__ void __launch_bounds__(256) auc_rank_sum_kernel(const uint32_t* __restrict__ sorted, long long n_sorted,
                                                           const long long* __restrict__ n_pos_local, const uint32_t* __restrict__ pos_keys,
                                                           long long pos_capacity, const long long* __restrict__ n_pos,
                                                           unsigned long long* __restrict__ sum2) {
  __shared__ uint32_t sh_split[kSplitters];
  __shared__ unsigned long long sh[8];
  const SortedKeys s = stage_splitters<kSplitters>(sorted, n_sorted - *n_pos_local, sh_split);
  long long count = *n_pos;
  if (count > pos_capacity) count = pos_capacity;
  unsigned long long local = rank_sum_span<false>(s, sh_split, pos_keys, count);
  // block reduction, one atomic per block (integer: order-independent, deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(kFull, local, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    if (t) atomicAdd(sum2, t);
  }
}

__global__ void auc_finalize_kernel(const unsigned long long* __restrict__ sum2, const long long* __restrict__ n_pos, long long n,
                                    double* __restrict__ out) {
  const double P = (double)*n_pos, N = (double)(n - *n_pos);
  out[0] = (P > 0 && N > 0) ? (double)*sum2 / (2.0 * P * N) : 0.0;
  out[1] = P, out[2] = N, out[3] = (double)*sum2;
}

static int grid_for(long long n, int block, int cap) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

int auc_build_keys(const float* preds, const uint8_t* labels, long long n, int sigmoid_mode, const int32_t* flags, uint32_t* neg_keys,
                   uint32_t* pos_keys, long long* n_pos, cudaStream_t stream) {
  int st = cuda_status(cudaMemsetAsync(n_pos, 0, sizeof(long long), stream), "cudaMemsetAsync(n_pos)");
  if (st != MB200_OK || n == 0) return st;
  auc_build_keys_kernel<<<grid_for(n, 256 * kBuildItems, 148 * 8), 256, 0, stream>>>(preds, labels, n, sigmoid_mode, flags, neg_keys, pos_keys,
                                                                                   (unsigned long long)n, reinterpret_cast<unsigned long long*>(n_pos));
  note_launch(1);
  return cuda_status(cudaGetLastError(), "auc_build_keys_kernel");
}

size_t auc_sort_workspace_bytes(long long n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 32);
  return bytes + 256;
}

int auc_sort_keys(const uint32_t* keys_in, uint32_t* keys_out, long long n, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n == 0) return MB200_OK;
  size_t need = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, need, keys_in, keys_out, n, 0, 32);
  if (workspace == nullptr || workspace_bytes < need) return MB200_ERR_WORKSPACE;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(workspace, workspace_bytes, keys_in, keys_out, n, 0, 32, stream);
  note_library_launch(1);
  return cuda_status(e, "cub::DeviceRadixSort::SortKeys");
}

int auc_rank_sum(const uint32_t* sorted, long long n_sorted, const long long* n_pos_local, const uint32_t* pos_keys, long long pos_capacity,
                 const long long* n_pos, unsigned long long* sum2, cudaStream_t stream) {
  if (pos_capacity <= 0) return MB200_OK;
  auc_rank_sum_kernel<<<grid_for(pos_capacity, 256, 148 * 4), 256, 0, stream>>>(sorted, n_sorted, n_pos_local, pos_keys, pos_capacity, n_pos,
                                                                               sum2);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "auc_rank_sum_kernel");
}

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

size_t pooled_auc_workspace_bytes(long long n) {
  // neg keys, sorted keys, positive keys, counters, cub scratch
  return 3 * al256((size_t)n * sizeof(uint32_t)) + 256 + auc_sort_workspace_bytes(n);
}

int pooled_auc(const float* preds, const uint8_t* labels, long long n, int sigmoid_mode, const int32_t* flags, void* workspace,
               size_t workspace_bytes, double* out, cudaStream_t stream) {
  if (workspace == nullptr || ((uintptr_t)workspace & 255) || workspace_bytes < pooled_auc_workspace_bytes(n)) return MB200_ERR_WORKSPACE;
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  const size_t keys_bytes = al256((size_t)n * sizeof(uint32_t));
  uint32_t* neg_keys = reinterpret_cast<uint32_t*>(w);
  uint32_t* sorted = reinterpret_cast<uint32_t*>(w + keys_bytes);
  uint32_t* pos_keys = reinterpret_cast<uint32_t*>(w + 2 * keys_bytes);
  long long* n_pos = reinterpret_cast<long long*>(w + 3 * keys_bytes);
  unsigned long long* sum2 = reinterpret_cast<unsigned long long*>(w + 3 * keys_bytes + 8);
  void* cub_ws = w + 3 * keys_bytes + 256;
  const size_t cub_bytes = workspace_bytes - (3 * keys_bytes + 256);

  int st = cuda_status(cudaMemsetAsync(sum2, 0, sizeof(unsigned long long), stream), "cudaMemsetAsync(sum2)");
  if (st != MB200_OK) return st;
  if ((st = auc_build_keys(preds, labels, n, sigmoid_mode, flags, neg_keys, pos_keys, n_pos, stream)) != MB200_OK) return st;
  if (n == 0) {
    st = cuda_status(cudaMemsetAsync(n_pos, 0, sizeof(long long), stream), "cudaMemsetAsync");
    if (st != MB200_OK) return st;
  }
  if ((st = auc_sort_keys(neg_keys, sorted, n, cub_ws, cub_bytes, stream)) != MB200_OK) return st;
  if ((st = auc_rank_sum(sorted, n, n_pos, pos_keys, n, n_pos, sum2, stream)) != MB200_OK) return st;
  auc_finalize_kernel<<<1, 1, 0, stream>>>(sum2, n_pos, n, out);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "auc_finalize_kernel");
}


// ---- pooled AUROC when the positives are few: sort THEM, stream the negatives ------------------------------------------------
// sum_p (#neg < p + #neg <= p)  ==  sum_n (#pos > n + #pos >= n)  =  sum_n (2 P - lower_bound_pos(n) - upper_bound_pos(n)).
// A click log has ~4 % positives: sorting 107 k keys instead of 2.67 M takes the radix sort from 0.12 ms to ~0.02 ms, and the
// 2.56 M negatives are ranked in one streaming pass: 12 search levels on 4 096 positives staged in shared memory, then a window of
// ~28 keys (two 32-byte sectors).  The caller must know an upper bound on the number of positives (the labels are host data
// wherever behaviours are uploaded from); *n_pos above the bound makes the result NaN.
constexpr int kStreamThreads = 512;

template <int S>
__device__ __forceinline__ unsigned long long stream_negatives(const float* __restrict__ preds, const uint8_t* __restrict__ labels, long long n,
                                                               bool sig, const SortedKeys& pos, const uint32_t* sh) {
  unsigned long long local = 0;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads * kSearchIlp) {
    uint32_t key[kSearchIlp];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < kSearchIlp; ++k) {
      const long long ii = i + k * nthreads;
      key[k] = 0u;
      if (ii < n && labels[ii] == 0) {
        float x = preds[ii];
        if (sig) x = sigmoid_f32(x);
        key[k] = orderable_key(x);
        valid |= 1u << k;
      }
    }
    if (valid) local += 2ull * (unsigned long long)pos.n * __popc(valid) - bounds_sum<S>(pos, sh, key, valid);
  }
  return local;
}

// block sum of `local` -> one atomic per block; returns true in the LAST block to finish (all blocks' sums are in *sum2 then)
__device__ __forceinline__ bool block_accumulate(unsigned long long local, unsigned long long* sum2, unsigned int* ticket) {
  __shared__ unsigned long long sh_sum[32];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(kFull, local, o);
  if ((threadIdx.x & 31) == 0) sh_sum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh_sum[w];
    if (t) atomicAdd(sum2, t);
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  return last;
}

__global__ void __launch_bounds__(kStreamThreads) auc_stream_negatives_kernel(const float* __restrict__ preds, const uint8_t* __restrict__ labels,
                                                                              long long n, int sigmoid_mode, const int32_t* __restrict__ flags,
                                                                              const uint32_t* __restrict__ pos_sorted, long long pos_capacity,
                                                                              const unsigned long long* __restrict__ n_pos,
                                                                              unsigned long long* __restrict__ sum2, unsigned int* __restrict__ ticket,
                                                                              double* __restrict__ out) {
  __shared__ uint32_t sh_split[kPosSplitters];
  const bool sig = sigmoid_mode == 1 || (sigmoid_mode == 2 && flags != nullptr && (*flags & MB200_FLAG_OUTSIDE_UNIT));
  const long long P_all = (long long)*n_pos;
  const long long P = min(P_all, pos_capacity);
  const SortedKeys pos = stage_splitters<kPosSplitters>(pos_sorted, P, sh_split);
  const unsigned long long local = stream_negatives<kPosSplitters>(preds, labels, n, sig, pos, sh_split);
  if (block_accumulate(local, sum2, ticket) && threadIdx.x == 0) {
    __threadfence();
    const unsigned long long s2 = atomicAdd(sum2, 0ull);
    const double Pd = (double)P_all, Nd = (double)(n - P_all);
    out[0] = (P_all > pos_capacity) ? __longlong_as_double(0x7ff8000000000000ll) : ((Pd > 0 && Nd > 0) ? (double)s2 / (2.0 * Pd * Nd) : 0.0);
    out[1] = Pd, out[2] = Nd, out[3] = (double)s2;
  }
}

size_t pooled_auc_bounded_workspace_bytes(long long n, long long pos_capacity) {
  if (pos_capacity > n) pos_capacity = n;
  if (pos_capacity < 1) pos_capacity = 1;
  // positive keys, sorted positive keys, counters, cub scratch
  return 2 * al256((size_t)pos_capacity * sizeof(uint32_t)) + 256 + auc_sort_workspace_bytes(pos_capacity);
}

int pooled_auc_bounded(const float* preds, const uint8_t* labels, long long n, long long pos_capacity, int sigmoid_mode, const int32_t* flags,
                       void* workspace, size_t workspace_bytes, double* out, cudaStream_t stream) {
  if (pos_capacity > n) pos_capacity = n;
  if (pos_capacity < 1) pos_capacity = 1;
  if (workspace == nullptr || ((uintptr_t)workspace & 255) || workspace_bytes < pooled_auc_bounded_workspace_bytes(n, pos_capacity)) return MB200_ERR_WORKSPACE;
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  const size_t keys_bytes = al256((size_t)pos_capacity * sizeof(uint32_t));
  uint32_t* pos_keys = reinterpret_cast<uint32_t*>(w);
  uint32_t* pos_sorted = reinterpret_cast<uint32_t*>(w + keys_bytes);
  unsigned long long* n_pos = reinterpret_cast<unsigned long long*>(w + 2 * keys_bytes);
  unsigned long long* sum2 = n_pos + 1;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(n_pos + 2);
  void* cub_ws = w + 2 * keys_bytes + 256;
  const size_t cub_bytes = workspace_bytes - (2 * keys_bytes + 256);

  // slots no positive lands in sort behind the real keys
  int st = cuda_status(cudaMemsetAsync(pos_keys, 0xff, keys_bytes, stream), "cudaMemsetAsync(pos_keys)");
  if (st != MB200_OK) return st;
  if ((st = cuda_status(cudaMemsetAsync(n_pos, 0, 32, stream), "cudaMemsetAsync(counters)")) != MB200_OK) return st;
  if (n > 0) {
    auc_build_keys_kernel<<<grid_for(n, 256 * kBuildItems, 148 * 8), 256, 0, stream>>>(preds, labels, n, sigmoid_mode, flags, nullptr, pos_keys,
                                                                                     (unsigned long long)pos_capacity, n_pos);
    note_launch(1);
    if ((st = cuda_status(cudaGetLastError(), "auc_build_keys_kernel")) != MB200_OK) return st;
    if ((st = auc_sort_keys(pos_keys, pos_sorted, pos_capacity, cub_ws, cub_bytes, stream)) != MB200_OK) return st;
  }
  auc_stream_negatives_kernel<<<grid_for(n, kStreamThreads * kSearchIlp, 148 * 3), kStreamThreads, 0, stream>>>(
      preds, labels, n, sigmoid_mode, flags, pos_sorted, pos_capacity, n_pos, sum2, ticket, out);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "auc_stream_negatives_kernel");
}


// ------------------------------------------------------------------------------------------------------
// Fused multi-GPU exchange: everything an evaluation has to tell the other GPUs -- the fp64 metric payload
// and the raw keys of its (few) positives -- is STORED straight into every peer's mailbox over NVLink / NVSwitch peer memory
// by one kernel; a second kernel waits for the peers' stores, reduces the payloads in rank order, ranks ALL ranks' positives
// against the local sorted negatives, posts the three additive integers and sums the peers'.  No NCCL call on the path: it
// replaces all_reduce(sums) + all_gather(positive keys) + all_reduce(stats) + one rank-sum launch per rank.
//
// The negatives are sorted by their RAW score keys before anything is exchanged; whether torchmetrics' AUROC would apply the
// sigmoid is only known globally (any score outside [0,1] on any rank), so the rank search compares sigmoid(key) when the
// reduced payload says so -- fp32 sigmoid is monotone, hence the raw order is a valid order for it (ties only merge).
//
// Mailbox of a rank = 2 (epoch parity) x n_ranks slots; slot = [256 B header][payload doubles][positive keys].
// ------------------------------------------------------------------------------------------------------

struct ExchangeHeader {
  uint32_t flag1;  // epoch once payload + positives of this slot are complete
  uint32_t flag2;  // epoch once the statistics below are complete
  long long n_pos;
  unsigned long long sum2;
  long long pos_total, neg_total;
  // local scratch of the OWNER of the mailbox (only used in its own slot [parity][my_rank])
  unsigned int ticket;
  unsigned int pad;
  unsigned long long acc;
};
static_assert(sizeof(ExchangeHeader) <= 256, "mailbox header");

struct ExchangeParams {
  unsigned char* mailbox[MB200_MAX_TABLE_SHARDS];
  int n_ranks, my_rank;
  uint32_t epoch;
  int n_payload;
  int outside_index;  // payload entry that is > 0 when some score of that rank was outside [0,1] (-1: never)
  long long pos_capacity;
  size_t slot_bytes, payload_off, keys_off;
  const double* payload;
  const uint32_t* pos_keys;
  const long long* n_pos;
  const uint32_t* sorted_neg;  // this rank's keys, sorted by RAW score key (negatives in front)
  uint32_t* sorted_sig;        // workspace [n_rows]: the same negatives as sigmoid keys (still sorted: the fp32 sigmoid is monotone)
  long long n_rows;
  double* out_payload;
  long long* out_stats;
  int32_t* flags;
};

__device__ __forceinline__ unsigned char* slot_of_mailbox(const ExchangeParams& p, int owner, int src) {
  return p.mailbox[owner] + ((size_t)(p.epoch & 1u) * p.n_ranks + src) * p.slot_bytes;
}

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// waits until *flag == epoch (bounded: 4 s, then reports instead of hanging the GPU)
__device__ bool wait_flag(const uint32_t* flag, uint32_t epoch) {
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(flag) != epoch) {
    __nanosleep(200);
    if (global_timer_ns() - t0 > 4000000000ull) return false;
  }
  return true;
}

// One launch, two jobs that run side by side: the first kPostBlocks CTAs STORE this rank's payload and positive keys into every
// rank's mailbox (NVLink stores, then the epoch flag); the other CTAs meanwhile prepare the sigmoid keys of this rank's sorted
// negatives -- the rank search after the exchange then compares plain integers whichever way AUROC's sigmoid rule (known only once
// all payloads are in) turns out.  (As two kernels back to back they cost 36 us at 2 GPUs; the stores are latency, the sigmoid is
// fp64 arithmetic.)
constexpr int kPostBlocks = 148;

__global__ void __launch_bounds__(256) exchange_post_kernel(const ExchangeParams p, int with_sigmoid) {
  if ((int)blockIdx.x >= kPostBlocks) {
    if (!with_sigmoid) return;
    const long long n_neg = p.n_rows - *p.n_pos;
    const long long nthreads = (long long)(gridDim.x - kPostBlocks) * blockDim.x;
    for (long long i = (long long)(blockIdx.x - kPostBlocks) * blockDim.x + threadIdx.x; i < n_neg; i += nthreads)
      p.sorted_sig[i] = orderable_key(sigmoid_f32(key_to_float(p.sorted_neg[i])));
    return;
  }
  const long long n_pos = min(*p.n_pos, p.pos_capacity);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)kPostBlocks * blockDim.x;
  if (tid == 0 && p.flags) *p.flags = 0, p.flags[1] = 0;  // the exchange's own flag word (an int64 slot of the result): set by the finish kernel only
  for (int r = 0; r < p.n_ranks; ++r) {
    unsigned char* slot = slot_of_mailbox(p, r, p.my_rank);
    double* pay = reinterpret_cast<double*>(slot + p.payload_off);
    for (long long i = tid; i < p.n_payload; i += nthreads) pay[i] = p.payload[i];
    uint32_t* keys = reinterpret_cast<uint32_t*>(slot + p.keys_off);
    for (long long i = tid; i < n_pos; i += nthreads) keys[i] = p.pos_keys[i];
    if (tid == 0) reinterpret_cast<ExchangeHeader*>(slot)->n_pos = *p.n_pos;  // the true count: the receiver reports an overflow
  }
  ExchangeHeader* own = reinterpret_cast<ExchangeHeader*>(slot_of_mailbox(p, p.my_rank, p.my_rank));
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(&own->ticket, 1u) == kPostBlocks - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) own->ticket = 0, own->acc = 0;
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p.n_ranks) st_release_sys(&reinterpret_cast<ExchangeHeader*>(slot_of_mailbox(p, threadIdx.x, p.my_rank))->flag1, p.epoch);
  }
}

__global__ void __launch_bounds__(256) exchange_finish_kernel(const ExchangeParams p) {
  __shared__ bool last;
  __shared__ int sh_good, sh_outside;
  __shared__ unsigned long long sh[8];
  __shared__ uint32_t sh_split[kSplitters];
  __shared__ unsigned long long sh_stats[MB200_MAX_TABLE_SHARDS][3];
  // one thread per rank waits for that rank's slot (a single thread walking eight system-scope polls in turn cost more than the search)
  if (threadIdx.x == 0) sh_good = 1, sh_outside = 0;
  __syncthreads();
  if ((int)threadIdx.x < p.n_ranks) {
    unsigned char* slot = slot_of_mailbox(p, p.my_rank, threadIdx.x);
    const bool good = wait_flag(&reinterpret_cast<ExchangeHeader*>(slot)->flag1, p.epoch);
    if (!good) atomicAnd(&sh_good, 0);
    else if (p.outside_index >= 0 && reinterpret_cast<const double*>(slot + p.payload_off)[p.outside_index] > 0.0) atomicOr(&sh_outside, 1);
  }
  __syncthreads();
  const bool ok = sh_good != 0, sig = sh_outside != 0;
  if (!ok) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.flags) atomicOr(p.flags, MB200_FLAG_EXCHANGE_TIMEOUT);
    return;
  }
  // metric payload: fixed rank order, every rank computes the same doubles
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < p.n_payload; i += blockDim.x) {
      double a = 0.0;
      for (int r = 0; r < p.n_ranks; ++r) a += reinterpret_cast<const double*>(slot_of_mailbox(p, p.my_rank, r) + p.payload_off)[i];
      p.out_payload[i] = a;
    }
  }
  // every rank's positives against MY sorted negatives
  const long long my_pos = *p.n_pos;
  const long long n_neg = p.n_rows - my_pos;
  const bool use_sig = sig;
  const SortedKeys negs = stage_splitters<kSplitters>(use_sig ? p.sorted_sig : p.sorted_neg, n_neg, sh_split);
  // all ranks' positives as ONE index space, so that a sweep of the grid is full whatever the number of ranks (one strided loop per
  // rank left four fifths of the threads idle in each of R latency-bound sweeps)
  unsigned long long local = 0;
  bool overflow = false;
  long long first[MB200_MAX_TABLE_SHARDS + 1];
  const uint32_t* keys_of[MB200_MAX_TABLE_SHARDS];
  first[0] = 0;
#pragma unroll
  for (int r = 0; r < MB200_MAX_TABLE_SHARDS; ++r) {
    long long cnt = 0;
    keys_of[r] = nullptr;
    if (r < p.n_ranks) {
      const unsigned char* slot = slot_of_mailbox(p, p.my_rank, r);
      cnt = reinterpret_cast<const ExchangeHeader*>(slot)->n_pos;
      if (cnt > p.pos_capacity) cnt = p.pos_capacity, overflow = true;
      keys_of[r] = reinterpret_cast<const uint32_t*>(slot + p.keys_off);
    }
    first[r + 1] = first[r] + cnt;
  }
  const long long total = first[MB200_MAX_TABLE_SHARDS];
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += nthreads * kSearchIlp) {
    uint32_t key[kSearchIlp];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < kSearchIlp; ++k) {
      const long long gg = g + k * nthreads;
      key[k] = 0u;
      if (gg < total) {
        const uint32_t* src = keys_of[0];
        long long base = 0;
#pragma unroll
        for (int r = 1; r < MB200_MAX_TABLE_SHARDS; ++r)
          if (gg >= first[r]) src = keys_of[r], base = first[r];
        key[k] = src[gg - base];
        if (use_sig) key[k] = orderable_key(sigmoid_f32(key_to_float(key[k])));
        valid |= 1u << k;
      }
    }
    local += bounds_sum<kSplitters>(negs, sh_split, key, valid);
  }
  if (overflow && blockIdx.x == 0 && threadIdx.x == 0 && p.flags) atomicOr(p.flags, MB200_FLAG_POS_OVERFLOW);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(kFull, local, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  ExchangeHeader* own = reinterpret_cast<ExchangeHeader*>(slot_of_mailbox(p, p.my_rank, p.my_rank));
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    if (t) atomicAdd(&own->acc, t);
    __threadfence();
    last = atomicAdd(&own->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // the last block posts this rank's three additive integers to every peer, then sums what the peers posted: one thread per peer, so
  // the R NVLink round trips (stores + fence + flag out, flag + statistics back) run side by side
  __shared__ unsigned long long sh_sum2;
  if (threadIdx.x == 0) {
    __threadfence();
    sh_sum2 = atomicAdd(&own->acc, 0ull);
    own->ticket = 0;
    sh_good = 1;
  }
  __syncthreads();
  if ((int)threadIdx.x < p.n_ranks) {
    const int r = threadIdx.x;
    ExchangeHeader* h = reinterpret_cast<ExchangeHeader*>(slot_of_mailbox(p, r, p.my_rank));
    h->sum2 = sh_sum2, h->pos_total = my_pos, h->neg_total = n_neg;
    __threadfence_system();
    st_release_sys(&h->flag2, p.epoch);
    const ExchangeHeader* in = reinterpret_cast<const ExchangeHeader*>(slot_of_mailbox(p, p.my_rank, r));
    if (!wait_flag(&in->flag2, p.epoch)) atomicAnd(&sh_good, 0);
    sh_stats[r][0] = *reinterpret_cast<const volatile unsigned long long*>(&in->sum2);
    sh_stats[r][1] = (unsigned long long)*reinterpret_cast<const volatile long long*>(&in->pos_total);
    sh_stats[r][2] = (unsigned long long)*reinterpret_cast<const volatile long long*>(&in->neg_total);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s2 = 0, P = 0, N = 0;
    for (int r = 0; r < p.n_ranks; ++r) s2 += sh_stats[r][0], P += sh_stats[r][1], N += sh_stats[r][2];
    if (!sh_good && p.flags) atomicOr(p.flags, MB200_FLAG_EXCHANGE_TIMEOUT);
    p.out_stats[0] = (long long)s2, p.out_stats[1] = (long long)P, p.out_stats[2] = (long long)N;
  }
}

static size_t exchange_layout(int n_payload, long long pos_capacity, size_t* payload_off, size_t* keys_off) {
  const size_t po = 256, ko = po + al256((size_t)n_payload * sizeof(double));
  if (payload_off) *payload_off = po;
  if (keys_off) *keys_off = ko;
  return ko + al256((size_t)pos_capacity * sizeof(uint32_t));
}

size_t exchange_mailbox_bytes(int n_ranks, int n_payload, long long pos_capacity) {
  if (n_ranks < 1 || n_ranks > MB200_MAX_TABLE_SHARDS || n_payload < 0 || pos_capacity < 0) return 0;
  return 2 * (size_t)n_ranks * exchange_layout(n_payload, pos_capacity, nullptr, nullptr);
}

size_t exchange_workspace_bytes(long long n_rows) { return n_rows < 0 ? 0 : al256((size_t)n_rows * sizeof(uint32_t)) + 256; }

static int exchange_params(const mb200_exchange_desc* d, ExchangeParams* p) {
  if (d == nullptr || d->struct_size != sizeof(mb200_exchange_desc)) return MB200_ERR_INVALID_ARG;
  if (d->n_ranks < 1 || d->n_ranks > MB200_MAX_TABLE_SHARDS || d->my_rank < 0 || d->my_rank >= d->n_ranks) return MB200_ERR_INVALID_ARG;
  if (d->n_payload < 0 || d->pos_capacity < 0 || d->n_rows < 0 || d->epoch == 0) return MB200_ERR_INVALID_ARG;
  if (d->outside_index >= d->n_payload) return MB200_ERR_INVALID_ARG;
  if (!d->n_pos || !d->out_payload || !d->out_stats || (d->n_payload > 0 && !d->payload)) return MB200_ERR_INVALID_ARG;
  if (d->n_rows > 0 && (!d->sorted_neg || !d->pos_keys)) return MB200_ERR_INVALID_ARG;
  if (d->workspace == nullptr || ((uintptr_t)d->workspace & 255) || d->workspace_bytes < exchange_workspace_bytes(d->n_rows)) return MB200_ERR_WORKSPACE;
  for (int r = 0; r < d->n_ranks; ++r)
    if (d->mailbox[r] == nullptr || ((uintptr_t)d->mailbox[r] & 255)) return MB200_ERR_INVALID_ARG;
  for (int r = 0; r < d->n_ranks; ++r) p->mailbox[r] = reinterpret_cast<unsigned char*>(d->mailbox[r]);
  p->n_ranks = d->n_ranks, p->my_rank = d->my_rank, p->epoch = d->epoch, p->n_payload = d->n_payload, p->outside_index = d->outside_index;
  p->pos_capacity = d->pos_capacity;
  p->slot_bytes = exchange_layout(d->n_payload, d->pos_capacity, &p->payload_off, &p->keys_off);
  p->payload = d->payload, p->pos_keys = d->pos_keys, p->n_pos = reinterpret_cast<const long long*>(d->n_pos);
  p->sorted_neg = d->sorted_neg, p->n_rows = d->n_rows, p->sorted_sig = reinterpret_cast<uint32_t*>(d->workspace);
  p->out_payload = d->out_payload, p->out_stats = reinterpret_cast<long long*>(d->out_stats), p->flags = d->flags;
  return MB200_OK;
}

int exchange_post(const mb200_exchange_desc* d, cudaStream_t stream) {
  ExchangeParams p{};
  int st = exchange_params(d, &p);
  if (st != MB200_OK) return st;
  if ((st = use_device_of(d->out_payload, nullptr)) != MB200_OK) return st;
  const int with_sigmoid = (d->n_rows > 0 && d->outside_index >= 0) ? 1 : 0;
  const int sig_blocks = with_sigmoid ? grid_for(d->n_rows, 256 * 4, 148 * 6) : 0;
  exchange_post_kernel<<<kPostBlocks + sig_blocks, 256, 0, stream>>>(p, with_sigmoid);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "exchange_post_kernel");
}

int exchange_finish(const mb200_exchange_desc* d, cudaStream_t stream) {
  ExchangeParams p{};
  int st = exchange_params(d, &p);
  if (st != MB200_OK) return st;
  if ((st = use_device_of(d->out_payload, nullptr)) != MB200_OK) return st;
  exchange_finish_kernel<<<148 * 4, 256, 0, stream>>>(p);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "exchange_finish_kernel");
}

// Loads every kernel of this file that can be launched behind a fused kernel that is still waiting for its pipelined upload (see
// force_load_eval_kernels in score_eval.cu), including the CUB radix-sort kernels: those are templates of the library, so they
// are loaded by running one small sort on scratch memory.
int force_load_auc_kernels(cudaStream_t stream) {
  cudaFuncAttributes a;
  int st = cuda_status(cudaFuncGetAttributes(&a, auc_build_keys_kernel), "cudaFuncGetAttributes");
  if (st == MB200_OK) st = cuda_status(cudaFuncGetAttributes(&a, auc_rank_sum_kernel), "cudaFuncGetAttributes");
  if (st == MB200_OK) st = cuda_status(cudaFuncGetAttributes(&a, auc_finalize_kernel), "cudaFuncGetAttributes");
  if (st == MB200_OK) st = cuda_status(cudaFuncGetAttributes(&a, auc_stream_negatives_kernel), "cudaFuncGetAttributes");
  if (st == MB200_OK) st = cuda_status(cudaFuncGetAttributes(&a, exchange_post_kernel), "cudaFuncGetAttributes");
  if (st == MB200_OK) st = cuda_status(cudaFuncGetAttributes(&a, exchange_finish_kernel), "cudaFuncGetAttributes");
  if (st != MB200_OK) return st;
  const long long n = 1 << 16;  // large enough for the multi-pass (onesweep) path the real sorts take
  const size_t cub_bytes = auc_sort_workspace_bytes(n);
  unsigned char* scratch = nullptr;
  if ((st = cuda_status(cudaMalloc(&scratch, 2 * n * sizeof(uint32_t) + cub_bytes), "cudaMalloc(warm-up scratch)")) != MB200_OK) return st;
  st = cuda_status(cudaMemsetAsync(scratch, 0x5a, 2 * n * sizeof(uint32_t), stream), "cudaMemsetAsync");
  if (st == MB200_OK)
    st = auc_sort_keys(reinterpret_cast<uint32_t*>(scratch), reinterpret_cast<uint32_t*>(scratch) + n, n, scratch + 2 * n * sizeof(uint32_t), cub_bytes, stream);
  if (st == MB200_OK) st = cuda_status(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  cudaFree(scratch);
  return st;
}

// ---- read-bandwidth probe (bench.py's L2 roofline denominator) ----------------------------------------------------------
// Every warp streams 3 KB rows (the gather's access shape: 6 x LDG.E.128 per lane and row, K rows in flight) of a buffer of
// 2^k rows, `repeats` passes; rows of one batch are far apart like gathered rows.  With a buffer that fits the 126 MB L2 this
// measures the L2 -> SM read bandwidth a row gather can reach at best, with a larger one the HBM read bandwidth.  Three shapes
// of the same bytes in flight per SM (mode 0: 16 warps x 4 rows, 1: 32 x 2, 2: 64 x 1); the caller takes the best.
template <int K>
__device__ __forceinline__ void read_probe_body(const uint4* __restrict__ buf, unsigned row_mask, int repeats, unsigned int* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const unsigned n_rows = row_mask + 1;
  unsigned int acc = 0;
  for (int it = 0; it < repeats; ++it) {
    for (unsigned r0 = gw * K; r0 < n_rows; r0 += nw * K) {
      uint4 v[K][6];
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const unsigned row = ((r0 + r) * 2654435761u + (unsigned)it * 40503u) & row_mask;  // odd multiplier: a permutation of the rows
        const uint4* src = buf + (size_t)row * 192 + lane;
#pragma unroll
        for (int k = 0; k < 6; ++k) v[r][k] = __ldg(src + 32 * k);
      }
#pragma unroll
      for (int r = 0; r < K; ++r)
#pragma unroll
        for (int k = 0; k < 6; ++k) acc ^= v[r][k].x ^ v[r][k].y ^ v[r][k].z ^ v[r][k].w;
    }
  }
  if (acc == 0x9e3779b9u) *sink = acc;  // keeps the loads alive
}
__global__ void __launch_bounds__(512, 1) read_probe_kernel4(const uint4* __restrict__ buf, unsigned row_mask, int repeats, unsigned int* __restrict__ sink) {
  read_probe_body<4>(buf, row_mask, repeats, sink);
}
__global__ void __launch_bounds__(1024, 1) read_probe_kernel2(const uint4* __restrict__ buf, unsigned row_mask, int repeats, unsigned int* __restrict__ sink) {
  read_probe_body<2>(buf, row_mask, repeats, sink);
}
__global__ void __launch_bounds__(1024, 2) read_probe_kernel1(const uint4* __restrict__ buf, unsigned row_mask, int repeats, unsigned int* __restrict__ sink) {
  read_probe_body<1>(buf, row_mask, repeats, sink);
}

int read_probe(const void* buf, size_t bytes, int repeats, int mode, void* sink, cudaStream_t stream) {
  if (buf == nullptr || sink == nullptr || bytes < 3072 || repeats < 1 || ((uintptr_t)buf & 15) || mode < 0 || mode > 2) return MB200_ERR_INVALID_ARG;
  size_t rows = bytes / 3072;
  if (rows & (rows - 1)) return MB200_ERR_INVALID_ARG;  // power of two
  if (rows > (1ull << 31)) return MB200_ERR_UNSUPPORTED;
  int device = 0, sms = 0;
  int st = use_device_of(buf, &device);
  if (st != MB200_OK) return st;
  if ((st = cuda_status(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute")) != MB200_OK) return st;
  const uint4* b = reinterpret_cast<const uint4*>(buf);
  unsigned int* sk = reinterpret_cast<unsigned int*>(sink);
  const unsigned mask = (unsigned)(rows - 1);
  if (mode == 0) read_probe_kernel4<<<sms, 512, 0, stream>>>(b, mask, repeats, sk);
  else if (mode == 1) read_probe_kernel2<<<sms, 1024, 0, stream>>>(b, mask, repeats, sk);
  else read_probe_kernel1<<<sms * 2, 1024, 0, stream>>>(b, mask, repeats, sk);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "read_probe_kernel");
}

}  // namespace mb200
