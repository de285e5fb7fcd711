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

// fp32 sigmoid with the reference's operation sequence (ATen sigmoid on a float tensor:
// e = exp(-x); t = 1 + e; s = 1 / t, each rounded to fp32), so that saturation -- s == 1 for x > ~16.6,
// s == 0 once exp overflows fp32 -- produces the same ties torchmetrics sees.  exp itself is evaluated in
// fp64 and rounded once (ATen's vectorised fp32 exp is within 2 ulp of that).
__device__ __forceinline__ float sigmoid_f32(float x) {
  const float e = (float)exp(-(double)x);
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

__global__ void __launch_bounds__(256) auc_build_keys_kernel(const float* __restrict__ preds, const uint8_t* __restrict__ labels,
                                                             long long n, int sigmoid_mode, const int32_t* __restrict__ flags,
                                                             uint32_t* __restrict__ neg_keys, uint32_t* __restrict__ pos_keys,
                                                             unsigned long long* __restrict__ n_pos) {
  const bool sig = sigmoid_mode == 1 || (sigmoid_mode == 2 && flags != nullptr && (*flags & MB200_FLAG_OUTSIDE_UNIT));
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // every lane of a warp runs the same number of iterations so the ballots below are convergent
  const long long warp_first = first - lane;
  for (long long base = warp_first; base < n; base += stride) {
    const long long i = base + lane;
    bool pos = false;
    uint32_t key = kPositiveSentinel;
    if (i < n) {
      float x = preds[i];
      if (sig) x = sigmoid_f32(x);
      key = orderable_key(x);
      pos = labels[i] != 0;
      neg_keys[i] = pos ? kPositiveSentinel : key;
    }
    const unsigned m = __ballot_sync(kFull, pos);
    if (m) {
      unsigned long long slot = 0;
      if (lane == __ffs(m) - 1) slot = atomicAdd(n_pos, (unsigned long long)__popc(m));
      slot = __shfl_sync(kFull, slot, __ffs(m) - 1);
      if (pos) pos_keys[slot + __popc(m & ((1u << lane) - 1u))] = key;
    }
  }
}

__global__ void __launch_bounds__(256) auc_rank_sum_kernel(const uint32_t* __restrict__ sorted, long long n_sorted,
                                                           const long long* __restrict__ n_pos_local, const uint32_t* __restrict__ pos_keys,
                                                           long long pos_capacity, const long long* __restrict__ n_pos,
                                                           unsigned long long* __restrict__ sum2) {
  const long long n_neg = n_sorted - *n_pos_local;
  long long count = *n_pos;
  if (count > pos_capacity) count = pos_capacity;
  unsigned long long local = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t key = pos_keys[i];
    long long lo = 0, hi = n_neg;  // lower_bound: first index with sorted[idx] >= key
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (sorted[mid] < key) lo = mid + 1; else hi = mid;
    }
    const long long lb = lo;
    hi = n_neg;  // upper_bound continues from lb
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (sorted[mid] <= key) lo = mid + 1; else hi = mid;
    }
    local += (unsigned long long)(lb + lo);
  }
  // block reduction, one atomic per block (integer: order-independent, deterministic)
  __shared__ unsigned long long sh[8];
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
  auc_build_keys_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, stream>>>(preds, labels, n, sigmoid_mode, flags, neg_keys, pos_keys,
                                                                       reinterpret_cast<unsigned long long*>(n_pos));
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
  auc_rank_sum_kernel<<<grid_for(pos_capacity, 256, 148 * 8), 256, 0, stream>>>(sorted, n_sorted, n_pos_local, pos_keys, pos_capacity, n_pos,
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

}  // namespace mb200
