// extern "C" surface of libmanner_b200.so (include/manner_b200.h).  Thin: argument checks, device
// selection, launch bookkeeping.  No torch types, no allocation of caller-visible memory.

#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace mb200 {

static thread_local char g_cuda_err[256] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_library_launches{0};

Tuning& tuning() {
  static Tuning t;
  return t;
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MB200_OK;
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  (void)cudaGetLastError();  // clear the sticky-free error so the next call starts clean
  return MB200_ERR_CUDA;
}

void note_launch(int n) { g_launches += n; }
void note_library_launch(int n) { g_library_launches += n; }

int use_device_of(const void* ptr, int* device_out) {
  cudaPointerAttributes attr;
  int st = cuda_status(cudaPointerGetAttributes(&attr, ptr), "cudaPointerGetAttributes");
  if (st != MB200_OK) return st;
  if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "pointer %p is not device memory", ptr);
    return MB200_ERR_INVALID_ARG;
  }
  st = cuda_status(cudaSetDevice(attr.device), "cudaSetDevice");
  if (st != MB200_OK) return st;
  if (device_out) *device_out = attr.device;
  return MB200_OK;
}

// implemented in score_eval.cu / pooled_auc.cu
float host_dcg_discount(int rank);
float last_score_kernel_ms();
float last_score_kernel_begin_after(cudaEvent_t since);
int last_hot_stats(int32_t out[4]);
size_t eval_workspace_bytes(const mb200_eval_desc* d);
int score_eval(const mb200_eval_desc* d, cudaStream_t stream);
int auc_build_keys(const float*, const uint8_t*, long long, int, const int32_t*, uint32_t*, uint32_t*, long long*, cudaStream_t);
size_t auc_sort_workspace_bytes(long long n);
int auc_sort_keys(const uint32_t*, uint32_t*, long long, void*, size_t, cudaStream_t);
int auc_rank_sum(const uint32_t*, long long, const long long*, const uint32_t*, long long, const long long*, unsigned long long*, cudaStream_t);
size_t pooled_auc_workspace_bytes(long long n);
int pooled_auc(const float*, const uint8_t*, long long, int, const int32_t*, void*, size_t, double*, cudaStream_t);
size_t pooled_auc_bounded_workspace_bytes(long long n, long long pos_capacity);
int pooled_auc_bounded(const float*, const uint8_t*, long long, long long, int, const int32_t*, void*, size_t, double*, cudaStream_t);
size_t retrieval_workspace_bytes(const mb200_retrieval_desc* d);
int retrieve_topk(const mb200_retrieval_desc* d, cudaStream_t stream);
int pool_users(const void*, int, int, long long, long long, const int32_t*, const int32_t*, long long, void*, int32_t*, cudaStream_t);
int merge_topk(const float*, const long long*, int, long long, int, float*, long long*, cudaStream_t);
int attention_logits(const void*, int, int, long long, long long, const float*, const float*, const float*, int, float*, cudaStream_t);
int step_loss(const float*, long long, int, int, const int32_t*, const uint8_t*, double*, cudaStream_t);
size_t exchange_mailbox_bytes(int n_ranks, int n_payload, long long pos_capacity);
size_t exchange_workspace_bytes(long long n_rows);
int exchange_post(const mb200_exchange_desc* d, cudaStream_t stream);
int exchange_finish(const mb200_exchange_desc* d, cudaStream_t stream);
int read_probe(const void* buf, size_t bytes, int repeats, int mode, void* sink, cudaStream_t stream);
int upload_begin(const mb200_upload_desc* d, cudaStream_t compute);
int upload_finish(const mb200_upload_desc* d);
size_t metrics_workspace_bytes(const mb200_metrics_desc* d);
int rank_metrics(const mb200_metrics_desc* d, cudaStream_t stream);

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb200_abi_version(void) { return MB200_ABI_VERSION; }

const char* mb200_status_str(int status) {
  switch (status) {
    case MB200_OK: return "ok";
    case MB200_ERR_INVALID_ARG: return "invalid argument";
    case MB200_ERR_UNSUPPORTED: return "unsupported shape (dim / row stride / max_cand outside what the kernels cover)";
    case MB200_ERR_CUDA: return "CUDA runtime error";
    case MB200_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
  }
  return "unknown status";
}

const char* mb200_last_cuda_error(void) { return g_cuda_err; }

size_t mb200_eval_workspace_bytes(const mb200_eval_desc* desc) { return eval_workspace_bytes(desc); }

int mb200_score_eval(const mb200_eval_desc* desc, void* stream) { return score_eval(desc, static_cast<cudaStream_t>(stream)); }

int mb200_upload_begin(const mb200_upload_desc* desc, void* compute_stream) { return upload_begin(desc, static_cast<cudaStream_t>(compute_stream)); }

int mb200_upload_finish(const mb200_upload_desc* desc) { return upload_finish(desc); }

int mb200_auc_build_keys(const float* preds, const uint8_t* labels, int64_t n, int sigmoid_mode, const int32_t* flags, uint32_t* neg_keys,
                         uint32_t* pos_keys, int64_t* n_pos, void* stream) {
  if (n < 0 || n_pos == nullptr || (n > 0 && (!preds || !labels || !neg_keys || !pos_keys))) return MB200_ERR_INVALID_ARG;
  if (sigmoid_mode < 0 || sigmoid_mode > 2 || (sigmoid_mode == 2 && flags == nullptr)) return MB200_ERR_INVALID_ARG;
  int st = use_device_of(n_pos, nullptr);
  if (st != MB200_OK) return st;
  return auc_build_keys(preds, labels, n, sigmoid_mode, flags, neg_keys, pos_keys, reinterpret_cast<long long*>(n_pos),
                        static_cast<cudaStream_t>(stream));
}

size_t mb200_auc_sort_workspace_bytes(int64_t n) { return n < 0 ? 0 : auc_sort_workspace_bytes(n); }

int mb200_auc_sort_keys(const uint32_t* keys_in, uint32_t* keys_out, int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || (n > 0 && (!keys_in || !keys_out))) return MB200_ERR_INVALID_ARG;
  if (n == 0) return MB200_OK;
  int st = use_device_of(keys_in, nullptr);
  if (st != MB200_OK) return st;
  return auc_sort_keys(keys_in, keys_out, n, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mb200_auc_rank_sum(const uint32_t* sorted_keys, int64_t n_sorted, const int64_t* n_pos_local, const uint32_t* pos_keys,
                       int64_t pos_capacity, const int64_t* n_pos, uint64_t* sum2, void* stream) {
  if (n_sorted < 0 || pos_capacity < 0 || !n_pos_local || !n_pos || !sum2) return MB200_ERR_INVALID_ARG;
  if (pos_capacity > 0 && !pos_keys) return MB200_ERR_INVALID_ARG;
  if (n_sorted > 0 && !sorted_keys) return MB200_ERR_INVALID_ARG;
  int st = use_device_of(sum2, nullptr);
  if (st != MB200_OK) return st;
  return auc_rank_sum(sorted_keys, n_sorted, reinterpret_cast<const long long*>(n_pos_local), pos_keys, pos_capacity,
                      reinterpret_cast<const long long*>(n_pos), reinterpret_cast<unsigned long long*>(sum2),
                      static_cast<cudaStream_t>(stream));
}

size_t mb200_pooled_auc_workspace_bytes(int64_t n) { return n < 0 ? 0 : pooled_auc_workspace_bytes(n); }

int mb200_pooled_auc(const float* preds, const uint8_t* labels, int64_t n, int sigmoid_mode, const int32_t* flags, void* workspace,
                     size_t workspace_bytes, double* out, void* stream) {
  if (n < 0 || out == nullptr || (n > 0 && (!preds || !labels))) return MB200_ERR_INVALID_ARG;
  if (sigmoid_mode < 0 || sigmoid_mode > 2 || (sigmoid_mode == 2 && flags == nullptr)) return MB200_ERR_INVALID_ARG;
  int st = use_device_of(out, nullptr);
  if (st != MB200_OK) return st;
  return pooled_auc(preds, labels, n, sigmoid_mode, flags, workspace, workspace_bytes, out, static_cast<cudaStream_t>(stream));
}

size_t mb200_pooled_auc_bounded_workspace_bytes(int64_t n, int64_t pos_capacity) {
  return (n < 0 || pos_capacity < 0) ? 0 : pooled_auc_bounded_workspace_bytes(n, pos_capacity);
}

int mb200_pooled_auc_bounded(const float* preds, const uint8_t* labels, int64_t n, int64_t pos_capacity, int sigmoid_mode, const int32_t* flags,
                             void* workspace, size_t workspace_bytes, double* out, void* stream) {
  if (n < 0 || pos_capacity < 0 || out == nullptr || (n > 0 && (!preds || !labels))) return MB200_ERR_INVALID_ARG;
  if (sigmoid_mode < 0 || sigmoid_mode > 2 || (sigmoid_mode == 2 && flags == nullptr)) return MB200_ERR_INVALID_ARG;
  int st = use_device_of(out, nullptr);
  if (st != MB200_OK) return st;
  return pooled_auc_bounded(preds, labels, n, pos_capacity, sigmoid_mode, flags, workspace, workspace_bytes, out, static_cast<cudaStream_t>(stream));
}

size_t mb200_retrieval_workspace_bytes(const mb200_retrieval_desc* desc) { return retrieval_workspace_bytes(desc); }

int mb200_retrieve_topk(const mb200_retrieval_desc* desc, void* stream) { return retrieve_topk(desc, static_cast<cudaStream_t>(stream)); }

int mb200_pool_users(const void* table, int dtype, int dim, int64_t row_stride, int64_t n_news, const int32_t* hist_offsets,
                     const int32_t* hist_ids, int64_t n_users, void* out_bf16, int32_t* flags, void* stream) {
  return pool_users(table, dtype, dim, row_stride, n_news, hist_offsets, hist_ids, n_users, out_bf16, flags, static_cast<cudaStream_t>(stream));
}

int mb200_merge_topk(const float* scores, const int64_t* ids, int shards, int64_t n_users, int k, float* out_scores, int64_t* out_ids,
                     void* stream) {
  return merge_topk(scores, reinterpret_cast<const long long*>(ids), shards, n_users, k, out_scores, reinterpret_cast<long long*>(out_ids),
                    static_cast<cudaStream_t>(stream));
}

int mb200_attention_logits(const void* table, int dtype, int dim, int64_t row_stride, int64_t n_rows, const float* weight, const float* bias,
                           const float* query, int q_dim, float* out, void* stream) {
  return attention_logits(table, dtype, dim, row_stride, n_rows, weight, bias, query, q_dim, out, static_cast<cudaStream_t>(stream));
}

int mb200_step_loss(const float* loss_per_impression, int64_t n_impressions, int step, int loss_kind, const int32_t* cand_offsets,
                    const uint8_t* labels, double* out, void* stream) {
  return step_loss(loss_per_impression, n_impressions, step, loss_kind, cand_offsets, labels, out, static_cast<cudaStream_t>(stream));
}

size_t mb200_metrics_workspace_bytes(const mb200_metrics_desc* desc) { return metrics_workspace_bytes(desc); }

int mb200_rank_metrics(const mb200_metrics_desc* desc, void* stream) { return rank_metrics(desc, static_cast<cudaStream_t>(stream)); }

size_t mb200_exchange_mailbox_bytes(int n_ranks, int n_payload, int64_t pos_capacity) { return exchange_mailbox_bytes(n_ranks, n_payload, pos_capacity); }

size_t mb200_exchange_workspace_bytes(int64_t n_rows) { return exchange_workspace_bytes(n_rows); }

int mb200_exchange_post(const mb200_exchange_desc* desc, void* stream) { return exchange_post(desc, static_cast<cudaStream_t>(stream)); }

int mb200_exchange_finish(const mb200_exchange_desc* desc, void* stream) { return exchange_finish(desc, static_cast<cudaStream_t>(stream)); }

int mb200_read_probe(const void* buf, size_t bytes, int repeats, int mode, void* sink, void* stream) {
  return read_probe(buf, bytes, repeats, mode, sink, static_cast<cudaStream_t>(stream));
}

int mb200_enable_peer_access(int device, int peer) {
  int can = 0;
  int st = cuda_status(cudaDeviceCanAccessPeer(&can, device, peer), "cudaDeviceCanAccessPeer");
  if (st != MB200_OK) return st;
  if (!can) return MB200_ERR_UNSUPPORTED;
  int prev = 0;
  cudaGetDevice(&prev);
  if ((st = cuda_status(cudaSetDevice(device), "cudaSetDevice")) != MB200_OK) return st;
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(prev);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return MB200_OK;
  }
  return cuda_status(e, "cudaDeviceEnablePeerAccess");
}

int mb200_ipc_export(const void* ptr, unsigned char handle[64], int64_t* offset) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle || !offset) return MB200_ERR_INVALID_ARG;
  int st = use_device_of(ptr, nullptr);
  if (st != MB200_OK) return st;
  typedef int (*RangeFn)(unsigned long long*, size_t*, unsigned long long);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || sym == nullptr)
    return MB200_ERR_CUDA;
  unsigned long long base = 0;
  size_t size = 0;
  if (reinterpret_cast<RangeFn>(sym)(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) return MB200_ERR_CUDA;
  cudaIpcMemHandle_t h;
  st = cuda_status(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)), "cudaIpcGetMemHandle");
  if (st != MB200_OK) return st;
  memcpy(handle, &h, 64);
  *offset = (int64_t)((unsigned long long)(uintptr_t)ptr - base);
  return MB200_OK;
}

int mb200_ipc_open(const unsigned char handle[64], int64_t offset, int device, void** out_ptr) {
  if (!handle || !out_ptr || offset < 0) return MB200_ERR_INVALID_ARG;
  int prev = 0;
  cudaGetDevice(&prev);
  int st = cuda_status(cudaSetDevice(device), "cudaSetDevice");
  if (st != MB200_OK) return st;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* base = nullptr;
  st = cuda_status(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  cudaSetDevice(prev);
  if (st != MB200_OK) return st;
  *out_ptr = static_cast<unsigned char*>(base) + offset;
  return MB200_OK;
}

int mb200_ipc_close(void* mapped_base, int device) {
  if (!mapped_base) return MB200_ERR_INVALID_ARG;
  int prev = 0;
  cudaGetDevice(&prev);
  int st = cuda_status(cudaSetDevice(device), "cudaSetDevice");
  if (st != MB200_OK) return st;
  st = cuda_status(cudaIpcCloseMemHandle(mapped_base), "cudaIpcCloseMemHandle");
  cudaSetDevice(prev);
  return st;
}

float mb200_dcg_discount(int rank) { return host_dcg_discount(rank); }
int64_t mb200_launch_count(void) { return g_launches.load(); }
int64_t mb200_library_launch_count(void) { return g_library_launches.load(); }

float mb200_last_score_kernel_ms(void) { return last_score_kernel_ms(); }

float mb200_last_score_kernel_begin_after(void* event) { return last_score_kernel_begin_after(static_cast<cudaEvent_t>(event)); }

int mb200_last_hot_stats(int32_t out[4]) { return out ? last_hot_stats(out) : MB200_ERR_INVALID_ARG; }

int mb200_set_tuning(int key, int value) {
  int prev = -1;
  switch (key) {
    case 0: prev = tuning().chunks_per_warp, tuning().chunks_per_warp = value; break;
    case 1: prev = tuning().variant, tuning().variant = value; break;
    case 2: prev = tuning().ctas_per_sm, tuning().ctas_per_sm = value; break;
    case 3: prev = tuning().time_kernel, tuning().time_kernel = value; break;
    case 4: prev = tuning().retrieval_diag, tuning().retrieval_diag = value; break;
    case 5: prev = tuning().retrieval_pair, tuning().retrieval_pair = value; break;
    case 6: prev = tuning().hot_kb_cap, tuning().hot_kb_cap = value; break;
    case 7: prev = tuning().static_chunks, tuning().static_chunks = value; break;
    case 8: prev = tuning().retrieval_window, tuning().retrieval_window = value; break;
  }
  return prev;
}

}  // extern "C"
