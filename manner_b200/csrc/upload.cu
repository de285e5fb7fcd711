// Pipelined host -> device upload of one behaviour set (mb200_upload_begin / mb200_upload_finish, include/manner_b200.h).
//
// The reference moves every batch to the device before its forward pass (Lightning's transfer_batch_to_device in front of
// cr_module.py:105 / ensemble_module.py:95); the epoch-granular path here has ONE 20 MB CSR set per pass, and copying it in
// front of the fused kernel costs 0.4 ms of a 2.2 ms end-to-end step.  Instead the offsets go first, the persistent kernel is
// launched at once, and the id / label arrays follow in segments of geometrically growing size on a copy stream; after each
// segment a 4-byte copy raises the device word `ready` to the number of leading impressions whose rows are resident.  The
// kernel's warps draw small chunks in impression order and wait on `ready` before they touch one, so the copy engine stays ahead
// of the SMs and only the offsets and the small first segment are exposed.
//
// Copies cover disjoint 128-byte aligned element ranges, so a 32-byte sector never holds bytes of two copies: a warp that
// reads up to the end of segment s cannot pull not-yet-written bytes of segment s + 1 into its L1.

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace mb200 {

int force_load_eval_kernels();               // score_eval.cu
int force_load_loss_kernels();               // attention.cu
int force_load_auc_kernels(cudaStream_t s);  // pooled_auc.cu

// once per device and process: see force_load_eval_kernels
static int force_load_once(int device, cudaStream_t stream) {
  static bool loaded[64] = {false};
  if (device < 0 || device >= 64 || loaded[device]) return MB200_OK;
  int st = force_load_eval_kernels();
  if (st == MB200_OK) st = force_load_loss_kernels();
  if (st == MB200_OK) st = force_load_auc_kernels(stream);
  if (st == MB200_OK) loaded[device] = true;
  return st;
}

struct UploadEvents {
  cudaEvent_t reset = nullptr, offsets = nullptr;
};
static UploadEvents g_upload_events[64];

static int upload_events_for(int device, UploadEvents** out) {
  if (device < 0 || device >= 64) return MB200_ERR_UNSUPPORTED;
  UploadEvents* e = &g_upload_events[device];
  if (e->reset == nullptr) {
    int st = cuda_status(cudaEventCreateWithFlags(&e->reset, cudaEventDisableTiming), "cudaEventCreateWithFlags");
    if (st != MB200_OK) return st;
    st = cuda_status(cudaEventCreateWithFlags(&e->offsets, cudaEventDisableTiming), "cudaEventCreateWithFlags");
    if (st != MB200_OK) return st;
  }
  *out = e;
  return MB200_OK;
}

static int validate_upload(const mb200_upload_desc* d) {
  if (d == nullptr || d->struct_size != sizeof(mb200_upload_desc)) return MB200_ERR_INVALID_ARG;
  if (d->n_segments < 1 || d->n_segments > MB200_MAX_UPLOAD_SEGMENTS || d->segments_first < 0 || d->segments_first > d->n_segments) return MB200_ERR_INVALID_ARG;
  if (d->n_impressions < 1 || d->n_impressions > 0x7ffffff0ll) return MB200_ERR_INVALID_ARG;
  if (!d->h_hist_offsets || !d->h_hist_ids || !d->h_cand_offsets || !d->h_cand_ids || !d->h_labels) return MB200_ERR_INVALID_ARG;
  if (!d->d_hist_offsets || !d->d_hist_ids || !d->d_cand_offsets || !d->d_cand_ids || !d->d_labels) return MB200_ERR_INVALID_ARG;
  if ((d->h_hist_pad == nullptr) != (d->d_hist_pad == nullptr) || (d->h_cand_pad == nullptr) != (d->d_cand_pad == nullptr)) return MB200_ERR_INVALID_ARG;
  if (!d->ready || ((uintptr_t)d->ready & 3) || !d->h_marks || !d->copy_stream) return MB200_ERR_INVALID_ARG;
  // the aligned-range argument above needs the destinations themselves on 128-byte boundaries
  if (((uintptr_t)d->d_hist_ids & 127) || ((uintptr_t)d->d_cand_ids & 127) || ((uintptr_t)d->d_labels & 127)) return MB200_ERR_INVALID_ARG;
  return MB200_OK;
}

// End of segment s (1-based) in impressions.  Segments grow geometrically -- the first holds 1/2^(S-1) of the work (rows gathered
// + 4 per impression, the partition_kernel measure), every later one as much as all before it -- so the fused kernel waits for a
// few hundred KB before its first chunk, and from then on each copy has the whole time the SMs spend on the data already there.
static long long segment_end(const mb200_upload_desc* d, int s) {
  const long long n = d->n_impressions;
  if (s >= d->n_segments) return n;
  const long long total = (long long)d->h_hist_offsets[n] + d->h_cand_offsets[n] + 4 * n;
  const long long target = total >> (d->n_segments - s);
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)d->h_hist_offsets[mid] + d->h_cand_offsets[mid] + 4 * mid >= target) hi = mid; else lo = mid + 1;
  }
  return lo;
}

static long long round_up(long long x, long long a) { return (x + a - 1) / a * a; }

static int copy_segments(const mb200_upload_desc* d, int first, int last) {
  const long long n = d->n_impressions;
  cudaStream_t cs = static_cast<cudaStream_t>(d->copy_stream);
  const long long nh = d->h_hist_offsets[n], nc = d->h_cand_offsets[n];
  long long b_prev = first > 0 ? segment_end(d, first) : 0;
  long long h_prev = first > 0 ? (first >= d->n_segments ? nh : std::min(round_up(d->h_hist_offsets[b_prev], 32), nh)) : 0;
  long long c_prev = first > 0 ? (first >= d->n_segments ? nc : std::min(round_up(d->h_cand_offsets[b_prev], 128), nc)) : 0;
  for (int s = first; s < last; ++s) {
    const long long b = segment_end(d, s + 1);
    const long long h_end = (s + 1 >= d->n_segments) ? nh : std::min(round_up(d->h_hist_offsets[b], 32), nh);
    const long long c_end = (s + 1 >= d->n_segments) ? nc : std::min(round_up(d->h_cand_offsets[b], 128), nc);
    int st = MB200_OK;
    if (h_end > h_prev)
      st = cuda_status(cudaMemcpyAsync(d->d_hist_ids + h_prev, d->h_hist_ids + h_prev, (size_t)(h_end - h_prev) * 4, cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(hist_ids)");
    if (st == MB200_OK && c_end > c_prev) {
      st = cuda_status(cudaMemcpyAsync(d->d_cand_ids + c_prev, d->h_cand_ids + c_prev, (size_t)(c_end - c_prev) * 4, cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(cand_ids)");
      if (st == MB200_OK)
        st = cuda_status(cudaMemcpyAsync(d->d_labels + c_prev, d->h_labels + c_prev, (size_t)(c_end - c_prev), cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(labels)");
    }
    if (st != MB200_OK) return st;
    d->h_marks[s] = (uint32_t)b;
    st = cuda_status(cudaMemcpyAsync(d->ready, d->h_marks + s, 4, cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(ready)");
    if (st != MB200_OK) return st;
    h_prev = h_end, c_prev = c_end;
  }
  return MB200_OK;
}

// The segments that mb200_upload_begin does not queue itself (segments_first < n_segments) are queued by ONE library thread,
// concurrently with the caller -- who goes straight on to launch the fused kernel.  A blocking launch on the caller's thread
// (CUDA_LAUNCH_BLOCKING=1) does not stop that thread; a tool that serialises every CUDA call behind a running kernel does (ncu's
// launch-list pass starved it until the kernel's 4 s time-out), which is why the Python layer queues everything itself by default.
struct UploadJob {
  mb200_upload_desc d;
  int first, last, device;
  unsigned long long ticket;
};

class UploadWorker {
 public:
  UploadWorker() : thread_([this] { run(); }) { thread_.detach(); }
  unsigned long long submit(const UploadJob& job) {
    std::lock_guard<std::mutex> lock(m_);
    UploadJob j = job;
    j.ticket = ++issued_;
    q_.push_back(j);
    cv_.notify_one();
    return j.ticket;
  }
  // blocks until every job submitted so far has been queued on its copy stream; returns the first error among them
  int drain() {
    std::unique_lock<std::mutex> lock(m_);
    done_cv_.wait(lock, [this] { return done_ == issued_; });
    const int st = status_;
    status_ = MB200_OK;
    return st;
  }

 private:
  void run() {
    for (;;) {
      UploadJob job;
      {
        std::unique_lock<std::mutex> lock(m_);
        cv_.wait(lock, [this] { return !q_.empty(); });
        job = q_.front();
        q_.pop_front();
      }
      int st = cuda_status(cudaSetDevice(job.device), "cudaSetDevice");
      if (st == MB200_OK) st = copy_segments(&job.d, job.first, job.last);
      {
        std::lock_guard<std::mutex> lock(m_);
        if (st != MB200_OK && status_ == MB200_OK) status_ = st;
        done_ = job.ticket;
      }
      done_cv_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  std::deque<UploadJob> q_;
  unsigned long long issued_ = 0, done_ = 0;
  int status_ = MB200_OK;
  std::thread thread_;
};

static UploadWorker& upload_worker() {
  static UploadWorker* w = new UploadWorker();  // never destroyed: the thread sleeps on its queue until the process ends
  return *w;
}

int upload_begin(const mb200_upload_desc* d, cudaStream_t compute) {
  int st = validate_upload(d);
  if (st != MB200_OK) return st;
  if (static_cast<cudaStream_t>(d->copy_stream) == compute) return MB200_ERR_INVALID_ARG;
  int device = 0;
  if ((st = use_device_of(d->ready, &device)) != MB200_OK) return st;
  UploadEvents* ev = nullptr;
  if ((st = upload_events_for(device, &ev)) != MB200_OK) return st;
  if ((st = force_load_once(device, compute)) != MB200_OK) return st;
  cudaStream_t cs = static_cast<cudaStream_t>(d->copy_stream);
  const long long n = d->n_impressions;
  // nothing is "ready" until this pass's copies say so; the copies start after everything already queued on the compute stream
  // (the previous pass may still read the buffers these destinations were recycled from)
  if ((st = cuda_status(cudaMemsetAsync(d->ready, 0, 4, compute), "cudaMemsetAsync(ready)")) != MB200_OK) return st;
  if ((st = cuda_status(cudaEventRecord(ev->reset, compute), "cudaEventRecord")) != MB200_OK) return st;
  if ((st = cuda_status(cudaStreamWaitEvent(cs, ev->reset, 0), "cudaStreamWaitEvent")) != MB200_OK) return st;
  const size_t off_bytes = (size_t)(n + 1) * 4;
  if ((st = cuda_status(cudaMemcpyAsync(d->d_hist_offsets, d->h_hist_offsets, off_bytes, cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(hist_offsets)")) != MB200_OK) return st;
  if ((st = cuda_status(cudaMemcpyAsync(d->d_cand_offsets, d->h_cand_offsets, off_bytes, cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(cand_offsets)")) != MB200_OK) return st;
  if (d->h_hist_pad && (st = cuda_status(cudaMemcpyAsync(d->d_hist_pad, d->h_hist_pad, (size_t)n * 4, cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(hist_pad)")) != MB200_OK) return st;
  if (d->h_cand_pad && (st = cuda_status(cudaMemcpyAsync(d->d_cand_pad, d->h_cand_pad, (size_t)n * 4, cudaMemcpyHostToDevice, cs), "cudaMemcpyAsync(cand_pad)")) != MB200_OK) return st;
  // the compute stream may run the partition + the fused kernel as soon as the offsets are in
  if ((st = cuda_status(cudaEventRecord(ev->offsets, cs), "cudaEventRecord")) != MB200_OK) return st;
  if ((st = cuda_status(cudaStreamWaitEvent(compute, ev->offsets, 0), "cudaStreamWaitEvent")) != MB200_OK) return st;
  if ((st = copy_segments(d, 0, d->segments_first)) != MB200_OK) return st;
  if (d->segments_first < d->n_segments) {
    UploadJob job;
    job.d = *d, job.first = d->segments_first, job.last = d->n_segments, job.device = device, job.ticket = 0;
    upload_worker().submit(job);
  }
  return MB200_OK;
}

int upload_finish(const mb200_upload_desc* d) {
  if (d == nullptr || d->struct_size != sizeof(mb200_upload_desc)) return MB200_ERR_INVALID_ARG;
  return upload_worker().drain();
}

}  // namespace mb200
