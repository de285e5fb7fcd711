// Shared host/device helpers of libmanner_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/manner_b200.h"

namespace mb200 {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr unsigned kFull = 0xffffffffu;

// 1 / log2(rank + 1) in fp32, exactly the values `1 / torch.log2(arange(k) + 2.0)` produces on the
// reference's CPU path (torchmetrics 0.11.4 `_dcg`); index = rank - 1.  Pinned by tests/test_abi.py.
#define MB200_INV_DISC_TABLE                                                                              \
  0x1.000000p+0f, 0x1.43093ap-1f, 0x1.000000p-1f, 0x1.b90348p-2f, 0x1.8c2324p-2f, 0x1.6cc194p-2f,         \
      0x1.555556p-2f, 0x1.43093ap-2f, 0x1.344136p-2f, 0x1.28009cp-2f, 0x1.1da338p-2f, 0x1.14b950p-2f,     \
      0x1.0cf400p-2f, 0x1.0619dcp-2f, 0x1.000000p-2f, 0x1.f50b58p-3f, 0x1.eb22cap-3f, 0x1.e21e10p-3f,     \
      0x1.d9dcd2p-3f, 0x1.d244c8p-3f, 0x1.cb4058p-3f, 0x1.c4bd96p-3f, 0x1.bead78p-3f, 0x1.b90348p-3f,     \
      0x1.b3b432p-3f, 0x1.aeb6f6p-3f, 0x1.aa038ep-3f, 0x1.a59304p-3f, 0x1.a15f4ep-3f, 0x1.9d630ep-3f,     \
      0x1.99999ap-3f, 0.0f

struct Tuning {
  int chunks_per_warp = 0;  // work-balanced impression chunks per resident warp of the fused kernel (0 = default: 1 static, 16 dynamic)
  int static_chunks = 0;    // chunk schedule of the fused kernel: 0 = dynamic hand-out iff the behaviours come through a pipelined upload,
                            // 1 = always static round-robin, 2 = always dynamic
  int variant = -1;         // reference-width kernel: -1 = by table type (fp32: 2, bf16: 3), 0 = 4 rows in flight / 3 CTAs per SM, 1 = same with L1::no_allocate loads,
                            // 2 = 3 rows / 4 CTAs (default: best on Zipf-shaped ids), 3 = 2 rows / 5 CTAs
  int ctas_per_sm = 0;      // CTAs (of kWarpsPerCta warps) per SM; 0 = as many as are resident (occupancy query)
  int time_kernel = 0;      // 1: bracket the fused kernel with CUDA events (mb200_last_score_kernel_ms)
  int hot_kb_cap = 0;       // hot-row cache variants: cap on the cache size in KB (0 = all the shared memory the per-warp areas leave)
  int retrieval_pair = 1;   // retrieve_topk_kernel: 1 (default) = CTA pairs (tcgen05 cta_group::2, UMMA M = 256) when there are >= 2 user tiles
  int retrieval_window = 48;  // retrieve_topk_kernel: catalogue tiles a CTA may run ahead of the slowest CTA of the sweep (0 = unthrottled)
  int retrieval_diag = 0;   // DIAGNOSTIC ONLY (results invalid when != 0): 1 = epilogue reads TMEM but selects nothing, 2 = epilogue only
                            // releases the accumulator (isolates the TMA + MMA pipeline when tuning retrieve_topk_kernel)
};
Tuning& tuning();

// records the CUDA error text for mb200_last_cuda_error(); returns MB200_ERR_CUDA or MB200_OK
int cuda_status(cudaError_t e, const char* what);
void note_launch(int n = 1);
void note_library_launch(int n = 1);

// Makes the device that owns `ptr` current for this thread; MB200_OK or an error code.
int use_device_of(const void* ptr, int* device_out);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

}  // namespace mb200
