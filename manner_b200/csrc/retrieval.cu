// Full-catalog retrieval (BASELINE.json configs[4], SURVEY 8(d) mode R): every user's late-fusion vector
// against every news row of the catalogue, top-k per user.  No reference counterpart (the reference only
// scores the ~37 candidates of an impression, models/cr_module.py:105-131); this is the one part of the
// path that really is a dense users x catalogue x D contraction, so it runs on the 5th-generation tensor
// cores:
//
//   TMA (cp.async.bulk.tensor.2d, 128B swizzle)  ->  4-stage shared-memory ring (A 128x64, B 256x64 bf16)
//   tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16), one elected thread, accumulators in TMEM
//   (2 x 256 columns, double buffered)  ->  tcgen05.ld in 4 epilogue warps  ->  per-row running top-k
//
// The U x N score matrix is never written: an epilogue thread owns one user row for the whole sweep over
// the catalogue, keeps the row's current k-th best score in a register, appends the (few) scores above it
// to a per-row candidate buffer and lets its warp compact the buffer to the best k when it fills up.
// Final order: score descending, catalogue id ascending on ties.
//
// mb200_pool_users builds the user matrix: mean of the history rows (cr_module.py:116-123), rounded to bf16.

#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"

namespace mb200 {

namespace rt {
constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
// One CTA per tile (cta_group::1): B = 256 catalogue rows per stage, 4 stages of 48 KB.  CTA pair (cta_group::2, UMMA M = 256):
// every CTA stages its own 128 users and HALF of the catalogue tile, 6 stages of 32 KB -- a third less L2 -> SM traffic and half
// the shared-memory reads of B per flop.
template <bool PAIR> struct Ring {
  static constexpr int STAGES = PAIR ? 6 : 4;
  static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};
constexpr int RING_BYTES = 4 * (A_BYTES + BLOCK_N * BLOCK_K * 2);  // 192 KB either way
static_assert(Ring<false>::STAGES * Ring<false>::STAGE_BYTES == RING_BYTES && Ring<true>::STAGES * Ring<true>::STAGE_BYTES == RING_BYTES, "ring size");
constexpr int MAX_STAGES = 6;
constexpr int CAP = 256;          // per-row candidate buffer entries (hard limit: a row is compacted before a chunk could overflow it)
constexpr int SOFT_CAP = 160;     // soft limit (> MAX_K): from here on a row is compacted between tiles, one row per tile and warp
constexpr int MAX_K = 128;        // top-k limit (k <= CAP / 2)
constexpr int EPI_GROUPS = 2;     // epilogue warpgroups: group g scans columns 128 g .. 128 g + 127 of every accumulator tile
constexpr int EPI_COLS = BLOCK_N / EPI_GROUPS;
constexpr int WS_HEADER = 1024;  // workspace header: error flag, diagnostic counters, per-CTA sweep progress [160]
constexpr int THREADS = 64 + EPI_GROUPS * 128;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int TMEM_COLS = 512;    // two 128 x 256 fp32 accumulators
constexpr int SCRATCH_BYTES = 0;
constexpr int SHARE_BYTES = EPI_GROUPS * BLOCK_M * 8;  // per (group, row): published admission threshold + list length
constexpr int SMEM_BYTES = 1024 /*alignment slack*/ + RING_BYTES + SHARE_BYTES + 256 /*barriers*/;
constexpr unsigned long long WAIT_LIMIT_NS = 2000ull * 1000 * 1000;
constexpr uint32_t WAIT_HINT_NS = 20000;  // upper bound of one hardware-suspended try_wait
}  // namespace rt

struct RetrievalParams {
  float* out_scores;       // [n_users, k]
  long long* out_ids;      // [n_users, k]
  float* cand_scores;      // workspace [grid][EPI_GROUPS][128][CAP]
  int* cand_ids;           // workspace [grid][EPI_GROUPS][128][CAP]
  float* debug_scores;     // optional [n_users, n_catalog]
  int* error_flag;
  int* progress;           // workspace [grid]: catalogue tiles this CTA's producer has issued so far (sweep throttle, see the TMA producer)
  int window;              // a producer may run at most this many catalogue tiles ahead of the slowest CTA (0 = no throttle)
  unsigned long long* stats;  // diag == 4: [0] cycles in compactions, [1] in the admission slow path, [2] waiting for an accumulator,
                              // [3] compactions, [4] slow chunks, [5] epilogue cycles in total (summed over epilogue warps, lane 0);
                              // [6] / [7] cycles the MMA thread waited for a free accumulator / for operands (summed over CTAs)
  long long n_users, n_catalog, id_offset;
  int dim, k, m_tiles, n_tiles;
  int diag;  // Tuning::retrieval_diag
  // fused exchange: gather buffers of the peer GPUs ([n_peers][peer_rows][k], this rank writes slot my_rank); 0 = off
  int n_peers, my_rank;
  long long peer_rows;
  float* peer_scores[MB200_MAX_TABLE_SHARDS];
  long long* peer_ids[MB200_MAX_TABLE_SHARDS];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must end in a trap (a reported CUDA error), never in a hung GPU.  No legitimate
// wait in this kernel is longer than one tile's MMA time (microseconds); the bound is 2 s of %globaltimer.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// `try_wait` with a suspend-time hint parks the thread in hardware until the phase completes (or the hint expires), so
// the ten mostly-waiting warps of this kernel do not burn issue slots and power spinning -- the kernel runs under the
// 1 kW power cap, where every wasted instruction lowers the tensor-core clock.
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag) {
  const uint32_t addr = smem_u32(bar);
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(rt::WAIT_HINT_NS)
        : "memory");
    if (done) return;
    if ((spin & 63u) == 63u) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > rt::WAIT_LIMIT_NS) {
        if (error_flag) atomicExch(error_flag, 1);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// CTA-pair variants.  In a cluster the 32-bit shared-window address carries the CTA's rank within the pair in bit 24; clearing it
// addresses the same offset in the leader CTA (rank 0) -- the convention of CUTLASS' SM100_TMA_2SM_LOAD (Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}

// K-major operand tile in shared memory with 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B
// apart.  Descriptor: start >> 4, LBO = 1 (unused for swizzled K-major), SBO = 1024 >> 4, version 1 (sm_100),
// layout SWIZZLE_128B.  Advancing by one UMMA_K (16 bf16 = 32 B) adds 2 to the start field.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// kind::f16 instruction descriptor: D fp32, A and B bf16, both K-major, N = 256, M = 128.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((rt::BLOCK_N >> 3) << 17) | ((rt::BLOCK_M >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
// CTA pair: one instruction drives both SMs' tensor cores (M = 256: 128 rows of A and 128 of the 256 rows of B from each CTA's
// shared memory, 128 x 256 accumulator rows into each CTA's TMEM); issued by the leader CTA only.
constexpr uint32_t kIdescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((rt::BLOCK_N >> 3) << 17) | (((2 * rt::BLOCK_M) >> 4) << 24);
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdescPair), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((unsigned short)3)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define MB200_R32(r) \
  r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], r[16], r[17], r[18], r[19], r[20], r[21], \
      r[22], r[23], r[24], r[25], r[26], r[27], r[28], r[29], r[30], r[31]
// Asynchronous TMEM -> register load of 32 consecutive columns of this thread's lane; the registers are valid only after
// tmem_ld_wait on the same array (which names them as in/out operands so no use can be scheduled above the wait).
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// ---- per-row candidate lists -------------------------------------------------------------------------
// (score, id) ordering of the final list: higher score first, lower id first on ties.  Scores are compared
// through an order-preserving uint32 key (s + 0.0f folds -0.0 into +0.0 so the key order equals the float order).
__device__ __forceinline__ uint32_t score_key(float s) {
  const uint32_t b = __float_as_uint(s + 0.0f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

constexpr int kPerLane = rt::CAP / 32;  // candidate entries a lane holds during a compaction

// Largest T with |{key >= T}| >= kk over the warp's 32 x kPerLane keys (kk >= 1; absent entries are 0 and never
// count): the kk-th largest key, found bit by bit -- one round = kPerLane compares + one REDUX.  The bits all present keys
// share (sign, exponent and a few mantissa bits of an already filtered list) are skipped, and with max_rounds < 32 the
// search stops early: T is then a lower bound of the kk-th largest key with its remaining low bits zero.
__device__ __forceinline__ uint32_t kth_largest(const uint32_t (&key)[kPerLane], int kk, int max_rounds = 32) {
  uint32_t all_or = 0, all_and = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) all_or |= key[j], all_and &= key[j] ? key[j] : 0xffffffffu;
  all_or = __reduce_or_sync(kFull, all_or), all_and = __reduce_and_sync(kFull, all_and);
  const uint32_t diff = all_or ^ all_and;
  if (diff == 0) return all_or;  // all present keys are equal
  const int hi = 31 - __clz(diff);
  uint32_t t = all_or & ~((2u << hi) - 1u);  // the common prefix (hi == 31: 2u << 31 == 0 wraps to an all-ones mask, t = 0)
  const int lo = max(0, hi - max_rounds + 1);
#pragma unroll 4
  for (int b = hi; b >= lo; --b) {
    const uint32_t trial = t | (1u << b);
    int c = 0;
#pragma unroll
    for (int j = 0; j < kPerLane; ++j) c += key[j] >= trial ? 1 : 0;
    if ((int)__reduce_add_sync(kFull, (unsigned)c) >= kk) t = trial;
  }
  return t;
}

// Warp-cooperative compaction of one row's candidate buffer (n <= CAP entries, unsorted, in global memory).
//   exact:   keeps exactly min(n, k) best entries (ties at the boundary resolved towards the lower id), still unsorted,
//            at the front; returns the k-th best score (-inf while fewer than k entries exist).
//   !exact:  the k-th largest key is only located to kApproxRounds bits below the keys' common prefix; every entry >= that
//            lower bound is kept (k entries or a few more) and the bound is returned -- a valid, slightly conservative
//            admission threshold at less than half the latency.  Used between tiles; the hard limit and the end of the
//            sweep use the exact form.
// Returns the new length through *n_out.
constexpr int kApproxRounds = 12;
__device__ float compact_row(float* cs, int* ci, int n, int k, int lane, bool exact, int* n_out) {
  __syncwarp();  // the owning lane's appends to cs / ci become visible to the whole warp
  float sc[kPerLane];
  int id[kPerLane];
  uint32_t key[kPerLane];
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int t = lane + 32 * j;
    sc[j] = t < n ? cs[t] : 0.f;
    id[j] = t < n ? ci[t] : 0;
    key[j] = t < n ? score_key(sc[j]) : 0u;
  }
  *n_out = n;
  if (n <= k) return -CUDART_INF_F;  // nothing to drop (warp-uniform)
  const uint32_t kth = kth_largest(key, k, exact ? 32 : kApproxRounds);
  if (!exact && kth == 0u) return -CUDART_INF_F;  // no usable bound at this precision: leave the list as it is
  uint32_t id_floor = 0;  // on (0x80000000 - id): larger = lower id
  if (exact) {
    int c_gt = 0, c_eq = 0;
#pragma unroll
    for (int j = 0; j < kPerLane; ++j) c_gt += key[j] > kth ? 1 : 0, c_eq += key[j] == kth ? 1 : 0;
    c_gt = (int)__reduce_add_sync(kFull, (unsigned)c_gt);
    c_eq = (int)__reduce_add_sync(kFull, (unsigned)c_eq);
    const int need = k - c_gt;  // 1 <= need <= c_eq entries of score == kth survive: the ones with the lowest ids
    if (c_eq > need) {
      uint32_t rid[kPerLane];
#pragma unroll
      for (int j = 0; j < kPerLane; ++j) rid[j] = key[j] == kth ? 0x80000000u - (uint32_t)id[j] : 0u;
      id_floor = kth_largest(rid, need);
    }
  }
  __syncwarp();
  int base = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const bool keep = key[j] > kth || (key[j] == kth && 0x80000000u - (uint32_t)id[j] >= id_floor);
    const unsigned m = __ballot_sync(kFull, keep);
    if (keep) {
      const int pos = base + __popc(m & lt);
      cs[pos] = sc[j], ci[pos] = id[j];
    }
    base += __popc(m);
  }
  __syncwarp();
  *n_out = base;
  // approximate form: entries equal to the bound were kept, so the strict admission test must use the float just below it
  uint32_t bound = kth;
  if (!exact) {
    bound = kth - 1u;
    if (bound == 0x7fffffffu) bound = 0x7ffffffeu;  // skip -0.0 (compares equal to +0.0)
  }
  return key_score(bound);
}

// Final pass over one row: its (already compacted, n <= k <= 128) entries are ranked by counting -- (score desc,
// id asc) -- and written in order; slots n .. k-1 get (-inf, -1).
__device__ void write_sorted_row(const float* cs, const int* ci, int n, int k, float* out_s, long long* out_i, long long id_offset, int lane) {
  constexpr int PER = rt::MAX_K / 32;
  unsigned long long key[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int t = lane + 32 * j;
    // composite key: larger = better; ids are unique, so keys are distinct and the ranks a permutation
    key[j] = t < n ? ((unsigned long long)score_key(cs[t]) << 32) | (0xffffffffu - (uint32_t)ci[t]) : 0ull;
  }
  int rank[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) rank[j] = 0;
  for (int src = 0; src < 32; ++src) {
#pragma unroll
    for (int jj = 0; jj < PER; ++jj) {
      if (src + 32 * jj >= n) break;  // warp-uniform
      const unsigned long long other = __shfl_sync(kFull, key[jj], src);
#pragma unroll
      for (int j = 0; j < PER; ++j) rank[j] += other > key[j] ? 1 : 0;
    }
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int t = lane + 32 * j;
    if (t < n) {
      out_s[rank[j]] = key_score((uint32_t)(key[j] >> 32));
      out_i[rank[j]] = (long long)(0xffffffffu - (uint32_t)key[j]) + id_offset;
    } else if (t < k) {
      out_s[t] = -CUDART_INF_F, out_i[t] = -1ll;
    }
  }
}

template <bool PAIR>
__global__ void __launch_bounds__(rt::THREADS, 1)
retrieve_topk_kernel(const __grid_constant__ CUtensorMap tmap_users, const __grid_constant__ CUtensorMap tmap_catalog, const RetrievalParams p) {
  using namespace rt;
  constexpr int STAGES = Ring<PAIR>::STAGES, STAGE_BYTES = Ring<PAIR>::STAGE_BYTES;
  // tiles of this CTA: unit u = blockIdx.x / UNIT + j * (gridDim.x / UNIT); a unit is one user tile, or a pair of them for a CTA pair
  constexpr int UNIT = PAIR ? 2 : 1;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader of the pair
  const int unit0 = (int)blockIdx.x / UNIT, unit_step = (int)gridDim.x / UNIT;
  const int n_units = (p.m_tiles + UNIT - 1) / UNIT;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_base = smem;                                   // STAGES x (A | B), 1024-aligned
  float* thr_sh = reinterpret_cast<float*>(smem + RING_BYTES);  // [EPI_GROUPS][BLOCK_M] published admission thresholds
  int* cnt_sh = reinterpret_cast<int*>(thr_sh + EPI_GROUPS * BLOCK_M);     // [EPI_GROUPS][BLOCK_M] list lengths at the end of a sweep
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES + SHARE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_blocks = p.dim / BLOCK_K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
    // the leader's tmem_empty collects the epilogue warps of both CTAs of a pair
    for (int s = 0; s < 2; ++s) mbar_init(&tmem_full[s], 1), mbar_init(&tmem_empty[s], 4 * EPI_GROUPS * UNIT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = threadIdx.x; t < EPI_GROUPS * BLOCK_M; t += THREADS) thr_sh[t] = -CUDART_INF_F;
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // pair: the peer's barriers must exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    // Every CTA sweeps the whole catalogue for each of its user tiles, all in the same direction: a catalogue tile is read from
    // DRAM once and served to the other CTAs from the L2 -- as long as they are close behind.  Unthrottled, the 148 producers
    // drift apart over the 4 883 tiles of a sweep until the spread exceeds the 126 MB L2 and every CTA fetches its own copy
    // (21.5 GB of DRAM reads per launch against 1.9 GB of catalogue, round 1).  So every 8 tiles a producer publishes how many
    // tiles it has issued and waits (bounded, a soft throttle) while it is more than `window` tiles (~19 MB) ahead of the slowest
    // CTA; the slowest never waits, and equal work per CTA means the leaders would only have idled at the end anyway.
    int stage = 0;
    uint32_t phase = 0;
    int issued = 0;
    volatile int* progress = p.progress;
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int mt = unit * UNIT + (int)rank;  // a pair's second tile may lie past the last user tile: TMA zero-fills it
      for (int nt = 0; nt < p.n_tiles; ++nt, ++issued) {
        if (p.window > 0 && (issued & 7) == 0) {
          if (lane == 0) progress[blockIdx.x] = issued;
          for (int spin = 0; spin < 48; ++spin) {
            int slowest = 0x7fffffff;
            for (int c = lane; c < (int)gridDim.x; c += 32) slowest = min(slowest, progress[c]);
            slowest = __reduce_min_sync(kFull, slowest);
            if (issued - slowest <= p.window) break;
            __nanosleep(1000);
          }
        }
        if (lane == 0) {
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait<64>(&empty_bar[stage], phase ^ 1, p.error_flag);
            unsigned char* a = stage_base + stage * STAGE_BYTES;
            if constexpr (PAIR) {
              // both CTAs' bytes are counted on the LEADER's barrier (the MMA is issued there)
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
              tma_load_2d_pair(a, &tmap_users, &full_bar[stage], kb * BLOCK_K, mt * BLOCK_M);
              tma_load_2d_pair(a + A_BYTES, &tmap_catalog, &full_bar[stage], kb * BLOCK_K, nt * BLOCK_N + (int)rank * (BLOCK_N / 2));
            } else {
              mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
              tma_load_2d(a, &tmap_users, &full_bar[stage], kb * BLOCK_K, mt * BLOCK_M);
              tma_load_2d(a + A_BYTES, &tmap_catalog, &full_bar[stage], kb * BLOCK_K, nt * BLOCK_N);
            }
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) progress[blockIdx.x] = 0x7fffffff;  // done: nobody waits for this CTA any more
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && rank == 0) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      const bool st_on = p.diag == 4;
      long long w_acc = 0, w_feed = 0, c0 = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          if (st_on) c0 = clock64();
          mbar_wait(&tmem_empty[as], aphase ^ 1, p.error_flag);
          if (st_on) w_acc += clock64() - c0;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BLOCK_N;
          for (int kb = 0; kb < k_blocks; ++kb) {
            if (st_on) c0 = clock64();
            mbar_wait(&full_bar[stage], phase, p.error_flag);
            if (st_on) w_feed += clock64() - c0;
            tc_fence_after();
            const uint32_t a_addr = smem_u32(stage_base + stage * STAGE_BYTES);
            const uint64_t adesc = umma_desc(a_addr), bdesc = umma_desc(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              if constexpr (PAIR) umma_f16_pair(d_tmem, adesc + (uint64_t)(k * (UMMA_K * 2 / 16)), bdesc + (uint64_t)(k * (UMMA_K * 2 / 16)), (kb | k) != 0);
              else umma_f16(d_tmem, adesc + (uint64_t)(k * (UMMA_K * 2 / 16)), bdesc + (uint64_t)(k * (UMMA_K * 2 / 16)), (kb | k) != 0);
            }
            // frees the stage (in both CTAs of a pair) once these MMAs have read it
            if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
          if constexpr (PAIR) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]);  // accumulator complete
          if (++as == 2) as = 0, aphase ^= 1;
        }
      }
      if (st_on) atomicAdd(p.stats + 6, (unsigned long long)w_acc), atomicAdd(p.stats + 7, (unsigned long long)w_feed);
    }
  } else {
    // ===== epilogue: 2 warpgroups x (4 warps x 32 lanes = the 128 accumulator rows); warp w may touch TMEM lanes
    // 32 (w % 4) .. + 31.  Group g scans one half of the columns of EVERY tile: the MMA of tile t + 2 reuses the accumulator of
    // tile t, so a tile's epilogue has to finish within ONE tile time (the MMA of tile t + 1) -- splitting the columns halves
    // its latency and gives every scheduler two epilogue warps.  A user row therefore has one candidate list per group; the
    // groups publish their admission thresholds to each other (an item below EITHER group's k-th best can not be in the
    // row's top k) and the two lists are merged at the end of the sweep. =====
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const size_t list0 = ((size_t)blockIdx.x * EPI_GROUPS + g) * BLOCK_M;
    float* my_cs = p.cand_scores + (list0 + row_in_tile) * CAP;
    int* my_ci = p.cand_ids + (list0 + row_in_tile) * CAP;
    volatile float* thr_mine = thr_sh + g * BLOCK_M + row_in_tile;
    volatile float* thr_other = thr_sh + (g ^ 1) * BLOCK_M + row_in_tile;
    int tile = 0;  // tiles this CTA has seen so far: stage = tile & 1, phase = (tile >> 1) & 1
    const bool st_on = p.diag == 4;
    long long st_comp = 0, st_slow = 0, st_wait = 0, st_ncomp = 0, st_nslow = 0;
    const long long st_t0 = clock64();
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int mt = unit * UNIT + (int)rank;
      const long long user = (long long)mt * BLOCK_M + row_in_tile;
      const bool row_valid = user < p.n_users;
      float thr = -CUDART_INF_F;  // admission threshold: the best k-th best score either group has established for this row
      int cnt = 0;
      for (int nt = 0; nt < p.n_tiles; ++nt, ++tile) {
        const int as = tile & 1;
        long long c0 = st_on ? clock64() : 0;
        mbar_wait<20>(&tmem_full[as], (uint32_t)(tile >> 1) & 1u, p.error_flag);
        if (st_on) st_wait += clock64() - c0;
        tc_fence_after();
        {
          // The other group's k-th best bounds this row too, but only NON-strictly: its lists may already hold items with
          // HIGHER catalogue ids than the ones seen here (the groups run concurrently), and an equal score with a lower id
          // must still be admitted.  "s >= other" == "s > the float just below other".
          const float other = *thr_other;
          if (other > -CUDART_INF_F) {
            uint32_t kk = score_key(other) - 1u;
            if (kk == 0x7fffffffu) kk = 0x7ffffffeu;  // skip -0.0: it compares equal to +0.0
            thr = fmaxf(thr, key_score(kk));
          }
        }
        const bool last_tile = nt == p.n_tiles - 1;
        // make room for a whole 32-column chunk in every row of the warp before looking at it (hard limit; see SOFT_CAP below)
        auto make_room = [&]() {
          unsigned need = __ballot_sync(kFull, cnt > CAP - 32);
          if (need == 0) return;
          if (st_on) c0 = clock64(), st_ncomp += __popc(need);
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const int n_src = __shfl_sync(kFull, cnt, src);
            float* cs = reinterpret_cast<float*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_cs), src));
            int* ci = reinterpret_cast<int*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_ci), src));
            int n_new;
            const float kth = compact_row(cs, ci, n_src, p.k, lane, true, &n_new);
            if (lane == src) {
              thr = fmaxf(thr, kth), cnt = n_new;
              *thr_mine = thr;
            }
          }
          if (st_on) st_comp += clock64() - c0;
        };
        // one chunk: 32 consecutive scores of this thread's user row
        auto process = [&](uint32_t (&v)[32], int c) {
          const int col0 = nt * BLOCK_N + g * EPI_COLS + c * 32;
          if (p.debug_scores != nullptr && row_valid) {
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (col0 + t < p.n_catalog) p.debug_scores[user * p.n_catalog + col0 + t] = __uint_as_float(v[t]);
          }
          if (last_tile) {  // columns past the catalogue hold 0 (TMA zero fill): they must never be admitted
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (col0 + t >= p.n_catalog) v[t] = 0xff800000u;  // -inf
          }
          // two-level filter: the maxima of the four 8-column groups, then only the groups that can hold an admission
          float gm[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // three-input max (FMNMX3): 4 instructions per 8 columns instead of 7
            const float a0 = fmax3(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]), __uint_as_float(v[8 * q + 2]));
            const float a1 = fmax3(__uint_as_float(v[8 * q + 3]), __uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
            const float a2 = fmaxf(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
            gm[q] = fmax3(a0, a1, a2);
          }
          float mx = fmaxf(fmax3(gm[0], gm[1], gm[2]), gm[3]);
          if (p.diag == 1 || !row_valid) mx = -CUDART_INF_F;
          const bool any_slow = st_on && __any_sync(kFull, mx > thr);
          if (any_slow) c0 = clock64(), ++st_nslow;
          if (mx > thr) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (gm[q] > thr) {
#pragma unroll
                for (int t = 8 * q; t < 8 * q + 8; ++t) {
                  const float sc = __uint_as_float(v[t]);
                  if (sc > thr) my_cs[cnt] = sc, my_ci[cnt] = col0 + t, ++cnt;
                }
              }
            }
          }
          __syncwarp();
          if (any_slow) st_slow += clock64() - c0;
        };
        if (p.diag != 2) {
          // software pipeline over this group's 4 chunks of the tile: the TMEM load of chunk c + 1 is in flight while chunk c is scanned
          const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + g * EPI_COLS);
          uint32_t va[32], vb[32];
          tmem_ld32_issue(t0, va);
          tmem_ld_wait(va);
#pragma unroll 1
          for (int c = 0; c < EPI_COLS / 32; c += 2) {
            tmem_ld32_issue(t0 + (uint32_t)((c + 1) * 32), vb);
            make_room();
            process(va, c);
            tmem_ld_wait(vb);
            if (c + 2 < EPI_COLS / 32) tmem_ld32_issue(t0 + (uint32_t)((c + 2) * 32), va);
            make_room();
            process(vb, c + 1);
            if (c + 2 < EPI_COLS / 32) tmem_ld_wait(va);
          }
        }
        // accumulator drained: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_remote(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]);
        }
        // Off the critical path (the accumulator is already released): compact at most ONE row per tile, the fullest one
        // above the soft limit.  Rows of a warp fill up at similar times; compacting them all when they hit the hard limit
        // (above) would hold the accumulator stage for ~32 compactions and stall the MMA pipeline.
        {
          const int worst = __reduce_max_sync(kFull, cnt);
          if (worst > SOFT_CAP) {
            if (st_on) c0 = clock64(), ++st_ncomp;
            const int src = __ffs(__ballot_sync(kFull, cnt == worst)) - 1;
            float* cs = reinterpret_cast<float*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_cs), src));
            int* ci = reinterpret_cast<int*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_ci), src));
            int n_new;
            const float kth = compact_row(cs, ci, worst, p.k, lane, false, &n_new);
            if (lane == src) {
              thr = fmaxf(thr, kth), cnt = n_new;
              *thr_mine = thr;
            }
            if (st_on) st_comp += clock64() - c0;
          }
        }
      }
      // end of the sweep for this user tile: cut every list to its best k, merge the two groups' lists, write in order
      __syncwarp();
      for (int src = 0; src < 32; ++src) {
        const int n_src = __shfl_sync(kFull, cnt, src);
        if (n_src > p.k) {  // warp-uniform
          float* cs = reinterpret_cast<float*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_cs), src));
          int* ci = reinterpret_cast<int*>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_ci), src));
          int n_new;
          compact_row(cs, ci, n_src, p.k, lane, true, &n_new);
        }
      }
      cnt_sh[g * BLOCK_M + row_in_tile] = min(cnt, p.k);
      __threadfence_block();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_GROUPS * 128) : "memory");  // both groups' lists are final
      // warp (g, quarter) finishes rows quarter * 32 + 16 g .. + 15
      for (int rr = 0; rr < 16; ++rr) {
        const int row = quarter * 32 + g * 16 + rr;
        const long long u_row = (long long)mt * BLOCK_M + row;
        if (u_row >= p.n_users) break;  // warp-uniform
        const int n0 = cnt_sh[row], n1 = cnt_sh[BLOCK_M + row];
        float* cs0 = p.cand_scores + ((size_t)blockIdx.x * EPI_GROUPS * BLOCK_M + row) * CAP;
        int* ci0 = p.cand_ids + ((size_t)blockIdx.x * EPI_GROUPS * BLOCK_M + row) * CAP;
        const float* cs1 = cs0 + (size_t)BLOCK_M * CAP;
        const int* ci1 = ci0 + (size_t)BLOCK_M * CAP;
        for (int t = lane; t < n1; t += 32) cs0[n0 + t] = cs1[t], ci0[n0 + t] = ci1[t];  // n0 + n1 <= 2 k <= CAP
        const int n = n0 + n1;
        int n_new;
        if (n > p.k) compact_row(cs0, ci0, n, p.k, lane, true, &n_new); else __syncwarp();
        write_sorted_row(cs0, ci0, min(n, p.k), p.k, p.out_scores + u_row * p.k, p.out_ids + u_row * p.k, p.id_offset, lane);
        __syncwarp();
        if (p.n_peers > 1) {
          // fused exchange: the finished list goes straight into slot `my_rank` of every other GPU's gather buffer (peer
          // stores over NVLink / NVSwitch), overlapped with the MMA of the user tiles this CTA still has to sweep
          for (int t = lane; t < p.k; t += 32) {
            const float sv = p.out_scores[u_row * p.k + t];
            const long long iv = p.out_ids[u_row * p.k + t];
            const size_t at = ((size_t)p.my_rank * p.peer_rows + u_row) * p.k + t;
            for (int r = 0; r < p.n_peers; ++r)
              if (r != p.my_rank) p.peer_scores[r][at] = sv, p.peer_ids[r][at] = iv;
          }
        }
      }
      *thr_mine = -CUDART_INF_F;  // the next user tile starts from scratch
      __threadfence_block();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_GROUPS * 128) : "memory");  // lists and thresholds may be reused
    }
    if (st_on && lane == 0) {
      atomicAdd(p.stats + 0, (unsigned long long)st_comp), atomicAdd(p.stats + 1, (unsigned long long)st_slow);
      atomicAdd(p.stats + 2, (unsigned long long)st_wait), atomicAdd(p.stats + 3, (unsigned long long)st_ncomp);
      atomicAdd(p.stats + 4, (unsigned long long)st_nslow), atomicAdd(p.stats + 5, (unsigned long long)(clock64() - st_t0));
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---- user matrix: mean of the history rows, bf16 (cr_module.py:116-123) ---------------------------------
__device__ __forceinline__ float2 load_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load_pair(const __nv_bfloat16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }

template <typename T>
__global__ void __launch_bounds__(256) pool_users_kernel(const T* __restrict__ table, long long row_stride, long long n_news, int dim,
                                                         const int32_t* __restrict__ hist_offsets, const int32_t* __restrict__ hist_ids,
                                                         long long n_users, __nv_bfloat16* __restrict__ out, int32_t* __restrict__ flags) {
  constexpr int MAXP = 16;  // column pairs per lane: dim <= 2 * 32 * 16 = 1024
  const int lane = threadIdx.x & 31;
  const long long user = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (user >= n_users) return;
  const int h0 = hist_offsets[user], h1 = hist_offsets[user + 1];
  float2 acc[MAXP];
#pragma unroll
  for (int j = 0; j < MAXP; ++j) acc[j] = make_float2(0.f, 0.f);
  for (int h = h0; h < h1; ++h) {
    long long id = hist_ids[h];
    if ((unsigned long long)id >= (unsigned long long)n_news) {
      id = 0;
      if (flags && lane == 0) atomicOr(flags, MB200_FLAG_BAD_ID);
    }
    const T* row = table + id * row_stride;
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      const int d0 = 2 * (lane + 32 * j);
      if (d0 < dim) {
        const float2 v = load_pair(row + d0);
        acc[j].x += v.x, acc[j].y += v.y;
      }
    }
  }
  const float hf = (float)(h1 - h0);  // true division, like torch.div(sum, hist_size)
#pragma unroll
  for (int j = 0; j < MAXP; ++j) {
    const int d0 = 2 * (lane + 32 * j);
    if (d0 < dim)
      *reinterpret_cast<__nv_bfloat162*>(out + user * dim + d0) = __floats2bfloat162_rn(__fdiv_rn(acc[j].x, hf), __fdiv_rn(acc[j].y, hf));
  }
}

// ---- multi-GPU merge: per-shard sorted top-k lists [shards][n_users][k] -> global top-k ------------------
__global__ void __launch_bounds__(256) merge_topk_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, int shards,
                                                         long long n_users, int k, float* __restrict__ out_scores,
                                                         long long* __restrict__ out_ids) {
  // one warp per user; `shards` sorted lists are merged by repeated selection of the best head (shards <= 32)
  const int lane = threadIdx.x & 31;
  const long long user = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (user >= n_users) return;
  int head = 0;  // lane s < shards walks list s
  for (int t = 0; t < k; ++t) {
    float s = -CUDART_INF_F;
    long long id = 0x7fffffffffffffffll;
    if (lane < shards && head < k) {
      const size_t at = ((size_t)lane * n_users + user) * k + head;
      s = scores[at], id = ids[at];
      if (id < 0) s = -CUDART_INF_F, id = 0x7fffffffffffffffll;
    }
    float bs = s;
    long long bi = id;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(kFull, bs, o);
      const long long oi = __shfl_xor_sync(kFull, bi, o);
      if (os > bs || (os == bs && oi < bi)) bs = os, bi = oi;
    }
    if (lane < shards && head < k && s == bs && id == bi) ++head;
    if (lane == 0) {
      out_scores[user * k + t] = bs;
      out_ids[user * k + t] = (bi == 0x7fffffffffffffffll) ? -1ll : bi;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, long long rows, int dim, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return MB200_ERR_CUDA;
  const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
  const cuuint32_t box[2] = {(cuuint32_t)rt::BLOCK_K, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MB200_OK : MB200_ERR_CUDA;
}

// persistent grid: one CTA per SM, never more CTAs than user tiles; a CTA-pair launch needs an even grid
static int retrieval_grid(int device, long long n_users, bool pair) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const long long m_tiles = (n_users + rt::BLOCK_M - 1) / rt::BLOCK_M;
  if (!pair) return (int)(m_tiles < sms ? m_tiles : sms);
  const long long units = (m_tiles + 1) / 2;
  const long long pairs = units < sms / 2 ? units : sms / 2;
  return (int)(2 * pairs);
}

size_t retrieval_workspace_bytes(const mb200_retrieval_desc* d) {
  if (d == nullptr || d->struct_size != sizeof(mb200_retrieval_desc) || d->n_users <= 0) return 0;
  const long long m_tiles = (d->n_users + rt::BLOCK_M - 1) / rt::BLOCK_M;
  const long long grid = m_tiles + 1 < 160 ? m_tiles + 1 : 160;  // + 1: a CTA pair rounds the tile count up to even
  return (size_t)grid * rt::EPI_GROUPS * rt::BLOCK_M * rt::CAP * 8 + rt::WS_HEADER;
}

int retrieve_topk(const mb200_retrieval_desc* d, cudaStream_t stream) {
  if (d == nullptr || d->struct_size != sizeof(mb200_retrieval_desc)) return MB200_ERR_INVALID_ARG;
  if (!d->users || !d->catalog || !d->out_scores || !d->out_ids || d->n_users <= 0 || d->n_catalog <= 0) return MB200_ERR_INVALID_ARG;
  if (d->k < 1 || d->k > rt::MAX_K) return MB200_ERR_UNSUPPORTED;
  if (d->dim < rt::BLOCK_K || d->dim % rt::BLOCK_K != 0) return MB200_ERR_UNSUPPORTED;
  if (d->n_catalog > 0x7fffff00ll || d->n_users > 0x7fffff00ll * (long long)rt::BLOCK_M) return MB200_ERR_UNSUPPORTED;
  if (((uintptr_t)d->users & 15) || ((uintptr_t)d->catalog & 15)) return MB200_ERR_INVALID_ARG;
  if (d->n_peers > 1) {
    if (d->n_peers > MB200_MAX_TABLE_SHARDS || d->my_rank < 0 || d->my_rank >= d->n_peers || d->peer_rows < d->n_users) return MB200_ERR_INVALID_ARG;
    for (int r = 0; r < d->n_peers; ++r)
      if (!d->peer_scores[r] || !d->peer_ids[r]) return MB200_ERR_INVALID_ARG;
  }
  int device = 0;
  int st = use_device_of(d->users, &device);
  if (st != MB200_OK) return st;
  const size_t need = retrieval_workspace_bytes(d);
  if (d->workspace == nullptr || ((uintptr_t)d->workspace & 255) || d->workspace_bytes < need) return MB200_ERR_WORKSPACE;

  // CTA pairs (cta_group::2) when there are at least two user tiles; tuning key 5: 0 = never, 1 = when possible
  const bool pair = tuning().retrieval_pair != 0 && d->n_users > rt::BLOCK_M;
  CUtensorMap map_u, map_c;
  if ((st = make_map(&map_u, d->users, d->n_users, d->dim, rt::BLOCK_M)) != MB200_OK) return st;
  if ((st = make_map(&map_c, d->catalog, d->n_catalog, d->dim, pair ? rt::BLOCK_N / 2 : rt::BLOCK_N)) != MB200_OK) return st;

  RetrievalParams p{};
  const int grid = retrieval_grid(device, d->n_users, pair);
  unsigned char* ws = reinterpret_cast<unsigned char*>(d->workspace);
  p.error_flag = reinterpret_cast<int*>(ws);
  p.stats = reinterpret_cast<unsigned long long*>(ws + 64);
  p.progress = reinterpret_cast<int*>(ws + 256);  // [160]
  p.window = tuning().retrieval_window;
  p.cand_scores = reinterpret_cast<float*>(ws + rt::WS_HEADER);
  p.cand_ids = reinterpret_cast<int*>(ws + rt::WS_HEADER + (size_t)grid * rt::EPI_GROUPS * rt::BLOCK_M * rt::CAP * 4);
  p.out_scores = d->out_scores, p.out_ids = reinterpret_cast<long long*>(d->out_ids), p.debug_scores = d->debug_scores;
  p.n_users = d->n_users, p.n_catalog = d->n_catalog, p.id_offset = d->catalog_id_offset;
  p.dim = d->dim, p.k = d->k;
  p.diag = tuning().retrieval_diag;
  p.n_peers = d->n_peers > 1 ? d->n_peers : 0, p.my_rank = d->my_rank, p.peer_rows = d->peer_rows;
  for (int r = 0; r < p.n_peers; ++r) p.peer_scores[r] = d->peer_scores[r], p.peer_ids[r] = reinterpret_cast<long long*>(d->peer_ids[r]);
  p.m_tiles = (int)((d->n_users + rt::BLOCK_M - 1) / rt::BLOCK_M);
  p.n_tiles = (int)((d->n_catalog + rt::BLOCK_N - 1) / rt::BLOCK_N);
  st = cuda_status(cudaMemsetAsync(p.error_flag, 0, rt::WS_HEADER, stream), "cudaMemsetAsync");
  if (st != MB200_OK) return st;
  if (pair) {
    st = cuda_status(cudaFuncSetAttribute(retrieve_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rt::SMEM_BYTES), "cudaFuncSetAttribute");
    if (st != MB200_OK) return st;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(rt::THREADS), cfg.dynamicSmemBytes = rt::SMEM_BYTES, cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    st = cuda_status(cudaLaunchKernelEx(&cfg, retrieve_topk_kernel<true>, map_u, map_c, p), "retrieve_topk_kernel<pair>");
    note_launch(1);
    return st;
  }
  st = cuda_status(cudaFuncSetAttribute(retrieve_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rt::SMEM_BYTES), "cudaFuncSetAttribute");
  if (st != MB200_OK) return st;
  retrieve_topk_kernel<false><<<grid, rt::THREADS, rt::SMEM_BYTES, stream>>>(map_u, map_c, p);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "retrieve_topk_kernel");
}

int pool_users(const void* table, int dtype, int dim, long long row_stride, long long n_news, const int32_t* hist_offsets,
               const int32_t* hist_ids, long long n_users, void* out_bf16, int32_t* flags, cudaStream_t stream) {
  if (!table || !hist_offsets || !hist_ids || !out_bf16 || n_users <= 0 || dim <= 0 || row_stride < dim) return MB200_ERR_INVALID_ARG;
  if (dim % 2 != 0 || dim > 1024 || row_stride % 2 != 0) return MB200_ERR_UNSUPPORTED;
  int st = use_device_of(table, nullptr);
  if (st != MB200_OK) return st;
  const int warps = 8;
  const unsigned grid = (unsigned)((n_users + warps - 1) / warps);
  if (dtype == MB200_F32)
    pool_users_kernel<float><<<grid, warps * 32, 0, stream>>>(reinterpret_cast<const float*>(table), row_stride, n_news, dim, hist_offsets, hist_ids,
                                                               n_users, reinterpret_cast<__nv_bfloat16*>(out_bf16), flags);
  else if (dtype == MB200_BF16)
    pool_users_kernel<__nv_bfloat16><<<grid, warps * 32, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(table), row_stride, n_news, dim,
                                                                       hist_offsets, hist_ids, n_users, reinterpret_cast<__nv_bfloat16*>(out_bf16), flags);
  else
    return MB200_ERR_INVALID_ARG;
  note_launch(1);
  return cuda_status(cudaGetLastError(), "pool_users_kernel");
}

int merge_topk(const float* scores, const long long* ids, int shards, long long n_users, int k, float* out_scores, long long* out_ids,
               cudaStream_t stream) {
  if (!scores || !ids || !out_scores || !out_ids || shards < 1 || shards > 32 || n_users <= 0 || k < 1) return MB200_ERR_INVALID_ARG;
  int st = use_device_of(scores, nullptr);
  if (st != MB200_OK) return st;
  const int warps = 8;
  merge_topk_kernel<<<(unsigned)((n_users + warps - 1) / warps), warps * 32, 0, stream>>>(scores, ids, shards, n_users, k, out_scores, out_ids);
  note_launch(1);
  return cuda_status(cudaGetLastError(), "merge_topk_kernel");
}

}  // namespace mb200
