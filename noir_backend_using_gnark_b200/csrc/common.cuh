// Context, workspace and error plumbing shared by the NTT / MSM translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../include/b200zk.h"

namespace b200zk {

struct NttDomain {
  bool ready = false;
  unsigned log2n = 0;
  // tw[i] = w^i (forward) / w^-i (inverse) for i < max(1, N/2): full tables, no extra multiplications per butterfly
  void* tw_fwd = nullptr;
  void* tw_inv = nullptr;
  // two-level coset tables: 5^e = lo[e & (LO-1)] * hi[e >> LO_BITS];  the inverse lo table carries the 1/n factor
  void* coset_lo = nullptr;
  void* coset_hi = nullptr;
  void* coset_inv_lo = nullptr;
  void* coset_inv_hi = nullptr;
  void* scalars = nullptr;  // [0] = 1/n (Montgomery)
};

// phases timed with CUDA events when profiling is enabled (b200zk_profile_*)
enum Phase {
  PH_MSM_DIGITS = 0, PH_MSM_SCAN = 1, PH_MSM_SCATTER = 2, PH_MSM_ACCUMULATE = 3, PH_MSM_BIG = 4, PH_MSM_REDUCE = 5,
  PH_MSM_FINAL = 6, PH_NTT_PASS = 7, PH_COUNT = 8
};
struct PhaseRecord {
  int phase;
  cudaEvent_t e0, e1;
};

struct DeviceBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace b200zk

namespace b200zk {
// MSM workspace (grown on demand, reused across calls).  A context owns MSM_LANES of them, each with its own stream, so
// that independent MSMs (the three commitments of one prover round) can be in flight together: the latency-bound front
// and back ends of one hide under the bucket accumulation of another.
static constexpr int MSM_LANES = 3;
struct MsmWorkspace {
  DeviceBuf msm_digits, msm_sorted, msm_counts, msm_starts, msm_cursor, msm_buckets, msm_tmp, msm_small, msm_scan_tmp,
      msm_big, msm_pairs, msm_keys;
  cudaStream_t stream = nullptr;   // lane 0: the context stream
  cudaEvent_t done = nullptr;      // recorded after the lane's last MSM (lanes > 0)
};
}  // namespace b200zk

struct b200zk_bases {
  const void* dev = nullptr;  // n x 64 B affine points
  size_t n = 0;
  bool owned = false;
  // optional window table (b200zk_bases_precompute): table[j*n + i] = 2^tab_wstart[j] * P_i, j < tab_W; row 0 = the bases
  void* table = nullptr;
  unsigned tab_c = 0, tab_W = 0;
  uint8_t tab_wstart[65] = {0};  // window j covers scalar bits [tab_wstart[j], tab_wstart[j+1])
  // narrow-window table of the first small_n bases (large base sets only): what small MSMs use
  void* small_table = nullptr;
  size_t small_n = 0;
  unsigned small_c = 0;
  uint8_t small_wstart[65] = {0};
};

struct b200zk_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;  // copy stream of the host-scalar MSM (chunk i+1 lands while chunk i is processed)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_chunk[8] = {};  // host-scalar MSM: chunk i of the scalars has landed (copy stream -> compute stream)
  int msm_host_chunks = 0;       // 0 = choose from n; 1 = never split (tests / tuning)
  int msm_chunk_log = 0;         // tests / tuning: force the running-sum chunk size of the bucket reduction (3 or 5)
  bool lane_pending = false;     // a forked commitment is in flight on MSM lane 1
  int msm_single_lane = 0;       // tests / tuning: the prover's commitment rounds run one MSM after the other
  int msm_no_tiny = 0;           // tests: force the bucket pipeline also for small table-mode MSMs
  int msm_pair_rounds = -1;      // batched-affine pair rounds before the bucket walk: -1 = from the size, 0 = off, r = forced
  int msm_pair_kmax = 0;         // tuning: additions per inversion per lane (0 = default)
  int msm_scatter_passes = 0;    // tests / tuning: bucket-range passes of the counting sort's scatter (0 = from the bucket count)
  uint64_t launches = 0;
  char cuda_err[256] = {0};
  int forced_window = 0;
  int ntt_radix2 = 0;  // tests: force the radix-2 pass kernel
  bool profiling = false;
  std::vector<b200zk::PhaseRecord> records;
  b200zk::NttDomain domains[B200ZK_MAX_LOG2N + 1];
  // staging buffer for host-pointer entry points
  b200zk::DeviceBuf stage;
  b200zk::MsmWorkspace ws[b200zk::MSM_LANES];
};

namespace b200zk {

inline int set_cuda_error(b200zk_ctx* ctx, cudaError_t e, const char* where) {
  if (ctx) snprintf(ctx->cuda_err, sizeof(ctx->cuda_err), "%s: %s", where, cudaGetErrorString(e));
  cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? B200ZK_ERR_OOM : B200ZK_ERR_CUDA;
}

#define B200ZK_CUDA(ctx, call)                                                \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return b200zk::set_cuda_error(ctx, e__, #call);   \
  } while (0)

#define B200ZK_LAUNCH_CHECK(ctx, name)                                        \
  do {                                                                        \
    (ctx)->launches++;                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) return b200zk::set_cuda_error(ctx, e__, name);    \
  } while (0)

#define B200ZK_TRY(expr)                \
  do {                                  \
    int rc__ = (expr);                  \
    if (rc__ != B200ZK_OK) return rc__; \
  } while (0)

inline int ensure(b200zk_ctx* ctx, DeviceBuf& b, size_t bytes, cudaStream_t user = nullptr) {
  if (b.cap >= bytes) return B200ZK_OK;
  if (b.p) {
    // the buffer may still be in use by work enqueued earlier on the stream
    B200ZK_CUDA(ctx, cudaStreamSynchronize(user ? user : ctx->stream));
    B200ZK_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + (bytes >> 3);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    b.p = nullptr;
    return set_cuda_error(ctx, e, "cudaMalloc(workspace)");
  }
  b.cap = want;
  return B200ZK_OK;
}

// RAII phase timer: records an event pair around a group of launches when ctx->profiling is on
struct PhaseTimer {
  b200zk_ctx* ctx;
  PhaseRecord rec;
  bool on;
  cudaStream_t st;
  PhaseTimer(b200zk_ctx* c, int phase, cudaStream_t stream = nullptr) : ctx(c), on(c->profiling), st(stream ? stream : c->stream) {
    if (!on) return;
    rec.phase = phase;
    if (cudaEventCreate(&rec.e0) != cudaSuccess || cudaEventCreate(&rec.e1) != cudaSuccess) {
      on = false;
      cudaGetLastError();
      return;
    }
    cudaEventRecord(rec.e0, st);
  }
  ~PhaseTimer() {
    if (!on) return;
    cudaEventRecord(rec.e1, st);
    ctx->records.push_back(rec);
  }
};

// entry points implemented in ntt.cu / msm.cu
int ntt_run(b200zk_ctx* ctx, void* a_dev, unsigned log2n, int inverse, int decimation, int coset);
int ntt_dist_run(b200zk_ctx* ctx, const void* src, void* dst, unsigned log2n, unsigned log2g, unsigned rank,
                 unsigned log2c, int half, int inverse, int decimation, int coset, void* const* peers = nullptr);
int bit_reverse_run(b200zk_ctx* ctx, void* a_dev, unsigned log2n);
void ntt_free_domains(b200zk_ctx* ctx);
int ntt_prepare(b200zk_ctx* ctx, unsigned log2n);  // builds the twiddle / coset tables of a domain if missing
// lane: which workspace / stream of the context runs it (0 = the context stream)
// part / shape_n: one MSM fed in several chunks (see msm.cu); MSM_PART_ALL = the whole MSM in one call
enum { MSM_PART_ALL = 0, MSM_PART_FIRST = 1, MSM_PART_MORE = 2, MSM_PART_LAST = 3 };
int msm_run(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_dev, size_t n,
            void* out_dev, int out_kind, int lane = 0, int part = MSM_PART_ALL, size_t shape_n = 0);
int msm_precompute_run(b200zk_ctx* ctx, b200zk_bases* bases, int c);
unsigned msm_window_count(const b200zk_ctx* ctx, const b200zk_bases* bases, size_t n);
// out_kind 0: canonical affine (64 B); 1: extended-Jacobian sum (128 B)
int g1_sum_run(b200zk_ctx* ctx, const void* partials_dev, size_t count, void* out_dev, int out_kind = 0);
int srs_decompress_run(b200zk_ctx* ctx, const void* compressed_host, size_t n, void* out_dev, unsigned* bad_count);
int srs_compress_run(b200zk_ctx* ctx, const void* points_dev, size_t n, void* out_host);
int microbench_run(b200zk_ctx* ctx, int which, double* out_ops_per_s);
int srs_generate_run(b200zk_ctx* ctx, const void* alpha_dev, size_t first, size_t n, void* out_dev);

}  // namespace b200zk
