// BN254 G1 multi-scalar multiplication for sm_100a: replaces gnark-crypto v0.9.1 ecc/bn254/multiexp.go
// ((*G1Affine).MultiExp: partitionScalars + processChunkG1* + msmReduceChunkG1Affine), reached in the reference
// through kzg.Commit inside plonk.Prove / plonk.Setup (/root/reference/gnark_backend_ffi/backend/plonk/plonk.go:67, :21).
//
// Pipeline (all on the context stream, no host synchronisation):
//   1 msm_digits_kernel     scalar: Montgomery -> regular, signed c-bit digits (2^(c-1) buckets per window),
//                           digit matrix [window][point] + bucket histogram (global atomics)
//   2 exclusive scan         bucket start offsets
//   3 msm_scatter_kernel    counting sort: (point index | sign) grouped by (window, bucket)
//   4 msm_accumulate_kernel one thread per bucket walks its run: extended-Jacobian mixed additions, next point
//                           prefetched while the current addition runs; over-long runs (skewed scalars) are left to
//   4b msm_big_* kernels    which split a run over many CTAs and tree-reduce the partial sums
//   5 msm_reduce_l1/l2      per-window sum_b b*B[b] by three levels of 32-wide running sums
//   6 msm_final_kernel      Horner over windows (c doublings each), then canonical affine (or XYZZ partial)
// The result is a canonical affine point, so it is bit-identical to gnark's for any window size / summation order.
#include <cub/device/device_scan.cuh>
#include "common.cuh"
#include "g1.cuh"

namespace b200zk {

static constexpr int CHUNK = 32;          // buckets per running-sum chunk
static constexpr int BIG_CHUNK = 8192;    // sorted entries per CTA in the long-run path
static constexpr int BIG_THREADS = 256;

struct MsmShape {
  unsigned c;         // window bits
  unsigned W;         // number of windows
  unsigned B;         // buckets per window = 2^(c-1)
  unsigned big_len;   // runs longer than this go to the cooperative path
};

// ---------------------------------------------------------------------------------------------------
// 1. digits + histogram
// ---------------------------------------------------------------------------------------------------
__global__ void msm_digits_kernel(const uint4* __restrict__ scalars, size_t n, MsmShape sh, int16_t* __restrict__ digits,
                                  unsigned* __restrict__ counts) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = fe_from_mont(fe_load<FrParams>(scalars + 2 * i));
  unsigned carry = 0;
  const unsigned mask = (1u << sh.c) - 1u;
  for (unsigned w = 0; w < sh.W; w++) {
    const unsigned bit = w * sh.c;
    const unsigned limb = bit >> 5, off = bit & 31;
    unsigned v = 0;
    if (limb < 8) {
      // 64-bit window over two limbs
      unsigned lo = s.l[limb];
      unsigned hi = limb + 1 < 8 ? s.l[limb + 1] : 0u;
      v = (unsigned)((((uint64_t)hi << 32) | lo) >> off) & mask;
    }
    int d = (int)(v + carry);
    carry = 0;
    if (d >= (int)sh.B) {  // d in [B, 2B] -> d - 2^c in [-B, 0]
      d -= (int)(2 * sh.B);
      carry = 1;
    }
    digits[(size_t)w * n + i] = (int16_t)d;
    if (d != 0) {
      unsigned mag = d < 0 ? (unsigned)(-d) : (unsigned)d;
      atomicAdd(&counts[w * sh.B + (mag - 1)], 1u);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// 3. scatter (counting sort)
// ---------------------------------------------------------------------------------------------------
__global__ void msm_scatter_kernel(const int16_t* __restrict__ digits, size_t n, MsmShape sh, unsigned* __restrict__ cursor,
                                   unsigned* __restrict__ sorted) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  unsigned w = blockIdx.y;
  if (i >= n) return;
  int d = digits[(size_t)w * n + i];
  if (d == 0) return;
  unsigned mag = d < 0 ? (unsigned)(-d) : (unsigned)d;
  unsigned pos = atomicAdd(&cursor[w * sh.B + (mag - 1)], 1u);
  sorted[pos] = (unsigned)i | (d < 0 ? 0x80000000u : 0u);
}

// ---------------------------------------------------------------------------------------------------
// 4. bucket accumulation
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ G1Affine load_signed(const void* bases, unsigned entry) {
  G1Affine p = g1_load_affine(bases, entry & 0x7fffffffu);
  if (entry & 0x80000000u) p.y = fe_neg(p.y);
  return p;
}

__global__ void __launch_bounds__(128) msm_accumulate_kernel(const void* __restrict__ bases, const unsigned* __restrict__ starts,
                                                             const unsigned* __restrict__ sorted, MsmShape sh,
                                                             unsigned nbuckets, void* __restrict__ buckets) {
  unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  unsigned lo = starts[b], hi = starts[b + 1];
  G1XYZZ acc = g1_xyzz_inf();
  if (hi - lo <= sh.big_len && hi > lo) {
    G1Affine cur = load_signed(bases, sorted[lo]);
    for (unsigned j = lo + 1; j < hi; j++) {
      G1Affine nxt = load_signed(bases, sorted[j]);
      g1_add_mixed(acc, cur);
      cur = nxt;
    }
    g1_add_mixed(acc, cur);
  }
  g1_store_xyzz(buckets, b, acc);
}

// ---- long runs: list them, split over CTAs, reduce --------------------------------------------------
struct BigPlan {
  unsigned nbig;          // number of long runs
  unsigned nchunks;       // total CTA chunks
};

// every long run is cut into BIG_CHUNK-entry chunks; chunk descriptors are (bucket, index within the run)
__global__ void msm_big_list_kernel(const unsigned* __restrict__ starts, unsigned nbuckets, MsmShape sh, BigPlan* plan,
                                    unsigned* __restrict__ big_bucket, unsigned* __restrict__ big_first_chunk,
                                    unsigned cap, unsigned* __restrict__ chunk_bucket, unsigned* __restrict__ chunk_idx) {
  unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  unsigned len = starts[b + 1] - starts[b];
  if (len > sh.big_len) {
    unsigned chunks = (len + BIG_CHUNK - 1) / BIG_CHUNK;
    unsigned slot = atomicAdd(&plan->nbig, 1u);
    unsigned first = atomicAdd(&plan->nchunks, chunks);
    if (slot < cap) {
      big_bucket[slot] = b;
      big_first_chunk[slot] = first;
      for (unsigned k = 0; k < chunks; k++) {
        chunk_bucket[first + k] = b;
        chunk_idx[first + k] = k;
      }
    }
  }
}

__device__ __forceinline__ void block_reduce_xyzz(G1XYZZ& acc, G1XYZZ* sh_pts) {
  // tree reduction over BIG_THREADS partial sums in shared memory
  sh_pts[threadIdx.x] = acc;
  __syncthreads();
  for (unsigned s = BIG_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      G1XYZZ a = sh_pts[threadIdx.x];
      g1_add(a, sh_pts[threadIdx.x + s]);
      sh_pts[threadIdx.x] = a;
    }
    __syncthreads();
  }
  acc = sh_pts[0];
  __syncthreads();
}

// grid-stride over chunk descriptors: one CTA sums one chunk of a long run
__global__ void __launch_bounds__(BIG_THREADS) msm_big_accumulate_kernel(const void* __restrict__ bases,
                                                                         const unsigned* __restrict__ starts,
                                                                         const unsigned* __restrict__ sorted,
                                                                         const BigPlan* __restrict__ plan,
                                                                         const unsigned* __restrict__ chunk_bucket,
                                                                         const unsigned* __restrict__ chunk_idx,
                                                                         void* __restrict__ partials) {
  extern __shared__ uint4 big_smem[];
  G1XYZZ* sh_pts = reinterpret_cast<G1XYZZ*>(big_smem);
  const unsigned nchunks = plan->nchunks;
  for (unsigned item = blockIdx.x; item < nchunks; item += gridDim.x) {
    const unsigned b = chunk_bucket[item];
    const unsigned lo = starts[b], hi = starts[b + 1];
    unsigned clo = lo + chunk_idx[item] * BIG_CHUNK;
    unsigned chi = clo + BIG_CHUNK < hi ? clo + BIG_CHUNK : hi;
    G1XYZZ acc = g1_xyzz_inf();
    for (unsigned j = clo + threadIdx.x; j < chi; j += BIG_THREADS) {
      G1Affine p = load_signed(bases, sorted[j]);
      g1_add_mixed(acc, p);
    }
    block_reduce_xyzz(acc, sh_pts);
    if (threadIdx.x == 0) g1_store_xyzz(partials, item, acc);
  }
}

// one CTA per long run: sum its chunk partials into the bucket
__global__ void __launch_bounds__(BIG_THREADS) msm_big_reduce_kernel(const unsigned* __restrict__ starts,
                                                                     const BigPlan* __restrict__ plan,
                                                                     const unsigned* __restrict__ big_bucket,
                                                                     const unsigned* __restrict__ big_first_chunk,
                                                                     unsigned cap, const void* __restrict__ partials,
                                                                     void* __restrict__ buckets) {
  extern __shared__ uint4 big_smem[];
  G1XYZZ* sh_pts = reinterpret_cast<G1XYZZ*>(big_smem);
  const unsigned nbig = plan->nbig < cap ? plan->nbig : cap;
  for (unsigned slot = blockIdx.x; slot < nbig; slot += gridDim.x) {
    const unsigned b = big_bucket[slot];
    const unsigned len = starts[b + 1] - starts[b];
    const unsigned chunks = (len + BIG_CHUNK - 1) / BIG_CHUNK;
    G1XYZZ acc = g1_xyzz_inf();
    for (unsigned ch = threadIdx.x; ch < chunks; ch += BIG_THREADS) {
      G1XYZZ q = g1_load_xyzz(partials, big_first_chunk[slot] + ch);
      g1_add(acc, q);
    }
    block_reduce_xyzz(acc, sh_pts);
    if (threadIdx.x == 0) g1_store_xyzz(buckets, b, acc);
  }
}

// ---------------------------------------------------------------------------------------------------
// 5. per-window bucket reduction: S_w = sum_{b=1..B} b * bucket[b-1]
//    level 1: chunk t of 32 buckets -> run1 = sum, acc1 = sum_j (j+1)*item_j
//    level 2: chunk u of 32 level-1 chunks -> asum = sum acc1, run2 = sum run1, acc2 = sum_j j*run1_j
//    final  : S_w = sum_u asum_u + 32*(sum_u acc2_u + 32 * sum_u u*run2_u)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) msm_reduce_l1_kernel(const void* __restrict__ buckets, unsigned nchunks_total,
                                                            void* __restrict__ run1, void* __restrict__ acc1) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks_total) return;
  G1XYZZ run = g1_xyzz_inf(), acc = g1_xyzz_inf();
  for (int j = CHUNK - 1; j >= 0; j--) {
    G1XYZZ q = g1_load_xyzz(buckets, (size_t)t * CHUNK + j);
    g1_add(run, q);
    g1_add(acc, run);
  }
  g1_store_xyzz(run1, t, run);
  g1_store_xyzz(acc1, t, acc);
}

__global__ void __launch_bounds__(128) msm_reduce_l2_kernel(const void* __restrict__ run1, const void* __restrict__ acc1,
                                                            unsigned n1_per_window, unsigned n2_per_window, unsigned W,
                                                            void* __restrict__ asum, void* __restrict__ run2,
                                                            void* __restrict__ acc2) {
  unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n2_per_window * W) return;
  unsigned w = id / n2_per_window, u = id % n2_per_window;
  G1XYZZ a = g1_xyzz_inf(), run = g1_xyzz_inf(), acc = g1_xyzz_inf();
  for (int j = CHUNK - 1; j >= 0; j--) {
    unsigned t = u * CHUNK + j;
    if (t >= n1_per_window) continue;
    size_t idx = (size_t)w * n1_per_window + t;
    G1XYZZ q = g1_load_xyzz(acc1, idx);
    g1_add(a, q);
    q = g1_load_xyzz(run1, idx);
    g1_add(run, q);
    if (j > 0) g1_add(acc, run);
  }
  g1_store_xyzz(asum, id, a);
  g1_store_xyzz(run2, id, run);
  g1_store_xyzz(acc2, id, acc);
}

// ---------------------------------------------------------------------------------------------------
// 6. window sums + Horner + normalisation
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) msm_final_kernel(const void* __restrict__ asum, const void* __restrict__ run2,
                                                       const void* __restrict__ acc2, unsigned n2_per_window, MsmShape sh,
                                                       void* __restrict__ out, int out_kind) {
  __shared__ G1XYZZ wsum[64];  // W <= 43 (c >= 6)
  const unsigned w = threadIdx.x;
  if (w < sh.W) {
    G1XYZZ A = g1_xyzz_inf(), Bs = g1_xyzz_inf(), run = g1_xyzz_inf(), C = g1_xyzz_inf();
    for (int u = (int)n2_per_window - 1; u >= 0; u--) {
      size_t idx = (size_t)w * n2_per_window + u;
      G1XYZZ q = g1_load_xyzz(asum, idx);
      g1_add(A, q);
      q = g1_load_xyzz(acc2, idx);
      g1_add(Bs, q);
      q = g1_load_xyzz(run2, idx);
      g1_add(run, q);
      if (u > 0) g1_add(C, run);
    }
    // S = A + 32*(Bs + 32*C)
    for (int k = 0; k < 5; k++) g1_double(C);
    g1_add(C, Bs);
    for (int k = 0; k < 5; k++) g1_double(C);
    g1_add(C, A);
    wsum[w] = C;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    G1XYZZ total = wsum[sh.W - 1];
    for (int ww = (int)sh.W - 2; ww >= 0; ww--) {
      for (unsigned k = 0; k < sh.c; k++) g1_double(total);
      g1_add(total, wsum[ww]);
    }
    if (out_kind == 0) {
      G1Affine r = g1_to_affine(total);
      g1_store_affine(out, 0, r);
    } else {
      g1_store_xyzz(out, 0, total);
    }
  }
}

__global__ void g1_sum_kernel(const void* __restrict__ partials, unsigned count, void* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G1XYZZ total = g1_xyzz_inf();
  for (unsigned i = 0; i < count; i++) {
    G1XYZZ q = g1_load_xyzz(partials, i);
    g1_add(total, q);
  }
  G1Affine r = g1_to_affine(total);
  g1_store_affine(out, 0, r);
}

__global__ void copy_u32_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------
static unsigned choose_window(size_t n) {
  unsigned lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  // fewer, larger windows as n grows; at least 2^5 buckets so the 32-wide reduction levels are full
  if (lg >= 23) return 16;
  if (lg >= 21) return 15;
  if (lg >= 19) return 14;
  if (lg >= 17) return 13;
  if (lg >= 15) return 12;
  if (lg >= 13) return 11;
  if (lg >= 11) return 10;
  if (lg >= 9) return 9;
  return 8;
}

int msm_run(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_dev, size_t n,
            void* out_dev, int out_kind) {
  if (!bases || !out_dev || (n && !scalars_dev)) return B200ZK_ERR_BAD_ARG;
  if (first_base > bases->n || n > bases->n - first_base) return B200ZK_ERR_BAD_ARG;
  if (n >= ((size_t)1 << 31)) return B200ZK_ERR_UNSUPPORTED;
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    // empty sum = point at infinity: affine (0,0) / XYZZ with ZZ = 0
    B200ZK_CUDA(ctx, cudaMemsetAsync(out_dev, 0, out_kind == 0 ? 64 : 128, st));
    return B200ZK_OK;
  }
  MsmShape sh;
  sh.c = ctx->forced_window ? (unsigned)ctx->forced_window : choose_window(n);
  if (sh.c < 6) sh.c = 6;
  if (sh.c > 16) sh.c = 16;
  sh.W = (255 + sh.c - 1) / sh.c;  // c*W >= 255: the top window never produces a carry
  sh.B = 1u << (sh.c - 1);
  const size_t total = n * sh.W;
  if (total >= ((size_t)1 << 32)) return B200ZK_ERR_UNSUPPORTED;
  const unsigned nbuckets = sh.W * sh.B;
  {
    size_t avg = total / nbuckets;
    size_t bl = 4 * avg + 256;
    sh.big_len = (unsigned)(bl > 0x7fffffff ? 0x7fffffff : bl);
  }
  const unsigned n1 = sh.B / CHUNK;                  // level-1 chunks per window
  const unsigned n2 = (n1 + CHUNK - 1) / CHUNK;      // level-2 chunks per window
  const unsigned big_cap = nbuckets / 4 + 16;
  const size_t max_big_chunks = total / BIG_CHUNK + big_cap + 1;

  B200ZK_TRY(ensure(ctx, ctx->msm_digits, total * sizeof(int16_t)));
  B200ZK_TRY(ensure(ctx, ctx->msm_sorted, total * sizeof(unsigned)));
  B200ZK_TRY(ensure(ctx, ctx->msm_counts, (size_t)(nbuckets + 1) * 4));
  B200ZK_TRY(ensure(ctx, ctx->msm_starts, (size_t)(nbuckets + 1) * 4));
  B200ZK_TRY(ensure(ctx, ctx->msm_cursor, (size_t)(nbuckets + 1) * 4));
  B200ZK_TRY(ensure(ctx, ctx->msm_buckets, (size_t)nbuckets * 128));
  B200ZK_TRY(ensure(ctx, ctx->msm_tmp, ((size_t)sh.W * n1 * 2 + (size_t)sh.W * n2 * 3) * 128));
  B200ZK_TRY(ensure(ctx, ctx->msm_small, sizeof(BigPlan) + (size_t)big_cap * 8 + max_big_chunks * 8));
  B200ZK_TRY(ensure(ctx, ctx->msm_big, max_big_chunks * 128));
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned*)nullptr, (unsigned*)nullptr, (int)(nbuckets + 1), st);
  B200ZK_TRY(ensure(ctx, ctx->msm_scan_tmp, scan_bytes));

  const void* base_ptr = (const char*)bases->dev + first_base * 64;
  int16_t* digits = (int16_t*)ctx->msm_digits.p;
  unsigned* sorted = (unsigned*)ctx->msm_sorted.p;
  unsigned* counts = (unsigned*)ctx->msm_counts.p;
  unsigned* starts = (unsigned*)ctx->msm_starts.p;
  unsigned* cursor = (unsigned*)ctx->msm_cursor.p;
  char* tmp = (char*)ctx->msm_tmp.p;
  void* run1 = tmp;
  void* acc1 = tmp + (size_t)sh.W * n1 * 128;
  void* asum = tmp + (size_t)sh.W * n1 * 256;
  void* run2 = (char*)asum + (size_t)sh.W * n2 * 128;
  void* acc2 = (char*)run2 + (size_t)sh.W * n2 * 128;
  BigPlan* plan = (BigPlan*)ctx->msm_small.p;
  unsigned* big_bucket = (unsigned*)((char*)ctx->msm_small.p + sizeof(BigPlan));
  unsigned* big_first = big_bucket + big_cap;
  unsigned* chunk_bucket = big_first + big_cap;
  unsigned* chunk_idx = chunk_bucket + max_big_chunks;

  B200ZK_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)(nbuckets + 1) * 4, st));
  B200ZK_CUDA(ctx, cudaMemsetAsync(plan, 0, sizeof(BigPlan), st));
  {
    PhaseTimer pt(ctx, PH_MSM_DIGITS);
    unsigned blocks = (unsigned)((n + 255) / 256);
    msm_digits_kernel<<<blocks, 256, 0, st>>>((const uint4*)scalars_dev, n, sh, digits, counts);
    B200ZK_LAUNCH_CHECK(ctx, "msm_digits_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_SCAN);
    B200ZK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->msm_scan_tmp.p, scan_bytes, counts, starts, (int)(nbuckets + 1), st));
    ctx->launches++;
    unsigned blocks = (nbuckets + 1 + 255) / 256;
    copy_u32_kernel<<<blocks, 256, 0, st>>>(starts, cursor, nbuckets + 1);
    B200ZK_LAUNCH_CHECK(ctx, "copy_u32_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_SCATTER);
    dim3 grid((unsigned)((n + 255) / 256), sh.W);
    msm_scatter_kernel<<<grid, 256, 0, st>>>(digits, n, sh, cursor, sorted);
    B200ZK_LAUNCH_CHECK(ctx, "msm_scatter_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_ACCUMULATE);
    unsigned blocks = (nbuckets + 127) / 128;
    msm_accumulate_kernel<<<blocks, 128, 0, st>>>(base_ptr, starts, sorted, sh, nbuckets, ctx->msm_buckets.p);
    B200ZK_LAUNCH_CHECK(ctx, "msm_accumulate_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_BIG);
    unsigned blocks = (nbuckets + 255) / 256;
    msm_big_list_kernel<<<blocks, 256, 0, st>>>(starts, nbuckets, sh, plan, big_bucket, big_first, big_cap, chunk_bucket,
                                                chunk_idx);
    B200ZK_LAUNCH_CHECK(ctx, "msm_big_list_kernel");
    const size_t shm = BIG_THREADS * sizeof(G1XYZZ);
    msm_big_accumulate_kernel<<<ctx->sm_count * 2, BIG_THREADS, shm, st>>>(base_ptr, starts, sorted, plan, chunk_bucket,
                                                                         chunk_idx, ctx->msm_big.p);
    B200ZK_LAUNCH_CHECK(ctx, "msm_big_accumulate_kernel");
    msm_big_reduce_kernel<<<ctx->sm_count, BIG_THREADS, shm, st>>>(starts, plan, big_bucket, big_first, big_cap,
                                                                   ctx->msm_big.p, ctx->msm_buckets.p);
    B200ZK_LAUNCH_CHECK(ctx, "msm_big_reduce_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_REDUCE);
    unsigned tot1 = sh.W * n1;
    msm_reduce_l1_kernel<<<(tot1 + 127) / 128, 128, 0, st>>>(ctx->msm_buckets.p, tot1, run1, acc1);
    B200ZK_LAUNCH_CHECK(ctx, "msm_reduce_l1_kernel");
    unsigned tot2 = sh.W * n2;
    msm_reduce_l2_kernel<<<(tot2 + 127) / 128, 128, 0, st>>>(run1, acc1, n1, n2, sh.W, asum, run2, acc2);
    B200ZK_LAUNCH_CHECK(ctx, "msm_reduce_l2_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_FINAL);
    msm_final_kernel<<<1, 64, 0, st>>>(asum, run2, acc2, n2, sh, out_dev, out_kind);
    B200ZK_LAUNCH_CHECK(ctx, "msm_final_kernel");
  }
  return B200ZK_OK;
}

int g1_sum_run(b200zk_ctx* ctx, const void* partials_dev, size_t count, void* out_affine_dev) {
  if (!partials_dev || !out_affine_dev) return B200ZK_ERR_BAD_ARG;
  g1_sum_kernel<<<1, 32, 0, ctx->stream>>>(partials_dev, (unsigned)count, out_affine_dev);
  B200ZK_LAUNCH_CHECK(ctx, "g1_sum_kernel");
  return B200ZK_OK;
}

}  // namespace b200zk
