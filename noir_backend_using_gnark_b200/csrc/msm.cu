// BN254 G1 multi-scalar multiplication for sm_100a: replaces gnark-crypto v0.9.1 ecc/bn254/multiexp.go
// ((*G1Affine).MultiExp: partitionScalars + processChunkG1* + msmReduceChunkG1Affine), reached in the reference
// through kzg.Commit inside plonk.Prove / plonk.Setup (/root/reference/gnark_backend_ffi/backend/plonk/plonk.go:67, :21).
//
// Pippenger with signed c-bit digits.  Two modes:
//   classic      W = ceil(255/c) windows, 2^(c-1) buckets per window, window sums combined by Horner (c doublings each)
//   precomputed  the bases are static (the KZG SRS), so b200zk_bases_precompute stores 2^(c*j) * P_i for every window j
//                once; every (point, window) pair then lands in ONE shared set of 2^(c-1) buckets, which removes the
//                per-window reduction and the Horner tail and lets c grow to ~log2(n)-2 (fewer additions per point)
// Pipeline (all on the context stream, no host synchronisation):
//   1 msm_hist_kernel        scalar: Montgomery -> regular, signed digits, bucket histogram (global atomics); for large
//                            sorts it also parks a (bucket | sign) record per (point, window) pair
//   2 exclusive scan          bucket start offsets (CUB)
//   3 msm_scatter_kernel     counting sort: (table index | sign) grouped by bucket; with more than 2^19 buckets as
//     msm_scatter_pass_kernel bucket-range passes over the parked records, so that the partially written sectors of a
//                            pass stay in L2 (the single pass is bound by DRAM read-modify-writes of its random 4-byte
//                            stores, not by its atomics).  Two other front ends were built and measured no better — a
//                            two-level partition + shared-memory scatter, a CUB radix sort: profiles/r02_msm_frontend.md
//   4a msm_pair_kernel       optional batched-affine pre-summation of the runs (off: slower on this part)
//   4 msm_accumulate_kernel  one thread per bucket walks its run: 64-byte affine gathers (next point prefetched),
//                            extended-Jacobian mixed additions; over-long runs (skewed scalars) are left to
//   4b msm_big_* kernels     which split a run over many CTAs and tree-reduce the partial sums in shared memory
//   5 msm_reduce_*           sum_b b*B[b]: running sums over 32-bucket chunks, then the chunk weights are split in
//                            bit planes which are summed in parallel (warp-shuffle trees) and recombined by doublings
//   6 msm_final_kernel       Horner over windows (classic mode only), canonical affine (or XYZZ partial) output
// The result is a canonical affine point, so it is bit-identical to gnark's for any window size / summation order.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"
#include "g1.cuh"

namespace b200zk {

static constexpr int CHUNK_LOG = 5;            // buckets per running-sum chunk: 2^5 when the reduction is throughput-bound,
static constexpr int CHUNK = 1 << CHUNK_LOG;   // 2^3 for small bucket counts, where the chunk's chain of additions is the latency
static constexpr int CHUNK_LOG_SMALL = 3;
static constexpr int BIG_CHUNK = 8192;         // sorted entries per CTA in the long-run path
static constexpr int BIG_THREADS = 256;
static constexpr size_t TINY_MAX_TERMS = (size_t)1 << 17;
static constexpr unsigned SMALL_TABLE_W = 43;     // 6-bit windows
static constexpr size_t SMALL_TABLE_N = 2048;     // head of a large base set that also gets the narrow-window table  // (points x windows) up to which the bucket pipeline is skipped
static constexpr int MAX_WINDOWS = 64;
static constexpr int SLICE = 4096;             // chunk results per CTA in the bit-plane sums (a power of two)
static constexpr int MAX_PLANES = 24;
static constexpr unsigned NO_KEY = 0x7fffffffu;                   // record of a zero digit in the scatter passes' key stream
// scatter passes: a pass should leave no more partially written 32-byte sectors open (one per bucket) than L2 keeps until
// they are complete.  Measured at 2^24 points, 2^21 buckets: 1 pass 5.75 ms, 5 passes 3.29, 7 passes 3.08, 9 passes 3.36,
// 13 passes 3.96, 17 passes 4.65 (every pass streams all records once more); 2^19 buckets: 2 passes 0.69 against 0.73 ms.
static constexpr size_t SCATTER_OPEN_BYTES = (size_t)10 << 20;
static constexpr unsigned SCATTER_MAX_PASSES = 16;
static constexpr size_t SCATTER_PASS_MIN_ENTRIES = (size_t)1 << 24;   // below: the scatter is a small part of a small MSM

struct MsmShape {
  unsigned c;           // window bits (the widest window in table mode)
  unsigned W;           // number of windows
  unsigned B;           // buckets per bucket set = 2^(c-1)
  unsigned nsets;       // bucket sets: W (classic) or 1 (precomputed)
  unsigned key_stride;  // bucket id = w * key_stride + |digit| - 1     (B or 0)
  unsigned tab_stride;  // table index = w * tab_stride + point index   (0 or table row length)
  unsigned first;       // index of the first base used
  unsigned big_len;     // runs longer than this go to the cooperative path
  unsigned resume;      // 1: the buckets hold the partial sums of an earlier chunk of the same MSM (host-scalar chunks)
  // window w covers scalar bits [wstart[w], wstart[w+1]); classic mode: uniform c-bit windows.  Table mode splits the
  // 255 bits (254 scalar bits + one spare bit that absorbs the last carry) EVENLY over W windows, so the top window
  // is as wide as the others instead of holding only 254 - c*(W-1) bits (which funnels n entries into a few buckets)
  uint8_t wstart[MAX_WINDOWS + 1];
};

static void uniform_windows(MsmShape& sh) {
  for (unsigned w = 0; w <= sh.W && w <= MAX_WINDOWS; w++) {
    unsigned b = w * sh.c;
    sh.wstart[w] = (uint8_t)(b > 255 ? 255 : b);
  }
}

// signed-digit recoding of one scalar; calls f(w, digit) for every window (digit in [-B, B-1])
template <class F>
__device__ __forceinline__ void for_each_digit(const Fr& s, const MsmShape& sh, F f) {
  unsigned carry = 0;
  for (unsigned w = 0; w < sh.W; w++) {
    const unsigned bit = sh.wstart[w];
    const unsigned width = (unsigned)sh.wstart[w + 1] - bit;   // 0 only for classic windows entirely above bit 255
    const unsigned limb = bit >> 5, off = bit & 31;
    unsigned v = 0;
    if (width) {
      unsigned lo = s.l[limb];
      unsigned hi = limb + 1 < 8 ? s.l[limb + 1] : 0u;
      v = (unsigned)((((uint64_t)hi << 32) | lo) >> off) & ((1u << width) - 1u);
    }
    // signed digit in [-2^(cw-1), 2^(cw-1)) for a cw-bit window (classic windows clipped at bit 255 keep cw = c)
    const unsigned cw = width && sh.tab_stride ? width : sh.c;
    const int half = 1 << (cw - 1);
    int d = (int)(v + carry);
    carry = 0;
    if (d >= half) {
      d -= 2 * half;
      carry = 1;
    }
    f(w, d);
  }
}

__device__ __forceinline__ Fr load_scalar_regular(const uint4* __restrict__ scalars, size_t i) {
  // the limbs are indexed dynamically by for_each_digit: keep them addressable
  return fe_from_mont(fe_load<FrParams>(scalars + 2 * i));
}

// ---------------------------------------------------------------------------------------------------
// 1. histogram, 3. scatter
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_hist_kernel(const uint4* __restrict__ scalars, size_t n, MsmShape sh,
                                                       unsigned* __restrict__ counts, unsigned* __restrict__ keys_out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const unsigned active = __ballot_sync(0xffffffffu, i < n);
  if (i >= n) return;
  const unsigned lane = threadIdx.x & 31;
  Fr s = load_scalar_regular(scalars, i);
  for_each_digit(s, sh, [&](unsigned w, int d) {
    const unsigned mag = d < 0 ? (unsigned)(-d) : (unsigned)d;
    const unsigned key = w * sh.key_stride + (mag - 1);
    // window-major (bucket | sign) record for the scatter passes; NO_KEY for a zero digit
    if (keys_out) keys_out[(size_t)w * n + i] = d == 0 ? NO_KEY : (key | (d < 0 ? 0x80000000u : 0u));
    if (w + 1 == sh.W) {
      // the top window holds only 254 - c*(W-1) scalar bits: few distinct buckets, so the lanes of a warp collide
      // on the same counters -> aggregate per warp (one atomic per distinct bucket)
      const unsigned grp = __match_any_sync(active, d != 0 ? key : 0xffffffffu);
      if (d != 0 && lane == (unsigned)(__ffs(grp) - 1)) atomicAdd(&counts[key], (unsigned)__popc(grp));
    } else if (d != 0) {
      atomicAdd(&counts[key], 1u);
    }
  });
}

// Single-pass scatter (bucket sets of up to 2^19 buckets, small sorts): every thread first issues ALL of its point's
// returning atomics (one per window, positions kept in registers) and only then the dependent stores — W atomics in
// flight per thread instead of one.
template <int MAXW>
__global__ void __launch_bounds__(256) msm_scatter_kernel(const uint4* __restrict__ scalars, size_t n, MsmShape sh,
                                                          unsigned* __restrict__ cursor, unsigned* __restrict__ sorted) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const unsigned active = __ballot_sync(0xffffffffu, i < n);
  if (i >= n) return;
  const unsigned lane = threadIdx.x & 31;
  Fr s = load_scalar_regular(scalars, i);
  unsigned pos[MAXW], entry[MAXW];
  unsigned carry = 0;
#pragma unroll
  for (int w = 0; w < MAXW; w++) {
    pos[w] = 0xffffffffu;
    entry[w] = 0;
    if ((unsigned)w < sh.W) {
      const unsigned bit = sh.wstart[w];
      const unsigned width = (unsigned)sh.wstart[w + 1] - bit;
      const unsigned limb = bit >> 5, off = bit & 31;
      unsigned v = 0;
      if (width) {
        unsigned lo = s.l[limb];
        unsigned hi = limb + 1 < 8 ? s.l[limb + 1] : 0u;
        v = (unsigned)((((uint64_t)hi << 32) | lo) >> off) & ((1u << width) - 1u);
      }
      const unsigned cw = width && sh.tab_stride ? width : sh.c;
      const int half = 1 << (cw - 1);
      int d = (int)(v + carry);
      carry = 0;
      if (d >= half) {
        d -= 2 * half;
        carry = 1;
      }
      const unsigned mag = d < 0 ? (unsigned)(-d) : (unsigned)d;
      const unsigned key = (unsigned)w * sh.key_stride + (mag - 1);
      entry[w] = ((unsigned)w * sh.tab_stride + sh.first + (unsigned)i) | (d < 0 ? 0x80000000u : 0u);
      if ((unsigned)w + 1 == sh.W) {
        // warp-aggregated cursor bump for the last window (classic mode: narrow top window, see msm_hist_kernel)
        const unsigned grp = __match_any_sync(active, d != 0 ? key : 0xffffffffu);
        const unsigned leader = (unsigned)(__ffs(grp) - 1);
        unsigned base = 0;
        if (d != 0 && lane == leader) base = atomicAdd(&cursor[key], (unsigned)__popc(grp));
        base = __shfl_sync(grp, base, leader);
        if (d != 0) pos[w] = base + __popc(grp & ((1u << lane) - 1u));
      } else if (d != 0) {
        pos[w] = atomicAdd(&cursor[key], 1u);
      }
    }
  }
#pragma unroll
  for (int w = 0; w < MAXW; w++)
    if (pos[w] != 0xffffffffu) sorted[pos[w]] = entry[w];
}

// Large sorts.  What bounds the scatter above is not its atomics but the 4-byte stores at random positions of an array
// far larger than L2: each becomes a 32-byte read-modify-write in DRAM (11.4 GB moved for 0.8 GB of entries at 2^24; with
// the stores confined to an L2-resident window the same kernel takes 2.2 instead of 5.75 ms).  So the entries are placed
// in gridDim.y PASSES: pass p takes only bucket range p, whose slice of the sorted array (contiguous, the buckets being
// laid out in order) stays in L2 until it is complete and is written back once.  A pass does not repeat the digit
// recoding: it streams the (bucket | sign) records msm_hist_kernel parked window-major (4 B per entry, coalesced).
template <int MAXW>
__global__ void __launch_bounds__(256) msm_scatter_pass_kernel(const unsigned* __restrict__ keys, size_t n, MsmShape sh,
                                                               unsigned keys_per_pass, unsigned* __restrict__ cursor,
                                                               unsigned* __restrict__ sorted) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const unsigned active = __ballot_sync(0xffffffffu, i < n);
  if (i >= n) return;
  const unsigned lane = threadIdx.x & 31;
  const unsigned key_lo = blockIdx.y * keys_per_pass;
  unsigned recs[MAXW];   // all of the point's records in flight at once (streamed: they must not push the slice out of L2)
#pragma unroll
  for (int w = 0; w < MAXW; w++) recs[w] = (unsigned)w < sh.W ? __ldcs(keys + (size_t)w * n + i) : NO_KEY;
#pragma unroll
  for (int w = 0; w < MAXW; w++) {
    if ((unsigned)w >= sh.W) break;
    const unsigned rec = recs[w];
    const unsigned key = rec & 0x7fffffffu;
    const bool hit = key - key_lo < keys_per_pass;             // NO_KEY is above every range
    unsigned pos = 0;
    if (w + 1 == sh.W && sh.tab_stride == 0) {
      // classic mode: the narrow top window has a handful of buckets -> one cursor bump per distinct bucket of the warp
      const unsigned grp = __match_any_sync(active, hit ? key : 0xffffffffu);
      const unsigned leader = (unsigned)(__ffs(grp) - 1);
      if (hit && lane == leader) pos = atomicAdd(&cursor[key], (unsigned)__popc(grp));
      pos = __shfl_sync(grp, pos, leader) + __popc(grp & ((1u << lane) - 1u));
    } else if (hit) {
      pos = atomicAdd(&cursor[key], 1u);
    }
    if (hit) sorted[pos] = (w * sh.tab_stride + sh.first + (unsigned)i) | (rec & 0x80000000u);
  }
}

// ---------------------------------------------------------------------------------------------------
// 4. bucket accumulation
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ G1Affine load_signed(const void* bases, unsigned entry) {
  G1Affine p = g1_load_affine(bases, entry & 0x7fffffffu);
  if (entry & 0x80000000u) p.y = fe_neg(p.y);
  return p;
}

// Element j of a bucket's run.  Gathered: the sorted entry names a (signed) point of the bases / window table.
// LINEAR: the run was pre-summed by the pair rounds below and lies as affine points in a staging buffer; run bounds are
// the padded bucket offsets shifted down by the number of rounds.
template <bool LINEAR>
__device__ __forceinline__ G1Affine run_point(const void* src, const unsigned* __restrict__ sorted, unsigned j) {
  if (LINEAR) return g1_load_affine(src, j);
  return load_signed(src, sorted[j]);
}

// Threads take buckets in order of decreasing run length (`order`), so the 32 runs of a warp have (almost) the same
// length and the longest runs start first: no lane idles while its neighbours finish.
template <bool LINEAR>
__global__ void __launch_bounds__(128, 4) msm_accumulate_kernel(const void* __restrict__ src, const unsigned* __restrict__ starts,
                                                             const unsigned* __restrict__ sorted,
                                                             const unsigned* __restrict__ order, MsmShape sh, unsigned shift,
                                                             unsigned nbuckets, void* __restrict__ buckets) {
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nbuckets) return;
  const unsigned b = order[tid];
  unsigned lo = starts[b] >> shift, hi = starts[b + 1] >> shift;
  G1XYZZ acc = sh.resume ? g1_load_xyzz(buckets, b) : g1_xyzz_inf();
  if (hi - lo <= sh.big_len && hi > lo) {
    // software pipeline of the random 64-byte gathers: the point for step j+1 is loaded into registers while step j
    // adds, and the line for step j+3 is pulled into L2 (the window table is far larger than L2 and the TLB reach)
    G1Affine cur = run_point<LINEAR>(src, sorted, lo);
    for (unsigned j = lo + 1; j < hi; j++) {
      if (!LINEAR && j + 2 < hi) {
        const char* ahead = reinterpret_cast<const char*>(src) + (size_t)(sorted[j + 2] & 0x7fffffffu) * 64;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ahead));
      }
      G1Affine nxt = run_point<LINEAR>(src, sorted, j);
      g1_add_mixed(acc, cur);
      cur = nxt;
    }
    g1_add_mixed(acc, cur);
  }
  g1_store_xyzz(buckets, b, acc);
}

// ---------------------------------------------------------------------------------------------------
// 4a. pair rounds (batched affine additions).  The counting sort pads every bucket's run to a multiple of 2^R entries
//     (pad entries = infinity), so R rounds of  out[o] = in[2o] + in[2o+1]  over the whole entry array pre-sum each run
//     to 1/2^R of its length without any bucket bookkeeping.  An affine addition needs 1/(x2-x1): every lane batches
//     the K denominators of its K slots with Montgomery's trick (forward pass: running product, parked in the output
//     slot; ONE inversion per lane; backward pass: 1/d_k, lambda, the sum) = 5 multiplications + 1 squaring per
//     addition + one Fermat inversion per K additions, against 8 + 2 for the extended-Jacobian mixed addition.
//     Lanes of a warp take consecutive slots (slot = base + 32k + lane): the entry pairs, the staged points of later
//     rounds, the parked products and the sums are all warp-contiguous.
// ---------------------------------------------------------------------------------------------------
static constexpr unsigned PAD_ENTRY = 0xffffffffu;   // never a table index: W*n < 2^31
static constexpr unsigned PAIR_THREADS = 128;
static constexpr unsigned PAIR_KMAX = 512;
enum { PAIR_ADD = 0, PAIR_DBL = 1, PAIR_TAKE_A = 2, PAIR_TAKE_B = 3, PAIR_INF = 4 };

__global__ void msm_pad_counts_kernel(const unsigned* __restrict__ counts, unsigned nbuckets, unsigned mask,
                                      unsigned* __restrict__ padded) {
  unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nbuckets) return;
  padded[b] = b < nbuckets ? (counts[b] + mask) & ~mask : 0u;
}

__global__ void msm_pad_fill_kernel(const unsigned* __restrict__ starts, const unsigned* __restrict__ counts,
                                    unsigned nbuckets, unsigned* __restrict__ sorted) {
  unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  for (unsigned j = starts[b] + counts[b]; j < starts[b + 1]; j++) sorted[j] = PAD_ENTRY;
}

__device__ __forceinline__ G1Affine load_entry(const void* bases, unsigned entry) {
  if (entry == PAD_ENTRY) {
    G1Affine p;
    p.x = fe_zero<FpParams>();
    p.y = fe_zero<FpParams>();
    return p;
  }
  return load_signed(bases, entry);
}

// denominator of the chord / tangent slope of a + b (1 when the sum needs none), and which formula applies
__device__ __forceinline__ int pair_classify(const G1Affine& a, const G1Affine& b, Fp& d) {
  const bool ia = g1_is_inf(a), ib = g1_is_inf(b);
  d = fe_one<FpParams>();
  if (ia || ib) return ia ? (ib ? PAIR_INF : PAIR_TAKE_B) : PAIR_TAKE_A;
  const Fp dx = fe_sub(b.x, a.x);
  if (fe_is_zero(dx)) {
    if (!fe_eq(a.y, b.y)) return PAIR_INF;   // b = -a
    d = fe_dbl(a.y);                          // no 2-torsion on this curve: y != 0
    return PAIR_DBL;
  }
  d = dx;
  return PAIR_ADD;
}

template <bool FIRST>
__device__ __forceinline__ void pair_load(const void* __restrict__ src, const unsigned* __restrict__ sorted, size_t o,
                                          G1Affine& a, G1Affine& b) {
  if (FIRST) {
    const uint2 e = __ldg(reinterpret_cast<const uint2*>(sorted) + o);
    a = load_entry(src, e.x);
    b = load_entry(src, e.y);
  } else {
    a = g1_load_affine(src, 2 * o);
    b = g1_load_affine(src, 2 * o + 1);
  }
}

// FIRST: src = bases / window table, gathered through the sorted entries; otherwise src = the previous round's sums.
// total_ptr -> padded number of sorted entries (device-side: the host never learns it); this round has total >> round slots.
template <bool FIRST>
__global__ void __launch_bounds__(PAIR_THREADS, 4) msm_pair_kernel(const void* __restrict__ src,
                                                                   const unsigned* __restrict__ sorted,
                                                                   const unsigned* __restrict__ total_ptr, unsigned round,
                                                                   unsigned K, void* __restrict__ dst) {
  const size_t nslots = (size_t)(*total_ptr >> round);
  const size_t warp = ((size_t)blockIdx.x * PAIR_THREADS + threadIdx.x) >> 5;
  const size_t base = warp * K * 32 + (threadIdx.x & 31);   // slot of step k = base + 32 k
  if (base >= nslots) return;
  const size_t avail = (nslots - base + 31) / 32;
  const unsigned steps = avail < K ? (unsigned)avail : K;
  char* out = reinterpret_cast<char*>(dst);
  // forward: park the product of the earlier denominators in the slot, extend it by this slot's
  Fp acc = fe_one<FpParams>();
  {
    G1Affine a, b;
    pair_load<FIRST>(src, sorted, base, a, b);
    for (unsigned k = 0; k < steps; k++) {
      const size_t o = base + (size_t)k * 32;
      G1Affine na = a, nb = b;
      if (k + 1 < steps) pair_load<FIRST>(src, sorted, o + 32, na, nb);   // in flight during the product below
      Fp d;
      pair_classify(a, b, d);
      fe_store(out + o * 64, acc);
      acc = fe_mul(acc, d);
      a = na;
      b = nb;
    }
  }
  Fp inv = fe_inv(acc);   // all lanes invert together; no denominator is zero
  {
    G1Affine a, b;
    pair_load<FIRST>(src, sorted, base + (size_t)(steps - 1) * 32, a, b);
    for (unsigned k = steps; k-- > 0;) {
      const size_t o = base + (size_t)k * 32;
      G1Affine na = a, nb = b;
      if (k > 0) pair_load<FIRST>(src, sorted, o - 32, na, nb);
      const Fp pre = fe_load<FpParams>(out + o * 64);
      Fp d;
      const int kind = pair_classify(a, b, d);
      const Fp dinv = fe_mul(inv, pre);   // 1 / d_k
      inv = fe_mul(inv, d);               // 1 / (d_0 ... d_{k-1})
      G1Affine r;
      if (kind <= PAIR_DBL) {
        Fp num, x2 = b.x;
        if (kind == PAIR_DBL) {
          const Fp xx = fe_sqr(a.x);
          num = fe_add(fe_dbl(xx), xx);
          x2 = a.x;
        } else {
          num = fe_sub(b.y, a.y);
        }
        const Fp lam = fe_mul(num, dinv);
        r.x = fe_sub(fe_sub(fe_sqr(lam), a.x), x2);
        r.y = fe_sub(fe_mul(lam, fe_sub(a.x, r.x)), a.y);
      } else if (kind == PAIR_TAKE_A) {
        r = a;
      } else if (kind == PAIR_TAKE_B) {
        r = b;
      } else {
        r.x = fe_zero<FpParams>();
        r.y = fe_zero<FpParams>();
      }
      g1_store_affine(dst, o, r);
      a = na;
      b = nb;
    }
  }
}

// ---- long runs: list them, split over CTAs, reduce --------------------------------------------------
struct BigPlan {
  unsigned nbig;          // number of long runs
  unsigned nchunks;       // total CTA chunks
};

// every long run is cut into BIG_CHUNK-entry chunks; chunk descriptors are (bucket, index within the run)
__global__ void msm_big_list_kernel(const unsigned* __restrict__ starts, unsigned nbuckets, MsmShape sh, unsigned shift,
                                    BigPlan* plan, unsigned* __restrict__ big_bucket, unsigned* __restrict__ big_first_chunk,
                                    unsigned cap, unsigned* __restrict__ chunk_bucket, unsigned* __restrict__ chunk_idx) {
  unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  unsigned len = (starts[b + 1] >> shift) - (starts[b] >> shift);
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

__device__ __forceinline__ G1XYZZ warp_sum_xyzz(G1XYZZ v) {
  // butterfly reduction with warp shuffles: every 32-bit limb of the point travels by __shfl_xor_sync
  for (int d = 16; d > 0; d >>= 1) {
    G1XYZZ o;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      o.x.l[i] = __shfl_xor_sync(0xffffffffu, v.x.l[i], d);
      o.y.l[i] = __shfl_xor_sync(0xffffffffu, v.y.l[i], d);
      o.zz.l[i] = __shfl_xor_sync(0xffffffffu, v.zz.l[i], d);
      o.zzz.l[i] = __shfl_xor_sync(0xffffffffu, v.zzz.l[i], d);
    }
    // both partners compute the same group element (the XYZZ representative may differ between partners,
    // which is fine: only lane 0's value is used)
    g1_add(v, o);
  }
  return v;
}

// grid-stride over chunk descriptors: one WARP sums one chunk of a long run (lane-serial mixed additions, then a
// warp-shuffle tree: 5 full additions of overhead per chunk instead of a shared-memory tree per CTA)
template <bool LINEAR>
__global__ void __launch_bounds__(BIG_THREADS) msm_big_accumulate_kernel(const void* __restrict__ bases,
                                                                         const unsigned* __restrict__ starts,
                                                                         const unsigned* __restrict__ sorted, unsigned shift,
                                                                         const BigPlan* __restrict__ plan,
                                                                         const unsigned* __restrict__ chunk_bucket,
                                                                         const unsigned* __restrict__ chunk_idx,
                                                                         void* __restrict__ partials) {
  const unsigned nchunks = plan->nchunks;
  const unsigned lane = threadIdx.x & 31;
  const unsigned warps_per_cta = BIG_THREADS / 32;
  for (unsigned item = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); item < nchunks; item += gridDim.x * warps_per_cta) {
    const unsigned b = chunk_bucket[item];
    const unsigned lo = starts[b] >> shift, hi = starts[b + 1] >> shift;
    unsigned clo = lo + chunk_idx[item] * BIG_CHUNK;
    unsigned chi = clo + BIG_CHUNK < hi ? clo + BIG_CHUNK : hi;
    G1XYZZ acc = g1_xyzz_inf();
    unsigned j = clo + lane;
    if (j < chi) {
      G1Affine cur = run_point<LINEAR>(bases, sorted, j);
      for (j += 32; j < chi; j += 32) {
        G1Affine nxt = run_point<LINEAR>(bases, sorted, j);  // in flight while the addition below runs
        g1_add_mixed(acc, cur);
        cur = nxt;
      }
      g1_add_mixed(acc, cur);
    }
    acc = warp_sum_xyzz(acc);
    if (lane == 0) g1_store_xyzz(partials, item, acc);
  }
}

// one CTA per long run: sum its chunk partials into the bucket
__global__ void __launch_bounds__(BIG_THREADS) msm_big_reduce_kernel(const unsigned* __restrict__ starts,
                                                                     const BigPlan* __restrict__ plan,
                                                                     const unsigned* __restrict__ big_bucket,
                                                                     const unsigned* __restrict__ big_first_chunk,
                                                                     unsigned cap, const void* __restrict__ partials,
                                                                     void* __restrict__ buckets, unsigned resume, unsigned shift) {
  extern __shared__ uint4 big_smem[];
  G1XYZZ* sh_pts = reinterpret_cast<G1XYZZ*>(big_smem);
  const unsigned nbig = plan->nbig < cap ? plan->nbig : cap;
  for (unsigned slot = blockIdx.x; slot < nbig; slot += gridDim.x) {
    const unsigned b = big_bucket[slot];
    const unsigned len = (starts[b + 1] >> shift) - (starts[b] >> shift);
    const unsigned chunks = (len + BIG_CHUNK - 1) / BIG_CHUNK;
    G1XYZZ acc = g1_xyzz_inf();
    for (unsigned ch = threadIdx.x; ch < chunks; ch += BIG_THREADS) {
      G1XYZZ q = g1_load_xyzz(partials, big_first_chunk[slot] + ch);
      g1_add(acc, q);
    }
    block_reduce_xyzz(acc, sh_pts);
    if (threadIdx.x == 0) {
      if (resume) {
        G1XYZZ prev = g1_load_xyzz(buckets, b);  // the accumulate kernel left the earlier chunks' sum untouched
        g1_add(acc, prev);
      }
      g1_store_xyzz(buckets, b, acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// 5. bucket reduction per bucket set:  S = sum_{b=1..B} b * bucket[b-1]
//    R1: chunk t (32 buckets):  run_t = sum_j X_{32t+j},  acc_t = sum_j (j+1) X_{32t+j}
//        => S = sum_t acc_t + 32 * sum_t t * run_t
//    R2: the weights t are split in bit planes: plane 0 sums acc_t, plane 1+k sums run_t over {t : bit k of t set};
//        each CTA sums one slice of one plane (thread-serial partial sums, then a shared-memory tree)
//    R3: one CTA per set: warp p finishes plane p (lane partial sums + warp-shuffle tree), then
//        S = P_0 + 32 * sum_k 2^k P_{1+k}  by doublings
// ---------------------------------------------------------------------------------------------------
template <int CL>
__global__ void __launch_bounds__(128) msm_reduce_r1_kernel(const void* __restrict__ buckets, unsigned nchunks_total,
                                                            void* __restrict__ run, void* __restrict__ acc_out) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks_total) return;
  G1XYZZ r = g1_xyzz_inf(), a = g1_xyzz_inf();
  for (int j = (1 << CL) - 1; j >= 0; j--) {
    G1XYZZ q = g1_load_xyzz(buckets, ((size_t)t << CL) + j);
    g1_add(r, q);
    g1_add(a, r);
  }
  g1_store_xyzz(run, t, r);
  g1_store_xyzz(acc_out, t, a);
}

// grid = (slices, planes, sets).  A thread visits only the chunk results that belong to its plane (those with bit
// plane-1 of the chunk index set), enumerated directly — no lane idles through the other half — and a CTA covers 4096
// chunk results, so the shared-memory tree (12 warp-wide additions with mostly idle lanes) is paid once per 8-16
// additions of every thread instead of once per 4.
__global__ void __launch_bounds__(BIG_THREADS) msm_reduce_r2_kernel(const void* __restrict__ run, const void* __restrict__ acc_in,
                                                                    unsigned chunks_per_set, unsigned nslices,
                                                                    void* __restrict__ partial) {
  extern __shared__ uint4 big_smem[];
  G1XYZZ* sh_pts = reinterpret_cast<G1XYZZ*>(big_smem);
  const unsigned slice = blockIdx.x, plane = blockIdx.y, set = blockIdx.z;
  const unsigned lo = slice * SLICE;
  const unsigned hi = lo + SLICE < chunks_per_set ? lo + SLICE : chunks_per_set;   // hi - lo is a power of two
  const unsigned len = hi - lo;
  // members of this plane inside [lo, hi): all of them (plane 0, or a bit above the slice that is set), none (such a bit
  // clear), or every second group of 2^b
  unsigned cnt = len, b = 0;
  bool spread = false;
  if (plane) {
    b = plane - 1;
    if ((1u << b) >= len) cnt = ((lo >> b) & 1u) ? len : 0u;
    else {
      cnt = len >> 1;
      spread = true;
    }
  }
  G1XYZZ acc = g1_xyzz_inf();
  const void* src = plane ? run : acc_in;
  for (unsigned i = threadIdx.x; i < cnt; i += BIG_THREADS) {
    const unsigned t = lo + (spread ? (((i >> b) << (b + 1)) | (1u << b) | (i & ((1u << b) - 1u))) : i);
    G1XYZZ q = g1_load_xyzz(src, (size_t)set * chunks_per_set + t);
    g1_add(acc, q);
  }
  block_reduce_xyzz(acc, sh_pts);
  if (threadIdx.x == 0) g1_store_xyzz(partial, ((size_t)set * gridDim.y + plane) * nslices + slice, acc);
}

// grid = sets, block = 8 warps; warp w finishes planes w, w+8, ...; then warp 0 weighs the planes — lane k doubles plane
// k up to its weight 2^(k-1+chunk_log) (all lanes step together, 2.6 us per doubling) and a shuffle tree adds them: about
// 20 dependent doublings + 5 additions instead of a Horner chain of one doubling AND one addition per plane on one thread
__global__ void __launch_bounds__(256) msm_reduce_r3_kernel(const void* __restrict__ partial, unsigned nplanes,
                                                            unsigned nslices, unsigned chunk_log, void* __restrict__ set_sums) {
  __shared__ G1XYZZ plane_sum[MAX_PLANES];
  const unsigned set = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (unsigned plane = warp; plane < nplanes; plane += 8) {
    G1XYZZ acc = g1_xyzz_inf();
    for (unsigned s = lane; s < nslices; s += 32) {
      G1XYZZ q = g1_load_xyzz(partial, ((size_t)set * nplanes + plane) * nslices + s);
      g1_add(acc, q);
    }
    acc = warp_sum_xyzz(acc);
    if (lane == 0) plane_sum[plane] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    G1XYZZ v = lane < nplanes ? plane_sum[lane] : g1_xyzz_inf();
    const unsigned mine = (lane == 0 || lane >= nplanes) ? 0u : lane - 1 + chunk_log;   // S = P_0 + 2^chunk_log sum_k 2^(k-1) P_k
    const unsigned most = nplanes >= 2 ? nplanes - 2 + chunk_log : 0u;
    for (unsigned j = 0; j < most; j++) {
      G1XYZZ d = v;
      g1_double(d);
      if (j < mine) v = d;
    }
    v = warp_sum_xyzz(v);
    if (lane == 0) g1_store_xyzz(set_sums, set, v);
  }
}

// ---------------------------------------------------------------------------------------------------
// 5b. small MSMs against a window table (the sizes of the reference's own test circuits: tens to thousands of
//     points).  Bucketing has nothing to amortise there and the pipeline above is ~17 latency-bound launches, so
//     every (point, window) pair becomes one thread: digit * T[window][point] by double-and-add over the <= 13-bit
//     digit (unsigned windows: no carries), summed by a shared-memory tree per CTA and by one final CTA.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIG_THREADS) msm_tiny_kernel(const void* __restrict__ table,
                                                               const uint4* __restrict__ scalars, unsigned n, MsmShape sh,
                                                               void* __restrict__ partials) {
  extern __shared__ uint4 big_smem[];
  G1XYZZ* sh_pts = reinterpret_cast<G1XYZZ*>(big_smem);
  const unsigned t = blockIdx.x * BIG_THREADS + threadIdx.x;
  G1XYZZ acc = g1_xyzz_inf();
  if (t < n * sh.W) {
    const unsigned w = t / n, i = t - w * n;
    const Fr s = load_scalar_regular(scalars, i);
    const unsigned bit = sh.wstart[w], width = (unsigned)sh.wstart[w + 1] - bit;
    const unsigned limb = bit >> 5, off = bit & 31;
    unsigned d = 0;
    if (width) {
      const unsigned lo = s.l[limb], hi = limb + 1 < 8 ? s.l[limb + 1] : 0u;
      d = (unsigned)((((uint64_t)hi << 32) | lo) >> off) & ((1u << width) - 1u);
    }
    if (d) {
      const G1Affine pt = g1_load_affine(table, (size_t)w * sh.tab_stride + sh.first + i);
      for (int b = 31 - __clz(d); b >= 0; b--) {
        g1_double(acc);
        if ((d >> b) & 1u) g1_add_mixed(acc, pt);
      }
    }
  }
  block_reduce_xyzz(acc, sh_pts);
  if (threadIdx.x == 0) g1_store_xyzz(partials, blockIdx.x, acc);
}

__global__ void __launch_bounds__(BIG_THREADS) msm_tiny_final_kernel(const void* __restrict__ partials, unsigned count,
                                                                     void* __restrict__ out, int out_kind) {
  extern __shared__ uint4 big_smem[];
  G1XYZZ* sh_pts = reinterpret_cast<G1XYZZ*>(big_smem);
  G1XYZZ acc = g1_xyzz_inf();
  for (unsigned k = threadIdx.x; k < count; k += BIG_THREADS) {
    G1XYZZ q = g1_load_xyzz(partials, k);
    g1_add(acc, q);
  }
  block_reduce_xyzz(acc, sh_pts);
  if (threadIdx.x == 0) {
    if (out_kind == 0) g1_store_affine(out, 0, g1_to_affine_single(acc));
    else g1_store_xyzz(out, 0, acc);
  }
}

// ---------------------------------------------------------------------------------------------------
// 6. window Horner (classic mode) + normalisation
// ---------------------------------------------------------------------------------------------------
__global__ void msm_final_kernel(const void* __restrict__ set_sums, MsmShape sh, void* __restrict__ out, int out_kind) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G1XYZZ total = g1_load_xyzz(set_sums, sh.nsets - 1);
  for (int w = (int)sh.nsets - 2; w >= 0; w--) {
    for (unsigned k = 0; k < sh.c; k++) g1_double(total);
    G1XYZZ q = g1_load_xyzz(set_sums, w);
    g1_add(total, q);
  }
  if (out_kind == 0) {
    G1Affine r = g1_to_affine_single(total);
    g1_store_affine(out, 0, r);
  } else {
    g1_store_xyzz(out, 0, total);
  }
}

__global__ void g1_sum_kernel(const void* __restrict__ partials, unsigned count, void* __restrict__ out, int out_kind) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G1XYZZ total = g1_xyzz_inf();
  for (unsigned i = 0; i < count; i++) {
    G1XYZZ q = g1_load_xyzz(partials, i);
    g1_add(total, q);
  }
  if (out_kind == 0) g1_store_affine(out, 0, g1_to_affine_single(total));
  else g1_store_xyzz(out, 0, total);
}

__global__ void iota_u32_kernel(unsigned* __restrict__ dst, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (unsigned)i;
}

__global__ void copy_u32_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------
// precomputation of window multiples: table[j*n + i] = 2^(c*j) * P_i  (affine), j < W; row 0 = the bases.
// One thread per point: c doublings per window, the W-1 results normalised with ONE inversion (Montgomery's trick;
// the per-window numerators and prefix products are parked in `tmp`).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) msm_precompute_kernel(const void* __restrict__ table, size_t n, size_t first,
                                                             size_t count, MsmShape sh, void* __restrict__ tmp,
                                                             void* __restrict__ table_out) {
  const unsigned W = sh.W;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= count) return;
  const size_t i = first + t;
  G1Affine p = g1_load_affine(table, i);
  if (g1_is_inf(p)) {
    for (unsigned j = 1; j < W; j++) g1_store_affine(table_out, j * n + i, p);
    return;
  }
  G1XYZZ acc;
  acc.x = p.x;
  acc.y = p.y;
  acc.zz = fe_one<FpParams>();
  acc.zzz = fe_one<FpParams>();
  Fp pre = fe_one<FpParams>();
  for (unsigned j = 1; j < W; j++) {
    const unsigned steps = (unsigned)sh.wstart[j] - (unsigned)sh.wstart[j - 1];   // row j = 2^wstart[j] * P
    for (unsigned k = 0; k < steps; k++) g1_double(acc);
    Fp zz3 = fe_mul(acc.zz, acc.zzz);
    G1XYZZ q;
    q.x = fe_mul(acc.x, acc.zzz);  // x = X/ZZ  = X*ZZZ / (ZZ*ZZZ)
    q.y = fe_mul(acc.y, acc.zz);   // y = Y/ZZZ = Y*ZZ  / (ZZ*ZZZ)
    q.zz = pre;                    // product of the earlier denominators
    q.zzz = zz3;
    g1_store_xyzz(tmp, (j - 1) * count + t, q);
    pre = fe_mul(pre, zz3);
  }
  Fp inv = fe_inv(pre);  // 1 / (all denominators)
  for (unsigned j = W - 1; j >= 1; j--) {
    G1XYZZ q = g1_load_xyzz(tmp, (j - 1) * count + t);
    Fp einv = fe_mul(inv, q.zz);  // 1 / denominator_j
    G1Affine a;
    a.x = fe_mul(q.x, einv);
    a.y = fe_mul(q.y, einv);
    g1_store_affine(table_out, j * n + i, a);
    inv = fe_mul(inv, q.zzz);
  }
}

// ---------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------
static unsigned choose_window(size_t n) {
  unsigned lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  // fewer, larger windows as n grows; at least 2^5 buckets so the 32-wide reduction chunks are full
  // window sizes whose TOP window still holds >= 7 scalar bits (254 - c*(W-1)): c = 12, 14, 18, 21, 23 leave 1-2 bits
  // there, which funnels n entries into <= 3 buckets (long-run path + atomic contention)
  if (lg >= 23) return 16;
  if (lg >= 19) return 15;
  if (lg >= 15) return 13;
  if (lg >= 11) return 10;
  return 8;
}

// table[j*n + i] = 2^wstart[j] * P_i for W windows splitting the 255 scalar bits evenly (+1 bit for the first 255 % W)
static int build_window_table(b200zk_ctx* ctx, const void* points, size_t n, unsigned W, void** table_out, uint8_t* wstart_out,
                              unsigned* c_out) {
  MsmShape tsh;
  memset(&tsh, 0, sizeof(tsh));
  tsh.W = W;
  {
    const unsigned base = 255 / W, rem = 255 % W;
    unsigned pos = 0;
    for (unsigned w = 0; w < W; w++) {
      tsh.wstart[w] = (uint8_t)pos;
      pos += base + (w < rem ? 1 : 0);
    }
    tsh.wstart[W] = 255;
    tsh.c = base + (rem ? 1 : 0);
  }
  if ((size_t)W * n >= ((size_t)1 << 31)) return B200ZK_ERR_UNSUPPORTED;
  void* table = nullptr;
  cudaError_t e = cudaMalloc(&table, (size_t)W * n * 64);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaMalloc(window table)");
  const size_t chunk = n < ((size_t)1 << 18) ? n : ((size_t)1 << 18);
  void* tmp = nullptr;
  e = cudaMalloc(&tmp, (size_t)(W - 1) * chunk * 128);
  if (e != cudaSuccess) {
    cudaFree(table);
    return set_cuda_error(ctx, e, "cudaMalloc(window table scratch)");
  }
  e = cudaMemcpyAsync(table, points, n * 64, cudaMemcpyDeviceToDevice, ctx->stream);
  for (size_t first = 0; first < n && e == cudaSuccess; first += chunk) {
    const size_t count = first + chunk <= n ? chunk : n - first;
    msm_precompute_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(table, n, first, count, tsh, tmp, table);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) {
    cudaFree(table);
    return set_cuda_error(ctx, e, "msm_precompute_kernel");
  }
  *table_out = table;
  memcpy(wstart_out, tsh.wstart, sizeof(tsh.wstart));
  *c_out = tsh.c;
  return B200ZK_OK;
}

int msm_precompute_run(b200zk_ctx* ctx, b200zk_bases* bases, int c_req) {
  if (!bases || !bases->dev) return B200ZK_ERR_BAD_ARG;
  if (bases->table) return B200ZK_OK;
  const size_t n = bases->n;
  if (n == 0) return B200ZK_OK;
  unsigned lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  // number of windows from the size (or from a requested maximum width c_req)
  unsigned W;
  if (c_req) W = (255 + (unsigned)c_req - 1) / (unsigned)c_req;
  else if (lg >= 23) W = 12;   // widths 22,22,22,21 x9  -> 2^21 buckets (2^23 points: 19.15 ms, 19.27 ms with W = 13)
  else if (lg >= 19) W = 13;   // widths 20 x8, 19 x5     -> 2^19 buckets
  else if (lg >= 15) W = 15;   // widths 17 x15           -> 2^16 buckets
  else if (lg >= 14) W = 16;   // widths 16 x15, 15       -> 2^15 buckets
  else if (lg >= 13) W = 17;   // widths 15 x17           -> 2^14 buckets
  else if (lg >= 11) W = 20;   // widths 13 x15, 12 x5    -> 2^12 buckets (mostly served by the small path)
  else W = SMALL_TABLE_W;      // widths 6               -> the small path multiplies by 6-bit digits: its latency is
                               //                           the double-and-add chain, the table is tiny anyway
  // (measured per size with scripts/table_window_sweep.py: below 2^19 points the bucket count decides the latency —
  //  too few buckets leave the accumulation with a handful of long serial runs, too many make the reduction dominate)
  if (W > MAX_WINDOWS) W = MAX_WINDOWS;
  void* table = nullptr;
  uint8_t wstart[65];
  unsigned c = 0;
  B200ZK_TRY(build_window_table(ctx, bases->dev, n, W, &table, wstart, &c));
  // A large SRS serving a small circuit (the reference's default: 10^6 points, tens of rows): the head of the bases
  // also gets the narrow-window table of the small path, whose latency is the width of the digit.
  if (!c_req && n > SMALL_TABLE_N) {
    void* st = nullptr;
    unsigned sc = 0;
    if (build_window_table(ctx, bases->dev, SMALL_TABLE_N, SMALL_TABLE_W, &st, bases->small_wstart, &sc) == B200ZK_OK) {
      bases->small_table = st;
      bases->small_n = SMALL_TABLE_N;
      bases->small_c = sc;
    }
  }
  bases->table = table;
  bases->tab_c = c;
  bases->tab_W = W;
  memcpy(bases->tab_wstart, wstart, sizeof(wstart));
  return B200ZK_OK;
}

// Number of pair rounds for an MSM with `total` (point, window) entries in `nbuckets` buckets.  Padding a run to a
// multiple of 2^R costs (2^R - 1)/2 infinity entries per bucket (ctx->msm_pair_rounds: -1 automatic, 0 off, r forced).
static unsigned pair_rounds(const b200zk_ctx* ctx, size_t total, unsigned nbuckets) {
  if (total + (size_t)nbuckets * 63 >= ((size_t)1 << 32)) return 0;   // slot arithmetic is 32-bit on the device
  if (ctx->msm_pair_rounds >= 0) return (unsigned)(ctx->msm_pair_rounds > 6 ? 6 : ctx->msm_pair_rounds);
  // Measured on B200 (profiles/r02_batch_affine.md): the rounds are bit-exact but SLOWER than letting the extended-Jacobian
  // walk do everything (2^24 points: 35.5 ms with 3 rounds against 31.2 ms) — the walk needs 64 B of HBM per 9.5
  // multiplications, a pair round 320-670 B per 6.3, and the first round's gathers are fetched twice at 128-byte
  // granularity (60 GB read for 104 M additions).  Automatic therefore means off; the path stays for other parts / sizes.
  (void)nbuckets;
  return 0;
}

// number of windows (= bucket additions per point) msm_run uses for n points of these bases; same rule as below
unsigned msm_window_count(const b200zk_ctx* ctx, const b200zk_bases* bases, size_t n) {
  if (bases->table && !ctx->forced_window) return bases->tab_W;
  unsigned c = ctx->forced_window ? (unsigned)ctx->forced_window : choose_window(n);
  if (c < 6) c = 6;
  if (c > 16) c = 16;
  return (255 + c - 1) / c;
}

int msm_run(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_dev, size_t n,
            void* out_dev, int out_kind, int lane, int part, size_t shape_n) {
  if (!bases || !out_dev || (n && !scalars_dev) || lane < 0 || lane >= MSM_LANES) return B200ZK_ERR_BAD_ARG;
  // part: one MSM fed in several chunks of points (host scalars arriving over PCIe) — MSM_PART_FIRST / MSM_PART_MORE stop
  // after the bucket accumulation, MSM_PART_MORE / MSM_PART_LAST continue from the buckets of the previous chunk; all
  // chunks take their shape (windows, buckets) from shape_n = the size of the whole MSM
  const bool resume = part == MSM_PART_MORE || part == MSM_PART_LAST;
  const bool stop_after_accumulate = part == MSM_PART_FIRST || part == MSM_PART_MORE;
  if (part == MSM_PART_ALL) shape_n = n;
  if (shape_n < n) return B200ZK_ERR_BAD_ARG;
  MsmWorkspace& ws = ctx->ws[lane];
  if (first_base > bases->n || n > bases->n - first_base) return B200ZK_ERR_BAD_ARG;
  if (n >= ((size_t)1 << 31)) return B200ZK_ERR_UNSUPPORTED;
  cudaStream_t st = ws.stream;
  if (n == 0) {
    // empty sum = point at infinity: affine (0,0) / XYZZ with ZZ = 0
    B200ZK_CUDA(ctx, cudaMemsetAsync(out_dev, 0, out_kind == 0 ? 64 : 128, st));
    return B200ZK_OK;
  }
  MsmShape sh;
  memset(&sh, 0, sizeof(sh));
  if (part == MSM_PART_ALL && bases->table && !ctx->forced_window && !ctx->msm_no_tiny &&
      n * (size_t)bases->tab_W <= TINY_MAX_TERMS) {
    // small problem: one thread per (point, window) term of the table, two launches
    const bool head = bases->small_table && first_base + n <= bases->small_n;  // narrow windows: shorter digit chains
    const void* tiny_table = head ? bases->small_table : bases->table;
    sh.c = head ? bases->small_c : bases->tab_c;
    sh.W = head ? SMALL_TABLE_W : bases->tab_W;
    sh.tab_stride = (unsigned)(head ? bases->small_n : bases->n);
    sh.first = (unsigned)first_base;
    memcpy(sh.wstart, head ? bases->small_wstart : bases->tab_wstart, sizeof(sh.wstart));
    const unsigned terms = (unsigned)(n * sh.W);
    const unsigned blocks = (terms + BIG_THREADS - 1) / BIG_THREADS;
    B200ZK_TRY(ensure(ctx, ws.msm_big, (size_t)blocks * 128, st));
    const size_t shm = (size_t)BIG_THREADS * sizeof(G1XYZZ);
    PhaseTimer pt(ctx, PH_MSM_ACCUMULATE, st);
    msm_tiny_kernel<<<blocks, BIG_THREADS, shm, st>>>(tiny_table, (const uint4*)scalars_dev, (unsigned)n, sh, ws.msm_big.p);
    B200ZK_LAUNCH_CHECK(ctx, "msm_tiny_kernel");
    msm_tiny_final_kernel<<<1, BIG_THREADS, shm, st>>>(ws.msm_big.p, blocks, out_dev, out_kind);
    B200ZK_LAUNCH_CHECK(ctx, "msm_tiny_final_kernel");
    return B200ZK_OK;
  }
  // a table is used whenever there is one: even for an MSM over a small part of a large SRS the fixed cost of reducing the
  // table's single bucket set (0.8 ms at 2^19 buckets) is below the classic path's per-window reductions plus its Horner
  // tail of W*c dependent doublings on one thread (1.1 ms)
  const bool use_table = bases->table && !ctx->forced_window;
  const void* base_ptr;
  if (use_table) {
    sh.c = bases->tab_c;
    sh.W = bases->tab_W;
    sh.B = 1u << (sh.c - 1);
    sh.nsets = 1;
    sh.key_stride = 0;
    sh.tab_stride = (unsigned)bases->n;
    memcpy(sh.wstart, bases->tab_wstart, sizeof(sh.wstart));
    base_ptr = bases->table;
  } else {
    sh.c = ctx->forced_window ? (unsigned)ctx->forced_window : choose_window(shape_n);
    if (sh.c < 6) sh.c = 6;
    if (sh.c > 16) sh.c = 16;
    sh.W = (255 + sh.c - 1) / sh.c;  // c*W >= 255: the top window never produces a carry
    sh.B = 1u << (sh.c - 1);
    sh.nsets = sh.W;
    sh.key_stride = sh.B;
    sh.tab_stride = 0;
    uniform_windows(sh);
    base_ptr = bases->dev;
  }
  sh.first = (unsigned)first_base;
  sh.resume = resume ? 1u : 0u;
  const size_t total = n * sh.W;
  if (total >= ((size_t)1 << 32) || sh.W > MAX_WINDOWS) return B200ZK_ERR_UNSUPPORTED;
  const unsigned nbuckets = sh.nsets * sh.B;
  // pair rounds (batched affine pre-summation of the runs, see msm_pair_kernel): R rounds shorten every run 2^R times
  const unsigned R = pair_rounds(ctx, total, nbuckets);
  const unsigned pad = (1u << R) - 1u;
  const size_t total_bound = total + (size_t)nbuckets * pad;   // upper bound of the padded entry count
  {
    size_t avg = (total / nbuckets) >> R;
    size_t bl = 4 * avg + 256;
    sh.big_len = (unsigned)(bl > 0x7fffffff ? 0x7fffffff : bl);
  }
  // short chunks while the bucket count is small enough for the bit-plane sums to stay cheap
  const unsigned chunk_log = (ctx->msm_chunk_log ? (unsigned)ctx->msm_chunk_log
                                                 : ((size_t)nbuckets <= ((size_t)1 << 17) ? CHUNK_LOG_SMALL
                                                    : (size_t)nbuckets <= ((size_t)1 << 19) ? 4 : CHUNK_LOG));
  const unsigned chunks_per_set = sh.B >> chunk_log;
  unsigned chunk_bits = 0;
  while ((1u << chunk_bits) < chunks_per_set) chunk_bits++;
  const unsigned nplanes = 1 + chunk_bits;
  const unsigned nslices = (chunks_per_set + SLICE - 1) / SLICE;
  const unsigned big_cap = nbuckets / 4 + 16;
  const size_t max_big_chunks = total / BIG_CHUNK + big_cap + 1;

  // scatter passes (ctx->msm_scatter_passes: 0 = from the bucket count, 1 = single pass, k = forced)
  unsigned passes = 1;
  if (ctx->msm_scatter_passes > 0) {
    passes = (unsigned)ctx->msm_scatter_passes;
  } else if (total >= SCATTER_PASS_MIN_ENTRIES) {
    const size_t want = ((size_t)nbuckets * 32 + SCATTER_OPEN_BYTES - 1) / SCATTER_OPEN_BYTES;
    passes = (unsigned)(want > SCATTER_MAX_PASSES ? SCATTER_MAX_PASSES : want);
  }
  if (passes > nbuckets) passes = nbuckets;
  if (passes < 1) passes = 1;
  const unsigned keys_per_pass = (nbuckets + passes - 1) / passes;
  unsigned* keys = nullptr;
  if (passes > 1) {
    B200ZK_TRY(ensure(ctx, ws.msm_keys, total * sizeof(unsigned), st));
    keys = (unsigned*)ws.msm_keys.p;
  }
  B200ZK_TRY(ensure(ctx, ws.msm_sorted, total_bound * sizeof(unsigned), st));
  if (R) B200ZK_TRY(ensure(ctx, ws.msm_pairs, ((total_bound >> 1) + (R > 1 ? (total_bound >> 2) : 0) + 2) * 64, st));
  B200ZK_TRY(ensure(ctx, ws.msm_counts, (size_t)(nbuckets + 1) * 4, st));
  B200ZK_TRY(ensure(ctx, ws.msm_starts, (size_t)(nbuckets + 1) * 4, st));
  B200ZK_TRY(ensure(ctx, ws.msm_cursor, (size_t)(nbuckets + 1) * 4, st));
  B200ZK_TRY(ensure(ctx, ws.msm_buckets, (size_t)nbuckets * 128, st));
  B200ZK_TRY(ensure(ctx, ws.msm_tmp,
                    ((size_t)sh.nsets * chunks_per_set * 2 + (size_t)sh.nsets * nplanes * nslices + sh.nsets) * 128));
  B200ZK_TRY(ensure(ctx, ws.msm_small, sizeof(BigPlan) + (size_t)big_cap * 8 + max_big_chunks * 8, st));
  B200ZK_TRY(ensure(ctx, ws.msm_big, max_big_chunks * 128, st));
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned*)nullptr, (unsigned*)nullptr, (int)(nbuckets + 1), st);
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                            (const unsigned*)nullptr, (unsigned*)nullptr, (int)nbuckets, 0, 32, st);
  if (sort_bytes > scan_bytes) scan_bytes = sort_bytes;
  B200ZK_TRY(ensure(ctx, ws.msm_scan_tmp, scan_bytes, st));
  // run-length ordering scratch: [iota | sorted lengths (unused) | order]
  B200ZK_TRY(ensure(ctx, ws.msm_digits, (size_t)nbuckets * 12, st));
  unsigned* iota = (unsigned*)ws.msm_digits.p;
  unsigned* len_sorted = iota + nbuckets;
  unsigned* order = len_sorted + nbuckets;

  unsigned* sorted = (unsigned*)ws.msm_sorted.p;
  unsigned* counts = (unsigned*)ws.msm_counts.p;
  unsigned* starts = (unsigned*)ws.msm_starts.p;
  unsigned* cursor = (unsigned*)ws.msm_cursor.p;
  char* tmp = (char*)ws.msm_tmp.p;
  void* run = tmp;
  void* acc = tmp + (size_t)sh.nsets * chunks_per_set * 128;
  void* partial = tmp + (size_t)sh.nsets * chunks_per_set * 256;
  void* set_sums = (char*)partial + (size_t)sh.nsets * nplanes * nslices * 128;
  BigPlan* plan = (BigPlan*)ws.msm_small.p;
  unsigned* big_bucket = (unsigned*)((char*)ws.msm_small.p + sizeof(BigPlan));
  unsigned* big_first = big_bucket + big_cap;
  unsigned* chunk_bucket = big_first + big_cap;
  unsigned* chunk_idx = chunk_bucket + max_big_chunks;

  B200ZK_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)(nbuckets + 1) * 4, st));
  B200ZK_CUDA(ctx, cudaMemsetAsync(plan, 0, sizeof(BigPlan), st));
  {
    PhaseTimer pt(ctx, PH_MSM_DIGITS, st);
    unsigned blocks = (unsigned)((n + 255) / 256);
    msm_hist_kernel<<<blocks, 256, 0, st>>>((const uint4*)scalars_dev, n, sh, counts, keys);
    B200ZK_LAUNCH_CHECK(ctx, "msm_hist_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_SCAN, st);
    unsigned blocks = (nbuckets + 1 + 255) / 256;
    const unsigned* lens = counts;
    if (R) {   // offsets of runs padded to a multiple of 2^R entries (the cursor array doubles as scratch)
      msm_pad_counts_kernel<<<blocks, 256, 0, st>>>(counts, nbuckets, pad, cursor);
      B200ZK_LAUNCH_CHECK(ctx, "msm_pad_counts_kernel");
      lens = cursor;
    }
    B200ZK_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ws.msm_scan_tmp.p, scan_bytes, lens, starts, (int)(nbuckets + 1), st));
    ctx->launches++;
    copy_u32_kernel<<<blocks, 256, 0, st>>>(starts, cursor, nbuckets + 1);
    B200ZK_LAUNCH_CHECK(ctx, "copy_u32_kernel");
    iota_u32_kernel<<<blocks, 256, 0, st>>>(iota, nbuckets);
    B200ZK_LAUNCH_CHECK(ctx, "iota_u32_kernel");
    size_t sb = ws.msm_scan_tmp.cap;
    B200ZK_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(ws.msm_scan_tmp.p, sb, (const unsigned*)counts, len_sorted,
                                                               (const unsigned*)iota, order, (int)nbuckets, 0, 32, st));
    ctx->launches += 4;
  }
  {
    PhaseTimer pt(ctx, PH_MSM_SCATTER, st);
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (keys) {   // x fastest: all CTAs of pass p are placed before those of pass p + 1
      if (sh.W <= 13)
        msm_scatter_pass_kernel<13><<<dim3(blocks, passes), 256, 0, st>>>(keys, n, sh, keys_per_pass, cursor, sorted);
      else if (sh.W <= 20)
        msm_scatter_pass_kernel<20><<<dim3(blocks, passes), 256, 0, st>>>(keys, n, sh, keys_per_pass, cursor, sorted);
      else
        msm_scatter_pass_kernel<MAX_WINDOWS><<<dim3(blocks, passes), 256, 0, st>>>(keys, n, sh, keys_per_pass, cursor, sorted);
    } else if (sh.W <= 13)
      msm_scatter_kernel<13><<<blocks, 256, 0, st>>>((const uint4*)scalars_dev, n, sh, cursor, sorted);
    else if (sh.W <= 20)
      msm_scatter_kernel<20><<<blocks, 256, 0, st>>>((const uint4*)scalars_dev, n, sh, cursor, sorted);
    else
      msm_scatter_kernel<MAX_WINDOWS><<<blocks, 256, 0, st>>>((const uint4*)scalars_dev, n, sh, cursor, sorted);
    B200ZK_LAUNCH_CHECK(ctx, "msm_scatter_kernel");
    if (R) {
      msm_pad_fill_kernel<<<(nbuckets + 255) / 256, 256, 0, st>>>(starts, counts, nbuckets, sorted);
      B200ZK_LAUNCH_CHECK(ctx, "msm_pad_fill_kernel");
    }
  }
  const void* run_src = base_ptr;   // what the run walkers read: the bases / table, or the last pair round's sums
  {
    PhaseTimer pt(ctx, PH_MSM_ACCUMULATE, st);
    if (R) {
      // staging: round 1 -> A, round 2: A -> B, round 3: B -> A, ...  (A holds total/2 points, B total/4)
      char* bufA = (char*)ws.msm_pairs.p;
      char* bufB = bufA + ((total_bound >> 1) + 1) * 64;
      const size_t cap_threads = (size_t)ctx->sm_count * 4 * PAIR_THREADS;   // one resident wave
      const unsigned kmax = ctx->msm_pair_kmax > 0 ? (unsigned)ctx->msm_pair_kmax : PAIR_KMAX;
      for (unsigned r = 1; r <= R; r++) {
        const size_t slots = (total_bound >> r) + 1;
        const size_t waves = (slots + cap_threads * kmax - 1) / (cap_threads * kmax);
        unsigned K = (unsigned)((slots + cap_threads * waves - 1) / (cap_threads * waves));
        if (K < 1) K = 1;
        const size_t warps = (slots + (size_t)K * 32 - 1) / ((size_t)K * 32);
        const unsigned blocks = (unsigned)((warps * 32 + PAIR_THREADS - 1) / PAIR_THREADS);
        void* dst = (r & 1) ? bufA : bufB;
        if (r == 1)
          msm_pair_kernel<true><<<blocks, PAIR_THREADS, 0, st>>>(base_ptr, sorted, starts + nbuckets, r, K, dst);
        else
          msm_pair_kernel<false><<<blocks, PAIR_THREADS, 0, st>>>(run_src, nullptr, starts + nbuckets, r, K, dst);
        B200ZK_LAUNCH_CHECK(ctx, "msm_pair_kernel");
        run_src = dst;
      }
    }
    unsigned blocks = (nbuckets + 127) / 128;
    if (R)
      msm_accumulate_kernel<true><<<blocks, 128, 0, st>>>(run_src, starts, sorted, order, sh, R, nbuckets, ws.msm_buckets.p);
    else
      msm_accumulate_kernel<false><<<blocks, 128, 0, st>>>(run_src, starts, sorted, order, sh, 0, nbuckets, ws.msm_buckets.p);
    B200ZK_LAUNCH_CHECK(ctx, "msm_accumulate_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_BIG, st);
    unsigned blocks = (nbuckets + 255) / 256;
    msm_big_list_kernel<<<blocks, 256, 0, st>>>(starts, nbuckets, sh, R, plan, big_bucket, big_first, big_cap, chunk_bucket,
                                                chunk_idx);
    B200ZK_LAUNCH_CHECK(ctx, "msm_big_list_kernel");
    const size_t shm = BIG_THREADS * sizeof(G1XYZZ);
    if (R)
      msm_big_accumulate_kernel<true><<<ctx->sm_count * 4, BIG_THREADS, 0, st>>>(run_src, starts, sorted, R, plan, chunk_bucket,
                                                                               chunk_idx, ws.msm_big.p);
    else
      msm_big_accumulate_kernel<false><<<ctx->sm_count * 4, BIG_THREADS, 0, st>>>(run_src, starts, sorted, 0, plan,
                                                                                chunk_bucket, chunk_idx, ws.msm_big.p);
    B200ZK_LAUNCH_CHECK(ctx, "msm_big_accumulate_kernel");
    msm_big_reduce_kernel<<<ctx->sm_count, BIG_THREADS, shm, st>>>(starts, plan, big_bucket, big_first, big_cap,
                                                                   ws.msm_big.p, ws.msm_buckets.p, sh.resume, R);
    B200ZK_LAUNCH_CHECK(ctx, "msm_big_reduce_kernel");
  }
  if (stop_after_accumulate) return B200ZK_OK;  // the next chunk of this MSM continues from ws.msm_buckets
  {
    PhaseTimer pt(ctx, PH_MSM_REDUCE, st);
    const unsigned tot1 = sh.nsets * chunks_per_set;
    if (chunk_log == CHUNK_LOG_SMALL)
      msm_reduce_r1_kernel<CHUNK_LOG_SMALL><<<(tot1 + 127) / 128, 128, 0, st>>>(ws.msm_buckets.p, tot1, run, acc);
    else if (chunk_log == 4)
      msm_reduce_r1_kernel<4><<<(tot1 + 127) / 128, 128, 0, st>>>(ws.msm_buckets.p, tot1, run, acc);
    else
      msm_reduce_r1_kernel<CHUNK_LOG><<<(tot1 + 127) / 128, 128, 0, st>>>(ws.msm_buckets.p, tot1, run, acc);
    B200ZK_LAUNCH_CHECK(ctx, "msm_reduce_r1_kernel");
    const size_t shm = BIG_THREADS * sizeof(G1XYZZ);
    dim3 grid2(nslices, nplanes, sh.nsets);
    msm_reduce_r2_kernel<<<grid2, BIG_THREADS, shm, st>>>(run, acc, chunks_per_set, nslices, partial);
    B200ZK_LAUNCH_CHECK(ctx, "msm_reduce_r2_kernel");
    msm_reduce_r3_kernel<<<sh.nsets, 256, 0, st>>>(partial, nplanes, nslices, chunk_log, set_sums);
    B200ZK_LAUNCH_CHECK(ctx, "msm_reduce_r3_kernel");
  }
  {
    PhaseTimer pt(ctx, PH_MSM_FINAL, st);
    msm_final_kernel<<<1, 32, 0, st>>>(set_sums, sh, out_dev, out_kind);
    B200ZK_LAUNCH_CHECK(ctx, "msm_final_kernel");
  }
  return B200ZK_OK;
}

int g1_sum_run(b200zk_ctx* ctx, const void* partials_dev, size_t count, void* out_dev, int out_kind) {
  if (!partials_dev || !out_dev) return B200ZK_ERR_BAD_ARG;
  g1_sum_kernel<<<1, 32, 0, ctx->stream>>>(partials_dev, (unsigned)count, out_dev, out_kind);
  B200ZK_LAUNCH_CHECK(ctx, "g1_sum_kernel");
  return B200ZK_OK;
}

}  // namespace b200zk
