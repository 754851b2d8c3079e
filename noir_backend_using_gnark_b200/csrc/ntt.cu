// fr NTT for sm_100a: replaces gnark-crypto v0.9.1 ecc/bn254/fr/fft (Domain.FFT / FFTInverse / BitReverse),
// reached in the reference from plonk.Prove / plonk.Setup
// (/root/reference/gnark_backend_ffi/backend/plonk/plonk.go:67, :21).
//
// Structure: the log2(N) radix-2 stages are grouped into passes of k <= 8..10 stages.  One CTA stages a tile of
// 2^k "rows" x 2^cb contiguous "columns" through shared memory, runs the k stages there, and writes the tile
// back; a transform is ceil(log2 N / k) passes over HBM.  Twiddles come from one full table w^i (i < N/2) per
// domain and direction, so a butterfly costs exactly one Montgomery multiplication (the kernel is bound by the
// integer multiplier, not by HBM: see DESIGN.md).  Coset pre-scaling (5^i), the 1/n post-scaling and the coset
// post-scaling (5^-i / n) are fused into the first / last pass; 5^e is formed from a two-level table.
#include "common.cuh"
#include "consts.cuh"
#include "field.cuh"

namespace b200zk {

static constexpr int COSET_LO_BITS = 12;

// ---------------------------------------------------------------------------------------------------
// domain setup
// ---------------------------------------------------------------------------------------------------
struct DomainSeeds {
  // pw[d][k] = base_d^(2^k): d = 0 w, 1 w^-1, 2 coset, 3 coset^-1
  Fr pw[4][32];
  Fr ninv;
};

__global__ void ntt_seed_kernel(DomainSeeds* out, unsigned log2n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fr w, wi, g, gi, h;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    w.l[i] = FR_ROOT28[i];
    wi.l[i] = FR_ROOT28_INV[i];
    g.l[i] = FR_COSET[i];
    gi.l[i] = FR_COSET_INV[i];
    h.l[i] = FR_INV2[i];
  }
  for (unsigned k = log2n; k < B200ZK_MAX_LOG2N; k++) {
    w = fe_sqr(w);
    wi = fe_sqr(wi);
  }
  Fr ninv = fe_one<FrParams>();
  for (unsigned k = 0; k < log2n; k++) ninv = fe_mul(ninv, h);
  out->ninv = ninv;
  for (int k = 0; k < 32; k++) {
    out->pw[0][k] = w;
    out->pw[1][k] = wi;
    out->pw[2][k] = g;
    out->pw[3][k] = gi;
    w = fe_sqr(w);
    wi = fe_sqr(wi);
    g = fe_sqr(g);
    gi = fe_sqr(gi);
  }
}

// table[i] = base^(i << shift) (* ninv if with_ninv), base given by its 2^k powers
__global__ void ntt_fill_pow_kernel(uint4* table, size_t count, const DomainSeeds* seeds, int which, int shift,
                                    int with_ninv) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fr acc = with_ninv ? seeds->ninv : fe_one<FrParams>();
  size_t e = i;
  for (int k = shift; e != 0; k++, e >>= 1) {
    if (e & 1) acc = fe_mul(acc, seeds->pw[which][k]);
  }
  fe_store(table + 2 * i, acc);
}

static int build_domain(b200zk_ctx* ctx, unsigned log2n) {
  NttDomain& d = ctx->domains[log2n];
  if (d.ready) return B200ZK_OK;
  const size_t N = (size_t)1 << log2n;
  const size_t ntw = N / 2 ? N / 2 : 1;
  const size_t nlo = (size_t)1 << (log2n < (unsigned)COSET_LO_BITS ? log2n : COSET_LO_BITS);
  const size_t nhi = log2n > (unsigned)COSET_LO_BITS ? (size_t)1 << (log2n - COSET_LO_BITS) : 1;
  DomainSeeds* seeds = nullptr;
  B200ZK_CUDA(ctx, cudaMalloc(&seeds, sizeof(DomainSeeds)));
  B200ZK_CUDA(ctx, cudaMalloc(&d.tw_fwd, ntw * 32));
  B200ZK_CUDA(ctx, cudaMalloc(&d.tw_inv, ntw * 32));
  B200ZK_CUDA(ctx, cudaMalloc(&d.coset_lo, nlo * 32));
  B200ZK_CUDA(ctx, cudaMalloc(&d.coset_inv_lo, nlo * 32));
  B200ZK_CUDA(ctx, cudaMalloc(&d.coset_hi, nhi * 32));
  B200ZK_CUDA(ctx, cudaMalloc(&d.coset_inv_hi, nhi * 32));
  B200ZK_CUDA(ctx, cudaMalloc(&d.scalars, 32));
  ntt_seed_kernel<<<1, 32, 0, ctx->stream>>>(seeds, log2n);
  B200ZK_LAUNCH_CHECK(ctx, "ntt_seed_kernel");
  auto fill = [&](void* t, size_t cnt, int which, int shift, int with_ninv) -> int {
    unsigned blocks = (unsigned)((cnt + 255) / 256);
    ntt_fill_pow_kernel<<<blocks, 256, 0, ctx->stream>>>((uint4*)t, cnt, seeds, which, shift, with_ninv);
    B200ZK_LAUNCH_CHECK(ctx, "ntt_fill_pow_kernel");
    return B200ZK_OK;
  };
  B200ZK_TRY(fill(d.tw_fwd, ntw, 0, 0, 0));
  B200ZK_TRY(fill(d.tw_inv, ntw, 1, 0, 0));
  B200ZK_TRY(fill(d.coset_lo, nlo, 2, 0, 0));
  B200ZK_TRY(fill(d.coset_inv_lo, nlo, 3, 0, 1));
  B200ZK_TRY(fill(d.coset_hi, nhi, 2, COSET_LO_BITS, 0));
  B200ZK_TRY(fill(d.coset_inv_hi, nhi, 3, COSET_LO_BITS, 0));
  B200ZK_TRY(fill(d.scalars, 1, 0, 0, 1));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  B200ZK_CUDA(ctx, cudaFree(seeds));
  d.log2n = log2n;
  d.ready = true;
  return B200ZK_OK;
}

int ntt_prepare(b200zk_ctx* ctx, unsigned log2n) {
  if (log2n > B200ZK_MAX_LOG2N) return B200ZK_ERR_BAD_ARG;
  return build_domain(ctx, log2n);
}

void ntt_free_domains(b200zk_ctx* ctx) {
  for (auto& d : ctx->domains) {
    if (!d.ready) continue;
    cudaFree(d.tw_fwd);
    cudaFree(d.tw_inv);
    cudaFree(d.coset_lo);
    cudaFree(d.coset_hi);
    cudaFree(d.coset_inv_lo);
    cudaFree(d.coset_inv_hi);
    cudaFree(d.scalars);
    d = NttDomain();
  }
}

// ---------------------------------------------------------------------------------------------------
// the pass kernel
// ---------------------------------------------------------------------------------------------------
enum ScaleMode { SCALE_NONE = 0, SCALE_PRE_COSET = 1, SCALE_POST_NINV = 2, SCALE_POST_COSET = 3 };

// A pass works on a LOCAL array of 2^nl elements which is a slice of the logical 2^n-point vector: the logical
// index is the local index with the `gap_bits`-wide field `gap_val` inserted at bit `gap_pos` (gap_bits = 0 on one
// GPU; the rank id for the two halves of the multi-GPU four-step transform).  Twiddle exponents and coset powers
// are functions of the logical index; tiles and addresses are functions of the local one.  Optionally the source
// or destination array is stored with the bit fields [high | mid(b bits) | low(a bits)] rotated to
// [mid | high | low] — the block layout an all-to-all delivers / expects.
struct PassParams {
  unsigned n;        // logical log2 N
  unsigned nl;       // local log2 size
  unsigned L;        // local index bits below the row bits
  unsigned k;        // stages in this pass = row bits
  unsigned cb;       // column bits; tile = 2^(k+cb) elements
  unsigned gap_pos, gap_bits, gap_val;
  unsigned perm_a, perm_b;
  int src_perm, dst_perm;
  int scale;         // ScaleMode
  int scale_bitrev;  // table index is the bit-reversed logical index
  const uint4* tw;
  const uint4* lo;
  const uint4* hi;
  const uint4* ninv;
  // fused four-step exchange: the tile is stored straight into the peers' exchange buffers over NVLink (peer-mapped
  // pointers, one per rank) instead of into `dst`; destination rank = top bits of the (permuted) local address
  unsigned peer_on, peer_chunk_log, peer_rank;
  uint4* peer_dst[8];
};

__device__ __forceinline__ uint4* store_target(const PassParams& p, uint4* dst, size_t addr) {
  if (!p.peer_on) return dst + 2 * addr;
  const size_t peer = addr >> p.peer_chunk_log;
  const size_t off = ((size_t)p.peer_rank << p.peer_chunk_log) | (addr & (((size_t)1 << p.peer_chunk_log) - 1));
  return p.peer_dst[peer] + 2 * off;
}

__device__ __forceinline__ size_t tile_to_local(const PassParams& p, size_t tile, unsigned r, unsigned c) {
  const unsigned Lc = p.L < p.cb ? p.L : p.cb;
  const size_t low_c = c & ((1u << Lc) - 1u);
  const size_t high_c = c >> Lc;
  const size_t tile_low = tile & (((size_t)1 << (p.L - Lc)) - 1);
  const size_t tile_high = tile >> (p.L - Lc);
  return ((tile_low << Lc) | low_c) | ((size_t)r << p.L) | (((tile_high << (p.cb - Lc)) | high_c) << (p.L + p.k));
}

__device__ __forceinline__ size_t local_to_logical(const PassParams& p, size_t l) {
  if (p.gap_bits == 0) return l;
  const size_t low = l & (((size_t)1 << p.gap_pos) - 1);
  return ((l >> p.gap_pos) << (p.gap_pos + p.gap_bits)) | ((size_t)p.gap_val << p.gap_pos) | low;
}

__device__ __forceinline__ size_t permuted(const PassParams& p, size_t l) {
  const size_t low = l & (((size_t)1 << p.perm_a) - 1);
  const size_t mid = (l >> p.perm_a) & (((size_t)1 << p.perm_b) - 1);
  const size_t high = l >> (p.perm_a + p.perm_b);
  return (mid << (p.nl - p.perm_b)) | (high << p.perm_a) | low;
}

__device__ __forceinline__ Fr coset_factor(const PassParams& p, size_t idx) {
  size_t e = idx;
  if (p.scale_bitrev) e = (size_t)(__brev((unsigned)idx) >> (32 - p.n));
  Fr f = fe_load_ro<FrParams>(p.lo + 2 * (e & ((1u << COSET_LO_BITS) - 1u)));
  if (p.n > (unsigned)COSET_LO_BITS) {
    Fr h = fe_load_ro<FrParams>(p.hi + 2 * (e >> COSET_LO_BITS));
    f = fe_mul(f, h);
  }
  return f;
}

// Shared layout: two planes of 16-byte half elements, so a warp's LDS.128 / STS.128 are conflict free.
template <bool DIT>
__global__ void __launch_bounds__(512) ntt_pass_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                       PassParams p) {
  extern __shared__ uint4 smem[];
  const unsigned tlog = p.k + p.cb;
  const unsigned T = 1u << tlog;
  uint4* s_lo = smem;
  uint4* s_hi = smem + T;
  const size_t tile = blockIdx.x;
  const unsigned nthreads = blockDim.x;  // T/2 (or 1)
  const unsigned cmask = (1u << p.cb) - 1u;

  // ---- load tile (with optional coset pre-scaling)
  for (unsigned e = threadIdx.x; e < T; e += nthreads) {
    const unsigned r = e >> p.cb, c = e & cmask;
    const size_t l = tile_to_local(p, tile, r, c);
    const size_t addr = p.src_perm ? permuted(p, l) : l;
    Fr v = fe_load<FrParams>(src + 2 * addr);
    if (p.scale == SCALE_PRE_COSET) v = fe_mul(v, coset_factor(p, local_to_logical(p, l)));
    s_lo[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    s_hi[e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
  }
  __syncthreads();

  // ---- k radix-2 stages in shared memory
  for (unsigned t = 0; t < p.k; t++) {
    // DIF walks row distance 2^(k-1) .. 1, DIT walks 1 .. 2^(k-1)
    const unsigned dbit = DIT ? t : (p.k - 1 - t);
    const unsigned hl = p.L + dbit;                                    // local half-distance bit
    const unsigned hbit = hl + (hl >= p.gap_pos ? p.gap_bits : 0u);    // logical half-distance = 2^hbit
    for (unsigned b = threadIdx.x; b < T / 2; b += nthreads) {
      const unsigned c = b & cmask;
      const unsigned rb = b >> p.cb;
      const unsigned r0 = ((rb >> dbit) << (dbit + 1)) | (rb & ((1u << dbit) - 1u));
      const unsigned r1 = r0 | (1u << dbit);
      const unsigned e0 = (r0 << p.cb) | c, e1 = (r1 << p.cb) | c;
      Fr u, v;
      {
        uint4 x = s_lo[e0], y = s_hi[e0];
        u.l[0] = x.x; u.l[1] = x.y; u.l[2] = x.z; u.l[3] = x.w;
        u.l[4] = y.x; u.l[5] = y.y; u.l[6] = y.z; u.l[7] = y.w;
        x = s_lo[e1]; y = s_hi[e1];
        v.l[0] = x.x; v.l[1] = x.y; v.l[2] = x.z; v.l[3] = x.w;
        v.l[4] = y.x; v.l[5] = y.y; v.l[6] = y.z; v.l[7] = y.w;
      }
      const size_t g0 = local_to_logical(p, tile_to_local(p, tile, r0, c));
      const size_t j = g0 & (((size_t)1 << hbit) - 1);
      const size_t ex = j << (p.n - 1 - hbit);  // exponent of w, < N/2
      Fr x0, x1;
      if (DIT) {
        if (hbit != 0) v = fe_mul(v, fe_load_ro<FrParams>(p.tw + 2 * ex));
        x0 = fe_add(u, v);
        x1 = fe_sub(u, v);
      } else {
        x0 = fe_add(u, v);
        x1 = fe_sub(u, v);
        if (hbit != 0) x1 = fe_mul(x1, fe_load_ro<FrParams>(p.tw + 2 * ex));
      }
      s_lo[e0] = make_uint4(x0.l[0], x0.l[1], x0.l[2], x0.l[3]);
      s_hi[e0] = make_uint4(x0.l[4], x0.l[5], x0.l[6], x0.l[7]);
      s_lo[e1] = make_uint4(x1.l[0], x1.l[1], x1.l[2], x1.l[3]);
      s_hi[e1] = make_uint4(x1.l[4], x1.l[5], x1.l[6], x1.l[7]);
    }
    __syncthreads();
  }

  // ---- store tile (with optional post-scaling)
  for (unsigned e = threadIdx.x; e < T; e += nthreads) {
    const unsigned r = e >> p.cb, c = e & cmask;
    const size_t l = tile_to_local(p, tile, r, c);
    uint4 x = s_lo[e], y = s_hi[e];
    Fr v;
    v.l[0] = x.x; v.l[1] = x.y; v.l[2] = x.z; v.l[3] = x.w;
    v.l[4] = y.x; v.l[5] = y.y; v.l[6] = y.z; v.l[7] = y.w;
    if (p.scale == SCALE_POST_NINV) v = fe_mul(v, fe_load_ro<FrParams>(p.ninv));
    if (p.scale == SCALE_POST_COSET) v = fe_mul(v, coset_factor(p, local_to_logical(p, l)));
    const size_t addr = p.dst_perm ? permuted(p, l) : l;
    fe_store(store_target(p, dst, addr), v);
  }
}

// ---------------------------------------------------------------------------------------------------
// radix-4 pass kernel: every thread keeps 4 elements in registers and runs TWO stages per shared-memory round trip
// (3 twiddle loads per 4 butterflies: the second stage's two butterflies share one twiddle).  The first round reads
// its elements straight from global memory and the last one writes straight back, so a pass of k stages makes
// ceil(k/2)-1 shared-memory exchanges instead of k+1.  Same PassParams / tiling as ntt_pass_kernel.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void smem_put(uint4* s_lo, uint4* s_hi, unsigned e, const Fr& v) {
  s_lo[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  s_hi[e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr smem_get(const uint4* s_lo, const uint4* s_hi, unsigned e) {
  uint4 x = s_lo[e], y = s_hi[e];
  Fr v;
  v.l[0] = x.x; v.l[1] = x.y; v.l[2] = x.z; v.l[3] = x.w;
  v.l[4] = y.x; v.l[5] = y.y; v.l[6] = y.z; v.l[7] = y.w;
  return v;
}

// twiddle for the butterfly whose lower element has logical index g, at local half-distance bit hl
__device__ __forceinline__ Fr stage_twiddle(const PassParams& p, size_t g, unsigned hl, bool& is_one) {
  const unsigned hbit = hl + (hl >= p.gap_pos ? p.gap_bits : 0u);
  is_one = hbit == 0;
  if (is_one) return fe_one<FrParams>();
  const size_t j = g & (((size_t)1 << hbit) - 1);
  return fe_load_ro<FrParams>(p.tw + 2 * (j << (p.n - 1 - hbit)));
}

template <bool DIT>
__global__ void __launch_bounds__(256, 3) ntt_pass4_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                        PassParams p) {
  extern __shared__ uint4 smem[];
  const unsigned tlog = p.k + p.cb;
  const unsigned T = 1u << tlog;
  uint4* s_lo = smem;
  uint4* s_hi = smem + T;
  const size_t tile = blockIdx.x;
  const unsigned u = threadIdx.x;  // T/4 threads
  const unsigned cmask = (1u << p.cb) - 1u;
  const unsigned c = u & cmask;
  const unsigned nrounds = (p.k + 1) >> 1;
  Fr a[4];
  unsigned rows[4];

  for (unsigned rd = 0; rd < nrounds; rd++) {
    // dbits handled by this round: a pair (q+1, q), or a single bit q when k is odd (last DIF / last DIT round)
    unsigned q;
    bool pair;
    if (DIT) {
      q = 2 * rd;
      pair = q + 1 < p.k;
    } else {
      pair = p.k >= 2 * (rd + 1);
      q = pair ? p.k - 2 * (rd + 1) : 0;
    }
    if (pair) {
      const unsigned rb = u >> p.cb;  // k-2 bits
      const unsigned base = ((rb >> q) << (q + 2)) | (rb & ((1u << q) - 1u));
#pragma unroll
      for (int m = 0; m < 4; m++) rows[m] = base | ((unsigned)m << q);
    } else {
      // two independent radix-2 butterflies per thread: (rows[0], rows[1]) and (rows[2], rows[3])
      const unsigned rb = (u >> p.cb) << 1;  // butterfly ids rb, rb+1 (k-1 bits each)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const unsigned b = rb + h;
        const unsigned r0 = ((b >> q) << (q + 1)) | (b & ((1u << q) - 1u));
        rows[2 * h] = r0;
        rows[2 * h + 1] = r0 | (1u << q);
      }
    }
    // ---- fetch the 4 elements
    if (rd == 0) {
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const size_t l = tile_to_local(p, tile, rows[m], c);
        const size_t addr = p.src_perm ? permuted(p, l) : l;
        a[m] = fe_load<FrParams>(src + 2 * addr);
        if (p.scale == SCALE_PRE_COSET) a[m] = fe_mul(a[m], coset_factor(p, local_to_logical(p, l)));
      }
    } else {
      __syncthreads();
#pragma unroll
      for (int m = 0; m < 4; m++) a[m] = smem_get(s_lo, s_hi, (rows[m] << p.cb) | c);
    }
    // ---- butterflies
    const size_t g0 = local_to_logical(p, tile_to_local(p, tile, rows[0], c));
    bool one;
    if (pair) {
      const size_t g1 = local_to_logical(p, tile_to_local(p, tile, rows[1], c));
      const unsigned h_lo = p.L + q, h_hi = p.L + q + 1;
      if (DIT) {
        // stage q: (a0,a1), (a2,a3) share one twiddle; stage q+1: (a0,a2) and (a1,a3)
        Fr t = stage_twiddle(p, g0, h_lo, one);
        Fr v1 = one ? a[1] : fe_mul(a[1], t);
        Fr v3 = one ? a[3] : fe_mul(a[3], t);
        Fr b0 = fe_add(a[0], v1), b1 = fe_sub(a[0], v1), b2 = fe_add(a[2], v3), b3 = fe_sub(a[2], v3);
        Fr t0 = stage_twiddle(p, g0, h_hi, one);
        Fr t1 = stage_twiddle(p, g1, h_hi, one);
        Fr w2 = fe_mul(b2, t0), w3 = fe_mul(b3, t1);
        a[0] = fe_add(b0, w2); a[2] = fe_sub(b0, w2);
        a[1] = fe_add(b1, w3); a[3] = fe_sub(b1, w3);
      } else {
        // stage q+1: (a0,a2) and (a1,a3); stage q: (a0,a1), (a2,a3) share one twiddle
        Fr t0 = stage_twiddle(p, g0, h_hi, one);
        Fr t1 = stage_twiddle(p, g1, h_hi, one);
        Fr b0 = fe_add(a[0], a[2]), b2 = fe_mul(fe_sub(a[0], a[2]), t0);
        Fr b1 = fe_add(a[1], a[3]), b3 = fe_mul(fe_sub(a[1], a[3]), t1);
        Fr t = stage_twiddle(p, g0, h_lo, one);
        a[0] = fe_add(b0, b1);
        a[1] = fe_sub(b0, b1);
        a[2] = fe_add(b2, b3);
        a[3] = fe_sub(b2, b3);
        if (!one) {
          a[1] = fe_mul(a[1], t);
          a[3] = fe_mul(a[3], t);
        }
      }
    } else {
      const size_t g2 = local_to_logical(p, tile_to_local(p, tile, rows[2], c));
      const unsigned hl = p.L + q;
      bool one2;
      Fr t0 = stage_twiddle(p, g0, hl, one);
      Fr t2 = stage_twiddle(p, g2, hl, one2);
      if (DIT) {
        Fr v1 = one ? a[1] : fe_mul(a[1], t0);
        Fr v3 = one2 ? a[3] : fe_mul(a[3], t2);
        Fr x0 = fe_add(a[0], v1), x1 = fe_sub(a[0], v1), x2 = fe_add(a[2], v3), x3 = fe_sub(a[2], v3);
        a[0] = x0; a[1] = x1; a[2] = x2; a[3] = x3;
      } else {
        Fr x0 = fe_add(a[0], a[1]), x1 = fe_sub(a[0], a[1]), x2 = fe_add(a[2], a[3]), x3 = fe_sub(a[2], a[3]);
        a[0] = x0; a[2] = x2;
        a[1] = one ? x1 : fe_mul(x1, t0);
        a[3] = one2 ? x3 : fe_mul(x3, t2);
      }
    }
    // ---- hand the elements to the next round, or write the tile back
    if (rd + 1 < nrounds) {
      if (rd != 0) __syncthreads();  // everyone has read its elements of this round
#pragma unroll
      for (int m = 0; m < 4; m++) smem_put(s_lo, s_hi, (rows[m] << p.cb) | c, a[m]);
    } else {
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const size_t l = tile_to_local(p, tile, rows[m], c);
        Fr v = a[m];
        if (p.scale == SCALE_POST_NINV) v = fe_mul(v, fe_load_ro<FrParams>(p.ninv));
        if (p.scale == SCALE_POST_COSET) v = fe_mul(v, coset_factor(p, local_to_logical(p, l)));
        const size_t addr = p.dst_perm ? permuted(p, l) : l;
        fe_store(store_target(p, dst, addr), v);
      }
    }
  }
}

__global__ void bit_reverse_kernel(uint4* a, unsigned log2n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >> log2n) return;
  size_t j = (size_t)(__brev((unsigned)i) >> (32 - log2n));
  if (i < j) {
    uint4 x0 = a[2 * i], x1 = a[2 * i + 1];
    uint4 y0 = a[2 * j], y1 = a[2 * j + 1];
    a[2 * i] = y0; a[2 * i + 1] = y1;
    a[2 * j] = x0; a[2 * j + 1] = x1;
  }
}

int bit_reverse_run(b200zk_ctx* ctx, void* a_dev, unsigned log2n) {
  if (log2n == 0) return B200ZK_OK;
  size_t N = (size_t)1 << log2n;
  unsigned blocks = (unsigned)((N + 255) / 256);
  bit_reverse_kernel<<<blocks, 256, 0, ctx->stream>>>((uint4*)a_dev, log2n);
  B200ZK_LAUNCH_CHECK(ctx, "bit_reverse_kernel");
  return B200ZK_OK;
}

static constexpr unsigned TILE_LOG = 10;   // 1024 elements = 32 KiB shared memory per CTA
static constexpr unsigned MIN_CB = 2;      // >= 4 contiguous elements (128 B) per row segment

// Description of the slice of the logical transform one call works on.
struct NttSlice {
  unsigned n, nl;                         // logical / local log2 sizes
  unsigned gap_pos, gap_bits, gap_val;    // logical = local with gap_val inserted at gap_pos
  unsigned stage_lo, stage_hi;            // LOCAL half-distance bits [stage_lo, stage_hi) to execute
  unsigned perm_a, perm_b;
  bool src_perm, dst_perm;
  bool first, last;                       // slice contains the first / last executed stage of the whole transform
  void* const* peers = nullptr;           // when set: the last pass scatters into these peer buffers (fused exchange)
  unsigned npeers = 0;
};

// Runs the stages of `sl` on src -> dst (src == dst allowed when no permutation is requested).
static int run_slice(b200zk_ctx* ctx, const void* src, void* dst, const NttSlice& sl, int inverse, bool dit, int coset) {
  const NttDomain& d = ctx->domains[sl.n];
  const unsigned nstages = sl.stage_hi - sl.stage_lo;
  const unsigned tlog = sl.nl < TILE_LOG ? sl.nl : TILE_LOG;
  unsigned npass, ks[8];
  if (nstages == 0) return B200ZK_ERR_BAD_ARG;
  if (nstages <= tlog && (sl.stage_lo == 0 || nstages <= tlog - MIN_CB || tlog < TILE_LOG)) {
    npass = 1;
    ks[0] = nstages;
  } else {
    const unsigned kmax = TILE_LOG - MIN_CB;
    npass = (nstages + kmax - 1) / kmax;
    unsigned base = nstages / npass, rem = nstages % npass;
    for (unsigned i = 0; i < npass; i++) ks[i] = base + (i < rem ? 1 : 0);
  }
  if ((sl.src_perm || sl.dst_perm) && src == dst && !sl.peers) return B200ZK_ERR_BAD_ARG;
  unsigned done = 0;
  for (unsigned pi = 0; pi < npass; pi++) {
    PassParams p;
    p.n = sl.n;
    p.nl = sl.nl;
    p.k = ks[pi];
    // DIF: stages from the top (largest distance first); DIT: from the bottom
    p.L = dit ? (sl.stage_lo + done) : (sl.stage_hi - done - p.k);
    p.cb = tlog - p.k;
    p.gap_pos = sl.gap_pos;
    p.gap_bits = sl.gap_bits;
    p.gap_val = sl.gap_val;
    p.perm_a = sl.perm_a;
    p.perm_b = sl.perm_b;
    p.src_perm = (sl.src_perm && pi == 0) ? 1 : 0;
    p.dst_perm = (sl.dst_perm && pi == npass - 1) ? 1 : 0;
    p.tw = (const uint4*)(inverse ? d.tw_inv : d.tw_fwd);
    p.peer_on = 0;
    p.peer_chunk_log = p.peer_rank = 0;
    for (int i = 0; i < 8; i++) p.peer_dst[i] = nullptr;
    if (sl.peers && pi == npass - 1) {
      p.peer_on = 1;
      p.peer_chunk_log = sl.nl - sl.gap_bits;
      p.peer_rank = sl.gap_val;
      for (unsigned i = 0; i < sl.npeers && i < 8; i++) p.peer_dst[i] = (uint4*)sl.peers[i];
    }
    p.lo = p.hi = nullptr;
    p.ninv = (const uint4*)d.scalars;
    p.scale = SCALE_NONE;
    p.scale_bitrev = 0;
    if (!inverse && coset && sl.first && pi == 0) {
      p.scale = SCALE_PRE_COSET;
      p.lo = (const uint4*)d.coset_lo;
      p.hi = (const uint4*)d.coset_hi;
      p.scale_bitrev = dit ? 1 : 0;  // DIT input is in bit-reversed order
    }
    if (inverse && sl.last && pi == npass - 1) {
      if (coset) {
        p.scale = SCALE_POST_COSET;
        p.lo = (const uint4*)d.coset_inv_lo;
        p.hi = (const uint4*)d.coset_inv_hi;
        p.scale_bitrev = dit ? 0 : 1;  // DIF output is in bit-reversed order
      } else {
        p.scale = SCALE_POST_NINV;
      }
    }
    // buffers: a permuted read happens on the first pass (src -> dst), a permuted write on the last one
    // (src -> dst); every other pass is in place on whichever buffer currently holds the data
    const void* in;
    void* out;
    if (sl.dst_perm) {
      in = src;
      out = (pi == npass - 1) ? dst : const_cast<void*>(src);
    } else {
      in = (pi == 0) ? src : dst;
      out = dst;
    }
    const unsigned T = 1u << tlog;
    const unsigned threads = T / 2 ? T / 2 : 1;
    const unsigned tiles = (unsigned)(((size_t)1 << sl.nl) >> tlog);
    const size_t shmem = (size_t)T * 32;
    PhaseTimer pt(ctx, PH_NTT_PASS);
    if (tlog >= 4 && p.k >= 2 && !ctx->ntt_radix2) {
      // radix-4 kernel: T/4 threads, 4 elements per thread (in-place tiles: every CTA reads its whole tile before
      // it writes, and tiles are disjoint, so src == dst is safe)
      if (dit)
        ntt_pass4_kernel<true><<<tiles, T / 4, shmem, ctx->stream>>>((const uint4*)in, (uint4*)out, p);
      else
        ntt_pass4_kernel<false><<<tiles, T / 4, shmem, ctx->stream>>>((const uint4*)in, (uint4*)out, p);
    } else if (dit) {
      ntt_pass_kernel<true><<<tiles, threads, shmem, ctx->stream>>>((const uint4*)in, (uint4*)out, p);
    } else {
      ntt_pass_kernel<false><<<tiles, threads, shmem, ctx->stream>>>((const uint4*)in, (uint4*)out, p);
    }
    B200ZK_LAUNCH_CHECK(ctx, "ntt_pass_kernel");
    done += p.k;
  }
  return B200ZK_OK;
}

int ntt_run(b200zk_ctx* ctx, void* a_dev, unsigned log2n, int inverse, int decimation, int coset) {
  if (log2n > B200ZK_MAX_LOG2N) return B200ZK_ERR_BAD_ARG;
  if (log2n == 0) return B200ZK_OK;
  B200ZK_TRY(build_domain(ctx, log2n));
  NttSlice sl;
  sl.n = sl.nl = log2n;
  sl.gap_pos = sl.gap_bits = sl.gap_val = 0;
  sl.stage_lo = 0;
  sl.stage_hi = log2n;
  sl.perm_a = sl.perm_b = 0;
  sl.src_perm = sl.dst_perm = false;
  sl.first = sl.last = true;
  return run_slice(ctx, a_dev, a_dev, sl, inverse, decimation == B200ZK_DIT, coset);
}

// Multi-GPU four-step transform, one half per call (the caller performs the all-to-all between the halves).
//   DIF: half 0 = strides >= C on the column-block shard [R][C/g] (in place); exchange;
//        half 1 = strides < C on the row-block shard, reading the all-to-all block layout [g][R/g][C/g] from src
//        and writing plain [R/g][C] to dst.
//   DIT: half 0 = strides < C on the row-block shard, last pass writes the block layout to dst; exchange;
//        half 1 = strides >= C on the column-block shard (in place).
int ntt_dist_run(b200zk_ctx* ctx, const void* src, void* dst, unsigned log2n, unsigned log2g, unsigned rank,
                 unsigned log2c, int half, int inverse, int decimation, int coset, void* const* peers) {
  if (log2n > B200ZK_MAX_LOG2N || log2g == 0 || log2c < log2g + MIN_CB || log2n < log2c + log2g || rank >> log2g)
    return B200ZK_ERR_BAD_ARG;
  B200ZK_TRY(build_domain(ctx, log2n));
  const bool dit = decimation == B200ZK_DIT;
  const unsigned nl = log2n - log2g;
  const unsigned cl = log2c - log2g;  // log2 of local columns in the column-block shard
  NttSlice sl;
  sl.n = log2n;
  sl.nl = nl;
  sl.gap_val = rank;
  sl.gap_bits = log2g;
  sl.perm_a = cl;
  sl.perm_b = log2g;
  sl.src_perm = sl.dst_perm = false;
  const bool high_half = dit ? (half == 1) : (half == 0);  // the half that runs strides >= C
  if (high_half) {
    sl.gap_pos = cl;  // column-block shard: logical = [r][rank][c_lo]
    sl.stage_lo = cl;
    sl.stage_hi = nl;
  } else {
    sl.gap_pos = nl;  // row-block shard: logical = [rank][r_lo][c]
    sl.stage_lo = 0;
    sl.stage_hi = log2c;
    if (dit) sl.dst_perm = true;
    else sl.src_perm = true;
  }
  sl.first = half == 0;
  sl.last = half == 1;
  if (peers) {
    if (half != 0 || log2g > 3) return B200ZK_ERR_BAD_ARG;
    sl.peers = peers;
    sl.npeers = 1u << log2g;
  }
  return run_slice(ctx, src, dst, sl, inverse, dit, coset);
}

}  // namespace b200zk
