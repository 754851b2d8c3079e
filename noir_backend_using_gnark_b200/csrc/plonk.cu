// Device-resident PLONK prover for BN254: the orchestration of gnark v0.8.0 backend/plonk/bn254 Setup / Prove
// (called at /root/reference/gnark_backend_ffi/backend/plonk/plonk.go:21 and :67) with every polynomial kept in HBM.
// Only the solution vector and 9 blinding scalars go up, and 9 G1 points + 8 field elements come back; between the
// NTT / MSM calls the elementwise stages of SURVEY.md §3.1.1 (P1, P3, P6, P8-P11, P13-P17) run as fused kernels:
//   gather L,R,O | blind | copy-constraint ratio (batch inversion + prefix product) | quotient on the 4n coset
//   (gate + permutation + L1 terms, division by X^n-1 fused) | Horner evaluations | division by (X - a) |
//   linearised polynomial | folds.  The Fiat-Shamir transcript (SHA-256) and the O(1) challenge scalars are host code.
#include <cstdlib>
#include <new>
#include <utility>
#include <vector>
#include "common.cuh"
#include "consts.cuh"
#include "g1.cuh"
#include "host_field.h"
#include "ffi/bn254_host.h"  // host G1 arithmetic (Straus) for the two homomorphic digests

namespace b200zk {

using host::Fe4;
using host::HFR;

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
struct FrArg {
  uint32_t l[8];
};
__device__ __forceinline__ Fr arg(const FrArg& a) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = a.l[i];
  return r;
}
static FrArg to_arg(const Fe4& v) {
  FrArg a;
  memcpy(a.l, v.l, 32);
  return a;
}

// w^j for the domain whose forward table tw holds w^i, i < n/2 (n >= 2)
__device__ __forceinline__ Fr omega_pow(const uint4* __restrict__ tw, size_t j, size_t half) {
  if (j < half) return fe_load_ro<FrParams>(tw + 2 * j);
  return fe_neg(fe_load_ro<FrParams>(tw + 2 * (j - half)));
}

__device__ __forceinline__ Fr coset_u() {
  Fr u;
#pragma unroll
  for (int i = 0; i < 8; i++) u.l[i] = FR_COSET[i];
  return u;
}

// identity-permutation support value for position p in [0, 3n): u^(p / n) * w^(p mod n)
__device__ __forceinline__ Fr ident_value(const uint4* __restrict__ tw, uint64_t p, unsigned log2n) {
  const size_t n = (size_t)1 << log2n;
  Fr v = omega_pow(tw, p & (n - 1), n >> 1);
  unsigned k = (unsigned)(p >> log2n);
  if (k >= 1) v = fe_mul(v, coset_u());
  if (k >= 2) v = fe_mul(v, coset_u());
  return v;
}

__global__ void k_gather_lro(const uint4* __restrict__ sol, const uint32_t* __restrict__ lro, size_t n, uint4* l,
                             uint4* r, uint4* o) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a = lro[i], b = lro[n + i], c = lro[2 * n + i];
  l[2 * i] = sol[2 * (size_t)a]; l[2 * i + 1] = sol[2 * (size_t)a + 1];
  r[2 * i] = sol[2 * (size_t)b]; r[2 * i + 1] = sol[2 * (size_t)b + 1];
  o[2 * i] = sol[2 * (size_t)c]; o[2 * i + 1] = sol[2 * (size_t)c + 1];
}

// spr.Solve's acceptance test, row by row: ql*l + qr*r + qm*l*r + qo*o + qk (+ public input on the placeholder rows)
// must vanish; the smallest failing row lands in *bad_row
__global__ void k_check_gates(const uint4* __restrict__ ql, const uint4* __restrict__ qr, const uint4* __restrict__ qm,
                              const uint4* __restrict__ qo, const uint4* __restrict__ lqk, const uint4* __restrict__ l,
                              const uint4* __restrict__ r, const uint4* __restrict__ o, const uint4* __restrict__ sol,
                              unsigned nb_public, size_t n, uint32_t* bad_row) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr a = fe_load<FrParams>(l + 2 * i), b = fe_load<FrParams>(r + 2 * i), c = fe_load<FrParams>(o + 2 * i);
  Fr v = fe_load<FrParams>(lqk + 2 * i);
  if (i < nb_public) v = fe_add(v, fe_load<FrParams>(sol + 2 * i));
  v = fe_add(v, fe_mul(fe_load<FrParams>(ql + 2 * i), a));
  v = fe_add(v, fe_mul(fe_load<FrParams>(qr + 2 * i), b));
  v = fe_add(v, fe_mul(fe_mul(fe_load<FrParams>(qm + 2 * i), a), b));
  v = fe_add(v, fe_mul(fe_load<FrParams>(qo + 2 * i), c));
  uint32_t nz = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) nz |= v.l[k];
  if (nz) atomicMin(bad_row, (uint32_t)i);
}

// DeserializeFelts on the device: 64 hex characters per element (32 bytes big-endian, regular form) ->
// fr.Element.SetBytes (reduce mod r, Montgomery form).  A non-hex character raises *bad.
__global__ void k_hex_to_fr(const uint4* __restrict__ hex, size_t n, uint4* out, unsigned* bad) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v;
  unsigned invalid = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint4 h = hex[4 * i + q];
    const uint32_t w[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t half = 0;  // 4 characters -> 16 bits, first character most significant
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t c = (w[j] >> (8 * k)) & 0xffu;
        const uint32_t dig = c - '0', let = (c | 0x20u) - 'a';
        invalid |= (dig >= 10u) & (let >= 6u);
        half = (half << 4) | (dig < 10u ? dig : let + 10u);
      }
      const int word = 4 * q + j;  // 0..15, most significant first
      if (word & 1) v.l[7 - word / 2] |= half;
      else v.l[7 - word / 2] = half << 16;
    }
  }
  if (invalid) atomicOr(bad, 1u);
#pragma unroll 1
  for (int k = 0; k < 5; k++) final_sub<FrParams>(v.l);  // 2^256 < 6r: five conditional subtractions reduce any input
  fe_store(out + 2 * i, fe_to_mont(v));
}

// BuildWitnesses on the device: wire k <- values[src[k]]
__global__ void k_gather_fr(const uint4* __restrict__ values, const uint32_t* __restrict__ src, size_t count, uint4* sol) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  const size_t s = src[i];
  sol[2 * i] = values[2 * s];
  sol[2 * i + 1] = values[2 * s + 1];
}

__global__ void k_fill(uint4* dst, size_t count, FrArg v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  dst[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  dst[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// s_k[i] = ident[perm[k*n + i]]  (Lagrange form of the permutation polynomials)
__global__ void k_perm_lagrange(const int64_t* __restrict__ perm, const uint4* __restrict__ tw, unsigned log2n,
                                uint4* s1, uint4* s2, uint4* s3) {
  const size_t n = (size_t)1 << log2n;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(s1 + 2 * i, ident_value(tw, (uint64_t)perm[i], log2n));
  fe_store(s2 + 2 * i, ident_value(tw, (uint64_t)perm[n + i], log2n));
  fe_store(s3 + 2 * i, ident_value(tw, (uint64_t)perm[2 * n + i], log2n));
}

// iop.Blind(order): p[i] -= b_i, p[n+i] += b_i  (p[n..] is zero before)
__global__ void k_blind(uint4* p, size_t n, const uint4* __restrict__ blinding, unsigned count) {
  unsigned i = threadIdx.x;
  if (i >= count) return;
  Fr b = fe_load<FrParams>(blinding + 2 * i);
  fe_store(p + 2 * i, fe_sub(fe_load<FrParams>(p + 2 * i), b));
  fe_store(p + 2 * (n + i), fe_add(fe_load<FrParams>(p + 2 * (n + i)), b));
}

// numerator / denominator of the copy-constraint ratio at row j
// (beta * u^k comes from the host for k = 0, 1, 2: beta * u^k * w^j is then ONE product per term instead of up to three)
struct BetaShift {
  FrArg bu[3];
};
__global__ void k_z_terms(const uint4* __restrict__ l, const uint4* __restrict__ r, const uint4* __restrict__ o,
                          const int64_t* __restrict__ perm, const uint4* __restrict__ tw, unsigned log2n, BetaShift bs,
                          FrArg gamma_a, uint4* num, uint4* den) {
  const size_t n = (size_t)1 << log2n;
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const Fr gamma = arg(gamma_a);
  const Fr wj = omega_pow(tw, j, n >> 1);
  Fr w[3] = {fe_load<FrParams>(l + 2 * j), fe_load<FrParams>(r + 2 * j), fe_load<FrParams>(o + 2 * j)};
  Fr a, b;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const Fr wg = fe_add(w[k], gamma);
    const uint64_t sp = (uint64_t)perm[(size_t)k * n + j];          // sigma: position in [0, 3n)
    const unsigned sk = (unsigned)(sp >> log2n);
    const Fr bsk = arg(sk == 0 ? bs.bu[0] : (sk == 1 ? bs.bu[1] : bs.bu[2]));
    const Fr ta = fe_add(wg, fe_mul(arg(bs.bu[k]), wj));                                  // w + gamma + beta * u^k * w^j
    const Fr tb = fe_add(wg, fe_mul(bsk, omega_pow(tw, sp & (n - 1), n >> 1)));           // w + gamma + beta * sigma
    a = k == 0 ? ta : fe_mul(a, ta);
    b = k == 0 ? tb : fe_mul(b, tb);
  }
  fe_store(num + 2 * j, a);
  fe_store(den + 2 * j, b);
}

// ratio[i] = num[i] / den[i] with one inversion per chunk (Montgomery's trick); result overwrites num, den is scratch
static constexpr int INV_CHUNK = 128;   // one Fermat inversion (267 multiplications) per chunk: 32 rows cost 8.3 of them per row
__global__ void __launch_bounds__(128) k_batch_ratio(uint4* num, uint4* den, uint4* scratch, size_t n) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * INV_CHUNK;
  if (lo >= n) return;
  size_t hi = lo + INV_CHUNK < n ? lo + INV_CHUNK : n;
  Fr acc = fe_one<FrParams>();
  for (size_t i = lo; i < hi; i++) {
    fe_store(scratch + 2 * i, acc);
    acc = fe_mul(acc, fe_load<FrParams>(den + 2 * i));
  }
  Fr inv = fe_inv(acc);
  for (size_t i = hi; i-- > lo;) {
    Fr d = fe_load<FrParams>(den + 2 * i);
    Fr di = fe_mul(inv, fe_load<FrParams>(scratch + 2 * i));
    inv = fe_mul(inv, d);
    fe_store(num + 2 * i, fe_mul(fe_load<FrParams>(num + 2 * i), di));
  }
}

// ---- exclusive prefix product: out[0] = 1, out[j+1] = out[j] * in[j]  (three kernels) -----------------
static constexpr int SCAN_CHUNK = 128;   // 64 left the one-CTA scan between the two chunk kernels with 128 chunks per thread: 0.35 ms
__global__ void __launch_bounds__(128) k_chunk_product(const uint4* __restrict__ in, size_t n, uint4* chunk_prod) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * SCAN_CHUNK;
  if (lo >= n) return;
  size_t hi = lo + SCAN_CHUNK < n ? lo + SCAN_CHUNK : n;
  Fr acc = fe_one<FrParams>();
  for (size_t i = lo; i < hi; i++) acc = fe_mul(acc, fe_load<FrParams>(in + 2 * i));
  fe_store(chunk_prod + 2 * t, acc);
}

// single CTA: chunk_prod[k] <- product of all chunk products before k
__global__ void __launch_bounds__(512) k_scan_chunk_products(uint4* chunk_prod, size_t nchunks) {
  __shared__ uint4 sh[2][512 * 2];
  const unsigned t = threadIdx.x;
  const size_t per = (nchunks + 511) / 512;
  const size_t lo = t * per, hi = lo + per < nchunks ? lo + per : nchunks;
  Fr tot = fe_one<FrParams>();
  for (size_t k = lo; k < hi; k++) tot = fe_mul(tot, fe_load<FrParams>(chunk_prod + 2 * k));
  int cur = 0;
  fe_store(&sh[cur][2 * t], tot);
  __syncthreads();
  for (unsigned d = 1; d < 512; d <<= 1) {  // inclusive Hillis-Steele
    Fr v = fe_load<FrParams>(&sh[cur][2 * t]);
    if (t >= d) v = fe_mul(v, fe_load<FrParams>(&sh[cur][2 * (t - d)]));
    fe_store(&sh[cur ^ 1][2 * t], v);
    cur ^= 1;
    __syncthreads();
  }
  Fr run = t == 0 ? fe_one<FrParams>() : fe_load<FrParams>(&sh[cur][2 * (t - 1)]);
  for (size_t k = lo; k < hi; k++) {
    Fr p = fe_load<FrParams>(chunk_prod + 2 * k);
    fe_store(chunk_prod + 2 * k, run);
    run = fe_mul(run, p);
  }
}

__global__ void __launch_bounds__(128) k_apply_prefix(const uint4* __restrict__ in, const uint4* __restrict__ chunk_prefix,
                                                      size_t n, uint4* out) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * SCAN_CHUNK;
  if (lo >= n) return;
  size_t hi = lo + SCAN_CHUNK < n ? lo + SCAN_CHUNK : n;
  Fr acc = fe_load<FrParams>(chunk_prefix + 2 * t);
  for (size_t i = lo; i < hi; i++) {
    Fr v = fe_load<FrParams>(in + 2 * i);  // read before write: in may alias out
    fe_store(out + 2 * i, acc);
    acc = fe_mul(acc, v);
  }
}

__global__ void k_set_public(uint4* qk, const uint4* __restrict__ sol, unsigned nb_public) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb_public) return;
  qk[2 * i] = sol[2 * i];
  qk[2 * i + 1] = sol[2 * i + 1];
}

// ---- the quotient numerator on the 4n coset, divided by X^n - 1 (bit-reversed layout) -------------------
struct QuotientArgs {
  const uint4 *el, *er, *eo, *ez, *eqk;
  const uint4 *ql, *qr, *qm, *qo, *s1, *s2, *s3, *lone;
  const uint4* tw_big;  // w_{4n}^i, i < N4/2
  uint4* out;
  unsigned log_big, log_ratio;
  // multi-GPU: this launch covers the contiguous range [base, base + count) of the bit-reversed coset evaluations; every
  // array above is indexed by i - base.  z(wX) of a position may live in the range of the rank whose id differs in the
  // lowest bit (8 ranks, 4n domain): ez_other holds that rank's range of ez.
  size_t base, count;
  unsigned log_local;
  const uint4* ez_other;
  FrArg alpha, beta, gamma, beta_u, beta_uu;
  FrArg xn_inv[8];
};

__global__ void __launch_bounds__(256) k_quotient(QuotientArgs q) {
  const size_t N = (size_t)1 << q.log_big;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= q.count) return;
  const size_t gi = q.base + i;  // position in the full bit-reversed vector
  const size_t nat = (size_t)(__brev((unsigned)gi) >> (32 - q.log_big));
  const size_t ratio = (size_t)1 << q.log_ratio;
  const size_t nat_s = (nat + ratio) & (N - 1);
  const size_t gshift = (size_t)(__brev((unsigned)nat_s) >> (32 - q.log_big));
  const size_t lmask = ((size_t)1 << q.log_local) - 1;
  const uint4* ez_s = (gshift >> q.log_local) == (q.base >> q.log_local) ? q.ez : q.ez_other;
  const size_t ishift = gshift & lmask;
  const Fr alpha = arg(q.alpha), beta = arg(q.beta), gamma = arg(q.gamma);
  const Fr L = fe_load_ro<FrParams>(q.el + 2 * i), R = fe_load_ro<FrParams>(q.er + 2 * i),
           O = fe_load_ro<FrParams>(q.eo + 2 * i);
  // gate: ql*l + qr*r + qm*l*r + qo*o + qk
  // (pairs of products share one Montgomery reduction: fe_mul2add)
  Fr ic = fe_mul2add(fe_load_ro<FrParams>(q.ql + 2 * i), L, fe_load_ro<FrParams>(q.qr + 2 * i), R);
  ic = fe_add(ic, fe_mul2add(fe_mul(fe_load_ro<FrParams>(q.qm + 2 * i), L), R, fe_load_ro<FrParams>(q.qo + 2 * i), O));
  ic = fe_add(ic, fe_load_ro<FrParams>(q.eqk + 2 * i));
  // permutation: zs * prod(w + beta*s + gamma) - z * prod(w + beta*u^k*x + gamma),  x = u * w_{4n}^nat
  const Fr x = fe_mul(coset_u(), omega_pow(q.tw_big, nat, N >> 1));
  const Fr Lg = fe_add(L, gamma), Rg = fe_add(R, gamma), Og = fe_add(O, gamma);
  const Fr z = fe_load_ro<FrParams>(q.ez + 2 * i);
  Fr a = fe_add(Lg, fe_mul(beta, x));
  a = fe_mul(a, fe_add(Rg, fe_mul(arg(q.beta_u), x)));
  a = fe_mul(a, fe_add(Og, fe_mul(arg(q.beta_uu), x)));
  Fr b = fe_add(Lg, fe_mul(beta, fe_load_ro<FrParams>(q.s1 + 2 * i)));
  b = fe_mul(b, fe_add(Rg, fe_mul(beta, fe_load_ro<FrParams>(q.s2 + 2 * i))));
  b = fe_mul(b, fe_add(Og, fe_mul(beta, fe_load_ro<FrParams>(q.s3 + 2 * i))));
  const Fr perm = fe_mul2add(b, fe_load_ro<FrParams>(ez_s + 2 * ishift), a, fe_neg(z));   // b * zs - a * z
  // (z - 1) * L1
  const Fr one_term = fe_mul(fe_sub(z, fe_one<FrParams>()), fe_load_ro<FrParams>(q.lone + 2 * i));
  Fr c = fe_add(fe_mul(one_term, alpha), perm);
  c = fe_add(fe_mul(c, alpha), ic);
  c = fe_mul(c, arg(q.xn_inv[nat & (ratio - 1)]));
  fe_store(q.out + 2 * i, c);
}

// ---- polynomial evaluation at a point: partial sums per block, then one CTA ---------------------------------
static constexpr int EVAL_CHUNK = 32;
struct EvalPowers {
  FrArg z;
  FrArg pw[32];  // (z^EVAL_CHUNK)^(2^k)
};

__device__ __forceinline__ Fr block_sum(Fr v, uint4* sh) {
  const unsigned t = threadIdx.x;
  fe_store(sh + 2 * t, v);
  __syncthreads();
  for (unsigned s = blockDim.x / 2; s > 0; s >>= 1) {
    if (t < s) fe_store(sh + 2 * t, fe_add(fe_load<FrParams>(sh + 2 * t), fe_load<FrParams>(sh + 2 * (t + s))));
    __syncthreads();
  }
  return fe_load<FrParams>(sh);
}

__global__ void __launch_bounds__(256) k_eval_partial(const uint4* __restrict__ p, size_t len, EvalPowers pw, uint4* partial) {
  __shared__ uint4 sh[512];
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * EVAL_CHUNK;
  Fr v = fe_zero<FrParams>();
  if (lo < len) {
    size_t hi = lo + EVAL_CHUNK < len ? lo + EVAL_CHUNK : len;
    const Fr z = arg(pw.z);
    for (size_t i = hi; i-- > lo;) v = fe_add(fe_mul(v, z), fe_load_ro<FrParams>(p + 2 * i));
    size_t e = t;  // multiply by z^(t * EVAL_CHUNK)
    for (int k = 0; e != 0; k++, e >>= 1)
      if (e & 1) v = fe_mul(v, arg(pw.pw[k]));
  }
  Fr s = block_sum(v, sh);
  if (threadIdx.x == 0) fe_store(partial + 2 * blockIdx.x, s);
}

__global__ void __launch_bounds__(256) k_eval_final(const uint4* __restrict__ partial, size_t count, uint4* out) {
  __shared__ uint4 sh[512];
  Fr v = fe_zero<FrParams>();
  for (size_t i = threadIdx.x; i < count; i += blockDim.x) v = fe_add(v, fe_load<FrParams>(partial + 2 * i));
  Fr s = block_sum(v, sh);
  if (threadIdx.x == 0) fe_store(out, s);
}

// ---- q = (f - f(a)) / (X - a):  q[i-1] = g_i,  g_i = f_i + a*g_{i+1}  (suffix Horner), three kernels -------------
static constexpr int DIV_CHUNK = 128;   // (64: the one-CTA carry kernel took 0.47 ms per division, four fifths of it)
__global__ void __launch_bounds__(128) k_div_chunk_horner(const uint4* __restrict__ f, size_t len, FrArg a_a, uint4* chunk_h) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * DIV_CHUNK;
  if (lo >= len) return;
  size_t hi = lo + DIV_CHUNK < len ? lo + DIV_CHUNK : len;
  const Fr a = arg(a_a);
  Fr v = fe_zero<FrParams>();
  for (size_t i = hi; i-- > lo;) v = fe_add(fe_mul(v, a), fe_load_ro<FrParams>(f + 2 * i));
  fe_store(chunk_h + 2 * t, v);
}

// single CTA: chunk_h[k] <- G_{k+1} = value of the suffix Horner entering chunk k from above
__global__ void __launch_bounds__(512) k_div_chunk_carry(uint4* chunk_h, size_t nchunks, FrArg a_chunk_a) {
  __shared__ uint4 sh[2][512 * 2];
  const unsigned t = threadIdx.x;
  const size_t per = (nchunks + 511) / 512;
  const size_t lo = t * per, hi = lo + per < nchunks ? lo + per : nchunks;
  const Fr A = arg(a_chunk_a);  // a^DIV_CHUNK
  // local suffix value of this thread's chunks, and M = A^per
  Fr loc = fe_zero<FrParams>();
  for (size_t k = hi; k-- > lo;) loc = fe_add(fe_mul(loc, A), fe_load<FrParams>(chunk_h + 2 * k));
  Fr M = fe_one<FrParams>();   // A^per by square and multiply
  for (int bit = 63 - __clzll((unsigned long long)(per | 1)); bit >= 0; bit--) {
    M = fe_sqr(M);
    if ((per >> bit) & 1) M = fe_mul(M, A);
  }
  int cur = 0;
  fe_store(&sh[cur][2 * t], loc);
  __syncthreads();
  Fr Md = M;
  for (unsigned d = 1; d < 512; d <<= 1) {  // T_t += M^d * T_{t+d}
    Fr v = fe_load<FrParams>(&sh[cur][2 * t]);
    if (t + d < 512) v = fe_add(v, fe_mul(Md, fe_load<FrParams>(&sh[cur][2 * (t + d)])));
    fe_store(&sh[cur ^ 1][2 * t], v);
    cur ^= 1;
    Md = fe_mul(Md, Md);
    __syncthreads();
  }
  Fr carry = t + 1 < 512 ? fe_load<FrParams>(&sh[cur][2 * (t + 1)]) : fe_zero<FrParams>();
  // threads whose range is empty (lo >= nchunks) hold zeros, so the carries are exact
  for (size_t k = hi; k-- > lo;) {
    Fr h = fe_load<FrParams>(chunk_h + 2 * k);
    fe_store(chunk_h + 2 * k, carry);
    carry = fe_add(fe_mul(carry, A), h);
  }
}

__global__ void __launch_bounds__(128) k_div_apply(const uint4* __restrict__ f, size_t len, FrArg a_a,
                                                   const uint4* __restrict__ chunk_carry, uint4* q) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * DIV_CHUNK;
  if (lo >= len) return;
  size_t hi = lo + DIV_CHUNK < len ? lo + DIV_CHUNK : len;
  const Fr a = arg(a_a);
  Fr g = fe_load<FrParams>(chunk_carry + 2 * t);
  for (size_t i = hi; i-- > lo;) {
    g = fe_add(fe_mul(g, a), fe_load_ro<FrParams>(f + 2 * i));  // g_i
    if (i >= 1) fe_store(q + 2 * (i - 1), g);
  }
}

// ---- linearised polynomial (computeLinearizedPolynomial) ----------------------------------------------------
struct LinArgs {
  const uint4 *bz, *s3, *qm, *ql, *qr, *qo, *cqk;
  uint4* out;
  size_t n, len;
  FrArg c_z, c_s3, alpha, rl, l, r, o, lag;
};
__global__ void __launch_bounds__(256) k_linpol(LinArgs a) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= a.len) return;
  const Fr z = fe_load_ro<FrParams>(a.bz + 2 * i);
  Fr v = fe_mul(z, arg(a.c_z));
  if (i < a.n) v = fe_add(v, fe_mul(fe_load_ro<FrParams>(a.s3 + 2 * i), arg(a.c_s3)));
  v = fe_mul(v, arg(a.alpha));
  if (i < a.n) {
    Fr t = fe_mul(fe_load_ro<FrParams>(a.qm + 2 * i), arg(a.rl));
    t = fe_add(t, fe_mul(fe_load_ro<FrParams>(a.ql + 2 * i), arg(a.l)));
    t = fe_add(t, fe_mul(fe_load_ro<FrParams>(a.qr + 2 * i), arg(a.r)));
    t = fe_add(t, fe_mul(fe_load_ro<FrParams>(a.qo + 2 * i), arg(a.o)));
    t = fe_add(t, fe_load_ro<FrParams>(a.cqk + 2 * i));
    v = fe_add(v, t);
  }
  v = fe_add(v, fe_mul(z, arg(a.lag)));
  fe_store(a.out + 2 * i, v);
}

// foldedH[i] = (h3[i]*zpm + h2[i])*zpm + h1[i]
__global__ void k_fold_h(const uint4* __restrict__ h, size_t m, FrArg zpm_a, uint4* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  const Fr zpm = arg(zpm_a);
  Fr v = fe_mul(fe_load_ro<FrParams>(h + 2 * (2 * m + i)), zpm);
  v = fe_add(v, fe_load_ro<FrParams>(h + 2 * (m + i)));
  v = fe_add(fe_mul(v, zpm), fe_load_ro<FrParams>(h + 2 * i));
  fe_store(out + 2 * i, v);
}

// kzg.BatchOpenSinglePoint fold: out[j] = sum_i gamma^i * p_i[j]
struct FoldArgs {
  const uint4* p[7];
  size_t len[7];
  FrArg gpow[7];
  uint4* out;
  size_t out_len;
};
__global__ void __launch_bounds__(256) k_fold7(FoldArgs a) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= a.out_len) return;
  Fr v = j < a.len[0] ? fe_load_ro<FrParams>(a.p[0] + 2 * j) : fe_zero<FrParams>();
#pragma unroll
  for (int i = 1; i < 7; i++)
    if (j < a.len[i]) v = fe_add(v, fe_mul(fe_load_ro<FrParams>(a.p[i] + 2 * j), arg(a.gpow[i])));
  fe_store(a.out + 2 * j, v);
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU prover (b200zk_plonk_join): the ranks' arenas are mapped into each other (CUDA IPC / same process), so a
// buffer of rank k is `base[k] + offset` with the SAME offset on every rank.  No collective library on the data path:
// barriers are epoch counters written into the peers' arenas, data moves by peer stores (NTT exchange, fused into the
// butterfly pass), peer loads (partial commitments) and peer DMA copies (polynomials).
// ---------------------------------------------------------------------------------------------------
struct PeerSet {
  unsigned world, rank;
  char* base[8];
};

// Stream-ordered barrier: thread t publishes `epoch` in slot [rank] of peer t's flag array, then waits until slot [t]
// of the own array has reached it.  Epochs only grow, so a peer that is already one barrier ahead also satisfies the
// wait.  A peer that never arrives (crashed process) raises *err after 20 s instead of hanging the device.
__global__ void k_group_barrier(PeerSet ps, size_t flags_off, size_t err_off, unsigned epoch) {  // flags_off: of the channel
  const unsigned t = threadIdx.x;
  if (t >= ps.world) return;
  __threadfence_system();
  unsigned* remote = reinterpret_cast<unsigned*>(ps.base[t] + flags_off) + ps.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const unsigned* mine = reinterpret_cast<const unsigned*>(ps.base[ps.rank] + flags_off) + t;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int)(v - epoch) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) {
      *reinterpret_cast<unsigned*>(ps.base[ps.rank] + err_off) = 1u;
      break;
    }
    __nanosleep(100);
  }
}

// commitment = canonical affine of the sum over ranks of the extended-Jacobian partials of one slot (block b: slot[b]);
// every rank reads all partials through the peer mappings and forms the same point
struct SumSlots {
  int slot[8];
};
__global__ void k_sum_partials_peers(PeerSet ps, size_t gather_off, SumSlots sl, void* points) {
  if (threadIdx.x != 0) return;
  const int slot = sl.slot[blockIdx.x];
  G1XYZZ total = g1_xyzz_inf();
  for (unsigned k = 0; k < ps.world; k++) {
    G1XYZZ q = g1_load_xyzz(ps.base[k] + gather_off, slot);
    g1_add(total, q);
  }
  g1_store_affine(points, slot, g1_to_affine_single(total));
}

// column-block shard of the zero-padded coefficient vector: S[r][c_lo] = poly[r*C + rank*C_loc + c_lo] (0 beyond len)
__global__ void k_build_colblock(const uint4* __restrict__ poly, size_t len, uint4* __restrict__ shard, unsigned log2c,
                                 unsigned cl, unsigned rank, size_t count) {
  size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (l >= count) return;
  const size_t r = l >> cl, c = l & (((size_t)1 << cl) - 1);
  const size_t g = (r << log2c) | ((size_t)rank << cl) | c;
  uint4 a = make_uint4(0, 0, 0, 0), b = a;
  if (g < len) {
    a = poly[2 * g];
    b = poly[2 * g + 1];
  }
  shard[2 * l] = a;
  shard[2 * l + 1] = b;
}

}  // namespace b200zk

// ---------------------------------------------------------------------------------------------------
// proving key (device resident) and the prover
// ---------------------------------------------------------------------------------------------------
using namespace b200zk;

struct b200zk_plonk_pk {
  unsigned log2n = 0, log_big = 0, nb_public = 0, nb_wires = 0;
  const b200zk_bases* bases = nullptr;
  char* arena = nullptr;  // one allocation; all pointers below point into it
  // static (circuit) data
  uint4 *ql, *qr, *qm, *qo, *cqk, *lqk, *s1, *s2, *s3;        // canonical (lqk: Lagrange), n each
  uint4 *lql, *lqr, *lqm, *lqo;                                // Lagrange selectors, for the per-row constraint check
  uint32_t* bad_row;                                           // first unsatisfied row of the last prove (device)
  long long last_bad_row = -1;
  uint32_t* sol_src = nullptr;   // optional (b200zk_plonk_set_solution_map): wire k <- values[sol_src[k]], own allocation
  size_t map_values = 0;
  uint4 *e_ql, *e_qr, *e_qm, *e_qo, *e_s1, *e_s2, *e_s3, *e_lone;  // Lagrange-coset bit-reversed, N4 each
  int64_t* perm;
  uint32_t* lro;
  // per-proof working set
  uint4 *sol, *blinding, *l, *r, *o, *bl, *br, *bo, *bz, *qk;   // n (+3) each
  uint4 *el, *er, *eo, *ez, *eqk, *t;                            // N4 each
  uint4 *lin, *folded_h, *folded, *quot, *chunks, *partials, *scal;
  void* points;       // 16 x 64 B result slots
  uint8_t vk_points[8 * 64];  // S0,S1,S2,Ql,Qr,Qm,Qo,Qk affine (Montgomery)
  // optional replacement for the local MSM of every commitment (multi-GPU: point-range-sharded MSM + NVLink gather)
  b200zk_commit_fn commit_hook = nullptr;
  void* commit_user = nullptr;
  // multi-GPU SPMD prover (b200zk_plonk_join): peers' arenas as seen from this process, barrier state
  unsigned rank = 0, world = 1, log2g = 0, dist_log2c = 0;
  size_t arena_bytes = 0;
  char* peer_base[8] = {};
  uint32_t* flags = nullptr;     // [world] arrival epochs written by the peers
  uint32_t* dist_err = nullptr;  // raised by a barrier that timed out
  void* gather = nullptr;        // 16 x 128 B: this rank's extended-Jacobian partial per commitment slot
  unsigned epoch[2] = {0, 0};    // barrier channel 0: context stream; 1: MSM lane 1 (forked commitments)
};

namespace {

struct Carver {
  char* base;
  size_t off = 0;
  template <class T>
  T* take(size_t bytes) {
    T* p = base ? (T*)(base + off) : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  }
};

void carve(b200zk_plonk_pk* pk, char* base, size_t* total) {
  Carver c{base};
  const size_t n = (size_t)1 << pk->log2n, N4 = (size_t)1 << pk->log_big;
  const size_t small = (n + 8) * 32, big = N4 * 32;
  uint4** smalls[] = {&pk->ql, &pk->qr, &pk->qm, &pk->qo, &pk->cqk, &pk->lqk, &pk->s1, &pk->s2, &pk->s3,
                      &pk->l, &pk->r, &pk->o, &pk->bl, &pk->br, &pk->bo, &pk->bz, &pk->qk,
                      &pk->lin, &pk->folded_h, &pk->folded, &pk->quot, &pk->lql, &pk->lqr, &pk->lqm, &pk->lqo};
  for (auto s : smalls) *s = c.take<uint4>(small);
  uint4** bigs[] = {&pk->e_ql, &pk->e_qr, &pk->e_qm, &pk->e_qo, &pk->e_s1, &pk->e_s2, &pk->e_s3, &pk->e_lone,
                    &pk->el, &pk->er, &pk->eo, &pk->ez, &pk->eqk, &pk->t};
  for (auto b : bigs) *b = c.take<uint4>(big);
  pk->perm = c.take<int64_t>(3 * n * 8);
  pk->lro = c.take<uint32_t>(3 * n * 4);
  pk->sol = c.take<uint4>(((size_t)pk->nb_wires + 1) * 32);
  pk->blinding = c.take<uint4>(16 * 32);
  pk->chunks = c.take<uint4>((N4 / 32 + 2048) * 32);
  pk->partials = c.take<uint4>((N4 / (32 * 256) + 64) * 32);
  pk->scal = c.take<uint4>(64 * 32);
  pk->bad_row = c.take<uint32_t>(256);
  pk->points = c.take<void>(24 * 64);  // 16 result slots + device copy of the 8 vk points
  pk->flags = c.take<uint32_t>(256);
  pk->dist_err = c.take<uint32_t>(256);
  pk->gather = c.take<void>(16 * 128);
  *total = c.off;
}

inline unsigned nblocks(size_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

// Lagrange (natural) -> canonical (natural): FFTInverse(DIF) + BitReverse
int to_canonical(b200zk_ctx* ctx, uint4* a, unsigned log2n) {
  B200ZK_TRY(ntt_run(ctx, a, log2n, 1, B200ZK_DIF, 0));
  return bit_reverse_run(ctx, a, log2n);
}

// canonical (len coefficients) -> Lagrange-coset on the big domain, bit-reversed layout
int to_coset(b200zk_ctx* ctx, const uint4* canonical, size_t len, uint4* out, unsigned log_big) {
  const size_t N4 = (size_t)1 << log_big;
  B200ZK_CUDA(ctx, cudaMemcpyAsync(out, canonical, len * 32, cudaMemcpyDeviceToDevice, ctx->stream));
  B200ZK_CUDA(ctx, cudaMemsetAsync((char*)out + len * 32, 0, (N4 - len) * 32, ctx->stream));
  return ntt_run(ctx, out, log_big, 0, B200ZK_DIF, 1);
}

// ---- multi-GPU helpers ------------------------------------------------------------------------------
inline bool is_dist(const b200zk_plonk_pk* pk) { return pk->world > 1; }

PeerSet peer_set(const b200zk_plonk_pk* pk) {
  PeerSet ps;
  ps.world = pk->world;
  ps.rank = pk->rank;
  for (int k = 0; k < 8; k++) ps.base[k] = pk->peer_base[k];
  return ps;
}

// rank k's copy of a buffer of this rank's arena (same offset in every arena)
template <class T>
T* peer_ptr(const b200zk_plonk_pk* pk, unsigned k, T* local) {
  return reinterpret_cast<T*>(pk->peer_base[k] + (reinterpret_cast<const char*>(local) - pk->arena));
}

// stream-ordered barrier over the ranks; all ranks issue the barriers of one channel in the same (program) order.
// Channel 0 lives on the context stream, channel 1 on MSM lane 1 (a forked commitment runs beside the context stream).
int group_barrier(b200zk_ctx* ctx, b200zk_plonk_pk* pk, int channel = 0) {
  pk->epoch[channel]++;
  cudaStream_t st = channel ? ctx->ws[1].stream : ctx->stream;
  k_group_barrier<<<1, 32, 0, st>>>(peer_set(pk), (size_t)((char*)pk->flags - pk->arena) + 32 * (size_t)channel,
                                    (size_t)((char*)pk->dist_err - pk->arena), pk->epoch[channel]);
  B200ZK_LAUNCH_CHECK(ctx, "k_group_barrier");
  return B200ZK_OK;
}

// `count` independent commitments, each sharded by point range: this rank's MSM over its range of every polynomial
// (on the MSM lanes), one barrier, then every rank sums all ranks' partials of every slot
int commit_dist(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const uint4* const* polys, const size_t* lens, const int* slots,
                int count) {
  if (count > 8) return B200ZK_ERR_BAD_ARG;
  const bool lanes = count > 1 && !ctx->msm_single_lane;
  if (lanes) {
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int l = 1; l < MSM_LANES && l < count; l++) B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->ws[l].stream, ctx->ev_fork, 0));
  }
  SumSlots sl;
  for (int k = 0; k < 8; k++) sl.slot[k] = 0;
  for (int k = 0; k < count; k++) {
    const size_t lo = lens[k] * pk->rank / pk->world, hi = lens[k] * (pk->rank + 1) / pk->world;
    sl.slot[k] = slots[k];
    B200ZK_TRY(msm_run(ctx, pk->bases, lo, polys[k] + 2 * lo, hi - lo, (char*)pk->gather + 128 * slots[k], 1,
                       lanes ? k % MSM_LANES : 0));
  }
  if (lanes)
    for (int l = 1; l < MSM_LANES && l < count; l++) {
      B200ZK_CUDA(ctx, cudaEventRecord(ctx->ws[l].done, ctx->ws[l].stream));
      B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ws[l].done, 0));
    }
  B200ZK_TRY(group_barrier(ctx, pk));
  k_sum_partials_peers<<<count, 32, 0, ctx->stream>>>(peer_set(pk), (size_t)((char*)pk->gather - pk->arena), sl, pk->points);
  B200ZK_LAUNCH_CHECK(ctx, "k_sum_partials_peers");
  return B200ZK_OK;
}

// buffers of the sharded 4n-domain work: the ranges [0, N4/g) of el, er, eo, ez, eqk, t hold this rank's contiguous
// range of the bit-reversed coset evaluations; the upper halves of el / er are the two exchange buffers of the
// four-step transforms (alternating, so one barrier per transform suffices), the upper half of eo stages the
// column-block shard, the upper half of ez receives the neighbour rank's range of ez
struct DistBufs {
  size_t local;      // N4 / g elements
  uint4 *xb[2], *stage, *ez_other;
};
DistBufs dist_bufs(const b200zk_plonk_pk* pk) {
  const size_t N4 = (size_t)1 << pk->log_big;
  DistBufs d;
  d.local = N4 >> pk->log2g;
  d.xb[0] = pk->el + N4;  // uint4 units: element N4/2
  d.xb[1] = pk->er + N4;
  d.stage = pk->eo + N4;
  d.ez_other = pk->ez + N4;
  return d;
}

// canonical (len coefficients) -> this rank's range of the Lagrange-coset form on the big domain (bit-reversed layout):
// four-step DIF, first half on the column-block shard with the exchange fused into its last pass (peer stores)
int to_coset_dist(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const uint4* canonical, size_t len, uint4* out, int seq) {
  const DistBufs d = dist_bufs(pk);
  const unsigned cl = pk->dist_log2c - pk->log2g;
  k_build_colblock<<<nblocks(d.local, 256), 256, 0, ctx->stream>>>(canonical, len, d.stage, pk->dist_log2c, cl, pk->rank, d.local);
  B200ZK_LAUNCH_CHECK(ctx, "k_build_colblock");
  void* peers[8] = {};
  for (unsigned k = 0; k < pk->world; k++) peers[k] = peer_ptr(pk, k, d.xb[seq & 1]);
  B200ZK_TRY(ntt_dist_run(ctx, d.stage, d.stage, pk->log_big, pk->log2g, pk->rank, pk->dist_log2c, 0, 0, B200ZK_DIF, 1, peers));
  B200ZK_TRY(group_barrier(ctx, pk));
  return ntt_dist_run(ctx, d.xb[seq & 1], out, pk->log_big, pk->log2g, pk->rank, pk->dist_log2c, 1, 0, B200ZK_DIF, 1);
}

int commit(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, const uint4* poly, size_t len, int slot) {
  void* out = (char*)pk->points + 64 * slot;
  if (is_dist(pk)) {
    const int slots[1] = {slot};
    return commit_dist(ctx, const_cast<b200zk_plonk_pk*>(pk), &poly, &len, slots, 1);
  }
  if (pk->commit_hook) {
    int rc = pk->commit_hook(pk->commit_user, poly, len, out);
    return rc == 0 ? B200ZK_OK : (rc < 0 ? rc : B200ZK_ERR_CUDA);
  }
  return msm_run(ctx, pk->bases, 0, poly, len, out, 0);
}

// several independent commitments (one prover round): each on its own MSM lane of the context, so the latency-bound
// phases of one MSM run under the bucket accumulation of another; everything is joined back into the context stream
int commit_many(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, const uint4* const* polys, const size_t* lens, const int* slots,
                int count) {
  if (is_dist(pk)) return commit_dist(ctx, const_cast<b200zk_plonk_pk*>(pk), polys, lens, slots, count);
  if (pk->commit_hook || count == 1 || ctx->msm_single_lane) {
    for (int k = 0; k < count; k++) B200ZK_TRY(commit(ctx, pk, polys[k], lens[k], slots[k]));
    return B200ZK_OK;
  }
  B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  for (int l = 1; l < MSM_LANES && l < count; l++) B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->ws[l].stream, ctx->ev_fork, 0));
  for (int k = 0; k < count; k++)
    B200ZK_TRY(msm_run(ctx, pk->bases, 0, polys[k], lens[k], (char*)pk->points + 64 * slots[k], 0, k % MSM_LANES));
  for (int l = 1; l < MSM_LANES && l < count; l++) {
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ws[l].done, ctx->ws[l].stream));
    B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ws[l].done, 0));
  }
  return B200ZK_OK;
}

// one commitment on MSM lane 1 while the context stream continues with work that neither needs its result nor
// rewrites its input; commit_join() orders the context stream after it
int commit_fork(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, const uint4* poly, size_t len, int slot) {
  if (pk->commit_hook || ctx->msm_single_lane) return commit(ctx, pk, poly, len, slot);
  B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->ws[1].stream, ctx->ev_fork, 0));
  if (is_dist(pk)) {
    // this rank's share on lane 1, the lane's own barrier channel, then the sum of all ranks' partials — all beside the
    // context stream
    b200zk_plonk_pk* mpk = const_cast<b200zk_plonk_pk*>(pk);
    const size_t lo = len * pk->rank / pk->world, hi = len * (pk->rank + 1) / pk->world;
    B200ZK_TRY(msm_run(ctx, pk->bases, lo, poly + 2 * lo, hi - lo, (char*)pk->gather + 128 * slot, 1, 1));
    B200ZK_TRY(group_barrier(ctx, mpk, 1));
    SumSlots sl;
    for (int k = 0; k < 8; k++) sl.slot[k] = slot;
    k_sum_partials_peers<<<1, 32, 0, ctx->ws[1].stream>>>(peer_set(pk), (size_t)((char*)pk->gather - pk->arena), sl, pk->points);
    B200ZK_LAUNCH_CHECK(ctx, "k_sum_partials_peers");
  } else {
    B200ZK_TRY(msm_run(ctx, pk->bases, 0, poly, len, (char*)pk->points + 64 * slot, 0, 1));
  }
  B200ZK_CUDA(ctx, cudaEventRecord(ctx->ws[1].done, ctx->ws[1].stream));
  ctx->lane_pending = true;
  return B200ZK_OK;
}
int commit_join(b200zk_ctx* ctx) {
  if (!ctx->lane_pending) return B200ZK_OK;
  ctx->lane_pending = false;
  B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ws[1].done, 0));
  return B200ZK_OK;
}

int fetch_points(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, int first, int count, uint8_t* out) {
  B200ZK_CUDA(ctx, cudaMemcpyAsync(out, (char*)pk->points + 64 * first, 64 * (size_t)count, cudaMemcpyDeviceToHost,
                                   ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

// enqueue p(z) -> scal[slot]
int eval_poly(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, const uint4* p, size_t len, const Fe4& z, int slot) {
  EvalPowers pw;
  pw.z = to_arg(z);
  Fe4 zc = host::pow_u64(HFR, z, EVAL_CHUNK);
  for (int k = 0; k < 32; k++) {
    pw.pw[k] = to_arg(zc);
    zc = host::mul(HFR, zc, zc);
  }
  const size_t threads = (len + EVAL_CHUNK - 1) / EVAL_CHUNK;
  const unsigned blocks = nblocks(threads, 256);
  k_eval_partial<<<blocks, 256, 0, ctx->stream>>>(p, len, pw, pk->partials);
  B200ZK_LAUNCH_CHECK(ctx, "k_eval_partial");
  k_eval_final<<<1, 256, 0, ctx->stream>>>(pk->partials, blocks, pk->scal + 2 * slot);
  B200ZK_LAUNCH_CHECK(ctx, "k_eval_final");
  return B200ZK_OK;
}

int fetch_scalars(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, int first, int count, Fe4* out) {
  B200ZK_CUDA(ctx, cudaMemcpyAsync(out, pk->scal + 2 * first, 32 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

// q = (f - f(a)) / (X - a), len(f) = len, len(q) = len - 1
int divide_x_minus_a(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, const uint4* f, size_t len, const Fe4& a, uint4* q) {
  const size_t nch = (len + DIV_CHUNK - 1) / DIV_CHUNK;
  k_div_chunk_horner<<<nblocks(nch, 128), 128, 0, ctx->stream>>>(f, len, to_arg(a), pk->chunks);
  B200ZK_LAUNCH_CHECK(ctx, "k_div_chunk_horner");
  k_div_chunk_carry<<<1, 512, 0, ctx->stream>>>(pk->chunks, nch, to_arg(host::pow_u64(HFR, a, DIV_CHUNK)));
  B200ZK_LAUNCH_CHECK(ctx, "k_div_chunk_carry");
  k_div_apply<<<nblocks(nch, 128), 128, 0, ctx->stream>>>(f, len, to_arg(a), pk->chunks, q);
  B200ZK_LAUNCH_CHECK(ctx, "k_div_apply");
  return B200ZK_OK;
}

// B200ZK_PROVE_TRACE=1: device time between the prover's stages (CUDA events on the context stream), printed to stderr
struct ProveTrace {
  bool on;
  cudaStream_t st;
  std::vector<std::pair<const char*, cudaEvent_t>> ev;
  explicit ProveTrace(cudaStream_t s) : on(getenv("B200ZK_PROVE_TRACE") != nullptr), st(s) { mark("start"); }
  void mark(const char* label) {
    if (!on) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    ev.emplace_back(label, e);
  }
  ~ProveTrace() {
    if (!on || ev.empty()) return;
    cudaEventSynchronize(ev.back().second);
    float total = 0;
    cudaEventElapsedTime(&total, ev.front().second, ev.back().second);
    fprintf(stderr, "[b200zk prove trace] total %.3f ms:", total);
    for (size_t i = 1; i < ev.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
      fprintf(stderr, " %s %.3f", ev[i].first, ms);
    }
    fprintf(stderr, "\n");
    for (auto& e : ev) cudaEventDestroy(e.second);
  }
};

struct Transcript {
  // fiatshamir.Transcript with sha256: challenge_i = H(name_i || challenge_{i-1} || bindings_i)
  host::Sha256 h;
  bool have_prev = false;
  uint8_t prev[32];
  void begin(const char* name) {
    h.reset();
    h.update(name, strlen(name));
    if (have_prev) h.update(prev, 32);
  }
  void bind(const void* p, size_t n) { h.update(p, n); }
  void bind_point(const uint8_t affine_mont[64]) {
    uint8_t b[64];
    host::marshal_g1(affine_mont, b);
    h.update(b, 64);
  }
  void bind_fr(const Fe4& v) {
    uint8_t b[32];
    host::marshal(HFR, v, b);
    h.update(b, 32);
  }
  Fe4 finish() {
    h.finish(prev);
    have_prev = true;
    return host::set_bytes(HFR, prev);
  }
};

}  // namespace

extern "C" {

int b200zk_plonk_setup(b200zk_ctx* ctx, const b200zk_bases* bases, unsigned log2n, unsigned log2n_big,
                       unsigned nb_public, unsigned nb_wires, const void* ql_l, const void* qr_l, const void* qm_l,
                       const void* qo_l, const void* qk_l, const int64_t* permutation, const uint32_t* lro,
                       b200zk_plonk_pk** out) {
  if (!ctx || !bases || !out || !ql_l || !qr_l || !qm_l || !qo_l || !qk_l || !permutation || !lro)
    return B200ZK_ERR_BAD_ARG;
  if (log2n < 1 || log2n_big < log2n + 2 || log2n_big > B200ZK_MAX_LOG2N) return B200ZK_ERR_BAD_ARG;
  const size_t n = (size_t)1 << log2n, N4 = (size_t)1 << log2n_big;
  if (bases->n < n + 3 || nb_public > n || nb_wires == 0) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  *out = nullptr;
  b200zk_plonk_pk* pk = new (std::nothrow) b200zk_plonk_pk();
  if (!pk) return B200ZK_ERR_OOM;
  pk->log2n = log2n;
  pk->log_big = log2n_big;
  pk->nb_public = nb_public;
  pk->nb_wires = nb_wires;
  pk->bases = bases;
  size_t total = 0;
  carve(pk, nullptr, &total);
  cudaError_t e = cudaMalloc((void**)&pk->arena, total);
  if (e != cudaSuccess) {
    delete pk;
    return set_cuda_error(ctx, e, "cudaMalloc(plonk arena)");
  }
  carve(pk, pk->arena, &total);
  pk->arena_bytes = total;
  cudaStream_t st = ctx->stream;
  auto fail = [&](int rc) {
    cudaStreamSynchronize(st);
    cudaFree(pk->arena);
    delete pk;
    return rc;
  };
#define PK_TRY(expr)                       \
  do {                                     \
    int rc__ = (expr);                     \
    if (rc__ != B200ZK_OK) return fail(rc__); \
  } while (0)
#define PK_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return fail(set_cuda_error(ctx, e__, #call));     \
  } while (0)

  // barrier state of the multi-GPU prover: zero BEFORE the arena can be shared with any peer
  PK_CUDA(cudaMemsetAsync(pk->flags, 0, 512, st));
  // upload the circuit description
  const void* hsrc[5] = {ql_l, qr_l, qm_l, qo_l, qk_l};
  uint4* hdst[5] = {pk->ql, pk->qr, pk->qm, pk->qo, pk->cqk};
  for (int i = 0; i < 5; i++) PK_CUDA(cudaMemcpyAsync(hdst[i], hsrc[i], n * 32, cudaMemcpyHostToDevice, st));
  PK_CUDA(cudaMemcpyAsync(pk->lqk, qk_l, n * 32, cudaMemcpyHostToDevice, st));
  PK_CUDA(cudaMemcpyAsync(pk->perm, permutation, 3 * n * 8, cudaMemcpyHostToDevice, st));
  PK_CUDA(cudaMemcpyAsync(pk->lro, lro, 3 * n * 4, cudaMemcpyHostToDevice, st));
  // selectors -> canonical (the Lagrange forms of ql, qr, qm, qo stay for the prover's constraint check)
  uint4* ldst[4] = {pk->lql, pk->lqr, pk->lqm, pk->lqo};
  for (int i = 0; i < 4; i++) PK_CUDA(cudaMemcpyAsync(ldst[i], hdst[i], n * 32, cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < 5; i++) PK_TRY(to_canonical(ctx, hdst[i], log2n));
  // permutation polynomials (needs the domain-n twiddles: built by the transforms above)
  const uint4* tw_n = (const uint4*)ctx->domains[log2n].tw_fwd;
  k_perm_lagrange<<<nblocks(n, 128), 128, 0, st>>>(pk->perm, tw_n, log2n, pk->s1, pk->s2, pk->s3);
  ctx->launches++;
  PK_CUDA(cudaGetLastError());
  PK_TRY(to_canonical(ctx, pk->s1, log2n));
  PK_TRY(to_canonical(ctx, pk->s2, log2n));
  PK_TRY(to_canonical(ctx, pk->s3, log2n));
  // verifying-key commitments: S1,S2,S3,Ql,Qr,Qm,Qo,Qk
  const uint4* cpoly[8] = {pk->s1, pk->s2, pk->s3, pk->ql, pk->qr, pk->qm, pk->qo, pk->cqk};
  {
    const size_t clen[8] = {n, n, n, n, n, n, n, n};
    const int cslot[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    PK_TRY(commit_many(ctx, pk, cpoly, clen, cslot, 8));
  }
  PK_TRY(fetch_points(ctx, pk, 0, 8, pk->vk_points));
  PK_CUDA(cudaMemcpyAsync((char*)pk->points + 64 * 16, pk->points, 8 * 64, cudaMemcpyDeviceToDevice, st));
  // Lagrange-coset forms on the big domain (gnark: computeLagrangeCosetPolys at key load)
  const uint4* csrc[7] = {pk->ql, pk->qr, pk->qm, pk->qo, pk->s1, pk->s2, pk->s3};
  uint4* cdst[7] = {pk->e_ql, pk->e_qr, pk->e_qm, pk->e_qo, pk->e_s1, pk->e_s2, pk->e_s3};
  for (int i = 0; i < 7; i++) PK_TRY(to_coset(ctx, csrc[i], n, cdst[i], log2n_big));
  // L_1 = (X^n - 1) / (n (X - 1)) = (1/n) * sum_i X^i
  Fe4 ninv = host::inv(HFR, host::from_u64(HFR, (uint64_t)n));
  k_fill<<<nblocks(n, 256), 256, 0, st>>>(pk->e_lone, n, to_arg(ninv));
  ctx->launches++;
  PK_CUDA(cudaGetLastError());
  PK_CUDA(cudaMemsetAsync((char*)pk->e_lone + n * 32, 0, (N4 - n) * 32, st));
  PK_TRY(ntt_run(ctx, pk->e_lone, log2n_big, 0, B200ZK_DIF, 1));
  PK_CUDA(cudaStreamSynchronize(st));
#undef PK_TRY
#undef PK_CUDA
  *out = pk;
  return B200ZK_OK;
}

// plonk.Setup(spr, srs) from the constraint system itself: lays out the rows [placeholders | constraints | padding],
// the wire columns and gnark's permutation (buildPermutation: every position points to the previous position holding
// the same wire, the first occurrence to the last) on the host, then runs b200zk_plonk_setup.
int b200zk_plonk_setup_r1cs(b200zk_ctx* ctx, const b200zk_bases* bases, unsigned nb_public, unsigned nb_secret,
                            size_t nb_constraints, const void* ql, const void* qr, const void* qm, const void* qo,
                            const void* qk, const uint32_t* wire_a, const uint32_t* wire_b, const uint32_t* wire_c,
                            b200zk_plonk_pk** out) {
  if (!ctx || !bases || !out) return B200ZK_ERR_BAD_ARG;
  if (nb_constraints && (!ql || !qr || !qm || !qo || !qk || !wire_a || !wire_b || !wire_c)) return B200ZK_ERR_BAD_ARG;
  const size_t size_system = nb_constraints + nb_public;
  unsigned log2n = 1;
  while (((size_t)1 << log2n) < size_system) log2n++;
  unsigned log_big = 0;
  while (((size_t)1 << log_big) < (size_system < 6 ? 8 : 4) * size_system) log_big++;
  if (log_big < log2n + 2) log_big = log2n + 2;
  if (log_big > B200ZK_MAX_LOG2N) return B200ZK_ERR_UNSUPPORTED;
  const size_t n = (size_t)1 << log2n;
  const unsigned nb_wires = nb_public + nb_secret ? nb_public + nb_secret : 1;
  std::vector<uint8_t> cols[5];
  const void* src[5] = {ql, qr, qm, qo, qk};
  for (int k = 0; k < 5; k++) {
    cols[k].assign(n * 32, 0);
    if (nb_constraints) memcpy(cols[k].data() + (size_t)nb_public * 32, src[k], nb_constraints * 32);
  }
  const Fe4 minus_one = host::neg(HFR, HFR.one);
  for (unsigned i = 0; i < nb_public; i++) memcpy(cols[0].data() + (size_t)i * 32, minus_one.l, 32);  // -PUB_i + qk_i = 0
  std::vector<uint32_t> lro(3 * n, 0);
  for (unsigned i = 0; i < nb_public; i++) lro[i] = i;
  for (size_t i = 0; i < nb_constraints; i++) {
    if (wire_a[i] >= nb_wires || wire_b[i] >= nb_wires || wire_c[i] >= nb_wires) return B200ZK_ERR_BAD_ARG;
    lro[nb_public + i] = wire_a[i];
    lro[n + nb_public + i] = wire_b[i];
    lro[2 * n + nb_public + i] = wire_c[i];
  }
  std::vector<int64_t> perm(3 * n, -1), cycle(nb_wires, -1);
  for (size_t i = 0; i < 3 * n; i++) {
    if (cycle[lro[i]] != -1) perm[i] = cycle[lro[i]];
    cycle[lro[i]] = (int64_t)i;
  }
  for (size_t i = 0; i < 3 * n; i++)
    if (perm[i] == -1) perm[i] = cycle[lro[i]];
  return b200zk_plonk_setup(ctx, bases, log2n, log_big, nb_public, nb_wires, cols[0].data(), cols[1].data(), cols[2].data(),
                            cols[3].data(), cols[4].data(), perm.data(), lro.data(), out);
}

long long b200zk_plonk_unsatisfied_row(const b200zk_plonk_pk* pk) { return pk ? pk->last_bad_row : -1; }

void b200zk_plonk_pk_free(b200zk_ctx* ctx, b200zk_plonk_pk* pk) {
  if (!pk) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  if (pk->arena) cudaFree(pk->arena);
  if (pk->sol_src) cudaFree(pk->sol_src);
  delete pk;
}

int b200zk_plonk_set_commit_hook(b200zk_plonk_pk* pk, b200zk_commit_fn fn, void* user) {
  if (!pk) return B200ZK_ERR_BAD_ARG;
  pk->commit_hook = fn;
  pk->commit_user = user;
  return B200ZK_OK;
}

int b200zk_plonk_arena(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, void** base_dev, size_t* bytes) {
  if (!ctx || !pk || !base_dev || !bytes) return B200ZK_ERR_BAD_ARG;
  *base_dev = pk->arena;
  *bytes = pk->arena_bytes;
  return B200ZK_OK;
}

int b200zk_plonk_join(b200zk_ctx* ctx, b200zk_plonk_pk* pk, unsigned rank, unsigned world, void* const* arena_ptrs) {
  if (!ctx || !pk || !arena_ptrs) return B200ZK_ERR_BAD_ARG;
  unsigned log2g = 0;
  while ((1u << log2g) < world) log2g++;
  if (world < 2 || world > 8 || (1u << log2g) != world || rank >= world) return B200ZK_ERR_BAD_ARG;
  if (arena_ptrs[rank] != (void*)pk->arena) return B200ZK_ERR_BAD_ARG;
  for (unsigned k = 0; k < world; k++)
    if (!arena_ptrs[k]) return B200ZK_ERR_BAD_ARG;
  // four-step shape of the big domain: C columns with at least 4 per rank, at least one row per rank
  const unsigned logb = pk->log_big;
  unsigned log2c = logb > 8 ? logb - 8 : 0;
  if (log2c < log2g + 2) log2c = log2g + 2;
  if (log2c < (logb + 1) / 2) log2c = (logb + 1) / 2;
  if (logb < log2c + log2g || logb - pk->log2n != 2) return B200ZK_ERR_UNSUPPORTED;  // circuits below ~2^6 rows: one GPU
  if (pk->commit_hook) return B200ZK_ERR_BAD_ARG;
  for (unsigned k = 0; k < 8; k++) pk->peer_base[k] = k < world ? (char*)arena_ptrs[k] : nullptr;
  pk->rank = rank;
  pk->world = world;
  pk->log2g = log2g;
  pk->dist_log2c = log2c;
  return B200ZK_OK;
}

int b200zk_plonk_leave(b200zk_ctx* ctx, b200zk_plonk_pk* pk) {
  if (!ctx || !pk) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  pk->world = 1;
  pk->rank = 0;
  pk->log2g = 0;
  return B200ZK_OK;
}

int b200zk_plonk_vk(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, void* out_8_points) {
  if (!ctx || !pk || !out_8_points) return B200ZK_ERR_BAD_ARG;
  memcpy(out_8_points, pk->vk_points, sizeof(pk->vk_points));
  return B200ZK_OK;
}

// copies one of the key's canonical polynomials back (for pk serialisation): which = 0..8 -> ql,qr,qm,qo,cqk,lqk,s1,s2,s3
int b200zk_plonk_pk_poly(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, int which, void* out_host) {
  if (!ctx || !pk || !out_host || which < 0 || which > 8) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint4* src[9] = {pk->ql, pk->qr, pk->qm, pk->qo, pk->cqk, pk->lqk, pk->s1, pk->s2, pk->s3};
  B200ZK_CUDA(ctx, cudaMemcpyAsync(out_host, src[which], ((size_t)32) << pk->log2n, cudaMemcpyDeviceToHost, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

}  // extern "C"

// The prover proper.  solution_host != null: the solution vector is copied up first; null: pk->sol has already been
// filled on the device by work enqueued on the context stream (b200zk_plonk_prove_hex).
static int prove_impl(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const void* solution_host, const void* blinding_host,
                      void* proof_out) {
  cudaStream_t st = ctx->stream;
  const unsigned log2n = pk->log2n, logb = pk->log_big;
  const size_t n = (size_t)1 << log2n, N4 = (size_t)1 << logb;
  const unsigned log_ratio = logb - log2n;
  if (log_ratio > 3) return B200ZK_ERR_UNSUPPORTED;
  uint8_t pts[16 * 64];
  Fe4 sc[16];

  // the kernels below read the twiddle tables of both domains directly (identity polynomial, permutation support)
  B200ZK_TRY(ntt_prepare(ctx, log2n));
  B200ZK_TRY(ntt_prepare(ctx, logb));
  const bool dist = is_dist(pk);
  const DistBufs db = dist_bufs(pk);
  if (dist && pk->rank != 0) solution_host = nullptr;  // ranks > 0 take the solution and the blinding from rank 0
  if (solution_host)
    B200ZK_CUDA(ctx, cudaMemcpyAsync(pk->sol, solution_host, (size_t)pk->nb_wires * 32, cudaMemcpyHostToDevice, st));
  if (!dist || pk->rank == 0)
    B200ZK_CUDA(ctx, cudaMemcpyAsync(pk->blinding, blinding_host, 9 * 32, cudaMemcpyHostToDevice, st));
  if (dist) {
    B200ZK_CUDA(ctx, cudaMemsetAsync(pk->dist_err, 0, 4, st));
    B200ZK_TRY(group_barrier(ctx, pk));
    if (pk->rank != 0) {
      B200ZK_CUDA(ctx, cudaMemcpyAsync(pk->sol, peer_ptr(pk, 0, pk->sol), (size_t)pk->nb_wires * 32, cudaMemcpyDeviceToDevice, st));
      B200ZK_CUDA(ctx, cudaMemcpyAsync(pk->blinding, peer_ptr(pk, 0, pk->blinding), 9 * 32, cudaMemcpyDeviceToDevice, st));
    }
  }
  std::vector<Fe4> pub(pk->nb_public);  // public inputs, bound into the transcript
  if (pk->nb_public) {
    if (solution_host) memcpy(pub.data(), solution_host, (size_t)pk->nb_public * 32);
    else B200ZK_CUDA(ctx, cudaMemcpyAsync(pub.data(), pk->sol, (size_t)pk->nb_public * 32, cudaMemcpyDeviceToHost, st));
  }

  ProveTrace tr(st);
  // P1-P4: L,R,O in Lagrange form, canonical, blinded, committed
  k_gather_lro<<<nblocks(n, 256), 256, 0, st>>>(pk->sol, pk->lro, n, pk->l, pk->r, pk->o);
  B200ZK_LAUNCH_CHECK(ctx, "k_gather_lro");
  // what spr.Solve would reject (plonk.Prove returns its error before committing to anything)
  B200ZK_CUDA(ctx, cudaMemsetAsync(pk->bad_row, 0xff, 4, st));
  k_check_gates<<<nblocks(n, 256), 256, 0, st>>>(pk->lql, pk->lqr, pk->lqm, pk->lqo, pk->lqk, pk->l, pk->r, pk->o, pk->sol,
                                                 pk->nb_public, n, pk->bad_row);
  B200ZK_LAUNCH_CHECK(ctx, "k_check_gates");
  uint32_t bad_row = 0xffffffffu;
  B200ZK_CUDA(ctx, cudaMemcpyAsync(&bad_row, pk->bad_row, 4, cudaMemcpyDeviceToHost, st));
  uint4* lag[3] = {pk->l, pk->r, pk->o};
  uint4* can[3] = {pk->bl, pk->br, pk->bo};
  for (int k = 0; k < 3; k++) {
    if (dist && (unsigned)k % pk->world != pk->rank) continue;  // multi-GPU: rank k mod g makes polynomial k, the others fetch it
    B200ZK_CUDA(ctx, cudaMemcpyAsync(can[k], lag[k], n * 32, cudaMemcpyDeviceToDevice, st));
    B200ZK_CUDA(ctx, cudaMemsetAsync((char*)can[k] + n * 32, 0, 8 * 32, st));
    B200ZK_TRY(to_canonical(ctx, can[k], log2n));
    k_blind<<<1, 32, 0, st>>>(can[k], n, pk->blinding + 2 * (2 * k), 2);
    B200ZK_LAUNCH_CHECK(ctx, "k_blind");
  }
  if (dist) {
    B200ZK_TRY(group_barrier(ctx, pk));
    for (int k = 0; k < 3; k++) {
      const unsigned owner = (unsigned)k % pk->world;
      if (owner != pk->rank)
        B200ZK_CUDA(ctx, cudaMemcpyAsync(can[k], peer_ptr(pk, owner, can[k]), (n + 8) * 32, cudaMemcpyDeviceToDevice, st));
    }
  }
  tr.mark("lro_canonical");
  {
    const size_t clen[3] = {n + 2, n + 2, n + 2};
    const int cslot[3] = {8, 9, 10};
    B200ZK_TRY(commit_many(ctx, pk, can, clen, cslot, 3));
  }
  tr.mark("lro_commit");
  B200ZK_TRY(fetch_points(ctx, pk, 8, 3, pts));  // pts[0..2] = LRO (the sync also lands bad_row)
  pk->last_bad_row = bad_row == 0xffffffffu ? -1 : (long long)bad_row;
  if (pk->last_bad_row >= 0) return B200ZK_ERR_UNSATISFIED;

  // P5: gamma, beta
  Transcript fs;
  fs.begin("gamma");
  for (int i = 0; i < 8; i++) fs.bind_point(pk->vk_points + 64 * i);
  for (auto& v : pub) fs.bind_fr(v);  // (landed with the synchronisation of the L,R,O commitments above)
  for (int i = 0; i < 3; i++) fs.bind_point(pts + 64 * i);
  const Fe4 gamma = fs.finish();
  fs.begin("beta");
  const Fe4 beta = fs.finish();

  // P6-P7: Z
  const uint4* tw_n = (const uint4*)ctx->domains[log2n].tw_fwd;
  uint4* num = pk->lin;       // scratch: lin / folded / quot are free until P15
  uint4* den = pk->folded;
  BetaShift bshift;
  {
    const Fe4 u5 = host::from_u64(HFR, 5);
    const Fe4 b1 = host::mul(HFR, beta, u5);
    bshift.bu[0] = to_arg(beta);
    bshift.bu[1] = to_arg(b1);
    bshift.bu[2] = to_arg(host::mul(HFR, b1, u5));
  }
  k_z_terms<<<nblocks(n, 128), 128, 0, st>>>(pk->l, pk->r, pk->o, pk->perm, tw_n, log2n, bshift, to_arg(gamma), num, den);
  B200ZK_LAUNCH_CHECK(ctx, "k_z_terms");
  k_batch_ratio<<<nblocks((n + INV_CHUNK - 1) / INV_CHUNK, 128), 128, 0, st>>>(num, den, pk->quot, n);
  B200ZK_LAUNCH_CHECK(ctx, "k_batch_ratio");
  {
    const size_t nch = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    k_chunk_product<<<nblocks(nch, 128), 128, 0, st>>>(num, n, pk->chunks);
    B200ZK_LAUNCH_CHECK(ctx, "k_chunk_product");
    k_scan_chunk_products<<<1, 512, 0, st>>>(pk->chunks, nch);
    B200ZK_LAUNCH_CHECK(ctx, "k_scan_chunk_products");
    k_apply_prefix<<<nblocks(nch, 128), 128, 0, st>>>(num, pk->chunks, n, pk->bz);
    B200ZK_LAUNCH_CHECK(ctx, "k_apply_prefix");
  }
  B200ZK_CUDA(ctx, cudaMemsetAsync((char*)pk->bz + n * 32, 0, 8 * 32, st));
  B200ZK_TRY(to_canonical(ctx, pk->bz, log2n));
  k_blind<<<1, 32, 0, st>>>(pk->bz, n, pk->blinding + 2 * 6, 3);
  B200ZK_LAUNCH_CHECK(ctx, "k_blind");
  // the commitment to Z runs on an MSM lane of its own while the context stream prepares what the quotient needs and
  // that does not depend on alpha (P8, P9)
  tr.mark("z_build");
  B200ZK_TRY(commit_fork(ctx, pk, pk->bz, n + 3, 11));

  // P8: qk completed with the public inputs, canonical
  B200ZK_CUDA(ctx, cudaMemcpyAsync(pk->qk, pk->lqk, n * 32, cudaMemcpyDeviceToDevice, st));
  if (pk->nb_public) {
    k_set_public<<<nblocks(pk->nb_public, 128), 128, 0, st>>>(pk->qk, pk->sol, pk->nb_public);
    B200ZK_LAUNCH_CHECK(ctx, "k_set_public");
  }
  B200ZK_TRY(to_canonical(ctx, pk->qk, log2n));
  tr.mark("qk");

  // P9: Lagrange-coset forms on the big domain
  if (dist) {
    B200ZK_TRY(to_coset_dist(ctx, pk, pk->bl, n + 2, pk->el, 0));
    B200ZK_TRY(to_coset_dist(ctx, pk, pk->br, n + 2, pk->er, 1));
    B200ZK_TRY(to_coset_dist(ctx, pk, pk->bo, n + 2, pk->eo, 2));
    B200ZK_TRY(to_coset_dist(ctx, pk, pk->bz, n + 3, pk->ez, 3));
    B200ZK_TRY(to_coset_dist(ctx, pk, pk->qk, n, pk->eqk, 4));
    // z(wX) at a position of this rank's range sits in the neighbour's range when there are more ranks than points of
    // the big domain per point of the small one; the neighbour finished ez before it entered the barrier of the
    // transform after it
    if (pk->log2g > log_ratio)
      B200ZK_CUDA(ctx, cudaMemcpyAsync(db.ez_other, peer_ptr(pk, pk->rank ^ 1u, pk->ez), db.local * 32, cudaMemcpyDeviceToDevice, st));
  } else {
    B200ZK_TRY(to_coset(ctx, pk->bl, n + 2, pk->el, logb));
    B200ZK_TRY(to_coset(ctx, pk->br, n + 2, pk->er, logb));
    B200ZK_TRY(to_coset(ctx, pk->bo, n + 2, pk->eo, logb));
    B200ZK_TRY(to_coset(ctx, pk->bz, n + 3, pk->ez, logb));
    B200ZK_TRY(to_coset(ctx, pk->qk, n, pk->eqk, logb));
  }

  tr.mark("coset_ntts");
  B200ZK_TRY(commit_join(ctx));
  tr.mark("z_commit_join");
  B200ZK_TRY(fetch_points(ctx, pk, 11, 1, pts + 64 * 3));  // pts[3] = Z
  fs.begin("alpha");
  fs.bind_point(pts + 64 * 3);
  const Fe4 alpha = fs.finish();

  // P10-P11: quotient numerator / (X^n - 1), then back to canonical
  {
    QuotientArgs q;
    q.el = pk->el; q.er = pk->er; q.eo = pk->eo; q.ez = pk->ez; q.eqk = pk->eqk;
    q.ql = pk->e_ql; q.qr = pk->e_qr; q.qm = pk->e_qm; q.qo = pk->e_qo;
    q.s1 = pk->e_s1; q.s2 = pk->e_s2; q.s3 = pk->e_s3; q.lone = pk->e_lone;
    q.tw_big = (const uint4*)ctx->domains[logb].tw_fwd;
    q.out = pk->t;
    q.log_big = logb;
    q.log_ratio = log_ratio;
    q.base = 0;
    q.count = N4;
    q.log_local = logb;
    q.ez_other = pk->ez;
    if (dist) {
      // this rank's contiguous range of the bit-reversed coset evaluations; the key's forms are read at the same range
      q.base = db.local * pk->rank;
      q.count = db.local;
      q.log_local = logb - pk->log2g;
      q.ez_other = db.ez_other;
      q.ql += 2 * q.base; q.qr += 2 * q.base; q.qm += 2 * q.base; q.qo += 2 * q.base;
      q.s1 += 2 * q.base; q.s2 += 2 * q.base; q.s3 += 2 * q.base; q.lone += 2 * q.base;
    }
    q.alpha = to_arg(alpha); q.beta = to_arg(beta); q.gamma = to_arg(gamma);
    const Fe4 u = host::from_u64(HFR, 5);
    const Fe4 bu = host::mul(HFR, beta, u);
    q.beta_u = to_arg(bu);
    q.beta_uu = to_arg(host::mul(HFR, bu, u));
    // (u * w_{4n}^i)^n - 1 = u^n * (w_{4n}^n)^i - 1: `ratio` distinct values
    const Fe4 un = host::pow_u64(HFR, u, (uint64_t)n);
    Fe4 g = host::from_u64(HFR, 1);
    {
      // w_{4n}^n is a primitive ratio-th root of unity: take it from the 2^28-th root constant
      Fe4 w;
      uint32_t root[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
      memcpy(w.l, root, 32);
      for (unsigned k = log_ratio; k < 28; k++) w = host::mul(HFR, w, w);  // order 2^log_ratio
      Fe4 wi = HFR.one;
      const size_t ratio = (size_t)1 << log_ratio;
      for (size_t i = 0; i < 8; i++) {
        if (i < ratio) {
          Fe4 v = host::sub(HFR, host::mul(HFR, un, wi), HFR.one);
          q.xn_inv[i] = to_arg(host::inv(HFR, v));
          wi = host::mul(HFR, wi, w);
        } else {
          q.xn_inv[i] = to_arg(g);
        }
      }
    }
    k_quotient<<<nblocks(q.count, 256), 256, 0, st>>>(q);
    B200ZK_LAUNCH_CHECK(ctx, "k_quotient");
  }
  tr.mark("quotient");
  if (dist) {
    // four-step DIT: strides < C on the range this rank holds (exchange fused into its last pass), strides >= C on the
    // column-block shard X[r][c_lo] in the exchange buffer; then every rank collects all column blocks (peer DMA, 2-D
    // copies) into the natural-order h — only the rows that hold the 3(n+2) coefficients
    void* peers[8] = {};
    for (unsigned k = 0; k < pk->world; k++) peers[k] = peer_ptr(pk, k, db.xb[1]);
    B200ZK_TRY(ntt_dist_run(ctx, pk->t, pk->t, logb, pk->log2g, pk->rank, pk->dist_log2c, 0, 1, B200ZK_DIT, 1, peers));
    B200ZK_TRY(group_barrier(ctx, pk));
    B200ZK_TRY(ntt_dist_run(ctx, db.xb[1], db.xb[1], logb, pk->log2g, pk->rank, pk->dist_log2c, 1, 1, B200ZK_DIT, 1));
    B200ZK_TRY(group_barrier(ctx, pk));
    const size_t C = (size_t)1 << pk->dist_log2c, C_loc = C >> pk->log2g;
    size_t rows = (3 * (n + 2) + C - 1) / C;
    if (rows > (N4 >> pk->dist_log2c)) rows = N4 >> pk->dist_log2c;
    for (unsigned k = 0; k < pk->world; k++)
      B200ZK_CUDA(ctx, cudaMemcpy2DAsync(pk->t + 2 * (k * C_loc), C * 32, peer_ptr(pk, k, db.xb[1]), C_loc * 32, C_loc * 32, rows,
                                         cudaMemcpyDeviceToDevice, st));
  } else {
    B200ZK_TRY(ntt_run(ctx, pk->t, logb, 1, B200ZK_DIT, 1));  // h, canonical, natural order
  }

  tr.mark("h_intt");
  // P12: commit h1, h2, h3
  const size_t m = n + 2;
  {
    const uint4* hp[3] = {pk->t, pk->t + 2 * m, pk->t + 4 * m};
    const size_t clen[3] = {m, m, m};
    const int cslot[3] = {12, 13, 14};
    B200ZK_TRY(commit_many(ctx, pk, hp, clen, cslot, 3));
  }
  tr.mark("h_commit");
  B200ZK_TRY(fetch_points(ctx, pk, 12, 3, pts + 64 * 4));  // pts[4..6] = H
  fs.begin("zeta");
  for (int i = 0; i < 3; i++) fs.bind_point(pts + 64 * (4 + i));
  const Fe4 zeta = fs.finish();

  // P13-P14: evaluations at zeta and at w*zeta, opening of Z at w*zeta
  Fe4 omega;
  {
    uint32_t root[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
    memcpy(omega.l, root, 32);
    for (unsigned k = log2n; k < 28; k++) omega = host::mul(HFR, omega, omega);
  }
  const Fe4 zeta_shift = host::mul(HFR, zeta, omega);
  B200ZK_TRY(eval_poly(ctx, pk, pk->bl, n + 2, zeta, 0));
  B200ZK_TRY(eval_poly(ctx, pk, pk->br, n + 2, zeta, 1));
  B200ZK_TRY(eval_poly(ctx, pk, pk->bo, n + 2, zeta, 2));
  B200ZK_TRY(eval_poly(ctx, pk, pk->s1, n, zeta, 3));
  B200ZK_TRY(eval_poly(ctx, pk, pk->s2, n, zeta, 4));
  B200ZK_TRY(eval_poly(ctx, pk, pk->bz, n + 3, zeta_shift, 5));
  tr.mark("evals");
  B200ZK_TRY(fetch_scalars(ctx, pk, 0, 6, sc));
  const Fe4 lz = sc[0], rz = sc[1], oz = sc[2], s1z = sc[3], s2z = sc[4], zu = sc[5];

  // P15/P16 scalars (host)
  const Fe4 u5 = host::from_u64(HFR, 5);
  auto M = [&](const Fe4& a, const Fe4& b) { return host::mul(HFR, a, b); };
  auto A = [&](const Fe4& a, const Fe4& b) { return host::add(HFR, a, b); };
  Fe4 c1 = M(A(A(M(s1z, beta), lz), gamma), A(A(M(s2z, beta), rz), gamma));
  c1 = M(M(c1, zu), beta);
  const Fe4 uz = M(zeta, u5), uuz = M(uz, u5);
  Fe4 c2 = M(A(A(M(beta, zeta), lz), gamma), A(A(M(beta, uz), rz), gamma));
  c2 = M(c2, A(A(M(beta, uuz), oz), gamma));
  c2 = host::neg(HFR, c2);
  Fe4 lagv = host::sub(HFR, host::pow_u64(HFR, zeta, (uint64_t)n), HFR.one);
  lagv = M(lagv, host::inv(HFR, host::sub(HFR, zeta, HFR.one)));
  lagv = M(M(M(lagv, alpha), alpha), host::inv(HFR, host::from_u64(HFR, (uint64_t)n)));
  const Fe4 zpm = host::pow_u64(HFR, zeta, (uint64_t)m);

  // P14: opening of Z at w*zeta
  B200ZK_TRY(divide_x_minus_a(ctx, pk, pk->bz, n + 3, zeta_shift, pk->quot));
  B200ZK_TRY(commit_fork(ctx, pk, pk->quot, n + 2, 15));  // ZShiftedOpening.H, on its own lane until the very end: the
                                                          // batched opening below neither needs it nor rewrites pk->quot

  // P15: linearised polynomial
  {
    LinArgs a;
    a.bz = pk->bz; a.s3 = pk->s3; a.qm = pk->qm; a.ql = pk->ql; a.qr = pk->qr; a.qo = pk->qo; a.cqk = pk->cqk;
    a.out = pk->lin;
    a.n = n;
    a.len = n + 3;
    a.c_z = to_arg(c2); a.c_s3 = to_arg(c1); a.alpha = to_arg(alpha); a.rl = to_arg(M(rz, lz));
    a.l = to_arg(lz); a.r = to_arg(rz); a.o = to_arg(oz); a.lag = to_arg(lagv);
    k_linpol<<<nblocks(n + 3, 256), 256, 0, st>>>(a);
    B200ZK_LAUNCH_CHECK(ctx, "k_linpol");
  }

  // P16: folded H polynomial
  k_fold_h<<<nblocks(m, 256), 256, 0, st>>>(pk->t, m, to_arg(zpm), pk->folded_h);
  B200ZK_LAUNCH_CHECK(ctx, "k_fold_h");

  // P17: batch opening at zeta of [foldedH, lin, L, R, O, S1, S2]
  B200ZK_TRY(eval_poly(ctx, pk, pk->folded_h, m, zeta, 6));
  B200ZK_TRY(eval_poly(ctx, pk, pk->lin, n + 3, zeta, 7));
  // While the device works through the queue above, the host forms the two homomorphic digests from commitments it
  // already holds (what the verifier does; equal to kzg.Commit of the polynomials because the commitment is linear,
  // so no (n+3)-point MSM and no kernel at all):
  //   lin     = l*[Ql] + r*[Qr] + l*r*[Qm] + o*[Qo] + [Qk] + alpha*c1*[S3] + (alpha*c2 + lag)*[Z]      -> pts[9]
  //   foldedH = [H0] + zeta^(n+2)*[H1] + zeta^(2(n+2))*[H2]                                            -> pts[10]
  {
    namespace hf = b200zk::ffi;
    const uint8_t* vkp = pk->vk_points;  // S0,S1,S2,Ql,Qr,Qm,Qo,Qk
    const std::vector<hf::G1> lp = {hf::g1_from_image(vkp + 64 * 3), hf::g1_from_image(vkp + 64 * 4),
                                    hf::g1_from_image(vkp + 64 * 5), hf::g1_from_image(vkp + 64 * 6),
                                    hf::g1_from_image(vkp + 64 * 2), hf::g1_from_image(pts + 64 * 3)};
    const std::vector<Fe4> ls = {lz, rz, M(lz, rz), oz, M(alpha, c1), A(M(alpha, c2), lagv)};
    hf::g1_to_image(hf::g1j_to_affine(hf::g1j_add_affine(hf::g1_msm_small(lp, ls), hf::g1_from_image(vkp + 64 * 7))),
                    pts + 64 * 9);
    const std::vector<hf::G1> hp = {hf::g1_from_image(pts + 64 * 5), hf::g1_from_image(pts + 64 * 6)};
    const std::vector<Fe4> hs = {zpm, M(zpm, zpm)};
    hf::g1_to_image(hf::g1j_to_affine(hf::g1j_add_affine(hf::g1_msm_small(hp, hs), hf::g1_from_image(pts + 64 * 4))),
                    pts + 64 * 10);
  }
  tr.mark("zs_div_lin_fold_evals");
  B200ZK_TRY(fetch_scalars(ctx, pk, 6, 2, sc + 6));
  const Fe4 claimed[7] = {sc[6], sc[7], lz, rz, oz, s1z, s2z};
  Transcript kz;
  kz.begin("gamma");
  kz.bind_fr(zeta);
  kz.bind_point(pts + 64 * 10);
  kz.bind_point(pts + 64 * 9);
  for (int i = 0; i < 3; i++) kz.bind_point(pts + 64 * i);
  kz.bind_point(pk->vk_points);
  kz.bind_point(pk->vk_points + 64);
  const Fe4 gk = kz.finish();
  {
    FoldArgs f;
    const uint4* polys[7] = {pk->folded_h, pk->lin, pk->bl, pk->br, pk->bo, pk->s1, pk->s2};
    const size_t lens[7] = {m, n + 3, n + 2, n + 2, n + 2, n, n};
    Fe4 acc = HFR.one;
    for (int i = 0; i < 7; i++) {
      f.p[i] = polys[i];
      f.len[i] = lens[i];
      f.gpow[i] = to_arg(acc);
      acc = host::mul(HFR, acc, gk);
    }
    f.out = pk->folded;
    f.out_len = n + 3;
    k_fold7<<<nblocks(n + 3, 256), 256, 0, st>>>(f);
    B200ZK_LAUNCH_CHECK(ctx, "k_fold7");
  }
  // the quotient goes to pk->lin (dead once k_fold7 has read it): pk->quot still feeds the MSM on the other lane, and the
  // two opening commitments are in flight together
  B200ZK_TRY(divide_x_minus_a(ctx, pk, pk->folded, n + 3, zeta, pk->lin));
  tr.mark("fold_div");
  B200ZK_TRY(commit(ctx, pk, pk->lin, n + 2, 2));  // slot 2: BatchedProof.H
  tr.mark("batched_commit");
  B200ZK_TRY(commit_join(ctx));  // the lane of ZShiftedOpening.H
  tr.mark("zs_commit_join");
  B200ZK_TRY(fetch_points(ctx, pk, 15, 1, pts + 64 * 8));  // pts[8] = ZShiftedOpening.H
  uint32_t dist_err = 0;
  if (dist) B200ZK_CUDA(ctx, cudaMemcpyAsync(&dist_err, pk->dist_err, 4, cudaMemcpyDeviceToHost, st));
  B200ZK_TRY(fetch_points(ctx, pk, 2, 1, pts + 64 * 7));
  if (dist_err) {
    snprintf(ctx->cuda_err, sizeof(ctx->cuda_err), "multi-GPU prover: a rank did not reach a barrier within 20 s");
    return B200ZK_ERR_CUDA;
  }

  // proof blob: LRO[3], Z, H[3], BatchedProof.H, ZShiftedOpening.H (64 B each) | claimed[7], Z(w*zeta) (32 B each)
  uint8_t* outp = (uint8_t*)proof_out;
  memcpy(outp, pts, 9 * 64);
  memcpy(outp + 9 * 64, claimed, 7 * 32);
  memcpy(outp + 9 * 64 + 7 * 32, &zu, 32);
  return B200ZK_OK;
}

extern "C" {

int b200zk_plonk_prove(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const void* solution_host, const void* blinding_host,
                       void* proof_out) {
  if (!ctx || !pk || !proof_out) return B200ZK_ERR_BAD_ARG;
  const bool follower = pk->world > 1 && pk->rank != 0;  // takes solution and blinding from rank 0 over NVLink
  if (!follower && (!solution_host || !blinding_host)) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  return prove_impl(ctx, pk, solution_host, blinding_host, proof_out);
}

int b200zk_plonk_set_solution_map(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const uint32_t* src_host, size_t nb_values) {
  if (!ctx || !pk || !src_host || nb_values == 0) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  for (unsigned i = 0; i < pk->nb_wires; i++)
    if (src_host[i] >= nb_values) return B200ZK_ERR_BAD_ARG;
  if (!pk->sol_src) B200ZK_CUDA(ctx, cudaMalloc((void**)&pk->sol_src, (size_t)pk->nb_wires * 4));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(pk->sol_src, src_host, (size_t)pk->nb_wires * 4, cudaMemcpyHostToDevice, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  pk->map_values = nb_values;
  return B200ZK_OK;
}

int b200zk_plonk_prove_hex(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const char* values_hex_host, size_t nb_values,
                           const void* blinding_host, void* proof_out) {
  if (!ctx || !pk || !values_hex_host || !blinding_host || !proof_out) return B200ZK_ERR_BAD_ARG;
  if (!pk->sol_src || nb_values != pk->map_values) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // staging: [flag 256 B | hex text 64 B per value | decoded values 32 B each]
  B200ZK_TRY(ensure(ctx, ctx->stage, 256 + nb_values * 96));
  char* base = (char*)ctx->stage.p;
  unsigned* bad = (unsigned*)base;
  uint4* hex = (uint4*)(base + 256);
  uint4* vals = (uint4*)(base + 256 + nb_values * 64);
  B200ZK_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, st));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(hex, values_hex_host, nb_values * 64, cudaMemcpyHostToDevice, st));
  k_hex_to_fr<<<nblocks(nb_values, 128), 128, 0, st>>>(hex, nb_values, vals, bad);
  B200ZK_LAUNCH_CHECK(ctx, "k_hex_to_fr");
  k_gather_fr<<<nblocks(pk->nb_wires, 256), 256, 0, st>>>(vals, pk->sol_src, pk->nb_wires, pk->sol);
  B200ZK_LAUNCH_CHECK(ctx, "k_gather_fr");
  unsigned bad_host = 0;
  B200ZK_CUDA(ctx, cudaMemcpyAsync(&bad_host, bad, 4, cudaMemcpyDeviceToHost, st));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(st));
  if (bad_host) return B200ZK_ERR_BAD_ARG;  // encoding/hex: invalid byte
  return prove_impl(ctx, pk, nullptr, blinding_host, proof_out);
}

}  // extern "C"
