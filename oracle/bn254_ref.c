/* CPU oracle in C (TEST INFRASTRUCTURE ONLY — never linked or loaded by the product path).
 *
 * A plain-C restatement of the arithmetic gnark-crypto v0.9.1 performs on the PLONK hot path of
 * /root/reference (call sites: gnark_backend_ffi/backend/plonk/plonk.go:21 plonk.Setup, :67 plonk.Prove):
 *   - fr / fp Montgomery arithmetic on 4 x u64 limbs      (ecc/bn254/fr/element.go, fp/element.go: CIOS "no-carry" mul)
 *   - fft.Domain.FFT / FFTInverse, DIF / DIT, coset        (ecc/bn254/fr/fft/fft.go: difFFT / ditFFT recursion,
 *                                                           parallel while stage < maxSplits, then sequential)
 *   - (*G1Affine).MultiExp                                  (ecc/bn254/multiexp.go: partitionScalars signed digits,
 *                                                           one task per window with extended-Jacobian buckets,
 *                                                           msmReduceChunk running sums, window Horner)
 * PARITY STATUS: parity unpinned against the real reference (its Go sources are not in /root/reference and no Go
 * toolchain exists here); pinned instead by the known-answer vectors in tests/test_oracle_kat.py and by
 * cross-checks against the independent Python big-int oracle (oracle/bn254.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * It is also the "port" CPU baseline: pthread-parallel like gnark's goroutines, but portable C with __int128
 * (slower per core than gnark's amd64 assembly by an unmeasured factor — never to be quoted as gnark's number).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
  uint64_t m[4];
  uint64_t ninv;
  fe one;  /* R mod m */
  fe r2;   /* R^2 mod m */
} field;

static const field FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}}};

static const field FP = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}}};

/* ---------------------------------------------------------------------------------------------- field */
static inline int fe_geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static inline void raw_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)t;
    borrow = (t >> 64) & 1;
  }
}
static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, 32) == 0; }

static inline void fe_add(const field* F, fe* r, const fe* a, const fe* b) {
  u128 c = 0;
  uint64_t t[4];
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  if (fe_geq(t, F->m)) raw_sub(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_sub(const field* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[4];
  u128 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - borrow;
    t[i] = (uint64_t)d;
    borrow = (d >> 64) & 1;
  }
  if (borrow) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)t[i] + F->m[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  memcpy(r->l, t, 32);
}
static inline void fe_neg(const field* F, fe* r, const fe* a) {
  if (fe_is_zero(a)) { *r = *a; return; }
  raw_sub(r->l, F->m, a->l);
}
/* CIOS Montgomery multiplication */
static inline void fe_mul(const field* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t q = t[0] * F->ninv;
    c = (u128)q * F->m[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)q * F->m[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  if (t[4] || fe_geq(t, F->m)) raw_sub(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_sqr(const field* F, fe* r, const fe* a) { fe_mul(F, r, a, a); }
static void fe_from_mont(const field* F, fe* r, const fe* a) {
  fe one = {{1, 0, 0, 0}};
  fe_mul(F, r, a, &one);
}
static void fe_to_mont(const field* F, fe* r, const fe* a) { fe_mul(F, r, a, &F->r2); }
static void fe_pow(const field* F, fe* r, const fe* a, const uint64_t e[4]) {
  fe acc = F->one, base = *a;
  for (int i = 0; i < 256; i++) {
    if ((e[i / 64] >> (i % 64)) & 1) fe_mul(F, &acc, &acc, &base);
    fe_sqr(F, &base, &base);
  }
  *r = acc;
}
static void fe_inv(const field* F, fe* r, const fe* a) {
  uint64_t e[4];
  memcpy(e, F->m, 32);
  e[0] -= 2;
  fe_pow(F, r, a, e);
}
static void fe_set_u64(const field* F, fe* r, uint64_t v) {
  fe t = {{v, 0, 0, 0}};
  fe_to_mont(F, r, &t);
}

/* exported scalar helpers (used by tests to cross-check against the Python oracle) */
void oracle_fr_mul(const void* a, const void* b, void* r) { fe_mul(&FR, (fe*)r, (const fe*)a, (const fe*)b); }
void oracle_fp_mul(const void* a, const void* b, void* r) { fe_mul(&FP, (fe*)r, (const fe*)a, (const fe*)b); }
void oracle_fr_inv(const void* a, void* r) { fe_inv(&FR, (fe*)r, (const fe*)a); }
void oracle_fp_inv(const void* a, void* r) { fe_inv(&FP, (fe*)r, (const fe*)a); }
void oracle_fr_add(const void* a, const void* b, void* r) { fe_add(&FR, (fe*)r, (const fe*)a, (const fe*)b); }
void oracle_fr_sub(const void* a, const void* b, void* r) { fe_sub(&FR, (fe*)r, (const fe*)a, (const fe*)b); }

/* ---------------------------------------------------------------------------------------------- threads */
typedef void (*range_fn)(void* arg, size_t lo, size_t hi, int tid);
typedef struct { range_fn fn; void* arg; size_t lo, hi; int tid; } range_task;
static void* range_tramp(void* p) {
  range_task* t = (range_task*)p;
  t->fn(t->arg, t->lo, t->hi, t->tid);
  return NULL;
}
/* parallel.Execute(n, fn): split [0, n) over nthreads */
static void parallel_for(size_t n, int nthreads, range_fn fn, void* arg) {
  if (nthreads <= 1 || n < 2) { fn(arg, 0, n, 0); return; }
  if ((size_t)nthreads > n) nthreads = (int)n;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  range_task* tk = (range_task*)malloc(sizeof(range_task) * nthreads);
  size_t per = n / nthreads, rem = n % nthreads, lo = 0;
  for (int i = 0; i < nthreads; i++) {
    size_t hi = lo + per + ((size_t)i < rem ? 1 : 0);
    tk[i].fn = fn; tk[i].arg = arg; tk[i].lo = lo; tk[i].hi = hi; tk[i].tid = i;
    pthread_create(&th[i], NULL, range_tramp, &tk[i]);
    lo = hi;
  }
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(th);
  free(tk);
}

/* ---------------------------------------------------------------------------------------------- NTT */
static const uint64_t FR_ROOT28[4] = {0x636e735580d13d9cULL, 0xa22bf3742445ffd6ULL, 0x56452ac01eb203d8ULL,
                                      0x1860ef942963f9e7ULL}; /* Montgomery form of g, order 2^28 */

static uint64_t bitrev(uint64_t i, unsigned log2n) {
  uint64_t r = 0;
  for (unsigned k = 0; k < log2n; k++) r |= ((i >> k) & 1) << (log2n - 1 - k);
  return r;
}

typedef struct {
  fe* a; const fe* tw; size_t n; unsigned log2n; unsigned stage; int dit;
} stage_arg;

/* one full radix-2 stage over the whole array (the part gnark runs with goroutines) */
static void stage_range(void* p, size_t lo, size_t hi, int tid) {
  (void)tid;
  stage_arg* s = (stage_arg*)p;
  /* butterfly index b in [0, n/2): block = b / half, j = b % half */
  size_t half = s->dit ? ((size_t)1 << s->stage) : (s->n >> (s->stage + 1));
  unsigned tshift = s->dit ? (s->log2n - 1 - s->stage) : s->stage;
  for (size_t b = lo; b < hi; b++) {
    size_t blk = b / half, j = b % half;
    fe* x = &s->a[blk * 2 * half + j];
    fe* y = x + half;
    const fe* w = &s->tw[j << tshift];
    if (s->dit) {
      fe t;
      fe_mul(&FR, &t, y, w);
      fe u = *x;
      fe_add(&FR, x, &u, &t);
      fe_sub(&FR, y, &u, &t);
    } else {
      fe u = *x, v = *y;
      fe_add(&FR, x, &u, &v);
      fe_sub(&FR, &v, &u, &v);
      fe_mul(&FR, y, &v, w);
    }
  }
}

/* sequential recursion on a sub-block (gnark: below maxSplits) */
static void dif_rec(fe* a, size_t n, const fe* tw, unsigned stage) {
  if (n == 1) return;
  size_t m = n >> 1;
  for (size_t i = 0; i < m; i++) {
    fe u = a[i], v = a[i + m];
    fe_add(&FR, &a[i], &u, &v);
    fe_sub(&FR, &v, &u, &v);
    fe_mul(&FR, &a[i + m], &v, &tw[i << stage]);
  }
  dif_rec(a, m, tw, stage + 1);
  dif_rec(a + m, m, tw, stage + 1);
}
static void dit_rec(fe* a, size_t n, const fe* tw, unsigned stage) {
  if (n == 1) return;
  size_t m = n >> 1;
  dit_rec(a, m, tw, stage + 1);
  dit_rec(a + m, m, tw, stage + 1);
  for (size_t i = 0; i < m; i++) {
    fe t, u = a[i];
    fe_mul(&FR, &t, &a[i + m], &tw[i << stage]);
    fe_add(&FR, &a[i], &u, &t);
    fe_sub(&FR, &a[i + m], &u, &t);
  }
}
typedef struct { fe* a; size_t sub; const fe* tw; unsigned stage; int dit; } rec_arg;
static void rec_range(void* p, size_t lo, size_t hi, int tid) {
  (void)tid;
  rec_arg* r = (rec_arg*)p;
  for (size_t k = lo; k < hi; k++) {
    if (r->dit) dit_rec(r->a + k * r->sub, r->sub, r->tw, r->stage);
    else dif_rec(r->a + k * r->sub, r->sub, r->tw, r->stage);
  }
}

typedef struct { fe* a; const fe* tab; const fe* extra; unsigned log2n; int rev; } scale_arg;
static void scale_range(void* p, size_t lo, size_t hi, int tid) {
  (void)tid;
  scale_arg* s = (scale_arg*)p;
  for (size_t i = lo; i < hi; i++) {
    if (s->tab) {
      size_t k = s->rev ? bitrev(i, s->log2n) : i;
      fe_mul(&FR, &s->a[i], &s->a[i], &s->tab[k]);
    }
    if (s->extra) fe_mul(&FR, &s->a[i], &s->a[i], s->extra);
  }
}

typedef struct {
  unsigned log2n;
  fe *tw, *twinv, *coset, *cosetinv;
  fe ninv;
} domain;
static domain* g_dom[29];
static pthread_mutex_t g_dom_lock = PTHREAD_MUTEX_INITIALIZER;

static void fill_powers(fe* t, size_t n, const fe* base) {
  t[0] = FR.one;
  for (size_t i = 1; i < n; i++) fe_mul(&FR, &t[i], &t[i - 1], base);
}
static domain* get_domain(unsigned log2n) {
  pthread_mutex_lock(&g_dom_lock);
  domain* d = g_dom[log2n];
  if (!d) {
    d = (domain*)calloc(1, sizeof(domain));
    d->log2n = log2n;
    size_t n = (size_t)1 << log2n, nt = n / 2 ? n / 2 : 1;
    fe w, wi, g, gi;
    memcpy(w.l, FR_ROOT28, 32);
    for (unsigned k = log2n; k < 28; k++) fe_sqr(&FR, &w, &w);
    fe_inv(&FR, &wi, &w);
    fe_set_u64(&FR, &g, 5);
    fe_inv(&FR, &gi, &g);
    fe nn;
    fe_set_u64(&FR, &nn, (uint64_t)n);
    fe_inv(&FR, &d->ninv, &nn);
    d->tw = (fe*)malloc(nt * 32);
    d->twinv = (fe*)malloc(nt * 32);
    d->coset = (fe*)malloc(n * 32);
    d->cosetinv = (fe*)malloc(n * 32);
    fill_powers(d->tw, nt, &w);
    fill_powers(d->twinv, nt, &wi);
    fill_powers(d->coset, n, &g);
    fill_powers(d->cosetinv, n, &gi);
    g_dom[log2n] = d;
  }
  pthread_mutex_unlock(&g_dom_lock);
  return d;
}

void oracle_domain_release(unsigned log2n) {
  pthread_mutex_lock(&g_dom_lock);
  domain* d = g_dom[log2n];
  if (d) {
    free(d->tw); free(d->twinv); free(d->coset); free(d->cosetinv); free(d);
    g_dom[log2n] = NULL;
  }
  pthread_mutex_unlock(&g_dom_lock);
}

/* domain.FFT (inverse = 0) / domain.FFTInverse (inverse = 1); decimation 0 = DIF, 1 = DIT */
int oracle_ntt(void* data, unsigned log2n, int inverse, int decimation, int coset, int nthreads) {
  if (log2n > 28) return -1;
  fe* a = (fe*)data;
  size_t n = (size_t)1 << log2n;
  domain* d = get_domain(log2n);
  if (nthreads < 1) nthreads = 1;
  if (!inverse && coset) {
    scale_arg s = {a, d->coset, NULL, log2n, decimation == 1};
    parallel_for(n, nthreads, scale_range, &s);
  }
  const fe* tw = inverse ? d->twinv : d->tw;
  /* maxSplits: stages run in parallel across the whole array until there are >= nthreads sub-blocks */
  unsigned splits = 0;
  while (((size_t)1 << splits) < (size_t)nthreads && splits < log2n) splits++;
  if (decimation == 0) {
    for (unsigned st = 0; st < splits; st++) {
      stage_arg s = {a, tw, n, log2n, st, 0};
      parallel_for(n / 2, nthreads, stage_range, &s);
    }
    rec_arg r = {a, n >> splits, tw, splits, 0};
    parallel_for((size_t)1 << splits, nthreads, rec_range, &r);
  } else {
    rec_arg r = {a, n >> splits, tw, splits, 1};
    parallel_for((size_t)1 << splits, nthreads, rec_range, &r);
    for (unsigned k = 0; k < splits; k++) {
      unsigned st = log2n - splits + k; /* half size 2^st */
      stage_arg s = {a, tw, n, log2n, st, 1};
      parallel_for(n / 2, nthreads, stage_range, &s);
    }
  }
  if (inverse) {
    scale_arg s = {a, coset ? d->cosetinv : NULL, &d->ninv, log2n, decimation == 0};
    parallel_for(n, nthreads, scale_range, &s);
  }
  return 0;
}

void oracle_bit_reverse(void* data, unsigned log2n) {
  fe* a = (fe*)data;
  size_t n = (size_t)1 << log2n;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log2n);
    if (i < j) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
  }
}

/* ---------------------------------------------------------------------------------------------- G1 */
typedef struct { fe x, y; } g1a;
typedef struct { fe x, y, zz, zzz; } g1x;

static int g1a_is_inf(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static void g1x_set_inf(g1x* p) { p->x = FP.one; p->y = FP.one; memset(&p->zz, 0, 32); memset(&p->zzz, 0, 32); }

static void g1x_double(g1x* p) {
  if (fe_is_zero(&p->zz)) return;
  fe U, V, W, S, M, X3, Y3, t;
  fe_add(&FP, &U, &p->y, &p->y);
  fe_sqr(&FP, &V, &U);
  fe_mul(&FP, &W, &U, &V);
  fe_mul(&FP, &S, &p->x, &V);
  fe_sqr(&FP, &t, &p->x);
  fe_add(&FP, &M, &t, &t);
  fe_add(&FP, &M, &M, &t);
  fe_sqr(&FP, &X3, &M);
  fe_sub(&FP, &X3, &X3, &S);
  fe_sub(&FP, &X3, &X3, &S);
  fe_sub(&FP, &t, &S, &X3);
  fe_mul(&FP, &Y3, &M, &t);
  fe_mul(&FP, &t, &W, &p->y);
  fe_sub(&FP, &Y3, &Y3, &t);
  p->x = X3; p->y = Y3;
  fe_mul(&FP, &p->zz, &V, &p->zz);
  fe_mul(&FP, &p->zzz, &W, &p->zzz);
}
static void g1x_double_mixed(g1x* p, const g1a* a) {
  p->x = a->x; p->y = a->y; p->zz = FP.one; p->zzz = FP.one;
  g1x_double(p);
}
/* g1JacExtended.addMixed; neg != 0 adds -a (subMixed) */
static void g1x_add_mixed(g1x* p, const g1a* a0, int neg) {
  if (g1a_is_inf(a0)) return;
  g1a a = *a0;
  if (neg) fe_neg(&FP, &a.y, &a.y);
  if (fe_is_zero(&p->zz)) { p->x = a.x; p->y = a.y; p->zz = FP.one; p->zzz = FP.one; return; }
  fe P, R, PP, PPP, Q, X3, Y3, t;
  fe_mul(&FP, &P, &a.x, &p->zz);
  fe_sub(&FP, &P, &P, &p->x);
  fe_mul(&FP, &R, &a.y, &p->zzz);
  fe_sub(&FP, &R, &R, &p->y);
  if (fe_is_zero(&P)) {
    if (fe_is_zero(&R)) g1x_double_mixed(p, &a);
    else g1x_set_inf(p);
    return;
  }
  fe_sqr(&FP, &PP, &P);
  fe_mul(&FP, &PPP, &P, &PP);
  fe_mul(&FP, &Q, &p->x, &PP);
  fe_sqr(&FP, &X3, &R);
  fe_sub(&FP, &X3, &X3, &PPP);
  fe_sub(&FP, &X3, &X3, &Q);
  fe_sub(&FP, &X3, &X3, &Q);
  fe_sub(&FP, &t, &Q, &X3);
  fe_mul(&FP, &Y3, &R, &t);
  fe_mul(&FP, &t, &p->y, &PPP);
  fe_sub(&FP, &Y3, &Y3, &t);
  p->x = X3; p->y = Y3;
  fe_mul(&FP, &p->zz, &p->zz, &PP);
  fe_mul(&FP, &p->zzz, &p->zzz, &PPP);
}
static void g1x_add(g1x* p, const g1x* q) {
  if (fe_is_zero(&q->zz)) return;
  if (fe_is_zero(&p->zz)) { *p = *q; return; }
  fe U1, U2, S1, S2, P, R, PP, PPP, Q, X3, Y3, t;
  fe_mul(&FP, &U1, &p->x, &q->zz);
  fe_mul(&FP, &U2, &q->x, &p->zz);
  fe_mul(&FP, &S1, &p->y, &q->zzz);
  fe_mul(&FP, &S2, &q->y, &p->zzz);
  fe_sub(&FP, &P, &U2, &U1);
  fe_sub(&FP, &R, &S2, &S1);
  if (fe_is_zero(&P)) {
    if (fe_is_zero(&R)) g1x_double(p);
    else g1x_set_inf(p);
    return;
  }
  fe_sqr(&FP, &PP, &P);
  fe_mul(&FP, &PPP, &P, &PP);
  fe_mul(&FP, &Q, &U1, &PP);
  fe_sqr(&FP, &X3, &R);
  fe_sub(&FP, &X3, &X3, &PPP);
  fe_sub(&FP, &X3, &X3, &Q);
  fe_sub(&FP, &X3, &X3, &Q);
  fe_sub(&FP, &t, &Q, &X3);
  fe_mul(&FP, &Y3, &R, &t);
  fe_mul(&FP, &t, &S1, &PPP);
  fe_sub(&FP, &Y3, &Y3, &t);
  p->x = X3; p->y = Y3;
  fe_mul(&FP, &t, &p->zz, &q->zz);
  fe_mul(&FP, &p->zz, &t, &PP);
  fe_mul(&FP, &t, &p->zzz, &q->zzz);
  fe_mul(&FP, &p->zzz, &t, &PPP);
}
static void g1x_to_affine(g1a* r, const g1x* p) {
  if (fe_is_zero(&p->zz)) { memset(r, 0, 64); return; }
  fe t, inv;
  fe_mul(&FP, &t, &p->zz, &p->zzz);
  fe_inv(&FP, &inv, &t);
  fe_mul(&FP, &t, &p->x, &p->zzz);
  fe_mul(&FP, &r->x, &t, &inv);
  fe_mul(&FP, &t, &p->y, &p->zz);
  fe_mul(&FP, &r->y, &t, &inv);
}

/* ---------------------------------------------------------------------------------------------- MSM */
typedef struct {
  const g1a* pts; const int32_t* digits; size_t n; unsigned c, W, S; g1x* wsum;
} msm_arg;

/* processChunk: one (point segment, window) task -> sum_b b * bucket[b].  gnark splits the point range when it has
 * more tasks (2*NumCPU) than windows and adds the sub-results (multiexp.go: nbSplits); S segments here. */
static void msm_window_range(void* p, size_t lo, size_t hi, int tid) {
  (void)tid;
  msm_arg* m = (msm_arg*)p;
  size_t B = (size_t)1 << (m->c - 1);
  g1x* buckets = (g1x*)malloc(sizeof(g1x) * B);
  for (size_t t = lo; t < hi; t++) {
    const size_t seg = t / m->W, w = t % m->W;
    const size_t i_lo = m->n * seg / m->S, i_hi = m->n * (seg + 1) / m->S;
    for (size_t b = 0; b < B; b++) g1x_set_inf(&buckets[b]);
    const int32_t* dg = m->digits + w * m->n;
    for (size_t i = i_lo; i < i_hi; i++) {
      int32_t d = dg[i];
      if (d == 0) continue;
      if (d > 0) g1x_add_mixed(&buckets[d - 1], &m->pts[i], 0);
      else g1x_add_mixed(&buckets[-d - 1], &m->pts[i], 1);
    }
    g1x run, tot;
    g1x_set_inf(&run);
    g1x_set_inf(&tot);
    for (size_t b = B; b-- > 0;) {
      g1x_add(&run, &buckets[b]);
      g1x_add(&tot, &run);
    }
    m->wsum[t] = tot;
  }
  free(buckets);
}

static unsigned msm_best_c(size_t n) {
  /* gnark's cost model: min over c of (255/c) * (n + 2^c) among its implemented window sizes */
  static const unsigned cs[] = {4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
  unsigned best = 4;
  double bestcost = 1e300;
  for (unsigned k = 0; k < sizeof(cs) / sizeof(cs[0]); k++) {
    double cost = (255.0 / cs[k]) * ((double)n + (double)((size_t)1 << cs[k]));
    if (cost < bestcost) { bestcost = cost; best = cs[k]; }
  }
  return best;
}

/* (*G1Affine).MultiExp(points, scalars): scalars Montgomery fr, result canonical affine (64 B) */
int oracle_msm(const void* points, const void* scalars, size_t n, void* out, int nthreads, int force_c) {
  const g1a* pts = (const g1a*)points;
  const fe* sc = (const fe*)scalars;
  if (n == 0) { memset(out, 0, 64); return 0; }
  unsigned c = force_c ? (unsigned)force_c : msm_best_c(n);
  unsigned W = (255 + c - 1) / c;
  int32_t* digits = (int32_t*)malloc(sizeof(int32_t) * n * W);
  const int32_t B = (int32_t)1 << (c - 1);
  for (size_t i = 0; i < n; i++) { /* partitionScalars */
    fe s;
    fe_from_mont(&FR, &s, &sc[i]);
    int32_t carry = 0;
    for (unsigned w = 0; w < W; w++) {
      unsigned bit = w * c;
      uint64_t v = 0;
      if (bit < 256) {
        unsigned limb = bit / 64, off = bit % 64;
        v = s.l[limb] >> off;
        if (off + c > 64 && limb + 1 < 4) v |= s.l[limb + 1] << (64 - off);
        v &= ((uint64_t)1 << c) - 1;
      }
      int32_t d = (int32_t)v + carry;
      carry = 0;
      if (d >= B) { d -= 2 * B; carry = 1; }
      digits[(size_t)w * n + i] = d;
    }
  }
  if (nthreads < 1) nthreads = 1;
  unsigned S = (unsigned)nthreads / W;
  if (S < 1) S = 1;
  if ((size_t)S > n) S = (unsigned)n;
  g1x* wsum = (g1x*)malloc(sizeof(g1x) * W * S);
  msm_arg m = {pts, digits, n, c, W, S, wsum};
  parallel_for((size_t)W * S, nthreads, msm_window_range, &m);
  g1x tot;
  g1x_set_inf(&tot);
  for (unsigned seg = 0; seg < S; seg++) {
    g1x part = wsum[seg * W + W - 1];
    for (int w = (int)W - 2; w >= 0; w--) {
      for (unsigned k = 0; k < c; k++) g1x_double(&part);
      g1x_add(&part, &wsum[seg * W + w]);
    }
    g1x_add(&tot, &part);
  }
  g1x_to_affine((g1a*)out, &tot);
  free(wsum);
  free(digits);
  return 0;
}

/* scalar multiplication / addition helpers for building test inputs quickly */
void oracle_g1_add_affine(const void* a, const void* b, void* out) {
  g1x p;
  g1x_set_inf(&p);
  g1x_add_mixed(&p, (const g1a*)a, 0);
  g1x_add_mixed(&p, (const g1a*)b, 0);
  g1x_to_affine((g1a*)out, &p);
}

/* out[i] = (a + i*b) * G for i < n, as affine points: repeated addition of step = b*G, batch normalised */
int oracle_g1_arith_progression(const void* first_affine, const void* step_affine, size_t n, void* out) {
  g1a* o = (g1a*)out;
  if (n == 0) return 0;
  g1x* acc = (g1x*)malloc(sizeof(g1x) * n);
  g1x cur;
  g1x_set_inf(&cur);
  g1x_add_mixed(&cur, (const g1a*)first_affine, 0);
  for (size_t i = 0; i < n; i++) {
    acc[i] = cur;
    g1x_add_mixed(&cur, (const g1a*)step_affine, 0);
  }
  /* batch inversion of zz*zzz */
  fe* pref = (fe*)malloc(32 * n);
  fe run = FP.one;
  for (size_t i = 0; i < n; i++) {
    fe t;
    if (fe_is_zero(&acc[i].zz)) { pref[i] = run; continue; }
    fe_mul(&FP, &t, &acc[i].zz, &acc[i].zzz);
    pref[i] = run;
    fe_mul(&FP, &run, &run, &t);
  }
  fe inv;
  fe_inv(&FP, &inv, &run);
  for (size_t i = n; i-- > 0;) {
    if (fe_is_zero(&acc[i].zz)) { memset(&o[i], 0, 64); continue; }
    fe t, zi;
    fe_mul(&FP, &zi, &inv, &pref[i]);
    fe_mul(&FP, &t, &acc[i].zz, &acc[i].zzz);
    fe_mul(&FP, &inv, &inv, &t);
    fe_mul(&FP, &t, &acc[i].x, &acc[i].zzz);
    fe_mul(&FP, &o[i].x, &t, &zi);
    fe_mul(&FP, &t, &acc[i].y, &acc[i].zz);
    fe_mul(&FP, &o[i].y, &t, &zi);
  }
  free(pref);
  free(acc);
  return 0;
}

/* uniform fr elements (Montgomery-form limbs are taken as the element, like fr.SetRandom): SplitMix64 stream */
void oracle_random_fr(void* out, size_t n, uint64_t seed) {
  fe* o = (fe*)out;
  uint64_t st = seed;
  for (size_t i = 0; i < n;) {
    fe v;
    for (int k = 0; k < 4; k++) {
      st += 0x9E3779B97F4A7C15ULL;
      uint64_t z = st;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
      v.l[k] = z ^ (z >> 31);
    }
    v.l[3] &= 0x3fffffffffffffffULL;
    if (!fe_geq(v.l, FR.m)) o[i++] = v;
  }
}

/* out[i] = scalars[i] * G (G = (1,2)), scalars Montgomery fr; affine outputs; parallel over nthreads.
 * Used to build kzg.NewSRS-style bases for the PLONK oracle ([alpha^i]G). */
typedef struct { const fe* sc; g1a* out; } mulgen_arg;
static void mulgen_range(void* p, size_t lo, size_t hi, int tid) {
  (void)tid;
  mulgen_arg* m = (mulgen_arg*)p;
  g1a G;
  G.x = FP.one;
  fe_add(&FP, &G.y, &FP.one, &FP.one);
  for (size_t i = lo; i < hi; i++) {
    fe s;
    fe_from_mont(&FR, &s, &m->sc[i]);
    g1x acc;
    g1x_set_inf(&acc);
    for (int b = 255; b >= 0; b--) {
      g1x_double(&acc);
      if ((s.l[b / 64] >> (b % 64)) & 1) g1x_add_mixed(&acc, &G, 0);
    }
    g1x_to_affine(&m->out[i], &acc);
  }
}
void oracle_g1_mul_gen_batch(const void* scalars, size_t n, void* out, int nthreads) {
  mulgen_arg m = {(const fe*)scalars, (g1a*)out};
  parallel_for(n, nthreads < 1 ? 1 : nthreads, mulgen_range, &m);
}
