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

/* ============================================================================================================
 * PLONK prover, C restatement (gnark v0.8.0 backend/plonk/bn254 Prove as called at
 * /root/reference/gnark_backend_ffi/backend/plonk/plonk.go:67; same steps as oracle/plonk.py prove(), which is the
 * readable statement of the algorithm — this port exists to give the CPU baseline of the prove metric a compiled,
 * multi-threaded arm and a third implementation for the byte-parity tests).  Parity status: unpinned vs real gnark.
 * ============================================================================================================ */
typedef struct { uint32_t h[8]; uint8_t buf[64]; uint64_t len; size_t fill; } sha256_ctx;
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};
static uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_block(sha256_ctx* s, const uint8_t* p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4*i] << 24) | ((uint32_t)p[4*i+1] << 16) | ((uint32_t)p[4*i+2] << 8) | p[4*i+3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = rotr32(w[i-15], 7) ^ rotr32(w[i-15], 18) ^ (w[i-15] >> 3);
    uint32_t s1 = rotr32(w[i-2], 17) ^ rotr32(w[i-2], 19) ^ (w[i-2] >> 10);
    w[i] = w[i-16] + s0 + w[i-7] + s1;
  }
  uint32_t a = s->h[0], b = s->h[1], c = s->h[2], d = s->h[3], e = s->h[4], f = s->h[5], g = s->h[6], h = s->h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = h + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
    uint32_t t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  s->h[0] += a; s->h[1] += b; s->h[2] += c; s->h[3] += d; s->h[4] += e; s->h[5] += f; s->h[6] += g; s->h[7] += h;
}
static void sha_init(sha256_ctx* s) {
  static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(s->h, iv, sizeof(iv)); s->len = 0; s->fill = 0;
}
static void sha_update(sha256_ctx* s, const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  s->len += n;
  while (n) {
    size_t take = 64 - s->fill < n ? 64 - s->fill : n;
    memcpy(s->buf + s->fill, p, take);
    s->fill += take; p += take; n -= take;
    if (s->fill == 64) { sha_block(s, s->buf); s->fill = 0; }
  }
}
static void sha_final(sha256_ctx* s, uint8_t out[32]) {
  uint64_t bits = s->len * 8;
  uint8_t pad = 0x80, z = 0, lb[8];
  sha_update(s, &pad, 1);
  while (s->fill != 56) sha_update(s, &z, 1);
  for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
  sha_update(s, lb, 8);
  for (int i = 0; i < 8; i++) { out[4*i] = s->h[i] >> 24; out[4*i+1] = s->h[i] >> 16; out[4*i+2] = s->h[i] >> 8; out[4*i+3] = s->h[i]; }
}

static void fe_marshal(const field* F, const fe* a_mont, uint8_t out[32]) {
  fe r; fe_from_mont(F, &r, a_mont);
  for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) out[31 - (8*i + b)] = (uint8_t)(r.l[i] >> (8*b));
}
static void g1_marshal(const g1a* p, uint8_t out[64]) {
  if (g1a_is_inf(p)) { memset(out, 0, 64); out[0] = 0x40; return; }
  fe_marshal(&FP, &p->x, out); fe_marshal(&FP, &p->y, out + 32);
}
static void fr_set_bytes(fe* r, const uint8_t in[32]) {
  fe v;
  for (int i = 0; i < 4; i++) { uint64_t w = 0; for (int b = 0; b < 8; b++) w |= (uint64_t)in[31 - (8*i + b)] << (8*b); v.l[i] = w; }
  while (fe_geq(v.l, FR.m)) raw_sub(v.l, v.l, FR.m);
  fe_to_mont(&FR, r, &v);
}
typedef struct { sha256_ctx s; int have_prev; uint8_t prev[32]; } transcript;
static void tr_begin(transcript* t, const char* name) { sha_init(&t->s); sha_update(&t->s, name, strlen(name)); if (t->have_prev) sha_update(&t->s, t->prev, 32); }
static void tr_point(transcript* t, const g1a* p) { uint8_t b[64]; g1_marshal(p, b); sha_update(&t->s, b, 64); }
static void tr_fr(transcript* t, const fe* v) { uint8_t b[32]; fe_marshal(&FR, v, b); sha_update(&t->s, b, 32); }
static void tr_finish(transcript* t, fe* out) { sha_final(&t->s, t->prev); t->have_prev = 1; fr_set_bytes(out, t->prev); }

#define FMUL(r, a, b) fe_mul(&FR, (r), (a), (b))
#define FADD(r, a, b) fe_add(&FR, (r), (a), (b))
#define FSUB(r, a, b) fe_sub(&FR, (r), (a), (b))

static void fr_pow_u64(fe* r, const fe* a, uint64_t e) {
  fe acc = FR.one, base = *a;
  while (e) { if (e & 1) FMUL(&acc, &acc, &base); FMUL(&base, &base, &base); e >>= 1; }
  *r = acc;
}
static void poly_eval(fe* r, const fe* p, size_t len, const fe* z) {
  fe acc; memset(&acc, 0, 32);
  for (size_t i = len; i-- > 0;) { FMUL(&acc, &acc, z); FADD(&acc, &acc, &p[i]); }
  *r = acc;
}
/* q = (f - f(a)) / (X - a); len(q) = len - 1 */
static void poly_div_x_minus_a(fe* q, const fe* f, size_t len, const fe* a) {
  fe g; memset(&g, 0, 32);
  for (size_t i = len; i-- > 1;) { FMUL(&g, &g, a); FADD(&g, &g, &f[i]); q[i - 1] = g; }
}
static void to_canonical(fe* a, unsigned log2n, int nthreads) {
  oracle_ntt(a, log2n, 1, 0, 0, nthreads);
  oracle_bit_reverse(a, log2n);
}
static void to_coset(fe* out, const fe* canonical, size_t len, unsigned log_big, int nthreads) {
  size_t N = (size_t)1 << log_big;
  memcpy(out, canonical, len * 32);
  memset(out + len, 0, (N - len) * 32);
  oracle_ntt(out, log_big, 0, 0, 1, nthreads);
}

typedef struct {
  unsigned log2n, log_big;
  const fe *el, *er, *eo, *ez, *eqk, *ql, *qr, *qm, *qo, *s1, *s2, *s3, *lone;
  const fe* tw_big;
  fe alpha, beta, gamma, beta_u, beta_uu, u;
  fe xn_inv[8];
  fe* out;
} quot_arg;
static void quot_range(void* p, size_t lo, size_t hi, int tid) {
  (void)tid;
  quot_arg* q = (quot_arg*)p;
  const size_t N = (size_t)1 << q->log_big, ratio = N >> q->log2n;
  for (size_t i = lo; i < hi; i++) {
    size_t nat = bitrev(i, q->log_big);
    size_t ishift = bitrev((nat + ratio) & (N - 1), q->log_big);
    const fe *L = &q->el[i], *R = &q->er[i], *O = &q->eo[i];
    fe ic, t, x, a, b, Lg, Rg, Og, one_t, c;
    FMUL(&ic, &q->ql[i], L);
    FMUL(&t, &q->qr[i], R); FADD(&ic, &ic, &t);
    FMUL(&t, &q->qm[i], L); FMUL(&t, &t, R); FADD(&ic, &ic, &t);
    FMUL(&t, &q->qo[i], O); FADD(&ic, &ic, &t);
    FADD(&ic, &ic, &q->eqk[i]);
    if (nat < N / 2) x = q->tw_big[nat]; else fe_neg(&FR, &x, &q->tw_big[nat - N / 2]);
    FMUL(&x, &x, &q->u);
    FADD(&Lg, L, &q->gamma); FADD(&Rg, R, &q->gamma); FADD(&Og, O, &q->gamma);
    FMUL(&t, &q->beta, &x); FADD(&a, &Lg, &t);
    FMUL(&t, &q->beta_u, &x); FADD(&t, &Rg, &t); FMUL(&a, &a, &t);
    FMUL(&t, &q->beta_uu, &x); FADD(&t, &Og, &t); FMUL(&a, &a, &t);
    FMUL(&a, &a, &q->ez[i]);
    FMUL(&t, &q->beta, &q->s1[i]); FADD(&b, &Lg, &t);
    FMUL(&t, &q->beta, &q->s2[i]); FADD(&t, &Rg, &t); FMUL(&b, &b, &t);
    FMUL(&t, &q->beta, &q->s3[i]); FADD(&t, &Og, &t); FMUL(&b, &b, &t);
    FMUL(&b, &b, &q->ez[ishift]);
    FSUB(&b, &b, &a);
    FSUB(&one_t, &q->ez[i], &FR.one); FMUL(&one_t, &one_t, &q->lone[i]);
    FMUL(&c, &one_t, &q->alpha); FADD(&c, &c, &b);
    FMUL(&c, &c, &q->alpha); FADD(&c, &c, &ic);
    FMUL(&q->out[i], &c, &q->xn_inv[nat & (ratio - 1)]);
  }
}

typedef struct { const fe *l, *r, *o; const int64_t* perm; const fe* ident; fe beta, gamma; size_t n; fe *num, *den; } zterm_arg;
static void zterm_range(void* p, size_t lo, size_t hi, int tid) {
  (void)tid;
  zterm_arg* z = (zterm_arg*)p;
  const fe* w[3] = {z->l, z->r, z->o};
  for (size_t j = lo; j < hi; j++) {
    fe a = FR.one, b = FR.one, t, wg;
    for (int k = 0; k < 3; k++) {
      FADD(&wg, &w[k][j], &z->gamma);
      FMUL(&t, &z->beta, &z->ident[k * z->n + j]); FADD(&t, &wg, &t); FMUL(&a, &a, &t);
      FMUL(&t, &z->beta, &z->ident[z->perm[k * z->n + j]]); FADD(&t, &wg, &t); FMUL(&b, &b, &t);
    }
    z->num[j] = a; z->den[j] = b;
  }
}

/* Inputs are gnark in-memory images (Montgomery).  pk polynomials: canonical ql,qr,qm,qo,cqk,s1,s2,s3 and Lagrange lqk
 * (n each); coset forms of ql,qr,qm,qo,s1,s2,s3 and L_1 on the big domain (bit-reversed layout, N4 each) as gnark
 * keeps them in the loaded key.  proof_out: same 832-byte layout as b200zk_plonk_prove. */
int oracle_plonk_prove(unsigned log2n, unsigned log_big, unsigned nb_public, unsigned nb_wires, const void* polys_n[9],
                       const void* cosets_big[8], const int64_t* perm, const uint32_t* lro, const void* vk_points,
                       const void* srs_g1, const void* solution, const void* blinding, int nthreads, void* proof_out) {
  const size_t n = (size_t)1 << log2n, N4 = (size_t)1 << log_big, m = n + 2;
  (void)nb_wires;
  const fe *ql = polys_n[0], *qr = polys_n[1], *qm = polys_n[2], *qo = polys_n[3], *cqk = polys_n[4], *lqk = polys_n[5],
           *s1 = polys_n[6], *s2 = polys_n[7], *s3 = polys_n[8];
  const fe* sol = (const fe*)solution;
  const fe* bl = (const fe*)blinding;
  const g1a* vk = (const g1a*)vk_points;
  domain* dn = get_domain(log2n);
  domain* db = get_domain(log_big);
  fe *l = malloc(n * 32), *r = malloc(n * 32), *o = malloc(n * 32);
  fe *cl = calloc(n + 8, 32), *cr = calloc(n + 8, 32), *co = calloc(n + 8, 32), *cz = calloc(n + 8, 32), *qk = malloc(n * 32);
  for (size_t i = 0; i < n; i++) { l[i] = sol[lro[i]]; r[i] = sol[lro[n + i]]; o[i] = sol[lro[2 * n + i]]; }
  fe* lag[3] = {l, r, o};
  fe* can[3] = {cl, cr, co};
  g1a pts[11]; /* LRO[3], Z, H[3], batched H, zshift H, lin digest, folded digest */
  for (int k = 0; k < 3; k++) {
    memcpy(can[k], lag[k], n * 32);
    to_canonical(can[k], log2n, nthreads);
    for (int i = 0; i < 2; i++) { FSUB(&can[k][i], &can[k][i], &bl[2 * k + i]); FADD(&can[k][n + i], &can[k][n + i], &bl[2 * k + i]); }
    oracle_msm(srs_g1, can[k], n + 2, &pts[k], nthreads, 0);
  }
  transcript fs; fs.have_prev = 0;
  fe gamma, beta, alpha, zeta;
  tr_begin(&fs, "gamma");
  for (int i = 0; i < 8; i++) tr_point(&fs, &vk[i]);
  for (unsigned i = 0; i < nb_public; i++) tr_fr(&fs, &sol[i]);
  for (int i = 0; i < 3; i++) tr_point(&fs, &pts[i]);
  tr_finish(&fs, &gamma);
  tr_begin(&fs, "beta"); tr_finish(&fs, &beta);
  /* identity support [w^i | u w^i | u^2 w^i] */
  fe* ident = malloc(3 * n * 32);
  fe u; fe_set_u64(&FR, &u, 5);
  fe uu; FMUL(&uu, &u, &u);
  for (size_t i = 0; i < n; i++) {
    fe w;
    if (i < n / 2 || n == 1) w = dn->tw[i < n / 2 ? i : 0]; else fe_neg(&FR, &w, &dn->tw[i - n / 2]);
    ident[i] = w; FMUL(&ident[n + i], &w, &u); FMUL(&ident[2 * n + i], &w, &uu);
  }
  fe *num = malloc(n * 32), *den = malloc(n * 32), *pre = malloc(n * 32);
  zterm_arg za = {l, r, o, perm, ident, beta, gamma, n, num, den};
  parallel_for(n, nthreads, zterm_range, &za);
  { /* batch inversion of den (Montgomery's trick), then z = exclusive prefix product of num/den */
    fe acc = FR.one;
    for (size_t i = 0; i < n; i++) { pre[i] = acc; FMUL(&acc, &acc, &den[i]); }
    fe inv; fe_inv(&FR, &inv, &acc);
    for (size_t i = n; i-- > 0;) { fe di; FMUL(&di, &inv, &pre[i]); FMUL(&inv, &inv, &den[i]); FMUL(&num[i], &num[i], &di); }
    acc = FR.one;
    for (size_t i = 0; i < n; i++) { cz[i] = acc; FMUL(&acc, &acc, &num[i]); }
  }
  to_canonical(cz, log2n, nthreads);
  for (int i = 0; i < 3; i++) { FSUB(&cz[i], &cz[i], &bl[6 + i]); FADD(&cz[n + i], &cz[n + i], &bl[6 + i]); }
  oracle_msm(srs_g1, cz, n + 3, &pts[3], nthreads, 0);
  tr_begin(&fs, "alpha"); tr_point(&fs, &pts[3]); tr_finish(&fs, &alpha);
  memcpy(qk, lqk, n * 32);
  for (unsigned i = 0; i < nb_public; i++) qk[i] = sol[i];
  to_canonical(qk, log2n, nthreads);
  fe *el = malloc(N4 * 32), *er = malloc(N4 * 32), *eo = malloc(N4 * 32), *ez = malloc(N4 * 32), *eqk = malloc(N4 * 32), *t = malloc(N4 * 32);
  to_coset(el, cl, n + 2, log_big, nthreads); to_coset(er, cr, n + 2, log_big, nthreads); to_coset(eo, co, n + 2, log_big, nthreads);
  to_coset(ez, cz, n + 3, log_big, nthreads); to_coset(eqk, qk, n, log_big, nthreads);
  quot_arg qa;
  qa.log2n = log2n; qa.log_big = log_big;
  qa.el = el; qa.er = er; qa.eo = eo; qa.ez = ez; qa.eqk = eqk;
  qa.ql = cosets_big[0]; qa.qr = cosets_big[1]; qa.qm = cosets_big[2]; qa.qo = cosets_big[3];
  qa.s1 = cosets_big[4]; qa.s2 = cosets_big[5]; qa.s3 = cosets_big[6]; qa.lone = cosets_big[7];
  qa.tw_big = db->tw; qa.alpha = alpha; qa.beta = beta; qa.gamma = gamma; qa.u = u;
  FMUL(&qa.beta_u, &beta, &u); FMUL(&qa.beta_uu, &qa.beta_u, &u);
  {
    fe un, w, wi = FR.one, v;
    fr_pow_u64(&un, &u, (uint64_t)n);
    fr_pow_u64(&w, &db->tw[N4 >= 2 ? 1 : 0], (uint64_t)n); /* w_{4n}^n */
    size_t ratio = N4 >> log2n;
    for (size_t i = 0; i < 8; i++) {
      if (i < ratio) { FMUL(&v, &un, &wi); FSUB(&v, &v, &FR.one); fe_inv(&FR, &qa.xn_inv[i], &v); FMUL(&wi, &wi, &w); }
      else qa.xn_inv[i] = FR.one;
    }
  }
  qa.out = t;
  parallel_for(N4, nthreads, quot_range, &qa);
  oracle_ntt(t, log_big, 1, 1, 1, nthreads); /* FFTInverse(DIT, coset): bit-reversed Lagrange-coset -> canonical */
  for (int k = 0; k < 3; k++) oracle_msm(srs_g1, t + k * m, m, &pts[4 + k], nthreads, 0);
  tr_begin(&fs, "zeta"); for (int i = 0; i < 3; i++) tr_point(&fs, &pts[4 + i]); tr_finish(&fs, &zeta);
  fe lz, rz, oz, s1z, s2z, zu, zeta_shift, omega;
  if (n >= 4) omega = dn->tw[1];
  else if (n == 2) fe_neg(&FR, &omega, &FR.one);
  else omega = FR.one;
  FMUL(&zeta_shift, &zeta, &omega);
  poly_eval(&lz, cl, n + 2, &zeta); poly_eval(&rz, cr, n + 2, &zeta); poly_eval(&oz, co, n + 2, &zeta);
  poly_eval(&s1z, s1, n, &zeta); poly_eval(&s2z, s2, n, &zeta); poly_eval(&zu, cz, n + 3, &zeta_shift);
  fe* quot = malloc((n + 8) * 32);
  poly_div_x_minus_a(quot, cz, n + 3, &zeta_shift);
  oracle_msm(srs_g1, quot, n + 2, &pts[8], nthreads, 0);
  /* linearised polynomial */
  fe c1, c2, lagv, rl, tt, t2, uz, uuz;
  FMUL(&c1, &s1z, &beta); FADD(&c1, &c1, &lz); FADD(&c1, &c1, &gamma);
  FMUL(&tt, &s2z, &beta); FADD(&tt, &tt, &rz); FADD(&tt, &tt, &gamma);
  FMUL(&c1, &c1, &tt); FMUL(&c1, &c1, &zu); FMUL(&c1, &c1, &beta);
  FMUL(&uz, &zeta, &u); FMUL(&uuz, &uz, &u);
  FMUL(&c2, &beta, &zeta); FADD(&c2, &c2, &lz); FADD(&c2, &c2, &gamma);
  FMUL(&tt, &beta, &uz); FADD(&tt, &tt, &rz); FADD(&tt, &tt, &gamma); FMUL(&c2, &c2, &tt);
  FMUL(&tt, &beta, &uuz); FADD(&tt, &tt, &oz); FADD(&tt, &tt, &gamma); FMUL(&c2, &c2, &tt);
  fe_neg(&FR, &c2, &c2);
  fr_pow_u64(&lagv, &zeta, (uint64_t)n); FSUB(&lagv, &lagv, &FR.one);
  FSUB(&tt, &zeta, &FR.one); fe_inv(&FR, &t2, &tt); FMUL(&lagv, &lagv, &t2);
  FMUL(&lagv, &lagv, &alpha); FMUL(&lagv, &lagv, &alpha); FMUL(&lagv, &lagv, &dn->ninv);
  FMUL(&rl, &rz, &lz);
  fe* lin = malloc((n + 8) * 32);
  for (size_t i = 0; i < n + 3; i++) {
    fe v, w;
    FMUL(&v, &cz[i], &c2);
    if (i < n) { FMUL(&w, &s3[i], &c1); FADD(&v, &v, &w); }
    FMUL(&v, &v, &alpha);
    if (i < n) {
      FMUL(&w, &qm[i], &rl); FADD(&v, &v, &w);
      FMUL(&w, &ql[i], &lz); FADD(&v, &v, &w);
      FMUL(&w, &qr[i], &rz); FADD(&v, &v, &w);
      FMUL(&w, &qo[i], &oz); FADD(&v, &v, &w);
      FADD(&v, &v, &cqk[i]);
    }
    FMUL(&w, &cz[i], &lagv); FADD(&lin[i], &v, &w);
  }
  oracle_msm(srs_g1, lin, n + 3, &pts[9], nthreads, 0);
  /* folded H */
  fe zpm; fr_pow_u64(&zpm, &zeta, (uint64_t)m);
  fe* fh = malloc((n + 8) * 32);
  for (size_t i = 0; i < m; i++) {
    fe v; FMUL(&v, &t[2 * m + i], &zpm); FADD(&v, &v, &t[m + i]); FMUL(&v, &v, &zpm); FADD(&fh[i], &v, &t[i]);
  }
  { /* folded digest = H0 + zpm*(H1 + zpm*H2) by double-and-add */
    fe zr; fe_from_mont(&FR, &zr, &zpm);
    g1x acc; g1x_set_inf(&acc); g1x_add_mixed(&acc, &pts[6], 0);
    for (int k = 1; k >= 0; k--) {
      g1a base; g1x_to_affine(&base, &acc);
      g1x_set_inf(&acc);
      for (int b = 255; b >= 0; b--) { g1x_double(&acc); if ((zr.l[b / 64] >> (b % 64)) & 1) g1x_add_mixed(&acc, &base, 0); }
      g1x_add_mixed(&acc, &pts[4 + k], 0);
    }
    g1x_to_affine(&pts[10], &acc);
  }
  fe claimed[7];
  poly_eval(&claimed[0], fh, m, &zeta); poly_eval(&claimed[1], lin, n + 3, &zeta);
  claimed[2] = lz; claimed[3] = rz; claimed[4] = oz; claimed[5] = s1z; claimed[6] = s2z;
  transcript kz; kz.have_prev = 0;
  fe gk;
  tr_begin(&kz, "gamma"); tr_fr(&kz, &zeta);
  tr_point(&kz, &pts[10]); tr_point(&kz, &pts[9]);
  for (int i = 0; i < 3; i++) tr_point(&kz, &pts[i]);
  tr_point(&kz, &vk[0]); tr_point(&kz, &vk[1]);
  tr_finish(&kz, &gk);
  const fe* polys[7] = {fh, lin, cl, cr, co, s1, s2};
  const size_t lens[7] = {m, n + 3, n + 2, n + 2, n + 2, n, n};
  fe* folded = calloc(n + 8, 32);
  fe gp = FR.one;
  for (int k = 0; k < 7; k++) {
    for (size_t j = 0; j < lens[k]; j++) { fe v; FMUL(&v, &polys[k][j], &gp); FADD(&folded[j], &folded[j], &v); }
    FMUL(&gp, &gp, &gk);
  }
  poly_div_x_minus_a(quot, folded, n + 3, &zeta);
  oracle_msm(srs_g1, quot, n + 2, &pts[7], nthreads, 0);
  uint8_t* out = (uint8_t*)proof_out;
  memcpy(out, pts, 9 * 64);
  memcpy(out + 576, claimed, 7 * 32);
  memcpy(out + 576 + 224, &zu, 32);
  free(l); free(r); free(o); free(cl); free(cr); free(co); free(cz); free(qk); free(ident); free(num); free(den); free(pre);
  free(el); free(er); free(eo); free(ez); free(eqk); free(t); free(quot); free(lin); free(fh); free(folded);
  return 0;
}
