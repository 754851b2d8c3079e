// BN254 G1 (y^2 = x^3 + 3 over fp) point arithmetic for sm_100a.
//
// Replaces (device side) gnark-crypto v0.9.1 ecc/bn254/g1.go: G1Affine (64 B, X||Y Montgomery limbs, (0,0) = infinity)
// and the extended-Jacobian bucket type g1JacExtended (X, Y, ZZ, ZZZ with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) that
// MultiExp accumulates into (reached from /root/reference/gnark_backend_ffi/backend/plonk/plonk.go:21,67 through
// kzg.Commit).  Formulas: EFD "xyzz" madd-2008-s / add-2008-s / dbl-2008-s-1 with a = 0; Y3 = R*(Q - X3) - Y1*PPP is
// formed as ONE two-product Montgomery sweep (fe_mul2add): 72 multiply-adds fewer per addition (1216 instead of 1288).
#pragma once
#include "field.cuh"

namespace b200zk {

struct G1Affine {
  Fp x, y;
};

struct G1XYZZ {
  Fp x, y, zz, zzz;
};

__device__ __forceinline__ bool g1_is_inf(const G1Affine& a) { return fe_is_zero(a.x) && fe_is_zero(a.y); }
__device__ __forceinline__ bool g1_is_inf(const G1XYZZ& p) { return fe_is_zero(p.zz); }

__device__ __forceinline__ G1XYZZ g1_xyzz_inf() {
  G1XYZZ p;
  p.x = fe_one<FpParams>();
  p.y = fe_one<FpParams>();
  p.zz = fe_zero<FpParams>();
  p.zzz = fe_zero<FpParams>();
  return p;
}

__device__ __forceinline__ G1Affine g1_load_affine(const void* base, size_t idx) {
  const uint4* q = reinterpret_cast<const uint4*>(base) + idx * 4;
  uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
  G1Affine r;
  r.x.l[0] = a.x; r.x.l[1] = a.y; r.x.l[2] = a.z; r.x.l[3] = a.w;
  r.x.l[4] = b.x; r.x.l[5] = b.y; r.x.l[6] = b.z; r.x.l[7] = b.w;
  r.y.l[0] = c.x; r.y.l[1] = c.y; r.y.l[2] = c.z; r.y.l[3] = c.w;
  r.y.l[4] = d.x; r.y.l[5] = d.y; r.y.l[6] = d.z; r.y.l[7] = d.w;
  return r;
}

__device__ __forceinline__ void g1_store_affine(void* base, size_t idx, const G1Affine& p) {
  char* q = reinterpret_cast<char*>(base) + idx * 64;
  fe_store(q, p.x);
  fe_store(q + 32, p.y);
}

__device__ __forceinline__ G1XYZZ g1_load_xyzz(const void* base, size_t idx) {
  const char* q = reinterpret_cast<const char*>(base) + idx * 128;
  G1XYZZ p;
  p.x = fe_load<FpParams>(q);
  p.y = fe_load<FpParams>(q + 32);
  p.zz = fe_load<FpParams>(q + 64);
  p.zzz = fe_load<FpParams>(q + 96);
  return p;
}

__device__ __forceinline__ void g1_store_xyzz(void* base, size_t idx, const G1XYZZ& p) {
  char* q = reinterpret_cast<char*>(base) + idx * 128;
  fe_store(q, p.x);
  fe_store(q + 32, p.y);
  fe_store(q + 64, p.zz);
  fe_store(q + 96, p.zzz);
}

// p = 2*a for affine a (not infinity, y != 0 always holds on this curve: no 2-torsion)
static __device__ __noinline__ G1XYZZ g1_double_mixed(G1Affine a) {
  G1XYZZ p;
  Fp U = fe_dbl(a.y);
  Fp V = fe_sqr(U);
  Fp W = fe_mul(U, V);
  Fp S = fe_mul(a.x, V);
  Fp XX = fe_sqr(a.x);
  Fp M = fe_add(fe_dbl(XX), XX);
  Fp X3 = fe_sub(fe_sqr(M), fe_dbl(S));
  Fp Y3 = fe_mul2add(M, fe_sub(S, X3), W, fe_neg(a.y));
  p.x = X3;
  p.y = Y3;
  p.zz = V;
  p.zzz = W;
  return p;
}

static __device__ __noinline__ G1XYZZ g1_double_v(G1XYZZ p) {
  if (g1_is_inf(p)) return p;
  Fp U = fe_dbl(p.y);
  Fp V = fe_sqr(U);
  Fp W = fe_mul(U, V);
  Fp S = fe_mul(p.x, V);
  Fp XX = fe_sqr(p.x);
  Fp M = fe_add(fe_dbl(XX), XX);
  Fp X3 = fe_sub(fe_sqr(M), fe_dbl(S));
  Fp Y3 = fe_mul2add(M, fe_sub(S, X3), W, fe_neg(p.y));
  p.x = X3;
  p.y = Y3;
  p.zz = fe_mul(V, p.zz);
  p.zzz = fe_mul(W, p.zzz);
  return p;
}
__device__ __forceinline__ void g1_double(G1XYZZ& p) { p = g1_double_v(p); }

// p += a (a affine).  Mirrors g1JacExtended.addMixed: handles a = inf, p = inf, a = p, a = -p.
__device__ __forceinline__ void g1_add_mixed(G1XYZZ& p, const G1Affine& a) {
  if (g1_is_inf(a)) return;
  if (g1_is_inf(p)) {
    p.x = a.x;
    p.y = a.y;
    p.zz = fe_one<FpParams>();
    p.zzz = fe_one<FpParams>();
    return;
  }
  Fp P = fe_sub(fe_mul(a.x, p.zz), p.x);
  Fp R = fe_sub(fe_mul(a.y, p.zzz), p.y);
  if (fe_is_zero(P)) {
    if (fe_is_zero(R)) {
      p = g1_double_mixed(a);
    } else {
      p = g1_xyzz_inf();
    }
    return;
  }
  Fp PP = fe_sqr(P);
  Fp PPP = fe_mul(P, PP);
  Fp Q = fe_mul(p.x, PP);
  Fp X3 = fe_sub(fe_sub(fe_sqr(R), PPP), fe_dbl(Q));
  Fp Y3 = fe_mul2add(R, fe_sub(Q, X3), fe_neg(p.y), PPP);   // one reduction for the two products
  p.x = X3;
  p.y = Y3;
  p.zz = fe_mul(p.zz, PP);
  p.zzz = fe_mul(p.zzz, PPP);
}

// p += q (both XYZZ)
static __device__ __noinline__ G1XYZZ g1_add_v(G1XYZZ p, G1XYZZ q) {
  if (g1_is_inf(q)) return p;
  if (g1_is_inf(p)) return q;
  Fp U1 = fe_mul(p.x, q.zz);
  Fp U2 = fe_mul(q.x, p.zz);
  Fp S1 = fe_mul(p.y, q.zzz);
  Fp S2 = fe_mul(q.y, p.zzz);
  Fp P = fe_sub(U2, U1);
  Fp R = fe_sub(S2, S1);
  if (fe_is_zero(P)) {
    if (fe_is_zero(R)) return g1_double_v(p);
    return g1_xyzz_inf();
  }
  Fp PP = fe_sqr(P);
  Fp PPP = fe_mul(P, PP);
  Fp Q = fe_mul(U1, PP);
  Fp X3 = fe_sub(fe_sub(fe_sqr(R), PPP), fe_dbl(Q));
  Fp Y3 = fe_mul2add(R, fe_sub(Q, X3), fe_neg(S1), PPP);
  p.x = X3;
  p.y = Y3;
  p.zz = fe_mul(fe_mul(p.zz, q.zz), PP);
  p.zzz = fe_mul(fe_mul(p.zzz, q.zzz), PPP);
  return p;
}
__device__ __forceinline__ void g1_add(G1XYZZ& p, const G1XYZZ& q) { p = g1_add_v(p, q); }

// canonical affine (gnark FromJacobian / fromJacExtended semantics: infinity -> (0,0))
static __device__ __noinline__ G1Affine g1_to_affine(G1XYZZ p) {
  G1Affine r;
  if (g1_is_inf(p)) {
    r.x = fe_zero<FpParams>();
    r.y = fe_zero<FpParams>();
    return r;
  }
  Fp inv = fe_inv(fe_mul(p.zz, p.zzz));
  r.x = fe_mul(fe_mul(p.x, p.zzz), inv);
  r.y = fe_mul(fe_mul(p.y, p.zz), inv);
  return r;
}

// the same for code in which a single thread normalises one point (see fe_inv_single)
static __device__ __noinline__ G1Affine g1_to_affine_single(G1XYZZ p) {
  G1Affine r;
  if (g1_is_inf(p)) {
    r.x = fe_zero<FpParams>();
    r.y = fe_zero<FpParams>();
    return r;
  }
  Fp inv = fe_inv_single(fe_mul(p.zz, p.zzz));
  r.x = fe_mul(fe_mul(p.x, p.zzz), inv);
  r.y = fe_mul(fe_mul(p.y, p.zz), inv);
  return r;
}

}  // namespace b200zk
