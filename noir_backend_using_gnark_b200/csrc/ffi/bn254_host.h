// Host-side BN254 for the FFI stand-in: what stays on the CPU in the reference too (gnark's plonk.Verify = two
// pairings + a 7-term MSM, SRS G2 elements, key / proof (de)serialisation).  O(1) work per call; not the hot path.
//   Fp / Fr   : csrc/host_field.h (4 x u64 Montgomery)
//   Fp2       : Fp[u]/(u^2+1)
//   G1 / G2   : affine + Jacobian (G1) host arithmetic, gnark compressed encodings (G1Affine.Bytes / G2Affine.Bytes)
//   pairing   : optimal ate, Fp12 = Fp[w]/(w^12 - 18 w^6 + 82) with u = w^6 - 9, generic affine line functions
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "../host_field.h"
#include "ffi_consts.h"

namespace b200zk {
namespace ffi {

using host::Fe4;
using host::HFP;
using host::HFR;

inline Fe4 fp_from_limbs(const uint64_t l[4]) { Fe4 r; memcpy(r.l, l, 32); return r; }
inline Fe4 fp_zero() { Fe4 r = {{0, 0, 0, 0}}; return r; }
inline bool fe_eq(const Fe4& a, const Fe4& b) { return memcmp(a.l, b.l, 32) == 0; }
inline Fe4 fp_mul(const Fe4& a, const Fe4& b) { return host::mul(HFP, a, b); }
inline Fe4 fp_add(const Fe4& a, const Fe4& b) { return host::add(HFP, a, b); }
inline Fe4 fp_sub(const Fe4& a, const Fe4& b) { return host::sub(HFP, a, b); }
inline Fe4 fp_neg(const Fe4& a) { return host::neg(HFP, a); }
inline Fe4 fp_inv(const Fe4& a) { return host::inv(HFP, a); }
inline Fe4 fp_small(uint64_t v) { return host::from_u64(HFP, v); }
inline Fe4 fp_pow(const Fe4& a, const uint64_t e[4]) {
  Fe4 acc = HFP.one, base = a;
  for (int i = 0; i < 256; i++) {
    if ((e[i / 64] >> (i % 64)) & 1) acc = fp_mul(acc, base);
    base = fp_mul(base, base);
  }
  return acc;
}
// regular-form value > (p-1)/2 ?
inline bool fp_lex_largest(const Fe4& a_mont) {
  Fe4 r = host::from_mont(HFP, a_mont);
  for (int i = 3; i >= 0; i--)
    if (r.l[i] != P_MINUS1_DIV2[i]) return r.l[i] > P_MINUS1_DIV2[i];
  return false;
}
inline void fp_to_be(const Fe4& a_mont, uint8_t out[32]) { host::marshal(HFP, a_mont, out); }
// 32-byte big-endian -> Montgomery fp; *ok = value < p
inline Fe4 fp_from_be(const uint8_t in[32], bool* ok) {
  Fe4 v;
  for (int i = 0; i < 4; i++) {
    uint64_t w = 0;
    for (int b = 0; b < 8; b++) w |= (uint64_t)in[31 - (8 * i + b)] << (8 * b);
    v.l[i] = w;
  }
  if (ok) *ok = !host::geq(v.l, HFP.m);
  return host::to_mont(HFP, v);
}

// ---------------------------------------------------------------------------------------------- G1
struct G1 { Fe4 x, y; bool inf; };
struct G1J { Fe4 x, y, z; };  // Jacobian, z = 0 -> infinity

inline G1 g1_from_image(const uint8_t b[64]) {
  G1 p;
  memcpy(p.x.l, b, 32);
  memcpy(p.y.l, b + 32, 32);
  p.inf = host::is_zero(p.x) && host::is_zero(p.y);
  return p;
}
inline void g1_to_image(const G1& p, uint8_t b[64]) {
  if (p.inf) { memset(b, 0, 64); return; }
  memcpy(b, p.x.l, 32);
  memcpy(b + 32, p.y.l, 32);
}
inline G1J g1j_inf() { G1J r; r.x = HFP.one; r.y = HFP.one; r.z = fp_zero(); return r; }
inline G1J g1j_double(const G1J& p) {
  if (host::is_zero(p.z)) return p;
  Fe4 A = fp_mul(p.x, p.x), B = fp_mul(p.y, p.y), C = fp_mul(B, B);
  Fe4 t = fp_add(p.x, B);
  Fe4 D = fp_sub(fp_sub(fp_mul(t, t), A), C);
  D = fp_add(D, D);
  Fe4 E = fp_add(fp_add(A, A), A), F = fp_mul(E, E);
  G1J r;
  r.x = fp_sub(F, fp_add(D, D));
  Fe4 c8 = fp_add(C, C); c8 = fp_add(c8, c8); c8 = fp_add(c8, c8);
  r.y = fp_sub(fp_mul(E, fp_sub(D, r.x)), c8);
  r.z = fp_mul(fp_add(p.y, p.y), p.z);
  return r;
}
inline G1J g1j_add_affine(const G1J& p, const G1& q) {
  if (q.inf) return p;
  if (host::is_zero(p.z)) { G1J r; r.x = q.x; r.y = q.y; r.z = HFP.one; return r; }
  Fe4 z2 = fp_mul(p.z, p.z), u2 = fp_mul(q.x, z2), s2 = fp_mul(fp_mul(q.y, p.z), z2);
  Fe4 h = fp_sub(u2, p.x), rr = fp_sub(s2, p.y);
  if (host::is_zero(h)) {
    if (host::is_zero(rr)) return g1j_double(p);
    return g1j_inf();
  }
  Fe4 hh = fp_mul(h, h), hhh = fp_mul(h, hh), v = fp_mul(p.x, hh);
  G1J r;
  r.x = fp_sub(fp_sub(fp_mul(rr, rr), hhh), fp_add(v, v));
  r.y = fp_sub(fp_mul(rr, fp_sub(v, r.x)), fp_mul(p.y, hhh));
  r.z = fp_mul(p.z, h);
  return r;
}
inline G1 g1j_to_affine(const G1J& p) {
  G1 r;
  if (host::is_zero(p.z)) { r.x = r.y = fp_zero(); r.inf = true; return r; }
  Fe4 zi = fp_inv(p.z), zi2 = fp_mul(zi, zi);
  r.x = fp_mul(p.x, zi2);
  r.y = fp_mul(p.y, fp_mul(zi2, zi));
  r.inf = false;
  return r;
}
inline G1 g1_neg(const G1& p) { G1 r = p; if (!p.inf) r.y = fp_neg(p.y); return r; }
// scalar given as Montgomery fr
inline G1J g1_mul_j(const G1& p, const Fe4& k_mont) {
  Fe4 k = host::from_mont(HFR, k_mont);
  G1J acc = g1j_inf();
  for (int b = 255; b >= 0; b--) {
    acc = g1j_double(acc);
    if ((k.l[b / 64] >> (b % 64)) & 1) acc = g1j_add_affine(acc, p);
  }
  return acc;
}
inline G1 g1_mul(const G1& p, const Fe4& k_mont) { return g1j_to_affine(g1_mul_j(p, k_mont)); }
inline G1 g1_add(const G1& a, const G1& b) {
  G1J j = g1j_inf();
  j = g1j_add_affine(j, a);
  j = g1j_add_affine(j, b);
  return g1j_to_affine(j);
}
// sum_k scalars[k] * points[k] for a handful of points (the verifier's linear combinations): Straus with 4-bit
// windows — the 252 doublings are shared, the multiples 1P..15P of every point are normalised to affine with one
// inversion (Montgomery's trick) so that every addition in the main loop is a mixed one.  Scalars: Montgomery fr.
inline G1J g1_msm_small(const std::vector<G1>& points, const std::vector<Fe4>& scalars_mont) {
  const size_t n = points.size();
  std::vector<G1J> jac(15 * n);
  for (size_t k = 0; k < n; k++) {
    G1J m = g1j_inf();
    for (int d = 0; d < 15; d++) {
      m = g1j_add_affine(m, points[k]);
      jac[15 * k + d] = m;
    }
  }
  // batch normalisation (entries with z = 0 stay at infinity)
  std::vector<Fe4> prefix(jac.size());
  Fe4 run = HFP.one;
  for (size_t i = 0; i < jac.size(); i++) {
    prefix[i] = run;
    if (!host::is_zero(jac[i].z)) run = fp_mul(run, jac[i].z);
  }
  Fe4 inv_run = fp_inv(run);
  std::vector<G1> table(jac.size());
  for (size_t i = jac.size(); i-- > 0;) {
    G1& t = table[i];
    if (host::is_zero(jac[i].z)) {
      t.x = t.y = fp_zero();
      t.inf = true;
      continue;
    }
    const Fe4 zi = fp_mul(inv_run, prefix[i]);
    inv_run = fp_mul(inv_run, jac[i].z);
    const Fe4 zi2 = fp_mul(zi, zi);
    t.x = fp_mul(jac[i].x, zi2);
    t.y = fp_mul(jac[i].y, fp_mul(zi2, zi));
    t.inf = false;
  }
  std::vector<Fe4> sc(n);
  for (size_t k = 0; k < n; k++) sc[k] = host::from_mont(HFR, scalars_mont[k]);
  G1J acc = g1j_inf();
  for (int w = 63; w >= 0; w--) {
    for (int t = 0; t < 4; t++) acc = g1j_double(acc);
    for (size_t k = 0; k < n; k++) {
      const unsigned d = (unsigned)(sc[k].l[w / 16] >> (4 * (w % 16))) & 15u;
      if (d) acc = g1j_add_affine(acc, table[15 * k + d - 1]);
    }
  }
  return acc;
}
inline G1 g1_generator() { G1 g; g.x = HFP.one; g.y = fp_add(HFP.one, HFP.one); g.inf = false; return g; }

// G1Affine.Bytes(): 32-byte BE X, flags 10 (smallest y) / 11 (largest y) / 01 (infinity)
inline void g1_compress(const G1& p, uint8_t out[32]) {
  if (p.inf) { memset(out, 0, 32); out[0] = 0x40; return; }
  fp_to_be(p.x, out);
  out[0] |= fp_lex_largest(p.y) ? 0xC0 : 0x80;
}
inline bool g1_decompress(const uint8_t in[32], G1* out) {
  const unsigned flag = in[0] >> 6;
  if (flag == 1) { out->x = out->y = fp_zero(); out->inf = true; return true; }
  if (flag == 0) return false;
  uint8_t b[32];
  memcpy(b, in, 32);
  b[0] &= 0x3f;
  bool ok;
  Fe4 x = fp_from_be(b, &ok);
  if (!ok) return false;
  Fe4 rhs = fp_add(fp_mul(fp_mul(x, x), x), fp_small(3));
  Fe4 y = fp_pow(rhs, EXP_P_PLUS1_DIV4);
  if (!fe_eq(fp_mul(y, y), rhs)) return false;
  if (fp_lex_largest(y) != (flag == 3)) y = fp_neg(y);
  out->x = x; out->y = y; out->inf = false;
  return true;
}

// ---------------------------------------------------------------------------------------------- Fp2, G2
struct Fp2 { Fe4 a0, a1; };
inline Fp2 f2_zero() { Fp2 r; r.a0 = r.a1 = fp_zero(); return r; }
inline Fp2 f2_one() { Fp2 r; r.a0 = HFP.one; r.a1 = fp_zero(); return r; }
inline bool f2_is_zero(const Fp2& a) { return host::is_zero(a.a0) && host::is_zero(a.a1); }
inline bool f2_eq(const Fp2& a, const Fp2& b) { return fe_eq(a.a0, b.a0) && fe_eq(a.a1, b.a1); }
inline Fp2 f2_add(const Fp2& a, const Fp2& b) { Fp2 r; r.a0 = fp_add(a.a0, b.a0); r.a1 = fp_add(a.a1, b.a1); return r; }
inline Fp2 f2_sub(const Fp2& a, const Fp2& b) { Fp2 r; r.a0 = fp_sub(a.a0, b.a0); r.a1 = fp_sub(a.a1, b.a1); return r; }
inline Fp2 f2_neg(const Fp2& a) { Fp2 r; r.a0 = fp_neg(a.a0); r.a1 = fp_neg(a.a1); return r; }
inline Fp2 f2_conj(const Fp2& a) { Fp2 r; r.a0 = a.a0; r.a1 = fp_neg(a.a1); return r; }
inline Fp2 f2_mul(const Fp2& a, const Fp2& b) {
  Fp2 r;
  r.a0 = fp_sub(fp_mul(a.a0, b.a0), fp_mul(a.a1, b.a1));
  r.a1 = fp_add(fp_mul(a.a0, b.a1), fp_mul(a.a1, b.a0));
  return r;
}
inline Fp2 f2_inv(const Fp2& a) {
  Fe4 d = fp_inv(fp_add(fp_mul(a.a0, a.a0), fp_mul(a.a1, a.a1)));
  Fp2 r; r.a0 = fp_mul(a.a0, d); r.a1 = fp_neg(fp_mul(a.a1, d));
  return r;
}
inline Fp2 f2_pow(const Fp2& a, const uint64_t e[4]) {
  Fp2 acc = f2_one(), base = a;
  for (int i = 0; i < 256; i++) {
    if ((e[i / 64] >> (i % 64)) & 1) acc = f2_mul(acc, base);
    base = f2_mul(base, base);
  }
  return acc;
}
// square root in Fp2 for p = 3 mod 4 (complex method); false if a is a non-residue
inline bool f2_sqrt(const Fp2& a, Fp2* out) {
  if (f2_is_zero(a)) { *out = a; return true; }
  Fp2 a1 = f2_pow(a, EXP_P_MINUS3_DIV4);
  Fp2 alpha = f2_mul(a1, f2_mul(a1, a));
  Fp2 a0 = f2_mul(f2_conj(alpha), alpha);
  Fp2 minus_one = f2_neg(f2_one());
  if (f2_eq(a0, minus_one)) return false;
  Fp2 x0 = f2_mul(a1, a);
  Fp2 x;
  if (f2_eq(alpha, minus_one)) {
    x.a0 = fp_neg(x0.a1);  // i * x0
    x.a1 = x0.a0;
  } else {
    Fp2 b = f2_pow(f2_add(f2_one(), alpha), EXP_P_MINUS1_DIV2);
    x = f2_mul(b, x0);
  }
  if (!f2_eq(f2_mul(x, x), a)) return false;
  *out = x;
  return true;
}

struct G2 { Fp2 x, y; bool inf; };
inline G2 g2_generator() {
  G2 g;
  g.x.a0 = fp_from_limbs(G2_X0); g.x.a1 = fp_from_limbs(G2_X1);
  g.y.a0 = fp_from_limbs(G2_Y0); g.y.a1 = fp_from_limbs(G2_Y1);
  g.inf = false;
  return g;
}
inline Fp2 twist_b() { Fp2 b; b.a0 = fp_from_limbs(TWIST_B0); b.a1 = fp_from_limbs(TWIST_B1); return b; }
inline G2 g2_add(const G2& a, const G2& b) {
  if (a.inf) return b;
  if (b.inf) return a;
  Fp2 lam;
  if (f2_eq(a.x, b.x)) {
    if (f2_is_zero(f2_add(a.y, b.y))) { G2 r; r.x = r.y = f2_zero(); r.inf = true; return r; }
    Fp2 xx = f2_mul(a.x, a.x);
    lam = f2_mul(f2_add(f2_add(xx, xx), xx), f2_inv(f2_add(a.y, a.y)));
  } else {
    lam = f2_mul(f2_sub(b.y, a.y), f2_inv(f2_sub(b.x, a.x)));
  }
  G2 r;
  r.x = f2_sub(f2_sub(f2_mul(lam, lam), a.x), b.x);
  r.y = f2_sub(f2_mul(lam, f2_sub(a.x, r.x)), a.y);
  r.inf = false;
  return r;
}
inline G2 g2_mul(const G2& p, const Fe4& k_mont) {
  Fe4 k = host::from_mont(HFR, k_mont);
  G2 acc; acc.x = acc.y = f2_zero(); acc.inf = true;
  for (int b = 255; b >= 0; b--) {
    acc = g2_add(acc, acc);
    if ((k.l[b / 64] >> (b % 64)) & 1) acc = g2_add(acc, p);
  }
  return acc;
}
// membership in the order-r subgroup of the twist (its cofactor is not 1): [r]Q = infinity.  gnark's G2 decoder performs
// this check on every decoded point; here it guards the two G2 elements read from srs.hex (two 254-bit ladders, < 1 ms).
inline bool g2_in_subgroup(const G2& q) {
  if (q.inf) return true;
  G2 acc; acc.x = acc.y = f2_zero(); acc.inf = true;
  for (int b = 255; b >= 0; b--) {
    acc = g2_add(acc, acc);
    if ((HFR.m[b / 64] >> (b % 64)) & 1) acc = g2_add(acc, q);
  }
  return acc.inf;
}
// G2Affine.Bytes(): X.A1 || X.A0 (32-byte BE each), flags in the first byte; y sign: A1 decides unless it is zero
inline bool f2_lex_largest(const Fp2& y) { return host::is_zero(y.a1) ? fp_lex_largest(y.a0) : fp_lex_largest(y.a1); }
inline void g2_compress(const G2& p, uint8_t out[64]) {
  if (p.inf) { memset(out, 0, 64); out[0] = 0x40; return; }
  fp_to_be(p.x.a1, out);
  fp_to_be(p.x.a0, out + 32);
  out[0] |= f2_lex_largest(p.y) ? 0xC0 : 0x80;
}
inline bool g2_decompress(const uint8_t in[64], G2* out) {
  const unsigned flag = in[0] >> 6;
  if (flag == 1) { out->x = out->y = f2_zero(); out->inf = true; return true; }
  if (flag == 0) return false;
  uint8_t b[32];
  memcpy(b, in, 32);
  b[0] &= 0x3f;
  bool ok1, ok0;
  Fp2 x;
  x.a1 = fp_from_be(b, &ok1);
  x.a0 = fp_from_be(in + 32, &ok0);
  if (!ok1 || !ok0) return false;
  Fp2 rhs = f2_add(f2_mul(f2_mul(x, x), x), twist_b());
  Fp2 y;
  if (!f2_sqrt(rhs, &y)) return false;
  if (f2_lex_largest(y) != (flag == 3)) y = f2_neg(y);
  out->x = x; out->y = y; out->inf = false;
  return true;
}

// ---------------------------------------------------------------------------------------------- Fp12 + pairing
struct F12 { Fe4 c[12]; };
inline F12 f12_zero() { F12 r; for (auto& x : r.c) x = fp_zero(); return r; }
inline F12 f12_one() { F12 r = f12_zero(); r.c[0] = HFP.one; return r; }
inline bool f12_eq(const F12& a, const F12& b) { for (int i = 0; i < 12; i++) if (!fe_eq(a.c[i], b.c[i])) return false; return true; }
inline F12 f12_add(const F12& a, const F12& b) { F12 r; for (int i = 0; i < 12; i++) r.c[i] = fp_add(a.c[i], b.c[i]); return r; }
inline F12 f12_sub(const F12& a, const F12& b) { F12 r; for (int i = 0; i < 12; i++) r.c[i] = fp_sub(a.c[i], b.c[i]); return r; }
inline F12 f12_neg(const F12& a) { F12 r; for (int i = 0; i < 12; i++) r.c[i] = fp_neg(a.c[i]); return r; }
inline F12 f12_scalar(const F12& a, uint64_t k) { Fe4 s = fp_small(k); F12 r; for (int i = 0; i < 12; i++) r.c[i] = fp_mul(a.c[i], s); return r; }
// w^12 = 18 w^6 - 82 applied to a product of degree <= 22
inline F12 f12_reduce(Fe4* t) {
  for (int k = 22; k >= 12; k--) {
    if (host::is_zero(t[k])) continue;
    // 18 t = 16 t + 2 t, 82 t = 64 t + 16 t + 2 t: doublings instead of multiplications by the small constants
    const Fe4 t2 = fp_add(t[k], t[k]), t4 = fp_add(t2, t2), t8 = fp_add(t4, t4), t16 = fp_add(t8, t8);
    const Fe4 t32 = fp_add(t16, t16), t64 = fp_add(t32, t32);
    t[k - 6] = fp_add(t[k - 6], fp_add(t16, t2));
    t[k - 12] = fp_sub(t[k - 12], fp_add(fp_add(t64, t16), t2));
  }
  F12 r;
  for (int i = 0; i < 12; i++) r.c[i] = t[i];
  return r;
}
// a * b where a has few non-zero coefficients (the lines of the Miller loop): schoolbook rows of a, zeros skipped
inline F12 f12_mul_sparse(const F12& a, const F12& b) {
  Fe4 t[23];
  for (auto& x : t) x = fp_zero();
  for (int i = 0; i < 12; i++) {
    if (host::is_zero(a.c[i])) continue;
    for (int j = 0; j < 12; j++) t[i + j] = fp_add(t[i + j], fp_mul(a.c[i], b.c[j]));
  }
  return f12_reduce(t);
}
// dense product: one level of Karatsuba on a = a0 + a1 w^6 (3 x 36 multiplications instead of 144)
inline F12 f12_mul(const F12& a, const F12& b) {
  auto mul6 = [](const Fe4* x, const Fe4* y, Fe4* out /* 11 */) {
    for (int i = 0; i < 11; i++) out[i] = fp_zero();
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) out[i + j] = fp_add(out[i + j], fp_mul(x[i], y[j]));
  };
  Fe4 lo[11], hi[11], mid[11], sa[6], sb[6];
  mul6(a.c, b.c, lo);
  mul6(a.c + 6, b.c + 6, hi);
  for (int i = 0; i < 6; i++) {
    sa[i] = fp_add(a.c[i], a.c[i + 6]);
    sb[i] = fp_add(b.c[i], b.c[i + 6]);
  }
  mul6(sa, sb, mid);
  Fe4 t[23];
  for (auto& x : t) x = fp_zero();
  for (int i = 0; i < 11; i++) {
    t[i] = fp_add(t[i], lo[i]);
    t[i + 6] = fp_add(t[i + 6], fp_sub(fp_sub(mid[i], lo[i]), hi[i]));
    t[i + 12] = fp_add(t[i + 12], hi[i]);
  }
  return f12_reduce(t);
}
inline F12 f12_sqr(const F12& a) {
  Fe4 t[23];
  for (auto& x : t) x = fp_zero();
  for (int i = 0; i < 12; i++) {
    if (host::is_zero(a.c[i])) continue;
    t[2 * i] = fp_add(t[2 * i], fp_mul(a.c[i], a.c[i]));
    for (int j = i + 1; j < 12; j++) {
      const Fe4 m = fp_mul(a.c[i], a.c[j]);
      t[i + j] = fp_add(t[i + j], fp_add(m, m));
    }
  }
  return f12_reduce(t);
}
inline F12 f12_pow_limbs(const F12& a, const uint64_t* e, int nlimbs) {
  F12 res = f12_one(), base = a;
  for (int i = 0; i < nlimbs * 64; i++) {
    if ((e[i / 64] >> (i % 64)) & 1) res = f12_mul(res, base);
    base = f12_mul(base, base);
  }
  return res;
}
inline int poly_deg(const Fe4* p, int n) {
  int d = n - 1;
  while (d >= 0 && host::is_zero(p[d])) d--;
  return d;
}
// inverse by the extended Euclidean algorithm in Fp[w] against w^12 - 18 w^6 + 82
inline F12 f12_inv(const F12& a) {
  Fe4 lm[13], hm[13], low[13], high[13];
  for (int i = 0; i < 13; i++) { lm[i] = hm[i] = low[i] = high[i] = fp_zero(); }
  lm[0] = HFP.one;
  for (int i = 0; i < 12; i++) low[i] = a.c[i];
  high[0] = fp_small(82);
  high[6] = fp_neg(fp_small(18));
  high[12] = HFP.one;
  while (poly_deg(low, 13) > 0) {
    const int dl = poly_deg(low, 13), dh = poly_deg(high, 13);
    Fe4 quo[13], temp[13];
    for (int i = 0; i < 13; i++) { quo[i] = fp_zero(); temp[i] = high[i]; }
    const Fe4 inv_lead = fp_inv(low[dl]);
    for (int i = dh - dl; i >= 0; i--) {
      Fe4 q = fp_mul(temp[dl + i], inv_lead);
      quo[i] = q;
      for (int c = 0; c <= dl; c++) temp[c + i] = fp_sub(temp[c + i], fp_mul(low[c], q));
    }
    Fe4 nm[13], nw[13];
    for (int i = 0; i < 13; i++) { nm[i] = hm[i]; nw[i] = high[i]; }
    for (int i = 0; i < 13; i++)
      for (int j = 0; j < 13 - i; j++) {
        nm[i + j] = fp_sub(nm[i + j], fp_mul(lm[i], quo[j]));
        nw[i + j] = fp_sub(nw[i + j], fp_mul(low[i], quo[j]));
      }
    for (int i = 0; i < 13; i++) { hm[i] = lm[i]; high[i] = low[i]; lm[i] = nm[i]; low[i] = nw[i]; }
  }
  const Fe4 inv0 = fp_inv(low[0]);
  F12 r;
  for (int i = 0; i < 12; i++) r.c[i] = fp_mul(lm[i], inv0);
  return r;
}
// (c0 + c1 u) * w^shift, u = w^6 - 9
inline F12 f12_from_f2(const Fp2& v, int shift) {
  F12 r = f12_zero();
  r.c[0] = fp_sub(v.a0, fp_mul(v.a1, fp_small(9)));
  r.c[6] = v.a1;
  if (shift) {
    F12 wp = f12_zero();
    wp.c[shift] = HFP.one;
    r = f12_mul(r, wp);
  }
  return r;
}
// ---- Frobenius.  Fp12 = Fp[w]/(w^12 - 18 w^6 + 82), w^6 = xi = 9 + u.  For c in Fp: (c w^i)^p = c w^i xi^(i (p-1)/6), and
// xi^(i (p-1)/6) = a_i + b_i u = (a_i - 9 b_i) + b_i w^6 lies in Fp2 = Fp[w^6].
struct FrobTable {
  Fp2 gamma[12];      // xi^(i (p-1)/6)
  Fp2 twist_x, twist_y;   // xi^((p-1)/3), xi^((p-1)/2): the p-power Frobenius on the twist
};
inline const FrobTable& frob_table() {
  static const FrobTable T = [] {
    FrobTable t;
    uint64_t e[4];  // (p - 1) / 6
    {
      uint64_t pm1[4];
      memcpy(pm1, P_LIMBS, 32);
      pm1[0] -= 1;  // p is odd
      unsigned __int128 rem = 0;
      for (int i = 3; i >= 0; i--) {
        unsigned __int128 cur = (rem << 64) | pm1[i];
        e[i] = (uint64_t)(cur / 6);
        rem = cur % 6;
      }
    }
    Fp2 xi;
    xi.a0 = fp_small(9);
    xi.a1 = HFP.one;
    const Fp2 g = f2_pow(xi, e);
    t.gamma[0] = f2_one();
    for (int i = 1; i < 12; i++) t.gamma[i] = f2_mul(t.gamma[i - 1], g);
    t.twist_x = t.gamma[2];
    t.twist_y = t.gamma[3];
    return t;
  }();
  return T;
}
inline F12 f12_frobenius(const F12& a) {
  const FrobTable& T = frob_table();
  const Fe4 c9 = fp_small(9), c18 = fp_small(18), c82 = fp_small(82);
  F12 r = f12_zero();
  for (int i = 0; i < 12; i++) {
    if (host::is_zero(a.c[i])) continue;
    const Fe4 lo = fp_mul(a.c[i], fp_sub(T.gamma[i].a0, fp_mul(T.gamma[i].a1, c9)));  // coefficient of w^i
    const Fe4 hi = fp_mul(a.c[i], T.gamma[i].a1);                                       // coefficient of w^(i+6)
    r.c[i] = fp_add(r.c[i], lo);
    if (i < 6) {
      r.c[i + 6] = fp_add(r.c[i + 6], hi);
    } else {  // w^(i+6) = w^(i-6) w^12 = 18 w^i - 82 w^(i-6)
      r.c[i] = fp_add(r.c[i], fp_mul(hi, c18));
      r.c[i - 6] = fp_sub(r.c[i - 6], fp_mul(hi, c82));
    }
  }
  return r;
}
// the p^6-power Frobenius: w -> -w (conjugation over Fp6 = Fp[w^2]); the inverse on the cyclotomic subgroup
inline F12 f12_conj6(const F12& a) {
  F12 r = a;
  for (int i = 1; i < 12; i += 2) r.c[i] = fp_neg(a.c[i]);
  return r;
}

// ---- optimal ate pairing.  Q stays on the twist E'(Fp2); the untwist is (x', y') -> (x' w^2, y' w^3), so a line of
// twist slope m through R' evaluated at P = (xP, yP) in G1 is   -yP + (m xP) w + (yR' - m xR') w^3.
inline F12 line_value(const Fp2& m, const G2& r, const G1& p) {
  F12 l = f12_zero();
  l.c[0] = fp_neg(p.y);
  Fp2 mx; mx.a0 = fp_mul(m.a0, p.x); mx.a1 = fp_mul(m.a1, p.x);
  const Fp2 c3 = f2_sub(r.y, f2_mul(m, r.x));
  const Fe4 c9 = fp_small(9);
  l.c[1] = fp_sub(mx.a0, fp_mul(mx.a1, c9)); l.c[7] = mx.a1;
  l.c[3] = fp_sub(c3.a0, fp_mul(c3.a1, c9)); l.c[9] = c3.a1;
  return l;
}
// The slopes of the Miller loop and the constant terms of its lines depend on Q only: a G2Prepared holds them per step
// (already in the Fp[w] basis), so that later pairings against the same Q do no G2 arithmetic and no inversion.
struct G2Prepared {
  struct Line { Fe4 m_lo, m_hi, c_lo, c_hi; };  // line(P) = -yP + (m_lo + m_hi w^6) xP w + (c_lo + c_hi w^6) w^3
  std::vector<Line> lines;
  bool usable = false;  // false: Q at infinity or a degenerate (vertical) step -> use the generic loop
};
// f <- f * line(R, S)(P), R <- R + S  (S == R doubles).  A vertical line contributes xP - xR' w^2.
// rec != null: the step's line is appended to it (or it is marked unusable on a degenerate step).
inline void miller_step(F12& f, G2& r, const G2& s, const G1& p, G2Prepared* rec = nullptr) {
  if (r.inf || s.inf) {
    if (r.inf) r = s;
    if (rec) rec->usable = false;
    return;
  }
  Fp2 m;
  if (f2_eq(r.x, s.x)) {
    if (!f2_eq(r.y, s.y) || f2_is_zero(r.y)) {
      F12 l = f12_zero();
      const Fe4 c9 = fp_small(9);
      l.c[0] = p.x;
      l.c[2] = fp_neg(fp_sub(r.x.a0, fp_mul(r.x.a1, c9)));
      l.c[8] = fp_neg(r.x.a1);
      f = f12_mul_sparse(l, f);
      r.inf = true;
      if (rec) rec->usable = false;
      return;
    }
    const Fp2 xx = f2_mul(r.x, r.x);
    m = f2_mul(f2_add(f2_add(xx, xx), xx), f2_inv(f2_add(r.y, r.y)));
  } else {
    m = f2_mul(f2_sub(s.y, r.y), f2_inv(f2_sub(s.x, r.x)));
  }
  f = f12_mul_sparse(line_value(m, r, p), f);
  if (rec) {
    const Fe4 c9 = fp_small(9);
    const Fp2 c3 = f2_sub(r.y, f2_mul(m, r.x));
    G2Prepared::Line l;
    l.m_lo = fp_sub(m.a0, fp_mul(m.a1, c9));
    l.m_hi = m.a1;
    l.c_lo = fp_sub(c3.a0, fp_mul(c3.a1, c9));
    l.c_hi = c3.a1;
    rec->lines.push_back(l);
  }
  G2 n;
  n.x = f2_sub(f2_sub(f2_mul(m, m), r.x), s.x);
  n.y = f2_sub(f2_mul(m, f2_sub(r.x, n.x)), r.y);
  n.inf = false;
  r = n;
}
// prod_i f_{6x+2,Q_i}(P_i) with the line corrections of the optimal ate pairing; the squaring of the accumulator is
// shared by all pairs
// record != null (one entry per pair): the lines of every Q are recorded on the way, at no extra G2 work
inline F12 miller_loop_product(const std::vector<std::pair<G1, G2>>& pairs, G2Prepared* record = nullptr) {
  std::vector<const G1*> ps;
  std::vector<const G2*> qs;
  std::vector<G2Prepared*> recs;
  for (size_t k = 0; k < pairs.size(); k++) {
    const auto& pq = pairs[k];
    if (record) {
      record[k].lines.clear();
      record[k].usable = false;
    }
    if (!pq.first.inf && !pq.second.inf) {
      ps.push_back(&pq.first);
      qs.push_back(&pq.second);
      recs.push_back(record ? &record[k] : nullptr);
      if (record) record[k].usable = true;
    }
  }
  F12 f = f12_one();
  std::vector<G2> r;
  for (auto q : qs) r.push_back(*q);
  // 6x+2 has 65 bits: bit 64 is the leading one, then ATE_LOOP_LO from bit 63 down
  for (int i = 63; i >= 0; i--) {
    f = f12_sqr(f);
    for (size_t k = 0; k < r.size(); k++) {
      miller_step(f, r[k], r[k], *ps[k], recs[k]);
      if ((ATE_LOOP_LO >> i) & 1) miller_step(f, r[k], *qs[k], *ps[k], recs[k]);
    }
  }
  const FrobTable& T = frob_table();
  for (size_t k = 0; k < r.size(); k++) {
    const G2& q = *qs[k];
    G2 q1, nq2;  // pi(Q) and -pi^2(Q) on the twist
    q1.x = f2_mul(f2_conj(q.x), T.twist_x);
    q1.y = f2_mul(f2_conj(q.y), T.twist_y);
    q1.inf = false;
    nq2.x = f2_mul(f2_conj(q1.x), T.twist_x);
    nq2.y = f2_neg(f2_mul(f2_conj(q1.y), T.twist_y));
    nq2.inf = false;
    miller_step(f, r[k], q1, *ps[k], recs[k]);
    miller_step(f, r[k], nq2, *ps[k], recs[k]);
  }
  return f;
}
inline F12 miller_loop(const G2& q, const G1& p) { return miller_loop_product({{p, q}}); }
// ---- fixed second argument: the verifier always pairs against the two G2 elements of the SRS
inline G2Prepared g2_prepare(const G2& q) {  // stand-alone preparation: a Miller loop against the generator, recorded
  G2Prepared out;
  if (q.inf) return out;
  miller_loop_product({{g1_generator(), q}}, &out);
  return out;
}
// prod_i f_{6x+2,Q_i}(P_i) for prepared Q_i (same value as miller_loop_product)
inline F12 miller_loop_prepared(const std::vector<std::pair<G1, const G2Prepared*>>& pairs) {
  std::vector<std::pair<G1, const G2Prepared*>> act;
  for (const auto& pq : pairs)
    if (!pq.first.inf) act.push_back(pq);
  F12 f = f12_one();
  std::vector<size_t> pos(act.size(), 0);
  auto apply = [&](size_t k) {
    const G2Prepared::Line& ln = act[k].second->lines[pos[k]++];
    const G1& p = act[k].first;
    F12 l = f12_zero();
    l.c[0] = fp_neg(p.y);
    l.c[1] = fp_mul(ln.m_lo, p.x);
    l.c[7] = fp_mul(ln.m_hi, p.x);
    l.c[3] = ln.c_lo;
    l.c[9] = ln.c_hi;
    f = f12_mul_sparse(l, f);
  };
  for (int i = 63; i >= 0; i--) {
    f = f12_sqr(f);
    for (size_t k = 0; k < act.size(); k++) {
      apply(k);
      if ((ATE_LOOP_LO >> i) & 1) apply(k);
    }
  }
  for (size_t k = 0; k < act.size(); k++) {
    apply(k);
    apply(k);
  }
  return f;
}

// f^((p^12 - 1) / r) = ((f^(p^6 - 1))^(p^2 + 1))^((p^4 - p^2 + 1) / r); the last exponent is exactly
// p^3 + (6x^2 + 1) p^2 - (36x^3 + 18x^2 + 12x - 1) p - (36x^3 + 30x^2 + 18x + 2) for the BN parameter x
inline F12 final_exponentiation(const F12& f) {
  // easy part: g = f^((p^6 - 1)(p^2 + 1)) lies in the cyclotomic subgroup, where inversion is the p^6-Frobenius
  F12 g = f12_mul(f12_conj6(f), f12_inv(f));
  g = f12_mul(f12_frobenius(f12_frobenius(g)), g);
  // hard part through the powers A = g^x, B = g^(x^2), C = g^(x^3) of the BN parameter (63 bits, weight 28):
  //   g^(6x^2+1)            = B^6 g
  //   g^(36x^3+18x^2+12x-1) = (C^2 B)^18 A^12 / g
  //   g^(36x^3+30x^2+18x+2) = (C^2 B)^18 B^12 A^18 g^2
  auto pow_x = [](const F12& a) {
    const uint64_t x = 4965661367192848881ULL;
    F12 r = a;
    for (int bit = 61; bit >= 0; bit--) {
      r = f12_sqr(r);
      if ((x >> bit) & 1) r = f12_mul(r, a);
    }
    return r;
  };
  const F12 A = pow_x(g), B = pow_x(A), C = pow_x(B);
  const F12 B3 = f12_mul(f12_sqr(B), B), B6 = f12_sqr(B3), B12 = f12_sqr(B6);
  const F12 T = f12_mul(f12_sqr(C), B);
  const F12 T2 = f12_sqr(T), T16 = f12_sqr(f12_sqr(f12_sqr(T2))), T18 = f12_mul(T16, T2);
  const F12 A3 = f12_mul(f12_sqr(A), A), A6 = f12_sqr(A3), A12 = f12_sqr(A6), A18 = f12_mul(A12, A6);
  const F12 m2 = f12_mul(B6, g);                                              // g^lambda2
  const F12 m1 = f12_mul(f12_mul(T18, A12), f12_conj6(g));                    // g^-lambda1
  const F12 m0 = f12_mul(f12_mul(T18, B12), f12_mul(A18, f12_sqr(g)));        // g^-lambda0
  const F12 g3 = f12_frobenius(f12_frobenius(f12_frobenius(g)));
  F12 r = f12_mul(f12_conj6(m0), f12_frobenius(f12_conj6(m1)));
  r = f12_mul(r, f12_frobenius(f12_frobenius(m2)));
  return f12_mul(r, g3);
}
// prod_i e(P_i, Q_i) == 1
inline bool pairing_product_is_one(const std::vector<std::pair<G1, G2>>& pairs, G2Prepared* record = nullptr) {
  return f12_eq(final_exponentiation(miller_loop_product(pairs, record)), f12_one());
}

// the same test against prepared second arguments
inline bool pairing_product_is_one_prepared(const std::vector<std::pair<G1, const G2Prepared*>>& pairs) {
  return f12_eq(final_exponentiation(miller_loop_prepared(pairs)), f12_one());
}

}  // namespace ffi
}  // namespace b200zk
