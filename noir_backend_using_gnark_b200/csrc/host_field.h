// Host-side scalar helpers for the prover orchestration (plonk.cu): the O(1) arithmetic gnark does between kernels —
// Fiat-Shamir challenges (fr.Element.SetBytes of a SHA-256 digest), the handful of challenge-dependent scalars of
// computeLinearizedPolynomial, and the byte encodings bound into the transcript (G1Affine.Marshal, fr.Element.Marshal).
// 4 x u64 Montgomery arithmetic with unsigned __int128; never used for bulk data.
#pragma once
#include <cstdint>
#include <cstring>

namespace b200zk {
namespace host {

typedef unsigned __int128 u128;

struct Fe4 {
  uint64_t l[4];
};

struct Field {
  uint64_t m[4];
  uint64_t ninv;
  Fe4 one;
  Fe4 r2;
};

static const Field HFR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}}};

static const Field HFP = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}}};

inline bool geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
inline void raw_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)t;
    borrow = (t >> 64) & 1;
  }
}
inline bool is_zero(const Fe4& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }

inline Fe4 add(const Field& F, const Fe4& a, const Fe4& b) {
  Fe4 r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (geq(r.l, F.m)) raw_sub(r.l, r.l, F.m);
  return r;
}
inline Fe4 sub(const Field& F, const Fe4& a, const Fe4& b) {
  Fe4 r;
  u128 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.l[i] - b.l[i] - borrow;
    r.l[i] = (uint64_t)d;
    borrow = (d >> 64) & 1;
  }
  if (borrow) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.l[i] + F.m[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}
inline Fe4 neg(const Field& F, const Fe4& a) {
  Fe4 z = {{0, 0, 0, 0}};
  return sub(F, z, a);
}
// Montgomery product (CIOS), fully unrolled so that the compiler keeps the accumulator in registers and, where the field
// is a compile-time constant at the call site (HFR / HFP), folds the modulus into immediates
inline Fe4 mul(const Field& F, const Fe4& a, const Fe4& b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  const uint64_t m0 = F.m[0], m1 = F.m[1], m2 = F.m[2], m3 = F.m[3], ninv = F.ninv;
  const uint64_t a0 = a.l[0], a1 = a.l[1], a2 = a.l[2], a3 = a.l[3];
#define B200ZK_CIOS_STEP(bi)                                   \
  {                                                            \
    u128 c = (u128)a0 * (bi) + t0;                             \
    t0 = (uint64_t)c;                                          \
    c = (c >> 64) + (u128)a1 * (bi) + t1;                      \
    t1 = (uint64_t)c;                                          \
    c = (c >> 64) + (u128)a2 * (bi) + t2;                      \
    t2 = (uint64_t)c;                                          \
    c = (c >> 64) + (u128)a3 * (bi) + t3;                      \
    t3 = (uint64_t)c;                                          \
    c = (c >> 64) + t4;                                        \
    t4 = (uint64_t)c;                                          \
    const uint64_t t5 = (uint64_t)(c >> 64);                   \
    const uint64_t q = t0 * ninv;                              \
    c = ((u128)q * m0 + t0) >> 64;                             \
    c += (u128)q * m1 + t1;                                    \
    t0 = (uint64_t)c;                                          \
    c = (c >> 64) + (u128)q * m2 + t2;                         \
    t1 = (uint64_t)c;                                          \
    c = (c >> 64) + (u128)q * m3 + t3;                         \
    t2 = (uint64_t)c;                                          \
    c = (c >> 64) + t4;                                        \
    t3 = (uint64_t)c;                                          \
    t4 = t5 + (uint64_t)(c >> 64);                             \
  }
  B200ZK_CIOS_STEP(b.l[0])
  B200ZK_CIOS_STEP(b.l[1])
  B200ZK_CIOS_STEP(b.l[2])
  B200ZK_CIOS_STEP(b.l[3])
#undef B200ZK_CIOS_STEP
  uint64_t t[4] = {t0, t1, t2, t3};
  if (t4 || geq(t, F.m)) raw_sub(t, t, F.m);
  Fe4 r;
  memcpy(r.l, t, 32);
  return r;
}
inline Fe4 from_mont(const Field& F, const Fe4& a) {
  Fe4 one = {{1, 0, 0, 0}};
  return mul(F, a, one);
}
inline Fe4 to_mont(const Field& F, const Fe4& a) { return mul(F, a, F.r2); }
inline Fe4 from_u64(const Field& F, uint64_t v) {
  Fe4 t = {{v, 0, 0, 0}};
  return to_mont(F, t);
}
inline Fe4 pow_u64(const Field& F, Fe4 base, uint64_t e) {
  Fe4 acc = F.one;
  while (e) {
    if (e & 1) acc = mul(F, acc, base);
    base = mul(F, base, base);
    e >>= 1;
  }
  return acc;
}
// a^-1 (Montgomery in, Montgomery out; inv(0) = 0 like gnark's Inverse): binary extended Euclid on the stored
// integer aR, which yields a^-1 R^-1, brought back to Montgomery form by one multiplication by R^3.
inline Fe4 inv(const Field& F, const Fe4& a) {
  if (is_zero(a)) return a;
  auto is_one = [](const uint64_t* x) { return x[0] == 1 && (x[1] | x[2] | x[3]) == 0; };
  auto shr1 = [](uint64_t* x, uint64_t top) {
    x[0] = (x[0] >> 1) | (x[1] << 63);
    x[1] = (x[1] >> 1) | (x[2] << 63);
    x[2] = (x[2] >> 1) | (x[3] << 63);
    x[3] = (x[3] >> 1) | (top << 63);
  };
  auto halve_mod = [&](uint64_t* x) {  // x / 2 mod m
    uint64_t top = 0;
    if (x[0] & 1) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) {
        c += (u128)x[i] + F.m[i];
        x[i] = (uint64_t)c;
        c >>= 64;
      }
      top = (uint64_t)c;
    }
    shr1(x, top);
  };
  auto sub_mod = [&](uint64_t* x, const uint64_t* y) {  // x - y mod m, both < m
    if (geq(x, y)) {
      raw_sub(x, x, y);
    } else {
      uint64_t t[4];
      raw_sub(t, y, x);
      raw_sub(x, F.m, t);
    }
  };
  uint64_t u[4], v[4], x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  memcpy(u, a.l, 32);
  memcpy(v, F.m, 32);
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1)) {
      shr1(u, 0);
      halve_mod(x1);
    }
    while (!(v[0] & 1)) {
      shr1(v, 0);
      halve_mod(x2);
    }
    if (geq(u, v)) {
      raw_sub(u, u, v);
      sub_mod(x1, x2);
    } else {
      raw_sub(v, v, u);
      sub_mod(x2, x1);
    }
  }
  Fe4 r;
  memcpy(r.l, is_one(u) ? x1 : x2, 32);
  return mul(F, r, mul(F, F.r2, F.r2));
}

// fr.Element.Marshal(): 32-byte big-endian regular form
inline void marshal(const Field& F, const Fe4& a_mont, uint8_t out[32]) {
  Fe4 r = from_mont(F, a_mont);
  for (int i = 0; i < 4; i++)
    for (int b = 0; b < 8; b++) out[31 - (8 * i + b)] = (uint8_t)(r.l[i] >> (8 * b));
}

// fr.Element.SetBytes(32-byte big-endian): reduce mod r, to Montgomery form
inline Fe4 set_bytes(const Field& F, const uint8_t in[32]) {
  Fe4 v;
  for (int i = 0; i < 4; i++) {
    uint64_t w = 0;
    for (int b = 0; b < 8; b++) w |= (uint64_t)in[31 - (8 * i + b)] << (8 * b);
    v.l[i] = w;
  }
  // value < 2^256 < 6r: subtract the modulus until reduced
  while (geq(v.l, F.m)) raw_sub(v.l, v.l, F.m);
  return to_mont(F, v);
}

// G1Affine.Marshal() = RawBytes(): X || Y big-endian regular form; infinity = 0x40 then zeros
inline void marshal_g1(const uint8_t affine_mont[64], uint8_t out[64]) {
  Fe4 x, y;
  memcpy(x.l, affine_mont, 32);
  memcpy(y.l, affine_mont + 32, 32);
  if (is_zero(x) && is_zero(y)) {
    memset(out, 0, 64);
    out[0] = 0x40;
    return;
  }
  marshal(HFP, x, out);
  marshal(HFP, y, out + 32);
}

// ---- SHA-256 (FIPS 180-4), for the Fiat-Shamir transcript ------------------------------------------------
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len;
  size_t fill;
  Sha256() { reset(); }
  void reset() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, iv, sizeof(iv));
    len = 0;
    fill = 0;
  }
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
        0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
        0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
        0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
        0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
        0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
        0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
        0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
      w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    len += n;
    while (n) {
      size_t take = 64 - fill < n ? 64 - fill : n;
      memcpy(buf + fill, p, take);
      fill += take;
      p += take;
      n -= take;
      if (fill == 64) {
        block(buf);
        fill = 0;
      }
    }
  }
  void finish(uint8_t out[32]) {
    uint64_t bits = len * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill != 56) update(&z, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    for (int i = 0; i < 8; i++) {
      out[4 * i] = (uint8_t)(h[i] >> 24);
      out[4 * i + 1] = (uint8_t)(h[i] >> 16);
      out[4 * i + 2] = (uint8_t)(h[i] >> 8);
      out[4 * i + 3] = (uint8_t)h[i];
    }
  }
};

}  // namespace host
}  // namespace b200zk
