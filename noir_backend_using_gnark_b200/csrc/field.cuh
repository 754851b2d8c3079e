// BN254 fp / fr Montgomery arithmetic for sm_100a.
//
// Replaces (device side) gnark-crypto v0.9.1 ecc/bn254/{fp,fr}/element.go — the arithmetic reached from
// /root/reference/gnark_backend_ffi/backend/plonk/plonk.go:67 (plonk.Prove) through kzg.Commit / fft.Domain.
// Memory layout is gnark's: 4 x u64 little-endian limbs, Montgomery form, R = 2^256, fully reduced (< modulus).
// In registers an element is 8 x u32.  The multiplier is a CIOS interleave written as two carry chains of
// IMAD.WIDE.U32(.X) per step: "even" products a[0,2,4,6]*b_i tile 64-bit slots aligned at bit 0, "odd" products
// a[1,3,5,7]*b_i tile slots aligned at bit 32, so no product ever straddles an accumulator boundary and every
// 32x32 product is exactly one IMAD.WIDE with the carry travelling in a predicate (ptxas fuses
// mad.lo.cc/madc.hi.cc pairs).  136 IMAD per multiplication, ~30 IADD3 on the other pipe.
#pragma once
#include <cstdint>

namespace b200zk {

struct FrParams {
  // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
  static constexpr uint32_t M0 = 0xf0000001u, M1 = 0x43e1f593u, M2 = 0x79b97091u, M3 = 0x2833e848u,
                            M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
  static constexpr uint32_t NINV = 0xefffffffu;  // -r^-1 mod 2^32
  // R mod r
  static __host__ __device__ __forceinline__ uint32_t one(int i) {
    constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return v[i];
  }
  // R^2 mod r
  static __host__ __device__ __forceinline__ uint32_t r2(int i) {
    constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return v[i];
  }
};

struct FpParams {
  // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
  static constexpr uint32_t M0 = 0xd87cfd47u, M1 = 0x3c208c16u, M2 = 0x6871ca8du, M3 = 0x97816a91u,
                            M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
  static constexpr uint32_t NINV = 0xe4866389u;  // -p^-1 mod 2^32
  static __host__ __device__ __forceinline__ uint32_t one(int i) {
    constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return v[i];
  }
  static __host__ __device__ __forceinline__ uint32_t r2(int i) {
    constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return v[i];
  }
};

template <class P>
struct Fe {
  uint32_t l[8];
};
using Fr = Fe<FrParams>;
using Fp = Fe<FpParams>;

// ---------------------------------------------------------------------------------------------------
// carry-chain building blocks (each asm statement is a complete chain: the PTX carry flag never has to
// survive between statements)
// ---------------------------------------------------------------------------------------------------

// acc[0..7] = sum_k a_k * b * 2^(64k)   (4 disjoint 64-bit products)
__device__ __forceinline__ void mulw4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                      uint32_t b) {
  asm("mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;"
      : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
        "=r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// acc[0..7] += sum_k a_k * b * 2^(64k); top += carry out
__device__ __forceinline__ void madw4_top(uint32_t* acc, uint32_t& top, uint32_t a0, uint32_t a1, uint32_t a2,
                                          uint32_t a3, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// acc[0..7] += sum_k a_k * b * 2^(64k)   (caller guarantees no carry out)
__device__ __forceinline__ void madw4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                      uint32_t b) {
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// carry = (x + y) >> 32 ; acc[0..7] += sum_k a_k * b * 2^(64k) + carry   (no carry out)
__device__ __forceinline__ void madw4_cin(uint32_t* acc, uint32_t x, uint32_t y, uint32_t a0, uint32_t a1,
                                          uint32_t a2, uint32_t a3, uint32_t b) {
  uint32_t scratch;
  (void)scratch;
  asm("add.cc.u32 %8, %9, %10;\n\t"
      "madc.lo.cc.u32 %0, %11, %15, %0;\n\t"
      "madc.hi.cc.u32 %1, %11, %15, %1;\n\t"
      "madc.lo.cc.u32 %2, %12, %15, %2;\n\t"
      "madc.hi.cc.u32 %3, %12, %15, %3;\n\t"
      "madc.lo.cc.u32 %4, %13, %15, %4;\n\t"
      "madc.hi.cc.u32 %5, %13, %15, %5;\n\t"
      "madc.lo.cc.u32 %6, %14, %15, %6;\n\t"
      "madc.hi.u32 %7, %14, %15, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
        "+r"(acc[7]), "=r"(scratch)
      : "r"(x), "r"(y), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// r = a + b (8 limbs), returns nothing: inputs < 2^255 so no carry out
__device__ __forceinline__ void add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
        "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}

// r = a - b (8 limbs); returns borrow mask (0xffffffff if a < b else 0)
__device__ __forceinline__ uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
        "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return borrow;
}

template <class P>
__device__ __forceinline__ void load_mod(uint32_t* m) {
  m[0] = P::M0; m[1] = P::M1; m[2] = P::M2; m[3] = P::M3;
  m[4] = P::M4; m[5] = P::M5; m[6] = P::M6; m[7] = P::M7;
}

// if x >= modulus: x -= modulus      (x < 2*modulus)
template <class P>
__device__ __forceinline__ void final_sub(uint32_t* x) {
  uint32_t m[8], t[8];
  load_mod<P>(m);
  uint32_t borrow = sub8(t, x, m);
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

// ---------------------------------------------------------------------------------------------------
// field operations
// ---------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ Fe<P> fe_zero() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0;
  return r;
}

template <class P>
__device__ __forceinline__ Fe<P> fe_one() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
  return r;
}

template <class P>
__device__ __forceinline__ bool fe_is_zero(const Fe<P>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.l[i];
  return o == 0;
}

template <class P>
__device__ __forceinline__ bool fe_eq(const Fe<P>& a, const Fe<P>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i];
  return o == 0;
}

template <class P>
__device__ __forceinline__ Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  add8(r.l, a.l, b.l);
  final_sub<P>(r.l);
  return r;
}

template <class P>
__device__ __forceinline__ Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  uint32_t m[8];
  load_mod<P>(m);
  uint32_t borrow = sub8(r.l, a.l, b.l);
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] &= borrow;
  add8(r.l, r.l, m);
  return r;
}

template <class P>
__device__ __forceinline__ Fe<P> fe_neg(const Fe<P>& a) {
  Fe<P> r;
  uint32_t m[8];
  load_mod<P>(m);
  sub8(r.l, m, a.l);
  bool z = fe_is_zero(a);
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = z ? 0u : r.l[i];
  return r;
}

template <class P>
__device__ __forceinline__ Fe<P> fe_dbl(const Fe<P>& a) {
  return fe_add(a, a);
}

// Montgomery product a*b/R mod modulus, fully reduced: schoolbook CIOS (136 IMAD.WIDE + 8 IMAD).
template <class P>
__device__ __forceinline__ Fe<P> fe_mul_schoolbook(const Fe<P>& a, const Fe<P>& b) {
  uint32_t E[8], O[8], top, c;
  // step 0
  mulw4(E, a.l[0], a.l[2], a.l[4], a.l[6], b.l[0]);
  mulw4(O, a.l[1], a.l[3], a.l[5], a.l[7], b.l[0]);
  top = 0;
  {
    uint32_t q = E[0] * P::NINV;
    madw4_top(E, top, P::M0, P::M2, P::M4, P::M6, q);
    // E[0] == 0 here, no pending limb: plain chain
    madw4(O, P::M1, P::M3, P::M5, P::M7, q);
  }
  // shift by one limb: T/2^32 = O + (E >> 32); E[1] stays pending in c
  c = E[1];
  {
    uint32_t nO[8] = {E[2], E[3], E[4], E[5], E[6], E[7], top, 0u};
#pragma unroll
    for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = nO[k]; }
  }
#pragma unroll
  for (int i = 1; i < 8; i++) {
    top = 0;
    madw4_top(E, top, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
    madw4(O, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
    uint32_t q = (E[0] + c) * P::NINV;
    madw4_top(E, top, P::M0, P::M2, P::M4, P::M6, q);
    // (E[0] + c) == 0 mod 2^32: its carry has weight 2^32 = limb 0 of the odd accumulator
    madw4_cin(O, E[0], c, P::M1, P::M3, P::M5, P::M7, q);
    c = E[1];
    uint32_t nO[8] = {E[2], E[3], E[4], E[5], E[6], E[7], top, 0u};
#pragma unroll
    for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = nO[k]; }
  }
  // result = E + c + 2^32 * O
  Fe<P> r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7])
      : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(c), "r"(O[0]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
  final_sub<P>(r.l);
  return r;
}

// (a*b + c*d)/R mod modulus, fully reduced: the CIOS sweep above with TWO operand rows accumulated per step before the
// reduction row, i.e. one reduction (64 + 8 multiply-adds) for two products instead of two: 200 IMAD.WIDE against 272.
// Bounds (checked limb by limb against big integers before it ran on a GPU): a7, c7, M7 < 2^30, so the odd accumulator's
// top slot takes three products without a carry out; the even accumulator's carries (at most 3 per step) go to `top`;
// the running value stays below 3*modulus*(1 + 2^-32) < 2^256, and the result is below (2*modulus^2/R + modulus) <
// 1.5 * modulus because modulus < R/4: one conditional subtraction.
template <class P>
__device__ __forceinline__ Fe<P> fe_mul2add(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c2, const Fe<P>& d) {
  uint32_t E[8], O[8], top, c;
  mulw4(E, a.l[0], a.l[2], a.l[4], a.l[6], b.l[0]);
  mulw4(O, a.l[1], a.l[3], a.l[5], a.l[7], b.l[0]);
  top = 0;
  madw4_top(E, top, c2.l[0], c2.l[2], c2.l[4], c2.l[6], d.l[0]);
  madw4(O, c2.l[1], c2.l[3], c2.l[5], c2.l[7], d.l[0]);
  {
    uint32_t q = E[0] * P::NINV;
    madw4_top(E, top, P::M0, P::M2, P::M4, P::M6, q);
    madw4(O, P::M1, P::M3, P::M5, P::M7, q);
  }
  c = E[1];
  {
    uint32_t nO[8] = {E[2], E[3], E[4], E[5], E[6], E[7], top, 0u};
#pragma unroll
    for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = nO[k]; }
  }
#pragma unroll
  for (int i = 1; i < 8; i++) {
    top = 0;
    madw4_top(E, top, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
    madw4(O, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
    madw4_top(E, top, c2.l[0], c2.l[2], c2.l[4], c2.l[6], d.l[i]);
    madw4(O, c2.l[1], c2.l[3], c2.l[5], c2.l[7], d.l[i]);
    uint32_t q = (E[0] + c) * P::NINV;
    madw4_top(E, top, P::M0, P::M2, P::M4, P::M6, q);
    madw4_cin(O, E[0], c, P::M1, P::M3, P::M5, P::M7, q);
    c = E[1];
    uint32_t nO[8] = {E[2], E[3], E[4], E[5], E[6], E[7], top, 0u};
#pragma unroll
    for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = nO[k]; }
  }
  Fe<P> r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7])
      : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(c), "r"(O[0]),
        "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
  final_sub<P>(r.l);
  return r;
}

// Montgomery reduction of a 16-limb value T < modulus * 2^256: R <- (R + q*m) / 2^32 + T[8+i] * 2^224, eight times (the
// same even/odd IMAD.WIDE chains as fe_mul with the upper limbs injected one per step); result fully reduced.
template <class P>
__device__ __forceinline__ Fe<P> mont_reduce16(const uint32_t* T) {
  uint32_t Ee[8], Oo[8], top, c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    Ee[k] = T[k];
    Oo[k] = 0;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    top = 0;
    uint32_t q = (Ee[0] + c) * P::NINV;
    madw4_top(Ee, top, P::M0, P::M2, P::M4, P::M6, q);
    madw4_cin(Oo, Ee[0], c, P::M1, P::M3, P::M5, P::M7, q);
    c = Ee[1];
    uint32_t s_lo, s_hi;
    asm("add.cc.u32 %0, %2, %3;\n\t"
        "addc.u32 %1, 0, 0;"
        : "=r"(s_lo), "=r"(s_hi)
        : "r"(top), "r"(T[8 + i]));
    uint32_t nO[8] = {Ee[2], Ee[3], Ee[4], Ee[5], Ee[6], Ee[7], s_lo, s_hi};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      Ee[k] = Oo[k];
      Oo[k] = nO[k];
    }
  }
  Fe<P> r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7])
      : "r"(Ee[0]), "r"(Ee[1]), "r"(Ee[2]), "r"(Ee[3]), "r"(Ee[4]), "r"(Ee[5]), "r"(Ee[6]), "r"(Ee[7]), "r"(c),
        "r"(Oo[0]), "r"(Oo[1]), "r"(Oo[2]), "r"(Oo[3]), "r"(Oo[4]), "r"(Oo[5]), "r"(Oo[6]));
  final_sub<P>(r.l);
  return r;
}

// ---- 4 x 4 limb product (128 x 128 -> 256 bits), 16 IMAD.WIDE: products x_k*y_j with k+j even fill 64-bit slots aligned at
// even limbs (E), those with k+j odd fill slots aligned at odd limbs (O); r = E + (O << 32)
__device__ __forceinline__ void kchain2_2(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
  asm("mad.lo.cc.u32 %0, %6, %8, %0;\n\t"
      "madc.hi.cc.u32 %1, %6, %8, %1;\n\t"
      "madc.lo.cc.u32 %2, %7, %9, %2;\n\t"
      "madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.u32 %5, %5, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void kchain3_1(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1,
                                          uint32_t b2) {
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %8, %11, %2;\n\t"
      "madc.hi.cc.u32 %3, %8, %11, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %9, %12, %5;\n\t"
      "addc.u32 %6, %6, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
      : "r"(a0), "r"(a1), "r"(a2), "r"(b0), "r"(b1), "r"(b2));
}
__device__ __forceinline__ void kchain1_3(uint32_t* acc, uint32_t a0, uint32_t b0) {
  asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.u32 %4, %4, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
      : "r"(a0), "r"(b0));
}

__device__ __forceinline__ void mul4x4(uint32_t* r, const uint32_t* x, const uint32_t* y) {
  auto put = [](uint32_t* dst, uint32_t u, uint32_t v) {
    uint64_t p = (uint64_t)u * v;
    dst[0] = (uint32_t)p;
    dst[1] = (uint32_t)(p >> 32);
  };
  uint32_t E[8], O[7];
  put(E + 0, x[0], y[0]); put(E + 2, x[0], y[2]); put(E + 4, x[1], y[3]); put(E + 6, x[3], y[3]);
  kchain2_2(E + 2, x[1], x[2], y[1], y[2]);
  kchain2_2(E + 2, x[2], x[3], y[0], y[1]);
  put(O + 0, x[0], y[1]); put(O + 2, x[0], y[3]); put(O + 4, x[2], y[3]);
  O[6] = 0;
  kchain3_1(O, x[1], x[1], x[3], y[0], y[2], y[2]);
  kchain1_3(O + 2, x[2], y[1]);
  kchain1_3(O + 2, x[3], y[0]);
  r[0] = E[0];
  asm("add.cc.u32 %0, %7, %14;\n\t"
      "addc.cc.u32 %1, %8, %15;\n\t"
      "addc.cc.u32 %2, %9, %16;\n\t"
      "addc.cc.u32 %3, %10, %17;\n\t"
      "addc.cc.u32 %4, %11, %18;\n\t"
      "addc.cc.u32 %5, %12, %19;\n\t"
      "addc.u32 %6, %13, %20;"
      : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]),
        "r"(O[4]), "r"(O[5]), "r"(O[6]));
}

// r[0..4] = a[0..3] + b[0..3] (r[4] = carry)
__device__ __forceinline__ void add4c(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  asm("add.cc.u32 %0, %5, %9;\n\t"
      "addc.cc.u32 %1, %6, %10;\n\t"
      "addc.cc.u32 %2, %7, %11;\n\t"
      "addc.cc.u32 %3, %8, %12;\n\t"
      "addc.u32 %4, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
}

// Montgomery product with one level of Karatsuba on the 8 x 8 limb product: 3 x 16 IMAD.WIDE for a*b (instead of 64),
// then the reduction sweep (64 + 8).  Same result as fe_mul_schoolbook (fully reduced; the whole GPU suite passes with it
// as the production multiplier).  Measured alternative, not the default: see fe_mul below.
template <class P>
__device__ __forceinline__ Fe<P> fe_mul_k(const Fe<P>& a, const Fe<P>& b) {
  uint32_t z0[8], z2[8], m[9], sa[5], sb[5];
  mul4x4(z0, a.l, b.l);
  mul4x4(z2, a.l + 4, b.l + 4);
  add4c(sa, a.l, a.l + 4);
  add4c(sb, b.l, b.l + 4);
  mul4x4(m, sa, sb);
  // m += (ca ? sb_lo : 0) << 128 + (cb ? sa_lo : 0) << 128 + (ca & cb) << 256
  {
    const uint32_t ma = 0u - sa[4], mb = 0u - sb[4];
    uint32_t u[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      u[i] = sb[i] & ma;
      v[i] = sa[i] & mb;
    }
    m[8] = sa[4] & sb[4];
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]), "+r"(m[8])
        : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]));
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]), "+r"(m[8])
        : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
  }
  // m -= z0 + z2   (the middle term a0*b1 + a1*b0 >= 0, < 2^257)
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const uint32_t* z = pass ? z2 : z0;
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, %8, 0;"
        : "+r"(m[0]), "+r"(m[1]), "+r"(m[2]), "+r"(m[3]), "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]), "+r"(m[8])
        : "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]));
  }
  // T = z0 + (m << 128) + (z2 << 256)
  uint32_t T[16];
#pragma unroll
  for (int i = 0; i < 4; i++) T[i] = z0[i];
  asm("add.cc.u32 %0, %12, %20;\n\t"
      "addc.cc.u32 %1, %13, %21;\n\t"
      "addc.cc.u32 %2, %14, %22;\n\t"
      "addc.cc.u32 %3, %15, %23;\n\t"
      "addc.cc.u32 %4, %16, %24;\n\t"
      "addc.cc.u32 %5, %17, %25;\n\t"
      "addc.cc.u32 %6, %18, %26;\n\t"
      "addc.cc.u32 %7, %19, %27;\n\t"
      "addc.cc.u32 %8, %29, %28;\n\t"
      "addc.cc.u32 %9, %30, 0;\n\t"
      "addc.cc.u32 %10, %31, 0;\n\t"
      "addc.u32 %11, %32, 0;"
      : "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]),
        "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
      : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]),
        "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(m[4]), "r"(m[5]), "r"(m[6]), "r"(m[7]), "r"(m[8]),
        "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
  return mont_reduce16<P>(T);
}

// acc += sum_k a_k*b_k * 2^(64k), carry rippling through 2 more limbs (caller guarantees no carry out)
__device__ __forceinline__ void sq_chain4_2(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  asm("mad.lo.cc.u32 %0, %10, %14, %0;\n\t"
      "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
      "madc.lo.cc.u32 %2, %11, %15, %2;\n\t"
      "madc.hi.cc.u32 %3, %11, %15, %3;\n\t"
      "madc.lo.cc.u32 %4, %12, %16, %4;\n\t"
      "madc.hi.cc.u32 %5, %12, %16, %5;\n\t"
      "madc.lo.cc.u32 %6, %13, %17, %6;\n\t"
      "madc.hi.cc.u32 %7, %13, %17, %7;\n\t"
      "addc.cc.u32 %8, %8, 0;\n\t"
      "addc.u32 %9, %9, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3));
}

// acc += sum_k a_k*b_k * 2^(64k), carry rippling through 4 more limbs (caller guarantees no carry out)
__device__ __forceinline__ void sq_chain2_4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
  asm("mad.lo.cc.u32 %0, %8, %10, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}

// acc += sum_k a_k*b_k * 2^(64k), carry rippling through 2 more limbs (caller guarantees no carry out)
__device__ __forceinline__ void sq_chain5_2(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t b4) {
  asm("mad.lo.cc.u32 %0, %12, %17, %0;\n\t"
      "madc.hi.cc.u32 %1, %12, %17, %1;\n\t"
      "madc.lo.cc.u32 %2, %13, %18, %2;\n\t"
      "madc.hi.cc.u32 %3, %13, %18, %3;\n\t"
      "madc.lo.cc.u32 %4, %14, %19, %4;\n\t"
      "madc.hi.cc.u32 %5, %14, %19, %5;\n\t"
      "madc.lo.cc.u32 %6, %15, %20, %6;\n\t"
      "madc.hi.cc.u32 %7, %15, %20, %7;\n\t"
      "madc.lo.cc.u32 %8, %16, %21, %8;\n\t"
      "madc.hi.cc.u32 %9, %16, %21, %9;\n\t"
      "addc.cc.u32 %10, %10, 0;\n\t"
      "addc.u32 %11, %11, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10]), "+r"(acc[11])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(b4));
}

// acc += sum_k a_k*b_k * 2^(64k), carry rippling through 4 more limbs (caller guarantees no carry out)
__device__ __forceinline__ void sq_chain3_4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1, uint32_t b2) {
  asm("mad.lo.cc.u32 %0, %10, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %10, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
      "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
      "madc.lo.cc.u32 %4, %12, %15, %4;\n\t"
      "madc.hi.cc.u32 %5, %12, %15, %5;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.cc.u32 %7, %7, 0;\n\t"
      "addc.cc.u32 %8, %8, 0;\n\t"
      "addc.u32 %9, %9, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8]), "+r"(acc[9])
      : "r"(a0), "r"(a1), "r"(a2), "r"(b0), "r"(b1), "r"(b2));
}

// acc += sum_k a_k*b_k * 2^(64k), carry rippling through 6 more limbs (caller guarantees no carry out)
__device__ __forceinline__ void sq_chain1_6(uint32_t* acc, uint32_t a0, uint32_t b0) {
  asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
      : "r"(a0), "r"(b0));
}

// Dedicated Montgomery squaring: 28 off-diagonal + 8 diagonal products (instead of 64), then a reduction-only CIOS
// sweep (the same even/odd IMAD.WIDE chains as fe_mul with the upper limbs of the square injected one per step):
// 100 IMAD.WIDE + 8 IMAD instead of 136 + 8.
template <class P>
__device__ __forceinline__ Fe<P> fe_sqr(const Fe<P>& x) {
  const uint32_t* a = x.l;
  uint32_t E[16], O[16];
  auto put = [](uint32_t* dst, uint32_t u, uint32_t v) {
    uint64_t p = (uint64_t)u * v;
    dst[0] = (uint32_t)p;
    dst[1] = (uint32_t)(p >> 32);
  };
  // off-diagonal products a_i*a_j (i<j): even i+j -> E (slot at limb i+j), odd i+j -> O (slot at limb i+j-1, weight 2^32)
  E[0] = E[1] = E[14] = E[15] = 0;
  put(E + 2, a[0], a[2]); put(E + 4, a[0], a[4]); put(E + 6, a[0], a[6]);
  put(E + 8, a[1], a[7]); put(E + 10, a[3], a[7]); put(E + 12, a[5], a[7]);
  sq_chain4_2(E + 4, a[1], a[1], a[2], a[4], a[3], a[5], a[6], a[6]);
  sq_chain2_4(E + 6, a[2], a[3], a[4], a[5]);
  O[14] = O[15] = 0;
  put(O + 0, a[0], a[1]); put(O + 2, a[0], a[3]); put(O + 4, a[0], a[5]); put(O + 6, a[0], a[7]);
  put(O + 8, a[2], a[7]); put(O + 10, a[4], a[7]); put(O + 12, a[6], a[7]);
  sq_chain5_2(O + 2, a[1], a[1], a[1], a[3], a[5], a[2], a[4], a[6], a[6], a[6]);
  sq_chain3_4(O + 4, a[2], a[2], a[4], a[3], a[5], a[5]);
  sq_chain1_6(O + 6, a[3], a[4]);
  // S = E + (O << 32)
  uint32_t S[16];
  S[0] = 0;
  S[1] = O[0];
  asm("add.cc.u32 %0, %14, %28;\n\t"
      "addc.cc.u32 %1, %15, %29;\n\t"
      "addc.cc.u32 %2, %16, %30;\n\t"
      "addc.cc.u32 %3, %17, %31;\n\t"
      "addc.cc.u32 %4, %18, %32;\n\t"
      "addc.cc.u32 %5, %19, %33;\n\t"
      "addc.cc.u32 %6, %20, %34;\n\t"
      "addc.cc.u32 %7, %21, %35;\n\t"
      "addc.cc.u32 %8, %22, %36;\n\t"
      "addc.cc.u32 %9, %23, %37;\n\t"
      "addc.cc.u32 %10, %24, %38;\n\t"
      "addc.cc.u32 %11, %25, %39;\n\t"
      "addc.cc.u32 %12, %26, %40;\n\t"
      "addc.u32 %13, %27, 0;"
      : "=r"(S[2]), "=r"(S[3]), "=r"(S[4]), "=r"(S[5]), "=r"(S[6]), "=r"(S[7]), "=r"(S[8]), "=r"(S[9]), "=r"(S[10]),
        "=r"(S[11]), "=r"(S[12]), "=r"(S[13]), "=r"(S[14]), "=r"(S[15])
      : "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]),
        "r"(E[12]), "r"(E[13]), "r"(O[13]), "r"(E[15]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]),
        "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(E[14]));
  // T = 2*S + diagonal
  uint32_t T[16];
#pragma unroll
  for (int k = 15; k >= 1; k--) T[k] = __funnelshift_l(S[k - 1], S[k], 1);
  T[0] = 0;
  uint32_t D[16];
#pragma unroll
  for (int i = 0; i < 8; i++) put(D + 2 * i, a[i], a[i]);
  asm("add.cc.u32 %0, %0, %16;\n\t"
      "addc.cc.u32 %1, %1, %17;\n\t"
      "addc.cc.u32 %2, %2, %18;\n\t"
      "addc.cc.u32 %3, %3, %19;\n\t"
      "addc.cc.u32 %4, %4, %20;\n\t"
      "addc.cc.u32 %5, %5, %21;\n\t"
      "addc.cc.u32 %6, %6, %22;\n\t"
      "addc.cc.u32 %7, %7, %23;\n\t"
      "addc.cc.u32 %8, %8, %24;\n\t"
      "addc.cc.u32 %9, %9, %25;\n\t"
      "addc.cc.u32 %10, %10, %26;\n\t"
      "addc.cc.u32 %11, %11, %27;\n\t"
      "addc.cc.u32 %12, %12, %28;\n\t"
      "addc.cc.u32 %13, %13, %29;\n\t"
      "addc.cc.u32 %14, %14, %30;\n\t"
      "addc.u32 %15, %15, %31;"
      : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]),
        "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
      : "r"(D[0]), "r"(D[1]), "r"(D[2]), "r"(D[3]), "r"(D[4]), "r"(D[5]), "r"(D[6]), "r"(D[7]), "r"(D[8]), "r"(D[9]),
        "r"(D[10]), "r"(D[11]), "r"(D[12]), "r"(D[13]), "r"(D[14]), "r"(D[15]));
  return mont_reduce16<P>(T);
}

// The product every kernel uses: the schoolbook CIOS form.  The Karatsuba form (-DB200ZK_KARATSUBA_MUL) has 112 instead
// of 127 IMAD.WIDE per product but measured SLOWER on B200 (63.5 vs 67.3 G/s; MSM 2^24 48.4 vs 40.3 ms; NTT 2^24 4.14 vs
// 3.52 ms): ptxas places 22 of its extra moves / carry adds on the multiplier pipe as IMAD.MOV / IMAD.IADD / IMAD.X and
// the hot kernels start to spill — profiles/r02_fieldmul.md.
template <class P>
__device__ __forceinline__ Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
#ifdef B200ZK_KARATSUBA_MUL
  return fe_mul_k(a, b);
#else
  return fe_mul_schoolbook(a, b);
#endif
}

// Montgomery form -> regular integer (multiply by 1)
template <class P>
__device__ __forceinline__ Fe<P> fe_from_mont(const Fe<P>& a) {
  Fe<P> one;
#pragma unroll
  for (int i = 0; i < 8; i++) one.l[i] = (i == 0) ? 1u : 0u;
  return fe_mul(a, one);
}

template <class P>
__device__ __forceinline__ Fe<P> fe_to_mont(const Fe<P>& a) {
  Fe<P> r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.l[i] = P::r2(i);
  return fe_mul(a, r2);
}

// a^(modulus-2): Fermat inversion, 4-bit fixed window.  inv(0) = 0 (gnark's Inverse convention).
template <class P>
__device__ __noinline__ Fe<P> fe_inv(const Fe<P>& a) {
  uint32_t e[8];
  load_mod<P>(e);
  e[0] -= 2;  // both moduli end in ...01 / ...47: no borrow
  Fe<P> tab[16];
  tab[0] = fe_one<P>();
  tab[1] = a;
  for (int i = 2; i < 16; i++) tab[i] = fe_mul(tab[i - 1], a);
  Fe<P> acc = fe_one<P>();
  for (int i = 7; i >= 0; i--) {
    for (int j = 28; j >= 0; j -= 4) {
      acc = fe_sqr(acc);
      acc = fe_sqr(acc);
      acc = fe_sqr(acc);
      acc = fe_sqr(acc);
      uint32_t d = (e[i] >> j) & 15u;
      acc = fe_mul(acc, tab[d]);
    }
  }
  return acc;
}

// a^-1 by the binary extended Euclidean algorithm: for the places where ONE thread normalises a point (the last step of
// every MSM), where the Fermat chain above is 318 dependent multiplications of latency (~0.1 ms) and nothing else runs.
// Data-dependent control flow: do not call it from code whose lanes work on different values.
// Works on the stored integer aR: the result (aR)^-1 is brought to Montgomery form by one multiplication by R^3.
template <class P>
__device__ __noinline__ Fe<P> fe_inv_single(const Fe<P>& a) {
  if (fe_is_zero(a)) return a;
  uint32_t m[8], u[8], v[8], x1[8], x2[8], t[8];
  load_mod<P>(m);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = a.l[i];
    v[i] = m[i];
    x1[i] = i == 0 ? 1u : 0u;
    x2[i] = 0u;
  }
  auto is_one = [](const uint32_t* x) { return x[0] == 1u && (x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) == 0u; };
  auto shr1 = [](uint32_t* x) {
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = __funnelshift_r(x[i], x[i + 1], 1);
    x[7] >>= 1;
  };
  auto halve_mod = [&](uint32_t* x) {  // x / 2 mod m  (x < m < 2^254: x + m does not overflow 8 limbs)
    if (x[0] & 1u) add8(x, x, m);
    shr1(x);
  };
  auto sub_mod = [&](uint32_t* x, const uint32_t* y) {  // x - y mod m
    if (sub8(t, x, y)) add8(t, t, m);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = t[i];
  };
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1u)) {
      shr1(u);
      halve_mod(x1);
    }
    while (!(v[0] & 1u)) {
      shr1(v);
      halve_mod(x2);
    }
    if (sub8(t, u, v) == 0u) {  // u >= v
#pragma unroll
      for (int i = 0; i < 8; i++) u[i] = t[i];
      sub_mod(x1, x2);
    } else {
      sub8(v, v, u);
      sub_mod(x2, x1);
    }
  }
  Fe<P> r, r2;
  const bool first = is_one(u);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.l[i] = first ? x1[i] : x2[i];
    r2.l[i] = P::r2(i);
  }
  return fe_mul(r, fe_mul(r2, r2));  // (aR)^-1 * R^3 / R = a^-1 R
}

// ---------------------------------------------------------------------------------------------------
// global memory access: one element = 32 bytes = two 16-byte vectors
// ---------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ Fe<P> fe_load(const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = q[0], hi = q[1];
  Fe<P> r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}

template <class P>
__device__ __forceinline__ Fe<P> fe_load_ro(const void* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = __ldg(q), hi = __ldg(q + 1);
  Fe<P> r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}

template <class P>
__device__ __forceinline__ void fe_store(void* p, const Fe<P>& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
  q[1] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
}

}  // namespace b200zk
