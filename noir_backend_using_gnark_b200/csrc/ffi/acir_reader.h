// Host-side payload handling of the FFI stand-in: a streaming reader for the ACIR JSON the Rust crate
// sends, threaded hex codecs for the felt / key payloads, and the reference's HandleValues + handleArithmeticOpcode
// restated as a reusable "wire plan" so that repeated calls on one circuit only touch the values.
//   [REF] /root/reference/gnark_backend_ffi/acir/acir.go:17-75, acir/opcode/*.go, acir/term/*.go   JSON shapes
//   [REF] backend/common.go:45-76 (HandleValues), backend/plonk/sparse_r1cs.go:27-107 (opcode -> gate), both bug-compatible
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>
#include "../host_field.h"

namespace b200zk {
namespace ffi {

using host::Fe4;
using host::HFR;

[[noreturn]] inline void fatal(const std::string& msg) {  // log.Fatal: message on stderr, exit status 1
  fprintf(stderr, "%s\n", msg.c_str());
  fflush(stderr);
  exit(1);
}

struct Span {
  const char* p;
  size_t n;
  bool is(const char* lit) const { return strlen(lit) == n && memcmp(p, lit, n) == 0; }
};

inline unsigned ffi_threads() {
  static unsigned t = [] {
    unsigned v = std::thread::hardware_concurrency();
    if (const char* e = getenv("B200ZK_FFI_THREADS")) v = (unsigned)atoi(e);
    if (v < 1) v = 1;
    if (v > 32) v = 32;
    return v;
  }();
  return t;
}

// fn(begin, end) over [0, n) on up to ffi_threads() host threads; small ranges run inline
inline void parallel_for(size_t n, size_t min_per_thread, const std::function<void(size_t, size_t)>& fn) {
  size_t parts = n / (min_per_thread ? min_per_thread : 1);
  if (parts > ffi_threads()) parts = ffi_threads();
  if (parts <= 1) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  const size_t step = (n + parts - 1) / parts;
  for (size_t b = step; b < n; b += step) th.emplace_back(fn, b, b + step < n ? b + step : n);
  fn(0, step < n ? step : n);
  for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------------------ hex
inline const int8_t* hex_table() {
  static int8_t t[256];
  static bool init = [] {
    memset(t, -1, sizeof(t));
    for (int c = '0'; c <= '9'; c++) t[c] = (int8_t)(c - '0');
    for (int c = 'a'; c <= 'f'; c++) t[c] = (int8_t)(c - 'a' + 10);
    for (int c = 'A'; c <= 'F'; c++) t[c] = (int8_t)(c - 'A' + 10);
    return true;
  }();
  (void)init;
  return t;
}
// hex.DecodeString into out (s.n / 2 bytes); false on an odd length or a non-hex character
inline bool hex_decode_into(Span s, uint8_t* out) {
  if (s.n % 2) return false;
  const int8_t* T = hex_table();
  int bad = 0;
  for (size_t i = 0; i < s.n / 2; i++) {
    const int a = T[(uint8_t)s.p[2 * i]], b = T[(uint8_t)s.p[2 * i + 1]];
    bad |= a | b;
    out[i] = (uint8_t)((a << 4) | (b & 15));
  }
  return bad >= 0;
}
inline std::vector<uint8_t> hex_decode(Span s) {  // errors are fatal in every caller (helpers.go:16-18, 28-30, ...)
  if (s.n % 2) fatal("encoding/hex: odd length hex string");
  std::vector<uint8_t> out(s.n / 2);
  if (!hex_decode_into(s, out.data())) fatal("encoding/hex: invalid byte");
  return out;
}
inline void hex_encode_into(const uint8_t* in, size_t n, char* out) {
  static const char* d = "0123456789abcdef";
  for (size_t i = 0; i < n; i++) {
    out[2 * i] = d[in[i] >> 4];
    out[2 * i + 1] = d[in[i] & 15];
  }
}
inline std::string hex_encode(const std::vector<uint8_t>& v) {
  std::string s(v.size() * 2, '0');
  hex_encode_into(v.data(), v.size(), &s[0]);
  return s;
}

// fr.Element.SetBytes on big-endian bytes of length <= 32
inline Fe4 felt_from_be(const uint8_t* b, size_t len) {
  uint8_t full[32] = {0};
  if (len > 32) fatal("field element longer than 32 bytes");
  memcpy(full + 32 - len, b, len);
  return host::set_bytes(HFR, full);
}
inline Fe4 felt_from_hex(Span h) {  // backend_helpers.DeserializeFelt (helpers.go:13-23)
  if (h.n % 2) fatal("encoding/hex: odd length hex string");
  if (h.n > 64) fatal("field element longer than 32 bytes");
  uint8_t raw[32];
  if (!hex_decode_into(h, raw)) fatal("encoding/hex: invalid byte");
  return felt_from_be(raw, h.n / 2);
}
// DeserializeFelts (helpers.go:25-33): hex of fr.Vector.MarshalBinary = u32-BE count || 32-byte BE elements.  The
// reference ignores UnmarshalBinary's error, so a short payload yields an empty vector.
inline std::vector<Fe4> felts_from_hex(Span h) {
  if (h.n % 2) fatal("encoding/hex: odd length hex string");
  std::vector<Fe4> out;
  uint8_t head[4];
  if (h.n < 8) {
    std::vector<uint8_t> tmp(h.n / 2 + 1);
    if (!hex_decode_into(h, tmp.data())) fatal("encoding/hex: invalid byte");
    return out;
  }
  if (!hex_decode_into(Span{h.p, 8}, head)) fatal("encoding/hex: invalid byte");
  const size_t n = ((size_t)head[0] << 24) | ((size_t)head[1] << 16) | ((size_t)head[2] << 8) | head[3];
  const size_t have = (h.n - 8) / 64;
  bool ok = true;
  if (have < n) {  // still validates the characters, then behaves like the ignored error
    std::vector<uint8_t> tmp(h.n / 2);
    if (!hex_decode_into(h, tmp.data())) fatal("encoding/hex: invalid byte");
    return out;
  }
  out.resize(n);
  std::atomic<bool> bad(false);
  parallel_for(n, 4096, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) {
      uint8_t raw[32];
      if (!hex_decode_into(Span{h.p + 8 + 64 * i, 64}, raw)) {
        bad.store(true);
        return;
      }
      out[i] = host::set_bytes(HFR, raw);
    }
  });
  if (bad.load()) ok = false;
  if (ok && h.n > 8 + 64 * n) {  // trailing bytes are ignored by UnmarshalBinary but must still be hex
    std::vector<uint8_t> tmp((h.n - 8 - 64 * n) / 2);
    ok = hex_decode_into(Span{h.p + 8 + 64 * n, h.n - 8 - 64 * n}, tmp.data());
  }
  if (!ok) fatal("encoding/hex: invalid byte");
  return out;
}

// hex.DecodeString's acceptance test without producing the bytes (threaded)
inline bool hex_is_valid(Span s) {
  if (s.n % 2) return false;
  const int8_t* T = hex_table();
  std::atomic<bool> bad(false);
  parallel_for(s.n, (size_t)1 << 20, [&](size_t b, size_t e) {
    int acc = 0;
    for (size_t i = b; i < e; i++) acc |= T[(uint8_t)s.p[i]];
    if (acc < 0) bad.store(true);
  });
  return !bad.load();
}

// ------------------------------------------------------------------------------------------------ content hash
// 128-bit non-cryptographic digest of a payload (cache key for parsed circuits and device keys).  A collision could
// only make this process prove with the key of another circuit it has itself submitted; the proof would then fail
// verification — soundness never depends on it.
struct Digest {
  uint64_t a, b;
  bool operator<(const Digest& o) const { return a != o.a ? a < o.a : b < o.b; }
};
inline uint64_t mix64(uint64_t x) {
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  x *= 0xd6e8feb86659fd93ULL;
  x ^= x >> 32;
  return x;
}
inline Digest digest_chunk(const char* p, size_t n) {
  uint64_t a = 0x9E3779B97F4A7C15ULL ^ n, b = 0xC2B2AE3D27D4EB4FULL + n;
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    uint64_t x, y;
    memcpy(&x, p + i, 8);
    memcpy(&y, p + i + 8, 8);
    a = (a ^ x) * 0xff51afd7ed558ccdULL;
    a = (a << 29) | (a >> 35);
    b = (b ^ y) * 0xc4ceb9fe1a85ec53ULL;
    b = (b << 31) | (b >> 33);
    a += b;
  }
  uint64_t tail[2] = {0, 0};
  memcpy(tail, p + i, n - i);
  a = (a ^ tail[0]) * 0xff51afd7ed558ccdULL;
  b = (b ^ tail[1]) * 0xc4ceb9fe1a85ec53ULL;
  return Digest{mix64(a ^ (b >> 7)), mix64(b + (a << 3) + 1)};
}
inline Digest digest(Span s) {
  const size_t chunk = (size_t)1 << 22;
  const size_t nchunks = (s.n + chunk - 1) / chunk;
  if (nchunks <= 1) return digest_chunk(s.p, s.n);
  std::vector<Digest> parts(nchunks);
  parallel_for(nchunks, 1, [&](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) parts[i] = digest_chunk(s.p + i * chunk, i + 1 < nchunks ? chunk : s.n - i * chunk);
  });
  return digest_chunk((const char*)parts.data(), parts.size() * sizeof(Digest));
}

// ------------------------------------------------------------------------------------------------ ACIR
// The circuit as handleArithmeticOpcode reads it: selector columns plus the ACIR witness ids of the three wires
// (set bit k = wire k was assigned from the opcode; an unassigned wire stays 0 without a map lookup).
struct Circuit {
  uint64_t current_witness = 0;
  std::vector<uint32_t> public_inputs;
  std::vector<Fe4> ql, qr, qm, qo, qk;
  std::vector<uint32_t> wa, wb, wc;
  std::vector<uint8_t> set;
  size_t size() const { return ql.size(); }
};

class AcirReader {
 public:
  explicit AcirReader(Span s) : p_(s.p), e_(s.p + s.n) {}

  Circuit read() {  // acir.ACIR.UnmarshalJSON (acir.go:17-75)
    Circuit c;
    bool have_ops = false, have_pub = false, have_cur = false;
    ws();
    if (!eat('{')) fatal("invalid character in ACIR JSON: expected an object");
    ws();
    if (!eat('}')) {
      for (;;) {
        Span key = string_token();
        ws();
        if (!eat(':')) fatal("invalid character in ACIR JSON: expected ':'");
        ws();
        if (key.is("opcodes")) {
          read_opcodes(c);
          have_ops = true;
        } else if (key.is("public_inputs")) {
          if (!eat('[')) fatal("Error: couldn't deserialize public inputs.");
          ws();
          if (!eat(']'))
            for (;;) {
              c.public_inputs.push_back(witness("Error: couldn't deserialize public inputs."));
              ws();
              if (eat(',')) { ws(); continue; }
              if (eat(']')) break;
              fatal("invalid character in ACIR JSON: public_inputs");
            }
          have_pub = true;
        } else if (key.is("current_witness_index")) {
          c.current_witness = witness("Error: couldn't deserialize current witness.");
          have_cur = true;
        } else {
          skip_value();
        }
        ws();
        if (eat(',')) { ws(); continue; }
        if (eat('}')) break;
        fatal("invalid character in ACIR JSON: expected ',' or '}'");
      }
    }
    ws();
    if (p_ != e_) fatal("invalid character in ACIR JSON: trailing data");
    if (!have_ops) fatal("Error: couldn't deserialize opcodes.");
    if (!have_pub) fatal("Error: couldn't deserialize public inputs.");
    if (!have_cur) fatal("Error: couldn't deserialize current witness.");
    return c;
  }

 private:
  const char* p_;
  const char* e_;

  void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) p_++; }
  bool eat(char c) {
    if (p_ < e_ && *p_ == c) { p_++; return true; }
    return false;
  }
  Span string_token() {  // raw contents between the quotes (keys and hex felts never carry escapes)
    if (!eat('"')) fatal("invalid character in ACIR JSON: expected a string");
    const char* st = p_;
    while (p_ < e_ && *p_ != '"') p_ += (*p_ == '\\' && p_ + 1 < e_) ? 2 : 1;
    if (p_ >= e_) fatal("unexpected end of JSON input");
    Span s{st, (size_t)(p_ - st)};
    p_++;
    return s;
  }
  uint32_t witness(const char* err) {  // json number -> float64 -> common.Witness (uint32)
    const char* st = p_;
    uint64_t v = 0;
    while (p_ < e_ && *p_ >= '0' && *p_ <= '9') v = v * 10 + (uint64_t)(*p_++ - '0');
    if (p_ < e_ && (*p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) {
      char* end = nullptr;
      std::string tmp(st, (size_t)((e_ - st) < 64 ? (e_ - st) : 64));
      double d = strtod(tmp.c_str(), &end);
      if (end == tmp.c_str()) fatal(err);
      p_ = st + (end - tmp.c_str());
      return (uint32_t)d;
    }
    if (p_ == st) fatal(err);
    return (uint32_t)v;
  }
  void skip_value() {
    ws();
    if (p_ >= e_) fatal("unexpected end of JSON input");
    if (*p_ == '"') { string_token(); return; }
    if (*p_ == '{' || *p_ == '[') {
      int depth = 0;
      while (p_ < e_) {
        const char ch = *p_;
        if (ch == '"') { string_token(); continue; }
        if (ch == '{' || ch == '[') depth++;
        if (ch == '}' || ch == ']') depth--;
        p_++;
        if (depth == 0) return;
      }
      fatal("unexpected end of JSON input");
    }
    while (p_ < e_ && *p_ != ',' && *p_ != '}' && *p_ != ']' && *p_ != ' ' && *p_ != '\n') p_++;  // number / literal
  }

  void read_opcodes(Circuit& c) {  // opcode.UnmarshalJSON trial order: Arithmetic, BlackBoxFuncCall, Directive (opcode.go:13-36)
    if (!eat('[')) fatal("Error: couldn't deserialize opcodes.");
    ws();
    if (eat(']')) return;
    for (;;) {
      if (!eat('{')) fatal("json: cannot unmarshal opcode");
      bool arith = false, blackbox = false, directive = false;
      ws();
      if (!eat('}'))
        for (;;) {
          Span key = string_token();
          ws();
          if (!eat(':')) fatal("invalid character in ACIR JSON: expected ':'");
          ws();
          if (key.is("Arithmetic") && !arith) {
            read_arithmetic(c);
            arith = true;
          } else if (key.is("BlackBoxFuncCall")) {
            read_blackbox();  // components.go:3-40: black-box functions add no constraints
            blackbox = true;
          } else {
            if (key.is("Directive")) directive = true;  // sparse_r1cs.go:36: skipped
            skip_value();
          }
          ws();
          if (eat(',')) { ws(); continue; }
          if (eat('}')) break;
          fatal("invalid character in ACIR JSON: opcode");
        }
      if (!arith && !blackbox && !directive) fatal("json: cannot unmarshal opcode: not Arithmetic, BlackBoxFuncCall or Directive");
      ws();
      if (eat(',')) { ws(); continue; }
      if (eat(']')) return;
      fatal("invalid character in ACIR JSON: opcodes");
    }
  }

  void read_blackbox() {
    if (!eat('{')) fatal("json: cannot unmarshal BlackBoxFuncCall opcode");
    bool in = false, name = false, out = false;
    ws();
    if (!eat('}'))
      for (;;) {
        Span key = string_token();
        ws();
        if (!eat(':')) fatal("invalid character in ACIR JSON: expected ':'");
        if (key.is("inputs")) in = true;
        if (key.is("name")) name = true;
        if (key.is("outputs")) out = true;
        skip_value();
        ws();
        if (eat(',')) { ws(); continue; }
        if (eat('}')) break;
        fatal("invalid character in ACIR JSON: BlackBoxFuncCall");
      }
    if (!in || !name || !out) fatal("json: cannot unmarshal BlackBoxFuncCall opcode");
  }

  // handleArithmeticOpcode (sparse_r1cs.go:44-107) applied while reading
  void read_arithmetic(Circuit& c) {
    if (!eat('{')) fatal("json: cannot unmarshal Arithmetic opcode");
    const Fe4 zero = {{0, 0, 0, 0}};
    Fe4 ql = zero, qr = zero, qm = zero, qo = zero, qk = zero;
    uint32_t w[3] = {0, 0, 0};
    uint8_t set = 0;
    bool have_mul = false, have_lin = false, have_qc = false, mul_first = false;
    Fe4 lin_c[3];
    uint32_t lin_w[3];
    size_t nlin = 0;
    Fe4 mul_c = zero;
    uint32_t mul_w[2] = {0, 0};
    ws();
    if (!eat('}'))
      for (;;) {
        Span key = string_token();
        ws();
        if (!eat(':')) fatal("invalid character in ACIR JSON: expected ':'");
        ws();
        if (key.is("mul_terms")) {
          if (!eat('[')) fatal("Error: couldn't deserialize mul terms.");
          ws();
          if (!eat(']'))
            for (;;) {
              if (!eat('[')) fatal("Error: couldn't deserialize mul term.");
              ws();
              if (p_ >= e_ || *p_ != '"') fatal("Error: couldn't deserialize coefficient.");
              Fe4 coeff = felt_from_hex(string_token());
              ws();
              if (!eat(',')) fatal("Error: couldn't deserialize multiplicand.");
              ws();
              uint32_t a = witness("Error: couldn't deserialize multiplicand.");
              ws();
              if (!eat(',')) fatal("Error: couldn't deserialize multiplier.");
              ws();
              uint32_t b = witness("Error: couldn't deserialize multiplier.");
              ws();
              if (!eat(']')) fatal("Error: couldn't deserialize mul term.");
              if (!mul_first) {  // only MulTerms[0] is read (sparse_r1cs.go:50)
                mul_first = true;
                mul_c = coeff;
                mul_w[0] = a;
                mul_w[1] = b;
              }
              ws();
              if (eat(',')) { ws(); continue; }
              if (eat(']')) break;
              fatal("invalid character in ACIR JSON: mul_terms");
            }
          have_mul = true;
        } else if (key.is("linear_combinations")) {
          if (!eat('[')) fatal("Error: couldn't deserialize linear combinations.");
          ws();
          if (!eat(']'))
            for (;;) {
              if (!eat('[')) fatal("Error: couldn't deserialize simple term.");
              ws();
              if (p_ >= e_ || *p_ != '"') fatal("Error: couldn't deserialize coefficient.");
              Fe4 coeff = felt_from_hex(string_token());
              ws();
              if (!eat(',')) fatal("Error: couldn't deserialize variable index.");
              ws();
              uint32_t v = witness("Error: couldn't deserialize variable index.");
              ws();
              if (!eat(']')) fatal("Error: couldn't deserialize simple term.");
              if (nlin < 3) {
                lin_c[nlin] = coeff;
                lin_w[nlin] = v;
              }
              nlin++;
              ws();
              if (eat(',')) { ws(); continue; }
              if (eat(']')) break;
              fatal("invalid character in ACIR JSON: linear_combinations");
            }
          have_lin = true;
        } else if (key.is("q_c")) {
          if (p_ >= e_ || *p_ != '"') fatal("Error: couldn't deserialize q_c.");
          qk = felt_from_hex(string_token());
          have_qc = true;
        } else {
          skip_value();
        }
        ws();
        if (eat(',')) { ws(); continue; }
        if (eat('}')) break;
        fatal("invalid character in ACIR JSON: Arithmetic");
      }
    if (!have_mul || !have_lin || !have_qc) fatal("json: cannot unmarshal Arithmetic opcode");
    if (mul_first) {  // qM1 = coeff, qM2 = 1; xa, xb = the two factors (:50-57)
      qm = mul_c;
      w[0] = mul_w[0];
      w[1] = mul_w[1];
      set |= 3;
    }
    if (nlin == 1) {  // :59-63
      qo = lin_c[0]; w[2] = lin_w[0]; set |= 4;
    } else if (nlin == 2) {  // :64-74: overwrites the mul term's wires
      ql = lin_c[0]; w[0] = lin_w[0];
      qr = lin_c[1]; w[1] = lin_w[1];
      set |= 3;
    } else if (nlin == 3) {  // :75-91
      ql = lin_c[0]; w[0] = lin_w[0];
      qr = lin_c[1]; w[1] = lin_w[1];
      qo = lin_c[2]; w[2] = lin_w[2];
      set |= 7;
    }  // more than 3 (or 0) linear terms: none is read
    c.ql.push_back(ql); c.qr.push_back(qr); c.qm.push_back(qm); c.qo.push_back(qo); c.qk.push_back(qk);
    c.wa.push_back(w[0]); c.wb.push_back(w[1]); c.wc.push_back(w[2]);
    c.set.push_back(set);
  }
};

// ------------------------------------------------------------------------------------------------ wire plan
// HandleValues (common.go:45-76) for a value vector of a given LENGTH: which value lands in which public / secret
// slot and which wire every ACIR witness maps to.  Bug-compatible: with P public inputs each non-public value is
// registered P times as a secret (public values P-1 extra times) and the index map keeps the last registration.
struct WirePlan {
  unsigned nb_public = 0, nb_secret = 0;
  std::vector<uint32_t> solution_src;  // wire k (publics then secrets, BuildWitnesses order) <- values[solution_src[k]]
  std::vector<uint32_t> a, b, c;       // per constraint: wire ids
};

inline WirePlan make_plan(const Circuit& cs, size_t nvalues) {
  WirePlan pl;
  std::vector<uint32_t> index_map(nvalues + 1, 0);
  std::vector<uint8_t> present(nvalues + 1, 0);
  std::vector<uint32_t> pub_src, sec_src;
  const auto& pubs = cs.public_inputs;
  for (size_t k = 0; k < nvalues; k++) {
    const uint32_t i = (uint32_t)k + 1;
    for (uint32_t p : pubs)
      if (i == p) {
        index_map[i] = pl.nb_public++;
        present[i] = 1;
        pub_src.push_back((uint32_t)k);
      }
  }
  for (size_t k = 0; k < nvalues; k++) {
    const uint32_t i = (uint32_t)k + 1;
    if (!pubs.empty()) {
      for (uint32_t p : pubs)
        if (i != p) {
          index_map[i] = pl.nb_public + pl.nb_secret++;
          present[i] = 1;
          sec_src.push_back((uint32_t)k);
        }
    } else {
      index_map[i] = pl.nb_public + pl.nb_secret++;
      present[i] = 1;
      sec_src.push_back((uint32_t)k);
    }
  }
  pl.solution_src = pub_src;
  pl.solution_src.insert(pl.solution_src.end(), sec_src.begin(), sec_src.end());
  auto wire = [&](uint32_t w) -> uint32_t { return (w <= nvalues && present[w]) ? index_map[w] : 0u; };  // Go map: missing -> 0
  const size_t m = cs.size();
  pl.a.resize(m); pl.b.resize(m); pl.c.resize(m);
  for (size_t g = 0; g < m; g++) {
    pl.a[g] = (cs.set[g] & 1) ? wire(cs.wa[g]) : 0;
    pl.b[g] = (cs.set[g] & 2) ? wire(cs.wb[g]) : 0;
    pl.c[g] = (cs.set[g] & 4) ? wire(cs.wc[g]) : 0;
  }
  return pl;
}

}  // namespace ffi
}  // namespace b200zk
