// C++ stand-in for gnark_backend_ffi's cgo exports (include/gnark_backend_ffi.h): same symbols, payload encodings and
// error behaviour as /root/reference/gnark_backend_ffi/main.go:24-78, with plonk.Setup / plonk.Prove served by the
// device-resident prover of libb200zk.so.  What is restated here is the reference's own Go glue [REF]:
//   acir/acir.go, acir/opcode/*.go, acir/term/*.go      JSON shapes (trial decoding Arithmetic -> BlackBox -> Directive)
//   backend/common.go:45-76   HandleValues              public / secret partition, bug-compatible
//   backend/plonk/sparse_r1cs.go:44-107                 one SparseR1C per arithmetic opcode, incl. its quirks
//   backend/common.go:78-144  SRS cache file
//   internal/backend/helpers.go                         hex + gnark binary (de)serialisers
// plus gnark v0.8.0's VerifyingKey / ProvingKey / Proof WriteTo layouts and plonk.Verify, as recalled in SURVEY.md
// Appendix C (their sources are not available here).  Verification is CPU work in the reference too.
#include <sys/stat.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../../include/b200zk.h"
#include "../../../include/gnark_backend_ffi.h"
#include "bn254_host.h"

using namespace b200zk;
using namespace b200zk::ffi;

namespace {

#define TRACE(msg) do { if (getenv("B200ZK_FFI_TRACE")) { fprintf(stderr, "[ffi] %s:%d %s\n", __func__, __LINE__, msg); fflush(stderr); } } while (0)

[[noreturn]] void fatal(const std::string& msg) {  // log.Fatal
  fprintf(stderr, "%s\n", msg.c_str());
  fflush(stderr);
  exit(1);
}

char* c_string(const std::string& s) {  // C.CString: malloc'ed copy, never freed by the Rust caller
  char* p = (char*)malloc(s.size() + 1);
  if (!p) fatal("out of memory");
  memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  return p;
}

std::string go_string(GoString s) { return std::string(s.p ? s.p : "", s.n > 0 ? (size_t)s.n : 0); }

// ---------------------------------------------------------------------------------------------- hex
int hex_val(char c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}
std::vector<uint8_t> hex_decode(const std::string& s) {  // hex.DecodeString; errors are fatal in every caller
  if (s.size() % 2) fatal("encoding/hex: odd length hex string");
  std::vector<uint8_t> out(s.size() / 2);
  for (size_t i = 0; i < out.size(); i++) {
    int a = hex_val(s[2 * i]), b = hex_val(s[2 * i + 1]);
    if (a < 0 || b < 0) fatal("encoding/hex: invalid byte");
    out[i] = (uint8_t)(a * 16 + b);
  }
  return out;
}
std::string hex_encode(const std::vector<uint8_t>& v) {
  static const char* d = "0123456789abcdef";
  std::string s(v.size() * 2, '0');
  for (size_t i = 0; i < v.size(); i++) {
    s[2 * i] = d[v[i] >> 4];
    s[2 * i + 1] = d[v[i] & 15];
  }
  return s;
}

// ---------------------------------------------------------------------------------------------- minimal JSON
struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  double num = 0;
  bool b = false;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;
  const Json* get(const std::string& k) const {
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
};
struct JsonParser {
  const std::string& s;
  size_t i = 0;
  bool ok = true;
  explicit JsonParser(const std::string& t) : s(t) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
  Json parse() {
    ws();
    Json j;
    if (i >= s.size()) { ok = false; return j; }
    char c = s[i];
    if (c == '{') {
      j.kind = Json::Obj;
      i++;
      ws();
      if (i < s.size() && s[i] == '}') { i++; return j; }
      while (ok) {
        ws();
        Json k = parse();
        if (k.kind != Json::Str) { ok = false; break; }
        ws();
        if (i >= s.size() || s[i] != ':') { ok = false; break; }
        i++;
        Json v = parse();
        j.obj.emplace_back(k.str, std::move(v));
        ws();
        if (i < s.size() && s[i] == ',') { i++; continue; }
        if (i < s.size() && s[i] == '}') { i++; break; }
        ok = false;
      }
    } else if (c == '[') {
      j.kind = Json::Arr;
      i++;
      ws();
      if (i < s.size() && s[i] == ']') { i++; return j; }
      while (ok) {
        j.arr.push_back(parse());
        ws();
        if (i < s.size() && s[i] == ',') { i++; continue; }
        if (i < s.size() && s[i] == ']') { i++; break; }
        ok = false;
      }
    } else if (c == '"') {
      j.kind = Json::Str;
      i++;
      while (i < s.size() && s[i] != '"') {
        if (s[i] == '\\' && i + 1 < s.size()) {
          char e = s[i + 1];
          j.str.push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e);
          i += 2;
        } else {
          j.str.push_back(s[i++]);
        }
      }
      if (i >= s.size()) ok = false;
      i++;
    } else if (c == 't' && s.compare(i, 4, "true") == 0) {
      j.kind = Json::Bool; j.b = true; i += 4;
    } else if (c == 'f' && s.compare(i, 5, "false") == 0) {
      j.kind = Json::Bool; i += 5;
    } else if (c == 'n' && s.compare(i, 4, "null") == 0) {
      i += 4;
    } else {
      size_t st = i;
      while (i < s.size() && (isdigit((unsigned char)s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.' || s[i] == 'e' || s[i] == 'E')) i++;
      if (i == st) { ok = false; return j; }
      j.kind = Json::Num;
      j.num = strtod(s.substr(st, i - st).c_str(), nullptr);
    }
    return j;
  }
};

// ---------------------------------------------------------------------------------------------- felts
Fe4 felt_from_be32(const uint8_t* b) { return host::set_bytes(HFR, b); }  // fr.Element.SetBytes: reduce, to Montgomery
Fe4 felt_from_hex(const std::string& h) {                                 // backend_helpers.DeserializeFelt
  std::vector<uint8_t> raw = hex_decode(h);
  // SetBytes interprets big-endian of any length; the reference always sends 32 bytes
  uint8_t b[32] = {0};
  if (raw.size() > 32) fatal("felt longer than 32 bytes");
  memcpy(b + 32 - raw.size(), raw.data(), raw.size());
  return felt_from_be32(b);
}
std::vector<Fe4> felts_from_hex(const std::string& h) {                   // DeserializeFelts: fr.Vector.UnmarshalBinary
  std::vector<uint8_t> raw = hex_decode(h);
  std::vector<Fe4> out;
  if (raw.size() < 4) return out;  // the reference ignores UnmarshalBinary's error (helpers.go:31)
  uint32_t n = ((uint32_t)raw[0] << 24) | ((uint32_t)raw[1] << 16) | ((uint32_t)raw[2] << 8) | raw[3];
  if (raw.size() < 4 + (size_t)n * 32) return out;
  out.resize(n);
  for (uint32_t i = 0; i < n; i++) out[i] = felt_from_be32(raw.data() + 4 + 32 * (size_t)i);
  return out;
}
void put_u64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 7; i >= 0; i--) v.push_back((uint8_t)(x >> (8 * i))); }
void put_u32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 3; i >= 0; i--) v.push_back((uint8_t)(x >> (8 * i))); }
void put_fr(std::vector<uint8_t>& v, const Fe4& a) { uint8_t b[32]; host::marshal(HFR, a, b); v.insert(v.end(), b, b + 32); }
void put_g1(std::vector<uint8_t>& v, const G1& p) { uint8_t b[32]; g1_compress(p, b); v.insert(v.end(), b, b + 32); }

// ---------------------------------------------------------------------------------------------- ACIR -> SparseR1CS
struct Gate { Fe4 ql, qr, qm, qo, qk; uint32_t a, b, c; };
struct R1CS {
  unsigned nb_public = 0, nb_secret = 0;
  std::vector<Gate> gates;
  std::vector<Fe4> public_vals, secret_vals;
};

bool as_u32(const Json& j, uint32_t* out) {
  if (j.kind != Json::Num) return false;
  *out = (uint32_t)j.num;
  return true;
}

R1CS build_sparse_r1cs(const std::string& acir_json, const std::vector<Fe4>& values) {
  JsonParser jp(acir_json);
  Json root = jp.parse();
  if (!jp.ok || root.kind != Json::Obj) fatal("invalid character in ACIR JSON");
  const Json* jops = root.get("opcodes");
  const Json* jpub = root.get("public_inputs");
  const Json* jcur = root.get("current_witness_index");
  if (!jops || jops->kind != Json::Arr) fatal("Error: couldn't deserialize opcodes.");
  if (!jpub || jpub->kind != Json::Arr) fatal("Error: couldn't deserialize public inputs.");
  if (!jcur || jcur->kind != Json::Num) fatal("Error: couldn't deserialize current witness.");
  std::vector<uint32_t> pubs;
  for (auto& p : jpub->arr) {
    uint32_t w;
    if (!as_u32(p, &w)) fatal("json: cannot unmarshal public input");
    pubs.push_back(w);
  }
  // HandleValues (common.go:45-76)
  R1CS cs;
  std::map<uint32_t, uint32_t> index_map;
  for (size_t k = 0; k < values.size(); k++) {
    const uint32_t i = (uint32_t)k + 1;
    for (uint32_t p : pubs)
      if (i == p) {
        index_map[i] = cs.nb_public++;
        cs.public_vals.push_back(values[k]);
      }
  }
  for (size_t k = 0; k < values.size(); k++) {
    const uint32_t i = (uint32_t)k + 1;
    if (!pubs.empty()) {
      for (uint32_t p : pubs)
        if (i != p) {
          index_map[i] = cs.nb_public + cs.nb_secret++;
          cs.secret_vals.push_back(values[k]);
        }
    } else {
      index_map[i] = cs.nb_public + cs.nb_secret++;
      cs.secret_vals.push_back(values[k]);
    }
  }
  auto wire = [&](uint32_t w) -> uint32_t {  // Go map lookup: missing key -> 0
    auto it = index_map.find(w);
    return it == index_map.end() ? 0u : it->second;
  };
  const Fe4 zero = {{0, 0, 0, 0}};
  for (auto& op : jops->arr) {
    if (op.kind != Json::Obj) fatal("json: cannot unmarshal opcode");
    if (const Json* ar = op.get("Arithmetic")) {
      const Json* mt = ar->kind == Json::Obj ? ar->get("mul_terms") : nullptr;
      const Json* lc = ar->kind == Json::Obj ? ar->get("linear_combinations") : nullptr;
      const Json* qc = ar->kind == Json::Obj ? ar->get("q_c") : nullptr;
      if (!mt || mt->kind != Json::Arr || !lc || lc->kind != Json::Arr || !qc || qc->kind != Json::Str)
        fatal("json: cannot unmarshal Arithmetic opcode");
      Gate g;
      g.ql = g.qr = g.qm = g.qo = zero;
      g.a = g.b = g.c = 0;
      if (!mt->arr.empty()) {  // only MulTerms[0] is read (sparse_r1cs.go:50)
        const Json& t = mt->arr[0];
        uint32_t w1, w2;
        if (t.kind != Json::Arr || t.arr.size() < 3 || t.arr[0].kind != Json::Str || !as_u32(t.arr[1], &w1) || !as_u32(t.arr[2], &w2))
          fatal("Error: couldn't deserialize mul term.");
        g.qm = felt_from_hex(t.arr[0].str);  // qM1 = coeff, qM2 = 1
        g.a = wire(w1);
        g.b = wire(w2);
      }
      std::vector<std::pair<Fe4, uint32_t>> lin;
      for (auto& t : lc->arr) {
        uint32_t w;
        if (t.kind != Json::Arr || t.arr.size() < 2 || t.arr[0].kind != Json::Str || !as_u32(t.arr[1], &w))
          fatal("Error: couldn't deserialize simple term.");
        lin.emplace_back(felt_from_hex(t.arr[0].str), w);
      }
      if (lin.size() == 1) { g.qo = lin[0].first; g.c = wire(lin[0].second); }
      if (lin.size() == 2) {
        g.ql = lin[0].first; g.a = wire(lin[0].second);  // overwrites the mul term's wires (:69, :73)
        g.qr = lin[1].first; g.b = wire(lin[1].second);
      }
      if (lin.size() == 3) {
        g.ql = lin[0].first; g.a = wire(lin[0].second);
        g.qr = lin[1].first; g.b = wire(lin[1].second);
        g.qo = lin[2].first; g.c = wire(lin[2].second);
      }
      g.qk = felt_from_hex(qc->str);
      cs.gates.push_back(g);
    } else if (const Json* bb = op.get("BlackBoxFuncCall")) {
      const Json* in = bb->kind == Json::Obj ? bb->get("inputs") : nullptr;
      const Json* nm = bb->kind == Json::Obj ? bb->get("name") : nullptr;
      const Json* ou = bb->kind == Json::Obj ? bb->get("outputs") : nullptr;
      if (!in || in->kind != Json::Arr || !nm || nm->kind != Json::Str || !ou || ou->kind != Json::Arr)
        fatal("json: cannot unmarshal BlackBoxFuncCall opcode");
      // components.go:3-40: black-box functions add no constraints
    } else if (op.get("Directive")) {
      // sparse_r1cs.go:36: skipped
    } else {
      fatal("json: cannot unmarshal opcode: not Arithmetic, BlackBoxFuncCall or Directive");
    }
  }
  return cs;
}

// ---------------------------------------------------------------------------------------------- device + SRS state
struct State {
  b200zk_ctx* ctx = nullptr;
  b200zk_bases* bases = nullptr;
  size_t srs_n = 0;
  G2 g2[2];
  std::vector<uint8_t> srs_file;  // u32 count || compressed G1 powers, until they are uploaded
  bool srs_ready = false, bases_ready = false;
  std::map<std::string, b200zk_plonk_pk*> keys;  // per circuit (the reference re-derives the key on every call)
};
State& state() {
  static State s;
  return s;
}
void check(int rc, const char* what) {
  if (rc != 0) {
    State& s = state();
    fatal(std::string(what) + ": " + b200zk_strerror(rc) + " " + (s.ctx ? b200zk_last_cuda_error(s.ctx) : ""));
  }
}
b200zk_ctx* context() {
  State& s = state();
  if (!s.ctx) {
    const char* dev = getenv("B200ZK_DEVICE");
    check(b200zk_init(dev ? atoi(dev) : 0, &s.ctx), "b200zk_init");
  }
  return s.ctx;
}

std::string srs_path() {  // os.UserConfigDir() + "/noir-lang/srs.hex"  (common.go:78-84)
  const char* xdg = getenv("XDG_CONFIG_HOME");
  std::string dir;
  if (xdg && *xdg) dir = xdg;
  else {
    const char* home = getenv("HOME");
    if (!home || !*home) fatal("neither $XDG_CONFIG_HOME nor $HOME are defined");
    dir = std::string(home) + "/.config";
  }
  return dir + "/noir-lang/srs.hex";
}

Fe4 random_fr(FILE* ur) {  // rand.Int(rand.Reader, r) / fr.SetRandom: uniform below r
  for (;;) {
    Fe4 v;
    if (fread(v.l, 1, 32, ur) != 32) fatal("cannot read /dev/urandom");
    v.l[3] &= 0x3fffffffffffffffULL;
    if (!host::geq(v.l, HFR.m)) return v;
  }
}

bool try_load_srs(State& s) {  // LoadSRS (common.go:86-105); the G1 part goes to the device only when a commitment needs it
  FILE* f = fopen(srs_path().c_str(), "rb");
  if (!f) return false;
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
  fclose(f);
  while (!text.empty() && (text.back() == '\n' || text.back() == ' ')) text.pop_back();
  if (text.size() % 2) return false;
  std::vector<uint8_t> raw(text.size() / 2);
  for (size_t i = 0; i < raw.size(); i++) {
    int a = hex_val(text[2 * i]), b = hex_val(text[2 * i + 1]);
    if (a < 0 || b < 0) return false;
    raw[i] = (uint8_t)(a * 16 + b);
  }
  if (raw.size() < 4 + 128) return false;
  size_t n = ((size_t)raw[0] << 24) | ((size_t)raw[1] << 16) | ((size_t)raw[2] << 8) | raw[3];
  if (raw.size() != 4 + 32 * n + 128 || n == 0) return false;
  if (!g2_decompress(raw.data() + 4 + 32 * n, &s.g2[0]) || !g2_decompress(raw.data() + 4 + 32 * n + 64, &s.g2[1])) return false;
  raw.resize(4 + 32 * n);
  s.srs_file = std::move(raw);
  s.srs_n = n;
  return true;
}

void save_srs(State& s) {  // SaveSRS (common.go:107-125); unlike the reference the directory is created first
  std::string path = srs_path();
  std::string dir = path.substr(0, path.rfind('/'));
  std::string parent = dir.substr(0, dir.rfind('/'));
  mkdir(parent.c_str(), 0755);
  mkdir(dir.c_str(), 0755);
  std::vector<uint8_t> raw;
  put_u32(raw, (uint32_t)s.srs_n);
  raw.resize(4 + 32 * s.srs_n);
  check(b200zk_bases_download_compressed(context(), s.bases, 0, s.srs_n, raw.data() + 4), "b200zk_bases_download_compressed");
  uint8_t g[64];
  g2_compress(s.g2[0], g); raw.insert(raw.end(), g, g + 64);
  g2_compress(s.g2[1], g); raw.insert(raw.end(), g, g + 64);
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return;  // the reference ignores SaveSRS's error too (common.go:141)
  std::string hex = hex_encode(raw);
  fwrite(hex.data(), 1, hex.size(), f);
  fclose(f);
}

void ensure_srs() {  // TryLoadSRS (common.go:127-144): the G2 pair (enough to verify)
  State& s = state();
  if (s.srs_ready) return;
  if (!try_load_srs(s)) {
    size_t n = 1000000;  // common.go:137
    if (const char* e = getenv("B200ZK_SRS_SIZE")) n = (size_t)strtoull(e, nullptr, 10);
    FILE* ur = fopen("/dev/urandom", "rb");
    if (!ur) fatal("cannot open /dev/urandom");
    Fe4 alpha_regular = random_fr(ur);
    fclose(ur);
    Fe4 alpha = host::to_mont(HFR, alpha_regular);
    check(b200zk_srs_generate(context(), alpha.l, 0, n, &s.bases), "b200zk_srs_generate");
    s.srs_n = n;
    s.g2[0] = g2_generator();
    s.g2[1] = g2_mul(s.g2[0], alpha);
    save_srs(s);
  }
  s.srs_ready = true;
}

void ensure_srs_bases() {  // + the G1 powers resident in HBM, with the window table of the static bases
  ensure_srs();
  TRACE("srs g2 ready");
  State& s = state();
  if (s.bases_ready) return;
  if (!s.bases) {
    int rc = b200zk_bases_upload_compressed(context(), s.srs_file.data() + 4, s.srs_n, &s.bases);
    if (rc == B200ZK_ERR_BAD_ARG) fatal("invalid point in the SRS file");
    check(rc, "b200zk_bases_upload_compressed");
    std::vector<uint8_t>().swap(s.srs_file);
  }
  TRACE("uploaded");
  b200zk_bases_precompute(context(), s.bases, 0);  // best effort: commitments use classic windows if memory is short
  s.bases_ready = true;
}

// ---------------------------------------------------------------------------------------------- keys
b200zk_plonk_pk* setup_key(const R1CS& cs, const std::string& cache_key) {
  State& s = state();
  auto it = s.keys.find(cache_key);
  if (it != s.keys.end()) return it->second;
  ensure_srs_bases();
  TRACE("bases ready");
  const size_t m = cs.gates.size();
  std::vector<Fe4> ql(m ? m : 1), qr(m ? m : 1), qm(m ? m : 1), qo(m ? m : 1), qk(m ? m : 1);
  std::vector<uint32_t> a(m ? m : 1), b(m ? m : 1), c(m ? m : 1);
  for (size_t i = 0; i < m; i++) {
    ql[i] = cs.gates[i].ql; qr[i] = cs.gates[i].qr; qm[i] = cs.gates[i].qm; qo[i] = cs.gates[i].qo; qk[i] = cs.gates[i].qk;
    a[i] = cs.gates[i].a; b[i] = cs.gates[i].b; c[i] = cs.gates[i].c;
  }
  b200zk_plonk_pk* pk = nullptr;
  int rc = b200zk_plonk_setup_r1cs(context(), s.bases, cs.nb_public, cs.nb_secret, m, ql.data(), qr.data(), qm.data(),
                                   qo.data(), qk.data(), a.data(), b.data(), c.data(), &pk);
  if (rc == B200ZK_ERR_BAD_ARG) fatal("kzg: the SRS is too small for this circuit (set B200ZK_SRS_SIZE)");
  check(rc, "plonk.Setup");
  s.keys[cache_key] = pk;
  return pk;
}

struct Sizes { uint64_t n, n_big; };
Sizes domain_sizes(const R1CS& cs) {
  const size_t sys = cs.gates.size() + cs.nb_public;
  uint64_t n = 2;
  while (n < sys) n <<= 1;
  uint64_t nb = 1;
  while (nb < (sys < 6 ? 8 : 4) * sys) nb <<= 1;
  if (nb < 4 * n) nb = 4 * n;
  return {n, nb};
}
Fe4 fr_root_of_unity(uint64_t n) {  // fft.NewDomain(n).Generator
  const uint32_t root[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
  Fe4 w;
  memcpy(w.l, root, 32);
  unsigned lg = 0;
  while (((uint64_t)1 << lg) < n) lg++;
  for (unsigned k = lg; k < 28; k++) w = host::mul(HFR, w, w);
  return w;
}

// VerifyingKey.WriteTo (gnark v0.8.0, recalled): Size | SizeInv | Generator | NbPublicVariables | S[0..2] | Ql Qr Qm Qo Qk
std::vector<uint8_t> serialize_vk(const R1CS& cs, const uint8_t vk_points[8 * 64]) {
  Sizes sz = domain_sizes(cs);
  std::vector<uint8_t> v;
  put_u64(v, sz.n);
  put_fr(v, host::inv(HFR, host::from_u64(HFR, sz.n)));
  put_fr(v, fr_root_of_unity(sz.n));
  put_u64(v, cs.nb_public);
  for (int i = 0; i < 8; i++) put_g1(v, g1_from_image(vk_points + 64 * i));
  return v;
}
void put_domain(std::vector<uint8_t>& v, uint64_t n) {  // fft.Domain.WriteTo
  Fe4 g = fr_root_of_unity(n), five = host::from_u64(HFR, 5);
  put_u64(v, n);
  put_fr(v, host::inv(HFR, host::from_u64(HFR, n)));
  put_fr(v, g);
  put_fr(v, host::inv(HFR, g));
  put_fr(v, five);
  put_fr(v, host::inv(HFR, five));
}

struct ParsedVk {
  uint64_t size = 0, nb_public = 0;
  Fe4 size_inv, generator;
  G1 S[3], Ql, Qr, Qm, Qo, Qk;
};
ParsedVk parse_vk(const std::vector<uint8_t>& v) {  // VerifyingKey.ReadFrom
  if (v.size() < 8 + 32 + 32 + 8 + 8 * 32) fatal("unexpected EOF reading the verifying key");
  ParsedVk k;
  auto u64 = [&](size_t off) { uint64_t x = 0; for (int i = 0; i < 8; i++) x = (x << 8) | v[off + i]; return x; };
  k.size = u64(0);
  k.size_inv = felt_from_be32(v.data() + 8);
  k.generator = felt_from_be32(v.data() + 40);
  k.nb_public = u64(72);
  G1* pts[8] = {&k.S[0], &k.S[1], &k.S[2], &k.Ql, &k.Qr, &k.Qm, &k.Qo, &k.Qk};
  for (int i = 0; i < 8; i++)
    if (!g1_decompress(v.data() + 80 + 32 * i, pts[i])) fatal("invalid point in the verifying key");
  return k;
}

// ---------------------------------------------------------------------------------------------- proof
struct ParsedProof {
  G1 LRO[3], Z, H[3], batched_H, zshift_H;
  Fe4 claimed[7], zshift_value;
};
std::vector<uint8_t> serialize_proof(const uint8_t blob[832]) {  // Proof.WriteTo (548 bytes)
  std::vector<uint8_t> v;
  for (int i = 0; i < 8; i++) put_g1(v, g1_from_image(blob + 64 * i));  // LRO[3], Z, H[3], BatchedProof.H
  put_u32(v, 7);
  for (int i = 0; i < 7; i++) { Fe4 c; memcpy(c.l, blob + 576 + 32 * i, 32); put_fr(v, c); }
  put_g1(v, g1_from_image(blob + 64 * 8));
  Fe4 zs; memcpy(zs.l, blob + 576 + 224, 32);
  put_fr(v, zs);
  return v;
}
ParsedProof parse_proof(const std::vector<uint8_t>& v) {  // Proof.ReadFrom
  if (v.size() < 8 * 32 + 4) fatal("unexpected EOF reading the proof");
  ParsedProof p;
  G1* pts[8] = {&p.LRO[0], &p.LRO[1], &p.LRO[2], &p.Z, &p.H[0], &p.H[1], &p.H[2], &p.batched_H};
  for (int i = 0; i < 8; i++)
    if (!g1_decompress(v.data() + 32 * i, pts[i])) fatal("invalid point in the proof");
  uint32_t k = ((uint32_t)v[256] << 24) | ((uint32_t)v[257] << 16) | ((uint32_t)v[258] << 8) | v[259];
  if (k != 7 || v.size() < 260 + 32 * 7 + 64) fatal("malformed proof");
  for (int i = 0; i < 7; i++) p.claimed[i] = felt_from_be32(v.data() + 260 + 32 * i);
  if (!g1_decompress(v.data() + 484, &p.zshift_H)) fatal("invalid point in the proof");
  p.zshift_value = felt_from_be32(v.data() + 516);
  return p;
}

// ---------------------------------------------------------------------------------------------- verifier
struct Transcript {
  host::Sha256 h;
  bool have_prev = false;
  uint8_t prev[32];
  void begin(const char* name) { h.reset(); h.update(name, strlen(name)); if (have_prev) h.update(prev, 32); }
  void point(const G1& p) { uint8_t img[64], b[64]; g1_to_image(p, img); host::marshal_g1(img, b); h.update(b, 64); }
  void fr(const Fe4& v) { uint8_t b[32]; host::marshal(HFR, v, b); h.update(b, 32); }
  Fe4 finish() { h.finish(prev); have_prev = true; return host::set_bytes(HFR, prev); }
};

bool kzg_verify(const G1& commitment, const Fe4& z, const Fe4& v, const G1& Hq, const G2 g2[2]) {
  // e(C - v G1 + z H, G2) * e(-H, alpha G2) == 1
  G1J acc = g1j_inf();
  acc = g1j_add_affine(acc, commitment);
  acc = g1j_add_affine(acc, g1_neg(g1_mul(g1_generator(), v)));
  acc = g1j_add_affine(acc, g1_mul(Hq, z));
  return pairing_product_is_one({{g1j_to_affine(acc), g2[0]}, {g1_neg(Hq), g2[1]}});
}

bool plonk_verify(const ParsedProof& pr, const ParsedVk& vk, const std::vector<Fe4>& pub, const G2 g2[2]) {  // plonk.Verify
  auto M = [](const Fe4& a, const Fe4& b) { return host::mul(HFR, a, b); };
  auto A = [](const Fe4& a, const Fe4& b) { return host::add(HFR, a, b); };
  auto S = [](const Fe4& a, const Fe4& b) { return host::sub(HFR, a, b); };
  Transcript fs;
  fs.begin("gamma");
  for (int i = 0; i < 3; i++) fs.point(vk.S[i]);
  fs.point(vk.Ql); fs.point(vk.Qr); fs.point(vk.Qm); fs.point(vk.Qo); fs.point(vk.Qk);
  for (auto& w : pub) fs.fr(w);
  for (int i = 0; i < 3; i++) fs.point(pr.LRO[i]);
  const Fe4 gamma = fs.finish();
  fs.begin("beta");
  const Fe4 beta = fs.finish();
  fs.begin("alpha");
  fs.point(pr.Z);
  const Fe4 alpha = fs.finish();
  fs.begin("zeta");
  for (int i = 0; i < 3; i++) fs.point(pr.H[i]);
  const Fe4 zeta = fs.finish();
  const Fe4 one = HFR.one, u = host::from_u64(HFR, 5), uu = M(u, u);
  const Fe4 zeta_n = host::pow_u64(HFR, zeta, vk.size);
  const Fe4 zz = S(zeta_n, one);
  Fe4 pi = {{0, 0, 0, 0}};
  Fe4 wi = one;
  for (size_t i = 0; i < pub.size(); i++) {  // PI(zeta) = sum_i L_i(zeta) w_i
    Fe4 li = M(M(M(wi, zz), vk.size_inv), host::inv(HFR, S(zeta, wi)));
    pi = A(pi, M(li, pub[i]));
    wi = M(wi, vk.generator);
  }
  const Fe4 lag1 = M(M(zz, vk.size_inv), host::inv(HFR, S(zeta, one)));
  const Fe4 &q = pr.claimed[0], &lin_z = pr.claimed[1], &l = pr.claimed[2], &r = pr.claimed[3], &o = pr.claimed[4],
            &s1 = pr.claimed[5], &s2 = pr.claimed[6], &zu = pr.zshift_value;
  const Fe4 f1 = A(A(M(s1, beta), l), gamma), f2 = A(A(M(s2, beta), r), gamma);
  Fe4 t1 = M(M(M(M(f1, f2), A(o, gamma)), alpha), zu);
  Fe4 lhs = S(A(A(lin_z, pi), t1), M(M(alpha, alpha), lag1));
  if (!fe_eq(lhs, M(q, zz))) return false;
  const Fe4 zpm = host::pow_u64(HFR, zeta, vk.size + 2);
  G1J fh = g1_mul_j(pr.H[2], zpm);
  fh = g1j_add_affine(fh, pr.H[1]);
  fh = g1_mul_j(g1j_to_affine(fh), zpm);
  fh = g1j_add_affine(fh, pr.H[0]);
  const G1 folded_h = g1j_to_affine(fh);
  const Fe4 c_s3 = M(M(M(M(f1, f2), beta), alpha), zu);
  Fe4 c_z = M(M(A(A(M(beta, zeta), l), gamma), A(A(M(M(beta, u), zeta), r), gamma)), A(A(M(M(beta, uu), zeta), o), gamma));
  c_z = A(host::neg(HFR, M(c_z, alpha)), M(M(alpha, alpha), lag1));
  const G1* lp[7] = {&vk.Ql, &vk.Qr, &vk.Qm, &vk.Qo, &vk.Qk, &vk.S[2], &pr.Z};
  const Fe4 ls[7] = {l, r, M(l, r), o, one, c_s3, c_z};
  G1J lin = g1j_inf();
  for (int i = 0; i < 7; i++) lin = g1j_add_affine(lin, g1_mul(*lp[i], ls[i]));
  const G1 lin_digest = g1j_to_affine(lin);
  const G1 digests[7] = {folded_h, lin_digest, pr.LRO[0], pr.LRO[1], pr.LRO[2], vk.S[0], vk.S[1]};
  Transcript kz;
  kz.begin("gamma");
  kz.fr(zeta);
  for (int i = 0; i < 7; i++) kz.point(digests[i]);
  const Fe4 gk = kz.finish();
  G1J fd = g1j_inf();
  Fe4 fe = {{0, 0, 0, 0}}, acc = one;
  for (int i = 0; i < 7; i++) {
    fd = g1j_add_affine(fd, g1_mul(digests[i], acc));
    fe = A(fe, M(pr.claimed[i], acc));
    acc = M(acc, gk);
  }
  if (!kzg_verify(g1j_to_affine(fd), zeta, fe, pr.batched_H, g2)) return false;
  return kzg_verify(pr.Z, M(zeta, vk.generator), zu, pr.zshift_H, g2);
}

std::string circuit_key(const std::string& acir, size_t nvalues) { return acir + "#" + std::to_string(nvalues); }

}  // namespace

// ================================================================================================ exports
extern "C" {

struct PlonkPreprocess_return PlonkPreprocess(GoString acirJSON, GoString encodedRandomValues) {  // main.go:58-78
  const std::string acir = go_string(acirJSON);
  // the Rust side sends a JSON-quoted hex string (plonk/mod.rs:197-203); main.go:66-72 un-quotes it
  const std::string quoted = go_string(encodedRandomValues);
  JsonParser jp(quoted);
  Json q = jp.parse();
  if (!jp.ok || q.kind != Json::Str) fatal("json: cannot unmarshal encoded values into Go value of type string");
  TRACE("parsed values");
  std::vector<Fe4> values = felts_from_hex(q.str);
  R1CS cs = build_sparse_r1cs(acir, values);
  TRACE("built r1cs");
  b200zk_plonk_pk* pk = setup_key(cs, circuit_key(acir, values.size()));
  TRACE("setup done");
  uint8_t vkp[8 * 64];
  check(b200zk_plonk_vk(context(), pk, vkp), "b200zk_plonk_vk");
  std::vector<uint8_t> vk = serialize_vk(cs, vkp);
  // ProvingKey.WriteTo (gnark v0.8.0, recalled): Vk | Domain[0] | Domain[1] | Ql Qr Qm Qo CQk LQk S1 S2 S3 | Permutation
  Sizes sz = domain_sizes(cs);
  std::vector<uint8_t> pkb = vk;
  put_domain(pkb, sz.n);
  put_domain(pkb, sz.n_big);
  std::vector<Fe4> poly(sz.n);
  for (int which = 0; which < 9; which++) {
    check(b200zk_plonk_pk_poly(context(), pk, which, poly.data()), "b200zk_plonk_pk_poly");
    put_u32(pkb, (uint32_t)sz.n);
    for (auto& c : poly) put_fr(pkb, c);
  }
  {
    // Permutation: gnark's buildPermutation over the same row layout
    std::vector<uint32_t> lro(3 * sz.n, 0);
    for (unsigned i = 0; i < cs.nb_public; i++) lro[i] = i;
    for (size_t i = 0; i < cs.gates.size(); i++) {
      lro[cs.nb_public + i] = cs.gates[i].a;
      lro[sz.n + cs.nb_public + i] = cs.gates[i].b;
      lro[2 * sz.n + cs.nb_public + i] = cs.gates[i].c;
    }
    const size_t nw = cs.nb_public + cs.nb_secret ? cs.nb_public + cs.nb_secret : 1;
    std::vector<int64_t> perm(3 * sz.n, -1), cycle(nw, -1);
    for (size_t i = 0; i < 3 * sz.n; i++) {
      if (cycle[lro[i]] != -1) perm[i] = cycle[lro[i]];
      cycle[lro[i]] = (int64_t)i;
    }
    for (size_t i = 0; i < 3 * sz.n; i++)
      if (perm[i] == -1) perm[i] = cycle[lro[i]];
    put_u32(pkb, (uint32_t)(3 * sz.n));
    for (auto v : perm) put_u64(pkb, (uint64_t)v);
  }
  struct PlonkPreprocess_return r;
  r.r0 = c_string(hex_encode(pkb));
  r.r1 = c_string(hex_encode(vk));
  return r;
}

char* PlonkProveWithPK(GoString acirJSON, GoString encodedValues, GoString encodedProvingKey) {  // main.go:24-37
  const std::string acir = go_string(acirJSON);
  std::vector<Fe4> values = felts_from_hex(go_string(encodedValues));
  // DeserializeProvingKey: the key is re-derived from the circuit and the cached SRS (the reference itself rebuilds the
  // constraint system and recomputes the key's derived data on every call: helpers.go:49-60, plonk.go:54); the payload
  // is only checked for well-formedness
  (void)hex_decode(go_string(encodedProvingKey));
  R1CS cs = build_sparse_r1cs(acir, values);
  b200zk_plonk_pk* pk = setup_key(cs, circuit_key(acir, values.size()));
  // BuildWitnesses (common.go:22-43): publics then secrets = wire order
  std::vector<Fe4> sol = cs.public_vals;
  sol.insert(sol.end(), cs.secret_vals.begin(), cs.secret_vals.end());
  if (sol.empty()) sol.push_back(Fe4{{0, 0, 0, 0}});
  // spr.Solve: every wire is an input here, so solving is checking (plonk.Prove fails -> log.Fatal, plonk.go:67-70)
  for (size_t k = 0; k < cs.gates.size(); k++) {
    const Gate& g = cs.gates[k];
    auto M = [](const Fe4& a, const Fe4& b) { return host::mul(HFR, a, b); };
    auto A = [](const Fe4& a, const Fe4& b) { return host::add(HFR, a, b); };
    Fe4 v = A(A(A(M(g.ql, sol[g.a]), M(g.qr, sol[g.b])), A(M(g.qo, sol[g.c]), M(M(g.qm, sol[g.a]), sol[g.b]))), g.qk);
    if (!host::is_zero(v)) fatal("constraint #" + std::to_string(k) + " is not satisfied");
  }
  Fe4 blinding[9];
  if (const char* seed = getenv("B200ZK_BLINDING_SEED")) {
    uint64_t st = strtoull(seed, nullptr, 0);
    for (int i = 0; i < 9;) {
      Fe4 v;
      for (int k = 0; k < 4; k++) {
        st += 0x9E3779B97F4A7C15ULL;
        uint64_t z = st;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        v.l[k] = z ^ (z >> 31);
      }
      v.l[3] &= 0x3fffffffffffffffULL;
      if (!host::geq(v.l, HFR.m)) blinding[i++] = v;
    }
  } else {
    FILE* ur = fopen("/dev/urandom", "rb");
    if (!ur) fatal("cannot open /dev/urandom");
    for (int i = 0; i < 9; i++) blinding[i] = random_fr(ur);  // limbs are the Montgomery form (fr.SetRandom)
    fclose(ur);
  }
  uint8_t blob[832];
  check(b200zk_plonk_prove(context(), pk, sol.data(), blinding, blob), "plonk.Prove");
  return c_string(hex_encode(serialize_proof(blob)));
}

uint8_t PlonkVerifyWithMeta(GoString, GoString, GoString) { return 0; }  // main.go:39-42

uint8_t PlonkVerifyWithVK(GoString acirJSON, GoString encodedProof, GoString encodedPublicInputs,
                          GoString encodedVerifyingKey) {  // main.go:44-56, plonk.go:29-51
  const std::string acir = go_string(acirJSON);
  ParsedProof proof = parse_proof(hex_decode(go_string(encodedProof)));
  std::vector<Fe4> values = felts_from_hex(go_string(encodedPublicInputs));
  ParsedVk vk = parse_vk(hex_decode(go_string(encodedVerifyingKey)));
  R1CS cs = build_sparse_r1cs(acir, values);  // only to learn which values are public (plonk.go:30)
  ensure_srs();                               // vk.InitKZG(srs): the G2 elements live in the SRS file
  return plonk_verify(proof, vk, cs.public_vals, state().g2) ? 1 : 0;
}

}  // extern "C"
