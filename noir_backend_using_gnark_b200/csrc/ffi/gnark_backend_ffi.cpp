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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "../../../include/b200zk.h"
#include "../../../include/gnark_backend_ffi.h"
#include "acir_reader.h"
#include "bn254_host.h"

using namespace b200zk;
using namespace b200zk::ffi;

namespace {

// B200ZK_FFI_TRACE=1: wall-clock trace of the stages of each call on stderr
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct Trace {
  bool on;
  double t0, last;
  const char* fn;
  explicit Trace(const char* f) : on(false), t0(now_ms()), last(t0), fn(f) {
    const char* e = getenv("B200ZK_FFI_TRACE");
    on = e && *e && strcmp(e, "0") != 0;
  }
  void operator()(const char* what) {
    if (!on) return;
    const double t = now_ms();
    fprintf(stderr, "[ffi] %s: %-28s %9.2f ms (+%.2f)\n", fn, what, t - t0, t - last);
    last = t;
  }
};

char* c_string(const std::string& s) {  // C.CString: malloc'ed copy, never freed by the Rust caller
  char* p = (char*)malloc(s.size() + 1);
  if (!p) fatal("out of memory");
  memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  return p;
}

Span span_of(GoString s) { return Span{s.p ? s.p : "", s.n > 0 ? (size_t)s.n : 0}; }  // borrowed for the call, not NUL-dependent

void put_u64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 7; i >= 0; i--) v.push_back((uint8_t)(x >> (8 * i))); }
void put_u32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 3; i >= 0; i--) v.push_back((uint8_t)(x >> (8 * i))); }
void put_fr(std::vector<uint8_t>& v, const Fe4& a) { uint8_t b[32]; host::marshal(HFR, a, b); v.insert(v.end(), b, b + 32); }
void put_g1(std::vector<uint8_t>& v, const G1& p) { uint8_t b[32]; g1_compress(p, b); v.insert(v.end(), b, b + 32); }
Fe4 felt_from_be32(const uint8_t* b) { return host::set_bytes(HFR, b); }

// ---------------------------------------------------------------------------------------------- device + SRS state
struct State {
  b200zk_ctx* ctx = nullptr;
  b200zk_bases* bases = nullptr;
  size_t srs_n = 0;
  G2 g2[2];
  G2Prepared g2_prepared[2];  // Miller-loop lines of the two SRS elements (every verification pairs against them)
  std::vector<uint8_t> srs_file;  // u32 count || compressed G1 powers, until they are uploaded
  bool srs_ready = false, bases_ready = false;
  // Parsed circuits and device-resident keys, found again by the digest of the ACIR text (the reference re-parses the
  // JSON and re-derives the key's coset forms on every call).  Least recently used entries are dropped.
  struct Entry {
    Digest id;
    std::shared_ptr<Circuit> circuit;
    std::map<size_t, std::shared_ptr<struct Keyed>> by_nvalues;
  };
  std::list<Entry> circuits;
  // at process exit the CUDA runtime may already be gone: leave the device allocations to the driver
  ~State() { ctx = nullptr; }
};
struct Keyed {  // everything that depends on (circuit, number of values) only
  WirePlan plan;
  b200zk_plonk_pk* pk = nullptr;
  std::vector<uint8_t> vk_bytes;
  uint64_t n = 0, n_big = 0;
  Fe4* solution = nullptr;  // page-locked staging for the solution vector (one per key, reused by every prove)
  bool map_on_device = false;  // b200zk_plonk_set_solution_map done
  ~Keyed();
};
State& state() {
  static State s;
  return s;
}
Keyed::~Keyed() {
  if (solution && state().ctx) b200zk_host_free(state().ctx, solution);
  if (pk && state().ctx) b200zk_plonk_pk_free(state().ctx, pk);
}
size_t cache_capacity() {
  const char* e = getenv("B200ZK_FFI_CACHE");
  const long v = e ? atol(e) : 4;
  return v < 1 ? 1 : (size_t)v;
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
  {
    std::atomic<bool> bad(false);
    parallel_for(raw.size(), (size_t)1 << 20, [&](size_t b, size_t e) {
      if (!hex_decode_into(Span{text.data() + 2 * b, 2 * (e - b)}, raw.data() + b)) bad.store(true);
    });
    if (bad.load()) return false;
  }
  if (raw.size() < 4 + 128) return false;
  size_t n = ((size_t)raw[0] << 24) | ((size_t)raw[1] << 16) | ((size_t)raw[2] << 8) | raw[3];
  if (raw.size() != 4 + 32 * n + 128 || n == 0) return false;
  if (!g2_decompress(raw.data() + 4 + 32 * n, &s.g2[0]) || !g2_decompress(raw.data() + 4 + 32 * n + 64, &s.g2[1])) return false;
  if (!g2_in_subgroup(s.g2[0]) || !g2_in_subgroup(s.g2[1])) return false;   // as gnark's decoder: points of the twist outside G2 are refused
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
  State& s = state();
  if (s.bases_ready) return;
  if (!s.bases) {
    int rc = b200zk_bases_upload_compressed(context(), s.srs_file.data() + 4, s.srs_n, &s.bases);
    if (rc == B200ZK_ERR_BAD_ARG) fatal("invalid point in the SRS file");
    check(rc, "b200zk_bases_upload_compressed");
    std::vector<uint8_t>().swap(s.srs_file);
  }
  b200zk_bases_precompute(context(), s.bases, 0);  // best effort: commitments use classic windows if memory is short
  s.bases_ready = true;
}

// ---------------------------------------------------------------------------------------------- keys
struct Sizes { uint64_t n, n_big; };
Sizes domain_sizes(size_t nb_constraints, unsigned nb_public) {
  const size_t sys = nb_constraints + nb_public;
  uint64_t n = 2;
  while (n < sys) n <<= 1;
  uint64_t nb = 1;
  while (nb < (sys < 6 ? 8 : 4) * sys) nb <<= 1;
  if (nb < 4 * n) nb = 4 * n;
  return {n, nb};
}

// acir.UnmarshalJSON + BuildSparseR1CS's structural part, remembered per ACIR text
State::Entry& circuit_entry(Span acir) {
  State& s = state();
  const Digest id = digest(acir);
  for (auto it = s.circuits.begin(); it != s.circuits.end(); ++it)
    if (it->id.a == id.a && it->id.b == id.b) {
      s.circuits.splice(s.circuits.begin(), s.circuits, it);
      return s.circuits.front();
    }
  State::Entry e;
  e.id = id;
  e.circuit = std::make_shared<Circuit>(AcirReader(acir).read());
  while (s.circuits.size() >= cache_capacity()) s.circuits.pop_back();
  s.circuits.push_front(std::move(e));
  return s.circuits.front();
}
Keyed& keyed(State::Entry& e, size_t nvalues) {
  auto it = e.by_nvalues.find(nvalues);
  if (it != e.by_nvalues.end()) return *it->second;
  auto k = std::make_shared<Keyed>();
  k->plan = make_plan(*e.circuit, nvalues);
  Sizes sz = domain_sizes(e.circuit->size(), k->plan.nb_public);
  k->n = sz.n;
  k->n_big = sz.n_big;
  e.by_nvalues[nvalues] = k;
  return *k;
}
std::vector<uint8_t> serialize_vk(const Keyed& k, const uint8_t vk_points[8 * 64]);
// plonk.Setup(spr, srs) (plonk.go:21) on the device, once per (circuit, number of values)
void ensure_key(const Circuit& cs, Keyed& k) {
  if (k.pk) return;
  ensure_srs_bases();
  State& s = state();
  const size_t m = cs.size();
  static const Fe4 zero = {{0, 0, 0, 0}};
  static const uint32_t zero_w = 0;
  int rc = b200zk_plonk_setup_r1cs(context(), s.bases, k.plan.nb_public, k.plan.nb_secret, m, m ? (const void*)cs.ql.data() : &zero,
                                   m ? (const void*)cs.qr.data() : &zero, m ? (const void*)cs.qm.data() : &zero,
                                   m ? (const void*)cs.qo.data() : &zero, m ? (const void*)cs.qk.data() : &zero,
                                   m ? k.plan.a.data() : &zero_w, m ? k.plan.b.data() : &zero_w, m ? k.plan.c.data() : &zero_w, &k.pk);
  if (rc == B200ZK_ERR_BAD_ARG) fatal("kzg: the SRS is too small for this circuit (set B200ZK_SRS_SIZE)");
  check(rc, "plonk.Setup");
  uint8_t vkp[8 * 64];
  check(b200zk_plonk_vk(context(), k.pk, vkp), "b200zk_plonk_vk");
  k.vk_bytes = serialize_vk(k, vkp);
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
std::vector<uint8_t> serialize_vk(const Keyed& k, const uint8_t vk_points[8 * 64]) {
  std::vector<uint8_t> v;
  put_u64(v, k.n);
  put_fr(v, host::inv(HFR, host::from_u64(HFR, k.n)));
  put_fr(v, fr_root_of_unity(k.n));
  put_u64(v, k.plan.nb_public);
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

// kzg.BatchVerifyMultiPoints for the two openings of plonk.Verify: both checks
//   e(C_i - v_i G1 + z_i H_i, G2) * e(-H_i, alpha G2) == 1
// are folded with a challenge lambda derived from everything they depend on, so one product of two pairings decides:
//   e(sum_i lambda^i (C_i + z_i H_i) - (sum_i lambda^i v_i) G1, G2) * e(-sum_i lambda^i H_i, alpha G2) == 1
bool kzg_verify_two(const G1 C[2], const Fe4 z[2], const Fe4 v[2], const G1 Hq[2], const G2 g2[2], G2Prepared* prepared) {
  Transcript t;
  t.begin("lambda");
  for (int i = 0; i < 2; i++) {
    t.point(C[i]);
    t.point(Hq[i]);
    t.fr(z[i]);
    t.fr(v[i]);
  }
  const Fe4 lambda = t.finish();
  const Fe4 vsum = host::add(HFR, v[0], host::mul(HFR, lambda, v[1]));
  G1J lhs = g1_msm_small({Hq[0], C[1], Hq[1], g1_generator()},
                         {z[0], lambda, host::mul(HFR, lambda, z[1]), host::neg(HFR, vsum)});
  lhs = g1j_add_affine(lhs, C[0]);
  G1J hs = g1_msm_small({Hq[1]}, {lambda});
  hs = g1j_add_affine(hs, Hq[0]);
  const G1 a = g1j_to_affine(lhs), b = g1_neg(g1j_to_affine(hs));
  if (prepared && prepared[0].usable && prepared[1].usable)
    return pairing_product_is_one_prepared({{a, &prepared[0]}, {b, &prepared[1]}});
  // first verification of the process: the generic loop, which records the lines of the two SRS elements on the way
  return pairing_product_is_one({{a, g2[0]}, {b, g2[1]}}, prepared);
}

bool plonk_verify(const ParsedProof& pr, const ParsedVk& vk, const std::vector<Fe4>& pub, const G2 g2[2],
                  G2Prepared* prepared) {  // plonk.Verify
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
  G1J fh = g1_msm_small({pr.H[1], pr.H[2]}, {zpm, M(zpm, zpm)});  // H0 + zeta^(n+2) H1 + zeta^(2(n+2)) H2
  fh = g1j_add_affine(fh, pr.H[0]);
  const G1 folded_h = g1j_to_affine(fh);
  const Fe4 c_s3 = M(M(M(M(f1, f2), beta), alpha), zu);
  Fe4 c_z = M(M(A(A(M(beta, zeta), l), gamma), A(A(M(M(beta, u), zeta), r), gamma)), A(A(M(M(beta, uu), zeta), o), gamma));
  c_z = A(host::neg(HFR, M(c_z, alpha)), M(M(alpha, alpha), lag1));
  const G1* lp[7] = {&vk.Ql, &vk.Qr, &vk.Qm, &vk.Qo, &vk.Qk, &vk.S[2], &pr.Z};
  const Fe4 ls[7] = {l, r, M(l, r), o, one, c_s3, c_z};
  std::vector<G1> lpts;
  std::vector<Fe4> lsc;
  for (int i = 0; i < 7; i++)
    if (i != 4) {  // [Qk] enters with coefficient one
      lpts.push_back(*lp[i]);
      lsc.push_back(ls[i]);
    }
  const G1 lin_digest = g1j_to_affine(g1j_add_affine(g1_msm_small(lpts, lsc), vk.Qk));
  const G1 digests[7] = {folded_h, lin_digest, pr.LRO[0], pr.LRO[1], pr.LRO[2], vk.S[0], vk.S[1]};
  Transcript kz;
  kz.begin("gamma");
  kz.fr(zeta);
  for (int i = 0; i < 7; i++) kz.point(digests[i]);
  const Fe4 gk = kz.finish();
  Fe4 fe = {{0, 0, 0, 0}}, acc = one;
  std::vector<G1> dpts;
  std::vector<Fe4> dsc;
  for (int i = 0; i < 7; i++) {
    if (i) {  // the first digest enters with gamma^0 = 1
      dpts.push_back(digests[i]);
      dsc.push_back(acc);
    }
    fe = A(fe, M(pr.claimed[i], acc));
    acc = M(acc, gk);
  }
  const G1 Cs[2] = {g1j_to_affine(g1j_add_affine(g1_msm_small(dpts, dsc), digests[0])), pr.Z};
  const Fe4 zs[2] = {zeta, M(zeta, vk.generator)}, vs[2] = {fe, zu};
  const G1 Hs[2] = {pr.batched_H, pr.zshift_H};
  return kzg_verify_two(Cs, zs, vs, Hs, g2, prepared);
}

// hex text written straight into the malloc'ed result (the pk of a 2^20-row circuit is 650 MB of hex)
struct HexOut {
  char* buf;
  size_t cap, len = 0;
  explicit HexOut(size_t bytes) : cap(2 * bytes) {
    buf = (char*)malloc(cap + 1);
    if (!buf) fatal("out of memory");
  }
  void put(const std::vector<uint8_t>& v) {
    if (len + 2 * v.size() > cap) fatal("internal: hex buffer overflow");
    hex_encode_into(v.data(), v.size(), buf + len);
    len += 2 * v.size();
  }
  char* reserve(size_t bytes) {
    if (len + 2 * bytes > cap) fatal("internal: hex buffer overflow");
    char* p = buf + len;
    len += 2 * bytes;
    return p;
  }
  char* finish() {
    buf[len] = 0;
    return buf;
  }
};

// ProvingKey.WriteTo size.  The nine polynomials are []fr.Element (u32 count + elements); pk.Permutation is a []int64
// that gnark-crypto's encoder hands to binary.Write as it is: 3n big-endian int64 with NO count (ReadFrom sizes it from
// Domain[0].Cardinality).
size_t pk_stream_bytes(const Keyed& k) {
  return k.vk_bytes.size() + 2 * (8 + 5 * 32) + 9 * (4 + 32 * k.n) + 8 * 3 * k.n;
}

// the 9 blinding scalars of plonk.Prove: fr.SetRandom draws (their limbs are the Montgomery form)
// Tests pin the draws through the explicit entry point b200zk_ffi_test_seed_blinding (below); nothing in the
// environment can switch the shipping library to predictable blinding.
bool g_seeded_blinding = false;
uint64_t g_blinding_seed = 0;
void draw_blinding(Fe4 blinding[9]) {
  if (g_seeded_blinding) {
    uint64_t st = g_blinding_seed;
    for (int i = 0; i < 9;) {
      Fe4 v;
      for (int j = 0; j < 4; j++) {
        st += 0x9E3779B97F4A7C15ULL;
        uint64_t z = st;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        v.l[j] = z ^ (z >> 31);
      }
      v.l[3] &= 0x3fffffffffffffffULL;
      if (!host::geq(v.l, HFR.m)) blinding[i++] = v;
    }
    return;
  }
  FILE* ur = fopen("/dev/urandom", "rb");
  if (!ur) fatal("cannot open /dev/urandom");
  for (int i = 0; i < 9; i++) blinding[i] = random_fr(ur);
  fclose(ur);
}

// hex(u32-BE count || 32-byte elements) with a complete body: the count, else false (short / empty payloads take the
// host path, which reproduces the reference's ignored UnmarshalBinary error)
bool felts_count(Span h, size_t* count) {
  if (h.n % 2 || h.n < 8) return false;
  uint8_t head[4];
  if (!hex_decode_into(Span{h.p, 8}, head)) return false;
  const size_t n = ((size_t)head[0] << 24) | ((size_t)head[1] << 16) | ((size_t)head[2] << 8) | head[3];
  if (n == 0 || (h.n - 8) / 64 < n) return false;
  *count = n;
  return true;
}

}  // namespace

// ================================================================================================ exports
// The Go exports are reentrant; this library keeps per-process state (SRS, circuits, device keys), so the exports
// serialise on one lock.
std::mutex g_export_lock;

extern "C" {

// TEST-ONLY entry point (not in the reference, not declared in include/gnark_backend_ffi.h's export list): pins the 9
// blinding draws of the following proofs to a splitmix64 stream so that proof bytes can be compared with the CPU checker's.
// enable = 0 returns to /dev/urandom.  Proofs made with a known seed are NOT zero-knowledge.
void b200zk_ffi_test_seed_blinding(uint64_t seed, int enable) {
  std::lock_guard<std::mutex> hold(g_export_lock);
  g_seeded_blinding = enable != 0;
  g_blinding_seed = seed;
}

struct PlonkPreprocess_return PlonkPreprocess(GoString acirJSON, GoString encodedRandomValues) {  // main.go:58-78
  std::lock_guard<std::mutex> hold(g_export_lock);
  Trace trace("PlonkPreprocess");
  // the Rust side sends a JSON-quoted hex string (plonk/mod.rs:197-203); main.go:66-72 un-quotes it
  Span quoted = span_of(encodedRandomValues);
  while (quoted.n && (quoted.p[0] == ' ' || quoted.p[0] == '\n' || quoted.p[0] == '\t')) { quoted.p++; quoted.n--; }
  while (quoted.n && (quoted.p[quoted.n - 1] == ' ' || quoted.p[quoted.n - 1] == '\n' || quoted.p[quoted.n - 1] == '\t')) quoted.n--;
  if (quoted.n < 2 || quoted.p[0] != '"' || quoted.p[quoted.n - 1] != '"')
    fatal("json: cannot unmarshal encoded values into Go value of type string");
  std::vector<Fe4> values = felts_from_hex(Span{quoted.p + 1, quoted.n - 2});
  trace("values decoded");
  State::Entry& entry = circuit_entry(span_of(acirJSON));
  const Circuit& cs = *entry.circuit;
  trace("circuit read");
  Keyed& k = keyed(entry, values.size());
  trace("wire plan");
  ensure_key(cs, k);
  trace("plonk.Setup (device)");
  // ProvingKey.WriteTo (gnark v0.8.0, recalled): Vk | Domain[0] | Domain[1] | Ql Qr Qm Qo CQk LQk S1 S2 S3 | Permutation
  HexOut out(pk_stream_bytes(k));
  out.put(k.vk_bytes);
  {
    std::vector<uint8_t> d;
    put_domain(d, k.n);
    put_domain(d, k.n_big);
    out.put(d);
  }
  std::vector<Fe4> poly(k.n);
  for (int which = 0; which < 9; which++) {
    check(b200zk_plonk_pk_poly(context(), k.pk, which, poly.data()), "b200zk_plonk_pk_poly");
    std::vector<uint8_t> len;
    put_u32(len, (uint32_t)k.n);
    out.put(len);
    char* dst = out.reserve(32 * k.n);
    parallel_for(k.n, 4096, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
        uint8_t be[32];
        host::marshal(HFR, poly[i], be);
        hex_encode_into(be, 32, dst + 64 * i);
      }
    });
  }
  trace("polynomials -> hex");
  {
    // Permutation: gnark's buildPermutation over the same row layout
    const size_t n = k.n, m = cs.size();
    const unsigned np = k.plan.nb_public;
    std::vector<uint32_t> lro(3 * n, 0);
    for (unsigned i = 0; i < np; i++) lro[i] = i;
    for (size_t i = 0; i < m; i++) {
      lro[np + i] = k.plan.a[i];
      lro[n + np + i] = k.plan.b[i];
      lro[2 * n + np + i] = k.plan.c[i];
    }
    const size_t nw = np + k.plan.nb_secret ? np + k.plan.nb_secret : 1;
    std::vector<int64_t> perm(3 * n, -1), cycle(nw, -1);
    for (size_t i = 0; i < 3 * n; i++) {
      if (cycle[lro[i]] != -1) perm[i] = cycle[lro[i]];
      cycle[lro[i]] = (int64_t)i;
    }
    for (size_t i = 0; i < 3 * n; i++)
      if (perm[i] == -1) perm[i] = cycle[lro[i]];
    char* dst = out.reserve(8 * 3 * n);
    parallel_for(3 * n, 1 << 16, [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
        uint8_t be[8];
        for (int j = 0; j < 8; j++) be[j] = (uint8_t)((uint64_t)perm[i] >> (8 * (7 - j)));
        hex_encode_into(be, 8, dst + 16 * i);
      }
    });
  }
  trace("permutation -> hex");
  struct PlonkPreprocess_return r;
  r.r0 = out.finish();
  r.r1 = c_string(hex_encode(k.vk_bytes));
  return r;
}

char* PlonkProveWithPK(GoString acirJSON, GoString encodedValues, GoString encodedProvingKey) {  // main.go:24-37
  std::lock_guard<std::mutex> hold(g_export_lock);
  Trace trace("PlonkProveWithPK");
  const Span payload = span_of(encodedValues);
  size_t nvalues = 0;
  const bool on_device = felts_count(payload, &nvalues);  // DeserializeFelts + BuildWitnesses run on the GPU
  std::vector<Fe4> values;
  if (!on_device) {
    values = felts_from_hex(payload);
    nvalues = values.size();
  } else if (payload.n > 8 + 64 * nvalues) {  // bytes after the vector are ignored by UnmarshalBinary but must be hex
    std::vector<uint8_t> tmp((payload.n - 8 - 64 * nvalues) / 2);
    if (!hex_decode_into(Span{payload.p + 8 + 64 * nvalues, payload.n - 8 - 64 * nvalues}, tmp.data()))
      fatal("encoding/hex: invalid byte");
  }
  trace("values header / host decode");
  State::Entry& entry = circuit_entry(span_of(acirJSON));
  const Circuit& cs = *entry.circuit;
  trace("circuit read / found");
  Keyed& k = keyed(entry, nvalues);
  ensure_key(cs, k);
  trace("key resident");
  // DeserializeProvingKey (helpers.go:49-60): the polynomials of the key are already resident on the device (derived
  // from the same circuit and SRS; the reference itself re-derives the key's coset forms on every call), so the
  // payload is checked, not re-read: its length must be that of ProvingKey.WriteTo for this circuit and the
  // verifying key it starts with must be the one derived here.
  {
    Span pk = span_of(encodedProvingKey);
    if (pk.n % 2) fatal("encoding/hex: odd length hex string");
    // (keys written by round-1 builds of this library carried a 4-byte count before Permutation: still accepted)
    if (pk.n != 2 * pk_stream_bytes(k) && pk.n != 2 * (pk_stream_bytes(k) + 4))
      fatal(pk.n < 2 * pk_stream_bytes(k) ? "unexpected EOF reading the proving key" : "proving key does not belong to this circuit (size mismatch)");
    std::vector<uint8_t> head = hex_decode(Span{pk.p, 2 * k.vk_bytes.size()});
    if (head != k.vk_bytes) fatal("proving key does not belong to this circuit and SRS");
  }
  trace("proving key checked");
  Fe4 blinding[9];
  draw_blinding(blinding);
  uint8_t blob[832];
  int rc;
  const size_t nw = k.plan.solution_src.size();
  if (on_device && nw) {
    if (!k.map_on_device) {
      check(b200zk_plonk_set_solution_map(context(), k.pk, k.plan.solution_src.data(), nvalues), "b200zk_plonk_set_solution_map");
      k.map_on_device = true;
    }
    rc = b200zk_plonk_prove_hex(context(), k.pk, payload.p + 8, nvalues, blinding, blob);
    if (rc == B200ZK_ERR_BAD_ARG) fatal("encoding/hex: invalid byte");
  } else {
    // BuildWitnesses (common.go:22-43) on the host: publics then secrets = wire order
    if (!k.solution) {
      void* p = nullptr;
      check(b200zk_host_alloc(context(), (nw ? nw : 1) * sizeof(Fe4), &p), "b200zk_host_alloc");
      k.solution = (Fe4*)p;
      k.solution[0] = Fe4{{0, 0, 0, 0}};
    }
    for (size_t i = 0; i < nw; i++) k.solution[i] = values[k.plan.solution_src[i]];
    rc = b200zk_plonk_prove(context(), k.pk, k.solution, blinding, blob);
  }
  // every wire is an input here, so spr.Solve amounts to checking each constraint: done on the device at the start of
  // the prover; plonk.Prove's error is fatal in the reference (plonk.go:67-70)
  if (rc == B200ZK_ERR_UNSATISFIED)
    fatal("constraint #" + std::to_string(b200zk_plonk_unsatisfied_row(k.pk) - (long long)k.plan.nb_public) + " is not satisfied");
  check(rc, "plonk.Prove");
  trace("plonk.Prove (device)");
  return c_string(hex_encode(serialize_proof(blob)));
}

uint8_t PlonkVerifyWithMeta(GoString, GoString, GoString) { return 0; }  // main.go:39-42

uint8_t PlonkVerifyWithVK(GoString acirJSON, GoString encodedProof, GoString encodedPublicInputs,
                          GoString encodedVerifyingKey) {  // main.go:44-56, plonk.go:29-51
  std::lock_guard<std::mutex> hold(g_export_lock);
  Trace trace("PlonkVerifyWithVK");
  ParsedProof proof = parse_proof(hex_decode(span_of(encodedProof)));
  ParsedVk vk = parse_vk(hex_decode(span_of(encodedVerifyingKey)));
  // DeserializeFelts: the whole payload must be hex, but only the values that turn out to be public are needed
  const Span payload = span_of(encodedPublicInputs);
  size_t nvalues = 0;
  std::vector<Fe4> values;
  const bool lazy = felts_count(payload, &nvalues);
  if (lazy) {
    if (!hex_is_valid(payload)) fatal("encoding/hex: invalid byte");
  } else {
    values = felts_from_hex(payload);
    nvalues = values.size();
  }
  trace("payloads decoded");
  State::Entry& entry = circuit_entry(span_of(acirJSON));
  Keyed& k = keyed(entry, nvalues);  // only to learn which values are public (plonk.go:30)
  std::vector<Fe4> pub(k.plan.nb_public);
  for (unsigned i = 0; i < k.plan.nb_public; i++) {
    const size_t src = k.plan.solution_src[i];
    pub[i] = lazy ? felt_from_hex(Span{payload.p + 8 + 64 * src, 64}) : values[src];
  }
  trace("circuit read / found");
  // plonk.Verify refuses a public witness whose length differs from vk.NbPublicVariables ("invalid witness size"):
  // the reference then returns false (plonk.go:47-49)
  if (vk.nb_public != pub.size()) return 0;
  ensure_srs();  // vk.InitKZG(srs): the G2 elements live in the SRS file
  const bool ok = plonk_verify(proof, vk, pub, state().g2, state().g2_prepared);
  trace("plonk.Verify (host pairing)");
  return ok ? 1 : 0;
}

}  // extern "C"
