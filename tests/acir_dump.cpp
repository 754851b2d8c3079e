// Test driver for csrc/ffi/acir_reader.h: reads an ACIR JSON file, applies the wire plan for <nvalues> values and prints
//   nb_public nb_secret
//   solution_src...
//   one line per gate: ql qr qm qo qk (64 hex chars, regular form) a b c
// so that tests/test_acir_reader.py can compare it with the oracle's restatement of the Go glue.
#include <fstream>
#include <iostream>
#include <sstream>
#include "../noir_backend_using_gnark_b200/csrc/ffi/acir_reader.h"

using namespace b200zk;
using namespace b200zk::ffi;

static std::string felt_hex(const Fe4& a) {
  uint8_t be[32];
  host::marshal(HFR, a, be);
  std::string s(64, '0');
  hex_encode_into(be, 32, &s[0]);
  return s;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  const size_t nvalues = strtoull(argv[2], nullptr, 10);
  Circuit c = AcirReader(Span{text.data(), text.size()}).read();
  WirePlan pl = make_plan(c, nvalues);
  Digest d = digest(Span{text.data(), text.size()});
  printf("%u %u %llu %016llx%016llx\n", pl.nb_public, pl.nb_secret, (unsigned long long)c.current_witness, (unsigned long long)d.a,
         (unsigned long long)d.b);
  for (uint32_t v : pl.solution_src) printf("%u ", v);
  printf("\n");
  for (size_t g = 0; g < c.size(); g++)
    printf("%s %s %s %s %s %u %u %u\n", felt_hex(c.ql[g]).c_str(), felt_hex(c.qr[g]).c_str(), felt_hex(c.qm[g]).c_str(),
           felt_hex(c.qo[g]).c_str(), felt_hex(c.qk[g]).c_str(), pl.a[g], pl.b[g], pl.c[g]);
  return 0;
}
