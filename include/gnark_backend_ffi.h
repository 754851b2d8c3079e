/* The OUTER drop-in boundary: the C symbols the Rust crate binds
 * (/root/reference/src/gnark_backend_wrapper/plonk/mod.rs:10-25) and the Go archive exports
 * (/root/reference/gnark_backend_ffi/main.go:24, :39, :44, :58; cgo header libgnark_backend.h).
 *
 * In the reference these are implemented in Go.  No Go toolchain exists in this build environment, so
 * lib/libgnark_backend_b200.so provides a C++ stand-in with the same names, argument meaning, payload encodings and
 * error behaviour, driving the device-resident prover of libb200zk.so:
 *   - GoString arguments are passed BY VALUE ({pointer, length}; not NUL-dependent), borrowed for the call
 *   - ACIR           : serde_json of acvm 0.5 Circuit (acir/acir.go:17-75, acir/opcode/, acir/term/)
 *   - values         : hex( u32-BE count || 32-byte big-endian felts )   (src/gnark_backend_wrapper/serialize.rs:33-47);
 *                      for PlonkPreprocess the hex string is additionally JSON-quoted (plonk/mod.rs:197-203, main.go:66-72)
 *   - pk / vk / proof: hex of gnark's binary WriteTo streams (internal/backend/helpers.go:75-94)
 *   - results        : malloc'ed NUL-terminated strings (C.CString); the Rust caller never frees them
 *   - any failure    : message on stderr and exit(1), like log.Fatal (main.go:29,49,64,71; plonk.go:18..69)
 *   - SRS            : $XDG_CONFIG_HOME|$HOME/.config + /noir-lang/srs.hex, hex of kzg.SRS.WriteTo, generated with a
 *                      fresh random alpha when unreadable (backend/common.go:78-144).  Size: 1_000_000 points as in
 *                      common.go:137 unless B200ZK_SRS_SIZE is set (needed for circuits above 2^19 rows).
 * Test hook (not in the reference, not part of this header's export list): the library also exports
 * b200zk_ffi_test_seed_blinding(uint64 seed, int enable), which makes the 9 blinding draws of the prover deterministic
 * (SplitMix64 stream standing in for crypto/rand.Reader) so proofs can be compared byte for byte.  Nothing in the
 * environment can enable it.
 */
#ifndef GNARK_BACKEND_FFI_H
#define GNARK_BACKEND_FFI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { const char* p; ptrdiff_t n; } GoString;                 /* c_go_structures.rs:5-10 */
struct PlonkPreprocess_return { char* r0; char* r1; };                   /* KeyPair{proving_key, verifying_key}, c_go_structures.rs:22-26 */

char* PlonkProveWithPK(GoString acirJSON, GoString encodedValues, GoString encodedProvingKey);            /* main.go:24 */
uint8_t PlonkVerifyWithMeta(GoString acirJSON, GoString encodedValues, GoString encodedProof);            /* main.go:39: always 0 */
uint8_t PlonkVerifyWithVK(GoString acirJSON, GoString encodedProof, GoString encodedPublicInputs,
                          GoString encodedVerifyingKey);                                                   /* main.go:44 */
struct PlonkPreprocess_return PlonkPreprocess(GoString acirJSON, GoString encodedRandomValues);           /* main.go:58 */

#ifdef __cplusplus
}
#endif
#endif
