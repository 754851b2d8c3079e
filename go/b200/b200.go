// Package b200 is the cgo binding of libb200zk (include/b200zk.h): the sm_100a implementation of the arithmetic that
// gnark_backend_ffi reaches through plonk.Setup / plonk.Prove
// (gnark_backend_ffi/backend/plonk/plonk.go:21, :67 of lambdaclass/noir_backend_using_gnark).
//
// It deliberately imports nothing from gnark / gnark-crypto: the patched gnark-crypto (go/patch) imports THIS package,
// so the API is expressed in unsafe.Pointer + sizes over gnark-crypto's in-memory layouts
// (fr.Element / fp.Element = 4 x uint64 Montgomery limbs, G1Affine = X || Y, infinity = 64 zero bytes).
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  The C ABI underneath is exercised by the
// test-suite through ctypes with the same call sequences.
package b200

/*
#cgo CFLAGS:  -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../noir_backend_using_gnark_b200/lib -lb200zk -Wl,-rpath,${SRCDIR}/../../noir_backend_using_gnark_b200/lib
#include <stdlib.h>
#include "b200zk.h"
*/
import "C"

import (
	"fmt"
	"log"
	"sync"
	"unsafe"
)

var (
	once sync.Once
	ctx  *C.b200zk_ctx
	// a context is single-threaded (one stream); gnark commits from three goroutines (commitToLRO), so every entry
	// point takes this lock.  Every entry point of the library re-selects its device, so no LockOSThread is needed.
	mu sync.Mutex
)

func context() *C.b200zk_ctx {
	once.Do(func() {
		if rc := C.b200zk_init(0, &ctx); rc != 0 {
			// same failure mode as the reference: log.Fatal (backend/plonk/plonk.go:69)
			log.Fatal("b200zk_init: ", C.GoString(C.b200zk_strerror(rc)))
		}
	})
	return ctx
}

func fatalIf(rc C.int, what string) {
	if rc != 0 {
		log.Fatal(what, ": ", C.GoString(C.b200zk_strerror(rc)), " ", C.GoString(C.b200zk_last_cuda_error(context())))
	}
}

// Bases is the device-resident copy of kzg.SRS.G1 (uploaded once per SRS, not per commitment) together with its
// window table.
type Bases struct {
	h *C.b200zk_bases
	N int
}

// Upload copies n G1Affine (64 bytes each, starting at g1) to the device and precomputes the window multiples.
// []bn254.G1Affine holds no Go pointers, so &srs.G1[0] may cross cgo for the duration of the call.
func Upload(g1 unsafe.Pointer, n int) *Bases {
	mu.Lock()
	defer mu.Unlock()
	var h *C.b200zk_bases
	fatalIf(C.b200zk_bases_upload(context(), g1, C.size_t(n), &h), "b200zk_bases_upload")
	fatalIf(C.b200zk_bases_precompute(context(), h, 0), "b200zk_bases_precompute")
	return &Bases{h, n}
}

// MultiExp == (*bn254.G1Affine).MultiExp(points[:n], scalars[:n], cfg) of gnark-crypto v0.9.1 ecc/bn254/multiexp.go as
// kzg.Commit calls it: scalars are Montgomery-form fr.Element, out receives the canonical G1Affine (64 bytes).
func (b *Bases) MultiExp(scalars unsafe.Pointer, n int, out unsafe.Pointer) error {
	if n > b.N {
		return fmt.Errorf("b200: %d scalars for %d bases", n, b.N)
	}
	mu.Lock()
	defer mu.Unlock()
	fatalIf(C.b200zk_msm_g1(context(), b.h, scalars, C.size_t(n), out), "b200zk_msm_g1")
	return nil
}

// NTT == (*fft.Domain).FFT (inverse = false) / FFTInverse (inverse = true) of gnark-crypto v0.9.1
// ecc/bn254/fr/fft/fft.go on 2^log2n fr.Element starting at a, in place.
func NTT(a unsafe.Pointer, log2n uint, inverse, dit, coset bool) {
	mu.Lock()
	defer mu.Unlock()
	b := func(v bool) C.int {
		if v {
			return 1
		}
		return 0
	}
	dec := C.int(C.B200ZK_DIF)
	if dit {
		dec = C.B200ZK_DIT
	}
	fatalIf(C.b200zk_ntt(context(), a, C.uint(log2n), b(inverse), dec, b(coset)), "b200zk_ntt")
}

// BitReverse == fft.BitReverse(a).
func BitReverse(a unsafe.Pointer, log2n uint) {
	mu.Lock()
	defer mu.Unlock()
	fatalIf(C.b200zk_bit_reverse(context(), a, C.uint(log2n)), "b200zk_bit_reverse")
}

// Key is a device-resident plonk.ProvingKey (b200zk_plonk_setup_r1cs).
type Key struct {
	h        *C.b200zk_plonk_pk
	NbPublic int
	// S[0..2], Ql, Qr, Qm, Qo, Qk as G1Affine images (64 bytes each): the commitments of plonk.VerifyingKey
	VkPoints [8 * 64]byte
}

// Setup == plonk.Setup(spr, srs) (backend/plonk/plonk.go:21) from the SparseR1CS columns: one entry per constraint
// qL*xa + qR*xb + qO*xc + qM*(xa*xb) + qC == 0 (backend/plonk/sparse_r1cs.go:98-106) — coefficient VALUES as fr.Element
// (qm = coeff(M[0]) * coeff(M[1])) and the wire ids of L, R, O (public wires first).
func Setup(b *Bases, nbPublic, nbSecret, nbConstraints int, ql, qr, qm, qo, qk unsafe.Pointer, a, bb, c *uint32) *Key {
	mu.Lock()
	defer mu.Unlock()
	k := &Key{NbPublic: nbPublic}
	fatalIf(C.b200zk_plonk_setup_r1cs(context(), b.h, C.uint(nbPublic), C.uint(nbSecret), C.size_t(nbConstraints),
		ql, qr, qm, qo, qk, (*C.uint32_t)(unsafe.Pointer(a)), (*C.uint32_t)(unsafe.Pointer(bb)), (*C.uint32_t)(unsafe.Pointer(c)), &k.h),
		"b200zk_plonk_setup_r1cs")
	fatalIf(C.b200zk_plonk_vk(context(), k.h, unsafe.Pointer(&k.VkPoints[0])), "b200zk_plonk_vk")
	return k
}

// Prove == plonk.Prove(spr, pk, witness) (backend/plonk/plonk.go:67): solution = the value of every wire (public wires
// first, then secret), blinding = the nine fr.SetRandom draws in gnark's order L,L,R,R,O,O,Z,Z,Z.  The 832-byte result
// holds LRO[3], Z, H[3], BatchedProof.H, ZShiftedOpening.H as G1Affine and BatchedProof.ClaimedValues[7],
// ZShiftedOpening.ClaimedValue as fr.Element.  A solution that violates a constraint is fatal with the message
// gnark's solver would give.
func (k *Key) Prove(solution, blinding unsafe.Pointer) (out [832]byte) {
	mu.Lock()
	defer mu.Unlock()
	rc := C.b200zk_plonk_prove(context(), k.h, solution, blinding, unsafe.Pointer(&out[0]))
	if rc == C.B200ZK_ERR_UNSATISFIED {
		log.Fatalf("constraint #%d is not satisfied", int64(C.b200zk_plonk_unsatisfied_row(k.h))-int64(k.NbPublic))
	}
	fatalIf(rc, "b200zk_plonk_prove")
	return
}

// Free releases the device memory of the key.
func (k *Key) Free() {
	mu.Lock()
	defer mu.Unlock()
	C.b200zk_plonk_pk_free(context(), k.h)
	k.h = nil
}
