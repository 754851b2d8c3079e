// Route B of INTEGRATION.md: drop this file into gnark_backend_ffi/backend/plonk/ (package plonk_backend) and switch the
// two calls of plonk.go — :21 `plonk.Setup(sparseR1CS, srs)` and :67 `plonk.Prove(sparseR1CS, provingKey, witness)` — to
// SetupB200 / ProveB200.  Everything else (ACIR decode, HandleValues, BuildWitnesses, serialisation, the four cgo
// exports, the Rust crate) stays as it is.
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no Go toolchain in the build image); written against gnark v0.8.0 /
// gnark-crypto v0.9.1 as pinned by gnark_backend_ffi/go.mod:5,23.
package plonk_backend

import (
	"bytes"
	"encoding/binary"
	"log"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark-crypto/kzg"
	kzg_bn254 "github.com/consensys/gnark-crypto/ecc/bn254/fr/kzg"
	"github.com/consensys/gnark/backend/plonk"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"

	"gnark_backend_ffi/b200"
)

var basesOf = map[*bn254.G1Affine]*b200.Bases{} // one upload per SRS (keyed by &srs.G1[0])

func bases(srs kzg.SRS) *b200.Bases {
	s := srs.(*kzg_bn254.SRS)
	if b, ok := basesOf[&s.G1[0]]; ok {
		return b
	}
	b := b200.Upload(unsafe.Pointer(&s.G1[0]), len(s.G1))
	basesOf[&s.G1[0]] = b
	return b
}

// columns flattens the SparseR1CS into coefficient values and wire ids, one entry per constraint.
func columns(spr *cs_bn254.SparseR1CS) (ql, qr, qm, qo, qk []fr.Element, a, b, c []uint32) {
	n := len(spr.Constraints)
	ql, qr, qm, qo, qk = make([]fr.Element, n), make([]fr.Element, n), make([]fr.Element, n), make([]fr.Element, n), make([]fr.Element, n)
	a, b, c = make([]uint32, n), make([]uint32, n), make([]uint32, n)
	for i, g := range spr.Constraints {
		ql[i] = spr.Coefficients[g.L.CoeffID()]
		qr[i] = spr.Coefficients[g.R.CoeffID()]
		qo[i] = spr.Coefficients[g.O.CoeffID()]
		qm[i].Mul(&spr.Coefficients[g.M[0].CoeffID()], &spr.Coefficients[g.M[1].CoeffID()])
		qk[i] = spr.Coefficients[g.K]
		a[i], b[i], c[i] = uint32(g.L.WireID()), uint32(g.R.WireID()), uint32(g.O.WireID())
	}
	return
}

// B200Key keeps the device-resident proving key next to gnark's verifying key.
type B200Key struct {
	dev *b200.Key
	Vk  plonk.VerifyingKey
}

// SetupB200 replaces plonk.Setup at plonk.go:21.
func SetupB200(spr *cs_bn254.SparseR1CS, srs kzg.SRS) *B200Key {
	if len(spr.Constraints) == 0 {
		log.Fatal("empty constraint system")
	}
	ql, qr, qm, qo, qk, a, b, c := columns(spr)
	dev := b200.Setup(bases(srs), len(spr.Public), len(spr.Secret), len(spr.Constraints),
		unsafe.Pointer(&ql[0]), unsafe.Pointer(&qr[0]), unsafe.Pointer(&qm[0]), unsafe.Pointer(&qo[0]), unsafe.Pointer(&qk[0]),
		&a[0], &b[0], &c[0])
	// plonk.VerifyingKey through its public ReadFrom: Size u64 | SizeInv | Generator | NbPublicVariables u64 |
	// S[0..2] Ql Qr Qm Qo Qk (compressed G1)
	size := uint64(ecc.NextPowerOfTwo(uint64(len(spr.Constraints) + len(spr.Public))))
	var sizeInv, gen fr.Element
	sizeInv.SetUint64(size).Inverse(&sizeInv)
	gen = generatorOf(size)
	var buf bytes.Buffer
	binary.Write(&buf, binary.BigEndian, size)
	buf.Write(sizeInv.Marshal())
	buf.Write(gen.Marshal())
	binary.Write(&buf, binary.BigEndian, uint64(len(spr.Public)))
	for i := 0; i < 8; i++ {
		p := (*bn254.G1Affine)(unsafe.Pointer(&dev.VkPoints[64*i]))
		cb := p.Bytes()
		buf.Write(cb[:])
	}
	vk := plonk.NewVerifyingKey(ecc.BN254)
	if _, err := vk.ReadFrom(&buf); err != nil {
		log.Fatal(err)
	}
	return &B200Key{dev, vk}
}

// generatorOf: fft.NewDomain(size).Generator (the 2^28-th root of unity raised to 2^(28 - log2 size)).
func generatorOf(size uint64) (g fr.Element) {
	g.SetString("19103219067921713944291392827692070036145651957329286315305642004821462161904")
	for s := uint64(1) << 28; s > size; s >>= 1 {
		g.Square(&g)
	}
	return
}

// ProveB200 replaces plonk.Prove at plonk.go:67.  solution = every wire's value, public wires first (what
// BuildWitnesses lays out, common.go:27-32).
func ProveB200(k *B200Key, solution []fr.Element) plonk.Proof {
	var blinding [9]fr.Element
	for i := range blinding {
		blinding[i].SetRandom() // same draws, same order as gnark: L,L,R,R,O,O,Z,Z,Z
	}
	blob := k.dev.Prove(unsafe.Pointer(&solution[0]), unsafe.Pointer(&blinding[0]))
	// proof.WriteTo stream (548 bytes): LRO[3], Z, H[3] compressed | BatchedProof.H | u32 7 | 7 fr | ZShifted.H | fr
	pt := func(i int) []byte {
		p := (*bn254.G1Affine)(unsafe.Pointer(&blob[64*i]))
		cb := p.Bytes()
		return cb[:]
	}
	sc := func(i int) []byte { return (*fr.Element)(unsafe.Pointer(&blob[576+32*i])).Marshal() }
	var buf bytes.Buffer
	for i := 0; i < 8; i++ {
		buf.Write(pt(i))
	}
	binary.Write(&buf, binary.BigEndian, uint32(7))
	for i := 0; i < 7; i++ {
		buf.Write(sc(i))
	}
	buf.Write(pt(8))
	buf.Write(sc(7))
	proof := plonk.NewProof(ecc.BN254)
	if _, err := proof.ReadFrom(&buf); err != nil {
		log.Fatal(err)
	}
	return proof
}
