// gnark_ref: checks this repository's golden vectors against real gnark and times the real CPU prover.
// UNTESTED SOURCE — written without a Go toolchain; see README.md in this directory.
package main

import (
	"bytes"
	"crypto/rand"
	"encoding/binary"
	"encoding/hex"
	"encoding/json"
	"flag"
	"fmt"
	"log"
	"math/big"
	"os"
	"path/filepath"
	"runtime"
	"time"

	"gnark_backend_ffi/acir"
	"gnark_backend_ffi/backend"
	plonk_backend "gnark_backend_ffi/backend/plonk"

	"github.com/consensys/gnark-crypto/ecc"
	bn254 "github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr/fft"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr/kzg"
	"github.com/consensys/gnark/backend/plonk"
)

// ---- SplitMix64 byte stream standing in for crypto/rand.Reader (oracle/bn254.py splitmix64) ----
type splitMix struct {
	state uint64
	buf   []byte
}

func (s *splitMix) next() uint64 {
	s.state += 0x9E3779B97F4A7C15
	z := s.state
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

func (s *splitMix) Read(p []byte) (int, error) {
	for i := range p {
		if len(s.buf) == 0 {
			s.buf = make([]byte, 8)
			binary.LittleEndian.PutUint64(s.buf, s.next())
		}
		p[i] = s.buf[0]
		s.buf = s.buf[1:]
	}
	return len(p), nil
}

// ---- in-memory images: 4 x u64 little-endian limbs, Montgomery form ----
func frFromImage(b []byte) (e fr.Element) {
	for i := 0; i < 4; i++ {
		e[i] = binary.LittleEndian.Uint64(b[8*i:])
	}
	return
}

func frToImage(e fr.Element, out []byte) {
	for i := 0; i < 4; i++ {
		binary.LittleEndian.PutUint64(out[8*i:], e[i])
	}
}

func g1FromImage(b []byte) (p bn254.G1Affine) {
	for i := 0; i < 4; i++ {
		p.X[i] = binary.LittleEndian.Uint64(b[8*i:])
		p.Y[i] = binary.LittleEndian.Uint64(b[32+8*i:])
	}
	return
}

func g1ToImage(p bn254.G1Affine, out []byte) {
	for i := 0; i < 4; i++ {
		binary.LittleEndian.PutUint64(out[8*i:], p.X[i])
		binary.LittleEndian.PutUint64(out[32+8*i:], p.Y[i])
	}
}

type nttVec struct {
	Log2n      int    `json:"log2n"`
	Inverse    int    `json:"inverse"`
	Decimation int    `json:"decimation"` // 0 = DIF, 1 = DIT
	Coset      int    `json:"coset"`
	In         string `json:"in"`
	Out        string `json:"out"`
}

type msmVec struct {
	N       int    `json:"n"`
	Points  string `json:"points"`
	Scalars string `json:"scalars"`
	Out     string `json:"out"`
}

func checkNTT(dir string) bool {
	raw, err := os.ReadFile(filepath.Join(dir, "ntt_small.json"))
	if err != nil {
		log.Fatal(err)
	}
	var vecs []nttVec
	if err := json.Unmarshal(raw, &vecs); err != nil {
		log.Fatal(err)
	}
	ok := true
	for k, v := range vecs {
		in, _ := hex.DecodeString(v.In)
		n := 1 << v.Log2n
		a := make([]fr.Element, n)
		for i := range a {
			a[i] = frFromImage(in[32*i:])
		}
		d := fft.NewDomain(uint64(n))
		dec := fft.DIF
		if v.Decimation == 1 {
			dec = fft.DIT
		}
		if v.Inverse == 1 {
			d.FFTInverse(a, dec, v.Coset == 1)
		} else {
			d.FFT(a, dec, v.Coset == 1)
		}
		out := make([]byte, 32*n)
		for i := range a {
			frToImage(a[i], out[32*i:])
		}
		pass := hex.EncodeToString(out) == v.Out
		ok = ok && pass
		fmt.Printf("ntt[%d] log2n=%d inverse=%d decimation=%d coset=%d: %s\n", k, v.Log2n, v.Inverse, v.Decimation, v.Coset, verdict(pass))
	}
	return ok
}

func checkMSM(dir string) bool {
	raw, err := os.ReadFile(filepath.Join(dir, "msm_small.json"))
	if err != nil {
		log.Fatal(err)
	}
	var vecs []msmVec
	if err := json.Unmarshal(raw, &vecs); err != nil {
		log.Fatal(err)
	}
	ok := true
	for k, v := range vecs {
		pb, _ := hex.DecodeString(v.Points)
		sb, _ := hex.DecodeString(v.Scalars)
		pts := make([]bn254.G1Affine, v.N)
		sc := make([]fr.Element, v.N)
		for i := 0; i < v.N; i++ {
			pts[i] = g1FromImage(pb[64*i:])
			sc[i] = frFromImage(sb[32*i:])
		}
		var res bn254.G1Affine
		if _, err := res.MultiExp(pts, sc, ecc.MultiExpConfig{}); err != nil {
			log.Fatal(err)
		}
		out := make([]byte, 64)
		g1ToImage(res, out)
		pass := hex.EncodeToString(out) == v.Out
		ok = ok && pass
		fmt.Printf("msm[%d] n=%d: %s\n", k, v.N, verdict(pass))
	}
	return ok
}

func verdict(pass bool) string {
	if pass {
		return "PASS"
	}
	return "FAIL"
}

// ---- the reference's three embedded circuits (main.go:233-247), seeded ----
const m1 = "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000000"
const one = "0000000000000000000000000000000000000000000000000000000000000001"
const zero = "0000000000000000000000000000000000000000000000000000000000000000"

func fixture(lastLin, lastQc, publicInputs string) string {
	return fmt.Sprintf(`{"current_witness_index":6,"opcodes":[{"Arithmetic":{"mul_terms":[],"linear_combinations":[["%s",1],["%s",2],["%s",3]],"q_c":"%s"}},`+
		`{"Directive":{"Invert":{"x":3,"result":4}}},`+
		`{"Arithmetic":{"mul_terms":[["%s",3,4]],"linear_combinations":[["%s",5]],"q_c":"%s"}},`+
		`{"Arithmetic":{"mul_terms":[["%s",3,5]],"linear_combinations":[["%s",3]],"q_c":"%s"}},`+
		`{"Arithmetic":{"mul_terms":[],"linear_combinations":[["%s",5]],"q_c":"%s"}}],"public_inputs":%s}`,
		one, m1, m1, zero, one, m1, zero, one, m1, zero, lastLin, lastQc, publicInputs)
}

type plonkVec struct {
	Acir         string   `json:"acir"`
	Values       []string `json:"values"` // decimal, possibly negative
	SrsAlpha     string   `json:"srs_alpha"`
	SrsSize      int      `json:"srs_size"`
	BlindingSeed string   `json:"blinding_seed"`
	Vk           string   `json:"vk"`
	Pk           string   `json:"pk"`
	Proof        string   `json:"proof"`
	Verifies     bool     `json:"verifies"`
}

func plonkGolden(dir string) {
	cases := []struct {
		js   string
		vals []int64
	}{
		{fixture(m1, one, "[2]"), []int64{0, 1, -1, -1, 1, 0}},
		{fixture(one, zero, "[2]"), []int64{2, 2, 0, 0, 0, 0}},
		{fixture(one, zero, "[]"), []int64{3, 3, 0, 0, 0, 0}},
	}
	alpha := new(big.Int).SetUint64(0xB2000005)
	var out []plonkVec
	for _, c := range cases {
		var a acir.ACIR
		if err := json.Unmarshal([]byte(c.js), &a); err != nil {
			log.Fatal(err)
		}
		values := make(fr.Vector, len(c.vals))
		strs := make([]string, len(c.vals))
		for i, v := range c.vals {
			values[i].SetInt64(v)
			strs[i] = fmt.Sprint(v)
		}
		spr, pub, sec := plonk_backend.BuildSparseR1CS(a, values)
		witness := backend.BuildWitnesses(spr.CurveID().ScalarField(), pub, sec, spr.GetNbPublicVariables(), spr.GetNbSecretVariables())
		srs, err := kzg.NewSRS(128, alpha)
		if err != nil {
			log.Fatal(err)
		}
		pk, vk, err := plonk.Setup(spr, srs)
		if err != nil {
			log.Fatal(err)
		}
		saved := rand.Reader
		rand.Reader = &splitMix{state: 0xB2000006}
		proof, err := plonk.Prove(spr, pk, witness)
		rand.Reader = saved
		if err != nil {
			log.Fatal(err)
		}
		pubW, _ := witness.Public()
		var bvk, bpk, bpr bytes.Buffer
		vk.WriteTo(&bvk)
		pk.WriteTo(&bpk)
		proof.WriteTo(&bpr)
		out = append(out, plonkVec{c.js, strs, "0xB2000005", 128, "0xB2000006", hex.EncodeToString(bvk.Bytes()),
			hex.EncodeToString(bpk.Bytes()), hex.EncodeToString(bpr.Bytes()), plonk.Verify(proof, vk, pubW) == nil})
	}
	raw, _ := json.Marshal(out)
	if err := os.WriteFile(filepath.Join(dir, "gnark_plonk.json"), raw, 0o644); err != nil {
		log.Fatal(err)
	}
	fmt.Println("wrote", filepath.Join(dir, "gnark_plonk.json"))
}

// ---- CPU timings with the bench.py seeds ----
func randomFr(n int, seed uint64) []fr.Element {
	s := &splitMix{state: seed}
	r := fr.Modulus()
	out := make([]fr.Element, n)
	for i := range out {
		for {
			var l [4]uint64
			for k := range l {
				l[k] = s.next()
			}
			l[3] &= 0x3fffffffffffffff
			v := new(big.Int)
			for k := 3; k >= 0; k-- {
				v.Lsh(v, 64).Or(v, new(big.Int).SetUint64(l[k]))
			}
			if v.Cmp(r) < 0 {
				out[i] = fr.Element(l) // limbs taken as the Montgomery image, like bench.py
				break
			}
		}
	}
	return out
}

func benchAll(logMSM, logNTT int) {
	emit := func(name string, log2n int, d time.Duration) {
		fmt.Printf(`{"impl":"gnark","what":"%s","log2n":%d,"ms":%.3f,"cores":%d}`+"\n", name, log2n, float64(d.Microseconds())/1e3, runtime.NumCPU())
	}
	if logMSM > 0 {
		n := 1 << logMSM
		alpha := new(big.Int).SetUint64(0xB2000005 * 0x9E3779B97F4A7C15 % (1 << 62))
		srs, err := kzg.NewSRS(uint64(n), alpha)
		if err != nil {
			log.Fatal(err)
		}
		sc := randomFr(n, 0xB2000001)
		var res bn254.G1Affine
		t := time.Now()
		res.MultiExp(srs.G1[:n], sc, ecc.MultiExpConfig{})
		emit("G1Affine.MultiExp", logMSM, time.Since(t))
	}
	if logNTT > 0 {
		n := 1 << logNTT
		a := randomFr(n, 0xB2000003)
		d := fft.NewDomain(uint64(n))
		t := time.Now()
		d.FFT(a, fft.DIF)
		emit("fft.Domain.FFT DIF", logNTT, time.Since(t))
	}
}

func main() {
	golden := flag.String("golden", "", "directory holding ntt_small.json / msm_small.json (tests/golden)")
	bench := flag.Bool("bench", false, "time MultiExp / FFT on all cores")
	lm := flag.Int("msm", 20, "log2 points for -bench")
	ln := flag.Int("ntt", 20, "log2 size for -bench")
	flag.Parse()
	if *golden != "" {
		ok := checkNTT(*golden)
		ok = checkMSM(*golden) && ok
		plonkGolden(*golden)
		if !ok {
			os.Exit(1)
		}
	}
	if *bench {
		benchAll(*lm, *ln)
	}
}
