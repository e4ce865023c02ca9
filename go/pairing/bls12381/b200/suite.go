//go:build b200

package b200

/*
#include <b2kyber.h>
*/
import "C"

import (
	"crypto/cipher"
	"crypto/sha256"
	"hash"
	"io"

	"go.dedis.ch/kyber/v4"
	"go.dedis.ch/kyber/v4/pairing"
	"go.dedis.ch/kyber/v4/util/random"
	"go.dedis.ch/kyber/v4/xof/blake2xb"
)

// Suite implements pairing.Suite (pairing/pairing.go:8-20), shaped like kilic.Suite (kilic/suite.go:17-106).
type Suite struct {
	domainG1, domainG2 []byte
}

// NewBLS12381Suite is the default suite; NewBLS12381SuiteWithDST fixes the hash-to-curve tags (kilic/suite.go:32-46).
func NewBLS12381Suite() pairing.Suite { return &Suite{} }
func NewBLS12381SuiteWithDST(dstG1, dstG2 []byte) pairing.Suite {
	return &Suite{domainG1: dstG1, domainG2: dstG2}
}

func (s *Suite) G1() kyber.Group { return NewGroupG1(s.domainG1...) }
func (s *Suite) G2() kyber.Group { return NewGroupG2(s.domainG2...) }
func (s *Suite) GT() kyber.Group { return NewGroupGT() }

// ValidatePairing: e(p1,p2) == e(inv1,inv2) (kilic/suite.go:57-68), one 2-pair Miller loop + one final exp.
func (s *Suite) ValidatePairing(p1, p2, inv1, inv2 kyber.Point) bool {
	return s.ValidatePairingBatch([]kyber.Point{p1}, []kyber.Point{p2}, []kyber.Point{inv1}, []kyber.Point{inv2})[0]
}

// Pair returns e(p1,p2) as a GT element (kilic/suite.go:70-75).
func (s *Suite) Pair(p1, p2 kyber.Point) kyber.Point {
	gt := newEmptyGT()
	a, b := p1.(*G1Elt).aff, p2.(*G2Elt).aff
	with(func(e *engine) {
		e.check(C.b2k_bls12381_pair(e.ctx, 1, ptr(a[:]), ptr(b[:]), ptr(gt.b[:])))
	})
	return gt
}

func (s *Suite) Read(r io.Reader, objs ...interface{}) error  { panic("Suite.Read(): deprecated in dedis") } // kilic/suite.go:78-90
func (s *Suite) Write(w io.Writer, objs ...interface{}) error { panic("Suite.Write(): deprecated in dedis") }
func (s *Suite) Hash() hash.Hash                             { return sha256.New() }
func (s *Suite) XOF(seed []byte) kyber.XOF                   { return blake2xb.New(seed) }
func (s *Suite) RandomStream() cipher.Stream                 { return random.New() }
