//go:build b200

package b200

/*
#include <b2kyber.h>
*/
import "C"

import (
	"go.dedis.ch/kyber/v4"
	"go.dedis.ch/kyber/v4/group/mod"
)

// BatchGroup is the batch extension the reference lacks (SURVEY.md F6): the loops in
// sign/bdn/mask.go:58-61, sign/bdn/bdn.go:128-154, share/poly.go:145-147,461-473 and
// util/test/benchmark.go:66-69 are re-pointed at these.
type BatchGroup interface {
	MulBatch(dst []kyber.Point, s []kyber.Scalar, p []kyber.Point) // dst[i] = s[i]*p[i]
	MSM(s []kyber.Scalar, p []kyber.Point) kyber.Point             // sum s[i]*p[i]
}

// BatchSuite adds batched pairing checks.
type BatchSuite interface {
	ValidatePairingBatch(p1, p2, inv1, inv2 []kyber.Point) []bool
}

func packG1(s []kyber.Scalar, p []kyber.Point) (sb, pb []byte) {
	n := len(s)
	sb, pb = make([]byte, 32*n), make([]byte, 96*n)
	for i := range s {
		b, _ := s[i].(*mod.Int).MarshalBinary()
		copy(sb[32*i:], b)
		copy(pb[96*i:], p[i].(*G1Elt).aff[:])
	}
	return
}

func (g *groupG1) MulBatch(dst []kyber.Point, s []kyber.Scalar, p []kyber.Point) {
	n := len(s)
	sb, pb := packG1(s, p)
	out := make([]byte, 96*n)
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	e.check(C.b2k_bls12381_g1_mul_batch_affine(e.ctx, C.size_t(n), ptr(sb), ptr(pb), ptr(out)))
	for i := range dst {
		copy(dst[i].(*G1Elt).aff[:], out[96*i:96*i+96])
	}
}

func (g *groupG1) MSM(s []kyber.Scalar, p []kyber.Point) kyber.Point {
	sb, pb := packG1(s, p)
	r := NullG1()
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	e.check(C.b2k_bls12381_g1_msm_affine(e.ctx, C.size_t(len(s)), ptr(sb), ptr(pb), ptr(r.aff[:])))
	return r
}

// ValidatePairingBatch runs n independent e(p1,p2) == e(inv1,inv2) checks in one launch
// (n x Suite.ValidatePairing, kilic/suite.go:57-68).
func (s *Suite) ValidatePairingBatch(p1, p2, inv1, inv2 []kyber.Point) []bool {
	n := len(p1)
	a1, a2, b1, b2, ok := make([]byte, 96*n), make([]byte, 192*n), make([]byte, 96*n), make([]byte, 192*n), make([]byte, n)
	for i := 0; i < n; i++ {
		copy(a1[96*i:], p1[i].(*G1Elt).aff[:])
		copy(a2[192*i:], p2[i].(*G2Elt).aff[:])
		copy(b1[96*i:], inv1[i].(*G1Elt).aff[:])
		copy(b2[192*i:], inv2[i].(*G2Elt).aff[:])
	}
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	e.check(C.b2k_bls12381_pairing_check(e.ctx, C.size_t(n), ptr(a1), ptr(a2), ptr(b1), ptr(b2), ptr(ok)))
	res := make([]bool, n)
	for i := range res {
		res[i] = ok[i] != 0
	}
	return res
}
