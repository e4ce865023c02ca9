//go:build b200

package b200

/*
#include <b2kyber.h>
*/
import "C"

import (
	"errors"

	"go.dedis.ch/kyber/v4"
	"go.dedis.ch/kyber/v4/group/mod"
)

// BatchGroup is the batch extension the reference lacks (SURVEY.md F6): the loops in
// sign/bdn/mask.go:58-61, sign/bdn/bdn.go:128-154, share/poly.go:145-147,461-473 and
// util/test/benchmark.go:66-69 are re-pointed at these.  groupBls implements it for G1 and G2.
type BatchGroup interface {
	MulBatch(dst []kyber.Point, s []kyber.Scalar, p []kyber.Point) // dst[i] = s[i]*p[i]
	MSM(s []kyber.Scalar, p []kyber.Point) kyber.Point             // sum s[i]*p[i]
}

// BatchSuite adds batched pairing checks and products of pairings with one final exponentiation.
type BatchSuite interface {
	ValidatePairingBatch(p1, p2, inv1, inv2 []kyber.Point) []bool
	PairingProductIsOne(g1, g2 []kyber.Point) bool
}

func packScalars(s []kyber.Scalar) []byte {
	sb := make([]byte, 32*len(s))
	for i := range s {
		b, _ := s[i].(*mod.Int).MarshalBinary()
		copy(sb[32*i:], b)
	}
	return sb
}

func (g *groupBls) isG1() bool { return g.str == "bls12-381.G1" }

func (g *groupBls) packPoints(p []kyber.Point) []byte {
	if g.isG1() {
		pb := make([]byte, 96*len(p))
		for i := range p {
			copy(pb[96*i:], p[i].(*G1Elt).aff[:])
		}
		return pb
	}
	pb := make([]byte, 192*len(p))
	for i := range p {
		copy(pb[192*i:], p[i].(*G2Elt).aff[:])
	}
	return pb
}

// MulBatch: n independent Point.Mul in one launch (kilic/g1.go:110-116 n times).
func (g *groupBls) MulBatch(dst []kyber.Point, s []kyber.Scalar, p []kyber.Point) {
	n := len(s)
	if n == 0 {
		return
	}
	sb, pb := packScalars(s), g.packPoints(p)
	if g.isG1() {
		out := make([]byte, 96*n)
		with(func(e *engine) {
			e.check(C.b2k_bls12381_g1_mul_batch_affine(e.ctx, C.size_t(n), ptr(sb), ptr(pb), ptr(out)))
		})
		for i := range dst {
			copy(dst[i].(*G1Elt).aff[:], out[96*i:96*i+96])
		}
		return
	}
	out := make([]byte, 192*n)
	with(func(e *engine) {
		e.check(C.b2k_bls12381_g2_mul_batch_affine(e.ctx, C.size_t(n), ptr(sb), ptr(pb), ptr(out)))
	})
	for i := range dst {
		copy(dst[i].(*G2Elt).aff[:], out[192*i:192*i+192])
	}
}

// MSM: sum s[i]*p[i] through the Pippenger pipeline (replaces the Mul+Add loops of share/poly.go:461-473).
func (g *groupBls) MSM(s []kyber.Scalar, p []kyber.Point) kyber.Point {
	n := len(s)
	sb, pb := packScalars(s), g.packPoints(p)
	if g.isG1() {
		r := NullG1()
		if n > 0 {
			with(func(e *engine) {
				e.check(C.b2k_bls12381_g1_msm_affine(e.ctx, C.size_t(n), ptr(sb), ptr(pb), ptr(r.aff[:])))
			})
		}
		return r
	}
	r := NullG2()
	if n > 0 {
		with(func(e *engine) {
			e.check(C.b2k_bls12381_g2_msm_affine(e.ctx, C.size_t(n), ptr(sb), ptr(pb), ptr(r.aff[:])))
		})
	}
	return r
}

// ValidatePairingBatch runs n independent e(p1,p2) == e(inv1,inv2) checks in one launch
// (n x Suite.ValidatePairing, kilic/suite.go:57-68).
func (s *Suite) ValidatePairingBatch(p1, p2, inv1, inv2 []kyber.Point) []bool {
	n := len(p1)
	res := make([]bool, n)
	if n == 0 {
		return res
	}
	a1, a2, b1, b2, ok := make([]byte, 96*n), make([]byte, 192*n), make([]byte, 96*n), make([]byte, 192*n), make([]byte, n)
	for i := 0; i < n; i++ {
		copy(a1[96*i:], p1[i].(*G1Elt).aff[:])
		copy(a2[192*i:], p2[i].(*G2Elt).aff[:])
		copy(b1[96*i:], inv1[i].(*G1Elt).aff[:])
		copy(b2[192*i:], inv2[i].(*G2Elt).aff[:])
	}
	with(func(e *engine) {
		e.check(C.b2k_bls12381_pairing_check(e.ctx, C.size_t(n), ptr(a1), ptr(a2), ptr(b1), ptr(b2), ptr(ok)))
	})
	for i := range res {
		res[i] = ok[i] != 0
	}
	return res
}

// PairingProductIsOne: prod_i e(g1[i], g2[i]) == 1 with n Miller loops and ONE final exponentiation
// (the use pointGT.Miller / Finalize are exported for, pairing/bn254/point.go:768-786).
func (s *Suite) PairingProductIsOne(g1, g2 []kyber.Point) bool {
	n := len(g1)
	if n == 0 {
		return true
	}
	a, b := make([]byte, 96*n), make([]byte, 192*n)
	for i := 0; i < n; i++ {
		copy(a[96*i:], g1[i].(*G1Elt).aff[:])
		copy(b[192*i:], g2[i].(*G2Elt).aff[:])
	}
	var ok [1]byte
	with(func(e *engine) {
		e.check(C.b2k_bls12381_pairing_product_check(e.ctx, C.size_t(n), ptr(a), ptr(b), ptr(ok[:])))
	})
	return ok[0] != 0
}

// MultiGPU shards one G1 MSM over ngpu devices of the box (BASELINE.json configs[4]; b2k_bls12381_g1_msm_multi_gpu):
// one context + one communicator per device, wired once with b2k_comm_connect_local; MSM may then be called repeatedly.
type MultiGPU struct {
	ctxs  []*C.b2k_ctx
	comms []*C.b2k_comm
}

func NewMultiGPU(ngpu int) (*MultiGPU, error) {
	m := &MultiGPU{ctxs: make([]*C.b2k_ctx, ngpu), comms: make([]*C.b2k_comm, ngpu)}
	for g := 0; g < ngpu; g++ {
		if rc := C.b2k_create(C.int(g), &m.ctxs[g]); rc != 0 {
			return nil, errors.New("b200: device unavailable")
		}
		if rc := C.b2k_comm_create(m.ctxs[g], C.int(ngpu), C.int(g), &m.comms[g]); rc != 0 {
			return nil, errors.New("b200: " + C.GoString(C.b2k_last_error(m.ctxs[g])))
		}
	}
	if rc := C.b2k_comm_connect_local(&m.comms[0], C.int(ngpu)); rc != 0 {
		return nil, errors.New("b200: peer access between the devices is not available")
	}
	return m, nil
}

// MSM: sum s[i]*p[i] over all devices; every device reduces its contiguous shard, the partial buckets are exchanged over
// NVLink inside the library, the 48-byte result is decoded like any UnmarshalBinary.
func (m *MultiGPU) MSM(s []kyber.Scalar, p []kyber.Point) (kyber.Point, error) {
	n := len(s)
	sb := packScalars(s)
	pb := make([]byte, 96*n)
	for i := range p {
		copy(pb[96*i:], p[i].(*G1Elt).aff[:])
	}
	var out [48]byte
	if rc := C.b2k_bls12381_g1_msm_multi_gpu(&m.comms[0], C.int(len(m.comms)), C.size_t(n), ptr(sb), ptr(pb), ptr(out[:])); rc != 0 {
		return nil, errors.New("b200: " + C.GoString(C.b2k_last_error(m.ctxs[0])))
	}
	r := NullG1()
	return r, r.UnmarshalBinary(out[:])
}

func (m *MultiGPU) Close() {
	for g := range m.comms {
		C.b2k_comm_destroy(m.comms[g])
		C.b2k_destroy(m.ctxs[g])
	}
}
