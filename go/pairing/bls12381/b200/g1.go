//go:build b200

package b200

/*
#include <b2kyber.h>
*/
import "C"

import (
	"bytes"
	"crypto/cipher"
	"encoding/hex"
	"errors"
	"io"

	"go.dedis.ch/kyber/v4"
	"go.dedis.ch/kyber/v4/group/internal/marshalling"
	"go.dedis.ch/kyber/v4/group/mod"
)

var domainG1 = []byte("BLS_SIG_BLS12381G1_XMD:SHA-256_SSWU_RO_NUL_") // kilic/g1.go:17

// G1Elt holds the point in operand form: affine x||y, 48-byte big-endian each, all-zero = infinity.
type G1Elt struct {
	aff [96]byte
	dst []byte
}

var g1Generator = [96]byte{ /* x||y of the standard generator; see kyber_b200/host/kyber_b200.hpp G1_GEN */ }

func NullG1(dst ...byte) *G1Elt { return &G1Elt{dst: dst} }

func (k *G1Elt) Equal(k2 kyber.Point) bool {
	k2g1, ok := k2.(*G1Elt)
	if !ok {
		return false // kilic/g1.go:45-51
	}
	return k.aff == k2g1.aff
}
func (k *G1Elt) Null() kyber.Point  { k.aff = [96]byte{}; return k }
func (k *G1Elt) Base() kyber.Point  { k.aff = g1Generator; return k }
func (k *G1Elt) Clone() kyber.Point { c := *k; return &c }
func (k *G1Elt) Set(q kyber.Point) kyber.Point { k.aff = q.(*G1Elt).aff; return k }

func (k *G1Elt) Pick(rand cipher.Stream) kyber.Point { // kilic/g1.go:61-65: 32 random bytes -> hash-to-curve
	var buf [32]byte
	rand.XORKeyStream(buf[:], buf[:])
	return k.Hash(buf[:])
}

var oneBE = func() []byte { b := make([]byte, 32); b[31] = 1; return b }()
var minusOneBE = func() []byte { // r - 1
	b, _ := hex.DecodeString("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000000")
	return b
}()

func (k *G1Elt) lin(a, b *G1Elt, second []byte) kyber.Point {
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	sb := append(append([]byte{}, oneBE...), second...)
	pb := append(append([]byte{}, a.aff[:]...), b.aff[:]...)
	var out [96]byte
	e.check(C.b2k_bls12381_g1_msm_affine(e.ctx, 2, ptr(sb), ptr(pb), ptr(out[:])))
	k.aff = out
	return k
}
func (k *G1Elt) Add(a, b kyber.Point) kyber.Point { return k.lin(a.(*G1Elt), b.(*G1Elt), oneBE) }
func (k *G1Elt) Sub(a, b kyber.Point) kyber.Point { return k.lin(a.(*G1Elt), b.(*G1Elt), minusOneBE) }
func (k *G1Elt) Neg(a kyber.Point) kyber.Point {
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	var out [96]byte
	src := a.(*G1Elt).aff
	e.check(C.b2k_bls12381_g1_mul_batch_affine(e.ctx, 1, ptr(minusOneBE), ptr(src[:]), ptr(out[:])))
	k.aff = out
	return k
}

// Mul implements kyber.Point.Mul (kilic/g1.go:110-116): nil q means the generator.
func (k *G1Elt) Mul(s kyber.Scalar, q kyber.Point) kyber.Point {
	src := g1Generator
	if q != nil {
		src = q.(*G1Elt).aff
	}
	sb, _ := s.(*mod.Int).MarshalBinary() // 32 B big-endian, < r
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	var out [96]byte
	e.check(C.b2k_bls12381_g1_mul_batch_affine(e.ctx, 1, ptr(sb), ptr(src[:]), ptr(out[:])))
	k.aff = out
	return k
}

// MarshalBinary returns the 48-byte ZCash compressed form (kilic/g1.go:119-124).
func (k *G1Elt) MarshalBinary() ([]byte, error) {
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	out := make([]byte, 48)
	src := k.aff
	if rc := C.b2k_bls12381_g1_mul_batch(e.ctx, 1, ptr(oneBE), ptr(src[:]), ptr(out)); rc != 0 {
		return nil, errors.New("b200: " + e.lastError())
	}
	return out, nil
}

// UnmarshalBinary = FromCompressed + subgroup check (kilic/g1.go:127-131).
func (k *G1Elt) UnmarshalBinary(buff []byte) error {
	if len(buff) != 48 {
		return errors.New("bls12-381: wrong buffer size for a G1 point")
	}
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	var out [96]byte
	var ok [1]byte
	if rc := C.b2k_bls12381_g1_decompress(e.ctx, 1, ptr(buff), ptr(out[:]), ptr(ok[:])); rc != 0 {
		return errors.New("b200: " + e.lastError())
	}
	if ok[0] == 0 {
		return errors.New("bls12-381: invalid G1 point encoding")
	}
	k.aff = out
	return nil
}

func (k *G1Elt) MarshalTo(w io.Writer) (int, error)     { return marshalling.PointMarshalTo(k, w) }
func (k *G1Elt) UnmarshalFrom(r io.Reader) (int, error) { return marshalling.PointUnmarshalFrom(k, r) }
func (k *G1Elt) MarshalSize() int                       { return 48 }
func (k *G1Elt) String() string                         { b, _ := k.MarshalBinary(); return "bls12-381.G1: " + hex.EncodeToString(b) }
func (k *G1Elt) EmbedLen() int                          { panic("bls12-381: unsupported operation") } // kilic/g1.go:78-88
func (k *G1Elt) Embed(data []byte, rand cipher.Stream) kyber.Point {
	panic("bls12-381: unsupported operation")
}
func (k *G1Elt) Data() ([]byte, error) { panic("bls12-381: unsupported operation") }

// Hash implements kyber.HashablePoint (kilic/g1.go:161-170).
func (k *G1Elt) Hash(m []byte) kyber.Point {
	dst := k.dst
	if len(dst) == 0 {
		dst = domainG1
	}
	e := getEngine()
	e.mu.Lock()
	defer e.mu.Unlock()
	offs := []C.uint32_t{0, C.uint32_t(len(m))}
	msg := m
	if len(msg) == 0 {
		msg = []byte{0}
	}
	var out [96]byte
	e.check(C.b2k_bls12381_hash_to_g1(e.ctx, 1, ptr(msg), &offs[0], ptr(dst), C.uint32_t(len(dst)), ptr(out[:])))
	k.aff = out
	return k
}

// IsInCorrectGroup implements kyber.SubGroupElement (group.go:191-194).
func (k *G1Elt) IsInCorrectGroup() bool {
	b, err := k.MarshalBinary()
	if err != nil {
		return false
	}
	return NullG1().UnmarshalBinary(b) == nil
}

var _ = bytes.Equal
