//go:build b200

package b200

/*
#include <b2kyber.h>
*/
import "C"

import (
	"crypto/cipher"
	"encoding/hex"
	"errors"
	"io"

	"go.dedis.ch/kyber/v4"
	"go.dedis.ch/kyber/v4/group/mod"
)

// GTElt is an element of the target group, held as its 576 MarshalBinary bytes (12 x 48 B big-endian, highest tower
// coefficient first: kilic/gt.go:115-117; convention pinned by encrypt/ibe/ibe_test.go:202-245).  GT is written
// additively like every kyber group: Add = Fp12 product, Neg = inverse, Mul = exponentiation, Null = 1
// (kilic/gt.go:33-38, 59-83); the arithmetic runs on the device (b2k_bls12381_gt_mul / _inv / _exp).
type GTElt struct {
	b [576]byte
}

// gtOne: the bytes of 1 (only the last coefficient, c0.c0.c0, is 1).
var gtOne = func() (o [576]byte) { o[575] = 1; return }()

func newEmptyGT() *GTElt { return &GTElt{b: gtOne} }

func (k *GTElt) Equal(kk kyber.Point) bool { return k.b == kk.(*GTElt).b } // canonical bytes (kilic/gt.go:29-31)
func (k *GTElt) Null() kyber.Point         { k.b = gtOne; return k }
func (k *GTElt) Base() kyber.Point         { panic("bls12-381.GT.Base(): unsupported operation") } // kilic/gt.go:40-46
func (k *GTElt) Pick(_ cipher.Stream) kyber.Point {
	panic("bls12-381.GT.Pick(): unsupported operation")
}
func (k *GTElt) Set(q kyber.Point) kyber.Point { k.b = q.(*GTElt).b; return k }
func (k *GTElt) Clone() kyber.Point            { c := *k; return &c }

func (k *GTElt) Add(a, b kyber.Point) kyber.Point {
	x, y := a.(*GTElt).b, b.(*GTElt).b
	var out [576]byte
	with(func(e *engine) { e.check(C.b2k_bls12381_gt_mul(e.ctx, 1, ptr(x[:]), ptr(y[:]), ptr(out[:]))) })
	k.b = out
	return k
}
func (k *GTElt) Neg(q kyber.Point) kyber.Point {
	x := q.(*GTElt).b
	var out [576]byte
	with(func(e *engine) { e.check(C.b2k_bls12381_gt_inv(e.ctx, 1, ptr(x[:]), ptr(out[:]))) })
	k.b = out
	return k
}
func (k *GTElt) Sub(a, b kyber.Point) kyber.Point { // kilic/gt.go:66-69
	nb := newEmptyGT().Neg(b)
	return k.Add(a, nb)
}
func (k *GTElt) Mul(s kyber.Scalar, q kyber.Point) kyber.Point {
	sb, _ := s.(*mod.Int).MarshalBinary()
	x := q.(*GTElt).b
	var out [576]byte
	with(func(e *engine) { e.check(C.b2k_bls12381_gt_exp(e.ctx, 1, ptr(sb), ptr(x[:]), ptr(out[:]))) })
	k.b = out
	return k
}

func (k *GTElt) MarshalBinary() ([]byte, error) { out := make([]byte, 576); copy(out, k.b[:]); return out, nil }
func (k *GTElt) MarshalTo(w io.Writer) (int, error) {
	buf, _ := k.MarshalBinary()
	return w.Write(buf)
}

// pBE is the base-field modulus; UnmarshalBinary refuses a coefficient >= p like the back-ends' FromBytes.
var pBE, _ = hex.DecodeString("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab")

func (k *GTElt) UnmarshalBinary(buf []byte) error {
	if len(buf) != 576 {
		return errors.New("bls12-381: wrong buffer size for a GT element")
	}
	for i := 0; i < 12; i++ {
		c := buf[48*i : 48*i+48]
		ge := true // c >= p ?
		for j := 0; j < 48; j++ {
			if c[j] != pBE[j] {
				ge = c[j] > pBE[j]
				break
			}
		}
		if ge {
			return errors.New("bls12-381: GT coefficient is not a canonical field element")
		}
	}
	copy(k.b[:], buf)
	return nil
}
func (k *GTElt) UnmarshalFrom(r io.Reader) (int, error) {
	buf := make([]byte, k.MarshalSize())
	n, err := io.ReadFull(r, buf)
	if err != nil {
		return n, err
	}
	return n, k.UnmarshalBinary(buf)
}
func (k *GTElt) MarshalSize() int { return 576 }
func (k *GTElt) String() string   { return "bls12-381.GT: " + hex.EncodeToString(k.b[:]) }
func (k *GTElt) EmbedLen() int    { panic("bls12-381.GT.EmbedLen(): unsupported operation") }
func (k *GTElt) Embed(_ []byte, _ cipher.Stream) kyber.Point {
	panic("bls12-381.GT.Embed(): unsupported operation")
}
func (k *GTElt) Data() ([]byte, error) { panic("bls12-381.GT.Data(): unsupported operation") }
