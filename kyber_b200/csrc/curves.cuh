// curves.cuh -- curve configurations: field, wire codecs and scalar range for each group on the path.
//
// Wire formats:
//   operand (engine input):  affine coordinates, big-endian, canonical (non-Montgomery), all-zero =
//                            infinity -- what a Go adapter holds after UnmarshalBinary.
//   result  (engine output): the reference's canonical MarshalBinary bytes:
//       BLS12-381 G1  48 B ZCash compressed      pairing/bls12381/kilic/g1.go:119-124,152-154
//       BLS12-381 G2  96 B ZCash compressed      pairing/bls12381/kilic/g2.go:118-123,151-153
//       bn254 G1      64 B x||y                  pairing/bn254/point.go:113-132
#pragma once
#include "constants.cuh"
#include "ec.cuh"

namespace b2k {

// 256-bit scalar as 8 little-endian limbs from the 32-byte big-endian mod.Int wire form
// (group/mod/int.go:334-349).
struct Scalar256 { uint32_t v[8]; };

B2K_D void scalar_load_be(Scalar256& s, const uint8_t* p) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint8_t* q = p + 4 * (7 - j);
    s.v[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
  }
}

// s < modulus of the scalar-field config R (mod.Int.UnmarshalBinary rejects s >= M, int.go:359-372)
template <class R>
B2K_D bool scalar_in_range(const Scalar256& s) {
  ptx::sub_cc(s.v[0], R::mod(0));
#pragma unroll
  for (int j = 1; j < 8; j++) ptx::subc_cc(s.v[j], R::mod(j));
  return ptx::subc(0, 0) != 0;
}

struct Bls381G1 {
  using FC = Bls381Fp;
  using F = Fp<Bls381Fp>;
  using ScalarField = Bls381Fr;
  static constexpr int SCALAR_BITS = 255;
  static constexpr int IN_BYTES = 96;
  static constexpr int OUT_BYTES = 48;

  // operand bytes -> Montgomery affine
  B2K_D static void load(Affine<F>& r, const uint8_t* p) {
    F x, y;
    fp_load_be(x, p);
    fp_load_be(y, p + 48);
    fp_to_mont(r.x, x);   // (0,0) stays (0,0): the infinity encoding survives the conversion
    fp_to_mont(r.y, y);
  }
  // Montgomery affine -> ZCash compressed bytes
  B2K_D static void store(uint8_t* out, const Affine<F>& p) {
    if (aff_is_inf(p)) {
      out[0] = 0xC0;
      for (int i = 1; i < 48; i++) out[i] = 0;
      return;
    }
    F x, y;
    fp_from_mont(x, p.x);
    fp_from_mont(y, p.y);
    fp_store_be(out, x);
    out[0] |= 0x80 | (fp_canon_gt_half(y) ? 0x20 : 0);
  }
  // Montgomery affine -> operand bytes (x||y)
  B2K_D static void store_affine(uint8_t* out, const Affine<F>& p) {
    F x, y;
    fp_from_mont(x, p.x);
    fp_from_mont(y, p.y);
    fp_store_be(out, x);
    fp_store_be(out + 48, y);
  }
  B2K_D static void generator(Affine<F>& g) {
#pragma unroll
    for (int j = 0; j < 12; j++) { g.x.v[j] = FC::gen_x(j); g.y.v[j] = FC::gen_y(j); }
  }
};

struct Bn254G1 {
  using FC = Bn254Fp;
  using F = Fp<Bn254Fp>;
  using ScalarField = Bn254Fr;
  static constexpr int SCALAR_BITS = 254;
  static constexpr int IN_BYTES = 64;
  static constexpr int OUT_BYTES = 64;

  B2K_D static void load(Affine<F>& r, const uint8_t* p) {
    F x, y;
    fp_load_be(x, p);
    fp_load_be(y, p + 32);
    fp_to_mont(r.x, x);
    fp_to_mont(r.y, y);
  }
  // pairing/bn254/point.go:113-132: x||y 32-byte big-endian each, infinity = all zeros
  B2K_D static void store(uint8_t* out, const Affine<F>& p) {
    F x, y;
    fp_from_mont(x, p.x);
    fp_from_mont(y, p.y);
    fp_store_be(out, x);
    fp_store_be(out + 32, y);
  }
  B2K_D static void store_affine(uint8_t* out, const Affine<F>& p) { store(out, p); }
  B2K_D static void generator(Affine<F>& g) {
#pragma unroll
    for (int j = 0; j < 8; j++) { g.x.v[j] = FC::gen_x(j); g.y.v[j] = FC::gen_y(j); }
  }
};

}  // namespace b2k
