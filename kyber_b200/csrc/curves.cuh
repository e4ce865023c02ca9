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

// ---- operand validation ---------------------------------------------------------------------------------------------
// The reference's UnmarshalBinary is where a malformed point is refused (kilic/g1.go:127-131, pairing/bn254/point.go:146-185);
// an adapter may hand the engine affine bytes it did not obtain that way, so every entry point that consumes operand points
// re-checks what is cheap: canonical coordinates (where the reference checks them) and the curve equation y^2 = x^3 + b
// (3 field products per point).  A violation raises FLAG_POINT (-> B2K_ERR_POINT) and the point is treated as infinity.
// NOT re-checked: membership in the order-r subgroup (a scalar multiplication per point) -- that stays the job of
// *_decompress / *_unmarshal_check, exactly as in the reference.
template <class FC, int NCOORD>
B2K_D bool wire_coords_canonical(const uint8_t* p) {            // NCOORD big-endian field elements of 4 FC::N bytes each
  bool ok = true;
#pragma unroll
  for (int c = 0; c < NCOORD; c++) {
    Fp<FC> t;
    fp_load_be(t, p + 4 * FC::N * c);
    ok = ok && fp_canon_lt_mod(t);
  }
  return ok;
}
template <class CV>
B2K_D bool aff_on_curve(const Affine<typename CV::F>& p) {       // infinity (0, 0) counts as valid
  if (aff_is_inf(p)) return true;
  typename CV::F y2, x3, b;
  f_sqr(y2, p.y);
  f_sqr(x3, p.x); f_mul(x3, x3, p.x);
  CV::curve_b(b);
  f_add(x3, x3, b);
  return f_eq(y2, x3);
}
// operand bytes -> Montgomery affine; false (and infinity) when the operand is malformed
template <class CV>
B2K_D bool load_checked(Affine<typename CV::F>& r, const uint8_t* w) {
  CV::load(r, w);
  const bool ok = CV::wire_canonical(w) && aff_on_curve<CV>(r);
  if (!ok) aff_set_inf(r);
  return ok;
}

struct Bls381G1 {
  using FC = Bls381Fp;
  using F = Fp<Bls381Fp>;
  using ScalarField = Bls381Fr;
  static constexpr int SCALAR_BITS = 255;
  static constexpr int IN_BYTES = 96;
  static constexpr int OUT_BYTES = 48;
  B2K_D static bool wire_canonical(const uint8_t* p) { return wire_coords_canonical<FC, 2>(p); }
  B2K_D static void curve_b(F& b) {
#pragma unroll
    for (int j = 0; j < 12; j++) b.v[j] = FC::curve_b(j);
  }

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
  B2K_D static bool wire_canonical(const uint8_t* p) { return wire_coords_canonical<FC, 2>(p); }   // gfP.Unmarshal, pairing/bn254/gfp.go:101-119
  B2K_D static void curve_b(F& b) {
#pragma unroll
    for (int j = 0; j < 8; j++) b.v[j] = FC::curve_b(j);
  }

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
