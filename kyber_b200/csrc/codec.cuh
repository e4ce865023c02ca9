// codec.cuh -- BLS12-381 G2 curve config, ZCash point decompression and subgroup membership tests.
//
// Replaces (reference call sites):
//   kilic.G1Elt.UnmarshalBinary -> FromCompressed   pairing/bls12381/kilic/g1.go:127-131
//   kilic.G2Elt.UnmarshalBinary -> FromCompressed   pairing/bls12381/kilic/g2.go:126-130
//   kilic.G2Elt.Mul / MarshalBinary                 pairing/bls12381/kilic/g2.go:109-123
// Accept/reject behaviour is pinned by the reference's 16 + 18 ZCash fixtures
// (pairing/bls12381/deserialization_tests, tests/golden/bls12381_deserialization.json):
//   flag bits (bit7 compressed, bit6 infinity, bit5 sign), x < p, on curve, in the r-torsion subgroup.
// Subgroup tests use the curve endomorphisms (validated against [r]P = inf in tests/):
//   G1:  [x^2]P + phi(P) == inf,   phi(x,y) = (beta x, y)
//   G2:  psi(P) == [x]P,           psi = twist o Frobenius o untwist
#pragma once
#include "curves.cuh"
#include "pairing.cuh"

namespace b2k {

// ---- G2 config (F = Fp2) ---------------------------------------------------------------------------
B2K_D bool fp2_lex_largest(const BFp2& y_mont) {
  BFp c0, c1;
  fp_from_mont(c1, y_mont.c1);
  if (!fp_is_zero(c1)) return fp_canon_gt_half<Bls381Fp>(c1);
  fp_from_mont(c0, y_mont.c0);
  return fp_canon_gt_half<Bls381Fp>(c0);
}

struct Bls381G2 {
  using FC = Bls381Fp;
  using F = BFp2;
  using ScalarField = Bls381Fr;
  static constexpr int SCALAR_BITS = 255;
  static constexpr int IN_BYTES = 192;
  static constexpr int OUT_BYTES = 96;
  B2K_D static void load(Affine<F>& r, const uint8_t* p) { g2_load(r, p); }
  B2K_D static bool wire_canonical(const uint8_t* p) { return wire_coords_canonical<FC, 4>(p); }
  B2K_D static void curve_b(F& b) {              // 4 (1 + u)
    BFp one, four;
    fp_set_one(one); fp_add(four, one, one); fp_add(four, four, four);
    b.c0 = four; b.c1 = four;
  }
  // 96 B ZCash compressed: x.c1 || x.c0, flags in the first byte (kilic/g2.go:118-123)
  B2K_D static void store(uint8_t* out, const Affine<F>& p) {
    if (aff_is_inf(p)) {
      out[0] = 0xC0;
      for (int i = 1; i < 96; i++) out[i] = 0;
      return;
    }
    BFp t;
    fp_from_mont(t, p.x.c1); fp_store_be(out, t);
    fp_from_mont(t, p.x.c0); fp_store_be(out + 48, t);
    out[0] |= 0x80 | (fp2_lex_largest(p.y) ? 0x20 : 0);
  }
  B2K_D static void store_affine(uint8_t* out, const Affine<F>& p) {
    BFp t;
    fp_from_mont(t, p.x.c1); fp_store_be(out, t);
    fp_from_mont(t, p.x.c0); fp_store_be(out + 48, t);
    fp_from_mont(t, p.y.c1); fp_store_be(out + 96, t);
    fp_from_mont(t, p.y.c0); fp_store_be(out + 144, t);
  }
  B2K_D static void generator(Affine<F>& g) {
#pragma unroll
    for (int j = 0; j < 12; j++) {
      g.x.c0.v[j] = Bls381Fp::g2x0(j); g.x.c1.v[j] = Bls381Fp::g2x1(j);
      g.y.c0.v[j] = Bls381Fp::g2y0(j); g.y.c1.v[j] = Bls381Fp::g2y1(j);
    }
  }
};

// ---- multiplication by |x| = 0xd201000000010000 (64-bit, weight 6) -------------------------------------
template <class F>
B2K_D void jac_mul_xabs(Jac<F>& r, const Affine<F>& p) {
  Jac<F> acc;
  jac_from_affine(acc, p);
  for (int b = 62; b >= 0; b--) {
    jac_dbl(acc, acc);
    if ((BLS_X_ABS >> b) & 1) jac_madd(acc, acc, p);
  }
  r = acc;
}

// same with a Jacobian base point (5 full additions; avoids an inversion between two chained ladders)
template <class F>
B2K_D void jac_mul_xabs_jac(Jac<F>& r, const Jac<F>& p) {
  Jac<F> acc = p;
  for (int b = 62; b >= 0; b--) {
    jac_dbl(acc, acc);
    if ((BLS_X_ABS >> b) & 1) jac_add(acc, acc, p);
  }
  r = acc;
}

// affine equality of a Jacobian point with an affine one (no inversion): X == x Z^2, Y == y Z^3
template <class F>
B2K_D bool jac_eq_affine(const Jac<F>& a, const Affine<F>& b) {
  if (jac_is_inf(a)) return aff_is_inf(b);
  if (aff_is_inf(b)) return false;
  F zz, t;
  f_sqr(zz, a.Z);
  f_mul(t, b.x, zz);
  if (!f_eq(t, a.X)) return false;
  f_mul(t, zz, a.Z);
  f_mul(t, t, b.y);
  return f_eq(t, a.Y);
}

// P in G1  <=>  [x^2]P == -phi(P)
B2K_NI bool g1_in_subgroup(const Affine<BFp>& p) {
  if (aff_is_inf(p)) return true;
  Jac<BFp> t, t2;
  jac_mul_xabs(t, p);
  jac_mul_xabs_jac(t2, t);
  t = t2;
  Affine<BFp> m;                       // -phi(P) = (beta x, -y)
  BFp beta;
#pragma unroll
  for (int j = 0; j < 12; j++) beta.v[j] = Bls381Fp::beta(j);
  fp_mul(m.x, p.x, beta);
  fp_neg(m.y, p.y);
  return jac_eq_affine(t, m);
}

// P in G2  <=>  psi(P) == [x]P = -[|x|]P
B2K_NI bool g2_in_subgroup(const Affine<BFp2>& p) {
  if (aff_is_inf(p)) return true;
  Jac<BFp2> t;
  jac_mul_xabs(t, p);
  BFp2 cx, cy;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    cx.c0.v[j] = Bls381Fp::psi_cx0(j); cx.c1.v[j] = Bls381Fp::psi_cx1(j);
    cy.c0.v[j] = Bls381Fp::psi_cy0(j); cy.c1.v[j] = Bls381Fp::psi_cy1(j);
  }
  Affine<BFp2> m;                      // -psi(P)
  BFp2 c;
  fp2_conj(c, p.x); fp2_mul(m.x, c, cx);
  fp2_conj(c, p.y); fp2_mul(c, c, cy); fp2_neg(m.y, c);
  return jac_eq_affine(t, m);
}

// ---- square roots -----------------------------------------------------------------------------------------
// p = 3 mod 4: candidate a^((p+1)/4); returns false when a is not a square
B2K_D bool fp_sqrt(BFp& r, const BFp& a) {
  BFp c, c2;
  fp_pow_const<Bls381Fp, Bls381Fp::ExpSqrt>(c, a);
  fp_sqr_c(c2, c);
  if (!fp_eq(c2, a)) return false;
  r = c;
  return true;
}

// Fp2 square root by the complex method (norm, then two Fp square roots); false for non-squares
B2K_NI bool fp2_sqrt(BFp2& r, const BFp2& a) {
  BFp t, alpha, delta, x0, x1, inv2;
  if (fp_is_zero(a.c1)) {
    if (fp_sqrt(x0, a.c0)) { r.c0 = x0; fp_set_zero(r.c1); return true; }
    fp_neg(t, a.c0);
    if (fp_sqrt(x0, t)) { fp_set_zero(r.c0); r.c1 = x0; return true; }
    return false;
  }
  fp_sqr_c(alpha, a.c0); fp_sqr_c(t, a.c1); fp_add(alpha, alpha, t);
  if (!fp_sqrt(alpha, alpha)) return false;
  // 1/2 in Montgomery form = (R mod p) halved
  fp_set_one(inv2);
  {  // inv2 = one / 2 : if odd add p, then shift right
    uint32_t odd = inv2.v[0] & 1u;
    uint32_t tt[12], carry;
    tt[0] = ptx::add_cc(inv2.v[0], odd ? Bls381Fp::mod(0) : 0u);
#pragma unroll
    for (int j = 1; j < 12; j++) tt[j] = ptx::addc_cc(inv2.v[j], odd ? Bls381Fp::mod(j) : 0u);
    carry = ptx::addc(0, 0);
#pragma unroll
    for (int j = 0; j < 11; j++) inv2.v[j] = (tt[j] >> 1) | (tt[j + 1] << 31);
    inv2.v[11] = (tt[11] >> 1) | (carry << 31);
  }
  fp_add(delta, a.c0, alpha); fp_mul_c(delta, delta, inv2);
  if (!fp_sqrt(x0, delta)) {
    fp_sub(delta, a.c0, alpha); fp_mul_c(delta, delta, inv2);
    if (!fp_sqrt(x0, delta)) return false;
  }
  fp_add(t, x0, x0); fp_inv(t, t); fp_mul_c(x1, a.c1, t);
  BFp2 c, c2;
  c.c0 = x0; c.c1 = x1;
  fp2_sqr(c2, c);
  if (!fp2_eq(c2, a)) return false;
  r = c;
  return true;
}

// ---- decompression ------------------------------------------------------------------------------------------
// returns true and the Montgomery affine point (infinity = (0,0)) for a valid encoding in the subgroup
B2K_D bool g1_decompress(Affine<BFp>& out, const uint8_t* in, bool check_subgroup) {
  const uint32_t f = in[0];
  if (!(f & 0x80)) return false;                           // uncompressed form not accepted at 48 bytes
  BFp x;
  fp_load_be(x, in);
  x.v[11] &= 0x1fffffffu;                                  // strip the three flag bits
  if (f & 0x40) {                                          // infinity: sign flag clear, x zero
    aff_set_inf(out);
    return !(f & 0x20) && fp_is_zero(x);
  }
  if (!fp_canon_lt_mod<Bls381Fp>(x)) return false;
  BFp xm, y2, y, b;
  fp_to_mont(xm, x);
  fp_sqr_c(y2, xm); fp_mul_c(y2, y2, xm);
#pragma unroll
  for (int j = 0; j < 12; j++) b.v[j] = Bls381Fp::curve_b(j);
  fp_add(y2, y2, b);
  if (!fp_sqrt(y, y2)) return false;                       // not on the curve
  BFp yc;
  fp_from_mont(yc, y);
  if (fp_canon_gt_half<Bls381Fp>(yc) != ((f & 0x20) != 0)) fp_neg(y, y);
  out.x = xm; out.y = y;
  return check_subgroup ? g1_in_subgroup(out) : true;
}

B2K_D bool g2_decompress(Affine<BFp2>& out, const uint8_t* in, bool check_subgroup) {
  const uint32_t f = in[0];
  if (!(f & 0x80)) return false;
  BFp x1, x0;
  fp_load_be(x1, in);
  fp_load_be(x0, in + 48);
  x1.v[11] &= 0x1fffffffu;
  if (f & 0x40) {
    aff_set_inf(out);
    return !(f & 0x20) && fp_is_zero(x1) && fp_is_zero(x0);
  }
  if (!fp_canon_lt_mod<Bls381Fp>(x1) || !fp_canon_lt_mod<Bls381Fp>(x0)) return false;
  BFp2 x, y2, y, b;
  fp_to_mont(x.c0, x0); fp_to_mont(x.c1, x1);
  fp2_sqr(y2, x); fp2_mul(y2, y2, x);
#pragma unroll
  for (int j = 0; j < 12; j++) { b.c0.v[j] = Bls381Fp::curve_b(j); b.c1.v[j] = Bls381Fp::curve_b(j); }   // 4 + 4u
  fp2_add(y2, y2, b);
  if (!fp2_sqrt(y, y2)) return false;
  if (fp2_lex_largest(y) != ((f & 0x20) != 0)) fp2_neg(y, y);
  out.x = x; out.y = y;
  return check_subgroup ? g2_in_subgroup(out) : true;
}

}  // namespace b2k
