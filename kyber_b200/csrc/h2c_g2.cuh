// h2c_g2.cuh -- RFC 9380 hash_to_curve for BLS12-381 G2 (suite BLS12381G2_XMD:SHA-256_SSWU_RO_), per thread.
//
// Replaces: kilic.G2Elt.Hash -> third-party HashToCurve(msg, dst)   pairing/bls12381/kilic/g2.go:160-169
// (default DST "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_", g2.go:18): the per-message step of bls.Sign/Verify
// for the scheme with signatures on G2 (sign/bls/bls.go:48-59), the drand default.
// Pipeline: expand_message_xmd (256 bytes) -> 2 elements of Fp2 -> simplified SWU on E2' -> derived 3-isogeny
// (tools/derive_isogeny.py; Velu, normalised by u = -3) evaluated projectively -> add -> clear cofactor with the
// endomorphism form [x^2-x-1]P + [x-1]psi(P) + psi^2(2P).  Pinned by the reference's drand KAT (tests/).
#pragma once
#include "h2c.cuh"

namespace b2k {

// expand_message_xmd with 256 output bytes (ell = 8)
B2K_NI void expand_message_xmd_256(uint8_t* out256, const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
  Sha256 s;
  uint8_t b0[32], bi[32], t[32];
  const uint8_t dl = (uint8_t)dst_len;
  const uint8_t lib[3] = {1, 0, 0};                 // I2OSP(256, 2) || I2OSP(0, 1)
  sha256_init(s);
  sha256_update_zero(s, 64);
  sha256_update(s, msg, msg_len);
  sha256_update(s, lib, 3);
  sha256_update(s, dst, dst_len);
  sha256_update(s, &dl, 1);
  sha256_final(s, b0);
  for (int i = 1; i <= 8; i++) {
    const uint8_t idx = (uint8_t)i;
    if (i == 1) { for (int k = 0; k < 32; k++) t[k] = b0[k]; }
    else { for (int k = 0; k < 32; k++) t[k] = b0[k] ^ bi[k]; }
    sha256_init(s);
    sha256_update(s, t, 32);
    sha256_update(s, &idx, 1);
    sha256_update(s, dst, dst_len);
    sha256_update(s, &dl, 1);
    sha256_final(s, bi);
    for (int k = 0; k < 32; k++) out256[32 * (i - 1) + k] = bi[k];
  }
}

B2K_D void fp2_load_table(BFp2& r, const uint32_t (*t)[12]) {
  fp_load_table(r.c0, t[0]);
  fp_load_table(r.c1, t[1]);
}

// sgn0 for Fp2: sign of c0, or of c1 when c0 is zero
B2K_D uint32_t fp2_sgn0(const BFp2& a_mont) {
  BFp c0, c1;
  fp_from_mont(c0, a_mont.c0);
  fp_from_mont(c1, a_mont.c1);
  const uint32_t s0 = c0.v[0] & 1u, z0 = fp_is_zero(c0) ? 1u : 0u, s1 = c1.v[0] & 1u;
  return s0 | (z0 & s1);
}

B2K_D void g2_curve_rhs(BFp2& r, const BFp2& x, const BFp2& A, const BFp2& B) {   // x^3 + A x + B
  BFp2 t;
  fp2_sqr(t, x); fp2_add(t, t, A); fp2_mul(t, t, x); fp2_add(r, t, B);
}

B2K_NI void map_to_curve_sswu_g2(BFp2& x, BFp2& y, const BFp2& u) {
  BFp2 A, B, Zc, zu2, tv1, x1, gx, c, yy;
  fp2_load_table(A, SSWU2_A); fp2_load_table(B, SSWU2_B); fp2_load_table(Zc, SSWU2_Z);
  fp2_sqr(zu2, u); fp2_mul(zu2, zu2, Zc);                 // Z u^2
  fp2_sqr(tv1, zu2); fp2_add(tv1, tv1, zu2);              // Z^2 u^4 + Z u^2
  if (fp2_is_zero(tv1)) {
    fp2_load_table(x1, SSWU2_B_OVER_ZA);
  } else {
    fp2_inv(tv1, tv1);
    fp2_set_one(c); fp2_add(tv1, tv1, c);
    fp2_load_table(c, SSWU2_NEG_B_OVER_A);
    fp2_mul(x1, c, tv1);
  }
  g2_curve_rhs(gx, x1, A, B);
  if (fp2_sqrt(yy, gx)) {
    x = x1;
  } else {
    fp2_mul(x, zu2, x1);
    g2_curve_rhs(gx, x, A, B);
    fp2_sqrt(yy, gx);
  }
  if (fp2_sgn0(u) != fp2_sgn0(yy)) fp2_neg(yy, yy);
  y = yy;
}

// 3-isogeny as a Jacobian point: d = x - x0, Z = 3d, X = x d^2 + v d + u, Y = -y (d^3 - v d - 2u)
B2K_NI void iso_map_g2(Jac<BFp2>& r, const BFp2& x, const BFp2& y) {
  BFp2 x0, v, u, u2, d, d2, t;
  fp2_load_table(x0, ISO_G2_X0); fp2_load_table(v, ISO_G2_V); fp2_load_table(u, ISO_G2_U); fp2_load_table(u2, ISO_G2_2U);
  fp2_sub(d, x, x0);
  if (fp2_is_zero(d)) { jac_set_inf(r); return; }
  fp2_sqr(d2, d);
  fp2_mul(r.X, x, d2); fp2_mul(t, v, d); fp2_add(r.X, r.X, t); fp2_add(r.X, r.X, u);
  fp2_mul(t, d2, d); BFp2 vd; fp2_mul(vd, v, d); fp2_sub(t, t, vd); fp2_sub(t, t, u2);
  fp2_mul(t, t, y); fp2_neg(r.Y, t);
  fp2_dbl(r.Z, d); fp2_add(r.Z, r.Z, d);
}

// psi on a Jacobian point: (conj X * cx, conj Y * cy, conj Z)
B2K_D void jac_psi(Jac<BFp2>& r, const Jac<BFp2>& p) {
  BFp2 cx, cy, c;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    cx.c0.v[j] = Bls381Fp::psi_cx0(j); cx.c1.v[j] = Bls381Fp::psi_cx1(j);
    cy.c0.v[j] = Bls381Fp::psi_cy0(j); cy.c1.v[j] = Bls381Fp::psi_cy1(j);
  }
  fp2_conj(c, p.X); fp2_mul(r.X, c, cx);
  fp2_conj(c, p.Y); fp2_mul(r.Y, c, cy);
  fp2_conj(r.Z, p.Z);
}

// [x^2 - x - 1]P + [x - 1]psi(P) + psi^2(2P),  x = -|x|  (multiplication by x = negated multiplication by |x|)
B2K_NI void g2_clear_cofactor(Jac<BFp2>& r, const Jac<BFp2>& p) {
  Jac<BFp2> t1, t2, t3, n;
  jac_mul_xabs_jac(t1, p); jac_neg(t1, t1);        // t1 = x P
  jac_psi(t2, p);                                  // t2 = psi(P)
  jac_dbl(t3, p); jac_psi(t3, t3); jac_psi(t3, t3);   // t3 = psi^2(2P)
  jac_neg(n, t2); jac_add(t3, t3, n);              // t3 -= t2
  jac_add(t2, t1, t2);                             // t2 = t1 + t2
  jac_mul_xabs_jac(n, t2); jac_neg(t2, n);         // t2 = x t2
  jac_add(t3, t3, t2);
  jac_neg(n, t1); jac_add(t3, t3, n);              // t3 -= t1
  jac_neg(n, p); jac_add(r, t3, n);                // - P
}

B2K_D void hash_to_g2(Affine<BFp2>& out, const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
  uint8_t uni[256];
  expand_message_xmd_256(uni, msg, msg_len, dst, dst_len);
  BFp2 u0, u1, x, y;
  fp_from_64_bytes(u0.c0, uni); fp_from_64_bytes(u0.c1, uni + 64);
  fp_from_64_bytes(u1.c0, uni + 128); fp_from_64_bytes(u1.c1, uni + 192);
  Jac<BFp2> q0, q1;
  map_to_curve_sswu_g2(x, y, u0); iso_map_g2(q0, x, y);
  map_to_curve_sswu_g2(x, y, u1); iso_map_g2(q1, x, y);
  jac_add(q0, q0, q1);
  g2_clear_cofactor(q1, q0);
  jac_to_affine(out, q1);
}

}  // namespace b2k
