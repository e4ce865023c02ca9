// pairing.cuh -- BLS12-381 optimal-ate pairing: Miller loop over |x| with inversion-free line
// functions, final exponentiation (exponent 3 (p^12-1)/r, the convention of the reference back-ends), GT codec.
//
// Replaces (reference call sites; the arithmetic itself is third-party in the reference):
//   kilic.Suite.Pair             pairing/bls12381/kilic/suite.go:70-75   (Engine.AddPair + Result)
//   kilic.Suite.ValidatePairing  pairing/bls12381/kilic/suite.go:57-68   (AddPair, AddPairInv, Check)
//   GT MarshalBinary             pairing/bls12381/kilic/gt.go:115-117    (576 bytes)
// Same structure as the in-tree bn254 pairing it is modelled on: miller / lineFunctionDouble /
// lineFunctionAdd / mulLine / finalExponentiation, pairing/bn254/optate.go:5-271.
//
// Lines.  The twist point T = (X,Y,Z) is Jacobian over Fp2, P = (xP,yP) in G1.  With the untwist
// (x',y') -> (x'/w^2, y'/w^3) a line evaluated at P, scaled by factors in Fp2 (killed by the final
// exponentiation), is  l0 + l2 w^2 + l3 w^3  with
//   doubling:  l0 = 3X^3 - 2Y^2,      l2 = -(3X^2 Z^2) xP,   l3 = (Z3 Z^2) yP      (Z3 = 2YZ)
//   addition:  l0 = r x2 - Z3 y2,     l2 = -r xP,            l3 = Z3 yP           (Z3 = Z H)
// In the tower Fp12 = Fp6[w]/(w^2 - v) that is the sparse element  c0 = (l0, l2, 0), c1 = (0, l3, 0).
#pragma once
#include "ec.cuh"

namespace b2k {

using BT = Bls381Tower;
using BFp = Fp<Bls381Fp>;
using BFp2 = Fp2<Bls381Fp>;
using BFp6 = Fp6<BT>;
using BFp12 = Fp12<BT>;

constexpr uint64_t BLS_X_ABS = 0xd201000000010000ULL;        // the curve parameter is -BLS_X_ABS
constexpr uint64_t BLS_E3 = 0x460055555555aaabULL;           // (BLS_X_ABS + 1) / 3

struct BlsLine { BFp2 l0, l2, l3; };

// f *= (l0, l2, 0) + (0, l3, 0) w        ("mul by 014"): 13 Fp2 multiplications
B2K_NI void fp12_mul_line(BFp12& f, const BlsLine& l) {
  BFp6 aa, bb, t;
  BFp2 o;
  fp6_mul_by_01(aa, f.c0, l.l0, l.l2);
  fp6_mul_by_1(bb, f.c1, l.l3);
  fp2_add(o, l.l2, l.l3);
  fp6_add(t, f.c1, f.c0);
  fp6_mul_by_01(t, t, l.l0, o);
  fp6_sub(t, t, aa);
  fp6_sub(f.c1, t, bb);
  fp6_mul_v(bb, bb);
  fp6_add(f.c0, bb, aa);
}

// T <- 2T and the tangent line at T evaluated at P
B2K_NI void miller_double_step(BlsLine& l, Jac<BFp2>& T, const Affine<BFp>& P) {
  BFp2 A, B, C, D, E, ZZ, t;
  fp2_sqr(A, T.X);
  fp2_sqr(B, T.Y);
  fp2_sqr(C, B);
  fp2_sqr(ZZ, T.Z);
  fp2_add(D, T.X, B); fp2_sqr(D, D); fp2_sub(D, D, A); fp2_sub(D, D, C); fp2_dbl(D, D);
  fp2_dbl(E, A); fp2_add(E, E, A);                       // 3X^2
  // line: l0 = E X - 2B ; l2 = -(E ZZ) xP ; l3 = (Z3 ZZ) yP
  fp2_mul(l.l0, E, T.X); fp2_sub(l.l0, l.l0, B); fp2_sub(l.l0, l.l0, B);
  fp2_mul(t, E, ZZ); fp2_mul_fp(t, t, P.x); fp2_neg(l.l2, t);
  fp2_mul(t, T.Y, T.Z); fp2_dbl(T.Z, t);                 // Z3 = 2YZ
  fp2_mul(t, T.Z, ZZ); fp2_mul_fp(l.l3, t, P.y);
  // point
  fp2_sqr(A, E); fp2_sub(A, A, D); fp2_sub(A, A, D);     // X3
  fp2_dbl(C, C); fp2_dbl(C, C); fp2_dbl(C, C);
  fp2_sub(D, D, A); fp2_mul(D, E, D); fp2_sub(T.Y, D, C);
  T.X = A;
}

// T <- T + Q (Q affine, Q != +-T inside the loop because |x| < r) and the chord line at P
B2K_NI void miller_add_step(BlsLine& l, Jac<BFp2>& T, const Affine<BFp2>& Q, const Affine<BFp>& P) {
  BFp2 ZZ, U2, S2, H, R, HH, HHH, V, t;
  fp2_sqr(ZZ, T.Z);
  fp2_mul(U2, Q.x, ZZ);
  fp2_mul(S2, Q.y, T.Z); fp2_mul(S2, S2, ZZ);
  fp2_sub(H, U2, T.X);
  fp2_sub(R, S2, T.Y);
  fp2_sqr(HH, H);
  fp2_mul(HHH, H, HH);
  fp2_mul(V, T.X, HH);
  fp2_mul(T.Z, T.Z, H);                                   // Z3 = Z H
  // line: l0 = R x2 - Z3 y2 ; l2 = -R xP ; l3 = Z3 yP
  fp2_mul(l.l0, R, Q.x); fp2_mul(t, T.Z, Q.y); fp2_sub(l.l0, l.l0, t);
  fp2_mul_fp(t, R, P.x); fp2_neg(l.l2, t);
  fp2_mul_fp(l.l3, T.Z, P.y);
  // point
  fp2_sqr(t, R); fp2_sub(t, t, HHH); fp2_sub(t, t, V); fp2_sub(t, t, V);   // X3
  fp2_sub(V, V, t); fp2_mul(V, R, V);
  fp2_mul(HHH, T.Y, HHH);
  fp2_sub(T.Y, V, HHH);
  T.X = t;
}

// f = prod_i f_{|x|,Q_i}(P_i), conjugated (x < 0).  Pairs with an infinity member contribute 1
// (like bn254 optimalAte, pairing/bn254/optate.go:267-269).
template <int NPAIRS>
B2K_D void miller_loop(BFp12& f, const Affine<BFp>* P, const Affine<BFp2>* Q) {
  Jac<BFp2> T[NPAIRS];
  bool live[NPAIRS];
#pragma unroll
  for (int i = 0; i < NPAIRS; i++) {
    live[i] = !(aff_is_inf(P[i]) || aff_is_inf(Q[i]));
    T[i].X = Q[i].x; T[i].Y = Q[i].y; fp2_set_one(T[i].Z);
  }
  fp12_set_one(f);
  BlsLine l;
  for (int b = 62; b >= 0; b--) {
    fp12_sqr(f, f);
#pragma unroll
    for (int i = 0; i < NPAIRS; i++) {
      if (!live[i]) continue;
      miller_double_step(l, T[i], P[i]);
      fp12_mul_line(f, l);
    }
    if ((BLS_X_ABS >> b) & 1) {
#pragma unroll
      for (int i = 0; i < NPAIRS; i++) {
        if (!live[i]) continue;
        miller_add_step(l, T[i], Q[i], P[i]);
        fp12_mul_line(f, l);
      }
    }
  }
  fp12_conj(f, f);
}

// ---- Frobenius ---------------------------------------------------------------------------------
// f = sum_k a_k w^k (a_k in Fp2; tower slots: w^0 c0.c0, w^1 c1.c0, w^2 c0.c1, w^3 c1.c1, w^4 c0.c2,
// w^5 c1.c2);  f^(p^j) = sum_k conj^j(a_k) * xi^(k (p^j-1)/6) w^k.
#define B2K_FROB_COEF(J, K, dst)                                                     \
  {                                                                                  \
    _Pragma("unroll") for (int q = 0; q < 12; q++) {                                 \
      (dst).c0.v[q] = Bls381Fp::frob##J##_##K##_c0(q);                               \
      (dst).c1.v[q] = Bls381Fp::frob##J##_##K##_c1(q);                               \
    }                                                                                \
  }

template <int J>
B2K_NI void fp12_frobenius(BFp12& r, const BFp12& f) {
  BFp2 g, a;
  // k = 0
  a = f.c0.c0; if (J & 1) fp2_conj(a, a); r.c0.c0 = a;
  // k = 1..5
  if (J == 1) {
    B2K_FROB_COEF(1, 1, g); fp2_conj(a, f.c1.c0); fp2_mul(r.c1.c0, a, g);
    B2K_FROB_COEF(1, 2, g); fp2_conj(a, f.c0.c1); fp2_mul(r.c0.c1, a, g);
    B2K_FROB_COEF(1, 3, g); fp2_conj(a, f.c1.c1); fp2_mul(r.c1.c1, a, g);
    B2K_FROB_COEF(1, 4, g); fp2_conj(a, f.c0.c2); fp2_mul(r.c0.c2, a, g);
    B2K_FROB_COEF(1, 5, g); fp2_conj(a, f.c1.c2); fp2_mul(r.c1.c2, a, g);
  } else if (J == 2) {
    B2K_FROB_COEF(2, 1, g); fp2_mul(r.c1.c0, f.c1.c0, g);
    B2K_FROB_COEF(2, 2, g); fp2_mul(r.c0.c1, f.c0.c1, g);
    B2K_FROB_COEF(2, 3, g); fp2_mul(r.c1.c1, f.c1.c1, g);
    B2K_FROB_COEF(2, 4, g); fp2_mul(r.c0.c2, f.c0.c2, g);
    B2K_FROB_COEF(2, 5, g); fp2_mul(r.c1.c2, f.c1.c2, g);
  } else {
    B2K_FROB_COEF(3, 1, g); fp2_conj(a, f.c1.c0); fp2_mul(r.c1.c0, a, g);
    B2K_FROB_COEF(3, 2, g); fp2_conj(a, f.c0.c1); fp2_mul(r.c0.c1, a, g);
    B2K_FROB_COEF(3, 3, g); fp2_conj(a, f.c1.c1); fp2_mul(r.c1.c1, a, g);
    B2K_FROB_COEF(3, 4, g); fp2_conj(a, f.c0.c2); fp2_mul(r.c0.c2, a, g);
    B2K_FROB_COEF(3, 5, g); fp2_conj(a, f.c1.c2); fp2_mul(r.c1.c2, a, g);
  }
}

// (squaring in the cyclotomic subgroup: fp12_cyclotomic_sqr, tower.cuh)

// ---- exponentiation by a 64-bit public exponent inside the cyclotomic subgroup ---------------------------
B2K_NI void fp12_pow_u64(BFp12& r, const BFp12& a, uint64_t e) {
  BFp12 acc = a;
  int top = 63;
  while (top > 0 && !((e >> top) & 1)) top--;
  for (int b = top - 1; b >= 0; b--) {
    fp12_cyclotomic_sqr(acc, acc);
    if ((e >> b) & 1) fp12_mul(acc, acc, a);
  }
  r = acc;
}

// Final exponentiation.  GT convention = the one every BLS12-381 back-end of the reference shares and the
// reference's only GT-dependent vector pins (encrypt/ibe/ibe_test.go:202-245, decrypts only with this value and
// the byte order of gt_store): exponent 3 (p^12-1)/r, i.e. the standard pairing cubed.  With
//   3 (p^4-p^2+1)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3
// the hard part is five exponentiations by |x| (x < 0: inverses in the cyclotomic subgroup are conjugates).
B2K_D void fp12_pow_x(BFp12& r, const BFp12& a) {        // a^x, x = -|x|
  fp12_pow_u64(r, a, BLS_X_ABS);
  fp12_conj(r, r);
}

B2K_D void final_exponentiation(BFp12& r, const BFp12& f) {
  BFp12 m, t, a, b, c;
  fp12_inv(t, f);
  fp12_conj(m, f);
  fp12_mul(m, m, t);                      // f^(p^6-1)
  fp12_frobenius<2>(t, m);
  fp12_mul(m, t, m);                      // ^(p^2+1)
  fp12_pow_x(b, m); fp12_conj(t, m); fp12_mul(b, b, t);              // b = m^(x-1)
  fp12_pow_x(a, b); fp12_conj(t, b); fp12_mul(a, a, t);              // a = m^((x-1)^2)
  fp12_pow_x(c, a); fp12_frobenius<1>(t, a); fp12_mul(c, c, t);      // c = a^(x+p)
  fp12_pow_x(b, c); fp12_pow_x(a, b);                                // a = c^(x^2)
  fp12_frobenius<2>(t, c); fp12_mul(a, a, t);
  fp12_conj(t, c); fp12_mul(a, a, t);                                // a = c^(x^2+p^2-1)
  fp12_cyclotomic_sqr(t, m); fp12_mul(t, t, m);                      // m^3
  fp12_mul(r, a, t);
}

// The same with the plain exponent (p^12-1)/r (kept for cross-checks: final_exponentiation == this cubed).
//   hard part  (p^4-p^2+1)/r = l0 + l1 p + l2 p^2 + l3 p^3,  l3 = (x-1)^2/3, l2 = l3 x, l1 = l2 x - l3, l0 = l1 x + 1
B2K_D void final_exponentiation_exact(BFp12& r, const BFp12& f) {
  BFp12 m, t, y3, y2, y1, y0;
  fp12_inv(t, f);
  fp12_conj(m, f);
  fp12_mul(m, m, t);
  fp12_frobenius<2>(t, m);
  fp12_mul(m, t, m);
  fp12_pow_u64(t, m, BLS_E3);
  fp12_pow_u64(y3, t, BLS_X_ABS);
  fp12_mul(y3, y3, t);
  fp12_pow_u64(y2, y3, BLS_X_ABS); fp12_conj(y2, y2);
  fp12_pow_u64(y1, y2, BLS_X_ABS); fp12_conj(y1, y1);
  fp12_conj(t, y3); fp12_mul(y1, y1, t);
  fp12_pow_u64(y0, y1, BLS_X_ABS); fp12_conj(y0, y0);
  fp12_mul(y0, y0, m);
  fp12_frobenius<1>(t, y1); fp12_mul(y0, y0, t);
  fp12_frobenius<2>(t, y2); fp12_mul(y0, y0, t);
  fp12_frobenius<3>(t, y3); fp12_mul(r, y0, t);
}

// ---- codecs ----------------------------------------------------------------------------------------
// G2 operand: x.c1 || x.c0 || y.c1 || y.c0, 48 B big-endian canonical each; all-zero = infinity
B2K_D void g2_load(Affine<BFp2>& r, const uint8_t* p) {
  BFp t;
  fp_load_be(t, p); fp_to_mont(r.x.c1, t);
  fp_load_be(t, p + 48); fp_to_mont(r.x.c0, t);
  fp_load_be(t, p + 96); fp_to_mont(r.y.c1, t);
  fp_load_be(t, p + 144); fp_to_mont(r.y.c0, t);
}

// GT, 576 B: 12 x 48 B big-endian, highest tower coefficient first (c1.c2.c1 ... c0.c0.c0),
// the kilic layout (gt.go:115-117).  Layout AND exponent are pinned by the reference vector
// encrypt/ibe/ibe_test.go:202-245 (reproduced in tests/).
B2K_D void gt_store(uint8_t* out, const BFp12& f) {
  const BFp2* order[6] = {&f.c1.c2, &f.c1.c1, &f.c1.c0, &f.c0.c2, &f.c0.c1, &f.c0.c0};
  for (int i = 0; i < 6; i++) {
    BFp t;
    fp_from_mont(t, order[i]->c1); fp_store_be(out + 96 * i, t);
    fp_from_mont(t, order[i]->c0); fp_store_be(out + 96 * i + 48, t);
  }
}

}  // namespace b2k
