// bn_pairing.cuh -- BN optimal-ate pairing for bn254 and its twin bn256 (D-type twist, signed-digit loop over 6u+2, two Frobenius
// steps, final exponentiation with the exact exponent (p^12-1)/n).
//
// Replaces the in-tree reference pairing:
//   miller                 pairing/bn254/optate.go:124-207   (lineFunctionDouble :54-94, lineFunctionAdd :5-52,
//                                                             mulLine :96-114, digit table :117-120)
//   finalExponentiation    pairing/bn254/optate.go:212-261   (easy part + the y0..y6 chain = exact exponent)
//   Suite.Pair / ValidatePairing   pairing/bn254/suite.go:133-144
//   GT MarshalBinary       pairing/bn254/point.go:625-656    (384 B, x.x.x first)
// Same kernel structure as the BLS12-381 pairing (pairing.cuh): Jacobian twist point over Fp2, lines scaled
// by Fp2 factors (killed by the final exponentiation), sparse line multiplication, cyclotomic squarings.
// D-type twist: psi(x',y') = (x' w^2, y' w^3), so a line at P is  l3 + l1 w + l0 w^3  with
//   doubling:  l0 = 3X^3 - 2Y^2,   l1 = -(3X^2 Z^2) xP,   l3 = (Z3 Z^2) yP
//   addition:  l0 = r x2 - Z3 y2,  l1 = -r xP,            l3 = Z3 yP
// i.e. the sparse Fp12 element  c0 = (l3, 0, 0),  c1 = (l1, l0, 0)  -- the reference's (a t + b) w + c with
// a ~ l0, b ~ l1, c ~ l3.  GT bytes equal those of a line-by-line restatement of the Go source (tests/).
#pragma once
#include "ec.cuh"
#include "curves.cuh"

namespace b2k {

// xi = 3 + u (pairing/bn256/gfp2.go:103-118):  (3 a0 - a1) + (3 a1 + a0) u
struct Bn256Tower {
  using Base = Bn256Fp;
  B2K_D static void mul_xi(Fp2<Base>& r, const Fp2<Base>& a) {
    Fp<Base> t0, t1, n0, n1;
    fp_add(t0, a.c0, a.c0); fp_add(t0, t0, a.c0);
    fp_add(t1, a.c1, a.c1); fp_add(t1, t1, a.c1);
    fp_sub(n0, t0, a.c1);
    fp_add(n1, t1, a.c0);
    r.c0 = n0; r.c1 = n1;
  }
};

// Pairing configurations: tower, u, digit table of 6u+2 (least significant first) and 32-byte codecs.
struct Bn254Pair {
  using T = Bn254Tower;
  using FC = Bn254Fp;
  static constexpr uint64_t U = 4965661367192848881ULL;
  static constexpr int NDIG = 65;
  B2K_D static int digit(int i) {   // pairing/bn254/optate.go:117-120 (data)
    const int8_t d[65] = {0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1,
                          1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1};
    return d[i];
  }
  B2K_D static void load32(Fp<FC>& r, const uint8_t* p) { Fp<FC> t; fp_load_be(t, p); fp_to_mont(r, t); }
  B2K_D static void store32(uint8_t* p, const Fp<FC>& a) { Fp<FC> t; fp_from_mont(t, a); fp_store_be(p, t); }
};

struct Bn256Pair {
  using T = Bn256Tower;
  using FC = Bn256Fp;
  static constexpr uint64_t U = 6518589491078791937ULL;
  static constexpr int NDIG = 66;
  B2K_D static int digit(int i) {   // pairing/bn256/optate.go:117-122 (data)
    const int8_t d[66] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, -1, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, -1, 0,
                          1, 0, 0, 0, 1, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 1};
    return d[i];
  }
  B2K_D static void load32(Fp<FC>& r, const uint8_t* p) {   // 32 bytes big-endian into 10 limbs
    Fp<FC> t;
    t.v[8] = 0; t.v[9] = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint8_t* q = p + 4 * (7 - j);
      t.v[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    fp_to_mont(r, t);
  }
  B2K_D static void store32(uint8_t* p, const Fp<FC>& a) {
    Fp<FC> t;
    fp_from_mont(t, a);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint8_t* q = p + 4 * (7 - j);
      q[0] = (uint8_t)(t.v[j] >> 24); q[1] = (uint8_t)(t.v[j] >> 16); q[2] = (uint8_t)(t.v[j] >> 8); q[3] = (uint8_t)t.v[j];
    }
  }
};

template <class P> using PFp = Fp<typename P::FC>;
template <class P> using PFp2 = Fp2<typename P::FC>;
template <class P> using PFp6 = Fp6<typename P::T>;
template <class P> using PFp12 = Fp12<typename P::T>;

// bn254 shorthands (used by the kernels of b2k_bn254_pairing.cu)
using NT = Bn254Tower;
using NFp = Fp<Bn254Fp>;
using NFp2 = Fp2<Bn254Fp>;
using NFp6 = Fp6<NT>;
using NFp12 = Fp12<NT>;

template <class P> struct BnLine { PFp2<P> l0, l1, l3; };

// f *= l3 + (l1 + l0 t) w
template <class P>
B2K_NI void fp12_mul_line_d(PFp12<P>& f, const BnLine<P>& l) {
  PFp6<P> A, B, C;
  PFp2<P> s;
  fp6_mul_by_01(A, f.c1, l.l1, l.l0);            // f1 * (l1 + l0 t)
  fp6_mul_fp2(B, f.c0, l.l3);                    // f0 * l3
  fp2_add(s, l.l1, l.l3);
  fp6_add(C, f.c0, f.c1);
  fp6_mul_by_01(C, C, s, l.l0);                  // (f0 + f1)(l3 + l1 + l0 t)
  fp6_sub(C, C, A);
  fp6_sub(f.c1, C, B);
  fp6_mul_v(A, A);
  fp6_add(f.c0, B, A);
}

template <class PC>
B2K_NI void bn_double_step(BnLine<PC>& l, Jac<PFp2<PC>>& T, const Affine<PFp<PC>>& P) {
  PFp2<PC> A, B, C, D, E, ZZ, t;
  fp2_sqr(A, T.X);
  fp2_sqr(B, T.Y);
  fp2_sqr(C, B);
  fp2_sqr(ZZ, T.Z);
  fp2_add(D, T.X, B); fp2_sqr(D, D); fp2_sub(D, D, A); fp2_sub(D, D, C); fp2_dbl(D, D);
  fp2_dbl(E, A); fp2_add(E, E, A);
  fp2_mul(l.l0, E, T.X); fp2_sub(l.l0, l.l0, B); fp2_sub(l.l0, l.l0, B);
  fp2_mul(t, E, ZZ); fp2_mul_fp(t, t, P.x); fp2_neg(l.l1, t);
  fp2_mul(t, T.Y, T.Z); fp2_dbl(T.Z, t);
  fp2_mul(t, T.Z, ZZ); fp2_mul_fp(l.l3, t, P.y);
  fp2_sqr(A, E); fp2_sub(A, A, D); fp2_sub(A, A, D);
  fp2_dbl(C, C); fp2_dbl(C, C); fp2_dbl(C, C);
  fp2_sub(D, D, A); fp2_mul(D, E, D); fp2_sub(T.Y, D, C);
  T.X = A;
}

template <class PC>
B2K_NI void bn_add_step(BnLine<PC>& l, Jac<PFp2<PC>>& T, const Affine<PFp2<PC>>& Q, const Affine<PFp<PC>>& P) {
  PFp2<PC> ZZ, U2, S2, H, R, HH, HHH, V, t;
  fp2_sqr(ZZ, T.Z);
  fp2_mul(U2, Q.x, ZZ);
  fp2_mul(S2, Q.y, T.Z); fp2_mul(S2, S2, ZZ);
  fp2_sub(H, U2, T.X);
  fp2_sub(R, S2, T.Y);
  fp2_sqr(HH, H);
  fp2_mul(HHH, H, HH);
  fp2_mul(V, T.X, HH);
  fp2_mul(T.Z, T.Z, H);
  fp2_mul(l.l0, R, Q.x); fp2_mul(t, T.Z, Q.y); fp2_sub(l.l0, l.l0, t);
  fp2_mul_fp(t, R, P.x); fp2_neg(l.l1, t);
  fp2_mul_fp(l.l3, T.Z, P.y);
  fp2_sqr(t, R); fp2_sub(t, t, HHH); fp2_sub(t, t, V); fp2_sub(t, t, V);
  fp2_sub(V, V, t); fp2_mul(V, R, V);
  fp2_mul(HHH, T.Y, HHH);
  fp2_sub(T.Y, V, HHH);
  T.X = t;
}

#define B2K_BN_COEF(J, K, dst)                                                       \
  {                                                                                  \
    _Pragma("unroll") for (int q = 0; q < PC::FC::N; q++) {                          \
      (dst).c0.v[q] = PC::FC::frob##J##_##K##_c0(q);                                 \
      (dst).c1.v[q] = PC::FC::frob##J##_##K##_c1(q);                                 \
    }                                                                                \
  }

// f = prod_i f_{6u+2,Q_i}(P_i) * l_{T,pi(Q_i)} * l_{T+pi(Q_i), -pi^2(Q_i)}; pairs with an infinity member give 1
template <class PC, int NPAIRS>
B2K_D void bn_miller_loop(PFp12<PC>& f, const Affine<PFp<PC>>* P, const Affine<PFp2<PC>>* Q) {
  using NFp2 = PFp2<PC>;
  Jac<NFp2> T[NPAIRS];
  Affine<NFp2> Qn[NPAIRS];
  bool live[NPAIRS];
#pragma unroll
  for (int i = 0; i < NPAIRS; i++) {
    live[i] = !(aff_is_inf(P[i]) || aff_is_inf(Q[i]));
    T[i].X = Q[i].x; T[i].Y = Q[i].y; fp2_set_one(T[i].Z);
    Qn[i].x = Q[i].x; fp2_neg(Qn[i].y, Q[i].y);
  }
  fp12_set_one(f);
  BnLine<PC> l;
  for (int i = PC::NDIG - 1; i > 0; i--) {
    if (i != PC::NDIG - 1) fp12_sqr(f, f);
#pragma unroll
    for (int k = 0; k < NPAIRS; k++) {
      if (!live[k]) continue;
      bn_double_step<PC>(l, T[k], P[k]);
      fp12_mul_line_d<PC>(f, l);
    }
    const int d = PC::digit(i - 1);
    if (d == 0) continue;
#pragma unroll
    for (int k = 0; k < NPAIRS; k++) {
      if (!live[k]) continue;
      bn_add_step<PC>(l, T[k], d > 0 ? Q[k] : Qn[k], P[k]);
      fp12_mul_line_d<PC>(f, l);
    }
  }
  // Frobenius steps: Q1 = (conj(x) xi^((p-1)/3), conj(y) xi^((p-1)/2)),  -Q2 = (x xi^((p^2-1)/3), y)
  NFp2 g13, g12, g23;
  B2K_BN_COEF(1, 2, g13);
  B2K_BN_COEF(1, 3, g12);
  B2K_BN_COEF(2, 2, g23);
#pragma unroll
  for (int k = 0; k < NPAIRS; k++) {
    if (!live[k]) continue;
    Affine<NFp2> q1, mq2;
    NFp2 c;
    fp2_conj(c, Q[k].x); fp2_mul(q1.x, c, g13);
    fp2_conj(c, Q[k].y); fp2_mul(q1.y, c, g12);
    fp2_mul(mq2.x, Q[k].x, g23);
    mq2.y = Q[k].y;
    bn_add_step<PC>(l, T[k], q1, P[k]);
    fp12_mul_line_d<PC>(f, l);
    bn_add_step<PC>(l, T[k], mq2, P[k]);
    fp12_mul_line_d<PC>(f, l);
  }
}

// f^(p^J) on the w-power basis (slots: w^0 c0.c0, w^1 c1.c0, w^2 c0.c1, w^3 c1.c1, w^4 c0.c2, w^5 c1.c2)
template <class PC, int J>
B2K_NI void bn_frobenius(PFp12<PC>& r, const PFp12<PC>& f) {
  PFp2<PC> g, a;
  a = f.c0.c0; if (J & 1) fp2_conj(a, a); r.c0.c0 = a;
#define B2K_BN_FROB_SLOT(K, slot)                                 \
  B2K_BN_COEF_SEL(K, g);                                          \
  a = f.slot; if (J & 1) fp2_conj(a, a); fp2_mul(r.slot, a, g);
#define B2K_BN_COEF_SEL(K, dst)                                   \
  if (J == 1) B2K_BN_COEF(1, K, dst) else if (J == 2) B2K_BN_COEF(2, K, dst) else B2K_BN_COEF(3, K, dst)
  B2K_BN_FROB_SLOT(1, c1.c0)
  B2K_BN_FROB_SLOT(2, c0.c1)
  B2K_BN_FROB_SLOT(3, c1.c1)
  B2K_BN_FROB_SLOT(4, c0.c2)
  B2K_BN_FROB_SLOT(5, c1.c2)
#undef B2K_BN_FROB_SLOT
#undef B2K_BN_COEF_SEL
}

// Granger-Scott squaring in the cyclotomic subgroup: fp12_cyclotomic_sqr (tower.cuh)
template <class PC>
B2K_NI void bn_pow_u(PFp12<PC>& r, const PFp12<PC>& a) {
  PFp12<PC> acc = a;
  for (int b = 61; b >= 0; b--) {            // u has 63 bits (both curves), top bit consumed by acc = a
    fp12_cyclotomic_sqr(acc, acc);
    if ((PC::U >> b) & 1) fp12_mul(acc, acc, a);
  }
  r = acc;
}

// easy part (p^6-1)(p^2+1), then the hard part  y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36  (optate.go:212-261)
template <class PC>
B2K_D void bn_final_exponentiation(PFp12<PC>& r, const PFp12<PC>& in) {
  PFp12<PC> t0, t1, fp1, fp2v, fp3, fu, fu2, fu3, y0, y2, y3, y4, y6, tmp;
  fp12_inv(t0, in);
  fp12_conj(t1, in);
  fp12_mul(t1, t1, t0);
  bn_frobenius<PC, 2>(t0, t1);
  fp12_mul(t1, t1, t0);                               // t1 = f^((p^6-1)(p^2+1))
  bn_frobenius<PC, 1>(fp1, t1);
  bn_frobenius<PC, 2>(fp2v, t1);
  bn_frobenius<PC, 1>(fp3, fp2v);
  bn_pow_u<PC>(fu, t1);
  bn_pow_u<PC>(fu2, fu);
  bn_pow_u<PC>(fu3, fu2);
  bn_frobenius<PC, 1>(y3, fu);   fp12_conj(y3, y3);
  bn_frobenius<PC, 1>(tmp, fu2); fp12_mul(y4, fu, tmp); fp12_conj(y4, y4);
  bn_frobenius<PC, 1>(tmp, fu3); fp12_mul(y6, fu3, tmp); fp12_conj(y6, y6);
  bn_frobenius<PC, 2>(y2, fu2);
  fp12_mul(y0, fp1, fp2v); fp12_mul(y0, y0, fp3);
  // y1 = conj(t1), y5 = conj(fu2)
  fp12_sqr(t0, y6); fp12_mul(t0, t0, y4); fp12_conj(tmp, fu2); fp12_mul(t0, t0, tmp);      // t0 = y6^2 y4 y5
  PFp12<PC> s;
  fp12_mul(s, y3, tmp); fp12_mul(s, s, t0);                                                // s = y3 y5 t0
  fp12_mul(t0, t0, y2);
  fp12_sqr(s, s); fp12_mul(s, s, t0); fp12_sqr(s, s);
  fp12_conj(tmp, t1);
  fp12_mul(t0, s, tmp);                                                                    // t0 = s y1
  fp12_mul(s, s, y0);
  fp12_sqr(t0, t0);
  fp12_mul(r, t0, s);
}

// ---- codecs (32-byte big-endian field elements; G2 imaginary part first; GT 384 B highest coefficient first) ------
template <class PC>
B2K_D void bn_g1_load(Affine<PFp<PC>>& r, const uint8_t* p) { PC::load32(r.x, p); PC::load32(r.y, p + 32); }
template <class PC>
B2K_D void bn_g2_load(Affine<PFp2<PC>>& r, const uint8_t* p) {
  PC::load32(r.x.c1, p); PC::load32(r.x.c0, p + 32); PC::load32(r.y.c1, p + 64); PC::load32(r.y.c0, p + 96);
}
template <class PC>
B2K_D void bn_gt_store(uint8_t* out, const PFp12<PC>& f) {
  const PFp2<PC>* order[6] = {&f.c1.c2, &f.c1.c1, &f.c1.c0, &f.c0.c2, &f.c0.c1, &f.c0.c0};
  for (int i = 0; i < 6; i++) { PC::store32(out + 64 * i, order[i]->c1); PC::store32(out + 64 * i + 32, order[i]->c0); }
}

// bn254 names kept for the existing call sites
template <int NPAIRS> B2K_D void bn254_miller_loop(NFp12& f, const Affine<NFp>* P, const Affine<NFp2>* Q) { bn_miller_loop<Bn254Pair, NPAIRS>(f, P, Q); }
B2K_D void bn254_final_exponentiation(NFp12& r, const NFp12& in) { bn_final_exponentiation<Bn254Pair>(r, in); }
B2K_D void bn254_g1_load(Affine<NFp>& r, const uint8_t* p) { bn_g1_load<Bn254Pair>(r, p); }
B2K_D void bn254_g2_load(Affine<NFp2>& r, const uint8_t* p) { bn_g2_load<Bn254Pair>(r, p); }
B2K_D void bn254_gt_store(uint8_t* out, const NFp12& f) { bn_gt_store<Bn254Pair>(out, f); }

// bn254 G2 (pairing/bn254/twist.go:167-181 twistPoint.Mul; wire format point.go:428-455)
struct Bn254G2 {
  using FC = Bn254Fp;
  using F = NFp2;
  using ScalarField = Bn254Fr;
  static constexpr int SCALAR_BITS = 254;
  static constexpr int IN_BYTES = 128;
  static constexpr int OUT_BYTES = 128;
  B2K_D static void load(Affine<F>& r, const uint8_t* p) { bn254_g2_load(r, p); }
  B2K_D static bool wire_canonical(const uint8_t* p) { return wire_coords_canonical<FC, 4>(p); }
  B2K_D static void curve_b(F& b) {
#pragma unroll
    for (int j = 0; j < 8; j++) { b.c0.v[j] = FC::twist_b_c0(j); b.c1.v[j] = FC::twist_b_c1(j); }
  }
  B2K_D static void store(uint8_t* out, const Affine<F>& p) {
    NFp t;
    fp_from_mont(t, p.x.c1); fp_store_be(out, t);
    fp_from_mont(t, p.x.c0); fp_store_be(out + 32, t);
    fp_from_mont(t, p.y.c1); fp_store_be(out + 64, t);
    fp_from_mont(t, p.y.c0); fp_store_be(out + 96, t);
  }
  B2K_D static void store_affine(uint8_t* out, const Affine<F>& p) { store(out, p); }
};

}  // namespace b2k
