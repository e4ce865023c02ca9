// h2c.cuh -- RFC 9380 hash_to_curve for BLS12-381 G1 (suite BLS12381G1_XMD:SHA-256_SSWU_RO_), per thread.
//
// Replaces: kilic.G1Elt.Hash -> third-party HashToCurve(msg, dst)   pairing/bls12381/kilic/g1.go:161-170
// (default DST "BLS_SIG_BLS12381G1_XMD:SHA-256_SSWU_RO_NUL_", g1.go:17; per-suite override suite.go:32-46),
// the per-message step of bls.Sign/Verify (sign/bls/bls.go:67-96).
// Pipeline: expand_message_xmd(SHA-256) -> 2 field elements -> simplified SWU on E' -> derived 11-isogeny
// (tools/derive_isogeny.py) evaluated projectively -> add -> clear cofactor by h_eff = 1 - x.
// Pinned end-to-end by the reference's signature KATs (tests/test_gpu_bls12381_h2c.py).
#pragma once
#include "codec.cuh"
#include "iso_g1_constants.cuh"

namespace b2k {

// ---- SHA-256 (FIPS 180-4), streaming over byte fragments -----------------------------------------------
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint32_t fill;
  uint64_t total;
};

B2K_D uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

B2K_NI void sha256_compress(uint32_t* h, const uint8_t* blk) {
  const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
      0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
      0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
      0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
      0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
      0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
      0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t w[64];
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = sha_rotr(w[i - 15], 7) ^ sha_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = sha_rotr(w[i - 2], 17) ^ sha_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + K[i] + w[i];
    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

B2K_D void sha256_init(Sha256& s) {
  const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  for (int i = 0; i < 8; i++) s.h[i] = iv[i];
  s.fill = 0;
  s.total = 0;
}

B2K_D void sha256_update(Sha256& s, const uint8_t* p, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) {
    s.buf[s.fill++] = p[i];
    if (s.fill == 64) { sha256_compress(s.h, s.buf); s.fill = 0; }
  }
  s.total += n;
}

B2K_D void sha256_update_zero(Sha256& s, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) {
    s.buf[s.fill++] = 0;
    if (s.fill == 64) { sha256_compress(s.h, s.buf); s.fill = 0; }
  }
  s.total += n;
}

B2K_D void sha256_final(Sha256& s, uint8_t* out32) {
  uint64_t bits = s.total * 8;
  s.buf[s.fill++] = 0x80;
  if (s.fill > 56) {
    while (s.fill < 64) s.buf[s.fill++] = 0;
    sha256_compress(s.h, s.buf);
    s.fill = 0;
  }
  while (s.fill < 56) s.buf[s.fill++] = 0;
  for (int i = 0; i < 8; i++) s.buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
  sha256_compress(s.h, s.buf);
  for (int i = 0; i < 8; i++) {
    out32[4 * i] = (uint8_t)(s.h[i] >> 24); out32[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
    out32[4 * i + 2] = (uint8_t)(s.h[i] >> 8); out32[4 * i + 3] = (uint8_t)s.h[i];
  }
}

// ---- expand_message_xmd, 128 output bytes (two field elements of L = 64 bytes) ----------------------------
// dst must be <= 255 bytes (longer DSTs are pre-hashed by the caller, RFC 9380 5.3.3)
B2K_NI void expand_message_xmd_128(uint8_t* out128, const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
  Sha256 s;
  uint8_t b0[32], bi[32], t[32];
  const uint8_t dl = (uint8_t)dst_len;
  const uint8_t lib[3] = {0, 128, 0};               // I2OSP(128, 2) || I2OSP(0, 1)
  sha256_init(s);
  sha256_update_zero(s, 64);
  sha256_update(s, msg, msg_len);
  sha256_update(s, lib, 3);
  sha256_update(s, dst, dst_len);
  sha256_update(s, &dl, 1);
  sha256_final(s, b0);
  for (int i = 1; i <= 4; i++) {
    const uint8_t idx = (uint8_t)i;
    if (i == 1) { for (int k = 0; k < 32; k++) t[k] = b0[k]; }
    else { for (int k = 0; k < 32; k++) t[k] = b0[k] ^ bi[k]; }
    sha256_init(s);
    sha256_update(s, t, 32);
    sha256_update(s, &idx, 1);
    sha256_update(s, dst, dst_len);
    sha256_update(s, &dl, 1);
    sha256_final(s, bi);
    for (int k = 0; k < 32; k++) out128[32 * (i - 1) + k] = bi[k];
  }
}

B2K_D void fp_load_table(BFp& r, const uint32_t* t) {
#pragma unroll
  for (int j = 0; j < 12; j++) r.v[j] = t[j];
}

// 64 big-endian bytes -> value mod p, Montgomery form:  hi * 2^256 + lo
B2K_D void fp_from_64_bytes(BFp& r, const uint8_t* b) {
  BFp hi, lo, c;
  fp_set_zero(hi); fp_set_zero(lo);
  for (int j = 0; j < 8; j++) {
    const uint8_t* q = b + 4 * (7 - j);
    hi.v[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    q = b + 32 + 4 * (7 - j);
    lo.v[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
  }
  fp_to_mont(hi, hi);
  fp_to_mont(lo, lo);
  fp_load_table(c, TWO_POW_256);
  fp_mul_c(hi, hi, c);
  fp_add(r, hi, lo);
}

B2K_D uint32_t fp_sgn0(const BFp& a_mont) {
  BFp c;
  fp_from_mont(c, a_mont);
  return c.v[0] & 1u;
}

// simplified SWU map onto E': y^2 = x^3 + A'x + B'  (RFC 9380 6.6.2, straight-line version with one inversion)
B2K_NI void map_to_curve_sswu(BFp& x, BFp& y, const BFp& u) {
  BFp A, B, Zc, t, tv1, x1, gx, u2, c;
  fp_load_table(A, SSWU_A); fp_load_table(B, SSWU_B); fp_load_table(Zc, SSWU_Z);
  fp_sqr_c(u2, u);
  fp_mul_c(t, Zc, u2);                  // Z u^2
  fp_sqr_c(tv1, t); fp_add(tv1, tv1, t);   // Z^2 u^4 + Z u^2
  if (fp_is_zero(tv1)) {
    fp_load_table(x1, SSWU_B_OVER_ZA);
  } else {
    fp_inv(tv1, tv1);
    fp_set_one(c); fp_add(tv1, tv1, c);
    fp_load_table(c, SSWU_NEG_B_OVER_A);
    fp_mul_c(x1, c, tv1);
  }
  fp_sqr_c(gx, x1); fp_add(gx, gx, A); fp_mul_c(gx, gx, x1); fp_add(gx, gx, B);   // x1^3 + A x1 + B
  BFp yy;
  if (fp_sqrt(yy, gx)) {
    x = x1;
  } else {
    fp_mul_c(x, t, x1);                 // x2 = Z u^2 x1
    fp_sqr_c(gx, x); fp_add(gx, gx, A); fp_mul_c(gx, gx, x); fp_add(gx, gx, B);
    fp_sqrt(yy, gx);                    // always a square when gx1 is not
  }
  if (fp_sgn0(u) != fp_sgn0(yy)) fp_neg(yy, yy);
  y = yy;
}

B2K_D void horner(BFp& r, const uint32_t (*coef)[12], int n, const BFp& x) {
  BFp acc, c;
  fp_load_table(acc, coef[n - 1]);
  for (int k = n - 2; k >= 0; k--) {
    fp_mul_c(acc, acc, x);
    fp_load_table(c, coef[k]);
    fp_add(acc, acc, c);
  }
  r = acc;
}

// iso_map(x', y') = (xn/xd, y' yn/yd) returned as a Jacobian point (Z = xd yd), no inversion
B2K_NI void iso_map_g1(Jac<BFp>& r, const BFp& x, const BFp& y) {
  BFp xn, xd, yn, yd, t;
  horner(xn, ISO_G1_XNUM, 12, x);
  horner(xd, ISO_G1_XDEN, 11, x);
  horner(yn, ISO_G1_YNUM, 16, x);
  horner(yd, ISO_G1_YDEN, 16, x);
  fp_mul_c(r.Z, xd, yd);
  if (fp_is_zero(r.Z)) { jac_set_inf(r); return; }      // kernel of the isogeny
  fp_sqr_c(t, yd);                                      // yd^2
  fp_mul_c(r.X, xn, xd); fp_mul_c(r.X, r.X, t);         // xn xd yd^2
  fp_mul_c(r.Y, y, yn); fp_mul_c(r.Y, r.Y, t);          // y yn yd^2
  fp_sqr_c(t, xd); fp_mul_c(t, t, xd);                  // xd^3
  fp_mul_c(r.Y, r.Y, t);
}

// full hash_to_curve: result affine (Montgomery) in G1
B2K_D void hash_to_g1(Affine<BFp>& out, const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
  uint8_t uni[128];
  expand_message_xmd_128(uni, msg, msg_len, dst, dst_len);
  BFp u0, u1, x, y;
  fp_from_64_bytes(u0, uni);
  fp_from_64_bytes(u1, uni + 64);
  Jac<BFp> q0, q1, acc;
  map_to_curve_sswu(x, y, u0); iso_map_g1(q0, x, y);
  map_to_curve_sswu(x, y, u1); iso_map_g1(q1, x, y);
  jac_add(q0, q0, q1);
  // clear cofactor: [h_eff] R, h_eff = |x| + 1 = 0xd201000000010001
  acc = q0;
  for (int b = 62; b >= 0; b--) {
    jac_dbl(acc, acc);
    if ((BLS_X_ABS >> b) & 1) jac_add(acc, acc, q0);
  }
  jac_add(acc, acc, q0);
  jac_to_affine(out, acc);
}

}  // namespace b2k
