// bn256.cuh -- curve configs for bn256 (the Cloudflare BN curve of pairing/bn256), G1 and G2.
//
// Replaces: bn256 curvePoint.Mul (pairing/bn256/curve.go:189-203), twistPoint.Mul (twist.go:162-175) and the
// Mul+Add loops of sign/bdn on this curve; wire formats of pairing/bn256/point.go:170-192 (G1: x||y, 32-byte
// big-endian each) and :423-452 (G2: x.imag||x.real||y.imag||y.real), infinity = all zeros.
// The base prime is 256 bits (> 2^255), so field elements use 10 limbs (R = 2^320): a+b and the Montgomery
// accumulator then stay inside the limb vector with the same code as the other curves.  The byte-exact BDN
// fixtures of the reference (sign/bdn/bdn_vartime_test.go:24-48, :90-135) run on these kernels.
#pragma once
#include "curves.cuh"

namespace b2k {

using B256Fp = Fp<Bn256Fp>;
using B256Fp2 = Fp2<Bn256Fp>;

B2K_D void bn256_load32(B256Fp& r, const uint8_t* p) {     // 32 bytes big-endian -> Montgomery (10 limbs)
  B256Fp t;
  t.v[8] = 0; t.v[9] = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint8_t* q = p + 4 * (7 - j);
    t.v[j] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
  }
  fp_to_mont(r, t);
}

B2K_D void bn256_store32(uint8_t* p, const B256Fp& a) {     // Montgomery -> 32 bytes big-endian canonical
  B256Fp t;
  fp_from_mont(t, a);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    uint8_t* q = p + 4 * (7 - j);
    q[0] = (uint8_t)(t.v[j] >> 24); q[1] = (uint8_t)(t.v[j] >> 16); q[2] = (uint8_t)(t.v[j] >> 8); q[3] = (uint8_t)t.v[j];
  }
}

struct Bn256G1 {
  using FC = Bn256Fp;
  using F = B256Fp;
  using ScalarField = Bn256Fr;
  static constexpr int SCALAR_BITS = 256;
  static constexpr int IN_BYTES = 64;
  static constexpr int OUT_BYTES = 64;
  B2K_D static void load(Affine<F>& r, const uint8_t* p) { bn256_load32(r.x, p); bn256_load32(r.y, p + 32); }
  B2K_D static bool wire_canonical(const uint8_t*) { return true; }      // bn256 reduces on Unmarshal (pairing/bn256/gfp.go:115-122)
  B2K_D static void curve_b(F& b) {
#pragma unroll
    for (int j = 0; j < 10; j++) b.v[j] = FC::curve_b(j);
  }
  B2K_D static void store(uint8_t* out, const Affine<F>& p) { bn256_store32(out, p.x); bn256_store32(out + 32, p.y); }
  B2K_D static void store_affine(uint8_t* out, const Affine<F>& p) { store(out, p); }
};

struct Bn256G2 {
  using FC = Bn256Fp;
  using F = B256Fp2;
  using ScalarField = Bn256Fr;
  static constexpr int SCALAR_BITS = 256;
  static constexpr int IN_BYTES = 128;
  static constexpr int OUT_BYTES = 128;
  // gfP2{x, y} = x*i + y (pairing/bn256/gfp2.go:13-15): imaginary part first on the wire
  B2K_D static bool wire_canonical(const uint8_t*) { return true; }
  B2K_D static void curve_b(F& b) {
#pragma unroll
    for (int j = 0; j < 10; j++) { b.c0.v[j] = FC::twist_b_c0(j); b.c1.v[j] = FC::twist_b_c1(j); }
  }
  B2K_D static void load(Affine<F>& r, const uint8_t* p) {
    bn256_load32(r.x.c1, p); bn256_load32(r.x.c0, p + 32);
    bn256_load32(r.y.c1, p + 64); bn256_load32(r.y.c0, p + 96);
  }
  B2K_D static void store(uint8_t* out, const Affine<F>& p) {
    bn256_store32(out, p.x.c1); bn256_store32(out + 32, p.x.c0);
    bn256_store32(out + 64, p.y.c1); bn256_store32(out + 96, p.y.c0);
  }
  B2K_D static void store_affine(uint8_t* out, const Affine<F>& p) { store(out, p); }
};

}  // namespace b2k
