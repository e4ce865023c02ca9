// host build of tools/probe/fp_dfma.cuh / ec_dfma.cuh for tests/test_dfma_model.py: limbs in/out as uint64 (52 bits each)
#include "ec_dfma.cuh"

namespace {
struct Rz {
  int old;
  Rz() : old(std::fegetround()) { std::fesetround(FE_TOWARDZERO); }
  ~Rz() { std::fesetround(old); }
};
dfma::Fp load(const uint64_t* a) { dfma::Fp x; for (int i = 0; i < dfma::L; i++) x.v[i] = (double)a[i]; return x; }
void store(uint64_t* o, const dfma::Fp& x) { for (int i = 0; i < dfma::L; i++) o[i] = (uint64_t)x.v[i]; }
}  // namespace

extern "C" {
int dfma_mont_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
  Rz rz;
  dfma::Fp r;
  dfma::mont_mul(r, load(a), load(b));
  store(out, r);
  return 0;
}
// library form in and out: 12 x 32-bit limbs of x 2^384 mod p
int dfma_mont_mul384_u32(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Rz rz;
  dfma::Fp x, y, r;
  dfma::from_u32(x, a); dfma::from_u32(y, b);
  dfma::mont_mul384(r, x, y);
  dfma::to_u32(out, r);
  return 0;
}
int dfma_add_sub(const uint64_t* a, const uint64_t* b, uint64_t* sum, uint64_t* diff) {
  Rz rz;
  dfma::Fp r;
  dfma::fp_add(r, load(a), load(b)); store(sum, r);
  dfma::fp_sub(r, load(a), load(b)); store(diff, r);
  return 0;
}
// acc = [X, Y, ZZ, ZZZ] (4 x 8 limbs, in/out), p = [x, y] (2 x 8 limbs), one = R mod p; returns xyzz_madd's case code
int dfma_xyzz_madd(uint64_t* acc, const uint64_t* p, const uint64_t* one) {
  Rz rz;
  dfma::Xyzz a{load(acc), load(acc + 8), load(acc + 16), load(acc + 24)};
  dfma::Affine q{load(p), load(p + 8)};
  const int rc = dfma::xyzz_madd<false>(a, q, load(one));
  store(acc, a.X); store(acc + 8, a.Y); store(acc + 16, a.ZZ); store(acc + 24, a.ZZZ);
  return rc;
}
// the same on LIBRARY values: acc = 4 x 12 u32 limbs (Xyzz<Fp> of ec.cuh, Montgomery radix 2^384), p = 2 x 12, one = 2^384 mod p
int dfma_xyzz_madd384_u32(uint32_t* acc, const uint32_t* p, const uint32_t* one) {
  Rz rz;
  dfma::Xyzz a;
  dfma::Affine q;
  dfma::Fp o;
  dfma::from_u32(a.X, acc); dfma::from_u32(a.Y, acc + 12); dfma::from_u32(a.ZZ, acc + 24); dfma::from_u32(a.ZZZ, acc + 36);
  dfma::from_u32(q.x, p); dfma::from_u32(q.y, p + 12);
  dfma::from_u32(o, one);
  const int rc = dfma::xyzz_madd<true>(a, q, o);
  dfma::to_u32(acc, a.X); dfma::to_u32(acc + 12, a.Y); dfma::to_u32(acc + 24, a.ZZ); dfma::to_u32(acc + 36, a.ZZZ);
  return rc;
}
}
