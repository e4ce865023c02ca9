// host build of tools/probe/fp_dfma.cuh for tests/test_dfma_model.py: limbs in/out as uint64 (52 bits each)
#include "fp_dfma.cuh"
extern "C" int dfma_mont_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
  const int old = std::fegetround();
  std::fesetround(FE_TOWARDZERO);
  dfma::Fp x, y, r;
  for (int i = 0; i < dfma::L; i++) { x.v[i] = (double)a[i]; y.v[i] = (double)b[i]; }
  dfma::mont_mul(r, x, y);
  for (int i = 0; i < dfma::L; i++) out[i] = (uint64_t)r.v[i];
  std::fesetround(old);
  return 0;
}
