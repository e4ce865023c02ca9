// Host emulation of the variable-time inversion (test infrastructure only).
#include "../../kyber_b200/csrc/constants.cuh"
#include "../../kyber_b200/csrc/fp_inv.cuh"
using namespace b2k;
extern "C" {
void emul_fp381_inv_vartime(const uint32_t* a, uint32_t* r) { Fp<Bls381Fp> x, z; for (int i = 0; i < 12; i++) x.v[i] = a[i]; fp_inv_vartime(z, x); for (int i = 0; i < 12; i++) r[i] = z.v[i]; }
void emul_fp254_inv_vartime(const uint32_t* a, uint32_t* r) { Fp<Bn254Fp> x, z; for (int i = 0; i < 8; i++) x.v[i] = a[i]; fp_inv_vartime(z, x); for (int i = 0; i < 8; i++) r[i] = z.v[i]; }
void emul_fp256_inv_vartime(const uint32_t* a, uint32_t* r) { Fp<Bn256Fp> x, z; for (int i = 0; i < 10; i++) x.v[i] = a[i]; fp_inv_vartime(z, x); for (int i = 0; i < 10; i++) r[i] = z.v[i]; }
}
// branch-free binary GCD on approximations (the inversion of the MSM's affine pair-tree rounds)
#define BINGCD_API(name, C)                                                                        \
  void emul_##name##_inv_bingcd(const uint32_t* a, uint32_t* r) {                                  \
    Fp<C> x, z; for (int i = 0; i < C::N; i++) x.v[i] = a[i];                                      \
    fp_inv_bingcd(z, x); for (int i = 0; i < C::N; i++) r[i] = z.v[i]; }
extern "C" {
BINGCD_API(fp381, Bls381Fp)
BINGCD_API(fp254, Bn254Fp)
BINGCD_API(fp256, Bn256Fp)
BINGCD_API(fp25519, Ed25519Fp)
}
