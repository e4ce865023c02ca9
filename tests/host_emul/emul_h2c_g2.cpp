// Host emulation of the hash-to-G2 device code (test infrastructure only).
#include "../../kyber_b200/csrc/h2c_g2.cuh"
using namespace b2k;
extern "C" void emul_bls12381_hash_to_g2(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dlen, uint8_t* out192) {
  Affine<BFp2> a;
  hash_to_g2(a, msg, len, dst, dlen);
  Bls381G2::store_affine(out192, a);
}
