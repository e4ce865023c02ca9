// Host emulation of the hash-to-G1 device code (test infrastructure only).
#include "../../kyber_b200/csrc/h2c.cuh"
using namespace b2k;
extern "C" {
void emul_bls12381_hash_to_g1(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dlen, uint8_t* out96) {
  Affine<BFp> a;
  hash_to_g1(a, msg, len, dst, dlen);
  Bls381G1::store_affine(out96, a);
}
void emul_expand_xmd_128(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dlen, uint8_t* out128) {
  expand_message_xmd_128(out128, msg, len, dst, dlen);
}
}
