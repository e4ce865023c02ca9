// Host emulation of the BN hash-to-G1 device code (test infrastructure only).
#include "../../kyber_b200/csrc/bn_hash.cuh"
using namespace b2k;
extern "C" {
void emul_keccak256(const uint8_t* msg, uint32_t len, uint8_t* out32) {
  Keccak256 k;
  keccak_init(k);
  keccak_update(k, msg, len);
  keccak_final(k, out32);
}
void emul_bn254_expand_96(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len, uint8_t* out96) {
  expand_message_xmd_keccak_96(out96, msg, len, dst, dst_len);
}
// the two field elements of hashToField, canonical 32-byte big-endian each
void emul_bn254_hash_to_field(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len, uint8_t* out64) {
  uint8_t uni[96];
  expand_message_xmd_keccak_96(uni, msg, len, dst, dst_len);
  NFp254 u, c;
  for (int k = 0; k < 2; k++) {
    bn254_fp_from_48_bytes(u, uni + 48 * k);
    fp_from_mont(c, u);
    fp_store_be(out64 + 32 * k, c);
  }
}
void emul_bn254_map_to_point(const uint8_t* u32be, uint8_t* out64) {
  NFp254 t, u;
  fp_load_be(t, u32be);
  fp_to_mont(u, t);
  Affine<NFp254> p;
  bn254_map_to_point(p, u);
  Bn254G1::store(out64, p);
}
void emul_bn254_hash_to_g1(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len, uint8_t* out64) {
  Affine<NFp254> p;
  bn254_hash_to_g1(p, msg, len, dst, dst_len);
  Bn254G1::store(out64, p);
}
void emul_bn256_hash_g1(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len, uint8_t* out64) {
  Affine<B256Fp> p;
  bn256_hash_g1(p, msg, len, dst, dst_len);
  Bn256G1::store(out64, p);
}
void emul_hmac_sha256(const uint8_t* key, uint32_t key_len, const uint8_t* msg, uint32_t len, uint8_t* out32) {
  hmac_sha256(out32, key, key_len, msg, len, nullptr, 0, nullptr, 0);
}
void emul_bn256_hash_to_g1(const uint8_t* msg, uint32_t len, uint8_t* out64) {
  Affine<B256Fp> p;
  bn256_hash_to_g1(p, msg, len);
  Bn256G1::store(out64, p);
}
}
