// Host emulation of the edwards25519 device code (test infrastructure only).
#include "../../kyber_b200/csrc/ed25519.cuh"
using namespace b2k;
extern "C" int emul_ed25519_mul(const uint8_t* k, const uint8_t* pt, uint8_t* out) {
  EdExt p, r;
  if (!ed_decode(p, pt)) return 0;
  ed_scalar_mul(r, k, p);
  ed_encode(out, r);
  return 1;
}
