// b2k_bdn.cu -- the host-side half of sign/bdn: the rogue-key coefficients.
//
// Replaces: bdn.hashPointToR   sign/bdn/bdn.go:29-63
//   blake2s.NewXOF(OutputLengthUnknown, nil) absorbs every public key's MarshalBinary in roster order, 16 bytes are
//   squeezed per key, reversed when the scalar type is big-endian (mod.Int: every pairing suite of the reference) and
//   SetBytes'd -- i.e. the 16 bytes are read LITTLE-endian.  Aggregation uses c_i + 1 (bdn.go:150-154, mask.go:58-61).
// The absorb phase is one sequential hash chain over n * pub_len bytes: inherently serial, it stays on the host here as
// in the reference (it is <1% of the curve work it feeds: the MSM over (c_i + 1, S_i) / (c_i + 1, PK_i) on the device).
// The squeeze phase (BLAKE2X: one independent compression per 32 output bytes) is spread over host threads.
// No device work, no context needed.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/b2kyber.h"

namespace {

constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
constexpr uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// BLAKE2s (RFC 7693) with an explicit 32-byte parameter block, streaming
struct Blake2s {
  uint32_t h[8];
  uint64_t t = 0;
  uint8_t buf[64];
  size_t fill = 0;

  explicit Blake2s(const uint8_t param[32]) {
    for (int i = 0; i < 8; i++) h[i] = IV[i] ^ le32(param + 4 * i);
  }
  void compress(const uint8_t* blk, bool last) {
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++) m[i] = le32(blk + 4 * i);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = IV[i]; }
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
      v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 16);
      v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 12);
      v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 8);
      v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 7);
    };
    for (int r = 0; r < 10; r++) {
      const uint8_t* s = SIGMA[r];
      G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
      G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
      G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
      G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
  }
  void update(const uint8_t* p, size_t n) {
    while (n > 0) {
      if (fill == 64) { t += 64; compress(buf, false); fill = 0; }   // a full buffer is only flushed when more follows
      size_t k = 64 - fill;
      if (k > n) k = n;
      std::memcpy(buf + fill, p, k);
      fill += k; p += k; n -= k;
    }
  }
  void final(uint8_t out[32]) {
    t += fill;
    std::memset(buf + fill, 0, 64 - fill);
    compress(buf, true);
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)h[i]; out[4 * i + 1] = (uint8_t)(h[i] >> 8); out[4 * i + 2] = (uint8_t)(h[i] >> 16); out[4 * i + 3] = (uint8_t)(h[i] >> 24); }
  }
};

// parameter block: digest_len, key_len, fanout, depth, leaf_len(4), node_offset(4), xof_len(2), node_depth, inner_len, salt(8), personal(8)
void param_block(uint8_t p[32], uint8_t digest, uint8_t fanout, uint8_t depth, uint32_t leaf, uint32_t node_offset, uint16_t xof_len,
                 uint8_t node_depth, uint8_t inner) {
  std::memset(p, 0, 32);
  p[0] = digest; p[2] = fanout; p[3] = depth;
  p[4] = (uint8_t)leaf; p[5] = (uint8_t)(leaf >> 8); p[6] = (uint8_t)(leaf >> 16); p[7] = (uint8_t)(leaf >> 24);
  p[8] = (uint8_t)node_offset; p[9] = (uint8_t)(node_offset >> 8); p[10] = (uint8_t)(node_offset >> 16); p[11] = (uint8_t)(node_offset >> 24);
  p[12] = (uint8_t)xof_len; p[13] = (uint8_t)(xof_len >> 8);
  p[14] = node_depth; p[15] = inner;
}

}  // namespace

extern "C" int b2k_bdn_coefficients(size_t n, const uint8_t* pubs, size_t pub_len, int add_one, uint8_t* out32n) {
  if ((n > 0 && (!pubs || !out32n)) || pub_len == 0) return B2K_ERR_ARG;
  if (n > (size_t)1 << 32) return B2K_ERR_ARG;           // BLAKE2Xs stream limit (2^32 blocks of 32 bytes) is far above
  uint8_t p[32], h0[32];
  param_block(p, 32, 1, 1, 0, 0, 0xFFFF, 0, 0);           // root: XOF length "unknown"
  Blake2s root(p);
  root.update(pubs, n * pub_len);
  root.final(h0);
  const size_t nblocks = (16 * n + 31) / 32;
  auto squeeze = [&](size_t lo, size_t hi) {
    for (size_t b = lo; b < hi; b++) {
      uint8_t pp[32], o[32];
      param_block(pp, 32, 0, 0, 32, (uint32_t)b, 0xFFFF, 0, 32);
      Blake2s node(pp);
      node.update(h0, 32);
      node.final(o);
      for (int half = 0; half < 2; half++) {
        const size_t i = 2 * b + half;
        if (i >= n) break;
        // 16 stream bytes, little-endian value c < 2^128; scalar wire format is 32 bytes big-endian
        uint8_t* dst = out32n + 32 * i;
        std::memset(dst, 0, 16);
        for (int k = 0; k < 16; k++) dst[31 - k] = o[16 * half + k];
        if (add_one) {
          for (int k = 31; k >= 0; k--) { if (++dst[k] != 0) break; }
        }
      }
    }
  };
  unsigned nt = std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  if (nt < 1 || nblocks < 4096) nt = 1;
  if (nt == 1) {
    squeeze(0, nblocks);
  } else {
    std::vector<std::thread> th;
    const size_t per = (nblocks + nt - 1) / nt;
    for (unsigned k = 0; k < nt; k++) {
      const size_t lo = k * per, hi = lo + per < nblocks ? lo + per : nblocks;
      if (lo < hi) th.emplace_back(squeeze, lo, hi);
    }
    for (auto& x : th) x.join();
  }
  return B2K_OK;
}
