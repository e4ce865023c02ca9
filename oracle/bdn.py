"""CPU ORACLE (test infrastructure, NOT product code) -- the host-side parts of sign/bdn restated.

  hashPointToR       sign/bdn/bdn.go:29-63: unkeyed BLAKE2Xs (unknown output length) over every public key's
                     MarshalBinary, 16 bytes per key, reversed when the scalar is big-endian (mod.Int), SetBytes
  AggregateSignatures  bdn.go:126-161: sum over enabled i of (c_i * S_i + S_i)
  AggregatePublicKeys  bdn.go:166-181 with the terms of mask.go:57-61: sum over enabled i of (c_i * PK_i + PK_i)
The curve arithmetic of the two sums is what the engine's MSM replaces; this module only derives coefficients
and states the expected combination.  Pinned by sign/bdn/bdn_vartime_test.go:24-48 (coefficients) in tests/.
"""
from __future__ import annotations
import struct

_IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
_SIGMA = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
          [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
          [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
          [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
          [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0]]
_M = 0xFFFFFFFF


def _rotr(x, n): return ((x >> n) | (x << (32 - n))) & _M


def _blake2s(data: bytes, param: bytes, outlen: int) -> bytes:
    h = [iv ^ p for iv, p in zip(_IV, struct.unpack("<8I", param))]
    blocks = [data[i:i + 64] for i in range(0, len(data), 64)] or [b""]
    t = 0
    for bi, blk in enumerate(blocks):
        t += len(blk)
        m = list(struct.unpack("<16I", blk.ljust(64, b"\x00")))
        v = h + _IV[:]
        v[12] ^= t & _M
        v[13] ^= (t >> 32) & _M
        if bi == len(blocks) - 1:
            v[14] ^= _M
        for r in range(10):
            s = _SIGMA[r]
            for i, (a, b, c, d) in enumerate(((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
                                              (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14))):
                v[a] = (v[a] + v[b] + m[s[2 * i]]) & _M; v[d] = _rotr(v[d] ^ v[a], 16)
                v[c] = (v[c] + v[d]) & _M; v[b] = _rotr(v[b] ^ v[c], 12)
                v[a] = (v[a] + v[b] + m[s[2 * i + 1]]) & _M; v[d] = _rotr(v[d] ^ v[a], 8)
                v[c] = (v[c] + v[d]) & _M; v[b] = _rotr(v[b] ^ v[c], 7)
        h = [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]
    return struct.pack("<8I", *h)[:outlen]


def _param(digest, fanout, depth, leaf, node_offset, xof_len, node_depth, inner):
    return struct.pack("<BBBBIIHBB16x", digest, 0, fanout, depth, leaf, node_offset, xof_len, node_depth, inner)


def blake2xs(data: bytes, n: int) -> bytes:
    """blake2s.NewXOF(OutputLengthUnknown, nil): first n bytes of the stream."""
    h0 = _blake2s(data, _param(32, 1, 1, 0, 0, 0xFFFF, 0, 0), 32)
    out, i = b"", 0
    while len(out) < n:
        out += _blake2s(h0, _param(32, 0, 0, 32, i, 0xFFFF, 0, 32), 32)
        i += 1
    return out[:n]


def hash_point_to_r(pub_bytes, order: int):
    """coefficients c_i for the marshalled public keys (big-endian scalar type, i.e. mod.Int)."""
    out = blake2xs(b"".join(pub_bytes), 16 * len(pub_bytes))
    return [int.from_bytes(out[16 * i:16 * i + 16][::-1], "big") % order for i in range(len(pub_bytes))]
