"""Deterministic synthetic inputs of BASELINE.md section 3.

PRNG = SHA256(seed || u64le(i)); scalars are masked to the bit length of the modulus and rejected
while >= modulus, mirroring random.Int (reference util/random/rand.go:19-46).  This is workload
generation shared by tests and bench -- it contains no curve arithmetic.
"""
from __future__ import annotations
import hashlib
import struct

R_BLS12381 = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
R_BN254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G1_BLS12381_AFFINE = bytes.fromhex(
    "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1")


def prng_scalars(seed: str, n: int, modulus: int, start: int = 0):
    """n integers in [0, modulus): counter i -> SHA256(seed||u64le(i)) masked, retry with (i, attempt)."""
    bits = modulus.bit_length()
    mask = (1 << bits) - 1
    sb = seed.encode()
    out = []
    sha = hashlib.sha256
    for i in range(start, start + n):
        v = int.from_bytes(sha(sb + struct.pack("<Q", i)).digest(), "big") & mask
        attempt = 0
        while v >= modulus:
            attempt += 1
            v = int.from_bytes(sha(sb + struct.pack("<QQ", i, attempt)).digest(), "big") & mask
        out.append(v)
    return out


def scalars_to_bytes(vals) -> bytes:
    return b"".join(v.to_bytes(32, "big") for v in vals)


def dot_mod(a, b, modulus: int) -> int:
    acc = 0
    for x, y in zip(a, b):
        acc += x * y
    return acc % modulus
