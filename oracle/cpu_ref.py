"""Loader for oracle/cpu_ref.c (the C restatement used as checker and CPU baseline).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this module.
"""
from __future__ import annotations
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "cpu_ref.c")
LIB = os.path.join(_HERE, "libcpu_ref.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -march=native would not travel to a different host CPU: build portable x86-64-v3 (BMI2/ADX-less safe)
        subprocess.run(["gcc", "-O3", "-mbmi2", "-shared", "-fPIC", "-pthread", SRC, "-o", LIB], check=True)
    return LIB


def load() -> C.CDLL:
    lib = C.CDLL(build())
    for name in ("cpu_g1_mul_batch", "cpu_g1_mul_batch_affine", "cpu_g1_msm_muladd", "cpu_g1_msm_pippenger"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    return lib


def _call(lib, name, scalars: bytes, points: bytes, out_len: int, threads: int) -> bytes:
    n = len(scalars) // 32
    out = C.create_string_buffer(out_len)
    rc = getattr(lib, name)(n, scalars, points, out, threads)
    if rc != 0:
        raise ValueError(f"{name} failed with {rc}")
    return out.raw


def g1_mul_batch(lib, scalars: bytes, points: bytes, threads: int = 1) -> bytes:
    return _call(lib, "cpu_g1_mul_batch", scalars, points, 48 * (len(scalars) // 32), threads)


def g1_mul_batch_affine(lib, scalars: bytes, points: bytes, threads: int = 1) -> bytes:
    return _call(lib, "cpu_g1_mul_batch_affine", scalars, points, 96 * (len(scalars) // 32), threads)


def g1_msm_muladd(lib, scalars: bytes, points: bytes, threads: int = 1) -> bytes:
    return _call(lib, "cpu_g1_msm_muladd", scalars, points, 48, threads)


def g1_msm_pippenger(lib, scalars: bytes, points: bytes, threads: int = 1) -> bytes:
    return _call(lib, "cpu_g1_msm_pippenger", scalars, points, 48, threads)


def pair(lib, g1: bytes, g2: bytes, threads: int = 1) -> bytes:
    n = len(g1) // 96
    out = C.create_string_buffer(576 * n)
    lib.cpu_pair.restype = C.c_int
    lib.cpu_pair.argtypes = [C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    assert lib.cpu_pair(n, g1, g2, out, threads) == 0
    return out.raw


def pairing_check(lib, a1: bytes, a2: bytes, b1: bytes, b2: bytes, threads: int = 1) -> bytes:
    n = len(a1) // 96
    out = C.create_string_buffer(n)
    lib.cpu_pairing_check.restype = C.c_int
    lib.cpu_pairing_check.argtypes = [C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    assert lib.cpu_pairing_check(n, a1, a2, b1, b2, out, threads) == 0
    return out.raw
