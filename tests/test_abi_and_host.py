"""CPU tests: the C-ABI library loads and exports every symbol include/b2kyber.h declares (no compute
calls without a GPU); host-side workload generation; device limb code under host emulation."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from kyber_b200 import build, capi
    build.build()
    lib = capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "b2kyber.h")).read()
    names = set(re.findall(r"\b(b2k_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/b2kyber.h but not exported"
    assert b"sm_100a" in lib.b2k_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kyber_b200 import Engine, B2KError
    with pytest.raises(B2KError):
        Engine(0)


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "kyber_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle" not in txt.replace("# oracle-free", ""), f"{f} mentions the oracle"


def test_prng_workload_is_deterministic_and_in_range():
    from kyber_b200 import workload as wl
    a = wl.prng_scalars("b2k/c2", 64, wl.R_BLS12381)
    b = wl.prng_scalars("b2k/c2", 32, wl.R_BLS12381, start=32)
    assert a[32:] == b and all(0 <= x < wl.R_BLS12381 for x in a) and len(set(a)) == 64
    assert wl.dot_mod([2, 3], [5, 7], 11) == (10 + 21) % 11


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    lib = os.path.join(ROOT, "tests", "host_emul", "libb2k_emul.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(
            os.path.getmtime(src), *(os.path.getmtime(os.path.join(ROOT, "kyber_b200", "csrc", f))
                                     for f in os.listdir(os.path.join(ROOT, "kyber_b200", "csrc")) if f.endswith(".cuh"))):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-DB2K_HOST_EMUL", src, "-o", lib], check=True)
    return ctypes.CDLL(lib)


def test_device_field_code_under_host_emulation(emul):
    import random
    rng = random.Random(1)
    fields = (("fp381", 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 12),
              ("fr381", 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 8),
              ("fp254", 21888242871839275222246405745257275088696311157297823662689037894645226208583, 8),
              ("fp256", 65000549695646603732796438742359905742825358107623003571877145026864184071783, 10),
              ("fp25519", 2**255 - 19, 8))
    for name, p, n in fields:
        R = 1 << (32 * n)
        Ri = pow(R, -1, p)

        def lim(v):
            return (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])

        def val(a):
            return sum(int(x) << (32 * i) for i, x in enumerate(a))

        vals = [0, 1, p - 1, p - 2, (p - 1) // 2, R % p] + [rng.randrange(p) for _ in range(60)]
        for _ in range(400):
            a, b = rng.choice(vals), rng.choice(vals)
            out = (ctypes.c_uint32 * n)()
            getattr(emul, f"emul_{name}_mul")(lim(a), lim(b), out)
            assert val(out) == a * b * Ri % p
            getattr(emul, f"emul_{name}_sqr")(lim(a), out)                          # dedicated squaring (wide_sqr + redc_wide)
            assert val(out) == a * a * Ri % p
            wide = (ctypes.c_uint32 * (2 * n))()
            getattr(emul, f"emul_{name}_wide_mul")(lim(a), lim(b), wide)
            assert val(wide) == a * b
            getattr(emul, f"emul_{name}_redc")(wide, out)
            assert val(out) == a * b * Ri % p
            getattr(emul, f"emul_{name}_wide_sqr")(lim(b), wide)
            assert val(wide) == b * b
            getattr(emul, f"emul_{name}_add")(lim(a), lim(b), out)
            assert val(out) == (a + b) % p
            getattr(emul, f"emul_{name}_sub")(lim(a), lim(b), out)
            assert val(out) == (a - b) % p
            getattr(emul, f"emul_{name}_addsub")(lim(a), lim(b), 0, out)              # add / subtract as one instruction stream
            assert val(out) == (a + b) % p
            getattr(emul, f"emul_{name}_addsub")(lim(a), lim(b), 1, out)
            assert val(out) == (a - b) % p
        for _ in range(200):                                                        # wide products are exact for ANY N-limb operands
            a, b = rng.choice((R - 1, rng.randrange(R), 2 * p - 2 if 2 * p - 2 < R else p - 1)), rng.randrange(R)
            wide = (ctypes.c_uint32 * (2 * n))()
            getattr(emul, f"emul_{name}_wide_mul")(lim(a), lim(b), wide)
            assert val(wide) == a * b
            getattr(emul, f"emul_{name}_wide_sqr")(lim(a), wide)
            assert val(wide) == a * a
        if name in ("fp381", "fp254", "fp256"):                                     # Fp2 product: three reduced products == lazily reduced form
            for _ in range(200):
                a0, a1, b0, b1 = (rng.choice(vals) for _ in range(4))
                r1, r2 = (ctypes.c_uint32 * (2 * n))(), (ctypes.c_uint32 * (2 * n))()
                A = (ctypes.c_uint32 * (2 * n))(*(list(lim(a0)) + list(lim(a1))))
                B = (ctypes.c_uint32 * (2 * n))(*(list(lim(b0)) + list(lim(b1))))
                getattr(emul, f"emul_{name}_fp2_mul_pair")(A, B, r1, r2)
                assert list(r1) == list(r2)
                assert val(r1[:n]) == (a0 * b0 - a1 * b1) * Ri % p and val(r1[n:]) == (a0 * b1 + a1 * b0) * Ri % p
        for a in vals[:12]:
            out = (ctypes.c_uint32 * n)()
            getattr(emul, f"emul_{name}_inv")(lim(a * R % p), out)                 # binary GCD (what the kernels use)
            assert val(out) == (pow(a, -1, p) * R % p if a else 0)
            getattr(emul, f"emul_{name}_inv_fermat")(lim(a * R % p), out)          # a^(p-2): the independent cross-check
            assert val(out) == (pow(a, -1, p) * R % p if a else 0)


def test_device_msm_bodies_under_host_emulation(emul):
    import random
    from oracle import bls12381 as o
    rng = random.Random(2)
    n = 21
    ks = [0, 1, o.R - 1] + [rng.randrange(o.R) for _ in range(n - 3)]
    pts = [o.g1_mul(rng.randrange(1, o.R)) for _ in range(n)]
    pts[4] = None
    pts[6], ks[6] = pts[5], ks[5]
    pts[8], ks[8] = o.g1_neg(pts[7]), ks[7]
    sb = b"".join(o.scalar_to_bytes(k) for k in ks)
    pb = b"".join(o.g1_to_affine_bytes(p) for p in pts)
    out = ctypes.create_string_buffer(48 * n)
    emul.emul_bls12381_g1_mul_batch(ctypes.c_size_t(n), sb, pb, out)
    for i in range(n):
        assert out.raw[48 * i:48 * i + 48] == o.g1_compress(o.g1_mul(ks[i], pts[i]))
    exp = o.g1_compress(o.g1_msm(ks, pts))
    for c, m in ((4, 2), (8, 8), (13, 32), (16, 32)):
        o48 = ctypes.create_string_buffer(48)
        assert emul.emul_bls12381_g1_msm(ctypes.c_size_t(n), sb, pb, c, m, o48) == 0
        assert o48.raw == exp, (c, m)


def test_glv_front_end_under_host_emulation(emul):
    """The endomorphism split of the BLS12-381 G1 MSM: k = s1 k1 + s2 k2 x^2 (mod r) with k1, k2 < 2^127 for edge and
    random scalars, phi(P) = (beta x, y) = [-x^2] P, and the whole split pipeline == the oracle MSM."""
    import random
    from oracle import bls12381 as o
    rng = random.Random(12)
    x2 = o.X_ABS ** 2
    edge = [0, 1, 2, x2 - 1, x2, x2 + 1, x2 // 2, x2 // 2 + 1, (o.R - 1) // 2, (o.R + 1) // 2, o.R - 1, o.R - 2, o.R - x2, 3 * x2 + x2 // 2,
            (1 << 254) - 1, 1 << 128, (1 << 128) - 1]
    k1b, k2b = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
    for k in edge + [rng.randrange(o.R) for _ in range(300)]:
        fl = emul.emul_glv_split_bls381(k.to_bytes(32, "big"), k1b, k2b)
        k1, k2 = int.from_bytes(k1b.raw, "big"), int.from_bytes(k2b.raw, "big")
        assert k1 < 1 << 127 and k2 < 1 << 127, hex(k)
        s1 = -1 if fl & 1 else 1
        s2 = -1 if fl & 2 else 1
        assert (s1 * k1 + s2 * k2 * x2 - k) % o.R == 0, hex(k)
    # the second point of the split is -phi(P) = [x^2] P
    n = 19
    ks = [0, 1, o.R - 1, x2, o.R - x2] + [rng.randrange(o.R) for _ in range(n - 5)]
    pts = [o.g1_mul(rng.randrange(1, o.R)) for _ in range(n)]
    pts[3] = None
    pts[6], ks[6] = pts[5], ks[5]
    pts[8], ks[8] = o.g1_neg(pts[7]), ks[7]
    sb = b"".join(o.scalar_to_bytes(k) for k in ks)
    pb = b"".join(o.g1_to_affine_bytes(p) for p in pts)
    exp = o.g1_compress(o.g1_msm(ks, pts))
    for c, m, L in ((4, 2, 0), (9, 8, 4), (16, 32, 3)):
        o48 = ctypes.create_string_buffer(48)
        assert emul.emul_bls12381_g1_msm_glv(ctypes.c_size_t(n), sb, pb, c, m, L, o48) == 0
        assert o48.raw == exp, (c, m, L)


def test_bdn_coefficients_host_function_matches_reference_vector_and_oracle():
    """b2k_bdn_coefficients (host C++ BLAKE2Xs in the library) == the reference vector of
    sign/bdn/bdn_vartime_test.go:24-48 (coefficients for G2 base, 2*base, 3*base on bn256) == the oracle, incl. ragged
    sizes around the 64-byte block and 32-byte squeeze boundaries and the multi-threaded squeeze."""
    import json
    import random
    from kyber_b200.capi import Engine
    from oracle import bdn, bn256 as o
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bdn_bn256_fixtures.json")))["hash_point_to_r"]
    pubs = b"".join(o.g2_marshal(o.g2_mul(k)) for k in (1, 2, 3))
    got = Engine.bdn_coefficients(pubs, 128)
    assert [got[32 * i:32 * i + 32].hex().lstrip("0") for i in range(3)] == [c.lstrip("0") for c in fx["coefs"]]
    plus = Engine.bdn_coefficients(pubs, 128, add_one=True)
    assert [int.from_bytes(plus[32 * i:32 * i + 32], "big") for i in range(3)] == [int(c, 16) + 1 for c in fx["coefs"]]
    rng = random.Random(5)
    for n, plen in ((0, 48), (1, 1), (1, 64), (2, 32), (3, 48), (5, 96), (9000, 48)):
        blob = bytes(rng.getrandbits(8) for _ in range(n * plen))
        want = bdn.hash_point_to_r([blob[plen * i:plen * (i + 1)] for i in range(n)], 1 << 255) if n < 100 else None
        got = Engine.bdn_coefficients(blob, plen)
        assert len(got) == 32 * n
        if want is not None:
            assert [int.from_bytes(got[32 * i:32 * i + 32], "big") for i in range(n)] == want
        else:      # large n: threaded squeeze == the stream prefix property (first coefficients do not depend on n's parity)
            stream = bdn.blake2xs(blob, 16 * 64)
            assert [got[32 * i + 16:32 * i + 32][::-1] for i in range(64)] == [stream[16 * i:16 * i + 16] for i in range(64)]
            tail = bdn.blake2xs(blob, 16 * n)[-16:]
            assert got[-16:][::-1] == tail
