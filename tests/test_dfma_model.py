"""CPU check of tools/probe/dfma_model.py: the FP64-pipe Montgomery product planned for the next round
is exact -- every hi/lo limb product equals the integer product, the uint64 column accumulators never lose a carry, and the
result equals a b R^-1 mod p -- on random and extreme operands.  (A model, not product code: nothing under kyber_b200/ uses it.)"""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("dfma_model", os.path.join(ROOT, "tools", "probe", "dfma_model.py"))
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)


def test_limb_product_is_exact():
    rng = random.Random(5)
    cases = [(0, 0), (1, 1), (m.MASK, m.MASK), (m.MASK, 1), (1 << 51, 1 << 51), (0, m.MASK)]
    cases += [(rng.randrange(1 << 52), rng.randrange(1 << 52)) for _ in range(2000)]
    for a, b in cases:
        h, l = m.limb_product(a, b)
        assert h >> 52 == 104 + 1023 and l >> 52 == 52 + 1023
        assert ((h & m.MASK) << 52) + (l & m.MASK) == a * b


def test_montgomery_product_matches_bigints():
    rng = random.Random(7)
    rinv = pow(m.R, -1, m.P)
    ops = [0, 1, m.P - 1, m.P - 2, (1 << 380), m.R % m.P, (m.P - 1) // 2]
    pairs = [(a, b) for a in ops for b in ops] + [(rng.randrange(m.P), rng.randrange(m.P)) for _ in range(300)]
    stats = {}
    for a, b in pairs:
        got = m.from_limbs(m.mont_mul(m.to_limbs(a), m.to_limbs(b), stats))
        assert got == a * b * rinv % m.P
    assert stats["limb_products"] == len(pairs) * 2 * m.LIMBS * m.LIMBS
    b = m.budget()
    assert b["fp64_pipe_cycles"] < b["imad_form_pipe_cycles"]      # the FP64 form alone is shorter on ITS pipe ...
    assert b["issue_slots"] > 3 * b["imad_form_issue_slots"]       # ... but needs several times the issue slots: hence warp specialisation


def _host_lib(tmp_path):
    import ctypes
    import subprocess
    so = str(tmp_path / "libfp_dfma_host.so")
    subprocess.run(["g++", "-O2", "-frounding-math", "-shared", "-fPIC", "-std=c++17",
                    os.path.join(ROOT, "tools", "probe", "fp_dfma_host.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def test_mixed_xyzz_addition_on_the_fp64_field(tmp_path):
    """tools/probe/ec_dfma.cuh: field add/sub and the mixed XYZZ addition of the accumulate pass on the FP64-pipe field give the
    oracle's group law, incl. accumulator at infinity, P + P (reported), P + (-P)."""
    import ctypes
    from oracle import bls12381 as o
    lib = _host_lib(tmp_path)
    A8, A16, A32 = ctypes.c_uint64 * 8, ctypes.c_uint64 * 16, ctypes.c_uint64 * 32
    rng = random.Random(13)
    mont = lambda x: m.to_limbs(x * m.R % m.P)
    unmont = lambda limbs: m.from_limbs(limbs) * pow(m.R, -1, m.P) % m.P
    for _ in range(200):
        a, b = rng.randrange(m.P), rng.randrange(m.P)
        s_, d_ = A8(), A8()
        lib.dfma_add_sub(A8(*m.to_limbs(a)), A8(*m.to_limbs(b)), s_, d_)
        assert m.from_limbs(list(s_)) == (a + b) % m.P and m.from_limbs(list(d_)) == (a - b) % m.P
    for a, b in ((0, 0), (m.P - 1, m.P - 1), (0, m.P - 1), (m.P - 1, 0), (5, 5)):
        s_, d_ = A8(), A8()
        lib.dfma_add_sub(A8(*m.to_limbs(a)), A8(*m.to_limbs(b)), s_, d_)
        assert m.from_limbs(list(s_)) == (a + b) % m.P and m.from_limbs(list(d_)) == (a - b) % m.P
    one = A8(*mont(1))

    def xyzz(pt, z):
        x, y = pt
        return mont(x * z * z % m.P) + mont(y * z * z * z % m.P) + mont(z * z % m.P) + mont(z * z * z % m.P)

    def affine_of(acc):
        X, Y, ZZ, ZZZ = (unmont(list(acc[8 * i:8 * i + 8])) for i in range(4))
        if ZZ == 0:
            return None
        return (X * pow(ZZ, -1, m.P) % m.P, Y * pow(ZZZ, -1, m.P) % m.P)

    for _ in range(25):
        p1, p2 = o.g1_mul(rng.randrange(1, o.R)), o.g1_mul(rng.randrange(1, o.R))
        acc = A32(*xyzz(p1, rng.randrange(1, m.P)))
        assert lib.dfma_xyzz_madd(acc, A16(*(mont(p2[0]) + mont(p2[1]))), one) == 0
        assert affine_of(acc) == o.g1_add(p1, p2)
    p1 = o.g1_mul(12345)
    acc = A32(*([0] * 32))                                            # accumulator at infinity
    assert lib.dfma_xyzz_madd(acc, A16(*(mont(p1[0]) + mont(p1[1]))), one) == 1 and affine_of(acc) == p1
    acc = A32(*xyzz(p1, 77))                                          # P + P: reported, left to the doubling path
    assert lib.dfma_xyzz_madd(acc, A16(*(mont(p1[0]) + mont(p1[1]))), one) == 2
    n1 = o.g1_neg(p1)
    assert lib.dfma_xyzz_madd(acc, A16(*(mont(n1[0]) + mont(n1[1]))), one) == 3 and affine_of(acc) is None


def test_cuda_header_host_build_matches_model(tmp_path):
    """tools/probe/fp_dfma.cuh (the header the sm_100a probe kernel is built from) compiled for the host -- fma() under
    FE_TOWARDZERO standing in for fma.rz.f64 -- gives a b R^-1 mod p limb for limb."""
    import ctypes
    lib = _host_lib(tmp_path)
    A = ctypes.c_uint64 * 8
    rng = random.Random(11)
    rinv = pow(m.R, -1, m.P)
    ops = [0, 1, m.P - 1, m.P - 2, 1 << 380]
    for a, b in [(a, b) for a in ops for b in ops] + [(rng.randrange(m.P), rng.randrange(m.P)) for _ in range(500)]:
        out = A()
        assert lib.dfma_mont_mul(A(*m.to_limbs(a)), A(*m.to_limbs(b)), out) == 0
        assert m.from_limbs(list(out)) == a * b * rinv % m.P


def test_radix_2_384_variant_shares_the_library_representation(tmp_path):
    """mont_mul384 (seven 52-bit reduction steps + one 20-bit step) takes and returns the library's own representation --
    12 x 32-bit limbs of x 2^384 mod p, re-packed, no conversion product -- and equals a b 2^-384 mod p; model and header agree."""
    import ctypes
    lib = _host_lib(tmp_path)
    A12 = ctypes.c_uint32 * 12
    rng = random.Random(17)
    rinv = pow(1 << 384, -1, m.P)
    u32 = lambda x: [(x >> (32 * i)) & 0xffffffff for i in range(12)]
    ops = [0, 1, m.P - 1, m.P - 2, 1 << 380, (1 << 381) - 1 - (1 << 200)]
    ops = [x % m.P for x in ops]
    for a, b in [(a, b) for a in ops for b in ops] + [(rng.randrange(m.P), rng.randrange(m.P)) for _ in range(600)]:
        want = a * b * rinv % m.P
        assert m.from_limbs(m.mont_mul_r384(m.to_limbs(a), m.to_limbs(b))) == want
        out = A12()
        assert lib.dfma_mont_mul384_u32(A12(*u32(a)), A12(*u32(b)), out) == 0
        assert sum(int(v) << (32 * i) for i, v in enumerate(out)) == want


def test_mixed_addition_on_library_values(tmp_path):
    """xyzz_madd<R384>: accumulator and operand in the LIBRARY's layout (12 x 32-bit limbs, Montgomery radix 2^384, the Xyzz / Affine
    of kyber_b200/csrc/ec.cuh) go through re-pack -> FP64-form addition -> re-pack and give the oracle's sum: what a warp of the
    future dual-pipe accumulate kernel does to a bucket."""
    import ctypes
    from oracle import bls12381 as o
    lib = _host_lib(tmp_path)
    R384 = 1 << 384
    u32 = lambda x: [(x >> (32 * i)) & 0xffffffff for i in range(12)]
    mont = lambda x: u32(x * R384 % m.P)
    val = lambda w: sum(int(v) << (32 * i) for i, v in enumerate(w)) * pow(R384, -1, m.P) % m.P
    A12, A24, A48 = ctypes.c_uint32 * 12, ctypes.c_uint32 * 24, ctypes.c_uint32 * 48
    one = A12(*mont(1))
    rng = random.Random(19)
    for _ in range(20):
        p1, p2 = o.g1_mul(rng.randrange(1, o.R)), o.g1_mul(rng.randrange(1, o.R))
        z = rng.randrange(1, m.P)
        acc = A48(*(mont(p1[0] * z * z % m.P) + mont(p1[1] * z * z * z % m.P) + mont(z * z % m.P) + mont(z * z * z % m.P)))
        assert lib.dfma_xyzz_madd384_u32(acc, A24(*(mont(p2[0]) + mont(p2[1]))), one) == 0
        X, Y, ZZ, ZZZ = (val(acc[12 * i:12 * i + 12]) for i in range(4))
        assert (X * pow(ZZ, -1, m.P) % m.P, Y * pow(ZZZ, -1, m.P) % m.P) == o.g1_add(p1, p2)
    acc = A48(*([0] * 48))
    p1 = o.g1_mul(99)
    assert lib.dfma_xyzz_madd384_u32(acc, A24(*(mont(p1[0]) + mont(p1[1]))), one) == 1
    assert (val(acc[0:12]), val(acc[12:24]), val(acc[24:36])) == (p1[0], p1[1], 1)


def test_fp64_slice_body_in_the_msm_pipeline(tmp_path):
    """tools/probe/msm_slice_fp64.cuh under host emulation (tests/host_emul/emul_fp64.cpp): the balanced-slice accumulate body on the
    FP64-form field, for all slices (mode 1) or for alternate groups of slices next to the IMAD-form body (mode 2, what a
    warp-specialised kernel does), gives the oracle's MSM -- incl. an operand at infinity, the same point three times in one
    bucket (the doubling path) and P, -P in one bucket."""
    import ctypes
    import subprocess
    from oracle import bls12381 as o
    so = str(tmp_path / "libemul_fp64.so")
    subprocess.run(["g++", "-O2", "-frounding-math", "-shared", "-fPIC", "-std=c++17", "-DB2K_HOST_EMUL", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "host_emul", "emul_fp64.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    rng = random.Random(41)
    n = 40
    pts = [o.g1_mul(rng.randrange(1, o.R)) for _ in range(n)]
    pts[3] = None
    pts[6] = pts[7] = pts[8]
    pts[10] = o.g1_neg(pts[11])
    pb = b"".join(o.g1_to_affine_bytes(p) for p in pts)
    for ks in ([rng.randrange(o.R) for _ in range(n)], [0x123456789ABCDEF] * n, [rng.randrange(1 << 20) for _ in range(n)]):
        sb = b"".join(o.scalar_to_bytes(k) for k in ks)
        want = o.g1_compress(o.g1_msm(ks, pts))
        for c, mm, L in ((4, 2, 1), (8, 1, 7), (13, 32, 5), (16, 32, 2)):
            for mode in (0, 1, 2):
                o48 = ctypes.create_string_buffer(48)
                assert lib.emul_bls12381_g1_msm_fp64(ctypes.c_size_t(n), sb, pb, c, mm, L, mode, o48) == 0 and o48.raw == want, (c, mm, L, mode)


def test_library_ec_templates_instantiate_on_the_fp64_field(tmp_path):
    """tools/probe/fpd_overloads.cuh puts the FP64-form field behind the library's generic field interface (f_mul, f_sub ...), so the
    EC templates of kyber_b200/csrc/ec.cuh -- xyzz_madd with both signs, xyzz_add, xyzz_dbl -- and the inversion instantiate on it
    unchanged; every result must equal, limb for limb after re-packing, the 12 x 32-bit instantiation (host emulation), incl. the
    accumulator or the operand at infinity, equal points and opposite points."""
    import ctypes
    import subprocess
    from oracle import bls12381 as o
    so = str(tmp_path / "libemul_fpd.so")
    subprocess.run(["g++", "-O2", "-frounding-math", "-shared", "-fPIC", "-std=c++17", "-DB2K_HOST_EMUL", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "host_emul", "emul_fpd.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    P, R384 = o.P, 1 << 384
    u32 = lambda x: [(x >> (32 * i)) & 0xffffffff for i in range(12)]
    mont = lambda x: u32(x * R384 % P)
    A24, A48 = ctypes.c_uint32 * 24, ctypes.c_uint32 * 48

    def xyzz(pt, z):
        if pt is None:
            return mont(1) + mont(1) + [0] * 24
        return mont(pt[0] * z * z % P) + mont(pt[1] * z ** 3 % P) + mont(z * z % P) + mont(z ** 3 % P)

    rng = random.Random(5)
    for k in range(30):
        p1, p2, p3 = (o.g1_mul(rng.randrange(1, o.R)) for _ in range(3))
        if k == 1: p2 = p1
        if k == 2: p2 = o.g1_neg(p1)
        if k == 3: p1 = None
        if k == 4: p2 = None
        if k == 5: p3 = p1
        q = [0] * 24 if p2 is None else mont(p2[0]) + mont(p2[1])
        rc = lib.emul_fpd_ec_agree(A48(*xyzz(p1, rng.randrange(1, P))), A24(*q), A48(*xyzz(p3, rng.randrange(1, P))))
        assert rc == 0, (k, rc)


def test_pair_tree_rounds_instantiate_on_the_fp64_field(tmp_path):
    """The affine pair-tree rounds of the accumulate pass (kyber_b200/csrc/msm_affine.cuh: msm_pairtree_forward / invert / backward,
    UNCHANGED templates) instantiated on the FP64-form field through tools/probe/fpd_overloads.cuh, followed by the library's XYZZ
    slices, give the oracle's MSM -- same exceptional inputs as the IMAD-form test (operand at infinity, one point four times,
    P and -P, points with x = 0, all-equal and short scalars)."""
    import ctypes
    import subprocess
    from oracle import bls12381 as o
    so = str(tmp_path / "libemul_fp64b.so")
    subprocess.run(["g++", "-O2", "-frounding-math", "-shared", "-fPIC", "-std=c++17", "-DB2K_HOST_EMUL", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "host_emul", "emul_fp64.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    rng = random.Random(11)
    n = 48
    pts = [o.g1_mul(rng.randrange(1, o.R)) for _ in range(n)]
    pts[5] = None
    pts[6] = pts[7] = pts[8] = pts[9]
    pts[10] = o.g1_neg(pts[11])
    pts[13] = (0, 2)
    pts[14] = (0, o.P - 2)
    pb = b"".join(o.g1_to_affine_bytes(p) for p in pts)
    for ks in ([rng.randrange(o.R) for _ in range(n)], [0x123456789ABCDEF] * n, [rng.randrange(1 << 9) for _ in range(n)]):
        sb = b"".join(o.scalar_to_bytes(k) for k in ks)
        want = o.g1_compress(o.g1_msm(ks, pts))
        for c, mm, L, rounds, pbatch in ((4, 2, 3, 1, 3), (4, 2, 2, 2, 9), (5, 4, 4, 3, 24), (8, 8, 5, 6, 64), (3, 2, 1, 9, 5)):
            o48 = ctypes.create_string_buffer(48)
            rc = lib.emul_bls12381_g1_msm_rounds_fp64(ctypes.c_size_t(n), sb, pb, c, mm, L, rounds, pbatch, o48)
            assert rc == 0 and o48.raw == want, (c, mm, L, rounds, pbatch)
