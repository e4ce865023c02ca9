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


def test_cuda_header_host_build_matches_model(tmp_path):
    """tools/probe/fp_dfma.cuh (the header the sm_100a probe kernel is built from) compiled for the host -- fma() under
    FE_TOWARDZERO standing in for fma.rz.f64 -- gives a b R^-1 mod p limb for limb."""
    import ctypes
    import subprocess
    so = str(tmp_path / "libfp_dfma_host.so")
    subprocess.run(["g++", "-O2", "-frounding-math", "-shared", "-fPIC", "-std=c++17",
                    os.path.join(ROOT, "tools", "probe", "fp_dfma_host.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    A = ctypes.c_uint64 * 8
    rng = random.Random(11)
    rinv = pow(m.R, -1, m.P)
    ops = [0, 1, m.P - 1, m.P - 2, 1 << 380]
    for a, b in [(a, b) for a in ops for b in ops] + [(rng.randrange(m.P), rng.randrange(m.P)) for _ in range(500)]:
        out = A()
        assert lib.dfma_mont_mul(A(*m.to_limbs(a)), A(*m.to_limbs(b)), out) == 0
        assert m.from_limbs(list(out)) == a * b * rinv % m.P
