"""CPU check of the warp-cooperative pairing's program tables (tools/gen_coop_pairing.py --check): the tower / Miller / final-exponentiation
formulas over plain integers, and the ENCODED, scheduled, slot-allocated program interpreted numerically (reads of a round before its
writes, like the lanes of a warp), must both reproduce the oracle's pairing -- for the 1-pair program (Suite.Pair) and the 2-pair program
(ValidatePairing, incl. a true e(aG, bH) e(-abG, H) = 1 instance) of BLS12-381, bn254 and bn256; and the generated
kyber_b200/csrc/coop_program_*.inc must be the generator's current output."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_program_tables_reproduce_the_oracle_and_are_up_to_date():
    inc = os.path.join(ROOT, "kyber_b200", "csrc", "coop_program_bn256.inc")
    if not os.path.exists(inc):                              # generated at build time (git-ignored)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_coop_pairing.py")], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_coop_pairing.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("encoded program ok") == 24 and "WRONG" not in r.stdout and "up to date" in r.stdout, r.stdout


def _limbs(v, n):
    return [(v >> (32 * i)) & 0xffffffff for i in range(n)]


def test_interpreter_core_under_host_emulation(tmp_path):
    """kyber_b200/csrc/coop_core.cuh (slot layout, the per-lane operation on Montgomery limbs, the constants tables) compiled for the host
    and run lane after lane over the generated programs: the BLS12-381 1-pair and 2-pair programs, the bn254 1-pair program and the Fp12
    square / product programs must give the oracle's values.  (Sequential lanes also expose a slot read in the round that writes it.)"""
    import ctypes
    import random
    sys.path.insert(0, ROOT)
    from oracle import bls12381 as o, bn254 as c4, bn254_pairing as b4
    inc = os.path.join(ROOT, "kyber_b200", "csrc", "coop_program_bls_gt.inc")
    if not os.path.exists(inc):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_coop_pairing.py")], check=True, stdout=subprocess.DEVNULL)
    so = str(tmp_path / "libemul_coop.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-DB2K_HOST_EMUL", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "host_emul", "emul_coop.cpp"), "-o", so], check=True)
    emul = ctypes.CDLL(so)
    rng = random.Random(9)

    def run(name, values, n):
        flat = [w for v in values for w in _limbs(v, n)]
        inp = (ctypes.c_uint32 * len(flat))(*flat)
        out = (ctypes.c_uint32 * (12 * n))()
        getattr(emul, name)(inp, out)
        return [sum(out[k * n + j] << (32 * j) for j in range(n)) for k in range(12)]

    def flat(f): return [f[h][k][c] for h in range(2) for k in range(3) for c in range(2)]

    def ins(pairs): return [x for p, q in pairs for x in (p[0], p[1], q[0][0], q[0][1], q[1][0], q[1][1])]
    p1, q1 = o.g1_mul(rng.randrange(1, o.R)), o.g2_mul(rng.randrange(1, o.R))
    assert run("emul_coop_bls_p1", ins([(p1, q1)]), 12) == flat(o.pairing_reference(p1, q1))
    a, b = rng.randrange(1, o.R), rng.randrange(1, o.R)
    pairs = [(o.g1_mul(a), o.g2_mul(b)), (o.g1_neg(o.g1_mul(a * b % o.R)), o.G2)]
    assert run("emul_coop_bls_p2", ins(pairs), 12) == flat(o.F12_ONE)             # a true ValidatePairing instance
    p4, q4 = c4.g1_mul(rng.randrange(1, c4.ORDER)), b4.g2_mul(rng.randrange(1, c4.ORDER))
    assert run("emul_coop_bn254_p1", ins([(p4, q4)]), 8) == flat(b4.pairing(p4, q4))
    x = tuple(tuple((rng.randrange(o.P), rng.randrange(o.P)) for _ in range(3)) for _ in range(2))
    y = tuple(tuple((rng.randrange(o.P), rng.randrange(o.P)) for _ in range(3)) for _ in range(2))
    assert run("emul_coop_bls_sqr", flat(x), 12) == flat(o.f12_mul(x, x))
    assert run("emul_coop_bls_mul", flat(x) + flat(y), 12) == flat(o.f12_mul(x, y))
