"""diagnostic (GPU): RecoverPubPoly on G2 for growing t; which commitments are wrong?"""
import random, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from kyber_b200 import Engine
from oracle import bls12381 as o
eng = Engine(0)
rng = random.Random(85)
for t in (1, 2, 3, 6):
    coeffs = [rng.randrange(o.R) for _ in range(t)]
    sc = b"".join(o.scalar_to_bytes(c) for c in coeffs)
    c2 = eng.commit_batch("bls12381_g2", sc)
    idx = sorted(rng.sample(range(50), t))
    sh2 = eng.bls12381_pubpoly_eval(2, c2, idx)
    got = eng.recover_pubpoly("bls12381_g2", idx, sh2)
    bad = [k for k in range(t) if got[192 * k:192 * k + 192] != c2[192 * k:192 * k + 192]]
    c1 = eng.commit_batch("bls12381_g1", sc)
    sh1 = eng.bls12381_pubpoly_eval(1, c1, idx)
    got1 = eng.recover_pubpoly("bls12381_g1", idx, sh1)
    print(f"t={t} idx={idx}: G2 wrong commitments {bad}; G1 ok: {got1 == c1}; G2 results on curve: "
          f"{[o.g2_is_on_curve(o.g2_from_affine_bytes(got[192*k:192*k+192])) for k in range(t)]}", flush=True)
