"""diagnostic (GPU): where does the bn254 product check go wrong?"""
import random, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kyber_b200 import Engine
from oracle import bn254 as c4, bn254_pairing as b4
eng = Engine(0)
rng = random.Random(204)
ks = [(rng.randrange(1, c4.ORDER), rng.randrange(1, c4.ORDER)) for _ in range(3)]
one = b4.gt_to_bytes(b4.F12_ONE)
for n in (1, 2, 3):
    sub = ks[:n]
    tot = sum(a * b for a, b in sub) % c4.ORDER
    g1 = b"".join(c4.g1_marshal(c4.g1_mul(a)) for a, _ in sub) + c4.g1_marshal(c4.g1_mul((c4.ORDER - tot) % c4.ORDER))
    g2 = b"".join(b4.g2_marshal(b4.g2_mul(b)) for _, b in sub) + b4.g2_marshal(b4.G2)
    f = eng.miller("bn254", g1, g2)
    e = eng.final_exp("bn254", f)
    pe = e[:384]
    pf = f[:384]
    for i in range(1, n + 1):
        pe = eng.gt_mul("bn254", pe, e[384 * i:384 * (i + 1)])
        pf = eng.gt_mul("bn254", pf, f[384 * i:384 * (i + 1)])
    print(n + 1, "pairs: prod of pairings == 1:", pe == one, "| finalize(prod of millers) == 1:", eng.final_exp("bn254", pf) == one,
          "| kernel:", eng.pairing_product_check("bn254", g1, g2), "| pair bytes ok:", e == eng.bn254_pair(g1, g2))
