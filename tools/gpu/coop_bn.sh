# Run under gpurun: cooperative pairing on the three curves -- parity tests, then one-call latencies
timeout 400 python -m pytest tests/test_gpu_coop_pairing.py tests/test_gpu_bn254_pairing.py tests/test_gpu_bdn_bn256.py tests/test_gpu_gt.py -m gpu -q 2>&1 | tail -8
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from kyber_b200 import Engine
eng = Engine(0)
for curve in ("bn254", "bn256"):
    if curve == "bn254":
        from oracle import bn254 as c, bn254_pairing as b
        pair, check = eng.bn254_pair, eng.bn254_pairing_check
    else:
        from oracle import bn256 as c, bn256_pairing as b
        pair, check = eng.bn256_pair, eng.bn256_pairing_check
    g1, g2 = c.g1_marshal(c.g1_mul(5)), b.g2_marshal(b.g2_mul(7))
    h1 = c.g1_marshal(c.g1_mul(35))
    for n in (1, 1024, 8192):
        for coop in (1 << 20, 0):
            eng.set_pairing_coop(coop)
            pair(g1 * n, g2 * n); check(g1 * n, g2 * n, h1 * n, b.g2_marshal(b.G2) * n)
            t0 = time.perf_counter(); pair(g1 * n, g2 * n); t1 = time.perf_counter(); ok = check(g1 * n, g2 * n, h1 * n, b.g2_marshal(b.G2) * n); t2 = time.perf_counter()
            assert ok == b"\x01" * n
            print(f"{curve} n={n} {'cooperative' if coop else 'per-thread '}: pair {1e3 * (t1 - t0):8.2f} ms, check {1e3 * (t2 - t1):8.2f} ms (host call incl. copies)", flush=True)
PY
