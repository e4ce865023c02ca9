# Run under gpurun: cooperative GT exponentiation -- parity tests, then one-call latency against the one-per-thread kernel
timeout 400 python -m pytest tests/test_gpu_coop_pairing.py tests/test_gpu_gt.py -m gpu -q 2>&1 | tail -6
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from kyber_b200 import Engine
from oracle import bls12381 as o
eng = Engine(0)
e = o.gt_to_bytes(o.pairing_reference(o.g1_mul(5), o.g2_mul(7)))
s = (o.R - 12345).to_bytes(32, "big")
for n in (1, 256, 2048):
    for coop in (1 << 20, 0):
        eng.set_pairing_coop(coop)
        eng.gt_exp("bls12381", s * n, e * n)
        t0 = time.perf_counter(); r = eng.gt_exp("bls12381", s * n, e * n); dt = time.perf_counter() - t0
        print(f"bls12381 GT.Mul n={n} {'cooperative' if coop else 'per-thread '}: {1e3 * dt:8.2f} ms (host call incl. copies)", flush=True)
PY
