"""GPU timing: warp-cooperative pairing check (one warp per check) against the one-check-per-thread kernel, by batch size.
Usage (under gpurun): python tools/perf_coop.py > gpurun_out/<tag>_coop.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kyber_b200 import Engine
from oracle import bls12381 as o

eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
N = 1 << 16
x, y = 12345, 6789
a1 = torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_mul(x)) * N), dtype=torch.uint8).cuda()
a2 = torch.frombuffer(bytearray(o.g2_to_affine_bytes(o.g2_mul(y)) * N), dtype=torch.uint8).cuda()
b1 = torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_mul(x * y % o.R)) * N), dtype=torch.uint8).cuda()
b2 = torch.frombuffer(bytearray(o.g2_to_affine_bytes(o.G2) * N), dtype=torch.uint8).cuda()
ok = torch.zeros(N, dtype=torch.uint8, device="cuda")
gt = torch.empty(N * 576, dtype=torch.uint8, device="cuda")
want_gt = o.gt_to_bytes(o.pairing_reference(o.g1_mul(x), o.g2_mul(y)))


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in (1, 2, 8, 32, 148, 512, 1024, 2048, 4096, 8192, 16384, 65536):
    row = []
    for coop in (1 << 20, 0):
        eng._check(eng.lib.b2k_set_pairing_coop(eng.h, coop))
        ok.zero_()
        ms = timed(lambda: eng._check(eng.lib.b2k_bls12381_pairing_check_dev(eng.h, n, a1.data_ptr(), a2.data_ptr(), b1.data_ptr(), b2.data_ptr(), ok.data_ptr())), 3 if n < 8192 else 2)
        assert int(ok[:n].sum()) == n, (n, coop)
        msp = timed(lambda: eng.call_dev("b2k_bls12381_pair_dev", n, a1.data_ptr(), a2.data_ptr(), gt.data_ptr()), 2)
        assert bytes(gt[576 * (n - 1):576 * n].cpu().tolist()) == want_gt, (n, coop)
        row.append((ms, msp))
    (c_ms, c_p), (t_ms, t_p) = row
    print(f"n={n:6d}  check: cooperative {c_ms:9.3f} ms ({2 * n / c_ms * 1e3:.3e} pairings/s) | per-thread {t_ms:9.3f} ms ({2 * n / t_ms * 1e3:.3e})"
          f"   pair: cooperative {c_p:9.3f} ms | per-thread {t_p:9.3f} ms", flush=True)
eng._check(eng.lib.b2k_set_pairing_coop(eng.h, 10240))
