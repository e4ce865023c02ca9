"""A/B: 2^20 independent BLS12-381 G1 Point.Mul (k_mul_batch): endomorphism + windows (default) vs plain windows, and the
resident-blocks variants of the kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kyber_b200 import workload as wl
from kyber_b200.capi import Engine
from oracle import bls12381 as o

n = 1 << 20
eng = Engine(0)
a = wl.prng_scalars("b2k/c2-a", n, o.R)
s = wl.prng_scalars("b2k/c2", n, o.R)
pts_b = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
pts = torch.frombuffer(bytearray(pts_b), dtype=torch.uint8).cuda()
sc = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).cuda()
out = torch.zeros(48 * n, dtype=torch.uint8, device="cuda")
ref = None
for glv, occ in ((1, 0), (1, 3), (1, 4), (0, 0), (0, 3)):
    eng.set_msm_glv(bool(glv))
    eng.set_mul_occupancy(occ)
    eng.call_dev("b2k_bls12381_g1_mul_batch_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr()); eng.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 3
    for _ in range(K):
        eng.call_dev("b2k_bls12381_g1_mul_batch_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr())
    eng.synchronize()
    dt = (time.perf_counter() - t0) / K
    got = bytes(out.cpu().numpy())
    if ref is None:
        ref = got
        for i in (0, 1, n // 3, n - 1):
            assert got[48 * i:48 * i + 48] == o.g1_compress(o.g1_mul(s[i] * a[i] % o.R)), i
    print("glv", glv, "blocks/SM", occ, "ms", round(dt * 1e3, 2), "muls/s %.3e" % (n / dt), "same" if got == ref else "DIFFERENT", flush=True)
