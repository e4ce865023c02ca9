"""Throughput probe: K independent MSMs issued round-robin on NC contexts (streams) vs one context."""
import sys, time
sys.path.insert(0, '.')
import torch
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o
n = 1 << 20
a = wl.prng_scalars("b2k/c2-a", n, o.R); s = wl.prng_scalars("b2k/c2", n, o.R)
e0 = Engine(0)
pts_h = e0.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
d_pts = torch.frombuffer(bytearray(pts_h), dtype=torch.uint8).cuda()
d_s = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).cuda()
exp = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
K = 12
for nc in (1, 2, 3):
    engs = [e0] + [Engine(0) for _ in range(nc - 1)]
    outs = [torch.zeros(64, dtype=torch.uint8, device='cuda') for _ in range(nc)]
    def run():
        for k in range(K):
            engs[k % nc].call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outs[k % nc].data_ptr())
    run(); torch.cuda.synchronize()
    for e in engs: e.synchronize()
    t0 = time.perf_counter(); run()
    for e in engs: e.synchronize()
    dt = time.perf_counter() - t0
    ok = all(bytes(x[:48].cpu().tolist()) == exp for x in outs)
    print(f"contexts={nc}: {dt/K*1e3:.3f} ms per MSM -> {n*K/dt:.3e} muls/s ok={ok}", flush=True)
