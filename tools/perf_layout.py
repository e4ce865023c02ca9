"""A/B on the GPU: field-product code layout (inlined vs compact by-value calls) for the G1 MSM and the pairing kernels.
Usage (under gpurun): python tools/perf_layout.py > gpurun_out/<tag>_layout_ab.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o

dev = torch.device("cuda", 0)
n = 1 << 20
eng = Engine(0)
a = wl.prng_scalars("b2k/c2-a", n, o.R)
s = wl.prng_scalars("b2k/c2", n, o.R)
d_a = torch.frombuffer(bytearray(wl.scalars_to_bytes(a)), dtype=torch.uint8).to(dev)
d_gen = torch.frombuffer(bytearray(wl.G1_BLS12381_AFFINE), dtype=torch.uint8).to(dev).repeat(n)
d_pts = torch.empty(n * 96, dtype=torch.uint8, device=dev)
eng.call_dev("b2k_bls12381_g1_mul_batch_affine_dev", n, d_a.data_ptr(), d_gen.data_ptr(), d_pts.data_ptr())
d_s = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).to(dev)
want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
NC = 4
streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]
engs = [Engine(0) for _ in range(NC)]
outs = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(NC)]
for e, st in zip(engs, streams):
    e.set_stream(st.cuda_stream)
torch.cuda.synchronize()
for layout in (0, 1, 0, 1):
    for e in engs:
        e._check(e.lib.b2k_set_msm_layout(e.h, layout))
    for k in range(2 * NC):
        engs[k % NC].call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outs[k % NC].data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(streams[0])
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(NC)]
    K = 24
    for k in range(K):
        engs[k % NC].call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outs[k % NC].data_ptr())
    for st, ev in zip(streams, ends):
        ev.record(st)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in ends) / K
    assert bytes(outs[0][:48].cpu().tolist()) == want
    tm = None
    engs[0].call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outs[0].data_ptr())
    tm = engs[0].last_timings()
    print(f"MSM 2^20 layout={layout}: {ms:.3f} ms per MSM pipelined ({n / ms * 1e3:.3e} muls/s); one MSM stages: accumulate {tm[4]:.3f} ms, rounds {tm[10]:.3f}, pipeline {tm[8]:.3f}", flush=True)

# independent Point.Mul (k_mul_batch), both layouts
outm = torch.empty(n * 48, dtype=torch.uint8, device=dev)
for layout in (0, 1, 0, 1):
    eng._check(eng.lib.b2k_set_msm_layout(eng.h, layout))
    eng.call_dev("b2k_bls12381_g1_mul_batch_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outm.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    eng.synchronize()
    import time
    t0 = time.perf_counter()
    eng.call_dev("b2k_bls12381_g1_mul_batch_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outm.data_ptr())
    eng.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    assert bytes(outm[:48].cpu().tolist()) == o.g1_compress(o.g1_mul(s[0] * a[0] % o.R))
    print(f"mul_batch 2^20 layout={layout}: {ms:.2f} ms ({n / ms * 1e3:.3e} muls/s)", flush=True)
eng._check(eng.lib.b2k_set_msm_layout(eng.h, 0))

# pairings: variant = shape + 4 * layout; shapes 64x4 / 64x8 / 64x6 (threads x min blocks per SM); layouts compact / inlined
m = 65536
g1 = torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_mul(12345)) * m), dtype=torch.uint8).to(dev)
g2 = torch.frombuffer(bytearray(o.g2_to_affine_bytes(o.g2_mul(6789)) * m), dtype=torch.uint8).to(dev)
gt = torch.empty(m * 576, dtype=torch.uint8, device=dev)
ok = torch.empty(m, dtype=torch.uint8, device=dev)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
wantgt = o.gt_to_bytes(o.pairing_reference(o.g1_mul(12345), o.g2_mul(6789)))
for v in (0, 1, 2, 4, 5, 6):
    eng._check(eng.lib.b2k_set_pairing_variant(eng.h, v))
    for name, fn in (("pair", lambda: eng.call_dev("b2k_bls12381_pair_dev", m, g1.data_ptr(), g2.data_ptr(), gt.data_ptr())),
                     ("check", lambda: eng._check(eng.lib.b2k_bls12381_pairing_check_dev(eng.h, m, g1.data_ptr(), g2.data_ptr(), g1.data_ptr(), g2.data_ptr(), ok.data_ptr())))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        per = 1 if name == "pair" else 2
        print(f"pairing variant {v} ({('compact', 'inlined')[v // 4]} {('64x4', '64x8', '64x6')[v % 4]}) {name}: n={m} {ms:.2f} ms -> {per * m / ms * 1e3:.3e} pairings/s", flush=True)
    assert bool(ok.min().item() == 1) and bytes(gt[:576].cpu().tolist()) == wantgt
print("all variants correct")
