import sys
sys.path.insert(0, '.')
import torch
from kyber_b200 import Engine
from oracle import bls12381 as o
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g1 = torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_mul(12345)) * n), dtype=torch.uint8).cuda()
g2 = torch.frombuffer(bytearray(o.g2_to_affine_bytes(o.g2_mul(6789)) * n), dtype=torch.uint8).cuda()
gt = torch.empty(n * 576, dtype=torch.uint8, device='cuda')
ok = torch.empty(n, dtype=torch.uint8, device='cuda')
eng.set_stream(torch.cuda.current_stream().cuda_stream)
want = o.gt_to_bytes(o.pairing_reference(o.g1_mul(12345), o.g2_mul(6789)))
variants = [0, 1, 2, 4, 16, 17, 18, 19, 0]      # fused compact shapes, fused inlined, the four split forms
for v in variants:
    eng._check(eng.lib.b2k_set_pairing_variant(eng.h, v))
    for name, fn in (("pair", lambda: eng.call_dev("b2k_bls12381_pair_dev", n, g1.data_ptr(), g2.data_ptr(), gt.data_ptr())),
                     ("check", lambda: eng._check(eng.lib.b2k_bls12381_pairing_check_dev(eng.h, n, g1.data_ptr(), g2.data_ptr(), g1.data_ptr(), g2.data_ptr(), ok.data_ptr())))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        per = 1 if name == "pair" else 2
        print(f"variant {v} {name}: n={n} {ms:.2f} ms -> {per*n/ms*1e3:.3e} pairings/s", flush=True)
    assert bool(ok.min().item() == 1) and bytes(gt[:576].cpu().tolist()) == want
print("all variants correct")
