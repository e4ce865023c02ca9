"""C3 mode A probe: n independent BLS verifications (sigs on G1) end to end on the device."""
import sys, time, struct
sys.path.insert(0, '.')
import torch
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o, h2c_bls12381 as h
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sk = 0x1f2e3d4c5b6a7988
msg = bytes(range(32))
pk = o.g2_compress(o.g2_mul(sk)); sig = o.g1_compress(o.g1_mul(sk, h.hash_to_g1(msg)))
d_pk = torch.frombuffer(bytearray(pk * n), dtype=torch.uint8).cuda()
d_sig = torch.frombuffer(bytearray(sig * n), dtype=torch.uint8).cuda()
d_msg = torch.frombuffer(bytearray(msg * n), dtype=torch.uint8).cuda()
d_off = torch.arange(0, 32 * (n + 1), 32, dtype=torch.int32).cuda()
d_dst = torch.frombuffer(bytearray(h.DST_G1), dtype=torch.uint8).cuda()
d_ok = torch.zeros(n, dtype=torch.uint8, device='cuda')
d_h = torch.zeros(n * 96, dtype=torch.uint8, device='cuda')
eng.set_stream(torch.cuda.current_stream().cuda_stream)
def verify():
    eng._check(eng.lib.b2k_bls12381_verify_g1sig_dev(eng.h, n, d_pk.data_ptr(), d_msg.data_ptr(), d_off.data_ptr(), d_dst.data_ptr(), len(h.DST_G1), d_sig.data_ptr(), d_ok.data_ptr()))
def hashonly():
    eng._check(eng.lib.b2k_bls12381_hash_to_g1_dev(eng.h, n, d_msg.data_ptr(), d_off.data_ptr(), d_dst.data_ptr(), len(h.DST_G1), d_h.data_ptr()))
for name, fn in (("hash_to_g1", hashonly), ("verify_g1sig", verify)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: n={n} {ms:.2f} ms -> {n/ms*1e3:.3e} /s", flush=True)
print("all verified:", int(d_ok.sum().item()) == n)
