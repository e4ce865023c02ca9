import sys, os, faulthandler
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o
dev = torch.device("cuda", 0)
n = 4096
eng = Engine(0)
a = wl.prng_scalars("b2k/cc-a", n, o.R); s = wl.prng_scalars("b2k/cc", n, o.R)
pts = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
d_p = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
d_s = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).to(dev)
out = torch.zeros(64, dtype=torch.uint8, device=dev)
outm = torch.zeros(n * 48, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
eng._check(eng.lib.b2k_set_msm_layout(eng.h, 1))
print("compact mul_batch ...", flush=True)
eng.call_dev("b2k_bls12381_g1_mul_batch_dev", n, d_s.data_ptr(), d_p.data_ptr(), outm.data_ptr()); eng.synchronize()
print("  ok:", bytes(outm[:48].cpu().tolist()) == o.g1_compress(o.g1_mul(s[0] * a[0] % o.R)), flush=True)
print("compact msm ...", flush=True)
eng.call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_p.data_ptr(), out.data_ptr()); eng.synchronize()
print("  ok:", bytes(out[:48].cpu().tolist()) == o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R))), flush=True)
