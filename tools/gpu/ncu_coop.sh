# Run under gpurun: ncu --set full of the cooperative pairing check, one warp per SM (profiles/r02u_coop_ncu_details.txt)
cat > /tmp/coop_probe.py <<'PY'
import sys
sys.path.insert(0, '.')
import torch
from kyber_b200 import Engine
from oracle import bls12381 as o
eng = Engine(0); n = 148
x, y = 12345, 6789
a1 = torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_mul(x)) * n), dtype=torch.uint8).cuda()
a2 = torch.frombuffer(bytearray(o.g2_to_affine_bytes(o.g2_mul(y)) * n), dtype=torch.uint8).cuda()
b1 = torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_mul(x * y % o.R)) * n), dtype=torch.uint8).cuda()
b2 = torch.frombuffer(bytearray(o.g2_to_affine_bytes(o.G2) * n), dtype=torch.uint8).cuda()
ok = torch.empty(n, dtype=torch.uint8, device='cuda')
eng._check(eng.lib.b2k_bls12381_pairing_check_dev(eng.h, n, a1.data_ptr(), a2.data_ptr(), b1.data_ptr(), b2.data_ptr(), ok.data_ptr()))
eng.synchronize(); print(int(ok.sum()))
PY
ncu --set full --clock-control none --import-source on -k regex:k_coop_pairing_check -c 1 -o gpurun_out/coop_r02u -f python /tmp/coop_probe.py > gpurun_out/ncu_coop_r02u.log 2>&1
ncu -i gpurun_out/coop_r02u.ncu-rep --page details > gpurun_out/r02u_coop_ncu_details.txt 2>&1
ncu -i gpurun_out/coop_r02u.ncu-rep --page source --csv > gpurun_out/r02u_coop_ncu_source.csv 2>&1
gzip -9 -f gpurun_out/r02u_coop_ncu_source.csv; rm -f gpurun_out/coop_r02u.ncu-rep; ls -la gpurun_out/r02u*; tail -2 gpurun_out/ncu_coop_r02u.log
