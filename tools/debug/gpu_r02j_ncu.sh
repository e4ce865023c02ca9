# ncu: pairing check kernel at n = 65536 in the default (compact) layout, variants 0 (255 regs) and 2 (168 regs); then the MSM pass
mkdir -p gpurun_out
cat > /tmp/pair_probe.py <<'PY'
import sys, ctypes
sys.path.insert(0, '.')
import torch
from kyber_b200 import Engine
from oracle import bls12381 as o
eng = Engine(0); n = 65536
a1 = torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_mul(12345)) * n), dtype=torch.uint8).cuda()
a2 = torch.frombuffer(bytearray(o.g2_to_affine_bytes(o.g2_mul(6789)) * n), dtype=torch.uint8).cuda()
ok = torch.empty(n, dtype=torch.uint8, device='cuda')
for v in (0, 2):
    eng.lib.b2k_set_pairing_variant(eng.h, v)
    eng._check(eng.lib.b2k_bls12381_pairing_check_dev(eng.h, n, a1.data_ptr(), a2.data_ptr(), a1.data_ptr(), a2.data_ptr(), ok.data_ptr()))
    eng.synchronize(); print(v, int(ok.sum()))
PY
ncu --set full --clock-control none --import-source on -k regex:k_bls_pairing_check -c 2 \
    -o gpurun_out/pairing_check_r02j -f python /tmp/pair_probe.py > gpurun_out/ncu_pairing_r02j.log 2>&1
ncu -i gpurun_out/pairing_check_r02j.ncu-rep --page details > gpurun_out/r02j_pairing_check_ncu_details.txt 2>&1
ncu -i gpurun_out/pairing_check_r02j.ncu-rep --page raw --csv > gpurun_out/r02j_pairing_check_ncu_raw.csv 2>&1
ncu -i gpurun_out/pairing_check_r02j.ncu-rep --page source --csv > gpurun_out/r02j_pairing_check_ncu_source.csv 2>&1
ls -la gpurun_out/pairing_check_r02j.ncu-rep gpurun_out/r02j_pairing_check_ncu_source.csv
gzip -9 -f gpurun_out/r02j_pairing_check_ncu_source.csv
ls -la gpurun_out/r02j_pairing_check_ncu_source.csv.gz
rm -f gpurun_out/pairing_check_r02j.ncu-rep
tail -3 gpurun_out/ncu_pairing_r02j.log
