timeout 400 python tools/perf_pairing.py > gpurun_out/r02l_pairing_variants.txt 2>&1; cat gpurun_out/r02l_pairing_variants.txt | tail -26
