timeout 400 python tools/perf_pairing.py > gpurun_out/r02n_pairing_split.txt 2>&1; tail -30 gpurun_out/r02n_pairing_split.txt
