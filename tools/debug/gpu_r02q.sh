timeout 300 python tools/perf_g2.py > gpurun_out/r02q_g2.txt 2>&1; cat gpurun_out/r02q_g2.txt | tail -8
