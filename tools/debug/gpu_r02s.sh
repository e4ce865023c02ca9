timeout 400 python tools/perf_stage.py > gpurun_out/r02s_stage_ab.txt 2>&1; tail -10 gpurun_out/r02s_stage_ab.txt
timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q 2>&1 | tail -3
