# Run under gpurun: warp-cooperative pairing -- parity tests, then cooperative vs per-thread kernel by batch size (profiles/r02t_coop.txt)
timeout 300 python -m pytest tests/test_gpu_coop_pairing.py tests/test_gpu_bls12381_pairing.py -m gpu -q -x 2>&1 | tail -8
timeout 400 python tools/perf_coop.py > gpurun_out/r02t_coop.txt 2>&1; tail -14 gpurun_out/r02t_coop.txt
