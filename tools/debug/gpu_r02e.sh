timeout 120 ./tools/debug/tree_debug > gpurun_out/r02e_tree_debug.txt 2>&1; cat gpurun_out/r02e_tree_debug.txt
timeout 60 python tools/debug/debug_pubpoly_g2.py 2>&1 | tail -6
timeout 300 python tools/perf_layout.py > gpurun_out/r02e_layout_ab.txt 2>&1; tail -22 gpurun_out/r02e_layout_ab.txt
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02e_gputests.txt; cat gpurun_out/r02e_gputests.txt
B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_SECTIONS=1 B2K_SKIP_SUSTAINED=1 timeout 200 python bench.py --steps 20 > gpurun_out/r02e_bench_1gpu.json 2> gpurun_out/r02e_bench_1gpu.err; tail -c 500 gpurun_out/r02e_bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_1gpu.json')); print({k:d.get(k) for k in ('value','ms_per_step','single_step_latency_ms')}, d['independent_muls']['value'], d['pairings']['value'], d['bls_verify']['value'], d['stages_ms'])"
