timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02h_gputests.txt; cat gpurun_out/r02h_gputests.txt
timeout 300 python tools/perf_layout.py > gpurun_out/r02h_layout_ab.txt 2>&1; tail -24 gpurun_out/r02h_layout_ab.txt
timeout 400 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 3000 gpurun_out/r02h_bench.json; tail -5 gpurun_out/r02h_bench.err
