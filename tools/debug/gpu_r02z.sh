timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02z_gputests.txt; cat gpurun_out/r02z_gputests.txt
timeout 900 bash tools/profile.sh r02z > gpurun_out/r02z_profile.log 2>&1; tail -5 gpurun_out/r02z_profile.log
timeout 500 python bench.py > gpurun_out/r02z_bench_1gpu.json 2> gpurun_out/r02z_bench.err; tail -c 600 gpurun_out/r02z_bench_1gpu.json; tail -3 gpurun_out/r02z_bench.err
