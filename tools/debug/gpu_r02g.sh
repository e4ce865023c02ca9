timeout 60 ./tools/debug/g2_fine_debug 2>&1 | tail -20
timeout 120 python tools/debug/compact_crash.py 2>&1 | tail -6
timeout 300 python tools/perf_layout.py > gpurun_out/r02g_layout_ab.txt 2>&1; tail -24 gpurun_out/r02g_layout_ab.txt
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02g_gputests.txt; cat gpurun_out/r02g_gputests.txt
