timeout 120 ./tools/debug/g2_scale_debug 2>&1 | tail -8
timeout 120 python tools/debug/compact_crash.py 2>&1 | tail -25
timeout 200 cuda-gdb -batch -ex run -ex bt --args python tools/debug/compact_crash.py 2>&1 | tail -40 > gpurun_out/r02f_gdb.txt; tail -30 gpurun_out/r02f_gdb.txt
