timeout 400 python tools/perf_stage.py > gpurun_out/r02m_stage_ab.txt 2>&1; tail -12 gpurun_out/r02m_stage_ab.txt
