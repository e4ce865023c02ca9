timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02r_gputests.txt; cat gpurun_out/r02r_gputests.txt
B2K_SKIP_PAIRINGS=1 B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_SECTIONS=1 B2K_SKIP_SUSTAINED=1 timeout 300 python bench.py > gpurun_out/r02r_bench4.json 2> gpurun_out/r02r_bench4.err
B2K_SKIP_PAIRINGS=1 B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_SECTIONS=1 B2K_SKIP_SUSTAINED=1 timeout 300 python bench.py --contexts 6 > gpurun_out/r02r_bench6.json 2> gpurun_out/r02r_bench6.err
python - <<'PY'
import json
for f in ('gpurun_out/r02r_bench4.json','gpurun_out/r02r_bench6.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.4e ms %.3f single %.3f e2e %.4e (%.3f ms) blocking %.3f'%(d['value'], d['ms_per_step'], d['single_step_latency_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['blocking_ms_per_step']))
PY
tail -3 gpurun_out/r02r_bench4.err
