timeout 300 python tools/perf_g2.py > gpurun_out/r02p_g2.txt 2>&1; cat gpurun_out/r02p_g2.txt | tail -8
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02p_gputests.txt; cat gpurun_out/r02p_gputests.txt
B2K_SKIP_PAIRINGS=1 B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_SUSTAINED=1 timeout 300 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print({k: (d[k].get('value'), d[k].get('parts_ms')) for k in ('recover_commit','bdn_aggregate','ed25519') if k in d})
PY
