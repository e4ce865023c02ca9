timeout 300 python tools/perf_layout.py > gpurun_out/r02k_layout_ab.txt 2>&1; grep -i "pairing\|correct" gpurun_out/r02k_layout_ab.txt | head -8
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02k_gputests.txt; cat gpurun_out/r02k_gputests.txt
B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_SECTIONS=1 B2K_SKIP_SUSTAINED=1 timeout 300 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), {k: d[k].get('value') for k in ('independent_muls','pairings','bls_verify') if k in d})
PY
