timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02v_gputests.txt; cat gpurun_out/r02v_gputests.txt
B2K_SKIP_CPU_BASELINE=1 timeout 500 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']))
print(d['pairings'].get('value'), d['pairings'].get('small_batch_latency'), d['bls_verify'].get('value'))
print(d['bdn_aggregate'].get('value'), d['bdn_aggregate'].get('parts_ms'))
PY
tail -3 gpurun_out/r02v_bench.err
