# Run under gpurun --gpus 2: the process-per-GPU parity tests, then the bench at N = 2 (trimmed)
export B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_PAIRINGS=1 B2K_SKIP_SECTIONS=1 B2K_SKIP_SUSTAINED=1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 \
  > gpurun_out/scale_bench_2gpu.json 2> gpurun_out/scale_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/scale_bench_2gpu.json').read().strip().splitlines()[-1])
print('N=2 value %.4e ms %.3f e2e %.4e'%(d['value'], d['ms_per_step'], d['e2e']['value']), d['multi_gpu_exchange']['ms_per_step'])
PY
tail -2 gpurun_out/scale_bench_2gpu.err
