# 8-GPU box: multi-rank parity test, then the C5 bench at N = 8 (trimmed: no CPU baseline / sections), N = 2 for the scaling point
export B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_PAIRINGS=1 B2K_SKIP_SECTIONS=1 B2K_SKIP_SUSTAINED=1
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r02i_multi_tests.txt; cat gpurun_out/r02i_multi_tests.txt
for N in 8 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02i_bench_${N}gpu.json 2> gpurun_out/r02i_bench_${N}gpu.err
  tail -c 2500 gpurun_out/r02i_bench_${N}gpu.json; tail -3 gpurun_out/r02i_bench_${N}gpu.err
done
