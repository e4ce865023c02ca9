# Run under gpurun --gpus 4: the C2-per-GPU bench at N = 4 (trimmed: no CPU baseline / sections)
export B2K_SKIP_CPU_BASELINE=1 B2K_SKIP_PAIRINGS=1 B2K_SKIP_SECTIONS=1 B2K_SKIP_SUSTAINED=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 \
  > gpurun_out/scale_bench_4gpu.json 2> gpurun_out/scale_bench_4gpu.err
tail -c 1500 gpurun_out/scale_bench_4gpu.json; tail -3 gpurun_out/scale_bench_4gpu.err
