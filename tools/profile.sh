#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the bucket-accumulate kernel.
# Usage: tools/profile.sh <round-tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
export B2K_SKIP_CPU_BASELINE=1
export B2K_SKIP_PAIRINGS=1
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# the top kernel, once
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate_slices -s 3 -c 1 \
    -o gpurun_out/accumulate_${TAG} -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
