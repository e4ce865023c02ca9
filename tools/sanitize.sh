#!/bin/bash
# compute-sanitizer memcheck over a small subset of the GPU tests (run under gpurun); slow, keep it small
set -u
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_bls12381_g1.py -m gpu -x -q -k "small_matches_oracle or mul_batch_matches or skewed or single_and" > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck exit: $?"
tail -5 gpurun_out/memcheck.log; tail -3 gpurun_out/memcheck_pytest.log
