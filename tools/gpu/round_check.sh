# Run under gpurun: the end-of-round verification on one B200 -- GPU parity suite, full bench line, the reference (CPU) arm, smoke().
# Usage: gpurun --timeout 2400 -- bash tools/gpu/round_check.sh   (results land in gpurun_out/)
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/round_gputests.txt; cat gpurun_out/round_gputests.txt
timeout 600 python bench.py > gpurun_out/round_bench_1gpu.json 2> gpurun_out/round_bench.err; tail -c 400 gpurun_out/round_bench_1gpu.json; tail -3 gpurun_out/round_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/round_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/round_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
