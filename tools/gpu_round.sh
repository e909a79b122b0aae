#!/bin/bash
# First GPU round trip: smoke, parity report, GPU tests, bench, ncu launch list.  Run via gpurun from the repo root.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
tail -3 gpurun_out/smoke.log
timeout 900 python tools/parity_report.py > gpurun_out/parity.md 2> gpurun_out/parity.err; echo "parity exit $?"
cat gpurun_out/parity.md; tail -5 gpurun_out/parity.err
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
