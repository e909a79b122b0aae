#!/bin/bash
# Round 2, first pass: GPU tests, parity report, bench (both arms).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/parity_report.py > gpurun_out/parity.md 2> gpurun_out/parity.err; echo "parity exit $?"; cat gpurun_out/parity.md; tail -3 gpurun_out/parity.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cut -c1-1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
