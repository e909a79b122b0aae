#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench.json; grep -v "^frame" gpurun_out/bench.err | tail -3
timeout 600 python bench.py --kind tvi2v --no-cpu-baseline --steps 5 > gpurun_out/bench_tvi2v.json 2> gpurun_out/bench_tvi2v.err; echo "bench tvi2v exit $?"; cut -c1-300 gpurun_out/bench_tvi2v.json
