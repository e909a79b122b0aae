#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/gemm_shapes.py tv2v > gpurun_out/shapes_tv2v.txt 2>&1; head -30 gpurun_out/shapes_tv2v.txt
timeout 600 python bench.py --kind tvi2v --no-cpu-baseline --steps 5 > gpurun_out/bench_tvi2v.json 2> gpurun_out/bench_tvi2v.err; echo "bench tvi2v exit $?"; cut -c1-200 gpurun_out/bench_tvi2v.json
