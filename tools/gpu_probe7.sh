#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x > gpurun_out/pytest_kern.log 2>&1; echo "pytest kern exit $?"; tail -12 gpurun_out/pytest_kern.log
timeout 300 python tools/gemm_shapes.py tv2v > gpurun_out/shapes_tv2v.txt 2>&1; head -24 gpurun_out/shapes_tv2v.txt
