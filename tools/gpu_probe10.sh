#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "linear or conv" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gemm.log
timeout 200 python tools/dev_gemm.py 2>&1 | head -6
