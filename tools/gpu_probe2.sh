#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "linear or conv or temporal_conv" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gemm.log
echo "== staged (default)"; timeout 200 python tools/dev_gemm.py 2>&1 | tee gpurun_out/dev_gemm_staged.txt
echo "== direct"; CCEDIT_GEMM_EPI=0 timeout 200 python tools/dev_gemm.py 2>&1 | tee gpurun_out/dev_gemm_direct.txt
CCEDIT_GEMM_TRACE=1 timeout 200 python tools/dev_gemm.py > gpurun_out/dev_gemm_trace2.txt 2>&1; echo "trace exit $?"
