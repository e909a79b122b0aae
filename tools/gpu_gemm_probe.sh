#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -m gpu -q -x > gpurun_out/pytest_probe.log 2>&1; echo "exit $?"; tail -4 gpurun_out/pytest_probe.log
for m in 1 0; do echo "== dev_gemm CCEDIT_GEMM_CLUSTER=$m"; CCEDIT_GEMM_CLUSTER=$m timeout 300 python tools/dev_gemm.py 2>&1 | tail -11; done
