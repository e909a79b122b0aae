#!/bin/bash
mkdir -p gpurun_out
for m in 2 1; do
  echo "== pytest kernels, CCEDIT_GEMM_CLUSTER=$m"
  CCEDIT_GEMM_CLUSTER=$m timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -m gpu -q -x > gpurun_out/pytest_cl$m.log 2>&1; echo "exit $?"; tail -6 gpurun_out/pytest_cl$m.log
done
for m in 0 1 2; do
  echo "== dev_gemm CCEDIT_GEMM_CLUSTER=$m"
  CCEDIT_GEMM_CLUSTER=$m timeout 300 python tools/dev_gemm.py 2>&1 | tail -12
done
