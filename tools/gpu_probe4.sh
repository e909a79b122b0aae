#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/dev_attn.py 2>&1 | tee gpurun_out/dev_attn_hpc.txt | grep -E "BAD|attn F|Error|error" | tail -12
echo "== hpc forced 1"; CCEDIT_ATTN_HPC=1 timeout 120 python tools/dev_attn.py 2>&1 | grep -E "BAD|cross-attn" | tail -4
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k attention > gpurun_out/pytest_kern.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_kern.log
