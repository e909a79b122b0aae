#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dev_attn.py 2>&1 | tee gpurun_out/dev_attn_gen.txt | grep -E "BAD|OK|attn F|Error|error" | tail -30
echo "== legacy"; CCEDIT_ATTN_LEGACY=1 timeout 300 python tools/dev_attn.py 2>&1 | grep -E "attn F" | tail -4
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k attention > gpurun_out/pytest_kern.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_kern.log
