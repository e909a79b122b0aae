#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 2; do echo "== EMU=$e"; CCEDIT_ATTN_EMU=$e timeout 200 python tools/dev_attn.py 2>&1 | tee gpurun_out/dev_attn_emu$e.txt | grep -E "BAD|attn F|Error|error" ; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x > gpurun_out/pytest_kern.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_kern.log
timeout 200 python tools/dev_gemm.py 2>&1 | tee gpurun_out/dev_gemm3.txt
