#!/bin/bash
mkdir -p gpurun_out
for st in 0 2; do for e in 0 1; do echo "== STAGGER=$st EMU=$e"; CCEDIT_ATTN_STAGGER=$st CCEDIT_ATTN_EMU=$e timeout 200 python tools/dev_attn.py 2>&1 | tee gpurun_out/dev_attn_s${st}_emu$e.txt | grep -E "BAD|attn F|Error|error" ; done; done
CCEDIT_ATTN_STAGGER=0 CCEDIT_ATTN_TRACE=1 timeout 200 python tools/dev_attn.py 2>&1 | tee gpurun_out/attn_trace_s0.txt | cut -c1-260 | head -10
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k attention > gpurun_out/pytest_kern.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_kern.log
