#!/bin/bash
# attention kernel A/B: folded (spare-channel) kernel vs the split-row kernel at d = 40
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn exit $?"; tail -5 gpurun_out/pytest_attn.log
for fold in 1 0; do for emu in 1 0; do
  echo "== FOLD=$fold EMU=$emu"
  CCEDIT_ATTN_FOLD=$fold CCEDIT_ATTN_EMU=$emu timeout 300 python tools/dev_attn.py 2>&1 | grep -E "BAD|attn F=34 L=6144|L=6144 Lkv=6144|scale=4" 
done; done
