#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "temporal_attention" 2>&1 | tail -3
for m in 1 0; do echo "== FIXED=$m"; CCEDIT_TA_FIXED=$m timeout 300 python tools/dev_norm.py 2>&1 | grep t_attn; done | tee gpurun_out/ta.txt
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -3
