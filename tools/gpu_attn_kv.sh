#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -3
cat > /tmp/akv.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_attn as da
da.bench(34, 6144, 8, 40, iters=6)
da.bench(17, 6144, 8, 40, iters=6)
da.bench(34, 1536, 8, 80, iters=10)
da.bench(34, 384, 8, 160, iters=10)
da.bench(34, 96, 8, 160, iters=10)
da.bench(66, 768, 8, 40, iters=10)
PY
timeout 300 python /tmp/akv.py 2>&1 | grep "attn F" | tee gpurun_out/attn_kv.txt
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -3
