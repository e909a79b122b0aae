#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -3
cat > /tmp/cx.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_attn as da
for (L, d) in ((6144, 40), (1536, 80), (384, 160), (96, 160)):
    da.bench_cross(34, L, 77, 8, d)
PY
for m in 1 0; do echo "== SHORT=$m"; CCEDIT_ATTN_SHORT=$m timeout 300 python /tmp/cx.py 2>&1 | grep cross; done | tee gpurun_out/cross.txt
