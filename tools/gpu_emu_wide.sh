#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/ew.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_attn as da
da.bench(34, 1536, 8, 80, iters=20)
da.bench(34, 384, 8, 160, iters=20)
da.bench(34, 6144, 8, 40, iters=4)
PY
for m in unset 1 unset 1; do echo "== EMU=$m"; if [ $m = unset ]; then timeout 300 python /tmp/ew.py 2>&1 | grep "attn F"; else CCEDIT_ATTN_EMU=$m timeout 300 python /tmp/ew.py 2>&1 | grep "attn F"; fi; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -2
