#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/em.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_attn as da
for _ in range(2):
    da.bench(34, 6144, 8, 40, iters=10)
da.bench(17, 6144, 8, 40, iters=10)
PY
for m in 1 0 1 0; do echo "== EMU=$m"; CCEDIT_ATTN_EMU=$m timeout 300 python /tmp/em.py 2>&1 | grep "attn F"; done | tee gpurun_out/emu.txt
