#!/bin/bash
cat > /tmp/tr.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import dev_gemm as dg
from ccedit_b200 import ops
dg.trace(0, 320, 320, taps=ops.conv_taps(), shape=(34, 64, 96))
dg.run(0, 320, 320, taps=ops.conv_taps(), shape=(34, 64, 96))
dg.run(13056, 1280, 1280, res=True)
dg.run(0, 1280, 1280, taps=ops.conv_taps(), shape=(34, 16, 24))
dg.run(52224, 640, 640, res=True)
PY
for m in 0 1; do echo "== CLUSTER=$m"; CCEDIT_GEMM_CLUSTER=$m timeout 300 python /tmp/tr.py 2>&1 | tail -20; done
