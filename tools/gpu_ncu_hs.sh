#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/hs1.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from ccedit_b200 import ops
torch.manual_seed(0)
Fr, H, W = 17, 512, 768
x0 = torch.randn(Fr, H, W, 8, device="cuda").half()
w0, b0 = torch.randn(16, 3, 3, 3) / 5, torch.randn(16)
w1, b1 = torch.randn(16, 16, 3, 3) / 12, torch.randn(16)
w2, b2 = torch.randn(32, 16, 3, 3) / 12, torch.randn(32)
w3, b3 = torch.randn(32, 32, 3, 3) / 17, torch.randn(32)
p0 = ops.pack_hint_stem_weight(w0, b0, "cuda", 8, 80); p1 = ops.pack_hint_stem_weight(w1, b1, "cuda", 16, 144)
p2 = ops.pack_hint_stem_weight(w2, b2, "cuda", 16, 144); p3 = ops.pack_hint_stem_weight(w3, b3, "cuda", 32, 288)
for _ in range(3):
    y = ops.hint_stem01(x0, *p0, *p1)
    z = ops.hint_stem23(y, *p2, *p3)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hint_stem -s 4 -c 2 -f -o gpurun_out/r02_hint_stem python /tmp/hs1.py > gpurun_out/ncu_hs.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_hs.log
