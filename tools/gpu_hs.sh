#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "hint_stem" 2>&1 | tail -3
cat > /tmp/hs.py <<'PY'
import os, sys, math, torch
sys.path.insert(0, os.getcwd())
from ccedit_b200 import ops
torch.manual_seed(0)
Fr, H, W = 17, 512, 768
x = torch.randn(Fr, H, W, 16, device="cuda").half()
w2, b2 = torch.randn(32, 16, 3, 3) / 12, torch.randn(32)
w3, b3 = torch.randn(32, 32, 3, 3) / 17, torch.randn(32)
p2 = ops.pack_hint_stem_weight(w2, b2, "cuda", 16, 144)
p3 = ops.pack_hint_stem_weight(w3, b3, "cuda", 32, 288)
for _ in range(3): y = ops.hint_stem23(x, *p2, *p3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = ops.hint_stem23(x, *p2, *p3)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"hint_stem23 17x512x768: {ms*1e3:.1f} us  {(x.numel()+y.numel())*2/ms/1e6:.0f} GB/s")
PY
timeout 300 python /tmp/hs.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -3
