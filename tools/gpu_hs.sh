#!/bin/bash
# fused hint-stem kernels: parity cases, isolated timing at the de-duplicated call's size (17 frames 512 x 768), network parity
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "hint_stem" 2>&1 | tail -3
cat > /tmp/hs.py <<'PY'
import os, sys, math, torch
sys.path.insert(0, os.getcwd())
from ccedit_b200 import ops
torch.manual_seed(0)
Fr, H, W = 17, 512, 768
x0 = torch.randn(Fr, H, W, 8, device="cuda").half()
ws = [(torch.randn(16, 3, 3, 3) / 5, torch.randn(16), 8, 80), (torch.randn(16, 16, 3, 3) / 12, torch.randn(16), 16, 144),
      (torch.randn(32, 16, 3, 3) / 12, torch.randn(32), 16, 144), (torch.randn(32, 32, 3, 3) / 17, torch.randn(32), 32, 288)]
p = [ops.pack_hint_stem_weight(w, b, "cuda", c, k) for w, b, c, k in ws]
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
y = ops.hint_stem01(x0, *p[0], *p[1])
print(f"hint_stem01 17x512x768: {timed(lambda: ops.hint_stem01(x0, *p[0], *p[1])):.1f} us")
print(f"hint_stem23 17x512x768: {timed(lambda: ops.hint_stem23(y, *p[2], *p[3])):.1f} us")
PY
timeout 300 python /tmp/hs.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_network_gpu.py -x -q -m gpu -k "golden or full" 2>&1 | tail -3
