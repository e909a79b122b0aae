#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "hint_stem" > gpurun_out/pytest_hs.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_hs.log

python - <<'PY'
import torch, math, sys
sys.path.insert(0, '.')
from ccedit_b200 import ops
x = torch.zeros(34, 512, 768, 8, dtype=torch.float16, device='cuda'); x[..., :3] = torch.randn(34, 512, 768, 3, device='cuda').half()
p0 = ops.pack_hint_stem_weight(torch.randn(16, 3, 3, 3) / 5, torch.randn(16), 'cuda', 8, 80)
p1 = ops.pack_hint_stem_weight(torch.randn(16, 16, 3, 3) / 12, torch.randn(16), 'cuda', 16, 144)
for _ in range(2): y = ops.hint_stem01(x, *p0, *p1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = ops.hint_stem01(x, *p0, *p1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"hint_stem01 34x512x768: {ms*1e3:.1f} us, {(x.numel()+y.numel())*2/ms/1e6:.0f} GB/s")
PY
