#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "groupnorm" > gpurun_out/pytest_gn.log 2>&1; echo "pytest gn exit $?"; tail -5 gpurun_out/pytest_gn.log
python - <<'PY'
import torch, sys, os
sys.path.insert(0, '.')
from ccedit_b200 import ops
def bench(F, HW, C, label):
    x = torch.randn(F, HW, C, device='cuda').half(); g = torch.randn(C, device='cuda'); b = torch.randn(C, device='cuda')
    y = torch.empty_like(x)
    for _ in range(3): ops.groupnorm_spatial(x, g, b, 1e-5, True, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.groupnorm_spatial(x, g, b, 1e-5, True, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{label} GN F={F} HW={HW} C={C}: {ms*1e3:.1f} us  {x.numel()*4/ms/1e6:.0f} GB/s (1R+1W)")
for args in [(34, 6144, 320), (34, 6144, 640), (34, 1536, 640), (34, 1536, 1280), (34, 384, 1280), (34, 96, 1280), (34, 6144, 960)]:
    bench(*args, os.environ.get("CCEDIT_GN_FUSED", "1"))
PY
