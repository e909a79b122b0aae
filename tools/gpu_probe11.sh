#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "temporal_attention" > gpurun_out/pytest_ta.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_ta.log
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from ccedit_b200 import ops
def bench(B, T, HW, heads, d):
    C = heads * d
    qkv = torch.randn(B, T, HW, 3 * C, device='cuda').half()
    q = torch.randn(B, T, HW, C, device='cuda').half()
    out = torch.empty(B, T, HW, C, dtype=torch.float16, device='cuda')
    run = lambda: ops.temporal_attention(q, qkv[..., C:2*C], qkv[..., 2*C:], heads, out)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"temporal attn B={B} T={T} HW={HW} d={d}: {ms*1e3:.1f} us {4*B*T*HW*C*2/ms/1e6:.0f} GB/s")
for a in [(2,17,6144,8,40),(2,17,1536,8,80),(2,17,384,8,160),(2,17,96,8,160),(2,33,1536,8,80)]: bench(*a)
PY
