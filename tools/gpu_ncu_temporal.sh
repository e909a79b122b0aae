#!/bin/bash
# ncu --set full of the temporal attention and the two GroupNorm kernels at the top level of the network
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/one_t.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from ccedit_b200 import ops
torch.manual_seed(0)
B, T, HW, heads, d = 2, 17, 6144, 8, 40
C = heads * d
qkv = torch.randn(B, T, HW, 3 * C, device="cuda").half()
out = torch.empty(B, T, HW, C, dtype=torch.float16, device="cuda")
x = torch.randn(B, T, HW, C, device="cuda").half()
y = torch.empty_like(x)
g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
for _ in range(3):
    ops.temporal_attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, out)
    ops.groupnorm_temporal(x, g, b, 1e-5, True, out=y)
    ops.groupnorm_spatial(x.view(B * T, HW, C), g, b, 1e-5, True, out=y.view(B * T, HW, C))
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"temporal_attn|gn_temporal|gn_spatial" -s 6 -c 3 -f -o gpurun_out/r02_hbm_kernels python /tmp/one_t.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_hbm.log
