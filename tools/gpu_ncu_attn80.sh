#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/one_attn80.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from ccedit_b200 import ops
torch.manual_seed(0)
Fr, L, heads, d = 34, 1536, 8, 80
C = heads * d
qkv = torch.randn(Fr, L, 3 * C, device="cuda").half()
out = torch.empty(Fr, L, C, dtype=torch.float16, device="cuda")
for _ in range(3):
    ops.attention(qkv[..., :C], [ops.KVSegment(qkv[..., C:2 * C], qkv[..., 2 * C:])], heads, out)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 2 -c 1 -f -o gpurun_out/r02_attn_d80 python /tmp/one_attn80.py > gpurun_out/ncu_attn_d80.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_attn_d80.log
