#!/bin/bash
# ncu --set full of the short-key attention kernel at the top level of the network (34 x 6144 rows, 77 keys, d = 40)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/one_short.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from ccedit_b200 import ops
torch.manual_seed(0)
Fr, L, heads, d = 34, 6144, 8, 40
C = heads * d
q = torch.randn(Fr, L, C, device="cuda").half()
kv = torch.randn(2, 77, 2 * C, device="cuda").half()
out = torch.empty(Fr, L, C, dtype=torch.float16, device="cuda")
for _ in range(3):
    ops.attention(q, [ops.KVSegment(kv[..., :C], kv[..., C:], div=Fr // 2)], heads, out)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:short_kv -s 2 -c 1 -f -o gpurun_out/r02_attn_short python /tmp/one_short.py > gpurun_out/ncu_attn_short.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_attn_short.log
