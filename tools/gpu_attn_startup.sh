#!/bin/bash
# how much of the top-level self-attention is CTA start-up: same query tiles, 1x / 2x / 0.5x the keys
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
cat > /tmp/su.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from ccedit_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def bench(Fr, L, Lkv, heads, d, iters=6):
    C = heads * d
    q = torch.randn(Fr, L, C, device=dev).half()
    kv = torch.randn(Fr, Lkv, 2 * C, device=dev).half()
    out = torch.empty(Fr, L, C, dtype=torch.float16, device=dev)
    run = lambda: ops.attention(q, [ops.KVSegment(kv[..., :C], kv[..., C:])], heads, out)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"attn F={Fr} Lq={L} Lkv={Lkv} d={d}: {ms:8.3f} ms  {4.0 * Fr * L * Lkv * C / ms / 1e9:8.1f} TFLOP/s", flush=True)
for lkv in (3072, 6144, 12288):
    bench(34, 6144, lkv, 8, 40)
for lkv in (768, 1536, 3072):
    bench(34, 1536, lkv, 8, 80)
PY
timeout 300 python /tmp/su.py 2>&1 | grep "attn F" | tee gpurun_out/attn_startup.txt
