"""Developer timing: LayerNorm-folded GEMMs with (mean, rstd) rows vs. partial sums finished in the epilogue."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccedit_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def run(M, K, N, geglu, slots):
    a = torch.randn(M, K, device=dev).half()
    w = torch.randn(N, K) / math.sqrt(K)
    pw = ops.pack_weight(w, torch.randn(N), dev, geglu=geglu, ln_gamma=torch.ones(K), ln_beta=torch.zeros(K))
    out = torch.empty(M, pw.n_out, dtype=torch.float16, device=dev)
    st = ops.layernorm_stats(a)
    if slots:
        sp = torch.zeros(M, slots, 2, device=dev)
        sp[:, 0, 0] = a.float().sum(1)
        sp[:, 0, 1] = (a.float() ** 2).sum(1)
        st = sp
    f = lambda: ops.gemm(a, pw, out, rowstats=st)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"M={M} K={K} N={N} geglu={geglu} slots={slots}: {ms * 1e3:8.1f} us", flush=True)


for slots in (0, 4):
    run(208896, 320, 2560, True, slots)
    run(208896, 320, 960, False, slots)
    run(208896, 320, 320, False, slots)
    run(52224, 640, 5120, True, slots)
