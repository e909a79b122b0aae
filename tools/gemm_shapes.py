"""Developer report: time of one network call split by kernel class and GEMM shape (CUDA events per launch)."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccedit_b200 import ops  # noqa: E402
from ccedit_b200.configs import build_network  # noqa: E402

dev = torch.device("cuda", 0)
kind = sys.argv[1] if len(sys.argv) > 1 else "tv2v"
wrap = build_network(kind, device=dev, use_cuda_graph=False, randomize_zero_init_seed=1)
g = torch.Generator().manual_seed(0)
T, h, w = 17, 64, 96
x = torch.randn(2, 4, T, h, w, generator=g).to(dev)
c = {"crossattn": torch.randn(2, 77, 768, generator=g).to(dev),
     "control_hint": (torch.rand(1, 3, T, 8 * h, 8 * w, generator=g) * 2 - 1).repeat(2, 1, 1, 1, 1).to(dev)}
if kind == "tvi2v":
    c["cond_feat"] = torch.randn(1, 4, h, w, generator=g).repeat(2, 1, 1, 1).to(dev)
t = torch.full((2,), 500, dtype=torch.long, device=dev)
wrap(x, t, c)
wrap(x, t, c)
torch.cuda.synchronize()
ops._PROF_SHAPES = True
ops.profile_start()
wrap(x, t, c)
recs = ops.profile_stop()
agg = collections.OrderedDict()
for name, fl, by, ms in recs:
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += fl; a[2] += by; a[3] += ms
tot = sum(a[3] for a in agg.values())
print(f"total {tot:.2f} ms over {len(recs)} launches")
for name, (n, fl, by, ms) in sorted(agg.items(), key=lambda kv: -kv[1][3])[:120]:
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}% n={n:3d} {ms / n * 1e3:8.1f} us/launch {fl / ms / 1e9 if fl else 0:7.1f} TF/s {by / ms / 1e6:7.0f} GB/s  {name}")
