"""One warm + one profiled network call (eager launches, headline shape by default) for ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/one_call.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccedit_b200.configs import build_network  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="tv2v")
ap.add_argument("--frames", type=int, default=17)
ap.add_argument("--h", type=int, default=64)
ap.add_argument("--w", type=int, default=96)
ap.add_argument("--plain", action="store_true", help="plain batch-2 call instead of the CFG de-duplicated forward_cfg")
a = ap.parse_args()
dev = torch.device("cuda", 0)
wrap = build_network(a.kind, device=dev, use_cuda_graph=False, randomize_zero_init_seed=1)
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 4, a.frames, a.h, a.w, generator=g).to(dev)
c = {"crossattn": torch.randn(2, 77, 768, generator=g).to(dev),
     "control_hint": (torch.rand(1, 3, a.frames, 8 * a.h, 8 * a.w, generator=g) * 2 - 1).repeat(2, 1, 1, 1, 1).to(dev)}
if a.kind == "tvi2v":
    c["cond_feat"] = torch.randn(1, 4, a.h, a.w, generator=g).repeat(2, 1, 1, 1).to(dev)
t = torch.full((2,), 500, dtype=torch.long, device=dev)
call = (lambda: wrap(x, t, c)) if a.plain else (lambda: wrap.forward_cfg(x[:1], t[:1], c))   # the bench's call
call()
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = call()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out.abs().mean()))
